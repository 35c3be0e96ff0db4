"""Device time of one training step (BASELINE config 5 per GPU: batch 8, L1 + BCE-with-logits, Adam) through the library.
Usage (GPU box): python tools/train_bench.py [batch] [steps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as e; e.build()
import torch
from bench import train_step_aux
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
print(json.dumps(train_step_aux(torch.device('cuda:0'), B, steps)))
