#!/usr/bin/env python
"""bench.py -- panoramas/sec of HorizonNet('resnet50', rnn).forward at 512x1024, batch 32 per GPU.

    python bench.py --gpus N --steps K --warmup W           # our CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N ...           # the reference's CPU implementation (oracle port)

One "step" = one forward over one batch of 32 synthetic panoramas per GPU (BASELINE.json configs[1];
N>1 = configs[3]: data-parallel shards + one NCCL all-gather of the [32,3,1024] outputs per step).
Prints ONE JSON line (rank 0).  Timing: W warm-up steps, then exactly K steps bracketed by
barrier + cuda synchronize, CUDA events on the launching stream, max over ranks.  The inputs rotate
between two 201 MB batches and every step streams >9 GB of activations, so nothing survives in the
126 MB L2 between steps ("inputs larger than L2").
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BATCH = 32
GFLOP_PER_PANO = 142.90            # BASELINE.md section 2 (algorithmic, all convs + LSTM + head)
WORKLOAD = 'batch-32 synthetic 512x1024 panoramas, resnet50_rnn forward (BASELINE configs[1])'


def _peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return {'hbm_gbs': p['hbm_gbs'], 'tf': p.get('bf16_tflops_sustained', p['bf16_tflops']),
                'tf_burst': p['bf16_tflops'], 'src': 'measured'}
    return {'hbm_gbs': 6650.0, 'tf': 1400.0, 'tf_burst': 1590.0, 'src': 'fallback'}


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU with NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.sm = []
        self.reasons = set()
        self.sm_max = None
        self.power = []

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                nv.nvmlClocksThrottleReasonHwSlowdown: 'hw_slowdown',
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: 'hw_thermal_slowdown',
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: 'sw_thermal_slowdown',
                nv.nvmlClocksThrottleReasonSwPowerCap: 'sw_power_cap',
                nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: 'hw_power_brake',
            }
            while not self.stop_flag.is_set():
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                except Exception:
                    pass
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(0.05)
        except Exception as e:              # NVML missing: report that, never fake numbers
            self.reasons.add(f'nvml_unavailable:{type(e).__name__}')

    def summary(self):
        sm = sorted(self.sm)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': self.sm_max,
                'power_w_max': max(self.power) if self.power else None, 'samples': len(sm),
                'reasons': sorted(self.reasons)}


def _cpu_oracle_rate(max_seconds, batch=1):
    """Panoramas/s of the CPU oracle (restatement of the reference forward) on this host."""
    import torch
    from oracle import horizonnet_ref
    from horizonnet_b200.weights import synthetic_state_dict, synthetic_panoramas
    sd = synthetic_state_dict(0, 'random')
    x = synthetic_panoramas(batch, seed=11)
    with torch.no_grad():
        horizonnet_ref.forward(sd, x)                       # warm-up
        t0 = time.perf_counter()
        n = 0
        while True:
            horizonnet_ref.forward(sd, x)
            n += 1
            if time.perf_counter() - t0 > max_seconds or n >= 16:
                break
        dt = time.perf_counter() - t0
    return n * batch / dt, n, torch.get_num_threads()


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (the oracle port:
    /root/reference cannot travel to the GPU box and has no compiled code), all host threads."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    from oracle import horizonnet_ref
    from horizonnet_b200.weights import synthetic_state_dict, synthetic_panoramas
    # torchrun exports OMP_NUM_THREADS=1 to its workers; this arm is the only rank doing work: use every host thread
    torch.set_num_threads(max(torch.get_num_threads(), os.cpu_count() or 1))
    sd = synthetic_state_dict(0, 'random')
    x = synthetic_panoramas(1, seed=11)
    with torch.no_grad():
        for _ in range(max(1, min(args.warmup, 2))):
            horizonnet_ref.forward(sd, x)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            horizonnet_ref.forward(sd, x)
        dt = time.perf_counter() - t0
    value = args.steps * 1 / dt
    cores = torch.get_num_threads()
    sample = '1 panorama per step (bounded sample of the batch-32 workload), fp32 CPU torch ops'
    emit_json({
        'impl': 'reference', 'metric': 'panoramas/sec', 'value': value, 'unit': 'panoramas/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'sample': sample},
        'cpu_baseline': {'value': value, 'unit': 'panoramas/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'panoramas/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    })


def run_ours(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    from horizonnet_b200 import _lib
    from horizonnet_b200.model import HorizonNet
    from horizonnet_b200.misc.panostretch import pano_stretch_batch
    from horizonnet_b200.parallel import gather_outputs
    from horizonnet_b200.weights import synthetic_state_dict, synthetic_panoramas

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- this framework has no CPU path (use --impl reference)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    if rank == 0:
        entry.build()
    if world > 1:
        dist.barrier()
    lib = _lib.lib()

    sd = synthetic_state_dict(0, 'random')
    net = HorizonNet('resnet50', True).eval()
    net.load_state_dict(sd, strict=True)
    net.use_tensor_cores(not args.fp32)
    net = net.to(dev)
    # two distinct input batches per rank, rotated (seeded per rank: SURVEY 8d config 4)
    xs = [synthetic_panoramas(BATCH, seed=1000 + rank + 100 * i).to(dev) for i in range(2)]

    def step(i):
        with torch.no_grad():
            bon, cor = net(xs[i & 1])
        if world > 1:
            bon, cor = gather_outputs(bon, cor)                       # NCCL all-gather of (y_cor, y_bon)
        return bon, cor

    for i in range(args.warmup):
        step(i)
    net.set_profile(True)
    net.read_profile(reset=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = lib.hn_kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.stop_flag.set()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    launches = torch.tensor([lib.hn_kernel_launches() - launches0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(launches, op=dist.ReduceOp.SUM)
    net.check()
    total_ms = float(ms.item())
    prof = net.read_profile(reset=True)
    net.set_profile(False)
    sampler.join(timeout=2)
    value = world * BATCH * args.steps / (total_ms * 1e-3)

    # ---- end-to-end through the C ABI with HOST buffers (H2D + forward + D2H inside the call)
    xh = [synthetic_panoramas(BATCH, seed=2000 + rank + 100 * i).pin_memory() for i in range(2)]
    e2e_steps = max(3, min(args.steps, 10))
    net.forward_host(xh[0], device=local)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    net.submit_host(xh[0], device=local)                 # pipelined host API: upload of batch i+1 overlaps forward i
    for i in range(e2e_steps):
        if i + 1 < e2e_steps:
            net.submit_host(xh[(i + 1) & 1], device=local)
        hb, hc = net.collect_host(device=local)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * BATCH * e2e_steps / float(e2e_s.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = _peaks()
    # ---- roofline of the dominant kernel family: the implicit-GEMM convolution kernel
    conv_cls = ('encoder_convs', 'height_reduction_convs', 'lstm_input_projection')
    conv_ms = sum(prof[c][0] for c in conv_cls)
    conv_flops = sum(prof[c][1] for c in conv_cls)
    conv_n = sum(prof[c][2] for c in conv_cls)
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    stage_ms = {k: round(v[0] / args.steps, 4) for k, v in prof.items()}
    traffic = None
    try:
        import glob
        files = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r*_conv_tc_traffic.json')))
        if files and not args.fp32:
            traffic = json.load(open(files[-1]))['dram_bytes_per_launch_avg']
    except Exception:
        traffic = None
    roofline = {
        'kernel': 'conv implicit-GEMM family (%s)' % ('conv_igemm_f32, fp32 CUDA cores' if args.fp32 else 'conv_tc_kernel, tcgen05 split-fp16 x3 products'),
        'bound': 'tensor', 'achieved': round(achieved, 3), 'peak': peaks['tf'], 'unit': 'TFLOP/s',
        'frac': round(achieved / peaks['tf'], 5), 'peak_source': peaks['src'] + ' bf16 dense, sustained',
        'traffic': traffic, 'traffic_source': 'dram__bytes_read+write per launch (avg of the 70 conv launches), ncu --set full, see profiles/' if traffic else None,
        'launches_per_step': conv_n / args.steps,
        'avg_launch_ms': round(conv_ms / max(conv_n, 1), 5),
        'algorithmic_gflop_per_step': round(conv_flops / args.steps / 1e9, 2),
        'share_of_step': round(conv_ms / total_ms, 4),
        'issued_over_algorithmic': 1 if args.fp32 else 3,
        'note': 'achieved = algorithmic conv FLOPs (2*M*N*K, single product) / summed CUDA-event time of the conv launches in the timed region; the tensor pipe issues 3x that (hi*hi + hi*lo + lo*hi)',
    }

    # ---- auxiliary: pano_stretch kernel (BASELINE configs[2]) against the HBM roofline
    aux = {'stage_ms_per_step': stage_ms}
    try:
        n_img = 64
        imgs = torch.rand(n_img, 512, 1024, 3, device=dev)
        grid = (0.5, 0.75, 1.0, 1.25, 1.5, 1.75, 2.0)
        pairs = [(a, b) for a in grid for b in grid]
        kx = [pairs[i % 49][0] for i in range(n_img)]
        ky = [pairs[i % 49][1] for i in range(n_img)]
        out = torch.empty_like(imgs)
        for _ in range(3):
            pano_stretch_batch(imgs, kx, ky, out=out)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        a0.record()
        for _ in range(reps):
            pano_stretch_batch(imgs, kx, ky, out=out)
        a1.record()
        torch.cuda.synchronize()
        ps_ms = a0.elapsed_time(a1) / reps
        gbs = n_img * 12582912 / (ps_ms * 1e-3) / 1e9
        aux['pano_stretch'] = {'panos_per_s': round(n_img / (ps_ms * 1e-3), 1), 'achieved_gbs': round(gbs, 1),
                               'peak_gbs': peaks['hbm_gbs'], 'frac': round(gbs / peaks['hbm_gbs'], 4),
                               'bytes_per_pano': 12582912, 'sample': '64 distinct 512x1024x3 fp32 panos (805 MB in+out > L2), 49-pair kx/ky grid'}
    except Exception as e:
        aux['pano_stretch'] = {'error': str(e)}
    # ---- auxiliary: single-panorama inference with device-side TTA (flip + 2 rotations = 4 views), "next" row f2
    try:
        from horizonnet_b200.inference import tta_forward
        xi = synthetic_panoramas(1, seed=77)
        for _ in range(2):
            tta_forward(net, xi, flip=True, rotate=[0.25, 0.5])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 10
        for _ in range(reps):
            tta_forward(net, xi, flip=True, rotate=[0.25, 0.5])     # host tensor in, numpy out: includes H2D + D2H
        dt = (time.perf_counter() - t0) / reps
        aux['tta_single_image'] = {'images_per_s': round(1.0 / dt, 2), 'ms_per_image': round(dt * 1e3, 3), 'views': 4,
                                   'api': 'horizonnet_b200.inference.tta_forward (hn_model_infer_tta), host in / host out'}
    except Exception as e:
        aux['tta_single_image'] = {'error': str(e)}
    lstm_ms = prof['lstm_recurrence'][0] / args.steps
    aux['lstm_recurrence'] = {'ms_per_step': round(lstm_ms, 4),
                              'achieved_gbs': round(151.4e6 / (lstm_ms * 1e-3) / 1e9, 2) if lstm_ms > 0 else None,
                              'peak_gbs': peaks['hbm_gbs'], 'note': 'algorithmic 151.4 MB/batch; latency-bound: 512 dependent steps'}

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle on this host's cores
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, n, cores = _cpu_oracle_rate(20.0)
        cpu = {'value': round(rate, 4), 'unit': 'panoramas/s', 'cores': cores, 'kind': 'port',
               'sample': f'{n} single-panorama forwards of the same random-init resnet50_rnn (oracle/horizonnet_ref.py, torch CPU fp32)'}

    line = {
        'metric': 'panoramas/sec', 'value': round(value, 3), 'unit': 'panoramas/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': round(total_ms / args.steps, 4),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32' if args.fp32 else 'f16x2-split (hi+lo fp16 planes, 3 tcgen05 products, fp32 accumulate; fp32-equivalent)',
        'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'global_batch': BATCH * world, 'parallelism': f'dp{world}',
                   'weights': 'random-init (synthetic_state_dict seed 0, randomised BN statistics)',
                   'l2': 'inputs larger than L2 (2 rotating 201 MB batches, >9 GB activations per step)',
                   'collective': 'all_gather of [32,3,1024] fp32 per rank per step' if world > 1 else 'none'},
        'clocks': sampler.summary(),
        'e2e': {'value': round(e2e_value, 3), 'unit': 'panoramas/s', 'steps': e2e_steps,
                'h2d_bytes_per_step': BATCH * 3 * 512 * 1024 * 4 * world, 'd2h_bytes_per_step': BATCH * 3 * 1024 * 4 * world,
                'api': 'hn_model_submit_host / hn_model_collect_host (pinned host buffers; every step uploads its 201 MB input and reads its outputs back; the upload of batch i+1 overlaps forward i)'},
        'gpu_launches': int(launches.item()),
        'roofline': roofline,
        'cpu_baseline': cpu,
        'tflops_algorithmic': round(value * GFLOP_PER_PANO / 1e3, 2),
        'aux': aux,
    }
    emit_json(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = None


def claim_stdout():
    """stdout must carry exactly ONE JSON line: keep a private duplicate of fd 1 for it and point fd 1 at stderr, so
    that anything else written to stdout at the C level (e.g. NCCL's "NCCL version ..." banner) lands on stderr."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit_json(line):
    data = (json.dumps(line) + '\n').encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--fp32', action='store_true', help='exact fp32 CUDA-core kernels instead of tcgen05')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
