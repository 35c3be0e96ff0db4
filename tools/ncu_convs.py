"""Summarise an `ncu --csv` capture of the 70 tensor-core conv launches of one forward into a per-conv JSON table.

usage: python tools/ncu_convs.py gpurun_out/convs.csv profiles/r01_conv_per_launch_final.json out.json
The second argument is an earlier table of the same 70 launches (names, shapes, algorithmic GFLOP per launch);
its `ms` column is carried along as `ms_prev` for comparison.
"""
import csv, json, sys

def main(csv_path, prev_path, out_path):
    rows = {}
    with open(csv_path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        d = rows.setdefault(int(r['ID']), {'kernel': r['Kernel Name']})
        d[r['Metric Name']] = (float(r['Metric Value'].replace(',', '')), r['Metric Unit'])
    prev = json.load(open(prev_path))
    assert len(rows) == len(prev), (len(rows), len(prev))
    out = []
    for (i, d), p in zip(sorted(rows.items()), prev):
        t, unit = d['gpu__time_duration.sum']
        ms = t / 1e6 if unit in ('ns', 'nsecond') else (t / 1e3 if unit in ('us', 'usecond') else t)
        kern = 'gemm_tc_kernel' if 'gemm_tc_kernel' in d['kernel'] else 'conv_tc_kernel'
        assert kern == p['kernel'], (i, kern, p)
        out.append({
            'conv': p['conv'], 'kernel': kern, 'M': p['M'], 'Cin': p['Cin'], 'Cout': p['Cout'], 'taps': p['taps'],
            'ms': round(ms, 4), 'ms_prev': p['ms'], 'gflop_alg': p['gflop_alg'],
            'tflops_alg': round(p['gflop_alg'] / ms, 1),
            'tensor_pipe_pct': round(d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'][0], 1),
            'dram_pct': round(d['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'][0], 1),
            'dram_mb': round((d['dram__bytes_read.sum'][0] + d['dram__bytes_write.sum'][0]) / 1e6, 1),
        })
        if 'lts__t_bytes.sum' in d:        # all L2 slice traffic (SM fills, stores, DRAM side): the large-K convs sit at the LTS cap
            out[-1]['l2_gb'] = round(d['lts__t_bytes.sum'][0] / 1e9, 3)
            out[-1]['l2_tbps'] = round(d['lts__t_bytes.sum'][0] / 1e9 / ms, 2)
    tot = sum(o['ms'] for o in out)
    tw = sum(o['ms'] * o['tensor_pipe_pct'] for o in out) / tot
    json.dump(out, open(out_path, 'w'), indent=0)
    print(f'{len(out)} launches, sum {tot:.3f} ms (prev {sum(o["ms_prev"] for o in out):.3f}), time-weighted tensor pipe {tw:.1f} %,'
          f' DRAM {sum(o["dram_mb"] for o in out) / 1e3:.2f} GB')
    for o in sorted(out, key=lambda o: -o['ms'])[:12]:
        print(f"  {o['conv']:10s} {o['kernel']:15s} {o['ms']:.4f} ms (prev {o['ms_prev']:.4f})  tensor {o['tensor_pipe_pct']:5.1f} %  dram {o['dram_pct']:5.1f} %  L2 {o.get('l2_tbps', 0):.1f} TB/s")

if __name__ == '__main__':
    main(*sys.argv[1:4])
