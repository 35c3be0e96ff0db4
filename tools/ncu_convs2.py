"""Round-2 per-launch table of the tensor-core conv family (conv_tc_kernel, gemm_tc_kernel, bott_tc_kernel) of one bs32 forward.

capture (under gpurun):
  ncu --clock-control none --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,\
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum \
      -k regex:"conv_tc_kernel|gemm_tc_kernel|bott_tc_kernel" -s 67 -c 67 --csv --log-file gpurun_out/convs2.csv python tools/fwd_once.py 2
usage: python tools/ncu_convs2.py gpurun_out/convs2.csv profiles/r02_conv_per_launch.json [B]
The launch order is the graph's (horizonnet_b200/csrc/model.cu forward_encoder / forward_rnn); names and algorithmic FLOPs
are regenerated here from the same topology.
"""
import csv, json, sys


def graph(B):
    """[(name, kernel, M, Cin, Cout, taps, gflop)] in launch order (tensor-core path, fused layer1 bottlenecks)."""
    out = []
    planes = (64, 128, 256, 512)
    nblk = (3, 4, 6, 3)
    H, W, inpl = 128, 256, 64

    def conv(name, kern, h, w, ci, co, taps, halo_rows):
        M = B * h * (w + 2) if halo_rows else B * h * w
        gflop = 2.0 * B * h * w * ci * co * taps / 1e9
        out.append((name, kern, M, ci, co, taps, gflop))

    for l in range(4):
        p = planes[l]
        for b in range(nblk[l]):
            s = 2 if (b == 0 and l > 0) else 1
            Ho, Wo = H // s, W // s
            n = f'l{l + 1}.{b}'
            gem1 = inpl <= 256
            conv(n + '.c1', 'gemm_tc_kernel' if gem1 else 'conv_tc_kernel', H, W, inpl, p, 1, True)
            if b == 0:
                conv(n + '.ds', 'gemm_tc_kernel' if (s == 1 and inpl <= 256) else 'conv_tc_kernel', Ho, Wo, inpl, 4 * p, 1, s == 1)
            if l == 0:
                M = B * Ho * Wo
                out.append((n + '.c2c3', 'bott_tc_kernel', M, p, 4 * p, 9, 2.0 * M * (p * p * 9 + p * 4 * p) / 1e9))
            else:
                conv(n + '.c2', 'conv_tc_kernel', Ho, Wo, p, p, 9, False)
                conv(n + '.c3', 'gemm_tc_kernel' if p <= 256 else 'conv_tc_kernel', Ho, Wo, p, 4 * p, 1, True)
            inpl = 4 * p
            H, W = Ho, Wo
    for s in range(4):
        c = planes[s] * 4
        ch = (c, c // 2, c // 2, c // 4, c // 8)
        h, w = 128 >> s, 256 >> s
        for j in range(4):
            conv(f'ghc{s}.{j}', 'conv_tc_kernel', h // 2, w, ch[j], ch[j + 1], 9, False)
            h //= 2
    for layer in range(2):
        out.append((f'xproj{layer}', 'conv_tc_kernel', 256 * B, 1024, 4096, 1, 2.0 * 256 * B * 1024 * 4096 / 1e9))
    return out


def main(csv_path, out_path, B=32):
    rows = {}
    with open(csv_path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        d = rows.setdefault(int(r['ID']), {'kernel': r['Kernel Name']})
        d[r['Metric Name']] = (float(r['Metric Value'].replace(',', '')), r['Metric Unit'])
    g = graph(int(B))
    assert len(rows) == len(g), (len(rows), len(g))
    out = []
    for (i, d), (name, kern, M, ci, co, taps, gflop) in zip(sorted(rows.items()), g):
        assert kern in d['kernel'], (i, name, kern, d['kernel'][:80])
        t, unit = d['gpu__time_duration.sum']
        ms = t / 1e6 if unit in ('ns', 'nsecond') else (t / 1e3 if unit in ('us', 'usecond') else t)
        o = {'conv': name, 'kernel': kern, 'M': M, 'Cin': ci, 'Cout': co, 'taps': taps, 'ms': round(ms, 4),
             'gflop_alg': round(gflop, 2), 'tflops_alg': round(gflop / ms, 1),
             'tensor_pipe_pct': round(d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'][0], 1),
             'dram_pct': round(d['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'][0], 1),
             'dram_mb': round((d['dram__bytes_read.sum'][0] + d['dram__bytes_write.sum'][0]) / 1e6, 1)}
        if 'lts__t_bytes.sum' in d:
            o['l2_tbps'] = round(d['lts__t_bytes.sum'][0] / 1e9 / ms, 2)
        out.append(o)
    tot = sum(o['ms'] for o in out)
    tw = sum(o['ms'] * o['tensor_pipe_pct'] for o in out) / tot
    dram = sum(o['dram_mb'] for o in out)
    json.dump(out, open(out_path, 'w'), indent=0)
    json.dump({'launches': len(out), 'sum_ms': round(tot, 3), 'time_weighted_tensor_pipe_pct': round(tw, 2), 'dram_gb': round(dram / 1e3, 2),
               'dram_bytes_per_launch_avg': dram * 1e6 / len(out), 'gflop_alg': round(sum(o['gflop_alg'] for o in out), 1)},
              open(out_path.replace('.json', '_summary.json'), 'w'), indent=1)
    print(f'{len(out)} launches, sum {tot:.3f} ms, time-weighted tensor pipe {tw:.1f} %, DRAM {dram / 1e3:.2f} GB')
    for o in sorted(out, key=lambda o: -o['ms'])[:14]:
        print(f"  {o['conv']:10s} {o['kernel']:15s} {o['ms']:.4f} ms  tensor {o['tensor_pipe_pct']:5.1f} %  dram {o['dram_pct']:5.1f} %  L2 {o.get('l2_tbps', 0):.1f} TB/s")


if __name__ == '__main__':
    main(*sys.argv[1:4])
