"""Per-launch table of wgrad_tc_kernel (tcgen05 weight gradients) for one batch-B training backward.

capture (under gpurun): tools/gpu_profile_train.sh  (ncu --clock-control none --metrics time, tensor-pipe %, DRAM %, DRAM bytes,
L2 bytes, grid size; -k regex:wgrad_tc_kernel -s 150 -c 75: the third step of `tools/train_bench.py 8 1`)
usage: python tools/ncu_wgrad.py gpurun_out/wgrad_launches.csv profiles/rNN_wgrad_tc_per_launch.json [B]
Launch order = horizonnet_b200/csrc/train_step.cu walk_backward: the LSTM weight gradients (layer 1 then 0, per direction
W_ih then W_hh), then the conv units in reverse graph order (units whose Cin or Cout is not a multiple of 64 -- the stem and
ghc_lst.0.layer.3 -- stay on the fp32 kernel and are not in this list).
"""
import csv, json, sys


def launches(B):
    """[(name, pixels, Cin, Cout, taps)] in launch order."""
    fwd = []
    planes, nblk = (64, 128, 256, 512), (3, 4, 6, 3)
    H, W, inpl = 128, 256, 64
    for l in range(4):
        p = planes[l]
        for b in range(nblk[l]):
            s = 2 if (b == 0 and l > 0) else 1
            Ho, Wo = H // s, W // s
            n = f'l{l + 1}.{b}'
            fwd.append((n + '.c1', B * H * W, inpl, p, 1))
            if b == 0:
                fwd.append((n + '.ds', B * Ho * Wo, inpl, 4 * p, 1))
            fwd.append((n + '.c2', B * Ho * Wo, p, p, 9))
            fwd.append((n + '.c3', B * Ho * Wo, p, 4 * p, 1))
            inpl, H, W = 4 * p, Ho, Wo
    for s in range(4):
        c = planes[s] * 4
        ch = (c, c // 2, c // 2, c // 4, c // 8)
        h, w = 128 >> s, 256 >> s
        for j in range(4):
            fwd.append((f'ghc{s}.{j}', B * (h // 2) * w, ch[j], ch[j + 1], 9))
            h //= 2
    out = []
    for layer in (1, 0):
        for d in ('fwd', 'rev'):
            out.append((f'lstm{layer}.{d}.w_ih', 256 * B, 1024, 2048, 1))
            out.append((f'lstm{layer}.{d}.w_hh', 256 * B, 512, 2048, 1))
    out += [u for u in reversed(fwd) if u[2] % 64 == 0 and u[3] % 64 == 0]
    return out


def main(csv_path, out_path, B=8):
    B = int(B)
    rows = {}
    with open(csv_path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        rows.setdefault(int(r['ID']), {})[r['Metric Name']] = (float(r['Metric Value'].replace(',', '')), r['Metric Unit'])
    ids = sorted(rows)
    names = launches(B)
    assert len(ids) == len(names), (len(ids), len(names))
    table = []
    tot_ms = tot_gf = wt = 0.0
    for i, (name, pix, ci, co, taps) in zip(ids, names):
        m = rows[i]
        t, u = m['gpu__time_duration.sum']
        ms = t / 1e6 if u in ('ns', 'nsecond') else (t / 1e3 if u in ('us', 'usecond') else t)
        gf = 2.0 * pix * ci * co * taps / 1e9
        tp = m['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'][0]
        table.append({'name': name, 'pixels': pix, 'Cin': ci, 'Cout': co, 'taps': taps, 'ctas': int(m['launch__grid_size'][0]),
                      'ms': round(ms, 4), 'gflop': round(gf, 2), 'tflops_algorithmic': round(gf / ms, 1), 'tensor_pipe_pct': round(tp, 1),
                      'dram_pct': round(m['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'][0], 1),
                      'dram_mb': round((m['dram__bytes_read.sum'][0] + m['dram__bytes_write.sum'][0]) / 1e6, 1),
                      'l2_tbs': round(m['lts__t_bytes.sum'][0] / (ms * 1e-3) / 1e12, 2)})
        tot_ms += ms; tot_gf += gf; wt += tp * ms
    summary = {'launches': len(table), 'sum_ms': round(tot_ms, 3), 'gflop': round(tot_gf, 1),
               'tflops_algorithmic': round(tot_gf / tot_ms, 1), 'tensor_pipe_pct_time_weighted': round(wt / tot_ms, 1), 'batch': B,
               'note': 'ncu --clock-control none, cold-cache serialised launches: compare columns, not absolutes; issued MMA work is 3x '
                       'the algorithmic FLOPs (hi*hi + hi*lo + lo*hi)'}
    json.dump({'summary': summary, 'launches': table}, open(out_path, 'w'), indent=1)
    print(summary)
    for r in sorted(table, key=lambda r: -r['ms'])[:25]:
        print(r)


if __name__ == '__main__':
    main(*sys.argv[1:4])
