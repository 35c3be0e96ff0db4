"""Read an .ncu-rep here (no GPU): kernel summary + hottest SASS instructions with stall reasons."""
import csv, subprocess, sys, io
rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.012
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2]))
for k in ['Kernel Name', 'Grid Size', 'gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
          'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
          'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
          'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
          'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']:
    print('%-70s %s' % (k, d.get(k, '')[:80]))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ia = hdr.index('Source'); isamp = hdr.index('# Samples'); iex = hdr.index('Instructions Executed')
body = [r for r in rows[2:] if len(r) > isamp and r[isamp].isdigit()]
tot = sum(int(r[isamp]) for r in body)
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: 0 for h in stall_cols}
for r in body:
    for h in stall_cols:
        v = r[hdr.index(h)]
        if v.isdigit(): agg[h] += int(v)
print('samples', tot, 'instrs', len(body), sorted(agg.items(), key=lambda kv: -kv[1])[:7])
for k, r in enumerate(body):
    s = int(r[isamp])
    if s > tot * thr:
        nz = [(h[6:], r[hdr.index(h)]) for h in stall_cols if r[hdr.index(h)] not in ('0', '', '-')]
        nz.sort(key=lambda kv: -int(kv[1]))
        print(k, s, r[iex], r[ia][:64].strip(), nz[:2])
