"""Extract the metrics the docs quote from an .ncu-rep (read here, no GPU) into a small tracked JSON.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rNN_ncu_<kernel>.json [note]"""
import csv, io, json, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg.pct_of_peak_sustained_active',
        'TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']


def main(rep, out, note=''):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        item = {'kernel': d.get('Kernel Name', '')[:120]}
        for k in KEYS:
            if d.get(k) not in (None, ''):
                try:
                    item[k] = {'value': float(d[k].replace(',', '')), 'unit': u.get(k, '')}
                except ValueError:
                    item[k] = {'value': d[k], 'unit': u.get(k, '')}
        for k in hdr:                                  # any tensor sub-pipe metric the report carries
            if 'subpipe_hmma' in k and k not in item and d.get(k):
                item[k] = {'value': d[k], 'unit': u.get(k, '')}
        res.append(item)
    json.dump({'source': rep.split('/')[-1], 'note': note, 'launches': res}, open(out, 'w'), indent=1)
    for it in res:
        print({k: (v['value'] if isinstance(v, dict) else v) for k, v in it.items()})


if __name__ == '__main__':
    main(*sys.argv[1:4])
