"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares of the summed device time).
usage: python tools/launch_summary.py gpurun_out/launches.csv profiles/rNN_launches_summary.json "<command>" """
import csv, json, re, sys


def main(path, out, command):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    agg = {}
    for r in csv.DictReader(lines):
        if r['Metric Name'] != 'gpu__time_duration.sum':
            continue
        k = re.sub(r'\(.*$', '', r['Kernel Name']).replace('void ', '').strip()
        t = float(r['Metric Value'].replace(',', ''))
        t = t / 1e6 if r['Metric Unit'] in ('ns', 'nsecond') else (t / 1e3 if r['Metric Unit'] in ('us', 'usecond') else t)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(v[1] for v in agg.values())
    ks = [{'kernel': k, 'launches': n, 'total_ms': round(t, 3), 'share': round(t / tot, 4)}
          for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    ours = sum(k['total_ms'] for k in ks if 'hn::' in k['kernel'])
    json.dump({'command': command, 'note': 'whole process (weight packing, warm-up + timed forwards of both timed regions, e2e forwards, '
               'pano_stretch 10k, augmentation aux); cold-cache serialised times under ncu: compare shares',
               'total_ms': round(tot, 2), 'share_of_library_kernels': round(ours / tot, 4), 'kernels': ks}, open(out, 'w'), indent=1)
    for k in ks[:14]:
        print(k)
    print('library kernels share', round(ours / tot, 4))


if __name__ == '__main__':
    main(*sys.argv[1:4])
