"""Summaries of ncu reports for profiles/:
     python tools/ncu_summary.py full  <rep.ncu-rep> <out.json>   # --set full capture -> per-kernel dram bytes / time / pipes
     python tools/ncu_summary.py list  <launches.csv> <out.txt>   # gpu__time_duration launch list -> per-kernel shares
"""
import collections
import csv
import io
import json
import subprocess
import sys


def full(rep, out):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = rows[0]
    col = {k: i for i, k in enumerate(h)}
    want = {'dram_read': 'dram__bytes_read.sum', 'dram_write': 'dram__bytes_write.sum', 'us': 'gpu__time_duration.sum',
            'tensor_pct': 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
            'dram_pct': 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'regs': 'launch__registers_per_thread',
            'ipc': 'sm__inst_executed.avg.per_cycle_active', 'grid': 'launch__grid_size'}
    units = rows[1]
    acc = collections.defaultdict(lambda: collections.defaultdict(list))
    for r in rows[2:]:
        name = r[col['Kernel Name']].split('(')[0].split('::')[-1]
        for k, m in want.items():
            if m in col and r[col[m]] != '':
                v = float(r[col[m]].replace(',', ''))
                u = units[col[m]].lower()
                if k.startswith('dram_r') or k.startswith('dram_w'):
                    v *= {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1)
                if k == 'us':
                    v *= {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'usecond': 1, 'nsecond': 1e-3, 'msecond': 1e3}.get(u, 1)
                acc[name][k].append(v)
    res = {}
    for name, d in acc.items():
        n = len(d['us'])
        e = {'launches': n, 'avg_us': sum(d['us']) / n}
        if d['dram_read']:
            e['avg_dram_bytes_per_launch'] = (sum(d['dram_read']) + sum(d['dram_write'])) / n
        for k in ('tensor_pct', 'dram_pct', 'ipc', 'regs', 'grid'):
            if d[k]:
                e[k] = sum(d[k]) / len(d[k])
        res[name] = e
    json.dump({'source': rep, 'note': 'ncu --set full --clock-control none; per-launch averages; cold-cache, serialised', 'kernels': res},
              open(out, 'w'), indent=1)
    print(json.dumps(res, indent=1))


def launch_list(path, out):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.reader(lines))
    h = rows[0]
    ik, iv, iu = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
    acc = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        name = r[ik].split('(')[0].split('::')[-1]
        v = float(r[iv].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[iu], 1e-3)
        acc[name][0] += 1
        acc[name][1] += v
    tot = sum(v[1] for v in acc.values())
    with open(out, 'w') as f:
        f.write(f'# ncu launch list summary (gpu__time_duration.sum, --clock-control none), {sum(v[0] for v in acc.values())} launches\n')
        f.write('# cold-cache, serialised: compare SHARES with the bench breakdown, not absolutes\n')
        f.write(f'{"kernel":44s} {"n":>5s} {"us":>10s} {"share":>7s}\n')
        for k, (n, us) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
            f.write(f'{k:44s} {n:5d} {us:10.1f} {100 * us / tot:6.1f}%\n')
    print(open(out).read())


if __name__ == '__main__':
    {'full': full, 'list': launch_list}[sys.argv[1]](sys.argv[2], sys.argv[3])
