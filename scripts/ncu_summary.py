"""Turns gpurun_out/*.ncu-rep + launches_*.csv of scripts/profile_final.sh into the text summaries
committed under profiles/ (and profiles/roofline_traffic.json, read by bench.py for roofline.traffic)."""
import collections
import csv
import io
import json
import subprocess
import sys

TAG = sys.argv[1] if len(sys.argv) > 1 else "r1j"
KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    return r[0], r[1], r[2]


def ops(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, data = rows[1], rows[2:]
    ia, ii, isamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    tot = sum(int(r[ii]) for r in data)
    ts = sum(int(r[isamp]) for r in data)
    op, sm = collections.Counter(), collections.Counter()
    for r in data:
        t = r[ia].split()
        o = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        op[o] += int(r[ii])
        sm[o] += int(r[isamp])
    return tot, [(k, 100 * v / tot, 100 * sm[k] / max(ts, 1)) for k, v in op.most_common(20)]


def to_bytes(v, u):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]


def main():
    traffic = None
    for name, note in (('walk', 'k_walk<false>, N=1e7 all-active step (relative criterion)'),
                       ('pass1', 'k_pass1_group, all-active launch: 1e7 queries in ~346k warp groups')):
        rep = f'gpurun_out/prof_{name}_{TAG}.ncu-rep'
        h, u, v = raw(rep)
        with open(f'profiles/{name}_{TAG}_ncu_summary.txt', 'w') as f:
            f.write(f'# {note}\n# command: python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e ; ncu --set full --clock-control none --import-source on\n# source: {rep}\n')
            for i, k in enumerate(h):
                if k in KEEP:
                    f.write(f'{k:75s} {v[i]:>20s} {u[i]}\n')
            tot, o = ops(rep)
            f.write(f'\n# SASS opcode mix (instructions executed = {tot}); columns: opcode, % of instructions, % of stall samples\n')
            for k, a, b in o:
                f.write(f'{k:10s} {a:6.1f} {b:6.1f}\n')
        if name == 'walk':
            d = dict(zip(h, zip(v, u)))
            traffic = to_bytes(*d['dram__bytes_read.sum']) + to_bytes(*d['dram__bytes_write.sum'])
    lines = [l for l in open(f'gpurun_out/launches_{TAG}.csv') if not l.startswith('==')]
    open(f'profiles/launches_{TAG}.csv', 'w').writelines(lines)
    r = list(csv.DictReader(lines))
    names = [row['Kernel Name'] for row in r]

    def ms(row):
        x = float(row['Metric Value'].replace(',', ''))
        return x / 1e3 if row['Metric Unit'] == 'us' else x / 1e6 if row['Metric Unit'] == 'ns' else x
    pred = [i for i, n in enumerate(names) if n.startswith('k_predict')]
    adv = [i for i, n in enumerate(names) if n.startswith('k_advance')]
    a, b = pred[-1], adv[-1]
    tot = collections.defaultdict(lambda: [0, 0.0])
    for row in r[a:b + 1]:
        k = row['Kernel Name'].split('(')[0][:70]
        tot[k][0] += 1
        tot[k][1] += ms(row)
    T = sum(x[1] for x in tot.values())
    with open(f'profiles/launches_{TAG}_summary.txt', 'w') as f:
        f.write('# one all-active step (k_predict .. k_advance) of: ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e (N=1e7)\n')
        f.write(f'# {b - a + 1} launches, {T:.2f} ms summed (cold-cache, serialised: compare shares)\n')
        for k, x in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f'{x[1]:8.3f} ms {x[0]:4d}x {100 * x[1] / T:5.1f}%  {k}\n')
    json.dump({"kernel": "k_walk", "particles": 10000000, "traffic_bytes_per_launch": traffic,
               "source": f"profiles/walk_{TAG}_ncu_summary.txt (dram__bytes_read.sum + dram__bytes_write.sum, one ncu --set full capture)"},
              open('profiles/roofline_traffic.json', 'w'), indent=1)
    print(open(f'profiles/launches_{TAG}_summary.txt').read()[:700])


if __name__ == '__main__':
    main()
