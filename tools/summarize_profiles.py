#!/usr/bin/env python
"""Turn gpurun_out/{launches_<tag>.csv, prof_*_<tag>.ncu-rep} into tracked text summaries under profiles/.
    python tools/summarize_profiles.py <tag> [<round label>]
Writes profiles/<label>_launches.txt, profiles/<label>_kernels.txt and updates profiles/traffic.json (DRAM bytes per
launch of each hot kernel from the `--set full` capture; bench.py copies it into roofline.traffic)."""
import collections
import csv
import json
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, 'gpurun_out')
PROF = os.path.join(REPO, 'profiles')

METRICS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
           'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
           'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
           'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
           'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
           'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
           'launch__waves_per_multiprocessor', 'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
           'sm__cycles_elapsed.max']


def to_us(v, u):
    v = float(v.replace(',', ''))
    return v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v * 1e6 if u == 's' else v


def to_bytes(v, u):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)


def short(name):
    m = re.search(r'(k_[a-z0-9_]+(<[^>]*>)?)', name)
    return m.group(1) if m else name[:60]


def launches(tag, label):
    path = os.path.join(OUT, 'launches_{}.csv'.format(tag))
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        k = short(r[ki])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += to_us(r[vi], r[ui])
    tot = sum(a[1] for a in agg.values())
    lines = ['# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ : python bench.py --steps 1 --warmup 1',
             '# (cold-cache, serialised launches of warm-up + timed step + e2e passes: compare SHARES, not absolutes)',
             '{:>12s} {:>9s} {:>10s} {:>7s}  kernel'.format('total_us', 'launches', 'avg_us', 'share')]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append('{:12.1f} {:9d} {:10.2f} {:6.1f}%  {}'.format(t, n, t / n, 100 * t / tot, k))
    open(os.path.join(PROF, '{}_launches.txt'.format(label)), 'w').write('\n'.join(lines) + '\n')
    return agg


def kernels(tag, label):
    traffic = {}
    lines = ['# ncu --set full --clock-control none --import-source on (one launch each, timed step of bench.py C2)']
    for rep in ('prof_k1k2_{}.ncu-rep'.format(tag), 'prof_fuse_{}.ncu-rep'.format(tag)):
        path = os.path.join(OUT, rep)
        if not os.path.exists(path):
            continue
        raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        seen = set()
        for r in rows[2:]:
            name = short(r[hdr.index('Kernel Name')])
            if name in seen:
                continue
            seen.add(name)
            lines.append('\n== {}'.format(name))
            vals = {}
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    vals[m] = (r[i], units[i])
                    lines.append('  {:70s} {:>18s} {}'.format(m, r[i][:18], units[i]))
            stall = [(hdr[i], r[i]) for i in range(len(hdr)) if 'issue_stalled' in hdr[i] and hdr[i].endswith('per_issue_active.ratio')]
            top = sorted(((float(v), h) for h, v in stall if v), reverse=True)[:6]
            lines.append('  top stalls (warps per issue-active cycle): ' + ', '.join(
                '{} {:.2f}'.format(h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v)
                for v, h in top))
            if 'dram__bytes_read.sum' in vals:
                b = to_bytes(*vals['dram__bytes_read.sum']) + to_bytes(*vals['dram__bytes_write.sum'])
                key = 'k1' if 'unproject' in name else 'k2' if 'grid_finalize' in name else 'fuse' if 'fuse' in name else 'blur'
                traffic[key] = b
                lines.append('  DRAM traffic per launch: {:.1f} MB'.format(b / 1e6))
    open(os.path.join(PROF, '{}_kernels.txt'.format(label)), 'w').write('\n'.join(lines) + '\n')
    tpath = os.path.join(PROF, 'traffic.json')
    json.dump(dict(traffic, source='{}_kernels.txt'.format(label)), open(tpath, 'w'), indent=1)


if __name__ == '__main__':
    tag = sys.argv[1]
    label = sys.argv[2] if len(sys.argv) > 2 else tag
    os.makedirs(PROF, exist_ok=True)
    launches(tag, label)
    kernels(tag, label)
    print(open(os.path.join(PROF, '{}_launches.txt'.format(label))).read())
