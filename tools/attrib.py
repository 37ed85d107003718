#!/usr/bin/env python
"""Attribute executed SASS instructions of one kernel (ncu report with --import-source) to CUDA source lines.
    python tools/attrib.py <report.ncu-rep> <kernel regex> <object file> <mangled-name substring> [units]
Uses nvdisasm -g line info of the object file; rows are matched by instruction order."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, kregex, obj, mangled = sys.argv[1:5]
    units = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
    dis = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split('\n')
    start = [i for i, l in enumerate(dis) if l.startswith('.text.') and mangled in l][0]
    sass, cur = [], None
    for l in dis[start + 1:]:
        if l.startswith('//--------------------- .text.'):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m:
            sass.append((m.group(2).strip(), cur))
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kregex],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[1]
    si, ei, wi = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
    end = [i for i, r in enumerate(rows) if i > 2 and r and r[0] == 'Kernel Name']
    body = [r for r in (rows[2:end[0]] if end else rows[2:]) if len(r) > ei and r[ei].isdigit()]
    print('static sass', len(sass), 'ncu rows', len(body))
    agg, stall = collections.Counter(), collections.Counter()
    tot = tots = 0
    for k in range(min(len(sass), len(body))):
        n = int(body[k][ei])
        s = int(body[k][wi]) if body[k][wi].isdigit() else 0
        agg[sass[k][1]] += n
        stall[sass[k][1]] += s
        tot += n
        tots += s
    print('total warp instr', tot, ' per unit', tot * 32 / units if units != 1 else '')
    cache = {}
    for (key, c) in agg.most_common(45):
        if key is None:
            continue
        f, l = key
        if f not in cache:
            try:
                cache[f] = open(f).read().split('\n')
            except Exception:
                cache[f] = []
        text = cache[f][l - 1].strip() if l - 1 < len(cache[f]) else ''
        print('{:9d} {:5.1f}% {:7.2f}/unit stall {:4.1f}%  {}:{}  {}'.format(
            c, 100 * c / tot, c * 32 / units, 100 * stall[key] / max(tots, 1), os.path.basename(f), l, text[:80]))


if __name__ == '__main__':
    main()
