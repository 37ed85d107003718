#!/usr/bin/env python
"""Static SASS instruction mix of the hot kernels in libvissat_b200.so's objects, grouped by issue pipe
(B300_MICROARCH.md: FFMA/FMUL/FADD/IMAD/FFMA2 on the fma pipe; IADD3/LOP3/SHF/PRMT/FMNMX/ISETP/FSETP/SEL on the alu
pipe, both one warp instruction per 2 cycles per sub-partition; conversions and MUFU on xu; DFMA/DADD/DMUL on fp64).
    python tools/sass_mix.py > profiles/<label>_sass_mix.txt
Static counts include rarely executed paths; the executed mix per source line is in *_source_attribution.txt."""
import collections
import os
import re
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(REPO, 'vissatsatellitestereo_b200', 'csrc', 'build')
KERNELS = [('rasterize.o', 'k_unproject_scatterILi3ELi1ELb1', 'K1 k_unproject_scatter<3,1,lean>'),
           ('rasterize.o', 'k_unproject_scatterILi3ELi1ELb0', 'K1 k_unproject_scatter<3,1,full> (audited / sparse mode)'),
           ('finalize.o', 'k_grid_finalizeIjN5vsfin6NoSink', 'K2 k_grid_finalize<u32>'),
           ('finalize.o', 'k_grid_finalizeIjN5vsfin8PeerSink', 'K2 k_grid_finalize<u32, PeerSink>'),
           ('finalize.o', 'k_median3x3', 'K4 k_median3x3'),
           ('fuse.o', 'k_fuse_mediumILi52', 'K3 k_fuse_medium<52>'),
           ('fuse.o', 'k_fuse_pairILi104', 'K3 k_fuse_pair<104>'),
           ('fuse.o', 'k_fuse_largeILi8ELi52', 'K3 k_fuse_large<8,52>'),
           ('stage_ab.o', 'k_stage_abILi3ELi1EN5vsfin6NoSink', 'K1||K2 k_stage_ab<3,1> (opt-in)')]
PIPE = {'fma': ('FFMA', 'FFMA2', 'FMUL', 'FMUL2', 'FADD', 'FADD2', 'IMAD', 'HFMA2'),
        'alu': ('IADD3', 'VIADD', 'LOP3', 'SHF', 'PRMT', 'FMNMX', 'FMNMX3', 'VIMNMX', 'VIMNMX3', 'ISETP', 'FSETP', 'SEL', 'FSEL',
                'PLOP3', 'LEA', 'POPC', 'FLO', 'BREV', 'IABS', 'MOV', 'DSETP'),
        'fp64': ('DFMA', 'DADD', 'DMUL'),
        'xu': ('MUFU', 'F2F', 'F2I', 'I2F', 'F2FP', 'I2FP', 'FRND'),
        'lsu': ('LDG', 'STG', 'LDS', 'STS', 'LDL', 'STL', 'LDC', 'LDCU', 'RED', 'ATOM', 'ATOMS', 'ATOMG', 'LDGSTS', 'LDGDEPBAR',
                'DEPBAR', 'LDSM'),
        'sync/branch': ('BAR', 'BRA', 'BSSY', 'BSYNC', 'WARPSYNC', 'EXIT', 'VOTE', 'SHFL', 'CALL', 'RET', 'NOP', 'MATCH', 'REDUX',
                        'ENDCOLLECTIVE', 'YIELD')}


def main():
    for obj, key, title in KERNELS:
        sass = subprocess.run(['cuobjdump', '-sass', os.path.join(BUILD, obj)], capture_output=True, text=True).stdout
        on, ops = False, collections.Counter()
        for line in sass.split('\n'):
            if 'Function :' in line:
                on = key in line
                continue
            if not on:
                continue
            m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)', line)
            if m:
                ops[m.group(1)] += 1
        total = sum(ops.values())
        print('== {}: {} SASS instructions'.format(title, total))
        seen = set()
        for pipe, names in PIPE.items():
            items = [(n, ops[n]) for n in names if ops[n]]
            seen.update(n for n, _ in items)
            cnt = sum(c for _, c in items)
            print('  {:12s} {:5d} {:5.1f}%  {}'.format(pipe, cnt, 100.0 * cnt / max(total, 1),
                                                       ' '.join('{}:{}'.format(n, c) for n, c in sorted(items, key=lambda t: -t[1]))))
        rest = [(n, c) for n, c in ops.items() if n not in seen]
        print('  {:12s} {:5d} {:5.1f}%  {}'.format('other', sum(c for _, c in rest), 100.0 * sum(c for _, c in rest) / max(total, 1),
                                                   ' '.join('{}:{}'.format(n, c) for n, c in sorted(rest, key=lambda t: -t[1])[:12])))
        print()


if __name__ == '__main__':
    main()
