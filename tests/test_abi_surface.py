"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header declares,
and the Python mirror keeps the reference's module paths and signatures.  No compute calls (no GPU here)."""
import inspect
import os
import re
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def native():
    so = os.path.join(REPO, 'vissatsatellitestereo_b200', 'libvissat_b200.so')
    if not os.path.exists(so):
        subprocess.run([sys.executable, '-c', 'import __graft_entry__ as g; g.build()'], cwd=REPO, check=True)
    from vissatsatellitestereo_b200 import _native
    return _native


def test_header_symbols_all_exported_and_bound(native):
    hdr = open(os.path.join(REPO, 'include', 'vissat_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(vs_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 18
    assert declared == set(native.SIGNATURES), declared ^ set(native.SIGNATURES)
    for name in declared:
        assert hasattr(native.lib, name), name
    assert native.lib.vs_abi_version() == native.ABI_VERSION


def test_library_is_sm100a_only(native):
    out = subprocess.run(['cuobjdump', '-lelf', native.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r'sm_(\d+a?)', out))
    assert archs == {'100a'}, archs


def test_no_gpu_means_loud_failure(native):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(native.VisSatError):
        native.Context(0)
    from vissatsatellitestereo_b200 import engine
    with pytest.raises(native.VisSatError):
        engine.require_cuda()


def test_bench_without_gpu_fails_loudly_and_prints_no_number():
    """bench.py's own arm must not fall back to anything when there is no CUDA device: non-zero exit, no JSON line."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    r = subprocess.run([sys.executable, os.path.join(REPO, 'bench.py'), '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert 'no CUDA device' in r.stderr
    assert not any(line.startswith('{') for line in r.stdout.splitlines())


def test_product_never_imports_oracle():
    pkg = os.path.join(REPO, 'vissatsatellitestereo_b200')
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), os.path.join(root, f)


def test_reference_signatures_kept(native):
    """Same names / positional parameters / defaults as the reference modules."""
    from vissatsatellitestereo_b200 import coordinate_system
    from vissatsatellitestereo_b200.colmap import read_dense
    from vissatsatellitestereo_b200.lib import latlon_utm_converter, latlonalt_enu_converter, proj_to_grid

    def params(fn):
        return [(p.name, p.default) for p in inspect.signature(fn).parameters.values()]
    E = inspect.Parameter.empty
    assert params(proj_to_grid.proj_to_grid) == [('points', E), ('xoff', E), ('yoff', E), ('xresolution', E),
                                                 ('yresolution', E), ('xsize', E), ('ysize', E), ('propagate', False)]
    assert [p[0] for p in params(latlonalt_enu_converter.latlonalt_to_enu)] == ['lat', 'lon', 'alt', 'lat0', 'lon0', 'alt0']
    assert [p[0] for p in params(latlonalt_enu_converter.enu_to_latlonalt)] == ['e', 'n', 'u', 'lat0', 'lon0', 'alt0']
    assert [p[0] for p in params(latlon_utm_converter.latlon_to_eastnorh)] == ['lat', 'lon']
    assert [p[0] for p in params(latlon_utm_converter.eastnorth_to_latlon)] == ['east', 'north', 'zone_number', 'hemisphere']
    assert [p[0] for p in params(coordinate_system.local_to_global)] == ['work_dir', 'xx', 'yy', 'zz']
    assert [p[0] for p in params(coordinate_system.global_to_local)] == ['work_dir', 'xx', 'yy', 'zz']
    assert [p[0] for p in params(read_dense.read_array)] == ['path']


def test_read_array_matches_reference_golden(golden, tmp_path):
    import numpy as np
    from vissatsatellitestereo_b200.colmap.read_dense import read_array, read_array_hw
    p = tmp_path / 'a.bin'
    p.write_bytes(golden['read_array_file'].tobytes())
    got = read_array(str(p))
    assert got.dtype == np.float32 and np.array_equal(got, golden['read_array_out'])
    assert np.array_equal(read_array_hw(str(p)), golden['read_array_out'])


def test_header_is_plain_c_and_struct_layouts_match_ctypes(native, tmp_path):
    """include/vissat_b200.h must compile as C (the boundary is a C ABI, no C++ or CUDA types) and the ctypes mirrors of
    its structs must have the same size and field offsets."""
    import ctypes as C
    import shutil
    import subprocess
    gcc = shutil.which('gcc')
    if gcc is None:
        pytest.skip('no gcc')
    structs = {'vs_aoi': native.vs_aoi, 'vs_fit_info': native.vs_fit_info, 'vs_exchange': native.vs_exchange}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "vissat_b200.h"', 'int main(void) {']
    for name, cls in structs.items():
        lines.append('  printf("{0} size %zu\\n", sizeof({0}));'.format(name))
        for field, _ in cls._fields_:
            lines.append('  printf("{0} {1} %zu\\n", offsetof({0}, {1}));'.format(name, field))
    lines.append('  printf("VS_MAX_RANKS %d\\nVS_IPC_HANDLE_BYTES %d\\nVS_NUM_STATS %d\\n", VS_MAX_RANKS, VS_IPC_HANDLE_BYTES, '
                 'VS_NUM_STATS);')
    lines += ['  return 0;', '}']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(lines) + '\n')
    exe = str(tmp_path / 'layout')
    subprocess.run([gcc, '-std=c99', '-Wall', '-Werror', '-pedantic', '-I', os.path.join(REPO, 'include'), str(src), '-o', exe],
                   check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split('\n')
    got = {tuple(l.split()[:-1]): int(l.split()[-1]) for l in out if l.strip()}
    for name, cls in structs.items():
        assert got[(name, 'size')] == C.sizeof(cls), name
        for field, _ in cls._fields_:
            assert got[(name, field)] == getattr(cls, field).offset, (name, field)
    assert got[('VS_MAX_RANKS',)] == native.VS_MAX_RANKS
    assert got[('VS_IPC_HANDLE_BYTES',)] == native.VS_IPC_HANDLE_BYTES
    assert got[('VS_NUM_STATS',)] == native.VS_NUM_STATS
