"""ctypes binding of libvissat_b200.so (C ABI: include/vissat_b200.h).

The library is the only compute backend.  If it has not been built (``python -c 'import __graft_entry__ as g;
g.build()'`` or ``make -C vissatsatellitestereo_b200/csrc``) importing this module raises ImportError; if there is
no CUDA device, creating a context raises VisSatError.  Nothing here falls back to a CPU implementation.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libvissat_b200.so')

VS_NUM_STATS = 4
STAT_VALID, STAT_INGRID, STAT_AMBIGUOUS, STAT_EXACT = 0, 1, 2, 3
ABI_VERSION = 3


class VisSatError(RuntimeError):
    pass


class vs_aoi(C.Structure):
    _fields_ = [('lat0', C.c_double), ('lon0', C.c_double), ('alt0', C.c_double),
                ('zone', C.c_int32), ('south', C.c_int32),
                ('ul_e', C.c_double), ('ul_n', C.c_double),
                ('row_res', C.c_double), ('col_res', C.c_double),
                ('xsize', C.c_int32), ('ysize', C.c_int32),
                ('alt_lo', C.c_double), ('alt_hi', C.c_double)]


class vs_fit_info(C.Structure):
    _fields_ = [('degree', C.c_int32), ('n_terms', C.c_int32), ('mixed', C.c_int32), ('reserved', C.c_int32),
                ('max_err_cells', C.c_double), ('max_err_alt_m', C.c_double),
                ('box_center', C.c_double * 3), ('box_half', C.c_double * 3)]


VS_MAX_RANKS = 16
VS_IPC_HANDLE_BYTES = 64


VS_TILE_W = 64
VS_TILE_H = 32


class vs_exchange(C.Structure):
    _fields_ = [('n_ranks', C.c_int32), ('rank', C.c_int32), ('halo', C.c_int32), ('occ_words', C.c_int32),
                ('view0', C.c_int64), ('n_views_total', C.c_int64), ('local_stack', C.c_void_p),
                ('band_stack', C.c_void_p * VS_MAX_RANKS), ('occ', C.c_void_p * VS_MAX_RANKS)]


_vp = C.c_void_p
_i32 = C.c_int32
_i64 = C.c_int64
_dbl = C.c_double

# every symbol include/vissat_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    'vs_abi_version': (C.c_int, []),
    'vs_last_error': (C.c_char_p, []),
    'vs_device_count': (C.c_int, [C.POINTER(C.c_int)]),
    'vs_ctx_create': (C.c_int, [C.c_int, C.POINTER(_vp)]),
    'vs_ctx_destroy': (C.c_int, [_vp]),
    'vs_set_aoi': (C.c_int, [_vp, C.POINTER(vs_aoi), C.c_int, C.POINTER(vs_fit_info)]),
    'vs_set_ambiguity_eps': (C.c_int, [_vp, _dbl]),
    'vs_fit_eval': (C.c_int, [_vp, C.POINTER(_dbl), _i64, C.POINTER(_dbl), C.POINTER(_dbl), C.POINTER(_dbl)]),
    'vs_unproject_rasterize': (C.c_int, [_vp, _vp, _i32, _i32, C.POINTER(_dbl), _vp, C.c_int, _vp, _vp, _vp]),
    'vs_keygrid_clear': (C.c_int, [_vp, _vp, _i64, C.c_int, _vp]),
    'vs_points_rasterize': (C.c_int, [_vp, _vp, _i64, _dbl, _dbl, _dbl, _dbl, _i32, _i32, _vp, C.c_int, _vp, _vp]),
    'vs_grid_finalize': (C.c_int, [_vp, _vp, _i32, _i32, _vp, C.c_int, _vp, _vp]),
    'vs_grid_finalize64': (C.c_int, [_vp, _vp, _i32, _i32, _vp, _vp, C.c_int, _vp]),
    'vs_views_to_dsm': (C.c_int, [_vp, _i32, C.POINTER(_vp), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_dbl), _vp, _vp,
                                  _i64, C.c_int, _vp, _vp, _vp]),
    'vs_set_streams': (C.c_int, [_vp, C.c_int]),
    'vs_set_coschedule': (C.c_int, [_vp, C.c_int]),
    'vs_set_timing': (C.c_int, [_vp, C.c_int]),
    'vs_get_timing': (C.c_int, [_vp, _i32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(_i32)]),
    'vs_fuse_views': (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    'vs_median3x3': (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, C.c_int, _vp, _vp]),
    'vs_enu_to_geodetic': (C.c_int, [_vp, _vp, _vp, _vp, _i64, _dbl, _dbl, _dbl, _vp, _vp, _vp, _vp]),
    'vs_geodetic_to_enu': (C.c_int, [_vp, _vp, _vp, _vp, _i64, _dbl, _dbl, _dbl, _vp, _vp, _vp, _vp]),
    'vs_geodetic_to_utm': (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    'vs_utm_to_geodetic': (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    'vs_enu_to_utm': (C.c_int, [_vp, _vp, _vp, _vp, _i64, _dbl, _dbl, _dbl, _i32, _i32, _vp, _vp, _vp, _vp]),
    'vs_launch_count': (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    'vs_peer_alloc': (C.c_int, [_vp, C.c_uint64, C.POINTER(_vp), C.POINTER(C.c_uint8)]),
    'vs_peer_open': (C.c_int, [_vp, C.POINTER(C.c_uint8), C.POINTER(_vp)]),
    'vs_peer_close': (C.c_int, [_vp, _vp]),
    'vs_peer_free': (C.c_int, [_vp, _vp]),
    'vs_set_exchange': (C.c_int, [_vp, C.POINTER(vs_exchange)]),
    'vs_set_occupancy': (C.c_int, [_vp, _vp, _i32, _vp, _i64]),
    'vs_fuse_views_sparse': (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _vp]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            'libvissat_b200.so is not built ({}). Build it with `make -C vissatsatellitestereo_b200/csrc` or '
            '`python -c "import __graft_entry__ as g; g.build()"`. There is no CPU fallback.'.format(LIB_PATH))
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    if lib.vs_abi_version() != ABI_VERSION:
        raise ImportError('libvissat_b200.so ABI {} != binding ABI {}; rebuild'.format(lib.vs_abi_version(), ABI_VERSION))
    return lib


lib = _load()


def check(rc, what=''):
    if rc != 0:
        msg = lib.vs_last_error()
        raise VisSatError('{} failed (code {}): {}'.format(what or 'libvissat_b200 call', rc,
                                                          msg.decode('utf-8', 'replace') if msg else ''))


def device_count():
    n = C.c_int(0)
    rc = lib.vs_device_count(C.byref(n))
    if rc != 0:
        return 0
    return n.value


class Context:
    """Owner of one vs_ctx (one CUDA device)."""

    def __init__(self, device=0):
        self._h = _vp()
        check(lib.vs_ctx_create(int(device), C.byref(self._h)), 'vs_ctx_create')
        self.device = int(device)
        self.fit = None

    def close(self):
        if getattr(self, '_h', None) is not None and self._h:
            lib.vs_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        if not self._h:
            raise VisSatError('context is closed')
        return self._h

    def set_aoi(self, aoi_struct, max_degree=5):
        info = vs_fit_info()
        check(lib.vs_set_aoi(self.handle, C.byref(aoi_struct), int(max_degree), C.byref(info)), 'vs_set_aoi')
        self.fit = {'degree': info.degree, 'n_terms': info.n_terms, 'mixed': info.mixed, 'max_err_cells': info.max_err_cells,
                    'max_err_alt_m': info.max_err_alt_m, 'box_center': list(info.box_center),
                    'box_half': list(info.box_half)}
        return self.fit

    def set_ambiguity_eps(self, eps):
        check(lib.vs_set_ambiguity_eps(self.handle, float(eps)), 'vs_set_ambiguity_eps')

    def launch_count(self):
        n = C.c_uint64(0)
        check(lib.vs_launch_count(self.handle, C.byref(n)), 'vs_launch_count')
        return n.value
