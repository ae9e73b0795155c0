"""ctypes binding of libavatarcap_b200.so (include/avatarcap_b200.h). Fails loudly when the library is missing:
there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libavatarcap_b200.so')

OK, EINVAL, ECUDA, ESTATE, ECAPACITY, EFORMAT, EVALUE = 0, -1, -2, -3, -4, -5, -6
IMPL_AUTO, IMPL_SIMT, IMPL_TC, IMPL_TC2 = 0, 1, 2, 3
IF_SDF, IF_OCCUPANCY = 0, 1
MAP_POSE, MAP_IMAGE = 0, 1
KIND_AVATAR, KIND_RECON = 0, 1
WEIGHT_SLOTS = 4
ABI_VERSION = 5
SHARD_HEADER_BYTES = 256
IPC_HANDLE_BYTES = 64
RASTER_CULL_BACK, RASTER_FLIP_X = 1, 2

_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
_F3 = C.c_float * 3
_F6 = C.c_float * 6
_I3 = C.c_int * 3

# name -> (restype, argtypes); this table is also what tests/test_abi.py checks against the header
SIGNATURES = {
    'avc_ctx_create': (_i, [_i, C.POINTER(_vp)]),
    'avc_ctx_destroy': (None, [_vp]),
    'avc_last_error': (C.c_char_p, [_vp]),
    'avc_abi_version': (_i, []),
    'avc_has_tensor_core_path': (_i, [_vp]),
    'avc_launch_count': (_i64, [_vp]),
    'avc_reset_launch_count': (None, [_vp]),
    'avc_debug_set_trace': (_i, [_vp, _vp, _i]),
    'avc_load_avatar_weights': (_i, [_vp, _vp, C.c_size_t]),
    'avc_load_recon_weights': (_i, [_vp, _vp, C.c_size_t]),
    'avc_load_weights_slot': (_i, [_vp, _i, _i, _vp, C.c_size_t]),
    'avc_select_weights': (_i, [_vp, _i, _i]),
    'avc_set_feature_map': (_i, [_vp, _i, _vp, _i, _i, _i, _vp]),
    'avc_set_feature_map_hwc': (_i, [_vp, _i, _vp, _i, _i, _i, _vp]),
    'avc_encoder_create': (_i, [_vp, _vp, _i64, _vp, C.c_size_t, _vp, _i64, C.POINTER(_vp)]),
    'avc_encoder_run': (_i, [_vp, _vp, _vp, _i, _vp]),
    'avc_encoder_shape': (_i, [_vp, C.POINTER(_i), C.POINTER(_i)]),
    'avc_encoder_read_buffer': (_i, [_vp, _i, _vp, _i64, _vp]),
    'avc_encoder_destroy': (None, [_vp]),
    'avc_eval_occupancy': (_i, [_vp, _vp, _i64, C.POINTER(_f), _vp, _vp, _vp, _vp, _i, _i, _vp]),
    'avc_eval_occupancy_grid': (_i, [_vp, C.POINTER(_f), C.POINTER(_i), _i, _i, C.POINTER(_f), _vp, _vp, _vp, _vp, _i, _i, _vp]),
    'avc_eval_recon_grid': (_i, [_vp, C.POINTER(_f), C.POINTER(_i), _i, _i, C.POINTER(_f), _vp, _i, _vp]),
    'avc_eval_warp': (_i, [_vp, _vp, _i64, C.POINTER(_f), _vp, _i, _vp]),
    'avc_eval_template': (_i, [_vp, _vp, _i64, _vp, _vp, _vp, _i, _i, _vp]),
    'avc_eval_recon': (_i, [_vp, _vp, _i64, C.POINTER(_f), _vp, _i, _vp]),
    'avc_eval_occupancy_host': (_i, [_vp, _vp, _i64, C.POINTER(_f), _vp, _vp, _vp, _vp, _i, _i]),
    'avc_eval_recon_host': (_i, [_vp, _vp, _i64, C.POINTER(_f), _vp, _i]),
    'avc_make_grid': (_i, [_vp, C.POINTER(_f), C.POINTER(_i), _i, _i, _vp, _vp]),
    'avc_scatter_fill': (_i, [_vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    'avc_mc_count': (_i, [_vp, _vp, C.POINTER(_i), _f, _i, _i, C.POINTER(_i64), C.POINTER(_i64), _vp]),
    'avc_mc_emit': (_i, [_vp, _vp, C.POINTER(_i), C.POINTER(_f), _f, _i, _i, _i, _i, _vp, _vp, _vp, _i64, _i64, _vp]),
    'avc_mc_extract': (_i, [_vp, _vp, C.POINTER(_i), C.POINTER(_f), _f, _i, _i, _i, _i, _vp, _vp, _vp, _i64, _i64, _vp, _vp]),
    'avc_mc_emit_counted': (_i, [_vp, _vp, C.POINTER(_i), C.POINTER(_f), _f, _i, _i, _i, _i, _vp, _vp, _vp, _i64, _i64, _vp]),
    'avc_shard_alloc': (_i, [_vp, C.c_size_t, C.POINTER(_vp), _vp]),
    'avc_shard_open': (_i, [_vp, _vp, C.POINTER(_vp)]),
    'avc_shard_close': (_i, [_vp, _vp]),
    'avc_shard_free': (_i, [_vp, _vp]),
    'avc_halo_push': (_i, [_vp, _vp, _vp, _vp, _i64, _i64, _i, _i, _i64, _i, _i64, C.c_uint64, _vp]),
    'avc_halo_wait': (_i, [_vp, _vp, _i, _i, C.c_uint64, _vp]),
    'avc_halo_ack': (_i, [_vp, _vp, _vp, C.c_uint64, _vp]),
    'avc_renumber_faces': (_i, [_vp, _vp, _i64, _i64, _i64, _i64, _vp]),
    'avc_knn': (_i, [_vp, _vp, _i64, _vp, _i, _i, _vp, _vp, _vp]),
    'avc_near_flag': (_i, [_vp, _vp, _i64, _vp, _i, C.c_double, _vp, _vp]),
    'avc_lbs_weights': (_i, [_vp, _vp, _i64, _vp, _i, _vp, _vp, _vp]),
    'avc_skin_points': (_i, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    'avc_skin_normals': (_i, [_vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    'avc_skin_mesh': (_i, [_vp, _vp, _vp, _i64, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    'avc_inside_volume': (_i, [_vp, _vp, _i, _vp, _i, C.POINTER(_f), C.POINTER(_i), _vp, _vp]),
    'avc_ray_samples': (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i, _vp, _vp, _vp, _vp]),
    'avc_nerf_raw': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(_f), _i64, _vp, _vp]),
    'avc_composite': (_i, [_vp, _vp, _vp, _i64, _i, _i, _vp, _vp, _vp, _vp]),
    'avc_rasterize': (_i, [_vp, _vp, _i64, _vp, _i64, _vp, C.POINTER(_f), _i, _i, C.POINTER(_f), _i, _i, _vp, _vp]),
    'avc_canonicalize_normals': (_i, [_vp, _vp, _vp, _i64, C.POINTER(_f), _f, _f, _f, _f, _vp, _i, _vp, _i, _i, _vp, _vp]),
    'avc_posed_to_cano': (_i, [_vp, _vp, _i64, _vp, _i, _vp, _vp, C.POINTER(_f), _vp, C.POINTER(_i), _vp, _vp, _vp]),
}

_lib = None


class AvcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__('avatarcap_b200 error %d: %s' % (code, msg))
        self.code = code


def load() -> C.CDLL:
    """Load the shared library (once). Raises ImportError with build instructions when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get('AVC_LIB_PATH') or LIB_PATH          # A/B of two builds of the library (tests/diag_*): not a fallback, it must exist
    if not os.path.exists(path):
        raise ImportError('%s not found: build it with `python -c "import __graft_entry__ as g; g.build()"` or '
                          '`make -C avatarcap_b200/csrc` (needs nvcc, sm_100a). There is no CPU fallback.' % path)
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def f3(v) -> C.Array:
    return _F3(*[float(x) for x in v])


def f6(bounds) -> C.Array:
    b = [float(x) for x in list(bounds[0]) + list(bounds[1])] if len(bounds) == 2 else [float(x) for x in bounds]
    return _F6(*b)


def i3(v) -> C.Array:
    return _I3(*[int(x) for x in v])
