"""In-process A/B of the field kernel between two builds of the library (e.g. the round-1 .so and the current one) on the bench
workload, interleaved A,B,A,B on the same box so that clock / power drift cancels:  python tests/diag_ab_field.py libA.so libB.so"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from avatarcap_b200 import packer, synth  # noqa: E402


def bind(path):
    lib = C.CDLL(path)
    vp, i, i64 = C.c_void_p, C.c_int, C.c_int64
    lib.avc_ctx_create.argtypes = [i, C.POINTER(vp)]
    lib.avc_load_avatar_weights.argtypes = [vp, vp, C.c_size_t]
    lib.avc_set_feature_map.argtypes = [vp, i, vp, i, i, i, vp]
    lib.avc_eval_occupancy.argtypes = [vp, vp, i64, C.POINTER(C.c_float), vp, vp, vp, vp, i, i, vp]
    lib.avc_make_grid.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(i), i, i, vp, vp]
    lib.avc_last_error.restype = C.c_char_p; lib.avc_last_error.argtypes = [vp]
    h = vp()
    assert lib.avc_ctx_create(0, C.byref(h)) == 0
    return lib, h


def main():
    paths = sys.argv[1:3]
    torch.cuda.set_device(0)
    body = synth.SynthBody(); frame = synth.make_frame(body, None)
    blob = packer.pack_avatar(synth.avatar_state_dict())
    fmap = torch.from_numpy(synth.feature_map(64, 256, 256, synth.SEED + 4)).cuda().contiguous()
    res = (256, 256, 256); n = 256 ** 3
    pts = torch.empty((n, 3), device='cuda'); occ = torch.empty(n, device='cuda'); off = torch.empty((n, 3), device='cuda')
    rgb = torch.empty((n, 3), device='cuda'); al = torch.empty(n, device='cuda')
    b = np.asarray(frame['cano_bounds'], np.float32).reshape(6)
    c = (C.c_float * 3)(*[float(x) for x in frame['cano_smpl_center']])
    libs = []
    for p in paths:
        lib, h = bind(p)
        assert lib.avc_load_avatar_weights(h, blob, len(blob)) == 0, lib.avc_last_error(h)
        assert lib.avc_set_feature_map(h, 0, C.c_void_p(fmap.data_ptr()), 64, 256, 256, None) == 0
        libs.append((lib, h))
    assert libs[0][0].avc_make_grid(libs[0][1], (C.c_float * 6)(*b), (C.c_int * 3)(*res), 0, 256, C.c_void_p(pts.data_ptr()), None) == 0
    torch.cuda.synchronize()

    def run(k):
        lib, h = libs[k]
        rc = lib.avc_eval_occupancy(h, C.c_void_p(pts.data_ptr()), n, c, C.c_void_p(occ.data_ptr()), C.c_void_p(off.data_ptr()), C.c_void_p(rgb.data_ptr()),
                                    C.c_void_p(al.data_ptr()), 0, 0, None)
        assert rc == 0, lib.avc_last_error(h)
    outs = []
    for k in range(2):
        run(k); torch.cuda.synchronize(); outs.append(occ.clone())
    print('max |occ_A - occ_B| = %.3g' % float((outs[0] - outs[1]).abs().max()))
    times = [[], []]
    for rep in range(6):
        for k in range(2):
            a = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
            a.record(); run(k); run(k); e.record(); torch.cuda.synchronize()
            times[k].append(a.elapsed_time(e) / 2)
    for k in range(2):
        t = np.array(times[k][1:])
        print('%s: %.2f ms median (%.2f..%.2f) -> %.1f Mpts/s' % (os.path.basename(paths[k]), np.median(t), t.min(), t.max(), n / np.median(t) / 1e3))


if __name__ == '__main__':
    main()
