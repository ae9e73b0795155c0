"""GPU diagnostic (not a pytest): per-op timeline of the tcgen05 kernel (CTA 0), from the avc_debug_set_trace hook.
Usage: python tests/diag_trace.py [texture(0/1)]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from helpers import tpose_scene  # noqa: E402
from avatarcap_b200.engine import Engine  # noqa: E402


def main():
    tex = bool(int(sys.argv[1])) if len(sys.argv) > 1 else True
    flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    impl = sys.argv[3] if len(sys.argv) > 3 else 'tc'
    eng = Engine()
    s = tpose_scene(256)
    eng.load_avatar(s['avatar_sd']); eng.set_pose_feature_map(s['pose_map'])
    fr = s['frame']
    pts = eng.make_grid(fr['cano_bounds'], (128, 128, 128))
    buf = torch.zeros(4 * 24 * 8, dtype=torch.int64, device=eng.device)
    eng.eval_occupancy(pts, fr['cano_smpl_center'], want_texture=tex, impl=impl); torch.cuda.synchronize()    # warm
    eng.lib.avc_debug_set_trace(eng._h, C.c_void_p(buf.data_ptr()), flags)
    eng.eval_occupancy(pts, fr['cano_smpl_center'], want_texture=tex, impl=impl); torch.cuda.synchronize()
    eng.lib.avc_debug_set_trace(eng._h, None, 0)
    t = buf.cpu().numpy().reshape(4, 24, 8)
    n_ops = 20 if tex else 17
    for tile in (1,):
        base = t[tile, 0, 0]
        print('tile %d (cycles relative to the tile\'s first MMA op start; tile period %d)' % (tile, t[tile + 1, 0, 0] - base if tile < 3 else -1))
        print(' op |  mma_start  issued_h0  issued_h1 | d0_seen  h0_done  d1_seen  epi_done | op period')
        for oi in range(n_ops):
            e = t[tile, oi]
            rel = [(int(x - base) if x else -1) for x in e]
            nxt = t[tile, oi + 1, 0] if oi + 1 < n_ops else (t[tile + 1, 0, 0] if tile < 3 else 0)
            print(' %2d | %9d %9d %9d | %8d %8d %8d %8d | %6d' % (oi, rel[0], rel[1], rel[2], rel[3], rel[4], rel[5], rel[6], int(nxt - e[0]) if nxt else -1))
    eng.close()


if __name__ == '__main__':
    main()
