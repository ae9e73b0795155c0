"""GPU diagnostic (not a pytest): where does the host-buffer entry lose time against the device-resident launch?
Times (a) one launch over 256^3 points, (b) the same points as 8 back-to-back chunk launches, no copies,
(c) H2D / D2H alone, (d) avc_eval_occupancy_host. Usage: python tests/diag_e2e.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avatarcap_b200 import synth  # noqa: E402
from avatarcap_b200.engine import Engine  # noqa: E402


def ev_ms(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    eng = Engine()
    body = synth.SynthBody(); frame = synth.make_frame(body, None)
    eng.load_avatar(synth.avatar_state_dict()); eng.set_pose_feature_map(synth.feature_map(64, 256, 256, synth.SEED + 4))
    pts = eng.make_grid(frame['cano_bounds'], (256, 256, 256)); n = pts.shape[0]; c = frame['cano_smpl_center']
    print('one launch            %.2f ms' % ev_ms(lambda: eng.eval_occupancy(pts, c, want_offsets=True, want_texture=True)))
    ch = 1 << 21
    print('8 chunk launches      %.2f ms' % ev_ms(lambda: [eng.eval_occupancy(pts[i:i + ch], c, want_offsets=True, want_texture=True) for i in range(0, n, ch)]))
    ph = torch.empty((n, 3), dtype=torch.float32).pin_memory(); ph.copy_(pts.cpu())
    oh = torch.empty((n, 8), dtype=torch.float32).pin_memory(); od = torch.empty((n, 8), dtype=torch.float32, device='cuda')
    print('H2D 201 MB alone      %.2f ms' % ev_ms(lambda: pts.copy_(ph, non_blocking=True)))
    print('D2H 537 MB alone      %.2f ms' % ev_ms(lambda: oh.copy_(od, non_blocking=True)))
    occ = np.asarray(torch.empty(n, dtype=torch.float32).pin_memory().numpy()); off = torch.empty((n, 3), dtype=torch.float32).pin_memory().numpy()
    rgb = torch.empty((n, 3), dtype=torch.float32).pin_memory().numpy(); al = torch.empty(n, dtype=torch.float32).pin_memory().numpy()
    p = ph.numpy()
    for label, outs in (('host entry, all outputs', (occ, off, rgb, al)), ('host entry, occ only   ', (occ, None, None, None))):
        eng.eval_occupancy_host(p, c, *outs)
        t0 = time.perf_counter()
        for _ in range(3):
            eng.eval_occupancy_host(p, c, *outs)
        torch.cuda.synchronize()
        print('%s %.2f ms' % (label, (time.perf_counter() - t0) / 3 * 1e3))
    # kernel under concurrent D2H traffic on another stream
    s2 = torch.cuda.Stream()
    def both():
        with torch.cuda.stream(s2):
            oh.copy_(od, non_blocking=True)
        eng.eval_occupancy(pts, c, want_offsets=True, want_texture=True)
    print('one launch + concurrent 537 MB D2H on a side stream  %.2f ms' % ev_ms(both))
    eng.close()


if __name__ == '__main__':
    main()
