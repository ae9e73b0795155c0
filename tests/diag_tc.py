"""GPU diagnostic (not a pytest): tcgen05 kernel vs the fp32 SIMT kernel, stage by stage, with error statistics.
Usage: python tests/diag_tc.py [n_points]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from helpers import golden_scene, load_golden  # noqa: E402
from avatarcap_b200.engine import Engine  # noqa: E402


def stats(name, a, b):
    a = a.detach().cpu().numpy().astype(np.float64); b = b.detach().cpu().numpy().astype(np.float64)
    d = np.abs(a - b)
    bad = ~np.isfinite(a)
    print('%-28s max|d|=%.3e mean|d|=%.3e  ref range [%.3g, %.3g]  nonfinite=%d  corr=%.6f' % (
        name, np.nanmax(d), np.nanmean(d), b.min(), b.max(), int(bad.sum()),
        np.corrcoef(np.nan_to_num(a).ravel(), b.ravel())[0, 1] if a.size > 2 else 1.0))
    return float(np.nanmax(d))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    impl = sys.argv[2] if len(sys.argv) > 2 else 'tc'
    eng = Engine()
    print('device', torch.cuda.get_device_name(0), 'tc path:', eng.has_tensor_core_path)
    s = golden_scene(); g = load_golden('avatar_golden.npz')
    eng.load_avatar(s['avatar_sd']); eng.load_recon(s['recon_sd'])
    eng.set_pose_feature_map(s['pose_map']); eng.set_image_feature_map(s['image_map'])
    pts = torch.from_numpy(np.tile(g['pts'], (max(1, n // len(g['pts']) + 1), 1))[:n]).cuda()
    c = g['center']
    for name, fn in (
        ('template rgb/alpha/occ', lambda impl: eng.eval_template(pts, impl=impl)),
        ('warp offsets', lambda impl: (eng.eval_warp(pts, c, impl=impl),)),
        ('recon ov', lambda impl: (eng.eval_recon(pts, c, impl=impl),)),
        ('query occ/off', lambda impl: tuple(eng.eval_occupancy(pts, c, impl=impl)[k] for k in ('occ', 'off'))),
        ('query+texture', lambda impl: tuple(eng.eval_occupancy(pts, c, want_texture=True, impl=impl)[k] for k in ('occ', 'off', 'rgb', 'alpha'))),
    ):
        try:
            ref = fn('simt'); torch.cuda.synchronize()
            out = fn(impl); torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print('%-28s FAILED: %r' % (name, e))
            break
        for i, (o, r) in enumerate(zip(out, ref)):
            stats('%s[%d]' % (name, i), o, r)
        if name.startswith('template'):
            print('   first occ tc  :', out[2][:6].cpu().numpy())
            print('   first occ simt:', ref[2][:6].cpu().numpy())
    # determinism / tile independence
    a = eng.eval_occupancy(pts, c, impl=impl)['occ']; b = eng.eval_occupancy(pts[:n // 2 + 7], c, impl=impl)['occ']
    print('tile independence: max|d| =', float((a[:n // 2 + 7] - b).abs().max()))
    eng.close()


if __name__ == '__main__':
    main()
