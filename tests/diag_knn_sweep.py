"""Sweep of the KNN grid knobs (AVC_KNN_CELL, AVC_KNN_RMAX; read per call) on three vertex sets: the masked-frame noise mesh (1.6 M vertices
within 10 cm of the body), the dense noise mesh (6.8 M vertices filling the bounding box) and a smooth body surface; plus the validity flag.
    python tests/diag_knn_sweep.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from avatarcap_b200 import pipeline, synth  # noqa: E402
from avatarcap_b200.engine import Engine  # noqa: E402

eng = Engine(); dev = eng.device
body = synth.SynthBody(); fr = synth.make_frame(body)
res = (256, 256, 256)
eng.load_avatar(synth.avatar_state_dict()); eng.set_pose_feature_map(synth.feature_map(64, 256, 256, synth.SEED + 4))
center = fr['cano_smpl_center']
cv = torch.from_numpy(fr['cano_smpl_v']).to(dev); sw = torch.from_numpy(fr['smpl_skinning_weights']).to(dev); jm = torch.from_numpy(fr['cano2live_jnt_mats']).to(dev)
pts = eng.make_grid(fr['cano_bounds'], res)
flag = pipeline.valid_points_flag(eng, pts, cv)
dense = eng.eval_occupancy_grid(fr['cano_bounds'], res, center, want_offsets=False)['occ'].reshape(res)
fill = torch.from_numpy(2.0 * synth.body_inside(pts[~flag].cpu().numpy(), synth.cano_pose()).astype(np.float32) - 1.0).to(dev)
masked = eng.scatter_fill(flag, dense.reshape(-1)[flag], fill).reshape(res)
smooth = torch.from_numpy(synth.body_sdf(pts.cpu().numpy(), synth.cano_pose())).to(dev).reshape(res)
meshes = {}
for name, vol in (('masked noise frame', masked), ('dense noise', dense), ('smooth body', smooth)):
    v, f, n = eng.extract_mesh(vol, fr['cano_bounds'], 0.0)
    meshes[name] = (v.clone(), n.clone())
    print('%s: %d vertices' % (name, v.shape[0]))


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))


ref = {k: eng.skin_mesh(v, n, cv, sw, jm)[0].clone() for k, (v, n) in meshes.items()}
print('%-6s %-5s | %s | near_flag' % ('cell', 'rmax', ' | '.join('%-18s' % k for k in meshes)))
quick = len(sys.argv) > 1 and sys.argv[1] == 'quick'          # only the default setting (A/B of two library builds through AVC_LIB_PATH)
for cell in (('0.04',) if quick else ('0.03', '0.04', '0.06')):
    for rmax in (('4',) if quick else ('1', '2', '3', '4', '6')):
        os.environ['AVC_KNN_CELL'] = cell; os.environ['AVC_KNN_RMAX'] = rmax
        row = []
        for k, (v, n) in meshes.items():
            t = timed(lambda: eng.skin_mesh(v, n, cv, sw, jm))
            same = torch.equal(eng.skin_mesh(v, n, cv, sw, jm)[0], ref[k])
            row.append('%8.3f ms %s' % (t, 'same' if same else 'DIFF'))
        tf = timed(lambda: eng.near_flag(pts, cv, 0.1), 3)
        print('%-6s %-5s | %s | %.3f ms' % (cell, rmax, ' | '.join('%-18s' % r for r in row), tf))
