"""A/B of the marching-cubes face pass (AVC_MC_FACES) and the KNN-4 block size (AVC_KNN_BLOCK) in ONE process (the knobs are read per
call) on the bench's masked-frame volume and on a smooth body surface.   python tests/diag_mesh_knobs.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avatarcap_b200 import pipeline, synth  # noqa: E402
from avatarcap_b200.engine import Engine  # noqa: E402

eng = Engine(); dev = eng.device
body = synth.SynthBody(); fr = synth.make_frame(body)
res = (256, 256, 256)
eng.load_avatar(synth.avatar_state_dict()); eng.set_pose_feature_map(synth.feature_map(64, 256, 256, synth.SEED + 4))
cv = torch.from_numpy(fr['cano_smpl_v']).to(dev); sw = torch.from_numpy(fr['smpl_skinning_weights']).to(dev); jm = torch.from_numpy(fr['cano2live_jnt_mats']).to(dev)
pts = eng.make_grid(fr['cano_bounds'], res)
flag = pipeline.valid_points_flag(eng, pts, cv)
fill = torch.from_numpy(2.0 * synth.body_inside(pts[~flag].cpu().numpy(), synth.cano_pose()).astype(np.float32) - 1.0).to(dev)
o = eng.eval_occupancy(pts[flag].contiguous(), fr['cano_smpl_center'])
vols = {'masked_noise': eng.scatter_fill(flag, o['occ'], fill).reshape(res),
        'dense_noise': eng.eval_occupancy(pts, fr['cano_smpl_center'])['occ'].reshape(res),
        'smooth_body': torch.from_numpy(synth.body_sdf(pts.cpu().numpy(), synth.cano_pose()).reshape(res)).to(dev)}


def t(fn, reps=7):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return round(float(np.median(ts)), 4)


out = {}
for name, vol in vols.items():
    r = {}
    ref = None
    for mode in ('triangle',):
        os.environ['AVC_MC_FACES'] = mode.split('+')[0]
        if mode.endswith('scalar'):
            os.environ['AVC_MC_SCALAR'] = '1'
        else:
            os.environ.pop('AVC_MC_SCALAR', None)
        r['mesh_ms_' + mode] = t(lambda: eng.extract_mesh(vol, fr['cano_bounds'], 0.0))
        m = eng.extract_mesh(vol, fr['cano_bounds'], 0.0)
        if ref is None:
            ref = [x.clone() for x in m]
        else:
            r['identical'] = all(torch.equal(a, b) for a, b in zip(ref, m))
    ref = ref or [x.clone() for x in m]
    os.environ.pop('AVC_MC_SCALAR', None); os.environ.pop('AVC_MC_FACES', None)
    v, f, n = ref
    r['verts'] = int(v.shape[0])
    for kb in ('256',):
        os.environ['AVC_KNN_BLOCK'] = kb
        r['skin_ms_b' + kb] = t(lambda: eng.skin_mesh(v, n, cv, sw, jm))
    os.environ.pop('AVC_KNN_BLOCK')
    for rm in ('2', '3', '4', '5', '6', '8'):
        os.environ['AVC_KNN_RMAX'] = rm
        r['skin_ms_rmax' + rm] = t(lambda: eng.skin_mesh(v, n, cv, sw, jm), 3)
    os.environ.pop('AVC_KNN_RMAX')
    out[name] = r
print(json.dumps(out))
