"""Workload for the ncu capture of the HBM-bound stages around the field kernel (one masked 256^3 frame, like bench.py's `frame`):
validity flag -> field -> scatter -> marching cubes -> LBS skinning -> avatar normal maps -> canonicalize_normal_map.

    ncu --set full --clock-control none -k regex:'mc_|raster_|skin_|canonicalize|scatter_fill|flag_count|near_flag|lbs_weights' \
        -o gpurun_out/r1_aux python profiles/capture_aux.py
    python profiles/ncu_aux_summary.py gpurun_out/r1_aux.ncu-rep profiles/r1_ncu_aux_kernels.md
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avatarcap_b200 import pipeline, render, synth  # noqa: E402
from avatarcap_b200.engine import Engine  # noqa: E402

eng = Engine(); dev = eng.device
body = synth.SynthBody(); fr = synth.make_frame(body)
res = (256, 256, 256)
eng.load_avatar(synth.avatar_state_dict()); eng.set_pose_feature_map(synth.feature_map(64, 256, 256, synth.SEED + 4))
center = fr['cano_smpl_center']
cv = torch.from_numpy(fr['cano_smpl_v']).to(dev); sw = torch.from_numpy(fr['smpl_skinning_weights']).to(dev); jm = torch.from_numpy(fr['cano2live_jnt_mats']).to(dev)
pts = eng.make_grid(fr['cano_bounds'], res)
flag = pipeline.valid_points_flag(eng, pts, cv)
fill = torch.from_numpy(2.0 * synth.body_inside(pts[~flag].cpu().numpy(), synth.cano_pose()).astype(np.float32) - 1.0).to(dev)
vpts = pts[flag].contiguous()
o = eng.eval_occupancy(vpts, center, want_offsets=True)
vol = eng.scatter_fill(flag, o['occ'], fill).reshape(res)
v, f, n = eng.extract_mesh(vol, fr['cano_bounds'], 0.0)
lv, ln = eng.skin_mesh(v, n, cv, sw, jm)
lbs = eng.lbs_weights(v, cv, sw); live_v, vmats = eng.skin_points(v, lbs, jm, return_pt_mats=True)
front, back = render.render_cano_mesh_device(eng, v, n, f, center, 512)
lc = 0.5 * (live_v.max(0)[0] + live_v.min(0)[0]).cpu().numpy()
w2c = np.identity(4, np.float32); w2c[:3, :3] = np.diag([1., -1., -1.]).astype(np.float32); w2c[:3, 3] = -(w2c[:3, :3] @ lc) + np.float32([0, 0, 2.6])
nmap = torch.zeros((512, 512, 3), device=dev); nmap[..., 2] = -1.0
render.canonicalize_normal_map_device(eng, v, live_v, f, nmap, vmats, w2c, 550., 550., 256., 256., center, 512)
torch.cuda.synchronize()
print('frame: %d valid points, %d vertices, %d faces, normal-map coverage %.1f%%' % (vpts.shape[0], v.shape[0], f.shape[0], 100 * float((front.norm(dim=-1) > 0).float().mean())))
