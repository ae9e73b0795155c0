"""One launch of the field kernel on the bench workload through a chosen entry point, for `ncu --set full` captures:
    ncu --set full --clock-control none --import-source on -k regex:field_tc2 -s 1 -c 1 -o gpurun_out/prof python tests/diag_ncu_field.py grid|list"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from avatarcap_b200 import synth  # noqa: E402
from avatarcap_b200.engine import Engine  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else 'grid'
eng = Engine()
fr = synth.make_frame(synth.SynthBody(), None)
eng.load_avatar(synth.avatar_state_dict()); eng.set_pose_feature_map(synth.feature_map(64, 256, 256, synth.SEED + 4))
res = (256, 256, 256)
pts = eng.make_grid(fr['cano_bounds'], res) if mode == 'list' else None
for _ in range(2):                                  # launch 0 = warm-up, launch 1 = the captured one (-s 1 -c 1)
    if mode == 'list':
        eng.eval_occupancy(pts, fr['cano_smpl_center'], want_offsets=True, want_texture=True)
    else:
        eng.eval_occupancy_grid(fr['cano_bounds'], res, fr['cano_smpl_center'], want_offsets=True, want_texture=True)
    torch.cuda.synchronize()
print('done', mode)
