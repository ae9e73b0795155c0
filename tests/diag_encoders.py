"""GPU diagnostic (not a pytest): encoder time and error against the reference goldens for the cuDNN switches.
Usage: python tests/diag_encoders.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import load_golden  # noqa: E402
from avatarcap_b200 import encoders, synth  # noqa: E402


def ms(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    g = load_golden('encoder_golden.npz')
    x = torch.from_numpy(synth.smpl_pos_map()).cuda(); y = torch.from_numpy(synth.normal_maps()).cuda()
    usd = synth.unet_state_dict(); hsd = synth.hgfilter_state_dict()
    for det, tf32, graph, cl, bm in ((True, False, True, True, False), (False, False, True, True, False), (False, False, True, False, False),
                                     (False, False, True, True, True), (False, False, True, False, True), (True, False, True, False, True),
                                     (False, True, True, True, False), (False, True, True, False, True)):
        pe = encoders.PoseFeatureEncoder(usd, device='cuda', use_graph=graph, allow_tf32=tf32, deterministic=det, channels_last=cl, benchmark=bm)
        ie = encoders.ImageFeatureEncoder(hsd, device='cuda', use_graph=graph, allow_tf32=tf32, deterministic=det, channels_last=cl, benchmark=bm)
        po = pe(x); io = ie(y)
        eu = float(np.abs(po[0].reshape(64, -1)[:, torch.as_tensor(g['pose_idx']).cuda()].cpu().numpy() - g['pose_feat']).max())
        eh = float(np.abs(io[0].reshape(32, -1)[:, torch.as_tensor(g['img_idx']).cuda()].cpu().numpy() - g['img_feat']).max())
        print('det=%-5s tf32=%-5s graph=%-5s channels_last=%-5s benchmark=%-5s | unet %.3f ms err %.2e | hgfilter %.3f ms err %.2e' % (
            det, tf32, graph, cl, bm, ms(lambda: pe(x)), eu, ms(lambda: ie(y)), eh))
        del pe, ie

if __name__ == '__main__':
    main()
