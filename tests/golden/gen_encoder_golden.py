"""Golden vectors for the per-frame encoders, produced by the REFERENCE's own modules on CPU (f32):

    network.unets.UnetNoCond7DS(input_nc=6, output_nc=64, nf=32, up_mode='upconv')      arch_avatar.py:95
    network.HGFilters.HGFilter(1, 4, 6, 32, 'group', 'no_down', False)                  arch_recon.py:28

Run in the build container only:   python tests/golden/gen_encoder_golden.py   -> tests/golden/encoder_golden.npz

Weights and inputs are the seeded ones of avatarcap_b200.synth (loaded with strict=True, which also pins the key names and
shapes of synth.unet_state_dict / hgfilter_state_dict to the reference's). Only a seeded subsample of the output pixels is
stored (all channels at 4096 positions) to keep the fixture small; the positions come from encoders.subsample_index.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('AVATARCAP_REFERENCE', '/root/reference')
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from avatarcap_b200 import encoders, synth  # noqa: E402


def tsd(sd):
    return {k: (torch.from_numpy(v) if v.shape else torch.tensor(v)) for k, v in sd.items()}


def main() -> None:
    torch.manual_seed(synth.SEED); torch.set_num_threads(max(1, os.cpu_count() or 1))
    from network.unets import UnetNoCond7DS
    from network.HGFilters import HGFilter
    unet = UnetNoCond7DS(input_nc=6, output_nc=64, nf=32, up_mode='upconv', use_dropout=False)
    unet.load_state_dict(tsd(synth.unet_state_dict()), strict=True); unet.eval()
    hg = HGFilter(1, 4, 6, 32, 'group', 'no_down', False)
    hg.load_state_dict(tsd(synth.hgfilter_state_dict()), strict=True); hg.eval()
    with torch.no_grad():
        pose = unet(torch.from_numpy(synth.smpl_pos_map()))[0].numpy()                       # (64,256,256)
        img = hg(torch.from_numpy(synth.normal_maps()))[0][-1][0].numpy()                    # (32,256,256)
    ip = encoders.subsample_index(64, 256, 256, 4096, synth.SEED + 20)
    ii = encoders.subsample_index(32, 256, 256, 4096, synth.SEED + 21)
    np.savez_compressed(os.path.join(HERE, 'encoder_golden.npz'),
                        pose_idx=ip, pose_feat=pose.reshape(64, -1)[:, ip].astype(np.float32),
                        pose_stats=np.array([pose.mean(), pose.std(), np.abs(pose).max()], np.float64),
                        img_idx=ii, img_feat=img.reshape(32, -1)[:, ii].astype(np.float32),
                        img_stats=np.array([img.mean(), img.std(), np.abs(img).max()], np.float64))
    print('pose feature map', pose.shape, 'std %.4f' % pose.std(), '| image feature map', img.shape, 'std %.4f' % img.std())


if __name__ == '__main__':
    main()
