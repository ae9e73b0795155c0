"""Pass-pruning study of the field kernel's split-precision scheme, emulated on the CPU in float64:
every layer is a product of fp16 hi/lo operands with 3 passes (hi*hi + a_hi*w_lo + a_lo*w_hi), 2 passes (drop a_lo*w_hi = fp16
activations, or drop a_hi*w_lo = fp16 weights) or 1 pass (plain fp16). Which layer groups can run with fewer passes inside the
stated tolerances (occupancy 1e-4 abs on an O(1) field; rgb is an 8-bit colour, 1/255 = 3.9e-3)?   python tests/diag_pass_pruning.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from avatarcap_b200 import packer, synth  # noqa: E402
from oracle import field_oracle as fo  # noqa: E402


def split(x):
    hi = x.astype(np.float16).astype(np.float64)
    lo = (x - hi).astype(np.float16).astype(np.float64)
    return hi, lo


def lin(L, x, mode):
    wh, wl = split(L.W.astype(np.float32)); xh, xl = split(x.astype(np.float32))
    acc = wh @ xh
    if mode in ('3', '2w'):
        acc = acc + wl @ xh                  # weight residual
    if mode in ('3', '2a'):
        acc = acc + wh @ xl                  # activation residual
    return (acc * L.scale[:, None].astype(np.float64) + L.bias[:, None].astype(np.float64)).astype(np.float32)


def softplus(v):
    return np.where(v > 20, v, np.log1p(np.exp(np.minimum(v, 20)))).astype(np.float32)


def run(layers, pts, fmap, center, modes):
    """modes: dict group -> '3' | '2a' (keep a_lo*w_hi) | '2w' (keep a_hi*w_lo) | '1'; groups: warp, shared, geo, clr"""
    p = torch.from_numpy(pts)
    pc = p - torch.from_numpy(center)[None]
    feat = fo.bilinear_border(torch.from_numpy(fmap), pc[:, 0], -pc[:, 1]).numpy().astype(np.float32)
    h0 = np.concatenate([pts.T, feat], 0).astype(np.float32)
    h = h0
    for i in range(7):
        inp = np.concatenate([h0, h], 0) if i == 4 else h
        h = softplus(lin(layers[i], inp, modes['warp']))
    off = lin(layers[7], h, '3').T                                  # the 256->3 head runs in fp32 on the CUDA cores
    q = (pts + off).astype(np.float32)
    e = fo.embed(torch.from_numpy(q), 10).numpy().T.astype(np.float32)
    x = e
    for i, L in enumerate(layers[8:15]):
        inp = np.concatenate([x, e], 0) if i == 4 else x
        x = lin(L, inp, modes['shared'])
        if i < 6:
            x = np.maximum(x, 0)
    g = lin(layers[15], x, modes['geo']); g = np.where(g > 0, g, g * np.float32(0.02))
    geo = lin(layers[16], g, '3')
    c = np.maximum(lin(layers[17], x, modes['clr']), 0)
    c = np.maximum(lin(layers[18], c, modes['clr']), 0)
    rgb = 1.0 / (1.0 + np.exp(-lin(layers[19], c, '3').astype(np.float64)))
    return off, geo[0], np.maximum(geo[1], 0), rgb.T


def main():
    sd = synth.avatar_state_dict()
    layers = packer.avatar_layers(sd)
    body = synth.SynthBody(); fr = synth.make_frame(body, None)
    fmap = synth.feature_map(64, 256, 256, synth.SEED + 4)
    rs = np.random.RandomState(3)
    b0, b1 = fr['cano_bounds']
    pts = (rs.uniform(0, 1, (20000, 3)) * (b1 - b0) + b0).astype(np.float32)
    ref = fo.occupancy_query(sd, pts, fmap, fr['cano_smpl_center'], dtype=torch.float64, with_texture=True)
    occ_ref = ref['cano_pts_ov'][:, 0]; rgb_ref = ref['rgb']; al_ref = ref['alpha'][:, 0]
    print('field range: occ %.3g .. %.3g, alpha max %.3g; 20000 points; errors are max-abs vs the f64 oracle' % (occ_ref.min(), occ_ref.max(), al_ref.max()))
    print('%-52s %10s %10s %10s %10s  MMA work' % ('passes (warp / shared / geo / clr)', 'offsets', 'occ', 'alpha/max', 'rgb'))
    flops = {'warp': 856576, 'shared': 850944, 'geo': 65536, 'clr': 196608}
    n = {'3': 3, '2a': 2, '2w': 2, '1': 1}
    for modes in ({'warp': '3', 'shared': '3', 'geo': '3', 'clr': '3'},
                  {'warp': '3', 'shared': '3', 'geo': '3', 'clr': '2a'},
                  {'warp': '3', 'shared': '3', 'geo': '3', 'clr': '2w'},
                  {'warp': '3', 'shared': '3', 'geo': '3', 'clr': '1'},
                  {'warp': '3', 'shared': '3', 'geo': '2a', 'clr': '1'},
                  {'warp': '3', 'shared': '2a', 'geo': '2a', 'clr': '2a'},
                  {'warp': '3', 'shared': '2w', 'geo': '2w', 'clr': '2w'},
                  {'warp': '3', 'shared': '1', 'geo': '1', 'clr': '1'},
                  {'warp': '2a', 'shared': '3', 'geo': '3', 'clr': '3'},
                  {'warp': '2w', 'shared': '3', 'geo': '3', 'clr': '3'},
                  {'warp': '1', 'shared': '3', 'geo': '3', 'clr': '3'}):
        off, occ, al, rgb = run(layers, pts, fmap, fr['cano_smpl_center'], modes)
        work = sum(flops[k] * n[modes[k]] for k in flops) / (3.0 * sum(flops.values()))
        print('%-52s %10.2e %10.2e %10.2e %10.2e  %5.1f %%' % (
            ' / '.join('%s' % modes[k] for k in ('warp', 'shared', 'geo', 'clr')), np.abs(off - ref['nonrigid_offset']).max(), np.abs(occ - occ_ref).max(),
            np.abs(al - al_ref).max() / max(1.0, al_ref.max()), np.abs(rgb - rgb_ref).max(), 100 * work))


if __name__ == '__main__':
    main()
