"""Stage-by-stage comparison of the tensor-core HGFilter (encoders.ImageFeatureEncoderTC, csrc/conv_tc.cu) with the functional
cuDNN restatement (encoders.ImageFeatureEncoder, f32) and the reference golden, plus timings. Run on a B200:  python tests/diag_encoder_tc.py"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from avatarcap_b200 import encoders, synth  # noqa: E402
from avatarcap_b200.engine import Engine  # noqa: E402
from helpers import load_golden  # noqa: E402


def main():
    torch.cuda.set_device(0)
    eng = Engine()
    sd = synth.hgfilter_state_dict()
    y = torch.from_numpy(synth.normal_maps()).cuda()
    ref = encoders.ImageFeatureEncoder(sd, device='cuda', use_graph=False, benchmark=False)
    tc = encoders.ImageFeatureEncoderTC(sd, engine=eng, use_graph=False)
    out = tc(y).clone(); torch.cuda.synchronize()
    # reference intermediates (NCHW f32, cuDNN without TF32)
    with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=False, allow_tf32=False):
        p = ref.p
        stem = F.conv2d(y, p['conv1.weight'], p['conv1.bias'], stride=2, padding=3)
        r = {'stem': stem}
        x = F.relu(ref._gn(stem, 'bn1')); r['bn1'] = x
        x = ref._block(x, 'conv2'); r['conv2'] = x
        x = ref._block(x, 'conv3'); r['conv3'] = x
        x = ref._block(x, 'conv4'); r['conv4'] = x
        x = ref._hourglass(4, x); r['hourglass'] = x
        x = ref._block(x, 'top_m_0'); r['top_m_0'] = x
        ll = F.conv2d(x, p['conv_last0.weight'], p['conv_last0.bias']); r['conv_last0'] = ll
        ll = F.relu(ref._gn(ll, 'bn_end0'))
        final = F.conv2d(ll, p['l0.weight'], p['l0.bias'])
    for name in ('stem', 'bn1', 'conv2', 'conv3', 'conv4', 'hourglass', 'top_m_0', 'conv_last0'):
        a = tc.intermediate(name).permute(2, 0, 1)[None]
        b = r[name]
        d = (a - b).abs()
        print('%-11s max-abs diff %.3e  (ref range %.3g .. %.3g, mean |ref| %.3g)  worst at %s' % (
            name, float(d.max()), float(b.min()), float(b.max()), float(b.abs().mean()), np.unravel_index(int(d.argmax()), d.shape)))
    d = (out - final).abs()
    print('%-11s max-abs diff %.3e  (ref range %.3g .. %.3g)' % ('output', float(d.max()), float(final.min()), float(final.max())))
    g = load_golden('encoder_golden.npz')
    samp = out[0].reshape(32, -1)[:, torch.from_numpy(g['img_idx']).cuda()].cpu().numpy()
    print('vs the reference golden (CPU PyTorch): max-abs %.3e' % float(np.abs(samp - g['img_feat']).max()))
    # timings
    def timed(fn, reps=10):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        return float(np.median(ts))
    tcg = encoders.ImageFeatureEncoderTC(sd, engine=eng, use_graph=True)
    refg = encoders.ImageFeatureEncoder(sd, device='cuda', use_graph=True)
    print('HGFilter: tcgen05 eager %.3f ms, tcgen05 graph %.3f ms, cuDNN f32 graph %.3f ms' % (timed(lambda: tc(y)), timed(lambda: tcg(y)), timed(lambda: refg(y))))
    o2 = tcg(y).clone(); o3 = tcg(y).clone()
    print('graph == eager: %s, replay deterministic: %s' % (bool(torch.equal(o2, out)), bool(torch.equal(o2, o3))))
    # UNet: the whole network as a library program (head='library') and cuDNN head + library tail (head='cudnn')
    xs = torch.from_numpy(synth.smpl_pos_map()).cuda()
    usd = synth.unet_state_dict()
    uref = encoders.PoseFeatureEncoder(usd, device='cuda', use_graph=True)
    ur = uref(xs).clone()
    for head in ('library', 'cudnn'):
        utc = encoders.PoseFeatureEncoderTC(usd, engine=eng, head=head)
        uo = utc(xs).clone()
        samp = uo[0].reshape(64, -1)[:, torch.from_numpy(g['pose_idx']).cuda()].cpu().numpy()
        print('UNet head=%s: vs cuDNN f32 max-abs %.3e, vs the reference golden %.3e (range %.3g .. %.3g); %.3f ms (all-cuDNN f32 graph %.3f ms)' % (
            head, float((uo - ur).abs().max()), float(np.abs(samp - g['pose_feat']).max()), float(ur.min()), float(ur.max()),
            timed(lambda: utc(xs)), timed(lambda: uref(xs))))
        utc.close()


if __name__ == '__main__':
    main()
