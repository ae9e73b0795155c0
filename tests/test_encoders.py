"""Per-frame encoders against golden outputs of the REFERENCE modules
(tests/golden/gen_encoder_golden.py: UnetNoCond7DS / HGFilter imported from the reference, CPU f32)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import load_golden  # noqa: E402
from avatarcap_b200 import encoders, synth  # noqa: E402


def _sampled(t, idx):
    c = t.shape[1]
    return t[0].reshape(c, -1)[:, torch.as_tensor(idx, device=t.device)].float().cpu().numpy()


def _report(name, got, ref):
    err = float(np.abs(got - ref).max()); rng = float(ref.max() - ref.min())
    print('%s: max|d| = %.3e (range %.3g, relative %.2e)' % (name, err, rng, err / rng))
    return err


def test_state_dict_keys_cover_reference_names():
    """The unused upconv4 keys exist (checkpoint compatibility) and both dicts carry the reference's module paths."""
    u = synth.unet_state_dict(); h = synth.hgfilter_state_dict()
    assert 'upconv4.up.weight' in u and u['upconv4.up.weight'].shape == (512, 128, 4, 4)
    assert u['conv1.conv.weight'].shape == (32, 6, 4, 4) and u['upconvC7.up.1.bias'].shape == (64,)
    assert 'conv2.bn.running_var' in u and 'conv1.bn.running_var' not in u and 'conv7.bn.running_var' not in u
    assert h['conv1.weight'].shape == (64, 6, 7, 7) and h['l0.weight'].shape == (32, 256, 1, 1)
    assert 'm0.b2_plus_1.conv1.weight' in h and 'conv2.downsample.2.weight' in h and 'conv3.downsample.2.weight' not in h


def test_pose_encoder_cpu_vs_reference_golden():
    g = load_golden('encoder_golden.npz')
    enc = encoders.PoseFeatureEncoder(synth.unet_state_dict(), device='cpu', use_graph=False)
    out = enc(synth.smpl_pos_map())
    assert tuple(out.shape) == (1, 64, 256, 256) and out.is_contiguous(memory_format=torch.channels_last)
    assert _report('unet cpu', _sampled(out, g['pose_idx']), g['pose_feat']) < 2e-5          # BN folded, f32
    with pytest.raises(ValueError):
        enc(np.zeros((1, 6, 100, 100), np.float32))


def test_image_encoder_cpu_vs_reference_golden():
    g = load_golden('encoder_golden.npz')
    enc = encoders.ImageFeatureEncoder(synth.hgfilter_state_dict(), device='cpu', use_graph=False)
    out = enc(synth.normal_maps())
    assert tuple(out.shape) == (1, 32, 256, 256)
    assert _report('hgfilter cpu', _sampled(out, g['img_idx']), g['img_feat']) < 1e-4


def test_prefix_lookup():
    sd = {'warping_field.unet.' + k: v for k, v in synth.unet_state_dict().items()}
    enc = encoders.PoseFeatureEncoder(sd, prefix='warping_field.unet.', device='cpu', use_graph=False)
    assert enc.c7[0].shape == (64, 64, 3, 3)


@pytest.mark.gpu
def test_encoders_gpu_graph_vs_golden_and_hwc_handoff():
    from avatarcap_b200.engine import Engine
    g = load_golden('encoder_golden.npz')
    pe = encoders.PoseFeatureEncoder(synth.unet_state_dict(), device='cuda', deterministic=True)
    ie = encoders.ImageFeatureEncoder(synth.hgfilter_state_dict(), device='cuda')
    x = torch.from_numpy(synth.smpl_pos_map()).cuda(); y = torch.from_numpy(synth.normal_maps()).cuda()
    po = pe(x); assert _report('unet cuda graph', _sampled(po, g['pose_idx']), g['pose_feat']) < 1e-4
    io = ie(y); assert _report('hgfilter cuda graph', _sampled(io, g['img_idx']), g['img_feat']) < 5e-4
    # replay with another input, then the first again: the captured graph must follow its static input
    first = po.clone()
    other = pe(x * 0.5 + 0.1); assert float((other - first).abs().max()) > 1e-3
    again = pe(x); print('replay difference', float((again - first).abs().max())); assert torch.equal(again, first)       # deterministic cuDNN algorithms
    # the channels_last output goes to the library without a transpose and must evaluate identically to the (C,H,W) route
    eng = Engine()
    eng.load_avatar(synth.avatar_state_dict())
    frame = synth.make_frame(synth.SynthBody(), None)
    pts = eng.make_grid(frame['cano_bounds'], (48, 48, 48))
    eng.set_pose_feature_map(first)                                            # channels_last -> avc_set_feature_map_hwc
    a = eng.eval_occupancy(pts, frame['cano_smpl_center'])
    eng.set_pose_feature_map(first.contiguous())                               # (C,H,W) -> transpose kernel
    b = eng.eval_occupancy(pts, frame['cano_smpl_center'])
    assert torch.equal(a['occ'], b['occ']) and torch.equal(a['off'], b['off'])
    eng.close()


def test_hgfilter_program_interpreted_on_the_cpu_vs_golden():
    """The op program + packed weights that encoders.build_hgfilter_program hands to the library, executed by a torch-CPU interpreter
    (tests/encoder_program_interp.py), must reproduce the reference's own HGFilter: pins the program structure, the (C_out, taps, C_in_pad)
    fp16 hi/lo weight packing with its scale exponents, and every parameter offset without a GPU."""
    import encoder_program_interp as interp
    g = load_golden('encoder_golden.npz')
    prog, wbytes, params = encoders.build_hgfilter_program(synth.hgfilter_state_dict())
    nb, npl, nops = int(prog[1]), int(prog[2]), int(prog[3])
    assert len(prog) == 16 + nb + 2 * npl + 16 * nops and (prog[4], prog[5], prog[6]) == (6, 512, 512) and tuple(prog[8:11]) == (32, 256, 256)
    ops = prog[16 + nb + 2 * npl:].reshape(nops, 16)
    assert (ops[:, 0] == interp.OP_CONV).sum() == 55 and (ops[:, 0] == interp.OP_STEM).sum() == 1         # 15 ConvBlocks x 3 + 3 downsample + 2 heads... = 55
    conv = ops[ops[:, 0] == interp.OP_CONV]
    assert (conv[:, 2] % 128 == 0).all() and set(conv[:, 8].tolist()) == {1, 9} and (conv[:, 6] % 64 == 0).all()     # aligned planes, taps, padded C_in
    out = interp.run_program(prog, wbytes, params, synth.normal_maps()[0])           # (256, 256, 32)
    got = out.reshape(-1, 32)[g['img_idx']].T                                        # (32, n_samples) like _sampled
    err = _report('hgfilter program on the CPU interpreter', got, g['img_feat'])
    assert err < 5e-6


def test_unet_tail_program_interpreted_on_the_cpu_vs_golden():
    """upconvC5 / C6 / C7 of the UNet as a library program (encoders.build_unet_tail_program: bilinear x2 + tcgen05 3x3 convolutions with
    the eval BatchNorm folded, skips copied into their channel slices), executed by the CPU interpreter on the functional head's tensors,
    against the reference's own UnetNoCond7DS output."""
    import torch.nn.functional as F
    import encoder_program_interp as interp
    g = load_golden('encoder_golden.npz')
    sd = synth.unet_state_dict()
    prog, wbytes, params, counts = encoders.build_unet_tail_program(sd)
    assert counts == (32 * 32 * 384, 64 * 64 * 64, 128 * 128 * 32) and tuple(prog[8:11]) == (64, 256, 256)
    pe = encoders.PoseFeatureEncoder(sd, device='cpu', use_graph=False)
    x = torch.from_numpy(synth.smpl_pos_map())
    with torch.no_grad():                                   # the head exactly as PoseFeatureEncoderTC._forward_head runs it
        a = []
        h = F.conv2d(x.contiguous(memory_format=pe.mf), pe.down[0][0], None, stride=2, padding=1)
        for w, b in pe.down[1:]:
            h = F.leaky_relu(h, 0.2); a.append(h); h = F.conv2d(h, w, b, stride=2, padding=1)
        for (w, b), skip in zip((pe.up[0], pe.up[1], pe.up[2], pe.up[2]), (a[5], a[4], a[3], a[2])):
            h = torch.cat([F.conv_transpose2d(F.relu(h), w, b, stride=2, padding=1), skip], 1)
        flat = torch.cat([t[0].permute(1, 2, 0).reshape(-1) for t in (h, a[1], a[0])]).numpy()
    out = interp.run_program(prog, wbytes, params, flat)      # (256, 256, 64)
    err = _report('unet tail program on the CPU interpreter', out.reshape(-1, 64)[g['pose_idx']].T, g['pose_feat'])
    assert err < 5e-6


def test_unet_whole_program_interpreted_on_the_cpu_vs_golden():
    """The whole UNet as one library program (encoders.build_unet_program): the 4x4 stride-2 and transposed convolutions as the gather-GEMM
    of conv4_gemm_kernel -- the interpreter transcribes its index formulas, it does not call F.conv2d for them -- with every skip written
    into the channel slice of its concatenated buffer, then the tcgen05 tail; against the reference's own UnetNoCond7DS output."""
    import encoder_program_interp as interp
    g = load_golden('encoder_golden.npz')
    prog, wbytes, params = encoders.build_unet_program(synth.unet_state_dict())
    assert tuple(prog[4:7]) == (6, 256, 256) and tuple(prog[8:11]) == (64, 256, 256)
    assert sum(1 for o in prog[16 + prog[1] + 2 * prog[2]:].reshape(-1, 16) if o[0] == encoders.OP_CONV4) == 11
    out = interp.run_program(prog, wbytes, params, synth.smpl_pos_map()[0])
    assert _report('unet program on the CPU interpreter', out.reshape(-1, 64)[g['pose_idx']].T, g['pose_feat']) < 5e-6


def test_unet_program_on_a_non_square_input():
    """The same program builder at 128 x 256 (conv7 ends on a 1 x 2 map, the first transposed convolution starts from it): interpreter vs the
    functional restatement, which test_encoders pins against the reference at 256 x 256."""
    import encoder_program_interp as interp
    sd = synth.unet_state_dict()
    x = np.random.RandomState(3).randn(1, 6, 128, 256).astype(np.float32)
    prog, wbytes, params = encoders.build_unet_program(sd, in_hw=(128, 256))
    out = interp.run_program(prog, wbytes, params, x[0])
    ref = encoders.PoseFeatureEncoder(sd, device='cpu', use_graph=False)(torch.from_numpy(x))[0].permute(1, 2, 0).numpy()
    assert out.shape == ref.shape == (128, 256, 64)
    assert _report('unet program at 128 x 256', out, ref) < 2e-6 * float(np.abs(ref).max())          # f32 rounding, relative to the +-10 range


def test_conv4_weight_packing_matches_torch():
    """pack_conv4_weight + the interpreter's gather-GEMM == F.conv2d / F.conv_transpose2d (4x4, stride 2, padding 1) on odd shapes: non-square
    maps, a 1-pixel-high input of the transposed convolution, channel slices of wider buffers, ReLU in / LeakyReLU out."""
    import torch.nn.functional as F
    import encoder_program_interp as interp
    rs = np.random.RandomState(5)
    for tr, hin, win, ci, co in ((False, 6, 10, 8, 12), (True, 1, 3, 4, 8), (True, 5, 2, 12, 4), (False, 2, 2, 20, 4)):
        pr = encoders._Program()
        w = rs.randn(*((ci, co, 4, 4) if tr else (co, ci, 4, 4))).astype(np.float32); b = rs.randn(co).astype(np.float32)
        ho, wo = (2 * hin, 2 * win) if tr else (hin // 2, win // 2)
        src = pr.buf(hin * win * (ci + 8)); dst = pr.buf(ho * wo * (co + 4)); pr.plane(1, 64)
        x = rs.randn(hin, win, ci + 8).astype(np.float32)
        pr.op(encoders.OP_INPUT, src, 0, x.size)
        pr.op(encoders.OP_CONV4, src, dst, hin, win, ci, ci + 8, 4, co, co + 4, 4, pr.param(encoders.pack_conv4_weight(w, tr)), pr.param(b), (1 if tr else 0) | 2 | 4)
        prog, wb, params = pr.pack((x.size, 1, 1), dst, (ho, wo, co + 4))
        got = interp.run_program(prog, wb, params, x.reshape(-1, 1, 1))[:, :, 4:]
        xin = torch.relu(torch.from_numpy(x[:, :, 4:4 + ci])).permute(2, 0, 1)[None]
        ref = (F.conv_transpose2d if tr else F.conv2d)(xin, torch.from_numpy(w), torch.from_numpy(b), stride=2, padding=1)
        ref = F.leaky_relu(ref, 0.2)[0].permute(1, 2, 0).numpy()
        assert np.abs(got - ref).max() < 1e-4 * max(1.0, np.abs(ref).max()), (tr, hin, win, ci, co)


@pytest.mark.gpu
@pytest.mark.parametrize('head', ['library', 'cudnn'])
def test_unet_tensor_core_tail_vs_golden(head):
    """PoseFeatureEncoderTC -- head='library': the whole UNet on kernels of this library (split-K fp32 gather-GEMMs for the 4x4 stride-2 /
    transposed convolutions, tcgen05 for the 3x3 stages); head='cudnn': cuDNN head + library tail -- against the reference's
    UnetNoCond7DS (golden) and the all-cuDNN restatement; deterministic; hands the (H,W,C) map to the field kernel unchanged."""
    from avatarcap_b200.engine import Engine
    g = load_golden('encoder_golden.npz')
    eng = Engine()
    if not eng.has_tensor_core_path:
        pytest.skip('needs sm_100')
    x = torch.from_numpy(synth.smpl_pos_map()).cuda()
    tc = encoders.PoseFeatureEncoderTC(synth.unet_state_dict(), engine=eng, deterministic=True, head=head)
    out = tc(x).clone()
    assert tuple(out.shape) == (1, 64, 256, 256) and out.is_contiguous(memory_format=torch.channels_last)
    assert _report('unet tcgen05 tail', _sampled(out, g['pose_idx']), g['pose_feat']) < 1e-5
    ref = encoders.PoseFeatureEncoder(synth.unet_state_dict(), device='cuda', use_graph=False, deterministic=True)(x)
    print('vs the cuDNN f32 restatement: max-abs %.3g' % float((out - ref).abs().max()))
    assert float((out - ref).abs().max()) < 1e-5
    other = tc(x * 0.5 + 0.1).clone(); assert float((other - out).abs().max()) > 1e-3
    again = tc(x); print('replay difference', float((again - out).abs().max())); assert torch.equal(again, out)
    eager = encoders.PoseFeatureEncoderTC(synth.unet_state_dict(), engine=eng, use_graph=False, deterministic=True, head=head)
    assert torch.equal(eager(x), out); eager.close()                    # graph replay == eager launches
    eng.load_avatar(synth.avatar_state_dict())
    frame = synth.make_frame(synth.SynthBody(), None)
    pts = eng.make_grid(frame['cano_bounds'], (32, 32, 32))
    eng.set_pose_feature_map(out)
    a = eng.eval_occupancy(pts, frame['cano_smpl_center'])
    eng.set_pose_feature_map(out.contiguous())
    b = eng.eval_occupancy(pts, frame['cano_smpl_center'])
    assert torch.equal(a['occ'], b['occ'])
    tc.close(); eng.close()


@pytest.mark.gpu
def test_hgfilter_tensor_core_vs_golden():
    """HGFilter on kernels of the library (csrc/conv_tc.cu: tcgen05 implicit-GEMM convolutions fed by TMA tensor loads, fp16 hi/lo split
    operands): within 1e-5 of the reference's own HGFilter (golden made on the CPU by tests/golden/gen_encoder_golden.py), as close as
    the cuDNN f32 restatement, deterministic, graph == eager, and the (H,W,C) result feeds the recon decoder unchanged."""
    from avatarcap_b200.engine import Engine
    g = load_golden('encoder_golden.npz')
    eng = Engine()
    if not eng.has_tensor_core_path:
        pytest.skip('needs sm_100')
    y = torch.from_numpy(synth.normal_maps()).cuda()
    tc = encoders.ImageFeatureEncoderTC(synth.hgfilter_state_dict(), engine=eng, use_graph=True)
    out = tc(y).clone()
    assert tuple(out.shape) == (1, 32, 256, 256) and out.is_contiguous(memory_format=torch.channels_last)
    err = _report('hgfilter tcgen05', _sampled(out, g['img_idx']), g['img_feat'])
    assert err < 1e-5                                                     # measured 7.4e-6 (cuDNN f32: 3.8e-6); feature range +-1.8, 55 convolutions deep
    ref = encoders.ImageFeatureEncoder(synth.hgfilter_state_dict(), device='cuda', use_graph=False, benchmark=False)(y)
    print('vs the cuDNN f32 restatement: max-abs %.3g' % float((out - ref).abs().max()))
    assert float((out - ref).abs().max()) < 1.5e-5
    again = tc(y * 0.5 + 0.1).clone(); assert float((again - out).abs().max()) > 1e-4
    assert torch.equal(tc(y), out)                                            # replays are bit-reproducible (fixed-order GroupNorm sums)
    eager = encoders.ImageFeatureEncoderTC(synth.hgfilter_state_dict(), engine=eng, use_graph=False)
    assert torch.equal(eager(y), out)
    # hand-off to the per-point decoder
    eng.load_recon(synth.recon_state_dict())
    frame = synth.make_frame(synth.SynthBody(), None)
    pts = eng.make_grid(frame['cano_bounds'], (32, 32, 32))
    eng.set_image_feature_map(out)
    a = eng.eval_recon(pts, frame['cano_smpl_center'])
    eng.set_image_feature_map(out.contiguous())
    assert torch.equal(a, eng.eval_recon(pts, frame['cano_smpl_center']))
    # what the encoder's error does to the quantity the tolerance is stated on: the reconstruction occupancy (north_star: 1e-4)
    eng.set_image_feature_map(ref)
    d_occ = float((a - eng.eval_recon(pts, frame['cano_smpl_center'])).abs().max())
    print('recon occupancy, tcgen05 features vs cuDNN f32 features: max-abs %.3g' % d_occ)
    assert d_occ < 2e-5
    tc.close(); eager.close(); eng.close()
