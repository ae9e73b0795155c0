"""avatarcap_b200.patch: the drop-in re-binding of the reference's call sites, exercised against
stand-in modules with the reference's class / attribute / key names (tests/fake_reference.py)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fake_reference as fr  # noqa: E402
from avatarcap_b200 import synth  # noqa: E402


@pytest.fixture()
def frame():
    return synth.make_frame(synth.SynthBody(), None)


def test_install_rebinds_and_uninstall_restores(frame):
    from avatarcap_b200 import patch
    mods = fr.make_modules(frame)
    aa = mods['network.arch_avatar']; ar = mods['network.arch_recon']
    orig = {(c, n): getattr(c, n) for c, n in ((aa.OccupancyNet, 'query'), (aa.WarpingField, 'query'), (aa.WarpingField, 'precompute_conv'),
                                              (aa.DoubleTNet, 'forward'), (aa.GeoTexAvatar, 'forward'), (ar.ReconNetwork, 'infer'),
                                              (ar.ReconNetwork, 'get_feat_maps'), (mods['utils.smpl_util'].SmplUtil, 'skinning'))}
    orig_rm = mods['utils.recon_util'].recon_mesh
    patch.install(engine=object(), modules=mods)            # the engine is only touched by calls under no_grad
    try:
        assert all(getattr(c, n) is not f for (c, n), f in orig.items())
        assert mods['utils.recon_util'].recon_mesh is not orig_rm
        # with autograd enabled every patched method falls through to the reference's own code (training, main.py:97-116)
        net = aa.GeoTexAvatar(frame).eval()                 # main.py:297
        with torch.enable_grad():
            with pytest.raises(fr.Fallthrough):
                aa.OccupancyNet(net).query({})
            with pytest.raises(fr.Fallthrough):
                net.warping_field.precompute_conv({'smpl_pos_map': torch.zeros(1, 6, 256, 256)})
            with pytest.raises(fr.Fallthrough):
                net(None, None, None, {}, 'cano')
        # host tensors never reach the CUDA encoders
        with torch.no_grad(), pytest.raises(fr.Fallthrough):
            net.warping_field.precompute_conv({'smpl_pos_map': torch.zeros(1, 6, 256, 256)})
        # a TRAIN-mode network under no_grad (finetune_tex's network_init, main.py:186-188, 226-231) evaluates BatchNorm with batch
        # statistics in the reference: the eval-mode-folding replacements must fall through
        net_init = aa.GeoTexAvatar(frame)                   # never .eval()'d, like the reference
        assert net_init.training and patch._mode_dependent(net_init)
        with torch.no_grad():
            with pytest.raises(fr.Fallthrough):
                aa.OccupancyNet(net_init).query({})
            with pytest.raises(fr.Fallthrough):
                net_init(None, None, None, {}, 'cano')
            with pytest.raises(fr.Fallthrough):
                net_init.warping_field.query(None, {})
            with pytest.raises(fr.Fallthrough):
                net_init.cano_template(None)
            # a sub-module that belongs to no packed network never evaluates with somebody else's weights
            with pytest.raises(fr.Fallthrough):
                net.eval().cano_template(None)
        # the stock ReconNetwork (GroupNorm + weight-norm) is mode-free: main.py never calls recon_net.eval() (main.py:300)
        assert not patch._mode_dependent(ar.ReconNetwork())
        # render stage: phong previews stay on the (fake) GL renderer and the GL-renderer call paths fall through
        R = mods['utils.renderer'].Renderer
        gl = R(512, 512, shader_name='phong_geometry', bg_color=(1, 1, 1), window_name='Phong')
        assert isinstance(gl, fr.GLRenderer) and gl.shader_name == 'phong_geometry'
        with pytest.raises(fr.Fallthrough):
            mods['utils.visualize_util'].render_cano_mesh(gl, None, None, None, np.zeros(3))
        with pytest.raises(fr.Fallthrough):
            mods['normal_fusion.normal_fusion'].canonicalize_normal_map(gl, gl, None)
        # the PLY writer needs no GPU: same bytes as the reference's (golden made by utils/obj_io.py itself)
        import tempfile
        from helpers import load_golden
        g = load_golden('ply_golden.npz')
        with tempfile.TemporaryDirectory() as td:
            mods['utils.obj_io'].save_mesh_as_ply(td + '/a.ply', g['v'], g['f'], g['n'], g['c'])
            assert np.array_equal(np.frombuffer(open(td + '/a.ply', 'rb').read(), np.uint8), g['bytes_nc'])
    finally:
        patch.uninstall()
    assert all(getattr(c, n) is f for (c, n), f in orig.items()) and mods['utils.recon_util'].recon_mesh is orig_rm
    assert mods['utils.renderer'].Renderer is fr.GLRenderer


def test_install_rebinds_driver_module_copies(frame):
    """main.py:19,21 copy `Renderer` / `canonicalize_normal_map` into the driver's namespace at import time"""
    import types
    from avatarcap_b200 import patch
    mods = fr.make_modules(frame)
    main = types.ModuleType('main')
    main.Renderer = mods['utils.renderer'].Renderer
    main.canonicalize_normal_map = mods['normal_fusion.normal_fusion'].canonicalize_normal_map
    had = sys.modules.get('main')
    sys.modules['main'] = main
    try:
        patch.install(engine=object(), modules=mods)
        assert main.Renderer is mods['utils.renderer'].Renderer and main.Renderer is not fr.GLRenderer
        assert main.canonicalize_normal_map is mods['normal_fusion.normal_fusion'].canonicalize_normal_map
        patch.uninstall()
        assert main.Renderer is fr.GLRenderer
    finally:
        patch.uninstall()
        if had is None:
            del sys.modules['main']
        else:
            sys.modules['main'] = had


@pytest.mark.gpu
def test_patched_call_sites_run_on_the_library(frame):
    from avatarcap_b200 import patch
    from avatarcap_b200.engine import Engine
    dev = torch.device('cuda', 0)
    eng = Engine(dev)
    mods = fr.make_modules(frame, dev)
    aa = mods['network.arch_avatar']; ar = mods['network.arch_recon']; su = mods['utils.smpl_util'].smpl_util
    patch.install(engine=eng, modules=mods)
    try:
        net = aa.GeoTexAvatar(frame).to(dev).eval(); occ_net = aa.OccupancyNet(net)          # main.py:296-298
        pts = torch.from_numpy(synth.volume_points(frame['cano_bounds'], (40, 40, 24)))[None].to(dev)
        batch = {'cano_pts': pts, 'smpl_pos_map': torch.from_numpy(synth.smpl_pos_map()).to(dev),
                 'cano_smpl_center': torch.from_numpy(frame['cano_smpl_center'])[None].to(dev),
                 'cano_bounds': torch.from_numpy(frame['cano_bounds'])[None].to(dev)}
        with torch.no_grad():
            net.warping_field.precompute_conv(batch)                                   # main.py:359
            fmap = net.warping_field.pose_feat_map
            assert tuple(fmap.shape) == (1, 64, 256, 256) and fmap.is_contiguous(memory_format=torch.channels_last)
            out = occ_net.query(batch)                                                 # main.py:360
            assert tuple(out['cano_pts_ov'].shape) == (1, pts.shape[1], 1) and tuple(out['nonrigid_offset'].shape) == (1, pts.shape[1], 3)
            eng.set_pose_feature_map(fmap.contiguous())
            ref = eng.eval_occupancy(pts[0], frame['cano_smpl_center'])
            assert torch.equal(out['cano_pts_ov'][0, :, 0], ref['occ']) and torch.equal(out['nonrigid_offset'][0], ref['off'])
            off = net.warping_field.query(pts, batch); assert torch.equal(off[0], ref['off'])
            # render stage through the patched names (main.py:330-331, 369, 408)
            R = mods['utils.renderer'].Renderer
            from avatarcap_b200 import render as render_mod
            nr = R(128, 128, shader_name='vertex_attribute', window_name='Normal'); pr = R(128, 128, shader_name='position', window_name='Position')
            assert isinstance(nr, render_mod.Renderer) and isinstance(pr, render_mod.Renderer) and nr.engine is eng
            tri_v = np.array([[-0.5, -0.5, 0], [0.5, -0.5, 0], [0, 0.5, 0]], np.float32); tri_n = np.array([[0, 0, 1]] * 3, np.float32)
            fimg, bimg = mods['utils.visualize_util'].render_cano_mesh(nr, tri_v, tri_n, np.array([[0, 1, 2]], np.int32), np.zeros(3))
            assert fimg.shape == (128, 128, 3) and fimg[..., 2].max() == 1.0 and not bimg.any()        # the back view culls the triangle
            rgb, alpha, occ = net.cano_template(pts + off)
            assert tuple(rgb.shape) == (1, pts.shape[1], 3) and float((occ[0, :, 0] - ref['occ']).abs().max()) < 1e-4
            # reconstruction network: HGFilter through the graph encoder, decoder through the field kernel
            rn = ar.ReconNetwork().to(dev)
            nm = torch.from_numpy(synth.normal_maps()).to(dev)
            items = {'cano_pts': pts, 'front_normal': nm[:, :3], 'back_normal': nm[:, 3:], 'cano_smpl_center': batch['cano_smpl_center']}
            ov = rn.infer(items)
            assert tuple(ov.shape) == (1, pts.shape[1]) and 0.0 <= float(ov.min()) and float(ov.max()) <= 1.0 and float(ov.std()) > 1e-3
            # two networks alternating per frame (main.py:307-315): each keeps its own resident weight slot, switching is a
            # pointer swap (no re-pack), and a sub-module call evaluates with ITS owner's weights
            net2 = aa.GeoTexAvatar(frame).to(dev).eval()
            with torch.no_grad():
                for p_ in net2.cano_template.parameters():
                    p_.mul_(1.01)
            net2.warping_field.pose_feat_map = fmap
            out2 = aa.OccupancyNet(net2).query(batch)
            assert float((out2['cano_pts_ov'] - out['cano_pts_ov']).abs().max()) > 1e-4
            cache = patch._cache('avatar')
            assert len(cache.entries) == 2 and cache.entries[id(net)][2] != cache.entries[id(net2)][2]
            import avatarcap_b200.packer as packer_mod
            calls = []
            orig_pack = packer_mod.pack_avatar
            packer_mod.pack_avatar = lambda sd: (calls.append(1), orig_pack(sd))[1]
            try:
                again = occ_net.query(batch); again2 = aa.OccupancyNet(net2).query(batch)
                r1, _, o1 = net.cano_template(pts + off); r2, _, o2 = net2.cano_template(pts + off)
            finally:
                packer_mod.pack_avatar = orig_pack
            assert not calls                                                           # cached blobs: nothing was re-packed
            assert torch.equal(again['cano_pts_ov'], out['cano_pts_ov']) and torch.equal(again2['cano_pts_ov'], out2['cano_pts_ov'])
            assert torch.equal(r1, rgb) and not torch.equal(o1, o2)
            # mesh + skinning call sites
            vol = out['cano_pts_ov'].reshape(40, 40, 24)
            v, f, n = mods['utils.recon_util'].recon_mesh(vol, (40, 40, 24), frame['cano_bounds'], 0.0)
            assert isinstance(v, np.ndarray) and v.dtype == np.float32 and f.dtype == np.int32 and v.shape == n.shape and len(f) > 0
            vt = torch.from_numpy(v)[None].to(dev)
            lbs = su.calculate_lbs(vt); assert tuple(lbs.shape) == (1, len(v), 24)
            jm = torch.from_numpy(frame['cano2live_jnt_mats'])[None].to(dev)
            lv, mats = su.skinning(vt, lbs, jm, return_pt_mats=True)
            ln = su.skinning_normal(torch.from_numpy(n)[None].to(dev), lbs, jm)
            assert tuple(lv.shape) == (1, len(v), 3) and tuple(mats.shape) == (1, len(v), 4, 4) and tuple(ln.shape) == (1, len(v), 3)
    finally:
        patch.uninstall()
        eng.close()
