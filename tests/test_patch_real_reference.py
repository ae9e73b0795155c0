"""CPU, build container only: avatarcap_b200.patch against the REAL reference modules (imported from /root/reference with the same
import stubs the golden generators use). Skipped where the reference is not mounted (the GPU box). Checks that the drop-in still
fits the reference as it is: every re-bound name exists with the same parameter list, and with autograd enabled (the training
path, main.py:97-116) the patched call sites fall through to the reference's own code bit for bit."""
import importlib.util
import inspect
import os
import sys
import tempfile

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('AVATARCAP_REFERENCE', '/root/reference')
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, 'main.py')), reason='reference checkout not mounted')

from avatarcap_b200 import synth  # noqa: E402


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m


@pytest.fixture(scope='module')
def ref_env():
    """Reference modules importable on the CPU: stubs for pytorch3d / dataset.smpl / skimage / trimesh / glfw / OpenGL."""
    saved = {k: sys.modules.get(k) for k in list(sys.modules)}
    saved_path = list(sys.path)
    gg = _load('_gen_golden', os.path.join(ROOT, 'tests', 'golden', 'gen_golden.py'))
    gr = _load('_gen_raster_golden', os.path.join(ROOT, 'tests', 'golden', 'gen_raster_golden.py'))
    body = synth.SynthBody(); frame = synth.make_frame(body, synth.random_pose(7, 0.4))
    tmp = tempfile.mkdtemp(prefix='avc_patch_real_')
    gg.install_stubs(body, tmp, synth.blend_weight_volume(frame))
    gr.install_stubs()                      # adds glfw / OpenGL / pytorch3d.transforms (keeps pytorch3d.ops from the line above? no:)
    gg.install_stubs(body, tmp, synth.blend_weight_volume(frame))     # ... re-install so that pytorch3d.ops is the KNN stub again
    import pytorch3d
    pytorch3d.transforms.axis_angle_to_matrix = gr.axis_angle_to_matrix
    yield {'frame': frame, 'gg': gg}
    # drop only what belongs to the reference tree and the stubs (third-party packages imported meanwhile, e.g. cv2, must stay whole)
    prefixes = ('config', 'network', 'utils', 'dataset', 'normal_fusion', 'pytorch3d', 'skimage', 'trimesh', 'glfw', 'OpenGL', '_gen_')
    for k in list(sys.modules):
        if k not in saved and (k in prefixes or k.split('.')[0] in prefixes or k.startswith('_gen_')):
            del sys.modules[k]
    sys.path[:] = saved_path


def test_patch_fits_the_real_reference(ref_env):
    from avatarcap_b200 import patch
    import network.arch_avatar as aa
    import network.arch_recon as ar
    import utils.recon_util as ru
    import utils.smpl_util as su
    import utils.renderer as rr
    import utils.visualize_util as vu
    import normal_fusion.normal_fusion as nf
    import utils.obj_io as oi
    sites = [(aa.OccupancyNet, 'query'), (aa.WarpingField, 'query'), (aa.WarpingField, 'precompute_conv'), (aa.DoubleTNet, 'forward'),
             (aa.GeoTexAvatar, 'forward'), (ar.ReconNetwork, 'infer'), (ar.ReconNetwork, 'get_feat_maps'), (ru, 'recon_mesh'),
             (su.SmplUtil, 'calculate_lbs'), (su.SmplUtil, 'skinning'), (su.SmplUtil, 'skinning_normal'), (vu, 'render_cano_mesh'),
             (oi, 'save_mesh_as_ply')]
    before = {(o, n): getattr(o, n) for o, n in sites}
    sig = {k: inspect.signature(f) for k, f in before.items()}
    frame = ref_env['frame']; gg = ref_env['gg']
    # reference outputs before patching
    net = aa.GeoTexAvatar().eval()
    net.load_state_dict(gg.to_torch_sd(synth.avatar_state_dict()), strict=False)
    net.warping_field.pose_feat_map = torch.from_numpy(synth.feature_map(64, 48, 40, synth.SEED + 2))[None]
    rs = np.random.RandomState(1)
    bmin, bmax = frame['cano_bounds']
    pts = torch.from_numpy((rs.uniform(0, 1, (257, 3)) * (bmax - bmin) + bmin).astype(np.float32))[None]
    batch = {'cano_pts': pts, 'cano_smpl_center': torch.from_numpy(frame['cano_smpl_center'])[None]}
    su.smpl_util.set_cano_smpl_vertices(torch.from_numpy(frame['cano_smpl_v']))
    with torch.no_grad():
        want = aa.OccupancyNet(net).query(batch)
        want_lbs = su.smpl_util.calculate_lbs(pts)
    patch.install(engine=object())                      # the engine is only touched under no_grad
    try:
        for (o, n), f in before.items():
            g = getattr(o, n)
            assert g is not f, (o, n)
            assert list(inspect.signature(g).parameters) == list(sig[(o, n)].parameters), (n, inspect.signature(g), sig[(o, n)])
        # defaults that callers rely on
        assert inspect.signature(ru.recon_mesh).parameters['iso_value'].default == 0.5                 # recon_util.py:51 / main.py:444
        assert inspect.signature(aa.GeoTexAvatar.forward).parameters['pts_space'].default == 'posed'   # arch_avatar.py:178
        assert inspect.signature(su.SmplUtil.skinning).parameters['return_pt_mats'].default is False   # smpl_util.py:58
        assert rr.Renderer is not None and nf.canonicalize_normal_map is not None
        # training path: autograd on -> the reference's own code runs, bit for bit
        with torch.enable_grad():
            got = aa.OccupancyNet(net).query(batch)
            p2 = pts.clone().requires_grad_()
            got_lbs = su.smpl_util.calculate_lbs(p2)
        assert torch.equal(got['cano_pts_ov'], want['cano_pts_ov']) and torch.equal(got['nonrigid_offset'], want['nonrigid_offset'])
        assert torch.equal(got_lbs.detach(), want_lbs)
        # attributes the patched methods read from the reference objects
        assert hasattr(net.warping_field, 'pose_feat_map') and hasattr(net.warping_field, 'unet') and hasattr(net, 'cano_template')
        assert tuple(net.cano_weight_volume.base_weight_volume.shape[:2]) == (1, 24)
        rn = ar.ReconNetwork()
        assert hasattr(rn, 'image_encoder') and hasattr(rn, 'image_decoder')
        assert hasattr(su.smpl_util, 'smpl_skinning_weights') and hasattr(su.smpl_util, 'cano_smpl_vertices')
        # the checkpoint keys the packers / encoders consume are the reference's
        from avatarcap_b200 import packer
        assert len(packer.pack_avatar(net.state_dict())) > 1000 and len(packer.pack_recon(rn.state_dict())) > 1000
        from avatarcap_b200 import encoders
        for k in synth.unet_state_dict():
            assert k in net.warping_field.unet.state_dict(), k
        for k in synth.hgfilter_state_dict():
            assert k in rn.image_encoder.state_dict(), k
        assert encoders.PoseFeatureEncoder is not None
    finally:
        patch.uninstall()
    assert all(getattr(o, n) is f for (o, n), f in before.items())


def test_projection_matrices_equal_the_reference(ref_env):
    """render.py / the oracle restate utils/renderer.py:296-323; compare with the reference's own functions"""
    import utils.renderer as rr
    from avatarcap_b200 import render
    from oracle import raster_oracle as ro
    for mod in (render, ro):
        assert np.array_equal(mod.gl_orthographic_projection_matrix(), rr.gl_orthographic_projection_matrix())
        assert np.array_equal(mod.gl_orthographic_projection_matrix(-50.0, -0.5), rr.gl_orthographic_projection_matrix(-50.0, -0.5))
        for gl_space in (False, True):
            a = mod.gl_perspective_projection_matrix(551.3, 548.9, 255.2, 260.7, 512, 480, gl_space=gl_space)
            b = rr.gl_perspective_projection_matrix(551.3, 548.9, 255.2, 260.7, 512, 480, gl_space=gl_space)
            assert a.dtype == b.dtype and np.array_equal(a, b)
        assert np.array_equal(mod.gl_perspective_projection_matrix(500, 500, 256, 256, 512, 512, 50.0, 0.2), rr.gl_perspective_projection_matrix(500, 500, 256, 256, 512, 512, 50.0, 0.2))


@pytest.mark.gpu
def test_patched_real_reference_runs_on_the_library(ref_env):
    """Where BOTH a B200 and a reference checkout are reachable (AVATARCAP_REFERENCE): the reference's own modules, moved to the
    device and patched, must produce the reference's results through the library -- OccupancyNet.query, ReconNetwork.infer's
    decoder, calculate_lbs / skinning, recon_mesh's vertex count -- and a train-mode twin must keep using PyTorch."""
    from avatarcap_b200 import patch
    from avatarcap_b200.engine import Engine
    import network.arch_avatar as aa
    import utils.smpl_util as su
    import config
    frame = ref_env['frame']; gg = ref_env['gg']
    dev = torch.device('cuda', 0)
    net = aa.GeoTexAvatar().eval()
    net.load_state_dict(gg.to_torch_sd(synth.avatar_state_dict()), strict=False)
    fmap = torch.from_numpy(synth.feature_map(64, 64, 64, synth.SEED + 2))[None]
    net.warping_field.pose_feat_map = fmap
    rs = np.random.RandomState(2)
    bmin, bmax = frame['cano_bounds']
    pts = torch.from_numpy((rs.uniform(0, 1, (5000, 3)) * (bmax - bmin) + bmin).astype(np.float32))[None]
    batch = {'cano_pts': pts, 'cano_smpl_center': torch.from_numpy(frame['cano_smpl_center'])[None]}
    su.smpl_util.set_cano_smpl_vertices(torch.from_numpy(frame['cano_smpl_v']))
    with torch.no_grad():
        want = aa.OccupancyNet(net).query(batch)                      # the reference, CPU f32
        want_lbs = su.smpl_util.calculate_lbs(pts)
    eng = Engine(dev)
    old_dev = config.device
    config.device = dev
    patch.install(engine=eng, encoders=False, render=False)
    try:
        net_d = net.to(dev); net_d.warping_field.pose_feat_map = fmap.to(dev)
        su.smpl_util.smpl_skinning_weights = su.smpl_util.smpl_skinning_weights.to(dev)
        su.smpl_util.set_cano_smpl_vertices(torch.from_numpy(frame['cano_smpl_v']).to(dev))
        batch_d = {k: v.to(dev) for k, v in batch.items()}
        launches = eng.launch_count
        with torch.no_grad():
            got = aa.OccupancyNet(net_d).query(batch_d)
            got_lbs = su.smpl_util.calculate_lbs(pts.to(dev))
        assert eng.launch_count > launches                            # it did go through the library
        assert float((got['cano_pts_ov'].cpu() - want['cano_pts_ov']).abs().max()) < 1e-4
        assert float((got['nonrigid_offset'].cpu() - want['nonrigid_offset']).abs().max()) < 2e-6
        assert float((got_lbs.cpu() - want_lbs).abs().max()) < 2e-6
        twin = aa.GeoTexAvatar().to(dev)                               # train mode, like finetune_tex's network_init
        twin.load_state_dict(net_d.state_dict()); twin.warping_field.pose_feat_map = fmap.to(dev)
        launches = eng.launch_count
        with torch.no_grad():
            aa.OccupancyNet(twin).query(batch_d)
        assert eng.launch_count == launches                           # batch-statistics BatchNorm: the reference's own path
    finally:
        patch.uninstall(); config.device = old_dev
        su.smpl_util.smpl_skinning_weights = su.smpl_util.smpl_skinning_weights.cpu()
        eng.close()
