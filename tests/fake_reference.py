"""Stand-ins for the reference's modules, just deep enough for avatarcap_b200.patch.install(modules=...):
same class / attribute / state_dict key names as network.arch_avatar, network.arch_recon, utils.recon_util and
utils.smpl_util, seeded synthetic weights, and ORIGINAL methods that raise -- so a test can tell whether a call went to
the CUDA library (returns) or fell through to "the reference" (raises Fallthrough). The real reference cannot be used in
the GPU tests: /root/reference does not exist on the GPU box."""
import types

import numpy as np
import torch
import torch.nn as nn

from avatarcap_b200 import synth


class Fallthrough(RuntimeError):
    pass


def bag(sd, prefix=''):
    """nn.Module tree whose state_dict() reproduces the dotted keys of `sd` below `prefix`."""
    root = nn.Module()
    for k, v in sd.items():
        if not k.startswith(prefix):
            continue
        parts = k[len(prefix):].split('.')
        m = root
        for p in parts[:-1]:
            if not hasattr(m, p):
                m.add_module(p, nn.Module())
            m = getattr(m, p)
        t = torch.from_numpy(np.ascontiguousarray(v)) if v.shape else torch.tensor(v)
        if t.dtype.is_floating_point and 'running' not in parts[-1]:
            m.register_parameter(parts[-1], nn.Parameter(t))
        else:
            m.register_buffer(parts[-1], t)
    return root


class WarpingField(nn.Module):
    def __init__(self, avatar_sd, unet_sd):
        super().__init__()
        self.unet = bag(unet_sd)
        self.mlp = bag(avatar_sd, 'warping_field.mlp.')
        self.out_layer_coord_affine = bag(avatar_sd, 'warping_field.out_layer_coord_affine.')
        self.pose_feat_map = None

    def precompute_conv(self, batch):
        raise Fallthrough('WarpingField.precompute_conv')

    def query(self, pts, batch):
        raise Fallthrough('WarpingField.query')


class DoubleTNet(nn.Module):
    def __init__(self, avatar_sd):
        super().__init__()
        for name in ('shared_mlp', 'geo_mlp', 'clr_mlp'):
            self.add_module(name, bag(avatar_sd, 'cano_template.%s.' % name))

    def forward(self, pts):
        raise Fallthrough('DoubleTNet.forward')


class _WeightVolume(nn.Module):
    def __init__(self, vol_xyzc):
        super().__init__()
        self.base_weight_volume = torch.from_numpy(vol_xyzc).permute(3, 0, 1, 2)[None].contiguous()      # (1,24,X,Y,Z) arch_avatar.py:174-176


class GeoTexAvatar(nn.Module):
    def __init__(self, frame):
        super().__init__()
        sd = synth.avatar_state_dict()
        self.cano_template = DoubleTNet(sd)
        self.warping_field = WarpingField(sd, synth.unet_state_dict())
        self.cano_weight_volume = _WeightVolume(synth.blend_weight_volume(frame))

    def forward(self, wpts, viewdirs, dists, batch, pts_space='posed'):
        raise Fallthrough('GeoTexAvatar.forward')


class OccupancyNet:
    def __init__(self, net):
        self.net = net

    def query(self, batch):
        raise Fallthrough('OccupancyNet.query')


class ReconNetwork(nn.Module):
    def __init__(self):
        super().__init__()
        self.image_encoder = bag(synth.hgfilter_state_dict())
        self.image_decoder = bag(synth.recon_state_dict(), 'image_decoder.')

    def get_feat_maps(self, image):
        raise Fallthrough('ReconNetwork.get_feat_maps')

    def infer(self, items):
        raise Fallthrough('ReconNetwork.infer')


class SmplUtil:
    def __init__(self, weights):
        self.smpl_skinning_weights = weights
        self.cano_smpl_vertices = None

    def set_cano_smpl_vertices(self, v):
        self.cano_smpl_vertices = v

    def calculate_lbs(self, points):
        raise Fallthrough('SmplUtil.calculate_lbs')

    def skinning(self, points, lbs, jnt_mats, return_pt_mats=False):
        raise Fallthrough('SmplUtil.skinning')

    def skinning_normal(self, normals, lbs, cano2live_jnt_mats):
        raise Fallthrough('SmplUtil.skinning_normal')


class GLRenderer:
    """utils/renderer.py Renderer: needs a GL context in the reference, so every use is a fall-through here."""
    def __init__(self, img_w, img_h, mvp=None, shader_name='vertex_attribute', bg_color=(0, 0, 0), window_name=''):
        self.img_w, self.img_h, self.shader_name = img_w, img_h, shader_name

    def render(self):
        raise Fallthrough('Renderer.render')


def make_render_modules():
    def render_cano_mesh(renderer, vertices, normals, faces, mesh_center=None, colors=None):
        raise Fallthrough('render_cano_mesh')

    def canonicalize_normal_map(pos_renderer, attri_renderer, *a, **k):
        raise Fallthrough('canonicalize_normal_map')

    def save_mesh_as_ply(path, vertices, faces=None, normals=None, colors=None):
        raise Fallthrough('save_mesh_as_ply')
    rm = types.ModuleType('utils.renderer'); rm.Renderer = GLRenderer
    vis = types.ModuleType('utils.visualize_util'); vis.render_cano_mesh = render_cano_mesh
    nf = types.ModuleType('normal_fusion.normal_fusion'); nf.canonicalize_normal_map = canonicalize_normal_map
    oi = types.ModuleType('utils.obj_io'); oi.save_mesh_as_ply = save_mesh_as_ply
    return {'utils.renderer': rm, 'utils.visualize_util': vis, 'normal_fusion.normal_fusion': nf, 'utils.obj_io': oi}


def make_modules(frame, device='cpu'):
    def recon_mesh(occ_volume, volume_res, bounds, iso_value=0.5):
        raise Fallthrough('recon_mesh')
    aa = types.ModuleType('network.arch_avatar')
    aa.OccupancyNet = OccupancyNet; aa.WarpingField = WarpingField; aa.DoubleTNet = DoubleTNet; aa.GeoTexAvatar = GeoTexAvatar
    ar = types.ModuleType('network.arch_recon'); ar.ReconNetwork = ReconNetwork
    ru = types.ModuleType('utils.recon_util'); ru.recon_mesh = recon_mesh
    su = types.ModuleType('utils.smpl_util'); su.SmplUtil = SmplUtil
    su.smpl_util = SmplUtil(torch.from_numpy(frame['smpl_skinning_weights']).to(device))
    su.smpl_util.set_cano_smpl_vertices(torch.from_numpy(frame['cano_smpl_v']).to(device))
    mods = {'network.arch_avatar': aa, 'network.arch_recon': ar, 'utils.recon_util': ru, 'utils.smpl_util': su}
    mods.update(make_render_modules())
    return mods
