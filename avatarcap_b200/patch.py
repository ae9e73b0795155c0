"""Drop-in installation into an unmodified AvatarCap checkout.

    import avatarcap_b200.patch as avc_patch
    avc_patch.install()            # before main.run_avatarcap(...); `python main.py -c cfg -m test` is unchanged

`install()` re-binds the reference's hot-path call sites to the CUDA library. The replacements are active only while
autograd is disabled (`torch.no_grad()` / inference), because the training loop (main.py:97-116) differentiates
through the same modules; with grad enabled every patched method falls through to the original PyTorch code.

    network.arch_avatar.OccupancyNet.query      (arch_avatar.py:356)
    network.arch_avatar.WarpingField.query      (arch_avatar.py:113)
    network.arch_avatar.DoubleTNet.forward      (arch_avatar.py:65)
    network.arch_avatar.GeoTexAvatar.forward    (arch_avatar.py:178)
    network.arch_recon.ReconNetwork.infer       (arch_recon.py:45)
    utils.recon_util.recon_mesh                 (recon_util.py:51)
    utils.smpl_util.SmplUtil.calculate_lbs / skinning / skinning_normal   (smpl_util.py:24,58,76)
    network.arch_avatar.WarpingField.precompute_conv   (arch_avatar.py:109)   -> encoders.PoseFeatureEncoder  (install(encoders=True))
    network.arch_recon.ReconNetwork.get_feat_maps      (arch_recon.py:41)     -> encoders.ImageFeatureEncoder
    utils.renderer.Renderer                     (renderer.py:326)   'vertex_attribute' / 'position' -> render.Renderer (no GL); phong_* -> the original
    utils.visualize_util.render_cano_mesh       (visualize_util.py:11)   when handed a render.Renderer
    normal_fusion.normal_fusion.canonicalize_normal_map   (normal_fusion.py:12)   when handed render.Renderer objects
    utils.obj_io.save_mesh_as_ply               (obj_io.py:223)
main.py binds `Renderer` and `canonicalize_normal_map` with `from ... import ...` (main.py:19,21): an already imported `main` /
`__main__` module is re-bound too.
"""
from __future__ import annotations

import importlib
import sys
from typing import Dict, Optional

import torch

from . import api, encoders as enc_mod
from .engine import Engine, default_engine

_originals: Dict[str, object] = {}
_state: Dict[str, object] = {}


def _engine() -> Engine:
    return _state.get('engine') or default_engine()


def _avatar_loaded(net) -> None:
    """(Re)pack the avatar weights when the module's parameters changed (main.py loads two checkpoints, :304-315)."""
    key = (id(net), sum(int(p._version) for p in net.parameters()))
    if _state.get('avatar_key') != key:
        _engine().load_avatar(net.state_dict()); _state['avatar_key'] = key


def _recon_loaded(net) -> None:
    key = (id(net), sum(int(p._version) for p in net.parameters()))
    if _state.get('recon_key') != key:
        _engine().load_recon(net.state_dict()); _state['recon_key'] = key


def _encoder(slot: str, module, cls):
    """CUDA-graph encoder rebuilt when the PyTorch module's parameters change (same keying as the weight packers)."""
    key = (id(module), sum(int(p._version) for p in module.parameters()))
    if _state.get(slot + '_key') != key:
        _state[slot] = cls(module.state_dict(), device=_engine().device); _state[slot + '_key'] = key
    return _state[slot]


def _rebind_importers(name: str, original, replacement) -> None:
    """`from module import name` copies the binding: re-point the driver script's copy (main.py:19,21) as well."""
    for mod_name in ('main', '__main__'):
        m = sys.modules.get(mod_name)
        if m is not None and getattr(m, name, None) is original:
            _originals['%s.%s' % (mod_name, name)] = (m, name, original)
            setattr(m, name, replacement)


def install(engine: Optional[Engine] = None, impl: Optional[str] = None, modules: Optional[Dict[str, object]] = None,
            encoders: bool = True, render: bool = True) -> None:
    """Patch the reference modules in sys.path (or the ones passed in `modules`, keyed by dotted name). `encoders=False`
    leaves the per-frame UNet / HGFilter on the reference's own nn.Modules; `render=False` leaves the OpenGL renderer, the
    normal canonicalisation and the PLY writer alone."""
    if _originals:
        return
    _state['engine'] = engine
    get = (lambda name: modules[name]) if modules else importlib.import_module

    def get_optional(name):
        if modules is not None:
            return modules.get(name)
        try:
            return importlib.import_module(name)
        except ImportError:                      # e.g. no glfw / PyOpenGL on a render-less box: nothing to re-bind there
            return None
    arch_avatar = get('network.arch_avatar'); arch_recon = get('network.arch_recon')
    recon_util = get('utils.recon_util'); smpl_util_mod = get('utils.smpl_util')

    def keep(owner, name):
        _originals[owner.__name__ + '.' + name] = (owner, name, getattr(owner, name))
        return getattr(owner, name)

    o_query = keep(arch_avatar.OccupancyNet, 'query')
    def query(self, batch):
        if torch.is_grad_enabled():
            return o_query(self, batch)
        _avatar_loaded(self.net)
        return api.occupancy_query(_engine(), batch, self.net.warping_field.pose_feat_map, impl=impl)
    arch_avatar.OccupancyNet.query = query

    o_wq = keep(arch_avatar.WarpingField, 'query')
    def wquery(self, pts, batch):
        if torch.is_grad_enabled() or _state.get('avatar_key') is None:
            return o_wq(self, pts, batch)
        return api.warping_field_query(_engine(), pts, batch, self.pose_feat_map, impl=impl)
    arch_avatar.WarpingField.query = wquery

    o_tf = keep(arch_avatar.DoubleTNet, 'forward')
    def tforward(self, pts):
        if torch.is_grad_enabled() or _state.get('avatar_key') is None:
            return o_tf(self, pts)
        return api.template_forward(_engine(), pts, impl=impl)
    arch_avatar.DoubleTNet.forward = tforward

    o_gf = keep(arch_avatar.GeoTexAvatar, 'forward')
    def gforward(self, wpts, viewdirs, dists, batch, pts_space='posed'):
        if torch.is_grad_enabled():
            return o_gf(self, wpts, viewdirs, dists, batch, pts_space)
        _avatar_loaded(self)
        su = smpl_util_mod.smpl_util
        wv = self.cano_weight_volume.base_weight_volume[0].permute(1, 2, 3, 0).contiguous()      # (X,Y,Z,24)
        return api.geotex_forward(_engine(), wpts, dists, batch, self.warping_field.pose_feat_map, su.smpl_skinning_weights,
                                  su.cano_smpl_vertices, wv, pts_space, impl=impl)
    arch_avatar.GeoTexAvatar.forward = gforward

    if encoders:
        o_pc = keep(arch_avatar.WarpingField, 'precompute_conv')
        def precompute_conv(self, batch):
            if torch.is_grad_enabled() or not batch['smpl_pos_map'].is_cuda:
                return o_pc(self, batch)
            # clone: the graph owns its output buffer and the reference keeps pose_feat_map across calls (arch_avatar.py:111)
            self.pose_feat_map = _encoder('unet', self.unet, enc_mod.PoseFeatureEncoder)(batch['smpl_pos_map']).clone(memory_format=torch.preserve_format)
        arch_avatar.WarpingField.precompute_conv = precompute_conv

        o_gfm = keep(arch_recon.ReconNetwork, 'get_feat_maps')
        def get_feat_maps(self, image):
            if torch.is_grad_enabled() or not image.is_cuda:
                return o_gfm(self, image)
            return [_encoder('hg', self.image_encoder, enc_mod.ImageFeatureEncoder)(image)]      # list like HGFilter's `outputs`
        arch_recon.ReconNetwork.get_feat_maps = get_feat_maps

    o_inf = keep(arch_recon.ReconNetwork, 'infer')
    def infer(self, items):
        with torch.no_grad():
            _recon_loaded(self)
            imgs = torch.cat([items['front_normal'], items['back_normal']], dim=1)            # arch_recon.py:51
            fmap = self.get_feat_maps(imgs)[-1]                                                 # HGFilter stays in PyTorch
            return api.recon_infer(_engine(), items, fmap, impl=impl)
    arch_recon.ReconNetwork.infer = infer

    o_rm = keep(recon_util, 'recon_mesh')
    def recon_mesh(occ_volume, volume_res, bounds, iso_value=0.5):
        return api.recon_mesh(_engine(), occ_volume, volume_res, bounds, iso_value)
    recon_util.recon_mesh = recon_mesh

    SU = smpl_util_mod.SmplUtil
    o_lbs = keep(SU, 'calculate_lbs'); o_sk = keep(SU, 'skinning'); o_sn = keep(SU, 'skinning_normal')
    def calculate_lbs(self, points):
        if torch.is_grad_enabled() and points.requires_grad:
            return o_lbs(self, points)
        if self.cano_smpl_vertices is None:
            raise ValueError('Canonical smpl vertices are invalid!')
        e = _engine()
        return torch.stack([e.lbs_weights(points[b], self.cano_smpl_vertices, self.smpl_skinning_weights) for b in range(points.shape[0])], 0)
    def skinning(self, points, lbs, jnt_mats, return_pt_mats=False):
        if torch.is_grad_enabled() and (points.requires_grad or lbs.requires_grad):
            return o_sk(self, points, lbs, jnt_mats, return_pt_mats)
        e = _engine()
        outs = [e.skin_points(points[b], lbs[b], jnt_mats[b], return_pt_mats) for b in range(points.shape[0])]
        if return_pt_mats:
            return torch.stack([o[0] for o in outs], 0), torch.stack([o[1] for o in outs], 0)
        return torch.stack(outs, 0)
    def skinning_normal(self, normals, lbs, cano2live_jnt_mats):
        if torch.is_grad_enabled() and (normals.requires_grad or lbs.requires_grad):
            return o_sn(self, normals, lbs, cano2live_jnt_mats)
        e = _engine()
        return torch.stack([e.skin_normals(normals[b], lbs[b], cano2live_jnt_mats[b]) for b in range(normals.shape[0])], 0)
    SU.calculate_lbs = calculate_lbs; SU.skinning = skinning; SU.skinning_normal = skinning_normal


    if render:
        _install_render(get_optional, keep)


def _install_render(get_optional, keep) -> None:
    from . import mesh_io, render as render_mod
    renderer_mod = get_optional('utils.renderer'); vis = get_optional('utils.visualize_util')
    nf = get_optional('normal_fusion.normal_fusion'); obj_io = get_optional('utils.obj_io')
    if renderer_mod is not None:
        o_R = keep(renderer_mod, 'Renderer')

        class Renderer:
            """utils/renderer.py:326 -- the two data-path shaders run in the CUDA library, the phong previews stay on OpenGL."""
            def __new__(cls, img_w, img_h, mvp=None, shader_name='vertex_attribute', bg_color=(0, 0, 0), window_name=''):
                if shader_name in ('vertex_attribute', 'position'):
                    return render_mod.Renderer(img_w, img_h, mvp, shader_name, bg_color, window_name, engine=_engine())
                args = (img_w, img_h) if mvp is None else (img_w, img_h, mvp)
                return o_R(*args, shader_name=shader_name, bg_color=bg_color, window_name=window_name)
        renderer_mod.Renderer = Renderer
        _rebind_importers('Renderer', o_R, Renderer)
    if vis is not None:
        o_rcm = keep(vis, 'render_cano_mesh')
        def render_cano_mesh(renderer, vertices, normals, faces, mesh_center=None, colors=None):
            if not isinstance(renderer, render_mod.Renderer):
                return o_rcm(renderer, vertices, normals, faces, *(() if mesh_center is None else (mesh_center,)), colors=colors)
            import numpy as np
            return render_mod.render_cano_mesh(renderer, vertices, normals, faces, np.zeros(3) if mesh_center is None else mesh_center, colors)
        vis.render_cano_mesh = render_cano_mesh
    if nf is not None:
        o_cnm = keep(nf, 'canonicalize_normal_map')
        def canonicalize_normal_map(pos_renderer, attri_renderer, *args, **kwargs):
            if isinstance(pos_renderer, render_mod.Renderer) and isinstance(attri_renderer, render_mod.Renderer):
                return render_mod.canonicalize_normal_map(pos_renderer, attri_renderer, *args, **kwargs)
            return o_cnm(pos_renderer, attri_renderer, *args, **kwargs)
        nf.canonicalize_normal_map = canonicalize_normal_map
        _rebind_importers('canonicalize_normal_map', o_cnm, canonicalize_normal_map)
    if obj_io is not None:
        keep(obj_io, 'save_mesh_as_ply')
        obj_io.save_mesh_as_ply = mesh_io.save_mesh_as_ply


def uninstall() -> None:
    for owner, name, fn in _originals.values():
        setattr(owner, name, fn)
    _originals.clear(); _state.clear()
