"""Drop-in installation into an unmodified AvatarCap checkout.

    import avatarcap_b200.patch as avc_patch
    avc_patch.install()            # before main.run_avatarcap(...); `python main.py -c cfg -m test` is unchanged

`install()` re-binds the reference's hot-path call sites to the CUDA library. The replacements are active only while
autograd is disabled (`torch.no_grad()` / inference) AND the module is in eval mode: the training loop (main.py:97-116)
differentiates through the same modules, and finetune_tex (main.py:174-231) queries a train-mode network under no_grad
(batch-statistics BatchNorm). In both cases every patched method falls through to the original PyTorch code.

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


def _version(module) -> int:
    return sum(int(p._version) for p in module.parameters()) + sum(int(b._version) for b in module.buffers())


class _SlotCache:
    """Per-module cache of the packed weights in the library's resident slots (avc_load_weights_slot / avc_select_weights):
    main.py's test loop alternates `network` and `network_finetuned` every frame (main.py:307-315), which with a single slot
    meant a re-pack + cudaFree + cudaMalloc + synchronous upload twice per frame. Entries are keyed by id(module) and hold a
    weakref, so a recycled id is never mistaken for the module that owned it; least-recently-used eviction."""

    def __init__(self, kind: str):
        self.kind = kind
        self.entries: Dict[int, list] = {}     # id -> [weakref, version, slot, last_use]
        self.clock = 0
        self.active = None

    def use(self, net) -> None:
        import weakref
        from . import _lib
        e = _engine()
        self.clock += 1
        ent = self.entries.get(id(net))
        ver = _version(net)
        if ent is not None and ent[0]() is net and ent[1] == ver:
            ent[3] = self.clock
            if self.active != ent[2]:
                (e.select_avatar if self.kind == 'avatar' else e.select_recon)(ent[2]); self.active = ent[2]
            return
        for k in [k for k, v in self.entries.items() if v[0]() is None]:          # dead modules free their slots
            del self.entries[k]
        if ent is not None and ent[0]() is net:
            slot = ent[2]                                                          # same module, new parameters: reload in place
        else:
            used = {v[2] for v in self.entries.values()}
            free = [s for s in range(_lib.WEIGHT_SLOTS) if s not in used]
            if free:
                slot = free[0]
            else:
                victim = min(self.entries, key=lambda k: self.entries[k][3])
                slot = self.entries.pop(victim)[2]
        (e.load_avatar if self.kind == 'avatar' else e.load_recon)(net.state_dict(), slot)
        self.entries[id(net)] = [weakref.ref(net), ver, slot, self.clock]
        self.active = slot
        if self.kind == 'avatar':
            # WarpingField / DoubleTNet have no back-reference to their GeoTexAvatar: remember whose they are so that a direct
            # no_grad call on a sub-module evaluates with ITS network's weights (or falls through when it belongs to none)
            owners = _state.setdefault('owners', {})
            for k in [k for k, v in owners.items() if v() is None or v() is net]:
                del owners[k]
            for sub in (getattr(net, 'warping_field', None), getattr(net, 'cano_template', None)):
                if sub is not None:
                    owners[id(sub)] = weakref.ref(net)
                    _state.setdefault('owner_subs', {})[id(sub)] = weakref.ref(sub)


def _cache(kind: str) -> _SlotCache:
    key = 'cache_' + kind
    if key not in _state:
        _state[key] = _SlotCache(kind)
    return _state[key]


def _avatar_loaded(net) -> None:
    """Make `net`'s packed weights the active avatar blob ((re)packing only when its parameters changed)."""
    _cache('avatar').use(net)


def _recon_loaded(net) -> None:
    _cache('recon').use(net)


def _owner_of(sub):
    """The GeoTexAvatar a WarpingField / DoubleTNet instance was packed with, or None."""
    ref = _state.get('owners', {}).get(id(sub))
    sref = _state.get('owner_subs', {}).get(id(sub))
    if ref is None or sref is None or sref() is not sub:
        return None
    return ref()


def _mode_dependent(module) -> bool:
    """True when train / eval mode changes the module's forward: BatchNorm with running statistics (recognised by its
    `running_mean` buffer) or dropout. GroupNorm and weight-norm are mode-free: main.py never puts `recon_net` into eval mode
    (main.py:300) and its GroupNorm HGFilter + weight-normed decoder evaluate identically either way."""
    import weakref
    cache = _state.setdefault('mode_dep', {})
    ent = cache.get(id(module))
    if ent is None or ent[0]() is not module:
        dep = any(n.rsplit('.', 1)[-1] == 'running_mean' for n, _ in module.named_buffers()) or \
            any(isinstance(m, torch.nn.modules.dropout._DropoutNd) for m in module.modules())
        ent = (weakref.ref(module), dep)
        cache[id(module)] = ent
    return ent[1]


def _inference(*modules) -> bool:
    """The replacements fold BatchNorm running statistics (eval mode) and have no autograd: they are only valid when grad is
    off AND no module involved is a train-mode module whose forward depends on the mode. finetune_tex (main.py:174-231)
    queries a train-mode `network_init` under torch.no_grad() -- batch-statistics BatchNorm -- and must keep the reference's
    PyTorch path."""
    if torch.is_grad_enabled():
        return False
    return all(not (m.training and _mode_dependent(m)) for m in modules if m is not None)


def _encoder(slot: str, module, cls):
    """CUDA-graph encoder per PyTorch module (dict keyed by id + weakref, rebuilt when the parameters change): two networks
    alternating per frame no longer rebuild + re-capture each other's encoder."""
    import weakref
    cache = _state.setdefault('enc_' + slot, {})
    ent = cache.get(id(module))
    ver = _version(module)
    if ent is None or ent[0]() is not module or ent[1] != ver:
        for k in [k for k, v in cache.items() if v[0]() is None]:
            del cache[k]
        ent = [weakref.ref(module), ver, cls(module.state_dict(), device=_engine().device)]
        cache[id(module)] = ent
    return ent[2]


def _weight_volume(cwv):
    """(X,Y,Z,24) copy of CanoBlendWeightVolume.base_weight_volume, cached per tensor version (the reference's
    NerfRenderer.render calls forward once per 2048-ray chunk -- ~100 full-volume permute copies per frame otherwise)."""
    t = cwv.base_weight_volume
    key = (t.data_ptr(), int(t._version), tuple(t.shape))
    if _state.get('wv_key') != key:
        _state['wv'] = t[0].permute(1, 2, 3, 0).contiguous(); _state['wv_key'] = key
    return _state['wv']


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
        if not _inference(self.net):
            return o_query(self, batch)
        _avatar_loaded(self.net)
        return api.occupancy_query(_engine(), batch, self.net.warping_field.pose_feat_map, impl=impl)
    arch_avatar.OccupancyNet.query = query

    o_wq = keep(arch_avatar.WarpingField, 'query')
    def wquery(self, pts, batch):
        owner = _owner_of(self)
        if owner is None or not _inference(self, owner):
            return o_wq(self, pts, batch)
        _avatar_loaded(owner)
        return api.warping_field_query(_engine(), pts, batch, self.pose_feat_map, impl=impl)
    arch_avatar.WarpingField.query = wquery

    o_tf = keep(arch_avatar.DoubleTNet, 'forward')
    def tforward(self, pts):
        owner = _owner_of(self)
        if owner is None or not _inference(self, owner):
            return o_tf(self, pts)
        _avatar_loaded(owner)
        return api.template_forward(_engine(), pts, impl=impl)
    arch_avatar.DoubleTNet.forward = tforward

    o_gf = keep(arch_avatar.GeoTexAvatar, 'forward')
    def gforward(self, wpts, viewdirs, dists, batch, pts_space='posed'):
        if not _inference(self):
            return o_gf(self, wpts, viewdirs, dists, batch, pts_space)
        _avatar_loaded(self)
        su = smpl_util_mod.smpl_util
        wv = _weight_volume(self.cano_weight_volume) if pts_space == 'posed' else None            # (X,Y,Z,24), only the posed warp reads it
        return api.geotex_forward(_engine(), wpts, dists, batch, self.warping_field.pose_feat_map, su.smpl_skinning_weights,
                                  su.cano_smpl_vertices, wv, pts_space, impl=impl)
    arch_avatar.GeoTexAvatar.forward = gforward

    if encoders:
        o_pc = keep(arch_avatar.WarpingField, 'precompute_conv')
        def precompute_conv(self, batch):
            if not _inference(self) or not batch['smpl_pos_map'].is_cuda:
                return o_pc(self, batch)
            # clone: the encoder owns its output buffer and the reference keeps pose_feat_map across calls (arch_avatar.py:111)
            x = batch['smpl_pos_map']
            e = _engine()
            if e.has_tensor_core_path and x.shape[0] == 1 and x.shape[2] % 128 == 0 and x.shape[3] % 128 == 0:
                # the UNet as one program of library kernels (split-K fp32 gather-GEMMs + tcgen05 3x3 convolutions), encoders.build_unet_program
                hw = (int(x.shape[2]), int(x.shape[3]))
                cls = lambda sd, device: enc_mod.PoseFeatureEncoderTC(sd, engine=e, in_hw=hw)      # noqa: E731
                self.pose_feat_map = _encoder('unettc%dx%d' % hw, self.unet, cls)(x).clone(memory_format=torch.preserve_format)
            else:
                self.pose_feat_map = _encoder('unet', self.unet, enc_mod.PoseFeatureEncoder)(x).clone(memory_format=torch.preserve_format)
        arch_avatar.WarpingField.precompute_conv = precompute_conv

        o_gfm = keep(arch_recon.ReconNetwork, 'get_feat_maps')
        def get_feat_maps(self, image):
            if not _inference(self) or not image.is_cuda:
                return o_gfm(self, image)
            e = _engine()
            if e.has_tensor_core_path and image.shape[0] == 1 and image.shape[2] == image.shape[3] and image.shape[2] % 32 == 0:
                # HGFilter on the library's own tcgen05 convolutions (csrc/conv_tc.cu), one program + CUDA graph per module and input size
                hw = (int(image.shape[2]), int(image.shape[3]))
                cls = lambda sd, device: enc_mod.ImageFeatureEncoderTC(sd, engine=e, in_hw=hw)      # noqa: E731
                return [_encoder('hgtc%dx%d' % hw, self.image_encoder, cls)(image)]
            return [_encoder('hg', self.image_encoder, enc_mod.ImageFeatureEncoder)(image)]      # list like HGFilter's `outputs`
        arch_recon.ReconNetwork.get_feat_maps = get_feat_maps

    o_inf = keep(arch_recon.ReconNetwork, 'infer')
    def infer(self, items):
        if self.training and _mode_dependent(self):         # see _inference; the stock ReconNetwork is mode-free
            return o_inf(self, items)
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
