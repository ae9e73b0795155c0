"""Host-side mirror of the reference's operator interface for the hot path (same names, arguments, return types,
dict keys and error behaviour). Everything here forwards to the CUDA library through
`Engine`; nothing is computed in PyTorch except trivial glue (reshape, 4x4 inverse of 24 matrices).

    reference symbol                                   mirror
    -------------------------------------------------  ------------------------------------------
    OccupancyNet.query(batch)   arch_avatar.py:356     OccupancyNet.query / occupancy_query
    WarpingField.query(pts, batch)          :113       warping_field_query
    DoubleTNet.forward(pts)                 :65        template_forward
    GeoTexAvatar.forward(...)               :178       geotex_forward
    ReconNetwork.infer(items)   arch_recon.py:45       recon_infer
    recon_util.recon_mesh(...)  recon_util.py:51       recon_mesh
    SmplUtil.calculate_lbs/skinning/skinning_normal    SmplUtil  (smpl_util.py:24,58,76)
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from .engine import Engine, default_engine


def _if_type(if_type: Optional[str]) -> str:
    if if_type is None:
        try:
            import config  # the reference's config module, when running inside the reference tree
            if_type = config.if_type
        except Exception:
            if_type = 'sdf'
    if if_type not in ('sdf', 'occupancy'):
        raise ValueError('Invalid config.if_type!')          # arch_avatar.py:82
    return if_type


# ------------------------------------------------------------------------------------------------------------
def occupancy_query(engine: Engine, batch: Dict, pose_feat_map: torch.Tensor, if_type: Optional[str] = None,
                    impl: Optional[str] = None) -> Dict[str, torch.Tensor]:
    """OccupancyNet.query (arch_avatar.py:356-381): {'cano_pts_ov': (B,N,1), 'nonrigid_offset': (B,N,3)}."""
    cano_pts = batch['cano_pts']
    B, N = cano_pts.shape[:2]
    occs, offs = [], []
    for b in range(B):
        engine.set_pose_feature_map(pose_feat_map[b:b + 1])
        o = engine.eval_occupancy(cano_pts[b], batch['cano_smpl_center'][b], want_offsets=True, if_type=_if_type(if_type), impl=impl)
        occs.append(o['occ'][None, :, None]); offs.append(o['off'][None])
    return {'cano_pts_ov': torch.cat(occs, 0), 'nonrigid_offset': torch.cat(offs, 0)}


def warping_field_query(engine: Engine, pts: torch.Tensor, batch: Dict, pose_feat_map: torch.Tensor, impl: Optional[str] = None) -> torch.Tensor:
    """WarpingField.query (arch_avatar.py:113-140): (B,N,3) -> (B,N,3)."""
    outs = []
    for b in range(pts.shape[0]):
        engine.set_pose_feature_map(pose_feat_map[b:b + 1])
        outs.append(engine.eval_warp(pts[b], batch['cano_smpl_center'][b], impl=impl)[None])
    return torch.cat(outs, 0)


def template_forward(engine: Engine, pts: torch.Tensor, if_type: Optional[str] = None, impl: Optional[str] = None):
    """DoubleTNet.forward (arch_avatar.py:65-83): (B,N,3) -> rgb (B,N,3), alpha (B,N,1), occ (B,N,1)."""
    B, N = pts.shape[:2]
    rgb, alpha, occ = engine.eval_template(pts.reshape(B * N, 3), _if_type(if_type), impl=impl)
    return rgb.reshape(B, N, 3), alpha.reshape(B, N, 1), occ.reshape(B, N, 1)


def recon_infer(engine: Engine, items: Dict, img_feat_map: torch.Tensor, impl: Optional[str] = None) -> torch.Tensor:
    """ReconNetwork.infer, per-point part (arch_recon.py:55-76). img_feat_map is get_feat_maps(imgs)[-1] (B,32,H,W).
    Returns (1,N) for B == 1 exactly like the reference ((B,1,N).squeeze(0), :74; the caller takes [0], main.py:442)."""
    cano_pts = items['cano_pts']
    outs = []
    for b in range(cano_pts.shape[0]):
        engine.set_image_feature_map(img_feat_map[b:b + 1])
        outs.append(engine.eval_recon(cano_pts[b], items['cano_smpl_center'][b], impl=impl)[None])
    out = torch.cat(outs, 0)[:, None, :]            # (B,1,N) like the decoder output
    return out.squeeze(0)


def recon_mesh(engine: Engine, occ_volume: torch.Tensor, volume_res, bounds: np.ndarray, iso_value: float = 0.5):
    """recon_util.recon_mesh (recon_util.py:51-70): -> vertices (V,3) f32, faces (F,3) i32, normals (V,3) f32 as numpy (host),
    exactly what the reference returns. Raises ValueError when iso lies outside the volume's range (skimage behaviour)."""
    vol = occ_volume.reshape(tuple(int(r) for r in volume_res))
    v, f, n = engine.extract_mesh(vol, np.asarray(bounds, dtype=np.float32), float(iso_value), with_normals=True)
    if v.shape[0] == 0:
        lo, hi = float(vol.min()), float(vol.max())
        if not (lo <= iso_value <= hi):
            raise ValueError('Surface level must be within volume data range.')
    return v.cpu().numpy(), f.cpu().numpy(), n.cpu().numpy()


class OccupancyNet:
    """Drop-in for network.arch_avatar.OccupancyNet (arch_avatar.py:352-381). `net` is the reference GeoTexAvatar module
    (or anything exposing .state_dict() and .warping_field.pose_feat_map)."""

    def __init__(self, net, engine: Optional[Engine] = None, impl: Optional[str] = None):
        self.net = net
        self.engine = engine or default_engine()
        self.impl = impl
        self.engine.load_avatar(net.state_dict())

    def query(self, batch):
        fmap = self.net.warping_field.pose_feat_map
        if fmap is None:
            raise RuntimeError('WarpingField.precompute_conv(batch) must run before OccupancyNet.query (main.py:359-360)')
        return occupancy_query(self.engine, batch, fmap, impl=self.impl)


class SmplUtil:
    """Drop-in for utils.smpl_util.SmplUtil (smpl_util.py:12-81); batch dimension must be 1 as in main.py."""

    def __init__(self, smpl_skinning_weights, engine: Optional[Engine] = None):
        self.engine = engine or default_engine()
        self.smpl_skinning_weights = self.engine._f32(smpl_skinning_weights, 24)
        self.cano_smpl_vertices = None

    def set_cano_smpl_vertices(self, cano_smpl_vertices: torch.Tensor):
        self.cano_smpl_vertices = self.engine._f32(cano_smpl_vertices, 3)

    def calculate_lbs(self, points: torch.Tensor) -> torch.Tensor:
        if self.cano_smpl_vertices is None:
            raise ValueError('Canonical smpl vertices are invalid!')     # smpl_util.py:30-31
        return torch.stack([self.engine.lbs_weights(points[b], self.cano_smpl_vertices, self.smpl_skinning_weights)
                            for b in range(points.shape[0])], 0)

    def skinning(self, points, lbs, jnt_mats, return_pt_mats: bool = False):
        outs = [self.engine.skin_points(points[b], lbs[b], jnt_mats[b], return_pt_mats) for b in range(points.shape[0])]
        if return_pt_mats:
            return torch.stack([o[0] for o in outs], 0), torch.stack([o[1] for o in outs], 0)
        return torch.stack(outs, 0)

    def skinning_normal(self, normals, lbs, cano2live_jnt_mats):
        return torch.stack([self.engine.skin_normals(normals[b], lbs[b], cano2live_jnt_mats[b]) for b in range(normals.shape[0])], 0)


def geotex_forward(engine: Engine, wpts: torch.Tensor, dists: torch.Tensor, batch: Dict, pose_feat_map: torch.Tensor,
                   smpl_skinning_weights: torch.Tensor, cano_smpl_vertices: torch.Tensor, weight_volume: torch.Tensor,
                   pts_space: str = 'posed', if_type: Optional[str] = None, impl: Optional[str] = None) -> Dict[str, torch.Tensor]:
    """GeoTexAvatar.forward (arch_avatar.py:178-237), inference only. weight_volume is the (X,Y,Z,24) array the reference
    loads from cano_base_blend_weight_volume.npy (:146-148). Like the reference, 'cano' mode adds the offsets to the caller's
    wpts tensor in place (:207,211-213)."""
    assert (pts_space == 'posed' or pts_space == 'cano' or pts_space == 'temp')          # :187
    B = wpts.shape[0]
    raws, occs, offs = [], [], []
    for b in range(B):
        w = engine._f32(wpts[b], 3)
        bounds = batch['cano_bounds'][b]
        if pts_space == 'posed':
            live2cano = torch.linalg.inv(batch['cano2live_jnt_mats'][b].to(engine.device, torch.float32))      # :199
            cano, near = engine.posed_to_cano(w, batch['live_smpl_v'][b], smpl_skinning_weights, live2cano, bounds, weight_volume)
        else:
            cano = w
            d2, _ = engine.knn(w, cano_smpl_vertices, 1)                                                       # :208
            near = d2[:, 0] < 0.08 * 0.08
        if pts_space in ('posed', 'cano'):
            engine.set_pose_feature_map(pose_feat_map[b:b + 1])
            o = engine.eval_occupancy(cano, batch['cano_smpl_center'][b], want_offsets=True, want_texture=True,
                                      if_type=_if_type(if_type), impl=impl)
            off = o['off']; rgb, alpha, occ = o['rgb'], o['alpha'], o['occ']
            cano = cano + off
            if pts_space == 'cano':
                wpts[b] += off.to(wpts.device)                                                                # in-place quirk
        else:
            off = torch.zeros_like(cano)
            rgb, alpha, occ = engine.eval_template(cano, _if_type(if_type), impl=impl)
        raw = engine.nerf_raw(cano, near, rgb, alpha, dists[b], bounds)                                       # :220-231
        raws.append(raw[None]); occs.append(occ[None, :, None]); offs.append(off[None])
    return {'raw': torch.cat(raws, 0), 'occ': torch.cat(occs, 0), 'nonrigid_offset': torch.cat(offs, 0)}


class NerfRenderer:
    """Mirror of network.arch_avatar.NerfRenderer (arch_avatar.py:240-349) for inference: ray sampling, the per-sample field
    (geotex_forward) and front-to-back compositing (utils/nerf_util.raw2outputs) all run in the CUDA library.
    `forward(wpts, dists, batch, pts_space)` is any callable with geotex_forward's result dict (see `for_engine`)."""

    def __init__(self, forward, engine: Optional[Engine] = None, n_samples: int = 64):
        self.forward = forward
        self.engine = engine or default_engine()
        self.n_samples = n_samples                     # config.N_samples (config.py:9)

    @classmethod
    def for_engine(cls, engine: Engine, pose_feat_map, smpl_skinning_weights, cano_smpl_vertices, weight_volume, n_samples: int = 64):
        def fwd(wpts, dists, batch, pts_space):
            return geotex_forward(engine, wpts, dists, batch, pose_feat_map, smpl_skinning_weights, cano_smpl_vertices, weight_volume, pts_space)
        return cls(fwd, engine, n_samples)

    def render(self, batch: Dict, pts_space: str = 'posed', near_dist: float = 0.05, far_dist: float = 0.05) -> Dict[str, torch.Tensor]:
        """render (arch_avatar.py:320-349) with B == 1 (as main.py:476 calls it). Like the reference, near/far of rays with a valid
        depth are overwritten IN PLACE in the caller's batch (get_pixel_value :285-287). Keys: rgb_map, acc_map, depth_map, raw, occ,
        nonrigid_offset."""
        e = self.engine
        ray_o, ray_d, near, far, depth = batch['ray_o'], batch['ray_d'], batch['near'], batch['far'], batch['depth']
        assert ray_o.shape[0] == 1, 'batch size 1'
        valid = depth > 1e-6
        near[valid] = depth[valid] - near_dist; far[valid] = depth[valid] + far_dist
        n_pixel = ray_o.shape[1]; S = self.n_samples
        outs = {k: [] for k in ('raw', 'occ', 'nonrigid_offset', 'rgb_map', 'acc_map', 'depth_map')}
        chunk = 1 << 15                                   # rays per call (the reference uses 2048, :330; chunking does not change results)
        for i in range(0, n_pixel, chunk):
            sl = slice(i, i + chunk)
            pts, z, dists = e.ray_samples(ray_o[0, sl], ray_d[0, sl], near[0, sl], far[0, sl], S)
            ret = self.forward(pts[None], dists[None, :, None], batch, pts_space)
            raw = ret['raw'][0]
            rgb, acc, dep = e.composite(raw, z)
            outs['raw'].append(raw); outs['occ'].append(ret['occ'][0]); outs['nonrigid_offset'].append(ret['nonrigid_offset'][0])
            outs['rgb_map'].append(rgb); outs['acc_map'].append(acc); outs['depth_map'].append(dep)
        return {k: torch.cat(v, 0)[None] for k, v in outs.items()}


def vertex_colors(renderer: NerfRenderer, batch: Dict, vertices: torch.Tensor, normals: torch.Tensor) -> torch.Tensor:
    """main.py:464-478: integrate the texture template along -normal through each avatar vertex; returns (V,3) colours in the
    reference's channel order ([:, [2,1,0]])."""
    items = dict(batch)
    items['ray_o'] = (vertices + normals)[None]; items['ray_d'] = -normals[None]
    items['depth'] = torch.ones((1, vertices.shape[0]), device=vertices.device, dtype=torch.float32)
    items['near'] = items['depth'] - 0.05; items['far'] = items['depth'] + 0.05
    out = renderer.render(items, pts_space='cano', near_dist=0.02, far_dist=0.05)
    return out['rgb_map'][0][:, [2, 1, 0]]


def transfer_colors(engine: Engine, vertices: torch.Tensor, src_vertices: torch.Tensor, src_colors: torch.Tensor) -> torch.Tensor:
    """main.py:480-484: nearest avatar vertex's colour for every reconstructed vertex (knn_points K=1 + knn_gather).
    The source set can be large (a mesh), so it is processed in tiles of the KNN kernel's reference capacity."""
    _, idx = engine.knn(vertices, src_vertices, 1)
    return src_colors[idx[:, 0]]
