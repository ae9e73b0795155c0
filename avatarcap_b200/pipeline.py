"""Per-frame geometry pipeline on the device: the hot-path part of main.run_avatarcap steps 1 and 3
(main.py:357-367, 383-389, 438-453) as one function each, built only from Engine calls.

    canonical avatar : OccupancyNet.query -> mask scatter (+-1 fill) -> recon_mesh(iso = config.iso_value = 0) -> LBS to live space
    reconstruction   : ReconNetwork.infer -> mask scatter -> recon_mesh(iso = 0.5) -> LBS (+ normals)

Everything stays in HBM between the stages (the reference copies the volume to the host for skimage and the vertices
back, recon_util.py:64, main.py:383). Also holds the dataset-side precompute of SURVEY.md section 8 row a15 (validity
flag by KNN-1, avatarcap_dataset.py:114-116).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from .engine import Engine


def valid_points_flag(engine: Engine, vol_pts: torch.Tensor, cano_smpl_v, thres: float = 0.1) -> torch.Tensor:
    """infer_pts_flag = knn_points(vol_pts, cano_smpl_v, K=1).dists < 0.1**2   (avatarcap_dataset.py:114-116)."""
    return engine.near_flag(vol_pts, cano_smpl_v, thres)


def invalid_points_fill(engine: Engine, smpl_verts, smpl_faces, bounds, vol_res, flag: torch.Tensor) -> torch.Tensor:
    """invalid_pts_ov = 2 * trimesh.contains(invalid_pts) - 1 for the grid points with flag == False, in grid order
    (avatarcap_dataset.py:120-124)."""
    inside = engine.inside_volume(smpl_verts, smpl_faces, bounds, vol_res).reshape(-1)
    return 2.0 * inside[~flag].to(torch.float32) - 1.0


def _mesh_to_live(engine: Engine, verts, normals, frame: Dict):
    return engine.skin_mesh(verts, normals, frame['cano_smpl_v'], frame['smpl_skinning_weights'], frame['cano2live_jnt_mats'])


def avatar_frame(engine: Engine, frame: Dict, pose_feat_map, vol_res, flag: Optional[torch.Tensor] = None,
                 pts: Optional[torch.Tensor] = None, fill: Optional[torch.Tensor] = None, iso: float = 0.0,
                 impl: Optional[str] = None) -> Dict[str, torch.Tensor]:
    """Step 1 of run_avatarcap for one frame. `frame` holds cano_bounds (2,3), cano_smpl_center (3,), cano_smpl_v,
    smpl_skinning_weights, cano2live_jnt_mats. Dense mode: flag/pts/fill None -> all Rx*Ry*Rz grid points are evaluated.
    Masked mode (the reference's): pts = grid[flag], fill = +-1 for the other voxels (main.py:362-363)."""
    bounds = np.asarray(frame['cano_bounds'], dtype=np.float32)
    engine.set_pose_feature_map(pose_feat_map)
    if flag is None:          # dense: coordinates from the grid index inside the kernel, no point list
        out = engine.eval_occupancy_grid(bounds, vol_res, frame['cano_smpl_center'], want_offsets=True, impl=impl)
    else:
        out = engine.eval_occupancy(pts, frame['cano_smpl_center'], want_offsets=True, impl=impl)
    vol = out['occ'] if flag is None else engine.scatter_fill(flag, out['occ'], fill)
    vol = vol.reshape(tuple(vol_res))
    v, f, n = engine.extract_mesh(vol, bounds, iso)
    lv, ln = _mesh_to_live(engine, v, n, frame)
    return {'volume': vol, 'offsets': out['off'], 'verts': v, 'faces': f, 'normals': n, 'live_verts': lv, 'live_normals': ln}


def recon_frame(engine: Engine, frame: Dict, img_feat_map, vol_res, flag: Optional[torch.Tensor] = None,
                pts: Optional[torch.Tensor] = None, fill: Optional[torch.Tensor] = None, iso: float = 0.5,
                impl: Optional[str] = None) -> Dict[str, torch.Tensor]:
    """Step 3 of run_avatarcap (main.py:438-453): decoder over the grid, scatter, recon_mesh(iso 0.5), skinning."""
    bounds = np.asarray(frame['cano_bounds'], dtype=np.float32)
    engine.set_image_feature_map(img_feat_map)
    if flag is None:
        ov = engine.eval_recon_grid(bounds, vol_res, frame['cano_smpl_center'], impl=impl)
    else:
        ov = engine.eval_recon(pts, frame['cano_smpl_center'], impl=impl)
    vol = ov if flag is None else engine.scatter_fill(flag, ov, fill)
    vol = vol.reshape(tuple(vol_res))
    v, f, n = engine.extract_mesh(vol, bounds, iso)
    lv, ln = _mesh_to_live(engine, v, n, frame)
    return {'volume': vol, 'verts': v, 'faces': f, 'normals': n, 'live_verts': lv, 'live_normals': ln}


def fused_normal_maps(engine: Engine, avatar: Dict[str, torch.Tensor], frame: Dict, normal_map, cam: Dict, w2c, img: int = 512,
                      integrate_manner: str = 'cover', neck_xy=None, iter_num: int = 100, merge_fn=None):
    """Step 2 of run_avatarcap (main.py:369, 400-428) on the device: the avatar's canonical front / back normal maps
    (render_cano_mesh), the image-observed normals brought to the canonical space (canonicalize_normal_map) and their fusion
    ('cover' :421, or 'merge' :416-420 -- the Adam rotation-grid optimiser is NOT part of this package: pass the reference's own
    normal_fusion.merge_normal_images as `merge_fn(src_img, tar_img, iter_num, neck_xy) -> np.ndarray`).
    `avatar` is avatar_frame()'s result; cam = {'fx','fy','cx','cy'}; w2c the (4,4) world->camera matrix (items['w2c_RT']).
    -> {'front_normal': (1,3,S,S), 'back_normal': (1,3,S,S)} ready for ReconNetwork.get_feat_maps (arch_recon.py:41-52)."""
    from . import render
    center = np.asarray(frame['cano_smpl_center'], np.float32)
    v, f, n = avatar['verts'], avatar['faces'], avatar['normals']
    front_avatar, back_avatar = render.render_cano_mesh_device(engine, v, n, f, center, img)                 # main.py:369
    lbs = engine.lbs_weights(v, frame['cano_smpl_v'], frame['smpl_skinning_weights'])                        # main.py:385
    live_v, vert_mats = engine.skin_points(v, lbs, frame['cano2live_jnt_mats'], return_pt_mats=True)        # main.py:386
    front_img, _, _ = render.canonicalize_normal_map_device(engine, v, live_v, f, normal_map, vert_mats, w2c, cam['fx'], cam['fy'], cam['cx'],
                                                            cam['cy'], center, img)                          # main.py:408-410
    if integrate_manner == 'cover':
        front = render.merge_normal_images_cover(front_avatar.clone(), front_img)
    elif integrate_manner == 'merge':
        if neck_xy is None or merge_fn is None:
            raise ValueError("integrate_manner='merge' needs neck_xy and merge_fn (the reference's normal_fusion.merge_normal_images)")
        front = torch.as_tensor(merge_fn(front_avatar.cpu().numpy(), front_img.cpu().numpy(), iter_num, neck_xy)).to(engine.device, torch.float32)
    else:
        raise ValueError('Invalid integration manner!')                                                      # main.py:423
    # "suppose that the performer is facing the camera": the back keeps the avatar normal (main.py:425-426)
    return {'front_normal': front.permute(2, 0, 1)[None].contiguous(), 'back_normal': back_avatar.permute(2, 0, 1)[None].contiguous(),
            'front_avatar_normal': front_avatar, 'front_image_normal': front_img, 'live_verts': live_v, 'vert_mats': vert_mats}


# ------------------------------------------------------------------------------------------------------------------
# Frame-parallel replicas (BASELINE config[5], SURVEY.md section 8e "Frame-parallel"): frame f runs on rank f mod world,
# no communication on the data path; only the per-frame results are gathered by the caller if it wants them.
def frames_for_rank(n_frames: int, world: int, rank: int):
    """Indices of the frames rank `rank` of `world` processes (round robin, like a DataLoader with a DistributedSampler
    without shuffling)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError('bad rank %d of world %d' % (rank, world))
    return list(range(rank, int(n_frames), world))


def run_frames(engine: Engine, frames, pose_feat_maps, vol_res, impl: Optional[str] = None, iso: float = 0.0):
    """Dense canonical-avatar frames, one after the other on this rank's GPU: per frame the feature map is (re)bound, the
    field evaluated over the whole grid, the mesh extracted and skinned to that frame's live pose. Returns the per-frame
    vertex / face counts (the meshes themselves are dropped frame by frame to bound memory)."""
    counts = []
    for fr, fmap in zip(frames, pose_feat_maps):
        out = avatar_frame(engine, fr, fmap, vol_res, iso=iso, impl=impl)
        counts.append((int(out['verts'].shape[0]), int(out['faces'].shape[0])))
        del out
    return counts
