"""Generate golden vectors by running the REFERENCE's own modules (imported from /root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/gen_golden.py            # writes tests/golden/*.npz

The reference needs three third-party packages that are absent here (pytorch3d, skimage, trimesh) and the
licensed SMPL pkl. They are replaced by import stubs injected into sys.modules -- the reference sources
are not edited or copied:
  * pytorch3d.ops.knn_points / knn_gather : brute-force (squared L2, ascending, int64 idx) stand-ins;
  * dataset.smpl                          : exposes smpl_params.weights / joint_num from the synthetic body;
  * skimage.measure, trimesh              : empty placeholders (marching cubes itself has no reference source here).
The per-frame encoders (UNet / HGFilter) are out of scope: their outputs are replaced by seeded feature maps
assigned to WarpingField.pose_feat_map / returned from ReconNetwork.get_feat_maps.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('AVATARCAP_REFERENCE', '/root/reference')
sys.path.insert(0, ROOT)

from avatarcap_b200 import synth  # noqa: E402


def install_stubs(body: synth.SynthBody, tmpdir: str, weight_volume: np.ndarray) -> None:
    def knn_points(p1, p2, K=1, **kw):
        d = ((p1[:, :, None, :] - p2[:, None, :, :]) ** 2).sum(-1)
        v, i = torch.topk(d, K, dim=2, largest=False, sorted=True)
        return v, i, None

    def knn_gather(x, idx):
        B, N, K = idx.shape
        return torch.stack([x[b][idx[b]] for b in range(B)], 0)

    p3d = types.ModuleType('pytorch3d'); ops = types.ModuleType('pytorch3d.ops')
    ops.knn_points = knn_points; ops.knn_gather = knn_gather; p3d.ops = ops
    tr = types.ModuleType('pytorch3d.transforms'); p3d.transforms = tr
    sys.modules.update({'pytorch3d': p3d, 'pytorch3d.ops': ops, 'pytorch3d.transforms': tr})

    smpl_mod = types.ModuleType('dataset.smpl')
    smpl_mod.smpl_params = types.SimpleNamespace(weights=body.weights, joint_num=synth.N_JOINTS, faces=None)
    smpl_mod.SmplModel = object
    sys.modules['dataset.smpl'] = smpl_mod

    sk = types.ModuleType('skimage'); skm = types.ModuleType('skimage.measure')

    def _no_mc(*a, **k):
        raise NotImplementedError('skimage is not installed; marching cubes has no reference source here')
    skm.marching_cubes = _no_mc; sk.measure = skm
    sys.modules.update({'skimage': sk, 'skimage.measure': skm})
    tm = types.ModuleType('trimesh'); tmp = types.ModuleType('trimesh.proximity'); tm.proximity = tmp
    sys.modules.setdefault('trimesh', tm); sys.modules.setdefault('trimesh.proximity', tmp)
    knn_mod = types.ModuleType('pytorch3d.ops.knn'); knn_mod.knn_points = knn_points; knn_mod.knn_gather = knn_gather
    sys.modules['pytorch3d.ops.knn'] = knn_mod

    sys.path.insert(0, REF)
    import config
    config.device = torch.device('cpu')
    config.cfg = {'model': {'cano_template': {'pos_encoding': 10}, 'warping_field': {'pos_encoding': 0}},
                  'training': {'training_data_dir': tmpdir}}
    np.save(os.path.join(tmpdir, 'cano_base_blend_weight_volume.npy'), weight_volume)


def to_torch_sd(sd):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}


def main() -> None:
    torch.manual_seed(synth.SEED); np.random.seed(synth.SEED)         # main.py:508-509
    torch.set_num_threads(os.cpu_count() or 1)
    body = synth.SynthBody()
    frame = synth.make_frame(body, synth.random_pose(7, 0.4))
    wvol = synth.blend_weight_volume(frame)
    tmpdir = tempfile.mkdtemp(prefix='avc_golden_')
    install_stubs(body, tmpdir, wvol)

    import config
    from network.arch_avatar import GeoTexAvatar, OccupancyNet
    from network.arch_recon import ReconNetwork
    from utils.net_util import get_embedder
    from utils.smpl_util import smpl_util
    from utils import recon_util

    rs = np.random.RandomState(synth.SEED + 17)
    N = 3000
    bmin, bmax = frame['cano_bounds']
    pts = (rs.uniform(0, 1, (N, 3)) * (bmax - bmin) * 1.1 + bmin - 0.05 * (bmax - bmin)).astype(np.float32)
    center = frame['cano_smpl_center']

    # ---------------- avatar network ----------------
    asd = synth.avatar_state_dict()
    net = GeoTexAvatar().eval()
    missing, unexpected = net.load_state_dict(to_torch_sd(asd), strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith('warping_field.unet.') for k in missing), [k for k in missing if 'unet' not in k]
    fmap = synth.feature_map(64, 48, 40, synth.SEED + 2)              # deliberately non-square, small
    net.warping_field.pose_feat_map = torch.from_numpy(fmap)[None]
    batch = {'cano_pts': torch.from_numpy(pts)[None],
             'cano_smpl_center': torch.from_numpy(center)[None],
             'cano_bounds': torch.from_numpy(frame['cano_bounds'])[None],
             'cano2live_jnt_mats': torch.from_numpy(frame['cano2live_jnt_mats'])[None],
             'live_smpl_v': torch.from_numpy(frame['live_smpl_v'])[None]}
    smpl_util.set_cano_smpl_vertices(torch.from_numpy(frame['cano_smpl_v']))
    with torch.no_grad():
        out = OccupancyNet(net).query(batch)
        off = out['nonrigid_offset'][0]
        rgb, alpha, occ = net.cano_template.forward(torch.from_numpy(pts)[None] + off[None])
        emb, emb_dim = get_embedder(10, input_dims=3)
        pe = emb(torch.from_numpy(pts))
        off_only = net.warping_field.query(torch.from_numpy(pts)[None], batch)[0]
    g_avatar = {
        'pts': pts, 'center': center, 'fmap_shape': np.array(fmap.shape), 'fmap_seed': np.array(synth.SEED + 2),
        'fmap_sum': np.array(float(fmap.astype(np.float64).sum())),
        'cano_pts_ov': out['cano_pts_ov'][0].numpy(), 'nonrigid_offset': off.numpy(),
        'warp_query': off_only.numpy(),
        'rgb': rgb[0].numpy(), 'alpha': alpha[0].numpy(), 'occ': occ[0].numpy(), 'pe': pe.numpy(),
    }
    assert emb_dim == 63

    # GeoTexAvatar.forward in the three spaces (N = 1024 points near the live / canonical body)
    M = 1024
    vid = rs.randint(0, synth.N_VERTS, M)
    wl = (frame['live_smpl_v'][vid] + rs.normal(0, 0.04, (M, 3))).astype(np.float32)
    wc = (frame['cano_smpl_v'][vid] + rs.normal(0, 0.04, (M, 3))).astype(np.float32)
    dists = rs.uniform(0.001, 0.02, (M, 1)).astype(np.float32)
    for space, w in (('posed', wl), ('cano', wc), ('temp', wc)):
        wt = torch.from_numpy(w.copy())[None]
        with torch.no_grad():
            o = net.forward(wt, None, torch.from_numpy(dists)[None], batch, pts_space=space)
        g_avatar['fwd_%s_raw' % space] = o['raw'][0].numpy()
        g_avatar['fwd_%s_occ' % space] = o['occ'][0].numpy()
        g_avatar['fwd_%s_off' % space] = o['nonrigid_offset'][0].numpy()
        g_avatar['fwd_%s_wpts_after' % space] = wt[0].numpy()     # 'cano' mutates its input (arch_avatar.py:207,213)
    # NerfRenderer.render as main.py:464-478 drives it (vertex colours): rays through surface-like points along -normal
    from network.arch_avatar import NerfRenderer
    R = 700
    rv = (frame['cano_smpl_v'][rs.randint(0, synth.N_VERTS, R)] + rs.normal(0, 0.01, (R, 3))).astype(np.float32)
    rn = rs.normal(0, 1, (R, 3)).astype(np.float32); rn /= np.linalg.norm(rn, axis=1, keepdims=True)
    items = dict(batch)
    items['ray_o'] = torch.from_numpy(rv + rn)[None]; items['ray_d'] = torch.from_numpy(-rn)[None]
    items['depth'] = torch.ones((1, R)); items['near'] = items['depth'] - 0.05; items['far'] = items['depth'] + 0.05
    items['occupancy'] = items['depth'].clone()
    with torch.no_grad():
        no = NerfRenderer(net).render(items, pts_space='cano', near_dist=0.02, far_dist=0.05)
    np.savez_compressed(os.path.join(HERE, 'nerf_golden.npz'), verts=rv, normals=rn, rgb_map=no['rgb_map'][0].numpy(),
                        acc_map=no['acc_map'][0].numpy(), depth_map=no['depth_map'][0].numpy(), raw=no['raw'][0].numpy(),
                        near_after=items['near'][0].numpy(), far_after=items['far'][0].numpy(), pose_seed=np.array(7))
    g_avatar['fwd_wpts_live'] = wl; g_avatar['fwd_wpts_cano'] = wc; g_avatar['fwd_dists'] = dists
    g_avatar['pose_seed'] = np.array(7)
    np.savez_compressed(os.path.join(HERE, 'avatar_golden.npz'), **g_avatar)

    # ---------------- recon network decoder ----------------
    rsd = synth.recon_state_dict()
    rnet = ReconNetwork()
    missing, unexpected = rnet.load_state_dict(to_torch_sd(rsd), strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith('image_encoder.') for k in missing)
    rmap = synth.feature_map(32, 36, 44, synth.SEED + 3)
    rnet.get_feat_maps = lambda image: [torch.from_numpy(rmap)[None]]      # HGFilter is out of scope
    items = dict(batch)
    items['front_normal'] = torch.zeros(1, 3, 8, 8); items['back_normal'] = torch.zeros(1, 3, 8, 8)
    ov = rnet.infer(items)
    np.savez_compressed(os.path.join(HERE, 'recon_golden.npz'), pts=pts, center=center,
                        fmap_shape=np.array(rmap.shape), fmap_seed=np.array(synth.SEED + 3),
                        fmap_sum=np.array(float(rmap.astype(np.float64).sum())), ov=ov.numpy())

    # ---------------- LBS ----------------
    V = 2000
    vid = rs.randint(0, synth.N_VERTS, V)
    mv = (frame['cano_smpl_v'][vid] + rs.normal(0, 0.03, (V, 3))).astype(np.float32)
    mn = rs.normal(0, 1, (V, 3)).astype(np.float32); mn /= np.linalg.norm(mn, axis=1, keepdims=True)
    with torch.no_grad():
        lbs = smpl_util.calculate_lbs(torch.from_numpy(mv)[None])
        live, mats = smpl_util.skinning(torch.from_numpy(mv)[None], lbs, batch['cano2live_jnt_mats'], True)
        ln = smpl_util.skinning_normal(torch.from_numpy(mn)[None], lbs, batch['cano2live_jnt_mats'])
    np.savez_compressed(os.path.join(HERE, 'lbs_golden.npz'), verts=mv, normals=mn, lbs=lbs[0].numpy(),
                        live=live[0].numpy(), mats=mats[0].numpy(), live_normals=ln[0].numpy(), pose_seed=np.array(7))

    # ---------------- Sobel normals (recon_util.py:9-48) + grid points ----------------
    res = (20, 24, 12)
    vol = rs.normal(0, 1, res).astype(np.float32)
    voxel = ((bmax - bmin) / np.array(res, dtype=np.float32)).astype(np.float32)
    gpts = rs.uniform(-1.05, 1.05, (500, 3)).astype(np.float32)
    with torch.no_grad():
        nv = recon_util.extract_normal_volume(torch.from_numpy(vol), voxel)
        nrm = recon_util.extract_normal_from_volume(torch.from_numpy(vol), voxel, torch.from_numpy(gpts))
    # generate_volume_points is a staticmethod of a class whose module needs trimesh/cv2/scipy.io at import;
    # all present or stubbed above.
    try:
        from dataset.avatarcap_dataset import AvatarCapDataset
        gp = AvatarCapDataset.generate_volume_points(frame['cano_bounds'], (7, 9, 5)).numpy()
        gp2 = AvatarCapDataset.generate_volume_points(frame['cano_bounds'], (64, 33, 128)).numpy()
    except Exception as e:  # pragma: no cover
        raise SystemExit('cannot import dataset.avatarcap_dataset: %r' % (e,))
    np.savez_compressed(os.path.join(HERE, 'mesh_golden.npz'), vol=vol, voxel=voxel, grid_pts=gpts,
                        normal_volume=nv.numpy(), normals=nrm.numpy(), bounds=frame['cano_bounds'],
                        vol_pts_7_9_5=gp, vol_pts_64_33_128_sub=gp2[::97].copy(),
                        vol_pts_64_33_128_sum=np.array(gp2.astype(np.float64).sum(0)))
    print('golden vectors written to', HERE)


if __name__ == '__main__':
    main()
