"""ORACLE (test infrastructure only) -- CPU restatement of AvatarCap's per-point implicit-field path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this file; the product path (avatarcap_b200/) never does.

Every function restates the reference algorithm on the CPU with torch tensors (float32 by default,
float64 on request so the tolerance budget is attributable) and cites the reference file:line it
follows. It consumes the reference's own ``state_dict`` key names, so the same
weights feed the reference modules (tests/golden/gen_golden.py), this oracle and the CUDA packer.

Pinning: the reference ships no tests or golden vectors. This oracle is pinned
against outputs of the reference's own modules executed in the build container
(tests/golden/gen_golden.py -> tests/golden/*.npz; tests/test_oracle_vs_golden.py).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

Tensor = torch.Tensor
SD = Dict[str, np.ndarray]


def _t(a, dtype) -> Tensor:
    if isinstance(a, torch.Tensor):
        return a.to(dtype)
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype)


# ------------------------------------------------------------------------------------------------
# utils/net_util.py
# ------------------------------------------------------------------------------------------------
def embed(x: Tensor, multires: int) -> Tensor:
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)]  (net_util.py:16-37, 40-55).
    multires == 0 -> only the identity term (freq_bands is empty)."""
    outs = [x]
    for k in range(multires):
        freq = torch.tensor(2.0 ** k, dtype=x.dtype)     # 2 ** linspace(0, L-1, L) is exact
        outs.append(torch.sin(x * freq))
        outs.append(torch.cos(x * freq))
    return torch.cat(outs, -1)


# ------------------------------------------------------------------------------------------------
# F.grid_sample restatements (bilinear / trilinear, padding_mode='border', align_corners=True)
# ------------------------------------------------------------------------------------------------
def _unnorm_clip(g: Tensor, size: int) -> Tensor:
    """ATen grid_sampler_compute_source_index: ((g+1)/2)*(size-1), then clip to [0,size-1] (border)."""
    x = ((g + 1) / 2) * (size - 1)
    return torch.clamp(x, 0, size - 1)


def bilinear_border(fmap: Tensor, gx: Tensor, gy: Tensor) -> Tensor:
    """fmap (C,H,W); gx,gy (N,) normalised coords -> (C,N). Call sites: arch_avatar.py:133, arch_recon.py:68."""
    C, H, W = fmap.shape
    ix = _unnorm_clip(gx, W); iy = _unnorm_clip(gy, H)
    x0 = torch.floor(ix); y0 = torch.floor(iy)
    tx = ix - x0; ty = iy - y0
    x0i = x0.long(); y0i = y0.long()
    x1i = x0i + 1; y1i = y0i + 1
    # ATen zeroes the weight of out-of-range taps; with border clipping the +1 tap is only out of range
    # when its weight is exactly 0, so clamping the index is equivalent.
    inx1 = (x1i <= W - 1); iny1 = (y1i <= H - 1)
    x1i = x1i.clamp(max=W - 1); y1i = y1i.clamp(max=H - 1)
    w_nw = (1 - tx) * (1 - ty); w_ne = tx * (1 - ty); w_sw = (1 - tx) * ty; w_se = tx * ty
    w_ne = w_ne * inx1; w_sw = w_sw * iny1; w_se = w_se * (inx1 & iny1)
    flat = fmap.reshape(C, H * W)
    out = (flat[:, y0i * W + x0i] * w_nw + flat[:, y0i * W + x1i] * w_ne +
           flat[:, y1i * W + x0i] * w_sw + flat[:, y1i * W + x1i] * w_se)
    return out


def trilinear_border(vol: Tensor, g: Tensor) -> Tensor:
    """vol (C,D,H,W); g (N,3) normalised (x->W, y->H, z->D) -> (C,N). Call sites: arch_avatar.py:159, recon_util.py:42."""
    C, D, H, W = vol.shape
    ix = _unnorm_clip(g[:, 0], W); iy = _unnorm_clip(g[:, 1], H); iz = _unnorm_clip(g[:, 2], D)
    x0 = torch.floor(ix); y0 = torch.floor(iy); z0 = torch.floor(iz)
    tx = ix - x0; ty = iy - y0; tz = iz - z0
    x0i = x0.long(); y0i = y0.long(); z0i = z0.long()
    flat = vol.reshape(C, -1)
    out = torch.zeros((C, g.shape[0]), dtype=vol.dtype)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                xi = x0i + dx; yi = y0i + dy; zi = z0i + dz
                ok = (xi <= W - 1) & (yi <= H - 1) & (zi <= D - 1)
                w = (tx if dx else 1 - tx) * (ty if dy else 1 - ty) * (tz if dz else 1 - tz) * ok
                xi = xi.clamp(max=W - 1); yi = yi.clamp(max=H - 1); zi = zi.clamp(max=D - 1)
                out = out + flat[:, (zi * H + yi) * W + xi] * w
    return out


# ------------------------------------------------------------------------------------------------
# network/mlp.py
# ------------------------------------------------------------------------------------------------
def _conv1x1(sd: SD, prefix: str, x: Tensor) -> Tensor:
    """nn.Conv1d(k=1) on (C,N): W x + b.  Handles weight-norm keys (mlp.py:24-28: W = g * v / ||v||, norm over dims 1,2)."""
    dt = x.dtype
    if prefix + '.weight_v' in sd:
        v = _t(sd[prefix + '.weight_v'], dt)[:, :, 0]
        g = _t(sd[prefix + '.weight_g'], dt)[:, 0, 0]
        W = v * (g / torch.sqrt((v * v).sum(1)))[:, None]
    else:
        W = _t(sd[prefix + '.weight'], dt)[:, :, 0]
    b = _t(sd[prefix + '.bias'], dt)
    return W @ x + b[:, None]


def _act(name: str, x: Tensor) -> Tensor:
    if name == 'relu':
        return torch.relu(x)
    if name == 'leaky_relu':
        return torch.where(x > 0, x, x * 0.02)           # nn.LeakyReLU(0.02), mlp.py:11
    if name == 'soft_plus':
        return torch.where(x > 20, x, torch.log1p(torch.exp(torch.clamp(x, max=20))))   # nn.Softplus(beta=1,threshold=20)
    raise ValueError(name)


def mlp_forward(sd: SD, prefix: str, x: Tensor, n_layers: int, res_layers: Sequence[int], nlactv: str,
                last_op: Optional[str]) -> Tensor:
    """MLP.forward (mlp.py:56-72). x is (C,N). Hidden layers are Sequential(conv, act) -> keys '<l>.0.*';
    the last layer is a bare conv -> keys '<l>.*'. Skip layers consume cat([x, input]) (activations FIRST)."""
    tmpx = x
    for i in range(n_layers):
        last = (i == n_layers - 1)
        p = '%s.fc_list.%d' % (prefix, i) + ('' if last else '.0')
        inp = torch.cat([x, tmpx], 0) if i in res_layers else x
        x = _conv1x1(sd, p, inp)
        if not last:
            x = _act(nlactv, x)
        elif last_op == 'sigmoid':
            x = torch.sigmoid(x)
        elif last_op == 'tanh':
            x = torch.tanh(x)
    return x


def offset_decoder(sd: SD, prefix: str, x: Tensor) -> Tensor:
    """OffsetDecoder.forward (mlp.py:101-112): 7 x (conv -> BatchNorm1d(eval) -> Softplus), skip cat([x, x4]) into conv5."""
    dt = x.dtype

    def bn(i: int, y: Tensor) -> Tensor:
        p = '%s.bn%d' % (prefix, i)
        mean = _t(sd[p + '.running_mean'], dt)[:, None]; var = _t(sd[p + '.running_var'], dt)[:, None]
        w = _t(sd[p + '.weight'], dt)[:, None]; b = _t(sd[p + '.bias'], dt)[:, None]
        return (y - mean) / torch.sqrt(var + 1e-5) * w + b      # F.batch_norm eval, eps 1e-5

    x0 = x
    h = x
    for i in range(1, 8):
        inp = torch.cat([x0, h], 0) if i == 5 else h          # mlp.py:106 (input FIRST)
        h = _act('soft_plus', bn(i, _conv1x1(sd, '%s.conv%d' % (prefix, i), inp)))
    return h


# ------------------------------------------------------------------------------------------------
# network/arch_avatar.py
# ------------------------------------------------------------------------------------------------
def warp_query(sd: SD, pts: Tensor, feat_map: Tensor, center: Tensor, pos_encoding: int = 0) -> Tensor:
    """WarpingField.query (arch_avatar.py:113-140). pts (N,3), feat_map (64,H,W), center (3,) -> offsets (N,3)."""
    pts_en = embed(pts, pos_encoding).t()                       # :121
    p_ = pts - center[None, :]                                  # :124
    feat = bilinear_border(feat_map, p_[:, 0], -p_[:, 1])       # :125-134
    h = offset_decoder(sd, 'warping_field.mlp', torch.cat([pts_en, feat], 0))   # :136-137
    off = _conv1x1(sd, 'warping_field.out_layer_coord_affine', h)               # :138
    return off.t()


def template_forward(sd: SD, pts: Tensor, if_type: str = 'sdf', pos_encoding: int = 10
                     ) -> Tuple[Tensor, Tensor, Tensor]:
    """DoubleTNet.forward (arch_avatar.py:65-83). pts (N,3) -> rgb (N,3), alpha (N,1), occ (N,1)."""
    e = embed(pts, pos_encoding).t()
    shared = mlp_forward(sd, 'cano_template.shared_mlp', e, 7, [4], 'relu', None)
    geo = mlp_forward(sd, 'cano_template.geo_mlp', shared, 2, [], 'leaky_relu', None)
    clr = mlp_forward(sd, 'cano_template.clr_mlp', shared, 3, [], 'relu', None)
    rgb = torch.sigmoid(clr).t()
    alpha = torch.relu(geo[1:2]).t()
    if if_type == 'occupancy':
        occ = torch.sigmoid(geo[0:1]).t()
    elif if_type == 'sdf':
        occ = geo[0:1].t()
    else:
        raise ValueError('Invalid config.if_type!')
    return rgb, alpha, occ


def occupancy_query(sd: SD, cano_pts: np.ndarray, feat_map: np.ndarray, center: np.ndarray,
                    dtype=torch.float32, chunk: int = 256 * 256 * 4, with_texture: bool = False):
    """OccupancyNet.query (arch_avatar.py:356-381) for B=1: off = warp(p); occ = template(p + off).
    Returns dict of numpy arrays: cano_pts_ov (N,1), nonrigid_offset (N,3) [+ rgb (N,3), alpha (N,1)]."""
    pts = _t(cano_pts, dtype); fm = _t(feat_map, dtype); c = _t(center, dtype)
    offs, occs, rgbs, alphas = [], [], [], []
    with torch.no_grad():
        for i in range(0, pts.shape[0], chunk):
            p = pts[i:i + chunk]
            off = warp_query(sd, p, fm, c)
            rgb, alpha, occ = template_forward(sd, p + off)
            offs.append(off); occs.append(occ); rgbs.append(rgb); alphas.append(alpha)
    out = {'cano_pts_ov': torch.cat(occs, 0).numpy(), 'nonrigid_offset': torch.cat(offs, 0).numpy()}
    if with_texture:
        out['rgb'] = torch.cat(rgbs, 0).numpy(); out['alpha'] = torch.cat(alphas, 0).numpy()
    return out


# ------------------------------------------------------------------------------------------------
# network/arch_recon.py
# ------------------------------------------------------------------------------------------------
def recon_infer(sd: SD, cano_pts: np.ndarray, img_feat_map: np.ndarray, center: np.ndarray,
                dtype=torch.float32, chunk: int = 256 * 256 * 4) -> np.ndarray:
    """ReconNetwork.infer decoder part (arch_recon.py:55-76; the HGFilter encoder at :51-52 is out of scope and its
    output img_feat_map (32,H,W) is an input here). -> (N,) sigmoid occupancy."""
    pts = _t(cano_pts, dtype); fm = _t(img_feat_map, dtype); c = _t(center, dtype)
    outs = []
    with torch.no_grad():
        for i in range(0, pts.shape[0], chunk):
            p_ = pts[i:i + chunk] - c[None, :]                                # :62
            feat = bilinear_border(fm, p_[:, 0], -p_[:, 1])                    # :63-68
            z = p_[:, 2][None, :]                                              # :69
            h0 = torch.cat([feat, z], 0)                                       # :70
            ov = mlp_forward(sd, 'image_decoder', h0, 4, [1, 2], 'leaky_relu', 'sigmoid')   # :71
            outs.append(ov[0])
    return torch.cat(outs, 0).numpy()


# ------------------------------------------------------------------------------------------------
# pytorch3d.ops.knn_points restatement (pytorch3d==0.6.0, un-vendored): squared L2, ascending, int64 idx
# ------------------------------------------------------------------------------------------------
def knn_points(q: Tensor, ref: Tensor, K: int = 1, chunk: int = 8192) -> Tuple[Tensor, Tensor]:
    """q (N,3), ref (M,3) -> (dists2 (N,K), idx (N,K)). Distances are sum((q-r)^2) evaluated directly (not the
    |q|^2+|r|^2-2qr expansion) as pytorch3d's kernels do."""
    ds, ids = [], []
    for i in range(0, q.shape[0], chunk):
        d = ((q[i:i + chunk, None, :] - ref[None, :, :]) ** 2).sum(-1)
        v, ix = torch.topk(d, K, dim=1, largest=False, sorted=True)
        ds.append(v); ids.append(ix)
    return torch.cat(ds, 0), torch.cat(ids, 0)


# ------------------------------------------------------------------------------------------------
# utils/smpl_util.py
# ------------------------------------------------------------------------------------------------
def calculate_lbs(points: np.ndarray, cano_smpl_v: np.ndarray, skin_w: np.ndarray, dtype=torch.float32) -> np.ndarray:
    """SmplUtil.calculate_lbs (smpl_util.py:24-39): KNN-4, w = exp(-d2/(2 r^2)), r = 0.05, normalise (+1e-16), blend."""
    p = _t(points, dtype); v = _t(cano_smpl_v, dtype); sw = _t(skin_w, dtype)
    d2, idx = knn_points(p, v, 4)
    w = torch.exp(-d2 / (2 * 0.05 * 0.05))
    w = w / (w.sum(-1, keepdim=True) + 1e-16)
    lbs = (sw[idx] * w[..., None]).sum(-2)
    return lbs.numpy()


def skinning(points: np.ndarray, lbs: np.ndarray, jnt_mats: np.ndarray, dtype=torch.float32):
    """SmplUtil.skinning (smpl_util.py:58-74) -> (live_pts (N,3), pt_mats (N,4,4))."""
    p = _t(points, dtype); l = _t(lbs, dtype); J = _t(jnt_mats, dtype)
    M = torch.einsum('nj,jxy->nxy', l, J)
    out = torch.einsum('nxy,ny->nx', M[:, :3, :3], p) + M[:, :3, 3]
    return out.numpy(), M.numpy()


def skinning_normal(normals: np.ndarray, lbs: np.ndarray, jnt_mats: np.ndarray, dtype=torch.float32) -> np.ndarray:
    """SmplUtil.skinning_normal (smpl_util.py:76-81): rotation block only, no inverse-transpose, no renormalise."""
    n = _t(normals, dtype); l = _t(lbs, dtype); J = _t(jnt_mats, dtype)
    M = torch.einsum('nj,jxy->nxy', l, J)
    return torch.einsum('nxy,ny->nx', M[:, :3, :3], n).numpy()


def cano_blend_weights(volume_xyzc: np.ndarray, pts01: Tensor) -> Tensor:
    """CanoBlendWeightVolume.forward (arch_avatar.py:143-165). volume (X,Y,Z,24) as stored on disk; pts in [0,1]^3.
    The reference permutes to (24,X,Y,Z) and samples with grid[:, [2,1,0]]: grid x <-> Z axis (W), y <-> Y (H), z <-> X (D)."""
    vol = _t(volume_xyzc, pts01.dtype).permute(3, 0, 1, 2)
    g = (2 * pts01 - 1)[:, [2, 1, 0]]
    return trilinear_border(vol, g).t()


def geotex_forward(sd: SD, wpts: np.ndarray, dists: np.ndarray, frame: Dict[str, np.ndarray], feat_map: np.ndarray,
                   weight_volume: np.ndarray, pts_space: str = 'posed', dtype=torch.float32):
    """GeoTexAvatar.forward (arch_avatar.py:178-237), B=1. wpts (N,3), dists (N,1).
    Returns dict raw (N,4), occ (N,1), nonrigid_offset (N,3), cano_pts (N,3) [the reference mutates wpts in 'cano' mode]."""
    assert pts_space in ('posed', 'cano', 'temp')
    w = _t(wpts, dtype); dd = _t(dists, dtype)
    bounds = _t(frame['cano_bounds'], dtype); center = _t(frame['cano_smpl_center'], dtype)
    skin_w = _t(frame['smpl_skinning_weights'], dtype)
    with torch.no_grad():
        if pts_space == 'posed':
            d2, idx = knn_points(w, _t(frame['live_smpl_v'], dtype), 1)                  # :190
            near = d2[:, 0] < 0.08 * 0.08
            pw = skin_w[idx[:, 0]]                                                        # :197-198
            live2cano = torch.linalg.inv(_t(frame['cano2live_jnt_mats'], dtype))          # :199
            M = torch.einsum('nj,jxy->nxy', pw, live2cano)
            cp = torch.einsum('nxy,ny->nx', M[:, :3, :3], w) + M[:, :3, 3]                # :200
            cp = (cp - bounds[0][None]) / (bounds[1] - bounds[0])[None]                   # :201-203
            pw2 = cano_blend_weights(weight_volume, cp)                                   # :204
            M = torch.einsum('nj,jxy->nxy', pw2, live2cano)
            cano = torch.einsum('nxy,ny->nx', M[:, :3, :3], w) + M[:, :3, 3]              # :205
        else:
            cano = w.clone()
            d2, idx = knn_points(w, _t(frame['cano_smpl_v'], dtype), 1)                   # :208
            near = d2[:, 0] < 0.08 * 0.08
        if pts_space in ('posed', 'cano'):
            off = warp_query(sd, cano, _t(feat_map, dtype), center)                       # :212
            cano = cano + off
        else:
            off = torch.zeros_like(cano)
        rgb, alpha, occ = template_forward(sd, cano)                                      # :218
        inside = (cano > bounds[0][None]) & (cano < bounds[1][None])                      # :221-222
        outside = inside.sum(1) != 3
        alpha = alpha.clone()
        alpha[outside] = 0; alpha[~near] = 0                                              # :224-225
        alpha = 1.0 - torch.exp(-alpha * dd)                                              # :227-229
        raw = torch.cat([rgb, alpha], -1)
    return {'raw': raw.numpy(), 'occ': occ.numpy(), 'nonrigid_offset': off.numpy(), 'cano_pts': cano.numpy()}


# ------------------------------------------------------------------------------------------------
# dataset/avatarcap_dataset.py grid + validity
# ------------------------------------------------------------------------------------------------
def generate_volume_points(bounds: np.ndarray, res) -> np.ndarray:
    """generate_volume_points (avatarcap_dataset.py:312-326) -- runs the same torch ops the reference runs."""
    xs = torch.linspace(0, 1, steps=res[0], dtype=torch.float32)
    ys = torch.linspace(0, 1, steps=res[1], dtype=torch.float32)
    zs = torch.linspace(0, 1, steps=res[2], dtype=torch.float32)
    xv, yv, zv = torch.meshgrid(xs, ys, zs, indexing='ij')
    pts = torch.cat([xv.reshape(-1, 1), yv.reshape(-1, 1), zv.reshape(-1, 1)], -1)
    b = torch.from_numpy(np.asarray(bounds, dtype=np.float32))
    return (pts * (b[1] - b[0]) + b[0]).numpy()


def valid_points_flag(vol_pts: np.ndarray, cano_smpl_v: np.ndarray, thres: float = 0.1) -> np.ndarray:
    """avatarcap_dataset.py:114-116: KNN-1 squared distance < 0.1**2."""
    d2, _ = knn_points(_t(vol_pts, torch.float32), _t(cano_smpl_v, torch.float32), 1)
    return (d2[:, 0] < thres ** 2).numpy()


def scatter_fill(flag: np.ndarray, vals: np.ndarray, fill: np.ndarray, res) -> np.ndarray:
    """main.py:357,362-364: vol = zeros; vol[flag] = vals; vol[~flag] = fill; reshape(vol_res)."""
    vol = np.zeros(int(np.prod(res)), dtype=np.float32)
    vol[flag] = vals
    vol[~flag] = fill
    return vol.reshape(res)


# ------------------------------------------------------------------------------------------------
# NerfRenderer + raw2outputs (vertex-colour driver, main.py:464-478)
# ------------------------------------------------------------------------------------------------
def raw2outputs(raw: Tensor, z_vals: Tensor, white_bkgd: bool = False):
    """utils/nerf_util.py:185-212 -> rgb_map, acc_map, depth_map."""
    rgb = raw[..., :-1]; alpha = raw[..., -1]
    weights = alpha * torch.cumprod(torch.cat([torch.ones((alpha.shape[0], 1), dtype=alpha.dtype), 1. - alpha + 1e-10], -1), -1)[:, :-1]
    rgb_map = torch.sum(weights[..., None] * rgb, -2)
    depth_map = torch.sum(weights * z_vals, -1)
    acc_map = torch.sum(weights, -1)
    if white_bkgd:
        rgb_map = rgb_map + (1. - acc_map[..., None])
    return rgb_map, acc_map, depth_map


def nerf_render(sd: SD, ray_o: np.ndarray, ray_d: np.ndarray, near: np.ndarray, far: np.ndarray, depth: np.ndarray,
                frame: Dict[str, np.ndarray], feat_map: np.ndarray, weight_volume: np.ndarray, pts_space: str = 'cano',
                near_dist: float = 0.05, far_dist: float = 0.05, n_samples: int = 64, dtype=torch.float32):
    """NerfRenderer.render / get_pixel_value / get_wsampling_points / get_density_color (arch_avatar.py:244-349), eval mode, B=1."""
    o = _t(ray_o, dtype); d = _t(ray_d, dtype); nr = _t(near, dtype).clone(); fr = _t(far, dtype).clone(); dp = _t(depth, dtype)
    valid = dp > 1e-6                                                                   # :285-287
    nr[valid] = dp[valid] - near_dist; fr[valid] = dp[valid] + far_dist
    t_vals = torch.linspace(0., 1., steps=n_samples).to(nr)                             # :249
    z_vals = nr[..., None] * (1. - t_vals) + fr[..., None] * t_vals                     # :250
    pts = o[:, None] + d[:, None] * z_vals[..., None]                                   # :262
    dists = z_vals[..., 1:] - z_vals[..., :-1]                                          # :276-277
    dists = torch.cat([dists, dists[..., -1:]], dim=1)
    ret = geotex_forward(sd, pts.reshape(-1, 3).numpy(), dists.reshape(-1, 1).numpy(), frame, feat_map, weight_volume, pts_space, dtype)
    raw = _t(ret['raw'], dtype).reshape(-1, n_samples, 4)
    rgb_map, acc_map, depth_map = raw2outputs(raw, z_vals)
    return {'rgb_map': rgb_map.numpy(), 'acc_map': acc_map.numpy(), 'depth_map': depth_map.numpy(), 'raw': raw.reshape(-1, 4).numpy(),
            'near': nr.numpy(), 'far': fr.numpy()}
