"""Deterministic synthetic inputs for the hot path.

The licensed SMPL pkl, the trained checkpoints and the example dataset are not
available offline, so every test / bench input is generated here from the
reference's seed 31359 (main.py:508-509):

* a 24-joint skeleton with the SMPL kinematic topology and 6 890 surface points on
  capsules around the bones, with 4-bone Gaussian skin weights (rows sum to 1) --
  the stand-in for ``dataset/smpl.py:SmplParams/SmplModel``;
* ``state_dict``-shaped weight dictionaries with exactly the reference's key names
  and shapes: PyTorch-default U(-1/sqrt(fan_in), +) init, the two
  ~0-initialised heads (network/arch_avatar.py:60,105) re-initialised to a
  non-degenerate scale, BatchNorm running stats randomised;
* seeded feature maps standing in for the per-frame encoder outputs
  (UnetNoCond7DS / HGFilter are out of scope).

Nothing here touches ``oracle/`` or the CUDA library: it only makes inputs.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np

SEED = 31359
N_VERTS = 6890
N_JOINTS = 24

# SMPL kinematic tree (parent of joint j), dataset/smpl.py:33-34 loads the same table.
PARENTS = np.array([-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14,
                    16, 17, 18, 19, 20, 21], dtype=np.int32)

# Approximate SMPL rest-pose joints (metres; y up, z forward).
_REST_JOINTS = np.array([
    [0.00, 0.00, 0.00],     # 0 pelvis
    [0.07, -0.09, 0.00],    # 1 l_hip
    [-0.07, -0.09, 0.00],   # 2 r_hip
    [0.00, 0.11, -0.01],    # 3 spine1
    [0.10, -0.47, 0.00],    # 4 l_knee
    [-0.10, -0.47, 0.00],   # 5 r_knee
    [0.00, 0.24, 0.00],     # 6 spine2
    [0.09, -0.87, -0.03],   # 7 l_ankle
    [-0.09, -0.87, -0.03],  # 8 r_ankle
    [0.00, 0.30, 0.01],     # 9 spine3
    [0.11, -0.93, 0.09],    # 10 l_foot
    [-0.11, -0.93, 0.09],   # 11 r_foot
    [0.00, 0.51, -0.02],    # 12 neck
    [0.08, 0.41, -0.01],    # 13 l_collar
    [-0.08, 0.41, -0.01],   # 14 r_collar
    [0.00, 0.60, 0.02],     # 15 head
    [0.18, 0.44, -0.02],    # 16 l_shoulder
    [-0.18, 0.44, -0.02],   # 17 r_shoulder
    [0.44, 0.44, -0.03],    # 18 l_elbow
    [-0.44, 0.44, -0.03],   # 19 r_elbow
    [0.69, 0.44, -0.03],    # 20 l_wrist
    [-0.69, 0.44, -0.03],   # 21 r_wrist
    [0.78, 0.43, -0.03],    # 22 l_hand
    [-0.78, 0.43, -0.03],   # 23 r_hand
], dtype=np.float64)

# capsule radius of the bone that ENDS at joint j (parent -> j)
_BONE_RADIUS = np.array([0.12, 0.10, 0.10, 0.12, 0.075, 0.075, 0.13, 0.055, 0.055, 0.13,
                         0.04, 0.04, 0.07, 0.09, 0.09, 0.09, 0.06, 0.06, 0.05, 0.05,
                         0.04, 0.04, 0.035, 0.035], dtype=np.float64)


def rodrigues(theta: np.ndarray) -> np.ndarray:
    """Axis-angle -> 3x3 rotation (what cv.Rodrigues does at dataset/smpl.py:84)."""
    theta = np.asarray(theta, dtype=np.float64).reshape(3)
    ang = float(np.linalg.norm(theta))
    if ang < 1e-12:
        return np.eye(3)
    k = theta / ang
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + math.sin(ang) * K + (1 - math.cos(ang)) * (K @ K)


def joint_affine_mats(pose75: np.ndarray, joints: np.ndarray = _REST_JOINTS) -> np.ndarray:
    """Forward kinematics exactly as dataset/smpl.py:79-101 (local = [R | (I-R) J])."""
    pose75 = np.asarray(pose75, dtype=np.float64).reshape(75)
    local = []
    for j in range(N_JOINTS):
        r = rodrigues(pose75[3 + 3 * j: 6 + 3 * j])
        m = np.eye(4)
        m[:3, :3] = r
        m[:3, 3] = pose75[0:3] if j == 0 else (np.eye(3) - r) @ joints[j]
        local.append(m)
    out = [local[0]]
    for j in range(1, N_JOINTS):
        out.append(out[PARENTS[j]] @ local[j])
    return np.stack(out, 0)


def cano_pose() -> np.ndarray:
    """Canonical pose: hips +-25 deg about z (utils/smpl_util.py:16-18)."""
    p = np.zeros(75, dtype=np.float64)
    p[3 + 3 * 1 + 2] = math.radians(25)
    p[3 + 3 * 2 + 2] = math.radians(-25)
    return p


def _point_segment_dist(p: np.ndarray, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    ab = b - a
    t = np.clip(((p - a) @ ab) / max(float(ab @ ab), 1e-12), 0.0, 1.0)
    return np.linalg.norm(p - (a + t[:, None] * ab), axis=1)


class SynthBody:
    """Stand-in for SmplParams + SmplModel: rest verts, skin weights, FK, posing."""

    def __init__(self, seed: int = SEED):
        rs = np.random.RandomState(seed)
        bones = [(int(PARENTS[j]), j) for j in range(1, N_JOINTS)]
        lens = np.array([np.linalg.norm(_REST_JOINTS[b] - _REST_JOINTS[a]) for a, b in bones])
        area = lens * _BONE_RADIUS[1:] + 2 * _BONE_RADIUS[1:] ** 2
        counts = np.floor(area / area.sum() * N_VERTS).astype(int)
        counts[0] += N_VERTS - counts.sum()
        verts = []
        for (a, b), n in zip(bones, counts):
            pa, pb = _REST_JOINTS[a], _REST_JOINTS[b]
            axis = pb - pa
            L = np.linalg.norm(axis)
            axis = axis / L
            tmp = np.array([1.0, 0, 0]) if abs(axis[0]) < 0.9 else np.array([0, 1.0, 0])
            u = np.cross(axis, tmp); u /= np.linalg.norm(u)
            v = np.cross(axis, u)
            t = rs.uniform(-0.15, 1.15, n)  # slight overshoot rounds the caps
            phi = rs.uniform(0, 2 * np.pi, n)
            r = _BONE_RADIUS[b] * np.sqrt(np.clip(1 - np.clip(np.abs(t - 0.5) - 0.5, 0, None) ** 2 / 0.0225, 0.05, 1))
            verts.append(pa + np.outer(np.clip(t, -0.15, 1.15) * L, axis) +
                         (r * np.cos(phi))[:, None] * u + (r * np.sin(phi))[:, None] * v)
        self.rest_verts = np.concatenate(verts, 0)
        assert self.rest_verts.shape == (N_VERTS, 3)
        # skin weights: Gaussian of the distance to each joint's outgoing bones, top-4
        d = np.full((N_VERTS, N_JOINTS), 1e9)
        for a, b in bones:
            dist = _point_segment_dist(self.rest_verts, _REST_JOINTS[a], _REST_JOINTS[b])
            d[:, a] = np.minimum(d[:, a], dist)
        leaf = [j for j in range(N_JOINTS) if j not in PARENTS]
        for j in leaf:
            d[:, j] = np.linalg.norm(self.rest_verts - _REST_JOINTS[j], axis=1)
        w = np.exp(-d ** 2 / (2 * 0.06 ** 2))
        kth = np.sort(w, axis=1)[:, -4][:, None]
        w = np.where(w >= kth, w, 0.0)
        w /= w.sum(1, keepdims=True)
        self.weights = w.astype(np.float32)          # smpl_params.weights (6890, 24)
        self.joints = _REST_JOINTS.astype(np.float32)

    def posed(self, pose75: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """-> (posed_vertices (6890,3) f32, jnt_affine_mats (24,4,4) f64), dataset/smpl.py:79-113."""
        mats = joint_affine_mats(pose75)
        vm = np.einsum('vj,jab->vab', self.weights.astype(np.float64), mats)
        pv = np.einsum('vab,vb->va', vm[:, :3, :3], self.rest_verts) + vm[:, :3, 3]
        return pv.astype(np.float32), mats


def body_inside(points: np.ndarray, pose75: np.ndarray) -> np.ndarray:
    """Analytic inside test of the capsule body in pose `pose75` -> bool (N,). Stand-in for the reference's
    `trimesh.contains` on the SMPL mesh (avatarcap_dataset.py:121-123), which needs the licensed SMPL faces."""
    mats = joint_affine_mats(pose75)
    pj = np.einsum('jab,jb->ja', mats[:, :3, :3], _REST_JOINTS) + mats[:, :3, 3]
    p = np.asarray(points, dtype=np.float64)
    inside = np.zeros(len(p), dtype=bool)
    for j in range(1, N_JOINTS):
        inside |= _point_segment_dist(p, pj[PARENTS[j]], pj[j]) < _BONE_RADIUS[j]
    return inside


def body_sdf(points: np.ndarray, pose75: np.ndarray) -> np.ndarray:
    """Smooth signed field of the capsule body (positive inside, the reference's SDF convention, avatarcap_dataset.py:124):
    max over bones of (radius - distance). Gives closed test meshes for the rasteriser / fusion stages."""
    mats = joint_affine_mats(pose75)
    pj = np.einsum('jab,jb->ja', mats[:, :3, :3], _REST_JOINTS) + mats[:, :3, 3]
    p = np.asarray(points, dtype=np.float64)
    f = np.full(len(p), -1e9)
    for j in range(1, N_JOINTS):
        f = np.maximum(f, _BONE_RADIUS[j] - _point_segment_dist(p, pj[PARENTS[j]], pj[j]))
    return f.astype(np.float32)


def random_pose(seed: int, max_abs: float = 0.5) -> np.ndarray:
    rs = np.random.RandomState(seed)
    p = np.zeros(75)
    p[6:] = rs.uniform(-max_abs, max_abs, 69)
    p[3 + 22 * 3:] = 0.0  # hands zeroed as avatarcap_dataset.py:197-198 does
    return p


def make_frame(body: SynthBody, live_pose75: np.ndarray | None = None) -> Dict[str, np.ndarray]:
    """Per-frame geometry dict with the keys of avatarcap_dataset.py:253-264 (numpy, no batch dim)."""
    cano_v, cano_mats = body.posed(cano_pose())
    live_pose75 = np.zeros(75) if live_pose75 is None else live_pose75
    live_v, live_mats = body.posed(live_pose75)
    mn, mx = cano_v.min(0), cano_v.max(0)
    center = (0.5 * (mn + mx)).astype(np.float32)           # avatarcap_dataset.py:65
    bmin, bmax = mn.copy(), mx.copy()
    bmin[:2] -= 0.05; bmax[:2] += 0.05; bmin[2] -= 0.15; bmax[2] += 0.15   # :90-97
    cano2live = (live_mats @ np.linalg.inv(cano_mats)).astype(np.float32)  # :198
    return {
        'cano_smpl_v': cano_v,
        'live_smpl_v': live_v,
        'cano_smpl_center': center,
        'cano_bounds': np.stack([bmin, bmax], 0).astype(np.float32),
        'cano2live_jnt_mats': cano2live,
        'smpl_skinning_weights': body.weights,
    }


def blend_weight_volume(frame: Dict[str, np.ndarray], voxel: float = 0.025) -> np.ndarray:
    """(X,Y,Z,24) nearest-vertex skin weights on a 2.5 cm grid over cano_bounds
    (stand-in for cano_base_blend_weight_volume.npy, gen_data/preprocess_training_data.py:426-460)."""
    from scipy.spatial import cKDTree
    bmin, bmax = frame['cano_bounds']
    dims = np.maximum(np.round((bmax - bmin) / voxel).astype(int), 2)
    axes = [np.linspace(bmin[i], bmax[i], dims[i]) for i in range(3)]
    g = np.stack(np.meshgrid(*axes, indexing='ij'), -1).reshape(-1, 3)
    _, idx = cKDTree(frame['cano_smpl_v']).query(g, k=1)
    return frame['smpl_skinning_weights'][idx].reshape(dims[0], dims[1], dims[2], N_JOINTS).astype(np.float32)


# ----------------------------------------------------------------------------------------
# weights (state_dict key names = SURVEY.md appendix A)
# ----------------------------------------------------------------------------------------

def _conv(rs: np.random.RandomState, cout: int, cin: int, sd: Dict[str, np.ndarray], prefix: str,
          wkey: str = 'weight') -> None:
    """nn.Conv1d default init: kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for W and b."""
    bound = 1.0 / math.sqrt(cin)
    sd[prefix + '.' + wkey] = rs.uniform(-bound, bound, (cout, cin, 1)).astype(np.float32)
    sd[prefix + '.bias'] = rs.uniform(-bound, bound, (cout,)).astype(np.float32)


def avatar_state_dict(seed: int = SEED, occ_scale: float = 1.0, off_scale: float = 0.02, density_scale: float = 300.0) -> Dict[str, np.ndarray]:
    """Per-point keys of GeoTexAvatar.state_dict() (the UNet keys are out of scope and omitted).

    occ_scale / off_scale set the magnitude of the two heads the reference initialises to ~0
    (arch_avatar.py:17-23,60,105): occ ~ O(occ_scale) like a trained clip(sdf,+-0.1)/0.1 target
    (main.py:104), offsets ~ O(off_scale) metres.
    """
    rs = np.random.RandomState(seed)
    sd: Dict[str, np.ndarray] = {}
    # cano_template.shared_mlp: 63 -> 256 x6 (skip at 4: 256+63) -> 256 linear   (arch_avatar.py:38-44)
    chans = [63, 256, 256, 256, 256, 256, 256]
    for l in range(6):
        cin = chans[l] + (63 if l == 4 else 0)
        _conv(rs, chans[l + 1], cin, sd, 'cano_template.shared_mlp.fc_list.%d.0' % l)
    _conv(rs, 256, 256, sd, 'cano_template.shared_mlp.fc_list.6')
    # geo_mlp 256 -> 128 -> 2 (arch_avatar.py:46-51)
    _conv(rs, 128, 256, sd, 'cano_template.geo_mlp.fc_list.0.0')
    _conv(rs, 2, 128, sd, 'cano_template.geo_mlp.fc_list.1')
    # clr_mlp 256 -> 256 -> 128 -> 3 (arch_avatar.py:53-58)
    _conv(rs, 256, 256, sd, 'cano_template.clr_mlp.fc_list.0.0')
    _conv(rs, 128, 256, sd, 'cano_template.clr_mlp.fc_list.1.0')
    _conv(rs, 3, 128, sd, 'cano_template.clr_mlp.fc_list.2')
    # warping_field.mlp = OffsetDecoder(67) (mlp.py:79-99)
    for i in range(1, 8):
        cin = 67 if i == 1 else (256 + 67 if i == 5 else 256)
        _conv(rs, 256, cin, sd, 'warping_field.mlp.conv%d' % i)
        p = 'warping_field.mlp.bn%d' % i
        sd[p + '.weight'] = rs.uniform(0.6, 1.4, 256).astype(np.float32)
        sd[p + '.bias'] = rs.uniform(-0.3, 0.3, 256).astype(np.float32)
        sd[p + '.running_mean'] = rs.uniform(-0.2, 0.2, 256).astype(np.float32)
        sd[p + '.running_var'] = rs.uniform(0.3, 1.5, 256).astype(np.float32)
        sd[p + '.num_batches_tracked'] = np.array(100, dtype=np.int64)
    _conv(rs, 3, 256, sd, 'warping_field.out_layer_coord_affine')
    # re-initialise the two ~0 heads to a non-degenerate scale (SURVEY.md section 8c hygiene)
    k = 'cano_template.geo_mlp.fc_list.1'
    sd[k + '.weight'] = (rs.uniform(-1, 1, (2, 128, 1)) * occ_scale * 6.0).astype(np.float32)
    sd[k + '.bias'] = (rs.uniform(-0.1, 0.1, 2) * occ_scale).astype(np.float32)
    # row 1 is the NeRF density: a trained field reaches sigma*delta ~ 1 over a ~1 mm sample spacing (arch_avatar.py:227-229)
    sd[k + '.weight'][1] *= density_scale; sd[k + '.bias'][1] *= density_scale
    k = 'warping_field.out_layer_coord_affine'
    sd[k + '.weight'] = (rs.uniform(-1, 1, (3, 256, 1)) * off_scale * 0.12).astype(np.float32)
    sd[k + '.bias'] = (rs.uniform(-0.2, 0.2, 3) * off_scale).astype(np.float32)
    return sd


def recon_state_dict(seed: int = SEED + 1) -> Dict[str, np.ndarray]:
    """image_decoder keys of ReconNetwork.state_dict() (weight-norm: weight_g, weight_v; arch_recon.py:18-39)."""
    rs = np.random.RandomState(seed)
    sd: Dict[str, np.ndarray] = {}
    spec = [(512, 33), (256, 512 + 33), (128, 256 + 33)]
    for l, (cout, cin) in enumerate(spec):
        p = 'image_decoder.fc_list.%d.0' % l
        _conv(rs, cout, cin, sd, p, wkey='weight_v')
        v = sd[p + '.weight_v']
        nrm = np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True))
        sd[p + '.weight_g'] = (nrm * rs.uniform(0.7, 1.6, (cout, 1, 1))).astype(np.float32)
    _conv(rs, 1, 128, sd, 'image_decoder.fc_list.3')
    sd['image_decoder.fc_list.3.weight'] = (sd['image_decoder.fc_list.3.weight'] * 12.0).astype(np.float32)
    return sd


def feature_map(channels: int, height: int, width: int, seed: int, scale: float = 0.5) -> np.ndarray:
    """(C,H,W) f32 smooth-ish seeded field standing in for an encoder output."""
    rs = np.random.RandomState(seed)
    lo_h, lo_w = max(height // 8, 2), max(width // 8, 2)
    lo = rs.normal(0, 1, (channels, lo_h, lo_w))
    yi = np.linspace(0, lo_h - 1, height); xi = np.linspace(0, lo_w - 1, width)
    y0 = np.floor(yi).astype(int).clip(0, lo_h - 2); x0 = np.floor(xi).astype(int).clip(0, lo_w - 2)
    fy = (yi - y0)[None, :, None]; fx = (xi - x0)[None, None, :]
    a = lo[:, y0][:, :, x0]; b = lo[:, y0][:, :, x0 + 1]
    c = lo[:, y0 + 1][:, :, x0]; d = lo[:, y0 + 1][:, :, x0 + 1]
    smooth = (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy
    out = scale * (smooth + 0.35 * rs.normal(0, 1, (channels, height, width)))
    return out.astype(np.float32)


def volume_points(bounds: np.ndarray, res) -> np.ndarray:
    """Restates generate_volume_points (avatarcap_dataset.py:312-326) in numpy f32: inclusive linspace per axis,
    z fastest, then pts * len + bmin. torch.linspace(0,1,n) in f32 computes i*step for the lower half and
    1 - (n-1-i)*step for the upper half; reproduced here so the points are bit-identical."""
    axes = []
    for n in res:
        step = np.float32(1.0) / np.float32(n - 1) if n > 1 else np.float32(0)
        i = np.arange(n)
        lo = (i.astype(np.float32) * step).astype(np.float32)
        hi = (np.float32(1.0) - (np.float32(n - 1) - i.astype(np.float32)) * step).astype(np.float32)
        axes.append(np.where(i < n // 2, lo, hi).astype(np.float32))
    xv, yv, zv = np.meshgrid(axes[0], axes[1], axes[2], indexing='ij')
    pts = np.stack([xv.reshape(-1), yv.reshape(-1), zv.reshape(-1)], -1).astype(np.float32)
    ln = (bounds[1] - bounds[0]).astype(np.float32)
    return (pts * ln + bounds[0].astype(np.float32)).astype(np.float32)


# ----------------------------------------------------------------------------------------
# per-frame encoders: seeded weights with the reference's key names
# ----------------------------------------------------------------------------------------

def _conv2d(rs, sd, prefix, cout, cin, k, bias=True, wshape=None, gain=1.0):
    """He-style normal init (keeps activations O(1) through 15-25 conv layers); wshape overrides for ConvTranspose2d (in,out,k,k)."""
    std = gain * math.sqrt(2.0 / (cin * k * k))
    sd[prefix + '.weight'] = rs.normal(0, std, wshape or (cout, cin, k, k)).astype(np.float32)
    if bias:
        sd[prefix + '.bias'] = rs.uniform(-0.1, 0.1, (cout,)).astype(np.float32)


def _bn_stats(rs, sd, prefix, c):
    """BatchNorm2d(affine=False) buffers (unets.py:17,46): non-trivial running stats so that the folding is exercised."""
    sd[prefix + '.running_mean'] = rs.uniform(-0.2, 0.2, c).astype(np.float32)
    sd[prefix + '.running_var'] = rs.uniform(0.5, 1.5, c).astype(np.float32)
    sd[prefix + '.num_batches_tracked'] = np.array(100, dtype=np.int64)


def unet_state_dict(seed: int = SEED + 10, prefix: str = '') -> Dict[str, np.ndarray]:
    """UnetNoCond7DS(input_nc=6, output_nc=64, nf=32, up_mode='upconv') keys (unets.py:169-199, arch_avatar.py:95), including the
    never-executed `upconv4` (checkpoint compatibility; forward calls upconv3 twice, unets.py:214-215)."""
    rs = np.random.RandomState(seed)
    sd: Dict[str, np.ndarray] = {}
    nf = 32
    down = [(6, nf), (nf, 2 * nf), (2 * nf, 4 * nf), (4 * nf, 8 * nf), (8 * nf, 8 * nf), (8 * nf, 8 * nf), (8 * nf, 8 * nf)]
    for i, (cin, cout) in enumerate(down, 1):
        _conv2d(rs, sd, prefix + 'conv%d.conv' % i, cout, cin, 4, bias=False)
        if 2 <= i <= 6:
            _bn_stats(rs, sd, prefix + 'conv%d.bn' % i, cout)
    for i, (cin, cout) in enumerate([(8 * nf, 8 * nf), (16 * nf, 8 * nf), (16 * nf, 8 * nf), (16 * nf, 4 * nf)], 1):
        # ConvTranspose2d weight is (in, out, k, k); stride 2 / kernel 4 -> every output sees cin*4 taps
        _conv2d(rs, sd, prefix + 'upconv%d.up' % i, cout, cin, 2, bias=False, wshape=(cin, cout, 4, 4))
        _bn_stats(rs, sd, prefix + 'upconv%d.bn' % i, cout)
    for name, cin, cout, bn in (('C5', 12 * nf, 2 * nf, True), ('C6', 4 * nf, nf, True), ('C7', 2 * nf, 64, False)):
        _conv2d(rs, sd, prefix + 'upconv%s.up.1' % name, cout, cin, 3, bias=True)
        if bn:
            _bn_stats(rs, sd, prefix + 'upconv%s.bn' % name, cout)
    return sd


def _gn(rs, sd, prefix, c):
    sd[prefix + '.weight'] = rs.uniform(0.7, 1.3, c).astype(np.float32)
    sd[prefix + '.bias'] = rs.uniform(-0.2, 0.2, c).astype(np.float32)


def _convblock(rs, sd, prefix, cin, cout):
    """HGFilters.ConvBlock with GroupNorm (HGFilters.py:33-75)."""
    _conv2d(rs, sd, prefix + '.conv1', cout // 2, cin, 3, bias=False)
    _conv2d(rs, sd, prefix + '.conv2', cout // 4, cout // 2, 3, bias=False)
    _conv2d(rs, sd, prefix + '.conv3', cout // 4, cout // 4, 3, bias=False)
    _gn(rs, sd, prefix + '.bn1', cin); _gn(rs, sd, prefix + '.bn2', cout // 2); _gn(rs, sd, prefix + '.bn3', cout // 4); _gn(rs, sd, prefix + '.bn4', cin)
    if cin != cout:
        # downsample = Sequential(bn4, ReLU, Conv1x1): bn4 is registered twice (as .bn4 and .downsample.0), same tensors
        sd[prefix + '.downsample.0.weight'] = sd[prefix + '.bn4.weight']; sd[prefix + '.downsample.0.bias'] = sd[prefix + '.bn4.bias']
        _conv2d(rs, sd, prefix + '.downsample.2', cout, cin, 1, bias=False, gain=0.7)


def hgfilter_state_dict(seed: int = SEED + 11, prefix: str = '') -> Dict[str, np.ndarray]:
    """HGFilter(stack=1, depth=4, in_ch=6, last_ch=32, norm='group', down_type='no_down', use_sigmoid=False) keys
    (HGFilters.py:124-175, arch_recon.py:28)."""
    rs = np.random.RandomState(seed)
    sd: Dict[str, np.ndarray] = {}
    _conv2d(rs, sd, prefix + 'conv1', 64, 6, 7, bias=True)
    _gn(rs, sd, prefix + 'bn1', 64)
    _convblock(rs, sd, prefix + 'conv2', 64, 128)
    _convblock(rs, sd, prefix + 'conv3', 128, 128)
    _convblock(rs, sd, prefix + 'conv4', 128, 256)

    def hourglass(level):
        _convblock(rs, sd, prefix + 'm0.b1_%d' % level, 256, 256)
        _convblock(rs, sd, prefix + 'm0.b2_%d' % level, 256, 256)
        if level > 1:
            hourglass(level - 1)
        else:
            _convblock(rs, sd, prefix + 'm0.b2_plus_%d' % level, 256, 256)
        _convblock(rs, sd, prefix + 'm0.b3_%d' % level, 256, 256)
    hourglass(4)
    _convblock(rs, sd, prefix + 'top_m_0', 256, 256)
    _conv2d(rs, sd, prefix + 'conv_last0', 256, 256, 1, bias=True)
    _gn(rs, sd, prefix + 'bn_end0', 256)
    _conv2d(rs, sd, prefix + 'l0', 32, 256, 1, bias=True, gain=0.5)
    return sd


def smpl_pos_map(seed: int = SEED + 12) -> np.ndarray:
    """(1,6,256,256) f32 stand-in for the rendered SMPL position map (N(0, 0.3^2), seeded)."""
    return np.random.RandomState(seed).normal(0, 0.3, (1, 6, 256, 256)).astype(np.float32)


def normal_maps(seed: int = SEED + 13) -> np.ndarray:
    """cat([front_normal, back_normal], 1): (1,6,512,512) f32 ~ U(0,1), seeded."""
    return np.random.RandomState(seed).uniform(0, 1, (1, 6, 512, 512)).astype(np.float32)
