"""state_dict -> device weight blob.

Takes the reference's own checkpoint keys (``GeoTexAvatar.state_dict()`` / ``ReconNetwork.state_dict()``, i.e. what
main.py:302-320 loads) and produces the flat blob ``avc_load_avatar_weights`` / ``avc_load_recon_weights`` expect
(layout: csrc/common.cuh ``AvcBlobHeader``):

* BatchNorm1d(eval) of the OffsetDecoder (network/mlp.py:90-97,102-110) and weight-norm of the recon decoder
  (network/mlp.py:24-28) are folded into a per-output-channel ``scale``/``bias`` applied after the GEMM, so the weight
  matrices themselves stay bit-identical to the checkpoint;
* skip-concat layers keep the reference's column order as two K segments (mlp.py:61 activations-first,
  mlp.py:106 input-first);
* float32 section: W^T [K][N] per layer for the CUDA-core kernel (heads with N<=4 stay [N][K]);
* fp16 section: every layer as 16-wide k-steps, each a (hi, lo) pair of N x 16 slabs in the tcgen05 canonical
  K-major no-swizzle core-matrix layout (8 rows x 16 B contiguous; SBO = 256 B between 8-row groups, LBO = 128 B
  between the two k-halves). hi = fp16(W * 2^s), lo = fp16(W * 2^s - hi); the power-of-two s keeps lo out of the
  fp16 subnormal range and is undone by the tensor-core scale vector.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Optional, Tuple

import numpy as np

MAGIC = 0x57435641
VERSION = 3
MAX_LAYERS = 24
KIND_AVATAR, KIND_RECON = 0, 1
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_SOFTPLUS, ACT_SIGMOID = 0, 1, 2, 3, 4
_HDR_BYTES = 16 + 32 + MAX_LAYERS * 48


def _np(sd, key) -> np.ndarray:
    v = sd[key]
    if hasattr(v, 'detach'):
        v = v.detach().cpu().numpy()
    return np.asarray(v)


def _pad16(k: int) -> int:
    return (k + 15) // 16 * 16


class _Layer:
    def __init__(self, W: np.ndarray, scale: np.ndarray, bias: np.ndarray, k0: int, k1: int, act: int):
        assert W.ndim == 2 and W.shape[1] == k0 + k1, (W.shape, k0, k1)
        self.W = W.astype(np.float32); self.scale = scale.astype(np.float32); self.bias = bias.astype(np.float32)
        self.k0, self.k1, self.act = k0, k1, act
        self.n = W.shape[0]
        self.tc_perm0 = None      # optional column permutation of segment 0 for the tensor-core section


def _plain(sd, prefix: str, k0: int, k1: int, act: int) -> _Layer:
    W = _np(sd, prefix + '.weight')[:, :, 0]
    b = _np(sd, prefix + '.bias')
    return _Layer(W, np.ones(W.shape[0], np.float32), b, k0, k1, act)


def _bn_folded(sd, conv: str, bn: str, k0: int, k1: int) -> _Layer:
    """y = gamma * (Wx + b - mean) / sqrt(var + 1e-5) + beta  ->  scale * (Wx) + bias'   (eval mode, mlp.py:102-110)."""
    W = _np(sd, conv + '.weight')[:, :, 0]
    b = _np(sd, conv + '.bias').astype(np.float64)
    g = _np(sd, bn + '.weight').astype(np.float64); beta = _np(sd, bn + '.bias').astype(np.float64)
    mu = _np(sd, bn + '.running_mean').astype(np.float64); var = _np(sd, bn + '.running_var').astype(np.float64)
    s = g / np.sqrt(var + 1e-5)
    return _Layer(W, s, (b - mu) * s + beta, k0, k1, ACT_SOFTPLUS)


def _weight_normed(sd, prefix: str, k0: int, k1: int, act: int) -> _Layer:
    """W = g * v / ||v|| (norm over dims 1,2; mlp.py:24-28) -> v as the matrix, g/||v|| as the scale."""
    v = _np(sd, prefix + '.weight_v')[:, :, 0]
    g = _np(sd, prefix + '.weight_g')[:, 0, 0].astype(np.float64)
    nrm = np.sqrt((v.astype(np.float64) ** 2).sum(1))
    return _Layer(v, g / nrm, _np(sd, prefix + '.bias'), k0, k1, act)


def avatar_layers(sd) -> List[_Layer]:
    """Fixed order: 0..6 warp conv1..7 | 7 warp out | 8..14 shared fc0..6 | 15,16 geo | 17..19 clr."""
    L: List[_Layer] = []
    p = 'warping_field.mlp'
    w1 = _np(sd, p + '.conv1.weight')
    if w1.shape[1] != 67:
        raise ValueError('warping_field expects pos_encoding 0 (67 input channels), got %d' % w1.shape[1])
    for i in range(1, 8):
        k0, k1 = (67, 0) if i == 1 else ((67, 256) if i == 5 else (256, 0))
        L.append(_bn_folded(sd, '%s.conv%d' % (p, i), '%s.bn%d' % (p, i), k0, k1))
        if k0 == 67:
            # tensor-core kernel stages h0 as [f0..f63, x, y, z] so the 64 gathered channels are 16-byte aligned k-groups
            L[-1].tc_perm0 = np.concatenate([np.arange(3, 67), np.arange(0, 3)])
    L.append(_plain(sd, 'warping_field.out_layer_coord_affine', 256, 0, ACT_NONE))
    p = 'cano_template.shared_mlp.fc_list'
    if _np(sd, p + '.0.0.weight').shape[1] != 63:
        raise ValueError('cano_template expects pos_encoding 10 (63 input channels)')
    for l in range(6):
        k0, k1 = (63, 0) if l == 0 else ((256, 63) if l == 4 else (256, 0))
        L.append(_plain(sd, '%s.%d.0' % (p, l), k0, k1, ACT_RELU))
    L.append(_plain(sd, p + '.6', 256, 0, ACT_NONE))
    L.append(_plain(sd, 'cano_template.geo_mlp.fc_list.0.0', 256, 0, ACT_LRELU))
    L.append(_plain(sd, 'cano_template.geo_mlp.fc_list.1', 128, 0, ACT_NONE))
    L.append(_plain(sd, 'cano_template.clr_mlp.fc_list.0.0', 256, 0, ACT_RELU))
    L.append(_plain(sd, 'cano_template.clr_mlp.fc_list.1.0', 256, 0, ACT_RELU))
    L.append(_plain(sd, 'cano_template.clr_mlp.fc_list.2', 128, 0, ACT_SIGMOID))
    expect = [256] * 7 + [3] + [256] * 7 + [128, 2, 256, 128, 3]
    assert [l.n for l in L] == expect, [l.n for l in L]
    return L


def recon_layers(sd) -> List[_Layer]:
    p = 'image_decoder.fc_list'
    L = [_weight_normed(sd, p + '.0.0', 33, 0, ACT_LRELU),
         _weight_normed(sd, p + '.1.0', 512, 33, ACT_LRELU),
         _weight_normed(sd, p + '.2.0', 256, 33, ACT_LRELU),
         _plain(sd, p + '.3', 128, 0, ACT_SIGMOID)]
    assert [l.n for l in L] == [512, 256, 128, 1], [l.n for l in L]
    return L


def split_hi_lo(x: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """x (float32) -> (hi, lo) float16 with hi + lo ~= x to ~2^-22 relative."""
    hi = x.astype(np.float16)
    lo = (x - hi.astype(np.float32)).astype(np.float16)
    return hi, lo


def _slab(mat_nk16: np.ndarray) -> np.ndarray:
    """(np, 16) fp16 -> canonical K-major core-matrix order [np/8][2][8][8] flattened."""
    n = mat_nk16.shape[0]
    return np.ascontiguousarray(mat_nk16.reshape(n // 8, 8, 2, 8).transpose(0, 2, 1, 3)).reshape(-1)


STAGE_KSTEPS = 2      # k-steps per shared-memory ring stage (csrc/field_tc2.cu)


def tc_pieces(kind: int, layer: int, npad: int):
    """[(row_off, n_rows, [(first weight k-step, k-steps), ...] in issue order)] for one layer -- the op program of
    csrc/field_tc2.cu build_ops(): skip-input (shared-memory) k-steps are issued before the TMEM-fed ones."""
    if kind == KIND_AVATAR:
        segs = {0: [(0, 5)], 4: [(0, 5), (5, 16)], 8: [(0, 4)], 12: [(16, 4), (0, 16)], 16: [(0, 8)], 19: [(0, 8)]}.get(layer, [(0, 16)])
        return [(0, npad, segs)]
    return {0: [(0, 256, [(0, 3)]), (256, 256, [(0, 3)])],
            1: [(0, 256, [(0, 16)]), (0, 256, [(32, 3), (16, 16)])],
            2: [(0, 128, [(16, 3), (0, 16)])],
            3: [(0, 16, [(0, 8)])]}[layer]


def pack(layers: List[_Layer], kind: int) -> bytes:
    f32: List[np.ndarray] = []
    f32_len = 0

    def add_f32(a: np.ndarray) -> int:
        nonlocal f32_len
        pad = (-f32_len) % 4
        if pad:
            f32.append(np.zeros(pad, np.float32)); f32_len += pad
        off = f32_len
        f32.append(np.ascontiguousarray(a, dtype=np.float32).reshape(-1)); f32_len += a.size
        return off

    f16: List[np.ndarray] = []
    f16_bytes = 0
    descs = []
    for L in layers:
        n = L.n
        npad = max(_pad16(n), 16)
        k0p, k1p = _pad16(L.k0), (_pad16(L.k1) if L.k1 else 0)
        wt_off = add_f32(L.W if n <= 4 else L.W.T)
        sb_off = add_f32(np.concatenate([L.scale, L.bias]))
        # ---- tensor-core section ----
        amax = float(np.abs(L.W).max())
        shift = int(np.clip(np.floor(np.log2(1000.0 / max(amax, 1e-30))), 0, 14))
        Wp = np.zeros((npad, k0p + k1p), np.float32)
        Wp[:n, :L.k0] = L.W[:, :L.k0] if L.tc_perm0 is None else L.W[:, :L.k0][:, L.tc_perm0]
        if L.k1:
            Wp[:n, k0p:k0p + L.k1] = L.W[:, L.k0:]
        Wp = Wp * np.float32(2.0 ** shift)
        hi, lo = split_hi_lo(Wp)
        # The tensor-core kernel consumes a layer as a STREAM: for each piece (row block of an op), each N-half of 128 rows,
        # each ring stage (<= 2 k-steps, in issue order: shared-memory segment first), the hi slab then the lo slab of those
        # rows. One cp.async.bulk per stage moves it; tc_pieces() mirrors build_ops() in csrc/field_tc2.cu.
        tc_w_off = f16_bytes
        for row_off, n_piece, segs in tc_pieces(kind, len(descs), npad):
            halves = 2 if n_piece == 256 else 1
            rows = n_piece // halves
            for h in range(halves):
                r0 = row_off + h * rows
                for w0, ks in segs:
                    for j in range(0, ks, STAGE_KSTEPS):
                        for u in range(min(STAGE_KSTEPS, ks - j)):
                            kk = 16 * (w0 + j + u)
                            for part in (hi, lo):
                                sl = _slab(part[r0:r0 + rows, kk:kk + 16])
                                f16.append(sl); f16_bytes += sl.size * 2
        tsc = np.zeros(npad, np.float32); tbi = np.zeros(npad, np.float32)
        tsc[:n] = L.scale * np.float32(2.0 ** -shift); tbi[:n] = L.bias
        tc_sb_off = add_f32(np.concatenate([tsc, tbi]))
        descs.append((L.k0, L.k1, k0p, k1p, n, npad, L.act, wt_off, sb_off, tc_w_off, tc_sb_off, shift))
    f32_blob = np.concatenate(f32).astype(np.float32).tobytes() if f32 else b''
    f16_blob = np.concatenate(f16).astype(np.float16).tobytes() if f16 else b''
    f32_off = (_HDR_BYTES + 127) // 128 * 128
    f16_off = (f32_off + len(f32_blob) + 127) // 128 * 128
    hdr = struct.pack('<4I4Q', MAGIC, VERSION, kind, len(layers), f32_off, len(f32_blob), f16_off, len(f16_blob))
    for d in descs:
        hdr += struct.pack('<12i', *d)
    hdr += b'\0' * (_HDR_BYTES - len(hdr))
    out = bytearray(f16_off + len(f16_blob))
    out[:len(hdr)] = hdr
    out[f32_off:f32_off + len(f32_blob)] = f32_blob
    out[f16_off:f16_off + len(f16_blob)] = f16_blob
    return bytes(out)


def pack_avatar(state_dict) -> bytes:
    """GeoTexAvatar.state_dict() (or any mapping with its per-point keys) -> blob. The UNet keys are ignored."""
    return pack(avatar_layers(state_dict), KIND_AVATAR)


def pack_recon(state_dict) -> bytes:
    """ReconNetwork.state_dict() -> blob. The image_encoder (HGFilter) keys are ignored."""
    return pack(recon_layers(state_dict), KIND_RECON)


def parse_header(blob: bytes) -> Dict:
    magic, version, kind, n_layers, f32_off, f32_bytes, f16_off, f16_bytes = struct.unpack_from('<4I4Q', blob, 0)
    layers = []
    names = ('k0', 'k1', 'k0p', 'k1p', 'n', 'np', 'act', 'wt_off', 'sb_off', 'tc_w_off', 'tc_sb_off', 'shift')
    for l in range(n_layers):
        layers.append(dict(zip(names, struct.unpack_from('<12i', blob, 48 + 48 * l))))
    return dict(magic=magic, version=version, kind=kind, n_layers=n_layers, f32_off=f32_off, f32_bytes=f32_bytes,
                f16_off=f16_off, f16_bytes=f16_bytes, layers=layers)
