"""Per-frame encoders feeding the per-point path ("next" row 1).

    PoseFeatureEncoder   == WarpingField.unet = UnetNoCond7DS(6, 64, nf=32)     (unets.py:169-229, arch_avatar.py:95,109-111)
    ImageFeatureEncoder  == ReconNetwork.image_encoder = HGFilter(1,4,6,32,'group','no_down',False)
                                                                                  (HGFilters.py:124-219, arch_recon.py:28,41-43)

Both are small conv nets that run ONCE per frame; the per-point kernels then gather from their output 16.8 M times.

On sm_100 both run as PROGRAMS of kernels of this library (csrc/conv_tc.cu; ImageFeatureEncoderTC / PoseFeatureEncoderTC below build the
op program from the reference's state_dict and hand it to avc_encoder_create): tcgen05 implicit-GEMM 3x3 / 1x1 convolutions fed by TMA
tensor loads, split-K fp32 gather-GEMMs for the UNet's 4x4 stride-2 / transposed convolutions, GroupNorm / pool / bicubic / bilinear /
stem kernels, one CUDA graph per network: HGFilter 2.8 ms, UNet 0.36 ms, 7.4e-6 / 4.7e-6 from the reference.

PoseFeatureEncoder / ImageFeatureEncoder are the same networks on cuDNN (library convolutions), re-stated functionally from the reference's
state_dict -- the A/B of the programs and the path for other devices -- so that

  * eval-mode BatchNorm(affine=False) is folded into the conv weights at load time (one kernel less per layer),
  * the whole forward is captured once into a CUDA graph and replayed per frame (about 60 / 200 tiny launches otherwise,
    several of them on 2x2 .. 8x8 images where launch latency is everything),
  * the result is returned channels_last, i.e. already in the (H,W,C) layout the gather kernels read
    (avc_set_feature_map_hwc: a straight device copy instead of a transpose); the UNet runs channels_last throughout.

cuDNN switches (measured on the B200, tests/diag_encoders.py; errors are max-abs against the reference goldens):
    f32 (default)           UNet 2.2 ms (1.6e-6)   HGFilter 16.5 ms (3.8e-6)      deterministic=True: 4.6 / 15.2 ms, bit-identical replays
    allow_tf32=True         UNet 0.35 ms (8.7e-4)  HGFilter 7.0 ms (2.6e-3)       opt-in: outside the 1e-4 occupancy budget

Reference quirks kept (parity is tested against the reference modules themselves, tests/golden/encoder_golden.npz):
  * `Conv2DBlock.relu` is LeakyReLU(0.2, inplace=True) (unets.py:18-23): it rewrites the previous block's output in place, so
    every skip connection d1..d6 carries the ACTIVATED tensor;
  * `forward` applies `upconv3` twice and never `upconv4` (unets.py:214-215); the unused keys are accepted and ignored;
  * `UpConv2DBlock(up_mode='upsample')` = bilinear x2 with align_corners=False, then a 3x3 conv WITH bias (unets.py:41-44);
  * HourGlass up-sampling is bicubic with align_corners=True (HGFilters.py:115).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5   # nn.BatchNorm2d default


def _t(sd, key, device):
    v = sd[key]
    t = torch.from_numpy(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else v.detach()
    return t.to(device=device, dtype=torch.float32)


class _GraphedForward:
    """Static-shape forward replayed from a CUDA graph (device tensors only). Falls back to eager on the CPU (tests)."""

    def __init__(self, fn, device: torch.device, use_graph: bool):
        self.fn = fn
        self.device = device
        self.use_graph = bool(use_graph) and device.type == 'cuda'
        self._graph = None
        self._in = None
        self._out = None

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        x = x.to(device=self.device, dtype=torch.float32)
        if not self.use_graph:
            with torch.no_grad():
                return self.fn(x)
        if self._graph is None or self._in.shape != x.shape:
            self._in = torch.empty_like(x, memory_format=torch.channels_last)
            self._in.copy_(x)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side), torch.no_grad():
                for _ in range(2):                       # warm-up outside the capture (cuDNN algorithm selection, workspace)
                    self.fn(self._in)
            torch.cuda.current_stream(self.device).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g), torch.no_grad():
                self._out = self.fn(self._in)
            self._graph = g
        self._in.copy_(x)
        self._graph.replay()
        return self._out                                  # owned by the graph: valid until the next call


class PoseFeatureEncoder:
    """pose_feat_map = unet(smpl_pos_map): (1,6,256,256) -> (1,64,256,256)   (WarpingField.precompute_conv, arch_avatar.py:109-111).

    `state_dict` uses the reference's key names below `prefix` (e.g. 'warping_field.unet.' inside GeoTexAvatar.state_dict()).
    The returned tensor is channels_last (memory order H,W,C) and, with graphs on, is overwritten by the next call."""

    def __init__(self, state_dict: Dict, prefix: str = '', device='cuda', use_graph: bool = True, allow_tf32: bool = False,
                 deterministic: bool = False, channels_last: bool = True, benchmark: bool = False):
        self.device = torch.device(device)
        self.allow_tf32 = allow_tf32
        self.deterministic = deterministic
        self.benchmark = benchmark
        self.mf = torch.channels_last if channels_last else torch.contiguous_format
        d = self.device
        g = lambda k: _t(state_dict, prefix + k, d)       # noqa: E731

        def folded(wkey, bnkey, bias=None, transpose=False):
            """conv followed by BatchNorm(affine=False, eval): W' = W * s, b' = (b - mean) * s with s = 1/sqrt(var+eps), per OUT channel."""
            w = g(wkey)
            if bnkey is None:
                return w.contiguous(memory_format=self.mf), bias
            s = torch.rsqrt(g(bnkey + '.running_var').double() + BN_EPS)
            shape = (1, -1, 1, 1) if transpose else (-1, 1, 1, 1)
            w = (w.double() * s.view(shape)).float()
            b0 = bias.double() if bias is not None else torch.zeros_like(s)
            b = ((b0 - g(bnkey + '.running_mean').double()) * s).float()
            return w.contiguous(memory_format=self.mf), b
        self.down = [folded('conv%d.conv.weight' % i, 'conv%d.bn' % i if 2 <= i <= 6 else None) for i in range(1, 8)]
        self.up = [folded('upconv%d.up.weight' % i, 'upconv%d.bn' % i, transpose=True) for i in (1, 2, 3)]
        self.c5 = folded('upconvC5.up.1.weight', 'upconvC5.bn', g('upconvC5.up.1.bias'))
        self.c6 = folded('upconvC6.up.1.weight', 'upconvC6.bn', g('upconvC6.up.1.bias'))
        self.c7 = folded('upconvC7.up.1.weight', None, g('upconvC7.up.1.bias'))
        self._run = _GraphedForward(self._forward, self.device, use_graph)

    def _forward(self, x: torch.Tensor) -> torch.Tensor:
        with torch.backends.cudnn.flags(enabled=True, benchmark=self.benchmark, deterministic=self.deterministic, allow_tf32=self.allow_tf32):
            x = x.contiguous(memory_format=self.mf)
            a = []                                                   # activated skips a1..a6 (the in-place LeakyReLU quirk)
            h = F.conv2d(x, self.down[0][0], None, stride=2, padding=1)
            for w, b in self.down[1:]:
                h = F.leaky_relu(h, 0.2)
                a.append(h)
                h = F.conv2d(h, w, b, stride=2, padding=1)
            # h = d7 (2x2); shared decoder: upconv1, upconv2, upconv3, upconv3 AGAIN (unets.py:211-215)
            for (w, b), skip in zip((self.up[0], self.up[1], self.up[2], self.up[2]), (a[5], a[4], a[3], a[2])):
                h = torch.cat([F.conv_transpose2d(F.relu(h), w, b, stride=2, padding=1), skip], 1)
            for (w, b), skip in ((self.c5, a[1]), (self.c6, a[0]), (self.c7, None)):
                h = F.interpolate(F.relu(h), scale_factor=2, mode='bilinear', align_corners=False)
                h = F.conv2d(h, w, b, stride=1, padding=1)
                if skip is not None:
                    h = torch.cat([h, skip], 1)
            return h.contiguous(memory_format=torch.channels_last)

    def __call__(self, smpl_pos_map) -> torch.Tensor:
        x = torch.as_tensor(smpl_pos_map)
        if x.dim() != 4 or x.shape[1] != 6 or x.shape[2] % 128 or x.shape[3] % 128:
            raise ValueError('smpl_pos_map must be (B,6,H,W) with H,W multiples of 128 (7 stride-2 levels), got %s' % (tuple(x.shape),))
        return self._run(x)


class ImageFeatureEncoder:
    """img_feat_map = HGFilter(cat([front_normal, back_normal], 1))[0][-1]: (1,6,512,512) -> (1,32,256,256)
    (ReconNetwork.get_feat_maps, arch_recon.py:41-43,51-52). GroupNorm(32, C) everywhere (per-sample statistics: nothing to fold).
    Runs NCHW by default: cuDNN's f32 (non-TF32) kernels and GroupNorm are 1.5x faster in that layout on the B200 (16.5 vs 24.7 ms,
    tests/diag_encoders.py); only the final 32-channel map is converted to channels_last for the hand-off."""

    def __init__(self, state_dict: Dict, prefix: str = '', device='cuda', use_graph: bool = True, allow_tf32: bool = False,
                 deterministic: bool = False, channels_last: bool = False, benchmark: bool = True):
        self.device = torch.device(device)
        self.allow_tf32 = allow_tf32
        self.deterministic = deterministic
        self.benchmark = benchmark
        self.mf = torch.channels_last if channels_last else torch.contiguous_format
        self.p = {}
        for k in state_dict:
            if k.startswith(prefix):
                t = _t(state_dict, k, self.device)
                self.p[k[len(prefix):]] = t.contiguous(memory_format=self.mf) if t.dim() == 4 else t
        self._run = _GraphedForward(self._forward, self.device, use_graph)

    def _gn(self, x, name):
        return F.group_norm(x, 32, self.p[name + '.weight'], self.p[name + '.bias'], 1e-5)

    def _block(self, x, name):
        """ConvBlock.forward (HGFilters.py:61-75)."""
        p = self.p
        o1 = F.conv2d(F.relu(self._gn(x, name + '.bn1')), p[name + '.conv1.weight'], None, padding=1)
        o2 = F.conv2d(F.relu(self._gn(o1, name + '.bn2')), p[name + '.conv2.weight'], None, padding=1)
        o3 = F.conv2d(F.relu(self._gn(o2, name + '.bn3')), p[name + '.conv3.weight'], None, padding=1)
        out = torch.cat([o1, o2, o3], 1)
        if (name + '.downsample.2.weight') in p:
            x = F.conv2d(F.relu(self._gn(x, name + '.bn4')), p[name + '.downsample.2.weight'], None)
        return out + x

    def _hourglass(self, level, x):
        """HourGlass._forward (HGFilters.py:97-118)."""
        up1 = self._block(x, 'm0.b1_%d' % level)
        low = self._block(F.avg_pool2d(x, 2, stride=2), 'm0.b2_%d' % level)
        low = self._hourglass(level - 1, low) if level > 1 else self._block(low, 'm0.b2_plus_%d' % level)
        low = self._block(low, 'm0.b3_%d' % level)
        return up1 + F.interpolate(low, scale_factor=2, mode='bicubic', align_corners=True)

    def _forward(self, x: torch.Tensor) -> torch.Tensor:
        p = self.p
        with torch.backends.cudnn.flags(enabled=True, benchmark=self.benchmark, deterministic=self.deterministic, allow_tf32=self.allow_tf32):
            x = x.contiguous(memory_format=self.mf)
            x = F.relu(self._gn(F.conv2d(x, p['conv1.weight'], p['conv1.bias'], stride=2, padding=3), 'bn1'))
            x = self._block(x, 'conv2')                               # down_type == 'no_down'
            x = self._block(self._block(x, 'conv3'), 'conv4')
            ll = self._block(self._hourglass(4, x), 'top_m_0')
            ll = F.relu(self._gn(F.conv2d(ll, p['conv_last0.weight'], p['conv_last0.bias']), 'bn_end0'))
            out = F.conv2d(ll, p['l0.weight'], p['l0.bias'])           # use_sigmoid=False: no tanh
            return out.contiguous(memory_format=torch.channels_last)

    def __call__(self, normals) -> torch.Tensor:
        x = torch.as_tensor(normals)
        if x.dim() != 4 or x.shape[1] != 6 or x.shape[2] % 32 or x.shape[3] % 32:
            raise ValueError('normal maps must be (B,6,H,W) with H,W multiples of 32, got %s' % (tuple(x.shape),))
        return self._run(x)


# ---------------------------------------------------------------------------------------------------------------------------------
# HGFilter on the tensor cores: the network as a PROGRAM of library ops (csrc/conv_tc.cu). This module only restates the STRUCTURE of
# HGFilter.forward / ConvBlock.forward / HourGlass._forward (HGFilters.py:61-75, 97-118, 177-219) as a list of ops over numbered
# buffers and packs the weights (fp16 hi / lo planes, (C_out, taps, C_in) K-major, power-of-two pre-scale); every arithmetic
# operation runs in the CUDA library: tcgen05 implicit-GEMM convolutions fed by TMA tensor loads, GroupNorm, pooling, bicubic
# up-sampling and the 7x7 stem as kernels of ours, all captured in one CUDA graph.
OP_STEM, OP_GN, OP_CONV, OP_ADD, OP_POOL, OP_UPADD, OP_INPUT, OP_UPSPLIT, OP_COPY, OP_CONV4 = 1, 2, 3, 4, 5, 6, 7, 8, 9, 10


class _Program:
    def __init__(self):
        self.f32_sizes, self.planes, self.plane_key, self.ops = [], [], {}, []
        self.weights = bytearray()
        self.params = []
        self.marks = {}                    # name -> (f32 buffer, (H, W, C)): intermediate tensors the tests compare one by one

    def buf(self, n_floats: int) -> int:
        self.f32_sizes.append(int(n_floats)); return len(self.f32_sizes) - 1

    def plane(self, pixels: int, c: int) -> int:
        """fp16 hi / lo plane pair (pixels, max(c, 64)); one per distinct shape (its only reader is the convolution right behind)"""
        key = (int(pixels), max(64, int(c)))
        if key not in self.plane_key:
            self.planes.append(key); self.plane_key[key] = len(self.planes) - 1
        return self.plane_key[key]

    def param(self, arr) -> int:
        off = sum(len(a) for a in self.params)
        self.params.append(np.asarray(arr, np.float32).reshape(-1)); return off

    def conv_weight_bias(self, w: np.ndarray, bias):
        """conv_weight + a bias vector (or None) -> (weight offset, scale exponent, C_in_pad, bias offset | -1)"""
        off, s, cpad = self.conv_weight(w)
        return off, s, cpad, (-1 if bias is None else self.param(bias))

    def conv_weight(self, w: np.ndarray):
        """(C_out, C_in, kh, kw) f32 -> byte offset of [hi plane | lo plane], each (C_out, taps, C_in_pad) fp16, and the scale exponent"""
        co, ci, kh, kw = w.shape
        cpad = max(64, ci)
        m = float(np.abs(w).max())
        s = int(np.floor(-np.log2(m))) if m > 0 else 0          # max |w| * 2^s in [0.5, 1): the fp16 lo parts stay out of the subnormals
        s = max(-40, min(40, s))
        ws = np.zeros((co, kh * kw, cpad), np.float32)
        ws[:, :, :ci] = np.ldexp(w.astype(np.float32), s).transpose(0, 2, 3, 1).reshape(co, kh * kw, ci)
        hi = ws.astype(np.float16); lo = (ws - hi.astype(np.float32)).astype(np.float16)
        while len(self.weights) % 128:
            self.weights += b'\0'
        off = len(self.weights)
        self.weights += hi.tobytes() + lo.tobytes()
        return off, s, cpad

    def op(self, *words):
        w = list(int(x) for x in words) + [0] * (16 - len(words)); assert len(w) == 16
        self.ops.append(w)

    def pack(self, in_chw, out_buf, out_hwc):
        hdr = [0x45435641, len(self.f32_sizes), len(self.planes), len(self.ops), in_chw[0], in_chw[1], in_chw[2], out_buf, out_hwc[2], out_hwc[0], out_hwc[1]]
        words = hdr + [0] * (16 - len(hdr)) + self.f32_sizes + [x for pl in self.planes for x in pl] + [x for o in self.ops for x in o]
        return (np.asarray(words, np.int32), bytes(self.weights), np.concatenate(self.params).astype(np.float32) if self.params else np.zeros(1, np.float32))


def build_hgfilter_program(state_dict: Dict, prefix: str = '', in_hw=(512, 512)):
    """HGFilter(1, 4, 6, 32, 'group', 'no_down', False).forward (HGFilters.py:177-219) -> (program int32, weights fp16 bytes, params f32)."""
    g = lambda k: np.asarray(state_dict[prefix + k].detach().cpu().numpy() if hasattr(state_dict[prefix + k], 'detach') else state_dict[prefix + k], np.float32)   # noqa: E731
    Hin, Win = in_hw
    if Hin % 32 or Win % 32 or Hin != Win:
        raise ValueError('normal maps must be square with an edge that is a multiple of 32')
    H, W = Hin // 2, Win // 2
    pr = _Program()

    def gn(src, P, C, ld, c_off, name, relu, plane=-1, dst32=-1):
        if name is None:
            pr.op(OP_GN, src, P, C, ld, c_off, -1, -1, relu, plane, dst32, C)
        else:
            pr.op(OP_GN, src, P, C, ld, c_off, pr.param(g(name + '.weight')), pr.param(g(name + '.bias')), relu, plane, dst32, C)

    def conv(plane, wname, out, h, w, n_out, c_off, ldc, accumulate=0, bias=None):
        wt = g(wname)
        off, s, cpad = pr.conv_weight(wt)
        assert wt.shape[0] == n_out
        pr.op(OP_CONV, plane, off, out, h, w, cpad, n_out, wt.shape[2] * wt.shape[3], c_off, ldc, accumulate, -1 if bias is None else pr.param(g(bias)), s)

    def block(x, h, w, cin, cout, name):
        """ConvBlock.forward (HGFilters.py:61-75): the three 3x3 convolutions write the channel slices of ONE output tensor (the
        reference's torch.cat), the residual is added last (identity: add; otherwise the 1x1 convolution accumulates into it)."""
        P = h * w
        out = pr.buf(P * cout)
        c1, c2 = cout // 2, cout // 4
        gn(x, P, cin, cin, 0, name + '.bn1', 1, plane=pr.plane(P, cin))
        conv(pr.plane(P, cin), name + '.conv1.weight', out, h, w, c1, 0, cout)
        gn(out, P, c1, cout, 0, name + '.bn2', 1, plane=pr.plane(P, c1))
        conv(pr.plane(P, c1), name + '.conv2.weight', out, h, w, c2, c1, cout)
        gn(out, P, c2, cout, c1, name + '.bn3', 1, plane=pr.plane(P, c2))
        conv(pr.plane(P, c2), name + '.conv3.weight', out, h, w, c2, c1 + c2, cout)
        if cin != cout:
            gn(x, P, cin, cin, 0, name + '.bn4', 1, plane=pr.plane(P, cin))
            conv(pr.plane(P, cin), name + '.downsample.2.weight', out, h, w, cout, 0, cout, accumulate=1)
        else:
            pr.op(OP_ADD, out, x, P * cout)
        return out

    def hourglass(level, x, h, w):
        """HourGlass._forward (HGFilters.py:97-118), 256 features."""
        up1 = block(x, h, w, 256, 256, 'm0.b1_%d' % level)
        low = pr.buf((h // 2) * (w // 2) * 256)
        pr.op(OP_POOL, x, low, h, w, 256)
        low = block(low, h // 2, w // 2, 256, 256, 'm0.b2_%d' % level)
        low = hourglass(level - 1, low, h // 2, w // 2) if level > 1 else block(low, h // 2, w // 2, 256, 256, 'm0.b2_plus_%d' % level)
        low = block(low, h // 2, w // 2, 256, 256, 'm0.b3_%d' % level)
        out = pr.buf(h * w * 256)
        pr.op(OP_UPADD, up1, low, out, h // 2, w // 2, 256)
        return out

    P = H * W
    stem = pr.buf(P * 64)
    pr.op(OP_STEM, pr.param(g('conv1.weight')), pr.param(g('conv1.bias')), stem, Hin, Win)        # conv1 7x7 s2 (HGFilters.py:180)
    pr.marks['stem'] = (stem, (H, W, 64))
    x = pr.buf(P * 64)
    gn(stem, P, 64, 64, 0, 'bn1', 1, dst32=x)                                                  # relu(bn1(.))
    pr.marks['bn1'] = (x, (H, W, 64))
    x = block(x, H, W, 64, 128, 'conv2')                                                       # down_type == 'no_down'
    pr.marks['conv2'] = (x, (H, W, 128))
    x = block(x, H, W, 128, 128, 'conv3')
    pr.marks['conv3'] = (x, (H, W, 128))
    x = block(x, H, W, 128, 256, 'conv4')
    pr.marks['conv4'] = (x, (H, W, 256))
    x = hourglass(4, x, H, W)
    pr.marks['hourglass'] = (x, (H, W, 256))
    x = block(x, H, W, 256, 256, 'top_m_0')
    pr.marks['top_m_0'] = (x, (H, W, 256))
    ll = pr.buf(P * 256)
    gn(x, P, 256, 256, 0, None, 0, plane=pr.plane(P, 256))                                     # conv_last0 reads the raw tensor
    conv(pr.plane(P, 256), 'conv_last0.weight', ll, H, W, 256, 0, 256, bias='conv_last0.bias')
    gn(ll, P, 256, 256, 0, 'bn_end0', 1, plane=pr.plane(P, 256))
    pr.marks['conv_last0'] = (ll, (H, W, 256))
    out = pr.buf(P * 32)
    conv(pr.plane(P, 256), 'l0.weight', out, H, W, 32, 0, 32, bias='l0.bias')                  # use_sigmoid=False: no tanh
    build_hgfilter_program.last_marks = pr.marks
    return pr.pack((6, Hin, Win), out, (H, W, 32))


def _unet_getter(state_dict, prefix):
    return lambda k: np.asarray(state_dict[prefix + k].detach().cpu().numpy() if hasattr(state_dict[prefix + k], 'detach') else state_dict[prefix + k], np.float64)   # noqa: E731


def _emit_unet_tail(pr: '_Program', g, u4: int, a2: int, a1: int, H: int, W: int) -> int:
    """upconvC5 / C6 / C7 (unets.py:217-219) on buffers u4 (H/8, W/8, 384), a2 (H/4, W/4, 64), a1 (H/2, W/2, 32) -> the output buffer (H, W, 64)"""
    def folded(name, bn):
        w = g(name + '.up.1.weight'); b = g(name + '.up.1.bias')
        if bn:
            s = 1.0 / np.sqrt(g(name + '.bn.running_var') + BN_EPS)
            w = w * s[:, None, None, None]; b = (b - g(name + '.bn.running_mean')) * s
        return w.astype(np.float32), b.astype(np.float32)

    h8, w8, h4, w4, h2, w2 = H // 8, W // 8, H // 4, W // 4, H // 2, W // 2
    # upconvC5: relu(u4) -> up -> conv 384 -> 64 (+bn) ; cat a2
    c5 = pr.buf(h4 * w4 * 128)
    pr.op(OP_UPSPLIT, u4, h8, w8, 384, 384, 0, 1, pr.plane(h4 * w4, 384))
    wt, b = folded('upconvC5', True); off, s_, cpad, boff = pr.conv_weight_bias(wt, b)
    pr.op(OP_CONV, pr.plane(h4 * w4, 384), off, c5, h4, w4, cpad, 64, 9, 0, 128, 0, boff, s_)
    pr.op(OP_COPY, a2, c5, h4 * w4, 64, 128, 64)
    # upconvC6: relu(c5) -> up -> conv 128 -> 32 (+bn) ; cat a1
    c6 = pr.buf(h2 * w2 * 64)
    pr.op(OP_UPSPLIT, c5, h4, w4, 128, 128, 0, 1, pr.plane(h2 * w2, 128))
    wt, b = folded('upconvC6', True); off, s_, cpad, boff = pr.conv_weight_bias(wt, b)
    pr.op(OP_CONV, pr.plane(h2 * w2, 128), off, c6, h2, w2, cpad, 32, 9, 0, 64, 0, boff, s_)
    pr.op(OP_COPY, a1, c6, h2 * w2, 32, 64, 32)
    # upconvC7: relu(c6) -> up -> conv 64 -> 64 + bias (no bn)
    out = pr.buf(H * W * 64)
    pr.op(OP_UPSPLIT, c6, h2, w2, 64, 64, 0, 1, pr.plane(H * W, 64))
    wt, b = folded('upconvC7', False); off, s_, cpad, boff = pr.conv_weight_bias(wt, b)
    pr.op(OP_CONV, pr.plane(H * W, 64), off, out, H, W, cpad, 64, 9, 0, 64, 0, boff, s_)
    return out


def build_unet_tail_program(state_dict: Dict, prefix: str = '', out_hw=(256, 256)):
    """The three `UpConv2DBlock(up_mode='upsample')` stages that end UnetNoCond7DS.forward (unets.py:217-219: upconvC5, upconvC6,
    upconvC7 = relu -> bilinear x2 -> 3x3 convolution (+ eval BatchNorm, folded) [-> cat skip]) as a library program: 7.8 of the UNet's
    10.5 GFLOP. Input = one flat f32 buffer [u4 (H/8, W/8, 384) | a2 (H/4, W/4, 64) | a1 (H/2, W/2, 32)], all (H, W, C) order; output
    (H, W, 64). -> (program, weights, params, input float counts)."""
    g = _unet_getter(state_dict, prefix)
    H, W = out_hw
    if H % 128 or W % 128:
        raise ValueError('the UNet needs H, W multiples of 128')
    pr = _Program()
    n_u4, n_a2, n_a1 = (H // 8) * (W // 8) * 384, (H // 4) * (W // 4) * 64, (H // 2) * (W // 2) * 32
    u4 = pr.buf(n_u4); a2 = pr.buf(n_a2); a1 = pr.buf(n_a1)
    pr.op(OP_INPUT, u4, 0, n_u4); pr.op(OP_INPUT, a2, n_u4, n_a2); pr.op(OP_INPUT, a1, n_u4 + n_a2, n_a1)
    out = _emit_unet_tail(pr, g, u4, a2, a1, H, W)
    return pr.pack((n_u4 + n_a2 + n_a1, 1, 1), out, (H, W, 64)) + ((n_u4, n_a2, n_a1),)


def pack_conv4_weight(w: np.ndarray, transposed: bool) -> np.ndarray:
    """4x4 stride-2 kernels in the K-major order csrc/conv_tc.cu conv4_gemm_kernel streams: Conv2d weight (Co, Ci, 4, 4) ->
    [(ky*4 + kx) * Ci + ci][co]; ConvTranspose2d weight (Ci, Co, 4, 4) -> four output-parity classes (ry, rx), each
    [(ty*2 + tx) * Ci + ci][co] holding kernel element (ky, kx) = (1 - ry + 2 ty, 1 - rx + 2 tx)."""
    w = np.asarray(w, np.float32)
    if not transposed:
        co, ci = w.shape[:2]
        return np.ascontiguousarray(w.transpose(2, 3, 1, 0)).reshape(16 * ci, co)
    ci, co = w.shape[:2]
    out = np.empty((4, 4 * ci, co), np.float32)
    for ry in range(2):
        for rx in range(2):
            for ty in range(2):
                for tx in range(2):
                    out[ry * 2 + rx, (ty * 2 + tx) * ci:(ty * 2 + tx + 1) * ci] = w[:, :, 1 - ry + 2 * ty, 1 - rx + 2 * tx]
    return out


def build_unet_program(state_dict: Dict, prefix: str = '', in_hw=(256, 256)):
    """All of UnetNoCond7DS.forward (unets.py:201-219) as ONE library program: conv1..7 (4x4 stride-2 convolutions, eval BatchNorm folded,
    in-place LeakyReLU), upconv1, 2, 3, 3 (transposed 4x4 stride-2 convolutions on relu(cat)) as split-K fp32 gather-GEMMs (OP_CONV4),
    then upconvC5 / C6 / C7 on the tcgen05 convolution. Every encoder output is written straight into the channel slice of the concatenated
    buffer its skip connection feeds (the reference's torch.cat), and read from there by the next encoder level.
    Input (6, H, W) as the reference feeds it, output (H, W, 64). -> (program, weights, params)."""
    g = _unet_getter(state_dict, prefix)
    H, W = in_hw
    if H % 128 or W % 128:
        raise ValueError('the UNet needs H, W multiples of 128')
    pr = _Program()

    def fold(wkey, bnkey, transposed=False):
        w = g(wkey)
        if bnkey is None:
            return w.astype(np.float32), None
        s = 1.0 / np.sqrt(g(bnkey + '.running_var') + BN_EPS)
        w = w * (s[None, :, None, None] if transposed else s[:, None, None, None])
        return w.astype(np.float32), (-g(bnkey + '.running_mean') * s).astype(np.float32)

    def conv4(src, dst, hin, win, ci, ld_src, c_off_src, co, ld_dst, c_off_dst, wb, transposed=False, in_relu=False, leaky=False):
        w, b = wb
        assert w.shape[:2] == ((ci, co) if transposed else (co, ci)), (w.shape, ci, co)
        pr.op(OP_CONV4, src, dst, hin, win, ci, ld_src, c_off_src, co, ld_dst, c_off_dst, pr.param(pack_conv4_weight(w, transposed)),
              -1 if b is None else pr.param(b), (1 if transposed else 0) | (2 if in_relu else 0) | (4 if leaky else 0))

    down = [fold('conv%d.conv.weight' % i, 'conv%d.bn' % i if 2 <= i <= 6 else None) for i in range(1, 8)]
    up = [fold('upconv%d.up.weight' % i, 'upconv%d.bn' % i, transposed=True) for i in (1, 2, 3)]
    hs = [H >> i for i in range(8)]; ws = [W >> i for i in range(8)]            # extent after i stride-2 levels
    a1 = pr.buf(hs[1] * ws[1] * 32)                 # d1 (skip of upconvC6)
    a2 = pr.buf(hs[2] * ws[2] * 64)                 # d2 (skip of upconvC5)
    cat4 = pr.buf(hs[3] * ws[3] * 384)              # [upconv3(u3) 256 | d3 128] = u4
    cat3 = pr.buf(hs[4] * ws[4] * 512)              # [upconv3(u2) 256 | d4 256]
    cat2 = pr.buf(hs[5] * ws[5] * 512)              # [upconv2(u1) 256 | d5 256]
    cat1 = pr.buf(hs[6] * ws[6] * 512)              # [upconv1(d7) 256 | d6 256]
    d7 = pr.buf(hs[7] * ws[7] * 256)
    conv4(-1, a1, H, W, 6, 0, 0, 32, 32, 0, down[0], leaky=True)
    conv4(a1, a2, hs[1], ws[1], 32, 32, 0, 64, 64, 0, down[1], leaky=True)
    conv4(a2, cat4, hs[2], ws[2], 64, 64, 0, 128, 384, 256, down[2], leaky=True)
    conv4(cat4, cat3, hs[3], ws[3], 128, 384, 256, 256, 512, 256, down[3], leaky=True)
    conv4(cat3, cat2, hs[4], ws[4], 256, 512, 256, 256, 512, 256, down[4], leaky=True)
    conv4(cat2, cat1, hs[5], ws[5], 256, 512, 256, 256, 512, 256, down[5], leaky=True)
    conv4(cat1, d7, hs[6], ws[6], 256, 512, 256, 256, 256, 0, down[6])
    conv4(d7, cat1, hs[7], ws[7], 256, 256, 0, 256, 512, 0, up[0], transposed=True, in_relu=True)
    conv4(cat1, cat2, hs[6], ws[6], 512, 512, 0, 256, 512, 0, up[1], transposed=True, in_relu=True)
    conv4(cat2, cat3, hs[5], ws[5], 512, 512, 0, 256, 512, 0, up[2], transposed=True, in_relu=True)
    conv4(cat3, cat4, hs[4], ws[4], 512, 512, 0, 256, 384, 0, up[2], transposed=True, in_relu=True)          # upconv3 AGAIN (unets.py:215)
    out = _emit_unet_tail(pr, g, cat4, a2, a1, H, W)
    return pr.pack((6, H, W), out, (H, W, 64))


class PoseFeatureEncoderTC(PoseFeatureEncoder):
    """PoseFeatureEncoder on kernels of this library: same call, same result layout.

    head='library' (default): the whole UNet is ONE program (build_unet_program): the 4x4 stride-2 encoder and the transposed
    convolutions of the shared decoder as split-K fp32 gather-GEMMs (weight streaming over 2^2 .. 128^2 pixels; cuDNN spends 1.35 ms there
    on grids of 8-16 blocks), the three final 3x3 stages (three quarters of the FLOPs) on the tcgen05 convolution kernel; one CUDA graph.
    head='cudnn': the stride-2 / transposed head on the cuDNN CUDA-graph replay of the parent class, only the tail in the library (A/B)."""

    def __init__(self, state_dict: Dict, prefix: str = '', engine=None, in_hw=(256, 256), use_graph: bool = True, deterministic: bool = False,
                 head: str = 'library'):
        import ctypes as C
        from .engine import default_engine
        if head not in ('library', 'cudnn'):
            raise ValueError("head must be 'library' or 'cudnn'")
        self.engine = engine if engine is not None else default_engine()
        self.head = head
        self.in_hw = tuple(in_hw)
        self.use_graph = bool(use_graph)
        if head == 'cudnn':
            # `deterministic` concerns the cuDNN head only (its transposed convolutions may pick atomics-based algorithms); the library is
            super().__init__(state_dict, prefix=prefix, device=self.engine.device, use_graph=False, channels_last=True, deterministic=deterministic)
            prog, wbytes, params, self._in_counts = build_unet_tail_program(state_dict, prefix, in_hw)
            self._tail_in = torch.empty(sum(self._in_counts), device=self.device, dtype=torch.float32)
            self._head = _GraphedForward(self._forward_head, self.device, use_graph)
        else:
            self.device = self.engine.device
            prog, wbytes, params = build_unet_program(state_dict, prefix, in_hw)
            self._in = torch.empty((6,) + self.in_hw, device=self.device, dtype=torch.float32)
        h = C.c_void_p()
        e = self.engine
        e._check(e.lib.avc_encoder_create(e._h, prog.ctypes.data_as(C.c_void_p), len(prog), wbytes, len(wbytes), params.ctypes.data_as(C.c_void_p),
                                          len(params), C.byref(h)))
        self._h = h
        self._out = torch.empty((in_hw[0], in_hw[1], 64), device=self.device, dtype=torch.float32)

    def _forward_head(self, x: torch.Tensor) -> torch.Tensor:
        """conv1..7 + the shared decoder (unets.py:201-215) on cuDNN; writes [u4 | a2 | a1] in (H, W, C) order into the tail's input buffer"""
        with torch.backends.cudnn.flags(enabled=True, benchmark=self.benchmark, deterministic=self.deterministic, allow_tf32=self.allow_tf32):
            x = x.contiguous(memory_format=self.mf)
            a = []
            h = F.conv2d(x, self.down[0][0], None, stride=2, padding=1)
            for w, b in self.down[1:]:
                h = F.leaky_relu(h, 0.2)
                a.append(h)
                h = F.conv2d(h, w, b, stride=2, padding=1)
            for (w, b), skip in zip((self.up[0], self.up[1], self.up[2], self.up[2]), (a[5], a[4], a[3], a[2])):
                h = torch.cat([F.conv_transpose2d(F.relu(h), w, b, stride=2, padding=1), skip], 1)
            n4, n2, n1 = self._in_counts
            self._tail_in[:n4].view(h.shape[2], h.shape[3], h.shape[1]).copy_(h[0].permute(1, 2, 0))
            self._tail_in[n4:n4 + n2].view(a[1].shape[2], a[1].shape[3], a[1].shape[1]).copy_(a[1][0].permute(1, 2, 0))
            self._tail_in[n4 + n2:].view(a[0].shape[2], a[0].shape[3], a[0].shape[1]).copy_(a[0][0].permute(1, 2, 0))
            return self._tail_in

    def close(self) -> None:
        if getattr(self, '_h', None) and getattr(self.engine, '_h', None):
            self.engine.lib.avc_encoder_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __call__(self, smpl_pos_map) -> torch.Tensor:
        import ctypes as C
        x = torch.as_tensor(smpl_pos_map)
        if x.dim() != 4 or x.shape[0] != 1 or x.shape[1] != 6 or tuple(x.shape[2:]) != self.in_hw:
            raise ValueError('smpl_pos_map must be (1,6,%d,%d), got %s' % (self.in_hw + (tuple(x.shape),)))
        if self.head == 'cudnn':
            self._head(x)
            src = self._tail_in
        else:
            self._in.copy_(x[0].to(device=self.device, dtype=torch.float32))          # (6,H,W) as the reference feeds it; a stable pointer for the graph
            src = self._in
        e = self.engine
        e._check(e.lib.avc_encoder_run(self._h, C.c_void_p(src.data_ptr()), C.c_void_p(self._out.data_ptr()), int(self.use_graph), e._stream()))
        return self._out.permute(2, 0, 1)[None]            # (1,64,H,W) view with channels_last strides; overwritten by the next call


class ImageFeatureEncoderTC:
    """ImageFeatureEncoder on kernels of this library (tcgen05 convolutions): same call, same result layout.
    img_feat_map = HGFilter(cat([front_normal, back_normal], 1))[0][-1]: (1,6,512,512) -> (1,32,256,256), channels_last."""

    def __init__(self, state_dict: Dict, prefix: str = '', engine=None, in_hw=(512, 512), use_graph: bool = True):
        import ctypes as C
        from .engine import default_engine
        self.engine = engine if engine is not None else default_engine()
        self.device = self.engine.device
        self.use_graph = bool(use_graph)
        self.in_hw = tuple(in_hw)
        prog, wbytes, params = build_hgfilter_program(state_dict, prefix, in_hw)
        self.marks = dict(build_hgfilter_program.last_marks)
        self._out_hwc = (in_hw[0] // 2, in_hw[1] // 2, 32)
        h = C.c_void_p()
        e = self.engine
        e._check(e.lib.avc_encoder_create(e._h, prog.ctypes.data_as(C.c_void_p), len(prog), wbytes, len(wbytes), params.ctypes.data_as(C.c_void_p),
                                          len(params), C.byref(h)))
        self._h = h
        self._in = torch.empty((6,) + self.in_hw, device=self.device, dtype=torch.float32)
        self._out = torch.empty(self._out_hwc, device=self.device, dtype=torch.float32)

    def close(self) -> None:
        if getattr(self, '_h', None) and getattr(self.engine, '_h', None):
            self.engine.lib.avc_encoder_destroy(self._h)
        self._h = None

    def intermediate(self, name: str) -> torch.Tensor:
        """(H, W, C) f32 copy of a marked intermediate tensor of the last run (tests / debugging)."""
        import ctypes as C
        buf, shape = self.marks[name]
        t = torch.empty(shape, device=self.device, dtype=torch.float32)
        e = self.engine
        e._check(e.lib.avc_encoder_read_buffer(self._h, int(buf), C.c_void_p(t.data_ptr()), t.numel(), e._stream()))
        return t

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __call__(self, normals) -> torch.Tensor:
        import ctypes as C
        x = torch.as_tensor(normals)
        if x.dim() != 4 or x.shape[0] != 1 or x.shape[1] != 6 or tuple(x.shape[2:]) != self.in_hw:
            raise ValueError('normal maps must be (1,6,%d,%d), got %s' % (self.in_hw + (tuple(x.shape),)))
        self._in.copy_(x[0].to(device=self.device, dtype=torch.float32))
        e = self.engine
        e._check(e.lib.avc_encoder_run(self._h, C.c_void_p(self._in.data_ptr()), C.c_void_p(self._out.data_ptr()), int(self.use_graph), e._stream()))
        return self._out.permute(2, 0, 1)[None]            # (1,32,H,W) view with channels_last strides; overwritten by the next call


def subsample_index(c: int, h: int, w: int, count: int, seed: int) -> np.ndarray:
    """Seeded flat (h*w) pixel positions used by the golden fixtures (all channels are kept at each position)."""
    return np.sort(np.random.RandomState(seed).choice(h * w, size=min(count, h * w), replace=False))
