"""TEST INFRASTRUCTURE: a CPU interpreter of the encoder op program that avatarcap_b200/encoders.py builds for the library
(csrc/conv_tc.cu interprets the same words on the device). It executes every op with plain torch-CPU float32 arithmetic on the
(H, W, C) buffers the program numbers, taking the convolution weights from the packed fp16 hi / lo planes exactly as the kernel sees
them (hi + lo, un-scaled by 2^-s). What it pins without a GPU: the STRUCTURE of the program (op order, buffer / slice bookkeeping,
residual paths), the weight packing ((C_out, taps, C_in_pad) layout, tap order, scale exponents, channel padding) and the parameter
offsets -- against the reference's own HGFilter output (tests/golden/encoder_golden.npz)."""
import numpy as np
import torch
import torch.nn.functional as F

OP_STEM, OP_GN, OP_CONV, OP_ADD, OP_POOL, OP_UPADD, OP_INPUT, OP_UPSPLIT, OP_COPY, OP_CONV4 = 1, 2, 3, 4, 5, 6, 7, 8, 9, 10


def run_program(prog: np.ndarray, weights: bytes, params: np.ndarray, x_chw: np.ndarray) -> np.ndarray:
    assert prog[0] == 0x45435641
    nb, npl, nops = int(prog[1]), int(prog[2]), int(prog[3])
    out_buf, out_c, out_h, out_w = int(prog[7]), int(prog[8]), int(prog[9]), int(prog[10])
    sizes = prog[16:16 + nb]
    planes = prog[16 + nb:16 + nb + 2 * npl].reshape(npl, 2)
    ops = prog[16 + nb + 2 * npl:].reshape(nops, 16)
    bufs = [torch.zeros(int(s), dtype=torch.float32) for s in sizes]
    plane = [torch.zeros((int(p), int(c)), dtype=torch.float32) for p, c in planes]       # what hi + lo represent
    wbytes = np.frombuffer(weights, dtype=np.float16)
    x = torch.from_numpy(np.ascontiguousarray(x_chw))
    flat_in = x.reshape(-1)
    for op in ops:
        k = int(op[0])
        if k == OP_STEM:
            w = torch.from_numpy(params[op[1]:op[1] + 64 * 6 * 49].reshape(64, 6, 7, 7).copy()); b = torch.from_numpy(params[op[2]:op[2] + 64].copy())
            y = F.conv2d(x[None], w, b, stride=2, padding=3)[0].permute(1, 2, 0).contiguous()
            bufs[op[3]][:y.numel()] = y.reshape(-1)
        elif k == OP_GN:
            src, P, C, ld, c_off, g_off, b_off, relu, pl, dst32, ld32 = (int(v) for v in op[1:12])
            v = bufs[src][:P * ld].view(P, ld)[:, c_off:c_off + C]
            if g_off >= 0:
                gamma = torch.from_numpy(params[g_off:g_off + C].copy()); beta = torch.from_numpy(params[b_off:b_off + C].copy())
                v = F.group_norm(v.t().reshape(1, C, P), 32, gamma, beta, 1e-5)[0].t()
            if relu:
                v = torch.relu(v)
            if pl >= 0:
                plane[pl][:, :C] = v                      # channels >= C keep whatever an earlier, wider use left there (finite; their weights are zero)
            if dst32 >= 0:
                bufs[dst32][:P * ld32].view(P, ld32)[:, :C] = v
        elif k == OP_CONV:
            pl, w_off, out, H, W, cin, N, taps, c_off, ldc, acc, b_off, sexp = (int(v) for v in op[1:14])
            n = N * taps * cin
            hi = wbytes[w_off // 2:w_off // 2 + n].astype(np.float32); lo = wbytes[w_off // 2 + n:w_off // 2 + 2 * n].astype(np.float32)
            wt = torch.from_numpy(np.ldexp(hi + lo, -sexp).astype(np.float32)).view(N, taps, cin)
            ksz = 3 if taps == 9 else 1
            wt = wt.view(N, ksz, ksz, cin).permute(0, 3, 1, 2).contiguous()
            a = plane[pl].view(H, W, cin).permute(2, 0, 1)[None]
            bias = torch.from_numpy(params[b_off:b_off + N].copy()) if b_off >= 0 else None
            y = F.conv2d(a, wt, bias, padding=ksz // 2)[0].permute(1, 2, 0).reshape(H * W, N)
            dst = bufs[out][:H * W * ldc].view(H * W, ldc)
            if acc:
                dst[:, c_off:c_off + N] += y
            else:
                dst[:, c_off:c_off + N] = y
        elif k == OP_ADD:
            bufs[op[1]][:op[3]] += bufs[op[2]][:op[3]]
        elif k == OP_POOL:
            src, dst, H, W, C = (int(v) for v in op[1:6])
            v = bufs[src][:H * W * C].view(H, W, C).permute(2, 0, 1)[None]
            y = F.avg_pool2d(v, 2, stride=2)[0].permute(1, 2, 0).contiguous()
            bufs[dst][:y.numel()] = y.reshape(-1)
        elif k == OP_UPADD:
            up1, low, dst, h, w, C = (int(v) for v in op[1:7])
            lo_t = bufs[low][:h * w * C].view(h, w, C).permute(2, 0, 1)[None]
            y = F.interpolate(lo_t, scale_factor=2, mode='bicubic', align_corners=True)[0].permute(1, 2, 0).reshape(-1)
            bufs[dst][:y.numel()] = bufs[up1][:y.numel()] + y
        elif k == OP_INPUT:
            bufs[op[1]][:op[3]] = flat_in[op[2]:op[2] + op[3]]
        elif k == OP_UPSPLIT:
            src, h, w, C, ld, c_off, relu, pl = (int(v) for v in op[1:9])
            v = bufs[src][:h * w * ld].view(h, w, ld)[:, :, c_off:c_off + C].permute(2, 0, 1)[None]
            if relu:
                v = torch.relu(v)
            y = F.interpolate(v, scale_factor=2, mode='bilinear', align_corners=False)[0].permute(1, 2, 0).reshape(4 * h * w, C)
            plane[pl][:, :C] = y
        elif k == OP_COPY:
            src, dst, P, C, ld, c_off = (int(v) for v in op[1:7])
            bufs[dst][:P * ld].view(P, ld)[:, c_off:c_off + C] = bufs[src][:P * C].view(P, C)
        elif k == OP_CONV4:
            # the gather-GEMM of csrc/conv_tc.cu conv4_gemm_kernel, index formulas transcribed (NOT F.conv2d): what is pinned here is the
            # tap -> input pixel map, the output-parity classes of the transposed convolution and the K-major weight packing
            src, dst, Hin, Win, Ci, lds, cos, Co, ldd, cod, w_off, b_off, flags = (int(v) for v in op[1:14])
            tr = flags & 1
            X = (bufs[src][:Hin * Win * lds].view(Hin, Win, lds)[:, :, cos:cos + Ci] if src >= 0 else x.permute(1, 2, 0)).contiguous()
            if flags & 2:
                X = torch.relu(X)
            Ho, Wo = (2 * Hin, 2 * Win) if tr else (Hin // 2, Win // 2)
            ntaps, ncls = (4, 4) if tr else (16, 1)
            K = ntaps * Ci
            Wk = torch.from_numpy(params[w_off:w_off + ncls * K * Co].reshape(ncls, K, Co).copy())
            hq, wq = (Hin, Win) if tr else (Ho, Wo)
            qy = torch.arange(hq)[:, None].expand(hq, wq); qx = torch.arange(wq)[None, :].expand(hq, wq)
            Y = torch.zeros(Ho, Wo, Co)
            for cls in range(ncls):
                ry, rx = cls >> 1, cls & 1
                A = torch.zeros(hq, wq, ntaps, Ci)
                for tap in range(ntaps):
                    if tr:
                        iy = qy + (ry - (tap >> 1)); ix = qx + (rx - (tap & 1))
                    else:
                        iy = 2 * qy + ((tap >> 2) - 1); ix = 2 * qx + ((tap & 3) - 1)
                    ok = (iy >= 0) & (iy < Hin) & (ix >= 0) & (ix < Win)
                    A[:, :, tap] = X[iy.clamp(0, Hin - 1), ix.clamp(0, Win - 1)] * ok[:, :, None]
                y = (A.reshape(hq * wq, K) @ Wk[cls]).view(hq, wq, Co)
                if tr:
                    Y[ry::2, rx::2] = y
                else:
                    Y = y
            if b_off >= 0:
                Y = Y + torch.from_numpy(params[b_off:b_off + Co].copy())
            if flags & 4:
                Y = F.leaky_relu(Y, 0.2)
            bufs[dst][:Ho * Wo * ldd].view(Ho * Wo, ldd)[:, cod:cod + Co] = Y.reshape(Ho * Wo, Co)
        else:
            raise ValueError('unknown op %d' % k)
    return bufs[out_buf][:out_h * out_w * out_c].view(out_h, out_w, out_c).numpy()
