"""Precision probe for the round-2 candidate: HGFilter with every convolution evaluated as a 3-pass fp16 hi/lo split product
(fp32 accumulate), against an f64 evaluation of the same network; plain f32 and single-pass fp16 / bf16 / tf32 for comparison."""
import sys, time
import numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from avatarcap_b200 import synth, encoders
torch.set_num_threads(8)
sd = synth.hgfilter_state_dict()
x = torch.from_numpy(synth.normal_maps())           # (1,6,512,512)

def rnd(t, kind):
    if kind == 'fp16': return t.to(torch.float16).to(t.dtype)
    if kind == 'bf16': return t.to(torch.bfloat16).to(t.dtype)
    if kind == 'tf32':
        i = t.to(torch.float32).view(torch.int32); i = (i + 0x1000) & ~0x1FFF
        return i.view(torch.float32).to(t.dtype)
    return t

class Probe(encoders.ImageFeatureEncoder):
    def __init__(self, mode, dtype):
        self.mode = mode; self.dtype = dtype
        super().__init__({k: (v.astype(np.float64) if dtype == torch.float64 else v) for k, v in sd.items()}, device='cpu', use_graph=False)
        self.p = {k: v.to(dtype) for k, v in self.p.items()}

def conv(mode, xin, w, b=None, **kw):
    if mode in ('f64', 'f32'):
        return torch.conv2d(xin, w, b, **kw)
    # operands rounded in the probe's precision; products and sums in f64 (an fp32 accumulator adds ~1e-7 relative, measured by the f32 row)
    xd, wd = xin.double(), w.double()
    if mode.startswith('split'):
        s = 2.0 ** np.floor(-np.log2(float(wd.abs().max()) + 1e-30))          # per-layer power of two: weights near 1 -> lo stays normal
        ws = wd * s
        wh = rnd(ws, 'fp16'); wl = rnd(ws - wh, 'fp16')
        xh = rnd(xd, 'fp16'); xl = rnd(xd - xh, 'fp16')
        y = (torch.conv2d(xh, wh, None, **kw) + torch.conv2d(xl, wh, None, **kw) + torch.conv2d(xh, wl, None, **kw)) / s
    else:
        y = torch.conv2d(rnd(xd, mode), rnd(wd, mode), None, **kw)
    if b is not None: y = y + b.double().view(1, -1, 1, 1)
    return y.to(xin.dtype)

def run(mode):
    dtype = torch.float64 if mode == 'f64' else torch.float32
    enc = Probe(mode, dtype)
    orig = F.conv2d
    def patched(xin, w, b=None, stride=1, padding=0, dilation=1, groups=1):
        return conv(mode, xin, w, b, stride=stride, padding=padding)
    encoders.F.conv2d = patched if mode not in ('f64', 'f32') else orig
    try:
        t = time.time(); out = enc._forward(x.to(dtype)); dt = time.time() - t
    finally:
        encoders.F.conv2d = orig
    return out.double(), dt

ref, dt = run('f64'); print('f64 reference: %.1f s, output range %.3g .. %.3g, rms %.3g' % (dt, ref.min(), ref.max(), ref.pow(2).mean().sqrt()))
for mode in ('f32', 'split', 'tf32', 'fp16', 'bf16'):
    out, dt = run(mode)
    e = (out - ref).abs()
    print('%-6s max-abs %.3g  rms %.3g   (%.1f s)' % (mode, e.max(), e.pow(2).mean().sqrt(), dt), flush=True)
