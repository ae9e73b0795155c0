"""CPU: host-side logic -- packer layout, ABI surface, slab sharding / halo exchange (gloo, world_size 2), MC tables."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from avatarcap_b200 import packer, synth, shard, mc_tables
from helpers import ROOT


def h_layers(blob):
    return packer.parse_header(blob)['layers']


def test_packer_layout_and_folding():
    sd = synth.avatar_state_dict()
    blob = packer.pack_avatar(sd)
    h = packer.parse_header(blob)
    assert h['magic'] == packer.MAGIC and h['n_layers'] == 20 and h['kind'] == packer.KIND_AVATAR
    f32 = np.frombuffer(blob, np.float32, h['f32_bytes'] // 4, h['f32_off'])
    # layer 4 = conv5: concat order input(67) FIRST then x4(256)  (mlp.py:106)
    L = h['layers'][4]
    assert (L['k0'], L['k1'], L['k0p'], L['k1p'], L['n']) == (67, 256, 80, 256, 256)
    W = sd['warping_field.mlp.conv5.weight'][:, :, 0]
    Wt = f32[L['wt_off']:L['wt_off'] + 323 * 256].reshape(323, 256)
    assert np.array_equal(Wt, W.T)
    # BN fold: scale*(Wx) + bias == BN(Wx + b)
    x = np.random.RandomState(0).normal(0, 1, 323).astype(np.float64)
    y_ref = (W.astype(np.float64) @ x + sd['warping_field.mlp.conv5.bias'] - sd['warping_field.mlp.bn5.running_mean']) / \
        np.sqrt(sd['warping_field.mlp.bn5.running_var'].astype(np.float64) + 1e-5) * sd['warping_field.mlp.bn5.weight'] + sd['warping_field.mlp.bn5.bias']
    sc = f32[L['sb_off']:L['sb_off'] + 256]; bi = f32[L['sb_off'] + 256:L['sb_off'] + 512]
    assert np.abs((W.astype(np.float64) @ x) * sc + bi - y_ref).max() < 1e-5
    # layer 12 = shared fc4: activations(256) first, then PE(63)  (mlp.py:60-61)
    L = h['layers'][12]
    assert (L['k0'], L['k1'], L['k0p'], L['k1p']) == (256, 63, 256, 64)
    # fp16 hi/lo slabs reconstruct W * 2^shift to ~2^-21 relative; canonical core-matrix order
    f16 = np.frombuffer(blob, np.float16, h['f16_bytes'] // 2, h['f16_off'])
    L = h['layers'][1]
    W = sd['warping_field.mlp.conv2.weight'][:, :, 0]
    # stream order (packer.tc_pieces): half h (128 rows) -> k-step ks -> hi slab (128 x 16) then lo slab
    rows, ks, h = 128, 3, 1
    off = L['tc_w_off'] // 2 + h * (16 * rows * 32) + ks * (rows * 32)
    hi = f16[off:off + rows * 16].reshape(rows // 8, 2, 8, 8).transpose(0, 2, 1, 3).reshape(rows, 16).astype(np.float64)
    lo = f16[off + rows * 16:off + 2 * rows * 16].reshape(rows // 8, 2, 8, 8).transpose(0, 2, 1, 3).reshape(rows, 16).astype(np.float64)
    ref = W[h * rows:(h + 1) * rows, 16 * ks:16 * ks + 16].astype(np.float64) * 2.0 ** L['shift']
    assert np.abs(hi + lo - ref).max() <= np.abs(ref).max() * 2.0 ** -20
    # every layer's stream has exactly np * (k0p + k1p) * 4 bytes (hi + lo) and the streams tile the f16 section
    pos = 0
    for Ld in h_layers(blob):
        assert Ld['tc_w_off'] == pos
        pos += Ld['np'] * (Ld['k0p'] + Ld['k1p']) * 4
    assert pos == packer.parse_header(blob)['f16_bytes']
    tsc = f32[L['tc_sb_off']:L['tc_sb_off'] + 256]
    assert np.allclose(tsc * 2.0 ** L['shift'], f32[L['sb_off']:L['sb_off'] + 256], rtol=1e-7)


def test_packer_recon_weight_norm():
    sd = synth.recon_state_dict()
    h = packer.parse_header(packer.pack_recon(sd))
    assert [l['n'] for l in h['layers']] == [512, 256, 128, 1]
    assert [(l['k0'], l['k1']) for l in h['layers']] == [(33, 0), (512, 33), (256, 33), (128, 0)]
    L = packer.recon_layers(sd)[1]
    v = sd['image_decoder.fc_list.1.0.weight_v'][:, :, 0].astype(np.float64); g = sd['image_decoder.fc_list.1.0.weight_g'][:, 0, 0]
    W = v * (g / np.sqrt((v * v).sum(1)))[:, None]
    assert np.abs(L.W * L.scale[:, None] - W).max() < 1e-6


def test_packer_rejects_other_architectures():
    sd = synth.avatar_state_dict()
    sd['cano_template.shared_mlp.fc_list.0.0.weight'] = np.zeros((256, 39, 1), np.float32)      # pos_encoding 6
    with pytest.raises(ValueError):
        packer.pack_avatar(sd)


def test_split_precision_scheme_meets_tolerance():
    """Emulates the tensor-core arithmetic on the CPU: fp16 hi/lo operands, 3 products (hi*hi + hi*lo + lo*hi), wide accumulate,
    activations re-split every layer. The scheme must keep the occupancy well inside the 1e-4 budget of the reference."""
    from oracle import field_oracle as fo
    sd = synth.avatar_state_dict()
    layers = packer.avatar_layers(sd)
    rs = np.random.RandomState(1)
    pts = rs.uniform(-0.8, 0.8, (4000, 3)).astype(np.float32)

    def split(x):
        hi = x.astype(np.float16).astype(np.float64)
        lo = (x - hi).astype(np.float16).astype(np.float64)
        return hi, lo

    def lin(L, x):      # x (K,N) float32
        wh, wl = split(L.W.astype(np.float32)); xh, xl = split(x.astype(np.float32))
        acc = wh @ xh + wh @ xl + wl @ xh
        return (acc * L.scale[:, None].astype(np.float64) + L.bias[:, None].astype(np.float64)).astype(np.float32)

    e = fo.embed(torch.from_numpy(pts), 10).numpy().T.astype(np.float32)
    x = e
    for i, L in enumerate(layers[8:15]):
        inp = np.concatenate([x, e], 0) if i == 4 else x
        x = lin(L, inp)
        if i < 6:
            x = np.maximum(x, 0)
    g = lin(layers[15], x); g = np.where(g > 0, g, g * np.float32(0.02))
    occ = lin(layers[16], g)[0]
    _, _, ref = fo.template_forward(sd, torch.from_numpy(pts).double())
    err = np.abs(occ - ref.numpy()[:, 0]).max()
    assert err < 3e-5, err


def test_abi_header_matches_binding():
    """every function the header declares is bound by _lib.SIGNATURES (and vice versa) and exported by the built library"""
    from avatarcap_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'avatarcap_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(avc_[a-z0-9_]+)\s*\(', hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip('library not built (python -c "import __graft_entry__ as g; g.build()")')
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.avc_abi_version() == 5


def test_header_is_plain_c99_and_the_c_client_compiles():
    """the drop-in boundary is a C ABI: include/avatarcap_b200.h and the C client tests/abi_smoke.c must compile as C99 (no C++ leaks)"""
    import shutil
    import subprocess
    if shutil.which('gcc') is None:
        pytest.skip('no gcc')
    cuda_inc = os.path.join(os.environ.get('CUDA_HOME', '/usr/local/cuda'), 'include')
    if not os.path.exists(os.path.join(cuda_inc, 'cuda_runtime_api.h')):
        pytest.skip('no CUDA headers')
    r = subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-pedantic', '-fsyntax-only', '-I', os.path.join(ROOT, 'include'), '-isystem', cuda_inc,
                        os.path.join(ROOT, 'tests', 'abi_smoke.c')], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_engine_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip('has a GPU')
    from avatarcap_b200.engine import Engine
    with pytest.raises((RuntimeError, ImportError)):
        Engine()


def test_mc_tables_consistency():
    assert mc_tables.MAX_TRI == 5 and int(mc_tables.NTRI.sum()) == 820
    for c in range(256):
        cut = 0
        for e, (a, b) in enumerate(mc_tables.EDGES):
            if ((c >> a) & 1) != ((c >> b) & 1):
                cut |= 1 << e
        assert mc_tables.EDGE_MASK[c] == cut
        assert mc_tables.EDGE_MASK[c] == mc_tables.EDGE_MASK[255 - c]     # complementary case cuts the same edges
    inc = open(os.path.join(ROOT, 'avatarcap_b200', 'csrc', 'mc_tables.inc')).read()
    row = ','.join(str(int(x)) for x in list(mc_tables.TRI[37]) + [-1])   # rows are padded to 16 bytes
    assert '{%s}' % row in inc                                             # the committed CUDA include is up to date
    nib = sum(int(a) << (4 * e) for e, a in enumerate(np.array(mc_tables.EDGES)[:, 0]))
    assert '0x%xull' % nib in inc and [int(x) for x in mc_tables.EDGE_AXIS] == [e >> 2 for e in range(12)]


def test_slab_ranges_and_merge():
    for rx in (5, 37, 256, 512):
        for world in (1, 2, 3, 8):
            if rx < world:
                continue
            r = [shard.slab_range(rx, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == rx and all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(e - s for s, e in r) - min(e - s for s, e in r) <= 1
    assert shard.halo_planes(512, 0, 64) == (0, 3) and shard.halo_planes(512, 448, 512) == (2, 0) and shard.halo_planes(512, 64, 128) == (2, 3)
    # merge: indices past a rank's own vertices point into the next rank
    a = (np.zeros((3, 3), np.float32), np.array([[0, 1, 2], [2, 3, 4]], np.int32), np.zeros((3, 3), np.float32))
    b = (np.ones((2, 3), np.float32), np.array([[0, 1, 1]], np.int32), np.ones((2, 3), np.float32))
    v, f, n = shard.merge_meshes([a, b])
    assert v.shape == (5, 3) and f.tolist() == [[0, 1, 2], [2, 3, 4], [3, 4, 4]]


_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
from avatarcap_b200 import shard
from oracle import mesh_oracle as mo
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%%s' %% os.environ['MASTER_PORT'], rank=rank, world_size=world)
rs = np.random.RandomState(0); res = (13, 6, 5)
s, e = shard.slab_range(res[0], world, rank)
lo, hi = shard.halo_planes(res[0], s, e)
# 1. functional form
vol = torch.from_numpy(rs.normal(0, 1, res).astype(np.float32))
out = shard.exchange_halo(vol[s:e].contiguous(), rank, world, res[0])
assert torch.equal(out, vol[s - lo:e + hi]), (rank, out.shape)
# 2. SlabVolume: the field output lands in `own`, the neighbours' planes in the halo regions, no concatenation; several epochs
sv = shard.SlabVolume(res, world, rank, device='cpu')
assert sv.mode == 'sendrecv' and (sv.x0, sv.x1, sv.lo, sv.hi) == (s, e, lo, hi)
assert sv.padded.data_ptr() + 4 * sv.lo * sv.plane == sv.own.data_ptr()            # one buffer, two views
for epoch in range(3):
    vol = torch.from_numpy(rs.normal(0, 1, res).astype(np.float32))
    sv.own.copy_(vol[s:e])
    sv.exchange()
    assert torch.equal(sv.padded, vol[s - lo:e + hi]), (rank, epoch)
    sv.release()
# 3. count-then-payload mesh gather: slab meshes cut out of the oracle's whole mesh by vertex / face ownership, numbered the way
#    the kernel numbers them (own vertices first, then the next slab's first-plane vertices)
vnp = vol.numpy()
v, f, vox, axis, cells = mo.marching_cubes(vnp, 0.0, return_owner=True)
n = np.random.RandomState(1).normal(0, 1, v.shape).astype(np.float32)
plane = res[1] * res[2]
vplane = vox // plane; fplane = cells // plane
ranges = [shard.slab_range(res[0], world, r) for r in range(world)]
vb = [int((vplane < a).sum()) for a, _ in ranges] + [len(v)]
mine_v = slice(vb[rank], vb[rank + 1])
fsel = (fplane >= s) & (fplane < e)
fl = f[fsel].astype(np.int64)
n_own = vb[rank + 1] - vb[rank]
local = np.where(fl < vb[rank + 1], fl - vb[rank], n_own + (fl - vb[rank + 1]))
assert local.min(initial=0) >= 0
pad = 7                                                                       # capacity buffers are larger than the counts
tv = torch.zeros((n_own + pad, 3)); tv[:n_own] = torch.from_numpy(v[mine_v])
tn = torch.zeros((n_own + pad, 3)); tn[:n_own] = torch.from_numpy(n[mine_v])
tf = torch.zeros((len(local) + pad, 3), dtype=torch.int32); tf[:len(local)] = torch.from_numpy(local.astype(np.int32))
gv, gf, gn, counts = shard.gather_mesh(tv, tf, tn, n_own, len(local), rank, world)
assert counts.shape == (world, 2) and int(counts[:, 0].sum()) == len(v) and int(counts[:, 1].sum()) == len(f)
if rank == 0:
    assert torch.equal(gv, torch.from_numpy(v)) and torch.equal(gn, torch.from_numpy(n))
    assert np.array_equal(gf.numpy(), f)
else:
    assert gv is None and gf is None
dist.barrier(); dist.destroy_process_group()
print('ok', rank)
'''


import pytest


@pytest.mark.parametrize('world,port', [(2, 29631), (3, 29641)])
def test_halo_exchange_and_mesh_gather_gloo(tmp_path, world, port):
    script = tmp_path / 'w.py'
    script.write_text(_WORKER % ROOT)
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_PORT=str(port), MASTER_ADDR='127.0.0.1')
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_frames_round_robin_partition():
    """config[5] frame-parallel replicas: every frame on exactly one rank, loads differ by at most one frame."""
    from avatarcap_b200 import pipeline
    for n_frames, world in ((16, 8), (16, 3), (5, 8), (0, 2), (7, 1)):
        parts = [pipeline.frames_for_rank(n_frames, world, r) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(n_frames))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
        assert all(f % world == r for r, p in enumerate(parts) for f in p)
    import pytest
    with pytest.raises(ValueError):
        pipeline.frames_for_rank(4, 2, 2)


def test_knn_shell_walk_covers_every_cell_exactly_once():
    """Model of the cell traversal of lbs.cu `knn_grid` (3x3x3 cube as z-runs, then shells r >= 2 as side-face runs + the two end cells of
    the interior rows): after the shells 0..R every cell of the clamped Chebyshev ball of radius R has been visited exactly once, for
    any grid shape and query cell -- the exactness argument of the search rests on that."""
    def walk(dims, q, R):
        dx, dy, dz = dims; cx, cy, cz = q
        seen = []
        def run(i, j, k0, k1):
            seen.extend((i, j, k) for k in range(k0, k1 + 1))
        for i in range(max(cx - 1, 0), min(cx + 1, dx - 1) + 1):
            for j in range(max(cy - 1, 0), min(cy + 1, dy - 1) + 1):
                run(i, j, max(cz - 1, 0), min(cz + 1, dz - 1))
        for r in range(2, R + 1):
            for i in range(max(cx - r, 0), min(cx + r, dx - 1) + 1):
                for j in range(max(cy - r, 0), min(cy + r, dy - 1) + 1):
                    if abs(i - cx) == r or abs(j - cy) == r:
                        run(i, j, max(cz - r, 0), min(cz + r, dz - 1))
                    else:
                        if cz - r >= 0:
                            run(i, j, cz - r, cz - r)
                        if cz + r <= dz - 1:
                            run(i, j, cz + r, cz + r)
        return seen
    rs = np.random.RandomState(0)
    for _ in range(300):
        dims = tuple(int(x) for x in rs.randint(1, 9, 3))
        q = tuple(int(rs.randint(0, d)) for d in dims)
        R = int(rs.randint(1, 9))
        seen = walk(dims, q, R)
        ball = {(i, j, k) for i in range(dims[0]) for j in range(dims[1]) for k in range(dims[2])
                if max(abs(i - q[0]), abs(j - q[1]), abs(k - q[2])) <= R}
        assert len(seen) == len(set(seen)), (dims, q, R)
        assert set(seen) == ball, (dims, q, R)
