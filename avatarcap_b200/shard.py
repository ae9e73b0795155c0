"""Multi-GPU host logic for the dense-grid path: slab sharding along the slowest-varying volume
axis (x in the reference layout flat=(i*Ry+j)*Rz+k, main.py:364), the halo exchange marching cubes needs, and the mesh
gather. One process per GPU. The arithmetic stays in the CUDA library: this module only places planes and offsets ids.

Seam rule: a vertex belongs to the rank that owns the lower voxel of its grid edge. A slab's faces may reference
vertices of the first plane of the next slab; the kernel numbers those right after the slab's own vertices (in the
next rank's canonical order), so global id = base[rank+1] + (local id - n_owned) and no welding is needed: the merged
mesh is IDENTICAL (faces bit-exact) to the single-GPU mesh of the same volume (recon_util.recon_mesh, :51-70).

Exchange step (`SlabVolume.exchange`): every rank holds ONE padded buffer [lo halo | own planes | hi halo]; the field
kernel writes its occupancy straight into the `own` view and marching cubes reads the padded view -- no torch.cat.
  mode 'p2p'      (GPUs): the buffer is library-allocated and exported over CUDA IPC; a kernel of ours stores the boundary
                  planes into the neighbours' buffers over NVLink, epochs + acks keep it stream-ordered (csrc/shard.cu).
  mode 'sendrecv' (fallback, and the gloo CPU tests): torch.distributed isend / irecv of the planes, neighbours only,
                  received directly into the halo regions.
Mesh gather (`gather_mesh`): counts all-gather -> exclusive scan -> every rank renumbers its own faces on the device ->
payload send / recv straight into rank 0's merged buffers (count-then-payload).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

HALO_LO = 2   # planes below a slab needed by the Sobel/trilinear normal stencil (recon_util.py:9-48 incl. the +0.5 voxel quirk)
HALO_HI = 3   # planes above: 1 for the cells, +2 for vertex ids of the next rank's first plane and the normal stencil


def slab_range(rx: int, world: int, rank: int) -> Tuple[int, int]:
    """Planes [start, end) of `rank`: contiguous, sizes differ by at most one, earlier ranks take the remainder."""
    base, rem = divmod(rx, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def halo_planes(rx: int, start: int, end: int) -> Tuple[int, int]:
    """(lo, hi) halo widths actually available for the slab [start, end) of a volume with rx planes."""
    return min(HALO_LO, start), min(HALO_HI, rx - end)


def merge_meshes(parts: Sequence[Tuple[np.ndarray, np.ndarray, np.ndarray]]):
    """Host-side statement of the seam rule (used by the tests as the checker of gather_mesh): concatenate per-rank
    (verts, faces, normals) into the single-volume mesh. Face indices >= the rank's own vertex count refer to the next
    rank's first vertices (see module docstring)."""
    counts = [p[0].shape[0] for p in parts]
    base = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    faces = []
    for r, (v, f, n) in enumerate(parts):
        f = f.astype(np.int64)
        own = counts[r]
        g = np.where(f < own, f + base[r], f - own + base[min(r + 1, len(parts))])
        faces.append(g)
    verts = np.concatenate([p[0] for p in parts], 0)
    normals = np.concatenate([p[2] for p in parts], 0) if parts[0][2] is not None else None
    return verts, np.concatenate(faces, 0).astype(np.int32), normals


class _DevBuf:
    """Raw device allocation exposed through __cuda_array_interface__ so that torch can alias it without owning it."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {'shape': (int(n_floats),), 'typestr': '<f4', 'data': (int(ptr), False), 'version': 2}


class SlabVolume:
    """This rank's x-slab of a (Rx,Ry,Rz) volume with room for the halo planes, plus the exchange step.

        sv = SlabVolume(res, world, rank, engine=eng)        # collective: every rank constructs it
        eng.eval_occupancy(pts, center, out_occ=sv.own.view(-1))
        sv.exchange()                                        # push my boundary planes, wait for the neighbours'
        v, f, n, counts = eng.extract_mesh_async(sv.padded, bounds, iso, cap_v, cap_f, True, sv.lo, sv.hi, sv.x0 - sv.lo, res[0])
        sv.release()                                         # neighbours may overwrite my halo from now on
    """

    def __init__(self, res, world: int, rank: int, engine=None, device=None, group=None, mode: str = 'auto'):
        import torch
        import torch.distributed as dist
        self.res = tuple(int(r) for r in res)
        self.world, self.rank, self.engine, self.group = int(world), int(rank), engine, group
        self.x0, self.x1 = slab_range(self.res[0], world, rank)
        self.nx = self.x1 - self.x0
        self.lo, self.hi = halo_planes(self.res[0], self.x0, self.x1)
        if world > 1 and self.nx < HALO_HI:
            raise ValueError('slab of %d planes is thinner than the halo (%d)' % (self.nx, HALO_HI))
        self.plane = self.res[1] * self.res[2]
        self.device = torch.device(device) if device is not None else (engine.device if engine is not None else torch.device('cpu'))
        self.epoch = 0
        self._base = None; self._peer_lo = None; self._peer_hi = None; self._holder = None
        n_floats = (HALO_LO + self.nx + HALO_HI) * self.plane
        if mode not in ('auto', 'p2p', 'sendrecv'):
            raise ValueError('mode must be auto, p2p or sendrecv')
        want_p2p = mode != 'sendrecv' and engine is not None and self.device.type == 'cuda' and world > 1
        self.mode = 'sendrecv'
        if mode == 'p2p' and not want_p2p and world > 1:
            raise ValueError("mode 'p2p' needs an Engine on a CUDA device")
        if want_p2p:
            ok, handle = self._alloc_p2p(n_floats)
            # all ranks must take the same path: agree on it
            state = [None] * world
            dist.all_gather_object(state, (bool(ok), handle), group=group)
            if all(s[0] for s in state):
                ok2 = self._open_peers(state)
                flags = [None] * world
                dist.all_gather_object(flags, bool(ok2), group=group)
                if all(flags):
                    self.mode = 'p2p'
            if self.mode != 'p2p':
                if mode == 'p2p':
                    raise RuntimeError('CUDA IPC peer mapping is not available on this box: %s' % getattr(self, '_p2p_error', 'a peer failed'))
                self._free_p2p()
        if self.mode == 'p2p':
            self._holder = _DevBuf(self._base + engine_header_bytes(), n_floats)
            flat = torch.as_tensor(self._holder, device=self.device)
        else:
            flat = torch.empty(n_floats, device=self.device, dtype=torch.float32)
        self._flat = flat
        all_planes = flat.view(HALO_LO + self.nx + HALO_HI, self.res[1], self.res[2])
        self.own = all_planes[HALO_LO:HALO_LO + self.nx]                                    # (nx, Ry, Rz): the field kernel's output
        self.padded = all_planes[HALO_LO - self.lo:HALO_LO + self.nx + self.hi]             # [lo | own | hi]: what marching cubes reads

    # ------------------------------------------------------------------ CUDA IPC plumbing
    def _alloc_p2p(self, n_floats: int):
        from . import _lib
        e = self.engine
        base = C.c_void_p(); handle = (C.c_ubyte * _lib.IPC_HANDLE_BYTES)()
        rc = e.lib.avc_shard_alloc(e._h, n_floats * 4, C.byref(base), handle)
        if rc:
            self._p2p_error = e.lib.avc_last_error(e._h).decode()
            return False, None
        self._base = int(base.value)
        return True, bytes(handle)

    def _open_peers(self, state) -> bool:
        e = self.engine
        for attr, peer in (('_peer_lo', self.rank - 1), ('_peer_hi', self.rank + 1)):
            if peer < 0 or peer >= self.world:
                continue
            hb = (C.c_ubyte * len(state[peer][1])).from_buffer_copy(state[peer][1])
            p = C.c_void_p()
            rc = e.lib.avc_shard_open(e._h, hb, C.byref(p))
            if rc:
                self._p2p_error = e.lib.avc_last_error(e._h).decode()
                return False
            setattr(self, attr, int(p.value))
        return True

    def _free_p2p(self) -> None:
        e = self.engine
        for attr in ('_peer_lo', '_peer_hi'):
            if getattr(self, attr):
                e.lib.avc_shard_close(e._h, C.c_void_p(getattr(self, attr))); setattr(self, attr, None)
        if self._base:
            e.lib.avc_shard_free(e._h, C.c_void_p(self._base)); self._base = None

    def close(self) -> None:
        """Collective in spirit: call it on every rank once no rank can still be pushing (after a barrier)."""
        self.own = self.padded = self._flat = None
        self._holder = None
        if self.mode == 'p2p' and self.engine is not None and getattr(self.engine, '_h', None):
            self._free_p2p()

    # ------------------------------------------------------------------ the exchange step
    def _neighbour_layout(self, peer: int):
        """(nx, lo, hi) of rank `peer` -- where its halo regions start inside ITS buffer."""
        s, e = slab_range(self.res[0], self.world, peer)
        lo, hi = halo_planes(self.res[0], s, e)
        return e - s, lo, hi

    def exchange(self) -> None:
        """My first HALO_HI planes -> the lower neighbour's hi halo, my last HALO_LO planes -> the upper neighbour's lo halo;
        returns with the wait for the neighbours' planes ENQUEUED (p2p) / completed on the stream (sendrecv)."""
        if self.world == 1:
            return
        self.epoch += 1
        has_lo, has_hi = self.rank > 0, self.rank < self.world - 1
        if self.mode == 'p2p':
            e = self.engine
            lo_off = hi_off = 0
            if has_lo:
                nx_l, _, hi_l = self._neighbour_layout(self.rank - 1)
                assert hi_l == HALO_HI
                lo_off = (HALO_LO + nx_l) * self.plane                      # the lower neighbour's hi-halo region
            if has_hi:
                _, lo_h, _ = self._neighbour_layout(self.rank + 1)
                assert lo_h == HALO_LO
                hi_off = 0                                                   # the upper neighbour's lo-halo region starts its data area
            st = e._stream()
            e._check(e.lib.avc_halo_push(e._h, C.c_void_p(self._base), C.c_void_p(self._peer_lo) if has_lo else None,
                                         C.c_void_p(self._peer_hi) if has_hi else None, self.plane, HALO_LO * self.plane, self.nx,
                                         HALO_HI, lo_off, HALO_LO, hi_off, self.epoch, st))
            e._check(e.lib.avc_halo_wait(e._h, C.c_void_p(self._base), int(has_lo), int(has_hi), self.epoch, st))
            return
        import torch.distributed as dist
        planes = self._flat.view(HALO_LO + self.nx + HALO_HI, self.res[1], self.res[2])
        ops = []
        if has_lo:
            ops.append(dist.P2POp(dist.isend, planes[HALO_LO:HALO_LO + HALO_HI], self.rank - 1, group=self.group))
            ops.append(dist.P2POp(dist.irecv, planes[:HALO_LO], self.rank - 1, group=self.group))
        if has_hi:
            ops.append(dist.P2POp(dist.isend, planes[self.nx:self.nx + HALO_LO], self.rank + 1, group=self.group))
            ops.append(dist.P2POp(dist.irecv, planes[HALO_LO + self.nx:], self.rank + 1, group=self.group))
        for r in dist.batch_isend_irecv(ops):
            r.wait()

    def release(self) -> None:
        """Enqueue after the kernels that read the halo: the neighbours may push the next epoch's planes."""
        if self.world == 1 or self.mode != 'p2p':
            return
        e = self.engine
        has_lo, has_hi = self.rank > 0, self.rank < self.world - 1
        e._check(e.lib.avc_halo_ack(e._h, C.c_void_p(self._peer_lo) if has_lo else None, C.c_void_p(self._peer_hi) if has_hi else None,
                                    self.epoch, e._stream()))


def engine_header_bytes() -> int:
    from . import _lib
    return _lib.SHARD_HEADER_BYTES


def exchange_halo(slab, rank: int, world: int, rx: int, group=None):
    """Functional form for callers that hold a plain (nx,Ry,Rz) slab tensor: returns [lo halo | slab | hi halo].
    Neighbour send / recv straight into a padded buffer (one copy of the slab; SlabVolume avoids that one too)."""
    if world == 1:
        return slab
    sv = SlabVolume((rx, slab.shape[1], slab.shape[2]), world, rank, device=slab.device, group=group, mode='sendrecv')
    if sv.nx != slab.shape[0]:
        raise ValueError('slab has %d planes, rank %d of %d owns %d' % (slab.shape[0], rank, world, sv.nx))
    sv.own.copy_(slab)
    sv.exchange()
    return sv.padded


def _all_counts(mine, world: int, group=None):
    """all-gather of a small int64 device tensor -> (world, k) numpy on the host (ONE synchronisation)."""
    import torch
    import torch.distributed as dist
    if world > 1:
        flat = mine.contiguous().view(-1)
        allc = torch.empty(world * flat.numel(), device=mine.device, dtype=torch.int64)
        dist.all_gather_into_tensor(allc, flat, group=group)
        allc = allc.view((world,) + tuple(mine.shape))
    else:
        allc = mine[None]
    return allc.cpu().numpy()


def gather_mesh(verts, faces, normals, n_verts: int, n_faces: int, rank: int, world: int, engine=None, group=None, dst: int = 0,
                counts=None):
    """Count-then-payload gather of the per-slab meshes into the single mesh the caller of recon_mesh expects
    (utils/recon_util.py:51-70, main.py:367), on the device.
      verts (>=n_verts,3) f32, faces (>=n_faces,3) i32 with LOCAL ids (see module docstring), normals like verts or None.
      counts: (world,2) host array of every rank's (n_verts, n_faces) when the caller already all-gathered them.
    -> on rank `dst`: (verts (V,3), faces (F,3), normals | None, per_rank_counts (world,2)); elsewhere (None, None, None, counts).
    One all-gather of the counts (the only host synchronisation), then every rank offsets its own faces with a kernel and the
    payloads travel point-to-point straight into their final position in rank dst's buffers."""
    import torch
    import torch.distributed as dist
    dev = verts.device
    if counts is None:
        counts = _all_counts(torch.tensor([int(n_verts), int(n_faces)], device=dev, dtype=torch.int64), world, group)
    counts = np.asarray(counts)[:, :2]
    vbase = np.concatenate([[0], np.cumsum(counts[:, 0])]).astype(np.int64)
    fbase = np.concatenate([[0], np.cumsum(counts[:, 1])]).astype(np.int64)
    if vbase[-1] > 0x7fffffff:
        raise ValueError('merged mesh has %d vertices: too many for int32 face indices' % vbase[-1])
    f_own = faces[:n_faces]
    if n_faces:
        if engine is not None and f_own.is_cuda:
            engine.renumber_faces(f_own, int(n_verts), int(vbase[rank]), int(vbase[min(rank + 1, world)]))
        else:           # host tensors (gloo tests of the plumbing): the same rule in torch
            f_own.copy_(torch.where(f_own < n_verts, f_own + int(vbase[rank]), f_own - int(n_verts) + int(vbase[min(rank + 1, world)])))
    if world == 1:
        return verts[:n_verts], f_own, (normals[:n_verts] if normals is not None else None), counts
    ops = []
    out_v = out_f = out_n = None
    if rank == dst:
        V, F = int(vbase[-1]), int(fbase[-1])
        out_v = torch.empty((V, 3), device=dev, dtype=torch.float32)
        out_f = torch.empty((F, 3), device=dev, dtype=torch.int32)
        out_n = torch.empty((V, 3), device=dev, dtype=torch.float32) if normals is not None else None
        out_v[vbase[rank]:vbase[rank + 1]].copy_(verts[:n_verts]); out_f[fbase[rank]:fbase[rank + 1]].copy_(f_own)
        if out_n is not None:
            out_n[vbase[rank]:vbase[rank + 1]].copy_(normals[:n_verts])
        for r in range(world):
            if r == dst:
                continue
            if counts[r, 0]:
                ops.append(dist.P2POp(dist.irecv, out_v[vbase[r]:vbase[r + 1]], r, group=group))
                if out_n is not None:
                    ops.append(dist.P2POp(dist.irecv, out_n[vbase[r]:vbase[r + 1]], r, group=group))
            if counts[r, 1]:
                ops.append(dist.P2POp(dist.irecv, out_f[fbase[r]:fbase[r + 1]], r, group=group))
    else:
        if n_verts:
            ops.append(dist.P2POp(dist.isend, verts[:n_verts], dst, group=group))
            if normals is not None:
                ops.append(dist.P2POp(dist.isend, normals[:n_verts], dst, group=group))
        if n_faces:
            ops.append(dist.P2POp(dist.isend, f_own, dst, group=group))
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    return out_v, out_f, out_n, counts


def extract_sharded_mesh(engine, sv: SlabVolume, bounds, iso: float, with_normals: bool = True, group=None, dst: int = 0,
                         cap_hint: Optional[Tuple[int, int]] = None):
    """exchange -> per-slab marching cubes (capacity-bounded, asynchronous) -> release -> gather on `dst`.
    Returns gather_mesh's tuple. The field values must already be in sv.own (same stream). ONE host synchronisation: the
    all-gathered device-side counts (which also carry every rank's overflow flag)."""
    sv.exchange()
    nvox = sv.padded.numel()
    cap_v, cap_f = cap_hint if cap_hint else (max(4096, nvox // 16), max(8192, nvox // 8))
    args = (with_normals, sv.lo, sv.hi, sv.x0 - sv.lo, sv.res[0])
    v, f, n, counts = engine.extract_mesh_async(sv.padded, bounds, iso, cap_v, cap_f, *args)
    allc = _all_counts(counts, sv.world, group)                     # (world, 4): n_verts, n_faces, n_scan, overflow
    if allc[:, 3].any():                                            # some rank overflowed: it re-runs with exact sizes (the halo is still in place)
        if allc[sv.rank, 3]:
            v, f, n, counts = engine.extract_mesh_async(sv.padded, bounds, iso, max(int(allc[sv.rank, 0]), 1), max(int(allc[sv.rank, 1]), 1), *args)
        allc = _all_counts(counts, sv.world, group)
    sv.release()
    return gather_mesh(v, f, n, int(allc[sv.rank, 0]), int(allc[sv.rank, 1]), sv.rank, sv.world, engine, group, dst, counts=allc)
