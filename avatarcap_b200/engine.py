"""Thin torch-tensor front end over the C ABI: one `Engine` per CUDA device.

PyTorch is used for device memory, streams and (in shard.py) torch.distributed only; all arithmetic of the hot path
runs in libavatarcap_b200.so. Every method enqueues on the caller's current CUDA stream
(`torch.cuda.current_stream()`), as the reference's single-stream code expects.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib, packer
from ._lib import AvcError, IMPL_AUTO, IMPL_SIMT, IMPL_TC, IF_SDF, IF_OCCUPANCY, MAP_POSE, MAP_IMAGE

_IMPL = {'auto': IMPL_AUTO, 'simt': IMPL_SIMT, 'tc': IMPL_TC, 'tc2': _lib.IMPL_TC2}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class Engine:
    def __init__(self, device: Optional[torch.device] = None, impl: str = 'auto'):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError('avatarcap_b200 needs a CUDA device (B200); there is no CPU fallback')
        dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.index is None:
            dev = torch.device('cuda', torch.cuda.current_device())
        self.device = dev
        self.impl = impl
        h = C.c_void_p()
        rc = self.lib.avc_ctx_create(dev.index, C.byref(h))
        if rc:
            raise AvcError(rc, self.lib.avc_last_error(None).decode())
        self._h = h
        self._keep: Dict[str, object] = {}

    # ------------------------------------------------------------------ plumbing
    def close(self) -> None:
        if getattr(self, '_h', None):
            self.lib.avc_ctx_destroy(self._h); self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int) -> None:
        if rc:
            msg = self.lib.avc_last_error(self._h).decode()
            if rc == _lib.EVALUE:
                raise ValueError(msg)            # same exception class as the reference (config.py:22, skimage)
            raise AvcError(rc, msg)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _f32(self, t, shape_last: Optional[int] = None) -> torch.Tensor:
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(np.asarray(t))
        t = t.to(device=self.device, dtype=torch.float32).contiguous()
        if shape_last is not None and t.shape[-1] != shape_last:
            raise ValueError('expected last dimension %d, got shape %s' % (shape_last, tuple(t.shape)))
        return t

    def _impl(self, impl: Optional[str]) -> int:
        return _IMPL[impl or self.impl]

    @property
    def has_tensor_core_path(self) -> bool:
        return bool(self.lib.avc_has_tensor_core_path(self._h))

    @property
    def launch_count(self) -> int:
        return int(self.lib.avc_launch_count(self._h))

    def reset_launch_count(self) -> None:
        self.lib.avc_reset_launch_count(self._h)

    # ------------------------------------------------------------------ weights / feature maps
    def load_avatar(self, state_dict, slot: Optional[int] = None) -> None:
        """Pack + upload the per-point avatar layers. `slot` (0..WEIGHT_SLOTS-1) keeps several networks resident: the
        reference's test loop alternates `network` / `network_finetuned` every frame (main.py:307-315) and
        select_avatar() then switches between them with a pointer swap. None = the active slot."""
        blob = packer.pack_avatar(state_dict)
        if slot is None:
            self._check(self.lib.avc_load_avatar_weights(self._h, blob, len(blob)))
        else:
            self._check(self.lib.avc_load_weights_slot(self._h, _lib.KIND_AVATAR, int(slot), blob, len(blob)))

    def load_recon(self, state_dict, slot: Optional[int] = None) -> None:
        blob = packer.pack_recon(state_dict)
        if slot is None:
            self._check(self.lib.avc_load_recon_weights(self._h, blob, len(blob)))
        else:
            self._check(self.lib.avc_load_weights_slot(self._h, _lib.KIND_RECON, int(slot), blob, len(blob)))

    def select_avatar(self, slot: int) -> None:
        self._check(self.lib.avc_select_weights(self._h, _lib.KIND_AVATAR, int(slot)))

    def select_recon(self, slot: int) -> None:
        self._check(self.lib.avc_select_weights(self._h, _lib.KIND_RECON, int(slot)))

    def set_feature_map(self, which: int, fmap) -> None:
        """(1,C,H,W) or (C,H,W). A channels_last (1,C,H,W) device tensor -- what encoders.py produces -- already has the
        (H,W,C) memory order the gather kernels read and is handed over without the transpose."""
        if (isinstance(fmap, torch.Tensor) and fmap.dim() == 4 and fmap.shape[0] == 1 and fmap.dtype == torch.float32
                and fmap.device == self.device and fmap.shape[1] > 1 and not fmap.is_contiguous()
                and fmap.is_contiguous(memory_format=torch.channels_last)):
            _, Cc, H, W = fmap.shape
            self._check(self.lib.avc_set_feature_map_hwc(self._h, which, _ptr(fmap), Cc, H, W, self._stream()))
            return
        t = self._f32(fmap)
        if t.dim() == 4:
            if t.shape[0] != 1:
                raise ValueError('feature map batch must be 1 per call')
            t = t[0]
        Cc, H, W = t.shape
        self._check(self.lib.avc_set_feature_map(self._h, which, _ptr(t), Cc, H, W, self._stream()))

    def set_pose_feature_map(self, fmap) -> None:
        """WarpingField.pose_feat_map (1,64,H,W) produced by precompute_conv (arch_avatar.py:109-111)."""
        self.set_feature_map(MAP_POSE, fmap)

    def set_image_feature_map(self, fmap) -> None:
        """HGFilter output (1,32,H,W) (arch_recon.py:51-52)."""
        self.set_feature_map(MAP_IMAGE, fmap)

    # ------------------------------------------------------------------ field evaluation (device tensors)
    def eval_occupancy(self, pts, center, want_offsets: bool = True, want_texture: bool = False, if_type: str = 'sdf',
                       impl: Optional[str] = None, out_occ: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """OccupancyNet.query body for one batch element: pts (N,3), center (3,) -> occ (N,), off (N,3)[, rgb (N,3), alpha (N,)].
        out_occ: write the occupancy into this contiguous (N,) device tensor (e.g. a slab of a larger volume) instead of a new one."""
        if if_type not in ('sdf', 'occupancy'):
            raise ValueError('Invalid config.if_type!')
        p = self._f32(pts, 3)
        n = p.shape[0]
        if out_occ is not None:
            if out_occ.device != self.device or out_occ.dtype != torch.float32 or out_occ.numel() != n or not out_occ.is_contiguous():
                raise ValueError('out_occ must be a contiguous float32 device tensor with one element per point')
            occ = out_occ.view(-1)
        else:
            occ = torch.empty(n, device=self.device, dtype=torch.float32)
        off = torch.empty((n, 3), device=self.device, dtype=torch.float32) if want_offsets else None
        rgb = torch.empty((n, 3), device=self.device, dtype=torch.float32) if want_texture else None
        alpha = torch.empty(n, device=self.device, dtype=torch.float32) if want_texture else None
        c = _lib.f3(center.detach().cpu().tolist() if isinstance(center, torch.Tensor) else center)
        self._check(self.lib.avc_eval_occupancy(self._h, _ptr(p), n, c, _ptr(occ), _ptr(off), _ptr(rgb), _ptr(alpha),
                                                IF_SDF if if_type == 'sdf' else IF_OCCUPANCY, self._impl(impl), self._stream()))
        out = {'occ': occ}
        if off is not None:
            out['off'] = off
        if want_texture:
            out['rgb'] = rgb; out['alpha'] = alpha
        return out

    def eval_occupancy_grid(self, bounds, res, center, x_first: int = 0, x_count: Optional[int] = None, want_offsets: bool = True,
                            want_texture: bool = False, if_type: str = 'sdf', impl: Optional[str] = None,
                            out_occ: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """OccupancyNet.query over the dense grid of generate_volume_points (avatarcap_dataset.py:312-326), i-planes
        [x_first, x_first + x_count): coordinates come from the point index inside the kernel -- no (N,3) point list exists.
        Bit-identical to eval_occupancy(make_grid(bounds, res, x_first, x_count), ...)."""
        if if_type not in ('sdf', 'occupancy'):
            raise ValueError('Invalid config.if_type!')
        res = tuple(int(r) for r in res)
        x_count = res[0] - x_first if x_count is None else int(x_count)
        n = x_count * res[1] * res[2]
        if out_occ is not None:
            if out_occ.device != self.device or out_occ.dtype != torch.float32 or out_occ.numel() != n or not out_occ.is_contiguous():
                raise ValueError('out_occ must be a contiguous float32 device tensor with one element per grid point')
            occ = out_occ.view(-1)
        else:
            occ = torch.empty(n, device=self.device, dtype=torch.float32)
        off = torch.empty((n, 3), device=self.device, dtype=torch.float32) if want_offsets else None
        rgb = torch.empty((n, 3), device=self.device, dtype=torch.float32) if want_texture else None
        alpha = torch.empty(n, device=self.device, dtype=torch.float32) if want_texture else None
        c = _lib.f3(center.detach().cpu().tolist() if isinstance(center, torch.Tensor) else center)
        b = np.asarray(bounds.detach().cpu().numpy() if isinstance(bounds, torch.Tensor) else bounds, dtype=np.float32).reshape(6)
        self._check(self.lib.avc_eval_occupancy_grid(self._h, _lib.f6(b), _lib.i3(res), int(x_first), x_count, c, _ptr(occ), _ptr(off), _ptr(rgb),
                                                     _ptr(alpha), IF_SDF if if_type == 'sdf' else IF_OCCUPANCY, self._impl(impl), self._stream()))
        out = {'occ': occ}
        if off is not None:
            out['off'] = off
        if want_texture:
            out['rgb'] = rgb; out['alpha'] = alpha
        return out

    def eval_recon_grid(self, bounds, res, center, x_first: int = 0, x_count: Optional[int] = None, impl: Optional[str] = None) -> torch.Tensor:
        """ReconNetwork.infer's decoder over the dense grid (main.py:438-440), coordinates from the index."""
        res = tuple(int(r) for r in res)
        x_count = res[0] - x_first if x_count is None else int(x_count)
        ov = torch.empty(x_count * res[1] * res[2], device=self.device, dtype=torch.float32)
        c = _lib.f3(center.detach().cpu().tolist() if isinstance(center, torch.Tensor) else center)
        b = np.asarray(bounds.detach().cpu().numpy() if isinstance(bounds, torch.Tensor) else bounds, dtype=np.float32).reshape(6)
        self._check(self.lib.avc_eval_recon_grid(self._h, _lib.f6(b), _lib.i3(res), int(x_first), x_count, c, _ptr(ov), self._impl(impl), self._stream()))
        return ov

    def eval_warp(self, pts, center, impl: Optional[str] = None) -> torch.Tensor:
        p = self._f32(pts, 3); n = p.shape[0]
        off = torch.empty((n, 3), device=self.device, dtype=torch.float32)
        c = _lib.f3(center.detach().cpu().tolist() if isinstance(center, torch.Tensor) else center)
        self._check(self.lib.avc_eval_warp(self._h, _ptr(p), n, c, _ptr(off), self._impl(impl), self._stream()))
        return off

    def eval_template(self, pts, if_type: str = 'sdf', impl: Optional[str] = None):
        """DoubleTNet.forward for (N,3) points -> rgb (N,3), alpha (N,), occ (N,)."""
        if if_type not in ('sdf', 'occupancy'):
            raise ValueError('Invalid config.if_type!')
        p = self._f32(pts, 3); n = p.shape[0]
        rgb = torch.empty((n, 3), device=self.device, dtype=torch.float32)
        alpha = torch.empty(n, device=self.device, dtype=torch.float32)
        occ = torch.empty(n, device=self.device, dtype=torch.float32)
        self._check(self.lib.avc_eval_template(self._h, _ptr(p), n, _ptr(rgb), _ptr(alpha), _ptr(occ),
                                               IF_SDF if if_type == 'sdf' else IF_OCCUPANCY, self._impl(impl), self._stream()))
        return rgb, alpha, occ

    def eval_recon(self, pts, center, impl: Optional[str] = None) -> torch.Tensor:
        p = self._f32(pts, 3); n = p.shape[0]
        ov = torch.empty(n, device=self.device, dtype=torch.float32)
        c = _lib.f3(center.detach().cpu().tolist() if isinstance(center, torch.Tensor) else center)
        self._check(self.lib.avc_eval_recon(self._h, _ptr(p), n, c, _ptr(ov), self._impl(impl), self._stream()))
        return ov

    # ------------------------------------------------------------------ host-buffer (end-to-end) entry points
    def eval_occupancy_host(self, pts_host: np.ndarray, center, out_occ: np.ndarray, out_off: Optional[np.ndarray] = None,
                            out_rgb: Optional[np.ndarray] = None, out_alpha: Optional[np.ndarray] = None,
                            if_type: str = 'sdf', impl: Optional[str] = None) -> None:
        assert pts_host.dtype == np.float32 and pts_host.flags.c_contiguous and out_occ.dtype == np.float32
        n = pts_host.shape[0]
        self._check(self.lib.avc_eval_occupancy_host(self._h, pts_host.ctypes.data_as(C.c_void_p), n, _lib.f3(center),
                                                     out_occ.ctypes.data_as(C.c_void_p),
                                                     None if out_off is None else out_off.ctypes.data_as(C.c_void_p),
                                                     None if out_rgb is None else out_rgb.ctypes.data_as(C.c_void_p),
                                                     None if out_alpha is None else out_alpha.ctypes.data_as(C.c_void_p),
                                                     IF_SDF if if_type == 'sdf' else IF_OCCUPANCY, self._impl(impl)))

    def eval_recon_host(self, pts_host: np.ndarray, center, out_ov: np.ndarray, impl: Optional[str] = None) -> None:
        assert pts_host.dtype == np.float32 and pts_host.flags.c_contiguous and out_ov.dtype == np.float32
        self._check(self.lib.avc_eval_recon_host(self._h, pts_host.ctypes.data_as(C.c_void_p), pts_host.shape[0], _lib.f3(center),
                                                 out_ov.ctypes.data_as(C.c_void_p), self._impl(impl)))

    # ------------------------------------------------------------------ grid / scatter
    def make_grid(self, bounds, res, x_first: int = 0, x_count: Optional[int] = None) -> torch.Tensor:
        x_count = res[0] - x_first if x_count is None else x_count
        out = torch.empty((x_count * res[1] * res[2], 3), device=self.device, dtype=torch.float32)
        b = np.asarray(bounds, dtype=np.float32).reshape(6)
        self._check(self.lib.avc_make_grid(self._h, _lib.f6(b), _lib.i3(res), x_first, x_count, _ptr(out), self._stream()))
        return out

    def scatter_fill(self, flag: torch.Tensor, vals: torch.Tensor, fill: torch.Tensor) -> torch.Tensor:
        flag = flag.to(self.device).contiguous()
        f8 = flag.view(torch.uint8) if flag.dtype == torch.bool else flag.to(torch.uint8)
        vals = self._f32(vals).reshape(-1); fill = self._f32(fill).reshape(-1)
        out = torch.empty(f8.numel(), device=self.device, dtype=torch.float32)
        self._check(self.lib.avc_scatter_fill(self._h, _ptr(f8), f8.numel(), _ptr(vals), _ptr(fill), _ptr(out), self._stream()))
        return out

    # ------------------------------------------------------------------ mesh extraction
    def mc_count(self, vol: torch.Tensor, iso: float, halo_lo: int = 0, halo_hi: int = 0) -> Tuple[int, int]:
        vol = self._f32(vol)
        nv, nf = C.c_int64(), C.c_int64()
        self._check(self.lib.avc_mc_count(self._h, _ptr(vol), _lib.i3(vol.shape), float(iso), halo_lo, halo_hi, C.byref(nv), C.byref(nf),
                                          self._stream()))
        return nv.value, nf.value

    def extract_mesh_async(self, vol: torch.Tensor, bounds, iso: float, cap_v: int, cap_f: int, with_normals: bool = True,
                           halo_lo: int = 0, halo_hi: int = 0, x_origin: int = 0, gres_x: Optional[int] = None):
        """avc_mc_extract: everything enqueued, nothing synchronised. -> (verts (cap_v,3), faces (cap_f,3), normals | None,
        counts (4,) int64 on the device = {n_verts, n_faces, n_verts incl. the next slab's first plane, overflow flags})."""
        vol = self._f32(vol)
        if vol.dim() != 3:
            raise ValueError('volume must be (Rx,Ry,Rz)')
        # (a zero-row torch tensor has a NULL data pointer: keep one spare row so that capacity 0 is a legal, always-overflowing request)
        verts = torch.empty((cap_v + 1, 3), device=self.device, dtype=torch.float32)[:cap_v]
        faces = torch.empty((cap_f + 1, 3), device=self.device, dtype=torch.int32)[:cap_f]
        normals = torch.empty((cap_v + 1, 3), device=self.device, dtype=torch.float32)[:cap_v] if with_normals else None
        counts = torch.empty(4, device=self.device, dtype=torch.int64)
        b = np.asarray(bounds, dtype=np.float32).reshape(6)
        gx = vol.shape[0] if gres_x is None else gres_x
        self._check(self.lib.avc_mc_extract(self._h, _ptr(vol), _lib.i3(vol.shape), _lib.f6(b), float(iso), halo_lo, halo_hi, x_origin, gx,
                                            _ptr(verts), _ptr(normals), _ptr(faces), int(cap_v), int(cap_f), _ptr(counts), self._stream()))
        return verts, faces, normals, counts

    def extract_mesh(self, vol: torch.Tensor, bounds, iso: float, with_normals: bool = True, halo_lo: int = 0, halo_hi: int = 0,
                     x_origin: int = 0, gres_x: Optional[int] = None):
        """-> verts (V,3) f32, faces (F,3) i32, normals (V,3) f32 | None, all on the device, in the reference's conventions.
        The sizes are data dependent: the kernels run against capacity buffers (sized from the last meshes this engine extracted,
        or 1/16 of the voxels the first time) and the ONE host synchronisation is the read of the counts after everything has been
        enqueued; an overflow re-runs with the exact sizes."""
        vol = self._f32(vol)
        if vol.dim() != 3:
            raise ValueError('volume must be (Rx,Ry,Rz)')
        nvox = int(vol.numel())
        hint = self._keep.get('mc_hint')
        cap_v, cap_f = hint if hint else (max(4096, nvox // 16), max(8192, nvox // 8))
        for _ in range(2):
            verts, faces, normals, counts = self.extract_mesh_async(vol, bounds, iso, cap_v, cap_f, with_normals, halo_lo, halo_hi, x_origin, gres_x)
            nv, nf, _, over = (int(x) for x in counts.tolist())            # the only synchronisation of the extraction
            if not over:
                break
            cap_v, cap_f = max(nv, 1), max(nf, 1)
        else:
            raise AvcError(_lib.ECAPACITY, 'marching cubes overflowed its exact-size retry')
        self._keep['mc_hint'] = (max(4096, nv + nv // 4), max(8192, nf + nf // 4))
        return verts[:nv], faces[:nf], (normals[:nv] if with_normals else None)

    def renumber_faces(self, faces: torch.Tensor, n_own: int, base_own: int, base_next: int) -> None:
        """Slab-mesh seam rule, in place on an (F,3) int32 device tensor (shard.gather_mesh)."""
        if faces.dtype != torch.int32 or not faces.is_contiguous() or faces.device != self.device:
            raise ValueError('faces must be a contiguous int32 tensor on the engine device')
        self._check(self.lib.avc_renumber_faces(self._h, _ptr(faces), faces.shape[0], int(n_own), int(base_own), int(base_next), self._stream()))

    # ------------------------------------------------------------------ KNN / LBS
    def knn(self, query, ref, K: int = 1):
        q = self._f32(query, 3); r = self._f32(ref, 3)
        n = q.shape[0]
        d2 = torch.empty((n, K), device=self.device, dtype=torch.float32)
        idx = torch.empty((n, K), device=self.device, dtype=torch.int64)
        self._check(self.lib.avc_knn(self._h, _ptr(q), n, _ptr(r), r.shape[0], K, _ptr(d2), _ptr(idx), self._stream()))
        return d2, idx

    def near_flag(self, query, ref, radius: float) -> torch.Tensor:
        q = self._f32(query, 3); r = self._f32(ref, 3)
        out = torch.empty(q.shape[0], device=self.device, dtype=torch.uint8)
        self._check(self.lib.avc_near_flag(self._h, _ptr(q), q.shape[0], _ptr(r), r.shape[0], float(radius), _ptr(out), self._stream()))
        return out.bool()

    def inside_volume(self, verts, faces, bounds, res) -> torch.Tensor:
        """(Rx,Ry,Rz) bool: grid point inside the closed mesh (trimesh.contains on the dense grid, avatarcap_dataset.py:120-124)."""
        v = self._f32(verts, 3); f = torch.as_tensor(np.asarray(faces) if not isinstance(faces, torch.Tensor) else faces).to(self.device, torch.int32).contiguous()
        out = torch.empty(tuple(int(r) for r in res), device=self.device, dtype=torch.uint8)
        b = np.asarray(bounds, dtype=np.float32).reshape(6)
        self._check(self.lib.avc_inside_volume(self._h, _ptr(v), v.shape[0], _ptr(f), f.shape[0], _lib.f6(b), _lib.i3(res), _ptr(out), self._stream()))
        return out.bool()

    def lbs_weights(self, pts, cano_verts, skin_weights) -> torch.Tensor:
        p = self._f32(pts, 3); v = self._f32(cano_verts, 3); w = self._f32(skin_weights, 24)
        out = torch.empty((p.shape[0], 24), device=self.device, dtype=torch.float32)
        self._check(self.lib.avc_lbs_weights(self._h, _ptr(p), p.shape[0], _ptr(v), v.shape[0], _ptr(w), _ptr(out), self._stream()))
        return out

    def skin_points(self, pts, lbs, jnt_mats, return_pt_mats: bool = False):
        p = self._f32(pts, 3); l = self._f32(lbs, 24); J = self._f32(jnt_mats).reshape(24, 16)
        out = torch.empty_like(p)
        mats = torch.empty((p.shape[0], 4, 4), device=self.device, dtype=torch.float32) if return_pt_mats else None
        self._check(self.lib.avc_skin_points(self._h, _ptr(p), _ptr(l), _ptr(J), p.shape[0], _ptr(out), _ptr(mats), self._stream()))
        return (out, mats) if return_pt_mats else out

    def skin_normals(self, normals, lbs, jnt_mats) -> torch.Tensor:
        nrm = self._f32(normals, 3); l = self._f32(lbs, 24); J = self._f32(jnt_mats).reshape(24, 16)
        out = torch.empty_like(nrm)
        self._check(self.lib.avc_skin_normals(self._h, _ptr(nrm), _ptr(l), _ptr(J), nrm.shape[0], _ptr(out), self._stream()))
        return out

    def skin_mesh(self, verts, normals, cano_verts, skin_weights, jnt_mats):
        v = self._f32(verts, 3); nrm = None if normals is None else self._f32(normals, 3)
        cv = self._f32(cano_verts, 3); w = self._f32(skin_weights, 24); J = self._f32(jnt_mats).reshape(24, 16)
        ov = torch.empty_like(v); on = None if nrm is None else torch.empty_like(nrm)
        self._check(self.lib.avc_skin_mesh(self._h, _ptr(v), _ptr(nrm), v.shape[0], _ptr(cv), cv.shape[0], _ptr(w), _ptr(J), _ptr(ov), _ptr(on),
                                           self._stream()))
        return ov, on

    def posed_to_cano(self, wpts, live_verts, skin_weights, live2cano_mats, bounds, weight_volume):
        p = self._f32(wpts, 3); lv = self._f32(live_verts, 3); w = self._f32(skin_weights, 24)
        J = self._f32(live2cano_mats).reshape(24, 16); vol = self._f32(weight_volume, 24)
        cano = torch.empty_like(p); near = torch.empty(p.shape[0], device=self.device, dtype=torch.uint8)
        b = np.asarray(bounds.detach().cpu().numpy() if isinstance(bounds, torch.Tensor) else bounds, dtype=np.float32).reshape(6)
        self._check(self.lib.avc_posed_to_cano(self._h, _ptr(p), p.shape[0], _ptr(lv), lv.shape[0], _ptr(w), _ptr(J), _lib.f6(b), _ptr(vol),
                                               _lib.i3(vol.shape[:3]), _ptr(cano), _ptr(near), self._stream()))
        return cano, near.bool()


    # ------------------------------------------------------------------ vertex-colour driver (NerfRenderer / raw2outputs)
    def ray_samples(self, ray_o, ray_d, near, far, n_samples: int):
        o = self._f32(ray_o, 3); d = self._f32(ray_d, 3); nr = self._f32(near).reshape(-1); fr = self._f32(far).reshape(-1)
        n = o.shape[0]
        pts = torch.empty((n * n_samples, 3), device=self.device, dtype=torch.float32)
        z = torch.empty((n, n_samples), device=self.device, dtype=torch.float32)
        dists = torch.empty((n * n_samples,), device=self.device, dtype=torch.float32)
        self._check(self.lib.avc_ray_samples(self._h, _ptr(o), _ptr(d), _ptr(nr), _ptr(fr), n, n_samples, _ptr(pts), _ptr(z), _ptr(dists), self._stream()))
        return pts, z, dists

    def nerf_raw(self, cano_q, near_flag, rgb, alpha_raw, dists, bounds) -> torch.Tensor:
        q = self._f32(cano_q, 3); nf = near_flag.to(self.device).contiguous()
        nf8 = nf.view(torch.uint8) if nf.dtype == torch.bool else nf.to(torch.uint8)
        c = self._f32(rgb, 3); al = self._f32(alpha_raw).reshape(-1); dd = self._f32(dists).reshape(-1)
        raw = torch.empty((q.shape[0], 4), device=self.device, dtype=torch.float32)
        b = np.asarray(bounds.detach().cpu().numpy() if isinstance(bounds, torch.Tensor) else bounds, dtype=np.float32).reshape(6)
        self._check(self.lib.avc_nerf_raw(self._h, _ptr(q), _ptr(nf8), _ptr(c), _ptr(al), _ptr(dd), _lib.f6(b), q.shape[0], _ptr(raw), self._stream()))
        return raw

    def composite(self, raw, z_vals, white_bkgd: bool = False):
        r = self._f32(raw, 4); z = self._f32(z_vals)
        n, S = z.shape
        rgb = torch.empty((n, 3), device=self.device, dtype=torch.float32)
        acc = torch.empty(n, device=self.device, dtype=torch.float32); dep = torch.empty(n, device=self.device, dtype=torch.float32)
        self._check(self.lib.avc_composite(self._h, _ptr(r), _ptr(z), n, S, int(white_bkgd), _ptr(rgb), _ptr(acc), _ptr(dep), self._stream()))
        return rgb, acc, dep


    # ------------------------------------------------------------------ off-screen rasteriser / normal canonicalisation
    def rasterize(self, verts, faces, attrs, mvp, width: int, height: int, bg=(0., 0., 0.), cull: bool = True, flip_x: bool = False,
                  channels: int = 4) -> torch.Tensor:
        """Renderer.render() (utils/renderer.py:428-451) -> (height, width, channels) float32 device tensor, row 0 = top.
        faces None = triangle soup (glDrawArrays); attrs None = the 'position' shader."""
        v = self._f32(verts, 3)
        f = None
        if faces is not None:
            f = (faces if isinstance(faces, torch.Tensor) else torch.as_tensor(np.asarray(faces))).to(self.device, torch.int32).contiguous()
            if f.dim() != 2 or f.shape[1] != 3:
                raise ValueError('faces must be (F,3)')
        a = None if attrs is None else self._f32(attrs, 3)
        if a is not None and a.shape[0] != v.shape[0]:
            raise ValueError('one attribute row per vertex expected')
        nf = f.shape[0] if f is not None else v.shape[0] // 3
        m = (C.c_float * 16)(*[float(x) for x in np.asarray(mvp, dtype=np.float32).reshape(16)])
        out = torch.empty((height, width, channels), device=self.device, dtype=torch.float32)
        flags = (_lib.RASTER_CULL_BACK if cull else 0) | (_lib.RASTER_FLIP_X if flip_x else 0)
        self._check(self.lib.avc_rasterize(self._h, _ptr(v), v.shape[0], _ptr(f), nf, _ptr(a), m, int(width), int(height), _lib.f3(bg), flags,
                                           int(channels), _ptr(out), self._stream()))
        return out

    def canonicalize_normals(self, live_verts, vert_mats, mv, fx, fy, cx, cy, position_map, normal_map) -> torch.Tensor:
        """Per-vertex part of canonicalize_normal_map (normal_fusion/normal_fusion.py:27-62) -> (V,3) canonical normals."""
        v = self._f32(live_verts, 3); vm = self._f32(vert_mats).reshape(-1, 16)
        pm = self._f32(position_map); nm = self._f32(normal_map, 3)
        if vm.shape[0] != v.shape[0] or pm.dim() != 3 or nm.dim() != 3 or pm.shape[:2] != nm.shape[:2] or pm.shape[2] not in (3, 4):
            raise ValueError('canonicalize_normals: inconsistent shapes')
        m = (C.c_float * 16)(*[float(x) for x in np.asarray(mv, dtype=np.float32).reshape(16)])
        out = torch.empty_like(v)
        self._check(self.lib.avc_canonicalize_normals(self._h, _ptr(v), _ptr(vm), v.shape[0], m, float(fx), float(fy), float(cx), float(cy), _ptr(pm),
                                                      int(pm.shape[2]), _ptr(nm), int(nm.shape[0]), int(nm.shape[1]), _ptr(out), self._stream()))
        return out


_default: Dict[int, Engine] = {}


def default_engine(device: Optional[torch.device] = None) -> Engine:
    """Process-wide engine per device (what patch.install() uses)."""
    idx = torch.cuda.current_device() if device is None or torch.device(device).index is None else torch.device(device).index
    if idx not in _default:
        _default[idx] = Engine(torch.device('cuda', idx))
    return _default[idx]
