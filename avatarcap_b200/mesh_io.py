"""Mesh output in the reference's file format (utils/obj_io.py:223-269 save_mesh_as_ply).

The reference packs one `struct` per vertex and per face in a Python loop (seconds for a 1.6 M-vertex frame); here the same
bytes are produced by one structured-array write. Byte-identical files are pinned by tests/golden/ply_golden.npz."""
from __future__ import annotations

import numpy as np


def _np(x):
    if x is None:
        return None
    if hasattr(x, 'detach'):
        x = x.detach().cpu().numpy()
    return np.asarray(x)


def save_mesh_as_ply(path, vertices, faces=None, normals=None, colors=None) -> None:
    """Same arguments, header and binary layout as the reference (binary little-endian; x y z [nx ny nz] [red green blue];
    faces as `list int int`); colours below 1.0 are scaled by 255 (:245-247). Unlike the reference, `colors` is not modified in place."""
    vertices = _np(vertices).astype(np.float32).reshape(-1, 3)
    faces = _np(faces); normals = _np(normals); colors = _np(colors)
    nv = vertices.shape[0]
    head = 'ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n' % nv
    fields = [('v', '<f4', (3,))]
    if normals is not None:
        head += 'property float nx\nproperty float ny\nproperty float nz\n'
        fields.append(('n', '<f4', (3,)))
    if colors is not None:
        head += 'property uchar red\nproperty uchar green\nproperty uchar blue\n'
        fields.append(('c', 'u1', (3,)))
    face_num = 0 if faces is None else faces.shape[0]
    head += 'element face %d\nproperty list int int vertex_indices\nend_header\n' % face_num
    rec = np.empty(nv, dtype=np.dtype(fields))          # packed (no padding), like struct.pack('6f3B', ...)
    rec['v'] = vertices
    if normals is not None:
        rec['n'] = normals.astype(np.float32).reshape(-1, 3)
    if colors is not None:
        c = colors.reshape(-1, 3)
        if c.size and c.max() < 1.:
            c = c * 255
        rec['c'] = c.astype(np.uint8)
    with open(path, 'wb') as fp:
        fp.write(head.encode('ascii'))
        rec.tofile(fp)
        if faces is not None and face_num:
            fr = np.empty((face_num, 4), '<i4')
            fr[:, 0] = 3
            fr[:, 1:] = faces.reshape(-1, 3)
            fr.tofile(fp)
