"""Shared scene construction for the parity tests (no GPU, no oracle imports here)."""
import os

import numpy as np

from avatarcap_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def golden_scene():
    """Everything gen_golden.py used, regenerated from seeds: body, frame (pose seed 7), feature maps, weights."""
    body = synth.SynthBody()
    frame = synth.make_frame(body, synth.random_pose(7, 0.4))
    return {
        'body': body, 'frame': frame,
        'avatar_sd': synth.avatar_state_dict(), 'recon_sd': synth.recon_state_dict(),
        'pose_map': synth.feature_map(64, 48, 40, synth.SEED + 2),
        'image_map': synth.feature_map(32, 36, 44, synth.SEED + 3),
    }


def tpose_scene(map_hw=256):
    """BASELINE configs 1-4: T-pose live body, 256x256 feature maps."""
    body = synth.SynthBody()
    frame = synth.make_frame(body, None)
    return {
        'body': body, 'frame': frame,
        'avatar_sd': synth.avatar_state_dict(), 'recon_sd': synth.recon_state_dict(),
        'pose_map': synth.feature_map(64, map_hw, map_hw, synth.SEED + 4),
        'image_map': synth.feature_map(32, map_hw, map_hw, synth.SEED + 5),
    }


def maxabs(a, b):
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)))) if np.size(a) else 0.0


def uv_sphere(center, radius, n_lat=24, n_lon=32):
    """Closed, outward-oriented UV sphere (verts (V,3) f32, faces (F,3) i32)."""
    th = np.linspace(0, np.pi, n_lat + 1)[1:-1]; ph = np.linspace(0, 2 * np.pi, n_lon, endpoint=False)
    v = [[0, 0, 1]] + [[np.sin(t) * np.cos(p), np.sin(t) * np.sin(p), np.cos(t)] for t in th for p in ph] + [[0, 0, -1]]
    v = np.asarray(v) * radius + np.asarray(center)
    f = []
    for j in range(n_lon):
        f.append([0, 1 + j, 1 + (j + 1) % n_lon])
    for i in range(n_lat - 2):
        for j in range(n_lon):
            a = 1 + i * n_lon + j; b = 1 + i * n_lon + (j + 1) % n_lon; c = a + n_lon; d = b + n_lon
            f += [[a, c, b], [b, c, d]]
    last = len(v) - 1; base = 1 + (n_lat - 2) * n_lon
    for j in range(n_lon):
        f.append([last, base + (j + 1) % n_lon, base + j])
    return v.astype(np.float32), np.asarray(f, np.int32)
