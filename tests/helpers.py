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
