"""One eager forward of the hybrid UNet (encoders.PoseFeatureEncoderTC) for an ncu launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/unet_launches.csv python tests/diag_unet_launches.py
The first forward (cuDNN algorithm selection, lazy module loads) runs before cudaProfilerStart; only the second is listed."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from avatarcap_b200 import encoders, synth  # noqa: E402
from avatarcap_b200.engine import Engine  # noqa: E402

if __name__ == '__main__':
    eng = Engine()
    x = torch.from_numpy(synth.smpl_pos_map()).cuda()
    enc = encoders.PoseFeatureEncoderTC(synth.unet_state_dict(), engine=eng, use_graph=False)
    enc(x); torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    enc(x); torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
