"""Multi-rank GPU check of the slab exchange + mesh gather (run under torchrun on a GPU box; pytest cannot launch ranks):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29711 tests/mgpu_check.py [res]

For several epochs with a DIFFERENT volume each (a stale halo plane would change the mesh): every rank fills its slab, the boundary
planes travel by our peer-memory kernels (mode p2p) and by NCCL send/recv (mode sendrecv), marching cubes runs per slab, the mesh
is gathered count-then-payload on rank 0 and must equal -- vertices, normals, faces, bit for bit -- the mesh rank 0 extracts from
the whole volume on one GPU. Uneven slabs (Rx not divisible by N) included. Prints one JSON line on rank 0.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from avatarcap_b200 import shard  # noqa: E402
from avatarcap_b200.engine import Engine  # noqa: E402


def main():
    rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    eng = Engine(dev)
    big = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    bounds = np.array([[-0.9, -1.0, -0.35], [0.95, 0.9, 0.3]], np.float32)
    report = {'world': world, 'cases': []}
    ok_all = True
    for res in ((8 * world + 3, 20, 24), (big + 1, big, big)):
        for mode in ('p2p', 'sendrecv'):
            try:
                sv = shard.SlabVolume(res, world, rank, engine=eng, mode=mode)
            except RuntimeError as ex:
                if rank == 0:
                    report['cases'].append({'res': res, 'mode': mode, 'error': str(ex)[:200]})
                ok_all = False
                continue
            times = []
            for epoch in range(4):
                g = torch.Generator(device='cpu'); g.manual_seed(1000 * epoch + 7)
                # a noisy sphere: many vertices on every slab boundary; the same generator on every rank -> the same whole volume
                ii = torch.arange(res[0], dtype=torch.float32)[:, None, None]; jj = torch.arange(res[1], dtype=torch.float32)[None, :, None]
                kk = torch.arange(res[2], dtype=torch.float32)[None, None, :]
                whole = (min(res) / 2.5 - torch.sqrt((ii - res[0] / 2) ** 2 + (jj - res[1] / 2) ** 2 + (kk - res[2] / 2) ** 2)
                         + 0.7 * torch.randn(res, generator=g)).to(dev)
                sv.own.copy_(whole[sv.x0:sv.x1])
                torch.cuda.synchronize(); dist.barrier()
                t0 = time.perf_counter()
                mv, mf, mn, counts = shard.extract_sharded_mesh(eng, sv, bounds, 0.0)
                torch.cuda.synchronize()
                times.append((time.perf_counter() - t0) * 1e3)
                assert torch.equal(sv.padded, whole[sv.x0 - sv.lo:sv.x1 + sv.hi]), 'halo planes differ (rank %d, epoch %d, mode %s)' % (rank, epoch, mode)
                if rank == 0:
                    wv, wf, wn = eng.extract_mesh(whole, bounds, 0.0)
                    same = bool(wv.shape == mv.shape and wf.shape == mf.shape and torch.equal(wv, mv) and torch.equal(wf, mf) and torch.equal(wn, mn))
                    ok_all = ok_all and same
                    if epoch == 3:
                        report['cases'].append({'res': res, 'mode': sv.mode, 'verts': int(wv.shape[0]), 'faces': int(wf.shape[0]), 'equal': same,
                                                'ms_last': round(times[-1], 3), 'per_rank_counts': counts.tolist()})
            dist.barrier()
            sv.close()
    flag = torch.tensor([1 if ok_all else 0], device=dev)
    dist.broadcast(flag, 0)
    if rank == 0:
        report['ok'] = bool(ok_all)
        print(json.dumps(report))
    dist.barrier(); dist.destroy_process_group()
    eng.close()
    sys.exit(0 if int(flag[0]) else 1)


if __name__ == '__main__':
    main()
