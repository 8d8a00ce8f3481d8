"""Run under torchrun on N GPUs (not collected by pytest): every rank scans its contiguous window range on its own GPU,
results are gathered with NCCL, and rank 0 checks that they are bit-identical to a single-GPU scan of the whole region.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/multi_gpu_check.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import fpt_b200.api as api
    import fpt_b200.synth as synth
    from fpt_b200.sharding import css_scan_sharded, fet_scan_sharded
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    api.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    regend, wsize, wstep, nsnp, asize, bsize, seed = 400000, 2500, 500, 9000, 20, 20, 11
    ch = synth.chromosome(5, regend, nsnp, asize, bsize)
    a, b, pos = ch["acodes"], ch["bcodes"], ch["pos"]
    s, d = fet_scan_sharded(a, b, pos, asize, bsize, regend, wsize, wstep, 0.95, rank, world, seed=seed)
    out = {}
    for mds in (0, 1, 2):
        out[mds] = css_scan_sharded(a, b, pos, asize, bsize, regend, wsize, wstep, 10, 300, rank, world, mds=mds, seed=seed)
    if rank == 0:
        s1, d1, _ = api.fet_scan(a, b, pos, asize, bsize, regend, wsize, wstep, 0.95, seed=seed)
        assert np.array_equal(s, s1) and np.array_equal(d, d1), "FET: sharded != single GPU"
        for mds in (0, 1, 2):
            c1, p1, _ = api.css_scan(a, b, pos, asize, bsize, regend, wsize, wstep, 10, 300, mds=mds, seed=seed)
            assert np.array_equal(out[mds][0], c1, equal_nan=True) and np.array_equal(out[mds][1], p1), "CSS mds=%d: sharded != single GPU" % mds
        print("multi-GPU check ok: %d ranks, %d windows, FET + CSS (mds 0,1,2) bit-identical to one GPU" % (world, regend // wstep))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
