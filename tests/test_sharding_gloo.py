"""world_size-2 gloo test (CPU) of the multi-GPU host logic: static particle shards + ONE all-reduce of
the raw tally + finalisation after the reduce reproduce the single-rank solve.  The per-rank compute is
the CPU oracle here (no GPU in this test); on the GPU box the same helpers wrap mcb_solve_raw_dev."""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from montecarlocpp_b200 import materials, sharding


def test_shard_ranges_partition_exactly():
    for n in (0, 1, 7, 1000, 12345, 10**9 + 7):
        for world in (1, 2, 3, 8):
            r = [sharding.shard_range(n, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:])) and all(b >= a for a, b in r)


def _worker(rank, world, port, matdir, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as orc
    from tests import cases
    mat = orc.Material(os.path.join(matdir, "grey_disp.txt"), os.path.join(matdir, "grey_relax2.txt"))
    dom = cases.slab(ncell=16)
    prob = orc.Problem(mat, dom, "multi", 30001, 40)
    b, e = sharding.shard_range(prob.nemit, world, rank)
    raw, st = prob.solve(rng=orc.RNG_PHILOX, seed=77, n_begin=b, n_end=e, nthreads=1, raw=True)
    t = torch.from_numpy(np.ascontiguousarray(raw))
    sharding.allreduce_raw_field(t)
    steps = torch.tensor([st["steps"], st["emitted"]], dtype=torch.int64)
    dist.all_reduce(steps)
    if rank == 0:
        fin = prob.finalize(t.numpy())
        full, fst = prob.solve(rng=orc.RNG_PHILOX, seed=77, nthreads=1)
        np.save(out, np.stack([fin, full]))
        assert steps[0].item() == fst["steps"] and steps[1].item() == prob.nemit
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_plus_allreduce_equal_full_solve():
    d = tempfile.mkdtemp()
    materials.write_grey(d)
    out = os.path.join(d, "res.npy")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, d, out), nprocs=2, join=True)
    fin, full = np.load(out)
    scale = np.abs(full).max(axis=1, keepdims=True)
    assert (np.abs(fin - full) <= 1e-10 * scale).all()
