"""The N>1 path on CPU: world_size-2 gloo process group, each rank holding a partial result
(here produced by the oracle on a disjoint half of the row sites, as a GPU rank would for its
part), merged with repeatresolver_b200.dist.merge_over_ranks.  The merge must reproduce the
full scan exactly, arg-max tie rule included."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, codes, mincov, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib as O
    from repeatresolver_b200.dist import merge_over_ranks, rank_part
    assert rank_part() == (rank, world)
    o = O.Oracle.from_codes(codes)
    M, A, P = o.scan(mincov, modulus=world, res_lo=rank, res_hi=rank + 1)  # this rank's rows
    Mg, Ag = merge_over_ranks(M, A)
    Pt = torch.tensor([P], dtype=torch.int64)
    dist.all_reduce(Pt)
    if rank == 0:
        np.savez(os.path.join(out_dir, "merged.npz"), M=Mg, A=Ag, P=int(Pt.item()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_merge_over_ranks_gloo(world, tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    import repeatresolver_b200 as rr
    g = rr.MsaGen(type="Tree", copies=4, coverage=16, repeat_len=500, diff=0.03, seed=77, flank=300, min_overlap=50)
    codes = g.codes()
    # duplicate a few columns so that exact ties between different partners exist
    codes[:, 300:320] = codes[:, 100:120]
    port = 29500 + (os.getpid() % 500) + world
    mp.spawn(_worker, args=(world, port, codes, 12, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "merged.npz")
    M0, A0, P0 = O.Oracle.from_codes(codes).scan(12)
    assert int(got["P"]) == P0
    assert (got["M"] == M0).all()
    assert (got["A"] == A0).all()


def test_merge_single_process_is_identity():
    from repeatresolver_b200.dist import merge_over_ranks
    M = np.array([0.0, 1.5]); A = np.array([-1, 7], dtype=np.int32)
    M2, A2 = merge_over_ranks(M, A)
    assert M2 is M and A2 is A
