"""run under torchrun (one process per GPU, NCCL): the ranks share the packing (dist.pack_over_ranks: each rank uploads a
slice of the rows, one all-reduce OR of the bitsets), scan their parts (dist.scan_part) and merge (dist.merge_over_ranks);
rank 0 compares with the single-GPU result of the same MSA and prints DIST_OK."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import repeatresolver_b200 as rr  # noqa: E402
from repeatresolver_b200.dist import merge_over_ranks, pack_over_ranks, scan_part  # noqa: E402

local_rank = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
g = rr.MsaGen(type="Tree", copies=30, coverage=40, repeat_len=1500, diff=0.02, seed=91, flank=700, min_overlap=100)
codes = g.codes()
msa = rr.MSA.from_cells(codes)
pk = pack_over_ranks(msa, local_rank)
st = scan_part(pk, 30)
M, A = pk.fetch()
M, A = merge_over_ranks(M, A)
pairs = torch.tensor([st["pair_tests"]], dtype=torch.int64, device="cuda")
dist.all_reduce(pairs)
pk.finalize(M, A)
if dist.get_rank() == 0:
    one = rr.Packed(msa, local_rank)
    st1 = one.scan(mincov=30)
    M1, A1 = one.fetch()
    one.finalize(M1, A1)
    gs, cv = pk.sizes()
    gs1, cv1 = one.sizes()
    assert (gs == gs1).all() and (cv == cv1).all()
    assert int(pairs.item()) == st1["pair_tests"], (int(pairs.item()), st1["pair_tests"])
    assert (M == M1).all() and (A == A1).all()
    print("DIST_OK world", dist.get_world_size(), "pairs", st1["pair_tests"], flush=True)
dist.barrier()
dist.destroy_process_group()
