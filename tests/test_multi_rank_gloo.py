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


def _free_port():
    """a port nobody listens on right now (the tests may run side by side under pytest-xdist)"""
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


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
    port = _free_port()
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


# ---- Cliquer over ranks: queries sliced cyclically, results all-gathered in query order --------------------------------
def _cliquer_worker(rank, world, port, codes, queries, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib as O
    from repeatresolver_b200.dist import cliquer_over_ranks
    o = O.Oracle.from_codes(codes)
    seen = []

    def batch(qs, mincov, maxclique, greedy):
        """stands in for Packed.cliquer_batch on a machine without a GPU: same signature and result layout"""
        seen.extend(int(q) for q in qs)
        members = np.full((len(qs), maxclique + 1), -1, dtype=np.int32)
        scores = np.zeros((len(qs), maxclique), dtype=np.float64)
        n = np.zeros(len(qs), dtype=np.int32)
        for k, q in enumerate(qs):
            m, z = o.cliquer(int(q), mincov, maxclique, greedy)
            members[k, :len(m)], scores[k, :len(z)], n[k] = m, z, len(m)
        return members, scores, n, {"pairs": len(qs) * 5 * codes.shape[1]}

    members, scores, n, st = cliquer_over_ranks(batch, queries, maxclique=8, mincov=12, greedy=2.0)
    assert seen == [int(q) for q in queries[rank::world]]
    np.savez(os.path.join(out_dir, f"clq{rank}.npz"), members=members, scores=scores, n=n)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nq", [(2, 7), (3, 10), (2, 1)])
def test_cliquer_over_ranks_gloo(world, nq, tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    import repeatresolver_b200 as rr
    g = rr.MsaGen(type="Tree", copies=4, coverage=16, repeat_len=500, diff=0.03, seed=78, flank=300, min_overlap=50)
    codes = g.codes()
    o = O.Oracle.from_codes(codes)
    M0, _, _ = o.scan(12)
    queries = np.argsort(-M0, kind="stable")[:nq].astype(np.int32)
    port = _free_port()
    mp.spawn(_cliquer_worker, args=(world, port, codes, queries, str(tmp_path)), nprocs=world, join=True)
    results = [np.load(tmp_path / f"clq{r}.npz") for r in range(world)]
    for k, q in enumerate(queries):
        m, z = o.cliquer(int(q), 12, 8, 2.0)
        for res in results:                                   # every rank holds the full result, in query order
            assert res["n"][k] == len(m)
            assert list(res["members"][k, :len(m)]) == list(m) and (res["members"][k, len(m):] == -1).all()
            assert np.array_equal(res["scores"][k, :len(z)], z)
    assert max(int(r["n"].max()) for r in results) > 1


def _refine_worker(rank, world, port, codes, M, cutoff, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib as O
    from repeatresolver_b200.dist import group_refinement_over_ranks
    o = O.Oracle.from_codes(codes)
    R = codes.shape[0]
    seen = []

    def refine(MaxCorrs, cutoff, mincov, maxclique, greedy):
        """stands in for Packed.group_refinement on a machine without a GPU: same signature and result layout"""
        want_M, want = O.group_refinement(o, codes, MaxCorrs, cutoff, mincov, maxclique, greedy)
        groups = np.array(sorted(want), dtype=np.int32)
        seen.extend(int(g) for g in groups)
        sc = R // 64 + 1
        res = {"MaxCorrs": want_M, "groups": groups, "Cliques": np.array([want[g]["clique"] for g in groups], dtype=np.int32).reshape(len(groups), maxclique + 1),
               "Sizes": np.array([want[g]["size"] for g in groups], dtype=np.int32), "Cutoffs": np.array([want[g]["cutoff"] for g in groups], dtype=np.int32),
               "Drop_Off": np.array([want[g]["drop_off"] for g in groups], dtype=np.float64),
               "C_Groups": np.zeros((len(groups), sc), dtype=np.uint64), "C_Coverage": None, "stats": {"rank": rank}}
        for k, g in enumerate(groups):
            if want[g]["size"] > 5:
                res["C_Groups"][k] = O.bitset_words(want[g]["group"])
        return res

    res = group_refinement_over_ranks(refine, M, cutoff, mincov=12, maxclique=10, greedy=2.0)
    q = np.flatnonzero(M > cutoff)
    assert seen == [int(g) for g in q[rank::world]]
    assert res["C_Coverage"] is None and res["stats"] == {"rank": rank}
    np.savez(os.path.join(out_dir, f"gr{rank}.npz"), **{k: v for k, v in res.items() if k not in ("C_Coverage", "stats")})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nq", [(2, 31), (3, 2)])             # three ranks, two groups: one rank has nothing to do
def test_group_refinement_over_ranks_gloo(world, nq, tmp_path):
    """Group_Refinement sharded by group over world ranks (no data-path collective, one all-gather per result array): every
    rank ends with the single-process result, bitsets with their top bit set included"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    import repeatresolver_b200 as rr
    g = rr.MsaGen(type="Tree", copies=4, coverage=16, repeat_len=500, diff=0.03, seed=78, flank=300, min_overlap=50)
    codes = g.codes()
    o = O.Oracle.from_codes(codes)
    M, _, _ = o.scan(12)
    vals = np.unique(M)[::-1]                                # the largest value that leaves at least nq groups above it
    cutoff = float(next(v for v in vals if np.count_nonzero(M > v) >= nq))
    nq = int(np.count_nonzero(M > cutoff))
    port = _free_port()
    mp.spawn(_refine_worker, args=(world, port, codes, M, cutoff, str(tmp_path)), nprocs=world, join=True)
    want_M, want = O.group_refinement(o, codes, M, cutoff, 12, 10, 2.0)
    assert len(want) == nq
    for r in range(world):
        res = np.load(tmp_path / f"gr{r}.npz")
        assert np.array_equal(res["MaxCorrs"], want_M) and [int(x) for x in res["groups"]] == sorted(want)
        for k, grp in enumerate(res["groups"]):
            w = want[int(grp)]
            assert list(res["Cliques"][k]) == list(w["clique"]) and res["Sizes"][k] == w["size"] and res["Cutoffs"][k] == w["cutoff"]
            assert res["Drop_Off"][k] == w["drop_off"]
            if w["size"] > 5:
                assert np.array_equal(res["C_Groups"][k], O.bitset_words(w["group"]))
            else:
                assert not res["C_Groups"][k].any()
    if nq > 4:
        assert sum(1 for w in want.values() if w["size"] > 5) >= 5
