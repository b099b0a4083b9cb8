"""Multi-process plumbing for the scan: one process per GPU (torch.distributed, NCCL on GPUs,
gloo on CPU for tests).  The hot path has no collective: every rank holds the whole packed MSA
and scans its own pair-balanced range of row sites (rr_scan part_index/part_count).  The exchanges
around it: the packed bitsets when the ranks share the packing (pack_over_ranks), the seeded
thresholds (scan_part), and the final element-wise max of the per-group results, the multi-process
form of the reference's thread merge (MaxCorrelation.c:882-891).
"""
import numpy as np

INT_MAX = 2 ** 31 - 1


def merge_over_ranks(M, A, device=None):
    """Element-wise max of the per-rank maxima M[5N] (float64); among ranks attaining the max the
    smallest partner id wins (A[5N] int32, -1 = none).  Works on any initialised process group;
    returns numpy arrays.  With no process group (single process) it returns its inputs."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return M, A
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    Mt = torch.from_numpy(np.ascontiguousarray(M)).to(dev)
    At = torch.from_numpy(np.ascontiguousarray(A)).to(dev)
    Mg = Mt.clone()
    dist.all_reduce(Mg, op=dist.ReduceOp.MAX)
    cand = torch.where((Mt == Mg) & (At >= 0) & (Mg > 0), At, torch.full_like(At, INT_MAX))
    dist.all_reduce(cand, op=dist.ReduceOp.MIN)
    cand = torch.where(cand == INT_MAX, torch.full_like(cand, -1), cand)
    return Mg.cpu().numpy(), cand.cpu().numpy()


class _DevBuf:
    """a device buffer of the library as a CUDA array (torch.as_tensor wraps it without a copy)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<i4", "data": (ptr, False), "version": 3}


def pack_over_ranks(msa, device):
    """Packed MSA on this rank's GPU with the upload and the packing shared by all ranks: rank r uploads rows
    [r R / world, (r + 1) R / world) of the MSA (every rank holds it in host memory), the covered spans of all rows are
    all-gathered (3 int32 per row), every rank packs its rows into full-size bitsets that are zero elsewhere, and ONE
    all-reduce SUM (= OR: the slices are disjoint) over NVLink gives every GPU the whole packed MSA - 1.4 GB at config 2
    instead of every rank pushing the same 1.8 GB through host memory and PCIe.  The exchange step of the multi-GPU
    Einlesen; the scan itself still needs no collective.  Single rank: plain Packed(msa, device)."""
    import torch
    import torch.distributed as dist
    from .maxcorr import Packed
    rank, world = rank_part()
    if world == 1:
        return Packed(msa, device)
    R = msa.rows
    lo, hi = R * rank // world, R * (rank + 1) // world
    pk = Packed(msa, device, rows=(lo, hi))
    nccl = dist.get_backend() == "nccl"
    per = (R + world - 1) // world + 1
    mine = np.zeros((3, per), dtype=np.int32)
    mine[:, :hi - lo] = pk.slice_spans()
    t = torch.from_numpy(mine)
    if nccl:
        t = t.cuda()
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    spans = np.zeros((3, R), dtype=np.int32)
    for r in range(world):
        a, b = R * r // world, R * (r + 1) // world
        spans[:, a:b] = parts[r].cpu().numpy()[:, :b - a]
    pk.set_spans(spans)
    ptr, nbytes = pk.bits_device()
    if nccl:
        buf = torch.as_tensor(_DevBuf(ptr, nbytes), device=torch.device("cuda", torch.cuda.current_device()))
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        torch.cuda.current_stream().synchronize()
    else:
        raise RuntimeError("pack_over_ranks moves device buffers: it needs the nccl backend")
    pk.finish()
    return pk


def rank_part():
    """(part_index, part_count) of this process"""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def scan_part(pk, mincov=30, variant="auto", flags=0):
    """This rank's part of the scan (pk: repeatresolver_b200.Packed of the whole MSA on this rank's GPU).
    With more than one rank: seeding pass on every rank, all-reduce MAX of the seeded maxima as common
    pruning thresholds (values only, 5N doubles), then the full pass.  Returns the stats of the full pass;
    results stay on the device (pk.fetch(), then merge_over_ranks)."""
    import torch
    import torch.distributed as dist
    from .maxcorr import FLAG_SEED_ONLY, FLAG_SKIP_SEED
    rank, world = rank_part()
    if world == 1:
        return pk.scan(mincov=mincov, variant=variant, flags=flags)
    st0 = pk.scan(mincov=mincov, variant=variant, flags=flags | FLAG_SEED_ONLY, part_index=rank, part_count=world)
    if dist.get_backend() == "nccl":
        # device to device: the library writes the maxima into a torch tensor, NCCL reduces it over NVLink,
        # the library reads it back as thresholds; nothing crosses PCIe
        Mt = torch.empty(5 * pk.cols, dtype=torch.float64, device=torch.device("cuda", torch.cuda.current_device()))
        pk.values_to_device(Mt.data_ptr())
        dist.all_reduce(Mt, op=dist.ReduceOp.MAX)
        torch.cuda.current_stream().synchronize()
        pk.set_thresholds_device(Mt.data_ptr())
    else:
        M, _ = pk.fetch()
        Mt = torch.from_numpy(M)
        dist.all_reduce(Mt, op=dist.ReduceOp.MAX)
        pk.set_thresholds(Mt.numpy())
    st = pk.scan(mincov=mincov, variant=variant, flags=flags | FLAG_SKIP_SEED, part_index=rank, part_count=world)
    st["kernel_ms"] += st0["kernel_ms"]  # both passes
    st["prepare_ms"] += st0["prepare_ms"]
    return st


def cliquer_over_ranks(cliquer_batch, query_groups, maxclique=30, **kw):
    """Cliquer (RepeatResolver.c:1179-1240) for a query list shared by all ranks.  Queries are independent, so the path
    shards without a data-path collective: rank r takes queries r, r + world, ... (cyclic, so that the significant
    groups of one region of the MSA spread over the GPUs), runs `cliquer_batch` on them (Packed.cliquer_batch of the
    rank's own packed copy of the MSA) and one all-gather returns every rank the full result in query order:
    (members [nq][maxclique+1], scores [nq][maxclique], n_members [nq], this rank's stats)."""
    q = np.ascontiguousarray(query_groups, dtype=np.int32).ravel()
    rank, world = rank_part()
    members, scores, n, st = cliquer_batch(q[rank::world], maxclique=maxclique, **kw)
    if world == 1:
        return members, scores, n, st
    return (_gather_cyclic(members, -1, len(q)), _gather_cyclic(scores, 0.0, len(q)), _gather_cyclic(n, 0, len(q)), st)


def _gather_cyclic(a, fill, nq):
    """rank r holds the rows r, r + world, ... of an [nq][...] array: one all-gather gives every rank the whole array"""
    import torch
    import torch.distributed as dist
    rank, world = rank_part()
    per = (nq + world - 1) // world                         # slices differ by at most one row: pad to the longest
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    pad = np.full((per,) + a.shape[1:], fill, dtype=a.dtype)
    pad[:len(a)] = a
    as_signed = {np.dtype(np.uint64): np.int64, np.dtype(np.uint32): np.int32}.get(pad.dtype)     # gloo / torch: signed views
    mine = torch.from_numpy(pad.view(as_signed) if as_signed else pad).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    out = np.full((nq,) + a.shape[1:], fill, dtype=a.dtype)
    for r in range(world):
        k = len(range(r, nq, world))
        part = parts[r][:k].cpu().numpy()
        out[r::world] = part.view(a.dtype) if as_signed else part
    return out


def group_refinement_over_ranks(group_refinement, MaxCorrs, cutoff, **kw):
    """Group_Refinement (RepeatResolver.c:1634-1693; the reference's own parallel form 1770-1821 deals the groups to threads
    by i % NTHREADS, 1717) for MaxCorrs shared by all ranks.  The groups above the cutoff are independent: rank r takes the
    r-th, (r + world)-th, ... of them - `group_refinement` is Packed.group_refinement of the rank's own packed copy of the MSA
    and sees MaxCorrs with every other entry lowered to -inf - and one all-gather per result array returns every rank the full
    result in ascending group order, in the layout of Packed.group_refinement (stats: this rank's)."""
    M = np.array(MaxCorrs, dtype=np.float64)
    rank, world = rank_part()
    if world == 1:
        return group_refinement(M, cutoff, **kw)
    q = np.flatnonzero(M > cutoff).astype(np.int32)
    masked = np.full_like(M, -np.inf)
    masked[q[rank::world]] = M[q[rank::world]]
    res = group_refinement(masked, cutoff, **kw)
    assert np.array_equal(res["groups"], q[rank::world])
    out = {"groups": q, "stats": res["stats"]}
    for key, fill in (("Cliques", -1), ("Sizes", 0), ("Cutoffs", 0), ("Drop_Off", 0.0), ("C_Groups", 0), ("C_Coverage", 0)):
        out[key] = None if res[key] is None else _gather_cyclic(res[key], fill, len(q))
    M[q[out["Sizes"] <= 5]] = 0.0                            # 1685
    out["MaxCorrs"] = M
    return out
