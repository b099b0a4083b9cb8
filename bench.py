#!/usr/bin/env python
"""bench.py -- MaxCorrelation scan throughput on B200 (site-group pair tests / second).

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE
JSON line on rank 0.  A step is one full all-pairs scan of the workload MSA:
  value  whole-job pair tests / s with the packed MSA already resident in HBM (strong scaling:
         the same MSA is split into N pair-balanced row ranges, one per GPU, no collective on
         the hot path);
  e2e    the same metric through the C ABI with HOST buffers: H2D of the cell matrix from pinned
         memory + device pack + scan + D2H of the per-group result (+ the max-merge over ranks);
  roofline / cpu_baseline as specified.  `--impl reference` times the reference's own CPU
  implementation (oracle/_ref, else the oracle port) on a bounded sample of the same workload.

Workloads (synthetic, DataSimulator-shaped, csrc/msagen.c): BASELINE.json configs[1]
"Tree_1perc_30000" = Tree, 100 copies, 40x, 30 kbp, 1 % (R ~ 13.7k rows, N ~ 133k columns).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: MsaGen kwargs
    "Tree_1perc_30000": dict(type="Tree", copies=100, coverage=40, repeat_len=30000, diff=0.01, seed=1002),
    "Distributed_1perc_30000": dict(type="Distributed", copies=100, coverage=40, repeat_len=30000, diff=0.01, seed=1003),
    "EquiDistant_1perc_30000": dict(type="EquiDistant", copies=100, coverage=40, repeat_len=30000, diff=0.01, seed=1003),
    "Tree_1perc_5000": dict(type="Tree", copies=10, coverage=40, repeat_len=5000, diff=0.01, seed=1001),
    "Tree_1perc_10000_25": dict(type="Tree", copies=25, coverage=40, repeat_len=10000, diff=0.01, seed=1004),
    # BASELINE.json configs[4] shape (~40k reads, 100 kbp repeat): 37 120 rows x 440 004 columns, 1.66e11 pair tests
    "Tree_1perc_100000_92": dict(type="Tree", copies=92, coverage=40, repeat_len=100000, diff=0.01, seed=1005),
}
MINCOV = 30
# dram__bytes_read.sum + dram__bytes_write.sum of the scan kernel, per launch, from the committed ncu capture
# (profiles/); None until measured for that workload
TRAFFIC = {"Tree_1perc_30000": 1.048e11}  # profiles/r2_final_ncu_full_summary.csv (full pass, --set full, cold L2): dram read 103.6 GB + write 1.2 GB, mxf4 operands, CTA pairs


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region, sampled in-process through NVML every
    100 ms (an `nvidia-smi -lms` child process was measured to delay CUDA API calls of the timed loop)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.sm = []
        self.sm_max = None
        self.reasons = set()
        self.stop_flag = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                     "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            while not self.stop_flag.is_set():
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for nm, bit in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                self.stop_flag.wait(0.1)
        except Exception as e:  # NVML missing: report that no sample was taken
            self.error = repr(e)

    def finish(self):
        self.stop_flag.set()
        if self.is_alive():
            self.join(timeout=2)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(sm)}


def make_msa(rr, workload):
    g = rr.MsaGen(**WORKLOADS[workload], threads=min(32, os.cpu_count() or 8))
    msa = rr.MSA.alloc(g.rows, g.cols, codes=True)  # page-locked host cells
    g.codes(out=msa.cells())
    return g, msa


def site_pair_counts(codes, mincov):
    """per site ii, the number of PositiveSignificance calls (MaxCorrelation.c:820) its row groups make, for a code matrix
    whose rows are single spans, from the filters alone (numpy + the product's first-break sweep; no intersections)"""
    import numpy as np
    import repeatresolver_b200 as rr
    R, N = codes.shape
    gs = np.zeros((N, 5), dtype=np.int64)
    start = np.full(R, 2 ** 31 - 1, dtype=np.int32)
    end = np.full(R, -1, dtype=np.int32)
    for r0 in range(0, R, 512):                      # in row slabs: the config-2 matrix is 1.8 GB
        blk = codes[r0:r0 + 512]
        for k in range(5):
            gs[:, k] += (blk == k).sum(0)
        covd = blk < 5
        anyc = covd.any(1)
        start[r0:r0 + 512] = np.where(anyc, covd.argmax(1), 2 ** 31 - 1)
        end[r0:r0 + 512] = np.where(anyc, N - 1 - covd[:, ::-1].argmax(1), -1)
    coverage = gs.sum(1)
    q = mincov // 4
    colok = (gs > q) & (gs < R)
    basey = gs[:, :4].sum(1) > coverage // 2
    nrow = (colok & basey[:, None]).sum(1)
    pref = np.concatenate([[0], np.cumsum(colok.sum(1))])
    brk = np.minimum(rr.breakcols_from_spans(start, end, N, mincov), N)
    lo = np.minimum(np.arange(N) + 20, N)
    return nrow * np.maximum(0, pref[np.maximum(brk, lo)] - pref[lo])


def count_pairs_host(codes, mincov):
    return int(site_pair_counts(codes, mincov).sum())


def reference_samples(rr, g, msa, workload, seconds_target, tmpdir, reps=1):
    """Time the reference's CPU implementation on bounded samples of the workload, BASELINE.md section 3.4: the WHOLE MSA
    (written as the text file the reference reads), an exact 1/k cyclic sample of its row sites (ii % k == thread,
    MaxCorrelation.c:796) on all host cores, through the unmodified reference's own Einlesen and HilfsMaxCorrsRechner
    (oracle/_ref/ref_driver_big); the oracle port when that binary is absent.  The pair tests of a sample are counted twice and
    must agree: by the oracle's instrumented loop (rr_oracle_count_pairs: the reference's loops and filters without the score)
    and from the product's own filters / first-break sweep.  Returns a list of dicts, one per repetition."""
    import numpy as np
    import oracle_lib as O
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 64))
    cells = msa.cells()
    R, N = cells.shape
    drv = os.path.join(ROOT, "oracle", "_ref", "ref_driver_big")
    kind = "reference" if os.path.exists(drv) else "port"
    per_site = site_pair_counts(cells, MINCOV)
    P_total = int(per_site.sum())
    # 1/k of the row sites per thread: ~0.45 M pair tests/s/core at R ~ 13.7k, cost per pair test ~ R
    rate_guess = 0.45e6 * (13700.0 / max(R, 300)) * threads
    want_pairs = rate_guess * seconds_target
    modulus = int(max(threads, round(P_total / max(want_pairs, 1.0) * threads)))
    oracle = O.Oracle.from_codes(cells)
    out = []
    reps_per_run = max(1, min(reps, (min(128, modulus) // threads))) if kind == "reference" else 1
    path = os.path.join(tmpdir, "workload.msa")
    if kind == "reference" and not os.path.exists(path):
        g.write(path)
    while len(out) < reps:
        n = min(reps_per_run, reps - len(out))
        if kind == "reference":
            run = subprocess.run([drv, path, str(MINCOV), str(modulus), "0", str(threads), "-", str(n)], capture_output=True, text=True)
            lines = [l.split() for l in run.stdout.splitlines() if l.startswith("REF ")]
            if run.returncode != 0 or len(lines) != n:
                raise RuntimeError("ref_driver failed: " + run.stderr[-500:])
            runs = [(float(l[4]), int(l[6])) for l in lines]
        else:
            runs = []
            for r in range(n):
                t0 = time.perf_counter()
                oracle.scan(MINCOV, modulus, r * threads, (r + 1) * threads)
                runs.append((time.perf_counter() - t0, r * threads))
        for scan_s, lo in runs:
            P_oracle = oracle.count_pairs(MINCOV, modulus, lo, lo + threads)
            sel = (np.arange(N) % modulus >= lo) & (np.arange(N) % modulus < lo + threads)
            P_product = int(per_site[sel].sum())
            if P_oracle != P_product:
                raise RuntimeError(f"pair-test count of the CPU sample: oracle {P_oracle}, product filters {P_product}")
            out.append({"value": P_oracle / scan_s, "unit": "pair tests/s", "cores": threads, "kind": kind,
                        "sample": f"row sites ii % {modulus} in [{lo},{lo + threads}) of the whole {workload} MSA ({R} x {N}), "
                                  f"-c {MINCOV}, {threads} threads = {threads}/{modulus} of the rows; pair tests counted by the "
                                  f"oracle's instrumented loop = product filters ({P_oracle})",
                        "pair_tests": P_oracle, "seconds": scan_s, "extrapolation_factor": modulus / threads,
                        "pair_tests_whole_msa": P_total})
    oracle.close()
    return out


def cliquer_queries(np, pk, R, nq):
    """the queries a Group_Refinement pass issues are the groups with MaxCorrs > cutoff (RepeatResolver.c:1647): minor
    groups of a plausible size; here every len/nq-th group with 30 < size < R/3"""
    gs, _ = pk.sizes()
    cand = np.flatnonzero((gs > 30) & (gs < R // 3))
    return cand[::max(1, len(cand) // nq)][:nq].astype(np.int32)


def cliquer_cpu_sample(np, codes, queries, seconds_target):
    """The oracle's Cliquer (oracle/maxcorr_oracle.c, pinned against the unmodified RepeatResolver.c) on the box's host
    cores, on a bounded sample of the same queries: one query per task, all cores."""
    import oracle_lib as O
    from concurrent.futures import ThreadPoolExecutor
    cores = os.cpu_count() or 1
    o = O.Oracle.from_codes(codes)
    t0 = time.perf_counter()
    o.cliquer(int(queries[0]), MINCOV, 30, 3.0)
    one = max(time.perf_counter() - t0, 1e-4)
    n = int(max(cores, min(len(queries), seconds_target * cores / one)))
    sample = [int(q) for q in queries[np.linspace(0, len(queries) - 1, n).astype(np.int64)]]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        sizes = list(ex.map(lambda q: len(o.cliquer(q, MINCOV, 30, 3.0)[0]), sample))
    dt = time.perf_counter() - t0
    pairs = len(sample) * (5 * codes.shape[1] - 1)
    return {"value": pairs / dt, "unit": "candidate pairs/s", "cores": cores, "kind": "port",
            "sample": f"{len(sample)} of the {len(queries)} queries, all candidate groups, -c {MINCOV}, maxclique 30, greedy 3.0",
            "seconds": dt, "mean_clique": sum(sizes) / len(sizes)}


def bench_cliquer(args, rr):
    """--path cliquer: the next row of the scope table.  A step = Cliquer (RepeatResolver.c:1179-1240)
    for every query group of one Group_Refinement pass against all groups of the MSA.  Unit: candidate pairs/s, one
    pair = one (query, candidate group) = one Schnitt of line 1213.  Queries are independent, so N GPUs take
    disjoint slices of the same query list (strong scaling, no collective on the data path)."""
    import numpy as np
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if rank != 0:
            return 0
        g = rr.MsaGen(**WORKLOADS[args.workload], threads=min(32, os.cpu_count() or 8))
        codes = g.codes()
        gs = np.stack([(codes == k).sum(0) for k in range(5)], 1).reshape(-1)
        cand = np.flatnonzero((gs > 30) & (gs < g.rows // 3))
        queries = cand[::max(1, len(cand) // args.queries)][:args.queries].astype(np.int32)
        vals = [cliquer_cpu_sample(np, codes, queries, args.cpu_seconds) for _ in range(args.warmup + args.steps)][args.warmup:]
        v = sum(x["value"] for x in vals) / len(vals)
        print(json.dumps({"impl": "reference", "metric": "Cliquer candidate pairs/sec (Group_Refinement)", "value": v,
                          "unit": "candidate pairs/s", "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup,
                          "ms_per_step": 1e3 * sum(x["seconds"] for x in vals) / len(vals), "higher_is_better": True,
                          "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                          "config": {"workload": args.workload, "path": "cliquer", "rows": g.rows, "cols": g.cols,
                                     "queries": int(len(queries))},
                          "cpu_baseline": {k: (v if k == "value" else vals[0][k]) for k in ("value", "unit", "cores", "kind", "sample")},
                          "e2e": {"value": v, "unit": "candidate pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return 0
    if rr.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: there is no CPU path")
    if world > 1:                                    # torch only as the process-group plumbing of the N > 1 launch
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():                                   # rr_cliquer_batch returns with its stream drained
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()

    def allred(x, op):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op={"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX}[op])
        return float(t.item())

    g, msa = make_msa(rr, args.workload)
    R, N = g.rows, g.cols
    pk = rr.Packed(msa, local_rank)
    queries = cliquer_queries(np, pk, R, args.queries)
    mine = queries[rank::world]                      # independent queries: a cyclic slice per GPU
    if world > 1:                                    # + one all-gather of the cliques, so that every rank holds all of them
        from repeatresolver_b200.dist import cliquer_over_ranks

        def step():
            return cliquer_over_ranks(pk.cliquer_batch, queries, maxclique=30, mincov=MINCOV, greedy=3.0)
    else:
        def step():
            return pk.cliquer_batch(queries, MINCOV, 30, 3.0)
    st = None
    for _ in range(args.warmup):
        st = step()[3]
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = rr.launch_count()
    k_ms, t0 = [], time.perf_counter()
    for _ in range(args.steps):
        members, scores, n, st = step()              # host query list in, cliques out: the C-ABI call
        k_ms.append(st["kernel_ms"])
    call_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    launches = rr.launch_count() - launches0
    barrier()
    clocks = sampler.finish()
    pairs = int(allred(st["pairs"], "sum"))
    kernel_ms = allred(sum(k_ms) / len(k_ms), "max")
    call_ms = allred(call_ms, "max")
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    peaks, peak_src = load_peaks()
    words = (R + 31) // 32
    achieved = pairs * words / (kernel_ms * 1e-3) / 1e12
    peak = 148 * 16 * (clocks["sm_mhz"] or 1965) * 1e6 / 1e12 * world
    w32 = 4 * ((R + 127) // 128)
    roof = {"bound": "popc-issue", "achieved": achieved, "peak": peak, "unit": "Tpopc32/s", "frac": achieved / peak, "traffic": None,
            "ops": "32-bit AND+POPC of the Schnitt at RepeatResolver.c:1213, ceil(R/32) per candidate pair (the reference touches "
                   "every word); words outside the shared coverage of query and candidate site are skipped exactly, so the "
                   "executed count is lower",
            "peak_source": "148 SMs x 16 POPC lanes/clk x median SM clock under load",
            "hbm_floor_bytes_per_step": 6 * N * w32 * 4,
            "hbm_floor_ms": 6 * N * w32 * 4 / (peaks["hbm_gbs"] * 1e9) * 1e3}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cliquer_cpu_sample(np, g.codes(), queries, args.cpu_seconds)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    # the step that follows Cliquer in Group_Refinement (1662-1664): CliqueGroup + CliqueCoverage of every clique found
    # (rr_clique_groups), cutoff = a third of the clique; HBM-bound: every member's bitset word is read once per kind
    cg = None
    if world == 1:
        cut = np.maximum(n // 3, 1).astype(np.int32)
        for _ in range(2):
            pk.clique_groups(members, cut)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            Gq, Vq = pk.clique_groups(members, cut)
        cg_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        cg_bytes = int(2 * (int(n.sum()) * w32 * 4 + len(n) * w32 * 4 * 2 + len(n) * (R // 64 + 1) * 8))
        cg = {"cliques": int(len(n)), "members": int(n.sum()), "ms_per_call": cg_ms, "algorithmic_bytes": cg_bytes,
              "gbs": cg_bytes / (cg_ms * 1e-3) / 1e9, "frac_of_hbm_peak": cg_bytes / (cg_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
              "note": "wall time of the C-ABI call (uploads of the clique lists, 4 kernels, download of the bitsets, host sync)"}
        if not args.no_cpu_baseline:
            import oracle_lib as O
            codes_h = g.codes()
            t0 = time.perf_counter()
            k_cpu = 0
            while time.perf_counter() - t0 < min(args.cpu_seconds, 3.0) and k_cpu < len(n):
                O.clique_group(codes_h, members[k_cpu, :n[k_cpu]], int(cut[k_cpu]))
                O.clique_coverage(codes_h, members[k_cpu, :n[k_cpu]], int(cut[k_cpu]))
                k_cpu += 1
            cg["cpu_port_cliques_per_s"] = k_cpu / (time.perf_counter() - t0)
            cg["gpu_cliques_per_s"] = len(n) / (cg_ms * 1e-3)
    # Group_Refinement as a whole (1634-1693) for the same query groups through rr_group_refinement: Cliquer batch, Sizes,
    # Dropoff_Cutoff on the device's member counts, CliqueGroup + CliqueCoverage at that cutoff
    gr = None
    if world == 1:
        Mq = np.zeros(5 * N)
        Mq[queries] = 50.0
        for _ in range(2):
            res = pk.group_refinement(Mq, 3.0, MINCOV, 30, 3.0)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = pk.group_refinement(Mq, 3.0, MINCOV, 30, 3.0)
        gr_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        assert np.array_equal(res["Cliques"], members) and np.array_equal(np.sort(queries), res["groups"])
        gr = {"groups_above_cutoff": int(len(res["groups"])), "refined": int((res["Sizes"] > 5).sum()), "ms_per_call": gr_ms,
              "groups_per_s": len(res["groups"]) / (gr_ms * 1e-3),
              "cutoff_histogram": {str(int(c)): int(k) for c, k in zip(*np.unique(res["Cutoffs"][res["Sizes"] > 5], return_counts=True))},
              "note": "wall time of the C-ABI call: Cliquer batch + member-count kernel + cutoff rule on the host + CliqueGroup / CliqueCoverage"}
        if not args.no_cpu_baseline:
            import oracle_lib as O
            codes_h = g.codes()
            o_h = O.Oracle.from_codes(codes_h)
            t0 = time.perf_counter()
            k_cpu = 0
            while time.perf_counter() - t0 < min(args.cpu_seconds, 5.0) and k_cpu < len(queries):
                Mk = np.zeros(5 * N)
                Mk[queries[k_cpu]] = 50.0
                O.group_refinement(o_h, codes_h, Mk, 3.0, MINCOV, 30, 3.0)
                k_cpu += 1
            gr["cpu_port_groups_per_s"] = k_cpu / (time.perf_counter() - t0)
            gr["cpu_port_cores"] = 1
    print(json.dumps({"metric": "Cliquer candidate pairs/sec (Group_Refinement)", "value": pairs / (kernel_ms * 1e-3),
                      "unit": "candidate pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                      "dtype": "u32 bitset counts, f64 score", "data": "synthetic",
                      "config": {"workload": args.workload, "path": "cliquer", "rows": R, "cols": N, "queries": int(len(queries)),
                                 "mincov": MINCOV, "maxclique": 30, "greedy": 3.0,
                                 "l2": "candidate bitsets larger than L2 at config 2; smaller workloads are L2-resident"},
                      "candidates": st["candidates"], "hits": st["hits"], "host_evals": st["host_evals"],
                      "mean_clique": float(n.mean()) if len(n) else 0.0, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
                      "clique_groups": cg, "group_refinement": gr,
                      "e2e": {"value": pairs / (call_ms * 1e-3), "unit": "candidate pairs/s", "ms_per_step": call_ms,
                              "h2d_bytes_per_step": int(4 * len(mine)), "d2h_bytes_per_step": int(32 * st["hits"] + 16)},
                      "gpu_launches": launches}))
    if world > 1:
        dist.destroy_process_group()
    return 0


def bench_rows(args, rr):
    """--path relvars | kmeans: the rows after Cliquer on one GPU.  The read partition is the reads'
    symbol at the most significant site of a scan of the same MSA (what the first split of RepeatResolver's clustering sees).
    relvars: a step = Relative_Vars (RepeatResolver.c:2424-2493) for every part, on the packed MSA with the part as a mask
             (rr_relative_vars_packed); unit: group pairs tested/s (one pair = one Triple_Schnitt + score of line 2465).
    kmeans:  a step = Kmeans (2604-2821) for every part on the groups Relative_Vars selected; unit: read pairs compared/s
             (GrMatch calls of the two sweeps, 2 x anzahl^2 per part).
    CPU baseline: the oracle port (pinned against the unmodified RepeatResolver.c) on one core, on the smallest part."""
    import numpy as np
    import oracle_lib as O
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if rr.device_count() < 1 and args.impl != "reference":
        raise SystemExit("bench.py needs a CUDA device: there is no CPU path")
    g, msa = make_msa(rr, args.workload)
    codes = msa.cells()
    R, N = codes.shape
    CUTOFF, MINGROUP = 3.0, 8

    def cpu_leg(M, ut, parts, pk):
        """bounded CPU sample: the smallest part, and for Relative_Vars only the groups of a window of columns (MaxCorrs zeroed
        outside it) sized for a few million group pairs - each costs the oracle two hypergeometric tails"""
        o = O.Oracle.from_codes(codes)
        u_no = min(parts, key=lambda u: int((ut == u).sum()))
        width = N
        Mw = M
        while True:
            Mw = np.where(np.arange(5 * N) // 5 < width, M, 0.0)
            vw, pw = pk.relative_vars(ut, u_no, Mw, CUTOFF, MINGROUP, with_pairs=True)
            if pw <= 3_000_000 or width <= 200:
                break
            width = max(200, width // 2)
        t0 = time.perf_counter()
        v = o.relative_vars(ut, u_no, Mw, CUTOFF, MINGROUP)
        t_rel = time.perf_counter() - t0
        assert list(v) == list(vw), "device selection differs from the oracle's"
        t0 = time.perf_counter()
        o.kmeans(ut, u_no, v, MINGROUP)
        t_km = time.perf_counter() - t0
        return u_no, width, pw, len(v), t_rel, t_km

    if args.impl == "reference":
        # partition from the oracle's own scan is too slow at bench sizes: the sample MSA is scanned by the library when a GPU
        # is present, else the reference arm is unavailable
        if rr.device_count() < 1:
            print(json.dumps({"impl": "reference", "unavailable": "the read partition of this path comes from a scan; no GPU here"}))
            return 0
    pk = rr.Packed(msa, 0)
    pk.scan(mincov=MINCOV)
    M, A = pk.fetch()
    site = int(np.argmax(M)) // 5
    ut = codes[:, site].astype(np.int32)
    parts = [u for u in sorted(set(int(x) for x in ut)) if int((ut == u).sum()) >= 40]
    sel = {u: pk.relative_vars(ut, u, M, CUTOFF, MINGROUP, with_pairs=True) for u in parts}
    pairs_rel = sum(p for _, p in sel.values())
    sizes = {u: int((ut == u).sum()) for u in parts}
    pairs_km = sum(2 * sizes[u] * sizes[u] for u in parts if len(sel[u][0]))
    words_rel = pairs_rel * ((R + 31) // 32)
    scv = {u: len(sel[u][0]) // 64 + 1 for u in parts}
    words_km = sum(2 * sizes[u] * sizes[u] * 2 * scv[u] for u in parts if len(sel[u][0]))   # 64-bit XOR+POPC = 2 x 32-bit

    def step():
        if args.path == "relvars":
            for u in parts:
                pk.relative_vars(ut, u, M, CUTOFF, MINGROUP)
        else:
            for u in parts:
                if len(sel[u][0]):
                    rr.Kmeans(msa, ut, u, sel[u][0], MINGROUP)

    cpu = None
    if args.impl == "reference" or not args.no_cpu_baseline:
        u_small, width, pw, nv, t_rel, t_km = cpu_leg(M, ut, parts, pk)
        cpu_pairs = pw if args.path == "relvars" else 2 * sizes[u_small] ** 2
        cpu = {"value": cpu_pairs / max(t_rel if args.path == "relvars" else t_km, 1e-9), "unit": "pairs/s", "cores": 1, "kind": "port",
               "sample": f"part {u_small} ({sizes[u_small]} reads) of the partition at site {site}, groups of columns [0, {width}) "
                         f"({pw} group pairs, {nv} groups selected), cutoff {CUTOFF}, mingroup {MINGROUP}"}
    metric = "Relative_Vars group pairs/sec" if args.path == "relvars" else "Kmeans read pairs/sec"
    if args.impl == "reference":
        print(json.dumps({"impl": "reference", "metric": metric, "value": cpu["value"], "unit": "pairs/s", "n_gpus": args.gpus, "steps": 1,
                          "warmup": 0, "ms_per_step": 1e3 * (t_rel if args.path == "relvars" else t_km), "higher_is_better": True,
                          "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                          "config": {"workload": args.workload, "path": args.path, "rows": R, "cols": N}, "cpu_baseline": cpu,
                          "e2e": {"value": cpu["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return 0
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(0)
    sampler.start()
    launches0 = rr.launch_count()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    ms = (time.perf_counter() - t0) * 1e3 / args.steps           # the C-ABI calls return with their streams drained
    launches = rr.launch_count() - launches0
    clocks = sampler.finish()
    pairs = pairs_rel if args.path == "relvars" else pairs_km
    if args.path == "relvars":        # per part: the partition up, the selected groups back
        h2d, d2h = int(4 * R * len(parts)), int(4 * 5 * N * len(parts))
    else:                             # per part: its rows and group ids up; signatures, centroids, first assignment back (+ the
        km = [u for u in parts if len(sel[u][0])]                            # dissolution's score table, reads x candidate clusters)
        h2d = int(sum(sizes[u] * N + 4 * len(sel[u][0]) for u in km))
        d2h = int(sum(2 * sizes[u] * scv[u] * 8 + 4 * sizes[u] for u in km))
    words = words_rel if args.path == "relvars" else words_km
    peak = 148 * 16 * (clocks["sm_mhz"] or 1965) * 1e6 / 1e12
    achieved = words / (ms * 1e-3) / 1e12
    print(json.dumps({"metric": metric, "value": pairs / (ms * 1e-3), "unit": "pairs/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                      "dtype": "u32 bitset counts, f64 score" if args.path == "relvars" else "u64 signatures, integer", "data": "synthetic",
                      "config": {"workload": args.workload, "path": args.path, "rows": R, "cols": N, "partition_site": site,
                                 "parts": {str(u): {"reads": sizes[u], "vars": int(len(sel[u][0]))} for u in parts}, "cutoff": CUTOFF, "mingroup": MINGROUP,
                                 "timing": "wall clock of the C-ABI calls (host selection / signatures / dissolution included)"},
                      "clocks": clocks,
                      "roofline": {"bound": "popc-issue", "achieved": achieved, "peak": peak, "unit": "Tpopc32/s", "frac": achieved / peak, "traffic": None,
                                   "ops": "32-bit AND+POPC per pair: ceil(R/32) (relvars, whole-MSA bitsets under the part's mask) / "
                                          "2 x (vars/64+1) per read pair and sweep (kmeans)",
                                   "peak_source": "148 SMs x 16 POPC lanes/clk x median SM clock under load"},
                      "cpu_baseline": cpu, "e2e": {"value": pairs / (ms * 1e-3), "unit": "pairs/s", "ms_per_step": ms,
                                                   "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                      "gpu_launches": launches}))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--path", default="scan", choices=["scan", "cliquer", "relvars", "kmeans"],
                    help="scan = the north-star hot path (default); cliquer / relvars / kmeans = the next scope rows (SURVEY.md 8f, 2-4)")
    ap.add_argument("--queries", type=int, default=1024, help="--path cliquer: query groups per step")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: Tree_1perc_30000 (BASELINE.json configs[1]) for --path scan / cliquer, Tree_1perc_10000_25 for "
                         "relvars / kmeans (their CPU legs and the all-pairs selection of whole parts take minutes at config-2 size)")
    ap.add_argument("--variant", default="auto", choices=["auto", "bitset", "umma", "umma_f4", "umma_mxf4"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e-file", dest="e2e_file", action="store_false", help="skip the run of the drop-in program")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.workload is None:
        args.workload = "Tree_1perc_10000_25" if args.path in ("relvars", "kmeans") else "Tree_1perc_30000"

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    import repeatresolver_b200 as rr
    if args.path == "cliquer":
        return bench_cliquer(args, rr)
    if args.path in ("relvars", "kmeans"):
        return bench_rows(args, rr)
    from repeatresolver_b200.dist import merge_over_ranks, pack_over_ranks, scan_part

    if args.impl == "reference":
        if rank != 0:
            return 0
        g, msa = make_msa(rr, args.workload)
        with tempfile.TemporaryDirectory() as d:
            vals = reference_samples(rr, g, msa, args.workload, args.cpu_seconds, d, reps=args.warmup + args.steps)[args.warmup:]
        tot_p = sum(v["pair_tests"] for v in vals)
        tot_s = sum(v["seconds"] for v in vals)
        v = tot_p / tot_s
        line = {"impl": "reference", "metric": "site-group pair tests/sec (MaxCorrelation)", "value": v,
                "unit": "pair tests/s", "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup,
                "ms_per_step": 1e3 * tot_s / len(vals), "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": {"workload": args.workload, "rows": g.rows, "cols": g.cols, "mincov": MINCOV},
                "cpu_baseline": {"value": v, "unit": "pair tests/s", "cores": vals[0]["cores"], "kind": vals[0]["kind"],
                                 "sample": vals[0]["sample"]},
                "e2e": {"value": v, "unit": "pair tests/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist

    if rr.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: there is no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    g, msa = make_msa(rr, args.workload)
    R, N = g.rows, g.cols
    pk = rr.Packed(msa, local_rank)
    variant = args.variant

    # ---- value: inputs resident in HBM ------------------------------------------------------
    st = None
    for _ in range(args.warmup):
        st = scan_part(pk, MINCOV, variant)
    barrier()
    sampler = ClockSampler(local_rank)
    if not os.environ.get("BENCH_NO_SAMPLER"):
        sampler.start()
    launches0 = rr.launch_count()
    kernel_ms = []
    pk.timer_start()
    for _ in range(args.steps):
        st = scan_part(pk, MINCOV, variant)   # N > 1: seeding pass, all-reduce MAX of thresholds, full pass
        kernel_ms.append(st["kernel_ms"])
    loop_ms = pk.timer_stop()
    launches = rr.launch_count() - launches0
    barrier()
    clocks = sampler.finish()
    loop_ms = allmax(loop_ms)
    P_total = int(allsum(st["pair_tests"]))
    executed_total = allsum(st["executed_ops"])          # every rank's own part, summed
    k_ms = allmax(sum(kernel_ms) / len(kernel_ms))
    ms_per_step = loop_ms / args.steps
    value = P_total / (ms_per_step * 1e-3)
    variant_used = rr.VARIANT_NAMES[st["variant"]]

    # ---- e2e: host buffers through the C ABI -------------------------------------------------
    # every rank holds the cell matrix in page-locked host memory; a step = upload + device pack + scan + D2H + merge.
    # With N ranks the upload and the packing are shared (rank r pushes R/N rows through PCIe, the packed bitsets are merged
    # with one NCCL all-reduce over NVLink: dist.pack_over_ranks); the bytes below are per rank.
    G = 5 * N
    e2e_ms = []
    rows_mine = R * (rank + 1) // world - R * rank // world
    h2d = rows_mine * N + 4 * R + 8 * (R + 2)
    d2h = 16 * G + 4 * 3 * rows_mine + 4 * 6 * N
    for s in range(args.e2e_steps + 1):
        barrier()
        t0 = time.perf_counter()
        pk2 = pack_over_ranks(msa, local_rank)                # H2D (pinned) of this rank's rows + device pack (+ all-reduce OR)
        scan_part(pk2, MINCOV, variant)
        M, A = pk2.fetch()                                    # D2H of the per-group result
        M, A = merge_over_ranks(M, A)                         # element-wise max over ranks (882-891)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        pk2.close()
        if s > 0:
            e2e_ms.append(allmax(dt * 1e3))
    e2e_value = P_total / (sum(e2e_ms) / len(e2e_ms) * 1e-3) if e2e_ms else None

    # ---- e2e_file: the drop-in program itself, text file in, MaxCorrsOf_* out (parse, upload, pack, scan on N GPUs, host
    # finalisation, "%f" text), wall clock of a fresh process; run by rank 0 while the other ranks wait
    e2e_file = None
    if args.e2e_file:
        barrier()
        if rank == 0:
            exe = os.path.join(ROOT, "repeatresolver_b200", "bin", "MaxCorrelation")
            with tempfile.TemporaryDirectory() as d:
                g.write(os.path.join(d, "MSAreal"))
                walls = []
                for _ in range(2):                            # the first run also warms the page cache
                    t0 = time.perf_counter()
                    r = subprocess.run([exe, "MSAreal", "-c", str(MINCOV), "-p", str(world)], cwd=d, capture_output=True, text=True)
                    walls.append(time.perf_counter() - t0)
                    if r.returncode != 0 or not os.path.exists(os.path.join(d, "MaxCorrsOf_MSAreal")):
                        raise SystemExit("bin/MaxCorrelation failed: " + r.stderr[-400:])
                e2e_file = {"value": P_total / walls[-1], "unit": "pair tests/s", "wall_s": walls[-1], "first_run_wall_s": walls[0],
                            "command": f"bin/MaxCorrelation MSAreal -c {MINCOV} -p {world}",
                            "input_bytes": os.path.getsize(os.path.join(d, "MSAreal")),
                            "output_bytes": os.path.getsize(os.path.join(d, "MaxCorrsOf_MSAreal")),
                            "includes": "process start, CUDA start-up, text parse, upload, pack, scan, host finalisation, text output"}
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel -------------------------------------------------------
    peaks, peak_src = load_peaks()
    if variant_used in ("umma", "umma_f4", "umma_mxf4"):
        # Hardware fraction: the MACs the kernel EXECUTES (padding, masked entries and K skipping included) against the
        # MEASURED rate of the tensor pipe for the MMA kind it uses - a bare tcgen05.mma loop of the same kind, tile shape
        # and operand layout in this library (rr_debug_mma_peak), taken here on the same GPU right after the timed loop.
        # Beside it the algorithmic figure of SURVEY.md 8d: 2*R 8-bit-rate ops per pair test (the reference touches all
        # R/64+1 words per intersection) against the dense INT8 peak measured with torch._int_mm 8192^3 (cuBLASLt).
        psamp = ClockSampler(local_rank)
        psamp.start()
        pipe = rr.debug.mma_peak(variant_used, local_rank)
        pipe_i8 = rr.debug.mma_peak("umma", local_rank) if variant_used != "umma" else pipe
        a8 = torch.randint(-2, 3, (8192, 8192), dtype=torch.int8, device="cuda")
        b8 = torch.randint(-2, 3, (8192, 8192), dtype=torch.int8, device="cuda")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        int8_ms = []
        for _ in range(12):
            e0.record()
            torch._int_mm(a8, b8)
            e1.record()
            e1.synchronize()
            int8_ms.append(e0.elapsed_time(e1))
        del a8, b8
        pclk = psamp.finish()
        int8_tops = 2.0 * 8192 ** 3 / (min(int8_ms[2:]) * 1e-3) / 1e12
        algo = 2.0 * R * P_total
        executed = executed_total
        achieved = executed / (k_ms * 1e-3) / 1e12
        peak = pipe["tflops"] * world
        roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": TRAFFIC.get(args.workload),
                "ops": "EXECUTED tensor-core multiply-adds x 2 (every K block issued: padding rows, masked entries and the exact "
                       "K-range skipping included) / kernel time, against the measured rate of the pipe used",
                "pipe": {"umma": "tcgen05.mma kind::i8 (int32 accumulate)", "umma_f4": "tcgen05.mma kind::f8f6f4 on e2m1 (fp32 accumulate)",
                         "umma_mxf4": "tcgen05.mma kind::mxf4.block_scale on packed e2m1, unit scales (fp32 accumulate)"}[variant_used],
                "peak_source": f"measured here: bare tcgen05.mma loop of this kind, M=128 N=240, all SMs (rr_debug_mma_peak): "
                               f"{pipe['tflops']:.1f} TFLOP/s per GPU; SM clock median {pclk['sm_mhz']} MHz while measuring",
                "peaks_measured": {"pipe_used_tflops": pipe["tflops"], "tcgen05_i8_loop_tops": pipe_i8["tflops"],
                                   "int8_cublaslt_8192_tops": int8_tops, "bf16_cublas_burst_tflops": peaks["bf16_tflops"],
                                   "bf16_source": peak_src, "clocks_while_measuring": pclk},
                "executed_ops_per_step": executed, "executed_frac_of_algorithmic": executed / algo if algo else None,
                "algorithmic": {"ops_per_step": algo, "tflops": algo / (k_ms * 1e-3) / 1e12,
                                "frac_of_int8_peak": algo / (k_ms * 1e-3) / 1e12 / (int8_tops * world),
                                "note": "2*R 8-bit-rate ops per pair test x pair tests / kernel time: a speed metric (K skipping and "
                                        "the 4-bit pipe make it exceed what an INT8 GEMM of the full shape could reach), not a "
                                        "hardware fraction"}}
    else:
        # AND+POPC variant: issue-bound on the POPC pipe (16 lanes/clk/SM); reported against that peak
        words = P_total * ((R + 31) // 32)
        achieved = words / (k_ms * 1e-3) / 1e12
        peak = 148 * 16 * (clocks["sm_mhz"] or 1965) * 1e6 / 1e12 * world
        roof = {"bound": "popc-issue", "achieved": achieved, "peak": peak, "unit": "Tpopc32/s", "frac": achieved / peak,
                "traffic": None, "ops": "32-bit AND+POPC, ceil(R/32) per pair test",
                "peak_source": "148 SMs x 16 POPC lanes/clk x median SM clock under load"}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        with tempfile.TemporaryDirectory() as d:
            cpu = reference_samples(rr, g, msa, args.workload, args.cpu_seconds, d)[0]
            if cpu["pair_tests_whole_msa"] != P_total:
                raise SystemExit(f"pair tests: device {P_total}, host filters {cpu['pair_tests_whole_msa']}")
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "extrapolation_factor")}

    line = {"metric": "site-group pair tests/sec (MaxCorrelation)", "value": value, "unit": "pair tests/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": {"umma": "int8->int32 counts, f64 score", "umma_f4": "e2m1(0/1)->f32 counts, f64 score",
                      "umma_mxf4": "e2m1(0/1) x unit block scale ->f32 counts, f64 score"}.get(
                variant_used, "u32 bitset counts, f64 score"),
            "data": "synthetic",
            "config": {"workload": args.workload, "rows": R, "cols": N, "mincov": MINCOV, "variant": variant_used,
                       "pair_tests": P_total, "l2": "inputs larger than L2 (packed operands >> 126 MB)",
                       "partition": f"{world} cost-balanced contiguous row ranges"},
            "kernel_ms": k_ms, "exact_evals": st["exact_evals"], "bound_evals": st["bound_evals"],
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "pair tests/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": sum(e2e_ms) / len(e2e_ms) if e2e_ms else None,
                    "path": "C ABI: rr_pack (N > 1: rr_pack_rows + NCCL all-reduce OR of the bitsets + rr_pack_finish), rr_scan, "
                            "rr_scan_fetch, max-merge"},
            "e2e_file": e2e_file,
            "gpu_launches": launches}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
