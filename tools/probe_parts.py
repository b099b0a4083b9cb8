import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import repeatresolver_b200 as rr
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "Tree_1perc_30000"
parts = int(sys.argv[2]) if len(sys.argv) > 2 else 2
g, msa = bench.make_msa(rr, wl)
pk = rr.Packed(msa, 0)
for p in range(parts):
    for rep in range(2):
        t0 = time.perf_counter()
        if os.environ.get("EXCHANGE"):
            st = pk.scan(mincov=30, variant="auto", part_index=p, part_count=parts, flags=rr.FLAG_SKIP_SEED) if rep else None
            if st is None:
                # thresholds = full-scan result scaled down is not available here; emulate the exchange with the
                # max over all parts' seeding passes
                import numpy as np
                seeds = []
                for q in range(parts):
                    pk.scan(mincov=30, variant="auto", part_index=q, part_count=parts, flags=rr.FLAG_SEED_ONLY)
                    seeds.append(pk.fetch()[0])
                thr = np.maximum.reduce(seeds)
                st0 = pk.scan(mincov=30, variant="auto", part_index=p, part_count=parts, flags=rr.FLAG_SEED_ONLY)
                print("   seed pass ms", round(st0["kernel_ms"], 2), end="")
                pk.set_thresholds(thr)
                st = pk.scan(mincov=30, variant="auto", part_index=p, part_count=parts, flags=rr.FLAG_SKIP_SEED)
                dt = 0.0
                break
        else:
            st = pk.scan(mincov=30, variant="auto", part_index=p, part_count=parts)
        dt = (time.perf_counter() - t0) * 1e3
    print("part", p, "of", parts, "wall_ms %.1f" % dt, {k: st[k] for k in ("kernel_ms", "prepare_ms", "pair_tests", "exact_evals", "bound_evals", "executed_ops", "work_units")}, flush=True)
