import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import repeatresolver_b200 as rr
import bench
parts = int(sys.argv[1]) if len(sys.argv) > 1 else 8
g, msa = bench.make_msa(rr, "Tree_1perc_30000")
pk = rr.Packed(msa, 0)
for rep in range(2):
    st = pk.scan(mincov=30, part_index=0, part_count=parts, flags=rr.FLAG_SEED_ONLY)
    print("seed-only", st["kernel_ms"], st["exact_evals"], st["bound_evals"])
