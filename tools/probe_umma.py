"""timing experiments on the scan kernels (debug flags: see include/rr_debug.h:
0x400 = RR_DEBUG_MMA_ONLY, the producer + MMA pipeline alone).  usage: probe_umma.py WORKLOAD flags,flags,... [variant]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import repeatresolver_b200 as rr
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "Tree_1perc_30000"
flag_list = [int(x, 0) for x in (sys.argv[2] if len(sys.argv) > 2 else "0,0x400").split(",")]
variant = sys.argv[3] if len(sys.argv) > 3 else "umma"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
g, msa = bench.make_msa(rr, wl)
pk = rr.Packed(msa, 0)
for flags in flag_list:
    for _ in range(reps):
        try:
            st = pk.scan(mincov=30, variant=variant, flags=flags)
        except rr.RRError as e:
            if flags & 0x400 and "mismatch" in str(e):
                print("flags", hex(flags), "(counts differ, expected in an experiment)", flush=True)
                continue
            raise
    print("flags", hex(flags), variant, {k: st[k] for k in ("kernel_ms", "prepare_ms", "pair_tests", "exact_evals", "bound_evals", "executed_ops")}, flush=True)
