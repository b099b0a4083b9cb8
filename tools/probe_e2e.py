"""where an end-to-end step (host cells -> packed -> scan -> result on the host) spends its time: RR_TRACE=1 makes the library
print one stderr line per host phase.  usage: RR_TRACE=1 python tools/probe_e2e.py [WORKLOAD] [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import repeatresolver_b200 as rr  # noqa: E402
import bench  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "Tree_1perc_30000"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mode = sys.argv[3] if len(sys.argv) > 3 else ""
g, msa = bench.make_msa(rr, wl)
if "torch" in mode:
    import torch
    torch.cuda.set_device(0)
    torch.cuda.synchronize()
if "keep" in mode:
    keep = rr.Packed(msa, 0)
    keep.scan(mincov=30)
if "nvml" in mode:
    smp = bench.ClockSampler(0)
    smp.start()
    time.sleep(0.5)
    print(smp.finish(), file=sys.stderr)
for r in range(reps):
    print(f"--- rep {r}", file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    pk = rr.Packed(msa, 0)
    t1 = time.perf_counter()
    st = pk.scan(mincov=30)
    t2 = time.perf_counter()
    M, A = pk.fetch()
    t3 = time.perf_counter()
    pk.close()
    t4 = time.perf_counter()
    print(f"rep {r}: pack {1e3 * (t1 - t0):.1f} ms, scan call {1e3 * (t2 - t1):.1f} ms (kernel {st['kernel_ms']:.1f}, prepare {st['prepare_ms']:.1f}, "
          f"h2d {st['h2d_ms']:.1f}, pack {st['pack_ms']:.1f}), fetch {1e3 * (t3 - t2):.1f} ms, close {1e3 * (t4 - t3):.1f} ms", flush=True)
