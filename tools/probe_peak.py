"""bare tcgen05.mma rates (rr_debug_mma_peak_shape): cta_group 1 (M=128, N=240) against CTA pairs at several N"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import repeatresolver_b200 as rr
for v in ("umma_mxf4", "umma_f4", "umma"):
    print(v, "cta_group 1 N=240", rr.debug.mma_peak(v), flush=True)
    for n in (256, 240, 224, 192, 128):
        print(v, "cta_group 2 N=%d" % n, rr.debug.mma_peak_shape(v, 2, n), flush=True)
