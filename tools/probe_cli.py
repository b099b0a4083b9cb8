"""end-to-end time of the drop-in program on a generated MSAreal text file (parse + pack + H2D + scan + finalize + write).
usage: probe_cli.py [WORKLOAD] [n_gpus] [reps]"""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import repeatresolver_b200 as rr
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "Tree_1perc_30000"
ngpu = sys.argv[2] if len(sys.argv) > 2 else "1"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
g = rr.MsaGen(**bench.WORKLOADS[wl], threads=min(32, os.cpu_count() or 8))
with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
    t0 = time.time()
    g.write(os.path.join(d, "MSAreal"))
    print(f"{wl}: {g.rows} x {g.cols}, text {os.path.getsize(os.path.join(d, 'MSAreal')) / 1e9:.2f} GB written in {time.time() - t0:.1f} s", flush=True)
    exe = os.path.join(ROOT, "repeatresolver_b200", "bin", "MaxCorrelation")
    for r in range(reps):
        t0 = time.time()
        p = subprocess.run([exe, "MSAreal", "-c", "30", "-p", ngpu], cwd=d, env=dict(os.environ, RR_TRACE="1"),
                           capture_output=True, text=True)
        dt = time.time() - t0
        print(f"run {r}: rc={p.returncode} wall {dt:.3f} s", flush=True)
        print(p.stdout[-1500:], p.stderr[-3000:], flush=True)
