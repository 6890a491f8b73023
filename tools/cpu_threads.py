"""Diagnostic: oracle-port training step time vs torch CPU thread count on this host (picks the reference arm's setting)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

if __name__ == "__main__":
    patch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    print("cpu_count", os.cpu_count(), "sched_affinity", len(os.sched_getaffinity(0)))
    for th in (8, 16, 32, 64, 128):
        if th > (os.cpu_count() or 1):
            break
        bench.cpu_step_time(th, reps=1, patch=patch)
        t = bench.cpu_step_time(th, reps=1, patch=patch)
        print(f"threads {th}: {t:.2f} s per step (patch {patch}^3)", flush=True)
