#!/usr/bin/env python
"""Local-BA micro-benchmark for kernel work: one prepared cfg4 problem (50+10 KF, 20k landmarks, ~100k observations).
  latency     median device time of ovs_local_ba_run alone on the GPU (one stream)
  throughput  S host threads, each with its own prepared problem, R runs each: BA calls per second
usage: python tools/ba_bench.py [runs=10] [streams=8]"""
import os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openvslam_b200 import optimize, synth, _lib  # noqa: E402

runs = int(sys.argv[1]) if len(sys.argv) > 1 else 10
S = int(sys.argv[2]) if len(sys.argv) > 2 else 8
q = synth.ba_problem(50, 10, 20000, model="equirectangular", seed=4)
args = (optimize.camera(**q["cam"]), True, q["poses"], q["fixed"], q["points"], q["obs_kf"], q["obs_lm"], q["obs_xy"], None, q["inv_sigma_sq"])
pbs = [optimize.prepared_local_ba(*args) for _ in range(S)]
if os.environ.get("BA_CLUSTER"):
    for pb in pbs:
        pb.set_cluster_width(int(os.environ["BA_CLUSTER"]))
for pb in pbs:
    pb.run()
us = []
for _ in range(runs):
    st = pbs[0].run()
    us.append(st["device_us"])
print("latency: median %.1f us  min %.1f  (trials %d, iterations %d, final chi2 %.6f, solver %.1f us over %d launches)"
      % (np.median(us), min(us), st["num_trials"], st["num_iterations"], st["final_chi2"], st["solver_us"], st["solver_launches"]))
_lib.lib().ovs_set_wait_mode(2)


def work(pb):
    for _ in range(runs):
        pb.run()
ths = [threading.Thread(target=work, args=(pb,)) for pb in pbs]
t0 = time.perf_counter()
for t in ths:
    t.start()
for t in ths:
    t.join()
dt = time.perf_counter() - t0
print("throughput: %d streams x %d runs in %.3f s = %.1f BA/s  (%.3f ms of GPU per BA)" % (S, runs, dt, S * runs / dt, 1e3 * dt / (S * runs)))
