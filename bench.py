#!/usr/bin/env python
"""bench.py -- frames/s of the OpenVSLAM hot path (extract + match + pose optimisation + local BA)
on synthetic 1920x960 equirectangular frames at 4000 keypoints (BASELINE.json configs[3]).

One "step" = one frame through the hot path on each of the --streams independent camera streams of a GPU
(default 8; every stream owns its handles, CUDA streams and host thread).  Per frame:
  orb_extractor::extract (1920x960, 4000 kp)
  match::robust::brute_force_match against the previous frame's descriptors (4000 x 4000 Hamming)
  pose_optimizer::optimize on 4000 matched landmarks (equirectangular, 4 x 10 LM iterations)
  local_bundle_adjuster::optimize on 50 free + 10 fixed keyframes, 20k landmarks, ~100k observations
Two measurements per run:
  value  device-resident: frames, descriptors and the BA problem already in HBM when the timed
         region starts (ovs_extract_device / *_topk_device / ovs_local_ba_run).
  e2e    through the host-buffer C-ABI entry points a reference caller would use, every
         host<->device copy inside the timed region.
`--impl reference` times the CPU oracle (the restated reference; the real one cannot be built here,
see DESIGN.md) on the same workload with all host threads, as independent streams.
Prints ONE JSON line on rank 0."""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, NKP = 1920, 960, 4000
K_FREE, K_FIXED, N_LM = 50, 10, 20000
METRIC = "frames/sec extract+match+local-BA @1920x960 4000kp"
WORKLOAD = "configs[3]: 1920x960 equirectangular stream, 4000 kp/frame, extract + brute-force match + pose_optimizer + local_bundle_adjuster (50+10 KF / 20k landmarks / ~100k obs)"


def host_cores():
    """CPU cores this process may actually use: the cgroup CPU quota when there is one (the GPU boxes show 128 logical
    CPUs but cap the container at 16 cores), else the affinity mask / cpu count."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            n = min(n, max(1, int(float(q) / float(per) + 0.5)))
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, int(q / per + 0.5)))
        except Exception:
            pass
    return n


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def max_over_ranks(elapsed, device, world):
    """MAX over ranks of a per-rank elapsed time (the contract's timing rule)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([elapsed], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_value(steps, elapsed, world):
    """Whole-job throughput: every rank processed `steps` frames of its own stream (weak scaling)."""
    return world * steps / elapsed


def make_workload(seed, ring_frames):
    from openvslam_b200 import synth
    base = [synth.frame(W, H, seed=seed * 100 + i) for i in range(6)]
    frames = []
    for i in range(ring_frames):
        frames.append(np.ascontiguousarray(np.roll(base[i % 6], 37 * (i // 6), axis=1)))  # equirectangular yaw
    ba = synth.ba_problem(K_FREE, K_FIXED, N_LM, model="equirectangular", seed=seed + 4)
    pose = synth.pose_problem(NKP, model="equirectangular", seed=seed + 3, stereo=False)
    return frames, ba, pose


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu,power.draw"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        def col(i):
            out = []
            for r in self.rows:
                try:
                    out.append(float(r[i]))
                except (ValueError, IndexError):
                    pass
            return out
        util, power = col(6), col(7)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "gpu_util_pct_median": float(np.median(util)) if util else None,
                "power_w_median": float(np.median(power)) if power else None}


class CameraStream:
    """One camera stream: its own extractor / matcher / optimiser handles (each with a private CUDA stream), as the
    reference owns them per tracking / mapping thread.  Streams are independent, so S of them per GPU is the same
    weak-scaling unit as one stream per rank."""

    def __init__(self, sid, local, dev, frames_dev, frames_host, ba, pose, ring, spec=4):
        import torch
        from openvslam_b200 import feature, match, optimize, _lib
        self.sid, self.ring, self.ba, self.pose = sid, ring, ba, pose
        self.L = _lib.lib()
        self._lib = _lib
        self.ext = feature.orb_extractor(feature.orb_params(max_num_keypts=NKP), device=local)
        self.mt = match.robust(lowe_ratio=0.75, device=local)
        self.po = optimize.pose_optimizer(device=local)
        self.cam = optimize.camera(**ba["cam"])
        self.pcam = optimize.camera(**pose["cam"])
        self.ba_args = (ba["poses"], ba["fixed"], ba["points"], ba["obs_kf"], ba["obs_lm"], ba["obs_xy"], None, ba["inv_sigma_sq"])
        self.lba = optimize.local_bundle_adjuster(device=local)
        self.pba = optimize.prepared_local_ba(self.cam, True, *self.ba_args, device=local)
        self.lba.set_speculation(spec); self.pba.set_speculation(spec)
        self.d_frames, self.h_frames = frames_dev, frames_host
        self.cap = self.L.ovs_extractor_max_keypoints(self.ext._h)
        self.d_kps = torch.zeros((2, self.cap, 28), dtype=torch.uint8, device=dev)
        self.d_desc = torch.zeros((2, self.cap, 32), dtype=torch.uint8, device=dev)
        self.d_keys = torch.zeros((self.cap, 8), dtype=torch.int32, device=dev)   # OVS_MATCH_TOPK keys per query
        self.e2e_ms = np.zeros(4)
        self.reset()
        self.n_prev, self.prev_desc = 0, None

    def reset(self):
        self.st = {"match_us": 0.0, "ext_us": np.zeros(8), "ba_us": 0.0, "pose_us": 0.0, "steps": 0,
                   "solver_us": 0.0, "solver_launches": 0, "solver_trials": 0, "reduced_dim": 0}
        self.e2e_ms[:] = 0

    def step_device(self, i):
        i += 11 * self.sid                      # streams walk the shared frame ring at different offsets
        cur = i & 1
        ext, st, L = self.ext, self.st, self.L
        n = ext.extract_device(self.d_frames[i % self.ring].data_ptr(), W, H, W, self.d_kps[cur].data_ptr(), self.d_desc[cur].data_ptr(), self.cap)
        t = ext.last_timings_us()
        if self.n_prev:
            self._lib.check(L.ovs_match_bruteforce_topk_device(self.mt._h, C.c_void_p(self.d_desc[cur].data_ptr()), n,
                                                               C.c_void_p(self.d_desc[cur ^ 1].data_ptr()), self.n_prev, C.c_void_p(self.d_keys.data_ptr())))
            st["match_us"] += self.mt.last_kernel_us()
        self.n_prev = n
        pose = self.pose
        _, _, _, pst = self.po.optimize(self.pcam, True, pose["pts_w"], pose["obs_xy"], None, pose["inv_sigma_sq"], pose["poses"][0])
        bst = self.pba.run()
        st["ext_us"] += np.array(list(t.values()))
        st["pose_us"] += pst["device_us"]; st["ba_us"] += bst["device_us"]; st["steps"] += 1
        st["solver_us"] += bst["solver_us"]; st["solver_launches"] += bst["solver_launches"]; st["solver_trials"] += bst["solver_trials"]
        st["reduced_dim"] = bst["reduced_dim"]
        return n

    def step_host(self, i):
        i += 11 * self.sid
        pose = self.pose
        t0 = time.perf_counter()
        kps, desc = self.ext.extract(self.h_frames[i % self.ring])
        t1 = time.perf_counter()
        if self.prev_desc is not None:
            self.mt.brute_force_match(desc, self.prev_desc)
        self.prev_desc = desc
        t2 = time.perf_counter()
        self.po.optimize(self.pcam, True, pose["pts_w"], pose["obs_xy"], None, pose["inv_sigma_sq"], pose["poses"][0])
        t3 = time.perf_counter()
        self.lba.optimize(self.cam, True, *self.ba_args)
        t4 = time.perf_counter()
        self.e2e_ms += np.array([t1 - t0, t2 - t1, t3 - t2, t4 - t3]) * 1e3
        return len(kps)

    def close(self):
        for h in (self.ext, self.mt, self.po, self.lba, self.pba):
            h.close()


def run_ours(args):
    import torch
    import torch.distributed as dist
    from openvslam_b200 import _lib
    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ring, S = args.ring, max(1, args.streams)
    frames, ba, pose = make_workload(rank, ring)
    # host wait mode: the GPU boxes give the container 16 cores for up to 8 GPUs x S driving threads
    wait = args.wait
    if wait == "auto":   # spin while every driving thread can own a core, yield-poll once they cannot
        wait = "spin" if S * world <= max(1, host_cores() - 2) else "yield"
    _lib.lib().ovs_set_wait_mode({"spin": 0, "block": 1, "yield": 2}[wait])

    # ---- device-resident inputs: ring of frames (> L2), shared read-only by the camera streams of this GPU
    d_frames = torch.empty((ring, H, W), dtype=torch.uint8, device=dev)
    h_frames = torch.empty((ring, H, W), dtype=torch.uint8).pin_memory()
    for i, f in enumerate(frames):
        h_frames[i].copy_(torch.from_numpy(f))
    d_frames.copy_(h_frames)
    torch.cuda.synchronize()
    h_frames_np = h_frames.numpy()
    spec = args.spec if args.spec > 0 else 4   # LM trials per launch sequence (measured at 8 streams: 1 -> 360, 2 -> 345, 3 -> 364, 4 -> 371 frames/s)
    cams = [CameraStream(sid, local, dev, d_frames, h_frames_np, ba, pose, ring, spec) for sid in range(S)]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_all(name, lo, hi):
        """every camera stream runs steps lo..hi-1 of `name` on its own host thread (the C ABI releases the GIL)"""
        errs = []

        def work(cs):
            try:
                torch.cuda.set_device(local)
                fn = getattr(cs, name)
                for i in range(lo, hi):
                    fn(i)
            except Exception as e:   # noqa: BLE001
                errs.append(e)
        if S == 1:
            work(cams[0])
        else:
            ths = [threading.Thread(target=work, args=(cs,)) for cs in cams]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
        if errs:
            raise errs[0]

    event_ms = {}

    def timed(name, steps, warmup, offset):
        run_all(name, offset, offset + warmup)
        for cs in cams:
            cs.reset()
        barrier()
        l0 = _lib.launch_count()
        # CUDA events bracket the region as well (every stream of the device is idle at both records, so the device
        # timeline between them is the region): reported next to the host clock as a cross-check
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        run_all(name, offset + warmup, offset + warmup + steps)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e1.record(); e1.synchronize()
        event_ms[name] = max_over_ranks(e0.elapsed_time(e1) * 1e-3, dev, world) * 1e3
        launches = _lib.launch_count() - l0
        return max_over_ranks(dt, dev, world), launches

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    t_dev, launches = timed("step_device", args.steps, args.warmup, 0)
    dev_state = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in cams[0].st.items()}   # per-kernel times: stream 0
    t_e2e, _ = timed("step_host", args.steps, args.warmup, 7)
    e2e_stage = {k: round(float(v) / args.steps, 3) for k, v in zip(("extract", "brute_force_match", "pose_optimizer", "local_ba"), cams[0].e2e_ms)}
    clocks = sampler.stop() if sampler else None

    # ---- per-frame latency of ONE stream alone on the GPU (what a live SLAM session sees): spin waits, BA launch
    #      sequences replayed as CUDA graphs.  Reported next to the throughput figures, not part of `value`.
    latency = None
    if not args.no_latency:
        cs = cams[0]
        _lib.lib().ovs_set_wait_mode(0)
        cs.pba.set_graphs(True); cs.lba.set_graphs(True)
        cs.pba.set_speculation(4); cs.lba.set_speculation(4)
        nlat = max(5, min(args.steps, 20))
        for i in range(3):
            cs.step_device(100 + i)
        barrier()
        t0 = time.perf_counter()
        for i in range(nlat):
            cs.step_device(103 + i)
        torch.cuda.synchronize()
        lat_dev = (time.perf_counter() - t0) / nlat
        for i in range(3):
            cs.step_host(100 + i)
        barrier()
        t0 = time.perf_counter()
        for i in range(nlat):
            cs.step_host(103 + i)
        torch.cuda.synchronize()
        lat_host = (time.perf_counter() - t0) / nlat
        lat_dev = max_over_ranks(lat_dev, dev, world); lat_host = max_over_ranks(lat_host, dev, world)
        latency = {"streams": 1, "frames": nlat, "ms_per_frame_device_resident": round(1e3 * lat_dev, 4), "ms_per_frame_e2e": round(1e3 * lat_host, 4),
                   "host_wait": "spin", "cuda_graphs": True}

    value = aggregate_value(args.steps * S, t_dev, world)
    e2e = aggregate_value(args.steps * S, t_e2e, world)
    h2d = W * H + 2 * NKP * 32 + NKP * (24 + 8 + 4) + 96 + len(ba["obs_kf"]) * 24 + (K_FREE + K_FIXED) * 100 + N_LM * 24
    d2h = NKP * (28 + 32) + NKP * 16 + NKP + 96 + (K_FREE + K_FIXED) * 96 + N_LM * 24 + len(ba["obs_kf"])

    out = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        s = max(dev_state["steps"], 1)
        ext_us = dev_state["ext_us"] / s
        names = ("upload", "pyramid", "fast_score", "cell_nms_compact", "host_tree", "orient_describe", "download", "total_wall")
        stages = {"extract_" + n: round(float(v), 1) for n, v in zip(names, ext_us)}
        stages.update(match_hamming_kernel=round(dev_state["match_us"] / max(s - 1, 1), 1), pose_optimizer_kernel=round(dev_state["pose_us"] / s, 1),
                      local_ba_device=round(dev_state["ba_us"] / s, 1))
        # Hamming kernel (the kernel BASELINE.json's metric names): algorithmic bytes = (N + M) * 32 + N * 8
        ham_us = dev_state["match_us"] / max(s - 1, 1)
        ham_bytes = (NKP + NKP) * 32 + NKP * 8
        ham_gbs = ham_bytes / (ham_us * 1e-6) / 1e9 if ham_us > 0 else 0.0
        # FAST score kernel: reads the pyramid once and writes the score map once
        lvl = [(1920, 960), (1600, 800), (1333, 667), (1111, 556), (926, 463), (772, 386), (643, 322), (536, 268)]
        fast_bytes = 2 * sum(w * h for w, h in lvl)
        fast_us = float(ext_us[2])
        fast_gbs = fast_bytes / (fast_us * 1e-6) / 1e9 if fast_us > 0 else 0.0
        # Dominant kernel of the step: the cluster Cholesky of the reduced camera system (FP64; DMMA trailing update and
        # panel GEMM).  Algorithmic flops per factorised system: n^3/3 (factorisation) + 2 n^2 (the two triangular solves).
        nred = int(dev_state["reduced_dim"])
        sol_launches = max(int(dev_state["solver_launches"]), 1)
        sol_us = dev_state["solver_us"] / sol_launches
        sol_flops = (nred ** 3 / 3.0 + 2.0 * nred ** 2) * dev_state["solver_trials"] / sol_launches
        sol_tf = sol_flops / (sol_us * 1e-6) / 1e12 if sol_us > 0 else 0.0
        # FP64 tensor peak: MEASURED_PEAKS.json holds no FP64 figure; tools/probe/fp64_probe.cu measured 17.3 clk per
        # independent m8n8k4 DMMA per warp scheduler on this pool's B200 = 59 FMA/clk/SM -> 148 SM x 1.965 GHz x 2
        fp64_peak = 148 * 1.965e9 * (256 / 17.3 * 4) * 2 / 1e12
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_dram_traffic.json")))
        except Exception:
            pass
        out = {
            "metric": METRIC, "value": round(value, 3), "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(1e3 * t_dev / args.steps, 4), "ms_per_step_cuda_events": round(event_ms.get("step_device", 0.0) / args.steps, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8 (extract, Hamming) + f64 (pose optimiser, local BA)", "data": "synthetic (seeded numpy frames and BA graph; no datasets offline)",
            "config": {"workload": WORKLOAD, "streams_per_gpu": S, "lm_speculation_width": spec, "host_wait": wait, "host_cores": host_cores(),
                       "step": "one frame on each of the %d independent camera streams of a GPU (own handles and CUDA streams, one host thread each)" % S,
                       "l2": "frame ring of %d x 1.84 MB = %.0f MB > 126 MB L2" % (ring, ring * W * H / 1e6),
                       "value_path": "device-resident (extract_device, topk_device, prepared local BA)", "e2e_path": "host-buffer C ABI",
                       "timing": "barrier + synchronize on both sides, max over ranks; host clock of the region (every C-ABI call returns with its "
                                 "stream drained) cross-checked by CUDA events recorded while the device is idle (ms_per_step_cuda_events)"},
            "e2e": {"value": round(e2e, 3), "unit": "frames/s", "ms_per_step": round(1e3 * t_e2e / args.steps, 4),
                    "ms_per_step_cuda_events": round(event_ms.get("step_host", 0.0) / args.steps, 4),
                    "h2d_bytes_per_step": int(h2d) * S, "d2h_bytes_per_step": int(d2h) * S, "stage_ms_per_frame_stream0": e2e_stage},
            "single_stream_latency": latency,
            "gpu_launches": int(launches),
            "stage_us_per_step": stages,
            "roofline": {"kernel": "k_ba_cholesky_solve", "bound": "tensor", "achieved": round(sol_tf, 4), "peak": round(fp64_peak, 1), "unit": "TFLOP/s",
                         "frac": round(sol_tf / fp64_peak, 5), "traffic": traffic.get("k_ba_cholesky_solve"),
                         "peak_source": "FP64 DMMA issue rate measured with tools/probe/fp64_probe.cu (MEASURED_PEAKS.json has no FP64 entry)",
                         "algorithmic_flops_per_launch": round(sol_flops), "avg_launch_us": round(sol_us, 2), "launches_per_step": round(sol_launches / s, 2),
                         "share_of_stream_time": round(dev_state["solver_us"] / s / (1e6 * t_dev / args.steps), 4), "reduced_dim": nred,
                         "systems_per_launch": round(dev_state["solver_trials"] / sol_launches, 2),
                         "note": "latency bound, not throughput bound: n dependent pivots (fma -> shuffle -> rsqrt -> mul, ~120 clk each "
                                 "measured) put a floor of n x 120 clk = %.1f us under every launch" % (nred * 120 / 1.965e3)},
            "roofline_fast_score": {"kernel": "k_fast_score", "bound": "hbm", "achieved": round(fast_gbs, 2), "peak": hbm_peak, "unit": "GB/s",
                                    "frac": round(fast_gbs / hbm_peak, 5), "traffic": traffic.get("k_fast_score"), "peak_source": peak_src,
                                    "algorithmic_bytes_per_launch": fast_bytes},
            "roofline_hamming": {"kernel": "k_hamming_topk+k_topk_merge", "bound": "hbm", "achieved": round(ham_gbs, 3), "peak": hbm_peak, "unit": "GB/s",
                                 "frac": round(ham_gbs / hbm_peak, 7), "algorithmic_bytes_per_launch": ham_bytes, "traffic": traffic.get("k_hamming_topk"),
                                 "operand_stream_gbs_not_hbm": round(NKP * NKP * 64 / (ham_us * 1e-6) / 1e9, 1) if ham_us > 0 else None,
                                 "popc32_per_s_not_hbm": round(8.0 * NKP * NKP / (ham_us * 1e-6), 0) if ham_us > 0 else None},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(frames, ba, pose, threads=1, budget_s=args.cpu_budget)
    for cs in cams:
        cs.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


# ------------------------------------------------------------------------------ CPU oracle legs
def oracle_step(O, frame, prev_desc, ba, pose, P):
    kps, desc, _ = O.extract(frame, P)
    if prev_desc is not None:
        O.robust_brute_force_match(desc, prev_desc, None, 0.75)
    O.pose_optimize(O.camera(**pose["cam"]), True, pose["pts_w"], pose["obs_xy"], None, pose["inv_sigma_sq"], pose["poses"][0])
    O.local_ba(O.camera(**ba["cam"]), True, ba["poses"], ba["fixed"], ba["points"], ba["obs_kf"], ba["obs_lm"], ba["obs_xy"], None, ba["inv_sigma_sq"])
    return desc


def cpu_run(frames, ba, pose, threads, steps_per_thread):
    """`threads` independent streams, each running `steps_per_thread` steps of the oracle."""
    from oracle import oracle as O
    O.build()
    O.lib()
    P = O.params(NKP)

    def worker(tid):
        prev = None
        for s in range(steps_per_thread + 1):  # first step primes prev_desc (untimed share is small and identical per thread)
            prev = oracle_step(O, frames[(tid + s) % len(frames)], prev, ba, pose, P)

    ths = [threading.Thread(target=worker, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    return threads * (steps_per_thread + 1) / dt, dt


def cpu_baseline(frames, ba, pose, threads=1, budget_s=20.0):
    fps, dt = cpu_run(frames, ba, pose, threads, 1)
    steps = 1
    if dt < budget_s / 3:
        steps = max(1, int(budget_s / (dt / 2)) - 1)
        fps, dt = cpu_run(frames, ba, pose, threads, steps)
    return {"value": round(fps, 4), "unit": "frames/s", "cores": threads, "kind": "port",
            "sample": "%d frame(s) of the same workload through oracle/ (restated CPU path, single thread), %.1f s" % ((steps + 1) * threads, dt),
            "host_cores_available": host_cores()}


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return None
    frames, ba, pose = make_workload(0, 6)
    threads = min(host_cores(), args.ref_threads) if args.ref_threads > 0 else host_cores()
    total = args.steps + args.warmup
    per_thread = max(1, (total + threads - 1) // threads)
    fps, dt = cpu_run(frames, ba, pose, threads, per_thread)
    return {
        "impl": "reference", "metric": METRIC, "value": round(fps, 4), "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1e3 / fps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8 + f64",
        "data": "synthetic", "config": {"workload": WORKLOAD},
        "cpu_baseline": {"value": round(fps, 4), "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": "%d independent streams x %d frames through oracle/ (restated CPU path; the reference itself cannot be built: no source in /root/reference), %.1f s"
                                   % (threads, per_thread + 1, dt)},
        "e2e": {"value": round(fps, 4), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ring", type=int, default=72, help="frames in the device ring (72 x 1.84 MB > L2)")
    ap.add_argument("--streams", type=int, default=8, help="independent camera streams per GPU (one host thread + private CUDA streams each)")
    ap.add_argument("--spec", type=int, default=0, help="local BA speculation width 1..4 (0 = default 4)")
    ap.add_argument("--wait", default="auto", choices=["auto", "spin", "block", "yield"], help="host wait mode (auto: spin while streams x ranks fit the usable cores, else yield-poll)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true", help="skip the single-stream latency pass")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--ref-threads", type=int, default=0, help="CPU arm: independent streams (0 = one per usable host core, cgroup quota respected)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    out = run_reference(args) if args.impl == "reference" else run_ours(args)
    if out is not None:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
