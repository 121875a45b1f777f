#!/usr/bin/env python
"""bench.py -- scans/sec of the per-scan hot path (SurfelMapping::processScan, loop closure off) on synthetic
HDL-64E-shaped scans: 64 x 2048 range images, exactly 10 Gauss-Newton iterations (BASELINE.json configs[1]).

  python bench.py --gpus N --steps K --warmup W            # N > 1: launched under torchrun, one rank per GPU
  python bench.py --impl reference --steps K --warmup W    # the CPU restatement of the reference on the host cores

One step = one scan through the whole path (K1-K3 preprocessing, model rendering, 10 x (K5 + GN step), post-ICP
rendering + statistics pass, K6 map update, model re-render). `value` is measured with the scans already resident in
HBM; `e2e` goes through the C ABI with pinned HOST buffers (H2D of the scan and D2H of the pose / sums inside the
timed region). N > 1 runs one independent sequence per GPU (weak scaling, no data-path collective).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from semantic_suma_b200 import synth  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[1]: KITTI-shaped 64x2048, geometric SuMa (no labels), exactly 10 ICP iterations
    "hdl64_2048_geometric": dict(width=2048, height=64, fov_up=3.0, fov_down=-25.0, semantic=False, iters=10),
    # configs[2]: semantic-weighted ICP + label-consistent fusion
    "hdl64_2048_semantic": dict(width=2048, height=64, fov_up=3.0, fov_down=-25.0, semantic=True, iters=10),
    # configs[0]: the plumbing case
    "hdl64_900_geometric": dict(width=900, height=64, fov_up=3.0, fov_down=-25.0, semantic=False, iters=10),
    # configs[3]: Ouster-128-style
    "ouster128_4096_geometric": dict(width=4096, height=128, fov_up=22.5, fov_down=-22.5, semantic=False, iters=15),
}


def param_kwargs(w):
    return dict(data_width=w["width"], data_height=w["height"], model_width=w["width"], model_height=w["height"],
                data_fov_up=w["fov_up"], data_fov_down=w["fov_down"], model_fov_up=w["fov_up"],
                model_fov_down=w["fov_down"], max_iterations=w["iters"], stopping_threshold=0.0, delta=0.0)


def _gen_one(job):
    w, seed, frame, pose = job
    sc = synth.Scene(width=w["width"], height=w["height"], fov_up=w["fov_up"], fov_down=w["fov_down"],
                     semantic=w["semantic"], seed=seed)
    return sc.scan(frame, pose)


def generate_scans(w, n_frames, seed):
    poses = synth.trajectory(n_frames)
    jobs = [(w, seed, f, poses[f]) for f in range(n_frames)]
    workers = max(1, min(32, (os.cpu_count() or 2) // max(1, int(os.environ.get("WORLD_SIZE", "1"))) - 1))
    if workers > 1 and n_frames > 4:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            return pool.map(_gen_one, jobs, chunksize=2)
    return [_gen_one(j) for j in jobs]


class ClockSampler(threading.Thread):
    """samples SM clock / power / throttle reasons during the timed region (B200_PROFILING.md, clocks line).
    Uses NVML in-process (nvidia_ml_py): spawning nvidia-smi every 100 ms perturbs a 50 ms timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        except Exception:  # noqa: BLE001
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)
        pw = n.nvmlDeviceGetPowerUsage(self.h) / 1000.0
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:  # noqa: BLE001
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        flags = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}
        return [sm, mx, pw] + [bool(r & v) for v in (flags["hw_slowdown"], flags["hw_thermal_slowdown"],
                                                     flags["sw_thermal_slowdown"], flags["sw_power_cap"])]

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout
        f = [x.strip() for x in out.strip().split(",")]
        return [float(f[0]), float(f[1]), float(f[2])] + [x.lower().startswith("active") for x in f[3:7]]

    def run(self):
        while not self.stop_flag:
            try:
                self.samples.append(self._sample_nvml() if self.nvml else self._sample_smi())
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.005 if self.nvml else 0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i] for s in self.samples)]
        return {"sm_mhz": statistics.median(s[0] for s in self.samples), "sm_max_mhz": max(s[1] for s in self.samples),
                "power_w_max": round(max(s[2] for s in self.samples), 1), "reasons": reasons,
                "samples": len(self.samples), "source": "nvml" if self.nvml else "nvidia-smi"}


def algorithmic_bytes(kernel, P, S, N, semantic):
    """compulsory HBM bytes of one launch (DESIGN.md 'Kernels and their rooflines')"""
    t = {
        "fill_u64": 8 * P, "project_scatter": 16 * N, "preprocess_tile": 8 * P + 16 * N + (48 + 16) * P,
        "normals_erode": 32 * P + 32 * P, "floodfill": 32 * P + 16 * P,
        "icp_fused": (96 if semantic else 64) * P, "icp_jacobian": (96 if semantic else 64) * P,
        "render_scatter": 48 * S, "render_resolve": 3 * 8 * P + 4 * 48 * P,
        "index_scatter": 48 * S, "radius": 48 * P, "update_surfels": 129 * S, "gen_surfels": 84 * P,
        "compact_scatter": 129 * S, "scan_blocks": 8 * (S // 256 + 1), "pose_products": 128 * 100,
    }
    return float(t.get(kernel, 0))


def run_native(args, w, rank, world, local_rank):
    import torch
    from semantic_suma_b200 import api

    n_frames = args.warmup + args.steps
    dbg = bool(os.environ.get("SUMA_B200_WATCHDOG"))
    if dbg:
        print("[r%d] generating %d scans" % (rank, n_frames), file=sys.stderr, flush=True)
    scans = generate_scans(w, n_frames, seed=1337 + (0 if args.striped else 1000 * rank))
    if dbg:
        print("[r%d] scans ready" % rank, file=sys.stderr, flush=True)
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pp = api.default_params(**param_kwargs(w))
    slam = api.SurfelMapping(pp, device=local_rank)
    ctx = slam.ctx
    if dist is not None and args.striped:
        from semantic_suma_b200 import stripes
        stripes.setup_comm(ctx, dist, fused=(args.comm == "fused"))
    stream = torch.cuda.ExternalStream(ctx.stream(), device=local_rank)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    sem = w["semantic"]

    dev = [(torch.from_numpy(p).cuda(), torch.from_numpy(l).cuda() if sem else None,
            torch.from_numpy(q).cuda() if sem else None) for p, l, q in scans]
    pin = [(torch.from_numpy(p).pin_memory(), torch.from_numpy(l).pin_memory() if sem else None,
            torch.from_numpy(q).pin_memory() if sem else None) for p, l, q in scans]
    h2d = sum(int(p.numel() * 4 + (l.numel() * 4 + q.numel() * 4 if sem else 0)) for p, l, q in pin[args.warmup:])
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def one_pass(bufs, on_device, sampler=None):
        prefetch = args.prefetch and not on_device
        slam.reset()
        for f in range(args.warmup):
            p, l, q = bufs[f]
            slam.process_scan_raw(p.data_ptr(), l.data_ptr() if sem else 0, q.data_ptr() if sem else 0, p.shape[0],
                                  on_device)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        launches0 = ctx.launch_count()
        barrier()
        if sampler:
            sampler.start()
        t_wall = time.time()
        for i in range(args.steps):
            p, l, q = bufs[args.warmup + i]
            with torch.cuda.stream(stream):
                flush.zero_()  # L2 flush between timed steps (not timed)
            ev[i][0].record(stream)
            if prefetch and i + 1 < args.steps:  # experimental: stage the next scan while this one is processed
                pn, ln, qn = bufs[args.warmup + i + 1]
                slam.prefetch_scan_raw(pn.data_ptr(), ln.data_ptr() if sem else 0, qn.data_ptr() if sem else 0, pn.shape[0])
            slam.process_scan_raw(p.data_ptr(), l.data_ptr() if sem else 0, q.data_ptr() if sem else 0, p.shape[0],
                                  on_device)
            if not on_device:
                slam.getCurrentPose()  # the step's result, read on the host
            ev[i][1].record(stream)
        barrier()
        t_wall = time.time() - t_wall
        if sampler:
            sampler.stop_flag = True
        ms = [a.elapsed_time(b) for a, b in ev]
        total_ms = float(sum(ms))
        if dist is not None:
            t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms, ctx.launch_count() - launches0, t_wall, ms

    if dbg:
        print("[r%d] buffers ready, starting passes" % rank, file=sys.stderr, flush=True)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_dev, launches, wall_dev, per_step = one_pass(dev, True, sampler)
    if dbg:
        print("[r%d] device-resident pass done" % rank, file=sys.stderr, flush=True)
    ms_e2e, _, wall_e2e, _ = one_pass(pin, False)
    surfels = slam.getMap().size()
    gt = np.linalg.inv(synth.trajectory(1)[0]) @ synth.trajectory(n_frames)[-1]
    drift = float(np.linalg.norm(slam.getCurrentPose()[:3, 3] - gt[:3, 3]))

    # per-kernel device time with CUDA events on the launching stream, over the same timed frames
    roofline = None
    kernel_table = {}
    if (rank == 0 or args.striped) and not args.no_profile:  # striped: every rank must walk the same scans
        slam.reset()
        for f in range(args.warmup):
            p, l, q = dev[f]
            slam.process_scan_raw(p.data_ptr(), l.data_ptr() if sem else 0, q.data_ptr() if sem else 0, p.shape[0], True)
        ctx.synchronize()
        ctx.profile(True)
        s_sum = 0
        for i in range(args.steps):
            p, l, q = dev[args.warmup + i]
            with torch.cuda.stream(stream):
                flush.zero_()
            s_sum += slam.getMap().size()
            slam.process_scan_raw(p.data_ptr(), l.data_ptr() if sem else 0, q.data_ptr() if sem else 0, p.shape[0], True)
        prof = ctx.profile_collect()
        ctx.profile(False)
        tot = sum(v[0] for v in prof.values())
        P = w["width"] * w["height"]
        S_avg = s_sum / max(1, args.steps)
        N_avg = float(np.mean([p.shape[0] for p, _, _ in scans[args.warmup:]]))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        for k, (ms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
            b = algorithmic_bytes(k, P, S_avg, N_avg, sem)
            us = 1e3 * ms / cnt
            kernel_table[k] = {"launches_per_step": round(cnt / args.steps, 2), "avg_us": round(us, 2),
                               "share": round(ms / tot, 4), "gbps": round(b / (us * 1e-6) / 1e9, 1) if us > 0 else None}
        top = max(prof.items(), key=lambda kv: kv[1][0])[0]
        us = 1e3 * prof[top][0] / prof[top][1]
        ach = algorithmic_bytes(top, P, S_avg, N_avg, sem) / (us * 1e-6) / 1e9
        traffic, traffic_src = None, None
        try:  # DRAM bytes of this kernel from the committed ncu --set full capture (profiles/ncu_traffic.py)
            tj = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))
            traffic = tj["kernels"][top]["dram_bytes_per_launch"]
            traffic_src = tj.get("source")
        except Exception:  # noqa: BLE001
            pass
        roofline = {"kernel": top, "bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(ach / peak, 4), "traffic": traffic, "traffic_source": traffic_src,
                    "algorithmic_bytes_per_launch": int(algorithmic_bytes(top, P, S_avg, N_avg, sem)),
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                    "avg_launch_us": round(us, 2), "surfels_avg": int(S_avg),
                    "step_share": kernel_table[top]["share"]}

    out = None
    if rank == 0:
        cpu = None if args.no_cpu_baseline else cpu_baseline(w, scans, budget_s=args.cpu_budget)
        seqs = 1 if args.striped else world
        value = seqs * args.steps / (ms_dev * 1e-3)
        e2e_value = seqs * args.steps / (ms_e2e * 1e-3)
        out = {
            "metric": "scans_per_sec", "value": round(value, 2), "unit": "scans/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_dev / args.steps, 4),
            "higher_is_better": True, "scaling": "strong" if args.striped else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "range_image": "%dx%d" % (w["height"], w["width"]),
                       "icp_iterations": w["iters"], "semantic": sem, "sequences": 1 if args.striped else world,
                       "parallelism": ("k5-row-stripes-x%d (%s)" % (world, "peer-memory all-reduce in-kernel" if args.comm == "fused"
                                                                   else "torch.distributed all-reduce per iteration")) if args.striped
                       else "one sequence per GPU",
                       "surfels_end": int(surfels), "pose_drift_m": round(drift, 4),
                       "l2": "flushed between timed steps (256 MiB memset, untimed)",
                       "step_ms_min_med_max": [round(min(per_step), 4), round(statistics.median(per_step), 4),
                                               round(max(per_step), 4)],
                       "timing": "CUDA events on the library's stream around every step, summed; max over ranks",
                       "reference_defaults": "config/default.xml except image size, max iterations, eps=delta=0"},
            "e2e": {"value": round(e2e_value, 2), "unit": "scans/s", "ms_per_step": round(ms_e2e / args.steps, 4),
                    "h2d_bytes_per_step": int(h2d / args.steps) + 128, "d2h_bytes_per_step": 536 + 256 + 16,
                    "wall_s": round(wall_e2e, 3), "input_prefetch": bool(args.prefetch)},
            "gpu_launches": int(launches),
            "clocks": sampler.summary() if sampler else None,
            "roofline": roofline, "kernels": kernel_table, "cpu_baseline": cpu,
        }
    slam.ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return out


def pick_oracle_threads(w, scans):
    """The oracle scales with cores only up to a point (per-thread raster targets are merged per pixel, the host may be
    a container with fewer usable cores than it reports): try a few thread counts on the first scans, keep the best."""
    from oracle import oracle as O
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (min(ncpu, 64), 32, 16, 8, 4) if c <= max(ncpu, 1)}, reverse=True)
    po = O.default_params(**param_kwargs(w))
    best, best_t, tried = cands[-1], float("inf"), {}
    for c in cands:
        O.set_threads(c)
        sl = O.Slam(po)
        for a in scans[:2]:
            sl.process_scan(*a)
        t0 = time.time()
        for a in scans[2:6]:
            sl.process_scan(*a)
        dt = time.time() - t0
        tried[c] = round(dt, 3)
        if dt < best_t:
            best, best_t = c, dt
    return O.set_threads(best), tried


def cpu_baseline(w, scans, budget_s=15.0, max_frames=None):
    """the CPU restatement of the reference (oracle/, OpenMP over all host cores) on a bounded sample of the same scans"""
    from oracle import oracle as O
    threads, tried = pick_oracle_threads(w, scans)
    po = O.default_params(**param_kwargs(w))
    sl = O.Slam(po)
    t0 = time.time()
    n = 0
    for p, l, q in scans:
        sl.process_scan(p, l, q)
        n += 1
        if time.time() - t0 > budget_s or (max_frames and n >= max_frames):
            break
    dt = time.time() - t0
    return {"value": round(n / dt, 3), "unit": "scans/s", "cores": threads, "kind": "port",
            "sample": "first %d scans of the same synthetic sequence (map grows from empty), oracle/ C port, "
                      "%d OpenMP threads (fastest of %s on scans 3-6)" % (n, threads, sorted(tried)),
            "seconds": round(dt, 2), "host_cpus": os.cpu_count(), "thread_calibration_s": tried}


def run_reference(args, w, rank, world):
    """--impl reference: the reference has no CPU (or buildable GL) path in this environment; the arm times the
    oracle's restatement of it on all host cores (C + OpenMP, thread-count independent results) on the same workload."""
    if rank != 0:
        return None
    n_frames = args.warmup + args.steps
    scans = generate_scans(w, n_frames, seed=1337)
    from oracle import oracle as O
    threads, tried = pick_oracle_threads(w, scans)
    po = O.default_params(**param_kwargs(w))
    sl = O.Slam(po)
    for f in range(args.warmup):
        sl.process_scan(*scans[f])
    t0 = time.time()
    done = 0
    for i in range(args.steps):
        sl.process_scan(*scans[args.warmup + i])
        done += 1
        if time.time() - t0 > args.ref_budget:
            break
    dt = time.time() - t0
    v = done / dt
    return {
        "impl": "reference", "metric": "scans_per_sec", "value": round(v, 3), "unit": "scans/s", "n_gpus": world,
        "steps": done, "warmup": args.warmup, "ms_per_step": round(1e3 * dt / done, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "range_image": "%dx%d" % (w["height"], w["width"]),
                   "icp_iterations": w["iters"], "semantic": w["semantic"],
                   "note": "reference OpenGL path not runnable here (no GL/EGL, glow/gtsam/rangenet_lib absent); "
                           "this is the CPU restatement in oracle/ (kind=port)"},
        "cpu_baseline": {"value": round(v, 3), "unit": "scans/s", "cores": threads, "kind": "port",
                         "sample": "%d scans after %d warm-up scans, %d OpenMP threads (fastest of %s)"
                                   % (done, args.warmup, threads, sorted(tried)), "thread_calibration_s": tried,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": round(v, 3), "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main():
    if os.environ.get("SUMA_B200_WATCHDOG"):  # debugging aid: dump the Python stacks if the run hangs
        import faulthandler
        faulthandler.enable()
        faulthandler.dump_traceback_later(float(os.environ["SUMA_B200_WATCHDOG"]), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="hdl64_2048_geometric", choices=sorted(WORKLOADS))
    ap.add_argument("--comm", choices=["fused", "callback"], default="fused",
                    help="--striped exchange: in-kernel over peer memory (default) or the torch.distributed baseline")
    ap.add_argument("--striped", action="store_true",
                    help="N > 1: all ranks process the SAME sequence, the K5 reduction is striped over image rows and "
                         "all-reduced inside the kernel over peer memory (BASELINE.json configs[3]); strong scaling")
    ap.add_argument("--prefetch", action="store_true",
                    help="experimental (not the default, not measured yet): in the e2e pass stage scan i+1 on a copy "
                         "stream while scan i is processed (sb_prefetch_scan); every copy is still inside the timed region")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--ref-budget", type=float, default=240.0)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        print("bench.py: --gpus %d needs torchrun (python -m torch.distributed.run --nproc-per-node %d ...)" %
              (args.gpus, args.gpus), file=sys.stderr)
        sys.exit(2)
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        out = run_reference(args, w, rank, world)
    else:
        out = run_native(args, w, rank, world, local_rank)
    if rank == 0 and out is not None:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
