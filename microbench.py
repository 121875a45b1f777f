#!/usr/bin/env python
"""Fixed-load micro-benchmark of the surfel passes ("S = 1 000 000 synthetic surfels for K4/K6").

bench.py measures whole scans; here the map is frozen at exactly S surfels so that the per-kernel
numbers of K4 (render_scatter / render_resolve) and K6 (index_scatter, update_surfels, compaction, generation) can be
compared between builds. The load is produced by the product itself: the bench sequence is processed until the map
holds at least S surfels, the first S records (buffer order) and all poses are copied into a fresh context, and every
repetition re-uploads them before the timed call. Per-kernel device time comes from the library's CUDA-event profiler.

    python microbench.py [--surfels 1000000] [--reps 20]

Prints one JSON line (round-2 results: profiles/r02_microbench_*.json).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--surfels", type=int, default=1000000)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--max-scans", type=int, default=110)
    args = ap.parse_args()
    import bench
    from semantic_suma_b200 import api
    w = bench.WORKLOADS["hdl64_2048_geometric"]
    pp = api.default_params(**bench.param_kwargs(w))
    scans = bench.generate_scans(w, args.max_scans, seed=1337)
    grow = api.SurfelMapping(pp, device=0)
    poses = []
    used = 0
    for p, l, q in scans:
        grow.processScan(p)
        poses.append(grow.getCurrentPose().astype(np.float32))
        used += 1
        if grow.getMap().size() >= args.surfels:
            break
    S = min(args.surfels, grow.getMap().size())
    surfels = grow.getMap().getAllSurfels()[:S].copy()
    t_map = grow.getMap().timestamp()
    frame_maps = grow.getCurrentFrame().maps()          # the last scan's vertex / normal / semantic images
    pose = poses[-1]
    ct = float(pp.confidence_threshold)
    grow.ctx.close()

    slam = api.SurfelMapping(pp, device=0)
    ctx, m = slam.ctx, slam.getMap()
    frame = api.Frame(ctx, pp.data_width, pp.data_height)
    frame.upload(frame_maps)
    model = api.Frame(ctx, pp.model_width, pp.model_height)

    def load():
        m.upload(surfels, t_map)
        m.updatePoses(poses)

    P = pp.data_width * pp.data_height
    out = {"surfels": int(S), "scans_to_reach": used, "reps": args.reps, "range_image": "%dx%d" % (pp.data_height, pp.data_width)}
    for name, call in (("render", lambda: m.render(pose, pose, model, ct)),
                       ("render_active", lambda: m.render_active(pose, ct)),
                       ("update", lambda: m.update(pose, frame))):
        load(); call(); ctx.synchronize()                 # warm-up
        tot = {}
        for _ in range(args.reps):
            load()
            ctx.synchronize()
            ctx.profile(True)
            call()
            prof = ctx.profile_collect()
            ctx.profile(False)
            for k, (ms, cnt) in prof.items():
                a = tot.setdefault(k, [0.0, 0])
                a[0] += ms
                a[1] += cnt
        table = {}
        for k, (ms, cnt) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
            us = 1e3 * ms / cnt
            b = bench.algorithmic_bytes(k, P, S, P, False)
            table[k] = {"launches_per_call": round(cnt / args.reps, 2), "avg_us": round(us, 2),
                        "algorithmic_gbps": round(b / (us * 1e-6) / 1e9, 1) if us > 0 and b > 0 else None}
        out[name] = table
    print(json.dumps(out))


if __name__ == "__main__":
    main()
