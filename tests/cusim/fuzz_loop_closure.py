"""TEST INFRASTRUCTURE: loop-closure fuzz on the CPU executor against the oracle twin: a circle driven past its start (120 / 100 / 90
scans per turn, random step, search distance, verification count, image width), with and without integrateLoopClosures on
request: poses after every scan, the loop counters from scan 95 on, the surfel records at the end.
usage: python tests/cusim/fuzz_loop_closure.py [runs=8]"""
import sys, os, random
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0, os.path.dirname(HERE))
from cusim import build_sim
path = build_sim.build()
from semantic_suma_b200 import build as B
B.LIB = path; B.build = lambda *a, **k: path
import numpy as np
from oracle import oracle as O
from semantic_suma_b200 import api, synth
from helpers import both_params, sized, assert_bits_equal, surfel_fields_equal
rnd = random.Random(11)
bad = 0; closed = 0
N_RUNS = int(sys.argv[1]) if len(sys.argv) > 1 else 8
for i in range(N_RUNS):
    W = rnd.choice((360, 450))
    yaw = rnd.choice((3.0, 3.6, 4.0)); n_circle = int(round(360.0 / yaw))
    step = rnd.uniform(0.2, 0.35)
    N = n_circle + rnd.randint(10, 24)
    lp = dict(search_distance=rnd.choice((2.0, 3.0, 5.0)), min_trajectory_distance=rnd.choice((10.0, 15.0)), min_verifications=rnd.choice((1, 2, 3)))
    integrate = rnd.random() < 0.5
    po, pp = both_params(**sized(W))
    scene = synth.Scene(width=W, height=64, seed=700 + i)
    poses = synth.trajectory(N, step=step, yaw_deg=yaw)
    osl = O.Slam(po); osl.enable_loop_closure(**lp)
    gsl = api.SurfelMapping(pp); gsl.enableLoopClosure(True, **lp)
    keys = ("loop_count", "candidates_tested", "loop_edges_added", "unverified", "already_verified", "found_candidate", "use_candidate",
            "optimisation_requested", "n_edges", "n_poses")
    try:
        for f in range(N):
            pts, _, _ = scene.scan(f, poses[f])
            if integrate and osl.loop_info()["optimisation_requested"]:
                assert gsl.getLoopInfo()["optimisation_requested"] == 1
                assert gsl.integrateLoopClosures() == osl.integrate_loop_closures()
            osl.process_scan(pts); gsl.processScan(pts)
            assert_bits_equal(gsl.getCurrentPose(), osl.pose(), "t=%d pose" % f)
            if f >= 95:
                a, b = gsl.getLoopInfo(), osl.loop_info()
                for k in keys:
                    assert a[k] == b[k], "t=%d %s: %r vs %r" % (f, k, a[k], b[k])
            assert gsl.getMap().size() == osl.map.size(), "t=%d surfel count" % f
        surfel_fields_equal(gsl.getMap().getAllSurfels(), osl.map.download(), "surfels")
        closed += 1 if osl.loop_info()["loop_edges_added"] > 0 else 0
    except AssertionError as e:
        bad += 1; print("FAIL run", i, W, yaw, step, lp, integrate, str(e)[:300])
    gsl.ctx.close()
print("loop-closure fuzz: %d runs (circle driven past its start; random radius, step, search distance, verification count, with / without integration), %d closed the loop, failures %d" % (N_RUNS, closed, bad))
