"""GPU tests written AFTER the round's GPU minutes were spent: no B200 has run them yet. The file name sorts last on
purpose, so that `pytest -x` reaches them only after the suite that was verified on B200s (tests/test_gpu_*.py,
tests/test_golden.py). Everything they compare against is CPU-verified: the digests were computed by the reference itself
(tests/golden/make_reference_golden.py), the loop-closure integration by the oracle twin, which tests/test_ref_full.py
holds to the reference's own SurfelMapping::integrateLoopClosures bit for bit. What they exercise of the CUDA path has run,
and passes, on the CPU executor for the library's CUDA sources (`pytest -m gpu --cusim`, tests/cusim; log in
profiles/r02_cusim_runs.txt) -- kernel logic, not a GPU run."""
import json
import os

import numpy as np
import pytest

from helpers import assert_bits_equal, both_params, sized, surfel_fields_equal

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.gpu


def test_cuda_matches_the_reference_generated_golden():
    """the CUDA path against digests of what THE REFERENCE ITSELF computed (oracle/_ref/libsuma_ref_full.so in the build
    container): preprocessing of four scans and a map update / rendering tour that shifts the submap window -- everything
    on the path that does not pass through the reference's fp32 blending of the 48 ICP values"""
    from semantic_suma_b200 import api
    from golden import make_reference_golden as RG
    from helpers import sized
    ref = json.load(open(os.path.join(HERE, "golden", "reference_golden.json")))

    class CudaEngine:
        def __init__(self, p):
            self.ctx = api.Context(api.default_params(**sized(900)))
            self.map = api.SurfelMap(self.ctx)

        def preprocess(self, pts, lab, prb, timestamp):
            f = api.Frame(self.ctx, 900, 64)
            api.Preprocessing(self.ctx).process(pts, f, lab, prb, timestamp)
            return f

        def maps(self, frame):
            return frame.maps()

        def map_update(self, T, frame):
            self.map.update(T, frame)

        def map_render(self, T, ct):
            out = api.Frame(self.ctx, 900, 64)
            self.map.render(T, T, out, ct)
            return out.maps()

        def map_surfels(self):
            return self.map.getAllSurfels()

        def close(self):
            self.ctx.close()

    got = RG.compute(CudaEngine, process_scan=False)
    for k in got:
        assert got[k] == ref[k], k


def test_loop_closure_integration_bit_exact():
    """sb_integrate_loop_closures = SurfelMapping::integrateLoopClosures (SurfelMapping.cpp:212-258): when the library
    raises optimisation_requested the host hands the (here: unchanged) graph poses back before the next scan, as the
    reference does once its gtsam run has finished; poses, flags, counters and surfels then stay bit-identical to the oracle
    twin through the following scans and a second request / integrate cycle"""
    from oracle import oracle as O
    from semantic_suma_b200 import api, synth
    po, pp = both_params(**sized(900))
    scene = synth.Scene(width=900, height=64)
    N = 126
    poses = synth.trajectory(N, step=0.2618, yaw_deg=3.0)
    lp = dict(search_distance=3.0, min_trajectory_distance=15.0, min_verifications=2)
    osl = O.Slam(po)
    osl.enable_loop_closure(**lp)
    gsl = api.SurfelMapping(pp)
    gsl.enableLoopClosure(True, **lp)
    integrations = 0
    for f in range(N):
        pts, _, _ = scene.scan(f, poses[f])
        if osl.loop_info()["optimisation_requested"]:
            assert gsl.getLoopInfo()["optimisation_requested"] == 1
            assert gsl.integrateLoopClosures() == osl.integrate_loop_closures() == f
            integrations += 1
        osl.process_scan(pts)
        gsl.processScan(pts)
        assert_bits_equal(gsl.getCurrentPose(), osl.pose(), "t=%d pose" % f)
        if f >= 100:
            a, b = gsl.getLoopInfo(), osl.loop_info()
            for k in ("loop_count", "loop_edges_added", "found_candidate", "use_candidate", "optimisation_requested",
                      "n_edges", "n_poses"):
                assert a[k] == b[k], "t=%d %s: %r vs %r" % (f, k, a[k], b[k])
        assert gsl.getMap().size() == osl.map.size(), "t=%d surfel count" % f
    assert integrations == 2
    surfel_fields_equal(gsl.getMap().getAllSurfels(), osl.map.download(), "surfels after two integrations")
    gsl.ctx.close()


def _pipeline_equal(po, pp, sc, what):
    """tests/test_gpu_parity.py::_pipeline_equal (verified on B200s), repeated here so that this file stands alone"""
    from oracle import oracle as O
    from semantic_suma_b200 import api
    osl = O.Slam(po)
    gsl = api.SurfelMapping(pp)
    try:
        for t, (pts, lab, prb) in enumerate(sc):
            osl.process_scan(pts, lab, prb)
            gsl.processScan(pts, lab, prb)
            assert_bits_equal(gsl.getCurrentPose(), osl.pose(), "%s t=%d pose" % (what, t))
            so, sg = osl.stats(), gsl.getStatistics()
            assert sg["num_iterations"] == so["iterations"], "%s t=%d iterations" % (what, t)
            assert (sg["F"], sg["inlier"], sg["outlier"], sg["invalid"]) == (so["F"], so["inlier"], so["outlier"], so["invalid"]), \
                "%s t=%d statistics pass" % (what, t)
            assert gsl.getMap().size() == osl.map.size(), "%s t=%d surfel count" % (what, t)
        surfel_fields_equal(gsl.getMap().getAllSurfels(), osl.map.download(), what + " surfels")
        for g, o in zip(gsl.getCurrentFrame().maps(), osl.frame(0)):
            assert_bits_equal(g, o, what + " current frame")
        for g, o in zip(gsl.getLastModelFrame().maps(), osl.frame(1)):
            assert_bits_equal(g, o, what + " model frame")
    finally:
        gsl.ctx.close()


def test_parameter_branches_bit_exact():
    """the 24 parameter variants tests/test_ref_full.py holds the oracle to the reference with (robust weighting, sampling,
    compose rendering on / off, initial guess, surfel weighting / averaging schemes, confidence modes, stability, thresholds,
    submap layout): the CUDA path against the oracle, four semantic scans at 64x450 each"""
    from semantic_suma_b200 import synth
    from test_ref_full import VARIANTS
    scene = synth.Scene(width=450, height=64, semantic=True)
    poses = synth.trajectory(4)
    sc = [scene.scan(t, poses[t]) for t in range(4)]
    for kw in VARIANTS:
        po, pp = both_params(**sized(450), **kw)
        _pipeline_equal(po, pp, sc, repr(kw))


def test_image_geometries_bit_exact():
    """model image size / field of view / depth range different from the data image's, through the whole pipeline"""
    from semantic_suma_b200 import synth
    from test_ref_full import GEOMETRIES
    for kw in GEOMETRIES:
        po, pp = both_params(**kw)
        scene = synth.Scene(width=kw["data_width"], height=kw["data_height"], fov_up=kw.get("data_fov_up", 3.0),
                            fov_down=kw.get("data_fov_down", -25.0), semantic=True)
        poses = synth.trajectory(3)
        _pipeline_equal(po, pp, [scene.scan(t, poses[t]) for t in range(3)], repr(kw))


@pytest.mark.parametrize("t_now,compose,seed", [(150, 1, 1), (40, 1, 2), (150, 0, 3), (150, 1, 4)])
def test_adversarial_random_clouds_bit_exact(t_now, compose, seed):
    """the adversarial inputs tests/test_ref_full.py feeds the reference's own SurfelMap class with (random surfels of any
    orientation, huge and tiny discs across the azimuth seam -- the warp-cooperative large-quad rasteriser --, negative
    confidences, old and new creation times, movable labels; points outside the field of view, duplicates, zeros), through the
    CUDA operators: preprocessing, render / render_active / render_inactive / render_composed, map update -- against the oracle"""
    from oracle import oracle as O
    from semantic_suma_b200 import api
    from test_ref_shaders import _random_cloud, _random_surfels
    rng = np.random.default_rng(1000 * seed + t_now + compose)
    kw = dict(sized(360, 32), compose_rendering=compose)
    po, pp = both_params(**kw)
    pts, lab, prob = _random_cloud(rng, 30000, po)
    lab[:] = rng.choice(np.array((0, 10, 30, 40, 50), np.float32), lab.shape[0])
    ctx = api.Context(pp)
    om, gm = O.Map(po), api.SurfelMap(ctx)
    S = _random_surfels(rng, 20000, t_now)
    pose = np.eye(4)
    pose[:3, 3] = (0.3, -0.2, 0.1)
    a = np.deg2rad(3.0)
    pose[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]
    om.upload(S, t_now)
    gm.upload(S, t_now)
    for t in range(0, t_now + 1, 7):
        T = np.eye(4); T[:3, 3] = (0.01 * t, 0.002 * t, 0.0)
        om.set_pose(t, T)
        gm.set_pose(t, T)
    try:
        data = O.preprocess(po, pts, lab, prob, timestamp=t_now)
        f = api.Frame(ctx, 360, 32)
        api.Preprocessing(ctx).process(pts, f, lab, prob, t_now)
        for g, o, name in zip(f.maps(), data, "vns"):
            assert_bits_equal(g, o, "preprocess " + name)
        out = api.Frame(ctx, 360, 32)
        gm.render(np.eye(4), pose, out, 0.5)
        for g, o, name in zip(out.maps(), om.render(np.eye(4), pose, 0.5), "vns"):
            assert_bits_equal(g, o, "render " + name)
        frames = (gm.oldMapFrame, gm.newMapFrame, gm.composedFrame)
        for w in range(3 if compose else 2):
            for g, o, name in zip(frames[w]().maps(), om.frame(w), "vns"):
                assert_bits_equal(g, o, "frame %d %s" % (w, name))
        om.render_active(pose, 0.2); gm.render_active(pose, 0.2)
        om.render_inactive(pose, 0.2); gm.render_inactive(pose, 0.2)
        om.render_composed(np.eye(4), pose, 0.2); gm.render_composed(np.eye(4), pose, 0.2)
        for w in range(3):
            for g, o, name in zip(frames[w]().maps()[:2], om.frame(w)[:2], "vn"):
                assert_bits_equal(g, o, "single-view frame %d %s" % (w, name))
        om.update(pose, data)
        gm.update(pose, f)
        assert gm.size() == om.size() > 0
        surfel_fields_equal(gm.getAllSurfels(), om.download(), "surfels after the update")
    finally:
        ctx.close()
