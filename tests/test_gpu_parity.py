"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.
Bar: bit-exact for images, surfels, fixed-point sums, poses (everything is a chain of correctly rounded IEEE ops)."""
import numpy as np
import pytest

from oracle import oracle as O
from semantic_suma_b200 import api
from helpers import assert_bits_equal, both_params, scans, sized, surfel_fields_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx900():
    po, pp = both_params(**sized(900))
    c = api.Context(pp)
    yield po, pp, c
    c.close()


def _prep_both(po, ctx, scan, timestamp):
    pts, lab, prb = scan
    ov = O.preprocess(po, pts, lab, prb, timestamp)
    f = api.Frame(ctx, po.data_width, po.data_height)
    api.Preprocessing(ctx).process(pts, f, lab, prb, timestamp)
    return ov, f


@pytest.mark.parametrize("width,semantic,timestamp", [(900, False, 100), (900, True, 100), (900, True, 3),
                                                      (2048, False, 100), (2048, True, 0)])
def test_preprocess_bit_exact(width, semantic, timestamp):
    po, pp = both_params(**sized(width))
    ctx = api.Context(pp)
    sc, _ = scans(width, n=2, semantic=semantic)
    for s in sc:
        ov, f = _prep_both(po, ctx, s, timestamp)
        gv, gn, gs = f.maps()
        assert_bits_equal(gv, ov[0], "vertex_map")
        assert_bits_equal(gn, ov[1], "normal_map")
        assert_bits_equal(gs, ov[2], "semantic_map")
    ctx.close()


def test_preprocess_edge_cases(ctx900):
    po, pp, ctx = ctx900
    pre = api.Preprocessing(ctx)
    f = api.Frame(ctx, 900, 64)
    # empty scan
    pre.process(np.zeros((0, 4), np.float32), f)
    ov = O.preprocess(po, np.zeros((0, 4), np.float32))
    for g, o in zip(f.maps(), ov):
        assert_bits_equal(g, o, "empty scan")
    # degenerate points: origin, out of range, behind, straight up/down, duplicates in one pixel
    pts = np.array([[0, 0, 0, 1], [1000, 0, 0, 1], [-5, 0, 0, 1], [0, 0, 10, 1], [0, 0, -10, 1], [5, 1, -0.5, 1],
                    [5, 1, -0.5, 1], [5.0001, 1, -0.5, 1], [1.0, 0, 0, 1], [-5, 1e-9, -0.3, 1], [-5, -1e-9, -0.3, 1]],
                   np.float32)
    pre.process(pts, f)
    ov = O.preprocess(po, pts)
    for g, o in zip(f.maps(), ov):
        assert_bits_equal(g, o, "degenerate points")


@pytest.mark.parametrize("bilinear,weighting,semantic", [(1, 1, False), (0, 1, False), (1, 2, True), (1, 0, True),
                                                         (0, 2, False)])
def test_icp_jacobian_exact(bilinear, weighting, semantic):
    kw = sized(900, bilinear_sampling=bilinear, weighting=weighting)
    po, pp = both_params(**kw)
    ctx = api.Context(pp)
    sc, poses = scans(900, n=2, semantic=semantic)
    o0, f0 = _prep_both(po, ctx, sc[0], 100)
    o1, f1 = _prep_both(po, ctx, sc[1], 100)
    obj = api.Frame2Model(ctx)
    obj.setData(f1, f0)
    T = np.linalg.inv(poses[0]) @ poses[1]
    T[0, 3] += 0.1
    for it, pose in enumerate([np.eye(4), T]):
        obj.initialize(pose)
        obj.iteration_ = it
        obj.jacobianProducts()
        o48, raw = O.icp_jacobian(po, o1, o0, pose, iteration=it)
        assert np.array_equal(obj.raw32, raw), "raw sums differ: %s vs %s" % (obj.raw32, raw)
        assert_bits_equal(obj.out48, o48, "out48")
        assert obj.valid() + obj.invalid() == 900 * 64
    # row stripes add up exactly (multi-GPU contract)
    obj.initialize(T)
    tot = np.zeros(32, np.int64)
    for r0, r1 in [(0, 16), (16, 17), (17, 64)]:
        obj.jacobianProducts(rows=(r0, r1))
        tot += obj.raw32
    obj.jacobianProducts()
    assert np.array_equal(tot, obj.raw32)
    ctx.close()


def test_icp_minimize_matches_oracle_and_host_loop():
    po, pp = both_params(**sized(900))
    ctx = api.Context(pp)
    sc, poses = scans(900, n=2)
    o0, f0 = _prep_both(po, ctx, sc[0], 100)
    o1, f1 = _prep_both(po, ctx, sc[1], 100)
    obj = api.Frame2Model(ctx)
    obj.setData(f1, f0)
    gn = api.LieGaussNewton(ctx)
    gn.minimize(obj, np.eye(4))
    pose_o, o48, k, hist = O.icp_minimize(po, o1, o0, np.eye(4))
    assert gn.iterationCount() == k
    assert_bits_equal(gn.pose(), pose_o, "pose")
    assert_bits_equal(gn.out48, o48, "out48")
    assert len(gn.history()) == len(hist)
    for a, b in zip(gn.history(), hist):
        assert_bits_equal(a, b, "history")
    # the reference's call pattern (one jacobianProducts per iteration, solve on the host) gives the same bits
    gn2 = api.LieGaussNewton(ctx)
    obj2 = api.Frame2Model(ctx)
    obj2.setData(f1, f0)
    gn2.minimize_host(obj2, np.eye(4))
    assert gn2.iterationCount() == k
    assert_bits_equal(gn2.pose(), pose_o, "host-loop pose")
    # converges to the simulated motion
    gt = np.linalg.inv(poses[0]) @ poses[1]
    assert np.linalg.norm(gn.pose()[:3, 3] - gt[:3, 3]) < 0.05
    ctx.close()


def _run_map_sequence(width, n_frames, semantic, **kw):
    po, pp = both_params(**sized(width, **kw))
    ctx = api.Context(pp)
    sc, poses = scans(width, n=n_frames, semantic=semantic)
    omap = O.Map(po)
    gmap = api.SurfelMap(ctx)
    out = api.Frame(ctx, width, 64)
    for t in range(n_frames):
        ov, f = _prep_both(po, ctx, sc[t], t)
        pose = (np.linalg.inv(poses[0]) @ poses[t]).astype(np.float32)
        ct = -2.0 + 0.2 * t
        # render before update (as preprocess() does)
        orr = omap.render(pose, pose, ct)
        gmap.render(pose, pose, out, ct)
        for g, o, name in zip(out.maps(), orr, ("vertex", "normal", "semantic")):
            assert_bits_equal(g, o, "t=%d render frame %s" % (t, name))
        for which, fr in ((0, gmap.oldMapFrame()), (1, gmap.newMapFrame()), (2, gmap.composedFrame())):
            for g, o, name in zip(fr.maps(), omap.frame(which), ("vertex", "normal", "semantic")):
                assert_bits_equal(g, o, "t=%d map frame %d %s" % (t, which, name))
        omap.update(pose, ov)
        gmap.update(pose, f)
        oi, orad, oint, onu, onn = omap.update_debug()
        gi, grad, gint, gnu, gnn = gmap.update_debug()
        assert_bits_equal(gi, oi, "t=%d index map" % t)
        assert_bits_equal(grad, orad, "t=%d radius map" % t)
        assert_bits_equal(gint, oint, "t=%d integrated flags" % t)
        assert (gnu, gnn) == (onu, onn), "t=%d counts %s vs %s" % (t, (gnu, gnn), (onu, onn))
        assert gmap.size() == omap.size()
        surfel_fields_equal(gmap.getAllSurfels(), omap.download(), "t=%d surfels" % t)
    # render_active / inactive / composed after the sequence
    pose = (np.linalg.inv(poses[0]) @ poses[n_frames - 1]).astype(np.float32)
    pose2 = pose.copy(); pose2[0, 3] += 0.3
    omap.render_active(pose2, 0.0); gmap.render_active(pose2, 0.0)
    omap.render_inactive(pose, 0.0); gmap.render_inactive(pose, 0.0)
    omap.render_composed(pose, pose2, 0.0); gmap.render_composed(pose, pose2, 0.0)
    for which, fr in ((0, gmap.oldMapFrame()), (1, gmap.newMapFrame()), (2, gmap.composedFrame())):
        for g, o, name in zip(fr.maps(), omap.frame(which), ("vertex", "normal", "semantic")):
            assert_bits_equal(g, o, "final map frame %d %s" % (which, name))
    # different old / new poses through the full render
    orr = omap.render(pose, pose2, 0.0)
    gmap.render(pose, pose2, out, 0.0)
    for g, o, name in zip(out.maps(), orr, ("vertex", "normal", "semantic")):
        assert_bits_equal(g, o, "two-pose render %s" % name)
    ctx.close()


def test_map_update_and_render_bit_exact_geometric():
    _run_map_sequence(900, 6, False)


def test_map_update_and_render_bit_exact_semantic():
    _run_map_sequence(900, 5, True)


def test_map_render_old_surfels():
    # surfels older than the compose age (100 scans) exercise the old / composed views
    po, pp = both_params(**sized(900))
    ctx = api.Context(pp)
    sc, poses = scans(900, n=2)
    omap = O.Map(po); gmap = api.SurfelMap(ctx)
    ov, f = _prep_both(po, ctx, sc[0], 100)
    omap.update(np.eye(4, dtype=np.float32), ov); gmap.update(np.eye(4, dtype=np.float32), f)
    s = omap.download()
    half = s.shape[0] // 2
    # pretend the map is 150 scans old, half of the surfels re-observed recently
    s["timestamp"][:half] = 140
    omap.upload(s, 150); gmap.upload(s, 150)
    pose = np.eye(4, dtype=np.float32); pose[0, 3] = 0.5
    out = api.Frame(ctx, 900, 64)
    orr = omap.render(pose, pose, -1.0)
    gmap.render(pose, pose, out, -1.0)
    for g, o in zip(out.maps(), orr):
        assert_bits_equal(g, o, "compose frame")
    for which, fr in ((0, gmap.oldMapFrame()), (1, gmap.newMapFrame()), (2, gmap.composedFrame())):
        v = fr.vertex_map
        assert (v[..., 3] > 0).sum() > 1000
        for g, o in zip(fr.maps(), omap.frame(which)):
            assert_bits_equal(g, o, "frame %d" % which)
    ctx.close()


@pytest.mark.parametrize("width,semantic,frames", [(900, False, 8), (900, True, 6), (2048, False, 4)])
def test_process_scan_pipeline_bit_exact(width, semantic, frames):
    po, pp = both_params(**sized(width))
    sc, poses = scans(width, n=frames, semantic=semantic)
    osl = O.Slam(po)
    gsl = api.SurfelMapping(pp)
    for t in range(frames):
        pts, lab, prb = sc[t]
        osl.process_scan(pts, lab, prb)
        gsl.processScan(pts, lab, prb)
        assert_bits_equal(gsl.getCurrentPose(), osl.pose(), "t=%d pose" % t)
        so, sg = osl.stats(), gsl.getStatistics()
        assert sg["num_iterations"] == so["iterations"], "t=%d iterations" % t
        assert sg["F"] == so["F"] and sg["inlier"] == so["inlier"] and sg["outlier"] == so["outlier"]
        assert gsl.getMap().size() == osl.map.size(), "t=%d surfel count" % t
    surfel_fields_equal(gsl.getMap().getAllSurfels(), osl.map.download())
    for g, o in zip(gsl.getCurrentFrame().maps(), osl.frame(0)):
        assert_bits_equal(g, o, "current frame")
    gt = np.linalg.inv(poses[0]) @ poses[frames - 1]
    assert np.linalg.norm(gsl.getCurrentPose()[:3, 3] - gt[:3, 3]) < 0.1
    gsl.ctx.close()


def test_submap_paging_shift_extract_and_reinsert():
    """SurfelMap::updateActiveSubmaps (SurfelMap.cpp:744-824): moving > 1.1 * extent shifts the active window, queues the
    leaving tiles for extraction (one per update with partial extraction) and re-inserts cached tiles when coming back."""
    po, pp = both_params(**sized(900))
    ctx = api.Context(pp)
    sc, poses = scans(900, n=3)
    omap = O.Map(po)
    gmap = api.SurfelMap(ctx)
    frames = [_prep_both(po, ctx, s, 100) for s in sc]

    def pose_at(x, y):
        T = np.eye(4, dtype=np.float32)
        T[0, 3], T[1, 3] = x, y
        return T

    # forward along +x far enough to drop tiles (window is 9x9 tiles of 20 m), sideways, then all the way back
    xs = [0, 6, 12, 24, 36, 48, 60, 72, 84, 96, 108, 120, 120, 120, 108, 84, 60, 36, 12, 0, -12]
    ys = [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 12, 24, 24, 24, 24, 12, 0, 0, 0]
    shifts = 0
    for t, (x, y) in enumerate(zip(xs, ys)):
        ov, f = frames[t % 3]
        T = pose_at(x, y)
        omap.update(T, ov)
        gmap.update(T, f)
        assert gmap.submap_origin() == omap.submap_origin(), "t=%d submap origin / pending" % t
        assert gmap.size() == omap.size(), "t=%d size %d vs %d" % (t, gmap.size(), omap.size())
        if omap.submap_origin()[:2] != (0, 0):
            shifts += 1
    assert shifts > 5
    surfel_fields_equal(gmap.getAllSurfels(), omap.download(), "surfels after paging")
    out = api.Frame(ctx, 900, 64)
    orr = omap.render(pose_at(-12, 0), pose_at(-12, 0), -5.0)
    gmap.render(pose_at(-12, 0), pose_at(-12, 0), out, -5.0)
    for g, o in zip(out.maps(), orr):
        assert_bits_equal(g, o, "render after paging")
    ctx.close()
