"""CPU tests: analytic known-answer tests that pin the oracle (the reference ships no golden vectors)."""
import numpy as np
import pytest

from oracle import oracle as O
from helpers import scans, sized


def _ulp_err(val, ref):
    ref32 = np.float32(ref)
    return abs(np.float64(val) - ref) / np.float64(np.spacing(np.abs(ref32)) if ref32 != 0 else 1e-45)


def test_transcendentals_are_accurate():
    rng = np.random.default_rng(0)
    for x in rng.uniform(-1, 1, 3000).astype(np.float32):
        assert _ulp_err(O.scalar("asinf", float(x)), np.arcsin(np.float64(x))) < 3.0
        assert _ulp_err(O.scalar("acosf", float(x)), np.arccos(np.float64(x))) < 3.0
    for y, x in zip(rng.uniform(-80, 80, 3000).astype(np.float32), rng.uniform(-80, 80, 3000).astype(np.float32)):
        assert _ulp_err(O.scalar("atan2f", float(y), float(x)), np.arctan2(np.float64(y), np.float64(x))) < 4.0
    for x in rng.uniform(0, np.pi, 2000).astype(np.float32):
        assert abs(O.scalar("sinf", float(x)) - np.sin(np.float64(x))) < 1.5e-7
    for x in rng.uniform(-6, 2, 2000).astype(np.float32):
        assert _ulp_err(O.scalar("expf", float(x)), np.exp(np.float64(x))) < 2.0
    for x in rng.uniform(0.01, 60, 2000).astype(np.float32):
        assert abs(O.scalar("logf", float(x)) - np.log(np.float64(x))) < 3e-7
    for x in np.linspace(-7, 7, 1501):
        s, c = O.sincos(float(x))
        assert abs(s - np.sin(x)) < 3e-16 and abs(c - np.cos(x)) < 3e-16
    # special values
    assert O.scalar("atan2f", 0.0, 0.0) == 0.0
    assert np.isnan(O.scalar("acosf", 1.5)) and np.isnan(O.scalar("asinf", -1.0001))
    assert O.scalar("acosf", 1.0) == 0.0
    assert abs(O.scalar("atan2f", 0.0, -1.0) - np.pi) < 1e-6


def test_se3_exp_log_roundtrip_and_ldlt():
    rng = np.random.default_rng(1)
    for _ in range(50):
        x = rng.normal(0, 0.3, 6)
        T = O.se3_exp(x)
        assert np.allclose(T[:3, :3] @ T[:3, :3].T, np.eye(3), atol=1e-14)
        assert np.allclose(O.se3_log(T), x, atol=1e-10)
    assert np.array_equal(O.se3_exp(np.array([1, 2, 3, 0, 0, 0.0])), np.array([[1, 0, 0, 1], [0, 1, 0, 2], [0, 0, 1, 3], [0, 0, 0, 1.0]]))
    for _ in range(20):
        A = rng.normal(size=(6, 6)); A = A @ A.T + 0.1 * np.eye(6); b = rng.normal(size=6)
        assert np.allclose(O.ldlt_solve6(A, b), np.linalg.solve(A, b), rtol=1e-9, atol=1e-11)


def test_projection_pixel_of_beam_centres():
    # a point on the optical axis of pixel (row r, column k) must land in that pixel (gen_vertexmap.vert:78-89)
    p = O.default_params(**sized(900))
    W, H = 900, 64
    rows, cols = np.array([0, 1, 31, 63]), np.array([0, 1, 449, 450, 899])
    fov = 28.0
    pts = []
    for r in rows:
        for k in cols:
            az = np.pi * (1 - 2 * (k + 0.5) / W)
            el = np.deg2rad(3.0 - fov * (1 - (r + 0.5) / H))
            pts.append([10 * np.cos(el) * np.cos(az), 10 * np.cos(el) * np.sin(az), 10 * np.sin(el), 1])
    v, n, s = O.preprocess(p, np.array(pts, np.float32))
    hit = np.argwhere(v[..., 3] > 0)
    assert sorted(map(tuple, hit)) == sorted((r, k) for r in rows for k in cols)
    # pixels without a vertex carry normal (0,0,0,1) (Q2)
    assert np.array_equal(n[5, 5], np.array([0, 0, 0, 1], np.float32))


def test_zbuffer_nearest_wins_and_ties_keep_first():
    p = O.default_params(**sized(900))
    d = np.array([1, 0.01, -0.2])
    d = d / np.linalg.norm(d)
    pts = np.array([np.r_[d * 20, 1], np.r_[d * 10, 1], np.r_[d * 10, 1], np.r_[d * 30, 1]], np.float32)
    pts[2, 3] = 1.0
    v, _, _ = O.preprocess(p, pts)
    hit = np.argwhere(v[..., 3] > 0)
    assert hit.shape[0] == 1
    assert np.allclose(v[tuple(hit[0])][:3], d * 10, atol=1e-5)
    # out of depth range [2, 75): clipped
    pts = np.array([np.r_[d * 1.5, 1], np.r_[d * 80, 1]], np.float32)
    v, _, _ = O.preprocess(p, pts)
    assert (v[..., 3] > 0).sum() == 0


def test_movable_classes_removed_in_first_ten_scans_and_label_offset_quirk():
    p = O.default_params(**sized(900))
    sc, _ = scans(900, n=1, semantic=True)
    pts, lab, prb = sc[0]
    v0, _, s0 = O.preprocess(p, pts, lab, prb, timestamp=3)
    v1, _, s1 = O.preprocess(p, pts, lab, prb, timestamp=30)
    removed = (v1[..., 3] > 0) & (v0[..., 3] == 0)
    assert removed.sum() > 50
    # Q1: the label stored for the winning point i is labels[i+4]
    p2 = O.default_params(**sized(900), label_offset_quirk=0)
    lab_shift = np.r_[lab[4:], np.zeros(4, np.float32)]
    prb_shift = np.r_[prb[5:], np.zeros(5, np.float32)]
    va, na, sa = O.preprocess(p, pts, lab, prb, timestamp=30)
    vb, nb, sb = O.preprocess(p2, pts, lab_shift, prb_shift, timestamp=30)
    assert np.array_equal(sa, sb) and np.array_equal(va, vb)


def _frames(width=900, n=2, semantic=False):
    p = O.default_params(**sized(width))
    sc, poses = scans(width, n=n, semantic=semantic)
    return p, [O.preprocess(p, *s) for s in sc], poses


def test_icp_invariants():
    p, fr, poses = _frames()
    o48, raw = O.icp_jacobian(p, fr[1], fr[0], np.eye(4))
    JtJ = o48[:36].reshape(6, 6)
    assert np.array_equal(JtJ, JtJ.T)
    assert np.all(np.linalg.eigvalsh(JtJ) > 0)
    assert o48[42] + o48[46] == 900 * 64          # n_valid + n_invalid = P
    assert o48[44] <= o48[42] and o48[45] <= o48[43] + 1e-12
    assert np.array_equal(O.icp_unpack(raw), o48)
    # zero motion with nearest sampling: every pixel is associated with itself -> zero residual, GN stays put
    pn = O.default_params(**sized(900), bilinear_sampling=0)
    o48, _ = O.icp_jacobian(pn, fr[0], fr[0], np.eye(4))
    assert o48[43] == 0.0 and np.abs(o48[36:42]).max() == 0.0 and o48[44] == 0
    pose, _, k, _ = O.icp_minimize(pn, fr[0], fr[0], np.eye(4))
    assert k == 0 and np.array_equal(pose, np.eye(4))
    # row stripes add up exactly
    tot = sum(O.icp_jacobian(p, fr[1], fr[0], np.eye(4), rows=r)[1] for r in [(0, 10), (10, 40), (40, 64)])
    assert np.array_equal(tot, raw)


def test_icp_recovers_known_motion():
    p, fr, poses = _frames()
    gt = np.linalg.inv(poses[0]) @ poses[1]      # 1.0 m + 0.5 deg
    pose, o48, k, hist = O.icp_minimize(p, fr[1], fr[0], np.eye(4))
    assert np.linalg.norm(pose[:3, 3] - gt[:3, 3]) < 0.03
    assert abs(np.arccos((np.trace(pose[:3, :3].T @ gt[:3, :3]) - 1) / 2)) < 2e-3
    assert len(hist) == k + 1
    # exactly max_iterations passes when the stop tests are disabled
    p2 = O.default_params(**sized(900), max_iterations=10, stopping_threshold=0.0, delta=0.0)
    pose, o48, k, hist = O.icp_minimize(p2, fr[1], fr[0], np.eye(4))
    assert k == 10 and len(hist) == 11


def test_reference_fp32_blending_lies_within_1e5_of_exact_sums():
    # the GL path sums in fp32 (64-pixel partial sums + ROP adds); the exact fixed-point sums must agree to 1e-5 rel.
    p, fr, poses = _frames(semantic=True)
    T = np.linalg.inv(poses[0]) @ poses[1]
    o48, _ = O.icp_jacobian(p, fr[1], fr[0], T)
    f48 = O.icp_jacobian_fp32gl(p, fr[1], fr[0], T)
    for i in list(range(0, 36, 7)) + [43, 45]:    # diagonal of JtJ, F, F_inlier
        assert abs(f48[i] - o48[i]) <= 1e-5 * abs(o48[i]) + 1e-6, (i, f48[i], o48[i])
    assert np.array_equal(f48[[42, 44, 46]].astype(np.float64), o48[[42, 44, 46]])


def test_first_scan_surfels_match_predicate_and_order():
    p, fr, _ = _frames()
    m = O.Map(p)
    m.update(np.eye(4, dtype=np.float32), fr[0])
    v, n, s = fr[0]
    idx, rad, integ, nu, nn = m.update_debug()
    view = -v[..., :3] / np.linalg.norm(v[..., :3], axis=-1, keepdims=True).clip(1e-30)
    pred = (v[..., 3] >= 1) & (n[..., 3] >= 1) & (rad[..., 3] >= 0.5) & ((n[..., :3] * view).sum(-1) > 0.01)
    assert nu == 0 and abs(nn - int(pred.sum())) <= 2 and m.size() == nn
    sf = m.download()
    # x-major order (SurfelMap.cpp:88-92): positions equal the vertex map walked column by column
    cols = np.argwhere(pred.T)                   # (x, y) sorted by x then y
    exp = v[cols[:, 1], cols[:, 0], :3]
    if exp.shape[0] == sf.shape[0]:
        assert np.array_equal(np.stack([sf["x"], sf["y"], sf["z"]], 1), exp)
    assert np.all(sf["timestamp"] == 0) and np.all(sf["count"] == 0) and np.all(sf["color"] == 255.0)
    assert np.allclose(sf["confidence"], 0.0)     # log odds of p_prior = 0.5
    assert sf["radius"].min() >= p.min_radius - 1e-7 and sf["radius"].max() <= p.max_radius + 1e-7


def test_map_render_reproduces_first_scan_and_update_integrates_second():
    p, fr, poses = _frames()
    m = O.Map(p)
    I = np.eye(4, dtype=np.float32)
    m.update(I, fr[0])
    rv, rn, rs = m.render(I, I, -10.0)
    v = fr[0][0]
    both = (rv[..., 3] > 0) & (v[..., 3] > 0)
    assert both.sum() > 0.6 * (v[..., 3] > 0).sum()
    d = np.linalg.norm(rv[..., :3] - v[..., :3], axis=-1)[both]
    assert np.median(d) < 0.05
    n0 = m.size()
    T = (np.linalg.inv(poses[0]) @ poses[1]).astype(np.float32)
    m.update(T, fr[1])
    idx, rad, integ, nu, nn = m.update_debug()
    assert integ.sum() > 0.3 * (fr[1][0][..., 3] > 0).sum()      # most measurements are explained by the map
    assert nu <= n0 and m.size() == nu + nn
    sf = m.download()
    assert (sf["timestamp"] == 1).sum() >= integ.sum() * 0.5
    # order preserved: surviving old surfels keep their relative order (count==0 block first, then new ones)
    assert np.all(np.diff((sf["count"] == 1).astype(int)) >= 0)


def test_submap_shift_and_extraction():
    p = O.default_params(**sized(900))
    _, fr, _ = _frames()
    m = O.Map(p)
    I = np.eye(4, dtype=np.float32)
    m.update(I, fr[0])
    assert m.submap_origin() == (0, 0, 0)
    T = I.copy(); T[0, 3] = 11.5                  # > 1.1 * submap extent (10 m)
    m.update(T, fr[0])
    oi, oj, pending = m.submap_origin()
    assert (oi, oj) == (1, 0) and pending == 8    # 9 tiles queued, one extracted per update (partial extraction)


def test_oracle_results_do_not_depend_on_the_thread_count():
    """The oracle's OpenMP loops (integer sums, per-thread raster targets merged in buffer order, ordered chunk
    compaction) must reproduce the sequential result bit for bit: it is the checker for the CUDA path and the timed
    CPU arm of bench.py at the same time."""
    import hashlib
    from helpers import scans, sized
    pp = O.default_params(**sized(450))
    sc, _ = scans(450, n=6)

    def digest(threads):
        O.set_threads(threads)
        s = O.Slam(pp)
        h = hashlib.sha256()
        for a in sc:
            s.process_scan(*a)
            h.update(s.pose().tobytes())
        m = O.Map(pp, handle=O.lib().orc_slam_map(s.h))
        h.update(m.download().tobytes())
        for k in (0, 1):
            for img in s.frame(k):
                h.update(img.tobytes())
        return h.hexdigest()

    try:
        ref = digest(1)
        assert digest(2) == ref
        assert digest(5) == ref
    finally:
        O.set_threads(0)


def _ground_plane_frame(p, W, H, height=2.0):
    """range image of the plane z = -height seen from the origin, built from the pixel-centre beam directions (the
    inverse of the shaders' projection: x = 0.5 (1 - yaw/pi), y = 1 - (pitch_deg + fov_up) / fov, pitch = -asin(z/d))"""
    fov_up = abs(p.data_fov_up)
    fov = fov_up + abs(p.data_fov_down)
    xs = (np.arange(W) + 0.5) / W
    ys = (np.arange(H) + 0.5) / H
    yaw = -(2 * xs - 1) * np.pi
    elev = -np.deg2rad((1 - ys) * fov - fov_up)
    YAW, EL = np.meshgrid(yaw, elev)
    d = np.stack([np.cos(EL) * np.cos(YAW), np.cos(EL) * np.sin(YAW), np.sin(EL)], -1)
    r = np.where(EL < -0.02, height / np.maximum(-np.sin(EL), 1e-9), 0.0)
    valid = (r > p.min_depth) & (r < p.max_depth)
    V = np.zeros((H, W, 4), np.float32)
    N = np.zeros((H, W, 4), np.float32)
    V[..., :3] = d * r[..., None]
    V[..., 3] = 1
    N[..., 2] = 1
    N[..., 3] = 1
    V[~valid] = 0
    N[~valid] = 0
    return (V, N, np.zeros((H, W, 4), np.float32)), int(valid.sum())


@pytest.mark.parametrize("dz", [0.01, -0.02])
def test_icp_point_to_plane_terms_on_a_plane_are_the_analytic_ones(dz):
    """Closed form for K5: data = model = a horizontal plane with normal (0,0,1), pose = pure z translation dz.
    Whatever pixel a point is associated with, its point-to-plane residual is n.(T p - q) = dz, its translational
    Jacobian row is n = (0,0,1): F = n dz^2, (JtJ)_zz = n, (Jtf)_z = n dz, no x/y terms, and the Gauss-Newton step
    solves to exactly -dz (Frame2Model_jacobians.geom:150-185, LieGaussNewton.cpp:60-68)."""
    W, H = 900, 64
    p = O.default_params(**sized(W), weighting=0, bilinear_sampling=0)
    frame, n_px = _ground_plane_frame(p, W, H)
    T = np.eye(4)
    T[2, 3] = dz
    o48, _ = O.icp_jacobian(p, frame, frame, T)
    n_valid, n_out, n_invalid = o48[42], o48[44], o48[46]
    assert n_valid + n_invalid == W * H and n_out == 0
    assert 0.9 * n_px <= n_valid <= n_px          # a band of rows leaves the image when the plane moves
    JtJ = o48[:36].reshape(6, 6)
    assert JtJ[2, 2] == n_valid                   # sum of n_z^2 with unit weights
    assert abs(o48[43] - n_valid * dz * dz) <= 1e-4 * n_valid * dz * dz
    assert abs(o48[45] - o48[43]) == 0            # every valid pixel is an inlier
    assert abs(o48[38] - n_valid * dz) <= 1e-4 * abs(n_valid * dz)
    assert np.abs(JtJ[:2, :]).max() == 0 and np.abs(o48[36:38]).max() == 0    # no x / y translation terms
    assert JtJ[5, 5] == 0 and o48[41] == 0                                     # rotation about z is unobservable
    step = O.ldlt_solve6(JtJ, -o48[36:42])
    assert abs(step[2] + dz) < 1e-5 * abs(dz) + 1e-7 and np.abs(np.delete(step, 2)).max() < 1e-6


def test_movable_label_set_is_the_reference_shaders():
    """tests/golden/reference_movable_labels.json is parsed out of the reference's shaders (generator beside it): the
    label ids gen_vertexmap.vert removes during the first ten scans, update_surfels.vert penalises and gen_surfels.geom
    starts with a lower confidence. The oracle's K1 must remove exactly those; the CUDA predicate is the same literal
    list (checked in the source, the GPU parity tests then compare behaviour)."""
    import json
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    fx = json.load(open(os.path.join(root, "tests", "golden", "reference_movable_labels.json")))["passes"]
    sets = [set(v["movable"]) for v in fx.values()]
    assert len(sets) == 3 and sets[0] == sets[1] == sets[2] and len(sets[0]) == 9
    movable = sets[0]
    # one point per label value 0..259, each in its own pixel column of the middle row
    W, H = 900, 64
    p = O.default_params(**sized(W), label_offset_quirk=0)
    n = 260
    az = -np.pi + (np.arange(n) * 3 + 1.5) * (2 * np.pi / W)
    el = np.deg2rad(-11.0)
    pts = np.zeros((n, 4), np.float32)
    pts[:, 0] = 10 * np.cos(el) * np.cos(az)
    pts[:, 1] = 10 * np.cos(el) * np.sin(az)
    pts[:, 2] = 10 * np.sin(el)
    lab = np.arange(n, dtype=np.float32)
    prb = np.full(n, 0.9, np.float32)
    v_first, _, _ = O.preprocess(p, pts, lab, prb, timestamp=0)
    v_later, _, s_later = O.preprocess(p, pts, lab, prb, timestamp=10)
    assert int((v_later[..., 3] > 0).sum()) == n                 # every point owns a pixel
    kept_first = (v_first[..., 3] > 0)
    labels_kept = set(np.rint(s_later[kept_first][:, 0] * 255.0).astype(int).tolist())
    labels_all = set(np.rint(s_later[v_later[..., 3] > 0][:, 0] * 255.0).astype(int).tolist())
    assert labels_all == set(range(n))
    assert labels_all - labels_kept == {int(m) for m in movable}
    # the CUDA predicate and the oracle predicate are the same literal list
    for path in ("semantic_suma_b200/csrc/sb_math.cuh", "oracle/orc_internal.h"):
        src = open(os.path.join(root, path)).read()
        body = re.search(r"is_movable\(float l\)\s*\{(.*?)\}", src, flags=re.S).group(1)
        assert {float(x) for x in re.findall(r"l == ([\d.]+)f", body)} == movable, path


def test_loop_closure_twin_bookkeeping_without_a_loop():
    """oracle/orc_loop.cpp: with checkLoopClosure attached, a short straight drive records one odometry edge per scan,
    never finds a candidate (nothing is 100 scans old) and leaves the poses untouched (SurfelMapping.cpp:460-471, 478-495)"""
    p = O.default_params(**sized(360, 32))
    sc, poses = scans(360, 32, n=5)
    a, b = O.Slam(p), O.Slam(p)
    b.enable_loop_closure(search_distance=3.0, min_trajectory_distance=1.0, min_verifications=1)
    for s in sc:
        a.process_scan(*s)
        b.process_scan(*s)
    assert np.array_equal(a.pose(), b.pose())
    li = b.loop_info()
    assert (li["n_edges"], li["n_poses"], li["candidates_tested"], li["found_candidate"]) == (4, 5, 0, 0)
    assert li["time_without_loop_closure"] == 4
    edges = b.loop_edges()
    assert [(f, t) for f, t, _ in edges] == [(0, 1), (1, 2), (2, 3), (3, 4)]
    chain = np.eye(4)
    for _, _, rel in edges:
        chain = chain @ rel
    assert np.allclose(chain, b.pose(), atol=1e-9)
