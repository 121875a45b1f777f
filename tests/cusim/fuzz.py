"""TEST INFRASTRUCTURE: fuzzing the CUDA sources (on the CPU executor) against the oracle, beyond the fixed cases of the suite:
 (a) adversarial random surfel / point clouds (tests/test_zz_gpu_late.py::test_adversarial_random_clouds_bit_exact) over many
     seeds and over image sizes incl. odd ones (361x33, 128x16, 513x48, 900x64, 2048x64);
 (b) random COMBINATIONS of the parameter variants (tests/test_ref_full.py::VARIANTS, 2-5 at a time), geometric / semantic,
     three widths, three speeds, with and without a jump that triggers the track-loss fallback: five scans of the whole
     pipeline each;
 (c) the Jacobian operator (raw 32 int64 sums) on pairs of random clouds: 2 sampling modes x 4 robust weightings x geometric /
     semantic x three image sizes, three poses each;
 (d) whole Gauss-Newton minimisations (persistent kernel): real scene pairs and random clouds (degenerate systems, early
     stops), 3 weightings x 2 sampling modes x max iterations 1 / 7 / 30 x stop thresholds 0 / 1e-4 x three image sizes.
 (e) submap paging: random walks (28 scans, random steps and turns, U-turns back over extracted tiles) with small submap
     windows (extent 2 / 3 / 5 m, dimension 1 / 2, partial extraction on / off): the whole pipeline against the oracle.
 (f) pathological scans: 2 % of the points carry NaN / +-inf / +-1e30 / denormal coordinates, NaN probabilities, NaN / huge /
     negative labels -- whole pipeline, four scans, three widths (the executor converts float -> integer the PTX way: NaN -> 0,
     saturation, where x86 returns 0x8000...; run it with CUSIM_ASAN=1 for the memcheck of these inputs).
usage: python tests/cusim/fuzz.py [n_cloud_seeds=40] [n_param_combos=40] [n_walks=12]"""
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from cusim import build_sim  # noqa: E402

path = build_sim.build()
from semantic_suma_b200 import build as product_build  # noqa: E402

product_build.LIB = path
product_build.build = lambda *a, **k: path
import test_zz_gpu_late as T  # noqa: E402
from helpers import both_params, sized  # noqa: E402
from semantic_suma_b200 import api, synth  # noqa: E402
from test_ref_full import VARIANTS  # noqa: E402


def clouds(n_seeds):
    real_sized, real_frame = T.sized, api.Frame
    n = bad = 0
    try:
        for (W, H) in ((360, 32), (361, 33), (128, 16), (513, 48), (900, 64), (2048, 64)):
            T.sized = lambda w, h=64, _W=W, _H=H, **kw: real_sized(_W, _H, **kw)
            api.Frame = lambda ctx, w=None, h=None, handle=None, _W=W, _H=H: (
                real_frame(ctx, _W, _H) if handle is None else real_frame(ctx, handle=handle))
            for seed in range(100, 100 + n_seeds):
                for (t_now, compose) in ((150, 1), (40, 0), (7, 1)):
                    n += 1
                    try:
                        T.test_adversarial_random_clouds_bit_exact(t_now, compose, seed)
                    except AssertionError as e:
                        bad += 1
                        print("FAIL clouds %dx%d seed %d t_now %d compose %d: %s" % (H, W, seed, t_now, compose, str(e)[:300]))
    finally:
        T.sized, api.Frame = real_sized, real_frame
    return n, bad


def params(n_combos):
    rnd = random.Random(7)
    bad = 0
    for i in range(n_combos):
        kw = {}
        for v in rnd.sample(VARIANTS, rnd.randint(2, 5)):
            kw.update(v)
        W = rnd.choice((450, 360, 512))
        sem = rnd.random() < 0.6
        scene = synth.Scene(width=W, height=64, semantic=sem, seed=100 + i)
        poses = [p.copy() for p in synth.trajectory(5, step=rnd.choice((0.1, 0.3, 0.6)), yaw_deg=rnd.choice((0.0, 1.0, 4.0)))]
        if rnd.random() < 0.3:
            poses[3][0, 3] += 1.2
            poses[4][0, 3] += 1.2
        sc = [scene.scan(t, poses[t]) for t in range(5)]
        po, pp = both_params(**sized(W), **kw)
        try:
            T._pipeline_equal(po, pp, sc, repr(kw))
        except AssertionError as e:
            bad += 1
            print("FAIL params W=%d semantic=%r %r: %s" % (W, sem, kw, str(e)[:300]))
    return n_combos, bad


def paging(n_walks):
    import numpy as np
    rnd = random.Random(3)
    bad = 0
    for i in range(n_walks):
        kw = dict(submap_extent=rnd.choice((2.0, 3.0, 5.0)), submap_dimension=rnd.choice((1, 2)), partial_extraction=rnd.choice((0, 1)))
        W = rnd.choice((360, 450))
        poses, yaw = [np.eye(4)], 0.0
        for t in range(1, 28):
            yaw += np.deg2rad(rnd.uniform(-6, 6))
            step = rnd.uniform(0.2, 0.7)
            P = poses[-1].copy()
            P[:2, :2] = [[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]]
            P[0, 3] += step * np.cos(yaw)
            P[1, 3] += step * np.sin(yaw)
            if rnd.random() < 0.1:
                yaw += np.pi * rnd.choice((0.5, 1.0))
            poses.append(P)
        scene = synth.Scene(width=W, height=64, semantic=rnd.random() < 0.5, seed=300 + i)
        sc = [scene.scan(t, poses[t]) for t in range(28)]
        po, pp = both_params(**sized(W), **kw)
        try:
            T._pipeline_equal(po, pp, sc, repr(kw))
        except AssertionError as e:
            bad += 1
            print("FAIL paging walk %d %r: %s" % (i, kw, str(e)[:300]))
    return n_walks, bad


def pathological():
    import numpy as np
    from helpers import bits
    from oracle import oracle as O
    rng = np.random.default_rng(3)
    n = bad = 0
    for rep in range(6):
        W = (360, 450, 361)[rep % 3]
        po, pp = both_params(**sized(W))
        scene = synth.Scene(width=W, height=64, semantic=True, seed=50 + rep)
        poses = synth.trajectory(4)
        g, o = api.SurfelMapping(pp), O.Slam(po)
        for t in range(4):
            pts, lab, prb = (a.copy() for a in scene.scan(t, poses[t]))
            m = pts.shape[0]
            idx = rng.integers(0, m, m // 50)
            vals = np.array([np.nan, np.inf, -np.inf, 1e30, -1e30, 1e-40, 0.0], np.float32)
            pts[idx, rng.integers(0, 3, idx.shape[0])] = rng.choice(vals, idx.shape[0])
            prb[rng.integers(0, m, 50)] = np.nan
            lab[rng.integers(0, m, 50)] = rng.choice(np.array([np.nan, 1e9, -5.0], np.float32), 50)
            g.processScan(pts, lab, prb)
            o.process_scan(pts, lab, prb)
            n += 1
            if not (np.array_equal(bits(g.getCurrentPose()), bits(o.pose())) and g.getMap().size() == o.map.size()):
                bad += 1
                print("FAIL pathological scan: width %d scan %d" % (W, t))
        g.ctx.close()
    return n, bad


def operators():
    import itertools

    import numpy as np
    from helpers import assert_bits_equal
    from oracle import oracle as O
    from test_gpu_parity import _prep_both
    from test_ref_shaders import _random_cloud
    rng = np.random.default_rng(11)
    n_j = bad_j = n_g = bad_g = 0
    for (W, H) in ((360, 32), (361, 33), (900, 64)):
        for bil, wt, sem in itertools.product((0, 1), (0, 1, 2, 3), (False, True)):
            po, pp = both_params(**sized(W, H, bilinear_sampling=bil, weighting=wt))
            ctx = api.Context(pp)
            for rep in range(3):
                o0, f0 = _prep_both(po, ctx, _random_cloud(rng, 25000, po, labels=sem), 100)
                o1, f1 = _prep_both(po, ctx, _random_cloud(rng, 25000, po, labels=sem), 100)
                obj = api.Frame2Model(ctx)
                obj.setData(f1, f0)
                T0 = np.eye(4); T0[:3, 3] = rng.normal(0, 0.3, 3)
                a = rng.normal(0, 0.05)
                T0[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]
                for it, pose in enumerate([np.eye(4), T0, T0]):
                    n_j += 1
                    obj.initialize(pose); obj.iteration_ = it; obj.jacobianProducts()
                    _, raw = O.icp_jacobian(po, o1, o0, pose, iteration=it)
                    if not np.array_equal(obj.raw32, raw):
                        bad_j += 1
                        print("FAIL jacobian %dx%d bilinear %d weighting %d semantic %r" % (H, W, bil, wt, sem))
            ctx.close()
    for (W, H) in ((360, 32), (450, 64), (361, 33)):
        for wt, bil, mi, eps in itertools.product((0, 1, 2), (0, 1), (1, 7, 30), (0.0, 1e-4)):
            po, pp = both_params(**sized(W, H, bilinear_sampling=bil, weighting=wt, max_iterations=mi, stopping_threshold=eps,
                                         delta=eps))
            ctx = api.Context(pp)
            for rep in range(2):
                if rep == 0:
                    scene = synth.Scene(width=W, height=H, seed=int(rng.integers(1, 1000)))
                    poses = synth.trajectory(2, step=float(rng.uniform(0.05, 0.8)), yaw_deg=float(rng.uniform(0, 3)))
                    c0, c1 = scene.scan(0, poses[0]), scene.scan(1, poses[1])
                else:
                    c0, c1 = _random_cloud(rng, 20000, po, labels=False), _random_cloud(rng, 20000, po, labels=False)
                o0, f0 = _prep_both(po, ctx, c0, 100)
                o1, f1 = _prep_both(po, ctx, c1, 100)
                obj = api.Frame2Model(ctx)
                obj.setData(f1, f0)
                gn = api.LieGaussNewton(ctx)
                T0 = np.eye(4); T0[:3, 3] = rng.normal(0, 0.05, 3)
                gn.minimize(obj, T0)
                pose_o, o48, k, hist = O.icp_minimize(po, o1, o0, T0)
                n_g += 1
                try:
                    assert gn.iterationCount() == k and len(gn.history()) == len(hist)
                    assert_bits_equal(gn.pose(), pose_o, "pose")
                    assert_bits_equal(gn.out48, o48, "out48")
                except AssertionError as e:
                    bad_g += 1
                    print("FAIL minimise %dx%d weighting %d bilinear %d max_iter %d eps %g: %s" % (H, W, wt, bil, mi, eps, str(e)[:200]))
            ctx.close()
    return n_j, bad_j, n_g, bad_g


def main():
    a = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    b = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    n1, bad1 = clouds(a)
    n2, bad2 = params(b)
    n_j, bad_j, n_g, bad_g = operators()
    n_w, bad_w = paging(int(sys.argv[3]) if len(sys.argv) > 3 else 12)
    n_p, bad_p = pathological()
    print("fuzz: %d adversarial cloud cases (6 image sizes), %d failures; %d parameter combinations x 5 scans, %d failures; "
          "%d Jacobian evaluations on random clouds, %d failures; %d Gauss-Newton minimisations, %d failures; "
          "%d paging random walks x 28 scans, %d failures; %d scans with NaN / inf / 1e30 points, %d failures"
          % (n1, bad1, n2, bad2, n_j, bad_j, n_g, bad_g, n_w, bad_w, n_p, bad_p))
    assert bad1 == 0 and bad2 == 0 and bad_j == 0 and bad_g == 0 and bad_w == 0 and bad_p == 0


if __name__ == "__main__":
    main()
