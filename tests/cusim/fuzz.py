"""TEST INFRASTRUCTURE: fuzzing the CUDA sources (on the CPU executor) against the oracle, beyond the fixed cases of the suite:
 (a) adversarial random surfel / point clouds (tests/test_zz_gpu_late.py::test_adversarial_random_clouds_bit_exact) over many
     seeds and over image sizes incl. odd ones (361x33, 128x16, 513x48, 900x64, 2048x64);
 (b) random COMBINATIONS of the parameter variants (tests/test_ref_full.py::VARIANTS, 2-5 at a time), geometric / semantic,
     three widths, three speeds, with and without a jump that triggers the track-loss fallback: five scans of the whole
     pipeline each.
usage: python tests/cusim/fuzz.py [n_cloud_seeds=40] [n_param_combos=40]"""
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


def main():
    a = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    b = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    n1, bad1 = clouds(a)
    n2, bad2 = params(b)
    print("fuzz: %d adversarial cloud cases (6 image sizes), %d failures; %d parameter combinations x 5 scans, %d failures"
          % (n1, bad1, n2, bad2))
    assert bad1 == 0 and bad2 == 0


if __name__ == "__main__":
    main()
