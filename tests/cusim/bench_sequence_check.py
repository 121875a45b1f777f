"""TEST INFRASTRUCTURE: the very sequence bench.py times (workload, seed, trajectory, parameters of `python bench.py`) through
the library's CUDA sources on the CPU executor against the oracle: poses after every scan, the statistics pass, the surfel
records at the end -- bit for bit, into the map sizes the benchmark runs at.

usage: python tests/cusim/bench_sequence_check.py [n_scans=40] [workload=hdl64_2048_geometric]"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

import numpy as np  # noqa: E402

from cusim import build_sim  # noqa: E402

path = build_sim.build()
from semantic_suma_b200 import build as product_build  # noqa: E402

product_build.LIB = path
product_build.build = lambda *a, **k: path
import bench  # noqa: E402
from helpers import assert_bits_equal, surfel_fields_equal  # noqa: E402
from oracle import oracle as O  # noqa: E402
from semantic_suma_b200 import api  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    wl = sys.argv[2] if len(sys.argv) > 2 else "hdl64_2048_geometric"
    w = bench.WORKLOADS[wl]
    scans = bench.generate_scans(w, n, seed=1337)  # rank 0's sequence
    kw = bench.param_kwargs(w)
    O.set_threads(min(8, os.cpu_count() or 1))
    osl, gsl = O.Slam(O.default_params(**kw)), api.SurfelMapping(api.default_params(**kw))
    t_sim = t_orc = 0.0
    for t, (p, l, q) in enumerate(scans):
        t0 = time.time(); osl.process_scan(p, l, q); t_orc += time.time() - t0
        t0 = time.time(); gsl.processScan(p, l, q); t_sim += time.time() - t0
        assert_bits_equal(gsl.getCurrentPose(), osl.pose(), "scan %d pose" % t)
        so, sg = osl.stats(), gsl.getStatistics()
        assert (sg["num_iterations"], sg["F"], sg["inlier"], sg["outlier"], sg["invalid"]) == \
            (so["iterations"], so["F"], so["inlier"], so["outlier"], so["invalid"]), "scan %d statistics" % t
        assert gsl.getMap().size() == osl.map.size(), "scan %d surfel count" % t
    surfel_fields_equal(gsl.getMap().getAllSurfels(), osl.map.download(), "surfels")
    n_surfels = gsl.getMap().size()
    # the calls of bench.py's end-to-end pass: reset, host buffers, the next scan staged ahead (sb_prefetch_scan)
    ref_poses = []
    osl2 = O.Slam(O.default_params(**kw))
    m = min(n, 8)
    for p, l, q in scans[:m]:
        osl2.process_scan(p, l, q)
        ref_poses.append(osl2.pose().copy())
    sem = w["semantic"]
    bufs = [(np.ascontiguousarray(p, np.float32), np.ascontiguousarray(l, np.float32) if sem else None,
             np.ascontiguousarray(q, np.float32) if sem else None) for p, l, q in scans[:m]]
    ptr = lambda a: a.ctypes.data if a is not None else 0  # noqa: E731
    for on_device in (False, True):  # "device" pointers are plain pointers on the executor
        gsl.reset()
        for t in range(m):
            if not on_device and t + 1 < m:
                pn, ln, qn = bufs[t + 1]
                gsl.prefetch_scan_raw(ptr(pn), ptr(ln), ptr(qn), pn.shape[0])
            p, l, q = bufs[t]
            gsl.process_scan_raw(ptr(p), ptr(l), ptr(q), p.shape[0], on_device)
            assert_bits_equal(gsl.getCurrentPose(), ref_poses[t], "after reset, scan %d (on_device=%r)" % (t, on_device))
    print("bench sequence ok: %s, %d scans, %d surfels at the end, every pose / statistic / surfel record bit-identical "
          "(executor %.1f s, oracle %.1f s); reset + host buffers + input staging: the first %d scans again, identical"
          % (wl, n, n_surfels, t_sim, t_orc, m))
    gsl.ctx.close()


if __name__ == "__main__":
    main()
