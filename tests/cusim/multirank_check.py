"""TEST INFRASTRUCTURE (run by tests/test_cusim.py in a child process): the row-striped Gauss-Newton with the IN-KERNEL peer
exchange on N "GPUs" of the CPU executor -- N contexts driven by N host threads, mailboxes exchanged through the library's
own sb_comm_export / sb_comm_init (a CUDA-IPC handle is a plain pointer here). Every rank must hold the bits of the
un-striped run after every scan: poses, iteration counts, surfel counts -- through a track-loss recovery as well.

usage: python tests/cusim/multirank_check.py N [width] [height] [scans] [jump]"""
import ctypes as C
import os
import sys
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import numpy as np  # noqa: E402

from cusim import build_sim  # noqa: E402

path = build_sim.build()
from semantic_suma_b200 import build as product_build  # noqa: E402

product_build.LIB = path
product_build.build = lambda *a, **k: path
from semantic_suma_b200 import api, stripes, synth  # noqa: E402


def main():
    n = int(sys.argv[1])
    width = int(sys.argv[2]) if len(sys.argv) > 2 else 900
    height = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    n_scans = int(sys.argv[4]) if len(sys.argv) > 4 else 4
    jump = len(sys.argv) > 5 and sys.argv[5] == "jump"
    kw = dict(data_width=width, model_width=width, data_height=height, model_height=height, max_iterations=6,
              stopping_threshold=0.0, delta=0.0)
    if height == 128:
        kw.update(data_fov_up=22.5, data_fov_down=-22.5, model_fov_up=22.5, model_fov_down=-22.5)
    pp = api.default_params(**kw)
    scene = synth.Scene(width=width, height=height, fov_up=kw.get("data_fov_up", 3.0), fov_down=kw.get("data_fov_down", -25.0))
    poses = synth.trajectory(n_scans)
    if jump:  # a 1.5 m jump: the track-loss test fires and the recovery minimisation runs striped as well
        poses = [p.copy() for p in poses]
        for t in range(n_scans - 1, len(poses)):
            poses[t][0, 3] += 1.5
    sc = [scene.scan(t, poses[t]) for t in range(n_scans)]

    solo = api.SurfelMapping(pp)
    ref = []
    for s in sc:
        solo.processScan(*s)
        st = solo.getStatistics()
        ref.append((solo.getCurrentPose().tobytes(), st["num_iterations"], solo.getMap().size(), st["track_loss"]))
    solo.ctx.close()

    L = api.lib()
    sl = [api.SurfelMapping(pp) for _ in range(n)]
    handles = np.zeros((n, 64), np.uint8)
    for r in range(n):
        sl[r].ctx.check(L.sb_comm_export(sl[r].ctx.h, C.c_void_p(handles[r].ctypes.data)), "export")
    for r in range(n):
        r0, r1 = stripes.row_stripe(r, n, height)
        sl[r].ctx.check(L.sb_comm_init(sl[r].ctx.h, r, n, C.c_void_p(handles.ctypes.data), r0, r1), "init")
    got = [[] for _ in range(n)]
    errors = []

    def run(r):
        try:
            for s in sc:
                sl[r].processScan(*s)
                st = sl[r].getStatistics()
                got[r].append((sl[r].getCurrentPose().tobytes(), st["num_iterations"], sl[r].getMap().size(),
                               st["track_loss"]))
        except Exception as e:  # noqa: BLE001
            errors.append("rank %d: %r" % (r, e))

    th = [threading.Thread(target=run, args=(r,)) for r in range(n)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors
    for r in range(n):
        for t in range(n_scans):
            assert got[r][t] == ref[t], "rank %d scan %d differs from the un-striped run: %r vs %r" % (r, t, got[r][t][1:], ref[t][1:])
    for r in range(n):
        sl[r].ctx.close()
    print("multirank ok: %d ranks, %dx%d, %d scans, track losses %r" % (n, height, width, n_scans, ref[-1][3]))


if __name__ == "__main__":
    main()
