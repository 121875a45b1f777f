"""TEST INFRASTRUCTURE: a peer that never shows up. Two ranks are set up for the in-kernel exchange, only rank 0 runs: its
persistent Gauss-Newton kernel must give up after the bounded spin (10 s of %globaltimer; the executor's clock runs 64 x faster
here because its lanes time out one after the other), the scan must come back as an error
(SB_ERR_STATE) instead of hanging the stream, and the context must still close. ("spins with no timeout".)"""
import ctypes as C
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
os.environ.setdefault("CUSIM_CLOCK_SCALE", "64")  # lanes time out one after the other on the executor (cusim_rt.cpp, globaltimer)
import numpy as np  # noqa: E402
from cusim import build_sim  # noqa: E402

path = build_sim.build()
from semantic_suma_b200 import build as product_build  # noqa: E402

product_build.LIB = path
product_build.build = lambda *a, **k: path
from semantic_suma_b200 import api, stripes, synth  # noqa: E402


def main():
    pp = api.default_params(data_width=450, model_width=450, max_iterations=4, stopping_threshold=0.0, delta=0.0)
    L = api.lib()
    sl = [api.SurfelMapping(pp) for _ in range(2)]
    handles = np.zeros((2, 64), np.uint8)
    for r in range(2):
        sl[r].ctx.check(L.sb_comm_export(sl[r].ctx.h, C.c_void_p(handles[r].ctypes.data)), "export")
    for r in range(2):
        r0, r1 = stripes.row_stripe(r, 2, 64)
        sl[r].ctx.check(L.sb_comm_init(sl[r].ctx.h, r, 2, C.c_void_p(handles.ctypes.data), r0, r1), "init")
    scene = synth.Scene(width=450, height=64)
    poses = synth.trajectory(2)
    sl[0].processScan(*scene.scan(0, poses[0]))  # first scan: no minimisation, no exchange
    t0 = time.time()
    try:
        sl[0].processScan(*scene.scan(1, poses[1]))
    except api.SumaError as e:
        dt = time.time() - t0
        assert 0.1 < dt < 120.0, dt
        print("peer timeout ok: the scan returned %r after %.1f s" % (str(e)[:120], dt))
    else:
        raise AssertionError("the scan succeeded without its peer")
    for s in sl:
        s.ctx.close()
    print("contexts closed")


if __name__ == "__main__":
    main()
