"""1-GPU loop-back of the fused peer all-reduce (SUMA_B200_SELF_COMM=1): same kernel path, mailbox = own memory."""
import faulthandler, os, sys, time, ctypes as C, collections, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["SUMA_B200_SELF_COMM"] = "1"
import numpy as np
faulthandler.enable(); faulthandler.dump_traceback_later(40, exit=True)
from semantic_suma_b200 import api
from helpers import scans, sized
size = int(sys.argv[1]) if len(sys.argv) > 1 else 900
pp = api.default_params(**sized(size), max_iterations=8, stopping_threshold=0.0, delta=0.0)
sc, _ = scans(size, n=4)
solo = api.SurfelMapping(pp, device=0)
for s in sc: solo.processScan(*s)
ref = solo.getCurrentPose().copy(); solo.ctx.close()
sl = api.SurfelMapping(pp, device=0)
h = np.zeros(64, np.uint8)
L = api.lib()
sl.ctx.check(L.sb_comm_export(sl.ctx.h, C.c_void_p(h.ctypes.data)), "export")
sl.ctx.check(L.sb_comm_init(sl.ctx.h, 0, 1, C.c_void_p(h.ctypes.data), 0, pp.data_height), "init")
def dump():
    time.sleep(10)
    t = np.zeros(256, np.uint64)
    L.sb_debug_icp_trace.argtypes = [C.c_void_p, C.c_void_p]
    rc = L.sb_debug_icp_trace(sl.ctx.h, C.c_void_p(t.ctypes.data))
    print("HANG? TRACE rc", rc, "dbg[240..251] =", [int(x) for x in t[240:252]], flush=True)
    b = np.zeros(2048, np.uint64)
    L.sb_debug_icp_block_states.argtypes = [C.c_void_p, C.c_void_p]
    L.sb_debug_icp_block_states(sl.ctx.h, C.c_void_p(b.ctypes.data))
    print(" main states", sorted(collections.Counter(b[:232].astype(np.int64).tolist()).items()),
          "top states", sorted(collections.Counter(b[1024:1024+232].astype(np.int64).tolist()).items()), flush=True)
if os.environ.get("SUMA_B200_ICP_TRACE"):
    threading.Thread(target=dump, daemon=True).start()
for rep in range(int(os.environ.get("REPS", "20"))):
    for i, s in enumerate(sc): sl.processScan(*s)
    if rep == 0: print("equal to solo:", np.array_equal(sl.getCurrentPose(), ref), flush=True)
    sl.reset() if hasattr(sl, "reset") else None
print("SELF-COMM OK", flush=True)
