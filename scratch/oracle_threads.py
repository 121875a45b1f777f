import sys, time, hashlib, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, ctypes as C
from oracle import oracle as O
import bench
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "hdl64_2048_geometric"]
po = O.default_params(**bench.param_kwargs(w))
sc = bench.generate_scans(w, int(sys.argv[2]) if len(sys.argv) > 2 else 14, seed=1337)
def run(T):
    T = O.set_threads(T)
    s = O.Slam(po); h = hashlib.sha256(); t0=time.time()
    for i,(p,l,q) in enumerate(sc):
        s.process_scan(p,l,q)
        h.update(s.pose().tobytes())
    dt=time.time()-t0
    st=s.stats()
    mp = O.lib().orc_slam_map(s.h)
    n = O.lib().orc_map_size(C.c_void_p(mp))
    buf = np.zeros((n,16), np.float32)
    O.lib().orc_map_download(C.c_void_p(mp), buf.ctypes.data_as(C.c_void_p), C.c_uint32(n))
    h.update(buf.tobytes())
    for k in (0,1):
        for a in s.frame(k): h.update(a.tobytes())
    print("threads %d time %.2fs  n=%d  digest %s  last-scan phases %.3f %.3f %.3f"%(T, dt, n, h.hexdigest()[:16], st['t_preprocess'],st['t_icp'],st['t_mapping']), flush=True)
    return h.hexdigest()
r=[run(T) for T in (1,8,3)]
print("IDENTICAL" if len(set(r))==1 else "DIFFERENT")
