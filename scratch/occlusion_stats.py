"""Offline estimate: which fraction of the surfels that pass stage A of the rendering could be rejected by a
conservative box test against the FINAL depth image (upper bound of what an in-kernel early-Z can reach)."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from oracle import oracle as O
import bench

w = bench.WORKLOADS["hdl64_2048_geometric"]
N = int(sys.argv[1]) if len(sys.argv) > 1 else 70
scans = bench.generate_scans(w, N, seed=1337)
p = O.default_params(**bench.param_kwargs(w))
s = O.Slam(p)
poses = []
for i in range(N):
    s.process_scan(*scans[i])
    poses.append(s.pose())
surf = s.map.download()
print("surfels", surf.shape[0], surf.dtype.names)
P = poses[-1]
v, n, sem = s.frame(1)            # model frame rendered after the update at the current pose
W, H = w["width"], w["height"]
depth_img = np.linalg.norm(v[..., :3], axis=2)
depth_img[v[..., 3] < 0.5] = np.inf   # nothing rendered: cannot reject there
names = surf.dtype.names
print(names)
np.save('/tmp/surf.npy', surf); np.save('/tmp/poses.npy', np.array(poses)); np.save('/tmp/depth.npy', depth_img)
