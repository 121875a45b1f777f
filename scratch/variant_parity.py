"""Round-2 starter: parity + speed of an opt-in kernel variant against the default build, in one process per setting.
Usage on a B200:  python scratch/variant_parity.py SUMA_B200_RENDER_VARIANT 1 2
Runs the 64x2048 geometric sequence (30 scans) with the variable unset and with each value; prints a digest of poses,
surfels and model frames (must be equal) and the device time per scan."""
import hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import hashlib, json, os, sys, time
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np, torch
from semantic_suma_b200 import api
import bench
w = bench.WORKLOADS["hdl64_2048_geometric"]
scans = bench.generate_scans(w, 30, seed=1337)
pp = api.default_params(**bench.param_kwargs(w))
slam = api.SurfelMapping(pp, device=0)
h = hashlib.sha256()
dev = [torch.from_numpy(p).cuda() for p, _, _ in scans]
torch.cuda.synchronize()
t = []
for i, p in enumerate(dev):
    t0 = time.perf_counter()
    slam.process_scan_raw(p.data_ptr(), 0, 0, p.shape[0], True)
    t.append(time.perf_counter() - t0)
    h.update(slam.getCurrentPose().tobytes())
h.update(slam.getMap().getAllSurfels().tobytes())
for f in (slam.getCurrentModelFrame(), slam.getCurrentFrame()):
    for img in f.maps():
        h.update(np.nan_to_num(np.asarray(img), nan=7.0).tobytes())
print(json.dumps({"digest": h.hexdigest()[:20], "ms_per_scan_last20": round(1e3 * sum(t[10:]) / 20, 4),
                  "surfels": int(slam.getMap().size())}))
''' % (ROOT, ROOT)


def run(env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=600)
    if out.returncode != 0:
        return {"error": out.stderr[-800:]}
    return json.loads(out.stdout.strip().splitlines()[-1])


if __name__ == "__main__":
    var, values = sys.argv[1], sys.argv[2:]
    base = run({})
    print("default      ", base, flush=True)
    for v in values:
        r = run({var: v})
        same = r.get("digest") == base.get("digest")
        print("%s=%s" % (var, v), r, "IDENTICAL" if same else "DIFFERENT", flush=True)
