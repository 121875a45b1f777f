"""Generates tests/golden/oracle_golden.json: SHA-256 digests (and a few scalars) of the oracle's outputs on seeded
synthetic inputs. The reference ships no golden vectors for this path; these digests pin the ORACLE (exact sums -- the
CUDA path's contract) against regressions. Digests of the outputs of the reference itself are in reference_golden.json
(make_reference_golden.py).

    python tests/golden/make_golden.py            # rewrites the JSON (run only when the oracle changes on purpose)
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as O  # noqa: E402
from helpers import scans, sized  # noqa: E402


def digest(a):
    a = np.ascontiguousarray(a)
    if a.dtype.kind == "f":  # canonical NaN payload
        v = a.view(np.uint32 if a.dtype == np.float32 else np.uint64).copy()
        v[np.isnan(a)] = 0x7fc00000 if a.dtype == np.float32 else 0x7ff8000000000000
        a = v
    return hashlib.sha256(a.tobytes()).hexdigest()


def surfel_digest(s):
    return digest(np.stack([s[f].astype(np.float32) if s[f].dtype != np.uint32 else s[f].view(np.float32)
                            for f in s.dtype.names], 1))


def compute():
    out = {}
    for semantic in (False, True):
        tag = "semantic" if semantic else "geometric"
        p = O.default_params(**sized(900))
        sc, poses = scans(900, n=4, semantic=semantic)
        fr = [O.preprocess(p, *s, timestamp=t * 7) for t, s in enumerate(sc)]
        out["preprocess_%s" % tag] = [digest(x) for f in fr for x in f]
        o48, raw = O.icp_jacobian(p, fr[1], fr[0], np.linalg.inv(poses[0]) @ poses[1], iteration=1)
        out["icp_raw32_%s" % tag] = [int(x) for x in raw]
        pose, o48, k, hist = O.icp_minimize(p, fr[1], fr[0], np.eye(4))
        out["icp_minimize_%s" % tag] = {"iterations": k, "pose": digest(pose), "F": float(o48[43])}
        sl = O.Slam(p)
        for s in sc:
            sl.process_scan(*s)
        out["slam_%s" % tag] = {"pose": digest(sl.pose()), "surfels": int(sl.map.size()),
                                "surfel_digest": surfel_digest(sl.map.download()),
                                "frame": [digest(x) for x in sl.frame(1)]}
    return out


if __name__ == "__main__":
    g = compute()
    with open(os.path.join(HERE, "oracle_golden.json"), "w") as f:
        json.dump(g, f, indent=1, sort_keys=True)
    print("wrote", len(g), "entries")
