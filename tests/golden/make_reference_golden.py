"""Generates tests/golden/reference_golden.json: SHA-256 digests of the outputs of THE REFERENCE ITSELF -- its core classes
and shaders, compiled where they lie under /root/reference into oracle/_ref/libsuma_ref_full.so and run on the software
GL -- on seeded synthetic inputs. Runs only where /root/reference exists (the build container); the
JSON is committed so that the checks travel: tests/test_golden.py holds the oracle to it on every box.

Comparable with these digests: everything the reference computes without its fp32 blending of the 48 ICP values
(preprocessing, map update / rendering at given poses -- also the CUDA path's results), and whole processScan runs with
the oracle adding those 48 values the GL way (O.gl_sums(1)).

    python tests/golden/make_reference_golden.py      # rewrites the JSON
"""
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
from golden.make_golden import digest, surfel_digest  # noqa: E402

TOUR = [(0, 0), (12, 0), (36, 0), (60, 0), (84, 0), (60, 24), (12, 0)]  # shifts the submap window, re-inserts tiles


def pose_at(x, y):
    T = np.eye(4, dtype=np.float32)
    T[0, 3], T[1, 3] = x, y
    return T


class OracleEngine:
    """the same calls on oracle/ (what tests/test_golden.py runs against the committed digests)"""

    def __init__(self, p):
        self.p = p
        self.map = O.Map(p)
        self.slam = None

    def preprocess(self, pts, lab, prb, timestamp):
        return O.preprocess(self.p, pts, lab, prb, timestamp=timestamp)

    def maps(self, frame):
        return frame

    def close(self):
        pass

    def map_update(self, T, frame):
        self.map.update(T, frame)

    def map_render(self, T, ct):
        return self.map.render(T, T, ct)

    def map_surfels(self):
        return self.map.download()

    def process_scan(self, pts, lab, prb):
        if self.slam is None:
            self.slam = O.Slam(self.p)
        self.slam.process_scan(pts, lab, prb)

    def slam_state(self):
        return self.slam.pose(), self.slam.map.download(), self.slam.stats()["iterations"]


class ReferenceEngine:
    def __init__(self, p):
        from oracle import ref as R
        self.f = R.Full(p)

    def preprocess(self, pts, lab, prb, timestamp):
        return self.f.preprocess(pts, lab, prb, timestamp=timestamp)

    def maps(self, frame):
        return frame

    def close(self):
        pass

    def map_update(self, T, frame):
        self.f.map_update(T, frame)

    def map_render(self, T, ct):
        return self.f.map_render(T, T, ct)

    def map_surfels(self):
        return self.f.map_download()

    def process_scan(self, pts, lab, prb):
        self.f.process_scan(pts, lab, prb)

    def slam_state(self):
        it = self.f.statistic("num_iterations")
        return self.f.pose(), self.f.map_download(), 0.0 if it != it else it


def compute(make_engine, process_scan=True):
    """process_scan=False: only what does not involve the 48 blended ICP values (what the CUDA path must reproduce too)"""
    out = {}
    for semantic in (False, True):
        tag = "semantic" if semantic else "geometric"
        p = O.default_params(**sized(900))
        sc, _ = scans(900, n=4, semantic=semantic)
        e = make_engine(p)
        frames = [e.preprocess(*s, timestamp=t * 7) for t, s in enumerate(sc)]
        out["preprocess_%s" % tag] = [digest(x) for f in frames for x in e.maps(f)]
        sizes = []
        for t, (x, y) in enumerate(TOUR):
            e.map_update(pose_at(x, y), frames[t % 4])
            sizes.append(int(e.map_surfels().shape[0]))
        out["map_%s" % tag] = {"sizes": sizes, "surfel_digest": surfel_digest(e.map_surfels()),
                               "render": [digest(x) for x in e.map_render(pose_at(10, 2), -5.0)]}
        e.close()
        if not process_scan:
            continue
        e = make_engine(p)
        run = []
        for s in sc:
            e.process_scan(*s)
            pose, surfels, it = e.slam_state()
            run.append({"pose": digest(pose), "surfels": int(surfels.shape[0]), "surfel_digest": surfel_digest(surfels),
                        "iterations": int(it)})
        out["process_scan_gl_sums_%s" % tag] = run
    return out


if __name__ == "__main__":
    g = compute(ReferenceEngine)
    with open(os.path.join(HERE, "reference_golden.json"), "w") as f:
        json.dump(g, f, indent=1, sort_keys=True)
    print("wrote reference_golden.json:", {k: (len(v) if isinstance(v, list) else "...") for k, v in g.items()})
