"""Dry run of tests/test_zz_gpu_late.py WITHOUT a GPU: the api classes those tests touch are replaced by stand-ins that run
the CPU oracle, so that the TEST LOGIC (keys, call order, counts, digests) is exercised end to end. It says nothing about the
CUDA path -- only that a correct CUDA path would pass these tests. Run from the repository root."""
import sys, types
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np
from oracle import oracle as O
from semantic_suma_b200 import api as real_api


class Ctx:
    def __init__(self, pp):
        self.pp = pp
        self.po = O.default_params()
        for name, _ in type(pp)._fields_:           # same field names in both parameter blocks
            if hasattr(self.po, name):
                setattr(self.po, name, getattr(pp, name))
        self.h = 1
    def close(self):
        self.h = None


class Frame:
    def __init__(self, ctx, w=None, h=None, arrays=None):
        self.ctx = ctx; self.a = arrays or [np.zeros((h, w, 4), np.float32) for _ in range(3)]
    def maps(self):
        return tuple(self.a)


class Preprocessing:
    def __init__(self, ctx): self.ctx = ctx
    def process(self, pts, frame, labels=None, probs=None, timestamp=100):
        frame.a = list(O.preprocess(self.ctx.po, pts, labels, probs, timestamp=timestamp))


class SurfelMap:
    def __init__(self, ctx, m=None): self.ctx = ctx; self.m = m or O.Map(ctx.po)
    def update(self, T, frame): self.m.update(T, frame.maps())
    def render(self, Po, Pn, out, ct): out.a = list(self.m.render(Po, Pn, ct))
    def getAllSurfels(self): return self.m.download()
    def size(self): return self.m.size()


class SurfelMapping:
    def __init__(self, pp, device=0):
        self.ctx = Ctx(pp); self.s = O.Slam(self.ctx.po)
    def enableLoopClosure(self, on=True, **kw): self.s.enable_loop_closure(**kw)
    def processScan(self, pts, lab=None, prb=None): self.s.process_scan(pts, lab, prb)
    def getCurrentPose(self): return self.s.pose()
    def getStatistics(self):
        st = self.s.stats()
        return {"num_iterations": st["iterations"], "F": st["F"], "inlier": st["inlier"], "outlier": st["outlier"],
                "invalid": st["invalid"], "track_loss": st["track_loss"]}
    def getMap(self): return SurfelMap(self.ctx, self.s.map)
    def getLoopInfo(self): return self.s.loop_info()
    def integrateLoopClosures(self, poses=None): return self.s.integrate_loop_closures(poses)
    def getCurrentFrame(self): return Frame(self.ctx, arrays=list(self.s.frame(0)))
    def getLastModelFrame(self): return Frame(self.ctx, arrays=list(self.s.frame(1)))


fake = types.ModuleType("semantic_suma_b200.api")
for k in ("default_params", "Params", "colmajor", "from_colmajor", "lib", "SURFEL_DTYPE", "XML_KEYS"):
    setattr(fake, k, getattr(real_api, k))
fake.Context, fake.Frame, fake.Preprocessing, fake.SurfelMap, fake.SurfelMapping = Ctx, Frame, Preprocessing, SurfelMap, SurfelMapping
sys.modules["semantic_suma_b200.api"] = fake
import semantic_suma_b200
semantic_suma_b200.api = fake
import helpers
helpers.api = fake

import test_zz_gpu_late as T
for name in ("test_cuda_matches_the_reference_generated_golden", "test_parameter_branches_bit_exact",
             "test_image_geometries_bit_exact", "test_loop_closure_integration_bit_exact"):
    getattr(T, name)()
    print("dry run ok:", name, flush=True)
