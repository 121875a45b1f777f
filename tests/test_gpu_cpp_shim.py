"""GPU test: a C++ caller written against the reference's class names (tests/cpp/shim_example.cpp, compiled against
include/suma_b200.hpp) gives the same bits as the Python mirror and therefore as the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

from semantic_suma_b200 import api
from helpers import scans

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _compile(tmp_path):
    exe = os.path.join(str(tmp_path), "shim_example")
    from semantic_suma_b200 import build as product_build
    api.lib()
    so = product_build.LIB  # libsuma_b200.so; under pytest --cusim the CPU executor's build of the same sources (conftest)
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-I" + os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "shim_example.cpp"), so, "-Wl,-rpath," + os.path.dirname(so)])
    return exe


def test_cpp_shim_compiles_without_gpu(tmp_path):
    _compile(tmp_path)


@pytest.mark.gpu
def test_cpp_shim_matches_python_mirror(tmp_path):
    exe = _compile(tmp_path)
    sc, _ = scans(900, n=5)
    path = os.path.join(str(tmp_path), "scans.bin")
    with open(path, "wb") as f:
        for pts, _, _ in sc:
            f.write(struct.pack("<I", pts.shape[0]))
            f.write(np.ascontiguousarray(pts, np.float32).tobytes())
    out = subprocess.check_output([exe, path, "900"], text=True)
    lines = {l.split()[0]: l.split()[1:] for l in out.strip().splitlines()}
    pose_cpp = api.from_colmajor(np.array([float(x) for x in lines["pose"]]))
    pp = api.default_params(data_width=900, model_width=900, max_iterations=10)
    sl = api.SurfelMapping(pp)
    for pts, _, _ in sc:
        sl.processScan(pts)
    assert np.array_equal(pose_cpp, sl.getCurrentPose())
    assert int(lines["surfels"][0]) == sl.getMap().size()
    # operator-level: frame-to-frame ICP of the last two scans
    ctx = sl.ctx
    a, b = api.Frame(ctx, 900, 64), api.Frame(ctx, 900, 64)
    pre = api.Preprocessing(ctx)
    pre.process(sc[-2][0], a, timestamp=100)
    pre.process(sc[-1][0], b, timestamp=100)
    obj = api.Frame2Model(ctx)
    obj.setData(b, a)
    gn = api.LieGaussNewton(ctx)
    gn.minimize(obj, np.eye(4))
    icp = lines["icp"]
    assert int(icp[0]) == gn.iterationCount()
    assert np.array_equal(api.from_colmajor(np.array([float(x) for x in icp[1:]])), gn.pose())
    assert int(lines["inlier"][0]) == obj.inlier()
    ctx.close()
