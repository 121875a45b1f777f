"""TEST INFRASTRUCTURE: ctypes binding of oracle/_ref -- the reference's own GLSL shaders, transpiled to C++ from
/root/reference/src/shader by oracle/ref_harness/glsl2cpp.py and driven by a minimal software GL
(oracle/ref_harness/ref_pipeline.cpp). Two builds:

  mode="pinned"   GLSL built-ins follow the rules the oracle pins (orc_math.h)  -> must equal oracle/ bit for bit
  mode="precise"  GLSL built-ins in fp64 / libm                                 -> an independent legal GL

Only tests/ (and __graft_entry__.build(), which compiles it) may touch this module; the product never does.
/root/reference exists only in the build container: on the GPU box the prebuilt oracle/_ref/*.so are used as shipped.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_HARNESS = os.path.join(_HERE, "ref_harness")
_OUT = os.path.join(_HERE, "_ref")
REFERENCE = os.environ.get("SUMA_REFERENCE_DIR", "/root/reference")


def lib_path(mode):
    return os.path.join(_OUT, "libsuma_ref_%s.so" % mode)


def have_reference():
    return os.path.isdir(os.path.join(REFERENCE, "src", "shader"))


def available():
    return have_reference() or all(os.path.exists(lib_path(m)) for m in ("pinned", "precise"))


def build(force=False):
    """(re)build oracle/_ref from the reference's sources where they lie; no-op when /root/reference is absent"""
    if not have_reference():
        return available()
    cmd = ["make", "-C", _HARNESS, "-s", "REFERENCE=" + REFERENCE] + (["-B"] if force else [])
    subprocess.check_call(cmd)
    return True


_libs = {}


def lib(mode="pinned"):
    if mode not in _libs:
        build()
        L = C.CDLL(lib_path(mode))
        L.ref_map_create.restype = C.c_void_p
        L.ref_map_size.restype = C.c_uint32
        L.ref_map_timestamp.restype = C.c_uint32
        L.ref_map_download.restype = C.c_uint32
        L.ref_math_mode.restype = C.c_char_p
        assert L.ref_math_mode().decode() == mode
        _libs[mode] = L
    return _libs[mode]


_p, _f32, colmajor = O._p, O._f32, O.colmajor


def preprocess(p, pts, labels=None, probs=None, timestamp=100, mode="pinned"):
    pts = _f32(pts); labels = _f32(labels); probs = _f32(probs)
    H, W = p.data_height, p.data_width
    v = np.empty((H, W, 4), np.float32); nm = np.empty_like(v); s = np.empty_like(v)
    lib(mode).ref_preprocess(C.byref(p), _p(pts), _p(labels), _p(probs), C.c_uint32(pts.shape[0]), C.c_uint32(timestamp),
                             _p(v), _p(nm), _p(s))
    return v, nm, s


def icp_jacobian(p, data, model, pose, iteration=0, max_distance=None, max_angle=None, entries_per_kernel=64,
                 semantic=True, mode="pinned"):
    """the 2x8 RGB32F blend texture of Frame2Model::jacobianProducts as 48 floats (fp32 additive blending)"""
    dv, dn, ds = [_f32(a) for a in data]
    mv, mn, ms = [_f32(a) for a in model]
    if not semantic:
        ds = ms = None
    out48 = np.zeros(48, np.float32)
    lib(mode).ref_icp_jacobian(C.byref(p), _p(dv), _p(dn), _p(ds), _p(mv), _p(mn), _p(ms),
                               _p(colmajor(pose, np.float64), C.c_double), C.c_int32(iteration),
                               C.c_float(p.icp_max_distance if max_distance is None else max_distance),
                               C.c_float(p.icp_max_angle if max_angle is None else max_angle),
                               C.c_int32(entries_per_kernel), _p(out48))
    return out48


class Map:
    """SurfelMap driven through the reference's shaders (no submap paging: that is host code, not a shader)"""

    def __init__(self, p, mode="pinned"):
        self.p, self.L = p, lib(mode)
        self.h = C.c_void_p(self.L.ref_map_create(C.byref(p)))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_map_destroy(self.h); self.h = None

    def size(self):
        return self.L.ref_map_size(self.h)

    def timestamp(self):
        return self.L.ref_map_timestamp(self.h)

    def download(self):
        n = self.size()
        a = np.zeros(max(n, 1), O.SURFEL_DTYPE)
        k = self.L.ref_map_download(self.h, a.ctypes.data_as(C.c_void_p), C.c_uint32(n))
        return a[:k]

    def upload(self, surfels, timestamp):
        a = np.ascontiguousarray(surfels, O.SURFEL_DTYPE)
        self.L.ref_map_upload(self.h, a.ctypes.data_as(C.c_void_p), C.c_uint32(a.shape[0]), C.c_uint32(timestamp))

    def set_pose(self, t, pose):
        self.L.ref_map_set_pose(self.h, C.c_uint32(t), _p(colmajor(pose, np.float32)))

    def derived(self):
        a = np.zeros(6, np.float32)
        self.L.ref_map_derived(self.h, _p(a))
        return dict(zip(("pixel_size", "p_unstable", "log_prior", "log_unstable", "radconf_angle_thresh",
                         "update_angle_thresh"), a.tolist()))

    def _mframe(self):
        H, W = self.p.model_height, self.p.model_width
        return [np.zeros((H, W, 4), np.float32) for _ in range(3)]

    def render(self, pose_old, pose_new, conf_thr):
        v, n, s = self._mframe()
        self.L.ref_map_render(self.h, _p(colmajor(pose_old, np.float32)), _p(colmajor(pose_new, np.float32)),
                              C.c_float(conf_thr), _p(v), _p(n), _p(s))
        return v, n, s

    def render_active(self, pose, conf_thr):
        self.L.ref_map_render_active(self.h, _p(colmajor(pose, np.float32)), C.c_float(conf_thr))

    def render_inactive(self, pose, conf_thr):
        self.L.ref_map_render_inactive(self.h, _p(colmajor(pose, np.float32)), C.c_float(conf_thr))

    def render_composed(self, pose_old, pose_new, conf_thr):
        self.L.ref_map_render_composed(self.h, _p(colmajor(pose_old, np.float32)), _p(colmajor(pose_new, np.float32)),
                                       C.c_float(conf_thr))

    def frame(self, which):
        v, n, s = self._mframe()
        self.L.ref_map_get_frame(self.h, C.c_int(which), _p(v), _p(n), _p(s))
        return v, n, s

    def update(self, pose, frame):
        fv, fn, fs = [_f32(a) for a in frame]
        self.L.ref_map_update(self.h, _p(colmajor(pose, np.float32)), _p(fv), _p(fn), _p(fs))

    def update_debug(self):
        H, W = self.p.data_height, self.p.data_width
        idx = np.zeros((H, W), np.uint32); rad = np.zeros((H, W, 4), np.float32); integ = np.zeros((H, W), np.uint8)
        nu, nn = C.c_uint32(0), C.c_uint32(0)
        self.L.ref_map_get_update_debug(self.h, _p(idx, C.c_uint32), _p(rad), _p(integ, C.c_uint8), C.byref(nu),
                                        C.byref(nn))
        return idx, rad, integ, nu.value, nn.value
