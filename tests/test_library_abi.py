"""CPU tests: the C-ABI library builds for sm_100a, loads, and exports every symbol include/suma_b200.h declares."""
import ctypes as C
import os
import re

import numpy as np

from semantic_suma_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "suma_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    L = api.lib()
    names = _declared()
    assert len(names) >= 40
    for n in names:
        assert hasattr(L, n), "libsuma_b200.so does not export %s" % n
    assert sorted(api.EXPORTED_SYMBOLS) == names


def test_params_struct_layout_matches_header():
    p = api.default_params()
    assert C.sizeof(api.Params) == 4 * 12 + 8 + 16 + 4 * 37 or C.sizeof(api.Params) % 8 == 0
    assert (p.data_width, p.data_height, p.max_iterations) == (900, 64, 33)
    assert abs(p.stopping_threshold - 1e-4) < 1e-9 and p.weighting == 1 and p.bilinear_sampling == 1
    assert p.render_after_update == 1 and p.label_offset_quirk == 1 and abs(p.submap_extent - 10.0) < 1e-6
    q = api.default_params(**{"icp-max-distance": 0.5, "max iterations": 10, "weighting": "turkey"})
    assert abs(q.icp_max_distance - 0.5) < 1e-7 and q.max_iterations == 10 and q.weighting == 2


def test_host_math_matches_oracle_bitwise():
    # sb_se3_exp / sb_ldlt_solve6 / sb_gn_step run on the host without a GPU
    from oracle import oracle as O
    L = api.lib()
    rng = np.random.default_rng(3)
    for _ in range(20):
        x = rng.normal(0, 0.2, 6)
        T = np.zeros(16)
        L.sb_se3_exp(x.ctypes.data_as(C.POINTER(C.c_double)), T.ctypes.data_as(C.POINTER(C.c_double)))
        assert np.array_equal(api.from_colmajor(T), O.se3_exp(x))
        A = rng.normal(size=(6, 6)); A = A @ A.T + np.eye(6); b = rng.normal(size=6)
        Ac = np.ascontiguousarray(A.T).reshape(36); xs = np.zeros(6)
        L.sb_ldlt_solve6(Ac.ctypes.data_as(C.POINTER(C.c_double)), b.ctypes.data_as(C.POINTER(C.c_double)),
                         xs.ctypes.data_as(C.POINTER(C.c_double)))
        assert np.array_equal(xs, O.ldlt_solve6(A, b))
    raw = rng.integers(-2**40, 2**40, 32).astype(np.int64)
    o = np.zeros(48)
    L.sb_icp_unpack(raw.ctypes.data_as(C.POINTER(C.c_int64)), o.ctypes.data_as(C.POINTER(C.c_double)))
    assert np.array_equal(o, O.icp_unpack(raw))


def test_no_gpu_means_loud_failure_not_fallback():
    if api.lib().sb_device_count() > 0:
        return
    try:
        api.Context(api.default_params())
    except api.SumaError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("sb_create must fail without a CUDA device")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "semantic_suma_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.replace("CPU oracle", "").replace("the oracle", "").lower() or f == "build.py", f
                assert "cusim" not in txt.lower(), f  # nor the CPU executor of tests/cusim: only tests/conftest.py swaps it in


def test_cpp_mirror_has_the_reference_class_surface():
    """include/suma_b200.hpp must offer every hot-path method of the reference's four classes:
    tests/cpp/mirror_surface.cpp names them all; it only has to compile (and link against the library)."""
    import shutil
    import subprocess
    import tempfile
    from semantic_suma_b200 import build
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    so = build.build()
    with tempfile.TemporaryDirectory() as td:
        obj = os.path.join(td, "mirror_surface.o")
        subprocess.check_call([gxx, "-std=c++17", "-fPIC", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-c",
                               os.path.join(ROOT, "tests", "cpp", "mirror_surface.cpp"), "-o", obj])
        # link into a shared object against the library: every sb_* the header uses must resolve
        subprocess.check_call([gxx, "-shared", "-o", os.path.join(td, "libsurface.so"), obj, "-L" + os.path.dirname(so),
                               "-lsuma_b200", "-Wl,--no-undefined"])
        # the end-to-end example (reader -> processScan -> pose export) builds into an executable; without a GPU it must
        # fail loudly (exit code 1 from the std::runtime_error), never fall back to anything
        exe = os.path.join(td, "run_sequence_example")
        subprocess.check_call([gxx, "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                               os.path.join(ROOT, "tests", "cpp", "run_sequence_example.cpp"), "-o", exe,
                               "-L" + os.path.dirname(so), "-lsuma_b200", "-Wl,-rpath," + os.path.dirname(so)])
        if api.lib().sb_device_count() == 0:
            os.makedirs(os.path.join(td, "seq", "velodyne"))
            r = subprocess.run([exe, os.path.join(td, "seq"), os.path.join(td, "out.txt")], capture_output=True, text=True)
            assert r.returncode == 1 and "sb_create failed" in r.stderr


def test_default_parameters_equal_the_reference_config():
    """tests/golden/reference_default_xml.json is the reference's config/default.xml (made by
    tests/golden/make_reference_params_fixture.py). Every XML key the C++ mirror's ParameterList maps onto sb_params
    must default -- in the library AND in the oracle -- to the reference's value."""
    import json
    from oracle import oracle as O
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_default_xml.json")))["params"]
    hpp = open(os.path.join(ROOT, "include", "suma_b200.hpp")).read()
    keys = dict(re.findall(r'SUMA_F\("([^"]+)",\s*([a-z_0-9]+),', hpp))
    assert len(keys) >= 45
    lib_p, orc_p = api.default_params(), O.default_params()
    not_in_xml = sorted(k for k in keys if k not in fx)
    # code-side defaults (SurfelMap.cpp:265-273) and the two switches this implementation adds
    assert not_in_xml == ["active_timestamps", "label_offset_quirk", "max_weight", "render_after_update"]
    for key, field in keys.items():
        if key not in fx:
            continue
        want = fx[key]["value"]
        got, got_o = getattr(lib_p, field), getattr(orc_p, field)
        if fx[key]["type"] == "float" and isinstance(got, float):
            assert got in (want, float(np.float32(want))), (key, want, got)
        else:
            assert got == int(want), (key, want, got)
        assert got == got_o, (key, got, got_o)
    weights = {"none": 0, "huber": 1, "turkey": 2, "stability": 3}   # Frame2Model.cpp:69-80
    assert lib_p.weighting == orc_p.weighting == weights[fx["weighting"]["value"]]


def test_surfel_record_is_the_reference_struct():
    """core/Surfel.h of the reference (fixture: tests/golden/reference_surfel_layout.json) fixes the 64-byte record of
    getAllSurfels(): sb_surfel in the header, the numpy dtype of the mirror and the oracle's dtype must have the same
    members, types and order."""
    import json
    from oracle import oracle as O
    fields = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_surfel_layout.json")))["fields"]
    assert len(fields) == 16
    src = open(os.path.join(ROOT, "include", "suma_b200.h")).read()
    body = re.search(r"typedef struct sb_surfel\s*\{(.*?)\}\s*sb_surfel;", src, flags=re.S).group(1)
    mine = [[n.strip(), t] for t, names in re.findall(r"\b(float|uint32_t|int32_t)\s+([^;]+);", body)
            for n in names.split(",")]
    assert mine == fields
    np_type = {"float": np.dtype(np.float32), "uint32_t": np.dtype(np.uint32)}
    for dt in (api.SURFEL_DTYPE, O.SURFEL_DTYPE):
        assert dt.itemsize == 64
        assert [[n, dt.fields[n][0]] for n in dt.names] == [[n, np_type[t]] for n, t in fields]
        assert [dt.fields[n][1] for n in dt.names] == [4 * i for i in range(16)]
