"""CPU tests of the KITTI I/O and odometry-evaluation helpers: file formats round-trip, the pose
export uses the camera frame like the visualizer, and the devkit's segment errors give the closed-form answers on
trajectories with a known drift."""
import numpy as np
import pytest

from semantic_suma_b200 import kitti, synth


def _yaw(deg):
    a = np.deg2rad(deg)
    R = np.eye(4)
    R[0, 0], R[0, 1], R[1, 0], R[1, 1] = np.cos(a), -np.sin(a), np.sin(a), np.cos(a)
    return R


def test_scan_and_label_files_roundtrip(tmp_path):
    rng = np.random.default_rng(3)
    xyz = rng.normal(0, 20, (1234, 3)).astype(np.float32)
    rem = rng.uniform(0, 0.8, 1234).astype(np.float32)
    f = tmp_path / "000000.bin"
    kitti.write_scan(f, xyz, rem)
    assert f.stat().st_size == 1234 * 16
    pts, r = kitti.read_scan(f)
    assert pts.shape == (1234, 4) and np.array_equal(pts[:, :3], xyz) and np.all(pts[:, 3] == 1.0)
    assert np.array_equal(r, rem / rem.max()) and r.max() == 1.0      # KITTIReader.cpp:168-170
    # a truncated file loses only the partial record
    with open(f, "ab") as fh:
        fh.write(b"\x00" * 7)
    assert kitti.read_scan(f)[0].shape[0] == 1234
    # labels: class id in the low 16 bits, instance id above
    cls = rng.integers(0, 260, 1234)
    inst = rng.integers(0, 1000, 1234)
    g = tmp_path / "000000.label"
    kitti.write_labels(g, cls, inst)
    lab, prob = kitti.read_labels(g, n_points=1234, probability=0.9)
    assert lab.dtype == np.float32 and np.array_equal(lab, cls.astype(np.float32)) and np.all(prob == np.float32(0.9))
    with pytest.raises(ValueError):
        kitti.read_labels(g, n_points=10)
    # empty scan
    e = tmp_path / "empty.bin"
    e.write_bytes(b"")
    pts, r = kitti.read_scan(e)
    assert pts.shape == (0, 4) and r.shape == (0,)


def test_calibration_and_pose_files(tmp_path):
    calib = tmp_path / "calib.txt"
    Tr = np.array([[0, -1, 0, 0.1], [0, 0, -1, -0.2], [1, 0, 0, 0.3], [0, 0, 0, 1]], np.float32)  # velo -> cam axes
    calib.write_text("P0: " + " ".join(["1"] * 12) + "\nbroken line\nTr: " +
                     " ".join(repr(float(x)) for x in Tr[:3].reshape(-1)) + "\nshort: 1 2 3\n")
    c = kitti.read_calibration(calib)
    assert set(c) == {"P0", "Tr"} and np.array_equal(c["Tr"], Tr)
    poses = [p.astype(np.float32) for p in synth.trajectory(25)]
    out = tmp_path / "00.txt"
    kitti.save_poses(out, poses, Tr=c["Tr"])
    back = kitti.load_poses(out)
    assert len(back) == 25
    want = kitti.poses_to_camera_frame(poses, Tr)
    for a, b in zip(back, want):
        assert np.array_equal(a, b)              # repr(float32 value) round-trips exactly
    # forward motion along velodyne x is forward motion along camera z (VisualizerWindow.cpp:859)
    step = np.linalg.inv(want[0]) @ want[1]
    assert abs(step[2, 3] - 1.0) < 1e-3 and abs(step[0, 3]) < 2e-2
    # lines with fewer than 12 numbers are skipped
    out.write_text(out.read_text() + "1 2 3\n\n")
    assert len(kitti.load_poses(out)) == 25


def test_segment_errors_on_known_drift():
    n = 420                                       # 1 m per frame: segments of 100..400 m exist
    gt = [p.astype(np.float32) for p in synth.trajectory(n)]
    dist = kitti.trajectory_distances(gt)
    assert abs(dist[-1] - (n - 1)) < 0.05 * (n - 1) and np.all(np.diff(dist) > 0)
    first, length = 20, 100.0
    last = kitti.last_frame_from_segment_length(dist, first, length)
    assert dist[last] > dist[first] + length >= dist[last - 1]
    assert kitti.last_frame_from_segment_length(dist, n - 5, 100.0) == -1
    # identical trajectories: no error, every start with a long enough tail is evaluated
    errs = kitti.calc_sequence_errors(gt, gt)
    assert len(errs) > 0 and max(e[1] for e in errs) < 1e-5 and max(e[2] for e in errs) < 1e-5
    assert {e[3] for e in errs} == {100.0, 200.0, 300.0, 400.0} and all(e[0] % 10 == 0 for e in errs)
    # result = ground truth re-expressed in a rotated / shifted world: relative motions, hence errors, unchanged
    W = _yaw(33.0).astype(np.float32)
    W[:3, 3] = (5, -7, 1)
    moved = [(W @ p).astype(np.float32) for p in gt]
    e2 = kitti.calc_sequence_errors(gt, moved)
    assert max(e[1] for e in e2) < 1e-4 and max(e[2] for e in e2) < 1e-3
    # closed forms on a straight 1 m/frame run: a uniform 2 % scale error of the estimate gives a translational error of
    # 2 % of the displacement (= segment length on a straight line) and no rotational error
    straight = []
    for i in range(n):
        q = np.eye(4, dtype=np.float32)
        q[0, 3] = i
        straight.append(q)
    scaled = [q.copy() for q in straight]
    for q in scaled:
        q[:3, 3] *= np.float32(1.02)
    errs_s = kitti.calc_sequence_errors(straight, scaled)
    t_err, r_err = kitti.sequence_stats(errs_s)
    assert abs(t_err - 0.02) < 5e-4 and r_err < 1e-6
    ev = kitti.evaluate(straight, scaled)
    assert abs(ev["t_err_percent"] - 2.0) < 0.05 and ev["segments"] == len(errs_s)
    # an extra yaw of 0.01 deg per frame on top of the motion: 0.01 deg of rotational error per metre travelled
    drift = [(q.astype(np.float64) @ _yaw(0.01 * i)).astype(np.float32) for i, q in enumerate(straight)]
    errs_d = kitti.calc_sequence_errors(straight, drift)
    r_deg_per_m = np.degrees(np.mean([e[1] for e in errs_d]))
    assert abs(r_deg_per_m - 0.01) < 5e-4
    with pytest.raises(ValueError):
        kitti.calc_sequence_errors(gt, gt[:-1])


def test_sequence_runner_plumbing_with_an_injected_engine(tmp_path):
    """reader -> processScan -> pose export -> odometry errors over a synthetic sequence on disk. The engine injected
    here is the CPU oracle behind the SurfelMapping surface (tests may use it; the runner itself defaults to the CUDA
    library and never imports the oracle), so the whole file plumbing is exercised without a GPU."""
    from oracle import oracle as O
    from semantic_suma_b200 import run_kitti
    from helpers import sized
    seq = str(tmp_path / "seq")
    run_kitti.make_synthetic_sequence(seq, 12, width=450, semantic=True)
    assert len(run_kitti.list_scans(seq)) == 12

    class OracleEngine:
        def __init__(self, params):
            self.s = O.Slam(params)

        def processScan(self, pts, labels, probs):
            self.s.process_scan(pts, labels, probs)

        def getCurrentPose(self):
            return self.s.pose()

    po = O.default_params(**sized(450))
    seen = []
    res = run_kitti.run_sequence(seq, po, make_engine=OracleEngine, semantic=True,
                                 on_scan=lambda i, P: seen.append(i))
    assert res["scans"] == 12 and seen == list(range(12))
    ev = run_kitti.evaluate_against_ground_truth(seq, res["poses"])
    assert ev is not None and ev["end_point_error_m"] < 0.15        # 11 m driven, 2 cm range noise
    # the exported file is in the camera frame: forward motion is +z there
    out = tmp_path / "est.txt"
    from semantic_suma_b200 import kitti as K
    Tr = K.read_calibration(seq + "/calib.txt")["Tr"]
    K.save_poses(out, res["poses"], Tr=Tr)
    est = K.load_poses(out)
    assert len(est) == 12 and est[-1][2, 3] > 10.0 and abs(est[-1][1, 3]) < 0.2
    # run_sequence on an empty directory fails loudly
    with pytest.raises(FileNotFoundError):
        run_kitti.run_sequence(str(tmp_path / "nothing"), po, make_engine=OracleEngine)


@pytest.mark.gpu
def test_sequence_runner_on_the_gpu_matches_the_oracle_run(tmp_path):
    """the same sequence directory through the CUDA engine: identical poses (bit for bit) to the oracle-driven run"""
    from oracle import oracle as O
    from semantic_suma_b200 import api, run_kitti
    from helpers import sized
    seq = str(tmp_path / "seq")
    run_kitti.make_synthetic_sequence(seq, 8, width=900, semantic=True)

    class OracleEngine:
        def __init__(self, params):
            self.s = O.Slam(params)

        def processScan(self, pts, labels, probs):
            self.s.process_scan(pts, labels, probs)

        def getCurrentPose(self):
            return self.s.pose()

    ref = run_kitti.run_sequence(seq, O.default_params(**sized(900)), make_engine=OracleEngine, semantic=True)
    got = run_kitti.run_sequence(seq, api.default_params(**sized(900)), semantic=True)
    assert got["scans"] == 8
    for a, b in zip(got["poses"], ref["poses"]):
        assert np.array_equal(a, b)
    ev = run_kitti.evaluate_against_ground_truth(seq, got["poses"])
    assert ev["end_point_error_m"] < 0.15


def test_cpp_reader_and_metrics_agree_with_the_python_ones(tmp_path):
    """include/suma_b200_io.hpp (KITTIReader, KITTICalibration, KITTI::Odometry) against semantic_suma_b200/kitti.py on
    the same files: scan/label contents, pose files and the devkit errors of a drifting estimate."""
    import os
    import shutil
    import subprocess
    from semantic_suma_b200 import kitti as K, run_kitti
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "kitti_io_example")
    subprocess.check_call([gxx, "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "kitti_io_example.cpp"), "-o", exe])
    seq = str(tmp_path / "seq")
    n = 130
    # tiny scans (the content is irrelevant for the metrics), long trajectory so that 100 m segments exist
    run_kitti.make_synthetic_sequence(seq, 3, width=90, semantic=True)
    Tr = K.read_calibration(seq + "/calib.txt")["Tr"]
    gt_velo = [np.linalg.inv(synth.trajectory(1)[0]) @ p for p in synth.trajectory(n)]
    K.save_poses(seq + "/poses.txt", gt_velo, Tr=Tr)
    est_velo = []
    for i, p in enumerate(gt_velo):
        q = p @ _yaw(0.004 * i)
        q[:3, 3] *= 1.01
        est_velo.append(q)
    est_file = str(tmp_path / "est.txt")
    K.save_poses(est_file, est_velo, Tr=Tr)
    out = subprocess.check_output([exe, seq, est_file], text=True).split()
    count, n_scans, n_points = int(out[0]), int(out[1]), int(out[2])
    files = run_kitti.list_scans(seq)
    scans = [K.read_scan(f) for f in files]
    labels = [K.read_labels(f.replace("velodyne", "labels").replace(".bin", ".label"))[0] for f in files]
    assert count == n_scans == 3 and n_points == sum(s[0].shape[0] for s in scans)
    assert abs(float(out[3]) - sum(float(s[0].astype(np.float64).sum()) for s in scans)) < 1e-3
    assert float(out[4]) == sum(float(l.sum()) for l in labels)
    assert abs(float(out[5]) - sum(float(s[1].astype(np.float64).sum()) for s in scans)) < 1e-2
    assert int(out[6]) == 1 and int(float(out[10])) == scans[1][0].shape[0]
    errs = K.calc_sequence_errors(K.load_poses(seq + "/poses.txt"), K.load_poses(est_file))
    t_py, r_py = K.sequence_stats(errs)
    assert int(out[7]) == len(errs) > 0
    assert abs(float(out[8]) - t_py) < 1e-5 * max(1.0, abs(t_py)) + 1e-6
    assert abs(float(out[9]) - r_py) < 1e-5 * max(1.0, abs(r_py)) + 1e-7
    # savePoses(velodyne poses, Tr) reproduces the camera-frame file it was derived from
    again = K.load_poses(est_file + ".roundtrip")
    for a, b in zip(again, K.load_poses(est_file)):
        assert np.allclose(a, b, atol=2e-4)


@pytest.mark.gpu
def test_cpp_end_to_end_example_matches_the_python_runner(tmp_path):
    """tests/cpp/run_sequence_example.cpp (KITTIReader -> SurfelMapping::processScan -> savePoses) and
    semantic_suma_b200.run_kitti over the same directory: the same trajectory file up to float32 text round-off."""
    import os
    import subprocess
    from semantic_suma_b200 import api, kitti as K, run_kitti
    from helpers import sized
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    from semantic_suma_b200 import build as product_build
    api.lib()
    so = product_build.LIB  # libsuma_b200.so; under pytest --cusim the CPU executor's build of the same sources (conftest)
    exe = str(tmp_path / "run_sequence_example")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "run_sequence_example.cpp"), "-o", exe, so,
                           "-Wl,-rpath," + os.path.dirname(so)])
    seq = str(tmp_path / "seq")
    run_kitti.make_synthetic_sequence(seq, 6, width=900, semantic=False)
    out_cpp = str(tmp_path / "cpp.txt")
    log = subprocess.check_output([exe, seq, out_cpp, "900"], text=True)
    assert "scan 5:" in log
    res = run_kitti.run_sequence(seq, api.default_params(**sized(900)))
    Tr = K.read_calibration(seq + "/calib.txt")["Tr"]
    out_py = str(tmp_path / "py.txt")
    K.save_poses(out_py, res["poses"], Tr=Tr)
    a, b = K.load_poses(out_cpp), K.load_poses(out_py)
    assert len(a) == len(b) == 6
    for x, y in zip(a, b):
        assert np.allclose(x, y, atol=2e-4)
    assert abs(a[-1][2, 3] - 5.0) < 0.2          # five 1 m steps, forward = +z in the camera frame
