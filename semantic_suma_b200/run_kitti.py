"""Run the hot path over a KITTI-odometry-style sequence directory and write / evaluate the trajectory
(the callers and data formats on either side of `SurfelMapping::processScan`).

    <seq>/velodyne/000000.bin ...      scans (N x 4 float32)                     io/KITTIReader.cpp:40-62, 136-170
    <seq>/labels/000000.label ...      optional per-point class ids (stand-in for the RangeNet++ call, :172-203)
    <seq>/calib.txt                    optional; "Tr" moves the exported poses to the camera frame
    <seq>/poses.txt                    optional ground truth (camera frame if calib.txt has Tr) -> odometry errors

    python -m semantic_suma_b200.run_kitti <seq> [--out poses_out.txt] [--semantic] [--max-scans N] [--width 2048]
    python -m semantic_suma_b200.run_kitti --make-synthetic <dir> --scans 50     # writes such a directory

The engine is the CUDA library (a B200 is required); tests inject another engine through `make_engine` to check the
file plumbing without a GPU.
"""
import argparse
import glob
import json
import os
import sys
import time

import numpy as np

from . import kitti


def list_scans(seq_dir):
    return sorted(glob.glob(os.path.join(seq_dir, "velodyne", "*.bin")))


def _cuda_engine(params):
    from . import api
    return api.SurfelMapping(params)


def run_sequence(seq_dir, params, make_engine=None, semantic=False, max_scans=None, on_scan=None):
    """-> dict(poses=[4x4 float64 in the velodyne frame], seconds, scans). `make_engine(params)` must return an object
    with processScan(points4, labels, probs) and getCurrentPose() -- the SurfelMapping surface."""
    files = list_scans(seq_dir)
    if max_scans is not None:
        files = files[:max_scans]
    if not files:
        raise FileNotFoundError("no scans under %s/velodyne" % seq_dir)
    engine = (make_engine or _cuda_engine)(params)
    poses = []
    t0 = time.time()
    for i, f in enumerate(files):
        pts, _ = kitti.read_scan(f)
        labels = probs = None
        if semantic:
            lf = os.path.join(seq_dir, "labels", os.path.splitext(os.path.basename(f))[0] + ".label")
            labels, probs = kitti.read_labels(lf, n_points=pts.shape[0])
        engine.processScan(pts, labels, probs)
        poses.append(np.asarray(engine.getCurrentPose(), np.float64).copy())
        if on_scan:
            on_scan(i, poses[-1])
    return {"poses": poses, "seconds": time.time() - t0, "scans": len(files)}


def evaluate_against_ground_truth(seq_dir, poses_velo):
    """odometry errors of `poses_velo` against <seq>/poses.txt, in the frame that file uses (camera if calib has Tr)"""
    gt_file = os.path.join(seq_dir, "poses.txt")
    if not os.path.exists(gt_file):
        return None
    gt = kitti.load_poses(gt_file)[:len(poses_velo)]
    calib_file = os.path.join(seq_dir, "calib.txt")
    Tr = kitti.read_calibration(calib_file).get("Tr") if os.path.exists(calib_file) else None
    est = [np.asarray(p, np.float32) for p in poses_velo]
    if Tr is not None:
        est = kitti.poses_to_camera_frame(est, Tr)
    # both trajectories start at identity in the devkit's convention
    g0, e0 = np.linalg.inv(gt[0]), np.linalg.inv(est[0])
    gt = [(g0 @ g).astype(np.float32) for g in gt]
    est = [(e0 @ e).astype(np.float32) for e in est]
    out = kitti.evaluate(gt, est[:len(gt)])
    out["end_point_error_m"] = float(np.linalg.norm(gt[-1][:3, 3] - est[len(gt) - 1][:3, 3]))
    return out


def make_synthetic_sequence(seq_dir, n_scans, width=2048, height=64, semantic=False, seed=1337):
    """writes a synthetic sequence (semantic_suma_b200/synth.py scene) in the KITTI layout, with calib.txt (the usual
    velodyne -> camera axis change) and ground-truth poses in the camera frame"""
    from . import synth
    os.makedirs(os.path.join(seq_dir, "velodyne"), exist_ok=True)
    if semantic:
        os.makedirs(os.path.join(seq_dir, "labels"), exist_ok=True)
    Tr = np.array([[0, -1, 0, 0], [0, 0, -1, 0], [1, 0, 0, 0], [0, 0, 0, 1]], np.float32)
    with open(os.path.join(seq_dir, "calib.txt"), "w") as f:
        f.write("Tr: " + " ".join(repr(float(x)) for x in Tr[:3].reshape(-1)) + "\n")
    sc = synth.Scene(width=width, height=height, seed=seed, semantic=semantic)
    poses = synth.trajectory(n_scans)
    for i in range(n_scans):
        pts, labels, _ = sc.scan(i, poses[i])
        kitti.write_scan(os.path.join(seq_dir, "velodyne", "%06d.bin" % i), pts[:, :3], np.full(pts.shape[0], 0.5))
        if semantic:
            kitti.write_labels(os.path.join(seq_dir, "labels", "%06d.label" % i), labels.astype(np.uint32))
    rel = [np.linalg.inv(poses[0]) @ p for p in poses]          # start at identity
    kitti.save_poses(os.path.join(seq_dir, "poses.txt"), rel, Tr=Tr)
    return seq_dir


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("seq", nargs="?")
    ap.add_argument("--out", default=None, help="pose file to write (KITTI format; camera frame if calib.txt has Tr)")
    ap.add_argument("--semantic", action="store_true")
    ap.add_argument("--max-scans", type=int, default=None)
    ap.add_argument("--width", type=int, default=2048)
    ap.add_argument("--make-synthetic", default=None, metavar="DIR")
    ap.add_argument("--scans", type=int, default=50)
    args = ap.parse_args(argv)
    if args.make_synthetic:
        make_synthetic_sequence(args.make_synthetic, args.scans, width=args.width, semantic=args.semantic)
        print("wrote", args.make_synthetic)
        if not args.seq:
            return 0
    if not args.seq:
        ap.error("need a sequence directory")
    from . import api
    params = api.default_params(data_width=args.width, model_width=args.width)
    res = run_sequence(args.seq, params, semantic=args.semantic, max_scans=args.max_scans)
    calib_file = os.path.join(args.seq, "calib.txt")
    Tr = kitti.read_calibration(calib_file).get("Tr") if os.path.exists(calib_file) else None
    if args.out:
        kitti.save_poses(args.out, res["poses"], Tr=Tr)
    report = {"scans": res["scans"], "seconds": round(res["seconds"], 3),
              "scans_per_sec_incl_file_io": round(res["scans"] / res["seconds"], 2),
              "odometry": evaluate_against_ground_truth(args.seq, res["poses"])}
    print(json.dumps(report))
    return 0


if __name__ == "__main__":
    sys.exit(main())
