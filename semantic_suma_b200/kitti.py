"""KITTI odometry I/O and evaluation around the hot path: the data formats on either side of
`SurfelMapping::processScan`.

* velodyne `.bin` scans: N x 4 float32 (x, y, z, remission)                 io/KITTIReader.cpp:136-170
* SemanticKITTI `.label` files (uint32, class id in the low 16 bits) as the stand-in for the labels the reference
  obtains from RangeNet++ at read time (io/KITTIReader.cpp:172-203; the network is outside the hot path)
* `calib.txt` ("name: 12 floats" -> 4x4)                                    util/kitti_utils.cpp:32-61
* pose files, 12 floats per line (row-major 3x4), written in the camera frame T_cam_velo * P * T_velo_cam
                                                                            util/kitti_utils.cpp:78-108, VisualizerWindow.cpp:848-868
* the odometry devkit's segment errors (lengths 100..800 m, a start every 10 frames)
                                                                            util/kitti_utils.cpp:111-191, 479-498

Plain numpy on the host; nothing here touches the GPU or the oracle.
"""
import numpy as np

SEGMENT_LENGTHS = (100.0, 200.0, 300.0, 400.0, 500.0, 600.0, 700.0, 800.0)  # util/kitti_utils.cpp:75
STEP_SIZE = 10                                                            # util/kitti_utils.cpp:155 ("every second")


# ------------------------------------------------------------------------------------------------ scans and labels
def read_scan(path):
    """-> (points4 [N,4] float32 with w = 1, remissions [N] float32 scaled to max 1) like KITTIReader::read.
    A trailing partial record is dropped (the reference sizes the read as floor(bytes / 16))."""
    raw = np.fromfile(path, dtype=np.float32)
    n = raw.size // 4
    v = raw[:4 * n].reshape(n, 4)
    pts = np.ones((n, 4), np.float32)
    pts[:, :3] = v[:, :3]
    rem = v[:, 3].copy()
    mx = float(rem.max()) if n else 0.0
    if mx > 0.0:
        rem /= np.float32(mx)
    return pts, rem


def write_scan(path, xyz, remissions=None):
    xyz = np.asarray(xyz, np.float32)
    out = np.zeros((xyz.shape[0], 4), np.float32)
    out[:, :3] = xyz[:, :3]
    if remissions is not None:
        out[:, 3] = np.asarray(remissions, np.float32)
    out.tofile(path)


def read_labels(path, n_points=None, probability=1.0):
    """SemanticKITTI label file -> (labels float32, probabilities float32): what `Laserscan::labels_float` /
    `labels_prob` hold after the reader ran the network (class id per point, confidence of that class)."""
    raw = np.fromfile(path, dtype=np.uint32)
    if n_points is not None and raw.size != n_points:
        raise ValueError("label file %s has %d entries for %d points" % (path, raw.size, n_points))
    labels = (raw & np.uint32(0xFFFF)).astype(np.float32)
    return labels, np.full(raw.size, probability, np.float32)


def write_labels(path, class_ids, instance_ids=None):
    sem = np.asarray(class_ids).astype(np.uint32) & np.uint32(0xFFFF)
    inst = np.zeros_like(sem) if instance_ids is None else (np.asarray(instance_ids).astype(np.uint32) << np.uint32(16))
    (sem | inst).astype(np.uint32).tofile(path)


# ------------------------------------------------------------------------------------------------ calibration, poses
def read_calibration(path):
    """{name: 4x4 float32}; lines that are not 'name: 12 numbers' are skipped, as in KITTICalibration::initialize."""
    out = {}
    with open(path) as f:
        for line in f:
            parts = line.split(":")
            if len(parts) != 2:
                continue
            vals = parts[1].split()
            if len(vals) != 12:
                continue
            m = np.eye(4, dtype=np.float32)
            m[:3, :4] = np.asarray([float(v) for v in vals], np.float32).reshape(3, 4)
            out[parts[0].strip()] = m
    return out


def load_poses(path):
    """list of 4x4 float32; lines with fewer than 12 entries are skipped (loadPoses)."""
    poses = []
    with open(path) as f:
        for line in f:
            vals = line.split()
            if len(vals) < 12:
                continue
            m = np.eye(4, dtype=np.float32)
            m[:3, :4] = np.asarray([float(v) for v in vals[:12]], np.float32).reshape(3, 4)
            poses.append(m)
    return poses


def poses_to_camera_frame(poses_velo, Tr):
    """VisualizerWindow.cpp:852-860: pose_cam = Tr * pose_velo(float) * Tr^-1."""
    Tr = np.asarray(Tr, np.float32)
    Tr_inv = np.linalg.inv(Tr).astype(np.float32)
    return [(Tr @ np.asarray(P, np.float32) @ Tr_inv).astype(np.float32) for P in poses_velo]


def save_poses(path, poses, Tr=None):
    """one line per pose: the 3x4 row-major block, space separated; with `Tr` the poses are first moved to the camera
    frame like the visualizer's "save poses" does"""
    if Tr is not None:
        poses = poses_to_camera_frame(poses, Tr)
    with open(path, "w") as f:
        for P in poses:
            f.write(" ".join(repr(float(x)) for x in np.asarray(P, np.float32)[:3, :4].reshape(-1)) + "\n")


# ------------------------------------------------------------------------------------------------ odometry metrics
def trajectory_distances(poses):
    """cumulative path length at every frame (float32 accumulation like the devkit)"""
    t = np.asarray([np.asarray(P, np.float32)[:3, 3] for P in poses], np.float32).reshape(-1, 3)
    dist = np.zeros(len(poses), np.float32)
    for i in range(1, len(poses)):
        d = t[i - 1] - t[i]
        dist[i] = dist[i - 1] + np.float32(np.sqrt(np.float32(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])))
    return dist


def last_frame_from_segment_length(dist, first_frame, length):
    for i in range(first_frame, len(dist)):
        if dist[i] > dist[first_frame] + np.float32(length):
            return i
    return -1


def rotation_error(pose_error):
    E = np.asarray(pose_error, np.float32)
    d = np.float32(0.5) * (E[0, 0] + E[1, 1] + E[2, 2] - np.float32(1.0))
    return float(np.arccos(np.clip(d, np.float32(-1.0), np.float32(1.0))))


def translation_error(pose_error):
    E = np.asarray(pose_error, np.float32)
    return float(np.sqrt(E[0, 3] * E[0, 3] + E[1, 3] * E[1, 3] + E[2, 3] * E[2, 3]))


def calc_sequence_errors(poses_gt, poses_result):
    """[(first_frame, r_err per metre [rad/m], t_err per metre [-], length, speed)] for every start frame (every 10th)
    and every segment length the ground-truth trajectory is long enough for"""
    if len(poses_gt) != len(poses_result):
        raise ValueError("need as many result poses as ground-truth poses")
    gt = [np.asarray(P, np.float32) for P in poses_gt]
    res = [np.asarray(P, np.float32) for P in poses_result]
    dist = trajectory_distances(gt)
    err = []
    for first in range(0, len(gt), STEP_SIZE):
        for length in SEGMENT_LENGTHS:
            last = last_frame_from_segment_length(dist, first, length)
            if last == -1:
                continue
            delta_gt = np.linalg.inv(gt[first]) @ gt[last]
            delta_res = np.linalg.inv(res[first]) @ res[last]
            pose_error = np.linalg.inv(delta_res) @ delta_gt
            num_frames = float(last - first + 1)
            speed = length / (0.1 * num_frames)
            err.append((first, rotation_error(pose_error) / length, translation_error(pose_error) / length, length, speed))
    return err


def sequence_stats(errors):
    """(mean translational error [fraction], mean rotational error [rad/m]): the two numbers of the devkit's stats.txt"""
    if not errors:
        return float("nan"), float("nan")
    t = float(np.mean([e[2] for e in errors]))
    r = float(np.mean([e[1] for e in errors]))
    return t, r


def evaluate(poses_gt, poses_result):
    """convenience: {'t_err_percent', 'r_err_deg_per_m', 'segments'}"""
    errs = calc_sequence_errors(poses_gt, poses_result)
    t, r = sequence_stats(errs)
    return {"t_err_percent": 100.0 * t, "r_err_deg_per_m": float(np.degrees(r)), "segments": len(errs)}
