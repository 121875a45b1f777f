"""2-GPU test of the row-striped Jacobian reduction: the one-shot peer-memory all-reduce fused into the GN kernel must
give every rank exactly the pose a single GPU computes (integer sums -> bit-identical)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from semantic_suma_b200 import api, stripes
    from helpers import scans, sized
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    pp = api.default_params(**sized(900), max_iterations=8, stopping_threshold=0.0, delta=0.0)
    sc, _ = scans(900, n=4)
    # single-GPU result on this rank's device
    solo = api.SurfelMapping(pp, device=rank)
    for s in sc:
        solo.processScan(*s)
    ref_pose = solo.getCurrentPose().copy()
    ref_n = solo.getMap().size()
    solo.ctx.close()
    # striped over both GPUs
    sl = api.SurfelMapping(pp, device=rank)
    r0, r1 = stripes.setup_comm(sl.ctx, dist)
    for s in sc:
        sl.processScan(*s)
    pose = sl.getCurrentPose()
    ok = bool(np.array_equal(pose, ref_pose)) and sl.getMap().size() == ref_n
    q.put((rank, ok, (r0, r1), pose.tobytes()))
    dist.barrier()
    sl.ctx.close()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpu_striped_icp_is_bit_identical():
    from semantic_suma_b200 import api
    if api.lib().sb_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    got.sort()
    assert got[0][1] and got[1][1], "striped result differs from the single-GPU result"
    assert got[0][3] == got[1][3]
    assert got[0][2] == (0, 32) and got[1][2] == (32, 64)
