"""Row-striped Jacobian reduction over two ranks (BASELINE.json configs[3]): every rank reduces its rows,
the 32 int64 sums are all-reduced, and every rank must end with exactly the pose a single GPU computes.

Two exchanges are covered: the host-callback baseline (torch.distributed all-reduce once per iteration; on a one-GPU
box both ranks share the GPU and use gloo, as there is no in-kernel waiting) and the fused one, where the last block of
the persistent Gauss-Newton kernel exchanges the sums over CUDA-IPC peer memory (needs two GPUs: the kernels of the two
ranks must run at the same time)."""
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


def _worker(rank, world, port, q, two_gpus, fused, size, n_scans, iters, jump=False):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import torch
        import torch.distributed as dist
        from semantic_suma_b200 import api, stripes
        from helpers import scans, sized
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dev = rank if two_gpus else 0
        torch.cuda.set_device(dev)
        if two_gpus:
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
        else:
            dist.init_process_group("gloo", rank=rank, world_size=world)
        pp = api.default_params(**sized(size), max_iterations=iters, stopping_threshold=0.0, delta=0.0)
        sc, _ = scans(size, n=n_scans)
        if jump:  # a pose jump after scan 3: the frame-to-frame recovery minimisation runs striped as well
            from semantic_suma_b200 import synth
            scene = synth.Scene(width=size, height=64)
            tr = synth.trajectory(n_scans + 1)
            J = synth.translate(0.8, 0.3, 0) @ synth.rot_z(np.deg2rad(8.0))
            sc = [scene.scan(f, tr[f] if f < 4 else tr[f] @ J) for f in range(n_scans)]
        solo = api.SurfelMapping(pp, device=dev)
        for s in sc:
            solo.processScan(*s)
        ref_pose = solo.getCurrentPose().copy()
        ref_n = solo.getMap().size()
        assert not jump or solo.getStatistics()["track_loss"] >= 1
        solo.ctx.close()
        sl = api.SurfelMapping(pp, device=dev)
        r0, r1 = stripes.setup_comm(sl.ctx, dist, fused=fused)
        for s in sc:
            sl.processScan(*s)
        pose = sl.getCurrentPose()
        ok = bool(np.array_equal(pose, ref_pose)) and sl.getMap().size() == ref_n
        q.put((rank, ok, (r0, r1), pose.tobytes()))
        dist.barrier()
        sl.ctx.close()
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        q.put((rank, False, repr(e), b""))
        raise


def _run_two_ranks(fused, two_gpus, size=900, n_scans=4, iters=8, jump=False):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, two_gpus, fused, size, n_scans, iters, jump)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        got = sorted(q.get(timeout=240) for _ in range(2))
        for p in procs:
            p.join(timeout=60)
    finally:
        for p in procs:  # never leave a rank behind on the GPU
            if p.is_alive():
                p.terminate()
                p.join(timeout=10)
                if p.is_alive():
                    p.kill()
    assert got[0][1] and got[1][1], "striped result differs from the single-GPU result: %r %r" % (got[0][2], got[1][2])
    assert got[0][3] == got[1][3]
    assert got[0][2] == (0, 32) and got[1][2] == (32, 64)


@pytest.mark.gpu
def test_two_rank_striped_icp_is_bit_identical():
    from semantic_suma_b200 import api
    _run_two_ranks(fused=False, two_gpus=api.lib().sb_device_count() >= 2)


@pytest.mark.gpu
@pytest.mark.parametrize("size,n_scans,iters,jump", [(900, 4, 8, False), (2048, 6, 10, False), (900, 7, 33, True)])
def test_two_gpu_fused_peer_allreduce_is_bit_identical(size, n_scans, iters, jump):
    from semantic_suma_b200 import api
    if api.lib().sb_device_count() < 2:
        pytest.skip("the in-kernel peer exchange needs two GPUs")
    _run_two_ranks(fused=True, two_gpus=True, size=size, n_scans=n_scans, iters=iters, jump=jump)
