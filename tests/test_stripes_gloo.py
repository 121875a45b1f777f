"""CPU test of the N>1 path: two gloo ranks each reduce their row stripe with the oracle; the exact int64 all-reduce
must reproduce the single-process sums bit for bit, and the Gauss-Newton step computed from them must be identical on
both ranks."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

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
    import torch.distributed as dist
    from oracle import oracle as O
    from semantic_suma_b200 import stripes
    from helpers import scans, sized
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = O.default_params(**sized(900))
    sc, poses = scans(900, n=2)
    f0, f1 = O.preprocess(p, *sc[0]), O.preprocess(p, *sc[1])
    pose = np.eye(4)
    results = []
    for it in range(3):
        r0, r1 = stripes.row_stripe(rank, world, 64)
        _, raw = O.icp_jacobian(p, f1, f0, pose, iteration=it, rows=(r0, r1))
        tot = stripes.allreduce_raw32(raw, dist)
        o48 = O.icp_unpack(tot)
        _, full = O.icp_jacobian(p, f1, f0, pose, iteration=it)
        assert np.array_equal(tot, full), "striped sums differ from the single pass"
        import ctypes as C
        pc = O.colmajor(pose, np.float64); dx = np.zeros(6)
        O.lib().orc_gn_step(O._p(o48, C.c_double), C.c_double(1e30), C.c_double(0.0), C.c_double(0.0),
                            O._p(pc, C.c_double), O._p(dx, C.c_double))
        pose = O.from_colmajor(pc)
        results.append(pose.copy())
    q.put((rank, np.stack(results).tobytes()))
    dist.barrier()
    dist.destroy_process_group()


def test_row_stripes_cover_and_balance():
    sys.path.insert(0, ROOT)
    from semantic_suma_b200 import stripes
    for H in (64, 128, 7):
        for world in (1, 2, 3, 4, 8):
            rows = [stripes.row_stripe(r, world, H) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == H
            assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
            sizes = [b - a for a, b in rows]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_exact_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] == got[1], "ranks disagree on the pose after the exact all-reduce"
