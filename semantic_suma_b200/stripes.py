"""Row-stripe sharding of the K5 Jacobian reduction across GPUs (BASELINE.json configs[3]).

Every data pixel is independent; rank r owns rows [r*H/G, (r+1)*H/G) of the data image, the model maps are replicated.
The 32 fixed-point sums are integers, so the all-reduce is exact and order independent: any number of ranks gives the
bits of a single pass. On GPUs the exchange is fused into the Jacobian kernel's last block (peer-mapped mailboxes over
NVLink, sb_comm_init); this module is the host-side plumbing: the partition rule and the handle exchange.
"""
import ctypes as C

import numpy as np


def row_stripe(rank, world, height):
    """rows [begin, end) of rank `rank`; contiguous, covering, balanced to within one row"""
    base, rem = divmod(height, world)
    begin = rank * base + min(rank, rem)
    end = begin + base + (1 if rank < rem else 0)
    return begin, end


def allreduce_raw32(raw32, dist):
    """exact all-reduce of the 32 int64 sums through torch.distributed (gloo on CPU, NCCL on GPU tensors)"""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(raw32, np.int64).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy()


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_int64))


def setup_comm(ctx, dist, height=None, fused=None):
    """Row-stripe the K5 reduction of `ctx` over the ranks of `dist`.

    fused=True (default with the nccl backend, i.e. one GPU per rank): the exchange runs inside the persistent
    Gauss-Newton kernel -- its last block stores the 32 int64 sums into every peer's CUDA-IPC mailbox over NVLink and
    spins on the peers' stamps (sb_comm_init); no host round trip per iteration.
    fused=False (SUMA_B200_FUSED_COMM=0, and the only choice when ranks share a GPU): baseline -- the library calls back
    once per iteration and the sums are all-reduced with torch.distributed. Both are exact (integer sums).
    """
    import os
    import torch
    from . import api
    if fused is None:
        fused = os.environ.get("SUMA_B200_FUSED_COMM", "1" if dist.get_backend() == "nccl" else "0") == "1"
    rank, world = dist.get_rank(), dist.get_world_size()
    H = ctx.params.data_height if height is None else height
    if not fused:
        r0, r1 = row_stripe(rank, world, H)
        on_gpu = dist.get_backend() == "nccl"

        def _cb(user, ptr):
            try:
                a = np.ctypeslib.as_array(ptr, shape=(32,))
                t = torch.from_numpy(a.copy())
                if on_gpu:
                    t = t.cuda()
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                a[:] = t.cpu().numpy()
                return 0
            except Exception:  # noqa: BLE001
                return 1

        ctx._allreduce_cb = ALLREDUCE_FN(_cb)  # keep the thunk alive as long as the context
        ctx.check(api.lib().sb_comm_set_callback(ctx.h, C.cast(ctx._allreduce_cb, C.c_void_p), None, r0, r1),
                  "comm_set_callback")
        dist.barrier()
        return r0, r1
    h = np.zeros(64, np.uint8)
    ctx.check(api.lib().sb_comm_export(ctx.h, C.c_void_p(h.ctypes.data)), "comm_export")
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    mine = torch.from_numpy(h).to(dev)
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    allh = np.ascontiguousarray(np.concatenate([g.cpu().numpy() for g in gathered]))
    r0, r1 = row_stripe(rank, world, H)
    ctx.check(api.lib().sb_comm_init(ctx.h, rank, world, C.c_void_p(allh.ctypes.data), r0, r1), "comm_init")
    dist.barrier()
    return r0, r1
