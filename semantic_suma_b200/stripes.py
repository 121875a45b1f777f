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


def setup_comm(ctx, dist, height=None):
    """exchange the IPC handles of the ranks' mailboxes and enable the fused all-reduce on `ctx`"""
    import torch
    from . import api
    rank, world = dist.get_rank(), dist.get_world_size()
    h = np.zeros(64, np.uint8)
    ctx.check(api.lib().sb_comm_export(ctx.h, C.c_void_p(h.ctypes.data)), "comm_export")
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    mine = torch.from_numpy(h).to(dev)
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    allh = np.ascontiguousarray(np.concatenate([g.cpu().numpy() for g in gathered]))
    H = ctx.params.data_height if height is None else height
    r0, r1 = row_stripe(rank, world, H)
    ctx.check(api.lib().sb_comm_init(ctx.h, rank, world, C.c_void_p(allh.ctypes.data), r0, r1), "comm_init")
    dist.barrier()
    return r0, r1
