"""TEST INFRASTRUCTURE: bench.py's native arm at N > 1 as a dry run in ONE process: N threads are the N ranks (the executor gives
every host thread its own "GPU"), torch.distributed is replaced by a thread rendezvous, torch's CUDA calls are stubbed as in
bench_dryrun.py. Exercises what no single-rank run reaches: the per-rank gathering, the `striped` record (128x4096, 15
iterations, row-striped Gauss-Newton with the in-kernel peer exchange against the un-striped run, poses compared across ranks
and modes) and its guard. Numbers mean nothing; the line and `poses_bit_identical_across_ranks_and_modes` do.

usage: python tests/cusim/bench_dryrun_multi.py N [bench.py arguments]"""
import contextlib
import io
import json
import os
import sys
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
N = int(sys.argv[1])
ARGS = sys.argv[2:] or ["--steps", "3", "--warmup", "3", "--preroll", "2", "--striped-deadline", "3000"]  # one process: a guard that fires would end every rank
sys.argv = sys.argv[:1]
os.environ["CUSIM_DEVICES"] = str(N)
import bench_dryrun  # noqa: E402,F401  (installs the executor build and the torch.cuda stubs)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402

tls = threading.local()
bar = threading.Barrier(N)
slots = [None] * N
lock = threading.Lock()


def _exchange(obj):
    slots[tls.rank] = obj
    bar.wait()
    got = list(slots)
    bar.wait()
    return got


class _Op:
    MAX, SUM = "max", "sum"


def _all_reduce(t, op=_Op.SUM):
    got = _exchange(t.clone())
    r = got[0].clone()
    for g in got[1:]:
        r = torch.maximum(r, g) if op == _Op.MAX else r + g
    t.copy_(r)


def _all_gather_object(out, obj):
    out[:] = _exchange(obj)


def _all_gather(out, t):
    for o, g in zip(out, _exchange(t.clone())):
        o.copy_(g)


dist.init_process_group = lambda *a, **k: None
dist.destroy_process_group = lambda *a, **k: None
dist.barrier = lambda *a, **k: bar.wait()
dist.all_reduce = _all_reduce
dist.all_gather_object = _all_gather_object
dist.all_gather = _all_gather
dist.get_rank = lambda *a, **k: tls.rank
dist.get_world_size = lambda *a, **k: N
dist.get_backend = lambda *a, **k: "threads"
dist.ReduceOp = _Op
torch.device = lambda *a, **k: "cpu"

_gen, _cache = bench.generate_scans, {}


def _generate(w, n_frames, seed):  # no fork from a threaded process; ranks share what they can
    with lock:
        key = (w["width"], w["height"], n_frames, seed)
        if key not in _cache:
            poses = bench.synth.trajectory(n_frames)
            _cache[key] = [bench._gen_one((w, seed, f, poses[f])) for f in range(n_frames)]
        return _cache[key]


bench.generate_scans = _generate
results, errors = [None] * N, []


def rank_main(r):
    tls.rank = r
    try:
        ap_args = bench_args()
        results[r] = bench.run_native(ap_args, bench.WORKLOADS[ap_args.workload], r, N, r)
    except BaseException as e:  # noqa: BLE001
        import traceback
        errors.append("rank %d: %s" % (r, traceback.format_exc()))
        try:
            bar.abort()
        except Exception:  # noqa: BLE001
            pass


def bench_args():
    # bench.main() without the launch: parse the same arguments
    import argparse
    real = argparse.ArgumentParser.parse_args
    holder = {}

    def grab(self, *a, **k):
        holder["ns"] = real(self, ARGS)
        raise SystemExit(0)
    argparse.ArgumentParser.parse_args = grab
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            try:
                bench.main()
            except SystemExit:
                pass
    finally:
        argparse.ArgumentParser.parse_args = real
    ns = holder["ns"]
    if ns.warmup < 3:
        ns.warmup = 3
    return ns


def main():
    with lock:
        pass
    th = [threading.Thread(target=rank_main, args=(r,)) for r in range(N)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, "\n".join(errors)
    out = results[0]
    json.dumps(out)
    assert out["n_gpus"] == N and len(out["config"]["per_rank"]) == N
    s = out.get("striped")
    assert s and "error" not in s, s
    assert s["poses_bit_identical_across_ranks_and_modes"] is True, s
    assert s["striped"]["rows"] != s["solo"]["rows"]
    print("bench dry run ok at N = %d: line complete, striped record %s" % (N, json.dumps(s)[:400]))


if __name__ == "__main__":
    main()
