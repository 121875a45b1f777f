"""TEST INFRASTRUCTURE: bench.py's native arm (N = 1), end to end, with the library's CUDA sources on the CPU executor and
torch's CUDA calls stubbed (tensors stay in host memory; events read the wall clock) -- a dry run of the benchmark's own
Python path (pre-roll, device-resident pass, end-to-end pass with input staging, per-kernel pass, CPU baseline, the JSON
line) in a container without a GPU. The numbers it prints mean nothing; that the line is complete does.

usage: python tests/cusim/bench_dryrun.py [bench.py arguments]"""
import contextlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from cusim import build_sim  # noqa: E402

path = build_sim.build()
from semantic_suma_b200 import build as product_build  # noqa: E402

product_build.LIB = path
product_build.build = lambda *a, **k: path

import torch  # noqa: E402


class _Event:
    def __init__(self, enable_timing=False):
        self.t = 0.0

    def record(self, stream=None):
        self.t = time.time()

    def elapsed_time(self, other):
        return 1e3 * (other.t - self.t)


def _strip_device(fn):
    def wrapped(*a, **k):
        k.pop("device", None)
        return fn(*a, **k)
    return wrapped


torch.cuda.set_device = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.empty_cache = lambda *a, **k: None
torch.cuda.ExternalStream = lambda *a, **k: object()
torch.cuda.stream = lambda s: contextlib.nullcontext()
torch.cuda.Event = _Event
torch.Tensor.cuda = lambda self, *a, **k: self
torch.Tensor.pin_memory = lambda self, *a, **k: self
torch.empty = _strip_device(torch.empty)
torch.zeros = _strip_device(torch.zeros)
torch.tensor = _strip_device(torch.tensor)

import bench  # noqa: E402


def main():
    argv = sys.argv[1:] or ["--steps", "3", "--warmup", "3", "--preroll", "4", "--cpu-budget", "2"]
    sys.argv = ["bench.py"] + argv
    import io
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        bench.main()
    line = [l for l in buf.getvalue().splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    need = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"]
    missing = [k for k in need if k not in out]
    assert not missing, missing
    assert out["metric"] == "scans_per_sec" and out["n_gpus"] == 1 and out["gpu_launches"] > 0
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in out["e2e"], k
    assert out["e2e"]["h2d_bytes_per_step"] > 500_000
    assert "error" not in (out["roofline"] or {}), out["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in out["roofline"], k
    assert out["cpu_baseline"] and "error" not in out["cpu_baseline"], out["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in out["cpu_baseline"], k
    assert out["config"]["pose_drift_m"] < 0.5, out["config"]["pose_drift_m"]
    print("bench dry run ok: line complete (%d keys), %d launches in %d steps, top kernel %s, kernels %s"
          % (len(out), out["gpu_launches"], out["steps"], out["roofline"]["kernel"], sorted(out["kernels"])))


if __name__ == "__main__":
    main()
