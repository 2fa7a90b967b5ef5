"""Round-2 probe (one GPU): per-step cost of the one-call execution path and a chunk-size sweep, Q6 SF-10 and Q1 SF-100.
    python tools/r2_probe.py [q6|q1|both] [steps=30]
Prints kernel ms (device events around the launch, from sd_plan_metrics) and whole-step ms (CUDA events around K steps,
each step = sd_plan_execute_store + final merge) per chunk size."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from snappydata_b200 import capi  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "both"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
torch.cuda.set_device(0)
api = capi.product_api()
api.check(api.init(0))
for q1 in ([False, True] if which == "both" else [which == "q1"]):
    total = bench.SF100_ROWS if q1 else bench.SF10_ROWS
    run = bench.QueryRun(api, torch, None, q1, total, 0, 1, 0)
    for chunk in (0, 2048, 4096, 8192, 16384, 32768):
        if chunk:
            os.environ["SD_TUNE_CHUNK_ROWS"] = str(chunk)
        run.plan = capi.Plan(api, run.desc)
        run.plan.set_stream(torch.cuda.current_stream().cuda_stream)
        run.merge_plan = run.plan
        run.launches = run.kernel_ns = run.algo_bytes = 0
        for _ in range(5):
            run.step_resident()
        run.launches = run.kernel_ns = run.algo_bytes = 0
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            run.step_resident()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        kms = run.kernel_ns / 1e6 / max(1, run.launches)
        gbs = run.algo_bytes / max(1, run.launches) / (kms / 1e3) / 1e9
        print(f"{'q1' if q1 else 'q6'} chunk_rows={chunk or 'default':>7}: kernel {kms:.4f} ms ({gbs:.0f} GB/s), step {ms:.4f} ms "
              f"(overhead {1e3 * (ms - kms):.1f} us), {total / ms / 1e6:.1f} G rows/s", flush=True)
    os.environ.pop("SD_TUNE_CHUNK_ROWS", None)
    del run
    torch.cuda.empty_cache()
