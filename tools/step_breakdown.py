"""Per-phase host timings of one resident step under torchrun (diagnostic)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import bench
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
from snappydata_b200 import capi
api = capi.product_api(); api.check(api.init(lr))
q1 = (sys.argv[1] if len(sys.argv) > 1 else "q1") == "q1"
run = bench.QueryRun(api, torch, dist, q1, bench.SF100_ROWS if q1 else bench.SF10_ROWS, rank, world, lr)
for _ in range(5): run.step_resident()
acc = {}
def t(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); acc[name] = acc.get(name, 0) + time.perf_counter() - t0; return r
N = 20
for _ in range(N):
    p = run.plan
    t("reset+lits", lambda: p.reset().set_literals(run.lits))
    t("scan_store(launch+kernel)", lambda: p.scan_store(run.store))
    raw = t("finish", lambda: p.finish_raw())
    t("metrics", lambda: p.metrics())
    if world > 1:
        raw2 = t("all_gather", lambda: run.exchange.all_gather(raw))
    else:
        raw2 = raw
    t("final_merge", lambda: run.merge_plan.final_merge(raw2))
if rank == 0:
    print({k: round(v / N * 1e6, 1) for k, v in acc.items()}, "us per step; kernel_ms", p.metrics()["aggTimeNs"] / 1e6)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
