"""Worker of tests/test_gpu_multirank.py (launched under torchrun, one rank per GPU): every rank scans its own contiguous batch
range on its GPU, sd_plan_exchange (ncclAllGather inside libsnappygpu.so) merges the partial rows on every rank, and the
final result must equal the oracle's over the WHOLE table -- dense-table, no-key, hash-table (thousands of groups: the gather
slot grows in lock step) and string-key plans."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from helpers import assert_rowsets_match  # noqa: E402
from oracle import oracle  # noqa: E402
from snappydata_b200 import capi, lineitem, plan as P  # noqa: E402
from snappydata_b200.column_format import SqlType as T  # noqa: E402
from snappydata_b200.exchange import shard_batches  # noqa: E402
from snappydata_b200.plan import PlanBuilder  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
api = capi.product_api()
api.check(api.init(local))


def bcast(b):
    box = [b]
    dist.broadcast_object_list(box, src=0)
    return box[0]


comm = capi.Comm(api, rank, world, local, bcast)
TOTAL, PER = 330_001, 20_000
table = lineitem.gen_table(TOTAL, PER, seed=9)
first_row, nrows, nb = shard_batches(TOTAL, PER, rank, world)
mine = table[first_row // PER: first_row // PER + nb]


def by_date():
    b = PlanBuilder()
    ship, qty = b.col(T.DATE, P.L_SHIPDATE), b.col(T.DOUBLE, P.L_QUANTITY)
    b.group_by(ship)
    b.count().sum(qty)
    return b.build()


def by_flag_and_date():
    b = PlanBuilder()
    rf, ship, price = b.col(T.STRING, P.L_RETURNFLAG), b.col(T.DATE, P.L_SHIPDATE), b.col(T.DOUBLE, P.L_EXTENDEDPRICE)
    b.filter(ship > b.lit(T.DATE))
    b.group_by(rf, ship)
    b.count().min(price).max(rf)
    return b.build()


def form(which):
    """which form this rank's blob takes in sd_plan_exchange: "dense" = dictionaries + raw device state (default for dense /
    no-key plans), "rows" = partial rows by value, "mixed" = rank 0 dense, the others by value (receivers decode per blob)"""
    if which == "rows" or (which == "mixed" and rank != 0):
        os.environ["SD_TUNE_EXCHANGE_ROWS"] = "1"
    else:
        os.environ.pop("SD_TUNE_EXCHANGE_ROWS", None)


for name, desc, lits, nk, forms in (("q1", P.q1_plan(), P.Q1_LITERALS, 2, ("dense", "rows", "mixed")), ("q6", P.q6_plan(), P.Q6_LITERALS, 0, ("dense", "rows")),
                                    ("by_date", by_date(), [], 1, ("dense",)), ("by_flag_and_date", by_flag_and_date(), [10000], 2, ("dense",))):
    op = oracle.plan(desc).set_literals(lits)
    for b in table:
        op.submit(b)
    want = oracle.final_merge(desc, op.finish_raw())
    for which in forms:
        form(which)
        gp = capi.Plan(api, desc).set_literals(lits)
        for b in mine:
            gp.submit(b)
        gp.exchange(comm)
        merged = gp.finish_raw()
        got = gp.final_merge(merged)
        assert_rowsets_match(got, want, nk)
        m = gp.metrics()
        assert m["numOutputRows"] >= (1 if nk == 0 else 0) and m["aggTimeNs"] > 0
        if rank == 0:
            print(name, which, "ok:", len(got), "groups", comm.info(), flush=True)
form("dense")
assert comm.info()["regrows"] >= 1
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("MULTIRANK OK", flush=True)
