"""N > 1 host logic on CPU (gloo, world_size 2): sharding by contiguous batch ranges, the all-gather of
partial rows, and the product's host-side final merge -- with the per-partition partial aggregation done
by the oracle so that no GPU is needed.  The 2-rank result must equal the 1-partition result."""
import os
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _by_date_plan():
    """group by l_shipdate (a non-dictionary key -> the hash path; thousands of groups): count, sum(l_quantity)"""
    from snappydata_b200 import plan as P
    from snappydata_b200.column_format import SqlType as T
    b = P.PlanBuilder()
    ship, qty = b.col(T.DATE, P.L_SHIPDATE), b.col(T.DOUBLE, P.L_QUANTITY)
    b.group_by(ship)
    b.count().sum(qty)
    return b.build()


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from oracle import oracle
    from snappydata_b200 import capi, lineitem, plan as P
    from snappydata_b200.exchange import PartialRowExchange, shard_batches
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    api = capi.product_api()      # host-side entry points only (sd_final_merge): no CUDA call
    total, per_batch = 90_001, 8_000
    out = {}
    for name, desc, lits in (("q1", P.q1_plan(), P.Q1_LITERALS), ("q6", P.q6_plan(), P.Q6_LITERALS), ("by_date", _by_date_plan(), [])):
        first_row, nrows, nb = shard_batches(total, per_batch, rank, world)
        assert first_row % per_batch == 0
        mine = lineitem.gen_table(total, per_batch, seed=4, batches=range(first_row // per_batch, first_row // per_batch + nb))
        assert sum(b.num_rows for b in mine) == nrows
        pl = oracle.plan(desc).set_literals(lits)
        for b in mine:
            pl.submit(b)
        ex = PartialRowExchange(torch, dist, world, "cpu")
        gathered = ex.all_gather(pl.finish_raw())
        if name == "by_date":   # ~2500 groups per rank: far beyond the initial slot -> it grew, identically on both ranks
            assert ex.regrows >= 1 and len(gathered) > 4096
        # the merged PARTIAL rows (what sd_plan_exchange hands to sd_plan_finish) feed the final merge to the same result
        merged = capi.partial_merge_raw(api, desc, gathered)
        out[name] = capi.final_merge(api, desc, gathered)
        via_partial = capi.final_merge(api, desc, merged)
        assert sorted(map(repr, via_partial)) == sorted(map(repr, out[name]))
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_exchange_and_merge_equals_single_partition():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import assert_rowsets_match
    from oracle import oracle
    from snappydata_b200 import capi, lineitem, plan as P
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    whole = lineitem.gen_table(90_001, 8_000, seed=4)
    for name, desc, lits, nk in (("q1", P.q1_plan(), P.Q1_LITERALS, 2), ("q6", P.q6_plan(), P.Q6_LITERALS, 0), ("by_date", _by_date_plan(), [], 1)):
        pl = oracle.plan(desc).set_literals(lits)
        for b in whole:
            pl.submit(b)
        want = oracle.final_merge(desc, pl.finish_raw())
        assert_rowsets_match(got[name], want, nk)


def test_shards_cover_the_table_exactly():
    from snappydata_b200.exchange import shard_batches
    for total in (1, 199_999, 200_000, 600_037_902):
        for world in (1, 2, 4, 8):
            rows = 0
            nxt = 0
            for r in range(world):
                first, n, nb = shard_batches(total, 200_000, r, world)
                assert first == nxt or n == 0
                nxt = first + n if n else nxt
                rows += n
            assert rows == total
