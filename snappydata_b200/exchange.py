"""The one exchange of the hot path: partial-aggregate rows of every partition (= GPU) are gathered and
merged (SnappyStrategies.scala:566-604 plans partial -> Exchange -> final; CollectAggregateExec.scala:67-121
merges on the driver).  Here: one `all_gather` over torch.distributed (NCCL over NVLink on the GPU box, gloo in
the CPU tests) of the length-prefixed partial-row bytes, then the host-side final merge on every rank.
Payload is a few hundred bytes, so the exchange is latency bound; nothing is fused with it.

On the GPU box the product's exchange is sd_plan_exchange (C ABI, ncclAllGather inside libsnappygpu.so; capi.Comm);
this module is the same protocol over torch.distributed for the CPU (gloo) tests and for hosts that already own a
process group: a fixed-capacity slot per rank whose header carries the real length, grown in lock step when some
rank's rows do not fit.
"""
from __future__ import annotations

from typing import Tuple

SLOT_BYTES = 4096   # initial gather slot per rank: [int64 length][partial rows]; doubles (on every rank) until all fit


def shard_batches(total_rows: int, rows_per_batch: int, rank: int, world: int) -> Tuple[int, int, int]:
    """Contiguous batch range of `rank`: -> (first_row, nrows, nbatches).  Batches never span ranks
    (a partition is a set of whole buckets/batches, JDBCSourceAsColumnarStore.scala:755-762)."""
    nb = (total_rows + rows_per_batch - 1) // rows_per_batch
    lo, hi = nb * rank // world, nb * (rank + 1) // world
    first_row = lo * rows_per_batch
    nrows = min(total_rows, hi * rows_per_batch) - first_row
    return first_row, max(0, nrows), hi - lo


class PartialRowExchange:
    """Reusable buffers for the all-gather of partial rows."""

    def __init__(self, torch, dist, world: int, device: str, slot_bytes: int = SLOT_BYTES):
        self.torch, self.dist, self.world, self.device = torch, dist, world, device
        self.cuda = device.startswith("cuda")
        self.regrows = 0
        self._alloc(slot_bytes)

    def _alloc(self, slot: int):
        torch = self.torch
        self.slot = slot
        self.inp = torch.zeros(slot, dtype=torch.uint8, device=self.device)
        self.out = torch.zeros(slot * self.world, dtype=torch.uint8, device=self.device)
        self.pin_in = torch.zeros(slot, dtype=torch.uint8)
        self.pin_out = torch.zeros(slot * self.world, dtype=torch.uint8)
        if self.cuda:
            self.pin_in, self.pin_out = self.pin_in.pin_memory(), self.pin_out.pin_memory()
        self.np_in = self.pin_in.numpy()      # views over the (pinned) staging buffers
        self.np_out = self.pin_out.numpy()
        self.len_view = self.np_in[:8].view("<i8")

    def all_gather(self, raw: bytes) -> bytes:
        """-> concatenation of every rank's partial rows, in rank order (any size: the slot grows on all ranks)."""
        n = len(raw)
        while True:
            slot = self.slot
            fit = min(n, slot - 8)
            self.len_view[0] = n
            if fit:
                self.np_in[8:8 + fit] = memoryview(raw)[:fit]
            self.inp.copy_(self.pin_in, non_blocking=True)
            self.dist.all_gather_into_tensor(self.out, self.inp)
            self.pin_out.copy_(self.out, non_blocking=True)
            if self.cuda:
                self.torch.cuda.current_stream().synchronize()
            o = self.np_out
            lens = [int(o[r * slot: r * slot + 8].view("<i8")[0]) for r in range(self.world)]
            if max(lens) + 8 <= slot:
                return b"".join(o[r * slot + 8: r * slot + 8 + lens[r]].tobytes() for r in range(self.world))
            while slot < max(lens) + 8:   # every rank sees the same lengths: same decision everywhere
                slot *= 2
            self.regrows += 1
            self._alloc(slot)
