"""The one exchange of the hot path: partial-aggregate rows of every partition (= GPU) are gathered and
merged (SnappyStrategies.scala:566-604 plans partial -> Exchange -> final; CollectAggregateExec.scala:67-121
merges on the driver).  Here: one `all_gather` over torch.distributed (NCCL over NVLink on the GPU box, gloo in
the CPU tests) of the length-prefixed partial-row bytes, then the host-side final merge on every rank.
Payload is a few hundred bytes, so the exchange is latency bound; nothing is fused with it.
"""
from __future__ import annotations

from typing import Tuple

SLOT_BYTES = 4096   # fixed-size gather slot per rank: [int64 length][partial rows]


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

    def __init__(self, torch, dist, world: int, device: str):
        self.torch, self.dist, self.world, self.device = torch, dist, world, device
        cuda = device.startswith("cuda")
        self.inp = torch.zeros(SLOT_BYTES, dtype=torch.uint8, device=device)
        self.out = torch.zeros(SLOT_BYTES * world, dtype=torch.uint8, device=device)
        self.pin_in = torch.zeros(SLOT_BYTES, dtype=torch.uint8)
        self.pin_out = torch.zeros(SLOT_BYTES * world, dtype=torch.uint8)
        if cuda:
            self.pin_in, self.pin_out = self.pin_in.pin_memory(), self.pin_out.pin_memory()
        self.np_in = self.pin_in.numpy()      # views over the (pinned) staging buffers
        self.np_out = self.pin_out.numpy()
        self.len_view = self.np_in[:8].view("<i8")
        self.cuda = cuda

    def all_gather(self, raw: bytes) -> bytes:
        """-> concatenation of every rank's partial rows, in rank order."""
        n = len(raw)
        if n + 8 > SLOT_BYTES:
            raise ValueError(f"partial rows of one partition ({n} bytes) exceed the {SLOT_BYTES}-byte gather slot")
        self.len_view[0] = n
        if n:
            self.np_in[8:8 + n] = memoryview(raw)
        self.inp.copy_(self.pin_in, non_blocking=True)
        self.dist.all_gather_into_tensor(self.out, self.inp)
        self.pin_out.copy_(self.out, non_blocking=True)
        if self.cuda:
            self.torch.cuda.current_stream().synchronize()
        o = self.np_out
        parts = []
        for r in range(self.world):
            base = r * SLOT_BYTES
            ln = int(o[base: base + 8].view("<i8")[0])
            parts.append(o[base + 8: base + 8 + ln].tobytes())
        return b"".join(parts)
