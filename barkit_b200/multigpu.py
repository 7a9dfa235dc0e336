"""Host-side sharding logic shared by bench.py (one process per GPU under torchrun) and the tests.

The extract path shards by batch: in bench.py and these helpers batch b of a run goes to rank b mod world.
(The `barkit` CLI shards by batch too, but inside one process and dynamically: bk_run.cc's per-GPU workers
take the next unclaimed batch with a fetch_add and an ordered writer restores batch order.)  There is no collective on the data path;
torch.distributed is used only to (a) line the ranks up before and after the timed region and
(b) take the max over ranks of the device time.  Outputs are reassembled in batch order, which makes
the N-rank bytes identical to the 1-rank bytes (run.rs:79,194 ordering).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def batches_of(n_records: int, batch_records: int) -> List[Tuple[int, int]]:
    """[(first, last)) record ranges of a run, in order."""
    return [(a, min(a + batch_records, n_records)) for a in range(0, n_records, batch_records)]


def batches_for_rank(n_batches: int, rank: int, world: int) -> List[int]:
    """Round-robin deal: the batch indices rank `rank` processes, ascending."""
    return list(range(rank, n_batches, world))


def rank_first_index(rank: int, slot: int, pairs_per_step: int) -> int:
    """bench.py: first synthetic record index of resident batch `slot` on `rank` (disjoint per rank)."""
    return (rank << 36) + slot * pairs_per_step


def reassemble(parts: Sequence[Tuple[int, bytes]]) -> bytes:
    """Ordered writer: (batch index, output bytes) from any rank in any order -> the run's bytes."""
    return b"".join(p for _, p in sorted(parts, key=lambda t: t[0]))


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """Device time of a multi-GPU step = max over ranks (never wall clock of one rank)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(dist=None, device=None) -> None:
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return
    import torch
    t = torch.zeros(1, device=device or "cpu")
    dist.all_reduce(t)
    if device is not None and str(device).startswith("cuda"):
        torch.cuda.synchronize()
