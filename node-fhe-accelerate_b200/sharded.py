"""Sharded encrypted-ballot tally over the GPUs of one box.

One process per GPU (torch.distributed, NCCL over NVLink).  Ballots are split into contiguous
ranges; each rank folds its range with fheb_tally (HBM-bound kernel), the per-rank partial
tallies ([2][N] words each) are exchanged with ONE all-gather, and every rank folds the
gathered partials with the modular-add kernel (fheb_tally_combine).  Modular addition on
canonical residues is associative and commutative, so the result words are identical to the
reference's EncryptionEngine::batch_add / tally_votes (cpp/src/encryption.cpp:1061-1067,
1327-1458) for any split.

The exchange is the only data-path collective of the hot path; NTT / polymul / bootstrap
batches shard with no communication (see shard_range).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

from . import api


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of `total` independent units: rank r gets [lo, hi)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("rank/world out of range")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedTally:
    """tally_votes over `world` ranks.

    local_fn / combine_fn default to the CUDA kernels; the CPU (gloo) tests of the host
    logic inject checkers there - the product path never does.
    """

    def __init__(self, degree: int, modulus: int, group=None,
                 local_fn: Optional[Callable] = None, combine_fn: Optional[Callable] = None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.degree, self.modulus = degree, modulus
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.local_fn = local_fn or (lambda cts: api.tally_votes(cts, degree, modulus))
        self.combine_fn = combine_fn or (lambda parts: api.tally_combine(parts, degree, modulus))
        self._gathered = None

    def tally(self, local_cts, local_count: Optional[int] = None):
        """local_cts: this rank's ballots [count_r][2][N] (torch tensor on this rank's device).
        Every rank must hold at least one ballot.  Returns the global tally [2][N] on every rank."""
        import torch

        partial = self.local_fn(local_cts)
        if self.world == 1:
            return partial
        if self._gathered is None or self._gathered.device != partial.device or self._gathered.dtype != partial.dtype:
            self._gathered = torch.empty(self.world * 2 * self.degree, dtype=partial.dtype, device=partial.device)
        self.dist.all_gather_into_tensor(self._gathered, partial.contiguous().view(-1), group=self.group)
        return self.combine_fn(self._gathered.view(self.world, 2, self.degree))
