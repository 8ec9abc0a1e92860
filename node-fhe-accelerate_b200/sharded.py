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

On NCCL groups whose ranks can map each other's memory (the GPUs of one NVSwitch box) the exchange and
the combine are FUSED into the tally kernel (fheb_tally_peers_*: NVLink stores into every peer's inbox,
a flag per column chunk, one launch per rank instead of three); the all-gather + combine-kernel form
stays as the general path and is what the first call is checked against.
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
                 local_fn: Optional[Callable] = None, combine_fn: Optional[Callable] = None, fused: Optional[bool] = None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.degree, self.modulus = degree, modulus
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.local_fn = local_fn or (lambda cts: api.tally_votes(cts, degree, modulus))
        self.combine_fn = combine_fn or (lambda parts: api.tally_combine(parts, degree, modulus))
        self._gathered = None
        self._peers = None            # fused peer-memory path (set up lazily on the first CUDA call)
        self._peers_tried = local_fn is not None or combine_fn is not None or fused is False
        self._peers_checked = False

    def _setup_peers(self, device):
        """Exchange the IPC handles of the per-rank inboxes with one all-gather; every rank must agree on success."""
        import ctypes as C

        import numpy as np
        import torch

        self._peers_tried = True
        ok = 1
        why = ""
        handle = np.zeros(64, np.uint8)
        h = C.c_void_p()
        try:
            api.check(api.lib().fheb_tally_peers_create(self.degree, self.modulus, self.world, self.rank, C.byref(h),
                                                        handle.ctypes.data_as(C.c_void_p)))
        except api.FheError as exc:
            ok = 0
            why = str(exc)
        mine = torch.from_numpy(handle).to(device)
        everyone = torch.empty(self.world * 64, dtype=torch.uint8, device=device)
        self.dist.all_gather_into_tensor(everyone, mine, group=self.group)
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        if ok:
            table = np.ascontiguousarray(everyone.cpu().numpy())
            if api.lib().fheb_tally_peers_connect(h, table.ctypes.data_as(C.c_void_p)) != 0:
                flag[0] = 0
                why = api.lib().fheb_last_error().decode()
        self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 1:
            self._peers = h
        else:
            import sys

            print(f"fheb200: rank {self.rank}: fused sharded tally unavailable ({why or 'a peer could not set it up'}); "
                  "using all-gather + combine", file=sys.stderr)
            if h:
                api.lib().fheb_tally_peers_destroy(h)

    def _fused(self, local_cts):
        import ctypes as C

        import torch

        cts = api.as_words(local_cts)
        out = torch.empty((2, self.degree), dtype=cts.dtype, device=cts.device)
        count = cts.numel() // (2 * self.degree)
        api.check(api.lib().fheb_tally_peers_run(self._peers, C.c_void_p(cts.data_ptr()), count, C.c_void_p(out.data_ptr()),
                                                 C.c_void_p(torch.cuda.current_stream(cts.device).cuda_stream)))
        # (fheb_tally_peers_run itself refuses to launch once an earlier exchange of this handle has timed out: the
        # status word is host-mapped, reading it costs nothing.  A timed-out call also leaves all-ones words in `out`.)
        return out

    def check(self):
        """After synchronising: raises if any exchange issued so far has timed out (a peer rank never arrived)."""
        import ctypes as C

        if self._peers is not None:
            timed_out = C.c_int(0)
            api.check(api.lib().fheb_tally_peers_status(self._peers, C.byref(timed_out)))
            if timed_out.value:
                raise api.FheError(4, f"fused sharded tally: a peer rank did not take part in exchange {timed_out.value} "
                                      "(its result and all later ones are invalid)")

    def __del__(self):
        h, self._peers = getattr(self, "_peers", None), None
        if h:
            try:
                api.lib().fheb_tally_peers_destroy(h)
            except Exception:
                pass

    def tally(self, local_cts, local_count: Optional[int] = None):
        """local_cts: this rank's ballots [count_r][2][N] (torch tensor on this rank's device).
        Every rank must hold at least one ballot.  Returns the global tally [2][N] on every rank."""
        import torch

        if self.world > 1 and not self._peers_tried and getattr(local_cts, "is_cuda", False) \
                and self.dist.get_backend(self.group) == "nccl":
            self._setup_peers(local_cts.device)
        if self._peers is not None and getattr(local_cts, "is_cuda", False):
            res = self._fused(local_cts)
            if not self._peers_checked:  # first call: the general path must give the same words on every rank
                self._peers_checked = True
                ref = self._general(local_cts)
                same = torch.tensor([int(torch.equal(res.view(-1), ref.view(-1)))], dtype=torch.int32, device=res.device)
                self.dist.all_reduce(same, op=self.dist.ReduceOp.MIN, group=self.group)
                if int(same.item()) != 1:  # never expected; keep the general GPU path and say so
                    import sys

                    print("fheb200: fused sharded tally disagrees with the all-gather path; using the general path", file=sys.stderr)
                    api.lib().fheb_tally_peers_destroy(self._peers)
                    self._peers = None
                    return ref
            return res
        return self._general(local_cts)

    def _general(self, local_cts):
        import torch

        partial = self.local_fn(local_cts)
        if self.world == 1:
            return partial
        if self._gathered is None or self._gathered.device != partial.device or self._gathered.dtype != partial.dtype:
            self._gathered = torch.empty(self.world * 2 * self.degree, dtype=partial.dtype, device=partial.device)
        self.dist.all_gather_into_tensor(self._gathered, partial.contiguous().view(-1), group=self.group)
        return self.combine_fn(self._gathered.view(self.world, 2, self.degree))


class TallyGroup:
    """The sharded tally driven from ONE process that owns several GPUs (fheb_tally_group_*): the shape the
    reference's single-process addon needs (src/native/lib.rs:23-133).  shards[i] lives on devices[i]."""

    def __init__(self, degree: int, modulus: int, devices=None, ndev: Optional[int] = None):
        import ctypes as C

        self.degree, self.modulus = degree, modulus
        self.devices = list(devices) if devices is not None else list(range(int(ndev)))
        arr = (C.c_int * len(self.devices))(*self.devices)
        self._h = C.c_void_p()
        api.check(api.lib().fheb_tally_group_create(degree, modulus, arr, len(self.devices), C.byref(self._h)))

    def tally(self, shards, out=None):
        """shards: one tensor [count_i][2][N] per device of the group (None / empty = no ballots there).
        Returns the global tally [2][N] as a numpy array (or fills `out`, host array or CUDA tensor)."""
        import ctypes as C

        import numpy as np

        n = len(self.devices)
        assert len(shards) == n
        ptrs = (C.c_void_p * n)()
        counts = (C.c_size_t * n)()
        keep = []
        for k, sh in enumerate(shards):
            if sh is None or sh.numel() == 0:
                ptrs[k], counts[k] = None, 0
                continue
            w = api.as_words(sh)
            assert w.is_cuda and w.device.index == self.devices[k], "shard k must live on device k of the group"
            keep.append(w)
            ptrs[k], counts[k] = w.data_ptr(), w.numel() // (2 * self.degree)
        if out is None:
            out = np.empty((2, self.degree), dtype=np.uint64)
        import torch

        for d in self.devices:  # the shards were produced on torch's streams
            torch.cuda.synchronize(d)
        api.check(api.lib().fheb_tally_sharded(self._h, ptrs, counts, C.c_void_p(api._ptr(out))))
        return out

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                api.lib().fheb_tally_group_destroy(h)
            except Exception:
                pass
