"""Sharded tally on real GPUs (needs >= 2): the fused peer-memory exchange (fheb_tally_peers_*) and the all-gather +
combine path must both give the oracle's words on every rank.  One process per GPU, NCCL, launched with torchrun."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import fheb200
from oracle_bindings import Oracle
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
n, q = 1024, 1099511678977
orc = Oracle()
for total in (world * 700 + 3, world, 5000):          # ragged split, one ballot per rank, several slabs per rank
    rng = np.random.default_rng(total)
    cts = rng.integers(0, q, size=(total, 2, n), dtype=np.uint64)
    cts[0, 0, :2] = [q + 1, 2**64 - 1]                  # unreduced words
    exp = orc.tally(cts, q)
    lo, hi = fheb200.shard_range(total, rank, world)
    mine = torch.from_numpy(cts[lo:hi].view(np.int64)).cuda()
    fused = fheb200.ShardedTally(n, q)
    general = fheb200.ShardedTally(n, q, fused=False)
    for _ in range(3):                                  # several epochs: both inbox parities
        a = fused.tally(mine)
        b = general.tally(mine)
        assert fused._peers is not None, "peer path was not set up"
        assert np.array_equal(a.cpu().numpy().view(np.uint64), exp), (rank, total, "fused")
        assert np.array_equal(b.cpu().numpy().view(np.uint64), exp), (rank, total, "general")
    fused.check()
    del fused, general
# many back-to-back epochs on a tiny count: a peer may run one call ahead and overwrite its flag with epoch k+1
# before this rank has looked at epoch k - the wait is "reached", not "equals" (no timeout, same words every time)
cts = np.random.default_rng(7).integers(0, q, size=(world * 2, 2, n), dtype=np.uint64)
exp = orc.tally(cts, q)
mine = torch.from_numpy(cts[rank * 2:rank * 2 + 2].view(np.int64)).cuda()
fused = fheb200.ShardedTally(n, q)
outs = [fused.tally(mine) for _ in range(400)]
torch.cuda.synchronize()
fused.check()
for o in outs[::37] + outs[-3:]:
    assert np.array_equal(o.cpu().numpy().view(np.uint64), exp), (rank, "back-to-back epochs")
del fused, outs
# a peer that never arrives: the waiting rank's words become all-ones, its status names the call, later calls fail
dist.barrier()
lone = fheb200.ShardedTally(n, q)
first = lone.tally(mine)            # sets the peers up (collective) and checks the first call
torch.cuda.synchronize()
import ctypes as C
fheb200.lib().fheb_tally_peers_set_timeout(lone._peers, C.c_double(0.2))
dist.barrier()
if rank == 0:
    bad = lone.tally(mine)          # nobody else takes part in this exchange
    torch.cuda.synchronize()
    assert bool((bad.cpu().numpy().view(np.uint64) == np.uint64(2**64 - 1)).all()), "timed-out exchange must poison its result"
    try:
        lone.check()
        raise SystemExit("status did not report the timed-out exchange")
    except fheb200.FheError:
        pass
    try:
        lone.tally(mine)
        raise SystemExit("a later call on the failed handle must be refused")
    except fheb200.FheError:
        pass
dist.barrier()
del lone
torch.cuda.synchronize()
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "sharded tally ok")
"""


def test_fused_and_general_sharded_tally_match_the_oracle(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    world = min(torch.cuda.device_count(), 8)
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29577", str(script)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.count("sharded tally ok") == world, r.stdout[-2000:] + r.stderr[-4000:]


def _group_case(fhe, torch, oracle, ndev, total, n=1024, q=1099511678977):
    import numpy as np

    rng = np.random.default_rng(total + ndev)
    cts = rng.integers(0, q, size=(total, 2, n), dtype=np.uint64)
    cts[0, 1, :3] = [q, q + 5, 2**64 - 1]  # unreduced words
    shards = []
    for r in range(ndev):
        lo, hi = fhe.shard_range(total, r, ndev)
        shards.append(torch.from_numpy(cts[lo:hi].view(np.int64)).to(f"cuda:{r}") if hi > lo else None)
    return cts, shards


@pytest.mark.parametrize("total", [1, 3, 700, 5000])
def test_single_process_group_on_the_visible_gpus(total):
    """fheb_tally_group_* / fheb_tally_sharded: ONE process drives every visible GPU (peer access, no IPC, no torch
    collectives).  On a 1-GPU box the group has one member and the kernel still runs its exchange stage (inbox, flag,
    epoch parity) against itself - the fused path is covered wherever the suite runs."""
    import numpy as np
    import torch

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fheb200
    from oracle_bindings import Oracle

    ndev = min(torch.cuda.device_count(), 8)
    n, q = 1024, 1099511678977
    grp = fheb200.TallyGroup(n, q, ndev=ndev)
    cts, shards = _group_case(fheb200, torch, Oracle(), ndev, total)
    exp = Oracle().tally(cts, q) if total > 1 else None
    for _ in range(3):  # both inbox parities
        got = grp.tally(shards)
        if total == 1:   # the group sums canonical partials: a lone ballot comes back reduced (batch_add would return it raw)
            assert np.array_equal(got, cts[0] % np.uint64(q))
        else:
            assert np.array_equal(got, exp)
    out_dev = torch.empty((2, n), dtype=torch.int64, device="cuda:0")
    grp.tally(shards, out=out_dev)
    assert np.array_equal(out_dev.cpu().numpy().view(np.uint64), exp if total > 1 else cts[0] % np.uint64(q))


def test_host_batches_spread_over_the_visible_gpus():
    """fheb_set_devices: one process, every visible GPU; host batches are split into contiguous shares, one host thread and
    one copy/compute pipeline per device, plans and keys replicated on first use.  Same words as the one-GPU path."""
    import numpy as np
    import torch

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fheb200
    from oracle_bindings import Oracle

    orc = Oracle()
    n, q = 4096, 4611686018326724609
    ndev = min(torch.cuda.device_count(), 8)
    rng = np.random.default_rng(11)
    batch = 96 * ndev + 5
    a = rng.integers(0, q, size=(batch, n), dtype=np.uint64)
    b = rng.integers(0, q, size=(batch, n), dtype=np.uint64)
    ring = fheb200.PolynomialRing(n, q)
    single_t, single_p = ring.to_ntt(a), ring.multiply(a, b)
    assert fheb200.set_devices() == torch.cuda.device_count()
    try:
        spread_t, spread_p = ring.to_ntt(a), ring.multiply(a, b)
        assert np.array_equal(spread_t, single_t) and np.array_equal(spread_p, single_p)
        fwd, inv, _, _, inv_n = orc.twiddles(n, q)
        for i in (0, batch // 2, batch - 1):
            assert np.array_equal(spread_p[i], orc.multiply(a[i:i + 1], b[i:i + 1], q, fwd, inv, inv_n)[0])
        cts = rng.integers(0, q, size=(batch, 2, n), dtype=np.uint64)
        assert np.array_equal(fheb200.tally_votes(cts, n, q), orc.tally(cts, q))
    finally:
        fheb200.set_devices([])
