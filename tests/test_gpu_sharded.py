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
    del fused, general
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
