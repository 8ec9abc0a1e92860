"""Phase timing of the sharded tally (local fold / all-gather / combine) with CUDA events:
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/prof_tally_sharded.py [ballots_total]"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fheb200  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
eager = "--eager" in sys.argv  # bench.py's form: communicator created eagerly on the rank's device
if world > 1:
    if eager:
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
    else:
        dist.init_process_group("nccl")
args = [a for a in sys.argv[1:] if not a.startswith("--")]
total = int(args[0]) if args else 1 << 20
n, q = 1024, 1099511678977
per = total // world
cts = torch.empty((per, 2, n), dtype=torch.int64, device="cuda")
fheb200.synth_ballots(cts, rank * per, per, n, q, 0xB200)
st = fheb200.ShardedTally(n, q)
gathered = torch.empty(world * 2 * n, dtype=torch.int64, device="cuda")
for _ in range(5):
    st.tally(cts)
torch.cuda.synchronize()
iters = 30
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(iters)]
if world > 1:
    dist.barrier()
for i in range(iters):
    ev[i][0].record()
    part = fheb200.tally_votes(cts, n, q)
    ev[i][1].record()
    if world > 1:
        dist.all_gather_into_tensor(gathered, part.view(-1))
    ev[i][2].record()
    res = fheb200.tally_combine(gathered.view(world, 2, n), n, q) if world > 1 else part
    ev[i][3].record()
torch.cuda.synchronize()
loc = sum(e[0].elapsed_time(e[1]) for e in ev) / iters
gat = sum(e[1].elapsed_time(e[2]) for e in ev) / iters
com = sum(e[2].elapsed_time(e[3]) for e in ev) / iters
whole = ev[0][0].elapsed_time(ev[-1][3]) / iters
if world > 1:  # the fused path (one launch per rank: local fold + peer-memory exchange + combine)
    fz = fheb200.ShardedTally(n, q)
    for _ in range(5):
        fz.tally(cts)
    torch.cuda.synchronize()
    dist.barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(iters):
        r2 = fz.tally(cts)
    f1.record()
    torch.cuda.synchronize()
    fused_us = f0.elapsed_time(f1) / iters * 1e3
    same = torch.equal(r2.view(-1), res.view(-1))
    print(f"rank {rank}/{world}: fused path {'ON' if fz._peers is not None else 'OFF'}: iteration {fused_us:.1f} us -> "
          f"{total / fused_us:.1f} M ballots/s, same words as the general path: {same}")
print(f"rank {rank}/{world}: {per} ballots/rank  local {loc * 1e3:.1f} us  all-gather {gat * 1e3:.1f} us  combine {com * 1e3:.1f} us  "
      f"iteration {whole * 1e3:.1f} us -> {total / whole / 1e3:.1f} M ballots/s")
