"""Host<->device copy bandwidth of every visible GPU, alone and together (one process, pinned buffers, CUDA events +
wall clock), and where each GPU hangs in the host: NUMA node and local CPUs of its PCIe function.  Names the limiter
of the end-to-end (host-buffer) numbers when several GPUs stream at once.

    python tools/prof_pcie.py [MB per buffer = 256]
"""
import glob
import os
import sys
import time

import torch

MB = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ndev = torch.cuda.device_count()
print(f"{ndev} GPUs, {os.cpu_count()} CPUs, NUMA nodes: {sorted(os.path.basename(p) for p in glob.glob('/sys/devices/system/node/node[0-9]*'))}")
for node in sorted(glob.glob('/sys/devices/system/node/node[0-9]*')):
    try:
        print(f"  {os.path.basename(node)}: cpus {open(node + '/cpulist').read().strip()}  mem {open(node + '/meminfo').read().split()[3]} kB")
    except Exception as exc:
        print("  ", node, exc)
for d in range(ndev):
    p = torch.cuda.get_device_properties(d)
    bdf = None
    try:
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        numa = open(base + "/numa_node").read().strip()
        cpus = open(base + "/local_cpulist").read().strip()
        speed = open(base + "/current_link_speed").read().strip()
        width = open(base + "/current_link_width").read().strip()
        print(f"  GPU{d} {bdf}: numa_node {numa}, local cpus {cpus}, link {speed} x{width}")
    except Exception as exc:
        print(f"  GPU{d} {bdf}: {exc}")

words = MB * (1 << 20) // 8
host_in = [torch.empty(words, dtype=torch.int64).pin_memory() for _ in range(ndev)]
host_out = [torch.empty(words, dtype=torch.int64).pin_memory() for _ in range(ndev)]
dev_a = [torch.empty(words, dtype=torch.int64, device=f"cuda:{d}") for d in range(ndev)]
dev_b = [torch.empty(words, dtype=torch.int64, device=f"cuda:{d}") for d in range(ndev)]
s_in = [torch.cuda.Stream(device=d) for d in range(ndev)]
s_out = [torch.cuda.Stream(device=d) for d in range(ndev)]


def run(devs, h2d, d2h, reps=6):
    for d in devs:
        torch.cuda.synchronize(d)
    t0 = time.perf_counter()
    for _ in range(reps):
        for d in devs:
            if h2d:
                with torch.cuda.stream(s_in[d]):
                    dev_a[d].copy_(host_in[d], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s_out[d]):
                    host_out[d].copy_(dev_b[d], non_blocking=True)
    for d in devs:
        torch.cuda.synchronize(d)
    dt = time.perf_counter() - t0
    per_dir = MB / 1024 * reps * len(devs) / dt
    return per_dir


run(list(range(ndev)), True, True, 2)
print(f"\n{MB} MB buffers; GB/s PER DIRECTION summed over the GPUs taking part")
for d in range(ndev):
    print(f"GPU{d} alone: H2D {run([d], True, False):6.1f}  D2H {run([d], False, True):6.1f}  both ways {run([d], True, True):6.1f} each")
if ndev > 1:
    for other in range(1, ndev):
        print(f"GPU0+GPU{other}: both ways {run([0, other], True, True):6.1f} each way in total")
    for k in (2, 4, 8):
        if k <= ndev:
            devs = list(range(k))
            print(f"GPUs 0..{k - 1}: H2D {run(devs, True, False):6.1f}  D2H {run(devs, False, True):6.1f}  both ways {run(devs, True, True):6.1f} each way in total")
