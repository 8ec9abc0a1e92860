"""Pipe rooflines of bench.py, built from two measurements (nothing here is a pasted constant):

* the B200's per-instruction-class issue rates, measured LIVE by tools/microbench/pipes (built by
  __graft_entry__.build() from pipes.cu) on the same GPU right before the timed region - or, when the binary is
  missing, read from the committed raw output profiles/r02_pipes.txt, saying so;
* the kernel's DYNAMIC instruction mix per work unit (executed warp instructions per opcode class, from the source
  page of an `ncu --set full --import-source on` capture, aggregated by tools/ncu_opmix.py into profiles/r02_opmix.json),
  together with that capture's DRAM traffic.

A kernel that keeps pipe P busy for t_P = sum_class(n_class / rate_class) seconds per unit cannot run faster than
1 / max_P t_P units per second: that is `peak`; `achieved` is units / CUDA-event time of this run; frac = achieved / peak.
Pipes: "fma" carries IMAD.WIDE (measured ~25.8 thread-ops/clk/SM), IMAD.HI and every other IMAD (~63); "fp64" carries
DFMA/DADD/DMUL (~63) and - measured - shares its issue slot with IMAD.WIDE (dfma_plus_imad_wide_tops), so the two
are added; "alu" carries IADD3/LOP3/SHF/SEL/ISETP.
"""
from __future__ import annotations

import json
import os
import subprocess

ROOT = os.path.dirname(os.path.abspath(__file__))


def pipe_peaks(device_index: int = 0) -> dict:
    exe = os.path.join(ROOT, "tools", "microbench", "pipes")
    if os.path.exists(exe):
        try:
            r = subprocess.run([exe, "--json", str(device_index)], capture_output=True, text=True, timeout=120)
            for line in reversed(r.stdout.strip().splitlines()):
                if line.startswith("{"):
                    d = json.loads(line)
                    d["source"] = "live: tools/microbench/pipes --json on this GPU"
                    return d
        except Exception:
            pass
    path = os.path.join(ROOT, "profiles", "r02_pipes.txt")
    with open(path) as f:
        for line in reversed(f.read().strip().splitlines()):
            if line.startswith("{"):
                d = json.loads(line)
                d["source"] = "profiles/r02_pipes.txt (pipes binary not available in this run)"
                return d
    raise RuntimeError("no pipe-rate measurement available")


def load_opmix() -> dict:
    with open(os.path.join(ROOT, "profiles", "r02_opmix.json")) as f:
        return json.load(f)


def pipe_roofline(mix_key: str, units: float, seconds: float, peaks: dict, unit_name: str, opmix: dict | None = None) -> dict:
    """units processed in `seconds` (this run) against the pipe ceiling of the kernel whose mix is `mix_key`."""
    opmix = opmix or load_opmix()
    m = opmix[mix_key]
    per = m["per_unit"]  # executed THREAD instructions per unit (= warp instructions per warp-unit)
    wide, hi, narrow = per.get("imad_wide", 0.0), per.get("imad_hi", 0.0), per.get("imad", 0.0)
    fp64, alu = per.get("fp64", 0.0), per.get("alu", 0.0)
    r_wide, r_imad, r_dfma = peaks["imad_wide_tops"] * 1e12, peaks["imad_tops"] * 1e12, peaks["dfma_tops"] * 1e12
    r_alu = peaks.get("alu_tops", peaks["imad_tops"]) * 1e12
    # IMAD.HI.U32 (the 32-bit kernels' Shoup quotient): measured directly when the microbenchmark has the mode
    r_hi = peaks.get("imad_hi_tops", peaks["umul64hi_tops"] * 6.0) * 1e12
    t = {
        # DFMA and IMAD.WIDE do not overlap (measured: together they take the sum of their times), so the FP64 work
        # is charged to the same issue path as the wide multiplies
        "fma+fp64": wide / r_wide + hi / max(r_hi, 1.0) + narrow / r_imad + fp64 / r_dfma,
        "alu": alu / r_alu,
    }
    pipe = max(t, key=t.get)
    peak = 1.0 / t[pipe]
    achieved = units / seconds
    kind = "fp64-pipe" if fp64 / r_dfma > 0.5 * t["fma+fp64"] else "integer-pipe"
    return {"bound": kind, "achieved": achieved / 1e9, "peak": peak / 1e9, "unit": f"G {unit_name}/s", "frac": achieved / peak,
            "limiting_pipe": pipe, "pipe_seconds_per_unit": t,
            "traffic": m.get("dram_bytes"), "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of the captured launch (ncu), per launch",
            "mix_per_unit": {k: per.get(k, 0.0) for k in ("imad_wide", "imad_hi", "imad", "fp64", "alu", "lsu", "other")},
            "mix_source": f"profiles/r02_opmix.json[{mix_key}] (kernel {m.get('kernel', '?')[:60]}, ncu source page)",
            "ncu_pipe_busy_pct": m.get("pipe_busy_pct"),
            "pipe_rates_tops": {k: peaks[k] for k in ("imad_wide_tops", "imad_hi_tops", "imad_tops", "dfma_tops", "alu_tops") if k in peaks},
            "pipe_rates_source": peaks.get("source")}
