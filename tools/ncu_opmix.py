"""Dynamic instruction mix of the launches in an .ncu-rep (read here, no GPU): executed WARP instructions per SASS opcode
class, from the report's source page, divided by the work units of the launch.

    python tools/ncu_opmix.py <report | source.csv.gz> <units per launch> [--json key] [--raw raw.csv] [kernel-index]

--raw adds the launch's DRAM traffic, duration and pipe utilisation from the raw page (exported CSV) to the entry.

The classes are what the integer / FP64 pipe rooflines of bench.py need:
  imad_wide   IMAD.WIDE*            (64-bit product of 32-bit words: quarter-rate on the FMA-heavy pipe, tools/microbench/pipes.cu)
  imad_hi     IMAD.HI*
  imad        every other IMAD* (IMAD, IMAD.X, IMAD.IADD, IMAD.MOV, IMAD.SHL ...): FMA pipe, 64 lanes/clk/SM
  fp64        DFMA, DADD, DMUL, DSETP, F2F.F64*, I2F.F64*, F2I*.F64
  alu         IADD3*, LOP3*, SHF*, SEL, ISETP*, PRMT, LEA*, MOV, IABS, FSEL ... (ALU pipe)
  lsu         LD*, ST*, ATOM*, RED*
  other       everything else (BAR, BRA, S2R, ...)
"""
import csv
import io
import json
import subprocess
import sys
from collections import Counter


def classify(op: str) -> str:
    if op.startswith("IMAD.WIDE"):
        return "imad_wide"
    if op.startswith("IMAD.HI"):
        return "imad_hi"
    if op.startswith("IMAD"):
        return "imad"
    if op[:4] in ("DFMA", "DADD", "DMUL", "DSET", "DMNM") or ".F64" in op:
        return "fp64"
    if op.split(".")[0] in ("LDG", "STG", "LDS", "STS", "LD", "ST", "LDL", "STL", "ATOM", "ATOMG", "ATOMS", "RED", "LDGSTS", "LDSM", "UBLKCP", "LDC", "LDCU"):
        return "lsu"
    if op.split(".")[0] in ("IADD3", "IADD", "LOP3", "SHF", "SEL", "ISETP", "PRMT", "LEA", "MOV", "IABS", "FSEL", "IMNMX", "VIADD", "VIMNMX", "SGXT", "BMSK", "FLO", "POPC", "PLOP3", "UIADD3", "ULOP3", "USHF", "UMOV", "UISETP", "USEL", "ULEA", "UIMAD", "R2UR", "P2R", "R2P", "CS2R"):
        return "alu"
    return "other"


def main():
    rep, units = sys.argv[1], float(sys.argv[2])
    key = None
    rest = sys.argv[3:]
    if rest and rest[0] == "--json":
        key, rest = rest[1], rest[2:]
    rawcsv = None
    if rest and rest[0] == "--raw":
        rawcsv, rest = rest[1], rest[2:]
    which = int(rest[0]) if rest else 0
    if rep.endswith(".gz"):  # source page already exported on the GPU box (tools/ncu_round2.sh)
        import gzip

        raw = gzip.open(rep, "rt").read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    # one table per profiled launch, separated by blank lines / kernel headers
    tables, cur = [], []
    for line in raw.splitlines():
        if line.startswith('"Kernel Name"') or line.startswith("Kernel Name"):
            if cur:
                tables.append(cur)
            cur = [line]
        elif cur:
            cur.append(line)
    if cur:
        tables.append(cur)
    if not tables:
        tables = [raw.splitlines()]
    rows = list(csv.reader(io.StringIO("\n".join(tables[which]))))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and any(c.startswith("Instructions Executed") for c in r))
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    kname = rows[0][1] if len(rows[0]) > 1 else ""
    mix, byop = Counter(), Counter()
    tot = 0
    for r in rows[hi + 1:]:
        if len(r) <= ix["Source"] or not r[ix["Source"]].strip():
            continue
        toks = r[ix["Source"]].split()
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        try:
            n = int(r[ix["Instructions Executed"]])
        except ValueError:
            continue
        mix[classify(op)] += n
        byop[op] += n
        tot += n
    per = {k: v / units for k, v in mix.items()}
    out = {"kernel": kname, "report": rep, "units_per_launch": units, "warp_instructions": tot, "warp_instructions_per_unit": tot / units,
           "per_unit": per, "top_opcodes_per_unit": {k: v / units for k, v in byop.most_common(16)}}
    if rawcsv:
        rr = list(csv.reader(open(rawcsv)))
        h, row = rr[0], rr[2 + which]
        g = lambda name: float(row[h.index(name)].replace(",", "")) if name in h and row[h.index(name)] not in ("", "n/a") else None
        u = lambda name: rr[1][h.index(name)] if name in h else ""
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd = g("dram__bytes_read.sum") * scale.get(u("dram__bytes_read.sum"), 1.0)
        wr = g("dram__bytes_write.sum") * scale.get(u("dram__bytes_write.sum"), 1.0)
        dur = g("gpu__time_duration.sum") * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(u("gpu__time_duration.sum"), 1.0)
        out.update({"dram_bytes": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr, "duration_us_under_ncu": dur,
                    "pipe_busy_pct": {"fmaheavy_cycles_active": g("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                                      "fp64_cycles_active": g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                                      "alu_inst": g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                                      "issue_active": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                                      "dram_throughput": g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")},
                    "raw_source": rawcsv})
    print(json.dumps(out, indent=1))
    if key:
        path = "profiles/r02_opmix.json"
        try:
            with open(path) as f:
                allm = json.load(f)
        except Exception:
            allm = {}
        allm[key] = out
        with open(path, "w") as f:
            json.dump(allm, f, indent=1)


if __name__ == "__main__":
    main()
