#!/bin/bash
# ncu captures of round 2 (run under gpurun on ONE GPU): each program first runs plain (must exit 0), then under ncu.
# gpurun brings back at most 64 MiB, and a --set full report with sources is ~20 MB: the raw page and the SASS source page
# of every report are exported to CSV on the box (gzip) and the report itself is dropped; tools/ncu_summary.py and
# tools/ncu_opmix.py read those CSVs here.
set -u
O=gpurun_out
mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on"
run() {  # name, kernel regex, skip, count, command...
    local name=$1 rx=$2 skip=$3 cnt=$4
    shift 4
    "$@" > $O/${name}_plain.log 2>&1 && $NCU -k regex:$rx -s $skip -c $cnt -f -o /tmp/$name "$@" > $O/${name}_ncu.log 2>&1
    local rc=$?
    if [ -f /tmp/$name.ncu-rep ]; then
        ncu -i /tmp/$name.ncu-rep --page raw --csv > $O/${name}_raw.csv 2>/dev/null
        ncu -i /tmp/$name.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip -9 > $O/${name}_source.csv.gz
        rm -f /tmp/$name.ncu-rep
    fi
    echo "$name rc=$rc"
}
if [ "${1:-all}" = "all" ]; then
# prof_ntt.py: 3 warm-ups + 20 timed launches of forward, then inverse, then polymul: profile one warmed-up launch of each
run r02_ntt14_q62_fwd 'ntt_forward_kernel' 5 1 python tools/prof_ntt.py 14
run r02_ntt14_q62_inv 'ntt_inverse_kernel' 5 1 python tools/prof_ntt.py 14
run r02_ntt14_q62_mul 'polymul_kernel' 5 1 python tools/prof_ntt.py 14
run r02_ntt14_q27_fwd 'ntt_forward_kernel' 5 1 env FHEB_NO_U32=1 python tools/prof_ntt.py 14 132120577
run r02_ntt10_qt_fwd 'ntt_forward_kernel' 5 1 python tools/prof_ntt.py 10 1099511678977 16384
run r02_boot_lean 'boot_kernel' 3 1 python tools/prof_boot.py 740
run r02_relin_n4096_l4 'relin_fused_kernel' 5 1 python tools/prof_relin.py
run r02_tensor_n4096 'tensor_fused_kernel' 5 1 python tools/prof_tensor.py
# the 32-bit kernels (4-byte slots) and the pair-layout fused product, q = 132120577
run r02_ntt14_q27_u32_fwd 'ntt_forward_kernel' 5 1 python tools/prof_ntt.py 14 132120577
run r02_ntt14_q27_u32_inv 'ntt_inverse_kernel' 5 1 python tools/prof_ntt.py 14 132120577
run r02_ntt14_q27_u32p_mul 'polymul_kernel' 5 1 python tools/prof_ntt.py 14 132120577
run r02_ntt10_q27_u32_fwd 'ntt_forward_kernel' 5 1 python tools/prof_ntt.py 10 132120577 16384
elif [ "$1" = "rest" ]; then
run r02_tally_128k 'tally_kernel' 4 1 python tools/prof_tally.py --one 131072
run r02_pipes 'k' 1 9 tools/microbench/pipes
else
    "$@"
fi
du -sh $O
