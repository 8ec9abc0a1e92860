#!/bin/bash
# profiles/r02_opmix.json from the exported ncu pages under profiles/r02_ncu/ (tools/ncu_round2.sh made them on the GPU box).
# Units per launch are WARP units (work units / 32): butterflies of the launch for the transforms (batch x N/2 x log2 N,
# three transforms for the fused product, levels + 2 / 7 transforms per ciphertext for relinearisation / the tensor
# product), ciphertext-steps for the blind rotation (740 x 742), ballots for the tally.
set -e
cd "$(dirname "$0")/.."
D=profiles/r02_ncu
mix() { python tools/ncu_opmix.py $D/$2_source.csv.gz $3 --json $1 --raw $D/$2_raw.csv > /dev/null; }
mix ntt_forward_n16384_q62      r02_ntt14_q62_fwd      3670016
mix ntt_inverse_n16384_q62      r02_ntt14_q62_inv      3670016
mix polymul_n16384_q62          r02_ntt14_q62_mul      11010048
mix ntt_forward_n16384_q27_fp64 r02_ntt14_q27_fwd      3670016
mix ntt_forward_n1024_qt_fp64   r02_ntt10_qt_fwd       2621440
mix boot_lean_tfhe128_step      r02_boot_lean          17158.75
mix relin_fused_n4096_l4        r02_relin_n4096_l4     9437184
mix tensor_fused_n4096          r02_tensor_n4096       11010048
mix tally_131072_ballot         r02_tally_128k         4096
mix ntt_forward_n16384_q27_u32  r02_ntt14_q27_u32_fwd  3670016
mix ntt_inverse_n16384_q27_u32  r02_ntt14_q27_u32_inv  3670016
mix polymul_n16384_q27_u32pair  r02_ntt14_q27_u32p_mul 11010048
mix ntt_forward_n1024_q27_u32   r02_ntt10_q27_u32_fwd  2621440
python - <<'PY'
import json
d = json.load(open("profiles/r02_opmix.json"))
for k, v in d.items():
    print(f"{k:32s} {v['warp_instructions_per_unit']:7.2f} instr/unit  " + " ".join(f"{c}={x:.2f}" for c, x in sorted(v["per_unit"].items())))
PY
