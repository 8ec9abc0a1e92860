#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Builds oracle/_ref/libref_oracle.so from the
# reference's own C++ sources WHERE THEY LIE under $REF (default
# /root/reference) plus oracle/ref_harness.cpp (extern "C" wrappers), and the
# reference's own stand-alone test programs.  Nothing is copied into the repo:
# outputs go only to oracle/_ref/ (git-ignored, but shipped to the GPU box).
#
# Deviations from a verbatim compile (all mechanical):
#   P1  -I oracle/shim supplies an x86 <arm_neon.h> stand-in.
#   P2  -include oracle/shim/pre_fhe_types.h hides the duplicate
#       SecurityLevel/ParameterSet definitions of fhe_types.h.
#   P3  modular_arithmetic.cpp is streamed through sed into the compiler so
#       that the file-local mod_inverse (cpp/src/modular_arithmetic.cpp:17,20)
#       gets AArch64 divide-by-zero semantics (q/0 = 0, a%0 = a) instead of
#       SIGFPE on x86.  It only affects the single-limb Montgomery q_inv, which
#       the NTT / polynomial / bootstrap path never uses.
#   P6  key_serializer.cpp (wire formats) is compiled with `-include cstring`:
#       key_serializer.h uses std::memset without including <cstring>.
# encryption.cpp is NOT compiled (it does not compile as shipped);
# the tally and tensor-product wrappers compose unmodified PolynomialRing calls.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/cpp/src" ]; then
    echo "build_ref: $REF/cpp/src not present; keeping any prebuilt $OUT" >&2
    exit 0
fi
mkdir -p "$OUT"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT

CXX="${CXX:-g++}"
FLAGS=(-std=c++17 -O2 -fPIC -w -I"$HERE/shim" -I"$REF/cpp/include" -include "$HERE/shim/pre_fhe_types.h")

compile() { # src obj
    "$CXX" "${FLAGS[@]}" -c "$1" -o "$2"
}

# P3: patched stream, never written to disk.
P3_SED=(-e 's|int64_t q = a / m;|int64_t q = (m != 0) ? (int64_t)(a / m) : 0;|'
        -e 's|^\(\s*\)m = a % m;|\1m = (m != 0) ? (a % m) : a;|')
if [ "$(sed "${P3_SED[@]}" "$REF/cpp/src/modular_arithmetic.cpp" | grep -c '(m != 0)')" != "2" ]; then
    echo "build_ref: P3 pattern did not match exactly twice" >&2
    exit 1
fi
sed "${P3_SED[@]}" "$REF/cpp/src/modular_arithmetic.cpp" |
    "$CXX" "${FLAGS[@]}" -I"$REF/cpp/src" -x c++ -c - -o "$TMP/modular_arithmetic.o"

pids=()
for f in ntt_processor polynomial_ring parameter_set key_manager bootstrap_engine adaptive_dispatcher; do
    compile "$REF/cpp/src/$f.cpp" "$TMP/$f.o" &
    pids+=($!)
done
"$CXX" "${FLAGS[@]}" -include cstring -c "$REF/cpp/src/key_serializer.cpp" -o "$TMP/key_serializer.o" &
pids+=($!)
"$CXX" "${FLAGS[@]}" -include cstring -c "$HERE/ref_harness.cpp" -o "$TMP/ref_harness.o" &
pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done

"$CXX" -shared -o "$OUT/libref_oracle.so" "$TMP"/*.o -lpthread
echo "build_ref: built $OUT/libref_oracle.so"

# The reference's own stand-alone tests for this path (cpp/tests/README.md:29-37).
CORE=("$TMP/modular_arithmetic.o" "$TMP/ntt_processor.o" "$TMP/polynomial_ring.o")
pids=()
for t in test_ntt_processor test_multi_limb test_polynomial_ring test_neon_correctness; do
    if [ -f "$REF/cpp/tests/$t.cpp" ]; then
        "$CXX" "${FLAGS[@]}" -I"$REF/cpp/tests" "$REF/cpp/tests/$t.cpp" "${CORE[@]}" -o "$OUT/$t" -lpthread &
        pids+=($!)
    fi
done
for p in "${pids[@]}"; do wait "$p"; done
echo "build_ref: built reference test programs in $OUT"
