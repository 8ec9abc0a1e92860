"""Exhaustive check that the shared-memory swizzle of csrc/ntt_core.cuh (index bit j -> alpha^j in
GF(16)) is bank-conflict-free for every pass geometry of every plan: each group of 16 consecutive
lanes must hit 16 distinct 8-byte word-banks.  Also checks the popcount masks used on the device."""
m = [1]
for i in range(1, 16):
    v = m[-1] << 1
    if v & 16:
        v ^= 0b10011
    m.append(v)
MASKS = [0xF59, 0x1EB, 0x3D6, 0x7AC]
for k in range(4):
    assert MASKS[k] == sum(1 << (j - 4) for j in range(4, 16) if (m[j] >> k) & 1), k


def swz(i):
    h = i >> 4
    return i ^ sum((bin(h & MASKS[k]).count("1") & 1) << k for k in range(4))


PLANS = {2: [2], 3: [3], 4: [4], 5: [3, 2], 6: [3, 3], 7: [4, 3], 8: [4, 4], 9: [3, 3, 3], 10: [4, 3, 3], 11: [4, 4, 3],
         12: [4, 4, 4], 13: [4, 3, 3, 3], 14: [4, 4, 3, 3]}


def bitrev(x, b):
    return int(format(x, "0%db" % b)[::-1], 2) if b else 0


bad = 0
for L, R in PLANS.items():
    N, s0 = 1 << L, 0
    for p, r in enumerate(R):
        EB, items = L - s0 - r, N >> r
        for mode in ("natural", "bitrev"):
            if mode == "bitrev" and p != len(R) - 1:
                continue
            for c in range(1 << r):
                for w0 in range(0, items, 16):
                    banks = []
                    for t in range(w0, min(w0 + 16, items)):
                        u = bitrev(t, L - r) if mode == "bitrev" else t
                        pos = ((u >> EB) << (EB + r)) | (u & ((1 << EB) - 1)) | (c << EB)
                        banks.append(swz(pos) & 15)
                    bad += len(set(banks)) != len(banks)
        s0 += r
print("conflicting half-warp accesses:", bad)
assert bad == 0

# ---- 4-byte slots (MODE_U32): one wavefront of 32 lanes over 32 banks; index bit j -> beta^j in GF(32), x^5 + x^3 + 1
m5 = [1]
for i in range(1, 31):
    v = m5[-1] << 1
    if v & 32:
        v ^= 0b101001
    m5.append(v)
MASKS5 = [0x375, 0x6EA, 0x5D4, 0x0DD, 0x1BA]
for k in range(5):
    assert MASKS5[k] == sum(1 << (j - 5) for j in range(5, 16) if (m5[j] >> k) & 1), k


def swz32(i):
    h = i >> 5
    return i ^ sum((bin(h & MASKS5[k]).count("1") & 1) << k for k in range(5))


bad = 0
for L, R in PLANS.items():
    if len(R) < 2:
        continue
    N, s0 = 1 << L, 0
    for p, r in enumerate(R):
        EB, items = L - s0 - r, N >> r
        for mode in ("natural", "bitrev"):
            if mode == "bitrev" and p != len(R) - 1:
                continue
            for c in range(1 << r):
                for w0 in range(0, items, 32):
                    banks = []
                    for t in range(w0, min(w0 + 32, items)):
                        u = bitrev(t, L - r) if mode == "bitrev" else t
                        pos = ((u >> EB) << (EB + r)) | (u & ((1 << EB) - 1)) | (c << EB)
                        banks.append(swz32(pos) & 31)
                    bad += len(set(banks)) != len(banks)
        s0 += r
print("conflicting warp accesses with 4-byte slots:", bad)
assert bad == 0
