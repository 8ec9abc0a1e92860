"""Host-side mirror of the reference's C++ classes for the hot path, over the C ABI.

Class and method names follow the reference (cpp/include/ntt_processor.h:49-303,
polynomial_ring.h:312-513, modular_arithmetic.h:124-194, bootstrap_engine.h:176-508,
encryption.h:192 tally slice) so that the parity tests read like the reference's own.
Data are flat unsigned 64-bit words:

* numpy ``uint64`` arrays  -> HOST buffers (staged through the device inside the call);
* torch CUDA tensors of dtype int64/uint64 -> DEVICE buffers, used in place, asynchronous
  on torch's current stream.

torch is used only for device memory and streams.  All arithmetic runs in libfheb200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _cabi
from ._cabi import BootParams, DeviceInfo, FheError, WireHeader, check, lib

try:  # torch is plumbing only (device tensors / streams); numpy-only use works without it
    import torch
except Exception:  # pragma: no cover
    torch = None


# ------------------------------------------------------------------------------ buffers --
def _is_torch(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor)


def _ptr(x) -> int:
    if x is None:
        return None
    if _is_torch(x):
        if not x.is_contiguous():
            raise FheError(_cabi.INVALID_PARAMETERS, "tensor must be contiguous")
        if x.element_size() != 8:
            raise FheError(_cabi.INVALID_PARAMETERS, "tensor must hold 64-bit words (int64/uint64)")
        return x.data_ptr()
    if isinstance(x, np.ndarray):
        if x.dtype != np.uint64 or not x.flags["C_CONTIGUOUS"]:
            raise FheError(_cabi.INVALID_PARAMETERS, "array must be C-contiguous uint64")
        return x.ctypes.data
    raise FheError(_cabi.INVALID_PARAMETERS, f"unsupported buffer type {type(x)!r}")


def _like(x, shape=None):
    if _is_torch(x):
        return torch.empty(x.shape if shape is None else shape, dtype=x.dtype, device=x.device)
    return np.empty(x.shape if shape is None else shape, dtype=np.uint64)


def _raw(a):
    """pointer to a host ndarray of any dtype (byte buffers, status arrays)"""
    return None if a is None or a.size == 0 else a.ctypes.data_as(C.c_void_p)


def _stream(*xs) -> Optional[int]:
    for x in xs:
        if _is_torch(x) and x.is_cuda:
            return torch.cuda.current_stream(x.device).cuda_stream
    return None


def _words(x) -> int:
    return int(x.numel()) if _is_torch(x) else int(x.size)


def as_words(x):
    """numpy uint64 view/copy of host data; torch tensors pass through."""
    if _is_torch(x):
        return x
    return np.ascontiguousarray(x, dtype=np.uint64)


# ------------------------------------------------------------------------------ library --
def initialize(device: int = -1) -> None:
    """initialize(): src/native/lib.rs:23-30."""
    check(lib().fheb_init(device))


def version() -> str:
    return lib().fheb_version().decode()


def detect_hardware() -> dict:
    """detect_hardware(): src/native/lib.rs:32-42 (Apple fields truthfully false)."""
    info = DeviceInfo()
    check(lib().fheb_device_info_get(C.byref(info)))
    return dict(has_sme=bool(info.has_sme), has_metal=bool(info.has_metal), has_neon=bool(info.has_neon),
                has_amx=bool(info.has_amx), has_cuda=bool(info.has_cuda), compute_capability=(info.cc_major, info.cc_minor),
                sm_count=info.sm_count, device_memory_bytes=info.device_memory_bytes, l2_bytes=info.l2_bytes,
                smem_per_block_optin=info.smem_per_block_optin, name=info.name.decode())


def launch_count(reset: bool = False) -> int:
    return int(lib().fheb_launch_count(1 if reset else 0))


def synchronize() -> None:
    check(lib().fheb_synchronize(None))


# --------------------------------------------------------------------- ModularArithmetic --
class ModularArithmetic:
    """The scalar class of the reference's addon (src/native/lib.rs:44-120): host-side, one value per call,
    word-for-word the reference's arithmetic including its Montgomery-constant quirk."""

    def __init__(self, modulus: int):
        if modulus <= 0:
            raise FheError(_cabi.INVALID_PARAMETERS, "Modulus must be positive")  # lib.rs:53-55
        self._h = C.c_void_p()
        self._destroy = lib().fheb_modarith_destroy
        check(lib().fheb_modarith_create(modulus, C.byref(self._h)))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and self._destroy is not None:
            self._destroy(h)

    @staticmethod
    def _nonneg(*vals):
        if any(v < 0 for v in vals):
            raise FheError(_cabi.INVALID_PARAMETERS, "Inputs must be non-negative")  # lib.rs:64-66

    def montgomery_mul(self, a: int, b: int) -> int:
        self._nonneg(a, b)
        return int(lib().fheb_modarith_montgomery_mul(self._h, a, b))

    def mod_add(self, a: int, b: int) -> int:
        self._nonneg(a, b)
        return int(lib().fheb_modarith_mod_add(self._h, a, b))

    def mod_sub(self, a: int, b: int) -> int:
        self._nonneg(a, b)
        return int(lib().fheb_modarith_mod_sub(self._h, a, b))

    def to_montgomery(self, a: int) -> int:
        self._nonneg(a)
        return int(lib().fheb_modarith_to_montgomery(self._h, a))

    def from_montgomery(self, a: int) -> int:
        self._nonneg(a)
        return int(lib().fheb_modarith_from_montgomery(self._h, a))

    def get_modulus(self) -> int:
        return int(lib().fheb_modarith_get_modulus(self._h))


# ------------------------------------------------------------------------- NTTProcessor --
class NTTProcessor:
    """cpp/include/ntt_processor.h:49-303."""

    def __init__(self, degree: int, modulus: int, fwd_table=None, inv_table=None, inv_n: int = 0):
        self._h = C.c_void_p()
        self._destroy = lib().fheb_ntt_plan_destroy  # bound now: module globals are gone at interpreter exit
        if fwd_table is None:
            check(lib().fheb_ntt_plan_create(degree, modulus, C.byref(self._h)))
        else:  # caller-supplied tables: fast_ntt_forward / MetalComputeContext::batch_ntt_forward shape
            f, v = as_words(fwd_table), as_words(inv_table)
            check(lib().fheb_ntt_plan_create_with_tables(degree, modulus, _ptr(f), _ptr(v), inv_n, C.byref(self._h)))
        self.degree, self.modulus = degree, modulus

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and self._destroy is not None:
            self._destroy(h)

    def get_degree(self) -> int:
        return int(lib().fheb_ntt_plan_degree(self._h))

    def get_modulus(self) -> int:
        return int(lib().fheb_ntt_plan_modulus(self._h))

    def get_twiddles(self):
        """TwiddleFactors: (forward[N], inverse[N], primitive_root, inv_primitive_root, inv_n)."""
        f = np.empty(self.degree, np.uint64)
        v = np.empty(self.degree, np.uint64)
        s = np.empty(3, np.uint64)
        check(lib().fheb_ntt_plan_get_tables(self._h, _ptr(f), _ptr(v), _ptr(s)))
        return f, v, int(s[0]), int(s[1]), int(s[2])

    def _run(self, fn, coeffs, out):
        coeffs = as_words(coeffs)
        if coeffs.shape[-1] != self.degree:
            raise FheError(_cabi.INVALID_PARAMETERS, "Size must match polynomial degree")
        out = _like(coeffs) if out is None else out
        batch = _words(coeffs) // self.degree
        check(fn(self._h, _ptr(coeffs), _ptr(out), batch, _stream(coeffs, out)))
        return out

    def forward_ntt(self, coeffs, out=None):
        """forward_ntt / forward_ntt_batch (leading dimensions are the batch); out=coeffs is in place."""
        return self._run(lib().fheb_ntt_forward_batch, coeffs, out)

    def inverse_ntt(self, coeffs, out=None):
        return self._run(lib().fheb_ntt_inverse_batch, coeffs, out)

    def inverse_ntt_forward_network(self, coeffs, out=None):
        """fast_ntt_inverse semantics: cpp/src/adaptive_dispatcher.cpp:171-205."""
        return self._run(lib().fheb_ntt_inverse_fwdnet_batch, coeffs, out)

    forward_ntt_batch = forward_ntt
    inverse_ntt_batch = inverse_ntt


# ----------------------------------------------------------------------- PolynomialRing --
def _elementwise(fn, a, b, modulus, out):
    a, b = as_words(a), as_words(b)
    out = _like(a) if out is None else out
    check(fn(_ptr(a), _ptr(b), _ptr(out), _words(a), modulus, _stream(a, b, out)))
    return out


def modadd_batch(a, b, modulus, out=None):
    return _elementwise(lib().fheb_modadd_batch, a, b, modulus, out)


def modsub_batch(a, b, modulus, out=None):
    return _elementwise(lib().fheb_modsub_batch, a, b, modulus, out)


def modmul_batch(a, b, modulus, out=None):
    """fast_modmul_batch / MetalComputeContext::batch_modmul."""
    return _elementwise(lib().fheb_modmul_batch, a, b, modulus, out)


def modneg_batch(a, modulus, out=None):
    a = as_words(a)
    out = _like(a) if out is None else out
    check(lib().fheb_modneg_batch(_ptr(a), _ptr(out), _words(a), modulus, _stream(a, out)))
    return out


def modmul_scalar_batch(a, scalar, modulus, out=None):
    a = as_words(a)
    out = _like(a) if out is None else out
    check(lib().fheb_modmul_scalar_batch(_ptr(a), scalar, _ptr(out), _words(a), modulus, _stream(a, out)))
    return out


class PolynomialRing:
    """cpp/include/polynomial_ring.h:312-513; polynomials are [..., N] word arrays."""

    def __init__(self, degree: int, modulus: int):
        self.ntt = NTTProcessor(degree, modulus)
        self.degree, self.modulus = degree, modulus

    def add(self, a, b, out=None):
        return modadd_batch(a, b, self.modulus, out)

    def subtract(self, a, b, out=None):
        return modsub_batch(a, b, self.modulus, out)

    def negate(self, a, out=None):
        return modneg_batch(a, self.modulus, out)

    def multiply_scalar(self, a, scalar, out=None):
        return modmul_scalar_batch(a, scalar, self.modulus, out)

    def pointwise_multiply(self, a, b, out=None):
        return modmul_batch(a, b, self.modulus, out)

    def to_ntt(self, a, out=None):
        return self.ntt.forward_ntt(a, out)

    def from_ntt(self, a, out=None):
        return self.ntt.inverse_ntt(a, out)

    def multiply(self, a, b, out=None):
        """PolynomialRing::multiply on coefficient-form operands (polynomial_ring.cpp:421-447)."""
        a, b = as_words(a), as_words(b)
        if a.shape[-1] != self.degree or _words(a) != _words(b):
            raise FheError(_cabi.INVALID_PARAMETERS, "Polynomial degree mismatch")
        out = _like(a) if out is None else out
        check(lib().fheb_polymul_batch(self.ntt._h, _ptr(a), _ptr(b), _ptr(out), _words(a) // self.degree,
                                       _stream(a, b, out)))
        return out

    def tensor_multiply(self, ct1, ct2, out=None):
        """EncryptionEngine::multiply tensor product: [batch][2][N] x [batch][2][N] -> [batch][3][N]."""
        ct1, ct2 = as_words(ct1), as_words(ct2)
        batch = _words(ct1) // (2 * self.degree)
        out = _like(ct1, tuple(ct1.shape[:-2]) + (3, self.degree)) if out is None else out
        check(lib().fheb_tensor_multiply_batch(self.ntt._h, _ptr(ct1), _ptr(ct2), _ptr(out), batch, _stream(ct1, ct2, out)))
        return out


class RnsPolynomialRing:
    """PolynomialRing(degree, moduli) (cpp/src/polynomial_ring.cpp:224-237) with a real residue-number-system meaning.

    The reference builds one NTTProcessor per modulus but every method then computes with moduli[0] only
    (polynomial_ring.cpp:245-539), so its modulus chains (bfv-128-simd, ckks-128-ml: parameter_set.cpp:193-259) are
    inert.  Here every limb is live: data are limb-major `[limbs][batch][N]`, limb l is an ordinary batch over
    moduli[l] run by that modulus's plan (same kernels, one launch per limb and method).  Limb 0 of every method is
    exactly the reference's result; the other limbs are the same reference method over their own modulus."""

    def __init__(self, degree: int, moduli: Sequence[int]):
        if len(moduli) == 0:
            raise FheError(_cabi.INVALID_PARAMETERS, "At least one modulus required")  # polynomial_ring.cpp:228-230
        self.degree, self.moduli = degree, [int(m) for m in moduli]
        self.rings = [PolynomialRing(degree, m) for m in self.moduli]

    def _each(self, name, *operands, out=None):
        ops = [as_words(o) for o in operands]
        limbs = len(self.rings)
        if any(o.shape[0] != limbs or o.shape[-1] != self.degree for o in ops):
            raise FheError(_cabi.INVALID_PARAMETERS, "operands must be [limbs][batch][N] with one limb per modulus")
        out = _like(ops[0]) if out is None else out
        for l, ring in enumerate(self.rings):
            getattr(ring, name)(*[o[l] for o in ops], out=out[l])
        return out

    def add(self, a, b, out=None):
        return self._each("add", a, b, out=out)

    def subtract(self, a, b, out=None):
        return self._each("subtract", a, b, out=out)

    def negate(self, a, out=None):
        return self._each("negate", a, out=out)

    def pointwise_multiply(self, a, b, out=None):
        return self._each("pointwise_multiply", a, b, out=out)

    def to_ntt(self, a, out=None):
        return self._each("to_ntt", a, out=out)

    def from_ntt(self, a, out=None):
        return self._each("from_ntt", a, out=out)

    def multiply(self, a, b, out=None):
        return self._each("multiply", a, b, out=out)


class RelinearizationKey:
    """EvaluationKey.relin_key (cpp/include/key_manager.h:92-111) resident on the device, both polynomials of every
    level pre-transformed.  keys = [key_count][2][N]: (a, b) of every pair in generation order."""

    def __init__(self, ring: "PolynomialRing", keys, decomp_base_log: int = 0, decomp_level: int = 0, key_id: int = 0):
        self.ring, self.key_id = ring, key_id
        keys = np.ascontiguousarray(as_words(keys)) if not _is_torch(keys) else as_words(keys)
        count = _words(keys) // (2 * ring.degree)
        h = C.c_void_p()
        check(lib().fheb_relin_key_create(ring.ntt._h, _ptr(keys) if count else None, count, decomp_base_log, decomp_level,
                                          key_id, C.byref(h)))
        self._h = h
        self.levels = int(lib().fheb_relin_key_levels(h))

    @classmethod
    def from_wire(cls, ring: "PolynomialRing", data: bytes):
        """KeySerializer::deserialize_eval_key (key_serializer.cpp:414-466) straight into a device key."""
        self = cls.__new__(cls)
        self.ring = ring
        buf = np.frombuffer(bytes(data), dtype=np.uint8)
        h = C.c_void_p()
        check(lib().fheb_relin_key_from_wire(ring.ntt._h, _raw(buf), buf.size, C.byref(h)))
        self._h = h
        self.key_id = wire_header(data).key_id
        self.levels = int(lib().fheb_relin_key_levels(h))
        return self

    def relinearize(self, cts, ct_key_id: Optional[int] = None, out=None):
        """EncryptionEngine::relinearize (cpp/src/encryption.cpp:904-993): [batch][3][N] -> [batch][2][N]."""
        cts = as_words(cts)
        n = self.ring.degree
        batch = _words(cts) // (3 * n)
        out = _like(cts, tuple(cts.shape[:-2]) + (2, n)) if out is None else out
        check(lib().fheb_relinearize_batch(self._h, _ptr(cts), self.key_id if ct_key_id is None else ct_key_id, _ptr(out),
                                           batch, _stream(cts, out)))
        return out

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                lib().fheb_relin_key_destroy(h)
            except Exception:
                pass


# ----------------------------------------------------------- MultiLimbModularArithmetic --
class MultiLimbModularArithmetic:
    """cpp/include/modular_arithmetic.h:124-194; integers are [count][limbs] little-endian words."""

    def __init__(self, q_limbs: Sequence[int]):
        self.q = np.ascontiguousarray(q_limbs, dtype=np.uint64)
        self.limbs = int(self.q.size)
        consts = np.zeros(1 + 2 * self.limbs, np.uint64)
        check(lib().fheb_mlimb_constants(_ptr(self.q), self.limbs, _ptr(consts)))
        self.q_inv = int(consts[0])
        self.r_mod_q = consts[1:1 + self.limbs].copy()
        self.r2_mod_q = consts[1 + self.limbs:].copy()

    def _bin(self, which, a, b, out):
        a, b = as_words(a), as_words(b)
        out = _like(a) if out is None else out
        count = _words(a) // self.limbs
        L = lib()
        if which == 0:
            check(L.fheb_mlimb_montmul_batch(_ptr(a), _ptr(b), _ptr(out), count, self.limbs, _ptr(self.q), self.q_inv,
                                             _stream(a, b, out)))
        elif which == 1:
            check(L.fheb_mlimb_add_batch(_ptr(a), _ptr(b), _ptr(out), count, self.limbs, _ptr(self.q), _stream(a, b, out)))
        else:
            check(L.fheb_mlimb_sub_batch(_ptr(a), _ptr(b), _ptr(out), count, self.limbs, _ptr(self.q), _stream(a, b, out)))
        return out

    def montgomery_mul(self, a, b, out=None):
        return self._bin(0, a, b, out)

    def mod_add(self, a, b, out=None):
        return self._bin(1, a, b, out)

    def mod_sub(self, a, b, out=None):
        return self._bin(2, a, b, out)

    def _const_like(self, a, row):
        a = as_words(a)
        count = _words(a) // self.limbs
        tiled = np.ascontiguousarray(np.broadcast_to(row, (count, self.limbs)))
        if _is_torch(a):
            return torch.from_numpy(tiled.view(np.int64)).to(a.device).view(a.dtype)
        return tiled

    def to_montgomery(self, a, out=None):
        """montgomery_mul(a, R^2 mod q): modular_arithmetic.cpp:673-683."""
        return self._bin(0, a, self._const_like(a, self.r2_mod_q), out)

    def from_montgomery(self, a, out=None):
        """montgomery_mul(a, 1): modular_arithmetic.cpp:685-693."""
        one = np.zeros(self.limbs, np.uint64)
        one[0] = 1
        return self._bin(0, a, self._const_like(a, one), out)

    montgomery_mul_neon, mod_add_neon, mod_sub_neon = montgomery_mul, mod_add, mod_sub


# ---------------------------------------------------------------------- BootstrapEngine --
class BootstrapEngine:
    """Deterministic part of cpp/include/bootstrap_engine.h:176-508 (key generation and
    encryption stay with the reference: they are RNG-bound host code)."""

    def __init__(self, poly_degree: int, modulus: int, lwe_dimension: int, glwe_dimension: int, decomp_base_log: int,
                 decomp_level: int, bsk, plaintext_modulus: int = 4):
        self.N, self.q, self.n, self.k = poly_degree, modulus, lwe_dimension, glwe_dimension
        self.base_log, self.level, self.t = decomp_base_log, decomp_level, plaintext_modulus
        self.ntt = NTTProcessor(poly_degree, modulus)
        self._h = C.c_void_p()
        self._destroy = lib().fheb_boot_key_destroy
        bsk = as_words(bsk)
        rows = (self.k + 1) * self.level
        if _words(bsk) != self.n * rows * (self.k + 1) * self.N:
            raise FheError(_cabi.INVALID_PARAMETERS, "bootstrap key has the wrong number of words")
        params = BootParams(self.n, self.k, self.base_log, self.level)
        check(lib().fheb_boot_key_create(self.ntt._h, C.byref(params), _ptr(bsk), C.byref(self._h)))
        self.n_out = None

    @classmethod
    def from_wire(cls, data: bytes, lwe_dimension: int, decomp_base_log: int, decomp_level: int, plaintext_modulus: int = 4):
        """KeySerializer::deserialize_bootstrap_key (key_serializer.cpp:545-615) straight into a device key
        (glwe_dimension 1; degree and modulus come from the container's header)."""
        hdr = wire_header(data)
        self = cls.__new__(cls)
        self.N, self.q, self.n, self.k = hdr.poly_degree, hdr.modulus, lwe_dimension, 1
        self.base_log, self.level, self.t = decomp_base_log, decomp_level, plaintext_modulus
        self.ntt = NTTProcessor(self.N, self.q)
        self._h = C.c_void_p()
        self._destroy = lib().fheb_boot_key_destroy
        buf = np.frombuffer(bytes(data), dtype=np.uint8)
        params = BootParams(self.n, self.k, self.base_log, self.level)
        check(lib().fheb_boot_key_from_wire(self.ntt._h, C.byref(params), _raw(buf), buf.size, C.byref(self._h)))
        self.n_out = None
        return self

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and self._destroy is not None:
            self._destroy(h)

    def set_key_switch_key(self, ksk, n_out: int, base_log: int, level: int):
        ksk = as_words(ksk)
        entries = _words(ksk) // (n_out + 1)
        check(lib().fheb_boot_key_set_ksk(self._h, _ptr(ksk), entries, n_out, base_log, level))
        self.n_out = n_out

    def _glwe_words(self):
        return (self.k + 1) * self.N

    def external_product(self, glwe, index: int, out=None):
        glwe = as_words(glwe)
        out = _like(glwe) if out is None else out
        check(lib().fheb_external_product_batch(self._h, index, _ptr(glwe), _ptr(out), _words(glwe) // self._glwe_words(),
                                                _stream(glwe, out)))
        return out

    def cmux(self, index: int, ct0, ct1, out=None):
        ct0, ct1 = as_words(ct0), as_words(ct1)
        out = _like(ct0) if out is None else out
        check(lib().fheb_cmux_batch(self._h, index, _ptr(ct0), _ptr(ct1), _ptr(out), _words(ct0) // self._glwe_words(),
                                    _stream(ct0, ct1, out)))
        return out

    def blind_rotate(self, lwe, test_poly, out=None):
        """blind_rotate of acc = (0,..,0,test_poly): lwe [batch][n+1] -> [batch][k+1][N]."""
        lwe, test_poly = as_words(lwe), as_words(test_poly)
        batch = _words(lwe) // (self.n + 1)
        out = _like(lwe, (batch, self.k + 1, self.N)) if out is None else out
        check(lib().fheb_blind_rotate_batch(self._h, _ptr(lwe), _ptr(test_poly), _ptr(out), batch, _stream(lwe, out)))
        return out

    def sample_extract(self, glwe, out=None):
        glwe = as_words(glwe)
        batch = _words(glwe) // self._glwe_words()
        out = _like(glwe, (batch, self.k * self.N + 1)) if out is None else out
        check(lib().fheb_sample_extract_batch(self._h, _ptr(glwe), _ptr(out), batch, _stream(glwe, out)))
        return out

    def key_switch(self, lwe, out=None):
        if self.n_out is None:
            raise FheError(_cabi.INVALID_PARAMETERS, "no key switching key has been set (set_key_switch_key)")
        lwe = as_words(lwe)
        batch = _words(lwe) // (self.k * self.N + 1)
        out = _like(lwe, (batch, self.n_out + 1)) if out is None else out
        check(lib().fheb_key_switch_batch(self._h, _ptr(lwe), _ptr(out), batch, _stream(lwe, out)))
        return out

    def bootstrap_with_test_poly(self, lwe, test_poly, out=None):
        lwe, test_poly = as_words(lwe), as_words(test_poly)
        batch = _words(lwe) // (self.n + 1)
        width = (self.n_out + 1) if self.n_out is not None else (self.k * self.N + 1)
        out = _like(lwe, (batch, width)) if out is None else out
        check(lib().fheb_bootstrap_batch(self._h, _ptr(lwe), _ptr(test_poly), _ptr(out), batch, _stream(lwe, out)))
        return out

    bootstrap = bootstrap_with_test_poly

    def _lut(self, kind, arg0, arg1=0):
        out = np.empty(self.N, np.uint64)
        check(lib().fheb_make_test_poly(self.ntt._h, kind, arg0, arg1, _ptr(out)))
        return out

    def get_default_test_poly(self):
        return self._lut(3, self.t)

    def create_identity_lut(self, modulus):
        return self._lut(0, modulus)

    def create_negation_lut(self, modulus):
        return self._lut(1, modulus)

    def create_threshold_lut(self, threshold, modulus):
        return self._lut(2, threshold, modulus)


# -------------------------------------------------------------------------------- tally --
def tally_votes(cts, degree: int, modulus: int, out=None):
    """EncryptionEngine::batch_add / batch_add_tree / tally_votes words: [count][2][N] -> [2][N]."""
    cts = as_words(cts)
    count = _words(cts) // (2 * degree)
    if count == 0:
        raise FheError(_cabi.INVALID_PARAMETERS, "Cannot add empty vector of ciphertexts")
    out = _like(cts, (2, degree)) if out is None else out
    check(lib().fheb_tally(_ptr(cts), count, degree, modulus, _ptr(out), _stream(cts, out)))
    return out


batch_add = tally_votes


# ------------------------------------------------------------------------- wire formats --
WIRE_OK, WIRE_TOO_SMALL, WIRE_BAD_MAGIC, WIRE_BAD_CHECKSUM, WIRE_SHAPE_MISMATCH = range(5)


def wire_crc32(data: bytes) -> int:
    """KeySerializer::compute_crc32 with the reference's table as shipped (key_serializer.cpp:21-40)."""
    buf = np.frombuffer(bytes(data), dtype=np.uint8)
    return int(lib().fheb_wire_crc32(_raw(buf), buf.size))


def wire_header(data: bytes) -> WireHeader:
    buf = np.frombuffer(bytes(data[:49]), dtype=np.uint8)
    h = WireHeader()
    check(lib().fheb_wire_header_read(_raw(buf), buf.size, C.byref(h)))
    return h


def serialize_ballot(choices, modulus: int, timestamp: int) -> bytes:
    """BallotSerializer::serialize_ballot (key_serializer.cpp:709-774): choices = [num_choices][2][N] host words."""
    choices = np.ascontiguousarray(as_words(choices))
    num, _, n = choices.shape if choices.size else (0, 2, 0)
    size = int(lib().fheb_ballot_wire_size(num, n))
    out = np.zeros(size, np.uint8)
    written = C.c_size_t()
    check(lib().fheb_ballot_serialize(_ptr(choices) if choices.size else None, num, n, modulus, timestamp, _raw(out), size,
                                      C.byref(written)))
    return out[:written.value].tobytes()


def ingest_ballots(wire, count: int, num_choices: int, degree: int, modulus: int, offsets=None, out=None, device=None):
    """A loop of BallotSerializer::deserialize_ballot (key_serializer.cpp:776-846) over `count` FHEV records, validated
    and unpacked on the device.  wire: bytes / uint8 ndarray (host) or a torch uint8 CUDA tensor.  Returns
    (cts [count][num_choices][2][N], status [count] uint8, timestamps [count] uint64); cts is a CUDA tensor when
    `device` is given (or wire / out are CUDA tensors), else a host array.  Rejected records are zero ciphertexts."""
    if isinstance(wire, (bytes, bytearray, memoryview)):
        wire = np.frombuffer(bytes(wire), dtype=np.uint8)
    nbytes = int(wire.numel()) if _is_torch(wire) else int(wire.size)
    if out is None:
        shape = (count, num_choices, 2, degree)
        if _is_torch(wire) or device is not None:
            out = torch.empty(shape, dtype=torch.int64, device=wire.device if _is_torch(wire) else device)
        else:
            out = np.empty(shape, np.uint64)
    status = np.zeros(count, np.uint8)
    stamps = np.zeros(count, np.uint64)
    offs = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.uint64)
    accepted = C.c_size_t()
    wptr = C.c_void_p(wire.data_ptr()) if _is_torch(wire) else _raw(np.ascontiguousarray(wire))
    check(lib().fheb_ballots_ingest(wptr, nbytes, None if offs is None else _ptr(offs), count, num_choices, degree, modulus,
                                    _ptr(out), _raw(status), _ptr(stamps), C.byref(accepted), _stream(out)))
    return out, status, stamps


def tally_combine(partials, degree: int, modulus: int, out=None):
    partials = as_words(partials)
    parts = _words(partials) // (2 * degree)
    out = _like(partials, (2, degree)) if out is None else out
    check(lib().fheb_tally_combine(_ptr(partials), parts, degree, modulus, _ptr(out), _stream(partials, out)))
    return out


def synth_ballots(cts_device, first_ballot: int, count: int, degree: int, modulus: int, seed: int):
    check(lib().fheb_synth_ballots(_ptr(cts_device), first_ballot, count, degree, modulus, seed, _stream(cts_device)))
    return cts_device


class CiphertextStreamAccumulator:
    """Running tally kept on the device: the accumulator of CiphertextStreamProcessor::stream_add
    (cpp/src/streaming_processor.cpp:460-526).  add() folds a chunk [count][2][N] (host or device);
    total() returns the running total [2][N]."""

    def __init__(self, degree: int, modulus: int):
        self.degree, self.modulus = degree, modulus
        self._h = C.c_void_p()
        self._destroy = lib().fheb_tally_stream_destroy
        check(lib().fheb_tally_stream_create(degree, modulus, C.byref(self._h)))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and self._destroy is not None:
            self._destroy(h)

    def add(self, cts):
        cts = as_words(cts)
        count = _words(cts) // (2 * self.degree)
        check(lib().fheb_tally_stream_add(self._h, _ptr(cts), count, _stream(cts)))
        return self

    @property
    def count(self) -> int:
        return int(lib().fheb_tally_stream_count(self._h))

    def total(self, out=None):
        out = np.empty((2, self.degree), np.uint64) if out is None else out
        check(lib().fheb_tally_stream_total(self._h, _ptr(out), _stream(out)))
        return out


def tally_noise_budget(budgets: Sequence[float], variant: str = "linear") -> float:
    """noise_budget metadata of the reference's tally variants (host-side) through fheb_tally_noise_budget:
    "linear" = batch_add: min - log2(count) (encryption.cpp:1359-1360); "tree" = batch_add_tree / tally_votes: min(pair) - 1
    per level with the odd element carried (:1413,1437); "add" = a left fold with EncryptionEngine::add (:613)."""
    b = np.ascontiguousarray(budgets, dtype=np.float64)
    out = C.c_double(0.0)
    check(lib().fheb_tally_noise_budget(b.ctypes.data_as(C.c_void_p), int(b.size), {"linear": 0, "tree": 1, "add": 2}[variant], C.byref(out)))
    return float(out.value)


def pinned_empty(shape) -> np.ndarray:
    """uint64 host array in page-locked memory from the library's own allocator (fheb_host_alloc): placed on the NUMA node
    of the current GPU, usable from every device.  Freed when the array (and every view of it) is garbage collected."""
    import weakref

    shape = tuple(int(x) for x in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    nbytes = int(np.prod(shape)) * 8
    ptr = C.c_void_p()
    check(lib().fheb_host_alloc(C.byref(ptr), max(nbytes, 8)))
    buf = (C.c_uint64 * (max(nbytes, 8) // 8)).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=np.uint64, count=int(np.prod(shape))).reshape(shape)
    weakref.finalize(buf, lib().fheb_host_free, C.c_void_p(ptr.value))
    return arr


def set_devices(devices=None) -> int:
    """One process, several GPUs: host-buffer batches are split over `devices` (None = every visible GPU, [] = off).
    Returns the number of devices configured."""
    if devices is None:
        check(lib().fheb_set_devices(None, -1))
    else:
        arr = (C.c_int * max(len(devices), 1))(*devices)
        check(lib().fheb_set_devices(arr, len(devices)))
    return int(lib().fheb_get_devices(None, 0))


def tally_wire(wire, count: int, num_choices: int, degree: int, modulus: int, offsets=None, out=None, device=None):
    """Validate `count` FHEV records and tally every choice of the accepted ones straight from the wire bytes
    (deserialize_ballot per record + tally_votes per choice).  Returns (tallies [num_choices][2][N], status [count])."""
    if isinstance(wire, (bytes, bytearray, memoryview)):
        wire = np.frombuffer(bytes(wire), dtype=np.uint8)
    nbytes = int(wire.numel()) if _is_torch(wire) else int(wire.size)
    if out is None:
        shape = (num_choices, 2, degree)
        if _is_torch(wire) or device is not None:
            out = torch.empty(shape, dtype=torch.int64, device=wire.device if _is_torch(wire) else device)
        else:
            out = np.empty(shape, np.uint64)
    status = np.zeros(max(count, 1), np.uint8)
    offs = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.uint64)
    accepted = C.c_size_t()
    wptr = C.c_void_p(wire.data_ptr()) if _is_torch(wire) else _raw(np.ascontiguousarray(wire))
    check(lib().fheb_tally_wire(wptr, nbytes, None if offs is None else _ptr(offs), count, num_choices, degree, modulus, _ptr(out),
                                _raw(status), C.byref(accepted), _stream(out)))
    return out, status[:count]
