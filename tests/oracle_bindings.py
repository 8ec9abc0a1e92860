"""ctypes bindings for the two CPU checkers (TEST INFRASTRUCTURE ONLY).

* ``Oracle``   - oracle/libfhe_oracle.so, the plain-C restatement (oracle/fhe_oracle.c).
* ``RefOracle``- oracle/_ref/libref_oracle.so, the reference's OWN C++ compiled here by
  oracle/build_ref.sh (absent on machines that never had /root/reference and did not
  receive the prebuilt file).

Nothing under node-fhe-accelerate_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libfhe_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libref_oracle.so")

_T = {
    "u32": C.c_uint32,
    "u64": C.c_uint64,
    "i32": C.c_int32,
    "int": C.c_int,
    "sz": C.c_size_t,
    "p": C.c_void_p,
}


def _ptr(a):
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"], "need a contiguous ndarray"
    return a.ctypes.data_as(C.c_void_p)


def u64(x):
    return np.ascontiguousarray(x, dtype=np.uint64)


def _bind(lib, name, sig, restype=C.c_int):
    fn = getattr(lib, name)
    fn.argtypes = [_T[s] for s in sig.split()] if sig else []
    fn.restype = restype
    return fn


def build_oracle():
    """Compile the C oracle if its shared object is missing or stale."""
    src = os.path.join(ORACLE_DIR, "fhe_oracle.c")
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "libfhe_oracle.so"], stdout=subprocess.DEVNULL)
    return ORACLE_SO


class BootParams(C.Structure):
    _fields_ = [
        ("N", C.c_uint32),
        ("k", C.c_uint32),
        ("n", C.c_uint32),
        ("base_log", C.c_uint32),
        ("level", C.c_uint32),
        ("q", C.c_uint64),
        ("t", C.c_uint64),
        ("fwd_table", C.c_void_p),
        ("inv_table", C.c_void_p),
        ("inv_n", C.c_uint64),
    ]


class Oracle:
    """numpy-facing wrapper over oracle/fhe_oracle.h."""

    def __init__(self):
        self.lib = C.CDLL(build_oracle())
        L = self.lib
        self._f = {
            "mod_add": _bind(L, "orc_mod_add", "u64 u64 u64", C.c_uint64),
            "mod_sub": _bind(L, "orc_mod_sub", "u64 u64 u64", C.c_uint64),
            "mod_mul": _bind(L, "orc_mod_mul", "u64 u64 u64", C.c_uint64),
            "mod_pow": _bind(L, "orc_mod_pow", "u64 u64 u64", C.c_uint64),
            "mod_inverse": _bind(L, "orc_mod_inverse", "u64 u64", C.c_uint64),
            "twiddles": _bind(L, "orc_precompute_twiddles", "u32 u64 p p p"),
            "fwd": _bind(L, "orc_forward_ntt_batch", "p sz sz u64 p", None),
            "inv": _bind(L, "orc_inverse_ntt_batch", "p sz sz u64 p u64", None),
            "fast_inv": _bind(L, "orc_fast_ntt_inverse", "p sz u64 p", None),
            "add": _bind(L, "orc_poly_add", "p p p sz u64", None),
            "sub": _bind(L, "orc_poly_sub", "p p p sz u64", None),
            "neg": _bind(L, "orc_poly_negate", "p p sz u64", None),
            "scalar": _bind(L, "orc_poly_scalar", "p u64 p sz u64", None),
            "pointwise": _bind(L, "orc_poly_pointwise", "p p p sz u64", None),
            "multiply": _bind(L, "orc_poly_multiply", "p p p sz u64 p p u64", None),
            "ml_consts": _bind(L, "orc_mlimb_constants", "p sz p"),
            "ml_mul": _bind(L, "orc_mlimb_montmul", "p p p sz sz p u64", None),
            "ml_add": _bind(L, "orc_mlimb_add", "p p p sz sz p", None),
            "ml_sub": _bind(L, "orc_mlimb_sub", "p p p sz sz p", None),
            "test_poly": _bind(L, "orc_default_test_poly", "p p", None),
            "lut": _bind(L, "orc_lookup_table", "p int u64 u64 p", None),
            "rotate": _bind(L, "orc_rotate_polynomial", "p u32 u64 i32 p", None),
            "decompose": _bind(L, "orc_decompose_polynomial", "p u32 u64 u32 u32 p", None),
            "ext": _bind(L, "orc_external_product", "p p p p", None),
            "cmux": _bind(L, "orc_cmux", "p p p p p", None),
            "blind": _bind(L, "orc_blind_rotate", "p p p p", None),
            "extract": _bind(L, "orc_sample_extract", "p u32 u32 u64 p", None),
            "ks": _bind(L, "orc_key_switch", "p sz u64 p sz u32 u32 p", None),
            "boot": _bind(L, "orc_bootstrap", "p p p p p sz u32 u32 p", None),
            "tally_linear": _bind(L, "orc_tally_linear", "p sz sz u64 p"),
            "tally_tree": _bind(L, "orc_tally_tree", "p sz sz u64 p"),
            "tensor": _bind(L, "orc_tensor_multiply", "p p p sz u64 p p u64", None),
            "relin": _bind(L, "orc_relinearize", "p p u32 u32 u32 p sz u64 p p u64", None),
            "crc32": _bind(L, "orc_crc32", "p sz", C.c_uint32),
            "ballot_parse": _bind(L, "orc_ballot_parse", "p sz u32 sz u64 p p"),
        }

    # -- scalars
    def mod_add(self, a, b, q):
        return self._f["mod_add"](a, b, q)

    def mod_sub(self, a, b, q):
        return self._f["mod_sub"](a, b, q)

    def mod_mul(self, a, b, q):
        return self._f["mod_mul"](a, b, q)

    def mod_pow(self, a, e, q):
        return self._f["mod_pow"](a, e, q)

    def mod_inverse(self, a, q):
        return self._f["mod_inverse"](a, q)

    # -- plan
    def twiddles(self, n, q):
        fwd = np.zeros(n, np.uint64)
        inv = np.zeros(n, np.uint64)
        sc = np.zeros(3, np.uint64)
        if self._f["twiddles"](n, q, _ptr(fwd), _ptr(inv), _ptr(sc)) != 0:
            raise ValueError("modulus is not NTT-friendly")
        return fwd, inv, int(sc[0]), int(sc[1]), int(sc[2])

    # -- transforms ([batch][N] arrays, returns new array)
    def forward(self, x, q, fwd):
        y = u64(x).copy().reshape(-1, x.shape[-1])
        self._f["fwd"](_ptr(y), y.shape[0], y.shape[1], q, _ptr(fwd))
        return y.reshape(x.shape)

    def inverse(self, x, q, inv, inv_n):
        y = u64(x).copy().reshape(-1, x.shape[-1])
        self._f["inv"](_ptr(y), y.shape[0], y.shape[1], q, _ptr(inv), inv_n)
        return y.reshape(x.shape)

    def fast_inverse(self, x, q, inv):
        y = u64(x).copy().reshape(-1, x.shape[-1])
        for row in y:
            self._f["fast_inv"](_ptr(row), row.shape[0], q, _ptr(inv))
        return y.reshape(x.shape)

    # -- ring
    def _bin(self, name, a, b, q):
        a, b = u64(a), u64(b)
        r = np.zeros_like(a)
        self._f[name](_ptr(a), _ptr(b), _ptr(r), a.size, q)
        return r

    def add(self, a, b, q):
        return self._bin("add", a, b, q)

    def sub(self, a, b, q):
        return self._bin("sub", a, b, q)

    def pointwise(self, a, b, q):
        return self._bin("pointwise", a, b, q)

    def negate(self, a, q):
        a = u64(a)
        r = np.zeros_like(a)
        self._f["neg"](_ptr(a), _ptr(r), a.size, q)
        return r

    def scalar(self, a, s, q):
        a = u64(a)
        r = np.zeros_like(a)
        self._f["scalar"](_ptr(a), s, _ptr(r), a.size, q)
        return r

    def multiply(self, a, b, q, fwd, inv, inv_n):
        a = u64(a).reshape(-1, a.shape[-1])
        b = u64(b).reshape(-1, b.shape[-1])
        r = np.zeros_like(a)
        for i in range(a.shape[0]):
            self._f["multiply"](_ptr(a[i]), _ptr(b[i]), _ptr(r[i]), a.shape[1], q, _ptr(fwd), _ptr(inv), inv_n)
        return r

    # -- multi-limb
    def mlimb_constants(self, q_limbs):
        q_limbs = u64(q_limbs)
        l = q_limbs.size
        out = np.zeros(1 + 2 * l, np.uint64)
        if self._f["ml_consts"](_ptr(q_limbs), l, _ptr(out)) != 0:
            raise ValueError("modulus must be odd and non-zero")
        return int(out[0]), out[1 : 1 + l].copy(), out[1 + l :].copy()

    def mlimb_montmul(self, a, b, q_limbs, q_inv):
        a, b, q_limbs = u64(a), u64(b), u64(q_limbs)
        r = np.zeros_like(a)
        self._f["ml_mul"](_ptr(a), _ptr(b), _ptr(r), a.shape[0], a.shape[1], _ptr(q_limbs), q_inv)
        return r

    def mlimb_add(self, a, b, q_limbs):
        a, b, q_limbs = u64(a), u64(b), u64(q_limbs)
        r = np.zeros_like(a)
        self._f["ml_add"](_ptr(a), _ptr(b), _ptr(r), a.shape[0], a.shape[1], _ptr(q_limbs))
        return r

    def mlimb_sub(self, a, b, q_limbs):
        a, b, q_limbs = u64(a), u64(b), u64(q_limbs)
        r = np.zeros_like(a)
        self._f["ml_sub"](_ptr(a), _ptr(b), _ptr(r), a.shape[0], a.shape[1], _ptr(q_limbs))
        return r

    # -- bootstrap
    def boot_params(self, N, q, n, k, base_log, level, t, fwd, inv, inv_n):
        p = BootParams(N, k, n, base_log, level, q, t, fwd.ctypes.data, inv.ctypes.data, inv_n)
        p._keep = (fwd, inv)
        return p

    def default_test_poly(self, p):
        out = np.zeros(p.N, np.uint64)
        self._f["test_poly"](C.addressof(p), _ptr(out))
        return out

    def lookup_table(self, p, kind, arg0, arg1=0):
        out = np.zeros(p.N, np.uint64)
        self._f["lut"](C.addressof(p), kind, arg0, arg1, _ptr(out))
        return out

    def rotate(self, poly, q, rotation):
        poly = u64(poly)
        out = np.zeros_like(poly)
        self._f["rotate"](_ptr(poly), poly.size, q, rotation, _ptr(out))
        return out

    def decompose(self, poly, q, base_log, level):
        poly = u64(poly)
        out = np.zeros((level, poly.size), np.uint64)
        self._f["decompose"](_ptr(poly), poly.size, q, base_log, level, _ptr(out))
        return out

    def external_product(self, p, glwe, ggsw):
        glwe, ggsw = u64(glwe), u64(ggsw)
        out = np.zeros_like(glwe)
        self._f["ext"](C.addressof(p), _ptr(glwe), _ptr(ggsw), _ptr(out))
        return out

    def cmux(self, p, ggsw, ct0, ct1):
        ggsw, ct0, ct1 = u64(ggsw), u64(ct0), u64(ct1)
        out = np.zeros_like(ct0)
        self._f["cmux"](C.addressof(p), _ptr(ggsw), _ptr(ct0), _ptr(ct1), _ptr(out))
        return out

    def blind_rotate(self, p, lwe, bsk, test_poly):
        """lwe [count][n+1] -> acc [count][k+1][N] starting from (0,..,0,test_poly)."""
        lwe, bsk = u64(lwe).reshape(-1, p.n + 1), u64(bsk)
        out = np.zeros((lwe.shape[0], p.k + 1, p.N), np.uint64)
        out[:, p.k, :] = u64(test_poly)
        for i in range(lwe.shape[0]):
            self._f["blind"](C.addressof(p), _ptr(out[i]), _ptr(lwe[i]), _ptr(bsk))
        return out

    def sample_extract(self, glwe, k, N, q):
        glwe = u64(glwe).reshape(-1, k + 1, N)
        out = np.zeros((glwe.shape[0], k * N + 1), np.uint64)
        for i in range(glwe.shape[0]):
            self._f["extract"](_ptr(glwe[i]), k, N, q, _ptr(out[i]))
        return out

    def key_switch(self, lwe, q, ksk, n_out, base_log, level):
        lwe, ksk = u64(lwe), u64(ksk)
        lwe = lwe.reshape(-1, lwe.shape[-1])
        out = np.zeros((lwe.shape[0], n_out + 1), np.uint64)
        for i in range(lwe.shape[0]):
            self._f["ks"](_ptr(lwe[i]), lwe.shape[1] - 1, q, _ptr(ksk), n_out, base_log, level, _ptr(out[i]))
        return out

    def bootstrap(self, p, lwe, bsk, test_poly, ksk=None, n_out=0, ksk_base_log=0, ksk_level=0):
        lwe, bsk, test_poly = u64(lwe).reshape(-1, p.n + 1), u64(bsk), u64(test_poly)
        width = (n_out + 1) if ksk is not None else (p.k * p.N + 1)
        out = np.zeros((lwe.shape[0], width), np.uint64)
        kp = _ptr(u64(ksk)) if ksk is not None else None
        for i in range(lwe.shape[0]):
            self._f["boot"](C.addressof(p), _ptr(lwe[i]), _ptr(bsk), _ptr(test_poly), kp, n_out, ksk_base_log,
                            ksk_level, _ptr(out[i]))
        return out

    # -- tally
    def tally(self, cts, q, tree=False):
        cts = u64(cts)
        m, two, n = cts.shape
        assert two == 2
        out = np.zeros((2, n), np.uint64)
        rc = self._f["tally_tree" if tree else "tally_linear"](_ptr(cts), m, n, q, _ptr(out))
        if rc != 0:
            raise ValueError("Cannot add empty vector of ciphertexts")
        return out

    def tensor_multiply(self, ct1, ct2, q, fwd, inv, inv_n):
        ct1, ct2 = u64(ct1), u64(ct2)
        n = ct1.shape[-1]
        out = np.zeros((3, n), np.uint64)
        self._f["tensor"](_ptr(ct1), _ptr(ct2), _ptr(out), n, q, _ptr(fwd), _ptr(inv), inv_n)
        return out

    def relinearize(self, ct, keys, key_base_log, key_level, q, fwd, inv, inv_n):
        ct, keys = u64(ct), u64(keys)
        n = ct.shape[-1]
        out = np.zeros((2, n), np.uint64)
        self._f["relin"](_ptr(ct), _ptr(keys), keys.shape[0] if keys.size else 0, key_base_log, key_level, _ptr(out), n, q,
                         _ptr(fwd), _ptr(inv), inv_n)
        return out

    def crc32(self, data: bytes):
        buf = np.frombuffer(bytes(data), dtype=np.uint8)
        return int(self._f["crc32"](_ptr(buf) if buf.size else None, buf.size))

    def ballot_parse(self, record: bytes, choices, n, q):
        """-> (status, cts [choices][2][n], timestamp)"""
        buf = np.frombuffer(bytes(record), dtype=np.uint8)
        out = np.zeros((choices, 2, n), np.uint64)
        ts = C.c_uint64()
        rc = self._f["ballot_parse"](_ptr(buf) if buf.size else None, buf.size, choices, n, q, _ptr(out), C.addressof(ts))
        return rc, out, int(ts.value)


def ref_available():
    return os.path.exists(REF_SO)


class RefError(RuntimeError):
    pass


class RefOracle:
    """numpy-facing wrapper over oracle/ref_harness.cpp (the reference's own classes)."""

    def __init__(self):
        if not ref_available():
            raise FileNotFoundError(REF_SO)
        self.lib = C.CDLL(REF_SO)
        L = self.lib
        L.ref_last_error.restype = C.c_char_p
        b = lambda n, s, r=C.c_int: _bind(L, n, s, r)
        self._f = {
            "ntt_create": b("ref_ntt_create", "u32 u64 p"),
            "ntt_destroy": b("ref_ntt_destroy", "p", None),
            "ntt_tables": b("ref_ntt_tables", "p p p p"),
            "ntt_fwd": b("ref_ntt_forward", "p p sz int"),
            "ntt_inv": b("ref_ntt_inverse", "p p sz int"),
            "fast_fwd": b("ref_fast_ntt_forward", "p sz u64 p sz"),
            "fast_inv": b("ref_fast_ntt_inverse", "p sz u64 p sz"),
            "fast_modmul": b("ref_fast_modmul_batch", "p p p sz u64"),
            "ring_create": b("ref_ring_create", "u32 u64 p"),
            "ring_destroy": b("ref_ring_destroy", "p", None),
            "ring_binary": b("ref_ring_binary", "p int p p u64 p sz int"),
            "tally_linear": b("ref_tally_linear", "p p sz p"),
            "tally_tree": b("ref_tally_tree", "p p sz p"),
            "tensor": b("ref_tensor_multiply", "p p p p"),
            "relin": b("ref_relinearize", "p p p u32 u32 u32 p"),
            "crc32": b("ref_crc32", "p sz", C.c_uint32),
            "ser_ballot": b("ref_serialize_ballot", "p u32 u32 u64 u64 p sz p"),
            "de_ballot": b("ref_deserialize_ballot", "p sz u32 u64 p u32 p p"),
            "ser_eval": b("ref_serialize_eval_key", "p u32 u32 u64 u32 u32 u64 p sz p"),
            "de_eval": b("ref_deserialize_eval_key", "p sz u32 p u32 p p"),
            "ser_boot": b("ref_serialize_bootstrap_key", "p u32 u32 p u32 u32 u32 u32 u64 u64 p sz p"),
            "scalar_create": b("ref_scalar_create", "u64 p"),
            "scalar_destroy": b("ref_scalar_destroy", "p", None),
            "scalar_op": b("ref_scalar_op", "p int u64 u64", C.c_uint64),
            "ml_create": b("ref_mlimb_create", "p sz p"),
            "ml_destroy": b("ref_mlimb_destroy", "p", None),
            "ml_consts": b("ref_mlimb_constants", "p p"),
            "ml_batch": b("ref_mlimb_batch", "p int p p p sz sz int"),
            "boot_create": b("ref_boot_create", "u32 u64 u32 u32 u32 u32 u64 p"),
            "boot_destroy": b("ref_boot_destroy", "p", None),
            "boot_test_poly": b("ref_boot_default_test_poly", "p p"),
            "boot_lut": b("ref_boot_lut", "p int u64 u64 p"),
            "boot_keygen": b("ref_boot_keygen", "p int p"),
            "boot_export_bsk": b("ref_boot_export_bsk", "p p"),
            "boot_import_bsk": b("ref_boot_import_bsk", "p p"),
            "boot_ksk_shape": b("ref_boot_ksk_shape", "p p p"),
            "boot_export_ksk": b("ref_boot_export_ksk", "p p"),
            "boot_import_ksk": b("ref_boot_import_ksk", "p p u64 u64 u32 u32"),
            "boot_encrypt": b("ref_boot_encrypt_lwe", "p p sz p p"),
            "boot_decrypt": b("ref_boot_decrypt_lwe", "p p sz sz p p"),
            "boot_decompose": b("ref_boot_decompose", "p p u32 u32 p"),
            "boot_rotate": b("ref_boot_rotate", "p p i32 p"),
            "boot_ext": b("ref_boot_external_product", "p p u32 p"),
            "boot_cmux": b("ref_boot_cmux", "p u32 p p p"),
            "boot_blind": b("ref_boot_blind_rotate", "p p sz p p int"),
            "boot_extract": b("ref_boot_sample_extract", "p p sz p"),
            "boot_ks": b("ref_boot_key_switch", "p p sz sz p"),
            "boot_bootstrap": b("ref_boot_bootstrap", "p p sz p p int"),
        }

    def _ck(self, rc):
        if rc != 0:
            raise RefError(self.lib.ref_last_error().decode())

    def call(self, name, *args):
        self._ck(self._f[name](*[_ptr(a) if isinstance(a, np.ndarray) else a for a in args]))

    # -- NTTProcessor
    def ntt_create(self, n, q):
        h = C.c_void_p()
        self._ck(self._f["ntt_create"](n, q, C.addressof(h)))
        return h

    def ntt_destroy(self, h):
        self._f["ntt_destroy"](h)

    def ntt_tables(self, h, n):
        fwd = np.zeros(n, np.uint64)
        inv = np.zeros(n, np.uint64)
        sc = np.zeros(3, np.uint64)
        self.call("ntt_tables", h, fwd, inv, sc)
        return fwd, inv, int(sc[0]), int(sc[1]), int(sc[2])

    def ntt_forward(self, h, x, threads=1):
        y = u64(x).copy().reshape(-1, x.shape[-1])
        self.call("ntt_fwd", h, y, y.shape[0], threads)
        return y.reshape(x.shape)

    def ntt_inverse(self, h, x, threads=1):
        y = u64(x).copy().reshape(-1, x.shape[-1])
        self.call("ntt_inv", h, y, y.shape[0], threads)
        return y.reshape(x.shape)

    def fast_ntt_forward(self, x, q, tw):
        y = u64(x).copy().reshape(-1, x.shape[-1])
        self.call("fast_fwd", y, y.shape[1], q, u64(tw), y.shape[0])
        return y.reshape(x.shape)

    def fast_ntt_inverse(self, x, q, tw):
        y = u64(x).copy().reshape(-1, x.shape[-1])
        self.call("fast_inv", y, y.shape[1], q, u64(tw), y.shape[0])
        return y.reshape(x.shape)

    def fast_modmul(self, a, b, q):
        a, b = u64(a), u64(b)
        r = np.zeros_like(a)
        self.call("fast_modmul", a, b, r, a.size, q)
        return r

    # -- PolynomialRing
    def ring_create(self, n, q):
        h = C.c_void_p()
        self._ck(self._f["ring_create"](n, q, C.addressof(h)))
        return h

    def ring_destroy(self, h):
        self._f["ring_destroy"](h)

    OPS = {"add": 0, "sub": 1, "pointwise": 2, "multiply": 3, "negate": 4, "scalar": 5}

    def ring_op(self, h, op, a, b=None, scalar=0, threads=1):
        a = u64(a)
        a2 = a.reshape(-1, a.shape[-1])
        b2 = u64(b).reshape(a2.shape) if b is not None else a2
        r = np.zeros_like(a2)
        self.call("ring_binary", h, self.OPS[op], a2, b2, scalar, r, a2.shape[0], threads)
        return r.reshape(a.shape)

    def tally(self, h, cts, tree=False):
        cts = u64(cts)
        out = np.zeros((2, cts.shape[-1]), np.uint64)
        self.call("tally_tree" if tree else "tally_linear", h, cts, cts.shape[0], out)
        return out

    def tensor_multiply(self, h, ct1, ct2):
        ct1, ct2 = u64(ct1), u64(ct2)
        out = np.zeros((3, ct1.shape[-1]), np.uint64)
        self.call("tensor", h, ct1, ct2, out)
        return out

    def relinearize(self, h, ct, keys, key_base_log, key_level):
        ct, keys = u64(ct), u64(keys)
        out = np.zeros((2, ct.shape[-1]), np.uint64)
        self.call("relin", h, ct, keys, keys.shape[0] if keys.size else 0, key_base_log, key_level, out)
        return out

    # -- wire formats (KeySerializer / BallotSerializer)
    def crc32(self, data: bytes):
        buf = np.frombuffer(bytes(data), dtype=np.uint8)
        return int(self._f["crc32"](_ptr(buf) if buf.size else None, buf.size))

    def _ser(self, name, cap, *args):
        out = np.zeros(cap, np.uint8)
        n = C.c_size_t()
        self.call(name, *args, out, cap, C.addressof(n))
        return out[:n.value].tobytes()

    def serialize_ballot(self, choices, q, timestamp):
        choices = u64(choices)
        num, _, n = choices.shape
        return self._ser("ser_ballot", 64 + 12 + num * (12 + 16 * n), choices, num, n, q, timestamp)

    def deserialize_ballot(self, record: bytes, n, q, max_choices=8):
        """-> (cts [num_choices][2][n], timestamp); raises RefError with the reference's message"""
        buf = np.frombuffer(bytes(record), dtype=np.uint8)
        out = np.zeros((max_choices, 2, n), np.uint64)
        num, ts = C.c_uint32(), C.c_uint64()
        self.call("de_ballot", buf if buf.size else None, buf.size, n, q, out, max_choices, C.addressof(num), C.addressof(ts))
        return out[:num.value], int(ts.value)

    def serialize_eval_key(self, keys, q, base_log, level, key_id):
        keys = u64(keys)
        count, _, n = keys.shape
        return self._ser("ser_eval", 64 + 12 + count * 2 * (4 + 8 * n), keys, count, n, q, base_log, level, key_id)

    def deserialize_eval_key(self, data: bytes, n, max_count=64):
        buf = np.frombuffer(bytes(data), dtype=np.uint8)
        out = np.zeros((max_count, 2, n), np.uint64)
        meta = np.zeros(3, np.uint32)
        kid = C.c_uint64()
        self.call("de_eval", buf, buf.size, n, out, max_count, meta, C.addressof(kid))
        return out[:int(meta[0])], int(meta[1]), int(meta[2]), int(kid.value)

    def serialize_bootstrap_key(self, bsk, ksk, ksk_base_log, ksk_level, q, key_id):
        bsk, ksk = u64(bsk), u64(ksk)
        n_lwe, rows, _, n = bsk.shape
        cap = 64 + 8 + n_lwe * (4 + rows * 2 * (4 + 8 * n)) + 12 + ksk.shape[0] * 2 * (4 + 8 * n)
        return self._ser("ser_boot", cap, bsk, n_lwe, rows, ksk if ksk.size else None, ksk.shape[0], ksk_base_log, ksk_level, n, q, key_id)

    # -- scalar ModularArithmetic (the reference addon's class)
    SCALAR_OPS = {"montgomery_mul": 0, "mod_add": 1, "mod_sub": 2, "to_montgomery": 3, "from_montgomery": 4, "get_modulus": 5}

    def scalar_create(self, modulus):
        h = C.c_void_p()
        self._ck(self._f["scalar_create"](modulus, C.addressof(h)))
        return h

    def scalar_destroy(self, h):
        self._f["scalar_destroy"](h)

    def scalar_op(self, h, op, a=0, b=0):
        return int(self._f["scalar_op"](h, self.SCALAR_OPS[op], a, b))

    # -- MultiLimbModularArithmetic
    def mlimb_create(self, q_limbs):
        q_limbs = u64(q_limbs)
        h = C.c_void_p()
        self._ck(self._f["ml_create"](_ptr(q_limbs), q_limbs.size, C.addressof(h)))
        return h

    def mlimb_destroy(self, h):
        self._f["ml_destroy"](h)

    def mlimb_constants(self, h, limbs):
        out = np.zeros(1 + 2 * limbs, np.uint64)
        self.call("ml_consts", h, out)
        return int(out[0]), out[1 : 1 + limbs].copy(), out[1 + limbs :].copy()

    ML_OPS = {"montmul": 0, "add": 1, "sub": 2, "to_mont": 3, "from_mont": 4}

    def mlimb_op(self, h, op, a, b=None, threads=1):
        a = u64(a)
        b = u64(b) if b is not None else a
        r = np.zeros_like(a)
        self.call("ml_batch", h, self.ML_OPS[op], a, b, r, a.shape[0], a.shape[1], threads)
        return r

    # -- BootstrapEngine
    def boot_create(self, N, q, n, k, base_log, level, t):
        h = C.c_void_p()
        self._ck(self._f["boot_create"](N, q, n, k, base_log, level, t, C.addressof(h)))
        h.shape = dict(N=N, q=q, n=n, k=k, base_log=base_log, level=level, t=t)
        return h

    def boot_destroy(self, h):
        self._f["boot_destroy"](h)

    def boot_default_test_poly(self, h):
        out = np.zeros(h.shape["N"], np.uint64)
        self.call("boot_test_poly", h, out)
        return out

    def boot_lut(self, h, kind, arg0, arg1=0):
        out = np.zeros(h.shape["N"], np.uint64)
        self.call("boot_lut", h, kind, arg0, arg1, out)
        return out

    def boot_keygen(self, h, with_ksk=False):
        sk = np.zeros(h.shape["n"], np.int64)
        self.call("boot_keygen", h, 1 if with_ksk else 0, sk)
        return sk

    def bsk_shape(self, h):
        s = h.shape
        return (s["n"], (s["k"] + 1) * s["level"], s["k"] + 1, s["N"])

    def boot_export_bsk(self, h):
        out = np.zeros(self.bsk_shape(h), np.uint64)
        self.call("boot_export_bsk", h, out)
        return out

    def boot_import_bsk(self, h, bsk):
        bsk = u64(bsk)
        assert bsk.shape == self.bsk_shape(h)
        self.call("boot_import_bsk", h, bsk)

    def boot_export_ksk(self, h):
        e = np.zeros(1, np.uint64)
        n_out = np.zeros(1, np.uint64)
        self.call("boot_ksk_shape", h, e, n_out)
        out = np.zeros((int(e[0]), int(n_out[0]) + 1), np.uint64)
        self.call("boot_export_ksk", h, out)
        return out

    def boot_import_ksk(self, h, ksk, base_log, level):
        ksk = u64(ksk)
        self.call("boot_import_ksk", h, ksk, ksk.shape[0], ksk.shape[1] - 1, base_log, level)

    def boot_encrypt_lwe(self, h, values, sk):
        values = u64(values)
        out = np.zeros((values.size, h.shape["n"] + 1), np.uint64)
        self.call("boot_encrypt", h, values, values.size, np.ascontiguousarray(sk, np.int64), out)
        return out

    def boot_decrypt_lwe(self, h, cts, sk):
        cts = u64(cts)
        out = np.zeros(cts.shape[0], np.uint64)
        self.call("boot_decrypt", h, cts, cts.shape[0], cts.shape[1] - 1, np.ascontiguousarray(sk, np.int64), out)
        return out

    def boot_decompose(self, h, poly, base_log, level):
        out = np.zeros((level, h.shape["N"]), np.uint64)
        self.call("boot_decompose", h, u64(poly), base_log, level, out)
        return out

    def boot_rotate(self, h, poly, rotation):
        out = np.zeros(h.shape["N"], np.uint64)
        self.call("boot_rotate", h, u64(poly), rotation, out)
        return out

    def boot_external_product(self, h, glwe, index):
        glwe = u64(glwe)
        out = np.zeros_like(glwe)
        self.call("boot_ext", h, glwe, index, out)
        return out

    def boot_cmux(self, h, index, ct0, ct1):
        ct0, ct1 = u64(ct0), u64(ct1)
        out = np.zeros_like(ct0)
        self.call("boot_cmux", h, index, ct0, ct1, out)
        return out

    def boot_blind_rotate(self, h, lwe, test_poly, threads=1):
        s = h.shape
        lwe = u64(lwe).reshape(-1, s["n"] + 1)
        out = np.zeros((lwe.shape[0], s["k"] + 1, s["N"]), np.uint64)
        self.call("boot_blind", h, lwe, lwe.shape[0], u64(test_poly), out, threads)
        return out

    def boot_sample_extract(self, h, glwe):
        s = h.shape
        glwe = u64(glwe).reshape(-1, s["k"] + 1, s["N"])
        out = np.zeros((glwe.shape[0], s["k"] * s["N"] + 1), np.uint64)
        self.call("boot_extract", h, glwe, glwe.shape[0], out)
        return out

    def boot_key_switch(self, h, lwe, n_out):
        lwe = u64(lwe)
        lwe = lwe.reshape(-1, lwe.shape[-1])
        out = np.zeros((lwe.shape[0], n_out + 1), np.uint64)
        self.call("boot_ks", h, lwe, lwe.shape[0], lwe.shape[1] - 1, out)
        return out

    def boot_bootstrap(self, h, lwe, test_poly, n_out, threads=1):
        s = h.shape
        lwe = u64(lwe).reshape(-1, s["n"] + 1)
        out = np.zeros((lwe.shape[0], n_out + 1), np.uint64)
        self.call("boot_bootstrap", h, lwe, lwe.shape[0], u64(test_poly), out, threads)
        return out


def mt19937_64_coeffs(seed, count, q):
    """`TestRandom(seed).next_coefficient(q)` of the reference's test harness
    (cpp/tests/test_harness.h:29-47): libstdc++ uniform_int_distribution<u64>(0, 2^64-1)
    returns the raw mt19937_64 output, then `% q`.  numpy's MT19937 is the 32-bit
    generator, so the 64-bit variant is restated here (standard MT19937-64 constants)."""
    nn, mm = 312, 156
    matrix_a, um, lm = 0xB5026F5AA96619E9, 0xFFFFFFFF80000000, 0x7FFFFFFF
    mask = (1 << 64) - 1
    mt = [0] * nn
    mt[0] = seed & mask
    for i in range(1, nn):
        mt[i] = (6364136223846793005 * (mt[i - 1] ^ (mt[i - 1] >> 62)) + i) & mask
    out = np.zeros(count, np.uint64)
    idx = nn
    for c in range(count):
        if idx >= nn:
            for i in range(nn):
                x = (mt[i] & um) | (mt[(i + 1) % nn] & lm)
                v = mt[(i + mm) % nn] ^ (x >> 1)
                if x & 1:
                    v ^= matrix_a
                mt[i] = v
            idx = 0
        x = mt[idx]
        idx += 1
        x ^= (x >> 29) & 0x5555555555555555
        x ^= (x << 17) & 0x71D67FFFEDA60000
        x ^= (x << 37) & 0xFFF7EEE000000000
        x ^= x >> 43
        out[c] = (x & mask) % q
    return out
