"""ctypes binding of libfheb200.so (include/fheb200.h).

The library is the product: there is no Python or CPU fallback.  Importing this module
without the built library raises; calling any compute entry point without an sm_100 GPU
raises FheError(HARDWARE_UNAVAILABLE).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, os.environ.get("FHEB_LIB", "libfheb200.so"))  # FHEB_LIB: kernel experiments only

OK, INVALID_PARAMETERS, KEY_MISMATCH, HARDWARE_UNAVAILABLE, NATIVE_ERROR, OUT_OF_MEMORY = range(6)
_CODE_NAMES = {
    1: "INVALID_PARAMETERS", 2: "KEY_MISMATCH", 3: "HARDWARE_UNAVAILABLE", 4: "NATIVE_ERROR", 5: "OUT_OF_MEMORY",
}


class FheError(RuntimeError):
    """Mirrors FHEError / FHEErrorCode of the reference's TS API (src/api/types.ts:140-166)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"{_CODE_NAMES.get(code, code)}: {message}")
        self.code = code
        self.code_name = _CODE_NAMES.get(code, str(code))
        self.message = message


class DeviceInfo(C.Structure):
    _fields_ = [
        ("has_sme", C.c_int32), ("has_metal", C.c_int32), ("has_neon", C.c_int32), ("has_amx", C.c_int32),
        ("has_cuda", C.c_int32), ("cc_major", C.c_int32), ("cc_minor", C.c_int32), ("sm_count", C.c_int32),
        ("device_memory_bytes", C.c_uint64), ("l2_bytes", C.c_uint64), ("smem_per_block_optin", C.c_uint64),
        ("name", C.c_char * 128),
    ]


class WireHeader(C.Structure):
    """fheb_wire_header: SerializationHeader of cpp/include/key_serializer.h:59-84."""
    _fields_ = [("magic", C.c_uint32), ("version", C.c_uint32), ("key_type", C.c_uint32), ("key_id", C.c_uint64),
                ("poly_degree", C.c_uint32), ("modulus", C.c_uint64), ("data_size", C.c_uint32), ("checksum_type", C.c_uint8),
                ("compression", C.c_uint8), ("checksum", C.c_uint32)]


class BootParams(C.Structure):
    _fields_ = [("lwe_dimension", C.c_uint32), ("glwe_dimension", C.c_uint32), ("decomp_base_log", C.c_uint32),
                ("decomp_level", C.c_uint32)]


u32, u64, sz, p, i = C.c_uint32, C.c_uint64, C.c_size_t, C.c_void_p, C.c_int

# name -> (argtypes, restype); every symbol declared in include/fheb200.h
SIGNATURES = {
    "fheb_init": ([i], i),
    "fheb_shutdown": ([], i),
    "fheb_device_count": ([], i),
    "fheb_set_devices": ([p, i], i),
    "fheb_get_devices": ([p, i], i),
    "fheb_version": ([], C.c_char_p),
    "fheb_last_error": ([], C.c_char_p),
    "fheb_device_info_get": ([p], i),
    "fheb_device_alloc": ([p, sz], i),
    "fheb_device_free": ([p], i),
    "fheb_host_alloc": ([p, sz], i),
    "fheb_host_free": ([p], i),
    "fheb_copy": ([p, p, sz, p], i),
    "fheb_modarith_create": ([u64, p], i),
    "fheb_modarith_destroy": ([p], i),
    "fheb_modarith_montgomery_mul": ([p, u64, u64], u64),
    "fheb_modarith_mod_add": ([p, u64, u64], u64),
    "fheb_modarith_mod_sub": ([p, u64, u64], u64),
    "fheb_modarith_to_montgomery": ([p, u64], u64),
    "fheb_modarith_from_montgomery": ([p, u64], u64),
    "fheb_modarith_get_modulus": ([p], u64),
    "fheb_synchronize": ([p], i),
    "fheb_ntt_plan_create": ([u32, u64, p], i),
    "fheb_ntt_plan_create_with_tables": ([u32, u64, p, p, u64, p], i),
    "fheb_ntt_plan_destroy": ([p], i),
    "fheb_ntt_plan_get_tables": ([p, p, p, p], i),
    "fheb_ntt_plan_degree": ([p], u32),
    "fheb_ntt_plan_modulus": ([p], u64),
    "fheb_ntt_forward_batch": ([p, p, p, sz, p], i),
    "fheb_ntt_inverse_batch": ([p, p, p, sz, p], i),
    "fheb_ntt_inverse_fwdnet_batch": ([p, p, p, sz, p], i),
    "fheb_polymul_batch": ([p, p, p, p, sz, p], i),
    "fheb_modadd_batch": ([p, p, p, sz, u64, p], i),
    "fheb_modsub_batch": ([p, p, p, sz, u64, p], i),
    "fheb_modmul_batch": ([p, p, p, sz, u64, p], i),
    "fheb_modneg_batch": ([p, p, sz, u64, p], i),
    "fheb_modmul_scalar_batch": ([p, u64, p, sz, u64, p], i),
    "fheb_mlimb_montmul_batch": ([p, p, p, sz, u32, p, u64, p], i),
    "fheb_mlimb_add_batch": ([p, p, p, sz, u32, p, p], i),
    "fheb_mlimb_sub_batch": ([p, p, p, sz, u32, p, p], i),
    "fheb_mlimb_constants": ([p, u32, p], i),
    "fheb_boot_key_create": ([p, p, p, p], i),
    "fheb_boot_key_set_ksk": ([p, p, sz, u32, u32, u32], i),
    "fheb_boot_key_destroy": ([p], i),
    "fheb_external_product_batch": ([p, u32, p, p, sz, p], i),
    "fheb_cmux_batch": ([p, u32, p, p, p, sz, p], i),
    "fheb_blind_rotate_batch": ([p, p, p, p, sz, p], i),
    "fheb_sample_extract_batch": ([p, p, p, sz, p], i),
    "fheb_key_switch_batch": ([p, p, p, sz, p], i),
    "fheb_bootstrap_batch": ([p, p, p, p, sz, p], i),
    "fheb_make_test_poly": ([p, i, u64, u64, p], i),
    "fheb_tally": ([p, sz, u32, u64, p, p], i),
    "fheb_tally_combine": ([p, sz, u32, u64, p, p], i),
    "fheb_tally_noise_budget": ([p, sz, i, p], i),
    "fheb_tally_peers_create": ([u32, u64, u32, u32, p, p], i),
    "fheb_tally_peers_connect": ([p, p], i),
    "fheb_tally_peers_run": ([p, p, sz, p, p], i),
    "fheb_tally_peers_status": ([p, p], i),
    "fheb_tally_peers_epoch": ([p], u32),
    "fheb_tally_peers_reset": ([p, u32], i),
    "fheb_tally_peers_set_timeout": ([p, C.c_double], i),
    "fheb_tally_peers_destroy": ([p], i),
    "fheb_tally_group_create": ([u32, u64, p, u32, p], i),
    "fheb_tally_sharded": ([p, p, p, p], i),
    "fheb_tally_group_size": ([p], u32),
    "fheb_tally_group_destroy": ([p], i),
    "fheb_tally_stream_create": ([u32, u64, p], i),
    "fheb_tally_stream_add": ([p, p, sz, p], i),
    "fheb_tally_stream_total": ([p, p, p], i),
    "fheb_tally_stream_count": ([p], u64),
    "fheb_tally_stream_destroy": ([p], i),
    "fheb_tensor_multiply_batch": ([p, p, p, p, sz, p], i),
    "fheb_relin_key_create": ([p, p, u32, u32, u32, u64, p], i),
    "fheb_relin_key_destroy": ([p], i),
    "fheb_relin_key_levels": ([p], u32),
    "fheb_relinearize_batch": ([p, p, u64, p, sz, p], i),
    "fheb_wire_header_read": ([p, sz, p], i),
    "fheb_wire_crc32": ([p, sz], u32),
    "fheb_ballot_wire_size": ([u32, u32], sz),
    "fheb_ballot_serialize": ([p, u32, u32, u64, u64, p, sz, p], i),
    "fheb_ballots_ingest": ([p, sz, p, sz, u32, u32, u64, p, p, p, p, p], i),
    "fheb_tally_wire": ([p, sz, p, sz, u32, u32, u64, p, p, p, p], i),
    "fheb_relin_key_from_wire": ([p, p, sz, p], i),
    "fheb_boot_key_from_wire": ([p, p, p, sz, p], i),
    "fheb_synth_ballots": ([p, sz, sz, u32, u64, u64, p], i),
    "fheb_launch_count": ([i], u64),
}

_lib = None


def lib() -> C.CDLL:
    """Loads libfheb200.so (fails loudly when it has not been built: run build.py)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python node-fhe-accelerate_b200/build.py` "
                "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (args, res) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError here means header and library disagree
            fn.argtypes = args
            fn.restype = res
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != OK:
        raise FheError(rc, lib().fheb_last_error().decode())
