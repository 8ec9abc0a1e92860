"""B200 (sm_100a) backend for node-fhe-accelerate's data-parallel hot path.

The compute lives in libfheb200.so (hand-written CUDA behind the C ABI of include/fheb200.h);
this package is the thin host-side mirror of the reference's C++ class surface.  The directory
name carries a hyphen, so import it through the repo-root shim: ``import fheb200``.
"""
from ._cabi import FheError, LIB_PATH, SIGNATURES, lib  # noqa: F401
from .api import (  # noqa: F401
    BootstrapEngine, CiphertextStreamAccumulator, ModularArithmetic, MultiLimbModularArithmetic, NTTProcessor, PolynomialRing, RelinearizationKey, RnsPolynomialRing, batch_add, detect_hardware, initialize,
    launch_count, modadd_batch, pinned_empty, modmul_batch, modmul_scalar_batch, modneg_batch, modsub_batch, synchronize,
    ingest_ballots, serialize_ballot, set_devices, synth_ballots, tally_combine, tally_noise_budget, tally_votes, tally_wire, version, wire_crc32, wire_header,
)
from .sharded import ShardedTally, TallyGroup, shard_range  # noqa: F401
