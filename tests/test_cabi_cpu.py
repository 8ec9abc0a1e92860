"""CPU-side checks of the boundary and the host logic (no GPU needed, no compute calls):

* libfheb200.so loads and exports every symbol include/fheb200.h declares, and the Python
  binding table covers exactly those symbols;
* without a GPU every compute entry point fails loudly (HARDWARE_UNAVAILABLE) - there is no
  CPU fallback in the product;
* shard partitioning, and the sharded tally's exchange step on a world_size-2 gloo group.
"""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fheb200.h")


@pytest.fixture(scope="module")
def fhe():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge

    if not os.path.exists(os.path.join(ROOT, "node-fhe-accelerate_b200", "libfheb200.so")):
        ge.build()
    import fheb200

    return fheb200


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"FHEB_API\s+[^;(]*?\b(fheb_\w+)\s*\(", text)))


def test_header_symbols_exported_and_bound(fhe):
    syms = declared_symbols()
    assert len(syms) >= 40
    lib = fhe.lib()
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/fheb200.h but not exported"
    assert sorted(fhe.SIGNATURES) == syms
    out = subprocess.run(["nm", "-D", "--defined-only", fhe.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (fheb_\w+)", out))
    assert exported == set(syms), exported ^ set(syms)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "node-fhe-accelerate_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "libfhe_oracle" not in text and "oracle_bindings" not in text and "libref_oracle" not in text, f


def test_no_cpu_fallback_without_gpu(fhe):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert fhe.version().endswith("b200")
    with pytest.raises(fhe.FheError) as e:
        fhe.initialize()
    assert e.value.code_name == "HARDWARE_UNAVAILABLE"
    x = np.arange(8, dtype=np.uint64)
    for call in (lambda: fhe.NTTProcessor(8, 97), lambda: fhe.modadd_batch(x, x, 97),
                 lambda: fhe.tally_votes(np.zeros((2, 2, 8), np.uint64), 8, 97),
                 lambda: fhe.MultiLimbModularArithmetic([0xFFFFFFFFFFFFFF43, 1]).mod_add(np.zeros((1, 2), np.uint64),
                                                                                         np.zeros((1, 2), np.uint64)),
                 lambda: fhe.ingest_ballots(fhe.serialize_ballot(np.zeros((1, 2, 8), np.uint64), 97, 1), 1, 1, 8, 97),
                 lambda: fhe.tally_wire(fhe.serialize_ballot(np.zeros((1, 2, 8), np.uint64), 97, 1) * 2, 2, 1, 8, 97)):
        with pytest.raises(fhe.FheError) as e:
            call()
        assert e.value.code_name == "HARDWARE_UNAVAILABLE", e.value


def test_multi_limb_constants_are_host_side(fhe, oracle):
    for q in ([0xFFFFFFFFFFFFFF43, 1], [0xFFFFFFFFFFFFFFC5], [0x1D, 0, 1]):
        ml = fhe.MultiLimbModularArithmetic(q)
        q_inv, r1, r2 = oracle.mlimb_constants(np.array(q, np.uint64))
        assert ml.q_inv == q_inv
        assert np.array_equal(ml.r_mod_q, r1) and np.array_equal(ml.r2_mod_q, r2)
    with pytest.raises(fhe.FheError):
        fhe.MultiLimbModularArithmetic([2, 1])  # even modulus


def test_shard_range_partitions(fhe):
    for total in (0, 1, 7, 4096, 10**6):
        for world in (1, 2, 3, 8):
            spans = [fhe.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        fhe.shard_range(10, 2, 2)


def test_noise_budget_metadata(fhe):
    assert fhe.tally_noise_budget([30.0] * 20000, "linear") == pytest.approx(30.0 - np.log2(20000))
    assert fhe.tally_noise_budget([30.0] * 20000, "tree") == 15.0
    assert fhe.tally_noise_budget([12.5], "tree") == 12.5
    # heterogeneous budgets: the tree subtracts 1 per level along each path and carries the odd element
    # (cpp/src/encryption.cpp:1413,1437,1449-1451); restated here level by level
    bud = [30.0, 28.5, 31.0, 29.25, 27.0, 33.0, 30.5]
    lvl = list(bud)
    while len(lvl) > 1:
        nxt = [min(lvl[i], lvl[i + 1]) - 1.0 for i in range(0, len(lvl) - 1, 2)]
        if len(lvl) % 2:
            nxt.append(lvl[-1])
        lvl = nxt
    assert fhe.tally_noise_budget(bud, "tree") == lvl[0]
    acc = bud[0]
    for b in bud[1:]:
        acc = min(acc, b) - 1.0   # EncryptionEngine::add, :613
    assert fhe.tally_noise_budget(bud, "add") == acc
    assert fhe.tally_noise_budget(bud, "linear") == min(bud) - np.log2(len(bud))
    with pytest.raises(fhe.FheError):
        fhe.tally_noise_budget([], "tree")


_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import fheb200
from oracle_bindings import Oracle
rank, world = int(sys.argv[1]), 2
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=sys.argv[2], RANK=str(rank), WORLD_SIZE=str(world))
dist.init_process_group("gloo", rank=rank, world_size=world)
orc = Oracle()
n, q, total = 64, 1099511678977, 101
cts = np.random.default_rng(3).integers(0, q, size=(total, 2, n), dtype=np.uint64)
lo, hi = fheb200.shard_range(total, rank, world)
# host logic under test: partition + all-gather + combine order.  The kernels are replaced by the
# CPU checker here only because this test runs without a GPU; the product default is the CUDA path.
to_t = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64))
to_n = lambda t: t.numpy().view(np.uint64)
st = fheb200.ShardedTally(n, q, local_fn=lambda c: to_t(orc.tally(to_n(c), q)),
                          combine_fn=lambda p: to_t(orc.tally(to_n(p), q)))
got = to_n(st.tally(to_t(cts[lo:hi])))
assert np.array_equal(got, orc.tally(cts, q)), "sharded tally differs"
dist.destroy_process_group()
print("ok", rank)
"""


def test_sharded_tally_exchange_gloo_world2(fhe, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), str(r), port], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"ok {r}" in o, o


def test_scalar_modular_arithmetic_matches_reference(fhe, ref):
    """The addon's existing scalar class (src/native/lib.rs:44-120): host-side code, word-for-word the
    reference's - including its (mathematically wrong) Montgomery constant."""
    rng = np.random.default_rng(17)
    for q in (17, 97, 132120577, 1099511678977, 4611686018326724609, 0xFFFFFFFFFFFFFFC5):
        h = ref.scalar_create(q)
        m = fhe.ModularArithmetic(q)
        try:
            assert m.get_modulus() == ref.scalar_op(h, "get_modulus") == q
            vals = [0, 1, q - 1, q, q + 1] + [int(v) for v in rng.integers(0, 2**63, size=200, dtype=np.uint64)]
            for a, b in zip(vals, reversed(vals)):
                for op in ("montgomery_mul", "mod_add", "mod_sub"):
                    assert getattr(m, op)(a, b) == ref.scalar_op(h, op, a, b), (q, op, a, b)
                for op in ("to_montgomery", "from_montgomery"):
                    assert getattr(m, op)(a) == ref.scalar_op(h, op, a), (q, op, a)
        finally:
            ref.scalar_destroy(h)
    for bad in (0, 16):
        with pytest.raises(fhe.FheError) as e:
            fhe.ModularArithmetic(bad)
        assert "Modulus must be" in str(e.value)
    with pytest.raises(fhe.FheError) as e:
        fhe.ModularArithmetic(17).mod_add(-1, 2)
    assert "non-negative" in str(e.value)


def test_node_addon_source_type_checks_and_links(fhe, tmp_path):
    """the Node-API addon cannot be built for real here (no Node toolchain).  It is type-checked against
    the stand-in declarations in addon/stub/node_api.h and linked against libfheb200.so: every symbol it leaves
    undefined must be an N-API function (resolved by the node binary at load time) or come from the C/C++ runtime."""
    src = os.path.join(ROOT, "addon", "fheb_addon.cc")
    out = str(tmp_path / "addon.node")
    cmd = ["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-fPIC", "-shared", "-I", os.path.join(ROOT, "addon", "stub"),
           "-I", os.path.join(ROOT, "include"), src, "-o", out, "-L", os.path.dirname(fhe.LIB_PATH), "-lfheb200"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    nm = subprocess.run(["nm", "-D", "--undefined-only", out], capture_output=True, text=True).stdout
    undefined = set(re.findall(r"\bU (\w+)", nm))
    napi = {s for s in undefined if s.startswith("napi_")}
    fheb = {s for s in undefined if s.startswith("fheb_")}
    assert napi and fheb
    assert fheb <= set(declared_symbols())          # only declared C-ABI entry points
    stub = open(os.path.join(ROOT, "addon", "stub", "node_api.h")).read()
    assert all(re.search(rf"\b{s}\(", stub) for s in napi)
    defined = subprocess.run(["nm", "-D", "--defined-only", out], capture_output=True, text=True).stdout
    assert "napi_register_module_v1" in defined     # the entry point node looks up


def _build_addon_and_host(fhe, out_dir):
    """The addon as the shared object node-gyp would produce (addon/binding.gyp), and the mock Node-API host that loads it."""
    node = os.path.join(out_dir, "fheb_addon.node")
    host = os.path.join(out_dir, "mock_napi")
    lib_dir = os.path.dirname(fhe.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "addon", "stub"),
                           "-I", os.path.join(ROOT, "include"), "-o", node, os.path.join(ROOT, "addon", "fheb_addon.cc"),
                           "-L" + lib_dir, "-lfheb200", "-Wl,-rpath," + lib_dir])
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "addon", "stub"), "-o", host,
                           os.path.join(ROOT, "tests", "napi_host", "mock_napi.cpp"), os.path.join(ROOT, "oracle", "fhe_oracle.c"),
                           "-ldl", "-rdynamic"])
    return host, node


def test_node_addon_runs_under_a_mock_node_api_host(fhe, tmp_path):
    """SURVEY 8f N1, as far as this image allows: the addon is built into the .node shared object, loaded with dlopen +
    napi_register_module_v1 exactly as Node loads it, and its exports are driven by tests/napi_host/mock_napi.cpp (a small
    Node-API host: values, typed arrays, classes, exceptions).  Without a GPU the scalar ModularArithmetic surface of
    index.d.ts:14-44 runs end to end and initialize() must fail with HARDWARE_UNAVAILABLE (no CPU fallback); the bulk
    entry points run in the GPU suite (tests/test_gpu_cpp.py)."""
    host, node = _build_addon_and_host(fhe, str(tmp_path))
    r = subprocess.run([host, node, "cpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "NAPI HOST CPU OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
