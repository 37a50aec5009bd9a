"""The C-ABI library loads without a GPU and exports every function the public headers declare
(include/dsstne_b200.h -- kernel-level drop-in boundary; include/dsstne_b200_engine.h -- network level).
No compute calls here: those need a B200 and live in the -m gpu tests."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dsb200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(dsb):
    lib = dsb.lib()
    for header in ("dsstne_b200.h", "dsstne_b200_engine.h"):
        names = declared_functions(header)
        assert len(names) > 20
        missing = [n for n in names if not hasattr(lib, n)]
        assert not missing, f"{header}: not exported: {missing}"


def test_version_and_host_only_helpers(dsb, orc):
    lib = dsb.lib()
    assert lib.dsb200_version() == 100
    a, b = ctypes.c_uint32(), ctypes.c_uint32()
    for N, P in [(27278, 8), (1000000, 8), (128, 3), (7, 4)]:
        covered = 0
        for r in range(P):
            lib.dsb200_shard_range(ctypes.c_uint32(N), ctypes.c_uint32(r), ctypes.c_uint32(P), ctypes.byref(a), ctypes.byref(b))
            assert (a.value, b.value) == orc.shard_range(N, r, P)          # E/NNLayer.cpp:108-112
            assert a.value == covered
            covered = b.value
        assert covered == N
    for i, o in [(27278, 128), (128, 128), (128, 27278), (1024, 1000000), (300, 200), (200, 300)]:
        assert bool(lib.dsb200_weight_outgoing_larger(ctypes.c_uint32(i), ctypes.c_uint32(o))) == orc.weight_outgoing_larger(i, o)
        assert orc.weight_outgoing_larger(i, o) == (o * 3 > i * 2)         # E/NNWeight.cpp:435-457


def test_no_gpu_means_loud_failure_not_cpu_fallback(dsb):
    import torch
    if torch.cuda.is_available():
        return
    h = ctypes.c_void_p()
    rc = dsb.lib().dsb200_ctx_create(ctypes.byref(h), ctypes.c_int(0))
    assert rc != 0 and not h.value


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing in the product package, the headers or the engine may reference it."""
    bad = []
    for base in ("amazon-dsstne_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dirpath:
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"liboracle|from oracle|import oracle|oracle/_ref|dsstne_oracle\.h", text):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_cpp_shim_with_reference_kernel_names_compiles(tmp_path):
    """include/dsstne_b200_kernels.hpp re-declares the E/kernels.h launcher names over the C ABI; it must compile as
    C++14 (the reference's dialect, Makefile.inc:64) and cover the launchers of the hot path."""
    import subprocess
    src = tmp_path / "shim.cpp"
    src.write_text('#include "dsstne_b200_kernels.hpp"\n'
                   "void use(float* f, uint64_t* u64, uint32_t* u32) {\n"
                   "  kCalculateSparseZ(0, 1, 4, f, u64, u64, u32, f, f, 1.0f);\n"
                   "  kCalculateSparseAnalogZ<unsigned char>(0, 1, 4, f, u64, u64, u32, f, (unsigned char*)0, f, 1.0f);\n"
                   "  kCalculateSparseTransposedWeightGradient(1.0f, 0.0f, 4, 4, u32, u32, u32, f, f);\n"
                   "  (void)kCalculateSparseScaledMarginalCrossEntropyError(0, 1, 4, f, u64, u64, u32, f, false);\n"
                   "  kCalculateSparseCrossEntropyOutputDelta(0, 0, 1, 4, f, f, u64, u64, u32, f, false);\n"
                   "  kAdamUpdateWeights(.1f, 0, 0, .9f, .999f, 1.f, 16, f, f, f, f);\n"
                   "  kCalculateTopK(f, f, u32, 1, 4, 2);\n"
                   "}\nint main() { return 0; }\n")
    r = subprocess.run(["/usr/bin/g++", "-std=c++14", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    text = open(os.path.join(ROOT, "include", "dsstne_b200_kernels.hpp")).read()
    for name in ["kCalculateSparseZ", "kCalculateIndexedSparseZ", "kCalculateSparseAnalogZ", "kCalculateSparseDenoisedZ",
                 "kCalculateSparseTransposedMatrix", "kCalculateSparseTransposedWeightGradient", "kCalculateSparseL2Error",
                 "kCalculateSparseCrossEntropyError", "kCalculateSparseScaledMarginalCrossEntropyError", "kCalculateSparseOutputDelta",
                 "kCalculateSparseCrossEntropyOutputDelta", "kCalculateSparseScaledMarginalCrossEntropyOutputDelta", "kSGDUpdateWeights",
                 "kMomentumUpdateWeights", "kAdaGradUpdateWeights", "kNesterovUpdateWeights", "kRMSPropUpdateWeights", "kAdaDeltaUpdateWeights",
                 "kAdamUpdateWeights", "kCalculateTopK", "kClearUnit", "kCalculateRegularizationError"]:
        assert re.search(r"\b" + name + r"\b", text), name
