"""not gpu: the C-ABI library loads on a CPU-only host and exports every symbol include/*.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "meshanything_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ma_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    names = _declared()
    for must in ("ma_decode_generate", "ma_linear_f16", "ma_attention_f16", "ma_layernorm", "ma_last_error"):
        assert must in names


def test_library_loads_and_exports_all_symbols():
    from meshanything_b200 import capi
    lib = capi.lib()
    assert lib.ma_abi_version() == 1
    raw = ctypes.CDLL(capi.lib_path())
    for name in _declared():
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
    assert set(capi.EXPORTS) <= set(_declared())


def test_size_queries_need_no_gpu():
    from meshanything_b200 import capi
    lib = capi.lib()
    # 24 layers x K,V x 16 heads x 64 x fp16 = 98304 bytes per cached position
    assert lib.ma_kv_cache_bytes(24, 1, 1000) == 98304 * 1000
    assert lib.ma_kv_cache_bytes(24, 64, 7459) == 98304 * 7459 * 64
    assert lib.ma_decoder_workspace_bytes(1, 7459) > 0
    assert lib.ma_attention_scratch_bytes(1, 16, 7459) > 0


def test_product_does_not_import_the_oracle():
    """The product packages must not reference oracle/ (it is test infrastructure)."""
    bad = []
    for pkg in ("meshanything_b200", "MeshAnything"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, pkg)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    txt = open(os.path.join(dirpath, f)).read()
                    if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "libma_oracle" in txt:
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_fhfma_variant_builds_and_exports_the_same_symbols():
    """The opt-in FHFMA build must keep compiling: same sources with -DMA_FHFMA into
    libmeshanything_b200_fhfma.so, every header symbol exported, and the mixed-precision FMA present in its SASS."""
    import shutil
    import subprocess
    import sys
    if shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"):
        import pytest
        pytest.skip("nvcc not available")
    env = dict(os.environ, MA_B200_FHFMA="1")
    r = subprocess.run([sys.executable, "-m", "meshanything_b200.build"], cwd=ROOT, env=env, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    path = r.stdout.strip().splitlines()[-1]
    assert path.endswith("libmeshanything_b200_fhfma.so") and os.path.exists(path)
    try:
        raw = ctypes.CDLL(path)
        for name in _declared():
            assert hasattr(raw, name), name
        cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
        if os.path.exists(cuobjdump):
            sass = subprocess.run([cuobjdump, "-sass", path], capture_output=True, text=True, timeout=600).stdout
            assert sass.count("FHFMA") > 2000     # gemm_canon + fast_gemv + attention + the persistent kernel
    finally:
        libdir = os.path.dirname(path)
        for f in os.listdir(libdir):
            if "_fhfma" in f:
                os.remove(os.path.join(libdir, f))
