"""The C-ABI library: builds, loads, exports every symbol include/vtamiq_b200.h declares.  No compute calls."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "vtamiq_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vtq_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(built_lib):
    lib = ctypes.CDLL(built_lib)
    names = _declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"


def test_python_binding_covers_header(built_lib):
    from vtamiq_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    lib = _lib.load_library()
    assert lib.vtq_abi_version() == _lib.ABI_VERSION


def test_sass_is_blackwell_native(built_lib):
    """tcgen05 / TMA evidence in the shipped cubin (B200_PROFILING.md: UTC*MMA, UTMALDG, LDTM)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", built_lib], capture_output=True, text=True).stdout
    for mnem in ("UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "STTM"):
        assert mnem in sass, mnem
    assert "HMMA." not in sass.replace("UTCHMMA.", ""), "legacy mma.sync path must not be present"


def test_no_device_fails_loudly(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vtamiq_b200 import _lib
    with pytest.raises(_lib.VtqError, match="no CUDA device|CPU fallback"):
        _lib.Context(0)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from vtamiq_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.VtqError, match="missing"):
        _lib.load_library()
