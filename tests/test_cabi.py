"""CPU: the C-ABI library loads, exports every symbol include/crick_cuda.h declares, and fails
loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "crick_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(crick_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from crick_b200 import _lib
    names = declared_symbols()
    assert len(names) > 60
    for n in names:
        assert hasattr(_lib.lib, n), n
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if l.strip()}
    assert set(names) <= exported
    # nothing but the C ABI leaks out of the shared object
    assert all(s.startswith("crick_") for s in exported if not s.startswith("_")), exported


def test_python_binding_covers_the_header():
    from crick_b200 import _lib
    assert set(_lib._SIG) == set(declared_symbols())


def test_no_cpu_fallback_without_device():
    import crick_b200 as cb
    if cb.device_count() > 0:
        pytest.skip("a CUDA device is present")
    for ctor in (cb.TDigest, cb.SummaryStats, lambda: cb.SpaceSaving(10, "i8"),
                 lambda: cb.TDigestGroup(4)):
        with pytest.raises(RuntimeError, match="no CUDA device"):
            ctor()


def test_product_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "crick_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "liboracle" not in src and "libcrick_ref" not in src, f


def test_version_and_exports():
    import crick
    import crick_b200
    assert crick.TDigest is crick_b200.TDigest
    assert crick.TDigest.__module__ == "crick.tdigest"
    assert crick.SummaryStats.__module__ == "crick.stats"
    assert crick.SpaceSaving.__module__ == "crick.space_saving"
    assert isinstance(crick.__version__, str) and isinstance(crick.__numpy_version__, bytes)
    assert crick_b200._lib.lib.crick_version().decode().startswith("crick_b200")
