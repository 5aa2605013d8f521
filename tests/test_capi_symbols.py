"""CPU test: the C-ABI shared library loads and exports every symbol include/btkb.h declares; without a GPU the product
refuses to run (no CPU fallback).  No compute calls here."""
import ctypes as ct
import os
import re

import pytest

from conftest import ROOT

LIB = os.path.join(ROOT, "distant_speech_recognition_b200", "libbtkb.so")
HDR = os.path.join(ROOT, "include", "btkb.h")


def _declared():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(btkb_[a-z_0-9]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import __graft_entry__
        __graft_entry__.build()
    return ct.CDLL(LIB)


def test_header_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "libbtkb.so does not export %s declared in include/btkb.h" % n


def test_every_entry_point_has_a_python_binding():
    """The ctypes layer (the host side the tests and bench.py drive) reaches every function include/btkb.h declares."""
    src = open(os.path.join(ROOT, "distant_speech_recognition_b200", "_capi.py")).read()
    missing = [n for n in _declared() if ("lib.%s(" % n) not in src and ("lib.%s." % n) not in src]
    assert not missing, missing


def test_no_cpu_fallback(lib):
    from distant_speech_recognition_b200 import _capi
    if _capi.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_capi.BtkbError) as ei:
        _capi.Pipeline(8, 512)
    assert ei.value.code == _capi.ERR_NO_DEVICE
    assert "no CPU path" in str(ei.value)


def test_kernels_are_sm100a_with_tma(lib):
    """The shipped cubin is sm_100a and the per-bin kernel stages tiles with tensor-map TMA loads (SASS UTMALDG)."""
    import shutil, subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN4btkb8k_perbinILi8ELi1ELi0ELb0EEEv14CUtensorMap_stNS_10PerBinArgsE", LIB], capture_output=True, text=True).stdout
    assert "UTMALDG" in sass and "SYNCS" in sass
    assert "FFMA2" not in sass          # the scalar variant (BTKB_PERBIN_PACKED=0) has no packed instructions ...
    for fun in ("_ZN4btkb8k_perbinILi8ELi1ELi0ELb1EEEv14CUtensorMap_stNS_10PerBinArgsE", "_ZN4btkb12k_perbin_rlsILi8ELb1EEEv14CUtensorMap_stNS_10PerBinArgsE",
                "_ZN4btkb13k_analysis_r1ILi512ELi4ELi16ELi2ELb1ELb0EEEvNS_12AnalysisArgsE", "_ZN4btkb13k_analysis_r1ILi512ELi4ELi12ELi2ELb1ELb1EEEvNS_12AnalysisArgsE",
                "_ZN4btkb16k_synthesis_fastILi512ELi16ELi2ELi4ELi2ELb1EEEvNS_13SynthesisArgsE"):
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", fun, LIB], capture_output=True, text=True).stdout
        assert "FFMA2" in sass and "F32x2.LO_HI" in sass, fun    # ... and the packed variants (the defaults, incl. the 16-bit PCM analysis kernel) really are packed (csrc/btkb_f2.cuh)
