"""C-ABI surface: the product library loads, exports every symbol include/twstft.h declares, and
fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re
import shutil

import pytest

from amaranth_twstft_b200 import _lib
from amaranth_twstft_b200.api import Correlator, TwstftError
import parity_common as pc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "twstft.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(twstft_[a-z0-9_]+)\s*\(", src)))


def have_gpu():
    if not shutil.which("nvidia-smi"):
        return False
    return os.system("nvidia-smi -L > /dev/null 2>&1") == 0


def test_header_declares_what_python_binds():
    assert set(_lib.EXPORTS) == set(header_symbols())


@pytest.mark.parametrize("path", [pc.CUDA_LIB, pc.EMU_LIB], ids=["cuda", "emu"])
def test_library_exports_every_declared_symbol(path):
    if not os.path.exists(path):
        pytest.skip(f"{path} not built (run __graft_entry__.build())")
    lib = C.CDLL(path)
    for sym in header_symbols():
        assert hasattr(lib, sym), f"{os.path.basename(path)} does not export {sym}"


def test_product_library_is_cuda_and_emulation_is_separate():
    if not os.path.exists(pc.CUDA_LIB):
        pytest.skip("product library not built")
    assert _lib.load(pc.CUDA_LIB).twstft_backend() == b"cuda"
    assert os.path.dirname(pc.EMU_LIB).startswith(os.path.join(ROOT, "tests"))
    # the product package must not know about the emulation build or the oracle
    for fn in os.listdir(os.path.join(ROOT, "amaranth_twstft_b200")):
        if fn.endswith(".py"):
            text = open(os.path.join(ROOT, "amaranth_twstft_b200", fn)).read()
            assert "oracle" not in text and "emu" not in text.lower().replace("enumerate", ""), fn


def test_missing_library_raises():
    with pytest.raises(_lib.LibraryMissing):
        _lib.load("/nonexistent/libtwstft.so")


def test_no_gpu_fails_loudly():
    if have_gpu():
        pytest.skip("a GPU is present")
    if not os.path.exists(pc.CUDA_LIB):
        pytest.skip("product library not built")
    with pytest.raises(TwstftError) as ei:
        Correlator(nsamples=10000, max_batch=1, lib=pc.CUDA_LIB)
    assert "no CPU fallback" in str(ei.value)
    out = (C.c_uint8 * 16)()
    assert _lib.load(pc.CUDA_LIB).twstft_prn_generate(0, 13, 27, 0, 1, 16, out) == -2


def test_bad_arguments():
    with pytest.raises(TwstftError):
        Correlator(nsamples=12345, max_batch=1, lib=pc.EMU_LIB)          # unsupported length
    with pytest.raises(TwstftError):
        Correlator(nsamples=10000, nint=2, max_batch=1, lib=pc.EMU_LIB)  # Nint>1 unsupported
    with Correlator(nsamples=10000, max_batch=1, lib=pc.EMU_LIB) as cor:
        import numpy as np
        with pytest.raises(TwstftError):
            cor.process(np.zeros(4 * 10000, dtype=np.int16))             # no code installed
        with pytest.raises(TwstftError):
            cor.set_code(np.zeros(123, dtype=np.uint8))                  # wrong code length
        with pytest.raises(ValueError):
            cor.set_code(np.zeros(5000, dtype=np.uint8)) or cor.process(np.zeros(17, dtype=np.int16))
