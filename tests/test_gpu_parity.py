"""GPU parity tests proper: the CUDA library through its C ABI against the oracle."""
import json
import os

import numpy as np
import pytest

import parity_common as pc
from amaranth_twstft_b200 import synth
from amaranth_twstft_b200.api import Correlator
from oracle import correlator as oc
from oracle import prn

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def test_backend_is_cuda():
    with Correlator(nsamples=10000, max_batch=1, lib=pc.CUDA_LIB) as cor:
        assert cor.backend == "cuda"


@pytest.mark.parametrize("entry", json.load(open(os.path.join(G, "prn_digests.json"))),
                         ids=lambda e: os.path.basename(e["file"]))
def test_gpu_lfsr_matches_reference_files(entry):
    import ctypes
    import hashlib
    from amaranth_twstft_b200 import _lib
    lib = _lib.load(pc.CUDA_LIB)
    n = entry["noiselen"]
    out = np.zeros(n * (2 if entry["taps_b"] else 1), dtype=np.uint8)
    rc = lib.twstft_prn_generate(0, entry["bitlen"], entry["taps_a"], entry["taps_b"] or 0, 1, n, out.ctypes.data)
    assert rc == 0
    assert hashlib.sha256(out.tobytes()).hexdigest() == entry["sha256"]


@pytest.mark.parametrize("n", [10000, 200000, 250000, 1000000])
def test_variant_g_small(n):
    pc.run_variant_g(pc.CUDA_LIB, n, seeds=[11, 12, 13])


def test_variant_g_full_size():
    """configs[1]: two-channel 1-s records at 5 MS/s, 2.5 Mchip/s 22-bit code."""
    pc.run_variant_g(pc.CUDA_LIB, 5000000, seeds=[20261017, 20261018], max_batch=2)
