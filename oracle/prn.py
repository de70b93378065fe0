"""ORACLE (test infrastructure): PRN replica, follows amaranth_twstft/common.py:15-73.

``lfsr_bits`` uses the C restatement (oracle/prn_lfsr.c, built into
oracle/_build/) when present and otherwise a pure-Python loop that is a direct
restatement of common.py:23-30 (fine for short codes).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_prn.so")
_lib = None


def build_c(force=False):
    """gcc the C restatement into oracle/_build/liboracle_prn.so."""
    src = os.path.join(_HERE, "prn_lfsr.c")
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", _SO, src])
    return _SO


def _load():
    global _lib
    if _lib is None:
        try:
            build_c()
            _lib = ctypes.CDLL(_SO)
            _lib.oracle_lfsr_bits.restype = ctypes.c_uint32
            _lib.oracle_lfsr_bits.argtypes = [ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32,
                                              ctypes.c_size_t, ctypes.c_void_p]
        except Exception:  # no gcc: python loop
            _lib = False
    return _lib


def unary_xor(x):
    """common.py:15-21: XOR of all bits of x."""
    return bin(x).count("1") & 1


def nextstate(current, taps, bitlen):
    """common.py:23-30."""
    bit = unary_xor(current & taps)
    return (current >> 1) | (bit << (bitlen - 1))


def lfsr_bits_py(bitlen, taps, noiselen, seed=1):
    out = np.empty(noiselen, dtype=np.uint8)
    s = seed
    for i in range(noiselen):
        out[i] = s & 1
        s = nextstate(s, taps, bitlen)
    return out


def lfsr_bits(bitlen, taps, noiselen, seed=1):
    """uint8[noiselen] of 0/1: state%2 emitted before each step (common.py:64-66),
    restarted from `seed` every `noiselen` chips (synchronizer.py:70-75)."""
    lib = _load()
    if not lib:
        return lfsr_bits_py(bitlen, taps, noiselen, seed)
    out = np.empty(noiselen, dtype=np.uint8)
    lib.oracle_lfsr_bits(bitlen, taps, seed, noiselen, out.ctypes.data)
    return out


def prn_file_bytes(bitlen, noiselen, taps_a, taps_b=None, seed=1):
    """Content of the file written by common.py:59-73 (QPSK: a,b interleaved)."""
    a = lfsr_bits(bitlen, taps_a, noiselen, seed)
    if taps_b is None:
        return a
    b = lfsr_bits(bitlen, taps_b, noiselen, seed)
    out = np.empty(2 * noiselen, dtype=np.uint8)
    out[0::2] = a
    out[1::2] = b
    return out
