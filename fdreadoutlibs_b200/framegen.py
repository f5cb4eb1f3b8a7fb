"""Synthetic WIBEth / WIB2 frames (include/swtpg_framegen.h, libswtpg_framegen.so): test / benchmark utility, NOT part of the
product path. Kept in its own library and module so that the CPU checkers (the oracle tests, bench.py's reference arm) can
generate frames without mapping the CUDA product library: this module imports nothing else of the package — load it on its own
with `importlib.util.spec_from_file_location` where even the package import is unwanted."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libswtpg_framegen.so")
WIBETH_FRAME_BYTES, WIB2_SUPERCHUNK_BYTES = 7200, 5664


class GenParams(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64),
        ("noise_q8", C.c_uint32),
        ("pulse_prob_q32", C.c_uint32),
        ("amp_min", C.c_uint16), ("amp_max", C.c_uint16),
        ("hw_min", C.c_uint16), ("hw_max", C.c_uint16),
        ("ped_base", C.c_uint16), ("ped_step", C.c_uint16), ("ped_mod", C.c_uint16),
        ("bipolar", C.c_uint16),
    ]


EXPORTS = {
    "swtpg_gen_default_params": (None, [C.POINTER(GenParams), C.c_uint64, C.c_double]),
    "swtpg_gen_wibeth_host": (C.c_int, [C.POINTER(GenParams), C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint64, C.c_void_p, C.c_int]),
    "swtpg_gen_wib2_host": (C.c_int, [C.POINTER(GenParams), C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_void_p, C.c_int]),
    "swtpg_gen_wibeth_device": (C.c_int, [C.POINTER(GenParams), C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p]),
    "swtpg_gen_wib2_device": (C.c_int, [C.POINTER(GenParams), C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p]),
}


def load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `make lib` (or __graft_entry__.build())")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = load()


class FramegenError(RuntimeError):
    pass


def gen_params(seed: int = 1, pulses_per_64_ticks: float = 0.02, **overrides) -> GenParams:
    p = GenParams()
    lib.swtpg_gen_default_params(C.byref(p), seed, pulses_per_64_ticks)
    for k, v in overrides.items():
        setattr(p, k, v)
    return p


def gen_wibeth_host(p: GenParams, n_links: int, n_units: int, link0: int = 0, unit0: int = 0, ts0: int = 1 << 40, n_threads: int = 8,
                    out: "np.ndarray | None" = None) -> np.ndarray:
    if out is None:
        out = np.zeros((n_links, n_units, WIBETH_FRAME_BYTES), dtype=np.uint8)
    assert out.dtype == np.uint8 and out.flags["C_CONTIGUOUS"] and out.size == n_links * n_units * WIBETH_FRAME_BYTES
    if lib.swtpg_gen_wibeth_host(C.byref(p), link0, n_links, unit0, n_units, ts0, out.ctypes.data, n_threads) != 0:
        raise FramegenError("swtpg_gen_wibeth_host")
    return out.reshape(n_links, n_units, WIBETH_FRAME_BYTES)


def gen_wib2_host(p: GenParams, n_links: int, n_units: int, link0: int = 0, unit0: int = 0, ts0: int = 1 << 40, adc_offset: int = 0,
                  n_threads: int = 8) -> np.ndarray:
    out = np.zeros((n_links, n_units, WIB2_SUPERCHUNK_BYTES), dtype=np.uint8)
    if lib.swtpg_gen_wib2_host(C.byref(p), link0, n_links, unit0, n_units, ts0, adc_offset, out.ctypes.data, n_threads) != 0:
        raise FramegenError("swtpg_gen_wib2_host")
    return out


def gen_wibeth_device(p: GenParams, d_ptr: int, n_links: int, n_units: int, link0: int = 0, unit0: int = 0, ts0: int = 1 << 40, stream: int = 0):
    if lib.swtpg_gen_wibeth_device(C.byref(p), link0, n_links, unit0, n_units, ts0, C.c_void_p(d_ptr), C.c_void_p(stream)) != 0:
        raise FramegenError("swtpg_gen_wibeth_device")


def gen_wib2_device(p: GenParams, d_ptr: int, n_links: int, n_units: int, link0: int = 0, unit0: int = 0, ts0: int = 1 << 40, adc_offset: int = 0,
                    stream: int = 0):
    if lib.swtpg_gen_wib2_device(C.byref(p), link0, n_links, unit0, n_units, ts0, adc_offset, C.c_void_p(d_ptr), C.c_void_p(stream)) != 0:
        raise FramegenError("swtpg_gen_wib2_device")
