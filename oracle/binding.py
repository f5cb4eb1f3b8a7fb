"""TEST INFRASTRUCTURE — ctypes bindings of the two CPU checkers.

* `Oracle`     : oracle/liboracle.so, the plain-C restatement (oracle/swtpg_oracle.c). Always available.
* `Reference*` : oracle/_ref/libswtpg_ref.so, the reference's own headers compiled in place (oracle/ref_wrapper.cpp).
                 Present where it was built (`make -C oracle ref` needs /root/reference); travels prebuilt to GPU boxes.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from fdreadoutlibs_b200 import frames as F
from fdreadoutlibs_b200._lib import SwtpgConfig

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_PATH = os.path.join(_HERE, "liboracle.so")
REF_PATH = os.path.join(_HERE, "_ref", "libswtpg_ref.so")

FLAVOUR_AVX2, FLAVOUR_NAIVE = 0, 1


class _OracleChan(C.Structure):
    _fields_ = [(n, C.c_int16) for n in ("median", "accum", "q25", "q75", "a25", "a75", "rs", "median_rs", "accum_rs")] + \
               [(n, C.c_uint16) for n in ("rs_factor", "prev_over", "charge", "tover", "peak_adc", "peak_time")] + \
               [("ring", C.c_int16 * 8)]


class _OracleLink(C.Structure):
    _fields_ = [("cfg", SwtpgConfig), ("flavour", C.c_int), ("n_channels", C.c_int), ("initialized", C.c_int), ("k0", C.c_uint),
                ("ch", _OracleChan * 256)]


def build_oracle():
    subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)


def _load_oracle():
    if not os.path.exists(ORACLE_PATH):
        build_oracle()
    lib = C.CDLL(ORACLE_PATH)
    lib.oracle_link_init.argtypes = [C.POINTER(_OracleLink), C.POINTER(SwtpgConfig), C.c_int]
    lib.oracle_link_set_memory_factor.argtypes = [C.POINTER(_OracleLink), C.c_void_p]
    lib.oracle_process.restype = C.c_long
    lib.oracle_process.argtypes = [C.POINTER(_OracleLink), C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    lib.oracle_unpack14.restype = C.c_uint16
    lib.oracle_unpack14.argtypes = [C.c_void_p, C.c_uint]
    lib.oracle_wibeth_expand.argtypes = [C.c_void_p, C.c_void_p]
    lib.oracle_wib2_expand.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p]
    lib.oracle_firwin_int.restype = C.c_int
    lib.oracle_firwin_int.argtypes = [C.c_int, C.c_double, C.c_int, C.c_void_p]
    lib.oracle_get_state.argtypes = [C.POINTER(_OracleLink), C.c_void_p]
    return lib


_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        _oracle = _load_oracle()
    return _oracle


def make_config(fmt="wibeth", algorithm=0, threshold=60, acc_limit=10, rs_memory_factor=8, rs_scale_factor=5, fir_taps=None,
                tap_exponent=6, wib2_adc_offset=0) -> SwtpgConfig:
    cfg = SwtpgConfig()
    cfg.struct_size = C.sizeof(SwtpgConfig)
    cfg.format = 1 if fmt == "wib2" else 0
    cfg.algorithm = algorithm
    cfg.n_links = 1
    cfg.max_units = 1
    cfg.threshold = threshold
    cfg.frugal_acc_limit = acc_limit
    cfg.rs_memory_factor = rs_memory_factor
    cfg.rs_scale_factor = rs_scale_factor
    if fir_taps is not None:
        for i, t in enumerate(fir_taps):
            cfg.fir_taps[i] = int(t)
    cfg.tap_exponent = tap_exponent
    cfg.wib2_adc_offset = wib2_adc_offset
    return cfg


class Oracle:
    """One link's worth of carried state + the scalar restatement of the algorithm."""

    def __init__(self, cfg: SwtpgConfig, flavour: int = FLAVOUR_AVX2, link_id: int = 0):
        self.lib = oracle_lib()
        self.link = _OracleLink()
        self.lib.oracle_link_init(C.byref(self.link), C.byref(cfg), flavour)
        self.link_id = link_id
        self.wib2 = cfg.format == 1
        self.channels = 256 if self.wib2 else 64
        self.ticks = 12 if self.wib2 else 64
        self.unit_bytes = F.WIB2_SUPERCHUNK_BYTES if self.wib2 else F.WIBETH_FRAME_BYTES

    def set_memory_factor(self, by_channel):
        a = np.ascontiguousarray(by_channel, dtype=np.uint16)
        assert a.size == self.channels
        self.lib.oracle_link_set_memory_factor(C.byref(self.link), a.ctypes.data)

    def process(self, units: np.ndarray, cap: int = 1 << 18, dump: bool = False):
        units = np.ascontiguousarray(units, dtype=np.uint8)
        n_units = units.size // self.unit_bytes
        out = np.zeros(cap, dtype=F.TP_DTYPE)
        ped = wav = None
        if dump:
            ped = np.zeros((n_units, self.ticks, self.channels), dtype=np.int16)
            wav = np.zeros_like(ped)
        n = self.lib.oracle_process(C.byref(self.link), units.ctypes.data, n_units, self.link_id, out.ctypes.data, cap,
                                    None if ped is None else ped.ctypes.data, None if wav is None else wav.ctypes.data)
        if n > cap:
            raise RuntimeError(f"oracle produced {n} TPs > cap {cap}")
        return (out[:n].copy(), ped, wav) if dump else out[:n].copy()

    def state(self) -> np.ndarray:
        out = np.zeros(self.channels, dtype=F.STATE_DTYPE)
        self.lib.oracle_get_state(C.byref(self.link), out.ctypes.data)
        return out


def oracle_process_links(cfg: SwtpgConfig, frames: np.ndarray, n_units=None, flavour: int = FLAVOUR_AVX2, oracles=None, cap_per_link: int = 1 << 18):
    """frames [n_links, stride, unit_bytes] -> concatenated TPs of all links (link field = index). Returns (tps, oracles)."""
    n_links = frames.shape[0]
    if oracles is None:
        oracles = [Oracle(cfg, flavour, l) for l in range(n_links)]
    parts = []
    for l in range(n_links):
        nu = frames.shape[1] if n_units is None else int(n_units[l])
        parts.append(oracles[l].process(frames[l, :nu], cap=cap_per_link))
    return np.concatenate(parts) if parts else np.zeros(0, dtype=F.TP_DTYPE), oracles


# ---- the reference's own code ---------------------------------------------------------------------------------------------
def reference_available() -> bool:
    return os.path.exists(REF_PATH)


_ref = None


def ref_lib():
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_PATH)
        lib.ref_wibeth_create.restype = C.c_void_p
        lib.ref_wibeth_create.argtypes = [C.c_int, C.c_uint16, C.c_int16, C.c_uint16, C.c_uint16]
        lib.ref_wibeth_destroy.argtypes = [C.c_void_p]
        lib.ref_wibeth_set_memory_factor.argtypes = [C.c_void_p, C.c_void_p]
        lib.ref_wibeth_process.restype = C.c_long
        lib.ref_wibeth_process.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p]
        lib.ref_wibeth_expand.argtypes = [C.c_void_p, C.c_void_p]
        lib.ref_wib2_create.restype = C.c_void_p
        lib.ref_wib2_create.argtypes = [C.c_int, C.c_uint16, C.c_int]
        lib.ref_wib2_create_taps.restype = C.c_void_p
        lib.ref_wib2_create_taps.argtypes = [C.c_int, C.c_uint16, C.c_int, C.c_void_p, C.c_int]
        lib.ref_wib2_destroy.argtypes = [C.c_void_p]
        lib.ref_wib2_process.restype = C.c_long
        lib.ref_wib2_process.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p]
        lib.ref_wib2_expand.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.ref_firwin_int.restype = C.c_int
        lib.ref_firwin_int.argtypes = [C.c_int, C.c_double, C.c_int, C.c_void_p]
        lib.ref_wibeth_bench.restype = C.c_double
        lib.ref_wibeth_bench.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_uint16, C.c_int16, C.c_int,
                                         C.POINTER(C.c_uint64)]
        _ref = lib
    return _ref


# impl codes of oracle/ref_wrapper.cpp
REF_ETH_SIMPLE_AVX2, REF_ETH_SIMPLE_NAIVE, REF_ETH_ABSRS_AVX2, REF_ETH_STDRS_AVX2 = range(4)
REF_WIB2_SIMPLE_AVX2, REF_WIB2_FIR_AVX2, REF_WIB2_FIR_NAIVE, REF_WIB2_ABSRS_AVX2, REF_WIB2_ABSRS_NAIVE = range(5)


class ReferenceWibEth:
    def __init__(self, impl=REF_ETH_SIMPLE_AVX2, threshold=60, acc_limit=10, memory_factor=8, scale_factor=5, link_id=0):
        self.lib = ref_lib()
        self.h = self.lib.ref_wibeth_create(impl, threshold, acc_limit, memory_factor, scale_factor)
        self.link_id = link_id

    def set_memory_factor(self, by_channel):
        a = np.ascontiguousarray(by_channel, dtype=np.uint16)
        self.lib.ref_wibeth_set_memory_factor(self.h, a.ctypes.data)

    def process(self, frames: np.ndarray, cap: int = 1 << 18, dump: bool = False):
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        n = frames.size // F.WIBETH_FRAME_BYTES
        out = np.zeros(cap, dtype=F.TP_DTYPE)
        ped = np.zeros((n, 2, 64), dtype=np.int16) if dump else None
        k = self.lib.ref_wibeth_process(self.h, frames.ctypes.data, n, self.link_id, out.ctypes.data, cap, None if ped is None else ped.ctypes.data)
        if k < 0:
            raise RuntimeError("reference TP buffer overflow")
        return (out[:k].copy(), ped) if dump else out[:k].copy()

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_wibeth_destroy(self.h)
            self.h = None


class ReferenceWib2:
    """Both register selectors of one link (two WIB2FrameHandlers, src/wib2/WIB2FrameProcessor.cpp:224-225)."""

    def __init__(self, impl=REF_WIB2_SIMPLE_AVX2, threshold=60, link_id=0, fir_taps=None, tap_exponent=6):
        self.lib = ref_lib()
        if fir_taps is None:
            self.h = [self.lib.ref_wib2_create(impl, threshold, sel) for sel in (0, 1)]
        else:
            taps = np.zeros(8, dtype=np.int16)
            taps[: len(fir_taps)] = fir_taps
            self.h = [self.lib.ref_wib2_create_taps(impl, threshold, sel, taps.ctypes.data, tap_exponent) for sel in (0, 1)]
        self.link_id = link_id

    def process(self, superchunks: np.ndarray, cap: int = 1 << 18, dump: bool = False):
        sc = np.ascontiguousarray(superchunks, dtype=np.uint8)
        n = sc.size // F.WIB2_SUPERCHUNK_BYTES
        parts, dumps = [], []
        for sel in (0, 1):
            out = np.zeros(cap, dtype=F.TP_DTYPE)
            st = np.zeros((n, 3, 128), dtype=np.int16) if dump else None
            k = self.lib.ref_wib2_process(self.h[sel], sc.ctypes.data, n, self.link_id, out.ctypes.data, cap, None if st is None else st.ctypes.data)
            if k < 0:
                raise RuntimeError("reference TP buffer overflow")
            parts.append(out[:k].copy())
            dumps.append(st)
        tps = np.concatenate(parts)
        return (tps, np.concatenate(dumps, axis=2)) if dump else tps

    def __del__(self):
        for h in getattr(self, "h", []):
            if h:
                self.lib.ref_wib2_destroy(h)
        self.h = []


def ref_wibeth_bench(frames: np.ndarray, n_threads: int, impl=REF_ETH_SIMPLE_AVX2, threshold=60, acc_limit=10, reps=3):
    """frames [n_links, n_frames, 7200]. Returns (best seconds, TPs in one pass)."""
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    n_links, n_frames = frames.shape[0], frames.shape[1]
    ntp = C.c_uint64(0)
    sec = ref_lib().ref_wibeth_bench(frames.ctypes.data, n_links, n_frames, n_threads, impl, threshold, acc_limit, reps, C.byref(ntp))
    return sec, ntp.value
