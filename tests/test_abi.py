"""CPU: the C-ABI shared library loads, exports every symbol the headers declare, has the documented record layouts,
and refuses to compute without a GPU (no CPU fallback). Host-only helpers (generator, sort, merge, taps) are exercised."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import fdreadoutlibs_b200 as S
from fdreadoutlibs_b200 import _lib, framegen
from fdreadoutlibs_b200 import frames as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(hdr):
    text = open(os.path.join(ROOT, "include", hdr)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(swtpg_[a-z0-9_]+)\s*\(", text))


def test_every_declared_symbol_is_exported_and_bound():
    """include/swtpg.h <-> libswtpg_b200.so, include/swtpg_framegen.h <-> libswtpg_framegen.so (the generator is a test / bench
    utility in its own library, so that CPU-side checkers never map the CUDA product library)."""
    names = declared_symbols("swtpg.h")
    assert len(names) >= 28
    for n in sorted(names):
        assert hasattr(_lib.lib, n), f"{n} declared in include/swtpg.h but not exported by libswtpg_b200.so"
    assert names == set(_lib.EXPORTS), f"binding and header differ: {names ^ set(_lib.EXPORTS)}"
    gen = declared_symbols("swtpg_framegen.h")
    for n in sorted(gen):
        assert hasattr(framegen.lib, n), f"{n} declared in include/swtpg_framegen.h but not exported by libswtpg_framegen.so"
        assert not hasattr(_lib.lib, n), f"{n}: the generator must not live in the product library"
    assert gen == set(framegen.EXPORTS), f"binding and header differ: {gen ^ set(framegen.EXPORTS)}"


def test_abi_version_and_layouts():
    assert _lib.lib.swtpg_abi_version() == 2
    assert F.TP_DTYPE.itemsize == 32
    assert C.sizeof(_lib.SwtpgConfig) == 72  # static_assert'ed on the C side (swtpg_capi.cu)
    assert C.sizeof(framegen.GenParams) == 32
    assert F.STATE_DTYPE.itemsize == 48
    assert _lib.lib.swtpg_status_string(3) == b"busy (back-pressure)"


@pytest.mark.skipif(S.device_available(), reason="this check is for GPU-less hosts")
def test_no_cpu_fallback():
    with pytest.raises(S.SwtpgError) as e:
        S.TPGenerator(1, 4)
    assert e.value.status == _lib.SWTPG_ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_create_rejects_bad_config():
    cfg = _lib.SwtpgConfig()
    h = C.c_void_p()
    assert _lib.lib.swtpg_create(C.byref(cfg), C.byref(h)) == _lib.SWTPG_ERR_INVALID_ARG  # struct_size == 0
    cfg.struct_size = C.sizeof(cfg)
    assert _lib.lib.swtpg_create(C.byref(cfg), C.byref(h)) == _lib.SWTPG_ERR_INVALID_ARG  # n_links == 0
    cfg.n_links, cfg.max_units, cfg.algorithm = 1, 1, 17
    assert _lib.lib.swtpg_create(C.byref(cfg), C.byref(h)) == _lib.SWTPG_ERR_UNSUPPORTED  # TPGAlgorithmInexistent
    with pytest.raises(S.TPGAlgorithmInexistent):
        S.TPGenerator(1, 1, algorithm="NoSuchAlgo")


def test_generator_is_deterministic_and_sharding_invariant():
    p = S.gen_params(7, 0.1)
    a = S.gen_wibeth_host(p, 6, 5, n_threads=1)
    b = S.gen_wibeth_host(p, 6, 5, n_threads=4)
    assert (a == b).all()
    part = S.gen_wibeth_host(p, 2, 3, link0=3, unit0=2)
    assert (part == a[3:5, 2:5]).all()  # any (link, unit) window reproduces the same bytes
    adc, ts = F.unpack_wibeth_frames(a[1])
    assert list(ts) == [(1 << 40) + 2048 * i for i in range(5)]
    assert 850 < adc.mean() < 1700 and adc.max() <= 16383
    w = S.gen_wib2_host(p, 2, 4)
    adc2, ts2 = F.unpack_wib2_superchunks(w[1])
    assert list(ts2[:3]) == [1 << 40, (1 << 40) + 32, (1 << 40) + 64] and adc2.shape == (48, 256)
    # WIB2 link l channel c tick t is the same waveform as global channel 256 l + c
    assert (S.gen_wib2_host(p, 1, 4, link0=1)[0] == w[1]).all()


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(0)
    adc = rng.integers(0, 16384, size=(3, 64, 64), dtype=np.uint16)
    fr = F.pack_wibeth_frames(adc, 12345)
    back, ts = F.unpack_wibeth_frames(fr)
    assert (back == adc).all() and list(ts) == [12345, 12345 + 2048, 12345 + 4096]
    adc2 = rng.integers(0, 16384, size=(24, 256), dtype=np.uint16)
    sc = F.pack_wib2_superchunks(adc2, 99)
    back2, ts2 = F.unpack_wib2_superchunks(sc)
    assert (back2 == adc2).all() and ts2[13] == 99 + 32 * 13


def test_sort_and_merge():
    rng = np.random.default_rng(1)
    tps = np.zeros(5000, dtype=F.TP_DTYPE)
    tps["time_start"] = rng.integers(0, 400, 5000)
    tps["link"] = rng.integers(0, 8, 5000)
    tps["channel"] = rng.integers(0, 64, 5000)
    tps["adc_integral"] = rng.integers(1, 60000, 5000)
    s = S.sort_tps(tps)
    key = list(zip(s["time_start"].tolist(), s["link"].tolist(), s["channel"].tolist()))
    assert key == sorted(key)
    parts = [S.sort_tps(tps[tps["link"] % 3 == r]) for r in range(3)]
    merged = S.merge_sorted(parts)
    assert merged.size == tps.size
    assert (F.sort_tps(merged) == F.sort_tps(tps)).all()
    mk = list(zip(merged["time_start"].tolist(), merged["link"].tolist(), merged["channel"].tolist()))
    assert mk == sorted(mk)
    assert S.merge_sorted([np.zeros(0, dtype=F.TP_DTYPE), parts[0]]).size == parts[0].size


@pytest.mark.parametrize("n,wide", [(3000, False), (120000, False), (120000, True)])
def test_sort_and_merge_match_a_lexicographic_sort(n, wide):
    """swtpg_sort_tps (std::stable_sort below 4096 records, a radix sort of packed keys above, std::stable_sort again when the keys
    do not fit 64 bits) and swtpg_merge_sorted order by (time_start, link, channel, time_over_threshold, adc_integral) — with
    plenty of ties on the first three — exactly like numpy's lexsort, and the merge of per-GPU lists equals the sort of the union."""
    rng = np.random.default_rng(n + wide)
    tps = np.zeros(n, dtype=F.TP_DTYPE)
    tps["time_start"] = 79554162068719943 + 32 * rng.integers(0, 5000, n)
    if wide:
        tps["time_start"][::3] += np.uint64(1) << np.uint64(62)  # key range beyond 64 bits together with link and channel
        tps["link"][::5] = 0xFFFFFFF0
    tps["link"] += rng.integers(0, 6000, n).astype(np.uint32)
    tps["channel"] = rng.integers(0, 256, n)
    tps["time_over_threshold"] = 32 * rng.integers(1, 3, n)
    tps["adc_integral"] = rng.integers(1, 4, n)
    tps["time_peak"] = rng.integers(0, 1 << 40, n)
    tps["adc_peak"] = rng.integers(0, 1 << 14, n)
    order = np.lexsort((tps["adc_integral"], tps["time_over_threshold"], tps["channel"], tps["link"], tps["time_start"]))
    want = tps[order]
    got = S.sort_tps(tps.copy())
    keys = ("time_start", "link", "channel", "time_over_threshold", "adc_integral")
    for f in keys:
        assert (got[f] == want[f]).all(), f
    assert (F.sort_tps(got) == F.sort_tps(tps)).all()  # same multiset of whole records
    parts = [S.sort_tps(tps[r::5].copy()) for r in range(5)]
    merged = S.merge_sorted(parts)
    for f in keys:
        assert (merged[f] == want[f]).all(), f
    assert (F.sort_tps(merged) == F.sort_tps(tps)).all()


@pytest.mark.parametrize("k", [1, 2, 3, 4, 7, 8])
def test_kway_merge_is_the_total_order(k):
    """swtpg_merge_sorted (tree of two-way merges) over k lists, some of them empty, equals numpy's lexsort of the union on ALL
    seven fields — the order is total, so records that tie on (time_start, link, channel) come out the same way whatever order
    the device emitted them in."""
    rng = np.random.default_rng(k)
    n = 30000
    tps = np.zeros(n, dtype=F.TP_DTYPE)
    tps["time_start"] = 10**15 + 32 * rng.integers(0, 300, n)
    tps["link"] = rng.integers(0, 8, n)
    tps["channel"] = rng.integers(0, 64, n)
    tps["time_over_threshold"] = 32 * rng.integers(1, 3, n)
    tps["adc_integral"] = rng.integers(1, 3, n)
    tps["time_peak"] = tps["time_start"] + 32 * rng.integers(0, 3, n).astype(np.uint64)
    tps["adc_peak"] = rng.integers(0, 3, n)
    order = np.lexsort((tps["adc_peak"], tps["time_peak"], tps["adc_integral"], tps["time_over_threshold"], tps["channel"], tps["link"], tps["time_start"]))
    want = tps[order]
    owner = rng.integers(0, k, n)
    owner[owner == k // 2] = 0 if k > 2 else owner[owner == k // 2]  # an empty list in the middle when there are enough lists
    parts = [S.sort_tps(tps[owner == r].copy()) for r in range(k)]
    merged = S.merge_sorted(parts)
    assert merged.size == n and (merged == want).all()
    assert (S.sort_tps(tps.copy()) == want).all()


def test_firwin_int_host():
    assert list(S.firwin_int(7, 0.1, 64)) == [1, 6, 15, 20, 15, 6, 1]


def test_header_is_plain_c_and_links(tmp_path):
    """include/swtpg.h is the boundary a C / cgo / JNI / ctypes binding sees: it must compile as C11 and as C++17 without any
    CUDA or C++ header, and a C program must link against the shared library and reach the host-only entry points."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "abi_probe.c"
    src.write_text(
        '#include "swtpg.h"\n#include "swtpg_framegen.h"\n#include <stdio.h>\n'
        "int main(void) {\n"
        "  int16_t taps[8] = {0};\n"
        "  swtpg_config cfg = {0};\n"
        "  cfg.struct_size = sizeof cfg;\n"
        "  if (swtpg_abi_version() != SWTPG_ABI_VERSION) return 1;\n"
        "  if (swtpg_firwin_int(7, 0.1, 64, taps) != 7 || taps[3] != 20) return 2;\n"
        "  if (sizeof(swtpg_tp) != 32 || sizeof cfg != 72) return 3;\n"
        '  printf("%s\\n", swtpg_status_string(SWTPG_ERR_BUSY));\n'
        "  return 0;\n}\n")
    lib_dir = os.path.join(root, "fdreadoutlibs_b200")
    exe = tmp_path / "abi_probe"
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(root, "include"), str(src), "-o", str(exe),
                    "-L", lib_dir, "-lswtpg_b200", f"-Wl,-rpath,{lib_dir}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert out.strip() != ""
    cpp = tmp_path / "abi_probe.cpp"
    cpp.write_text('#include "swtpg.h"\n#include "swtpg_framegen.h"\nint f() { return int(swtpg_abi_version()); }\n')
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-c", "-I", os.path.join(root, "include"), str(cpp), "-o", str(tmp_path / "abi_probe.o")],
                   check=True)
