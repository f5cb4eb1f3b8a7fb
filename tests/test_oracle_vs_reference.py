"""CPU: differential test of the C restatement against the reference's own headers compiled in this container
(oracle/_ref/libswtpg_ref.so). Skipped where that library was not built (it needs /root/reference at build time)."""
import numpy as np
import pytest

import fdreadoutlibs_b200 as S
from oracle import binding as B
from util import assert_same_tps

pytestmark = pytest.mark.skipif(not B.reference_available(), reason="oracle/_ref/libswtpg_ref.so not built")


@pytest.mark.parametrize("seed,rate,thr", [(11, 0.02, 60), (12, 0.3, 20), (13, 0.9, 8), (14, 0.9, 0), (15, 0.05, 3)])
@pytest.mark.parametrize("algo,impl,flav", [(0, B.REF_ETH_SIMPLE_AVX2, 0), (0, B.REF_ETH_SIMPLE_NAIVE, 1), (1, B.REF_ETH_ABSRS_AVX2, 0),
                                            (2, B.REF_ETH_STDRS_AVX2, 0)])
def test_wibeth_matches_reference(seed, rate, thr, algo, impl, flav):
    fr = S.gen_wibeth_host(S.gen_params(seed, rate), 1, 150)[0]
    cfg = B.make_config(algorithm=algo, threshold=thr, rs_memory_factor=8, rs_scale_factor=5)
    o = B.Oracle(cfg, flav)
    r = B.ReferenceWibEth(impl, thr, 10, 8, 5)
    to, ped, _ = o.process(fr, dump=True)
    tr, pr = r.process(fr, dump=True)
    assert_same_tps(to, tr, f"algo {algo} impl {impl}")
    assert (ped[:, 63, :] == pr[:, 0, :]).all()          # pedestal after every frame
    assert (o.state()["accum"] == pr[-1, 1, :]).all()


@pytest.mark.parametrize("L", [1, 3, 10, 50, 0, -2])
def test_wibeth_accumulator_limits(L):
    """AVX2 SimpleThreshold honours the configured limit, including the degenerate L <= 0 cases of SURVEY A5."""
    fr = S.gen_wibeth_host(S.gen_params(21, 0.2), 1, 60)[0]
    o = B.Oracle(B.make_config(threshold=25, acc_limit=L))
    r = B.ReferenceWibEth(B.REF_ETH_SIMPLE_AVX2, 25, L, 8, 5)
    to, ped, _ = o.process(fr, dump=True)
    tr, pr = r.process(fr, dump=True)
    assert_same_tps(to, tr, f"L={L}")
    assert (ped[:, 63, :] == pr[:, 0, :]).all()


def test_wibeth_rs_memory_factor_per_channel():
    """Collection channels get R = 0 under enable_simple_threshold_on_collection (src/wibeth/WIBEthFrameProcessor.cpp:441-450)."""
    fr = S.gen_wibeth_host(S.gen_params(22, 0.2), 1, 80)[0]
    fac = np.where(np.arange(64) % 3 == 0, 0, 8).astype(np.uint16)
    o = B.Oracle(B.make_config(algorithm=1, threshold=30))
    o.set_memory_factor(fac)
    r = B.ReferenceWibEth(B.REF_ETH_ABSRS_AVX2, 30, 10, 8, 5)
    r.set_memory_factor(fac)
    assert_same_tps(o.process(fr), r.process(fr), "per-channel R")


@pytest.mark.parametrize("seed,rate", [(31, 0.05), (32, 0.6)])
@pytest.mark.parametrize("algo,impl,flav,thr", [(0, B.REF_WIB2_SIMPLE_AVX2, 0, 100), (0, B.REF_WIB2_SIMPLE_AVX2, 0, 30),
                                                (3, B.REF_WIB2_FIR_AVX2, 0, 5), (3, B.REF_WIB2_FIR_NAIVE, 1, 5), (3, B.REF_WIB2_FIR_AVX2, 0, 2),
                                                (1, B.REF_WIB2_ABSRS_AVX2, 0, 60), (1, B.REF_WIB2_ABSRS_AVX2, 0, 5), (1, B.REF_WIB2_ABSRS_AVX2, 0, 1),
                                                (1, B.REF_WIB2_ABSRS_AVX2, 0, 500)])
def test_wib2_matches_reference(seed, rate, algo, impl, flav, thr):
    sc = S.gen_wib2_host(S.gen_params(seed, rate), 1, 250)[0]
    o = B.Oracle(B.make_config(fmt="wib2", algorithm=algo, threshold=thr), flav)
    r = B.ReferenceWib2(impl, thr)
    tr, sd = r.process(sc, dump=True)
    assert_same_tps(o.process(sc), tr, f"wib2 algo {algo} impl {impl} thr {thr}")
    st = o.state()
    assert (st["pedestal"] == sd[-1, 0]).all()
    if algo in (1, 3):
        assert (st["quantile25"] == sd[-1, 1]).all() and (st["quantile75"] == sd[-1, 2]).all()


ANY_TAPS = [([2, 5, 11, 17, 9, 4, 1], 6), ([-3, 7, 20, 31, 20, 7, -3], 6), ([1, 3, 8, 10, 8, 3, 1], 5), ([300, -700, 1200, 2000, 1200, -700, 300], 6),
            ([0, 0, 0, 64, 0, 0, 0], 6)]


@pytest.mark.parametrize("taps,exponent", ANY_TAPS)
@pytest.mark.parametrize("impl,flav", [(B.REF_WIB2_FIR_AVX2, 0)])
def test_fir_arbitrary_taps_match_reference(taps, exponent, impl, flav):
    """The reference's ProcessingInfo takes any taps / exponent (its frame processor only ever passes firwin_int(7, 0.1, 64) and
    6): the oracle's multiply-add chain, its 16-bit wrapping (large taps) and the exponent-dependent clamps must follow it."""
    sc = S.gen_wib2_host(S.gen_params(34, 0.5), 1, 200)[0]
    o = B.Oracle(B.make_config(fmt="wib2", algorithm=3, threshold=5, fir_taps=taps, tap_exponent=exponent), flav)
    r = B.ReferenceWib2(impl, 5, fir_taps=taps, tap_exponent=exponent)
    tr, sd = r.process(sc, dump=True)
    assert_same_tps(o.process(sc), tr, f"taps {taps} exponent {exponent}")
    st = o.state()
    assert (st["pedestal"] == sd[-1, 0]).all() and (st["quantile25"] == sd[-1, 1]).all() and (st["quantile75"] == sd[-1, 2]).all()


def test_fir_threshold_64bit_lane_product():
    """`sigma * multiplier * threshold` is a 4 x int64 multiply; with a large threshold the product of one
    16-bit lane carries into its neighbour. The oracle must follow the AVX2 code there too."""
    sc = S.gen_wib2_host(S.gen_params(33, 0.4, noise_q8=40 * 256), 1, 200)[0]  # wide noise -> sigma at its clamp of 102
    for thr in (11, 40, 700):  # 102*64*11 = 71808 > 65535: lanes overflow into each other
        o = B.Oracle(B.make_config(fmt="wib2", algorithm=3, threshold=thr))
        r = B.ReferenceWib2(B.REF_WIB2_FIR_AVX2, thr)
        assert_same_tps(o.process(sc), r.process(sc), f"thr {thr}")


def test_wib2_naive_rs_is_a_different_algorithm(capfd):
    """wib2/tpg/ProcessNaiveRS.hpp:22-228 is wrapped like every other processor of the reference, but it is not a scalar twin of
    wib2/tpg/ProcessRSAVX2.hpp: float running sum with R = 0.8 and scale 2 (:31-33,:122-130), quartiles of the RUNNING SUM
    (:146-152), threshold hard-wired to 5 * sigma (:175), charge = sum of the pedestal-subtracted SAMPLE >> exponent (:179).
    So there is nothing to pin to it; this test documents by how much the two disagree on ordinary input, so that nobody
    mistakes it for a parity target (the CUDA path follows the AVX2 processor, like every production configuration)."""
    sc = S.gen_wib2_host(S.gen_params(5, 0.05), 1, 100)[0]
    avx = B.ReferenceWib2(B.REF_WIB2_ABSRS_AVX2, threshold=5).process(sc)
    naive = B.ReferenceWib2(B.REF_WIB2_ABSRS_NAIVE, threshold=5).process(sc)
    capfd.readouterr()  # the header prints "Found N hits" per call
    assert avx.size == 228 and naive.size == 202
    key = lambda t: set(zip(t["channel"].tolist(), t["time_start"].tolist()))
    assert len(key(avx) & key(naive)) == 16  # 7 % of the hits share channel and start time; none of those shares its charge
    both = {k: None for k in key(avx) & key(naive)}
    ca = {(c, t): q for c, t, q in zip(avx["channel"].tolist(), avx["time_start"].tolist(), avx["adc_integral"].tolist())}
    cn = {(c, t): q for c, t, q in zip(naive["channel"].tolist(), naive["time_start"].tolist(), naive["adc_integral"].tolist())}
    assert all(ca[k] != cn[k] for k in both)
