import numpy as np

from fdreadoutlibs_b200 import frames as F

FIELDS = ("time_start", "time_peak", "time_over_threshold", "adc_integral", "adc_peak", "channel", "link")


def assert_same_tps(got: np.ndarray, want: np.ndarray, what: str = ""):
    """Bit-exact comparison of two TP lists as sorted field tuples."""
    g, w = F.sort_tps(np.asarray(got, dtype=F.TP_DTYPE)), F.sort_tps(np.asarray(want, dtype=F.TP_DTYPE))
    assert g.size == w.size, f"{what}: {g.size} TPs, expected {w.size}"
    for f in FIELDS:
        bad = np.nonzero(g[f] != w[f])[0]
        assert bad.size == 0, f"{what}: field {f} differs at sorted index {bad[0]}: got {g[bad[0]]}, want {w[bad[0]]}"


def oracle_config(case, B):
    from fdreadoutlibs_b200.api import ALGORITHMS

    return B.make_config(fmt=case["fmt"], algorithm=ALGORITHMS[case["algorithm"]], threshold=case["threshold"],
                         acc_limit=case.get("acc_limit", 10), rs_memory_factor=case.get("rs_memory_factor", 8),
                         rs_scale_factor=case.get("rs_scale_factor", 5))
