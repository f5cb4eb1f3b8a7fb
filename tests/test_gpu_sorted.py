"""GPU (B200): SWTPG_FLAG_SORTED_TPS — the TP list of a batch ordered on the device (csrc/swtpg_sort.cu) must be, record for
record and in order, what swtpg_sort_tps makes of the unordered list on the host: (time_start, link, channel), the order
TriggerPrimitiveTypeAdapter::operator< imposes downstream (include/fdreadoutlibs/TriggerPrimitiveTypeAdapter.hpp:26-29)."""
import numpy as np
import pytest

import fdreadoutlibs_b200 as S
from fdreadoutlibs_b200 import frames as F
from oracle import binding as B

pytestmark = pytest.mark.gpu


def both(units, n_links, max_units, cap=1 << 20, **kw):
    """The same batches through an ordering handle and a plain one: (device-ordered lists, unordered lists, sort stats)."""
    out = []
    for flag in (True, False):
        with S.TPGenerator(n_links, max_units, sorted_tps=flag, **kw) as g:
            g.start()
            parts = [g.process_host(np.ascontiguousarray(units[:, u:u + max_units]), cap=cap) for u in range(0, units.shape[1], max_units)]
            out.append((parts, g.sort_stats()))
    return out[0][0], out[1][0], out[0][1]


@pytest.mark.parametrize("fmt,algorithm,thr", [("wibeth", "SimpleThreshold", 25), ("wibeth", "FIR", 4), ("wibeth", "AbsRS", 40),
                                               ("wib2", "SimpleThreshold", 25), ("wib2", "FIR", 4)])
def test_device_order_equals_host_order(fmt, algorithm, thr):
    n_links, n_units, step = 37, 10, 4  # three batches, the last one short
    gen = S.gen_wib2_host if fmt == "wib2" else S.gen_wibeth_host
    units = gen(S.gen_params(77, 0.4), n_links, n_units)
    dev, plain, st = both(units, n_links, step, fmt=fmt, algorithm=algorithm, threshold=thr)
    assert sum(p.size for p in dev) > 2000
    for d, p in zip(dev, plain):
        want = S.sort_tps(p)
        assert d.size == want.size and (d == want).all()
    assert st["lists"] == 3 and st["finished_on_host"] == 0 and st["last_ms"] > 0
    # and the multiset is the oracle's
    cfg = B.make_config(fmt=fmt, algorithm=S.ALGORITHMS[algorithm], threshold=thr)
    want, _ = B.oracle_process_links(cfg, units)
    a, b = F.sort_tps(np.concatenate(dev)), F.sort_tps(want)
    assert a.size == b.size and (a == b).all()


def test_many_tiles_and_wide_tiles():
    """Dense hits: 2.6 M records per batch = tiles wider than the 512-record minimum, every digit pass with thousands of tiles."""
    n_links, n_units = 600, 24
    units = S.gen_wibeth_host(S.gen_params(5, 0.5), n_links, n_units)
    dev, plain, st = both(units, n_links, n_units, cap=1 << 23, threshold=8, tp_capacity=1 << 23)
    assert dev[0].size > (1 << 21)
    want = S.sort_tps(plain[0])
    assert dev[0].size == want.size and (dev[0] == want).all() and st["finished_on_host"] == 0


def test_small_lists():
    """0, 1, 2 and a handful of records (below one warp round)."""
    n_links, n_units = 2, 1
    for rate, thr in ((0.0, 4000), (0.05, 200), (0.3, 60)):
        units = S.gen_wibeth_host(S.gen_params(3, rate), n_links, n_units)
        dev, plain, _ = both(units, n_links, n_units, threshold=thr)
        want = S.sort_tps(plain[0])
        assert dev[0].size == want.size and (dev[0] == want).all()


def test_unrelated_timestamps_are_finished_on_the_host():
    """Links whose clocks are 2^62 ticks apart: the packed key would need more than 64 bits, the host orders the copy."""
    n_links, n_units = 6, 4
    units = S.gen_wibeth_host(S.gen_params(9, 0.4), n_links, n_units).copy()
    ts = units[0].reshape(n_units, 7200)[:, 8:16].copy().view(np.uint64)
    units[0].reshape(n_units, 7200)[:, 8:16] = (ts + np.uint64(1 << 62)).view(np.uint8)
    dev, plain, st = both(units, n_links, n_units, threshold=25)
    want = S.sort_tps(plain[0])
    assert dev[0].size == want.size > 100 and (dev[0] == want).all() and st["finished_on_host"] == 1


def test_equal_keys_get_the_hosts_tie_break():
    """A link that delivers the same frames (same timestamps) twice produces records with equal (time_start, link, channel);
    the device order is then completed by the host's tie-break, so both orderings stay identical."""
    n_links, n_units = 3, 3
    one = S.gen_wibeth_host(S.gen_params(11, 0.5), n_links, n_units)
    units = np.concatenate([one, one], axis=1)
    dev, plain, st = both(units, n_links, 2 * n_units, threshold=25)
    want = S.sort_tps(plain[0])
    k = np.stack([want["time_start"], want["link"].astype(np.uint64), want["channel"].astype(np.uint64)], axis=1)
    assert (k[1:] == k[:-1]).all(axis=1).any(), "the construction should produce equal keys"
    assert st["finished_on_host"] == 1
    assert dev[0].size == want.size
    assert (dev[0] == want).all()


def test_streaming_batches_come_back_ordered():
    n_links, n_units = 12, 16
    units = S.gen_wibeth_host(S.gen_params(21, 0.4), n_links, n_units)
    want, _ = B.oracle_process_links(B.make_config(threshold=30), units)
    with S.TPGenerator(n_links, 4, threshold=30, sorted_tps=True) as g:
        g.start()
        for u in range(n_units):
            for l in range(n_links):
                assert g.submit(l, units[l, u], wait_us=1_000_000)
        got = g.drain()
        batches = g.counters()["batches"]
        st = g.sort_stats()
    a, b = F.sort_tps(got), F.sort_tps(want)
    assert a.size == b.size and (a == b).all()
    # the drained stream is a concatenation of per-batch ordered lists: at most one descent per batch boundary
    descents = int((got["time_start"][1:] < got["time_start"][:-1]).sum())
    assert descents <= batches - 1 and st["lists"] >= 1
