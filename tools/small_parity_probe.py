"""Small run of every kernel family, checked against the oracle (quick sanity probe; also usable under compute-sanitizer where that is allowed).
usage: python tools/small_parity_probe.py"""
import sys
sys.path.insert(0, '.')
import numpy as np
import fdreadoutlibs_b200 as S
from fdreadoutlibs_b200 import frames as F
from oracle import binding as B

ok = True
for fmt, algo, aid, thr in [("wibeth", "SimpleThreshold", 0, 20), ("wibeth", "AbsRS", 1, 30), ("wibeth", "StandardRS", 2, 30), ("wibeth", "FIR", 3, 5),
                            ("wibeth", "SimpleThreshold", 0, 40000), ("wib2", "SimpleThreshold", 0, 30), ("wib2", "FIR", 3, 5),
                            ("wib2", "AbsRS", 1, 30)]:
    n_links, n_units = (5, 6) if fmt == "wibeth" else (3, 10)
    gen = S.gen_wibeth_host if fmt == "wibeth" else S.gen_wib2_host
    units = gen(S.gen_params(81, 0.6), n_links, n_units)
    want, _ = B.oracle_process_links(B.make_config(fmt=fmt, algorithm=aid, threshold=thr), units)
    with S.TPGenerator(n_links, 4, fmt=fmt, algorithm=algo, threshold=thr, tp_capacity=1 << 18) as g:
        g.start()
        got = [g.process_host(np.ascontiguousarray(units[:, u:u + 4])) for u in range(0, n_units, 4)]
    a, b = F.sort_tps(np.concatenate(got)), F.sort_tps(want)
    good = a.size == b.size and (a == b).all()
    ok &= good
    print(fmt, algo, thr, a.size, "OK" if good else "MISMATCH", flush=True)
print("ALL OK" if ok else "FAILURES")
sys.exit(0 if ok else 1)
