"""Quick GPU parity probe (development aid; the real checks live in tests/)."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import fdreadoutlibs_b200 as S
from fdreadoutlibs_b200 import frames as F
from oracle import binding as B

def same(a, b, what):
    a, b = F.sort_tps(a), F.sort_tps(b)
    ok = a.size == b.size and (a == b).all()
    print(what, a.size, b.size, "OK" if ok else "MISMATCH", flush=True)
    if not ok:
        n = min(a.size, b.size)
        bad = np.nonzero(a[:n] != b[:n])[0]
        if bad.size:
            print(" first diff:", a[bad[0]], b[bad[0]])
    return ok

p = S.gen_params(1, 0.05)
n_links, n_units = 6, 40
fr = S.gen_wibeth_host(p, n_links, n_units)
allok = True
for algo_name, algo in [("SimpleThreshold", 0), ("AbsRS", 1), ("StandardRS", 2), ("FIR", 3)]:
    for thr, L in [(20, 10), (60, 10), (5 if algo == 3 else 40000, 10), (30, 0), (30, -3)]:
        if algo == 3 and L != 10:
            continue
        cfg = B.make_config(algorithm=algo, threshold=thr, acc_limit=L)
        ref, oracles = B.oracle_process_links(cfg, fr)
        with S.TPGenerator(n_links, 16, algorithm=algo_name, threshold=thr, acc_limit=L, tp_capacity=1 << 20) as g:
            g.start()
            parts = []
            for u0 in range(0, n_units, 16):   # 16,16,8: state carried across batches, last one short
                chunk = np.ascontiguousarray(fr[:, u0:u0 + 16])
                parts.append(g.process_host(chunk, units_stride=chunk.shape[1]))
            got = np.concatenate(parts)
            ok = same(got, ref, f"{algo_name} thr={thr} L={L}")
            st = g.dump_state(3); so = oracles[3].state()
            for f_ in ("pedestal", "accum", "prev_was_over", "hit_charge", "hit_tover"):
                if not (st[f_] == so[f_]).all():
                    print("  state mismatch", f_); ok = False
            allok &= ok
# debug dump parity (packed path)
cfg = B.make_config(algorithm=0, threshold=60)
o = B.Oracle(cfg)
to, ped_o, wav_o = o.process(fr[0], dump=True)
with S.TPGenerator(1, n_units, threshold=60) as g:
    g.start()
    tg, ped_g, wav_g = g.process_host(fr[:1], debug=True)
    print("dump ped", (ped_g[0] == ped_o).all(), "wav", (wav_g[0] == wav_o).all())
    allok &= bool((ped_g[0] == ped_o).all() and (wav_g[0] == wav_o).all())
# throughput probe
import torch
n_links, n_units = 5920, 64
buf = torch.empty(n_links * n_units * 7200, dtype=torch.uint8, device='cuda')
S.gen_wibeth_device(S.gen_params(2, 0.02), buf.data_ptr(), n_links, n_units)
torch.cuda.synchronize()
with S.TPGenerator(n_links, n_units, threshold=60) as g:
    g.start()
    for i in range(5):
        g.process_device(buf.data_ptr(), n_units)
        n = g.fetch_count()
        ms = g.last_kernel_ms()
        samples = n_links * n_units * 4096
        print(f"iter {i}: {ms:.3f} ms, {samples/ms/1e6:.1f} Gsamples/s, {n_links*n_units*7200/ms/1e6:.1f} GB/s, tps={n}", flush=True)
print("ALL OK" if allok else "FAILURES")
