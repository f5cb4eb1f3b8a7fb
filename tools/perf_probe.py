"""Kernel-time probe for the fused kernels on HBM-resident frames (tuning aid).
usage: python tools/perf_probe.py [links] [units] [algorithm] [threshold] [wibeth|wib2]"""
import sys
sys.path.insert(0, '.')
import torch
import fdreadoutlibs_b200 as S
n_links = int(sys.argv[1]) if len(sys.argv) > 1 else 5920
n_units = int(sys.argv[2]) if len(sys.argv) > 2 else 64
algo = sys.argv[3] if len(sys.argv) > 3 else "SimpleThreshold"
thr = int(sys.argv[4]) if len(sys.argv) > 4 else 60
fmt = sys.argv[5] if len(sys.argv) > 5 else "wibeth"
ub = 5664 if fmt == "wib2" else 7200
spu = 256 * 12 if fmt == "wib2" else 4096
buf = torch.empty(n_links * n_units * ub, dtype=torch.uint8, device='cuda')
if fmt == "wib2":
    S.gen_wib2_device(S.gen_params(2, 0.02), buf.data_ptr(), n_links, n_units)
else:
    S.gen_wibeth_device(S.gen_params(2, 0.02), buf.data_ptr(), n_links, n_units)
torch.cuda.synchronize()
import os
taps = [int(t) for t in os.environ["SWTPG_TAPS"].split(",")] if os.environ.get("SWTPG_TAPS") else None  # FIR: other taps than firwin_int's
with S.TPGenerator(n_links, n_units, fmt=fmt, algorithm=algo, threshold=thr, tp_capacity=1 << 22, fir_taps=taps) as g:
    g.start()
    ms = []
    for i in range(8):
        g.process_device(buf.data_ptr(), n_units)
        n = g.fetch_count()
        ms.append(g.last_kernel_ms())
    best = min(ms[2:]); avg = sum(ms[2:]) / len(ms[2:])
    samples = n_links * n_units * spu
    print(f"{fmt} {algo} links={n_links} units={n_units}: best {best:.3f} ms avg {avg:.3f} ms  {samples/best/1e6:.0f} Gsamples/s  "
          f"{n_links*n_units*ub/best/1e6:.0f} GB/s ({n_links*n_units*ub/best/1e6/6452.5*100:.1f}% of 6452.5)  tps={n}", flush=True)
