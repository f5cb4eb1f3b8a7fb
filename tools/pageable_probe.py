"""swtpg_process_host from pageable vs pinned host memory (tuning aid)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import fdreadoutlibs_b200 as S
n_links, n_units = 2960, 64
nbytes = n_links * n_units * 7200
pin = S.PinnedBuffer(nbytes)
S.gen_wibeth_host(S.gen_params(2, 0.02), n_links, n_units, out=pin.array.reshape(n_links, n_units, 7200)) if False else None
h = S.gen_wibeth_host(S.gen_params(2, 0.02), n_links, n_units, n_threads=16)
pin.array[:] = h.reshape(-1)
with S.TPGenerator(n_links, n_units, threshold=60, tp_capacity=1 << 22) as g:
    g.start()
    for name, buf in (("pageable", h), ("pinned", pin.array.reshape(n_links, n_units, 7200))):
        g.process_host(buf)
        t0 = time.perf_counter()
        for _ in range(3):
            n = g.process_host(buf).size
        dt = (time.perf_counter() - t0) / 3
        print(f"{name}: {dt*1e3:.1f} ms per batch of {nbytes/1e9:.2f} GB = {nbytes/dt/1e9:.1f} GB/s, {n} TPs", flush=True)
