"""Plug-in streaming path probe (tuning aid): python tools/plugin_probe.py LINKS SUPERCHUNK ZERO_COPY [SLOTS]"""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import fdreadoutlibs_b200 as S
from fdreadoutlibs_b200 import hostshim as H
links, sc, zc = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
units, passes = 2048, 4
h = S.gen_wibeth_host(S.gen_params(2, 0.02), links, units, n_threads=8)
with H.FrameProcessors(links, sc, threshold=60, emulator_mode=True, block_on_backpressure=True) as fp:
    if zc:
        fp.register_buffer(h)
    fp.start()
    fp.push_parallel(h[:, :256].copy())
    c0, t0 = time.process_time(), time.perf_counter()
    n = 0
    for _ in range(passes):
        fp.push_parallel(h)
        n += sum(fp.take_tps(l, cap=1 << 15).size for l in range(links))
    fp.stop()
    dt, cpu = time.perf_counter() - t0, time.process_time() - c0
    if zc:
        fp.register_buffer(h, on=False)
print(f"links={links} sc={sc} zero_copy={zc}: {passes*links*units*7200/dt/1e9:.1f} GB/s  wall {dt*1e3:.0f} ms  cpu {cpu*1e3:.0f} ms  ({cpu/dt:.1f} cores busy)  "
      f"{dt/(passes*units*links)*1e6*min(links,16):.2f} us/frame/core", flush=True)
