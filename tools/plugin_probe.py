"""Plug-in streaming path probe (tuning aid / profiles/r02_plugin_probe.txt):
  python tools/plugin_probe.py LINKS SUPERCHUNK ZERO_COPY THREADS [BURST] [PACE] [UNITS] [PASSES]
THREADS feeder threads serve LINKS links round-robin through WIBEthFrameProcessor (sequence_check, timestamp_check, find_hits);
ZERO_COPY=1 registers the payload array as the latency buffer; PACE > 0 runs against the clock at that multiple of real time."""
import resource
import sys
import time

sys.path.insert(0, '.')
import numpy as np

import fdreadoutlibs_b200 as S
from fdreadoutlibs_b200 import hostshim as H

links, sc, zc, threads = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
burst = int(sys.argv[5]) if len(sys.argv) > 5 else 16
pace = float(sys.argv[6]) if len(sys.argv) > 6 else 0.0
units = int(sys.argv[7]) if len(sys.argv) > 7 else 1024
passes = int(sys.argv[8]) if len(sys.argv) > 8 else 4
import os
EMU = os.environ.get('SWTPG_PROBE_EMULATOR', '0') != '0'  # emulator mode rewrites every header (and, as in the reference, flags a sequence error per frame)
buf = S.PinnedBuffer(links * units * 7200) if zc == 2 else None  # 2: cudaHostAlloc'ed latency buffer instead of cudaHostRegister
h = S.gen_wibeth_host(S.gen_params(2, 0.02), links, units, n_threads=8, out=None if buf is None else buf.array)
SLOTS = int(os.environ.get('SWTPG_PROBE_SLOTS', '3'))
with H.FrameProcessors(links, sc, threshold=60, n_slots=SLOTS, emulator_mode=EMU, block_on_backpressure=pace == 0, count_only_sink=True) as fp:
    if zc:
        fp.register_buffer(h)
    fp.start()
    warm = h[:, :128].copy()
    fp.push_feeders(warm, n_threads=threads, burst=burst)  # warm-up: engine creation, first launches (from unregistered memory)
    time.sleep(0.05)
    r0, t0 = resource.getrusage(resource.RUSAGE_SELF), time.perf_counter()
    per_pass = []
    if pace > 0:  # one call per pass, so that drops can be attributed (the clock restarts with every call)
        st = {"feeder_cpu_s": 0.0, "late_bursts": 0}
        for _ in range(passes):
            s1 = fp.push_feeders(h, n_threads=threads, burst=burst, pace=pace, passes=1)
            st["feeder_cpu_s"] = s1["feeder_cpu_s"]  # thread CPU clocks are per call: the last call's threads only
            st["late_bursts"] += s1["late_bursts"]
            per_pass.append(sum(fp.get_info(l)["num_frames_dropped_busy"] for l in range(links)))
        st["feeder_cpu_s"] *= passes
    else:
        st = fp.push_feeders(h, n_threads=threads, burst=burst, pace=pace, passes=passes)
    t_feed = time.perf_counter() - t0
    fp.stop()
    dt = time.perf_counter() - t0
    r1 = resource.getrusage(resource.RUSAGE_SELF)
    cpu = (r1.ru_utime - r0.ru_utime) + (r1.ru_stime - r0.ru_stime)
    dropped = sum(per_pass) + sum(fp.get_info(l)["num_frames_dropped_busy"] for l in range(links))
    tps = fp.tp_count()
    cnt = fp.counters()
    tim = fp.stream_timing()
    if zc:
        fp.register_buffer(h, on=False)
n = passes * links * units
apas = links / 40
print(f"links={links} sc={sc} zero_copy={zc} threads={threads} burst={burst} pace={pace}: {n*7200/dt/1e9:.1f} GB/s ({n*4096/dt/1e9:.1f} Gsamples/s = "
      f"{n*4096/dt/5e9:.2f} real-time APAs)  wall {dt*1e3:.0f} ms (feed {t_feed*1e3:.0f})  process cpu {cpu*1e3:.0f} ms = {cpu/dt:.2f} cores busy, "
      f"feeders {st['feeder_cpu_s']/dt:.2f} cores = {st['feeder_cpu_s']/n*1e6:.3f} us/frame  host core-s per APA-s {cpu/dt/ (n*4096/dt/5e9):.3f}  "
      f"tps={tps} dropped={dropped} late_bursts={st['late_bursts']} by_address={cnt['units_zero_copy']} by_copy={cnt['units_staged']} "
      f"batches={cnt['batches']} gather {tim['gather_ms']:.1f} ms = {cnt['h2d_bytes']/max(tim['gather_ms'],1e-9)/1e6:.1f} GB/s while active, "
      f"kernel {tim['kernel_ms']:.1f} ms" + (f" dropped_per_pass={per_pass}" if per_pass else ""), flush=True)
