#!/usr/bin/env python
"""bench.py — SWTPG hot-path throughput on B200 (contract: see the task brief / DESIGN.md §6).

  python bench.py [--gpus N] [--steps K] [--warmup W]            native arm (CUDA kernels through the C ABI)
  python bench.py --impl reference [--steps K] [--warmup W]      the reference's own AVX2 code on the host cores

A "step" = one superchunk batch: every link of this GPU's shard advances by `frames` WIBEth frames through the fused
unpack -> pedestal -> hit-finding kernel, state carried from the previous step. Under torchrun each rank owns one GPU
and an independent block of links (no data-path collective); the line printed by rank 0 carries the whole-job aggregate.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SAMPLES_PER_FRAME = 64 * 64
FRAME_BYTES = 7200
APA_SAMPLES_PER_S = 2560 * 62.5e6 / 32  # 2560 channels x 1.953125 MHz = 5.0e9
STATE_BYTES_PER_CHANNEL = 14           # SimpleThreshold: 7 x 16-bit carried values
TP_BYTES = 32
METRIC = "adc_samples_per_sec"
UNIT = "samples/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--links", type=int, default=5920, help="WIBEth links per GPU (5920 = 148 APAs: config[2]'s module-scale shard)")
    ap.add_argument("--frames", type=int, default=64, help="frames per link per step (superchunk length)")
    ap.add_argument("--threshold", type=int, default=60)
    ap.add_argument("--pulse-rate", type=float, default=0.02, help="pulses per channel per 64 ticks")
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the FIR / running-sum / WIB2 / stress side measurements")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons, sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_run(threshold: int, steps: int, warmup: int, pulse_rate: float):
    """The reference's own AVX2 SWTPG (oracle/_ref/libswtpg_ref.so: its headers compiled unmodified) on all host cores:
    expand_wibeth_adcs + process_window_avx2 + hit decode, one pinned worker per core, links dealt round-robin.
    Bounded sample: 8 links per core x 256 frames of the same synthetic workload. Returns (samples/s, dict)."""
    import fdreadoutlibs_b200 as S
    from oracle import binding as B

    cores = host_cores()
    kind = "reference" if B.reference_available() else "port"
    n_links, n_frames = 8 * cores, 256
    frames = S.gen_wibeth_host(S.gen_params(2, pulse_rate), n_links, n_frames, n_threads=cores)
    samples = n_links * n_frames * SAMPLES_PER_FRAME
    if kind == "reference":
        for _ in range(max(1, warmup)):
            B.ref_wibeth_bench(frames, cores, B.REF_ETH_SIMPLE_AVX2, threshold, 10, reps=1)
        secs = [B.ref_wibeth_bench(frames, cores, B.REF_ETH_SIMPLE_AVX2, threshold, 10, reps=1)[0] for _ in range(max(1, steps))]
    else:  # reference library not shipped: time the C restatement instead (single thread)
        cores = 1
        cfg = B.make_config(threshold=threshold)
        secs = []
        for _ in range(max(1, min(steps, 3))):
            t0 = time.perf_counter()
            B.oracle_process_links(cfg, frames)
            secs.append(time.perf_counter() - t0)
    total = sum(secs)
    value = samples * len(secs) / total
    extra = {}
    if kind == "reference":  # also one worker alone (per-core figure), AVX2 and the scalar "naive" processor
        one = frames[:8]
        t_avx = min(B.ref_wibeth_bench(one, 1, B.REF_ETH_SIMPLE_AVX2, threshold, 10, reps=1)[0] for _ in range(3))
        t_naive = B.ref_wibeth_bench(one[:2], 1, B.REF_ETH_SIMPLE_NAIVE, threshold, 10, reps=1)[0]
        per_core = 8 * n_frames * SAMPLES_PER_FRAME / t_avx
        extra = {"avx2_one_core": per_core, "naive_one_core": 2 * n_frames * SAMPLES_PER_FRAME / t_naive,
                 "links_real_time_per_core": per_core / 125.0e6}
    info = {"value": value, "unit": UNIT, "cores": cores, "kind": kind, **extra,
            "sample": f"{n_links} links x {n_frames} frames ({samples / 1e6:.0f} Msamples, {n_links * n_frames * FRAME_BYTES / 1e6:.0f} MB) per pass, "
                      f"{len(secs)} timed passes, threshold {threshold}, one pinned worker per core",
            "ms_per_pass": 1e3 * total / len(secs)}
    return value, info


class OneLineStdout:
    """Everything any library writes to fd 1 during the run (NCCL's version banner, the reference's setState printout, ...)
    goes to stderr; the JSON line is the only thing that reaches the real stdout."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line: str):
        sys.stdout.flush()
        os.write(self.real, (line + "\n").encode())


def main():
    out = OneLineStdout()
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    workload = (f"BASELINE config[2] per-GPU shard: {args.links} WIBEth links ({args.links / 40:.0f} APAs, {args.links * 64} channels) x "
                f"{args.frames} frames per step, synthetic noise+pulses, SimpleThreshold thr {args.threshold}")

    if args.impl == "reference":
        if rank != 0:
            return 0
        value, info = cpu_reference_run(args.threshold, args.steps, args.warmup, args.pulse_rate)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": info["ms_per_pass"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "int16", "data": "synthetic",
                "config": {"workload": workload, "note": "reference arm: bounded sample of the same workload on the host cores"},
                "cpu_baseline": {k: info[k] for k in ("value", "unit", "cores", "kind", "sample", "avx2_one_core", "naive_one_core",
                                                       "links_real_time_per_core") if k in info},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "real_time_apas": value / APA_SAMPLES_PER_S}
        out.emit(json.dumps(line))
        return 0

    import numpy as np
    import torch

    import fdreadoutlibs_b200 as S

    if not torch.cuda.is_available() or not S.device_available():
        print("bench.py: no CUDA device (this framework has no CPU fallback)", file=sys.stderr)
        return 2
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n_links, frames = args.links, args.frames
    link0 = rank * n_links  # weak scaling: every rank owns its own block of links (global link numbers differ)
    nbytes = n_links * frames * FRAME_BYTES
    samples_per_step = n_links * frames * SAMPLES_PER_FRAME

    # --- inputs resident in HBM before the timed region; 2.7 GB per step >> 126 MB L2, so no step re-reads from L2 ---
    d_frames = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    gp = S.gen_params(2, args.pulse_rate)
    S.gen_wibeth_device(gp, d_frames.data_ptr(), n_links, frames, link0=link0)
    torch.cuda.synchronize()

    gen = S.TPGenerator(n_links, frames, algorithm="SimpleThreshold", threshold=args.threshold, acc_limit=10, device=local_rank,
                        tp_capacity=1 << 22)
    gen.start()
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        gen.process_device(d_frames.data_ptr(), frames, stream=stream)
    tps_per_step = gen.fetch_count()

    # --- timed region: exactly K steps, device time by CUDA events on the launching stream ---
    kernel_ms = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with ClockSampler(local_rank) as clocks:
        ev0.record()
        for _ in range(args.steps):
            gen.process_device(d_frames.data_ptr(), frames, stream=stream)
        ev1.record()
        torch.cuda.synchronize()
        # long enough for nvidia-smi to see the load: repeat untimed launches for ~1 s, collecting per-launch times
        t_end = time.time() + 1.0
        while time.time() < t_end:
            gen.process_device(d_frames.data_ptr(), frames, stream=stream)
            kernel_ms.append(gen.last_kernel_ms())
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    local_ms_per_step = total_ms / args.steps
    tps_per_step = gen.fetch_count()
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * samples_per_step / (ms_per_step * 1e-3)

    # --- roofline of the fused kernel (rank-local): algorithmic bytes per launch / mean launch duration ---
    hbm_peak, peak_src = peaks()
    # launch duration = this rank's timed region / K (the region holds nothing but the K fused-kernel launches and their
    # 4-byte counter resets); the per-launch CUDA-event mean over the follow-on launches is kept as a cross-check.
    k_ms = local_ms_per_step
    algo_bytes = n_links * frames * FRAME_BYTES + tps_per_step * TP_BYTES + 2 * STATE_BYTES_PER_CHANNEL * n_links * 64
    achieved = algo_bytes / (k_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
                "peak_source": peak_src, "kernel": "wibeth_kernel<PackedSimpleWibEth>", "kernel_ms": k_ms,
                "kernel_ms_per_launch_events": sum(kernel_ms) / len(kernel_ms), "algorithmic_bytes_per_launch": algo_bytes}
    prof = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(prof):  # bytes per launch from the committed ncu --set full capture of this same workload
        with open(prof) as f:
            tr = json.load(f)
        if tr.get("links") == n_links and tr.get("frames") == frames:
            roofline["traffic"] = tr["dram_bytes_per_launch"]

    # --- end to end through the public API: pinned host frames in, TP list out (pinned), copies inside the timed region ---
    e2e_links = n_links
    # SWTPG_BENCH_WC=1: write-combined pinned pages (swtpg_alloc_pinned) — measured: no difference, not even with 8 GPUs ingesting at once
    wc = os.environ.get("SWTPG_BENCH_WC", "0") != "0"
    h_buf = S.PinnedBuffer(nbytes, write_combined=wc)
    h_frames = torch.from_numpy(h_buf.array)
    h_frames.copy_(d_frames)
    torch.cuda.synchronize()
    h_np = h_buf.array.reshape(e2e_links, frames, FRAME_BYTES)
    tp_cap = 1 << 22
    h_tps = torch.empty(tp_cap * TP_BYTES, dtype=torch.uint8, pin_memory=True).numpy().view(S.frames.TP_DTYPE)
    # the host link's own ceiling, measured here: the same pinned buffer copied H2D with nothing else going on
    h2d_ms = []
    for _ in range(3):
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        d_frames.copy_(h_frames, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        h2d_ms.append(c0.elapsed_time(c1))
    h2d_peak_gbs = nbytes / (min(h2d_ms) * 1e-3) / 1e9
    gen.stop()
    gen.start()
    n_tp = 0
    for _ in range(2):
        n_tp = gen.process_host(h_np, out=h_tps).size
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        n_tp = gen.process_host(h_np, out=h_tps).size
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = world * samples_per_step / e2e_s
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": n_tp * TP_BYTES + 4,
           "ms_per_step": e2e_s * 1e3, "h2d_gbs_per_gpu": nbytes / e2e_s / 1e9, "real_time_apas": e2e_value / APA_SAMPLES_PER_S,
           "ingest_roofline": {"bound": "host link (H2D)", "achieved": nbytes / e2e_s / 1e9, "peak": h2d_peak_gbs, "unit": "GB/s",
                               "frac": (nbytes / e2e_s / 1e9) / h2d_peak_gbs,
                               "peak_source": "plain pinned H2D copy of the same buffer, timed in this run (best of 3)"}}

    # --- BASELINE config[1]: ONE APA (40 links) on one GPU: latency/occupancy-limited, reported as a real-time multiple ---
    single = None
    if rank == 0:
        sl, sf = 40, 2048
        d1 = torch.empty(sl * sf * FRAME_BYTES, dtype=torch.uint8, device="cuda")
        S.gen_wibeth_device(gp, d1.data_ptr(), sl, sf)
        torch.cuda.synchronize()
        with S.TPGenerator(sl, sf, threshold=args.threshold, device=local_rank, tp_capacity=1 << 21) as g1:
            g1.start()
            ms1 = []
            for _ in range(4):
                g1.process_device(d1.data_ptr(), sf)
                g1.fetch_count()
                ms1.append(g1.last_kernel_ms())
            s1 = sl * sf * SAMPLES_PER_FRAME / (min(ms1[1:]) * 1e-3)
            single = {"workload": "BASELINE config[1]: one APA = 40 links x 2048 frames resident in HBM", "value": s1, "unit": UNIT,
                      "kernel_ms": min(ms1[1:]), "real_time_multiple": s1 / APA_SAMPLES_PER_S}
        del d1

    # --- the other kernels of the path, same box, same run (rank 0): FIR + IQR, running sums, WIB2, high-occupancy stress ---
    others = None
    if rank == 0 and not args.no_variants:
        others = {}

        def timed(fmt, algorithm, thr, d_ptr, links, units, unit_bytes, samples_per_unit, state_bytes, tp_cap=1 << 22, fir_taps=None):
            with S.TPGenerator(links, units, fmt=fmt, algorithm=algorithm, threshold=thr, device=local_rank, tp_capacity=tp_cap,
                               fir_taps=fir_taps) as g:
                g.start()
                for _ in range(3):
                    g.process_device(d_ptr, units)
                g.fetch_count()
                ms = []
                for _ in range(5):
                    g.process_device(d_ptr, units)
                    ntp = g.fetch_count()
                    ms.append(g.last_kernel_ms())
                k = sum(ms) / len(ms)
                by = links * units * unit_bytes + ntp * TP_BYTES + 2 * state_bytes * links * (256 if fmt == "wib2" else 64)
                return {"value": links * units * samples_per_unit / (k * 1e-3), "unit": UNIT, "kernel_ms": k, "tps_per_step": ntp,
                        "roofline": {"bound": "hbm", "achieved": by / (k * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                     "frac": by / (k * 1e-3) / 1e9 / hbm_peak}}

        # BASELINE config[1]/[2] data already resident (SimpleThreshold workload): FIR + IQR matched filter, AbsRS, StandardRS
        others["wibeth_fir_iqr_thr5"] = timed("wibeth", "FIR", 5, d_frames.data_ptr(), n_links, frames, FRAME_BYTES, SAMPLES_PER_FRAME, 38)
        # taps other than firwin_int(7, 0.1, 64): the packed multiply-add policy (here firwin_int's taps at multiplier 32, doubled)
        others["wibeth_fir_iqr_other_taps"] = timed("wibeth", "FIR", 5, d_frames.data_ptr(), n_links, frames, FRAME_BYTES, SAMPLES_PER_FRAME, 38,
                                                    fir_taps=[2, 6, 16, 20, 16, 6, 2])
        others["wibeth_abs_rs"] = timed("wibeth", "AbsRS", args.threshold, d_frames.data_ptr(), n_links, frames, FRAME_BYTES, SAMPLES_PER_FRAME, 22)
        others["wibeth_standard_rs"] = timed("wibeth", "StandardRS", args.threshold, d_frames.data_ptr(), n_links, frames, FRAME_BYTES,
                                             SAMPLES_PER_FRAME, 22)
        # BASELINE config[3]: high-occupancy stress — threshold 8 ADC (1.6 sigma), dense pulses
        S.gen_wibeth_device(S.gen_params(4, 0.5), d_frames.data_ptr(), n_links, frames, link0=link0)
        torch.cuda.synchronize()
        others["wibeth_simple_stress_thr8"] = timed("wibeth", "SimpleThreshold", 8, d_frames.data_ptr(), n_links, frames, FRAME_BYTES,
                                                    SAMPLES_PER_FRAME, 14, tp_cap=1 << 25)
        # BASELINE config[4]: legacy WIB2 superchunks (256 channels x 12 ticks), SimpleThreshold and FIR + IQR
        w_links, w_units = n_links // 4, 340
        d_w = torch.empty(w_links * w_units * 5664, dtype=torch.uint8, device="cuda")
        S.gen_wib2_device(gp, d_w.data_ptr(), w_links, w_units)
        torch.cuda.synchronize()
        others["wib2_simple"] = timed("wib2", "SimpleThreshold", args.threshold, d_w.data_ptr(), w_links, w_units, 5664, 256 * 12, 10)
        others["wib2_fir_iqr_thr5"] = timed("wib2", "FIR", 5, d_w.data_ptr(), w_links, w_units, 5664, 256 * 12, 38)
        others["wib2_abs_rs"] = timed("wib2", "AbsRS", args.threshold, d_w.data_ptr(), w_links, w_units, 5664, 256 * 12, 24)
        del d_w

    # --- the drop-in path as a readout application drives it: one C++ thread per link calling the frame processor's
    #     find_hits (swtpg_submit: copy into the pinned staging ring, superchunk dispatch, swtpg_poll -> TriggerPrimitives) ---
    plugin = None
    if rank == 0 and not args.no_variants:
        from fdreadoutlibs_b200 import hostshim as H

        p_links, p_units, p_sc, p_passes = 2 * host_cores(), 2048, 64, 4
        h_units = S.gen_wibeth_host(gp, p_links, p_units, n_threads=host_cores())

        def run_plugin(zero_copy):
            with H.FrameProcessors(p_links, p_sc, threshold=args.threshold, device=local_rank, emulator_mode=True, block_on_backpressure=True) as fp:
                if zero_copy:  # the payload array plays the latency buffer the constframeptrs point into
                    fp.register_buffer(h_units)
                fp.start()
                fp.push_parallel(h_units[:, :256].copy())  # warm-up: staging ring allocation, first launches
                n_tp = 0
                c0 = time.process_time()
                t0 = time.perf_counter()
                for _ in range(p_passes):  # the same buffer again and again (emulator mode keeps the timestamps running)
                    fp.push_parallel(h_units)
                    n_tp += sum(fp.take_tps(l, cap=1 << 15).size for l in range(p_links))
                fp.stop()
                dt = time.perf_counter() - t0
                cpu_s = time.process_time() - c0
                n_tp += sum(fp.take_tps(l, cap=1 << 20).size for l in range(p_links))
                if zero_copy:
                    fp.register_buffer(h_units, on=False)
            return {"value": p_passes * p_links * p_units * SAMPLES_PER_FRAME / dt, "unit": UNIT, "tps": n_tp,
                    "host_gbs": p_passes * p_links * p_units * FRAME_BYTES / dt / 1e9, "host_cpu_seconds": cpu_s, "wall_seconds": dt,
                    "real_time_apas": p_passes * p_links * p_units * SAMPLES_PER_FRAME / dt / APA_SAMPLES_PER_S}

        plugin = run_plugin(False)
        registered = run_plugin(True)
        plugin.update({"threads": p_links, "links": p_links, "frames_per_link": p_units * p_passes, "superchunk_frames": p_sc,
                       "note": "WIBEthFrameProcessor::find_hits per frame from one thread per link (C++ host shim), TriggerPrimitives out; "
                               "every frame is copied into the pinned staging ring by its link's thread (the ABI's default). With this few "
                               "links the path is bounded by the per-link serial speed of the kernel (one warp per link, ~2.1 GB/s = 9x "
                               "real time per link; batches of one handle run in order) and by 2 threads per host core, not by the copy, "
                               "the host link or the GPU (profiles/r01_plugin_probe.txt)",
                       "zero_copy_registered": dict(registered, note="same run with the payload array registered as the latency buffer "
                                                    "(swtpg_register_buffer): find_hits hands over pointers and the copy engine reads the "
                                                    "frames where they lie, one async copy per link and superchunk. Not faster on this box "
                                                    "(32 copies per batch instead of one); it takes the per-frame memcpy off the host cores")})
        del h_units

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        _, cpu = cpu_reference_run(args.threshold, steps=5, warmup=1, pulse_rate=args.pulse_rate)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "avx2_one_core", "naive_one_core", "links_real_time_per_core") if k in cpu}

    gen.close()
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16",
            "data": "synthetic",
            "config": {"workload": workload, "links_per_gpu": n_links, "frames_per_step": frames, "bytes_per_step_per_gpu": nbytes,
                       "l2_policy": "inputs larger than L2 (2.7 GB per step vs 126 MB)", "tps_per_step_per_gpu": tps_per_step,
                       "parallelism": f"links sharded over {world} GPU(s), no collective"},
            "real_time_apas": value / APA_SAMPLES_PER_S,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": e2e,
            "single_apa": single,
            "other_kernels": others,
            "plugin_streaming": plugin,
            "gpu_launches": args.steps,
            "clocks": clocks.summary(),
        }
        out.emit(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
