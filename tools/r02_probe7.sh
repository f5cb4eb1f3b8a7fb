#!/bin/bash
# round-2 GPU call 7: ALU-light FIR trackers; SimpleThreshold launch selection; bench + reference arm
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe7.txt
{
echo "== gpu tests"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "== FIR family"
for c in 3 4 5; do echo -n "[ctas/sm=$c] "; SWTPG_CTAS_PER_SM=$c python tools/perf_probe.py 5920 64 FIR 5 2>&1 | tail -1; done
for w in 1776 2368; do echo -n "[warp form warps=$w] "; SWTPG_WIBETH_KERNEL=warp SWTPG_WARPS=$w python tools/perf_probe.py 5920 64 FIR 5 2>&1 | tail -1; done
echo -n "[anytaps] "; SWTPG_TAPS=2,6,16,20,16,6,2 python tools/perf_probe.py 5920 64 FIR 5 2>&1 | tail -1
echo -n "[wib2 FIR] "; python tools/perf_probe.py 1480 340 FIR 5 wib2 2>&1 | tail -1
echo -n "[wib2 AbsRS] "; python tools/perf_probe.py 1480 340 AbsRS 60 wib2 2>&1 | tail -1
echo -n "[FIR 3000 links] "; python tools/perf_probe.py 3000 64 FIR 5 2>&1 | tail -1
echo -n "[FIR 750 links] "; python tools/perf_probe.py 750 64 FIR 5 2>&1 | tail -1
echo "== SimpleThreshold default selection"
for a in "5920 64" "6000 64" "5328 64" "4440 64" "3000 64" "750 64" "40 2048"; do echo -n "[auto] "; python tools/perf_probe.py $a SimpleThreshold 60 2>&1 | tail -1; done
echo -n "[auto stress] "; python tools/perf_probe.py 5920 64 SimpleThreshold 8 2>&1 | tail -1
} > $OUT 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; echo "bench rc=$?" >> $OUT; tail -5 gpurun_out/bench_r02b.err >> $OUT
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r02b_ref.json 2> gpurun_out/bench_r02b_ref.err; echo "ref rc=$?" >> $OUT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wibeth_ -s 3 -c 1 -f -o gpurun_out/r02e_wibeth_fir_full python tools/perf_probe.py 5920 64 FIR 5 > gpurun_out/ncu_r02e_fir.log 2>&1
tail -30 $OUT
