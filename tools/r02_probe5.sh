#!/bin/bash
# round-2 GPU call 5: FIR family with the merged quartile step; SimpleThreshold ring geometry with more warps; emulator app
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe5.txt
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== gpu tests"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15
echo "== FIR family"
for a in "5920 64 FIR 5" "1480 340 FIR 5 wib2" "1480 340 AbsRS 60 wib2"; do tools/variant_probe.sh "$a" base firu1 firu4; done
echo -n "[anytaps] "; SWTPG_TAPS=2,6,16,20,16,6,2 python tools/perf_probe.py 5920 64 FIR 5 2>&1 | tail -1
echo -n "[forced exact] "; SWTPG_FIR_FORCE_EXACT=1 python tools/perf_probe.py 5920 64 FIR 5 2>&1 | tail -1
echo -n "[thr 7: exact tier by configuration] "; python tools/perf_probe.py 5920 64 FIR 7 2>&1 | tail -1
for c in 4 5 6; do echo -n "[ctas/sm=$c] "; SWTPG_CTAS_PER_SM=$c python tools/perf_probe.py 5920 64 FIR 5 2>&1 | tail -1; done
echo "== SimpleThreshold: ring geometry x warps (straight-line form = pipe 0, pipelined = pipe 1)"
for v in base g2x16 g3x16; do for pipe in 0 1; do for w in 2368 2960 3552 4144; do
  if [ "$v" = base ]; then unset SWTPG_LIB; else export SWTPG_LIB=$PWD/build/variants/libswtpg_$v.so; fi
  echo -n "[$v pipe=$pipe warps=$w] "; SWTPG_SIMPLE_PIPE=$pipe SWTPG_WARPS=$w timeout 60 python tools/perf_probe.py 5920 64 SimpleThreshold 60 2>&1 | tail -1
done; done; done
unset SWTPG_LIB
echo "== default"
for a in "5920 64 SimpleThreshold 60" "6000 64 SimpleThreshold 60" "40 2048 SimpleThreshold 60" "5920 64 AbsRS 60" "5920 64 StandardRS 60" "5920 64 SimpleThreshold 8" "1480 340 SimpleThreshold 60 wib2"; do
  echo -n "[auto] "; python tools/perf_probe.py $a 2>&1 | tail -1; done
} > $OUT 2>&1
tail -40 $OUT
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err; echo "bench rc=$?" >> $OUT; tail -5 gpurun_out/bench_r02a.err >> $OUT
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02a_ref.json 2> gpurun_out/bench_r02a_ref.err; echo "ref rc=$?" >> $OUT
tail -12 $OUT
