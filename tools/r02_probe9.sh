#!/bin/bash
# round-2 GPU call 9: integer frugal accumulators (SWTPG_FIR_INT / SWTPG_SIMPLE_INT) against the fp16-subnormal forms
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe9.txt
{
echo "== gpu tests (integer forms)"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
echo "== kernels: base = integer forms (7-instruction SimpleThreshold step), s2 = short-chain step, old = fp16 forms (merged FIR), firsep = fp16 forms (separate FIR)"
for v in base s2 old firsep; do
  if [ "$v" = base ]; then unset SWTPG_LIB; else export SWTPG_LIB=$PWD/build/variants/libswtpg_$v.so; fi
  for pipe in 0 1; do
    for l in "5920 64" "3000 64" "750 64" "40 2048"; do echo -n "[$v pipe=$pipe] "; SWTPG_SIMPLE_PIPE=$pipe python tools/perf_probe.py $l SimpleThreshold 60 2>&1 | tail -1; done
  done
  echo -n "[$v pipe=0 warps=2368] "; SWTPG_SIMPLE_PIPE=0 SWTPG_WARPS=2368 python tools/perf_probe.py 5920 64 SimpleThreshold 60 2>&1 | tail -1
  echo -n "[$v pipe=1 warps=2960] "; SWTPG_SIMPLE_PIPE=1 SWTPG_WARPS=2960 python tools/perf_probe.py 5920 64 SimpleThreshold 60 2>&1 | tail -1
  echo -n "[$v stress] "; python tools/perf_probe.py 5920 64 SimpleThreshold 8 2>&1 | tail -1
  for c in 4 5 6; do echo -n "[$v ctas/sm=$c] "; SWTPG_CTAS_PER_SM=$c python tools/perf_probe.py 5920 64 FIR 5 2>&1 | tail -1; done
  echo -n "[$v other taps] "; SWTPG_TAPS=2,6,16,20,16,6,2 python tools/perf_probe.py 5920 64 FIR 5 2>&1 | tail -1
  echo -n "[$v] "; python tools/perf_probe.py 5920 64 AbsRS 60 2>&1 | tail -1
  echo -n "[$v] "; python tools/perf_probe.py 5920 64 StandardRS 60 2>&1 | tail -1
  echo -n "[$v] "; python tools/perf_probe.py 1480 340 SimpleThreshold 60 wib2 2>&1 | tail -1
  echo -n "[$v] "; python tools/perf_probe.py 1480 340 FIR 5 wib2 2>&1 | tail -1
  echo -n "[$v] "; python tools/perf_probe.py 1480 340 AbsRS 60 wib2 2>&1 | tail -1
done
unset SWTPG_LIB
} > $OUT 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wibeth_ -s 3 -c 1 -f -o gpurun_out/r02e_wibeth_fir_full python tools/perf_probe.py 5920 64 FIR 5 > gpurun_out/ncu_r02e_fir.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wibeth_ -s 3 -c 1 -f -o gpurun_out/r02e_wibeth_simple_full python tools/perf_probe.py 5920 64 > gpurun_out/ncu_r02e_simple.log 2>&1
cat $OUT
