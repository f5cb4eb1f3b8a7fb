#!/bin/bash
# round-2 GPU call 6: where does the FIR kernel's time go now; more SimpleThreshold geometry; fixed tests
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe6.txt
{
echo "== gpu tests"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "== FIR: forms and rings"
for v in base q3x32 q4x16; do
  if [ "$v" = base ]; then unset SWTPG_LIB; else export SWTPG_LIB=$PWD/build/variants/libswtpg_$v.so; fi
  for c in 3 4 5; do echo -n "[$v ctas/sm=$c] "; SWTPG_CTAS_PER_SM=$c python tools/perf_probe.py 5920 64 FIR 5 2>&1 | tail -1; done
done
unset SWTPG_LIB
for w in 1776 2368 2960; do echo -n "[warp form warps=$w] "; SWTPG_WIBETH_KERNEL=warp SWTPG_WARPS=$w python tools/perf_probe.py 5920 64 FIR 5 2>&1 | tail -1; done
echo "== SimpleThreshold straight-line: small rings, many warps"
for v in g2x16 g2x16m g2x8; do export SWTPG_LIB=$PWD/build/variants/libswtpg_$v.so
  for w in 3552 4144 4736; do echo -n "[$v pipe=0 warps=$w] "; SWTPG_SIMPLE_PIPE=0 SWTPG_WARPS=$w timeout 60 python tools/perf_probe.py 5920 64 SimpleThreshold 60 2>&1 | tail -1; done
  echo -n "[$v pipe=0 warps=4144 stress] "; SWTPG_SIMPLE_PIPE=0 SWTPG_WARPS=4144 timeout 60 python tools/perf_probe.py 5920 64 SimpleThreshold 8 2>&1 | tail -1
  echo -n "[$v pipe=0 warps=4144 4440 links] "; SWTPG_SIMPLE_PIPE=0 SWTPG_WARPS=4144 timeout 60 python tools/perf_probe.py 4440 64 SimpleThreshold 60 2>&1 | tail -1
  echo -n "[$v pipe=0 warps=4144 6000 links] "; SWTPG_SIMPLE_PIPE=0 SWTPG_WARPS=4144 timeout 60 python tools/perf_probe.py 6000 64 SimpleThreshold 60 2>&1 | tail -1
done
unset SWTPG_LIB
} > $OUT 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wibeth_ -s 3 -c 1 -f -o gpurun_out/r02d_wibeth_fir_full python tools/perf_probe.py 5920 64 FIR 5 > gpurun_out/ncu_r02d_fir.log 2>&1
tail -2 gpurun_out/ncu_r02d_fir.log >> $OUT
tail -40 $OUT
