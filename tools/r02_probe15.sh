#!/bin/bash
# round-2 GPU call: SimpleThreshold (integer accumulators) in the CTA form (4 links per CTA in lock-step + producer warp, the WIB2
# kernel's structure, which reaches 84 % issue-slot use there) against the one-warp-per-CTA form; ring of 2 x 32 and 2 x 16 ticks;
# 4-7 CTAs per SM. FIR + IQR on the 2 x 16 ring with 4-6 CTAs per SM (56 registers instead of 64).
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe15.txt
{
echo -n "[base simple] "; python tools/perf_probe.py 5920 64 SimpleThreshold 60 2>&1 | tail -1
for v in q7 q7c16; do
  export SWTPG_LIB=$PWD/build/variants/libswtpg_$v.so
  for c in 4 5 6 7; do echo -n "[$v ctas=$c simple] "; SWTPG_CTAS_PER_SM=$c timeout 60 python tools/perf_probe.py 5920 64 SimpleThreshold 60 2>&1 | tail -1; done
done
unset SWTPG_LIB
echo -n "[base fir] "; python tools/perf_probe.py 5920 64 FIR 5 2>&1 | tail -1
export SWTPG_LIB=$PWD/build/variants/libswtpg_q7c16.so
for c in 4 5 6; do echo -n "[q7c16 ctas=$c fir] "; SWTPG_CTAS_PER_SM=$c timeout 60 python tools/perf_probe.py 5920 64 FIR 5 2>&1 | tail -1; done
for c in 4 5 6; do echo -n "[q7c16 ctas=$c absrs] "; SWTPG_CTAS_PER_SM=$c timeout 60 python tools/perf_probe.py 5920 64 AbsRS 60 2>&1 | tail -1; done
unset SWTPG_LIB
echo -n "[base absrs] "; python tools/perf_probe.py 5920 64 AbsRS 60 2>&1 | tail -1
} > $OUT 2>&1
cat $OUT
