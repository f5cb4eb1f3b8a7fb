#!/bin/bash
# round-2 GPU call 11: how much of the full-GPU time is the tail of the last round of links (work unit = one link x the whole batch)?
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe11.txt
{
echo "== default launch selection, base = integer forms + FMA-pipe adds, nofma = integer forms with plain adds, old = fp16 forms"
for v in base nofma old; do
  if [ "$v" = base ]; then unset SWTPG_LIB; else export SWTPG_LIB=$PWD/build/variants/libswtpg_$v.so; fi
  for l in 5920 4144 8288 6000; do echo -n "[$v default] "; python tools/perf_probe.py $l 64 SimpleThreshold 60 2>&1 | tail -1; done
  for l in 5920 2368 4736; do echo -n "[$v default] "; python tools/perf_probe.py $l 64 FIR 5 2>&1 | tail -1; done
  echo -n "[$v default] "; python tools/perf_probe.py 1480 340 SimpleThreshold 60 wib2 2>&1 | tail -1
  echo -n "[$v default] "; python tools/perf_probe.py 740 340 SimpleThreshold 60 wib2 2>&1 | tail -1
done
unset SWTPG_LIB
echo "== exactly one / two rounds per persistent warp (pipe=0, 2x32 ring, 20 warps per SM = 2960)"
for l in 2960 5920 4440; do echo -n "[base pipe=0] "; SWTPG_SIMPLE_PIPE=0 python tools/perf_probe.py $l 64 SimpleThreshold 60 2>&1 | tail -1; done
echo "== same links, longer batches (tail unchanged, per-link prologue amortised)"
for u in 32 128 256; do echo -n "[base default] "; python tools/perf_probe.py 5920 $u SimpleThreshold 60 2>&1 | tail -1; done
} > $OUT 2>&1
cat $OUT
