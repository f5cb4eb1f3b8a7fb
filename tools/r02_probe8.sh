#!/bin/bash
# round-2 GPU call 8: FIR tracker forms x producer waits; where the streaming path's time goes
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe8.txt
{
echo "== FIR: merged (base) vs separate (fsep) quartile steps, producer wait variants, 3-stage ring"
for v in base fsep fmpw1 fmpw2 fmpw2s1000 fmq3; do
  if [ "$v" = base ]; then unset SWTPG_LIB; else export SWTPG_LIB=$PWD/build/variants/libswtpg_$v.so; fi
  for c in 4 5; do echo -n "[$v ctas/sm=$c] "; SWTPG_CTAS_PER_SM=$c python tools/perf_probe.py 5920 64 FIR 5 2>&1 | tail -1; done
done
for v in base fmpw2 fmpw2s1000; do
  if [ "$v" = base ]; then unset SWTPG_LIB; else export SWTPG_LIB=$PWD/build/variants/libswtpg_$v.so; fi
  echo -n "[$v] "; python tools/perf_probe.py 1480 340 FIR 5 wib2 2>&1 | tail -1
  echo -n "[$v] "; python tools/perf_probe.py 1480 340 SimpleThreshold 60 wib2 2>&1 | tail -1
done
unset SWTPG_LIB
echo "== streaming path: where the time goes"
P="timeout 300 python tools/plugin_probe.py"
$P 240 64 1 4 16 0 512 16
SWTPG_PROBE_SLOTS=4 $P 240 64 1 4 16 0 512 16
SWTPG_PROBE_SLOTS=6 $P 240 64 1 4 16 0 512 16
SWTPG_GATHER_CTAS=16 $P 240 64 1 4 16 0 512 16
SWTPG_GATHER_CTAS=32 $P 240 64 1 4 16 0 512 16
SWTPG_GATHER_CTAS=148 $P 240 64 1 4 16 0 512 16
SWTPG_GATHER_MODE=1 $P 240 64 1 4 16 0 512 16
$P 240 128 1 4 16 0 512 16
$P 240 32 1 4 16 0 512 16
$P 480 64 1 4 16 0 256 16
$P 120 64 1 4 16 0 1024 16
} > $OUT 2>&1
tail -40 $OUT
