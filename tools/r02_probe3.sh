#!/bin/bash
# round-2 GPU call 3: pipelined SimpleThreshold kernel (IADD3 median step, prefetch, deferred quiet test) + zero-copy fix
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe3.txt
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== SimpleThreshold"
for a in "5920 64 SimpleThreshold 60" "6000 64 SimpleThreshold 60" "3000 64 SimpleThreshold 60" "1500 64 SimpleThreshold 60" "750 64 SimpleThreshold 60" "240 64 SimpleThreshold 60" "40 2048 SimpleThreshold 60"; do
  tools/variant_probe.sh "$a" base pipe0 gu2 gu8
done
echo "== warps per SM"
for w in 1776 2368 2960 3256; do echo -n "[warps=$w] "; SWTPG_WARPS=$w python tools/perf_probe.py 5920 64 SimpleThreshold 60 2>&1 | tail -1; done
echo "== others"
for a in "1480 340 SimpleThreshold 60 wib2" "5920 64 AbsRS 60" "5920 64 StandardRS 60" "5920 64 SimpleThreshold 8"; do
  tools/variant_probe.sh "$a" base pipe0
done
echo "== plug-in path"
P="timeout 300 python tools/plugin_probe.py"
$P 240 64 1 8 16
$P 240 64 0 8 16
$P 240 64 2 8 16
$P 240 64 1 4 16
$P 240 64 1 2 16
$P 240 64 0 4 16
$P 240 16 1 4 16
$P 240 256 1 4 16
SWTPG_GATHER_MODE=1 $P 240 64 1 4 16
SWTPG_PROBE_EMULATOR=1 $P 240 64 1 4 16
$P 1000 64 1 4 16 0 256
echo "== paced"
$P 200 64 1 4 16 1.0 1024 6
$P 240 64 1 4 16 1.0 1024 6
$P 240 64 1 8 16 1.0 1024 6
$P 240 16 1 4 8 1.0 1024 6
$P 40 64 1 1 16 1.0 2048 6
} > $OUT 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wibeth_ -s 3 -c 1 -f -o gpurun_out/r02b_wibeth_simple_full python tools/perf_probe.py 5920 64 > gpurun_out/ncu_r02b_simple.log 2>&1
tail -2 gpurun_out/ncu_r02b_simple.log >> $OUT
tail -30 $OUT
