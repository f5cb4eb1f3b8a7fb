#!/bin/bash
# round-2 GPU call 2: tests on the new streaming engine, instruction latencies, plug-in path matrix, FIR kernel capture
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe2.txt
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader; nproc
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== latency"; timeout 120 build/bin/latency
echo "== plug-in path"
P="timeout 300 python tools/plugin_probe.py"
$P 240 64 1 8 16
$P 240 64 0 8 16
$P 240 64 2 8 16
$P 240 64 1 4 16
$P 240 64 1 2 16
$P 240 64 1 8 64
$P 240 64 1 8 4
$P 240 16 1 8 16
$P 240 256 1 8 16
SWTPG_GATHER_MODE=1 $P 240 64 1 8 16
SWTPG_GATHER_CTAS=16 $P 240 64 1 8 16
SWTPG_GATHER_CTAS=148 $P 240 64 1 8 16
$P 40 64 1 4 16
$P 1000 64 1 8 16 0 256
echo "== paced"
$P 200 64 1 8 16 1.0 1024 6
$P 240 64 1 8 16 1.0 1024 6
$P 200 16 1 8 8 1.0 1024 6
$P 40 64 1 2 16 1.0 2048 6
} > $OUT 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wibeth_ -s 3 -c 1 -f -o gpurun_out/r02a_wibeth_fir_full python tools/perf_probe.py 5920 64 FIR 5 > gpurun_out/ncu_r02a_fir.log 2>&1
tail -2 gpurun_out/ncu_r02a_fir.log >> $OUT
tail -30 $OUT
