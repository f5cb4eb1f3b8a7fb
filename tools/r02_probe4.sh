#!/bin/bash
# round-2 GPU call 4: policy form x ring depth against link count; paced plug-in path with per-pass drop counts
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe4.txt
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15
echo "== SimpleThreshold: pipe x stages"
for a in "40 2048" "240 64" "240 256" "750 64" "1500 64" "2220 64" "3000 64" "4440 64" "5920 64" "6000 64"; do
  for pipe in 0 1; do for st in 2 4; do
    echo -n "[pipe=$pipe stages=$st] "; SWTPG_SIMPLE_PIPE=$pipe SWTPG_SIMPLE_STAGES=$st timeout 60 python tools/perf_probe.py $a SimpleThreshold 60 2>&1 | tail -1
  done; done
done
echo "== warps per SM at full load"
for pipe in 0 1; do for w in 2368 2960; do echo -n "[pipe=$pipe warps=$w] "; SWTPG_SIMPLE_PIPE=$pipe SWTPG_SIMPLE_STAGES=2 SWTPG_WARPS=$w python tools/perf_probe.py 5920 64 SimpleThreshold 60 2>&1 | tail -1; done; done
echo "== default selection"
for a in "40 2048" "240 64" "750 64" "1500 64" "3000 64" "5920 64"; do echo -n "[auto] "; python tools/perf_probe.py $a SimpleThreshold 60 2>&1 | tail -1; done
echo "== paced plug-in path"
P="timeout 300 python tools/plugin_probe.py"
$P 200 64 1 4 16 1.0 1024 6
SWTPG_PROBE_SLOTS=8 $P 200 64 1 4 16 1.0 1024 6
$P 200 64 1 4 16 1.0 1024 20
$P 160 64 1 4 16 1.0 1024 6
SWTPG_PROBE_SLOTS=8 $P 240 64 1 4 16 1.0 1024 6
$P 240 64 1 4 16 0.9 1024 6
$P 40 64 1 1 16 1.0 2048 6
echo "== unpaced"
$P 240 64 1 4 16
$P 240 64 1 2 16
$P 240 64 0 8 16
$P 240 128 1 4 16
} > $OUT 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wibeth_ -s 3 -c 1 -f -o gpurun_out/r02c_wibeth_simple_full python tools/perf_probe.py 5920 64 > gpurun_out/ncu_r02c_simple.log 2>&1
SWTPG_SIMPLE_PIPE=1 SWTPG_SIMPLE_STAGES=4 timeout 300 ncu --set full --clock-control none --import-source on -k regex:wibeth_ -s 3 -c 1 -f -o gpurun_out/r02c_wibeth_simple_40links_full python tools/perf_probe.py 40 2048 > gpurun_out/ncu_r02c_simple40.log 2>&1
tail -2 gpurun_out/ncu_r02c_simple.log >> $OUT
tail -30 $OUT
