#!/bin/bash
# round-2 GPU call 21: the software-pipelined SimpleThreshold policy at full load (its chain is one instruction shorter since call 19):
# persistent warps per SM x slices (equal / halving) x link count
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe21.txt
S=$(date +%s)
{
pp() { echo -n "[$1] "; shift; timeout 120 env SWTPG_SIMPLE_PIPE=1 "$@" 2>&1 | tail -1; }
echo "== pipelined policy, 20 warps per SM (2960)"
pp "whole links" SWTPG_WARPS=2960 SWTPG_PARTS=1 python tools/perf_probe.py 5920 64 SimpleThreshold 60
pp "equal x4" SWTPG_WARPS=2960 SWTPG_PARTS=4 python tools/perf_probe.py 5920 64 SimpleThreshold 60
pp "halving x4" SWTPG_WARPS=2960 SWTPG_PARTS=4 SWTPG_SLICE_GEOM=1 python tools/perf_probe.py 5920 64 SimpleThreshold 60
pp "halving x8" SWTPG_WARPS=2960 SWTPG_PARTS=8 SWTPG_SLICE_GEOM=1 python tools/perf_probe.py 5920 64 SimpleThreshold 60
for l in 6000 8288 4440 3000; do
  pp "equal x4" SWTPG_WARPS=2960 SWTPG_PARTS=4 python tools/perf_probe.py $l 64 SimpleThreshold 60
  pp "halving x4" SWTPG_WARPS=2960 SWTPG_PARTS=4 SWTPG_SLICE_GEOM=1 python tools/perf_probe.py $l 64 SimpleThreshold 60
done
pp "halving x4 stress" SWTPG_WARPS=2960 SWTPG_PARTS=4 SWTPG_SLICE_GEOM=1 python tools/perf_probe.py 5920 64 SimpleThreshold 8
pp "halving x4, 256 units" SWTPG_WARPS=2960 SWTPG_PARTS=4 SWTPG_SLICE_GEOM=1 python tools/perf_probe.py 5920 256 SimpleThreshold 60
echo "== 16 warps per SM (2368)"
pp "halving x4" SWTPG_WARPS=2368 SWTPG_PARTS=4 SWTPG_SLICE_GEOM=1 python tools/perf_probe.py 8288 64 SimpleThreshold 60
pp "halving x4" SWTPG_WARPS=2368 SWTPG_PARTS=4 SWTPG_SLICE_GEOM=1 python tools/perf_probe.py 3000 64 SimpleThreshold 60
echo "== elapsed $(( $(date +%s)-S )) s"
} > $OUT 2>&1
cat $OUT
