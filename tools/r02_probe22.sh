#!/bin/bash
# round-2 GPU call 22: where the pipelined SimpleThreshold policy should switch from 16 to 20 persistent warps per SM, and the
# running sums with 20 warps per SM / slices for whole rounds (their dependent chain lost an instruction too)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe22.txt
S=$(date +%s)
{
pp() { echo -n "[$1] "; shift; timeout 120 env "$@" 2>&1 | tail -1; }
for l in 2500 2960 3500 4000; do
  pp "16 warps/SM" SWTPG_WARPS=2368 python tools/perf_probe.py $l 64 SimpleThreshold 60
  pp "20 warps/SM" SWTPG_WARPS=2960 python tools/perf_probe.py $l 64 SimpleThreshold 60
done
pp "AbsRS 16 warps/SM (default)" python tools/perf_probe.py 5920 64 AbsRS 60
pp "AbsRS 20 warps/SM" SWTPG_WARPS=2960 python tools/perf_probe.py 5920 64 AbsRS 60
pp "AbsRS 20 warps/SM sliced" SWTPG_WARPS=2960 SWTPG_PARTS=4 python tools/perf_probe.py 5920 64 AbsRS 60
pp "StandardRS 20 warps/SM whole links (default)" python tools/perf_probe.py 5920 64 StandardRS 60
pp "StandardRS 20 warps/SM sliced" SWTPG_PARTS=4 python tools/perf_probe.py 5920 64 StandardRS 60
echo "== elapsed $(( $(date +%s)-S )) s"
} > $OUT 2>&1
cat $OUT
