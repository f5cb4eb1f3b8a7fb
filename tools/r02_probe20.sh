#!/bin/bash
# round-2 GPU call 20: sliced hand-out with HALVING slices (n/2, n/4, ..., rest: SWTPG_SLICE_GEOM=1) against equal ones — what the
# last round leaves idle is at most one LAST slice, so small last slices shorten the tail without more state round trips per link —
# and the software-pipelined SimpleThreshold policy at full load now that its dependent chain is one instruction shorter.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe20.txt
S=$(date +%s)
{
echo "== sliced parity with halving slices (forced 2 and 8 slices in the sub-processes, ragged and whole batches)"
SWTPG_SLICE_GEOM=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "slices or persistent_warps or full_size" 2>&1 | tail -12
echo "== elapsed $(( $(date +%s)-S )) s"
pp() { echo -n "[$1] "; shift; timeout 120 env "$@" 2>&1 | tail -1; }
for l in 5920 6000 4440; do
  pp "equal x4" SWTPG_SLICE_GEOM=0 python tools/perf_probe.py $l 64 SimpleThreshold 60
  pp "halving x4" SWTPG_SLICE_GEOM=1 python tools/perf_probe.py $l 64 SimpleThreshold 60
  pp "halving x8" SWTPG_SLICE_GEOM=1 SWTPG_PARTS=8 python tools/perf_probe.py $l 64 SimpleThreshold 60
done
echo "== pipelined policy (16 warps per SM) at full load"
for l in 5920 6000; do
  pp "pipe equal x4" SWTPG_SIMPLE_PIPE=1 python tools/perf_probe.py $l 64 SimpleThreshold 60
  pp "pipe halving x4" SWTPG_SIMPLE_PIPE=1 SWTPG_SLICE_GEOM=1 python tools/perf_probe.py $l 64 SimpleThreshold 60
done
pp "pipe 20 warps/SM halving x4" SWTPG_SIMPLE_PIPE=1 SWTPG_WARPS=2960 SWTPG_SLICE_GEOM=1 SWTPG_PARTS=4 python tools/perf_probe.py 5920 64 SimpleThreshold 60
echo "== stress, running sums"
pp "equal x4 stress" python tools/perf_probe.py 5920 64 SimpleThreshold 8
pp "halving x4 stress" SWTPG_SLICE_GEOM=1 python tools/perf_probe.py 5920 64 SimpleThreshold 8
for a in AbsRS StandardRS; do
  pp "equal x4" python tools/perf_probe.py 5920 64 $a 60
  pp "halving x4" SWTPG_SLICE_GEOM=1 python tools/perf_probe.py 5920 64 $a 60
done
echo "== elapsed $(( $(date +%s)-S )) s"
} > $OUT 2>&1
cat $OUT
