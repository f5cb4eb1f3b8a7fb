#!/bin/bash
# round-2 GPU call 18: sliced hand-out through the ready queue (a warp takes the slice that has been ready longest) against the
# NOTE: needs the build of tools/r02_ready_queue.diff (SWTPG_QUEUE does not exist in the committed library); result: profiles/r02_ready_queue_probe.txt
# static order (a slice waits for its predecessor): parity under forced slicing, then kernel times against slices per link.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe18.txt
S=$(date +%s)
{
echo "== gpu tests, default"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
echo "== elapsed $(( $(date +%s)-S )) s"
pp() { echo -n "[$1] "; shift; timeout 120 env "$@" 2>&1 | tail -1; }
echo "== SimpleThreshold: static order (queue=0) against the ready queue, slices per link"
for l in 5920 6000 4440; do
  pp "queue=0 parts=4" SWTPG_QUEUE=0 SWTPG_PARTS=4 python tools/perf_probe.py $l 64 SimpleThreshold 60
  for s in 4 8 16; do pp "queue=1 parts=$s" SWTPG_PARTS=$s python tools/perf_probe.py $l 64 SimpleThreshold 60; done
done
pp "default" python tools/perf_probe.py 5920 64 SimpleThreshold 60
pp "default" python tools/perf_probe.py 8288 64 SimpleThreshold 60
pp "default" python tools/perf_probe.py 3000 64 SimpleThreshold 60
pp "default" python tools/perf_probe.py 5920 256 SimpleThreshold 60
echo "== stress, running sums"
pp "queue=0 parts=4 stress" SWTPG_QUEUE=0 SWTPG_PARTS=4 python tools/perf_probe.py 5920 64 SimpleThreshold 8
pp "queue=1 parts=4 stress" SWTPG_PARTS=4 python tools/perf_probe.py 5920 64 SimpleThreshold 8
pp "queue=1 parts=8 stress" SWTPG_PARTS=8 python tools/perf_probe.py 5920 64 SimpleThreshold 8
for a in AbsRS StandardRS; do
  pp "queue=0 parts=4" SWTPG_QUEUE=0 SWTPG_PARTS=4 python tools/perf_probe.py 5920 64 $a 60
  pp "queue=1 parts=4" SWTPG_PARTS=4 python tools/perf_probe.py 5920 64 $a 60
  pp "queue=1 parts=8" SWTPG_PARTS=8 python tools/perf_probe.py 5920 64 $a 60
done
echo "== elapsed $(( $(date +%s)-S )) s"
} > $OUT 2>&1
cat $OUT
