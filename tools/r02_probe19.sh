#!/bin/bash
# round-2 GPU call 19: integer frugal step with the accumulator reset as ONE IMAD (-L from the constant bank; 59 instead of 62
# instructions per quiet 4-tick group) against the previous form (variant oldstep), and the sliced hand-out's release with one
# fence instead of two (variant fence1). Variants: tools/build_variant.sh oldstep -DSWTPG_STEP_NEGL=0; fence1 -DSWTPG_SLICE_FENCES=1
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe19.txt
S=$(date +%s)
V=$PWD/build/variants
{
echo "== gpu tests, default build"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
echo "== elapsed $(( $(date +%s)-S )) s"
pp() { echo -n "[$1] "; shift; timeout 120 env "$@" 2>&1 | tail -1; }
for l in "5920 64" "6000 64" "8288 64" "3000 64" "40 2048"; do
  pp "new" python tools/perf_probe.py $l SimpleThreshold 60
  pp "old" SWTPG_LIB=$V/libswtpg_oldstep.so python tools/perf_probe.py $l SimpleThreshold 60
done
pp "new stress" python tools/perf_probe.py 5920 64 SimpleThreshold 8
pp "old stress" SWTPG_LIB=$V/libswtpg_oldstep.so python tools/perf_probe.py 5920 64 SimpleThreshold 8
for a in AbsRS StandardRS; do
  pp "new" python tools/perf_probe.py 5920 64 $a 60
  pp "old" SWTPG_LIB=$V/libswtpg_oldstep.so python tools/perf_probe.py 5920 64 $a 60
done
pp "new" python tools/perf_probe.py 1480 340 SimpleThreshold 60 wib2
pp "old" SWTPG_LIB=$V/libswtpg_oldstep.so python tools/perf_probe.py 1480 340 SimpleThreshold 60 wib2
echo "== one fence per slice (fence1) against two (new)"
for l in 5920 6000 4440; do pp "fence1" SWTPG_LIB=$V/libswtpg_fence1.so python tools/perf_probe.py $l 64 SimpleThreshold 60; done
pp "new" python tools/perf_probe.py 4440 64 SimpleThreshold 60
pp "fence1" SWTPG_LIB=$V/libswtpg_fence1.so python tools/perf_probe.py 5920 64 AbsRS 60
echo "== sliced parity under fence1"
SWTPG_LIB=$V/libswtpg_fence1.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "slices or persistent_warps or full_size" 2>&1 | tail -3
echo "== elapsed $(( $(date +%s)-S )) s"
} > $OUT 2>&1
cat $OUT
