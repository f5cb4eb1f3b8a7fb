#!/bin/bash
# round-2 GPU call 13: sliced hand-out, second implementation (out-of-line wait: the tick loop keeps its 72 registers)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe13.txt
{
echo "== gpu tests, default"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for s in 2 8; do echo "== gpu tests, every wibeth_kernel launch sliced in $s"; SWTPG_PARTS=$s timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3; done
echo "== SimpleThreshold, default launch selection, slices per link"
for l in 5920 6000 8288 4440; do
  for s in 1 2 4 8; do echo -n "[parts=$s] "; SWTPG_PARTS=$s python tools/perf_probe.py $l 64 SimpleThreshold 60 2>&1 | tail -1; done
done
for l in 5920 3000 750; do echo -n "[default] "; python tools/perf_probe.py $l 64 SimpleThreshold 60 2>&1 | tail -1; done
echo -n "[default] "; python tools/perf_probe.py 40 2048 SimpleThreshold 60 2>&1 | tail -1
echo "== stress, running sums"
for s in 1 4; do echo -n "[parts=$s stress] "; SWTPG_PARTS=$s python tools/perf_probe.py 5920 64 SimpleThreshold 8 2>&1 | tail -1; done
for s in 1 2 4; do echo -n "[parts=$s] "; SWTPG_PARTS=$s python tools/perf_probe.py 5920 64 AbsRS 60 2>&1 | tail -1; done
for s in 1 2 4; do echo -n "[parts=$s] "; SWTPG_PARTS=$s python tools/perf_probe.py 5920 64 StandardRS 60 2>&1 | tail -1; done
} > $OUT 2>&1
cat $OUT
