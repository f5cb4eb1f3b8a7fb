#!/bin/bash
# round-2 GPU call: do the gathers of consecutive batches share the host link better one after the other (one gather stream) than side by side?
cd "${GRAFT_REPO_ROOT:-.}"
OUT=gpurun_out/r02_probe17.txt
{
for rep in 1 2; do
for ser in 0 1; do for slots in 3 4 6; do
  echo -n "[serial=$ser slots=$slots] "; SWTPG_GATHER_SERIAL=$ser SWTPG_PROBE_SLOTS=$slots timeout 100 python tools/plugin_probe.py 240 64 1 4 2>&1 | tail -1 | cut -c1-330
done; done
done
for ser in 0 1; do echo -n "[serial=$ser slots=4 ctas=32] "; SWTPG_GATHER_CTAS=32 SWTPG_GATHER_SERIAL=$ser SWTPG_PROBE_SLOTS=4 timeout 100 python tools/plugin_probe.py 240 64 1 4 2>&1 | tail -1 | cut -c1-330; done
echo "== streaming tests with the serial gather stream"; SWTPG_GATHER_SERIAL=1 timeout 300 python -m pytest tests -m gpu -q -x -k "stream or plugin or shim or zero_copy or thread" 2>&1 | tail -2
} > $OUT 2>&1
cat $OUT
