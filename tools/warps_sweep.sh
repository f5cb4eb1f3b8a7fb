#!/bin/bash
# usage (GPU box): tools/warps_sweep.sh "<perf_probe args>" W...   — kernel time against the number of persistent warps (SWTPG_WARPS)
ARGS=$1; shift
for w in "$@"; do
  echo -n "[warps=$w] "; SWTPG_WARPS=$w python tools/perf_probe.py $ARGS 2>&1 | tail -1
done
