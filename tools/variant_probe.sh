#!/bin/bash
# usage (on the GPU box): tools/variant_probe.sh "<perf_probe args>" NAME...   — runs tools/perf_probe.py against build/variants/libswtpg_NAME.so ("base" = the in-tree library)
ARGS=$1; shift
for v in "$@"; do
  if [ "$v" = base ]; then unset SWTPG_LIB; else export SWTPG_LIB=$PWD/build/variants/libswtpg_$v.so; fi
  echo -n "[$v] "; timeout 60 python tools/perf_probe.py $ARGS 2>&1 | tail -1; echo
done
