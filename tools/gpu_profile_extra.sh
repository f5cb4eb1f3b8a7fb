#!/bin/bash
# usage (GPU box): bash tools/gpu_profile_extra.sh TAG — full ncu captures of the remaining kernels (tools/make_profiles.py summarises them)
TAG=${1:-r01}
O=gpurun_out
cap() { timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o $O/${TAG}_$1_full python tools/perf_probe.py $3 > $O/ncu_$1.log 2>&1; tail -1 $O/ncu_$1.log; }
cap wibeth_stdrs wibeth_ "5920 64 StandardRS 60"
cap wibeth_stress wibeth_ "5920 64 SimpleThreshold 8"
cap wib2_fir wib2_kernel "1480 340 FIR 5 wib2"
cap wib2_absrs wib2_kernel "1480 340 AbsRS 60 wib2"
