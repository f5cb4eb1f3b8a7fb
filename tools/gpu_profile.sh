#!/bin/bash
# usage (GPU box, one GPU): bash tools/gpu_profile.sh TAG — bench line, reference arm, ncu launch list and full captures of the main kernels
# into gpurun_out/ (tools/make_profiles.py turns them into profiles/). Every step only after the plain run exited 0.
TAG=${1:-r01}
O=gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_$TAG.json 2> $O/bench_$TAG.err || { echo "bench failed"; tail -5 $O/bench_$TAG.err; exit 1; }
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_${TAG}_ref.json 2> $O/bench_${TAG}_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/ncu_list.log 2>&1
cap() { # name, kernel regex, perf_probe args
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o $O/${TAG}_$1_full python tools/perf_probe.py $3 > $O/ncu_$1.log 2>&1
  tail -1 $O/ncu_$1.log
}
cap wibeth_simple wibeth_ "5920 64"
cap wibeth_fir wibeth_ "5920 64 FIR 5"
cap wibeth_absrs wibeth_ "5920 64 AbsRS 60"
cap wib2_simple wib2_kernel "1480 340 SimpleThreshold 60 wib2"
ls -la $O/${TAG}_*_full.ncu-rep $O/${TAG}_launches.csv
