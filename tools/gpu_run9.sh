python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r01c.json 2> gpurun_out/bench_r01c.err; tail -c 300 gpurun_out/bench_r01c.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r01c_ref.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wibeth_kernel -s 3 -c 1 -o gpurun_out/r01_wibeth_simple_full -f python tools/perf_probe.py 5920 64 > gpurun_out/ncu_g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wib2_kernel -s 3 -c 1 -o gpurun_out/r01_wib2_simple_full -f python tools/perf_probe.py 1480 340 SimpleThreshold 60 wib2 > gpurun_out/ncu_h.log 2>&1
tail -2 gpurun_out/ncu_g.log gpurun_out/ncu_h.log
