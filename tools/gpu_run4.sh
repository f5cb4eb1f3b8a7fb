ncu --set full --clock-control none --import-source on -k regex:wibeth_kernel -s 3 -c 1 -o gpurun_out/r01c_prof_v3 -f python tools/perf_probe.py 5920 64 > gpurun_out/ncu_d.log 2>&1
tail -3 gpurun_out/ncu_d.log
