python tools/perf_probe.py 5920 64
python tools/perf_probe.py 6000 64
python tools/perf_probe.py 7104 64
python tools/perf_probe.py 4000 64
python tools/perf_probe.py 2960 64
python tools/perf_probe.py 1480 64
python tools/perf_probe.py 40 2048
echo CTAS4; SWTPG_CTAS_PER_SM=4 python tools/perf_probe.py 4736 64
echo CTAS3; SWTPG_CTAS_PER_SM=3 python tools/perf_probe.py 3552 64
echo CTAS2; SWTPG_CTAS_PER_SM=2 python tools/perf_probe.py 2368 64
ncu --set full --clock-control none --import-source on -k regex:wibeth_kernel -s 3 -c 1 -o gpurun_out/r01d_prof_v4 -f python tools/perf_probe.py 5920 64 > gpurun_out/ncu_e.log 2>&1
tail -2 gpurun_out/ncu_e.log
