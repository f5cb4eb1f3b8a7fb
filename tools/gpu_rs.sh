python -m pytest tests -m gpu -q 2>&1 | tail -12
python tools/gpu_quick.py 2>&1 | grep -v "^iter" | tail -22
python tools/perf_probe.py 5920 64 AbsRS 60 wibeth
python tools/perf_probe.py 5920 64 StandardRS 60 wibeth
