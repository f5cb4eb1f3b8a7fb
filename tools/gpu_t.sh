python -m pytest tests -m gpu -q 2>&1 | tail -3
python tools/perf_probe.py 5920 64
python tools/perf_probe.py 6000 64
python tools/perf_probe.py 3200 64
python tools/perf_probe.py 4500 64
python tools/perf_probe.py 12000 64
