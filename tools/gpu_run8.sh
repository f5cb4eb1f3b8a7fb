python -m pytest tests -m gpu -q 2>&1 | tail -15
python tools/perf_probe.py 5920 64
python tools/perf_probe.py 1480 340 SimpleThreshold 60 wib2
python tools/perf_probe.py 888 340 SimpleThreshold 60 wib2
python tools/perf_probe.py 1480 64 FIR 5 wib2
python tools/perf_probe.py 2960 16 FIR 5 wibeth
python tools/perf_probe.py 2960 16 AbsRS 60 wibeth
