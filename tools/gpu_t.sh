python -m pytest tests -m gpu -q 2>&1 | tail -3
SWTPG_FIR_FORCE_EXACT=1 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "FIR or fir" 2>&1 | tail -2
python tools/perf_probe.py 5920 64 FIR 5 wibeth
python tools/perf_probe.py 1480 340 FIR 5 wib2
