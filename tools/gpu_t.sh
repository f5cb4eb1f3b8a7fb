python -m pytest tests/test_gpu_parity.py -m gpu -q -k "config1 or reference_itself" 2>&1 | tail -12
