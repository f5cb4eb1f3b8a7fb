python -m pytest tests -m gpu -q 2>&1 | tail -40
python tools/gpu_quick.py 2>&1 | head -60
