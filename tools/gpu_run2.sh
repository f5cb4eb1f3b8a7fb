set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/perf_probe.py 5920 64
for lib in imad u2 u2imad; do echo LIB=$lib; SWTPG_LIB=$PWD/build/libswtpg_$lib.so python tools/perf_probe.py 5920 64; done
for c in 4 3; do echo CTAS=$c; SWTPG_CTAS_PER_SM=$c python tools/perf_probe.py 5920 64; done
for g in 1 3 14; do echo GEO=$g; SWTPG_GEO=$g python tools/perf_probe.py 5920 64; done
python tools/perf_probe.py 40 2048
python tools/perf_probe.py 11840 64
python tools/perf_probe.py 5920 256
