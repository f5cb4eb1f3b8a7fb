python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/perf_probe.py 5920 64
for lib in cap128 cap128u2 cap128u8 cap128imad intacc; do echo LIB=$lib; SWTPG_LIB=$PWD/build/libswtpg_$lib.so python tools/perf_probe.py 5920 64; done
echo LIB=cap128 links 7104; SWTPG_LIB=$PWD/build/libswtpg_cap128.so python tools/perf_probe.py 7104 64
