for lib in pushcall u2 u1; do echo LIB=$lib; SWTPG_LIB=$PWD/build/libswtpg_$lib.so python tools/perf_probe.py 5920 64; done
for g in 1 4 14 13; do echo GEO=$g; SWTPG_GEO=$g python tools/perf_probe.py 5920 64; done
