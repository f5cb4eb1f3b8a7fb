"""Host-side TP ordering speed (tuning aid): python tools/sort_probe.py [n]"""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import fdreadoutlibs_b200 as S
from fdreadoutlibs_b200 import frames as F
n = int(sys.argv[1]) if len(sys.argv) > 1 else 458000
rng = np.random.default_rng(1)
tps = np.zeros(n, dtype=F.TP_DTYPE)
tps["time_start"] = 10**15 + 32 * rng.integers(0, 64 * 64 + 200, n)
tps["link"] = rng.integers(0, 5920, n)
tps["channel"] = rng.integers(0, 64, n)
a = tps.copy()
best = 1e9
for _ in range(6):
    a[:] = tps
    t0 = time.perf_counter(); S.sort_tps(a); best = min(best, time.perf_counter() - t0)
parts = [S.sort_tps(tps[i::8].copy()) for i in range(8)]
bm = 1e9
for _ in range(4):
    t0 = time.perf_counter(); S.merge_sorted(parts); bm = min(bm, time.perf_counter() - t0)
print(f"swtpg_sort_tps {n} records: {best*1e3:.1f} ms ({n/best/1e6:.1f} M TPs/s); swtpg_merge_sorted 8 lists: {bm*1e3:.1f} ms", flush=True)
