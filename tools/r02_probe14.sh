#!/bin/bash
# round-2 GPU call 14: full test suite + bench line with the current build; streaming path against slots / gather CTAs
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02b_gputests.log 2>&1; tail -3 gpurun_out/r02b_gputests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err || { echo "bench failed"; tail -20 gpurun_out/bench_r02b.err; }
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02b.json'))
print('value',d['value'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],d['e2e']['h2d_gbs_per_gpu'],'cpu',d['cpu_baseline']['value'])
print('module_split',d['module_split']['roofline_frac_per_gpu'],'single_apa',d['single_apa']['real_time_multiple'])
print({k:round(v['roofline']['frac'],3) for k,v in d['other_kernels'].items()})
print({k:round(v['roofline_frac'],3) for k,v in d['link_count_sweep'].items()})
PY
OUT=gpurun_out/r02_probe14.txt
{
echo "== streaming path: slots / gather CTAs / superchunk"
P="timeout 300 python tools/plugin_probe.py"
$P 240 64 1 4 16 0 512 16
SWTPG_PROBE_SLOTS=2 $P 240 64 1 4 16 0 512 16
SWTPG_PROBE_SLOTS=4 $P 240 64 1 4 16 0 512 16
SWTPG_PROBE_SLOTS=6 $P 240 64 1 4 16 0 512 16
SWTPG_GATHER_CTAS=16 $P 240 64 1 4 16 0 512 16
SWTPG_GATHER_CTAS=32 $P 240 64 1 4 16 0 512 16
SWTPG_GATHER_CTAS=148 $P 240 64 1 4 16 0 512 16
SWTPG_GATHER_MODE=1 $P 240 64 1 4 16 0 512 16
$P 240 128 1 4 16 0 512 16
$P 240 32 1 4 16 0 512 16
} > $OUT 2>&1
cat $OUT
