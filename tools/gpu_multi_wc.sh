# usage: bash tools/gpu_multi_wc.sh N — e2e with regular vs write-combined pinned host buffers at N ranks (tuning aid)
N=${1:-8}
for wc in 0 1 0 1; do
  SWTPG_BENCH_WC=$wc python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$wc bench.py --gpus $N --steps 5 --warmup 3 --no-variants --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('wc=$wc', d['value'], 'e2e', d['e2e']['value'], d['e2e']['h2d_gbs_per_gpu'], d['e2e']['ingest_roofline']['peak'])"
done
