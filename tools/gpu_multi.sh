N=${1:-4}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-variants > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_n$N.json"))
print("N", d["n_gpus"], "value", d["value"], "apas", d["real_time_apas"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"]["h2d_gbs_per_gpu"], d["clocks"])
PY
tail -3 gpurun_out/bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>/dev/null | tail -c 300
nproc
