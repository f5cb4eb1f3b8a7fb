# usage: bash tools/gpu_multi.sh N   (under gpurun --gpus N): torchrun bench at N ranks, native then reference arm
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_n${N}_ref.json 2>/dev/null
wc -l gpurun_out/bench_n$N.json gpurun_out/bench_n${N}_ref.json; tail -2 gpurun_out/bench_n$N.err; nproc; free -g | head -2
