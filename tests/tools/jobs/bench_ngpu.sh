# bench.py under torchrun on N GPUs of one box (gpurun --gpus N -- 'bash tests/tools/jobs/bench_ngpu.sh N')
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 4 --warmup 3 --no-traffic > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo done
