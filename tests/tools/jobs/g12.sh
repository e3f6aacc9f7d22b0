mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 4 --warmup 3 --no-traffic > gpurun_out/r4q_bench_8gpu.json 2> gpurun_out/r4q_bench_8gpu.err
echo done
