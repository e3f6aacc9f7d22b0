mkdir -p gpurun_out
python bench.py --steps 8 --warmup 3 > gpurun_out/r4z_bench.json 2> gpurun_out/r4z_bench.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r4z_bench_ref.json 2> gpurun_out/r4z_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r4z_launches.csv python bench.py --steps 2 --warmup 1 --no-traffic --no-c5 --no-c4 --no-nwd --no-e2e --no-single --no-cpu-baseline > gpurun_out/r4z_ncu.log 2>&1
echo done
