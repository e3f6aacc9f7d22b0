mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-traffic --no-c5 --no-c4 --no-nwd --no-e2e --no-single --no-cpu-baseline > gpurun_out/r4x_plain.json 2> gpurun_out/r4x_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r4x_launches.csv python bench.py --steps 2 --warmup 1 --no-traffic --no-c5 --no-c4 --no-nwd --no-e2e --no-single --no-cpu-baseline > gpurun_out/r4x_ncu.log 2>&1
echo done
