mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r4y_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r4y_pytest.log
python bench.py --steps 4 --warmup 3 --no-traffic --no-c5 --no-nwd --no-single --no-cpu-baseline > gpurun_out/r4y_bench.json 2> gpurun_out/r4y_bench.err
echo done
