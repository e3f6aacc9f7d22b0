mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r4c_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r4c_pytest.log
python tests/tools/dbg_time.py 1000 10000 10 1 50 > gpurun_out/r4c_dbg_c3.txt 2>&1
python tests/tools/dbg_time.py 5000 100000 10 1 50 > gpurun_out/r4c_dbg_c5.txt 2>&1
python tests/tools/dbg_time.py 1000 10000 10 296 50 > gpurun_out/r4c_dbg_c3_b296.txt 2>&1
python bench.py --steps 4 --warmup 3 > gpurun_out/r4c_bench.json 2> gpurun_out/r4c_bench.err
echo done
