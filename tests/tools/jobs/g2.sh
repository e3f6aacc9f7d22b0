mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r4b_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r4b_pytest.log
python tests/tools/dbg_time.py 1000 10000 10 1 50 > gpurun_out/r4b_dbg_c3.txt 2>&1
python tests/tools/dbg_time.py 5000 100000 10 1 50 > gpurun_out/r4b_dbg_c5.txt 2>&1
echo done
