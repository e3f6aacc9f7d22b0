mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nwd_gpu.py -m gpu -x -q > gpurun_out/r4s_pytest_nwd.log 2>&1; echo "rc=$?" >> gpurun_out/r4s_pytest_nwd.log
timeout 300 python tests/tools/dbg_nwd_mt.py > gpurun_out/r4s_nwd_dbg.txt 2>&1
echo done
