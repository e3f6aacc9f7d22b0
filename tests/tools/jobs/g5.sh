mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_caviar_gpu.py tests/test_caviar_parity_gpu.py -m gpu -x -q > gpurun_out/r4g_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r4g_pytest.log
timeout 120 python tests/tools/dbg_time.py 1000 10000 10 1 50 > gpurun_out/r4g_dbg_c3.txt 2>&1
CM_DBG=4097 timeout 120 python tests/tools/dbg_time.py 1000 10000 10 1 50 > gpurun_out/r4g_dbg_c3_oldinv.txt 2>&1
timeout 120 python tests/tools/dbg_time.py 1000 10000 10 296 50 > gpurun_out/r4g_dbg_c3_b296.txt 2>&1
timeout 200 python tests/tools/dbg_time.py 5000 100000 10 1 50 > gpurun_out/r4g_dbg_c5.txt 2>&1
echo done
