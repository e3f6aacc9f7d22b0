mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_caviar_gpu.py tests/test_caviar_parity_gpu.py -m gpu -x -q > gpurun_out/last_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/last_pytest.log
timeout 100 python tests/tools/dbg_time.py 1000 10000 10 296 50 > gpurun_out/last_c3_b296.txt 2>&1
echo done
