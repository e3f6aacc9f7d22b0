mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_caviar_gpu.py tests/test_caviar_parity_gpu.py -m gpu -x -q > gpurun_out/nt_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/nt_pytest.log
timeout 120 python tests/tools/dbg_time.py 1000 10000 10 296 50 > gpurun_out/nt_c3_b296.txt 2>&1
timeout 120 python tests/tools/dbg_time.py 500 5000 10 1024 50 > gpurun_out/nt_c4.txt 2>&1
timeout 120 python tests/tools/dbg_time.py 1000 10000 10 1 50 > gpurun_out/nt_c3.txt 2>&1
echo done
