mkdir -p gpurun_out
python -m pytest tests/test_caviar_gpu.py tests/test_caviar_parity_gpu.py -m gpu -x -q > gpurun_out/r4v_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r4v_pytest.log
python tests/tools/dbg_time.py 1000 10000 10 1 50 > gpurun_out/r4v_dbg_c3.txt 2>&1
python tests/tools/dbg_time.py 1000 10000 10 296 50 > gpurun_out/r4v_dbg_c3_b296.txt 2>&1
echo done
