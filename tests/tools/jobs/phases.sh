# CAVIaR parity tests + per-phase cycle counters of the three configurations the docs quote (single C3, 296 x C3, single C5)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_caviar_gpu.py tests/test_caviar_parity_gpu.py -m gpu -x -q > gpurun_out/phases_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/phases_pytest.log
timeout 120 python tests/tools/dbg_time.py 1000 10000 10 1 50 > gpurun_out/phases_c3.txt 2>&1
timeout 120 python tests/tools/dbg_time.py 1000 10000 10 296 50 > gpurun_out/phases_c3_b296.txt 2>&1
timeout 200 python tests/tools/dbg_time.py 5000 100000 10 1 50 > gpurun_out/phases_c5.txt 2>&1
echo done
