mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_caviar_gpu.py tests/test_caviar_parity_gpu.py -m gpu -x -q > gpurun_out/r4i_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r4i_pytest.log
for h in 15 31 63; do echo "== C3 helpers $h"; CM_CAVIAR_HELPERS=$h timeout 120 python tests/tools/dbg_time.py 1000 10000 10 1 50 2>&1 | grep -E "iters=50|^  (a|h|i|t)" | tail -19; done > gpurun_out/r4i_helpers_c3.txt 2>&1
timeout 200 python tests/tools/dbg_time.py 5000 100000 10 1 50 > gpurun_out/r4i_dbg_c5.txt 2>&1
echo done
