# demixer GPU tests + per-stage cycle counters of the multi-trace tcgen05 kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nwd_gpu.py -m gpu -x -q > gpurun_out/nwd_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/nwd_pytest.log
timeout 300 python tests/tools/dbg_nwd_mt.py > gpurun_out/nwd_dbg.txt 2>&1
echo done
