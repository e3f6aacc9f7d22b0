mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_caviar_gpu.py -m gpu -x -q -k "chain_teams or helper_ctas" > gpurun_out/r4p_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r4p_pytest.log
echo done
