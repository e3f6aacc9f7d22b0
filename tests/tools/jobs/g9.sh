mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r4k_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r4k_pytest.log
python bench.py --steps 8 --warmup 3 > gpurun_out/r4k_bench.json 2> gpurun_out/r4k_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4k_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r4k_smoke.log
echo done
