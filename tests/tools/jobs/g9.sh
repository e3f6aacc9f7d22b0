mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r4w_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r4w_pytest.log
python bench.py --steps 8 --warmup 3 > gpurun_out/r4w_bench.json 2> gpurun_out/r4w_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4w_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r4w_smoke.log
echo done
