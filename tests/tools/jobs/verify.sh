# full GPU verification of a build: every GPU test, the default bench line, smoke()  (gpurun -- 'bash tests/tools/jobs/verify.sh')
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/verify_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/verify_pytest.log
python bench.py --steps 8 --warmup 3 > gpurun_out/verify_bench.json 2> gpurun_out/verify_bench.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/verify_bench_ref.json 2> gpurun_out/verify_bench_ref.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/verify_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/verify_smoke.log
echo done
