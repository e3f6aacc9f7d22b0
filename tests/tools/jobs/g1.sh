mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-c4 --no-nwd > gpurun_out/r4a_bench.json 2> gpurun_out/r4a_bench.err
for h in 15 31 63; do echo "== C3 helpers $h"; CM_CAVIAR_HELPERS=$h python tests/tools/dbg_time.py 1000 10000 10 1 50 2>&1 | grep -v "^  \(a\|h\|i\|s\|n\|c\|t\)" | tail -2; done > gpurun_out/r4a_helpers_c3.txt 2>&1
for h in 47 95 127; do echo "== C5 helpers $h"; CM_CAVIAR_HELPERS=$h python tests/tools/dbg_time.py 5000 100000 10 1 50 2>&1 | tail -28; done > gpurun_out/r4a_helpers_c5.txt 2>&1
echo done
