mkdir -p gpurun_out
for h in 15 31 63 127; do echo "== C3 helpers $h"; CM_CAVIAR_HELPERS=$h timeout 120 python tests/tools/dbg_time.py 1000 10000 10 1 50 2>&1 | grep -E "iters=50|^  (a|h|i|t)" | tail -19; done > gpurun_out/r4h_helpers_c3.txt 2>&1
echo done
