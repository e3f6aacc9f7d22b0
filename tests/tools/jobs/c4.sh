mkdir -p gpurun_out
timeout 300 python tests/tools/dbg_time.py 500 5000 10 1024 50 > gpurun_out/phases_c4.txt 2>&1
echo done
