mkdir -p gpurun_out
python tests/tools/prof_single.py > gpurun_out/r4j_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:caviar_fit_kernel -s 1 -c 1 -f -o gpurun_out/r4j_single python tests/tools/prof_single.py > gpurun_out/r4j_ncu.log 2>&1
python tests/tools/prof_cmd.py 296 > gpurun_out/r4j_plain296.log 2>&1 || exit 1
ncu --set full --clock-control none -k regex:caviar_fit_kernel -s 1 -c 1 -f -o gpurun_out/r4j_batch python tests/tools/prof_cmd.py 296 > gpurun_out/r4j_ncu296.log 2>&1
echo done
