mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:caviar_fit_kernel -s 0 -c 1 -f -o gpurun_out/r4t_batch python tests/tools/prof_cmd.py 296 > gpurun_out/r4t_ncu296.log 2>&1
echo done
