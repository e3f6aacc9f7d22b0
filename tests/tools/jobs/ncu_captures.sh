# the ncu captures of profiles/: single C3 fit and the 296-fit launch with source correlation, the launch list of a short bench
mkdir -p gpurun_out
python tests/tools/prof_single.py > gpurun_out/ncu_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:caviar_fit_kernel -s 1 -c 1 -f -o gpurun_out/ncu_single python tests/tools/prof_single.py > gpurun_out/ncu_single.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:caviar_fit_kernel -s 0 -c 1 -f -o gpurun_out/ncu_batch python tests/tools/prof_cmd.py 296 > gpurun_out/ncu_batch.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launches.csv python bench.py --steps 2 --warmup 1 --no-traffic --no-c5 --no-c4 --no-nwd --no-e2e --no-single --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo done
