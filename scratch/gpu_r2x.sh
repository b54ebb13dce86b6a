set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 1 -o gpurun_out/r2x_nn_tc_step python bench.py --workload nn --steps 3 --warmup 3 --no-cpu-baseline --no-c5 --iters-per-launch 1 --sustained-seconds 0.1 > gpurun_out/r2x_ncu.log 2>&1
tail -3 gpurun_out/r2x_ncu.log
