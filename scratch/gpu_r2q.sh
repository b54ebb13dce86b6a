timeout 300 python -m pytest tests/test_nn_tensorcore_gpu.py -m gpu -q -x 2>&1 | tail -4
for wl in nn_tc; do python bench.py --workload $wl --steps 10 --no-cpu-baseline > gpurun_out/r2q_$wl.json 2> gpurun_out/r2q_$wl.err; python -c "
import json; d=json.load(open('gpurun_out/r2q_$wl.json')); print('$wl', d['value'], d['ms_per_step'], d['single_launch']['value'], 'xu', d['roofline']['xu']['frac'], 'fp32', d['roofline']['fp32']['frac'], d['final_mean_cost'])"; done
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 1 -o gpurun_out/r2q_nn_tc_step python bench.py --workload nn_tc --steps 3 --warmup 3 --no-cpu-baseline --iters-per-launch 1 --sustained-seconds 0.1 > gpurun_out/r2q_ncu.log 2>&1
