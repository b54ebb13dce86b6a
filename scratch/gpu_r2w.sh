set -x
timeout 900 python -m pytest tests/test_nn_tensorcore_gpu.py tests/test_fit_gpu.py tests/test_dropin_scripts.py -m gpu -x -q 2>&1 | tail -5
for wl in nn nn_fp32; do timeout 200 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-c5 > gpurun_out/r2w_$wl.json 2> gpurun_out/r2w_$wl.err; python -c "
import json; d=json.loads(open('gpurun_out/r2w_$wl.json').read().strip().splitlines()[-1]); print('$wl', d['value'], d['ms_per_step'], d['single_launch']['value'], d['roofline']['frac'], d['roofline']['xu']['frac'], d['final_mean_cost'])"; done
timeout 200 python bench.py --workload nn --voxels 4000000 --steps 10 --warmup 3 --no-cpu-baseline --no-c5 > gpurun_out/r2w_nn4m.json 2> gpurun_out/r2w_nn4m.err; python -c "
import json; d=json.loads(open('gpurun_out/r2w_nn4m.json').read().strip().splitlines()[-1]); print('nn4m', d['value'], d['ms_per_step'])"
