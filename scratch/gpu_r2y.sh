set -x
timeout 600 python -m pytest tests/test_nn_tensorcore_gpu.py -m gpu -x -q 2>&1 | tail -4
for wl in nn; do timeout 200 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-c5 > gpurun_out/r2y_$wl.json 2> gpurun_out/r2y_$wl.err; python -c "
import json; d=json.loads(open('gpurun_out/r2y_$wl.json').read().strip().splitlines()[-1]); print('$wl', d['value'], d['ms_per_step'], d['single_launch']['value'], d['roofline']['frac'], d['roofline']['xu']['frac'], d['final_mean_cost'])"; done
