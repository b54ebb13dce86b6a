set -x
timeout 900 python -m pytest tests/test_kernel_parity.py -m gpu -x -q -k "disp" 2>&1 | tail -5
for wl in disp; do timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-c5 > gpurun_out/r3a_$wl.json 2> gpurun_out/r3a_$wl.err; python -c "
import json; d=json.loads(open('gpurun_out/r3a_$wl.json').read().strip().splitlines()[-1]); print('$wl', d['value'], d['ms_per_step'], d['single_launch'], d['roofline']['frac'], d['final_mean_cost'])"; done
