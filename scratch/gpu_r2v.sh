set -x
timeout 900 python -m pytest tests/test_kernel_parity.py tests/test_fit_gpu.py -m gpu -x -q 2>&1 | tail -5
for wl in sim_art real_like pvc t1 nn; do timeout 200 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-c5 > gpurun_out/r2v_$wl.json 2> gpurun_out/r2v_$wl.err; python -c "
import json; d=json.loads(open('gpurun_out/r2v_$wl.json').read().strip().splitlines()[-1]); print('$wl', d['value'], d['ms_per_step'], d['single_launch']['value'], d['roofline']['frac'], d['final_mean_cost'])"; done
