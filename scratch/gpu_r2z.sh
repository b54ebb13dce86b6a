set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; tail -c 600 gpurun_out/r2z_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_ref.json 2> gpurun_out/r2z_ref.err; tail -c 400 gpurun_out/r2z_ref.json
for wl in real_like spatial disp; do timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-c5 > gpurun_out/r2z_$wl.json 2> gpurun_out/r2z_$wl.err; python -c "
import json; d=json.loads(open('gpurun_out/r2z_$wl.json').read().strip().splitlines()[-1]); print('$wl', d['value'], d['ms_per_step'], d['roofline']['frac'], d['final_mean_cost'])"; done
