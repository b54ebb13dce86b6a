timeout 300 python -m pytest tests -m gpu -q -k "nn" 2>&1 | tail -2
python bench.py --workload nn_fp32 --steps 10 --no-cpu-baseline > gpurun_out/r2r_nn.json 2> gpurun_out/r2r_nn.err; python -c "
import json; d=json.load(open('gpurun_out/r2r_nn.json')); print('nn_fp32 2-row', d['value'], d['ms_per_step'], d['single_launch']['value'], 'xu', d['roofline']['xu']['frac'], 'fp32', d['roofline']['fp32']['frac'])"
