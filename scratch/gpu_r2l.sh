python -m pytest tests -m gpu -q -k "disp or nn" 2>&1 | tail -2
python scratch/dbg_disp2.py 2>&1 | tail -3
python bench.py --workload disp --steps 8 --no-cpu-baseline > gpurun_out/r2l_disp.json 2>gpurun_out/r2l_disp.err; python -c "
import json; d=json.load(open('gpurun_out/r2l_disp.json')); print('disp', d['value'], d['ms_per_step'])"
python bench.py --workload nn --steps 10 --no-cpu-baseline > gpurun_out/r2l_nn.json 2> gpurun_out/r2l_nn.err; python -c "
import json; d=json.load(open('gpurun_out/r2l_nn.json')); print('nn', d['value'], d['ms_per_step'], d['single_launch']['value'], d['roofline']['xu']['frac'], d['roofline']['fp32']['frac'])"
ncu --set full --clock-control none --import-source on -k regex:disp_warp -s 4 -c 1 -o gpurun_out/r2l_disp_warp python bench.py --workload disp --steps 5 --no-cpu-baseline > gpurun_out/r2l_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 1 -o gpurun_out/r2l_nn_step python bench.py --workload nn --steps 3 --warmup 3 --no-cpu-baseline --iters-per-launch 1 --sustained-seconds 0.1 > gpurun_out/r2l_ncu2.log 2>&1
