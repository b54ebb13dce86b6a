python -m pytest tests -m gpu -q -k "disp or nn or device_resident" 2>&1 | tail -2
for wv in 8 16 4; do SVBASL_DW_WARPS=$wv python bench.py --workload disp --steps 8 --no-cpu-baseline > gpurun_out/r2k_disp_$wv.json 2>gpurun_out/r2k_disp_$wv.err; python -c "
import json; d=json.load(open('gpurun_out/r2k_disp_$wv.json')); print('disp warps $wv', d['value'], d['ms_per_step'])"; done
python bench.py --workload nn --steps 10 --no-cpu-baseline > gpurun_out/r2k_nn.json 2> gpurun_out/r2k_nn.err; python -c "
import json; d=json.load(open('gpurun_out/r2k_nn.json')); print('nn', d['value'], d['ms_per_step'], d['single_launch']['value'], d['roofline']['xu']['frac'], d['roofline']['fp32']['frac'])"
ncu --set full --clock-control none --import-source on -k regex:disp_warp -s 4 -c 1 -o gpurun_out/r2k_disp_warp python bench.py --workload disp --steps 5 --no-cpu-baseline > gpurun_out/r2k_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 1 -o gpurun_out/r2k_nn_step python bench.py --workload nn --steps 3 --warmup 3 --no-cpu-baseline --iters-per-launch 1 --sustained-seconds 0.1 > gpurun_out/r2k_ncu2.log 2>&1
