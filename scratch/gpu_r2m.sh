for t in 15 0 8 9 12 1; do SVBASL_DW_TUNE=$t python bench.py --workload disp --steps 8 --no-cpu-baseline > gpurun_out/r2m_disp_$t.json 2>gpurun_out/r2m_disp.err; python -c "
import json; d=json.load(open('gpurun_out/r2m_disp_$t.json')); print('disp tune $t', d['value'], d['ms_per_step'])"; done
