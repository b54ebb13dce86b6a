set -x
timeout 300 python -m pytest tests/test_nn_tensorcore_gpu.py -m gpu -x -q 2>&1 | tail -15
timeout 200 python bench.py --workload nn_tc --steps 10 --warmup 3 --no-cpu-baseline --no-c5 > gpurun_out/r2t_nn_tc.json 2> gpurun_out/r2t_nn_tc.err
tail -c 1500 gpurun_out/r2t_nn_tc.json
