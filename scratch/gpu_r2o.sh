timeout 300 python -m pytest tests/test_nn_tensorcore_gpu.py -m gpu -q -x 2>&1 | tail -15
