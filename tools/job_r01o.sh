# r01o (1 GPU): the part of the GPU suite that the -x run of r01n did not reach
(timeout -s KILL 200 python -m pytest tests/test_slepc_gpu.py -m gpu -q 2>&1 | tail -8)
