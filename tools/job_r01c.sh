mkdir -p gpurun_out
(timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q 2>&1 | tail -8)
(timeout -s KILL 300 python tools/kbench.py 2>&1 | grep -E "multvec|gs_update|Error|error" | tail -60)
