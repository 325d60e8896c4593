mkdir -p gpurun_out
(timeout -s KILL 110 python -m pytest tests/test_z_examples.py tests/test_slepc_gpu.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r02t_pytest.log 2>&1
tail -5 gpurun_out/r02t_pytest.log
