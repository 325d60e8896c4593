mkdir -p gpurun_out
(timeout -s KILL 40 python -m pytest tests/test_slepc_gpu.py -m gpu -x -q -k "test11 or bv_test2 or bv_test1" 2>&1 | tail -3) > gpurun_out/r02u_pytest.log 2>&1
tail -3 gpurun_out/r02u_pytest.log
