mkdir -p gpurun_out
(timeout -s KILL 35 python -m pytest tests/test_slepc_gpu.py -m gpu -x -q -k "svd or ex2 or smoke or test4" 2>&1 | tail -3) > gpurun_out/r02v_pytest.log 2>&1
tail -3 gpurun_out/r02v_pytest.log
