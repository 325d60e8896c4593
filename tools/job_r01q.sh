(timeout -s KILL 60 python -m pytest tests/test_slepc_gpu.py -m gpu -q -k full_size 2>&1 | tail -15)
