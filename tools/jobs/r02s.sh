mkdir -p gpurun_out
(timeout -s KILL 400 python -m pytest tests/test_spmv_parity_gpu.py tests/test_kernels_gpu.py tests/test_kernel_switches_gpu.py tests/test_slepc4py_compat_gpu.py -m gpu -x -q 2>&1 | tail -6) > gpurun_out/r02s_pytest.log 2>&1
(B2K_TIMING=1 timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-tts --no-cpu --no-lib --no-latency 2> gpurun_out/r02s_stderr.log | tail -1) > gpurun_out/r02s_bench_n1.log
tail -4 gpurun_out/r02s_pytest.log
grep b2k_csr_create_global gpurun_out/r02s_stderr.log | tail -4
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02s_bench_n1.log").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["phases_s_rank0"])
PY
