mkdir -p gpurun_out
(timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q 2>&1 | tail -8)
(timeout -s KILL 300 python tools/kbench.py 2>&1 | grep -E "spmv|mult_inplace|Error|error" | tail -10)
(B2K_VQ_TMA=0 B2K_SPMV_SELL=0 timeout -s KILL 300 python tools/kbench.py 2>&1 | grep -E "spmv|mult_inplace|Error|error" | tail -10)
(timeout -s KILL 300 python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_r01_d.json 2> gpurun_out/bench_r01_d.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r01_d.json')); print(d['value'], d['ms_per_step']); print({k:(round(v['avg_ms'],4), round(v['achieved_gbs'])) for k,v in d['kernels'].items()}); print(d['time_to_solution'])"; tail -5 gpurun_out/bench_r01_d.err)
