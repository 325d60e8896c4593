mkdir -p gpurun_out
(timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "spmv or laplacian or mult" 2>&1 | tail -8)
(timeout -s KILL 300 python tools/kbench.py 2>&1 | grep -E "spmv|mult_inplace|Error|error" | tail -10)
(B2K_SPMV_PIPE=0 timeout -s KILL 300 python tools/kbench.py 2>&1 | grep -E "spmv|Error|error" | tail -10)
(timeout -s KILL 300 python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e --no-tts > gpurun_out/bench_r01_e.json 2> gpurun_out/bench_r01_e.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r01_e.json')); print(d['value'], d['ms_per_step']); print({k:(round(v['avg_ms'],4), round(v['achieved_gbs'])) for k,v in d['kernels'].items()})"; tail -5 gpurun_out/bench_r01_e.err)
