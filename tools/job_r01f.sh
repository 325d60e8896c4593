mkdir -p gpurun_out
(timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dmma_probe tools/dmma_probe.cu && /tmp/dmma_probe
(timeout -s KILL 300 python tools/kbench.py 2>&1 | grep -E "spmv|mult_inplace|Error|error" | tail -10)
B2K_SPMV_SELL=0 timeout -s KILL 300 python tools/kbench.py 2>&1 | grep -E "spmv" | tail -4
timeout -s KILL 300 python tools/tts.py --case c4 --scale 0.3163 2>&1 | tail -2 | tee gpurun_out/tts_c4_2000.json
(timeout -s KILL 300 python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/bench_r01_c.json 2> gpurun_out/bench_r01_c.err; tail -c 2500 gpurun_out/bench_r01_c.json; tail -5 gpurun_out/bench_r01_c.err)
