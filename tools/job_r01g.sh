mkdir -p gpurun_out
(timeout -s KILL 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "spmv or laplacian" 2>&1 | tail -3)
(timeout -s KILL 300 python tools/kbench.py 2>&1 | grep -E "spmv|Error|error" | tail -10)
(timeout -s KILL 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -8)
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout -s KILL 300 $R tools/tts.py --case c5 --scale 0.02 2>&1 | grep -E "^\{|Error|error" | tail -3 | tee gpurun_out/tts_c5_s002_n2.json
timeout -s KILL 300 $R tools/tts.py --case c3 --scale 0.5 2>&1 | grep -E "^\{|Error|error" | tail -3 | tee gpurun_out/tts_c3_256_n2.json
timeout -s KILL 300 $R tools/tts.py --case c4 --scale 0.1 2>&1 | grep -E "^\{|Error|error" | tail -3 | tee gpurun_out/tts_c4_632_n2.json
