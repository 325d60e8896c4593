mkdir -p gpurun_out
(timeout -s KILL 300 python -m pytest tests/test_slepc_gpu.py -m gpu -x -q 2>&1 | tail -3)
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu"
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_gs_rt -s 100 -c 2 -o gpurun_out/r01_gs_rt -f $B > gpurun_out/ncu1.log 2>&1; tail -2 gpurun_out/ncu1.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_dotvec -s 50 -c 1 -o gpurun_out/r01_dotvec -f $B > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_spmv_csr -s 20 -c 1 -o gpurun_out/r01_spmv -f $B > gpurun_out/ncu3.log 2>&1; tail -2 gpurun_out/ncu3.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_vq -c 1 -o gpurun_out/r01_vq -f $B > gpurun_out/ncu4.log 2>&1; tail -2 gpurun_out/ncu4.log
timeout -s KILL 300 python tools/tts.py --case c2 --scale 0.125 2>&1 | tail -2 | tee gpurun_out/tts_c2_512.json
timeout -s KILL 300 python tools/tts.py --case c2 --scale 0.25 2>&1 | tail -2 | tee gpurun_out/tts_c2_1024.json
timeout -s KILL 300 python tools/tts.py --case c3 --scale 0.25 2>&1 | tail -2 | tee gpurun_out/tts_c3_128.json
timeout -s KILL 300 python tools/tts.py --case c4 --scale 0.1 2>&1 | tail -2 | tee gpurun_out/tts_c4_632.json
timeout -s KILL 600 python tools/tts.py --case c2 2>&1 | tail -2 | tee gpurun_out/tts_c2_4096.json
