mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu"
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_gs_tma -s 100 -c 2 -o gpurun_out/r01_gs_tma -f $B > gpurun_out/ncu5.log 2>&1; tail -2 gpurun_out/ncu5.log
timeout -s KILL 400 python tools/tts.py --case c3 2>&1 | tail -2 | tee gpurun_out/tts_c3_512.json
timeout -s KILL 600 python tools/tts.py --case c4 2>&1 | tail -2 | tee gpurun_out/tts_c4_6324.json
timeout -s KILL 400 python tools/tts.py --case c1 2>&1 | tail -2 | tee gpurun_out/tts_c1.json
