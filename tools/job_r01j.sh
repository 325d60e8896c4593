mkdir -p gpurun_out
(timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "mult" 2>&1 | tail -3)
(timeout -s KILL 300 python tools/kbench.py 2>&1 | grep -E "mult_inplace|Error|error" | tail -10)
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_vq_tma -s 3 -c 1 -o gpurun_out/r01_vq_tma -f python tools/kbench.py > gpurun_out/ncu6.log 2>&1; tail -2 gpurun_out/ncu6.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_spmv_sell_pipe -s 13 -c 1 -o gpurun_out/r01_spmv_pipe3d -f python tools/kbench.py > gpurun_out/ncu7.log 2>&1; tail -2 gpurun_out/ncu7.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_spmv_sell_pipe -s 3 -c 1 -o gpurun_out/r01_spmv_pipe2d -f python tools/kbench.py > gpurun_out/ncu8.log 2>&1; tail -2 gpurun_out/ncu8.log
