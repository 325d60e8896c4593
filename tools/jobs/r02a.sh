# round 2, call a (1 GPU): whole gpu suite incl. the new SpMV parity tests, the opt-in narrow restart GEMM, latency baselines
mkdir -p gpurun_out
(timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/r02a_pytest.log 2>&1
(B2K_TEST_EXPERIMENTAL=1 timeout -s KILL 300 python -m pytest tests/test_experimental_gpu.py -m gpu -q -s 2>&1 | tail -25) > gpurun_out/r02a_experimental.log 2>&1
(timeout -s KILL 200 python tools/tts.py --case c4 --scale 0.1 2>&1 | tail -3) > gpurun_out/r02a_tts_c4_m632.log 2>&1
(timeout -s KILL 200 python tools/tts.py --case c2 --scale 0.25 2>&1 | tail -3) > gpurun_out/r02a_tts_c2_1024.log 2>&1
tail -5 gpurun_out/r02a_pytest.log gpurun_out/r02a_experimental.log gpurun_out/r02a_tts_c4_m632.log gpurun_out/r02a_tts_c2_1024.log
