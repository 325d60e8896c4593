# round 2, call b (1 GPU): gpu suite, opt-in narrow restart GEMM, one-sync vs two-sync latency, quick bench with every new leg
mkdir -p gpurun_out
(timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/r02b_pytest.log 2>&1
(B2K_TEST_EXPERIMENTAL=1 timeout -s KILL 300 python -m pytest tests/test_experimental_gpu.py -m gpu -q -s 2>&1 | tail -25) > gpurun_out/r02b_experimental.log 2>&1
for one in 1 0; do
  (B2K_BV_ONESYNC=$one timeout -s KILL 200 python tools/tts.py --case c2 --scale 0.25 2>&1 | tail -1) > gpurun_out/r02b_tts_c2_1024_onesync$one.log 2>&1
  (B2K_BV_ONESYNC=$one timeout -s KILL 200 python tools/tts.py --case c3 --scale 0.25 2>&1 | tail -1) > gpurun_out/r02b_tts_c3_128_onesync$one.log 2>&1
done
(timeout -s KILL 600 python bench.py --steps 5 --warmup 3 --tts c3small 2>&1 | tail -3) > gpurun_out/r02b_bench_quick.log 2>&1
for f in gpurun_out/r02b_*.log; do echo "== $f"; tail -c 1500 $f; echo; done
