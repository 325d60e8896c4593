# round 2, call h (1 GPU): gpu suite (TSQR kernels, asynchronous Krylov cycle vs step-by-step path), latency-regime solves with
# B2K_BV_ASYNC on/off, default bench without the long legs
mkdir -p gpurun_out
(timeout -s KILL 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40) > gpurun_out/r02h_pytest.log 2>&1
for a in 1 0; do
  (B2K_BV_ASYNC=$a timeout -s KILL 200 python tools/tts.py --case c2 --scale 0.25 2>&1 | grep '^{' | tail -1) > gpurun_out/r02h_tts_c2_1024_async$a.log 2>&1
  (B2K_BV_ASYNC=$a timeout -s KILL 200 python tools/tts.py --case c4 --scale 0.1 2>&1 | grep '^{' | tail -1) > gpurun_out/r02h_tts_c4_m632_async$a.log 2>&1
  (B2K_BV_ASYNC=$a timeout -s KILL 200 python tools/tts.py --case c3 --scale 0.25 2>&1 | grep '^{' | tail -1) > gpurun_out/r02h_tts_c3_128_async$a.log 2>&1
done
(timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-tts --no-cpu 2>&1 | tail -2) > gpurun_out/r02h_bench_quick.log 2>&1
for f in gpurun_out/r02h_*; do echo "== $f"; tail -c 2500 $f | cut -c1-1500; echo; done
