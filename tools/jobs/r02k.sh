# round 2, call k (1 GPU): gpu suite (register-resident TSQR kernels, adaptive asynchronous cycle), TSQR timings, latency solves
mkdir -p gpurun_out
(timeout -s KILL 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40) > gpurun_out/r02k_pytest.log 2>&1
(timeout -s KILL 300 python tools/kbench3.py 2>&1 | tail -30) > gpurun_out/r02k_kbench3.jsonl 2>&1
(timeout -s KILL 200 python tools/tts.py --case c4 --scale 0.1 2>&1 | grep '^{' | tail -1) > gpurun_out/r02k_tts_c4_m632.log 2>&1
(timeout -s KILL 200 python tools/tts.py --case c2 --scale 0.25 2>&1 | grep '^{' | tail -1) > gpurun_out/r02k_tts_c2_1024.log 2>&1
for f in gpurun_out/r02k_*; do echo "== $f"; tail -c 3000 $f | cut -c1-1500; echo; done
