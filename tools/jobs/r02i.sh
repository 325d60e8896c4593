# round 2, call i (1 GPU): full gpu suite after the test fixes (TSQR kernels, asynchronous Krylov cycle), TSQR timings
mkdir -p gpurun_out
(timeout -s KILL 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40) > gpurun_out/r02i_pytest.log 2>&1
(timeout -s KILL 300 python tools/kbench3.py 2>&1 | tail -30) > gpurun_out/r02i_kbench3.jsonl 2>&1
for f in gpurun_out/r02i_*; do echo "== $f"; tail -c 3000 $f | cut -c1-1500; echo; done
