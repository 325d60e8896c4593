# round 2, call m (8 GPUs): the default bench exactly as the driver's scaling run launches it at N = 8 (C2 weak, e2e, latency leg,
# C3 512^3 strong-scaled to convergence with in-run checks, parity leg on 8 ranks), then C5 at full size (5e7 x 1e7, 8 GPUs) and the
# per-step time of C4 at full size (2e7 rows; bounded number of restarts)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
(timeout -s KILL 900 $TR --master-port 29601 bench.py --gpus 8 --steps 10 --warmup 3 2>&1 | tail -3) > gpurun_out/r02m_bench_n8.log 2>&1
(timeout -s KILL 400 $TR --master-port 29602 tools/tts.py --case c5 2>&1 | grep '^{' | tail -1) > gpurun_out/r02m_tts_c5_full_n8.log 2>&1
(timeout -s KILL 300 $TR --master-port 29603 tools/tts.py --case c4 --maxits 4000 2>&1 | grep '^{' | tail -1) > gpurun_out/r02m_tts_c4_full_n8_4000its.log 2>&1
for f in gpurun_out/r02m_*.log; do echo "== $f"; tail -c 2500 $f | cut -c1-2400; echo; done
