# round 2, call c (2 GPUs): multi-GPU tests (mailbox reductions, both halo transports, closed-form slab SpMV), latency-bound
# solves on 2 GPUs with both halos, one-sided vs two-sided SVD, quick bench with the parity leg
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(timeout -s KILL 900 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -30) > gpurun_out/r02c_pytest_mgpu.log 2>&1
(timeout -s KILL 200 $TR --master-port 29511 tools/tts.py --case c4 --scale 0.1 2>&1 | grep '^{' | tail -1) > gpurun_out/r02c_tts_c4_m632_2gpu_p2phalo.log 2>&1
(B2K_HALO_P2P=0 timeout -s KILL 200 $TR --master-port 29512 tools/tts.py --case c4 --scale 0.1 2>&1 | grep '^{' | tail -1) > gpurun_out/r02c_tts_c4_m632_2gpu_ncclhalo.log 2>&1
(timeout -s KILL 200 python tools/tts.py --case c4 --scale 0.1 2>&1 | grep '^{' | tail -1) > gpurun_out/r02c_tts_c4_m632_1gpu.log 2>&1
(timeout -s KILL 200 $TR --master-port 29513 tools/tts.py --case c2 --scale 0.25 2>&1 | grep '^{' | tail -1) > gpurun_out/r02c_tts_c2_1024_2gpu_p2phalo.log 2>&1
(B2K_HALO_P2P=0 timeout -s KILL 200 $TR --master-port 29514 tools/tts.py --case c2 --scale 0.25 2>&1 | grep '^{' | tail -1) > gpurun_out/r02c_tts_c2_1024_2gpu_ncclhalo.log 2>&1
(timeout -s KILL 300 $TR --master-port 29515 tools/tts.py --case c5 --scale 0.04 2>&1 | grep '^{' | tail -1) > gpurun_out/r02c_tts_c5_twoside_2gpu.log 2>&1
(timeout -s KILL 300 $TR --master-port 29516 tools/tts.py --case c5 --scale 0.04 --oneside 2>&1 | grep '^{' | tail -1) > gpurun_out/r02c_tts_c5_oneside_2gpu.log 2>&1
(timeout -s KILL 600 $TR --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --tts c3small 2>&1 | tail -4) > gpurun_out/r02c_bench_2gpu.log 2>&1
for f in gpurun_out/r02c_*.log; do echo "== $f"; tail -c 1200 $f; echo; done
