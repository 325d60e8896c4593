# round 2, call l (2 GPUs): multi-GPU tests with the reverse peer-memory halo (svd cases), C5-shaped SVD with both halos, the adaptive
# asynchronous cycle in the latency-bound solves on 2 GPUs, quick bench with the parity leg
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(timeout -s KILL 900 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -30) > gpurun_out/r02l_pytest_mgpu.log 2>&1
(timeout -s KILL 300 $TR --master-port 29541 tools/tts.py --case c5 --scale 0.04 2>&1 | grep '^{' | tail -1) > gpurun_out/r02l_tts_c5_twoside_2gpu.log 2>&1
(timeout -s KILL 300 $TR --master-port 29542 tools/tts.py --case c5 --scale 0.04 --oneside 2>&1 | grep '^{' | tail -1) > gpurun_out/r02l_tts_c5_oneside_2gpu.log 2>&1
(B2K_HALO_P2P=0 timeout -s KILL 300 $TR --master-port 29544 tools/tts.py --case c5 --scale 0.04 2>&1 | grep '^{' | tail -1) > gpurun_out/r02l_tts_c5_twoside_2gpu_ncclhalo.log 2>&1
for a in 1 0; do
  (B2K_BV_ASYNC=$a timeout -s KILL 200 $TR --master-port 2952$a tools/tts.py --case c4 --scale 0.1 2>&1 | grep '^{' | tail -1) > gpurun_out/r02l_tts_c4_m632_2gpu_async$a.log 2>&1
  (B2K_BV_ASYNC=$a timeout -s KILL 200 $TR --master-port 2953$a tools/tts.py --case c4 --scale 0.3 2>&1 | grep '^{' | tail -1) > gpurun_out/r02l_tts_c4_m1897_2gpu_async$a.log 2>&1
done
(timeout -s KILL 200 $TR --master-port 29513 tools/tts.py --case c2 --scale 0.25 2>&1 | grep '^{' | tail -1) > gpurun_out/r02l_tts_c2_1024_2gpu.log 2>&1
(timeout -s KILL 600 $TR --master-port 29543 bench.py --gpus 2 --steps 5 --warmup 3 --tts c3small 2>&1 | tail -4) > gpurun_out/r02l_bench_2gpu.log 2>&1
for f in gpurun_out/r02l_*.log; do echo "== $f"; tail -c 1500 $f | cut -c1-1200; echo; done
