# r01m (2 GPUs): peer-memory one-shot reductions (k_reduce_partials_xg) against the NCCL path
mkdir -p gpurun_out
(timeout -s KILL 240 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -6)
(timeout -s KILL 120 python -m pytest tests/test_slepc_gpu.py -m gpu -x -q -k "oneside" 2>&1 | tail -3)
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout -s KILL 120 $R bench.py --gpus 2 --steps 4 --warmup 3 --no-tts --no-cpu --no-e2e > gpurun_out/bench_r01_n2_p2p.json 2> gpurun_out/bench_r01_n2_p2p.err
B2K_COMM_P2P=0 timeout -s KILL 120 $R bench.py --gpus 2 --steps 4 --warmup 3 --no-tts --no-cpu --no-e2e > gpurun_out/bench_r01_n2_nccl.json 2> gpurun_out/bench_r01_n2_nccl.err
python -c "
import json
for f in ('p2p','nccl'):
    try:
        d=[json.loads(l) for l in open('gpurun_out/bench_r01_n2_%s.json'%f) if l.startswith('{')][0]; print(f, d['value'], d['ms_per_step'], d['collectives'][:40], d['gpu_launches'])
    except Exception as e: print(f, 'FAILED', e)"
grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/bench_r01_n2_p2p.err | tail -5
timeout -s KILL 100 $R tools/tts.py --case c4 --scale 0.1 2>&1 | grep -E "^\{|Error|error" | tail -3 | tee gpurun_out/tts_c4_632_n2_p2p.json
B2K_COMM_P2P=0 timeout -s KILL 100 $R tools/tts.py --case c4 --scale 0.1 2>&1 | grep -E "^\{|Error|error" | tail -3 | tee gpurun_out/tts_c4_632_n2_nccl.json
