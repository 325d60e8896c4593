# r01l (8 GPUs): weak-scaled C2 bench, strong-scaled C3 512^3 full solve, C5 (5e7 x 1e7 SVD) full solve, C4 full solve
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8; nproc
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
(time timeout -s KILL 400 $R bench.py --gpus 8 --steps 6 --warmup 3 --no-tts > gpurun_out/bench_r01_n8.json 2> gpurun_out/bench_r01_n8.err) 2>&1 | grep real
python -c "
import json
for l in open('gpurun_out/bench_r01_n8.json'):
    if l.startswith('{'):
        d=json.loads(l); print('c2 N=8', d['value'], d['ms_per_step'], d['e2e'], d['clocks']); print({k:(round(v['avg_ms'],4), round(v['achieved_gbs'])) for k,v in d['kernels'].items()})"; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/bench_r01_n8.err | tail -5
timeout -s KILL 200 $R tools/tts.py --case c5 --scale 0.02 2>&1 | grep -E "^\{|Error|error" | tail -3 | tee gpurun_out/tts_c5_s002_n8.json
(time timeout -s KILL 500 $R tools/tts.py --case c3 2>&1 | grep -E "^\{|Error|error" | tail -3 | tee gpurun_out/tts_c3_512_n8.json) 2>&1 | grep -v "^user\|^sys\|^$"
(time timeout -s KILL 600 $R tools/tts.py --case c5 2>&1 | grep -E "^\{|Error|error|Killed" | tail -3 | tee gpurun_out/tts_c5_full_n8.json) 2>&1 | grep -v "^user\|^sys\|^$"
(time timeout -s KILL 400 $R tools/tts.py --case c4 2>&1 | grep -E "^\{|Error|error|Killed" | tail -3 | tee gpurun_out/tts_c4_full_n8.json) 2>&1 | grep -v "^user\|^sys\|^$"
