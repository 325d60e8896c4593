# round 2, call n (4 GPUs): the world = 4 variants of the multi-GPU tests (never run before), and the e2e leg with its phase timings
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
(timeout -s KILL 300 $TR --master-port 29701 bench.py --gpus 4 --steps 10 --warmup 3 --no-tts --no-latency --no-parity --no-cpu --no-lib 2>&1 | tail -2) > gpurun_out/r02n_bench_n4_e2e.log 2>&1
(timeout -s KILL 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -k "4" 2>&1 | tail -30) > gpurun_out/r02n_pytest_world4.log 2>&1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02n_bench_n4_e2e.log").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"])
PY
tail -5 gpurun_out/r02n_pytest_world4.log
