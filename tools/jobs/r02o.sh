# round 2, call o (2 GPUs): the world = 2 multi-GPU tests with the local column numbering built in HBM, the new kernel-level tests,
# and the e2e leg with its phases (matrix set-up was 0.35 s of host work per rank)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(timeout -s KILL 300 python -m pytest tests/test_spmv_parity_gpu.py -m gpu -x -q -k "global_columns or transpose" 2>&1 | tail -15) > gpurun_out/r02o_pytest_new.log 2>&1
(timeout -s KILL 300 $TR --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 --no-tts --no-latency --no-parity --no-cpu --no-lib 2>&1 | tail -2) > gpurun_out/r02o_bench_n2_e2e.log 2>&1
(timeout -s KILL 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -k "2" 2>&1 | tail -30) > gpurun_out/r02o_pytest_world2.log 2>&1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02o_bench_n2_e2e.log").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"])
PY
tail -6 gpurun_out/r02o_pytest_new.log; tail -6 gpurun_out/r02o_pytest_world2.log
