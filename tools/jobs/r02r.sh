mkdir -p gpurun_out
(B2K_TIMING=1 timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-tts --no-cpu --no-lib --no-latency 2> gpurun_out/r02r_stderr.log | tail -1) > gpurun_out/r02r_bench_n1.log
grep b2k_csr_create_global gpurun_out/r02r_stderr.log | tail -8
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02r_bench_n1.log").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["phases_s_rank0"])
PY
