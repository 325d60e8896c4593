# round 2, call p (1 GPU): the final tree as the driver will run it: smoke(), the gpu suite, the bench line (without the 282 s C3 leg)
mkdir -p gpurun_out
(timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > gpurun_out/r02p_smoke.log 2>&1
(timeout -s KILL 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r02p_pytest.log 2>&1
(timeout -s KILL 400 python bench.py --no-tts 2>&1 | tail -1) > gpurun_out/r02p_bench_n1_notts.log 2>&1
cat gpurun_out/r02p_smoke.log; tail -4 gpurun_out/r02p_pytest.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02p_bench_n1_notts.log").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["phases_s_rank0"], "roofline", d["roofline"]["frac"], "cpu", d["cpu_baseline"]["value"], "lat", d["latency_leg"]["seconds"])
PY
