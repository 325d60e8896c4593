# r01n (1 GPU): final check of the round — full GPU suite, smoke(), default bench.py
mkdir -p gpurun_out
(timeout -s KILL 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6)
(timeout -s KILL 120 python __graft_entry__.py smoke 2>&1 | tail -2)
(time timeout -s KILL 300 python bench.py > gpurun_out/bench_r01_n.json 2> gpurun_out/bench_r01_n.err) 2>&1 | grep real
python -c "
import json; d=json.load(open('gpurun_out/bench_r01_n.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac']); print(d['e2e']['value'], d['cpu_baseline']['value'], d['collectives'])"; tail -3 gpurun_out/bench_r01_n.err
