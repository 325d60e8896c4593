# r01k (1 GPU): full GPU test suite, ncu --set full of the newest kernels, launch list of the restart cycle, the default
# bench.py run (all legs) with its wall time, the reference arm, per-kernel C3 breakdown
mkdir -p gpurun_out
nproc; free -g | head -2
(time timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) 2>&1 | grep -v "^$" | grep -v "^user\|^sys"
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_vq_tma -s 3 -c 1 -o gpurun_out/r01_vq_tma -f python tools/kbench.py > gpurun_out/ncu6.log 2>&1; tail -1 gpurun_out/ncu6.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_spmv_sell_pipe -s 13 -c 1 -o gpurun_out/r01_spmv_pipe3d -f python tools/kbench.py > gpurun_out/ncu7.log 2>&1; tail -1 gpurun_out/ncu7.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_spmv_sell_pipe -s 3 -c 1 -o gpurun_out/r01_spmv_pipe2d -f python tools/kbench.py > gpurun_out/ncu8.log 2>&1; tail -1 gpurun_out/ncu8.log
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-tts"
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_gs_tma -s 150 -c 2 -o gpurun_out/r01_gs_tma -f $B > gpurun_out/ncu9.log 2>&1; tail -1 gpurun_out/ncu9.log
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv --log-file gpurun_out/launches_r01k.csv $B > gpurun_out/ncu10.log 2>&1; tail -1 gpurun_out/ncu10.log
(time timeout -s KILL 900 python bench.py > gpurun_out/bench_r01_k.json 2> gpurun_out/bench_r01_k.err) 2>&1 | grep real
python -c "
import json; d=json.load(open('gpurun_out/bench_r01_k.json')); print(d['value'], d['ms_per_step'], d['roofline']); print({k:(round(v['avg_ms'],4), round(v['achieved_gbs'])) for k,v in d['kernels'].items()}); print(d['cpu_baseline']); print(d['e2e']); print(d['time_to_solution']); print(d['clocks'])"; tail -3 gpurun_out/bench_r01_k.err
(time timeout -s KILL 900 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_r01_k_ref.json 2> gpurun_out/bench_r01_k_ref.err) 2>&1 | grep real
cut -c1-600 gpurun_out/bench_r01_k_ref.json; tail -3 gpurun_out/bench_r01_k_ref.err
timeout -s KILL 600 python bench.py --workload c3 --steps 6 --warmup 3 --no-e2e --no-cpu --no-tts > gpurun_out/bench_r01_k_c3.json 2> gpurun_out/bench_r01_k_c3.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r01_k_c3.json')); print('c3', d['value'], d['ms_per_step'], d['lanczos_steps']); print({k:(v['launches'], round(v['avg_ms'],4), round(v['achieved_gbs']), round(v['share_of_step'],3)) for k,v in d['kernels'].items()})"; tail -3 gpurun_out/bench_r01_k_c3.err
