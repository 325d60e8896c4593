# round 2, call f (1 GPU): gpu suite (Gram chains, SpMM through the pipeline, generalized problems on the GPU), kernel timings,
# then the DEFAULT bench run (C2 legs + C3 512^3 to convergence + baselines) exactly as the driver will run it
mkdir -p gpurun_out
(timeout -s KILL 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/r02f_pytest.log 2>&1
(timeout -s KILL 300 python tools/kbench2.py 2>&1 | tail -30) > gpurun_out/r02f_kbench2.jsonl 2>&1
(timeout -s KILL 1500 python bench.py --gpus 1 --steps 20 --warmup 5 2>&1 | tail -3) > gpurun_out/r02f_bench_n1.log 2>&1
(timeout -s KILL 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2>&1 | tail -2) > gpurun_out/r02f_bench_ref.log 2>&1
for f in gpurun_out/r02f_*; do echo "== $f"; tail -c 1500 $f | cut -c1-700; echo; done
