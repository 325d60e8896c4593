mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
(timeout -s KILL 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -15)
(timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/bench_r01_n2.json 2> gpurun_out/bench_r01_n2.err; tail -c 3000 gpurun_out/bench_r01_n2.json; tail -5 gpurun_out/bench_r01_n2.err)
(timeout -s KILL 300 python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/bench_r01_n1b.json 2> gpurun_out/bench_r01_n1b.err; tail -c 3000 gpurun_out/bench_r01_n1b.json; tail -5 gpurun_out/bench_r01_n1b.err)
