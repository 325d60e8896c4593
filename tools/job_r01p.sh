# r01p (1 GPU): full GPU suite of the final tree + the library baseline (what SLEPc's CUDA BV back-end would run: cuBLAS / cuSPARSE)
mkdir -p gpurun_out
(timeout -s KILL 170 python -m pytest tests -m gpu -x -q 2>&1 | tail -4)
(timeout -s KILL 100 python tools/cublas_ref.py 2>/dev/null > gpurun_out/cublas_ref.jsonl; cat gpurun_out/cublas_ref.jsonl | cut -c1-200)
