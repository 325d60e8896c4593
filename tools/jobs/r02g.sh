# round 2, call g (1 GPU): gpu suite (device-built transpose, relaxed ex13 check, slepc4py binding), SVD time-to-solution with the
# transpose built in HBM, ncu launch list of the bench step and ONE --set full capture of the round-2 kernels
mkdir -p gpurun_out
(timeout -s KILL 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/r02g_pytest.log 2>&1
(timeout -s KILL 300 python tools/tts.py --case c5 --scale 0.04 2>&1 | grep '^{' | tail -1) > gpurun_out/r02g_tts_c5_s004_twoside.log 2>&1
(timeout -s KILL 300 python tools/tts.py --case c5 --scale 0.04 --oneside 2>&1 | grep '^{' | tail -1) > gpurun_out/r02g_tts_c5_s004_oneside.log 2>&1
(timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02g_launches.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu --no-lib --no-tts --no-latency --no-e2e --no-parity > gpurun_out/r02g_bench_under_ncu.log 2>&1)
(timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:'k_gram_tma|k_vq_tma|k_spmv_sell_pipe|k_gs_tma|k_dotvec' \
   -c 12 -f -o gpurun_out/r02g_kernels python tools/ncu_targets.py > gpurun_out/r02g_ncu_full.log 2>&1)
ls -la gpurun_out/ | tail -12
for f in gpurun_out/r02g_pytest.log gpurun_out/r02g_tts_*.log gpurun_out/r02g_ncu_full.log; do echo "== $f"; tail -c 1500 $f | cut -c1-900; echo; done
