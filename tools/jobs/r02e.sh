# round 2, call e (1 GPU): gpu suite with the new kernels (Gram, SpMM, gated sweep, C1), PDL on/off in the latency regime,
# kernel timings
mkdir -p gpurun_out
(timeout -s KILL 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/r02e_pytest.log 2>&1
for pdl in 1 0; do
  (B2K_PDL=$pdl timeout -s KILL 200 python tools/tts.py --case c2 --scale 0.25 2>&1 | grep '^{' | tail -1) > gpurun_out/r02e_tts_c2_1024_pdl$pdl.log 2>&1
  (B2K_PDL=$pdl timeout -s KILL 200 python tools/tts.py --case c4 --scale 0.1 2>&1 | grep '^{' | tail -1) > gpurun_out/r02e_tts_c4_m632_pdl$pdl.log 2>&1
  (B2K_PDL=$pdl timeout -s KILL 200 python tools/tts.py --case c3 --scale 0.25 2>&1 | grep '^{' | tail -1) > gpurun_out/r02e_tts_c3_128_pdl$pdl.log 2>&1
done
(timeout -s KILL 300 python tools/kbench2.py 2>&1 | tail -30) > gpurun_out/r02e_kbench2.jsonl 2>&1
(B2K_GRAM_TMA=0 timeout -s KILL 300 python tools/kbench2.py 2>&1 | tail -30) > gpurun_out/r02e_kbench2_nogram.jsonl 2>&1
for f in gpurun_out/r02e_*; do echo "== $f"; tail -c 2500 $f | cut -c1-600; echo; done
