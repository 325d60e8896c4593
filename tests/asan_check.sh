#!/bin/bash
# Memory/UB check of the C host side on CPU: host/*.c + the CPU oracle plug-in + the example programs (through the test-only
# shim tests/ex_cpu_shim.h) built with -fsanitize=address,undefined and run with LeakSanitizer on.  Exits non-zero on any report.
#     bash tests/asan_check.sh
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$(mktemp -d)
OB=$(python -c "import scipy,os,glob;print(os.path.realpath(glob.glob(os.path.join(os.path.dirname(scipy.__file__),'..','scipy.libs','libscipy_openblas*.so'))[0]))")
SAN="-g -O1 -fsanitize=address,undefined -fno-omit-frame-pointer -std=gnu11 -Wno-unused-function -fopenmp -I$ROOT/include -I$ROOT/slepc_b200/host -I$ROOT/examples"
make -C "$ROOT" -s slepc_b200/lib/libb200krylov.so
for f in "$ROOT"/slepc_b200/host/*.c "$ROOT"/oracle/oracle_cpu.c; do gcc $SAN -c "$f" -o "$OUT/$(basename "$f" .c).o"; done
for ex in ex2 ex3 ex5 svd_test3 svd_ex8 bv_test1 bv_test2 eps_test4; do
  gcc $SAN -include "$ROOT/tests/ex_cpu_shim.h" -o "$OUT/$ex" "$ROOT/examples/$ex.c" "$OUT"/*.o -L"$ROOT/slepc_b200/lib" -lb200krylov \
      -Wl,-rpath,"$ROOT/slepc_b200/lib" "$OB" -Wl,-rpath,"$(dirname "$OB")" -lm -ldl
done
export ASAN_OPTIONS=detect_leaks=1:halt_on_error=1 UBSAN_OPTIONS=halt_on_error=1:print_stacktrace=1 OMP_NUM_THREADS=2
fail=0
run() { echo "== $*"; if ! "$@" > "$OUT/log" 2>&1 || grep -qE "ERROR: |runtime error" "$OUT/log"; then cat "$OUT/log"; fail=1; fi; }
run "$OUT/ex2" -n 72 -eps_nev 4 -eps_ncv 20 -terse
run "$OUT/ex2" -n 30 -eps_nev 3
run "$OUT/ex3" -n 72 -eps_nev 4 -eps_ncv 20 -terse
run "$OUT/ex5" -m 15 -eps_nev 4 -eps_largest_real -terse
run "$OUT/bv_test1" -verbose
run "$OUT/bv_test2"
run "$OUT/bv_test2" -bv_orthog_type mgs
run "$OUT/eps_test4"
run "$OUT/svd_ex8"
run "$OUT/svd_test3" -svd_nsv 4
run "$OUT/svd_test3" -svd_nsv 4 -svd_trlanczos_locking 0
run "$OUT/svd_test3" -svd_nsv 4 -svd_trlanczos_oneside
run "$OUT/svd_test3" -svd_nsv 4 -svd_trlanczos_oneside -bv_orthog_type mgs
run "$OUT/svd_test3" -svd_nsv 4 -svd_trlanczos_oneside -bv_orthog_refine always
rm -rf "$OUT"
[ $fail = 0 ] && echo "asan_check: clean" || { echo "asan_check: FAILED"; exit 1; }
