/* CPU model of the one-shot cross-GPU reduction of slepc_b200/csrc/b2k_bv.cu (k_reduce_partials_xg): every "GPU" is a thread,
   the mailboxes are shared memory, flags are C11 atomics.  Checks, under random delays and for up to 8 ranks, that two mailbox
   parities are enough (a rank that runs ahead never overwrites a contribution that a slow rank has not read yet) and that every
   rank obtains the same, correct sum in every reduction.  TEST INFRASTRUCTURE (tests/test_proto_cpu.py); not part of the product.
       gcc -O2 -pthread reduce_proto.c -o reduce_proto && ./reduce_proto <ranks> <reductions> <seed> */
#include <pthread.h>
#include <sched.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define MAXR 8
#define NC 3                              /* columns per reduction */

typedef struct {
  double data[2][MAXR][NC];               /* [parity][source rank][column]           */
  _Atomic uint64_t flag[2][MAXR];         /* sequence number of the source's last write */
} Box;

static Box box[MAXR];
static int S, NX;
static _Atomic int failures;

static double contrib(int rank, uint64_t s, int c) { return (double)(rank + 1) * 1000.0 + (double)(s % 977) * (double)(c + 1); }
static void jitter(unsigned *st) { *st = *st * 1103515245u + 12345u; int n = (*st >> 16) % 9; for (int i = 0; i < n; i++) sched_yield(); }

static void *rank_main(void *arg)
{
  const int me = (int)(intptr_t)arg;
  unsigned st = 4242u * (unsigned)(me + 1);
  for (uint64_t s = 1; s <= (uint64_t)NX; s++) {
    const int par = (int)(s & 1);
    for (int p = 0; p < S; p++) {                              /* deliver my partial sums to every rank (myself included) */
      for (int c = 0; c < NC; c++) box[p].data[par][me][c] = contrib(me, s, c);
      if (p == S / 2) jitter(&st);
    }
    for (int p = 0; p < S; p++) atomic_store_explicit(&box[p].flag[par][me], s, memory_order_release);
    for (int p = 0; p < S; p++)                                /* wait for everybody's contribution in MY mailbox */
      while (atomic_load_explicit(&box[me].flag[par][p], memory_order_acquire) < s) sched_yield();
    jitter(&st);
    for (int c = 0; c < NC; c++) {
      double t = 0.0, ref = 0.0;
      for (int p = 0; p < S; p++) { t += box[me].data[par][p][c]; ref += contrib(p, s, c); }
      if (t != ref && atomic_fetch_add(&failures, 1) < 5) fprintf(stderr, "rank %d reduction %llu column %d: %.1f != %.1f\n", me, (unsigned long long)s, c, t, ref);
    }
    jitter(&st);
  }
  return NULL;
}

int main(int argc, char **argv)
{
  S = argc > 1 ? atoi(argv[1]) : 4;
  NX = argc > 2 ? atoi(argv[2]) : 5000;
  if (S < 1 || S > MAXR) return 2;
  pthread_t th[MAXR];
  for (int a = 0; a < S; a++) pthread_create(&th[a], NULL, rank_main, (void *)(intptr_t)a);
  for (int a = 0; a < S; a++) pthread_join(th[a], NULL);
  printf("ranks %d reductions %d: %s (%d mismatches)\n", S, NX, failures ? "FAILED" : "ok", (int)failures);
  return failures ? 1 : 0;
}
