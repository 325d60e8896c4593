/* CPU model of the peer-memory halo protocol of slepc_b200/csrc/b2k_halo.cu (k_halo_push / k_halo_wait): every "GPU" is a
   thread, peer memory is shared memory, flags are C11 atomics.  The model runs the SAME sequence of protocol steps with
   random delays and checks that every exchange delivers exactly the values of its own sequence number — i.e. that the
   double buffering + acknowledgement rule never lets a sender overwrite a ghost buffer that is still being read, and that
   no wait can deadlock.  TEST INFRASTRUCTURE (tests/test_proto_cpu.py); not part of the product.
       gcc -O2 -pthread halo_proto.c -o halo_proto && ./halo_proto <ranks> <exchanges> <seed> <barrier_every> */
#include <pthread.h>
#include <sched.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define MAXR 8
#define CNT 5                             /* entries per (sender, receiver) pair */

typedef struct {
  int nrecv, nsend, recvrank[MAXR], sendrank[MAXR];
  int nghost;
  double *data;                           /* [2][nghost] */
  _Atomic uint64_t arrived[MAXR], ack[MAXR];
} Rank;

static Rank R[MAXR];
static int S, NX, BARRIER_EVERY;
static _Atomic int failures;
static pthread_barrier_t bar;

static double encode(int sender, uint64_t s, int i) { return (double)sender * 1e9 + (double)s * 16.0 + (double)i; }
static int slot_recv(const Rank *r, int from) { for (int p = 0; p < r->nrecv; p++) if (r->recvrank[p] == from) return p; return -1; }
static int slot_send(const Rank *r, int to) { for (int q = 0; q < r->nsend; q++) if (r->sendrank[q] == to) return q; return -1; }
static void jitter(unsigned *st) { *st = *st * 1103515245u + 12345u; int n = (*st >> 16) % 7; for (int i = 0; i < n; i++) sched_yield(); }

static void *rank_main(void *arg)
{
  const int me = (int)(intptr_t)arg;
  Rank *r = &R[me];
  unsigned st = 777u * (unsigned)(me + 1);
  for (uint64_t s = 1; s <= (uint64_t)NX; s++) {
    /* ---- push(s): acknowledgements, then per destination: wait for its ack of s-2, copy, publish ---- */
    for (int p = 0; p < r->nrecv; p++) {
      Rank *src = &R[r->recvrank[p]];
      atomic_store_explicit(&src->ack[slot_send(src, me)], s - 1, memory_order_release);
    }
    for (int q = 0; q < r->nsend; q++) {
      Rank *dst = &R[r->sendrank[q]];
      while (atomic_load_explicit(&r->ack[q], memory_order_acquire) + 2 < s) sched_yield();
      const int slot = slot_recv(dst, me);
      double *buf = dst->data + (s & 1) * (uint64_t)dst->nghost + (uint64_t)slot * CNT;
      for (int i = 0; i < CNT; i++) { buf[i] = encode(me, s, i); if (i == 2) jitter(&st); }
      atomic_store_explicit(&dst->arrived[slot], s, memory_order_release);
    }
    jitter(&st);
    /* ---- wait(s) ---- */
    for (int p = 0; p < r->nrecv; p++)
      while (atomic_load_explicit(&r->arrived[p], memory_order_acquire) < s) sched_yield();
    /* ---- the SpMV of s reads buffer s&1 (twice, with a pause in between: an overwrite in flight would show) ---- */
    for (int pass = 0; pass < 2; pass++) {
      for (int p = 0; p < r->nrecv; p++) {
        const double *buf = r->data + (s & 1) * (uint64_t)r->nghost + (uint64_t)p * CNT;
        for (int i = 0; i < CNT; i++)
          if (buf[i] != encode(r->recvrank[p], s, i)) {
            if (atomic_fetch_add(&failures, 1) < 5) fprintf(stderr, "rank %d exchange %llu: slot %d entry %d holds %.1f\n", me, (unsigned long long)s, p, i, buf[i]);
          }
      }
      jitter(&st);
    }
    /* now and then a global reduction (= barrier) sits between two MatMults, as in the Lanczos loop; in between the ranks
       run free, as in a BVMatMult column loop */
    if (BARRIER_EVERY > 0 && s % (uint64_t)BARRIER_EVERY == 0) pthread_barrier_wait(&bar);
  }
  return NULL;
}

int main(int argc, char **argv)
{
  S = argc > 1 ? atoi(argv[1]) : 4;
  NX = argc > 2 ? atoi(argv[2]) : 2000;
  unsigned seed = argc > 3 ? (unsigned)atoi(argv[3]) : 1u;
  BARRIER_EVERY = argc > 4 ? atoi(argv[4]) : 0;
  if (S < 2 || S > MAXR) return 2;
  srand(seed);
  /* a random directed communication graph that contains the slab neighbours (every rank keeps at least one peer) */
  for (int a = 0; a < S; a++)
    for (int b = 0; b < S; b++) {
      if (a == b) continue;
      const int neighbour = (b == a + 1 || b == a - 1);
      if (neighbour || rand() % 3 == 0) { R[a].sendrank[R[a].nsend++] = b; R[b].recvrank[R[b].nrecv++] = a; }
    }
  for (int a = 0; a < S; a++) {
    R[a].nghost = R[a].nrecv * CNT;
    R[a].data = (double *)calloc((size_t)(2 * R[a].nghost + 1), sizeof(double));
  }
  pthread_barrier_init(&bar, NULL, (unsigned)S);
  pthread_t th[MAXR];
  for (int a = 0; a < S; a++) pthread_create(&th[a], NULL, rank_main, (void *)(intptr_t)a);
  for (int a = 0; a < S; a++) pthread_join(th[a], NULL);
  printf("ranks %d exchanges %d seed %u barrier_every %d: %s (%d mismatches)\n", S, NX, seed, BARRIER_EVERY, failures ? "FAILED" : "ok", (int)failures);
  return failures ? 1 : 0;
}
