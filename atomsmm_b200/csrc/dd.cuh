// Device-side primitives of the peer-memory domain decomposition (dist.cu, integrate.cu).
//
// One process per GPU; every rank maps the position array and a small SIGNAL BLOCK of every peer
// into its own address space (cudaIpc handles, NVLink / NVSwitch peer access).  All cross-GPU
// traffic on the hot path is plain loads and stores on those mapped pointers from inside the
// engine's own kernels -- no library collective, no host round trip, capturable in the step graph:
//
//   positions   readers PULL: after the owners have posted "my positions of exchange e are final",
//               every rank copies the atoms of its halo (or, when a neighbour-list rebuild is due,
//               of everybody) straight out of the owners' memory, then acknowledges; an owner waits
//               for the acknowledgements before it moves its atoms again;
//   sums        every rank PUSHES its partial into a slot of every peer's signal block, then each
//               rank adds the slots in rank order (deterministic, identical on all ranks).
//
// Signals are 64-bit words written with st.release.sys and polled with ld.acquire.sys; a wait gives
// up after DD_TIMEOUT_NS and raises a flag that b2_synchronize reports (a dead peer must not hang
// the GPU).
#pragma once

#include <stdint.h>

#include "ctx.h"

#define DD_POST 0                      // [rank]  (epoch << 1) | rebuild flag, written by `rank`
#define DD_ACK 16                      // [rank]  epoch whose pull `rank` has finished
#define DD_REDF 32                     // [slot][rank] epoch of the partial in DD_REDV[slot][rank]
#define DD_REDV 64                     // [slot][rank] partial sum (double bits)
#define DD_SIG_WORDS 128
#define DD_TIMEOUT_NS 60000000000ull

struct DDPeers {
    int rank, nranks;
    int range[B2_MAX_RANKS + 1];               // ownership boundaries (atoms, engine order)
    double* x[B2_MAX_RANKS];                   // position arrays (own entry: local)
    unsigned long long* sig[B2_MAX_RANKS];     // signal blocks (own entry: local)
};

__device__ __forceinline__ unsigned long long dd_ld_acquire(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void dd_st_release(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long dd_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// spin until (*p >> shift) >= want; returns the word read (0 after a time-out, which sets state[2])
__device__ __forceinline__ unsigned long long dd_wait(const unsigned long long* p, int shift, unsigned long long want,
                                                      unsigned long long* state) {
    unsigned long long v = dd_ld_acquire(p);
    if ((v >> shift) >= want) return v;
    const unsigned long long t0 = dd_now();
    while (true) {
        v = dd_ld_acquire(p);
        if ((v >> shift) >= want) return v;
        if (dd_now() - t0 > DD_TIMEOUT_NS) {
            state[2] = 1ull;
            return 0ull;
        }
        __nanosleep(64);
    }
}

__device__ __forceinline__ int dd_owner(const DDPeers& P, int atom) {
    int o = 0;
    while (o + 1 < P.nranks && atom >= P.range[o+1]) o++;
    return o;
}

// Sum of one double over all ranks, called by ALL threads of ONE block on every rank (`lane` =
// threadIdx.x; the first P.nranks threads do the work).  Returns the total in every thread: partials are added
// in rank order, so all ranks obtain the same bits.  Double buffered by the parity of the epoch: a
// peer can be at most one reduction ahead (it cannot finish reduction e+1 without this rank's
// partial, which is sent after this rank has consumed reduction e).
__device__ __forceinline__ double dd_allreduce_sum(const DDPeers& P, unsigned long long* state, double mine, int lane,
                                                   double* shared /* [B2_MAX_RANKS] */, unsigned long long* shared_epoch) {
    if (lane == 0) *shared_epoch = state[1] + 1ull;
    __syncthreads();
    const unsigned long long e = *shared_epoch;
    const int slot = (int)(e & 1ull);
    if (lane < P.nranks) {
        unsigned long long* dst = P.sig[lane];
        dst[DD_REDV + 16*slot + P.rank] = (unsigned long long)__double_as_longlong(mine);
        __threadfence_system();
        dd_st_release(&dst[DD_REDF + 16*slot + P.rank], e);
        const unsigned long long* src = P.sig[P.rank];
        dd_wait(&src[DD_REDF + 16*slot + lane], 0, e, state);
        shared[lane] = __longlong_as_double((long long)dd_ld_acquire(&src[DD_REDV + 16*slot + lane]));
    }
    __syncthreads();
    double total = 0;
    for (int r = 0; r < P.nranks; r++) total += shared[r];
    if (lane == 0) state[1] = e;
    return total;
}
