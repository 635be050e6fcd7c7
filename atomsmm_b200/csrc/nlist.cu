// Cell binning and cluster neighbour lists (kernels K1, K2, K8 of SURVEY 2.1).
//
// Replaces the neighbour search OpenMM performs for CustomNonbondedForce / NonbondedForce
// (reference call sites: forces.py:225,153; exclusions forces.py:310-312).  The Reference
// platform rebuilds an exact list at every evaluation; here a Verlet list with a skin is kept and
// rebuilt on the device, without host synchronisation, when any atom has moved more than skin/2.
//
// List format: for every i-group of 8 consecutive atoms (storage order is spatially sorted, so a
// group is a compact cluster) the list holds every atom j that lies within cutoff+skin of ANY of
// the 8 atoms (exact union of spheres, tested in float64), as 32-bit entries
//     entry = (exclusion mask over the 8 i-atoms) << 24 | j
// so all exclusion logic (forces.py:310-312; self pairs; padding) is resolved at build time and
// costs the pair kernel one shift + test.  Entries are in deterministic order (cells in x-fastest
// order, atoms by index inside a cell).
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <type_traits>

#include "ctx.h"
#include "packed.cuh"

#define FULL 0xffffffffu

struct BuildArgs {
    int nlists;
    double rlist2[B2_MAX_LISTS];
    double rlist[B2_MAX_LISTS];
    double rcore[B2_MAX_LISTS];      // entries beyond this distance from every atom of the i-group go to the list's tail
    float rl2f[B2_MAX_LISTS], rc2f[B2_MAX_LISTS];   // (rlist + margin)^2, (rcore + margin)^2 in the sweep's fp32 arithmetic
    int* entries[B2_MAX_LISTS];
    int* counts[B2_MAX_LISTS];
    unsigned char* gflags[B2_MAX_LISTS];
    int cap[B2_MAX_LISTS];
};

struct Grid {
    double box[3], inv[3], cs_inv[3];
    int nc[3];
    double rmax;
};

__device__ __forceinline__ double wrap1(double x, double L, double invL) {
    double w = x - L*floor(x*invL);
    return w >= L ? w - L : w;
}

__device__ __forceinline__ bool is_excluded(int oi, int oj, unsigned long long mask, const int* excl_ptr,
                                            const int* excl_idx) {
    int d = oj - oi;
    if (d >= -32 && d < 32) return (mask >> (d + 32)) & 1ull;
    if (excl_ptr) {
        for (int k = excl_ptr[oi]; k < excl_ptr[oi+1]; k++)
            if (excl_idx[k] == oj) return true;
    }
    return false;
}

// K8: skin test, one pass over x and xref (48 B/atom), before every pair-force evaluation.  The same pass
// refreshes the fixed-point copy of the positions that the pair tiles read (16 B/atom written).
__global__ void k_skin_check(int lo, int n, const double* __restrict__ x, const double* __restrict__ xref, double limit2,
                             int* flags, int have_ref, int4* __restrict__ xq, double sx, double sy, double sz) {
    int i = lo + blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double px = x[3*i], py = x[3*i+1], pz = x[3*i+2];
    xq[i] = make_int4(b2_to_fixed(px, sx), b2_to_fixed(py, sy), b2_to_fixed(pz, sz), 0);
    if (have_ref) {
        double dx = px - xref[3*i], dy = py - xref[3*i+1], dz = pz - xref[3*i+2];
        if (dx*dx + dy*dy + dz*dz > limit2) flags[0] = 1;
    } else {
        flags[0] = 1;
    }
}

// Geometry of every 8-atom group (8 lanes per group): bounding box of the group's atoms unwrapped
// around its first atom -> centre (wrapped into the box, fp32), half extents, atom positions
// RELATIVE to the centre (fp32 of a < 1 nm vector: 6e-8 nm regardless of the box size), and the
// cell of the centre.  Groups, not atoms, are binned: 8x less sorting work.
__global__ void k_group_geom(int n, int ngroups, const double* __restrict__ x, Grid g, float4* __restrict__ prel,
                             float4* __restrict__ gcen, float4* __restrict__ ghalf, int* __restrict__ gcell,
                             int* cell_count, int* hmax_bits, float fat_limit, int* __restrict__ fat_list,
                             int fat_capacity, int* flags) {
    if (!flags[0]) return;
    const int t = blockIdx.x*blockDim.x + threadIdx.x;
    const int grp = t >> 3, a = t & 7;
    const bool live = grp < ngroups;
    const int i = min(min(grp, ngroups - 1)*B2_GROUP + a, n - 1);    // padding duplicates the last atom
    const int lane = threadIdx.x & 31, seg = lane & ~7;
    double u[3];
    float lo[3], hi[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const double p = x[3*i+d];
        const double r0 = __shfl_sync(FULL, p, seg);
        double dd = p - r0;
        dd -= g.box[d]*rint(dd*g.inv[d]);
        u[d] = dd;
        float l = (float)dd, h = (float)dd;
        // round outwards so that the fp32 box contains the fp64 point
        l = l > dd ? nextafterf(l, -1e30f) : l;
        h = h < dd ? nextafterf(h, 1e30f) : h;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            l = fminf(l, __shfl_xor_sync(FULL, l, o));
            h = fmaxf(h, __shfl_xor_sync(FULL, h, o));
        }
        lo[d] = l; hi[d] = h;
    }
    double r0[3];
#pragma unroll
    for (int d = 0; d < 3; d++) r0[d] = __shfl_sync(FULL, x[3*i+d], seg);
    float c[3], h[3];
    int cell[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const double cu = 0.5*((double)lo[d] + (double)hi[d]);
        h[d] = 0.5f*(hi[d] - lo[d]) + 2e-6f;
        u[d] -= cu;
        const double cw = wrap1(r0[d] + cu, g.box[d], g.inv[d]);
        c[d] = (float)cw;
        cell[d] = min(max((int)(cw*g.cs_inv[d]), 0), g.nc[d] - 1);
    }
    float hm[3] = {0.f, 0.f, 0.f};          // half extents of this thread's group if it goes into the grid
    if (live) {
        prel[grp*B2_GROUP + a] = make_float4((float)u[0], (float)u[1], (float)u[2], 0.f);
        if (a == 0) {
            gcen[grp] = make_float4(c[0], c[1], c[2], 0.f);
            ghalf[grp] = make_float4(h[0], h[1], h[2], 0.f);
            // "fat" groups (atoms of molecules that have drifted apart, or a large molecule) stay out of the
            // cells: every i-group tests them directly, so they do not inflate everybody's search region
            if (fmaxf(h[0], fmaxf(h[1], h[2])) > fat_limit) {
                gcell[grp] = -1;
                const int slot = atomicAdd(&flags[11], 1);       // arrival order; sorted by k_cell_scan
                if (slot < fat_capacity) fat_list[slot] = grp;
            } else {
                const int cidx = (cell[2]*g.nc[1] + cell[1])*g.nc[0] + cell[0];
                gcell[grp] = cidx;
                atomicAdd(&cell_count[cidx], 1);
                hm[0] = h[0]; hm[1] = h[1]; hm[2] = h[2];
            }
        }
    }
    // largest half extent of any gridded group: reduced over the block first (one atomic per block and
    // axis instead of one per group -- 1.6 M atomics on three addresses at 4.2 M atoms)
    __shared__ float smax[8][3];
#pragma unroll
    for (int d = 0; d < 3; d++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) hm[d] = fmaxf(hm[d], __shfl_xor_sync(FULL, hm[d], o));
    if (lane == 0)
        for (int d = 0; d < 3; d++) smax[threadIdx.x >> 5][d] = hm[d];
    __syncthreads();
    if (threadIdx.x < 3) {
        float m = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) m = fmaxf(m, smax[w][threadIdx.x]);
        // m > 0: int order = float order.  Read first: after the first few blocks hardly any block raises the maximum,
        // and 16 k blocks hammering three addresses with atomics serialise in the L2 (the kernel's former bottleneck)
        if (m > 0.f && __float_as_int(m) > *(volatile int*)&hmax_bits[threadIdx.x])
            atomicMax(&hmax_bits[threadIdx.x], __float_as_int(m));
    }
}

// single-block exclusive scan (ncells <= a few 1e5); also resets the fill cursors
__global__ void k_cell_scan(int ncells, int* cell_count, int* cell_start, int ngroups, const int* __restrict__ gcell,
                            int* __restrict__ fat_list, int* flags) {
    if (!flags[0]) return;
    __shared__ int part[1024];
    int t = threadIdx.x;
    int per = (ncells + blockDim.x - 1)/blockDim.x;
    int lo = t*per, hi = min(lo + per, ncells);
    int s = 0;
    for (int k = lo; k < hi; k++) s += cell_count[k];
    part[t] = s;
    __syncthreads();
    for (int off = 1; off < blockDim.x; off <<= 1) {
        int v = t >= off ? part[t-off] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    int run = part[t] - s;
    for (int k = lo; k < hi; k++) {
        int c = cell_count[k];
        cell_start[k] = run;
        cell_count[k] = 0;      // reused as cursor by k_cell_fill
        run += c;
    }
    if (t == blockDim.x - 1) cell_start[ncells] = part[t];
    // the fat groups arrive in arbitrary order: sort them by index (deterministic lists).  They are rare
    // (< 0.1 % of the groups), so an odd-even transposition sort by one block is enough; in the unlikely
    // case that there are more than the block can sort, fall back to a compaction pass over all groups.
    const int nfat = flags[11];
    if (nfat <= 2*(int)blockDim.x) {
        // rank sort in shared memory (keys are distinct group ids): position = number of smaller keys
        __shared__ int keys[2048];
        for (int k = t; k < nfat; k += blockDim.x) keys[k] = fat_list[k];
        __syncthreads();
        for (int k = t; k < nfat; k += blockDim.x) {
            const int key = keys[k];
            int rank = 0;
            for (int m = 0; m < nfat; m++) rank += keys[m] < key ? 1 : 0;
            fat_list[rank] = key;
        }
        if (t == 0) { flags[8] = nfat; flags[11] = 0; }
        return;
    }
    __shared__ int wsum[32];
    __shared__ int base;
    if (t == 0) base = 0;
    __syncthreads();
    for (int g0 = 0; g0 < ngroups; g0 += blockDim.x) {
        const int grp = g0 + t;
        const bool fat = grp < ngroups && gcell[grp] < 0;
        const unsigned ballot = __ballot_sync(FULL, fat);
        const int lane = t & 31, w = t >> 5;
        if (lane == 0) wsum[w] = __popc(ballot);
        __syncthreads();
        int before = 0, total = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) {
            if (k < w) before += wsum[k];
            total += wsum[k];
        }
        if (fat) fat_list[base + before + __popc(ballot & ((1u << lane) - 1u))] = grp;
        __syncthreads();
        if (t == 0) base += total;
        __syncthreads();
    }
    if (t == 0) { flags[8] = base; flags[11] = 0; }
}

__global__ void k_cell_fill(int ngroups, const int* gcell, const int* cell_start, int* cell_count, int* cell_groups,
                            const int* flags) {
    if (!flags[0]) return;
    int grp = blockIdx.x*blockDim.x + threadIdx.x;
    if (grp >= ngroups) return;
    int c = gcell[grp];
    if (c < 0) return;
    int slot = atomicAdd(&cell_count[c], 1);
    cell_groups[cell_start[c] + slot] = grp;
}

// deterministic order inside each cell (the atomics above fill in arbitrary order), then the
// cell-ordered copies of the group geometry that the list build streams; resets the counters
__global__ void k_cell_sort_pack(int ncells, const int* __restrict__ cell_start, int* cell_groups, int* cell_count,
                                 const float4* __restrict__ gcen, const float4* __restrict__ ghalf,
                                 float4* __restrict__ cgc, float4* __restrict__ cgh, const int* flags) {
    if (!flags[0]) return;
    int c = blockIdx.x*blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    cell_count[c] = 0;          // ready for the next rebuild
    int lo = cell_start[c], hi = cell_start[c+1];
    for (int a = lo + 1; a < hi; a++) {
        int key = cell_groups[a];
        int b = a - 1;
        while (b >= lo && cell_groups[b] > key) { cell_groups[b+1] = cell_groups[b]; b--; }
        cell_groups[b+1] = key;
    }
    for (int a = lo; a < hi; a++) {
        const int grp = cell_groups[a];
        float4 cc = gcen[grp];
        cc.w = __int_as_float(grp);
        cgc[a] = cc;
        cgh[a] = ghalf[grp];
    }
}

// K2: one warp per i-group, two levels.
//  (1) group level: walk the cells whose group centres can matter (the i-group's box grown by the list
//      radius and the largest half extent of any group), one candidate j-GROUP per lane, box-box
//      distance test -> survivors go to a small per-warp queue;
//  (2) atom level: four queued j-groups (32 atoms) per sweep, each lane tests its j-atom against the 8
//      i-atoms in fp32 on centre-relative coordinates with a safety margin (the list only has to be
//      a superset of the pairs within cutoff+skin; the interacting set is decided exactly by the pair
//      kernels), resolves the exclusion mask and appends to each list by ballot.
#define NL_MARGIN 3e-4f
#define NL_WARPS 4
#define NL_QUEUE 40
// tuning knobs (measured on B200: profiles/round2_pair_variants.txt)
#ifndef B2_NL_MINB
#define B2_NL_MINB 6            // resident blocks per SM the register budget is sized for
#endif
#ifndef B2_NL_FLAT
#define B2_NL_FLAT 1            // group level: run table + flattened candidate walk (0 = nested row walk)
#endif
#ifndef B2_NL_PACKED
#define B2_NL_PACKED 1          // f32x2 distance test in the atom-level sweep
#endif
__global__ void __launch_bounds__(32*NL_WARPS, B2_NL_MINB) k_build_lists(int n, int g_lo, int ngroups, Grid g,
                                                             const int* __restrict__ cell_start,
                                                             const float4* __restrict__ cgc,
                                                             const float4* __restrict__ cgh,
                                                             const float4* __restrict__ gcen,
                                                             const float4* __restrict__ ghalf,
                                                             const float4* __restrict__ prel,
                                                             const int* __restrict__ hmax_bits,
                                                             const int* __restrict__ fat_list,
                                                             const int* __restrict__ orig,
                                                             const unsigned long long* __restrict__ exmask,
                                                             const int* __restrict__ excl_ptr,
                                                             const int* __restrict__ excl_idx, int excl_span,
                                                             BuildArgs a, int* flags,
                                                             unsigned char* __restrict__ halo_mark, int own_lo, int own_hi) {
    if (!flags[0]) return;
    const int warp = g_lo + ((blockIdx.x*blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    if (warp >= ngroups) return;
    __shared__ float4 sxi[NL_WARPS][B2_GROUP];       // one 16-byte broadcast read per i-atom in the sweep
    __shared__ int soi[NL_WARPS][B2_GROUP];
    __shared__ unsigned long long smask[NL_WARPS][B2_GROUP];
    __shared__ float4 queue[NL_WARPS][NL_QUEUE];
    // the NEGATED i positions as four atom PAIRS {(-x0,-x1,-y0,-y1), (-z0,-z1)} for the packed f32x2 distance test
    __shared__ float4 sP[NL_WARPS][B2_GROUP/2];
    __shared__ float2 sQ[NL_WARPS][B2_GROUP/2];
    // run table of the group-level search: 32 rows per batch, up to two runs per row
    __shared__ int seg_cb[NL_WARPS][64];
    __shared__ int seg_pre[NL_WARPS][65];
    __shared__ float4 seg_shift[NL_WARPS][64];
    __shared__ int* s_ebase[NL_WARPS][B2_MAX_LISTS];     // this group's slice of every list
    __shared__ int s_good[NL_WARPS][B2_MAX_LISTS];
    const int i0 = warp*B2_GROUP;
    const float box[3] = {(float)g.box[0], (float)g.box[1], (float)g.box[2]};
    const float ibox[3] = {(float)g.inv[0], (float)g.inv[1], (float)g.inv[2]};
    if (lane < B2_GROUP) {
        const float4 pi = prel[i0 + lane];
        sxi[wib][lane] = pi;
        float* P = reinterpret_cast<float*>(&sP[wib][lane >> 1]) + (lane & 1);
        P[0] = -pi.x; P[2] = -pi.y;
        reinterpret_cast<float*>(&sQ[wib][lane >> 1])[lane & 1] = -pi.z;
        const int i = i0 + lane;
        soi[wib][lane] = i < n ? orig[i] : -1;
        smask[wib][lane] = i < n ? exmask[i] : 0ull;
    }
    __syncwarp();
    const float4 ci4 = gcen[warp], hi4 = ghalf[warp];
    const float ci[3] = {ci4.x, ci4.y, ci4.z}, hi[3] = {hi4.x, hi4.y, hi4.z};
    unsigned padmask = 0;
    for (int k = 0; k < B2_GROUP; k++)
        if (soi[wib][k] < 0) padmask |= 1u << k;
    // only atoms within excl_span positions (engine order) of the group can be excluded partners
    // or the atoms themselves: excl_span is the largest distance, in the engine's order, between
    // the two atoms of any exclusion
    const int j_first = i0 - excl_span, j_last = i0 + B2_GROUP - 1 + excl_span;
    const float rmax = (float)g.rmax + NL_MARGIN;
    const float cs[3] = {(float)(g.box[0]/g.nc[0]), (float)(g.box[1]/g.nc[1]), (float)(g.box[2]/g.nc[2])};
    // the i-box grown by the largest half extent any j-group can have: a j-group can only matter if
    // its CENTRE is within rmax of this grown box
    float lo[3], up[3];
    int c_lo[3], c_n[3];
    bool all[3];
    for (int d = 0; d < 3; d++) {
        const float grow = __int_as_float(hmax_bits[d]) + 1e-5f;
        lo[d] = ci[d] - hi[d] - grow;
        up[d] = ci[d] + hi[d] + grow;
        const int a0 = (int)floorf((lo[d] - rmax)/cs[d]);
        const int a1 = (int)floorf((up[d] + rmax)/cs[d]);
        const int cover = a1 - a0 + 1;
        all[d] = cover >= g.nc[d];
        if (all[d]) { c_lo[d] = 0; c_n[d] = g.nc[d]; }
        else { c_lo[d] = a0; c_n[d] = cover; }
    }
    if (lane == 0) {
        for (int k = 0; k < a.nlists; k++) {
            unsigned char f = 0;
            for (int d = 0; d < 3; d++)
                if ((double)(2.f*hi[d]) + a.rlist[k] > 0.5*g.box[d] - 1e-4) f = 1;
            a.gflags[k][warp] = f;
        }
    }
    // Every list is written in two parts: CORE entries (within cutoff + delta of some atom of the i-group at build
    // time) from the front, SHELL entries (only within cutoff + skin) from the back of the group's capacity, joined
    // at the end.  All entries are tested by the pair tiles as before -- but the shell entries, which are outside
    // the cutoff of all eight i-atoms until something has moved by delta, sit together at the tail, where whole
    // tile steps find no pair inside the cutoff and skip the force arithmetic (a warp-uniform branch).
    // (squared radii come precomputed from the host and the lists' base pointers live in shared memory: with 80
    // registers the compiler otherwise re-derives both at every emission)
    int scount[B2_MAX_LISTS] = {0, 0, 0, 0};
    if (lane < B2_MAX_LISTS) {
        const int k = lane < a.nlists ? lane : 0;
        s_ebase[wib][lane] = a.entries[k] + (size_t)warp*a.cap[k];
        s_good[wib][lane] = -1;          // core entries in place when the list first overflowed (-1: it has not)
    }
    __syncwarp();
    const float rmax2 = rmax*rmax;
    int count[B2_MAX_LISTS] = {0, 0, 0, 0};
    const unsigned lt = (1u << lane) - 1u;
    int qn = 0;

    // atom-level sweep over queue[0 .. nq): four j-groups per pass
    bool wide = false;       // queue entries are minimum-image centre differences of fat groups
    // MI: apply the minimum image per atom pair (small boxes, fat groups)
    auto sweep_t = [&](int nq, auto mi_tag) {
        constexpr bool MI = decltype(mi_tag)::value;
        for (int q0 = 0; q0 < nq; q0 += 4) {
            const int q = q0 + (lane >> 3);
            const bool valid = q < nq;
            const float4 e = queue[wib][valid ? q : 0];
            const int j = __float_as_int(e.w)*B2_GROUP + (lane & 7);
            const bool have = valid && j < n;
            float d2min = 1e30f;
            unsigned m = padmask;
            if (have) {
                const float4 pj = prel[j];
                const float xj = e.x + pj.x, yj = e.y + pj.y, zj = e.z + pj.z;
                if (MI || !B2_NL_PACKED) {
#pragma unroll
                    for (int k = 0; k < B2_GROUP; k++) {
                        const float4 pi = sxi[wib][k];
                        float dx = xj - pi.x, dy = yj - pi.y, dz = zj - pi.z;
                        if (MI) {
                            if (all[0] || wide) dx -= box[0]*rintf(dx*ibox[0]);
                            if (all[1] || wide) dy -= box[1]*rintf(dy*ibox[1]);
                            if (all[2] || wide) dz -= box[2]*rintf(dz*ibox[2]);
                        }
                        d2min = fminf(d2min, dx*dx + dy*dy + dz*dz);
                    }
                } else {
                    // the common case: two i-atoms per instruction issue (FADD2 / FMUL2 / FFMA2)
                    const F2 xj2 = f2(xj), yj2 = f2(yj), zj2 = f2(zj);
#pragma unroll
                    for (int k = 0; k < B2_GROUP/2; k++) {
                        const float4 P = sP[wib][k];
                        const float2 Q = sQ[wib][k];
                        const F2 dx = xj2 + f2(P.x, P.y), dy = yj2 + f2(P.z, P.w), dz = zj2 + f2(Q.x, Q.y);
                        const F2 d2 = fma2(dz, dz, fma2(dy, dy, dx*dx));
                        d2min = fminf(d2min, fminf(d2.v.x, d2.v.y));
                    }
                }
                if (j >= j_first && j <= j_last) {
                    const int oj = orig[j];
                    for (int k = 0; k < B2_GROUP; k++) {
                        const int oi = soi[wib][k];
                        if (oi < 0) continue;
                        const int dd = oj - oi;
                        bool ex = dd == 0;
                        if (dd >= -32 && dd < 32) ex = ex || ((smask[wib][k] >> (dd + 32)) & 1ull);
                        else if (excl_ptr) ex = is_excluded(oi, oj, 0ull, excl_ptr, excl_idx);
                        if (ex) m |= 1u << k;
                    }
                }
            }
            const int entry = (int)((m << 24) | (unsigned)j);
#pragma unroll
            for (int k = 0; k < B2_MAX_LISTS; k++) {
                if (k >= a.nlists) break;
                const bool in = have && d2min < a.rl2f[k];
                const bool core = d2min < a.rc2f[k];           // rc2 <= rl2: core implies in (for lanes that have an atom)
                const unsigned ballot = __ballot_sync(FULL, in);
                if (ballot == 0u) continue;                    // warp-uniform: nothing of this sweep belongs to list k
                const unsigned cballot = __ballot_sync(FULL, in && core);
                const unsigned sballot = ballot & ~cballot;
                const int nc = __popc(cballot), ns = __popc(sballot);
                if (count[k] + scount[k] + nc + ns <= a.cap[k]) {
                    if (in) {
                        const int pos = core ? count[k] + __popc(cballot & lt)
                                             : a.cap[k] - 1 - (scount[k] + __popc(sballot & lt));
                        s_ebase[wib][k][pos] = entry;
                    }
                } else if (lane == 0 && s_good[wib][k] < 0) {
                    // capacity exceeded (reported through flags[1]; the host re-fits the lists): from here on nothing
                    // is written, and the list handed to the pair tiles ends with the core entries written so far --
                    // a contiguous, fully initialised prefix (the gap between the two parts never is)
                    s_good[wib][k] = count[k];
                }
                count[k] += nc;
                scount[k] += ns;
            }
        }
    };
    const bool any_all = all[0] || all[1] || all[2];
    auto sweep = [&](int nq) {
        if (any_all || wide) sweep_t(nq, std::true_type());
        else sweep_t(nq, std::false_type());
    };

#if B2_NL_FLAT
    // Group level.  The search region is a set of ROWS of cells (fixed y/z cell, a range of x cells trimmed to the
    // sphere cross-section); a row is one or two contiguous runs of the cell-ordered arrays (two when it crosses the
    // periodic boundary).  The rows' runs are tabulated first -- one row per lane, so the ~140 instructions of index
    // and trimming arithmetic per row are paid once per 32 rows instead of once per row by the whole warp -- and the
    // candidates are then walked as ONE flattened index space over all runs: every step tests 32 candidates with all
    // lanes busy, however short the individual runs are (which is what lets the cells be small in y and z too).
    const int nrows = c_n[1]*c_n[2];
    for (int r0 = 0; r0 < nrows; r0 += 32) {
        const int r = r0 + lane;
        int cb0 = 0, len0 = 0, cb1 = 0, len1 = 0;
        float sx0 = 0.f, sx1 = 0.f, sy = 0.f, sz = 0.f;
        if (r < nrows) {
            const int cz = r/c_n[1], cy = r - cz*c_n[1];
            const int uz = c_lo[2] + cz;
            int iz = uz % g.nc[2]; if (iz < 0) iz += g.nc[2];
            sz = all[2] ? 0.f : (float)((uz - iz)/g.nc[2])*box[2];      // periodic shift of this cell layer
            const float gz = all[2] ? 0.f : fmaxf(0.f, fmaxf(lo[2] - (uz + 1)*cs[2], uz*cs[2] - up[2]));
            const int uy = c_lo[1] + cy;
            int iy = uy % g.nc[1]; if (iy < 0) iy += g.nc[1];
            sy = all[1] ? 0.f : (float)((uy - iy)/g.nc[1])*box[1];
            const float gy = all[1] ? 0.f : fmaxf(0.f, fmaxf(lo[1] - (uy + 1)*cs[1], uy*cs[1] - up[1]));
            const float rem2 = rmax2 - gz*gz - gy*gy;
            if (rem2 >= 0.f) {
                // trim the x-range of this row of cells to the sphere cross-section
                int u_first = c_lo[0], u_last = c_lo[0] + c_n[0] - 1;
                if (!all[0]) {
                    const float reach = sqrtf(rem2);
                    u_first = max(u_first, (int)floorf((lo[0] - reach)/cs[0]));
                    u_last = min(u_last, (int)floorf((up[0] + reach)/cs[0]));
                }
                if (u_first <= u_last) {
                    const int row = (iz*g.nc[1] + iy)*g.nc[0];
                    const int wrap = u_first >= 0 ? u_first/g.nc[0] : -((-u_first + g.nc[0] - 1)/g.nc[0]);
                    const int ix0 = u_first - wrap*g.nc[0];
                    const int run_end = min(u_last, u_first + (g.nc[0] - 1 - ix0));
                    cb0 = cell_start[row + ix0];
                    len0 = cell_start[row + ix0 + (run_end - u_first) + 1] - cb0;
                    sx0 = all[0] ? 0.f : (float)wrap*box[0];
                    if (run_end < u_last) {          // the row continues on the other side of the periodic boundary
                        const int run_end2 = min(u_last, run_end + g.nc[0]);
                        cb1 = cell_start[row];
                        len1 = cell_start[row + (run_end2 - run_end - 1) + 1] - cb1;
                        sx1 = (float)(wrap + 1)*box[0];
                    }
                }
            }
        }
        __syncwarp();        // the previous batch's table has been consumed by every lane
        seg_cb[wib][2*lane] = cb0; seg_cb[wib][2*lane+1] = cb1;
        seg_shift[wib][2*lane] = make_float4(sx0, sy, sz, 0.f);
        seg_shift[wib][2*lane+1] = make_float4(sx1, sy, sz, 0.f);
        int incl = len0 + len1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up_ = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += up_;
        }
        const int excl = incl - (len0 + len1);
        seg_pre[wib][2*lane] = excl; seg_pre[wib][2*lane+1] = excl + len0;
        if (lane == 31) seg_pre[wib][64] = incl;
        const int total = __shfl_sync(FULL, incl, 31);
        __syncwarp();
        int seg = 0;                                     // this lane's cursor into the run table (k only grows)
        for (int kbase = 0; kbase < total; kbase += 32) {
            const int k = kbase + lane;
            bool pass = false;
            float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < total) {
                while (seg_pre[wib][seg + 1] <= k) seg++;
                const int idx = seg_cb[wib][seg] + (k - seg_pre[wib][seg]);
                const float4 sh = seg_shift[wib][seg];
                const float4 cj = cgc[idx], hj = cgh[idx];
                float dx = cj.x + sh.x - ci[0], dy = cj.y + sh.y - ci[1], dz = cj.z + sh.z - ci[2];
                if (all[0]) dx -= box[0]*rintf(dx*ibox[0]);
                if (all[1]) dy -= box[1]*rintf(dy*ibox[1]);
                if (all[2]) dz -= box[2]*rintf(dz*ibox[2]);
                const float gx = fmaxf(fabsf(dx) - (hi[0] + hj.x), 0.f);
                const float gyy = fmaxf(fabsf(dy) - (hi[1] + hj.y), 0.f);
                const float gzz = fmaxf(fabsf(dz) - (hi[2] + hj.z), 0.f);
                pass = gx*gx + gyy*gyy + gzz*gzz < rmax2;
                e = make_float4(dx, dy, dz, cj.w);
                // domain decomposition: a candidate group that this rank does not own entirely is halo
                if (halo_mark && pass) {
                    const int jg = __float_as_int(cj.w);
                    if (jg*B2_GROUP < own_lo || (jg + 1)*B2_GROUP > own_hi) halo_mark[jg] = 1;
                }
            }
            const unsigned ballot = __ballot_sync(FULL, pass);
            if (pass) queue[wib][qn + __popc(ballot & lt)] = e;
            qn += __popc(ballot);
            __syncwarp();
            if (qn >= 4) {
                const int full = qn & ~3;
                sweep(full);
                // move the (< 4) left-over entries to the head of the queue
                const int rest = qn - full;
                float4 keep = make_float4(0.f, 0.f, 0.f, 0.f);
                if (lane < rest) keep = queue[wib][full + lane];
                __syncwarp();
                if (lane < rest) queue[wib][lane] = keep;
                qn = rest;
                __syncwarp();
            }
        }
    }
#else
    // the plain nested walk (kept for A/B measurements: -DB2_NL_FLAT=0)
    for (int cz = 0; cz < c_n[2]; cz++) {
        const int uz = c_lo[2] + cz;
        int iz = uz % g.nc[2]; if (iz < 0) iz += g.nc[2];
        const float sz = all[2] ? 0.f : (float)((uz - iz)/g.nc[2])*box[2];      // periodic shift of this cell layer
        const float gz = all[2] ? 0.f : fmaxf(0.f, fmaxf(lo[2] - (uz + 1)*cs[2], uz*cs[2] - up[2]));
        for (int cy = 0; cy < c_n[1]; cy++) {
            const int uy = c_lo[1] + cy;
            int iy = uy % g.nc[1]; if (iy < 0) iy += g.nc[1];
            const float sy = all[1] ? 0.f : (float)((uy - iy)/g.nc[1])*box[1];
            const float gy = all[1] ? 0.f : fmaxf(0.f, fmaxf(lo[1] - (uy + 1)*cs[1], uy*cs[1] - up[1]));
            const float rem2 = rmax2 - gz*gz - gy*gy;
            if (rem2 < 0.f) continue;
            // trim the x-range of this row of cells to the sphere cross-section, then walk it in runs
            // of cells that share one periodic shift: a run is a contiguous range of the cell-ordered
            // arrays
            int u_first = c_lo[0], u_last = c_lo[0] + c_n[0] - 1;
            if (!all[0]) {
                const float reach = sqrtf(rem2);
                u_first = max(u_first, (int)floorf((lo[0] - reach)/cs[0]));
                u_last = min(u_last, (int)floorf((up[0] + reach)/cs[0]));
            }
            const int row = (iz*g.nc[1] + iy)*g.nc[0];
            int u = u_first;
            while (u <= u_last) {
                int wrap = u >= 0 ? u/g.nc[0] : -((-u + g.nc[0] - 1)/g.nc[0]);
                const int ix0 = u - wrap*g.nc[0];
                const int run_end = min(u_last, u + (g.nc[0] - 1 - ix0));
                const int ix1 = ix0 + (run_end - u);
                const float sx = all[0] ? 0.f : (float)wrap*box[0];
                const int cb = cell_start[row + ix0], ce = cell_start[row + ix1 + 1];
                for (int base = cb; base < ce; base += 32) {
                    const int idx = base + lane;
                    bool pass = false;
                    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (idx < ce) {
                        const float4 cj = cgc[idx], hj = cgh[idx];
                        float dx = cj.x + sx - ci[0], dy = cj.y + sy - ci[1], dz = cj.z + sz - ci[2];
                        if (all[0]) dx -= box[0]*rintf(dx*ibox[0]);
                        if (all[1]) dy -= box[1]*rintf(dy*ibox[1]);
                        if (all[2]) dz -= box[2]*rintf(dz*ibox[2]);
                        const float gx = fmaxf(fabsf(dx) - (hi[0] + hj.x), 0.f);
                        const float gyy = fmaxf(fabsf(dy) - (hi[1] + hj.y), 0.f);
                        const float gzz = fmaxf(fabsf(dz) - (hi[2] + hj.z), 0.f);
                        pass = gx*gx + gyy*gyy + gzz*gzz < rmax2;
                        e = make_float4(dx, dy, dz, cj.w);
                        // domain decomposition: a candidate group that this rank does not own entirely is halo
                        if (halo_mark && pass) {
                            const int jg = __float_as_int(cj.w);
                            if (jg*B2_GROUP < own_lo || (jg + 1)*B2_GROUP > own_hi) halo_mark[jg] = 1;
                        }
                    }
                    const unsigned ballot = __ballot_sync(FULL, pass);
                    if (pass) queue[wib][qn + __popc(ballot & lt)] = e;
                    qn += __popc(ballot);
                    __syncwarp();
                    if (qn >= 4) {
                        const int full = qn & ~3;
                        sweep(full);
                        // move the (< 4) left-over entries to the head of the queue
                        const int rest = qn - full;
                        float4 keep = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (lane < rest) keep = queue[wib][full + lane];
                        __syncwarp();
                        if (lane < rest) queue[wib][lane] = keep;
                        qn = rest;
                        __syncwarp();
                    }
                }
                u = run_end + 1;
            }
        }
    }
#endif
    if (qn > 0) sweep(qn);
    // fat groups: tested directly, minimum image per atom (valid because list radius < L/2)
    const int nfat = flags[8];
    if (nfat > 0) {
        wide = true;
        qn = 0;
        for (int base = 0; base < nfat; base += 32) {
            const int idx = base + lane;
            bool pass = false;
            float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < nfat) {
                const int jg = fat_list[idx];
                const float4 cj = gcen[jg], hj = ghalf[jg];
                float dx = cj.x - ci[0], dy = cj.y - ci[1], dz = cj.z - ci[2];
                dx -= box[0]*rintf(dx*ibox[0]);
                dy -= box[1]*rintf(dy*ibox[1]);
                dz -= box[2]*rintf(dz*ibox[2]);
                const float gx = fmaxf(fabsf(dx) - (hi[0] + hj.x), 0.f);
                const float gyy = fmaxf(fabsf(dy) - (hi[1] + hj.y), 0.f);
                const float gzz = fmaxf(fabsf(dz) - (hi[2] + hj.z), 0.f);
                pass = gx*gx + gyy*gyy + gzz*gzz < rmax2;
                e = make_float4(dx, dy, dz, __int_as_float(jg));
                if (halo_mark && pass && (jg*B2_GROUP < own_lo || (jg + 1)*B2_GROUP > own_hi)) halo_mark[jg] = 1;
            }
            const unsigned ballot = __ballot_sync(FULL, pass);
            if (pass) queue[wib][qn + __popc(ballot & lt)] = e;
            qn += __popc(ballot);
            __syncwarp();
            if (qn >= 4) {
                const int full = qn & ~3;
                sweep(full);
                const int rest = qn - full;
                float4 keep = make_float4(0.f, 0.f, 0.f, 0.f);
                if (lane < rest) keep = queue[wib][full + lane];
                __syncwarp();
                if (lane < rest) queue[wib][lane] = keep;
                qn = rest;
                __syncwarp();
            }
        }
        if (qn > 0) sweep(qn);
    }
    // join the two parts: the shell entries move down behind the core entries.  Ascending addresses, a whole warp
    // step read before it is written: safe even when the two regions overlap (the destination is never above the
    // source).
    for (int k = 0; k < a.nlists; k++) {
        const int total = count[k] + scount[k];
        if (total <= a.cap[k] && scount[k] > 0 && count[k] < a.cap[k] - scount[k]) {
            int* base = a.entries[k] + (size_t)warp*a.cap[k];
            const int src0 = a.cap[k] - scount[k];
            __syncwarp();
            for (int m0 = 0; m0 < scount[k]; m0 += 32) {
                const int m = m0 + lane;
                int v = 0;
                if (m < scount[k]) v = base[src0 + m];
                __syncwarp();
                if (m < scount[k]) base[count[k] + m] = v;
                __syncwarp();
            }
        }
        count[k] = total;
    }
    __syncwarp();
    if (lane == 0) {
        for (int k = 0; k < a.nlists; k++) {
            a.counts[k][warp] = count[k] <= a.cap[k] ? count[k] : max(s_good[wib][k], 0);
            if (count[k] > a.cap[k]) flags[1] = 1;
            if (count[k] > flags[3]) atomicMax(&flags[3], count[k]);
        }
    }
}

// reference positions for the next skin test; the last block to finish closes the rebuild
__global__ void k_save_ref(int lo3, int n3, const double* __restrict__ x, double* __restrict__ xref, int* flags, int* hmax_bits) {
    if (!flags[0]) return;
    int i = lo3 + blockIdx.x*blockDim.x + threadIdx.x;
    if (i < n3) xref[i] = x[i];
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(&flags[4], 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        flags[4] = 0;
        flags[2] += 1;
        hmax_bits[0] = hmax_bits[1] = hmax_bits[2] = 0;
        __threadfence();
        flags[0] = 0;
    }
}

// ---------------------------------------------------------------------------------------------

static Grid make_grid(b2_context* ctx) {
    Grid g;
    double rmax = 0;
    for (int k = 0; k < ctx->nlists; k++) rmax = std::max(rmax, ctx->lists[k].cutoff + ctx->skin);
    g.rmax = rmax;
    for (int d = 0; d < 3; d++) {
        g.box[d] = ctx->box[d];
        g.inv[d] = 1.0/ctx->box[d];
        g.nc[d] = ctx->ncell[d];
        g.cs_inv[d] = ctx->ncell[d]/ctx->box[d];
    }
    return g;
}

int nl_setup(b2_context* ctx) {
    if (ctx->nlists == 0) return B2_OK;
    double rmax = 0;
    for (int k = 0; k < ctx->nlists; k++) rmax = std::max(rmax, ctx->lists[k].cutoff + ctx->skin);
    for (int d = 0; d < 3; d++) {
        for (int k = 0; k < ctx->nlists; k++)
            if (ctx->lists[k].cutoff > 0.5*ctx->box[d] + 1e-9)
                return b2_fail(ctx, B2_ERR_ARG, "cutoff %g nm exceeds half the box length %g nm (minimum image)",
                               ctx->lists[k].cutoff, ctx->box[d]);
        // cells are short along x (the contiguous direction of the cell-ordered arrays) and as long as the list
        // radius along y and z.  Measured (profiles/round2_build_variants.txt): halving the y/z size shrinks the
        // searched volume but quadruples the cell count -- the single-block scan gives back what the search gains
        static const double yz = getenv("B2_CELL_YZ") ? atof(getenv("B2_CELL_YZ")) : 1.0;
        double target = d == 0 ? std::max(0.5*rmax, 0.35) : std::max(yz*rmax, 0.7*yz);
        int nc = std::max(1, (int)floor(ctx->box[d]/target));
        nc = std::min(nc, 160);
        ctx->ncell[d] = nc;
        ctx->cellsize[d] = ctx->box[d]/nc;
    }
    int ncells = ctx->ncell[0]*ctx->ncell[1]*ctx->ncell[2];
    if (ncells != ctx->ncells || ctx->cell_count == nullptr) {
        cudaFree(ctx->cell_count); cudaFree(ctx->cell_start);
        ctx->ncells = ncells;
        B2_CUDA(cudaMalloc(&ctx->cell_count, sizeof(int)*(ncells + 1)));
        B2_CUDA(cudaMalloc(&ctx->cell_start, sizeof(int)*(ncells + 1)));
        B2_CUDA(cudaMemsetAsync(ctx->cell_count, 0, sizeof(int)*(ncells + 1), ctx->stream));
    }
    if (ctx->cell_groups == nullptr) {
        const int ng = ctx->ngroups;
        B2_CUDA(cudaMalloc(&ctx->cell_groups, sizeof(int)*ng));
        B2_CUDA(cudaMalloc(&ctx->gcell, sizeof(int)*ng));
        B2_CUDA(cudaMalloc(&ctx->fat_list, sizeof(int)*ng));
        B2_CUDA(cudaMalloc(&ctx->gcen, sizeof(float4)*ng));
        B2_CUDA(cudaMalloc(&ctx->ghalf, sizeof(float4)*ng));
        B2_CUDA(cudaMalloc(&ctx->cgc, sizeof(float4)*ng));
        B2_CUDA(cudaMalloc(&ctx->cgh, sizeof(float4)*ng));
        B2_CUDA(cudaMalloc(&ctx->prel, sizeof(float4)*(size_t)ng*B2_GROUP));
    }
    if (ctx->nl_flags == nullptr) {
        // [0] rebuild needed [1] overflow [2] rebuild counter [3] max count [4] ticket [5..7] max half extent
        // [8] number of fat groups  [9] constraint failure  [10] band overflow  [11] fat arrival counter
        B2_CUDA(cudaMalloc(&ctx->nl_flags, sizeof(int)*16));
        B2_CUDA(cudaMemsetAsync(ctx->nl_flags, 0, sizeof(int)*16, ctx->stream));
    }
    return B2_OK;
}

static int alloc_lists(b2_context* ctx, int k, int cap) {
    NList& L = ctx->lists[k];
    cudaFree(L.entries);
    L.entries = nullptr;
    L.cap = cap;
    B2_CUDA(cudaMalloc(&L.entries, sizeof(int)*(size_t)cap*ctx->ngroups));
    if (L.counts == nullptr) {
        B2_CUDA(cudaMalloc(&L.counts, sizeof(int)*ctx->ngroups));
        B2_CUDA(cudaMalloc(&L.gflags, ctx->ngroups));
    }
    return B2_OK;
}

// enqueue: skin test, then the (device-conditional) rebuild pipeline
int nl_prepare(b2_context* ctx, bool force) {
    if (ctx->nlists == 0) {
        return B2_OK;
    }
    // the lists were already checked (and rebuilt if necessary) for these very positions: RESPA evaluates
    // the near and the far force back to back at the end of a step
    if (!force && ctx->lists_built && ctx->nl_checked_version == ctx->pos_version) return B2_OK;
    ctx->nl_checked_version = ctx->pos_version;
    const int n = ctx->n, ng = ctx->ngroups, T = 256;
    static const double fat_factor = getenv("B2_FAT_FACTOR") ? atof(getenv("B2_FAT_FACTOR")) : 1.2;
    Grid g = make_grid(ctx);
    cudaStream_t s = ctx->stream;
    double limit = 0.5*ctx->skin;
    int* hmax = ctx->nl_flags + 5;
    // Peer-memory domain decomposition: every atom is tested by its owner, the verdicts are OR-ed over the
    // ranks inside the halo exchange, which also brings in the positions the pair kernels (or, when the
    // verdict is "rebuild", the list build) are about to read.  Otherwise the test runs over all atoms
    // (single GPU; NCCL mode, where the caller has all-gathered the positions before).
    // (forced builds follow b2_set_positions / a box change: positions are complete on every rank and nothing is
    // pulled, so the pass covers all atoms then, too)
    phase_mark(ctx, B2_PHASE_EXCHANGE);
    const bool owned_only = ctx->p2p && !force;
    const int t_lo = owned_only ? ctx->a_lo : 0, t_hi = owned_only ? ctx->a_hi : n;
    k_skin_check<<<std::max(1, (t_hi - t_lo + T - 1)/T), T, 0, s>>>(t_lo, t_hi, ctx->x, ctx->xref, limit*limit, ctx->nl_flags,
                                                                  (ctx->lists_built && !force) ? 1 : 0, ctx->xq,
                                                                  4294967296.0/ctx->box[0], 4294967296.0/ctx->box[1],
                                                                  4294967296.0/ctx->box[2]);
    B2_LAUNCH_CHECK();
    if (ctx->p2p) {
        if (!force) B2_TRY(dist_exchange_halo(ctx));     // forced builds follow b2_set_positions: x is complete everywhere
        else B2_CUDA(cudaMemsetAsync(ctx->halo_count, 0, sizeof(int), s));
    }
    phase_mark(ctx, B2_PHASE_REBUILD);
    k_group_geom<<<(8*ng + T - 1)/T, T, 0, s>>>(n, ng, ctx->x, g, ctx->prel, ctx->gcen, ctx->ghalf, ctx->gcell,
                                                 ctx->cell_count, hmax, (float)(fat_factor*ctx->cellsize[0]),
                                                 ctx->fat_list, ng, ctx->nl_flags);
    B2_LAUNCH_CHECK();
    k_cell_scan<<<1, 1024, 0, s>>>(ctx->ncells, ctx->cell_count, ctx->cell_start, ng, ctx->gcell, ctx->fat_list,
                                   ctx->nl_flags);
    B2_LAUNCH_CHECK();
    k_cell_fill<<<(ng + T - 1)/T, T, 0, s>>>(ng, ctx->gcell, ctx->cell_start, ctx->cell_count, ctx->cell_groups,
                                               ctx->nl_flags);
    B2_LAUNCH_CHECK();
    k_cell_sort_pack<<<(ctx->ncells + T - 1)/T, T, 0, s>>>(ctx->ncells, ctx->cell_start, ctx->cell_groups,
                                                            ctx->cell_count, ctx->gcen, ctx->ghalf, ctx->cgc, ctx->cgh,
                                                            ctx->nl_flags);
    B2_LAUNCH_CHECK();
    BuildArgs a;
    a.nlists = ctx->nlists;
    for (int k = 0; k < B2_MAX_LISTS; k++) {
        const NList& L = ctx->lists[k < ctx->nlists ? k : 0];
        double r = L.cutoff + ctx->skin;
        a.rlist[k] = r; a.rlist2[k] = r*r;
        // B2_SHELL_DELTA (nm): 0 puts every entry into the core part (the lists of round 1)
        static const double delta = getenv("B2_SHELL_DELTA") ? atof(getenv("B2_SHELL_DELTA")) : 0.02;
        a.rcore[k] = delta > 0 ? std::min(r, L.cutoff + delta) : r + 1.0;
        const float rl = (float)a.rlist[k] + NL_MARGIN, rc = (float)a.rcore[k] + NL_MARGIN;
        a.rl2f[k] = rl*rl; a.rc2f[k] = rc*rc;
        a.entries[k] = L.entries; a.counts[k] = L.counts; a.gflags[k] = L.gflags; a.cap[k] = L.cap;
    }
    k_build_lists<<<std::max(1, (ctx->g_hi - ctx->g_lo + NL_WARPS - 1)/NL_WARPS), 32*NL_WARPS, 0, s>>>(
        n, ctx->g_lo, ctx->g_hi, g, ctx->cell_start, ctx->cgc, ctx->cgh, ctx->gcen, ctx->ghalf, ctx->prel, hmax,
        ctx->fat_list, ctx->orig, ctx->exmask, ctx->excl_far ? ctx->excl_ptr : nullptr, ctx->excl_idx, ctx->excl_span, a, ctx->nl_flags,
        ctx->p2p ? ctx->halo_mark : nullptr, ctx->a_lo, ctx->a_hi);
    B2_LAUNCH_CHECK();
    B2_TRY(dist_halo_compact(ctx));
    k_save_ref<<<std::max(1, (3*(t_hi - t_lo) + T - 1)/T), T, 0, s>>>(3*t_lo, 3*t_hi, ctx->x, ctx->xref, ctx->nl_flags, hmax);
    B2_LAUNCH_CHECK();
    ctx->lists_built = true;
    phase_mark(ctx, B2_PHASE_OTHER);
    return B2_OK;
}

// first build after positions are (re)set: fit the per-group capacity, synchronising as needed
int nl_initial_build(b2_context* ctx) {
    if (ctx->nlists == 0) return B2_OK;
    B2_TRY(nl_setup(ctx));
    const double volume = ctx->box[0]*ctx->box[1]*ctx->box[2];
    const double rho = ctx->n/volume;
    for (int k = 0; k < ctx->nlists; k++) {
        double r = ctx->lists[k].cutoff + ctx->skin + 0.25;
        int guess = (int)(1.3*rho*4.18879*r*r*r) + 64;
        guess = std::min(guess, ctx->n + 32);
        guess = ((guess + 31)/32)*32;
        if (ctx->lists[k].entries == nullptr || (!ctx->lists_fitted && ctx->lists[k].cap < guess))
            B2_TRY(alloc_lists(ctx, k, guess));
    }
    // First build of a context: fit the capacity to 1.3 x the largest list (density fluctuations during a
    // run).  Later builds (re-ordering at setPositions) keep the buffers and only grow them, by 1.5 x, when
    // a list actually overflowed: no allocation, one synchronisation.
    const bool first = !ctx->lists_fitted;
    for (int attempt = 0; attempt < 6; attempt++) {
        int zero[8] = {0};
        int flags[8];
        B2_CUDA(cudaMemcpyAsync(flags, ctx->nl_flags, sizeof(flags), cudaMemcpyDeviceToHost, ctx->stream));
        B2_CUDA(cudaStreamSynchronize(ctx->stream));
        zero[2] = flags[2];
        B2_CUDA(cudaMemcpyAsync(ctx->nl_flags, zero, sizeof(zero), cudaMemcpyHostToDevice, ctx->stream));
        B2_TRY(nl_prepare(ctx, true));
        B2_CUDA(cudaMemcpyAsync(flags, ctx->nl_flags, sizeof(flags), cudaMemcpyDeviceToHost, ctx->stream));
        B2_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->counters[4] = flags[3];
        double rbig = 0;
        for (int k = 0; k < ctx->nlists; k++) rbig = std::max(rbig, ctx->lists[k].cutoff + ctx->skin);
        bool grown = false;
        if (first || flags[1]) {
            const double margin = first ? 1.3 : 1.5;
            for (int k = 0; k < ctx->nlists; k++) {
                // scaled per list by the ratio of list volumes
                const double ratio = pow((ctx->lists[k].cutoff + ctx->skin)/rbig, 3.0);
                int cap = (int)(margin*flags[3]*std::min(1.0, ratio*1.15)) + 64;
                cap = std::min(((cap + 31)/32)*32, ((ctx->n + 63)/32)*32);
                if (cap > ctx->lists[k].cap) {
                    B2_TRY(alloc_lists(ctx, k, cap));
                    grown = true;
                }
            }
        }
        if (!flags[1] && !grown) {
            ctx->counters[3] = ctx->lists[0].cap;
            ctx->lists_fitted = true;
            return B2_OK;
        }
    }
    return b2_fail(ctx, B2_ERR_OVERFLOW, "neighbour list capacity could not be fitted");
}
