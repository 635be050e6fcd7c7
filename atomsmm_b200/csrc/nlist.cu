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

#include <algorithm>

#include "ctx.h"

#define FULL 0xffffffffu

struct BuildArgs {
    int nlists;
    double rlist2[B2_MAX_LISTS];
    double rlist[B2_MAX_LISTS];
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

// K8: skin test + fp32 wrapped copy used by the list build (the pair kernels read the float64
// master positions directly), one pass over x.
__global__ void k_wrap_check(int n, const double* __restrict__ x, const double* __restrict__ xref,
                             float4* __restrict__ pos4, Grid g, double limit2, int* flags, int have_ref) {
    int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    double px = x[3*i], py = x[3*i+1], pz = x[3*i+2];
    pos4[i] = make_float4((float)wrap1(px, g.box[0], g.inv[0]), (float)wrap1(py, g.box[1], g.inv[1]),
                          (float)wrap1(pz, g.box[2], g.inv[2]), 0.f);
    if (have_ref) {
        double dx = px - xref[3*i], dy = py - xref[3*i+1], dz = pz - xref[3*i+2];
        if (dx*dx + dy*dy + dz*dz > limit2) flags[0] = 1;
    } else {
        flags[0] = 1;
    }
}

__global__ void k_cell_count(int n, const double* __restrict__ x, Grid g, int* cell_of, int* cell_count,
                             const int* flags) {
    if (!flags[0]) return;
    int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        double w = wrap1(x[3*i+d], g.box[d], g.inv[d]);
        int k = (int)(w*g.cs_inv[d]);
        c[d] = min(max(k, 0), g.nc[d]-1);
    }
    int cell = (c[2]*g.nc[1] + c[1])*g.nc[0] + c[0];
    cell_of[i] = cell;
    atomicAdd(&cell_count[cell], 1);
}

// single-block exclusive scan (ncells <= a few 1e5); also resets the fill cursors
__global__ void k_cell_scan(int ncells, int* cell_count, int* cell_start, const int* flags) {
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
}

__global__ void k_cell_fill(int n, const int* cell_of, const int* cell_start, int* cell_count, int* cell_atoms,
                            const int* flags) {
    if (!flags[0]) return;
    int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = cell_of[i];
    int slot = atomicAdd(&cell_count[c], 1);
    cell_atoms[cell_start[c] + slot] = i;
}

// deterministic order inside each cell (atomics above fill in arbitrary order)
__global__ void k_cell_sort(int ncells, const int* cell_start, int* cell_atoms, const int* flags) {
    if (!flags[0]) return;
    int c = blockIdx.x*blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    int lo = cell_start[c], hi = cell_start[c+1];
    for (int a = lo + 1; a < hi; a++) {
        int key = cell_atoms[a];
        int b = a - 1;
        while (b >= lo && cell_atoms[b] > key) { cell_atoms[b+1] = cell_atoms[b]; b--; }
        cell_atoms[b+1] = key;
    }
}

// cell-ordered copies so that the list build streams candidates with coalesced loads
__global__ void k_cell_pack(int n, const int* __restrict__ cell_atoms, const float4* __restrict__ pos4,
                            const int* __restrict__ orig, float4* __restrict__ cpos, int* __restrict__ corig,
                            const int* flags) {
    if (!flags[0]) return;
    int k = blockIdx.x*blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int j = cell_atoms[k];
    float4 p = pos4[j];
    p.w = __int_as_float(j);
    cpos[k] = p;
    corig[k] = orig[j];
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

// K2: one warp per i-group.  Membership is tested in fp32 on wrapped coordinates with a safety
// margin (the list only has to be a superset of the pairs within cutoff+skin; the interacting set is
// decided exactly by the pair kernels).  Cells farther than the list radius from the group's
// bounding box are culled before their atoms are touched.
#define NL_MARGIN 3e-4f
__global__ void __launch_bounds__(128) k_build_lists(int n, int g_lo, int ngroups, const float4* __restrict__ pos4, Grid g,
                                                    const int* __restrict__ cell_start,
                                                    const float4* __restrict__ cpos,
                                                    const int* __restrict__ corig,
                                                    const int* __restrict__ orig,
                                                    const unsigned long long* __restrict__ exmask,
                                                    const int* __restrict__ excl_ptr,
                                                    const int* __restrict__ excl_idx, BuildArgs a, int* flags) {
    if (!flags[0]) return;
    const int warp = g_lo + ((blockIdx.x*blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    if (warp >= ngroups) return;
    __shared__ float sxi[4][B2_GROUP][3];
    __shared__ int soi[4][B2_GROUP];
    __shared__ unsigned long long smask[4][B2_GROUP];
    const int i0 = warp*B2_GROUP;
    const float bx = (float)g.box[0], by = (float)g.box[1], bz = (float)g.box[2];
    const float ibx = (float)g.inv[0], iby = (float)g.inv[1], ibz = (float)g.inv[2];
    const float4 r0 = pos4[i0];
    if (lane < B2_GROUP) {
        const int i = min(i0 + lane, n - 1);
        const float4 p = pos4[i];
        float dx = p.x - r0.x, dy = p.y - r0.y, dz = p.z - r0.z;      // unwrap relative to atom 0
        dx -= bx*rintf(dx*ibx); dy -= by*rintf(dy*iby); dz -= bz*rintf(dz*ibz);
        sxi[wib][lane][0] = r0.x + dx; sxi[wib][lane][1] = r0.y + dy; sxi[wib][lane][2] = r0.z + dz;
        soi[wib][lane] = (i0 + lane < n) ? orig[i] : -1;
        smask[wib][lane] = exmask[i];
    }
    __syncwarp();
    float lo[3], hi[3];
    for (int d = 0; d < 3; d++) {
        lo[d] = 1e30f; hi[d] = -1e30f;
        for (int k = 0; k < B2_GROUP; k++)
            if (soi[wib][k] >= 0) { lo[d] = fminf(lo[d], sxi[wib][k][d]); hi[d] = fmaxf(hi[d], sxi[wib][k][d]); }
    }
    unsigned padmask = 0;
    int omin = 0x7fffffff, omax = -1;
    for (int k = 0; k < B2_GROUP; k++) {
        const int oi = soi[wib][k];
        if (oi < 0) padmask |= 1u << k;
        else { omin = min(omin, oi); omax = max(omax, oi); }
    }
    const float rmax = (float)g.rmax + NL_MARGIN;
    const float cs[3] = {(float)(g.box[0]/g.nc[0]), (float)(g.box[1]/g.nc[1]), (float)(g.box[2]/g.nc[2])};
    const float box[3] = {bx, by, bz};
    int c_lo[3], c_n[3];
    bool all[3];
    for (int d = 0; d < 3; d++) {
        const int a0 = (int)floorf((lo[d] - rmax)/cs[d]);
        const int a1 = (int)floorf((hi[d] + rmax)/cs[d]);
        const int cover = a1 - a0 + 1;
        all[d] = cover >= g.nc[d];
        if (all[d]) { c_lo[d] = 0; c_n[d] = g.nc[d]; }
        else { c_lo[d] = a0; c_n[d] = cover; }
    }
    if (lane == 0) {
        for (int k = 0; k < a.nlists; k++) {
            unsigned char f = 0;
            for (int d = 0; d < 3; d++)
                if ((double)(hi[d] - lo[d]) + a.rlist[k] > 0.5*g.box[d] - 1e-4) f = 1;
            a.gflags[k][warp] = f;
        }
    }
    float rl2[B2_MAX_LISTS];
    for (int k = 0; k < B2_MAX_LISTS; k++) { const float r = (float)a.rlist[k] + NL_MARGIN; rl2[k] = r*r; }
    const float rmax2 = rmax*rmax;
    int count[B2_MAX_LISTS] = {0, 0, 0, 0};
    const unsigned lt = (1u << lane) - 1u;
    for (int cz = 0; cz < c_n[2]; cz++) {
        const int uz = c_lo[2] + cz;
        int iz = uz % g.nc[2]; if (iz < 0) iz += g.nc[2];
        const float sz = all[2] ? 0.f : (float)((uz - iz)/g.nc[2])*box[2];      // periodic shift of this cell layer
        const float gz = all[2] ? 0.f : fmaxf(0.f, fmaxf(lo[2] - (uz + 1)*cs[2], uz*cs[2] - hi[2]));
        for (int cy = 0; cy < c_n[1]; cy++) {
            const int uy = c_lo[1] + cy;
            int iy = uy % g.nc[1]; if (iy < 0) iy += g.nc[1];
            const float sy = all[1] ? 0.f : (float)((uy - iy)/g.nc[1])*box[1];
            const float gy = all[1] ? 0.f : fmaxf(0.f, fmaxf(lo[1] - (uy + 1)*cs[1], uy*cs[1] - hi[1]));
            const float rem2 = rmax2 - gz*gz - gy*gy;
            if (rem2 < 0.f) continue;
            // trim the x-range of this row of cells to the sphere cross-section, then walk it in runs
            // of cells that share one periodic shift: a run is a contiguous range of the cell-ordered
            // arrays
            int u_first = c_lo[0], u_last = c_lo[0] + c_n[0] - 1;
            if (!all[0]) {
                const float reach = sqrtf(rem2);
                u_first = max(u_first, (int)floorf((lo[0] - reach)/cs[0]));
                u_last = min(u_last, (int)floorf((hi[0] + reach)/cs[0]));
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
                    const bool have = idx < ce;
                    float d2min = 1e30f;
                    unsigned m = 0;
                    int j = 0;
                    if (have) {
                        const float4 pj = cpos[idx];
                        j = __float_as_int(pj.w);
                        const float xj = pj.x + sx, yj = pj.y + sy, zj = pj.z + sz;
                        const int oj = corig[idx];
                        // distance to the nearest of the 8 i-atoms (padding slots duplicate a real atom)
#pragma unroll
                        for (int k = 0; k < B2_GROUP; k++) {
                            float dx = xj - sxi[wib][k][0], dy = yj - sxi[wib][k][1], dz = zj - sxi[wib][k][2];
                            if (all[0]) dx -= bx*rintf(dx*ibx);
                            if (all[1]) dy -= by*rintf(dy*iby);
                            if (all[2]) dz -= bz*rintf(dz*ibz);
                            d2min = fminf(d2min, dx*dx + dy*dy + dz*dz);
                        }
                        // exclusion mask: only candidates whose caller index is close to the group's
                        // own index range (same molecule) can be excluded or be the atom itself
                        m = padmask;
                        if (excl_ptr != nullptr || (oj >= omin - 32 && oj <= omax + 32)) {
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
#pragma unroll
                    for (int k = 0; k < B2_MAX_LISTS; k++) {
                        if (k >= a.nlists) break;
                        const bool in = have && d2min < rl2[k];
                        const unsigned ballot = __ballot_sync(FULL, in);
                        if (in) {
                            const int pos = count[k] + __popc(ballot & lt);
                            if (pos < a.cap[k]) a.entries[k][(size_t)warp*a.cap[k] + pos] = (int)((m << 24) | (unsigned)j);
                        }
                        count[k] += __popc(ballot);
                    }
                }
                u = run_end + 1;
            }
        }
    }
    if (lane == 0) {
        for (int k = 0; k < a.nlists; k++) {
            a.counts[k][warp] = min(count[k], a.cap[k]);
            if (count[k] > a.cap[k]) flags[1] = 1;
            if (count[k] > flags[3]) atomicMax(&flags[3], count[k]);
        }
    }
}

__global__ void k_save_ref(int n3, const double* __restrict__ x, double* __restrict__ xref, int* flags) {
    if (!flags[0]) return;
    int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i < n3) xref[i] = x[i];
}

__global__ void k_finish_rebuild(int* flags) {
    if (flags[0]) { flags[0] = 0; flags[2] += 1; }
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
        if (ctx->lists[0].cutoff > 0.5*ctx->box[d] + 1e-9)
            return b2_fail(ctx, B2_ERR_ARG, "cutoff %g nm exceeds half the box length %g nm", rmax - ctx->skin,
                           ctx->box[d]);
        double target = std::max(0.5*rmax, 0.35);
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
    }
    if (ctx->cell_atoms == nullptr) {
        B2_CUDA(cudaMalloc(&ctx->cell_atoms, sizeof(int)*ctx->n));
        B2_CUDA(cudaMalloc(&ctx->cell_of, sizeof(int)*ctx->n));
        B2_CUDA(cudaMalloc(&ctx->cpos, sizeof(float4)*ctx->n));
        B2_CUDA(cudaMalloc(&ctx->corig, sizeof(int)*ctx->n));
    }
    if (ctx->nl_flags == nullptr) {
        B2_CUDA(cudaMalloc(&ctx->nl_flags, sizeof(int)*8));
        B2_CUDA(cudaMemsetAsync(ctx->nl_flags, 0, sizeof(int)*8, ctx->stream));
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

// enqueue: wrap + skin test, then the (device-conditional) rebuild pipeline
int nl_prepare(b2_context* ctx, bool force) {
    if (ctx->nlists == 0) {
        return B2_OK;
    }
    const int n = ctx->n, T = 256;
    Grid g = make_grid(ctx);
    cudaStream_t s = ctx->stream;
    double limit = 0.5*ctx->skin;
    k_wrap_check<<<(n + T - 1)/T, T, 0, s>>>(n, ctx->x, ctx->xref, ctx->pos4, g, limit*limit, ctx->nl_flags,
                                               (ctx->lists_built && !force) ? 1 : 0);
    B2_LAUNCH_CHECK();
    B2_CUDA(cudaMemsetAsync(ctx->cell_count, 0, sizeof(int)*(ctx->ncells + 1), s));
    k_cell_count<<<(n + T - 1)/T, T, 0, s>>>(n, ctx->x, g, ctx->cell_of, ctx->cell_count, ctx->nl_flags);
    B2_LAUNCH_CHECK();
    k_cell_scan<<<1, 1024, 0, s>>>(ctx->ncells, ctx->cell_count, ctx->cell_start, ctx->nl_flags);
    B2_LAUNCH_CHECK();
    k_cell_fill<<<(n + T - 1)/T, T, 0, s>>>(n, ctx->cell_of, ctx->cell_start, ctx->cell_count, ctx->cell_atoms,
                                              ctx->nl_flags);
    B2_LAUNCH_CHECK();
    k_cell_sort<<<(ctx->ncells + T - 1)/T, T, 0, s>>>(ctx->ncells, ctx->cell_start, ctx->cell_atoms, ctx->nl_flags);
    B2_LAUNCH_CHECK();
    k_cell_pack<<<(n + T - 1)/T, T, 0, s>>>(n, ctx->cell_atoms, ctx->pos4, ctx->orig, ctx->cpos, ctx->corig, ctx->nl_flags);
    B2_LAUNCH_CHECK();
    BuildArgs a;
    a.nlists = ctx->nlists;
    for (int k = 0; k < B2_MAX_LISTS; k++) {
        const NList& L = ctx->lists[k < ctx->nlists ? k : 0];
        double r = L.cutoff + ctx->skin;
        a.rlist[k] = r; a.rlist2[k] = r*r;
        a.entries[k] = L.entries; a.counts[k] = L.counts; a.gflags[k] = L.gflags; a.cap[k] = L.cap;
    }
    const int warps_per_block = 4;
    k_build_lists<<<std::max(1, (ctx->g_hi - ctx->g_lo + warps_per_block - 1)/warps_per_block), 32*warps_per_block, 0, s>>>(
        n, ctx->g_lo, ctx->g_hi, ctx->pos4, g, ctx->cell_start, ctx->cpos, ctx->corig, ctx->orig, ctx->exmask,
        ctx->excl_far ? ctx->excl_ptr : nullptr, ctx->excl_idx, a, ctx->nl_flags);
    B2_LAUNCH_CHECK();
    k_save_ref<<<(3*n + T - 1)/T, T, 0, s>>>(3*n, ctx->x, ctx->xref, ctx->nl_flags);
    B2_LAUNCH_CHECK();
    k_finish_rebuild<<<1, 1, 0, s>>>(ctx->nl_flags);
    B2_LAUNCH_CHECK();
    ctx->lists_built = true;
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
        if (ctx->lists[k].entries == nullptr || ctx->lists[k].cap < guess) B2_TRY(alloc_lists(ctx, k, guess));
    }
    for (int attempt = 0; attempt < 4; attempt++) {
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
        if (!flags[1]) {
            ctx->counters[3] = ctx->lists[0].cap;
            return B2_OK;
        }
        for (int k = 0; k < ctx->nlists; k++) {
            int cap = ((int)(flags[3]*1.25) + 63)/32*32;
            cap = std::min(cap, ((ctx->n + 63)/32)*32);
            B2_TRY(alloc_lists(ctx, k, std::max(cap, ctx->lists[k].cap)));
        }
    }
    return b2_fail(ctx, B2_ERR_OVERFLOW, "neighbour list capacity could not be fitted");
}
