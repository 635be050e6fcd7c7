// Distance constraints: SHAKE for positions and RATTLE for velocities (kernel K9 of SURVEY 2.1).
//
// Replaces what OpenMM does for CustomIntegrator.addConstrainPositions / addConstrainVelocities
// (reference call sites: propagators.py:245-252 TranslationPropagator, :270-273 VelocityBoostPropagator,
// :1121-1133 VelocityVerletPropagator; constraints come from app.ForceField.createSystem with
// rigidWater / HBonds).  OpenMM picks SETTLE, SHAKE or CCMA per cluster; all of them solve
// the same equations -- displacements along the constraint directions of the last constrained
// configuration such that every distance is restored -- so one converged solver reproduces them
// to the tolerance.
//
// Constraints never couple different molecules, and the engine keeps whole molecules contiguous,
// so the constraint graph falls apart into small clusters (a rigid water = 3 constraints, a heavy
// atom with its hydrogens = 1-3).  One thread owns one cluster and sweeps its constraints
// Gauss-Seidel fashion until all are satisfied: no inter-thread communication, no atomics, and the
// result does not depend on scheduling (bit-reproducible).
#include <math.h>

#include <algorithm>
#include <numeric>

#include "ctx.h"

#define CON_MAX_SWEEPS 2000

// reference configuration for the next position constraint: the atoms of every cluster
__global__ void k_con_snapshot(int nclusters, const int* __restrict__ ptr, const int2* __restrict__ pairs,
                               const double* __restrict__ x, double* __restrict__ xcon) {
    const int c = blockIdx.x*blockDim.x + threadIdx.x;
    if (c >= nclusters) return;
    for (int k = ptr[c]; k < ptr[c+1]; k++) {
        const int2 p = pairs[k];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            xcon[3*p.x+d] = x[3*p.x+d];
            xcon[3*p.y+d] = x[3*p.y+d];
        }
    }
}

// SHAKE: x_i += g/m_i r_ref, x_j -= g/m_j r_ref with g = (d^2 - |s|^2) / (2 s.r_ref (1/m_i + 1/m_j)),
// r_ref = xcon_i - xcon_j (the last constrained configuration), s = x_i - x_j; afterwards the
// constrained configuration becomes the new reference.
__global__ void k_shake(int nclusters, const int* __restrict__ ptr, const int2* __restrict__ pairs,
                        const double* __restrict__ d2, const double* __restrict__ mass, double* __restrict__ x,
                        double* __restrict__ xcon, double tol, int* flags) {
    const int c = blockIdx.x*blockDim.x + threadIdx.x;
    if (c >= nclusters) return;
    const int k0 = ptr[c], k1 = ptr[c+1];
    bool done = false;
    for (int sweep = 0; sweep < CON_MAX_SWEEPS && !done; sweep++) {
        done = true;
        for (int k = k0; k < k1; k++) {
            const int2 p = pairs[k];
            double s[3], r[3];
#pragma unroll
            for (int d = 0; d < 3; d++) {
                s[d] = x[3*p.x+d] - x[3*p.y+d];
                r[d] = xcon[3*p.x+d] - xcon[3*p.y+d];
            }
            const double target = d2[k];
            const double diff = target - (s[0]*s[0] + s[1]*s[1] + s[2]*s[2]);
            if (fabs(diff) > 2.0*tol*target) {
                done = false;
                const double mi = mass[p.x], mj = mass[p.y];
                const double wi = mi > 0 ? 1.0/mi : 0.0, wj = mj > 0 ? 1.0/mj : 0.0;
                const double sr = s[0]*r[0] + s[1]*r[1] + s[2]*r[2];
                const double g = diff/(2.0*sr*(wi + wj));
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    x[3*p.x+d] += g*wi*r[d];
                    x[3*p.y+d] -= g*wj*r[d];
                }
            }
        }
    }
    if (!done) flags[9] = 1;
    for (int k = k0; k < k1; k++) {
        const int2 p = pairs[k];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            xcon[3*p.x+d] = x[3*p.x+d];
            xcon[3*p.y+d] = x[3*p.y+d];
        }
    }
}

// RATTLE (velocity stage): remove the relative velocity along every constraint,
// v_i -= g/m_i r, v_j += g/m_j r with g = r.(v_i - v_j) / (|r|^2 (1/m_i + 1/m_j)), r = x_i - x_j.
// Converged when every |d ln(r)/dt| = |r.v_rel|/d^2 is below tol per ps.
__global__ void k_rattle(int nclusters, const int* __restrict__ ptr, const int2* __restrict__ pairs,
                         const double* __restrict__ d2, const double* __restrict__ mass,
                         const double* __restrict__ x, double* __restrict__ v, double tol, int* flags) {
    const int c = blockIdx.x*blockDim.x + threadIdx.x;
    if (c >= nclusters) return;
    const int k0 = ptr[c], k1 = ptr[c+1];
    bool done = false;
    for (int sweep = 0; sweep < CON_MAX_SWEEPS && !done; sweep++) {
        done = true;
        for (int k = k0; k < k1; k++) {
            const int2 p = pairs[k];
            double r[3], u[3];
#pragma unroll
            for (int d = 0; d < 3; d++) {
                r[d] = x[3*p.x+d] - x[3*p.y+d];
                u[d] = v[3*p.x+d] - v[3*p.y+d];
            }
            const double rv = r[0]*u[0] + r[1]*u[1] + r[2]*u[2];
            const double r2 = r[0]*r[0] + r[1]*r[1] + r[2]*r[2];
            if (fabs(rv) > tol*d2[k]) {
                done = false;
                const double mi = mass[p.x], mj = mass[p.y];
                const double wi = mi > 0 ? 1.0/mi : 0.0, wj = mj > 0 ? 1.0/mj : 0.0;
                const double g = rv/(r2*(wi + wj));
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    v[3*p.x+d] -= g*wi*r[d];
                    v[3*p.y+d] += g*wj*r[d];
                }
            }
        }
    }
    if (!done) flags[9] = 1;
}

extern "C" int b2_set_constraints(b2_context* ctx, int count, const int* pairs, const double* distances, double tolerance) {
    if (!ctx || ctx->n == 0) return b2_fail(ctx, B2_ERR_STATE, "set particles first");
    if (count < 0 || (count > 0 && (!pairs || !distances)) || !(tolerance > 0))
        return b2_fail(ctx, B2_ERR_ARG, "bad constraint arguments");
    for (int k = 0; k < count; k++) {
        const int i = pairs[2*k], j = pairs[2*k+1];
        if (i < 0 || j < 0 || i >= ctx->n || j >= ctx->n || i == j || !(distances[k] > 0))
            return b2_fail(ctx, B2_ERR_ARG, "bad constraint %d", k);
        if (ctx->h_mol[i] != ctx->h_mol[j])
            return b2_fail(ctx, B2_ERR_ARG, "constraint %d couples atoms of different molecules", k);
    }
    ctx->h_con_atoms.assign(pairs, pairs + 2*(size_t)count);
    ctx->h_con_dist.assign(distances, distances + count);
    ctx->con_tol = tolerance;
    ctx->con_built = false;
    program_release(ctx);
    return B2_OK;
}

// clusters = connected components of the constraint graph, in the engine's order, owned clusters only
int con_prepare(b2_context* ctx) {
    if (ctx->con_built) return B2_OK;
    ctx->con_built = true;
    cudaFree(ctx->con_ptr); cudaFree(ctx->con_pairs); cudaFree(ctx->con_d2);
    ctx->con_ptr = nullptr; ctx->con_pairs = nullptr; ctx->con_d2 = nullptr;
    ctx->nclusters = 0;
    const int nc = (int)ctx->h_con_dist.size();
    if (nc == 0 || ctx->h_orig.empty()) return B2_OK;
    const int n = ctx->n;
    std::vector<int> inv(n), parent(n);
    for (int s = 0; s < n; s++) inv[ctx->h_orig[s]] = s;
    std::iota(parent.begin(), parent.end(), 0);
    auto find = [&](int a) {
        while (parent[a] != a) { parent[a] = parent[parent[a]]; a = parent[a]; }
        return a;
    };
    // union in the engine's numbering; the root of a cluster is its smallest index
    for (int k = 0; k < nc; k++) {
        int a = find(inv[ctx->h_con_atoms[2*k]]), b = find(inv[ctx->h_con_atoms[2*k+1]]);
        if (a != b) parent[std::max(a, b)] = std::min(a, b);
    }
    std::vector<int> order(nc);
    std::iota(order.begin(), order.end(), 0);
    std::vector<int> root(nc);
    for (int k = 0; k < nc; k++) root[k] = find(inv[ctx->h_con_atoms[2*k]]);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return root[a] < root[b]; });
    std::vector<int> ptr;
    std::vector<int2> recs;
    std::vector<double> d2;
    int last = -1;
    for (int k : order) {
        if (root[k] < ctx->a_lo || root[k] >= ctx->a_hi) continue;     // another rank's cluster
        if (root[k] != last) { ptr.push_back((int)recs.size()); last = root[k]; }
        recs.push_back(make_int2(inv[ctx->h_con_atoms[2*k]], inv[ctx->h_con_atoms[2*k+1]]));
        d2.push_back(ctx->h_con_dist[k]*ctx->h_con_dist[k]);
    }
    ptr.push_back((int)recs.size());
    ctx->nclusters = (int)ptr.size() - 1;
    if (ctx->nclusters == 0) return B2_OK;
    B2_CUDA(cudaMalloc(&ctx->con_ptr, sizeof(int)*ptr.size()));
    B2_CUDA(cudaMalloc(&ctx->con_pairs, sizeof(int2)*recs.size()));
    B2_CUDA(cudaMalloc(&ctx->con_d2, sizeof(double)*d2.size()));
    B2_CUDA(cudaMemcpy(ctx->con_ptr, ptr.data(), sizeof(int)*ptr.size(), cudaMemcpyHostToDevice));
    B2_CUDA(cudaMemcpy(ctx->con_pairs, recs.data(), sizeof(int2)*recs.size(), cudaMemcpyHostToDevice));
    B2_CUDA(cudaMemcpy(ctx->con_d2, d2.data(), sizeof(double)*d2.size(), cudaMemcpyHostToDevice));
    if (ctx->xcon == nullptr) B2_CUDA(cudaMalloc(&ctx->xcon, sizeof(double)*3*n));
    return B2_OK;
}

int con_snapshot(b2_context* ctx) {
    if (ctx->nclusters == 0) return B2_OK;
    const int T = 128;
    k_con_snapshot<<<(ctx->nclusters + T - 1)/T, T, 0, ctx->stream>>>(ctx->nclusters, ctx->con_ptr, ctx->con_pairs,
                                                                        ctx->x, ctx->xcon);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

int con_positions(b2_context* ctx) {
    // the position version advances on EVERY rank whenever the system has constraints, also on a rank that
    // owns no cluster: versions drive the exchange protocol and must stay identical across ranks
    if (ctx->nclusters == 0) {
        if (!ctx->h_con_dist.empty()) ctx->pos_version++;
        return B2_OK;
    }
    const int T = 128;
    k_shake<<<(ctx->nclusters + T - 1)/T, T, 0, ctx->stream>>>(ctx->nclusters, ctx->con_ptr, ctx->con_pairs, ctx->con_d2,
                                                                 ctx->massd, ctx->x, ctx->xcon, ctx->con_tol, ctx->nl_flags);
    B2_LAUNCH_CHECK();
    ctx->pos_version++;
    return B2_OK;
}

int con_velocities(b2_context* ctx) {
    if (ctx->nclusters == 0) return B2_OK;
    const int T = 128;
    k_rattle<<<(ctx->nclusters + T - 1)/T, T, 0, ctx->stream>>>(ctx->nclusters, ctx->con_ptr, ctx->con_pairs, ctx->con_d2,
                                                                  ctx->massd, ctx->x, ctx->v, ctx->con_tol, ctx->nl_flags);
    B2_LAUNCH_CHECK();
    return B2_OK;
}
