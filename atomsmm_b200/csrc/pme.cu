// Smooth particle-mesh Ewald reciprocal space (kernel K10 of SURVEY 2.1), float64 throughout.
//
// Replaces the reciprocal-space part of openmm.NonbondedForce with nonbondedMethod = PME as
// atomsmm uses it (forces.py:185-187 copies the PME parameters; systems.py:74-75 puts direct and
// reciprocal space in group 2; systems.py:904-912 uses the charges-only force as the Coulomb
// virial).  Algorithm (Essmann et al. 1995, OpenMM Reference conventions): order-5
// cardinal B-splines, charges spread to grid points (floor(u)+k) mod K, E = 1/2 sum_m eterm(m)
// |Q^(m)|^2 with eterm = Kc exp(-pi^2 m^2/alpha^2)/(pi V m^2 B1 B2 B3), forces by interpolating the
// gradient of the convolved grid.  The 3-D transforms are cuFFT D2Z / Z2D; spreading, convolution
// (+ energy) and gathering are hand-written kernels.  The self term -Kc alpha/sqrt(pi) sum q^2 is
// added as the force's constant energy by the caller.
#include <cufft.h>
#include <math.h>

#include <vector>

#include "ctx.h"

#define PME_ORDER 5
#define FULL 0xffffffffu

struct PmeGrid {
    int K[3];
    double box[3];
};

// order-5 B-spline weights and derivatives at fractional offset w (same recursion as the oracle)
__device__ __forceinline__ void bspline5(double w, double (&th)[PME_ORDER], double (&dth)[PME_ORDER]) {
    th[PME_ORDER-1] = 0.0;
    th[1] = w;
    th[0] = 1.0 - w;
#pragma unroll
    for (int j = 3; j < PME_ORDER; j++) {
        const double div = 1.0/(j - 1);
        th[j-1] = div*w*th[j-2];
#pragma unroll
        for (int k = 1; k < j - 1; k++) th[j-k-1] = div*((w + k)*th[j-k-2] + (j - k - w)*th[j-k-1]);
        th[0] = div*(1.0 - w)*th[0];
    }
    dth[0] = -th[0];
#pragma unroll
    for (int j = 1; j < PME_ORDER; j++) dth[j] = th[j-1] - th[j];
    const double div = 1.0/(PME_ORDER - 1);
    th[PME_ORDER-1] = div*w*th[PME_ORDER-2];
#pragma unroll
    for (int k = 1; k < PME_ORDER - 1; k++)
        th[PME_ORDER-k-1] = div*((w + k)*th[PME_ORDER-k-2] + (PME_ORDER - k - w)*th[PME_ORDER-k-1]);
    th[0] = div*(1.0 - w)*th[0];
}

__device__ __forceinline__ void locate(const double* __restrict__ x, int i, const PmeGrid& g, int (&base)[3],
                                       double (&w)[3]) {
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const double s = x[3*i+d]/g.box[d];
        double u = (s - floor(s))*g.K[d];
        int b = (int)floor(u);
        if (b >= g.K[d]) { b = 0; u -= g.K[d]; }
        base[d] = b;
        w[d] = u - b;
    }
}

__global__ void k_pme_spread(int n, const double* __restrict__ x, const double* __restrict__ pard, PmeGrid g,
                             double* __restrict__ Q) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double q = pard[3*i];
    if (q == 0.0) return;
    int base[3];
    double w[3], tx[PME_ORDER], ty[PME_ORDER], tz[PME_ORDER], d[PME_ORDER];
    locate(x, i, g, base, w);
    bspline5(w[0], tx, d);
    bspline5(w[1], ty, d);
    bspline5(w[2], tz, d);
    for (int a = 0; a < PME_ORDER; a++) {
        const int ia = (base[0] + a) % g.K[0];
        for (int b = 0; b < PME_ORDER; b++) {
            const int ib = (base[1] + b) % g.K[1];
            const double qab = q*tx[a]*ty[b];
            for (int c = 0; c < PME_ORDER; c++) {
                const int ic = (base[2] + c) % g.K[2];
                atomicAdd(&Q[((size_t)ia*g.K[1] + ib)*g.K[2] + ic], qab*tz[c]);
            }
        }
    }
}

// multiply the half-spectrum by eterm and accumulate E = 1/2 sum eterm |Q^|^2 (Hermitian weights)
__global__ void k_pme_convolve(size_t total, int nzh, int nz, cufftDoubleComplex* __restrict__ F,
                               const double* __restrict__ eterm, double* __restrict__ acc, int want_energy) {
    double e = 0;
    for (size_t k = blockIdx.x*(size_t)blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x*blockDim.x) {
        const double t = eterm[k];
        cufftDoubleComplex f = F[k];
        if (want_energy) {
            const int kz = (int)(k % nzh);
            const double weight = (kz == 0 || (nz % 2 == 0 && kz == nz/2)) ? 1.0 : 2.0;
            e += 0.5*weight*t*(f.x*f.x + f.y*f.y);
        }
        f.x *= t; f.y *= t;
        F[k] = f;
    }
    if (want_energy) {
        for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(FULL, e, o);
        if ((threadIdx.x & 31) == 0 && e != 0.0) atomicAdd(acc, e);
    }
}

__global__ void k_pme_gather(int n, const double* __restrict__ x, const double* __restrict__ pard, PmeGrid g,
                             const double* __restrict__ C, float4* __restrict__ out, double* __restrict__ out64) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double q = pard[3*i];
    if (q == 0.0) return;
    int base[3];
    double w[3], tx[PME_ORDER], ty[PME_ORDER], tz[PME_ORDER], dx[PME_ORDER], dy[PME_ORDER], dz[PME_ORDER];
    locate(x, i, g, base, w);
    bspline5(w[0], tx, dx);
    bspline5(w[1], ty, dy);
    bspline5(w[2], tz, dz);
    double fx = 0, fy = 0, fz = 0;
    for (int a = 0; a < PME_ORDER; a++) {
        const int ia = (base[0] + a) % g.K[0];
        for (int b = 0; b < PME_ORDER; b++) {
            const int ib = (base[1] + b) % g.K[1];
            for (int c = 0; c < PME_ORDER; c++) {
                const int ic = (base[2] + c) % g.K[2];
                const double v = C[((size_t)ia*g.K[1] + ib)*g.K[2] + ic];
                fx += v*dx[a]*ty[b]*tz[c];
                fy += v*tx[a]*dy[b]*tz[c];
                fz += v*tx[a]*ty[b]*dz[c];
            }
        }
    }
    if (out64) {
        out64[3*i] -= q*fx*g.K[0]/g.box[0];
        out64[3*i+1] -= q*fy*g.K[1]/g.box[1];
        out64[3*i+2] -= q*fz*g.K[2]/g.box[2];
        return;
    }
    float4 f = out[i];
    f.x -= (float)(q*fx*g.K[0]/g.box[0]);
    f.y -= (float)(q*fy*g.K[1]/g.box[1]);
    f.z -= (float)(q*fz*g.K[2]/g.box[2]);
    out[i] = f;
}

// ---------------------------------------------------------------------------------------------
static void bspline_moduli(int K, std::vector<double>& mod) {
    double data[PME_ORDER] = {0}, w = 0.0;
    // weights at w = 0 (host copy of the recursion)
    data[PME_ORDER-1] = 0; data[1] = w; data[0] = 1 - w;
    for (int j = 3; j < PME_ORDER; j++) {
        double div = 1.0/(j - 1);
        data[j-1] = div*w*data[j-2];
        for (int k = 1; k < j - 1; k++) data[j-k-1] = div*((w + k)*data[j-k-2] + (j - k - w)*data[j-k-1]);
        data[0] = div*(1 - w)*data[0];
    }
    double div = 1.0/(PME_ORDER - 1);
    data[PME_ORDER-1] = div*w*data[PME_ORDER-2];
    for (int k = 1; k < PME_ORDER - 1; k++)
        data[PME_ORDER-k-1] = div*((w + k)*data[PME_ORDER-k-2] + (PME_ORDER - k - w)*data[PME_ORDER-k-1]);
    data[0] = div*(1 - w)*data[0];
    mod.assign(K, 0.0);
    for (int i = 0; i < K; i++) {
        double sc = 0, ss = 0;
        for (int j = 0; j < PME_ORDER && j < K; j++) {
            const double arg = 2.0*M_PI*i*j/K;
            sc += data[j]*cos(arg);
            ss += data[j]*sin(arg);
        }
        mod[i] = sc*sc + ss*ss;
    }
    for (int i = 0; i < K; i++)
        if (mod[i] < 1e-7) mod[i] = 0.5*(mod[(i - 1 + K) % K] + mod[(i + 1) % K]);
}

int pme_setup(b2_context* ctx, PmeForce& pf) {
    const int nx = pf.K[0], ny = pf.K[1], nz = pf.K[2], nzh = nz/2 + 1;
    const size_t real = (size_t)nx*ny*nz, half = (size_t)nx*ny*nzh;
    B2_CUDA(cudaMalloc(&pf.grid, sizeof(double)*real));
    B2_CUDA(cudaMalloc(&pf.spectrum, sizeof(cufftDoubleComplex)*half));
    B2_CUDA(cudaMalloc(&pf.eterm, sizeof(double)*half));
    std::vector<double> mod[3];
    for (int d = 0; d < 3; d++) bspline_moduli(pf.K[d], mod[d]);
    std::vector<double> e(half);
    const double V = ctx->box[0]*ctx->box[1]*ctx->box[2];
    for (int kx = 0; kx < nx; kx++) {
        const double mx = (kx < (nx + 1)/2 ? kx : kx - nx)/ctx->box[0];
        for (int ky = 0; ky < ny; ky++) {
            const double my = (ky < (ny + 1)/2 ? ky : ky - ny)/ctx->box[1];
            for (int kz = 0; kz < nzh; kz++) {
                const double mz = (kz < (nz + 1)/2 ? kz : kz - nz)/ctx->box[2];
                const double m2 = mx*mx + my*my + mz*mz;
                const size_t k = ((size_t)kx*ny + ky)*nzh + kz;
                if (kx == 0 && ky == 0 && kz == 0) { e[k] = 0.0; continue; }
                const double denom = M_PI*V*mod[0][kx]*mod[1][ky]*mod[2][kz]*m2;
                e[k] = pf.kc*exp(-M_PI*M_PI*m2/(pf.alpha*pf.alpha))/denom;
            }
        }
    }
    B2_CUDA(cudaMemcpy(pf.eterm, e.data(), sizeof(double)*half, cudaMemcpyHostToDevice));
    cufftHandle fwd, inv;
    if (cufftPlan3d(&fwd, nx, ny, nz, CUFFT_D2Z) != CUFFT_SUCCESS || cufftPlan3d(&inv, nx, ny, nz, CUFFT_Z2D) != CUFFT_SUCCESS)
        return b2_fail(ctx, B2_ERR_CUDA, "cuFFT plan creation failed for a %dx%dx%d grid", nx, ny, nz);
    pf.plan_fwd = (int)fwd;
    pf.plan_inv = (int)inv;
    return B2_OK;
}

// adds reciprocal-space forces into `out` (if not null) and the reciprocal energy into *acc
int pme_eval(b2_context* ctx, PmeForce& pf, float4* out, double* acc, double* out64) {
    const int n = ctx->n, T = 128;
    PmeGrid g;
    for (int d = 0; d < 3; d++) { g.K[d] = pf.K[d]; g.box[d] = ctx->box[d]; }
    const int nx = pf.K[0], ny = pf.K[1], nz = pf.K[2], nzh = nz/2 + 1;
    const size_t real = (size_t)nx*ny*nz, half = (size_t)nx*ny*nzh;
    cufftSetStream((cufftHandle)pf.plan_fwd, ctx->stream);
    cufftSetStream((cufftHandle)pf.plan_inv, ctx->stream);
    B2_CUDA(cudaMemsetAsync(pf.grid, 0, sizeof(double)*real, ctx->stream));
    k_pme_spread<<<(n + T - 1)/T, T, 0, ctx->stream>>>(n, ctx->x, ctx->pard[pf.set], g, pf.grid);
    B2_LAUNCH_CHECK();
    if (cufftExecD2Z((cufftHandle)pf.plan_fwd, pf.grid, (cufftDoubleComplex*)pf.spectrum) != CUFFT_SUCCESS)
        return b2_fail(ctx, B2_ERR_CUDA, "cufftExecD2Z failed");
    ctx->counters[0]++;
    const int blocks = (int)std::min<size_t>((half + 255)/256, 1184);
    k_pme_convolve<<<blocks, 256, 0, ctx->stream>>>(half, nzh, nz, (cufftDoubleComplex*)pf.spectrum, pf.eterm, acc,
                                                     acc != nullptr);
    B2_LAUNCH_CHECK();
    if (out || out64) {
        if (cufftExecZ2D((cufftHandle)pf.plan_inv, (cufftDoubleComplex*)pf.spectrum, pf.grid) != CUFFT_SUCCESS)
            return b2_fail(ctx, B2_ERR_CUDA, "cufftExecZ2D failed");
        ctx->counters[0]++;
        k_pme_gather<<<(n + T - 1)/T, T, 0, ctx->stream>>>(n, ctx->x, ctx->pard[pf.set], g, pf.grid, out, out64);
        B2_LAUNCH_CHECK();
    }
    return B2_OK;
}

void pme_release(PmeForce& pf) {
    if (pf.plan_fwd >= 0) cufftDestroy((cufftHandle)pf.plan_fwd);
    if (pf.plan_inv >= 0) cufftDestroy((cufftHandle)pf.plan_inv);
    cudaFree(pf.grid); cudaFree(pf.spectrum); cudaFree(pf.eterm);
    pf.grid = nullptr; pf.spectrum = nullptr; pf.eterm = nullptr;
    pf.plan_fwd = pf.plan_inv = -1;
}
