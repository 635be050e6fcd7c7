// Monte Carlo barostat behind the UpdateContextState hook.
//
// The reference has no barostat of its own: its integrators emit `addUpdateContextState()` first in
// every step program (reference: integrators.py:115-122), which is where OpenMM lets a
// MonteCarloBarostat force of the System act.  This file is that force on the engine, with
// OpenMM's algorithm (MonteCarloBarostatImpl::updateContextState): every `frequency` steps propose
// V' = V + dV, dV uniform in +-volumeScale, scale the CENTRE of every molecule (molecules move
// rigidly) and the box isotropically, and accept with probability
//     min(1, exp(-(E' - E + P dV - N_mol kT ln(V'/V)) / kT)),
// adapting volumeScale every 10 attempts towards 25-75 % acceptance.
//
// Energies are the float64 report-cadence energies of b2_eval.  The box is a kernel argument of the
// pair / list / bonded kernels, i.e. baked into the captured step graph: an ACCEPTED move releases
// the graph (one eager step + one capture follow), re-fits cells and lists, and rescales the
// long-range corrections (~1/V).  Random numbers: SplitMix64 of (seed, attempt counter) -- a pure
// function that the oracle interpreter restates, so accept/reject sequences can be compared.
#include <math.h>

#include <algorithm>
#include <vector>

#include "ctx.h"

__global__ void k_scale_molecules(int nmol, const int* __restrict__ mol_start, int a_lo, int a_hi, double* __restrict__ x,
                                  double scale) {
    const int m = blockIdx.x*blockDim.x + threadIdx.x;
    if (m >= nmol) return;
    const int lo = mol_start[m], hi = mol_start[m+1];
    if (lo < a_lo || lo >= a_hi) return;          // molecules are owned whole
    double c[3] = {0, 0, 0};
    for (int i = lo; i < hi; i++)
        for (int k = 0; k < 3; k++) c[k] += x[3*i+k];
    const double inv = 1.0/(double)(hi - lo);
    for (int k = 0; k < 3; k++) c[k] = c[k]*inv*(scale - 1.0);
    for (int i = lo; i < hi; i++)
        for (int k = 0; k < 3; k++) x[3*i+k] += c[k];
}

static double splitmix_uniform(unsigned long long seed, unsigned long long counter) {
    unsigned long long z = seed + 0x9e3779b97f4a7c15ull*(counter + 1ull);
    z = (z ^ (z >> 30))*0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27))*0x94d049bb133111ebull;
    z ^= z >> 31;
    return (double)(z >> 11)*(1.0/9007199254740992.0);
}

extern "C" int b2_barostat_uniform(unsigned long long seed, unsigned long long counter, double* out) {
    if (!out) return B2_ERR_ARG;
    *out = splitmix_uniform(seed, counter);
    return B2_OK;
}

extern "C" int b2_set_barostat(b2_context* ctx, double pressure, double kT, int frequency, unsigned long long seed) {
    if (!ctx) return B2_ERR_ARG;
    if (frequency < 0 || !(kT > 0)) return b2_fail(ctx, B2_ERR_ARG, "bad barostat parameters");
    ctx->baro_on = frequency > 0;
    ctx->baro_pressure = pressure; ctx->baro_kT = kT; ctx->baro_frequency = frequency; ctx->baro_seed = seed;
    ctx->baro_counter = 0;
    ctx->baro_vscale = 0;
    ctx->baro_attempts = ctx->baro_accepted = 0;
    ctx->baro_total_attempts = ctx->baro_total_accepted = 0;
    ctx->baro_steps = 0;
    return B2_OK;
}

extern "C" int b2_get_box(b2_context* ctx, double out[3]) {
    if (!ctx || !out) return B2_ERR_ARG;
    for (int d = 0; d < 3; d++) out[d] = ctx->box[d];
    return B2_OK;
}

extern "C" int b2_get_barostat_stats(b2_context* ctx, long long out[2], double* volume_scale) {
    if (!ctx || !out) return B2_ERR_ARG;
    out[0] = ctx->baro_total_attempts; out[1] = ctx->baro_total_accepted;
    if (volume_scale) *volume_scale = ctx->baro_vscale;
    return B2_OK;
}

// New box for a context that has positions: cells, lists and everything derived from the volume follow.
// Positions must already be consistent with the new box on every rank.
int box_update(b2_context* ctx, const double box[3]) {
    const double v_old = ctx->box[0]*ctx->box[1]*ctx->box[2];
    for (int d = 0; d < 3; d++) ctx->box[d] = box[d];
    const double v_new = ctx->box[0]*ctx->box[1]*ctx->box[2];
    for (PairForce& pf : ctx->pair_forces) pf.econst *= v_old/v_new;       // long-range corrections ~ N^2/V
    for (PmeForce& pm : ctx->pme_forces) B2_TRY(pme_setup(ctx, pm));       // influence function depends on the box
    program_release(ctx);
    for (int g = 0; g < B2_FSLOTS; g++) ctx->fvalid[g] = -1;
    ctx->deriv_version = -1;
    ctx->lists_built = false;
    if (ctx->have_positions && ctx->nlists > 0) B2_TRY(nl_initial_build(ctx));
    return B2_OK;
}

static int total_energy(b2_context* ctx, double* e) {
    return b2_eval(ctx, 0xffffffffu, B2_EVAL_ENERGY, nullptr, e, nullptr);
}

// one volume move (called between MD steps, never inside a graph capture)
int barostat_attempt(b2_context* ctx) {
    const int n = ctx->n;
    if (ctx->mol_start == nullptr) return b2_fail(ctx, B2_ERR_STATE, "molecule table missing");
    if (ctx->xbackup == nullptr) B2_CUDA(cudaMalloc(&ctx->xbackup, sizeof(double)*3*n));
    const double volume = ctx->box[0]*ctx->box[1]*ctx->box[2];
    if (ctx->baro_vscale == 0) ctx->baro_vscale = 0.01*volume;
    double e0 = 0, e1 = 0;
    B2_TRY(total_energy(ctx, &e0));
    B2_TRY(dist_sync_positions(ctx));                // complete positions on every rank: the backup, and the list rebuild
    B2_CUDA(cudaMemcpyAsync(ctx->xbackup, ctx->x, sizeof(double)*3*n, cudaMemcpyDeviceToDevice, ctx->stream));
    const double dv = ctx->baro_vscale*2.0*(splitmix_uniform(ctx->baro_seed, ctx->baro_counter++) - 0.5);
    const double new_volume = volume + dv;
    const double scale = cbrt(new_volume/volume);
    const double old_box[3] = {ctx->box[0], ctx->box[1], ctx->box[2]};
    const double new_box[3] = {old_box[0]*scale, old_box[1]*scale, old_box[2]*scale};
    std::vector<double> econst;                      // restored bit for bit after a rejected move
    for (const PairForce& pf : ctx->pair_forces) econst.push_back(pf.econst);
    const int T = 128;
    B2_TRY(dist_before_move(ctx));
    // every rank scales ALL molecules (positions are complete): no exchange needed afterwards
    k_scale_molecules<<<(ctx->nmol + T - 1)/T, T, 0, ctx->stream>>>(ctx->nmol, ctx->mol_start, 0, n, ctx->x, scale);
    B2_LAUNCH_CHECK();
    ctx->pos_version++;
    ctx->x_synced = ctx->pos_version;
    B2_TRY(box_update(ctx, new_box));
    B2_TRY(total_energy(ctx, &e1));
    const double w = e1 - e0 + ctx->baro_pressure*dv - ctx->nmol*ctx->baro_kT*log(new_volume/volume);
    bool accept = true;
    if (w > 0) accept = splitmix_uniform(ctx->baro_seed, ctx->baro_counter) <= exp(-w/ctx->baro_kT);
    ctx->baro_counter++;                              // the acceptance deviate is consumed either way
    if (!accept) {
        B2_CUDA(cudaMemcpyAsync(ctx->x, ctx->xbackup, sizeof(double)*3*n, cudaMemcpyDeviceToDevice, ctx->stream));
        ctx->pos_version++;
        ctx->x_synced = ctx->pos_version;
        B2_TRY(box_update(ctx, old_box));
        for (size_t k = 0; k < econst.size(); k++) ctx->pair_forces[k].econst = econst[k];
    } else {
        ctx->baro_accepted++;
        ctx->baro_total_accepted++;
        // the spatial order was built for the old box; a uniform scaling keeps it equally good
        B2_CUDA(cudaMemcpyAsync(ctx->xsort, ctx->x, sizeof(double)*3*n, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    ctx->baro_attempts++;
    ctx->baro_total_attempts++;
    if (ctx->baro_attempts >= 10) {
        if (ctx->baro_accepted < 0.25*ctx->baro_attempts) {
            ctx->baro_vscale /= 1.1;
            ctx->baro_attempts = ctx->baro_accepted = 0;
        } else if (ctx->baro_accepted > 0.75*ctx->baro_attempts) {
            ctx->baro_vscale = std::min(ctx->baro_vscale*1.1, 0.3*ctx->box[0]*ctx->box[1]*ctx->box[2]);
            ctx->baro_attempts = ctx->baro_accepted = 0;
        }
    }
    return B2_OK;
}

// Context.setPeriodicBoxVectors on a live context (OpenMM semantics: the atoms are not moved)
extern "C" int b2_update_box(b2_context* ctx, const double box[3]) {
    if (!ctx || !box) return B2_ERR_ARG;
    for (int d = 0; d < 3; d++)
        if (!(box[d] > 0)) return b2_fail(ctx, B2_ERR_ARG, "box length must be positive");
    if (ctx->have_positions) B2_TRY(dist_sync_positions(ctx));
    return box_update(ctx, box);
}
