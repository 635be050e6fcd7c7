// Internal context of the atomsmm_b200 engine (not part of the C ABI).
//
// Data layout in HBM (all per-atom arrays are in the engine's *spatial* order: whole molecules
// contiguous, molecules ordered along a Morton curve over cells; `orig`/`inv` map to and from
// the caller's numbering):
//
//   x, v        double [n][3]   master state, never wrapped (molecules stay whole)
//   xq          int4   [n]      x as 32-bit fixed-point fractions of the box, refreshed by the skin test
//   xref        double [n][3]   positions at the last neighbour-list build (skin test)
//   prel        float4 [n]      position relative to the centre of the atom's 8-atom group (list build)
//   par[s]      float4 [n]      per parameter set: {charge, sigma/2, sqrt(epsilon), 0}
//   massd       double [n]      mass;  invm: float [n] 1/mass (0 for massless)
//   fbuf[g]     float4 [n]      force of group g (slot 32 = all groups, "f")
//   perdof[k]   double [n][3]   user per-DOF integrator variables
//   lists       int [ngroups][cap]  per 8-atom i-group: (excl_mask<<24 | j) entries
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/atomsmm_b200.h"
#include "program.h"

#define B2_GROUP 8            // atoms per i-group (one warp = 8 i-atoms x 4 j-lanes)
#define B2_MAX_SETS 4
#define B2_MAX_LISTS 4
#define B2_FSLOTS 33          // force buffers: groups 0..31 and 32 = total
#define B2_MAX_PAIR_PARAMS 16
#ifndef B2_CHUNK
#define B2_CHUNK 96            // atoms per molecule chunk of the fused inner-loop kernel (measured: 96 beats 128, 64, 32)
#endif
#define B2_INNER_MAX_FORCES 16
#define B2_MAX_RANKS 16

struct PairParams {           // passed by value to kernels
    int family;
    float rc2;                // effective cutoff^2 (min of list cutoff and potential range)
    double p[B2_MAX_PAIR_PARAMS];
};

struct PairForce {
    int family, group, set, list;
    double cutoff;
    int nparams;
    double params[B2_MAX_PAIR_PARAMS];
    double econst;
    int bind[B2_MAX_PAIR_PARAMS];     // global-variable index a parameter is read from at run time, or -1
};

struct BondedForce {
    int family, group, nterms, arity, stride, periodic;
    int* atoms = nullptr;         // device, caller numbering
    std::vector<int> h_atoms;     // host copy (chunk construction)
    double* params = nullptr;     // device
    double gparams[8] = {0};
    int* code_e = nullptr; int ncode_e = 0;
    int* code_de = nullptr; int ncode_de = 0;
    double* consts = nullptr;
};

struct PmeForce {
    int group = 0, set = 0;
    int K[3] = {0, 0, 0};
    double alpha = 0, kc = 0, eself = 0;
    double* grid = nullptr;       // real-space charge / potential grid [nx][ny][nz]
    void* spectrum = nullptr;     // cufftDoubleComplex [nx][ny][nz/2+1]
    double* eterm = nullptr;      // influence function on the half spectrum
    int plan_fwd = -1, plan_inv = -1;
};

struct NList {
    double cutoff = 0;            // interaction cutoff
    int cap = 0;                  // entries per group
    int* entries = nullptr;       // [ngroups][cap]
    int* counts = nullptr;        // [ngroups]
    unsigned char* gflags = nullptr;  // [ngroups] bit0: group needs per-pair minimum image
};

// caller-order tables and scratch of the device-side ordering (order.cu)
struct OrderDevice {
    bool valid = false;
    int nmol = 0, nmol_used = 0, nsets = 0, nexcl = 0;
    int *mol_ptr = nullptr, *mol_atoms = nullptr;         // molecules in caller order (CSR)
    unsigned long long* keys = nullptr;                   // [2][nmol] Hilbert keys, unsorted / sorted
    int *ids = nullptr, *sizes = nullptr, *offsets = nullptr;
    double* mass_user = nullptr;
    unsigned long long* exmask_user = nullptr;
    double* sets_user[B2_MAX_SETS] = {nullptr, nullptr, nullptr, nullptr};
    int* excl_pairs = nullptr;
    int* span = nullptr;
    void* temp = nullptr;
    size_t temp_bytes = 0;
};

struct b2_context {
    int device = 0;
    cudaStream_t stream = 0;
    cudaStream_t own_stream = 0;
    cudaStream_t side_stream = 0;                 // second lane: two independent pair forces run concurrently
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    std::string err;
    long long counters[8] = {0};

    // ---- description -----------------------------------------------------------------------
    int n = 0, ngroups = 0;
    double box[3] = {0, 0, 0};
    int periodic = 1;
    double skin = 0.15;           // measured optimum on B200 for RESPA water (profiles/round2_skin_sweep.txt)
    std::vector<double> h_mass;
    std::vector<int> h_mol;
    std::vector<int> h_mol_ptr, h_mol_atoms;      // molecules in caller order (CSR), built once
    std::vector<std::vector<double>> h_sets;      // each 3*n: q, sigma, eps
    std::vector<int> h_excl;                      // pairs
    std::vector<PairForce> pair_forces;
    std::vector<BondedForce> bonded_forces;
    unsigned long long* bond_acc = nullptr;       // [n][3] fixed-point force accumulators of the explicit-list kernels
    std::vector<PmeForce> pme_forces;
    bool excl_far = false;                        // some exclusion spans > 31 in index
    int excl_span = 0;                            // largest distance in the engine's order between excluded atoms

    // ---- device state ----------------------------------------------------------------------
    bool have_order = false, have_positions = false, force_resort = false;
    int steps_since_order_check = 0;
    std::vector<int> h_orig;                      // sorted -> caller index
    OrderDevice order;                            // device-side ordering (order.cu)
    double *x = nullptr, *v = nullptr, *xref = nullptr, *xsort = nullptr;
    int4* xq = nullptr;                           // positions as 32-bit fixed-point fractions of the box (pair tiles)
    float4* par[B2_MAX_SETS] = {nullptr};
    double* pard[B2_MAX_SETS] = {nullptr};        // double [n][3]: charge, sigma, epsilon
    double* massd = nullptr;
    float* invm = nullptr;
    int *orig = nullptr, *inv = nullptr;
    unsigned long long* exmask = nullptr;         // per sorted atom: window mask over caller index
    int *excl_ptr = nullptr, *excl_idx = nullptr; // CSR by caller index (fallback for wide spans)
    float4* fbuf[B2_FSLOTS] = {nullptr};
    long long fvalid[B2_FSLOTS];                  // position version for which fbuf[g] is valid
    long long pos_version = 1;
    long long deriv_version = -1;                 // position version the parameter derivatives belong to
    // sum(m v.v) carried across velocity rescalings: v_version counts modifications of v; when mvv_version ==
    // v_version, globals[mvv_index] holds the sum of the CURRENT velocities (after v <- s v it is s^2 times
    // its old value), so a thermostat block that follows another one -- also across the step boundary --
    // needs neither a sweep over v nor, with several ranks, a reduction
    long long v_version = 1, mvv_version = -1;
    int mvv_index = -1, mvv_factor_hint = -1;     // hint: B2_OP_MVV_FACTOR seen right before the current op
    std::vector<double*> perdof;
    double* scratch3 = nullptr;                   // [n][3] staging for permuted copies
    std::vector<double*> carry_tmp;               // staging of v / per-DOF variables across a re-ordering
    double* order_tmp = nullptr;                  // positions in caller order for the in-run re-ordering

    // ---- neighbour lists -------------------------------------------------------------------
    int nlists = 0;
    NList lists[B2_MAX_LISTS];
    int ncell[3] = {0, 0, 0}, ncells = 0;
    double cellsize[3] = {0, 0, 0};
    int *cell_count = nullptr, *cell_start = nullptr;   // cells of GROUP centres
    int *cell_groups = nullptr, *gcell = nullptr;      // group ids in cell order; cell of every group (-1: fat)
    int* fat_list = nullptr;                           // groups too extended for the cells, in index order
    float4 *gcen = nullptr, *ghalf = nullptr;          // per group: box centre (wrapped), half extents
    float4 *cgc = nullptr, *cgh = nullptr;             // the same in cell order (w of cgc = group id)
    float4* prel = nullptr;                            // per atom: position relative to its group's centre
    int* nl_flags = nullptr;     // [0] rebuild needed, [1] overflow, [2] rebuild counter, [3] max count
    bool lists_built = false, lists_fitted = false;
    long long nl_checked_version = -1;            // position version of the last skin test

    // ---- domain decomposition (dist.cu) ------------------------------------------------------
    int rank = 0, nranks = 1;
    void* comm = nullptr;                         // ncclComm_t
    std::vector<int> range;                       // ownership boundaries in sorted order, size nranks+1
    int a_lo = 0, a_hi = 0, g_lo = 0, g_hi = 0;   // owned atoms [a_lo, a_hi), i-groups [g_lo, g_hi)
    long long x_synced = 0;                       // pos_version for which x is complete and consistent on all ranks
    // peer-memory halo exchange over NVLink (dist.cu): readers PULL the owners' positions
    bool p2p = false;                             // peers' x and signal blocks are mapped into this process
    double* peer_x[B2_MAX_RANKS] = {nullptr};     // peer r's position array (own entry: ctx->x)
    unsigned long long* peer_sig[B2_MAX_RANKS] = {nullptr};   // peer r's signal block (own entry: ctx->sig)
    unsigned long long* sig = nullptr;            // this rank's signal block: written by the peers
    unsigned long long* dd_state = nullptr;       // [0] exchange epoch [1] reduction epoch [2] time-out flag [3] ticket
    unsigned char* halo_mark = nullptr;           // [ngroups] j-group seen by the list build and not (entirely) owned
    int* halo_groups = nullptr;                   // compacted marks: the groups whose atoms are pulled
    int* halo_count = nullptr;                    // device counter of halo_groups
    bool acks_pending = false;                    // peers may still be reading x: wait before the next move

    // ---- fused RESPA inner loop (integrate.cu) ----------------------------------------------
    bool inner_built = false, inner_ok = false;
    int nchunks = 0;
    int *chunk_start = nullptr, *chunk_term_ptr = nullptr;
    int2* chunk_terms = nullptr;

    // ---- distance constraints (constraint.cu) -----------------------------------------------
    std::vector<int> h_con_atoms;                 // pairs, caller numbering
    std::vector<double> h_con_dist;
    double con_tol = 1e-5;
    bool con_built = false;
    int nclusters = 0;
    int* con_ptr = nullptr;                       // [nclusters+1] first constraint of every cluster
    int2* con_pairs = nullptr;                    // atoms in the engine's order
    double* con_d2 = nullptr;                     // squared target distances
    double* xcon = nullptr;                       // [n][3] last constrained configuration

    // ---- Monte Carlo barostat (barostat.cu) ------------------------------------------------------
    bool baro_on = false;
    double baro_pressure = 0, baro_kT = 0, baro_vscale = 0;     // kJ/mol/nm^3, kJ/mol, nm^3
    int baro_frequency = 0, baro_steps = 0, baro_attempts = 0, baro_accepted = 0;
    long long baro_total_attempts = 0, baro_total_accepted = 0;
    unsigned long long baro_seed = 0, baro_counter = 0;
    int nmol = 0;
    int* mol_start = nullptr;                     // [nmol+1] first atom of every molecule, engine order
    double* xbackup = nullptr;                    // positions before a trial move

    // ---- cutoff-band pairs settled in float64 (pair.cu) -----------------------------------
    int* band_pairs = nullptr;                    // [2 lanes][capacity] int2
    unsigned* band_count = nullptr;               // [2 lanes]
    unsigned band_capacity = 1u << 16;            // per lane; sized from n by b2_set_particles
    int* band_slot = nullptr;                     // [2 lanes][n] owner entry of every atom's accumulator (-1: none)
    long long* band_acc = nullptr;                // [2 lanes][capacity][3] fixed-point accumulators
    unsigned* band_ticket = nullptr;              // [2 lanes]

    // ---- profiling (eager mode only): CUDA-event pairs around every pair-force launch -------
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events;         // start, stop, start, stop ...
    std::vector<int> prof_tags;                   // pair force handle per pair of events
    std::vector<cudaEvent_t> phase_events;        // phase boundaries of the step (b2_get_phase_profile)
    std::vector<int> phase_tags;                  // tag of the phase that STARTS at the event

    // ---- energies --------------------------------------------------------------------------
    double* d_energy = nullptr;  // [32] energy + [32] virial + [2] dE/dlambda + pad
    double h_energy[32] = {0}, h_virial[32] = {0}, h_dlambda[2] = {0};

    // ---- integrator program ----------------------------------------------------------------
    std::vector<b2_op> ops;
    int *code = nullptr; int ncode = 0;
    std::vector<int> h_code;                      // host copy (kick term tables live in the code pool)
    // run-time compiled per-DOF / sum steps (jit.cu): module, one kernel per op (or null: interpreter), statistics
    void* jit_module = nullptr;
    std::vector<void*> jit_fn;
    bool jit_ready = false;
    bool jit_enabled = true;                      // b2_set_jit
    int jit_compiled = 0;
    long long jit_launches = 0;
    double *consts = nullptr; int nconsts = 0;
    double *globals = nullptr; int nglobals = 0;
    double* sum_partial = nullptr;                // per-block partial sums (sized by ensure_partials)
    int sum_partial_size = 0;
    unsigned* ticket = nullptr;                   // last-block-done counter of the velocity kernel
    unsigned long long* rng_state = nullptr;      // [0] seed, [1] draw counter
    bool program_loaded = false;
    bool uses_random = false;                     // the program draws random numbers (needs the step counter)
    bool prologue_valid = false;                  // the invariant coefficient prologue has been evaluated
    cudaGraphExec_t graph_exec = nullptr;
    bool graph_ready = false;
    int eager_steps = 0;
    long long graph_dpos = 0;
    unsigned long long graph_entry_mask = 0, graph_exit_mask = 0;
    bool graph_entry_synced = true, graph_exit_synced = true;   // replicated positions consistent on all ranks
    bool graph_entry_mvv = false, graph_exit_mvv = false;       // carried sum(m v.v) valid at entry / exit of the graph
    long long graph_dv = 0;
};

// position -> 32-bit fixed-point fraction of the box (scale = 2^32 / L); wraps like the periodic box
#ifdef __CUDACC__
__device__ __forceinline__ int b2_to_fixed(double x, double scale) { return (int)__double2ll_rn(x*scale); }
#endif

// ---- error helpers --------------------------------------------------------------------------
int b2_fail(b2_context* ctx, int code, const char* fmt, ...);

#define B2_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return b2_fail(ctx, B2_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, \
                           cudaGetErrorString(e_));                                            \
    } while (0)

#define B2_TRY(call)              \
    do {                          \
        int r_ = (call);          \
        if (r_ != B2_OK) return r_; \
    } while (0)

#define B2_LAUNCH_CHECK()                                                                      \
    do {                                                                                       \
        ctx->counters[0]++;                                                                    \
        cudaError_t e_ = cudaGetLastError();                                                   \
        if (e_ != cudaSuccess)                                                                 \
            return b2_fail(ctx, B2_ERR_CUDA, "kernel launch failed at %s:%d: %s", __FILE__, __LINE__, \
                           cudaGetErrorString(e_));                                            \
    } while (0)

template <typename T>
int b2_alloc(b2_context* ctx, T** ptr, size_t count);
int b2_free_all(b2_context* ctx);

// ---- cross-file entry points (host) ---------------------------------------------------------
int nl_setup(b2_context* ctx);                        // cells + list allocation for current box
int nl_prepare(b2_context* ctx, bool force);          // wrap + skin test + conditional rebuild
int nl_initial_build(b2_context* ctx);                // sized build with capacity fitting (syncs)
int pair_eval_forces(b2_context* ctx, const PairForce& pf, float4* out, bool accumulate, int lane = 0);
int pair_eval_energy(b2_context* ctx, const PairForce& pf, int group);
int pair_count_set(b2_context* ctx, const PairForce& pf, long long* count, unsigned long long* checksum,
                   int* pairs_dev, long long capacity);
int bonded_eval(b2_context* ctx, const BondedForce& bf, float4* out, bool want_force, bool want_energy);
int bonded_eval_forces(b2_context* ctx, uint32_t mask, float4* out);
int pme_setup(b2_context* ctx, PmeForce& pf);
int pme_eval(b2_context* ctx, PmeForce& pf, float4* out, double* acc, double* out64 = nullptr);
int pair_eval_forces64(b2_context* ctx, const PairForce& pf, double* out);
int bonded_eval_forces64(b2_context* ctx, uint32_t mask, double* out);
void pme_release(PmeForce& pf);
int dist_partition(b2_context* ctx);
int dist_sync_positions(b2_context* ctx);           // every rank ends up with the complete position array
int dist_exchange_halo(b2_context* ctx);            // peer-memory mode: post + pull (halo, or everything on a rebuild)
int dist_before_move(b2_context* ctx);              // peer-memory mode: wait until the peers have finished reading x
int dist_halo_compact(b2_context* ctx);             // after a list build: marks -> halo_groups
int dist_reduce_value(b2_context* ctx, double* value);   // sum of one device double over the ranks, on the stream
void dd_fill_peers(const b2_context* ctx, void* peers_out);   // DDPeers record for kernels outside dist.cu
int dist_gather3(b2_context* ctx, double* array);
int dist_gather_forces(b2_context* ctx, float4* array);
int dist_allreduce(b2_context* ctx, double* values, int count);
void dist_release(b2_context* ctx);
int forces_ensure(b2_context* ctx, uint32_t mask, int slot);
int inner_prepare(b2_context* ctx);
int order_refresh(b2_context* ctx);
// jit.cu: NVRTC compilation of the generic per-DOF / sum steps of the loaded program
int jit_prepare(b2_context* ctx);
void jit_release(b2_context* ctx);
int jit_launch(b2_context* ctx, int op_index, unsigned blocks, unsigned threads, void** args);
// order.cu: spatial order computed on the device from caller-order positions (orig, inv, static tables; h_orig on host)
int order_compute_device(b2_context* ctx, const double* x_user);
int order_refresh_param_set(b2_context* ctx, int k);
void order_release(b2_context* ctx);
int order_hilbert_bits(const double box[3]);
int con_prepare(b2_context* ctx);
int con_snapshot(b2_context* ctx);
int con_positions(b2_context* ctx);
int con_velocities(b2_context* ctx);
int program_run(b2_context* ctx, int nsteps);
int barostat_attempt(b2_context* ctx);
int box_update(b2_context* ctx, const double box[3]);
int program_release(b2_context* ctx);
// profiling mode: the work enqueued from here on belongs to phase `tag` (B2_PHASE_*)
enum { B2_PHASE_OTHER = 0, B2_PHASE_EXCHANGE = 1, B2_PHASE_REBUILD = 2, B2_PHASE_PAIR = 3, B2_PHASE_INTEGRATE = 4,
       B2_PHASE_REDUCE = 5, B2_PHASE_COUNT = 6 };
void phase_mark(b2_context* ctx, int tag);
int state_permute_to_sorted(b2_context* ctx, const double* user, double* sorted);
int state_permute_to_user(b2_context* ctx, const double* sorted, double* user);
