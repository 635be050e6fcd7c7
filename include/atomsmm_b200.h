/*
 * atomsmm_b200 -- C ABI of the B200-native engine for atomsmm's hot path.
 *
 * atomsmm (reference) has no FFI seam of its own: it hands *descriptions* to OpenMM's public
 * object model and OpenMM executes them.  This header is the seam one level up,
 * exactly where those descriptions are handed over.  Each entry point cites the reference
 * interface it replaces (file:line in /root/reference/src/atomsmm unless stated).
 *
 * Conventions
 *   - units: nm, ps, dalton, kJ/mol, elementary charge, K, rad
 *   - every call returns 0 on success or a negative B2_ERR_* code; b2_last_error() has the text
 *   - "host" pointers are plain host memory; "dev" pointers are CUDA device pointers (the Python
 *     side takes them from torch tensors with .data_ptr()); the library never returns memory
 *     it owns
 *   - atom indices are the caller's (original) numbering everywhere; the internal spatial
 *     ordering is never visible
 *   - one host thread per context; work is enqueued on the context's stream and calls are
 *     asynchronous unless they return host scalars
 */
#ifndef ATOMSMM_B200_H
#define ATOMSMM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B2_API __attribute__((visibility("default")))
#else
#define B2_API
#endif

typedef struct b2_context b2_context;

enum {
    B2_OK = 0,
    B2_ERR_CUDA = -1,        /* a CUDA runtime call failed */
    B2_ERR_ARG = -2,         /* invalid argument / call order */
    B2_ERR_UNSUPPORTED = -3, /* description outside the supported closed set */
    B2_ERR_OVERFLOW = -4,    /* neighbour-list capacity exceeded during a run */
    B2_ERR_STATE = -5        /* positions not set, no program loaded, ... */
};

/* ---- pair-potential families (closed set recognised from atomsmm's energy strings) -------- */
enum {
    /* NearNonbondedForce / NearForce._expressions (forces.py:539-567,655-670), also the
     * "-step(rc0-r)*near" discount of FarNonbondedForce (forces.py:713-722) and RESPASystem
     * group 31 (systems.py:73).  params: variant(0 none,1 shift,2 force-switch), rs0, rc0, Kc,
     * sign, use_coulomb */
    B2_PAIR_NEAR = 1,
    /* DampedSmoothedForce (forces.py:445-466).  params: alpha, rswitch, rcut, degree, Kc */
    B2_PAIR_DAMPED = 2,
    /* openmm.NonbondedForce direct space as configured by _AtomsMM_NonbondedForce
     * (forces.py:152-190): LJ (optional OpenMM switch) + Coulomb.  params: Kc, coulomb_kind
     * (0 none, 1 plain, 2 reaction field, 3 erfc/Ewald direct), krf, crf, alpha, use_switch,
     * rswitch, rcut */
    B2_PAIR_LJC = 3,
    /* ComputingSystem dispersion virial "24*epsilon*(2*(sigma/r)^12-(sigma/r)^6)"
     * (systems.py:894-898).  params: use_switch, rswitch, rcut */
    B2_PAIR_LJ_VIRIAL = 4,
    /* SoftcoreForce / SoftcoreLennardJonesForce (forces.py:727-793).  params: Kc, lambda_vdw,
     * lambda_coul, use_switch, rswitch, rcut, group_mode (1: the charges are +1/-1 labels of an
     * interaction group A x complement(A), systems.py:392; only unlike pairs interact) */
    B2_PAIR_SOFTCORE = 5
};

/* ---- explicit-list (bonded) families ------------------------------------------------------- */
enum {
    B2_BOND_HARMONIC = 1,    /* openmm.HarmonicBondForce; per term: r0, k */
    B2_ANGLE_HARMONIC = 2,   /* openmm.HarmonicAngleForce; per term: theta0, k */
    B2_TORSION_PERIODIC = 3, /* openmm.PeriodicTorsionForce; per term: n, phase, k */
    /* NonbondedExceptionsForce "4*epsilon*x*(x-1) + Kc*chargeprod/r" (forces.py:405-407) and
     * the exception part of openmm.NonbondedForce; per term: chargeprod, sigma, epsilon,
     * qi*qj (for the Ewald erf correction); globals: Kc, alpha (0 = no Ewald correction) */
    B2_BOND_LJC = 4,
    /* any other CustomBondForce / CustomAngleForce energy (NearExceptionForce forces.py:673-680,
     * redefine_bond/angle systems.py:168,228, ComputingSystem virials systems.py:914,934):
     * evaluated from bytecode of E and dE/d(r|theta) */
    B2_BOND_CUSTOM = 5,
    B2_ANGLE_CUSTOM = 6
};

/* flags for b2_eval */
enum { B2_EVAL_FORCES = 1, B2_EVAL_ENERGY = 2,
       B2_EVAL_DOUBLE = 4   /* with B2_EVAL_FORCES: every contribution evaluated and accumulated in float64
                               (OpenMM's Precision=double for State forces; report cadence only) */ };

/* ---- life cycle ---------------------------------------------------------------------------- */
/* replaces openmm.Context construction (computers.py:69, utils.py:155,226) */
B2_API int b2_create(int device, b2_context** out);
B2_API int b2_destroy(b2_context* ctx);
B2_API const char* b2_last_error(const b2_context* ctx);
B2_API const char* b2_version(void);
/* all work is enqueued on this stream (pass torch.cuda.current_stream().cuda_stream) */
B2_API int b2_set_stream(b2_context* ctx, void* cuda_stream);
B2_API int b2_synchronize(b2_context* ctx);

/* ---- system description (host pointers) ---------------------------------------------------- */
/* System.setDefaultPeriodicBoxVectors / Context.setPeriodicBoxVectors (computers.py:243);
 * orthorhombic only, like the reference's PressureComputer (computers.py:86-88) */
B2_API int b2_set_box(b2_context* ctx, const double box[3], int periodic);
/* System.addParticle + Context.getMolecules (computers.py:24): mass[n], molecule id [n] */
B2_API int b2_set_particles(b2_context* ctx, int n, const double* mass, const int* molecule);
/* CustomNonbondedForce.addParticle as used by importFrom (forces.py:299-301): per-particle
 * charge, sigma, epsilon.  Returns a parameter-set id shared by the pair forces built from the
 * same NonbondedForce. */
B2_API int b2_add_param_set(b2_context* ctx, const double* charge, const double* sigma, const double* epsilon,
                     int* set_id);
/* CustomNonbondedForce.addExclusion (forces.py:310-312): npairs (i,j) */
B2_API int b2_set_exclusions(b2_context* ctx, int npairs, const int* pairs);
/* openmm.CustomNonbondedForce / NonbondedForce added to a System (forces.py:35-36,108-110).
 * energy_constant: position-independent energy of the force (long-range correction). */
B2_API int b2_add_pair_force(b2_context* ctx, int family, int group, int param_set, double cutoff,
                      const double* params, int nparams, double energy_constant, int* handle);
/* updateParametersInContext / Context.setParameter for a pair force */
B2_API int b2_update_pair_force(b2_context* ctx, int handle, const double* params, int nparams,
                         double energy_constant);
/* Force.updateParametersInContext after setParticleParameters (openmm API used by the reference at
 * systems.py reset_coulomb_scaling_factor-style rescaling and by user scripts): replaces the
 * per-particle (charge, sigma, epsilon) table of ONE pair force; host arrays of n values each, caller's
 * atom order.  A parameter set shared with other forces is split so that they keep their values. */
B2_API int b2_update_pair_particles(b2_context* ctx, int handle, const double* charge, const double* sigma,
                             const double* epsilon);
/* A context parameter that the integrator itself moves (AFED extended variables,
 * integrators.py:670-744: `lambda <- lambda + 0.5*dt*v_lambda` is a ComputeGlobal on a context
 * parameter): from now on parameter `param` of pair force `handle` (1 = lambda_vdw, 2 = lambda_coul of
 * B2_PAIR_SOFTCORE) is read on the device from global variable `global_index` of the loaded program
 * at every evaluation.  Call after b2_load_program (which clears all bindings). */
B2_API int b2_bind_pair_parameter(b2_context* ctx, int handle, int param, int global_index);
/* HarmonicBondForce / HarmonicAngleForce / PeriodicTorsionForce / CustomBondForce.addBond
 * (forces.py:384-388).  atoms: arity*nterms indices; params: stride*nterms doubles;
 * gparams: family globals. */
B2_API int b2_add_bonded_force(b2_context* ctx, int family, int group, int nterms, const int* atoms,
                        const double* params, int stride, int periodic, const double* gparams,
                        int ngparams, int* handle);
/* CustomBondForce / CustomAngleForce with an arbitrary energy string, pre-compiled by the
 * Python front end to VM bytecode for E(s) and dE/ds (s = r or theta).  Variables: index 0 = s,
 * 1..stride = per-term parameters. */
B2_API int b2_add_custom_bonded_force(b2_context* ctx, int family, int group, int nterms, const int* atoms,
                               const double* params, int stride, int periodic,
                               const int* code_e, int ncode_e, const int* code_de, int ncode_de,
                               const double* consts, int nconsts, int* handle);
/* Reciprocal space of openmm.NonbondedForce with nonbondedMethod = PME (forces.py:185-187,
 * systems.py:74-75): smooth PME, order-5 B-splines, grid nx x ny x nz, Ewald parameter alpha.
 * self_energy = -Kc alpha/sqrt(pi) sum q^2 is added to the group energy. */
B2_API int b2_add_pme(b2_context* ctx, int group, int param_set, double alpha, int nx, int ny, int nz, double kc,
                      double self_energy, int* handle);
/* System.addConstraint (app.ForceField.createSystem with rigidWater / HBonds ...) and
 * Integrator.setConstraintTolerance: `count` pairs with their distances (nm); both atoms of a
 * constraint must belong to the same molecule.  Enforced by the CustomIntegrator steps
 * ConstrainPositions / ConstrainVelocities (propagators.py:245-252,270-273). */
B2_API int b2_set_constraints(b2_context* ctx, int count, const int* pairs, const double* distances,
                              double tolerance);
/* neighbour-list skin (nm); lists are rebuilt when an atom moved more than skin/2 */
B2_API int b2_set_skin(b2_context* ctx, double skin);

/* ---- state (device pointers, double [n][3], caller's atom order) --------------------------- */
/* Context.setPositions / setVelocities / getState (computers.py:74-84,244-245) */
B2_API int b2_set_positions(b2_context* ctx, const double* x_dev);
B2_API int b2_set_velocities(b2_context* ctx, const double* v_dev);
B2_API int b2_get_positions(b2_context* ctx, double* x_dev);
B2_API int b2_get_velocities(b2_context* ctx, double* v_dev);

/* ---- single-point evaluation --------------------------------------------------------------- */
/* Context.getState(getEnergy/getForces, groups=mask) (utils.py:164, computers.py:74-81).
 * forces_dev: double [n][3] or NULL; energy_host/virial_host: sums over the groups in the mask,
 * or NULL.  The virial is sum over pair-like terms of r.F = -r dE/dr (PressureComputer's
 * per-pair virial); per-group values via b2_get_group_energies. */
B2_API int b2_eval(b2_context* ctx, uint32_t group_mask, int flags, double* forces_dev,
            double* energy_host, double* virial_host);
/* kinetic energy 0.5 sum m v.v of the current velocities (State.getKineticEnergy, computers.py:86-88), reduced on the
 * device in a fixed order (and over the ranks of a decomposed system in rank order) */
B2_API int b2_kinetic_energy(b2_context* ctx, double* out_host);
B2_API int b2_get_group_energies(b2_context* ctx, double energy_host[32], double virial_host[32]);
/* dE/dlambda of the softcore pair forces evaluated by the last b2_eval with B2_EVAL_ENERGY:
 * out[0] = d/d lambda_vdw, out[1] = d/d lambda_coul (Context.getState(getParameterDerivatives)) */
B2_API int b2_get_parameter_derivatives(b2_context* ctx, double out_host[2]);
/* number of atom pairs (i<j) inside the cutoff of pair force `handle`, not excluded, and a
 * 64-bit order-independent checksum of the pair set (parity test: neighbour lists bit-exact).
 * pairs_dev (int2 [capacity], may be NULL) receives the pairs in caller numbering, unordered. */
B2_API int b2_pair_set(b2_context* ctx, int handle, long long* count_host, unsigned long long* checksum_host,
                int* pairs_dev, long long capacity);

/* ---- integrator program -------------------------------------------------------------------- */
/* One lowered CustomIntegrator program (integrators.py:113-145 emit it, integrators.py:163 runs
 * it).  ops: nops records of B2_OP_WORDS ints (layout in csrc/program.h); code/consts: shared
 * bytecode and constant pools; globals: initial values of all global variables (index 0 = dt). */
#define B2_OP_WORDS 8
B2_API int b2_load_program(b2_context* ctx, const int* ops, int nops, const int* code, int ncode,
                    const double* consts, int nconsts, const double* globals, int nglobals,
                    int nperdof, uint64_t seed);
B2_API int b2_set_globals(b2_context* ctx, int first, int count, const double* values_host);
B2_API int b2_get_globals(b2_context* ctx, int first, int count, double* values_host);
/* CustomIntegrator.set/getPerDofVariableByName (integrators.py:155-160); var >= 0 user slot */
B2_API int b2_set_perdof(b2_context* ctx, int var, const double* values_dev);
B2_API int b2_get_perdof(b2_context* ctx, int var, double* values_dev);
/* CustomIntegrator.step(n) (integrators.py:163).  Asynchronous. */
B2_API int b2_run(b2_context* ctx, int nsteps);
/* Generic per-DOF / sum steps (ComputePerDof / ComputeSum expressions that are not a recognised kick, drift or
 * rescaling: integrators.py:129,145) are compiled at run time -- bytecode -> CUDA C -> NVRTC for sm_100a -> one kernel
 * per step (csrc/jit.cu); without libnvrtc / libcuda, or with B2_NO_JIT=1, the device-side bytecode interpreter runs
 * them.  out_host[0] = steps of the loaded program that run as compiled kernels, out_host[1] = their launches so far. */
B2_API int b2_get_jit_stats(b2_context* ctx, long long out_host[2]);
/* enabled = 0: keep the bytecode interpreter for this context (A/B checks: both paths give identical bits) */
B2_API int b2_set_jit(b2_context* ctx, int enabled);
/* pure host check (no GPU): translate one per-DOF expression (VM bytecode, two ints per instruction) and compile it
 * with NVRTC for sm_100a; `out` receives the generated statements, or the compiler's log on failure */
B2_API int b2_jit_check(const int* code, int len, char* out, int out_size);
/* counters: [0] kernel launches since creation, [1] neighbour-list rebuilds, [2] pair-kernel
 * launches, [3] list capacity (entries per 8-atom group, largest list), [4] largest count seen */
B2_API int b2_get_counters(b2_context* ctx, long long out_host[8]);

/* neighbour-list diagnostics: [0] rebuilds so far, [1] largest list (entries per 8-atom group),
 * [2] groups too extended for the cell grid at the last rebuild ("fat": every list build tests them
 * directly), [3] number of groups */
B2_API int b2_get_list_stats(b2_context* ctx, long long out_host[4]);

/* profiling aid for bench.py: while on, steps run eagerly (no CUDA graph) and every pair-force
 * launch is bracketed by CUDA events on the context's stream.  b2_get_profile returns the summed
 * duration and launch count of pair force `handle`, and the current number of list entries it
 * reads per launch (for the algorithmic-bytes figure). */
B2_API int b2_set_profiling(b2_context* ctx, int on);
B2_API int b2_get_profile(b2_context* ctx, int handle, double* total_ms, long long* launches, long long* entries);
/* the same eager pass split into phases of the step, milliseconds accumulated since profiling was switched
 * on: [0] other, [1] skin test + halo exchange (includes waiting for the slowest peer), [2] list rebuild
 * pipeline, [3] pair-force kernels, [4] integrator kernels, [5] cross-rank reductions + scalar programs */
B2_API int b2_get_phase_profile(b2_context* ctx, double out_ms[6]);

/* ---- NPT: Monte Carlo barostat behind the UpdateContextState hook ---------------------------- *
 * The reference emits `addUpdateContextState()` as the first computation of every step program
 * (integrators.py:115-122) and has no barostat of its own; this is openmm.MonteCarloBarostat (the force a
 * user adds to the System) with OpenMM's algorithm: every `frequency` steps a trial volume change with
 * rigid molecule-centre scaling, Metropolis acceptance on E' - E + P dV - N_mol kT ln(V'/V), step size
 * adapted every 10 attempts.  pressure in kJ/mol/nm^3, kT in kJ/mol; frequency 0 switches it off.
 * Random numbers are SplitMix64(seed, counter) -- b2_barostat_uniform exposes the stream (pure host
 * function) so that a float64 oracle can replay the accept/reject sequence. */
B2_API int b2_set_barostat(b2_context* ctx, double pressure, double kT, int frequency, unsigned long long seed);
B2_API int b2_get_barostat_stats(b2_context* ctx, long long attempts_accepted_host[2], double* volume_scale);
B2_API int b2_barostat_uniform(unsigned long long seed, unsigned long long counter, double* out);
/* current periodic box (a barostat moves it) / Context.setPeriodicBoxVectors on a live context (OpenMM
 * semantics: atoms are not moved; cells, lists, PME influence function and long-range corrections follow) */
B2_API int b2_get_box(b2_context* ctx, double out[3]);
B2_API int b2_update_box(b2_context* ctx, const double box[3]);

/* ---- multi-GPU: spatial decomposition of ONE system over the GPUs of a node ---------------- *
 * The reference is single-process; these calls are the engine's own.  One process
 * (one context) per GPU.  Rank 0 obtains a 128-byte NCCL id, the host side broadcasts it (e.g.
 * torch.distributed), and every rank calls b2_comm_init BEFORE b2_set_positions.  Afterwards every
 * rank makes the same calls with the same (full-system) arguments; each rank integrates the atoms
 * of its ownership range and the getters return the full, gathered state on every rank. */
B2_API int b2_comm_unique_id(char out128[128]);
B2_API int b2_comm_init(b2_context* ctx, int nranks, int rank, const char id128[128]);
/* Pure host helper (no GPU needed): the ownership ranges the engine uses.  molecule_sorted[n] is the
 * molecule id of each atom in the engine's spatial order; out_ranges[nranks+1] receives boundaries
 * that fall on molecule boundaries nearest to k*n/nranks. */
B2_API int b2_partition_ranges(int n, const int* molecule_sorted, int nranks, int* out_ranges);
/* Pure host helper: index along the Hilbert curve (one-molecule resolution) by which the engine
 * orders molecules; consecutive indices are face-adjacent cells, so runs of consecutive atoms -- the
 * 8-atom i-groups, the molecule chunks, the ranks' ownership ranges -- are spatially compact. */
B2_API int b2_hilbert_index(const double position[3], const double box[3], unsigned long long* out_key);
/* diagnostic: the engine's current spatial order, orig_host[s] = caller index of the atom at position s (n ints).
 * The order is computed on the device (csrc/order.cu: Hilbert keys, radix sort, gathers); tests compare it with the
 * order b2_hilbert_index defines.  It never leaks through any other entry point. */
B2_API int b2_get_order(b2_context* ctx, int* orig_host);
/* Peer-memory halo exchange over NVLink / NVSwitch (the default when the GPUs of the node can map each
 * other's memory).  After b2_comm_init every rank exports a 256-byte record (cudaIpc handles of its
 * position array and of its signal block), the host side all-gathers the records (rank order) and every
 * rank imports the table.  From then on the ranks keep only their HALO current: before each pair-force
 * evaluation a rank tests the skin criterion on its own atoms, posts the verdict to its peers and pulls
 * the positions of the groups its last neighbour-list build saw (or of everybody, when any rank wants a
 * rebuild) directly out of the owners' memory; sums go through the peers' signal blocks.  No library
 * collective is left in the step graph.  If b2_comm_import fails (no peer access), the context stays in
 * the NCCL mode: full replicas, grouped ncclBroadcast all-gather, ncclAllReduce; b2_comm_import(ctx, -1, NULL)
 * puts a rank back into that mode (all ranks must run the same mode).
 * b2_comm_mode: peer_memory = 1 in the peer-memory mode; halo_atoms = atoms pulled per exchange. */
B2_API int b2_comm_export(b2_context* ctx, char out256[256]);
B2_API int b2_comm_import(b2_context* ctx, int nranks, const char* all256);
B2_API int b2_comm_mode(b2_context* ctx, int* peer_memory, long long* halo_atoms);
/* ownership range [lo, hi) of this rank in the engine's spatial order, and the number of
 * position exchanges performed so far */
B2_API int b2_comm_info(b2_context* ctx, int* rank, int* nranks, int* lo, int* hi, long long* exchanges);
/* device-side clock of the peer-memory exchange, accumulated since the context was created (measurement aid,
 * no reference counterpart): out[0] = seconds spent waiting for the peers' "positions final" posts,
 * out[1] = seconds copying halo positions out of the owners' memory, out[2] = number of exchanges,
 * out[3] = seconds owners waited for the readers' acknowledgements before moving their atoms */
B2_API int b2_comm_timing(b2_context* ctx, double out[4]);

#ifdef __cplusplus
}
#endif
#endif
