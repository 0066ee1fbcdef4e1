/*
 * chemlab_b200.h -- C-ABI of the B200-native engine behind chemlab's reactive-MD hot path.
 *
 * The reference (cgchemlab/chemlab) has no C interface: its hot path lives in the external,
 * un-vendored `espressopp` Boost.Python package and is reached from Python constructors
 * (SURVEY.md section 8b).  Every entry point below therefore replaces one espressopp call
 * site inside the reference; the call site is cited as `file:line` relative to the reference
 * root.  The Python package `chemlab_b200.espressopp` is the only intended caller; it mirrors
 * the espressopp names one-to-one on top of this ABI.
 *
 * Conventions
 *   - every function returns CLB_OK (0) or a negative error code; clb_last_error() gives text.
 *   - plain pointers and sizes only; all pointers are HOST pointers unless suffixed _dev.
 *   - particle ids are the caller's ids (any distinct non-negative int64, e.g. the .gro ids).
 *   - no CPU fallback: if no CUDA device is usable clb_create fails with CLB_ERR_CUDA.
 *   - one engine = one GPU = one host thread.  Multi-GPU: one engine per process/rank,
 *     joined with clb_comm_init (slab domain decomposition, NCCL halo exchange).
 */
#ifndef CHEMLAB_B200_H
#define CHEMLAB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLB_ABI_VERSION 1

#define CLB_OK 0
#define CLB_ERR_ARG (-1)      /* bad argument / unknown handle / unknown particle id   */
#define CLB_ERR_CUDA (-2)     /* CUDA runtime error (message holds cudaGetErrorString) */
#define CLB_ERR_STATE (-3)    /* call not valid in the current engine state            */
#define CLB_ERR_RANGE (-4)    /* table index out of range / particle lost / overflow   */
#define CLB_ERR_UNSUPPORTED (-5)
#define CLB_ERR_COMM (-6)     /* NCCL error */

typedef struct clb_engine clb_engine;

/* ---- lifecycle ------------------------------------------------------------------------- */
/* System() + storage.DomainDecomposition(system, nodeGrid, cellGrid) + bc.OrthorhombicBC
 * + VerletList(system, cutoff) : src/start_simulation.py:148-163,193-197.
 * box = orthorhombic edge lengths; rc_max = Verlet cutoff (max_cutoff, :77-80); skin = system.skin. */
int clb_create(clb_engine **out, int device, const double box[3], double rc_max, double skin,
               uint64_t seed);
void clb_destroy(clb_engine *e);
const char *clb_last_error(const clb_engine *e); /* e may be NULL: error of the failed clb_create */
int clb_abi_version(void);
/* Device blocks released by destroyed engines are cached per process (allocation on a busy device costs milliseconds);
 * this hands them back to the CUDA driver.  No reference counterpart (memory management of the engine). */
int clb_trim_cache(void);

/* Generic numeric options (name -> double).  Known names:
 *   "resort_criterion"  0 = reference rule (sum over steps of the per-step max displacement > skin/2,
 *                           VelocityVerlet::run [EXT]); 1 = true max displacement since last rebuild (default)
 *   "block_cells"       cells per row block of the tile kernels (default 8)
 *   "list_capacity"     neighbour entries per particle (0 = automatic)
 *   "fuse_integrator"   1 = fuse second half-kick + Langevin + first half-kick + drift (default 1)
 *   "sync_chunk"        steps enqueued between host checks of the rebuild flag (default 0 = adaptive) */
int clb_set_option(clb_engine *e, const char *name, double value);
int clb_get_option(clb_engine *e, const char *name, double *value);

/* ---- particles ------------------------------------------------------------------------- */
/* storage.addParticles(list, 'id','type','pos','mass','q','res_id','state',...) + decompose():
 * src/start_simulation.py:169-171 (props: src/chemlab/gromacs_topology.py:1426).
 * pos/vel are [n][3] row-major; vel, q, state, res_id may be NULL (zeros).
 * Positions are folded into the box and stored on a 2^32-per-edge fixed-point lattice
 * (resolution L/2^32; DESIGN.md "state representation"); image counters keep the unfolding. */
int clb_set_particles(clb_engine *e, int64_t n, const int64_t *id, const int32_t *type,
                      const double *pos, const double *vel, const double *mass, const double *q,
                      const int32_t *state, const int32_t *res_id);
int64_t clb_num_particles(const clb_engine *e);

/* storage.getParticle(pid).{pos,v,f,type,state,mass,q,res_id,imageBox}: src/start_simulation.py:855-873,
 * src/chemlab/files_io.py:268-279.  ids == NULL -> all particles in ascending-id order.
 * Any output pointer may be NULL.  pos is the folded position (lattice value, exactly what the
 * kernels use); image is the periodic image count per dimension. */
int clb_get_particles(clb_engine *e, int64_t n, const int64_t *ids, double *pos, int32_t *image,
                      double *vel, double *force, int32_t *type, int32_t *state, double *mass,
                      double *q, int32_t *res_id);

/* storage.modifyParticle(pid, prop, value): examples/atrp_lj/hooks.py:47-80.
 * field: 0 type, 1 state, 2 mass, 3 q, 4 res_id, 5 pos (3 values), 6 vel (3 values). */
int clb_modify_particle(clb_engine *e, int64_t id, int field, const double *value);
/* bulk overwrite of velocities / positions in ascending-id order (tools.velocities.gaussian, :136-146) */
int clb_set_velocities(clb_engine *e, int64_t n, const double *vel);
int clb_set_positions(clb_engine *e, int64_t n, const double *pos);

/* ---- exclusions ------------------------------------------------------------------------ */
/* DynamicExcludeList(integrator, exclusions): src/start_simulation.py:189.  pairs = [n][2] ids. */
int clb_set_exclusions(clb_engine *e, int64_t n, const int64_t *pairs);
int64_t clb_num_exclusions(const clb_engine *e);
int clb_get_exclusions(clb_engine *e, int64_t cap, int64_t *pairs, int64_t *n_out);
/* DynamicExcludeList.observe_tuple/triple/quadruple(list): src/start_simulation.py:378-391,428-441:
 * tuples appended to `list` later (reactions) add the exclusion (first id, last id). */
int clb_exclusions_observe(clb_engine *e, int list);

/* ---- tabulated functions ---------------------------------------------------------------- */
/* interaction.Tabulated(itype, filename, cutoff) / TabulatedAngular / TabulatedDihedral:
 * src/chemlab/gromacs_topology.py:696-707,919-925,1074-1080,1192-1198.  The caller parses the
 * `.pot` file (rows "r e f", tools/convert_gromacs2espp.py:84,107) and passes the columns.
 * x must be uniformly spaced.  interp: 1 linear, 2 Akima, 3 cubic spline (itype). */
int clb_add_table(clb_engine *e, int64_t n, const double *x, const double *energy,
                  const double *force, int interp, int *table_out);

/* ---- non-bonded interactions over the Verlet list -------------------------------------- */
#define CLB_NB_TABULATED 1       /* interaction.VerletListTabulated        gromacs_topology.py:512 */
#define CLB_NB_LENNARD_JONES 2   /* interaction.VerletListLennardJones     gromacs_topology.py:511 */
#define CLB_NB_MIXED_TABULATED 3 /* interaction.VerletListMixedTabulated   gromacs_topology.py:757-790 */
/* system.addInteraction(interaction, label): the returned handle is the index used by
 * clb_energy (analysis.PotentialEnergy(system, interaction), src/start_simulation.py:470-480). */
int clb_add_nonbonded(clb_engine *e, int kind, int *interaction_out);
/* .setPotential(type1, type2, potential) for each kind.  A type pair may be owned by one
 * non-bonded interaction only (chemlab never assigns two: gromacs_topology.py:678-790). */
int clb_nb_set_tabulated(clb_engine *e, int interaction, int type1, int type2, int table,
                         double cutoff);
int clb_nb_set_lj(clb_engine *e, int interaction, int type1, int type2, double epsilon,
                  double sigma, double cutoff, int shift_auto);
/* MixedTabulated(itype, tab1, tab2, mix_value, cutoff): U = x*tab1 + (1-x)*tab2.  If
 * conv_type >= 0 the mixing value follows analysis.ChemicalConversion(system, conv_type, conv_total)
 * = N(type)/total, re-evaluated after every reaction step; else x = mix_value is constant. */
int clb_nb_set_mixed(clb_engine *e, int interaction, int type1, int type2, int table1, int table2,
                     double mix_value, int conv_type, double conv_total, double cutoff);

/* ---- fixed tuple lists and bonded interactions ----------------------------------------- */
/* FixedPairList / FixedTripleList / FixedQuadrupleList(storage): gromacs_topology.py:1019,1143,1272;
 * reaction_setup.py:449.  arity = 2, 3 or 4. */
int clb_add_list(clb_engine *e, int arity, int *list_out);
/* fpl.addBonds / addTriples / addQuadruples: ids = [n][arity]. */
int clb_list_add(clb_engine *e, int list, int64_t n, const int64_t *ids);
/* fpl.totalSize(), fpl.getAllBonds()  (src/start_simulation.py:885-960) */
int64_t clb_list_size(clb_engine *e, int list);
int clb_list_get(clb_engine *e, int list, int64_t cap, int64_t *ids, int64_t *n_out);

#define CLB_POT_HARMONIC 1           /* Harmonic(K, r0): U = K (r-r0)^2            gromacs_topology.py:918,949 */
#define CLB_POT_TABULATED 2          /* Tabulated(itype, file)                     gromacs_topology.py:919-925 */
#define CLB_POT_ANGULAR_HARMONIC 3   /* AngularHarmonic(K, theta0): K (th-th0)^2   gromacs_topology.py:1073,1086 */
#define CLB_POT_TABULATED_ANGULAR 4  /* TabulatedAngular(itype, file)              gromacs_topology.py:1074-1080 */
#define CLB_POT_TABULATED_DIHEDRAL 5 /* TabulatedDihedral(itype, file)             gromacs_topology.py:1192-1198 */
#define CLB_POT_COSINE 6             /* Cosine(K, theta0): K (1 + cos(th-th0))     gromacs_topology.py:1082 */
#define CLB_POT_FENE 7               /* FENE(K, r0, rMax)                          gromacs_topology.py:949-961 */
#define CLB_POT_DIHEDRAL_HARMONIC 8  /* DihedralHarmonic(K, phi0): K (phi-phi0)^2  gromacs_topology.py:1206-1224 */
#define CLB_POT_FENE_LJ 9            /* FENELennardJones(K, r0, rMax, sigma, epsilon) gromacs_topology.py:935-961  */
#define CLB_POT_LENNARD_JONES 10     /* LennardJones(epsilon, sigma, cutoff) on a pair list: the 1-4 [ pairs ]        gromacs_topology.py:1314-1411;
                                      * params {epsilon, sigma, cutoff, shift}: U = 4 eps [(s/r)^12 - (s/r)^6] - shift for r <= cutoff, 0 beyond */
/* interaction.FixedPairList<Pot>(system, fpl, pot) (typed=0: one potential for the whole list) or
 * interaction.FixedPairListTypes<Pot>(system, fpl) (typed=1: potential chosen by particle types);
 * same for Triple/Quadruple lists.  Returns the interaction handle (system.addInteraction order). */
int clb_add_bonded(clb_engine *e, int list, int typed, int *interaction_out);
/* setPotential(type1, type2[, type3[, type4]], potential); for typed=0 the types are ignored.
 * params: HARMONIC {K, r0}; ANGULAR_HARMONIC {K, theta0}; COSINE {K, theta0}; FENE {K, r0, rMax};
 * DIHEDRAL_HARMONIC {K, phi0}; FENE_LJ {K, r0, rMax, sigma, epsilon}; LENNARD_JONES {epsilon, sigma, cutoff, shift}; TABULATED* {} with table handle. Unused types = -1. */
int clb_bonded_set_potential(clb_engine *e, int interaction, int t1, int t2, int t3, int t4,
                             int pot_kind, const double *params, int nparams, int table);

/* analysis.PotentialEnergy(system, interaction).compute(): src/start_simulation.py:470-480. */
int clb_energy(clb_engine *e, int interaction, double *energy_out);
/* analysis.KineticEnergy / Temperature / NPart: src/start_simulation.py:449-462.
 * out = {Ekin, T (2 Ekin / (3 N kB=1)), N}. */
int clb_kinetics(clb_engine *e, double out[3]);
/* number of particles of `type` (analysis.ChemicalConversion numerator, :500-520); state<0: any state */
int clb_count_type(clb_engine *e, int type, int state, int64_t *count_out);

/* ---- integrator ------------------------------------------------------------------------ */
/* integrator.VelocityVerlet(system); integrator.dt: src/start_simulation.py:165-167 */
int clb_set_dt(clb_engine *e, double dt);
/* integrator.LangevinThermostat(system); .temperature (= T*kB), .gamma, .add_valid_types:
 * src/start_simulation.py:330-336,353-354.  ntypes == 0: all types thermalised. */
int clb_set_langevin(clb_engine *e, int enabled, double kT, double gamma, int ntypes,
                     const int32_t *types);
/* integrator.CapForce(system, capForce) added with --max_force (src/start_simulation.py:320-324): after every force evaluation a
 * force vector longer than cap is scaled back to that length (before the thermostat acts).  cap <= 0: off. */
int clb_set_cap_force(clb_engine *e, double cap);
/* integrator.run(n): src/start_simulation.py:780.  Synchronous. */
int clb_run(clb_engine *e, int64_t nsteps);
/* Continuation of the previous clb_run inside ONE integrator.run(n) of the reference: ExtAnalyze and
 * ATRPActivator fire from signals inside VelocityVerlet::run (src/start_simulation.py:566-569, :780;
 * src/chemlab/reaction_post_process.py:393-424), so the run-entry force recalculation and the thermostat
 * heat-up happen once per integrator.run.  Falls back to clb_run when the last forces are not reusable
 * (particles, lists, exclusions or potentials were changed by the caller in between). */
int clb_run_continue(clb_engine *e, int64_t nsteps);
int64_t clb_step(const clb_engine *e); /* integrator.step */
/* force a decompose()+VerletList rebuild / a force evaluation now (storage.decompose(), :171,205,295) */
int clb_decompose(clb_engine *e);
int clb_compute_forces(clb_engine *e);

/* ---- reactions -------------------------------------------------------------------------- */
/* integrator.ChemicalReaction(system, vl, storage, tm, interval) + .nearest_mode, .max_per_interval:
 * src/chemlab/reaction_setup.py:417-427.  enabled: integrator.addExtension(ar) / ar.disconnect()
 * (src/start_simulation.py:737,777). */
int clb_reaction_general(clb_engine *e, int enabled, int interval, int nearest_mode,
                         int max_per_interval);

typedef struct clb_reaction_spec {
    int32_t type_1, type_2;           /* integrator.Reaction(type_1, type_2, ...) reaction_setup.py:81-93 */
    int32_t delta_1, delta_2;
    int32_t min_state_1, max_state_1; /* half-open window min <= state < max */
    int32_t min_state_2, max_state_2;
    double rate;
    double cutoff;
    double min_cutoff;     /* reaction.get_reaction_cutoff().min_cutoff, reaction_setup.py:110-111 */
    int32_t list;          /* fpl= : tuple list receiving the new bond                           */
    int32_t intramolecular; /* reaction_setup.py:101 */
    int32_t intraresidual;  /* :103 */
    int32_t is_virtual;     /* :104 */
    int32_t active;         /* :112-113 */
} clb_reaction_spec;
int clb_add_reaction(clb_engine *e, const clb_reaction_spec *spec, int *reaction_out);
/* r.rate = ..., r.active = ... mid-run (src/start_simulation.py:785-796) */
int clb_reaction_set_rate(clb_engine *e, int reaction, double rate);
int clb_reaction_set_active(clb_engine *e, int reaction, int active);
/* integrator.RestrictReaction(...).define_connection(b1, b2) for every line of the group's `connectivity_map`
 * (src/chemlab/reaction_setup.py:74-75,115-126; examples/dacron/restrict/reaction.cfg:26, connections.list): the
 * reaction forms a bond only between particle pairs named in the map (either order).  pairs = [n][2] particle ids;
 * the call REPLACES the map of this reaction; n = 0 leaves a restricted reaction that accepts no pair.  A reaction
 * that never received this call is unrestricted. */
int clb_reaction_define_connections(clb_engine *e, int reaction, int64_t n, const int64_t *pairs);

/* PostProcessChangeProperty().add_change_property(old_type, TopologyParticleProperties(type, mass, q
 * [, state | incr_state])) + reaction.add_postprocess(pp, 'type_1'|'type_2'|both):
 * reaction_setup.py:137-163.  side: 1, 2 or 3 (both).  nb_level == 0: applies to the reactant itself;
 * nb_level > 0: PostProcessChangeNeighboursProperty(tm).add_change_property(type, props, nb_level)
 * (src/chemlab/reaction_post_process.py:76-115): every particle exactly nb_level bonds away whose
 * type == old_type.  mass/q < 0 (mass) or NaN (q): keep.  state_mode: 0 keep, 1 set, 2 increment. */
int clb_reaction_add_change(clb_engine *e, int reaction, int side, int nb_level, int old_type,
                            int new_type, double new_mass, double new_q, int state_mode,
                            int state_value);

/* integrator.TopologyManager(system): observe_tuple(fpl), register_triplet(ftl, t1,t2,t3),
 * register_quadruplet(fql, t1..t4), initialize_topology(): src/start_simulation.py:395-444. */
int clb_topology_observe(clb_engine *e, int pair_list);
int clb_topology_register_triplet(clb_engine *e, int triple_list, int t1, int t2, int t3);
int clb_topology_register_quadruplet(clb_engine *e, int quad_list, int t1, int t2, int t3, int t4);
int clb_topology_initialize(clb_engine *e);

/* Run one ChemicalReaction::React pass now at the current state (parity tests and
 * hook_init_reaction-style drivers).  events_out = number of applied (A,B) events. */
int clb_react_now(clb_engine *e, int64_t *events_out);
/* ar.save_reaction_counters / get counters: src/start_simulation.py:1028-1036.
 * out[reaction] = cumulated number of events. */
int clb_reaction_counters(clb_engine *e, int cap, int64_t *out);

/* integrator.ATRPActivator(system, interval, num_particles, ratio_activator, ratio_deactivator, delta_catalyst, k_activate,
 * k_deactivate) and .add_reactive_center(type_id, state, is_activator, new_property, delta_state):
 * src/chemlab/reaction_post_process.py:380-426.  The caller fires clb_atrp_now every `interval` steps (the extension's signal
 * inside VelocityVerlet::run).  One pass: of the particles that sit on a reactive centre (type, state), the num_particles with
 * the smallest counter-based random keys (seed, step, particle) are drawn; a centre written X(s,A) reacts with probability
 * k_activate*ratio_activator, X(s,DA) (needs_deactivator) with k_deactivate*ratio_deactivator; state += delta_state, type /
 * mass / charge follow new_property (new_type < 0, new_mass <= 0, new_q NaN: keep).  counts_out = {activated, deactivated} of
 * this pass, ratios_out = {activator, deactivator} ratios after it.  clb_atrp_configure clears the registered centres. */
int clb_atrp_configure(clb_engine *e, int num_particles, double ratio_activator, double ratio_deactivator,
                       double delta_catalyst, double k_activate, double k_deactivate);
int clb_atrp_add_center(clb_engine *e, int type, int state, int needs_deactivator, int new_type, double new_mass,
                        double new_q, int delta_state);
int clb_atrp_now(clb_engine *e, int64_t counts_out[2], double ratios_out[2]);

/* ---- parity / introspection -------------------------------------------------------------- */
/* The current Verlet pair set as (min id, max id) rows sorted ascending (VerletList pairs, :193-197). */
int clb_get_pairs(clb_engine *e, int64_t cap, int64_t *pairs, int64_t *n_out);
/* candidates of the last reaction pass: rows (A id, B id, reaction, accepted) sorted; d2 separately */
int clb_get_last_candidates(clb_engine *e, int64_t cap, int64_t *rows, double *d2, int64_t *n_out);

/* integrator.getTimers() / verletlist.get_timers() buckets (src/start_simulation.py:1040-1076):
 * out[0..7] = seconds in {pair force, bonded, neighbour build, integrate, comm, reaction, other, total};
 * counters[0..7] = {steps, rebuilds, kernel launches, pair-list entries (full), reaction passes,
 * reaction events, ghosts, reserved}. */
int clb_timers(clb_engine *e, double out[8], int64_t counters[8]);
int clb_reset_timers(clb_engine *e);

/* Raw device pointers for zero-copy hand-off to torch (bench / e2e staging only):
 * which: 0 pos (int4 lattice), 1 vel (float4, w = mass), 2 force (double[3][n] SoA). */
int clb_device_ptr(clb_engine *e, int which, void **ptr_dev, int64_t *n_out);
/* the CUDA stream all engine work is issued on (cudaStream_t), for CUDA-event timing by the caller */
int clb_stream(clb_engine *e, void **stream_out);

/* ---- multi-GPU --------------------------------------------------------------------------- */
/* Slab domain decomposition along z over `nranks` engines (one per GPU / process); replaces
 * storage.DomainDecomposition's MPI node grid (src/start_simulation.py:152-163).
 * nccl_id = 128-byte ncclUniqueId created by rank 0 with clb_nccl_unique_id and broadcast by
 * the caller (torch.distributed).  Must be called before clb_set_particles; every rank then
 * passes the FULL particle set and keeps the ones it owns. */
int clb_nccl_unique_id(void *id128_out);
int clb_comm_init(clb_engine *e, int rank, int nranks, const void *nccl_id128);

#ifdef __cplusplus
}
#endif
#endif /* CHEMLAB_B200_H */
