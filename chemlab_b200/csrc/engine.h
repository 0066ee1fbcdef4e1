// engine.h -- host-side state of one engine (one GPU, one stream).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/chemlab_b200.h"
#include "clb_common.cuh"
#include "clb_kernels.cuh"

// Process-wide cache of released device blocks (per device): cudaMalloc / cudaFree synchronise the device and cost from
// 0.1 ms to (measured on the B200 box, after large host allocations were freed) > 100 ms per call, so a block released by
// one engine is handed to the next request of similar size instead of going back to the driver.  clb_trim_cache() frees it.
cudaError_t clb_cache_alloc(void** p, size_t bytes, size_t* got);
void clb_cache_free(void* p, size_t bytes);

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    size_t bytes = 0;      // size of the underlying block (>= n * sizeof(T))
    // grow-only; contents are NOT preserved
    cudaError_t ensure(size_t need) {
        if (need <= n && p) return cudaSuccess;
        // regrowth is geometric: device allocation is expensive
        size_t cap = p ? need + need / 2 + 64 : (need ? need : 1);
        release();
        size_t got = 0;
        cudaError_t e = clb_cache_alloc((void**)&p, cap * sizeof(T), &got);
        if (e == cudaSuccess) { bytes = got; n = got / sizeof(T); }
        return e;
    }
    // grow and keep the first `keep` elements
    cudaError_t ensure_keep(size_t need, size_t keep, cudaStream_t st) {
        if (need <= n && p) return cudaSuccess;
        T* q = nullptr;
        size_t cap = need + need / 2 + 16, got = 0;
        cudaError_t e = clb_cache_alloc((void**)&q, cap * sizeof(T), &got);
        if (e != cudaSuccess) return e;
        if (p && keep) { e = cudaMemcpyAsync(q, p, keep * sizeof(T), cudaMemcpyDeviceToDevice, st); if (e != cudaSuccess) return e; }
        e = cudaStreamSynchronize(st);
        release();
        p = q; bytes = got; n = got / sizeof(T);
        return e;
    }
    void release() { if (p) clb_cache_free(p, bytes); p = nullptr; n = 0; bytes = 0; }
};

struct HostTable { int n, interp; double x0, dx; std::vector<double> e, f; };
struct HostPairPot {
    int kind = 0, inter = -1, tab1 = -1, tab2 = -1, conv_type = -1;
    double rc = 0, eps = 0, sig = 0, shift = 0, mix = 1.0, conv_total = 1.0;
};
struct HostInter { int kind; int bonded; };
struct HostList { int arity = 2; long long n = 0; DevBuf<int> d; int bonded = -1; int tm_observed = 0; int excl_observed = 0; };
struct HostBonded { int list, typed, inter; std::vector<ClbBPot> pots; };
struct HostChange { int reaction, side, nb_level, old_type, new_type, state_mode, state_value; double new_mass, new_q; };
struct HostTmReg { int list; int t[4]; };

enum { CLB_B_PAIR = 0, CLB_B_BONDED, CLB_B_NEIGH, CLB_B_INTEG, CLB_B_COMM, CLB_B_REACT, CLB_B_OTHER, CLB_B_TOTAL, CLB_NBUCKET };

struct clb_engine {
    std::string err;
    int fail(int code, const char* fmt, ...);

    int device = 0, nsm = 148, smem_optin = 48 * 1024;
    cudaStream_t stream = nullptr;
    double box[3], rc, skin, dt = 0.001;
    uint64_t seed = 0;
    ClbGrid grid;
    ClbGeom geo;
    int criterion = 1, criterion_user = -1, fuse = 1, chunk_user = 0, timers_on = 0, tabs_smem_user = 1;

    // particles
    int n = 0, nstored = 0, ncap = 0, own0 = 0, own1 = 0, ntypes = 1;
    std::vector<int64_t> ids;                 // slot -> caller id (ascending)
    std::unordered_map<int64_t, int> id2slot;
    int slot_of(int64_t id) const;
    bool ids_dense = false; int64_t id_base = 0;
    DevBuf<int4> pos, pos2, xref;
    DevBuf<ClbVel> vel, vel2;
    DevBuf<int> slot, slot2, id2idx, image, resid, mol;
    DevBuf<int> wslot;                        // replicated per-slot type|state word (reaction decisions; clb_react.cuh)
    DevBuf<double> force, charge;
    DevBuf<int> key, key2, val, val2, cell_start;
    DevBuf<int4> blk_table; DevBuf<int> blk_row_n, blk_row_off;     // row-block table (k_blocks_*)
    int blk_p1 = 0, blk_pl = 0, block_target_user = 0;
    DevBuf<float2> cell_disp;                   // resort criterion 2: largest displacement per cell (k_cell_disp)
    // default 1; 2 (per-cell bound) is exact as well but saves only 3 % of the rebuilds of the 1M-bead melt (66 vs 68 per 400 steps:
    // the neighbourhood of a fast bead holds 2500 others, one of which is almost as fast) -- kept as an option, tested
    int resort_criterion() const { return criterion_user >= 0 ? criterion_user : 1; }
    DevBuf<unsigned> qsub;                      // per stored particle: position inside its cell, 84 units per edge (k_qsub, k_build_lists2)
    int build_kernel_user = 0, build_kernel_active = 1;
    int make_blocks();
    DevBuf<unsigned char> cubtmp, cubtmp2, stage;
    DevBuf<double> partial;
    DevBuf<unsigned long long> partial_u64;
    ClbCtl *d_ctl = nullptr, *h_ctl = nullptr;
    void *d_scalar = nullptr, *h_scalar = nullptr;

    // neighbour lists
    DevBuf<unsigned short> nl_entries;
    DevBuf<int> nl_count;
    DevBuf<unsigned short> nl_perm;           // work order of the home particles of every block (clb_tile.cuh block_perm)
    int pair_perm_user = 1, pair_pipe_user = -1, pair_pipe = 0, pair_tile_cap = 0;
    int nl_cap = 0, nl_cap_user = 0, nl_cap_user_seen = 0, nl_max = 0, tile_max = 0, home_max = 0;
    unsigned long long nl_total = 0, last_interacting = 0;
    int pair_grid = 0, pair_threads = 128, pair_smem = 0, tabs_smem = 1, pair_split = 1, pair_split_user = 0, pair_npw = 1, build_threads = 256;
    int ugrid_on = 0, all_tab = 0, branchfree_user = 1, pair_warps_user = 0;
    ClbTabMeta ugrid_meta;
    bool lists_valid = false, forces_valid = false, cont_ok = false, bx_user = false, bx_auto_done = false;

    // exclusions
    DevBuf<int2> excl_pairs;
    long long nexcl = 0;
    DevBuf<int> excl_off, excl_ids, ekey, ekey2, eval;
    bool excl_dirty = true;

    // potentials
    std::vector<HostTable> tables;
    std::vector<HostInter> inters;
    HostPairPot pp[CLB_MAX_TYPES][CLB_MAX_TYPES];
    bool pots_dirty = true, has_mixed = false;
    int nt_dev = 1, ntabs_dev = 0, nrows_dev = 0;
    DevBuf<ClbPairDesc> d_pd, d_pd_e;
    DevBuf<double2> d_plj;
    DevBuf<ClbPairDescE> d_pe;
    DevBuf<ClbTabMeta> d_tm, d_tm_e;
    DevBuf<double2> d_frows, d_erows, d_rows2, d_pd2;
    int tab2_ok = 0, tab2_onepd = 0, tab2_one_off = 0, pair_kernel_user = 0, pair_kernel_active = 1, pair_ni = 4;
    unsigned tab2_nm1 = 0;
    int pair_nv = 1, pair_nv_user = 0, pair_vc_bytes = 0;
    // windowed multi-table kernel (k_pair_forces_tab3)
    struct T3Slot { int off, n, w0, w1; double weight; int srow; int shift; };   // shift: rows between the common grid origin and this table's first row
    std::vector<T3Slot> t3_slots;               // one per uploaded table slot (premixed rows)
    std::vector<int> t3_pair_slot;              // [nt*nt] slot of the type pair, -1: no potential
    std::vector<double> t3_pair_rc2;            // [nt*nt] cutoff^2 in lattice^2
    std::vector<double2> t3_rows;               // host copy of all {A_i, B_i} rows
    DevBuf<ClbPairDesc3> d_pd3; DevBuf<int2> d_gmeta; DevBuf<double2> d_swin; DevBuf<unsigned long long> d_hist;
    int tab3_fb = 0, pair_fb_user = -1, pair_ni_user = 0, tab3_ok = 0, tab3_onepd = 0, tab3_nsrows = 0, tab3_rlog = 0, tab3_resident = 0, pair_rep_user = -1, pair_table_kb_user = -1;
    double tab3_resident_weight = 0;
    bool t3_dirty = true;
    ClbPairDesc3 tab3_one; int2 tab3_one_g;
    int configure_tables(size_t budget_bytes, int rlog);
    double tab2_invdx = 0, tab2_cmagic = 0, tab2_one_rc2 = 0;

    // tuple lists and bonded interactions
    std::vector<HostList> lists;
    std::vector<HostBonded> bondeds;
    DevBuf<ClbBondedDesc> d_bdesc;
    DevBuf<ClbBPot> d_bpots;
    DevBuf<ClbBTabMeta> d_btm;
    DevBuf<double> d_bcf, d_bce;
    DevBuf<const int*> d_list_ptrs;
    bool terms_dirty = true, lists_ptr_dirty = true, rt_valid = false;
    long long nterms = 0;
    DevBuf<int> term_off, term_meta, term_tuple, tkey, tkey2;
    DevBuf<unsigned long long> tval, tval2;
    DevBuf<int> rt_off, rt_cnt, rt_meta;
    DevBuf<int4> rt_mem;

    // integrator
    int64_t step = 0;
    int lang_on = 0;
    double kT = 1.0, gamma = 1.0;
    double cap_force = -1.0;      // integrator.CapForce: <= 0 off
    unsigned long long lang_mask = ~0ull;
    int last_interval = 0;
    int64_t last_rebuild_step = 0;

    // reactions / topology
    int react_on = 0, react_interval = 1, react_nearest = 1, react_max_per_interval = 0;
    std::vector<clb_reaction_spec> reactions;
    std::vector<HostChange> changes;
    std::vector<HostTmReg> tmregs;
    bool react_dirty = true, topo_dirty = true, topo_initialized = false;
    std::vector<int64_t> react_counters;
    std::vector<std::vector<int64_t>> react_conn;   // RestrictReaction: connectivity map per reaction, [n][2] particle ids
    std::vector<char> react_restricted;
    struct ReactDev;
    ReactDev* rd = nullptr;
    int R_prealloc_n = -1;

    // ATRPActivator
    struct HostAtrpCenter { int type, state, deactivator, new_type, delta_state; double new_mass, new_q; };
    std::vector<HostAtrpCenter> atrp_centers;
    int atrp_num = 0; double atrp_ratio_act = 0, atrp_ratio_deact = 0, atrp_delta = 0, atrp_k_act = 0, atrp_k_deact = 0;
    DevBuf<unsigned long long> atrp_keys, atrp_keys2, atrp_scal;
    DevBuf<unsigned char> atrp_cen;

    // timers / counters
    cudaEvent_t ev_a[CLB_NBUCKET], ev_b[CLB_NBUCKET];
    int bucket_open[CLB_NBUCKET] = {0};
    double bucket_s[CLB_NBUCKET] = {0};
    int64_t nsteps_total = 0, nrebuild = 0, launches = 0, nreact_pass = 0, nreact_events = 0, pair_launches_total = 0;
    int pair_event_timing = 0;
    std::vector<cudaEvent_t> pair_events, pair_events2;   // second pair: boundary-block launch of an overlapped step
    std::vector<char> pair_event_has2;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_int = nullptr, ev_comm = nullptr;
    int overlap_user = -1;    // -1 auto: on for >= 4 ranks (measured on B200, 1M beads: N=2 0.562 -> 0.608 ms/step, N=8 0.2525 -> 0.2437)
    std::vector<char> pair_event_valid;
    size_t pair_event_used = 0;
    double pair_ms = 0;
    int64_t pair_launches = 0;
    void bucket_begin(int b);
    void bucket_end(int b);
    cudaEvent_t next_pair_event();
    void collect_timers();

    // multi-GPU
    int rank = 0, nranks = 1;
    struct CommDev;
    CommDev* cd = nullptr;
    int comm_migrate();
    int comm_exchange_ghosts();
    int comm_halo_positions(cudaStream_t st);
    int comm_max_displacement(cudaStream_t st);
    int comm_step(cudaStream_t st);
    int comm_group_user = 1;
    int peer_user = -1;       // -1 auto (peer mailboxes when cudaIpc works), 0 = NCCL on the step path
    int comm_peer_setup();
    void comm_peer_teardown();
    int comm_step_peer(cudaStream_t st, int step_index);
    int comm_migrate_peer();
    int comm_exchange_ghosts_peer();
    bool peer_active() const;
    int pending_step_index = 0;
    int comm_allreduce_sum(double* v, int n);
    int comm_allreduce_sum_dev(double* d, size_t n);
    int comm_allgatherv(const void* dsend, size_t bytes, void** dout, size_t* total);
    int comm_gather_candidates(long long* nc);
    void comm_destroy();

    // methods
    void set_block_cells(int bx);
    int alloc_particles(int nlocal_cap);
    int get_particles_gathered(int64_t nq, const int64_t* ids, double* pos, int32_t* image, double* vel, double* force, int32_t* type,
                               int32_t* state, double* mass, double* q, int32_t* res_id);
    int build_excl_csr();
    int upload_potentials();
    int list_reserve(int li, long long need);
    int build_term_csr();
    int upload_list_ptrs();
    int resolve_terms();
    int read_ctl();
    int setup_sync();
    int rebuild();
    int configure_pair_launch();
    void enqueue_forces(bool overlap_halo = false, bool resort_checked = false);
    void launch_pair(int b0, int seg0, int b1, int nidx);
    void enqueue_integrate(int mode, uint64_t key_step);
    ClbIntegParams integ_params(uint64_t key_step) const;
    int check_device_errors(const char* where);
    int build_topology();
    int upload_reactions();
    int upload_list_descs();
    bool pending_rebuild = false;
    int react_pass(int64_t* events_out);
    int update_mixing();
    void free_all();
    void react_free();
};
