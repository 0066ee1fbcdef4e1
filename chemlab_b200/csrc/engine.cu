// engine.cu -- host side of the B200 engine: C-ABI (include/chemlab_b200.h), device state,
// rebuild pipeline, stall-flag step loop.  One engine = one GPU = one stream.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include <cub/cub.cuh>

#include "../../include/chemlab_b200.h"
#include "clb_common.cuh"
#include "clb_kernels.cuh"
#include "clb_tile.cuh"
#include "engine.h"
#include "clb_react.cuh"

static std::string g_create_error;

// ---- device block cache (engine.h) ----
#include <map>
#include <mutex>
static std::mutex g_cache_mu;
static std::multimap<size_t, void*> g_cache[16];        // per device: block size -> pointer
static size_t g_cache_bytes[16] = {0};
cudaError_t clb_cache_alloc(void** p, size_t bytes, size_t* got) {
    int dev = 0; cudaGetDevice(&dev);
    bytes = (bytes + 511) & ~(size_t)511;
    if (dev >= 0 && dev < 16) {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        auto it = g_cache[dev].lower_bound(bytes);
        // reuse a cached block of at most twice the size (bounded internal fragmentation)
        if (it != g_cache[dev].end() && it->first <= 2 * bytes + (1 << 20)) {
            *p = it->second; *got = it->first; g_cache_bytes[dev] -= it->first; g_cache[dev].erase(it);
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess && dev >= 0 && dev < 16) {
        // out of memory: give the cache back to the driver and retry once
        { std::lock_guard<std::mutex> lk(g_cache_mu);
          for (auto& kv : g_cache[dev]) cudaFree(kv.second);
          g_cache[dev].clear(); g_cache_bytes[dev] = 0; }
        (void)cudaGetLastError();
        e = cudaMalloc(p, bytes);
    }
    *got = bytes;
    return e;
}
void clb_cache_free(void* p, size_t bytes) {
    if (!p) return;
    int dev = 0; cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16 || bytes == 0) { cudaFree(p); return; }
    std::lock_guard<std::mutex> lk(g_cache_mu);
    g_cache[dev].insert({bytes, p}); g_cache_bytes[dev] += bytes;
}
extern "C" int clb_trim_cache(void) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    int dev0 = 0; cudaGetDevice(&dev0);
    for (int d = 0; d < 16; ++d) {
        if (g_cache[d].empty()) continue;
        cudaSetDevice(d);
        for (auto& kv : g_cache[d]) cudaFree(kv.second);
        g_cache[d].clear(); g_cache_bytes[d] = 0;
    }
    cudaSetDevice(dev0);
    return CLB_OK;
}

#define CK(call)                                                                                 \
    do {                                                                                         \
        cudaError_t _e = (call);                                                                 \
        if (_e != cudaSuccess) return e->fail(CLB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)
#define TRY(expr)                       \
    do {                                \
        int _r = (expr);                \
        if (_r != CLB_OK) return _r;    \
    } while (0)

int clb_engine::fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    err = buf;
    return code;
}

static inline int ceil_div(long long a, int b) { return (int)((a + b - 1) / b); }

// CLB_TRACE=1: host wall-clock per stage (each mark synchronises the stream; debugging aid only)
struct ClbTrace {
    bool on; cudaStream_t st; const char* tag; std::chrono::steady_clock::time_point t0; std::string log;
    ClbTrace(cudaStream_t s, const char* tg, bool enable = true) : on(enable && getenv("CLB_TRACE") != nullptr), st(s), tag(tg), t0(std::chrono::steady_clock::now()) {}
    void mark(const char* name) {
        if (!on) return;
        cudaStreamSynchronize(st);
        auto t1 = std::chrono::steady_clock::now();
        char b[128]; snprintf(b, sizeof(b), " %s=%.3fms", name, std::chrono::duration<double, std::milli>(t1 - t0).count());
        log += b; t0 = t1;
    }
    ~ClbTrace() { if (on && !log.empty()) fprintf(stderr, "[clb %s]%s\n", tag, log.c_str()); }
};

typedef void (*PairKernel)(ClbGrid, ClbGeom, ClbPairArgs);
template <bool C, bool S, bool U>
static PairKernel pair_kernel_split(int split) {
    switch (split) { case 1: return k_pair_forces<C, S, U, 1>; case 2: return k_pair_forces<C, S, U, 2>; default: return k_pair_forces<C, S, U, 4>; }
}
template <bool C, bool S, bool U>
static PairKernel pair_kernel_tab_split(int split) {
    switch (split) { case 1: return k_pair_forces_tab<C, S, U, 1>; case 2: return k_pair_forces_tab<C, S, U, 2>; default: return k_pair_forces_tab<C, S, U, 4>; }
}
static PairKernel pair_kernel_tab(int cubic, int smem, int ugrid, int split) {
    if (cubic) { if (smem) return ugrid ? pair_kernel_tab_split<true, true, true>(split) : pair_kernel_tab_split<true, true, false>(split);
                 return ugrid ? pair_kernel_tab_split<true, false, true>(split) : pair_kernel_tab_split<true, false, false>(split); }
    if (smem) return ugrid ? pair_kernel_tab_split<false, true, true>(split) : pair_kernel_tab_split<false, true, false>(split);
    return ugrid ? pair_kernel_tab_split<false, false, true>(split) : pair_kernel_tab_split<false, false, false>(split);
}
typedef void (*PairKernel2)(ClbGrid, ClbPairArgs2);
static PairKernel2 pair_kernel_tab2(int smem, int onepd, int ni) {
    if (ni >= 4) { if (smem) return onepd ? k_pair_forces_tab2<true, true, 4> : k_pair_forces_tab2<true, false, 4>;
                   return onepd ? k_pair_forces_tab2<false, true, 4> : k_pair_forces_tab2<false, false, 4>; }
    if (smem) return onepd ? k_pair_forces_tab2<true, true, 2> : k_pair_forces_tab2<true, false, 2>;
    return onepd ? k_pair_forces_tab2<false, true, 2> : k_pair_forces_tab2<false, false, 2>;
}
typedef void (*PairKernel3)(ClbGrid, ClbPairArgs3);
// variants: one descriptor in registers (ONEPD, optionally 8-way replicated rows) with the rare out-of-line global path;
// many tables with the out-of-line path (all hot windows resident) or with the inline predicated global gather (fb = 1)
static PairKernel3 pair_kernel_tab3(int onepd, int rlog, int ni, int fb) {
    if (onepd) { if (rlog) return ni >= 4 ? k_pair_forces_tab3<true, 3, 4, false> : k_pair_forces_tab3<true, 3, 2, false>;
                 return ni >= 4 ? k_pair_forces_tab3<true, 0, 4, false> : k_pair_forces_tab3<true, 0, 2, false>; }
    if (fb) return ni >= 4 ? k_pair_forces_tab3<false, 0, 4, true> : k_pair_forces_tab3<false, 0, 2, true>;   // (4 spills at 64 registers: 2 is the default)
    return ni >= 4 ? k_pair_forces_tab3<false, 0, 4, false> : k_pair_forces_tab3<false, 0, 2, false>;
}
static PairKernel pair_kernel(int cubic, int smem, int ugrid, int split) {
    if (cubic) { if (smem) return ugrid ? pair_kernel_split<true, true, true>(split) : pair_kernel_split<true, true, false>(split);
                 return ugrid ? pair_kernel_split<true, false, true>(split) : pair_kernel_split<true, false, false>(split); }
    if (smem) return ugrid ? pair_kernel_split<false, true, true>(split) : pair_kernel_split<false, true, false>(split);
    return ugrid ? pair_kernel_split<false, false, true>(split) : pair_kernel_split<false, false, false>(split);
}

// ------------------------------------------------------------------------------------------
extern "C" int clb_abi_version(void) { return CLB_ABI_VERSION; }

extern "C" const char* clb_last_error(const clb_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

extern "C" int clb_create(clb_engine** out, int device, const double box[3], double rc_max, double skin, uint64_t seed) {
    if (!out || !box || rc_max <= 0 || skin < 0) { g_create_error = "clb_create: bad argument"; return CLB_ERR_ARG; }
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        g_create_error = std::string("clb_create: no usable CUDA device (") + cudaGetErrorString(ce) + "); this engine has no CPU fallback";
        return CLB_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { g_create_error = "clb_create: bad device index"; return CLB_ERR_ARG; }
    clb_engine* e = new clb_engine();
    e->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { g_create_error = "cudaSetDevice failed"; delete e; return CLB_ERR_CUDA; }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    e->nsm = prop.multiProcessorCount;
    e->smem_optin = (int)prop.sharedMemPerBlockOptin;
    cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
    for (int d = 0; d < 3; ++d) e->box[d] = box[d];
    e->rc = rc_max; e->skin = skin; e->seed = seed;
    double rl = rc_max + skin;
    ClbGrid& g = e->grid;
    g.ncx = (int)floor(box[0] / rl); g.ncy = (int)floor(box[1] / rl); g.ncz = (int)floor(box[2] / rl);
    if (g.ncx < 3 || g.ncy < 3 || g.ncz < 3) {
        g_create_error = "clb_create: the box must hold at least 3 cells of edge rc+skin per dimension";
        delete e; return CLB_ERR_UNSUPPORTED;
    }
    if ((long long)g.ncx * g.ncy * g.ncz > (1ll << 30)) { g_create_error = "too many cells"; delete e; return CLB_ERR_UNSUPPORTED; }
    g.cz0 = 0; g.nczl = g.ncz; g.zoff = 0; g.nplanes = g.ncz; g.ghost = 0;
    g.blk = nullptr; g.nblk_d = nullptr; g.target = 0; g.nblocks = 0;
    e->geo.rl2 = rl * rl;
    for (int d = 0; d < 3; ++d) {
        e->geo.box[d] = box[d];
        e->geo.q[d] = box[d] / 4294967296.0;
        e->geo.cut[d] = (int)std::min(2147483000.0, ceil(rl / e->geo.q[d]) + 1.0);
    }
    e->geo.cubic = (box[0] == box[1] && box[1] == box[2]);
    e->geo.q2 = e->geo.q[0] * e->geo.q[0];
    e->set_block_cells(8);
    (void)cudaGetLastError();   // clear a non-sticky error left by an earlier call in this process
    ce = cudaMalloc(&e->d_ctl, sizeof(ClbCtl));
    if (ce == cudaSuccess) ce = cudaMemset(e->d_ctl, 0, sizeof(ClbCtl));
    if (ce == cudaSuccess) ce = cudaMallocHost(&e->h_ctl, sizeof(ClbCtl));
    if (ce == cudaSuccess) ce = cudaMalloc(&e->d_scalar, 64);
    if (ce == cudaSuccess) ce = cudaMallocHost(&e->h_scalar, 64);
    if (ce != cudaSuccess) {
        g_create_error = std::string("clb_create: ") + cudaGetErrorString(ce);
        delete e; return CLB_ERR_CUDA;
    }
    memset(e->h_ctl, 0, sizeof(ClbCtl));
    // opt-in shared memory for the tile kernels
    // static shared memory of the tile kernels (offset tables, reaction specs, reduction scratch) stays below 8 KB
    e->smem_optin -= 8192;
    int mx = e->smem_optin;
    cudaFuncSetAttribute(k_build_lists<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    cudaFuncSetAttribute(k_build_lists<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    cudaFuncSetAttribute(k_build_lists2, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    cudaFuncSetAttribute(k_build_lists2, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(k_build_lists<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(k_build_lists<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    for (int c = 0; c < 2; ++c) for (int sm = 0; sm < 2; ++sm) for (int ug = 0; ug < 2; ++ug) for (int sp = 0; sp < 3; ++sp)
    { cudaFuncSetAttribute(pair_kernel_tab(c, sm, ug, 1 << sp), cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
          cudaFuncSetAttribute(pair_kernel_tab(c, sm, ug, 1 << sp), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
          cudaFuncSetAttribute(pair_kernel(c, sm, ug, 1 << sp), cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
          cudaFuncSetAttribute(pair_kernel(c, sm, ug, 1 << sp), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); }
    for (int sm = 0; sm < 2; ++sm) for (int op = 0; op < 2; ++op) for (int ni = 2; ni <= 4; ni += 2) {
        cudaFuncSetAttribute(pair_kernel_tab2(sm, op, ni), cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        cudaFuncSetAttribute(pair_kernel_tab2(sm, op, ni), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    for (int op = 0; op < 2; ++op) for (int rl = 0; rl <= 3; rl += 3) for (int ni = 2; ni <= 4; ni += 2) for (int fb = 0; fb < 2; ++fb) {
        cudaFuncSetAttribute(pair_kernel_tab3(op, rl, ni, fb), cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        cudaFuncSetAttribute(pair_kernel_tab3(op, rl, ni, fb), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    cudaFuncSetAttribute(k_pair_energy<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    cudaFuncSetAttribute(k_pair_energy<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    cudaFuncSetAttribute(k_decode_pairs, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    cudaFuncSetAttribute(k_react_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    for (int i = 0; i < CLB_NBUCKET; ++i) { cudaEventCreate(&e->ev_a[i]); cudaEventCreate(&e->ev_b[i]); }
    ce = cudaGetLastError();
    if (ce != cudaSuccess) { g_create_error = std::string("clb_create: ") + cudaGetErrorString(ce); delete e; return CLB_ERR_CUDA; }
    *out = e;
    return CLB_OK;
}

void clb_engine::set_block_cells(int bx) {
    bx = std::max(1, std::min(bx, CLB_MAX_BX));
    grid.bx = bx;
    grid.ncell = grid.ncx * grid.ncy * grid.nplanes;
    grid.nblocks = 0;                  // known after the next rebuild (make_blocks)
}
// Row-block table of the current cell lists (clb_kernels.cuh).  target: home particles per block = the lanes of the warps that
// work on one tile (whole warps; about 160 keeps the tile near 1800 particles at melt densities).
int clb_engine::make_blocks() {
    clb_engine* e = this;
    const int nrows = grid.ncy * grid.nczl;
    CK(blk_table.ensure((size_t)std::max(1, nrows * grid.ncx)));
    CK(blk_row_n.ensure(nrows + 1)); CK(blk_row_off.ensure(nrows + 1));
    grid.blk = blk_table.p; grid.nblk_d = &d_ctl->nblocks;
    if (block_target_user != 0) grid.target = block_target_user;
    else {
        // measured (profiles/r2k_pipe_c2.log, r2t_sweep_c{3,5}.log): one-table melts are flat between 160 and 256 (160 keeps 5 tiles
        // resident); many-table systems, whose table windows take half of the shared memory, run best with 7-warp blocks
        const double ppc = (double)n / ((double)grid.ncx * grid.ncy * grid.ncz);
        int t = t3_slots.size() > 1 ? 224 : 160;
        while (t > 64 && 9.0 * (t + 2.0 * ppc) > 2900.0) t -= 32;      // tile = 9 rows of (home + 2 cells), at most ~46 KB
        grid.target = t;
    }
    k_blocks_rows<false><<<ceil_div(nrows, 128), 128, 0, stream>>>(grid, cell_start.p, blk_row_n.p, nullptr, nullptr);
    k_blocks_scan<<<1, 1024, 0, stream>>>(nrows, grid.ncy, grid.nczl, blk_row_n.p, blk_row_off.p, d_ctl);
    k_blocks_rows<true><<<ceil_div(nrows, 128), 128, 0, stream>>>(grid, cell_start.p, nullptr, blk_row_off.p, blk_table.p);
    launches += 3;
    return CLB_OK;
}

extern "C" void clb_destroy(clb_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    e->free_all();
    delete e;
}

extern "C" int clb_set_option(clb_engine* e, const char* name, double v) {
    if (!e || !name) return CLB_ERR_ARG;
    std::string s(name);
    if (s == "resort_criterion") { e->criterion_user = (int)v; e->criterion = e->resort_criterion(); }
    else if (s == "step") e->step = (int64_t)v;               // integrator.step (restart): keys the thermostat and reaction draws
    else if (s == "block_cells") { e->set_block_cells((int)v); e->bx_user = (int)v > 0; e->lists_valid = false; }
    else if (s == "block_target") { e->block_target_user = (int)v; e->lists_valid = false; }
    else if (s == "build_kernel") { e->build_kernel_user = (int)v; e->lists_valid = false; }
    else if (s == "pair_perm") e->pair_perm_user = (int)v;
    else if (s == "pair_pipe") { e->pair_pipe_user = (int)v; e->lists_valid = false; }
    else if (s == "list_capacity") { e->nl_cap_user = (int)v; e->lists_valid = false; }
    else if (s == "fuse_integrator") e->fuse = (int)v;
    else if (s == "sync_chunk") e->chunk_user = (int)v;
    else if (s == "timers") e->timers_on = (int)v;
    else if (s == "tables_in_smem") { e->tabs_smem_user = (int)v; if (e->lists_valid) return e->configure_pair_launch(); }
    else if (s == "pair_event_timing") e->pair_event_timing = (int)v;
    else if (s == "overlap_halo") e->overlap_user = (int)v;
    else if (s == "comm_group") e->comm_group_user = (int)v;
    else if (s == "comm_peer") e->peer_user = (int)v;
    else if (s == "pair_split") { e->pair_split_user = (int)v; if (e->lists_valid) return e->configure_pair_launch(); }
    else if (s == "build_threads") e->build_threads = (int)v;
    else if (s == "pair_warps") { e->pair_warps_user = (int)v; if (e->lists_valid) return e->configure_pair_launch(); }
    else if (s == "pair_kernel") { e->pair_kernel_user = (int)v; if (e->lists_valid) return e->configure_pair_launch(); }
    else if (s == "pair_nv") { e->pair_nv_user = (int)v; if (e->lists_valid) return e->configure_pair_launch(); }
    else if (s == "pair_ni") { e->pair_ni = (int)v >= 4 ? 4 : 2; e->pair_ni_user = (int)v > 0; if (e->lists_valid) return e->configure_pair_launch(); }
    else if (s == "pair_rep") { e->pair_rep_user = (int)v; e->t3_dirty = true; if (e->lists_valid) return e->configure_pair_launch(); }
    else if (s == "pair_inline_fallback") { e->pair_fb_user = (int)v; if (e->lists_valid) return e->configure_pair_launch(); }
    else if (s == "pair_table_kb") { e->pair_table_kb_user = (int)v; e->t3_dirty = true; if (e->lists_valid) return e->configure_pair_launch(); }
    else if (s == "pair_branchfree") { e->branchfree_user = (int)v; if (e->lists_valid) return e->configure_pair_launch(); }
    else return e->fail(CLB_ERR_ARG, "unknown option %s", name);
    return CLB_OK;
}
extern "C" int clb_get_option(clb_engine* e, const char* name, double* v) {
    if (!e || !name || !v) return CLB_ERR_ARG;
    std::string s(name);
    if (s == "resort_criterion") *v = e->resort_criterion();
    else if (s == "block_cells") *v = e->grid.bx;
    else if (s == "block_target") *v = e->grid.target;
    else if (s == "blocks") *v = e->grid.nblocks;
    else if (s == "build_kernel") *v = e->build_kernel_active;
    else if (s == "pair_pipe") *v = e->pair_pipe;
    else if (s == "list_capacity") *v = e->nl_cap;
    else if (s == "fuse_integrator") *v = e->fuse;
    else if (s == "sync_chunk") *v = e->chunk_user;
    else if (s == "timers") *v = e->timers_on;
    else if (s == "tile_max") *v = e->tile_max;
    else if (s == "home_max") *v = e->home_max;
    else if (s == "nl_max") *v = e->nl_max;
    else if (s == "nl_total") *v = (double)e->nl_total;
    else if (s == "ncx") *v = e->grid.ncx;
    else if (s == "pair_grid") *v = e->pair_grid;
    else if (s == "comm_peer") *v = e->peer_active() ? 1 : 0;
    else if (s == "pair_threads") *v = e->pair_threads;
    else if (s == "pair_split") *v = e->pair_split;
    else if (s == "pair_kernel") *v = e->pair_kernel_active;
    else if (s == "pair_onepd") *v = e->tab2_onepd;
    else if (s == "pair_nv") *v = e->pair_nv;
    else if (s == "uniform_grid") *v = e->ugrid_on;
    else if (s == "pair_smem") *v = e->pair_smem;
    else if (s == "tables_in_smem") *v = e->tabs_smem;
    else if (s == "pair_kernel_ms") *v = e->pair_ms;
    else if (s == "pair_rep") *v = 1 << e->tab3_rlog;
    else if (s == "pair_inline_fallback") *v = e->tab3_fb;
    else if (s == "pair_table_rows") *v = e->tab3_nsrows;
    else if (s == "pair_tables_resident") *v = e->tab3_resident;
    else if (s == "pair_tables_resident_weight") *v = e->tab3_resident_weight;
    else if (s == "pair_kernel_launches") *v = (double)e->pair_launches;
    else return e->fail(CLB_ERR_ARG, "unknown option %s", name);
    return CLB_OK;
}

// ------------------------------------------------------------------------------------------ particles
int clb_engine::slot_of(int64_t id) const {
    if (ids_dense) { int64_t s = id - id_base; return (s >= 0 && s < (int64_t)n) ? (int)s : -1; }
    auto it = id2slot.find(id);
    return it == id2slot.end() ? -1 : it->second;
}

// ---- host arrays <-> device state.  The conversion (fold into the box, 2^32 lattice, image counters, packing) runs on
// the device: the host only moves the caller's arrays (pinned buffers are copied by DMA, pageable ones through the driver)
struct ClbIngestArgs {
    const double *pos, *vel, *mass, *q; const int *type, *state, *resid, *order;
    double box[3];
};
__global__ void k_ingest(int n, ClbIngestArgs A, int4* __restrict__ P, ClbVel* __restrict__ V, int* __restrict__ image, int* __restrict__ res_o,
                         double* __restrict__ charge, int* __restrict__ wslot, int* __restrict__ flags) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int i = A.order ? A.order[s] : s;
    int xi[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        // same arithmetic as the host formula of round 1 (IEEE division, floor, rint): bit-identical lattice values
        double fr = A.pos[3 * (size_t)i + d] / A.box[d];
        double fl = floor(fr);
        double u = rint((fr - fl) * 4294967296.0);
        int im = (int)fl;
        if (u >= 4294967296.0) { u -= 4294967296.0; im += 1; }
        xi[d] = (int)(unsigned)(unsigned long long)u;
        image[3 * (size_t)s + d] = im;
    }
    const int t = A.type[i];
    if (t < 0 || t >= CLB_MAX_TYPES) { atomicMin(flags + 1, i); }
    atomicMax(flags, t);
    const int w = pw_pack(t, A.state ? A.state[i] : 0);
    P[s] = make_int4(xi[0], xi[1], xi[2], w);
    V[s] = clb_make_vel(A.vel ? A.vel[3 * (size_t)i] : 0.0, A.vel ? A.vel[3 * (size_t)i + 1] : 0.0, A.vel ? A.vel[3 * (size_t)i + 2] : 0.0, A.mass[i]);
    wslot[s] = w;
    res_o[s] = A.resid ? A.resid[i] : 0;
    charge[s] = A.q ? A.q[i] : 0.0;
}
__global__ void k_identity_slots(int n, int* __restrict__ slot, int* __restrict__ id2idx) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) { slot[s] = s; id2idx[s] = s; }
}
// multi-GPU: flag the particles whose cell plane this rank owns
__global__ void k_owned_flags(int n, const int4* __restrict__ P, ClbGrid g, unsigned char* __restrict__ flag) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int cz = __umulhi((unsigned)P[s].z, (unsigned)g.ncz);
    int l = cz - g.cz0; if (l < 0) l += g.ncz;
    flag[s] = l < g.nczl ? 1 : 0;
}
__global__ void k_take_owned(int nloc, const int* __restrict__ sel, const int4* __restrict__ P, const ClbVel* __restrict__ V,
                             int4* __restrict__ pos, ClbVel* __restrict__ vel, int* __restrict__ slot, int* __restrict__ id2idx) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nloc) return;
    const int s = sel[k];
    pos[k] = P[s]; vel[k] = V[s]; slot[k] = s; id2idx[s] = k;
}

// staging area on the device for whole-array transfers (grow-only, reused by set/get)
static cudaError_t stage_reserve(clb_engine* e, size_t bytes) { return e->stage.ensure(bytes + 256); }
struct StageCursor {
    unsigned char* base; size_t off = 0;
    template <typename T> T* take(size_t count) { off = (off + 255) & ~(size_t)255; T* p = reinterpret_cast<T*>(base + off); off += count * sizeof(T); return p; }
};

extern "C" int clb_set_particles(clb_engine* e, int64_t n, const int64_t* id, const int32_t* type, const double* pos,
                                 const double* vel, const double* mass, const double* q, const int32_t* state,
                                 const int32_t* res_id) {
    if (!e || n <= 0 || !id || !type || !pos || !mass) return e ? e->fail(CLB_ERR_ARG, "clb_set_particles: bad argument") : CLB_ERR_ARG;
    if (n >= (1ll << 28)) return e->fail(CLB_ERR_UNSUPPORTED, "more than 2^28 particles");
    cudaSetDevice(e->device);
    // slots = rank in ascending id order.  Ids given as a dense ascending run (the usual .gro numbering) need neither a
    // sort nor a hash map: slot = id - first id.
    bool dense = true;
    for (int64_t i = 1; i < n && dense; ++i) dense = id[i] == id[0] + i;
    std::vector<int> order;
    e->ids.assign(id, id + n);
    e->id2slot.clear();
    e->ids_dense = dense; e->id_base = id[0];
    if (!dense) {
        order.resize(n);
        for (int64_t i = 0; i < n; ++i) order[i] = (int)i;
        std::sort(order.begin(), order.end(), [&](int a, int b) { return id[a] < id[b]; });
        e->id2slot.reserve(n * 2);
        for (int64_t s = 0; s < n; ++s) {
            e->ids[s] = id[order[s]];
            if (s && e->ids[s] == e->ids[s - 1]) return e->fail(CLB_ERR_ARG, "duplicate particle id %lld", (long long)e->ids[s]);
            e->id2slot[e->ids[s]] = (int)s;
        }
    }
    e->n = (int)n;
    for (char c : e->react_restricted) if (c) e->react_dirty = true;     // connectivity maps are resolved to slots at upload
    ClbTrace tr(e->stream, "set_particles");
    tr.mark("ids");
    // per-slot arrays are full size on every rank
    CK(e->id2idx.ensure(n)); CK(e->image.ensure(3 * (size_t)n)); CK(e->resid.ensure(n)); CK(e->charge.ensure(n)); CK(e->mol.ensure(n)); CK(e->wslot.ensure(n));
    const size_t N = (size_t)n;
    CK(stage_reserve(e, N * (24 + 24 + 8 + 8 + 4 + 4 + 4 + 4 + 16 + 32 + 4 + 1) + 16 * 256));
    StageCursor sc{e->stage.p};
    double* d_pos = sc.take<double>(3 * N); double* d_vel = vel ? sc.take<double>(3 * N) : nullptr; double* d_mass = sc.take<double>(N);
    double* d_q = q ? sc.take<double>(N) : nullptr; int* d_type = sc.take<int>(N); int* d_state = state ? sc.take<int>(N) : nullptr;
    int* d_res = res_id ? sc.take<int>(N) : nullptr; int* d_order = dense ? nullptr : sc.take<int>(N);
    int4* d_P = sc.take<int4>(N); ClbVel* d_V = sc.take<ClbVel>(N); int* d_sel = sc.take<int>(N); unsigned char* d_flag = sc.take<unsigned char>(N);
    int* d_flags = sc.take<int>(4);
    cudaStream_t st = e->stream;
    tr.mark("alloc_stage");
    CK(cudaMemcpyAsync(d_pos, pos, 24 * N, cudaMemcpyHostToDevice, st));
    if (vel) CK(cudaMemcpyAsync(d_vel, vel, 24 * N, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_mass, mass, 8 * N, cudaMemcpyHostToDevice, st));
    if (q) CK(cudaMemcpyAsync(d_q, q, 8 * N, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_type, type, 4 * N, cudaMemcpyHostToDevice, st));
    if (state) CK(cudaMemcpyAsync(d_state, state, 4 * N, cudaMemcpyHostToDevice, st));
    if (res_id) CK(cudaMemcpyAsync(d_res, res_id, 4 * N, cudaMemcpyHostToDevice, st));
    if (!dense) CK(cudaMemcpyAsync(d_order, order.data(), 4 * N, cudaMemcpyHostToDevice, st));
    const int hflags0[4] = {0, 0x7fffffff, 0, 0};
    CK(cudaMemcpyAsync(d_flags, hflags0, 16, cudaMemcpyHostToDevice, st));
    ClbIngestArgs A; A.pos = d_pos; A.vel = d_vel; A.mass = d_mass; A.q = d_q; A.type = d_type; A.state = d_state; A.resid = d_res; A.order = d_order;
    for (int d = 0; d < 3; ++d) A.box[d] = e->box[d];
    const bool multi = e->nranks > 1;
    tr.mark("h2d");
    if (!multi) TRY(e->alloc_particles(e->n));
    tr.mark("alloc_particles");
    // single rank: the ingest kernel writes the particle arrays directly (slot order = initial order)
    k_ingest<<<ceil_div(n, 256), 256, 0, st>>>((int)n, A, multi ? d_P : e->pos.p, multi ? d_V : e->vel.p, e->image.p, e->resid.p, e->charge.p, e->wslot.p, d_flags);
    int64_t nloc = n;
    if (multi) {
        // every rank receives the full set and keeps the particles of its own cell planes (stable compaction)
        k_owned_flags<<<ceil_div(n, 256), 256, 0, st>>>((int)n, d_P, e->grid, d_flag);
        size_t tb = 0;
        cub::DeviceSelect::Flagged(nullptr, tb, cub::CountingInputIterator<int>(0), d_flag, d_sel, d_flags + 2, (int)n, st);
        CK(e->cubtmp2.ensure(tb + 256));
        cub::DeviceSelect::Flagged(e->cubtmp2.p, tb, cub::CountingInputIterator<int>(0), d_flag, d_sel, d_flags + 2, (int)n, st);
    }
    int hflags[4];
    CK(cudaMemcpyAsync(hflags, d_flags, 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (hflags[1] != 0x7fffffff) return e->fail(CLB_ERR_ARG, "particle type %d out of range [0,%d)", type[hflags[1]], CLB_MAX_TYPES);
    e->ntypes = std::max(e->ntypes, hflags[0] + 1);
    if (multi) {
        nloc = hflags[2];
        const ClbGrid& g = e->grid;
        double frac = (double)(g.nczl + 2) / g.ncz;
        int64_t cap = std::min<int64_t>(n, (int64_t)(n * frac * 1.5) + 16384);
        TRY(e->alloc_particles((int)std::max<int64_t>(cap, nloc)));
        TRY(e->comm_peer_setup());
        CK(cudaMemsetAsync(e->id2idx.p, 0xff, N * sizeof(int), st));
        if (nloc) k_take_owned<<<ceil_div(nloc, 256), 256, 0, st>>>((int)nloc, d_sel, d_P, d_V, e->pos.p, e->vel.p, e->slot.p, e->id2idx.p);
    } else {
        k_identity_slots<<<ceil_div(n, 256), 256, 0, st>>>((int)n, e->slot.p, e->id2idx.p);
    }
    CK(cudaMemsetAsync(e->force.p, 0, 3 * (size_t)e->ncap * sizeof(double), st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    tr.mark("ingest");
    e->nstored = (int)nloc; e->own0 = 0; e->own1 = (int)nloc;
    e->lists_valid = false; e->forces_valid = false; e->cont_ok = false; e->excl_dirty = true; e->terms_dirty = true; e->topo_dirty = true;
    return CLB_OK;
}
extern "C" int64_t clb_num_particles(const clb_engine* e) { return e ? e->n : 0; }

int clb_engine::alloc_particles(int nlocal_cap) {
    clb_engine* e = this;
    ncap = nlocal_cap + 64;
    CK(pos.ensure(ncap)); CK(pos2.ensure(ncap)); CK(vel.ensure(ncap)); CK(vel2.ensure(ncap));
    CK(slot.ensure(ncap)); CK(slot2.ensure(ncap)); CK(xref.ensure(ncap));
    CK(force.ensure(3 * (size_t)ncap));
    CK(id2idx.ensure(n)); CK(image.ensure(3 * (size_t)n)); CK(resid.ensure(n)); CK(charge.ensure(n)); CK(mol.ensure(n)); CK(wslot.ensure(n));
    CK(key.ensure(ncap)); CK(key2.ensure(ncap)); CK(val.ensure(ncap)); CK(val2.ensure(ncap));
    CK(cell_start.ensure((size_t)grid.ncell + 2));
    CK(nl_count.ensure(ncap)); CK(nl_perm.ensure(ncap));
    CK(partial.ensure(65536)); CK(partial_u64.ensure(65536));
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, key.p, key2.p, val.p, val2.p, ncap, 0, 32, stream);
    size_t tb2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb2, val.p, val2.p, ncap, stream);
    CK(cubtmp.ensure(std::max(tb, tb2) + 256));
    return CLB_OK;
}

// query slots of a get/modify call: NULL ids -> all particles in ascending-id order
static int query_slots(clb_engine* e, int64_t n, const int64_t* ids, std::vector<int>& qs) {
    if (!ids) return CLB_OK;
    qs.resize(n);
    for (int64_t k = 0; k < n; ++k) {
        int s = e->slot_of(ids[k]);
        if (s < 0) return e->fail(CLB_ERR_ARG, "unknown particle id %lld", (long long)ids[k]);
        qs[k] = s;
    }
    return CLB_OK;
}
struct ClbExportArgs {
    double *pos, *vel, *force, *mass, *q; int *image, *type, *state, *resid;
    double qd[3];
};
__global__ void k_export(int cnt, const int* __restrict__ qslot, const int* __restrict__ id2idx, const int4* __restrict__ pos,
                         const ClbVel* __restrict__ vel, const double* __restrict__ force, int fstride, const int* __restrict__ image,
                         const double* __restrict__ charge, const int* __restrict__ resid, ClbExportArgs O, int* __restrict__ missing) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cnt) return;
    const int s = qslot ? qslot[k] : k;
    const int i = id2idx[s];
    if (i < 0) { atomicAdd(missing, 1); return; }
    const size_t k3 = 3 * (size_t)k;
    if (O.pos || O.type || O.state) {
        const int4 p = pos[i];
        if (O.pos) { O.pos[k3] = (double)(unsigned)p.x * O.qd[0]; O.pos[k3 + 1] = (double)(unsigned)p.y * O.qd[1]; O.pos[k3 + 2] = (double)(unsigned)p.z * O.qd[2]; }
        if (O.type) O.type[k] = pw_type(p.w);
        if (O.state) O.state[k] = pw_state(p.w);
    }
    if (O.vel || O.mass) {
        const ClbVel v = vel[i];
        if (O.vel) { O.vel[k3] = v.x; O.vel[k3 + 1] = v.y; O.vel[k3 + 2] = v.z; }
        if (O.mass) O.mass[k] = v.w;
    }
    if (O.force) { O.force[k3] = force[i]; O.force[k3 + 1] = force[i + fstride]; O.force[k3 + 2] = force[i + 2 * (size_t)fstride]; }
    if (O.image) { O.image[k3] = image[3 * (size_t)s]; O.image[k3 + 1] = image[3 * (size_t)s + 1]; O.image[k3 + 2] = image[3 * (size_t)s + 2]; }
    if (O.q) O.q[k] = charge[s];
    if (O.resid) O.resid[k] = resid[s];
}

extern "C" int clb_get_particles(clb_engine* e, int64_t n, const int64_t* ids, double* pos, int32_t* image, double* vel,
                                 double* force, int32_t* type, int32_t* state, double* mass, double* q, int32_t* res_id) {
    if (!e) return CLB_ERR_ARG;
    cudaSetDevice(e->device);
    if (e->nranks > 1) return e->get_particles_gathered(n, ids, pos, image, vel, force, type, state, mass, q, res_id);
    const int64_t cnt = ids ? n : e->n;
    if (cnt <= 0) return CLB_OK;
    std::vector<int> qs;
    TRY(query_slots(e, cnt, ids, qs));
    const size_t N = (size_t)cnt;
    CK(stage_reserve(e, N * (24 * 3 + 8 * 2 + 12 + 4 * 4) + 16 * 256));
    StageCursor sc{e->stage.p};
    ClbExportArgs O;
    O.pos = pos ? sc.take<double>(3 * N) : nullptr; O.vel = vel ? sc.take<double>(3 * N) : nullptr; O.force = force ? sc.take<double>(3 * N) : nullptr;
    O.mass = mass ? sc.take<double>(N) : nullptr; O.q = q ? sc.take<double>(N) : nullptr; O.image = image ? sc.take<int>(3 * N) : nullptr;
    O.type = type ? sc.take<int>(N) : nullptr; O.state = state ? sc.take<int>(N) : nullptr; O.resid = res_id ? sc.take<int>(N) : nullptr;
    int* d_qs = ids ? sc.take<int>(N) : nullptr; int* d_missing = sc.take<int>(1);
    for (int d = 0; d < 3; ++d) O.qd[d] = e->geo.q[d];
    cudaStream_t st = e->stream;
    if (ids) CK(cudaMemcpyAsync(d_qs, qs.data(), 4 * N, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(d_missing, 0, 4, st));
    k_export<<<ceil_div(cnt, 256), 256, 0, st>>>((int)cnt, d_qs, e->id2idx.p, e->pos.p, e->vel.p, e->force.p, e->ncap, e->image.p, e->charge.p, e->resid.p, O, d_missing);
    if (pos) CK(cudaMemcpyAsync(pos, O.pos, 24 * N, cudaMemcpyDeviceToHost, st));
    if (vel) CK(cudaMemcpyAsync(vel, O.vel, 24 * N, cudaMemcpyDeviceToHost, st));
    if (force) CK(cudaMemcpyAsync(force, O.force, 24 * N, cudaMemcpyDeviceToHost, st));
    if (mass) CK(cudaMemcpyAsync(mass, O.mass, 8 * N, cudaMemcpyDeviceToHost, st));
    if (q) CK(cudaMemcpyAsync(q, O.q, 8 * N, cudaMemcpyDeviceToHost, st));
    if (image) CK(cudaMemcpyAsync(image, O.image, 12 * N, cudaMemcpyDeviceToHost, st));
    if (type) CK(cudaMemcpyAsync(type, O.type, 4 * N, cudaMemcpyDeviceToHost, st));
    if (state) CK(cudaMemcpyAsync(state, O.state, 4 * N, cudaMemcpyDeviceToHost, st));
    if (res_id) CK(cudaMemcpyAsync(res_id, O.resid, 4 * N, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    return CLB_OK;
}

// multi-rank read-back: every rank fills the rows of the particles it owns, the rows are summed over the ranks
// (exact: one non-zero contribution per row), so all ranks return the full, identical state
__global__ void k_pack_state(int no, int K, const int4* __restrict__ pos, const ClbVel* __restrict__ vel, const double* __restrict__ force,
                             int fstride, const int* __restrict__ slot, const int* __restrict__ image, double* __restrict__ M) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= no) return;
    const int s = slot[i];
    const int4 p = pos[i]; const ClbVel v = vel[i];
    double* r = M + (size_t)s * K;
    r[0] = (double)(unsigned)p.x; r[1] = (double)(unsigned)p.y; r[2] = (double)(unsigned)p.z;
    r[3] = v.x; r[4] = v.y; r[5] = v.z;
    r[6] = force[i]; r[7] = force[i + fstride]; r[8] = force[i + 2 * (size_t)fstride];
    r[9] = image[3 * s]; r[10] = image[3 * s + 1]; r[11] = image[3 * s + 2];
    r[12] = pw_type(p.w); r[13] = pw_state(p.w); r[14] = v.w;
}
__global__ void k_export_matrix(int cnt, const int* __restrict__ qslot, int K, const double* __restrict__ M, const double* __restrict__ charge,
                                const int* __restrict__ resid, ClbExportArgs O) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cnt) return;
    const int s = qslot ? qslot[k] : k;
    const double* r = M + (size_t)s * K;
    const size_t k3 = 3 * (size_t)k;
    if (O.pos) { O.pos[k3] = r[0] * O.qd[0]; O.pos[k3 + 1] = r[1] * O.qd[1]; O.pos[k3 + 2] = r[2] * O.qd[2]; }
    if (O.vel) { O.vel[k3] = r[3]; O.vel[k3 + 1] = r[4]; O.vel[k3 + 2] = r[5]; }
    if (O.force) { O.force[k3] = r[6]; O.force[k3 + 1] = r[7]; O.force[k3 + 2] = r[8]; }
    if (O.image) { O.image[k3] = (int)r[9]; O.image[k3 + 1] = (int)r[10]; O.image[k3 + 2] = (int)r[11]; }
    if (O.type) O.type[k] = (int)r[12];
    if (O.state) O.state[k] = (int)r[13];
    if (O.mass) O.mass[k] = r[14];
    if (O.q) O.q[k] = charge[s];
    if (O.resid) O.resid[k] = resid[s];
}
int clb_engine::get_particles_gathered(int64_t nq, const int64_t* ids, double* pos_o, int32_t* image_o, double* vel_o, double* force_o,
                                       int32_t* type_o, int32_t* state_o, double* mass_o, double* q_o, int32_t* res_o) {
    clb_engine* e = this;
    const int K = 15;
    const int no = own1;
    const int64_t cnt = ids ? nq : n;
    if (cnt <= 0) return CLB_OK;
    std::vector<int> qs;
    TRY(query_slots(e, cnt, ids, qs));
    // owned rows are packed on the device into a zeroed [n x K] matrix and summed over the ranks (exact: one non-zero
    // contribution per row); the requested fields are then unpacked on the device and copied straight into the caller's arrays
    const size_t N = (size_t)cnt;
    CK(stage_reserve(e, (size_t)n * K * 8 + N * (24 * 3 + 8 * 2 + 12 + 4 * 4) + 16 * 256));
    StageCursor sc{stage.p};
    double* M = sc.take<double>((size_t)n * K);
    ClbExportArgs O;
    O.pos = pos_o ? sc.take<double>(3 * N) : nullptr; O.vel = vel_o ? sc.take<double>(3 * N) : nullptr; O.force = force_o ? sc.take<double>(3 * N) : nullptr;
    O.mass = mass_o ? sc.take<double>(N) : nullptr; O.q = q_o ? sc.take<double>(N) : nullptr; O.image = image_o ? sc.take<int>(3 * N) : nullptr;
    O.type = type_o ? sc.take<int>(N) : nullptr; O.state = state_o ? sc.take<int>(N) : nullptr; O.resid = res_o ? sc.take<int>(N) : nullptr;
    int* d_qs = ids ? sc.take<int>(N) : nullptr;
    for (int d = 0; d < 3; ++d) O.qd[d] = geo.q[d];
    CK(cudaMemsetAsync(M, 0, (size_t)n * K * sizeof(double), stream));
    if (no) k_pack_state<<<ceil_div(no, 256), 256, 0, stream>>>(no, K, this->pos.p, this->vel.p, this->force.p, ncap, slot.p, image.p, M);
    TRY(comm_allreduce_sum_dev(M, (size_t)n * K));
    if (ids) CK(cudaMemcpyAsync(d_qs, qs.data(), 4 * N, cudaMemcpyHostToDevice, stream));
    k_export_matrix<<<ceil_div(cnt, 256), 256, 0, stream>>>((int)cnt, d_qs, K, M, charge.p, resid.p, O);
    if (pos_o) CK(cudaMemcpyAsync(pos_o, O.pos, 24 * N, cudaMemcpyDeviceToHost, stream));
    if (vel_o) CK(cudaMemcpyAsync(vel_o, O.vel, 24 * N, cudaMemcpyDeviceToHost, stream));
    if (force_o) CK(cudaMemcpyAsync(force_o, O.force, 24 * N, cudaMemcpyDeviceToHost, stream));
    if (mass_o) CK(cudaMemcpyAsync(mass_o, O.mass, 8 * N, cudaMemcpyDeviceToHost, stream));
    if (q_o) CK(cudaMemcpyAsync(q_o, O.q, 8 * N, cudaMemcpyDeviceToHost, stream));
    if (image_o) CK(cudaMemcpyAsync(image_o, O.image, 12 * N, cudaMemcpyDeviceToHost, stream));
    if (type_o) CK(cudaMemcpyAsync(type_o, O.type, 4 * N, cudaMemcpyDeviceToHost, stream));
    if (state_o) CK(cudaMemcpyAsync(state_o, O.state, 4 * N, cudaMemcpyDeviceToHost, stream));
    if (res_o) CK(cudaMemcpyAsync(res_o, O.resid, 4 * N, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    CK(cudaGetLastError());
    return CLB_OK;
}

static inline int lattice_of(double x, double L, int* im) {
    double fr = x / L, fl = floor(fr), u = rint((fr - fl) * 4294967296.0);
    int i = (int)fl;
    if (u >= 4294967296.0) { u -= 4294967296.0; i += 1; }
    if (im) *im = i;
    return (int)(uint32_t)(uint64_t)u;
}

extern "C" int clb_modify_particle(clb_engine* e, int64_t id, int field, const double* value) {
    if (!e || !value) return CLB_ERR_ARG;
    cudaSetDevice(e->device);
    int s = e->slot_of(id);
    if (s < 0) return e->fail(CLB_ERR_ARG, "unknown particle id %lld", (long long)id);
    int i;
    CK(cudaMemcpy(&i, e->id2idx.p + s, 4, cudaMemcpyDeviceToHost));
    if (e->nranks > 1 && (field == 5)) return e->fail(CLB_ERR_UNSUPPORTED, "moving a particle on a multi-rank engine is not supported");
    // replicated per-slot properties first (all ranks make the same call), then the local copy if the particle is stored here
    int w;
    CK(cudaMemcpy(&w, e->wslot.p + s, 4, cudaMemcpyDeviceToHost));
    switch (field) {
        case 0: { int t = (int)value[0]; if (t < 0 || t >= CLB_MAX_TYPES) return e->fail(CLB_ERR_ARG, "type out of range");
                  w = pw_pack(t, pw_state(w)); e->ntypes = std::max(e->ntypes, t + 1); e->pots_dirty = true; break; }
        case 1: w = pw_pack(pw_type(w), (int)value[0]); break;
        case 3: { double qq = value[0]; CK(cudaMemcpy(e->charge.p + s, &qq, 8, cudaMemcpyHostToDevice)); break; }
        case 4: { int r = (int)value[0]; CK(cudaMemcpy(e->resid.p + s, &r, 4, cudaMemcpyHostToDevice)); break; }
        case 2: case 5: case 6: break;
        default: return e->fail(CLB_ERR_ARG, "unknown field %d", field);
    }
    CK(cudaMemcpy(e->wslot.p + s, &w, 4, cudaMemcpyHostToDevice));
    e->forces_valid = false;
    if (i < 0) return CLB_OK;
    int4 p; ClbVel v;
    CK(cudaMemcpy(&p, e->pos.p + i, sizeof(p), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&v, e->vel.p + i, sizeof(v), cudaMemcpyDeviceToHost));
    p.w = w;
    switch (field) {
        case 2: v.w = value[0]; break;
        case 5: { int im[3]; p.x = lattice_of(value[0], e->box[0], &im[0]); p.y = lattice_of(value[1], e->box[1], &im[1]); p.z = lattice_of(value[2], e->box[2], &im[2]);
                  CK(cudaMemcpy(e->image.p + 3 * s, im, 12, cudaMemcpyHostToDevice)); e->lists_valid = false; e->cont_ok = false; break; }
        case 6: v.x = value[0]; v.y = value[1]; v.z = value[2]; break;
        default: break;
    }
    CK(cudaMemcpy(e->pos.p + i, &p, sizeof(p), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->vel.p + i, &v, sizeof(v), cudaMemcpyHostToDevice));
    return CLB_OK;
}
__global__ void k_set_velocities(int n, const int* __restrict__ id2idx, const double* __restrict__ v, ClbVel* __restrict__ vel) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int i = id2idx[s];
    if (i < 0) return;
    ClbVel o = vel[i];
    o.x = v[3 * (size_t)s]; o.y = v[3 * (size_t)s + 1]; o.z = v[3 * (size_t)s + 2];
    vel[i] = o;
}
__global__ void k_set_positions(int n, const int* __restrict__ id2idx, const double* __restrict__ x, double bx, double by, double bz,
                                int4* __restrict__ pos, int* __restrict__ image) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int i = id2idx[s];
    if (i < 0) return;
    const double box[3] = {bx, by, bz};
    int xi[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double fr = x[3 * (size_t)s + d] / box[d], fl = floor(fr), u = rint((fr - fl) * 4294967296.0);
        int im = (int)fl;
        if (u >= 4294967296.0) { u -= 4294967296.0; im += 1; }
        xi[d] = (int)(unsigned)(unsigned long long)u;
        image[3 * (size_t)s + d] = im;
    }
    int4 p = pos[i];
    pos[i] = make_int4(xi[0], xi[1], xi[2], p.w);
}
extern "C" int clb_set_velocities(clb_engine* e, int64_t n, const double* vel) {
    if (!e || n != e->n || !vel) return e ? e->fail(CLB_ERR_ARG, "clb_set_velocities: n mismatch") : CLB_ERR_ARG;
    cudaSetDevice(e->device);
    CK(stage_reserve(e, 24 * (size_t)n));
    CK(cudaMemcpyAsync(e->stage.p, vel, 24 * (size_t)n, cudaMemcpyHostToDevice, e->stream));
    k_set_velocities<<<ceil_div(n, 256), 256, 0, e->stream>>>((int)n, e->id2idx.p, (const double*)e->stage.p, e->vel.p);
    CK(cudaStreamSynchronize(e->stream));
    return CLB_OK;
}
extern "C" int clb_set_positions(clb_engine* e, int64_t n, const double* pos) {
    if (!e || n != e->n || !pos) return e ? e->fail(CLB_ERR_ARG, "clb_set_positions: n mismatch") : CLB_ERR_ARG;
    if (e->nranks > 1) return e->fail(CLB_ERR_UNSUPPORTED, "clb_set_positions on a multi-rank engine: call clb_set_particles again");
    cudaSetDevice(e->device);
    CK(stage_reserve(e, 24 * (size_t)n));
    CK(cudaMemcpyAsync(e->stage.p, pos, 24 * (size_t)n, cudaMemcpyHostToDevice, e->stream));
    k_set_positions<<<ceil_div(n, 256), 256, 0, e->stream>>>((int)n, e->id2idx.p, (const double*)e->stage.p, e->box[0], e->box[1], e->box[2], e->pos.p, e->image.p);
    CK(cudaStreamSynchronize(e->stream));
    e->lists_valid = false; e->forces_valid = false; e->cont_ok = false;
    return CLB_OK;
}

// ------------------------------------------------------------------------------------------ exclusions
// caller ids -> slots on the device for DENSE ids (slot = id - first id); flags[0] counts ids out of range
__global__ void k_ids_to_slots(long long m, const long long* __restrict__ ids, long long base, int n, int* __restrict__ out, int* __restrict__ flags) {
    long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= m) return;
    const long long s = ids[k] - base;
    if (s < 0 || s >= n) { atomicAdd(flags, 1); out[k] = 0; return; }
    out[k] = (int)s;
}
__global__ void k_excl_keys(long long m, const int* __restrict__ slots, unsigned long long* __restrict__ keys) {
    long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= m) return;
    const int a = slots[2 * k], b = slots[2 * k + 1];
    // self pairs sort to the end and are cut off
    keys[k] = a == b ? ~0ull : (((unsigned long long)(unsigned)min(a, b) << 32) | (unsigned)max(a, b));
}
__global__ void k_excl_unpack(long long m, const unsigned long long* __restrict__ keys, int2* __restrict__ pairs, int* __restrict__ nvalid) {
    long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= m) return;
    const unsigned long long v = keys[k];
    if (v == ~0ull) return;
    pairs[k] = make_int2((int)(v >> 32), (int)(v & 0xffffffffu));
    if (k + 1 == m || keys[k + 1] == ~0ull) *nvalid = (int)(k + 1);
}
// slots of a host id array, on the device (dense ids) or through the host map (arbitrary ids); result in d_out[0..m)
static int slots_to_device(clb_engine* e, long long m, const int64_t* ids, int* d_out, const char* what) {
    if (m == 0) return CLB_OK;
    if (e->ids_dense) {
        CK(stage_reserve(e, (size_t)m * 8 + 512));
        long long* d_ids = reinterpret_cast<long long*>(e->stage.p);
        int* d_flag = reinterpret_cast<int*>(e->stage.p + (((size_t)m * 8 + 255) & ~(size_t)255));
        CK(cudaMemcpyAsync(d_ids, ids, (size_t)m * 8, cudaMemcpyHostToDevice, e->stream));
        CK(cudaMemsetAsync(d_flag, 0, 4, e->stream));
        k_ids_to_slots<<<ceil_div(m, 256), 256, 0, e->stream>>>(m, d_ids, (long long)e->id_base, e->n, d_out, d_flag);
        int bad = 0;
        CK(cudaMemcpyAsync(&bad, d_flag, 4, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        if (bad) return e->fail(CLB_ERR_ARG, "%s names %d unknown particle id(s)", what, bad);
        return CLB_OK;
    }
    std::vector<int> h((size_t)m);
    for (long long k = 0; k < m; ++k) { int s = e->slot_of(ids[k]); if (s < 0) return e->fail(CLB_ERR_ARG, "%s names unknown particle id %lld", what, (long long)ids[k]); h[k] = s; }
    CK(cudaMemcpy(d_out, h.data(), (size_t)m * 4, cudaMemcpyHostToDevice));
    return CLB_OK;
}

extern "C" int clb_set_exclusions(clb_engine* e, int64_t n, const int64_t* pairs) {
    if (!e || n < 0 || (n && !pairs)) return CLB_ERR_ARG;
    cudaSetDevice(e->device);
    // canonical (min slot, max slot) pairs as 64-bit keys, sorted and made unique on the device
    CK(e->excl_pairs.ensure((size_t)n * 2 + (size_t)e->n + 1024));
    e->nexcl = 0;
    if (n > 0) {
        DevBuf<int> dsl; DevBuf<unsigned long long> k1, k2; DevBuf<int> dcount;
        CK(dsl.ensure(2 * (size_t)n)); CK(k1.ensure(n)); CK(k2.ensure(n)); CK(dcount.ensure(4));
        int r = slots_to_device(e, 2 * (long long)n, pairs, dsl.p, "exclusion list");
        if (r != CLB_OK) { dsl.release(); k1.release(); k2.release(); dcount.release(); return r; }
        k_excl_keys<<<ceil_div(n, 256), 256, 0, e->stream>>>(n, dsl.p, k1.p);
        size_t tb = 0;
        cub::DeviceRadixSort::SortKeys(nullptr, tb, k1.p, k2.p, (int)n, 0, 64, e->stream);
        CK(e->cubtmp2.ensure(tb + 256));
        cub::DeviceRadixSort::SortKeys(e->cubtmp2.p, tb, k1.p, k2.p, (int)n, 0, 64, e->stream);
        size_t tb2 = 0;
        cub::DeviceSelect::Unique(nullptr, tb2, k2.p, k1.p, dcount.p, (int)n, e->stream);
        CK(e->cubtmp2.ensure(tb2 + 256));
        cub::DeviceSelect::Unique(e->cubtmp2.p, tb2, k2.p, k1.p, dcount.p, (int)n, e->stream);
        int nu = 0;
        CK(cudaMemcpyAsync(&nu, dcount.p, 4, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaMemsetAsync(dcount.p + 1, 0, 4, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        if (nu > 0) k_excl_unpack<<<ceil_div(nu, 256), 256, 0, e->stream>>>(nu, k1.p, e->excl_pairs.p, dcount.p + 1);
        int nvalid = 0;
        CK(cudaMemcpyAsync(&nvalid, dcount.p + 1, 4, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        e->nexcl = nvalid;
        dsl.release(); k1.release(); k2.release(); dcount.release();
    }
    e->excl_dirty = true; e->lists_valid = false; e->cont_ok = false;
    return CLB_OK;
}
extern "C" int64_t clb_num_exclusions(const clb_engine* e) { return e ? e->nexcl : 0; }
extern "C" int clb_get_exclusions(clb_engine* e, int64_t cap, int64_t* pairs, int64_t* n_out) {
    if (!e) return CLB_ERR_ARG;
    cudaSetDevice(e->device);
    std::vector<int2> h(e->nexcl);
    if (e->nexcl) CK(cudaMemcpy(h.data(), e->excl_pairs.p, h.size() * sizeof(int2), cudaMemcpyDeviceToHost));
    std::sort(h.begin(), h.end(), [](const int2& x, const int2& y) { return x.x != y.x ? x.x < y.x : x.y < y.y; });
    h.erase(std::unique(h.begin(), h.end(), [](const int2& x, const int2& y) { return x.x == y.x && x.y == y.y; }), h.end());
    if (n_out) *n_out = (int64_t)h.size();
    for (size_t k = 0; k < h.size() && (int64_t)k < cap; ++k) { pairs[2 * k] = e->ids[h[k].x]; pairs[2 * k + 1] = e->ids[h[k].y]; }
    return CLB_OK;
}
extern "C" int clb_exclusions_observe(clb_engine* e, int list) {
    if (!e || list < 0 || list >= (int)e->lists.size()) return e ? e->fail(CLB_ERR_ARG, "bad list handle") : CLB_ERR_ARG;
    e->lists[list].excl_observed = 1; e->react_dirty = true;
    return CLB_OK;
}

__global__ void k_excl_expand(long long n, const int2* __restrict__ pairs, int* __restrict__ key, int* __restrict__ val) {
    long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= n) return;
    int2 p = pairs[k];
    key[2 * k] = p.x; val[2 * k] = p.y; key[2 * k + 1] = p.y; val[2 * k + 1] = p.x;
}
int clb_engine::build_excl_csr() {
    clb_engine* e = this;
    size_t m = (size_t)nexcl * 2;
    size_t mcap = std::max(m, excl_pairs.n * 2) + 16;     // sized for the capacity of the pair buffer: no regrowth per reaction pass
    CK(excl_off.ensure((size_t)n + 2));
    CK(excl_ids.ensure(mcap));
    if (m == 0) { CK(cudaMemsetAsync(excl_off.p, 0, ((size_t)n + 2) * 4, stream)); excl_dirty = false; return CLB_OK; }
    CK(ekey.ensure(mcap)); CK(ekey2.ensure(mcap)); CK(eval.ensure(mcap));
    k_excl_expand<<<ceil_div(nexcl, 256), 256, 0, stream>>>(nexcl, excl_pairs.p, ekey.p, eval.p);
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, ekey.p, ekey2.p, eval.p, excl_ids.p, (int)mcap, 0, 32, stream);
    CK(cubtmp2.ensure(tb + 256));
    cub::DeviceRadixSort::SortPairs(cubtmp2.p, tb, ekey.p, ekey2.p, eval.p, excl_ids.p, (int)m, 0, 32, stream);
    k_lower_bounds<<<ceil_div(n + 1, 256), 256, 0, stream>>>((int)m, ekey2.p, n, excl_off.p);
    CK(cudaGetLastError());
    excl_dirty = false;
    return CLB_OK;
}

// ------------------------------------------------------------------------------------------ tables & potentials
static void akima_coeffs(int n, double h, const double* y, std::vector<double>& c) {
    std::vector<double> m(n + 3);
    double* mm = m.data() + 2;
    for (int i = 0; i < n - 1; ++i) mm[i] = (y[i + 1] - y[i]) / h;
    mm[-1] = 2 * mm[0] - mm[1]; mm[-2] = 2 * mm[-1] - mm[0];
    mm[n - 1] = 2 * mm[n - 2] - mm[n - 3]; mm[n] = 2 * mm[n - 1] - mm[n - 2];
    std::vector<double> t(n);
    for (int i = 0; i < n; ++i) {
        double w1 = fabs(mm[i + 1] - mm[i]), w2 = fabs(mm[i - 1] - mm[i - 2]);
        t[i] = (w1 + w2 == 0.0) ? 0.5 * (mm[i - 1] + mm[i]) : (w1 * mm[i - 1] + w2 * mm[i]) / (w1 + w2);
    }
    c.assign(4 * (size_t)(n - 1), 0.0);
    for (int i = 0; i < n - 1; ++i) {
        c[4 * i] = y[i]; c[4 * i + 1] = t[i];
        c[4 * i + 2] = (3 * mm[i] - 2 * t[i] - t[i + 1]) / h;
        c[4 * i + 3] = (t[i] + t[i + 1] - 2 * mm[i]) / (h * h);
    }
}
static void spline_coeffs(int n, double h, const double* y, std::vector<double>& c) {
    std::vector<double> y2(n, 0.0), u(n, 0.0);
    for (int i = 1; i < n - 1; ++i) {
        double p = 0.5 * y2[i - 1] + 2.0;
        y2[i] = -0.5 / p;
        u[i] = (y[i + 1] - 2 * y[i] + y[i - 1]) / h;
        u[i] = (3.0 * u[i] / h - 0.5 * u[i - 1]) / p;
    }
    y2[n - 1] = 0;
    for (int k = n - 2; k >= 0; --k) y2[k] = y2[k] * y2[k + 1] + u[k];
    c.assign(4 * (size_t)(n - 1), 0.0);
    for (int i = 0; i < n - 1; ++i) {
        c[4 * i] = y[i];
        c[4 * i + 1] = (y[i + 1] - y[i]) / h - h * (2 * y2[i] + y2[i + 1]) / 6.0;
        c[4 * i + 2] = 0.5 * y2[i];
        c[4 * i + 3] = (y2[i + 1] - y2[i]) / (6.0 * h);
    }
}
static void linear_coeffs(int n, double h, const double* y, std::vector<double>& c) {
    c.assign(4 * (size_t)(n - 1), 0.0);
    for (int i = 0; i < n - 1; ++i) { c[4 * i] = y[i]; c[4 * i + 1] = (y[i + 1] - y[i]) / h; }
}

extern "C" int clb_add_table(clb_engine* e, int64_t n, const double* x, const double* energy, const double* force, int interp,
                             int* table_out) {
    if (!e || n < 2 || !x || !energy || !force || !table_out) return e ? e->fail(CLB_ERR_ARG, "clb_add_table: bad argument") : CLB_ERR_ARG;
    if (interp < 1 || interp > 3) return e->fail(CLB_ERR_ARG, "interp must be 1 (linear), 2 (Akima) or 3 (cubic)");
    HostTable t;
    t.n = (int)n; t.x0 = x[0]; t.dx = (x[n - 1] - x[0]) / (double)(n - 1); t.interp = interp;
    if (!(t.dx > 0)) return e->fail(CLB_ERR_ARG, "table abscissa must increase");
    for (int64_t i = 0; i < n; ++i)
        // `.pot` files carry 8 significant digits ("%15.8g", tools/convert_gromacs2espp.py:84): angle/dihedral abscissae in radians
        // are off the exact grid by up to 1e-7; the knots used are x0 + i*dx (U12), as in the oracle
        if (fabs(x[i] - (t.x0 + i * t.dx)) > 1e-3 * t.dx + 1e-7 * fabs(x[i]) + 5e-9)
            return e->fail(CLB_ERR_ARG, "table abscissa is not uniform at row %lld", (long long)i);
    t.e.assign(energy, energy + n); t.f.assign(force, force + n);
    e->tables.push_back(std::move(t));
    *table_out = (int)e->tables.size() - 1;
    e->pots_dirty = true;
    return CLB_OK;
}

extern "C" int clb_add_nonbonded(clb_engine* e, int kind, int* out) {
    if (!e || !out || kind < 1 || kind > 3) return e ? e->fail(CLB_ERR_ARG, "bad non-bonded kind") : CLB_ERR_ARG;
    HostInter it; it.kind = kind; it.bonded = -1;
    e->inters.push_back(it);
    *out = (int)e->inters.size() - 1;
    return CLB_OK;
}
static int check_nb(clb_engine* e, int inter, int t1, int t2, int kind, double cutoff) {
    if (inter < 0 || inter >= (int)e->inters.size() || e->inters[inter].bonded >= 0) return e->fail(CLB_ERR_ARG, "bad non-bonded interaction handle");
    if (e->inters[inter].kind != kind) return e->fail(CLB_ERR_ARG, "potential kind does not match the interaction kind");
    if (t1 < 0 || t2 < 0 || t1 >= CLB_MAX_TYPES || t2 >= CLB_MAX_TYPES) return e->fail(CLB_ERR_ARG, "type out of range");
    if (cutoff > e->rc * (1 + 1e-12)) return e->fail(CLB_ERR_ARG, "potential cutoff %g exceeds the Verlet cutoff %g", cutoff, e->rc);
    HostPairPot& p = e->pp[t1][t2];
    if (p.kind != 0 && p.inter != inter) return e->fail(CLB_ERR_UNSUPPORTED, "type pair (%d,%d) already belongs to interaction %d", t1, t2, p.inter);
    e->ntypes = std::max(e->ntypes, std::max(t1, t2) + 1);
    return CLB_OK;
}
extern "C" int clb_nb_set_tabulated(clb_engine* e, int inter, int t1, int t2, int table, double cutoff) {
    if (!e) return CLB_ERR_ARG;
    TRY(check_nb(e, inter, t1, t2, CLB_NB_TABULATED, cutoff));
    if (table < 0 || table >= (int)e->tables.size()) return e->fail(CLB_ERR_ARG, "bad table handle");
    if (e->tables[table].interp != 1) return e->fail(CLB_ERR_UNSUPPORTED, "non-bonded tables use linear interpolation (itype=1) as chemlab does (gromacs_topology.py:696-707)");
    HostPairPot p; p.kind = 1; p.inter = inter; p.tab1 = table; p.tab2 = -1; p.rc = cutoff;
    e->pp[t1][t2] = e->pp[t2][t1] = p;
    e->pots_dirty = true; e->forces_valid = false; e->cont_ok = false;
    return CLB_OK;
}
extern "C" int clb_nb_set_lj(clb_engine* e, int inter, int t1, int t2, double eps, double sig, double cutoff, int shift_auto) {
    if (!e) return CLB_ERR_ARG;
    TRY(check_nb(e, inter, t1, t2, CLB_NB_LENNARD_JONES, cutoff));
    HostPairPot p; p.kind = 2; p.inter = inter; p.eps = eps; p.sig = sig; p.rc = cutoff;
    double sr6 = pow(sig / cutoff, 6);
    p.shift = shift_auto ? 4 * eps * (sr6 * sr6 - sr6) : 0.0;
    e->pp[t1][t2] = e->pp[t2][t1] = p;
    e->pots_dirty = true; e->forces_valid = false; e->cont_ok = false;
    return CLB_OK;
}
extern "C" int clb_nb_set_mixed(clb_engine* e, int inter, int t1, int t2, int table1, int table2, double mix, int conv_type,
                                double conv_total, double cutoff) {
    if (!e) return CLB_ERR_ARG;
    TRY(check_nb(e, inter, t1, t2, CLB_NB_MIXED_TABULATED, cutoff));
    if (table1 < 0 || table1 >= (int)e->tables.size() || table2 < 0 || table2 >= (int)e->tables.size()) return e->fail(CLB_ERR_ARG, "bad table handle");
    const HostTable &a = e->tables[table1], &b = e->tables[table2];
    if (a.n != b.n || fabs(a.x0 - b.x0) > 1e-12 || fabs(a.dx - b.dx) > 1e-12 || a.interp != 1 || b.interp != 1)
        return e->fail(CLB_ERR_UNSUPPORTED, "mixed tables must share one grid and use linear interpolation");
    HostPairPot p; p.kind = 3; p.inter = inter; p.tab1 = table1; p.tab2 = table2; p.mix = mix; p.conv_type = conv_type; p.conv_total = conv_total; p.rc = cutoff;
    e->pp[t1][t2] = e->pp[t2][t1] = p;
    e->pots_dirty = true; e->forces_valid = false; e->cont_ok = false; e->has_mixed = true;
    return CLB_OK;
}

// Upload pair descriptors and tables.  Every type pair with a table gets its own row slot so that
// mixed tables (x*tab1 + (1-x)*tab2, linear in the rows) are premixed on the host; identical
// (table, mix) combinations share one slot.
int clb_engine::upload_potentials() {
    clb_engine* e = this;
    int nt = std::max(ntypes, 1);
    std::vector<ClbPairDesc> pd((size_t)nt * nt); std::vector<ClbPairDescE> pe((size_t)nt * nt); std::vector<double2> plj((size_t)nt * nt);
    std::vector<ClbTabMeta> tm; std::vector<double2> frows, erows;
    struct Key { int a, b; double mix; };
    std::vector<Key> keys;
    for (int a = 0; a < nt; ++a) for (int b = 0; b < nt; ++b) {
        const HostPairPot& p = pp[a][b];
        ClbPairDesc d; memset(&d, 0, sizeof(d)); ClbPairDescE de; memset(&de, 0, sizeof(de));
        d.rc2 = -1.0;
        de.inter = p.inter;
        plj[(size_t)a * nt + b] = make_double2(0.0, 0.0);
        if (p.kind == 2) {
            d.kind = 2; d.rc2 = p.rc * p.rc;
            double s6 = pow(p.sig, 6), s12 = s6 * s6;
            plj[(size_t)a * nt + b] = make_double2(48 * p.eps * s12, 24 * p.eps * s6);
            de.e12 = 4 * p.eps * s12; de.e6 = 4 * p.eps * s6; de.shift = p.shift;
        } else if (p.kind == 1 || p.kind == 3) {
            d.kind = 1; d.rc2 = p.rc * p.rc;
            double mix = p.kind == 3 ? p.mix : 1.0;
            int t2 = p.kind == 3 ? p.tab2 : -1;
            int slot_i = -1;
            for (size_t k = 0; k < keys.size(); ++k) if (keys[k].a == p.tab1 && keys[k].b == t2 && keys[k].mix == mix) slot_i = (int)k;
            if (slot_i < 0) {
                const HostTable& A = tables[p.tab1];
                const HostTable* B = t2 >= 0 ? &tables[t2] : nullptr;
                ClbTabMeta m; m.x0 = A.x0; m.dx = A.dx; m.invdx = 1.0 / A.dx; m.c_t = -A.x0 / A.dx - 0.5; m.c_idx = 0; m.n = A.n; m.off = (int)frows.size();
                if (A.x0 + A.dx * (A.n - 1) < p.rc * (1 - 1e-12)) return fail(CLB_ERR_ARG, "table for types (%d,%d) ends at %g before the cutoff %g", a, b, A.x0 + A.dx * (A.n - 1), p.rc);
                for (int i = 0; i < A.n; ++i) {
                    auto F = [&](int k) { return B ? mix * A.f[k] + (1 - mix) * B->f[k] : A.f[k]; };
                    auto E = [&](int k) { return B ? mix * A.e[k] + (1 - mix) * B->e[k] : A.e[k]; };
                    int i1 = std::min(i + 1, A.n - 1);
                    double df = F(i1) - F(i), de_ = E(i1) - E(i);
                    frows.push_back(make_double2(F(i) + 0.5 * df, df));
                    erows.push_back(make_double2(E(i), de_));
                }
                keys.push_back({p.tab1, t2, mix}); tm.push_back(m);
                slot_i = (int)keys.size() - 1;
            }
            d.tab = slot_i;
        }
        pd[(size_t)a * nt + b] = d; pe[(size_t)a * nt + b] = de;
    }
    // uniform-grid fast path: every table on the same (x0, dx, n) -> grid constants travel as kernel
    // arguments and the descriptor carries the first row directly
    all_tab = !tm.empty();
    for (auto& d : pd) if (d.kind == 2) all_tab = 0;
    ugrid_on = !tm.empty();
    for (size_t k = 1; k < tm.size(); ++k) if (tm[k].n != tm[0].n || tm[k].x0 != tm[0].x0 || tm[k].dx != tm[0].dx) ugrid_on = 0;
    if (!tm.empty()) ugrid_meta = tm[0];
    CK(d_tm_e.ensure(std::max<size_t>(tm.size(), 1)));
    if (!tm.empty()) CK(cudaMemcpyAsync(d_tm_e.p, tm.data(), tm.size() * sizeof(ClbTabMeta), cudaMemcpyHostToDevice, stream));   // real units
    CK(d_pd_e.ensure(pd.size()));
    CK(cudaMemcpyAsync(d_pd_e.p, pd.data(), pd.size() * sizeof(ClbPairDesc), cudaMemcpyHostToDevice, stream));   // energy kernel: slot ids
    if (ugrid_on) for (auto& d : pd) if (d.kind == 1) d.tab = tm[d.tab].off;
    if (geo.cubic) {
        // the pair-force kernel of cubic boxes works in lattice units (clb_tile.cuh)
        const double q = geo.q[0];
        for (auto& d : pd) if (d.kind) d.rc2 /= geo.q2;
        for (auto& m : tm) { m.invdx *= q; }
        for (auto& c : plj) { c.x /= pow(q, 13); c.y /= pow(q, 7); }
        if (!tm.empty()) ugrid_meta = tm[0];
    }
    {   // the last row of every table is read when r equals the table end: {f[n-1], 0}
        for (auto& m : tm) { frows[m.off + m.n - 1] = make_double2(frows[m.off + m.n - 1].x, 0.0); }
    }
    // third-generation kernel (k_pair_forces_tab2): rows {A_i, B_i} with F = A_i + r_lat * B_i, descriptors {rc2, first row}
    tab2_ok = 0; tab2_onepd = 0;
    if (geo.cubic && ugrid_on && all_tab) {
        const ClbTabMeta& m0 = tm[0];
        const double k0 = m0.x0 / m0.dx;
        if (fabs(k0 - rint(k0)) < 1e-9 && rint(k0) < 1e6) {
            const double q = geo.q[0];
            std::vector<double2> rows2(frows.size()), pd2(pd.size());
            for (auto& m : tm)
                for (int i = 0; i < m.n; ++i) {
                    const double2 fr = frows[m.off + i];               // {f_i + df_i/2, df_i}
                    const double fi = fr.x - 0.5 * fr.y, sl = fr.y / m.dx, xi = m.x0 + i * m.dx;
                    rows2[m.off + i] = make_double2(fi - xi * sl, q * sl);
                }
            bool one = true;
            for (size_t k = 0; k < pd.size(); ++k) {
                pd2[k] = make_double2(pd[k].kind ? pd[k].rc2 : -1.0, 6755399441055744.0 + (double)pd[k].tab);   // rc2 in lattice^2; first row in the LOW WORD (1.5*2^52 + n)
                one = one && pd[k].kind == 1 && pd[k].rc2 == pd[0].rc2 && pd[k].tab == pd[0].tab;
            }
            CK(d_rows2.ensure(rows2.size())); CK(d_pd2.ensure(pd2.size()));
            CK(cudaMemcpyAsync(d_rows2.p, rows2.data(), rows2.size() * sizeof(double2), cudaMemcpyHostToDevice, stream));
            CK(cudaMemcpyAsync(d_pd2.p, pd2.data(), pd2.size() * sizeof(double2), cudaMemcpyHostToDevice, stream));
            CK(cudaStreamSynchronize(stream));
            tab2_ok = 1; tab2_onepd = one ? 1 : 0;
            tab2_invdx = q / m0.dx; tab2_cmagic = 6755399441055744.0 - rint(k0); tab2_nm1 = (unsigned)m0.n - 1u;
            tab2_one_rc2 = pd[0].rc2; tab2_one_off = pd[0].tab;
        }
    }
    // fourth-generation kernel (k_pair_forces_tab3): cubic box, every pair tabulated, ONE (x0, dx) for all tables (lengths may
    // differ); rows {A_i, B_i} as for tab2, windows chosen in configure_tables()
    tab3_ok = 0;
    {
        // one dx for all tables; first abscissae may differ by whole rows (shipped dacron tables start at r = 0.002 or at r = 0)
        bool same_grid = !tm.empty() && geo.cubic && all_tab;
        double x0min = same_grid ? tm[0].x0 : 0.0;
        for (size_t k = 1; k < tm.size() && same_grid; ++k) { same_grid = tm[k].dx == tm[0].dx; x0min = std::min(x0min, tm[k].x0); }
        for (size_t k = 0; k < tm.size() && same_grid; ++k) { const double sh = (tm[k].x0 - x0min) / tm[0].dx; same_grid = fabs(sh - rint(sh)) < 1e-9 && rint(sh) < 1024; }
        const double k0 = same_grid ? x0min / tm[0].dx : 0.5;
        if (same_grid && fabs(k0 - rint(k0)) < 1e-9 && rint(k0) < 1e6) {
            const double q = geo.q[0];
            t3_rows.assign(frows.size(), make_double2(0.0, 0.0));
            t3_slots.clear();
            for (size_t k = 0; k < tm.size(); ++k) {
                const ClbTabMeta& m = tm[k];
                if (m.n > 65000) { same_grid = false; break; }
                double emin = 1e300;
                for (int i = 0; i < m.n; ++i) {
                    const double2 fr = frows[m.off + i];               // {f_i + df_i/2, df_i}
                    const double fi = fr.x - 0.5 * fr.y, sl = fr.y / m.dx, xi = m.x0 + i * m.dx;
                    t3_rows[m.off + i] = make_double2(fi - xi * sl, q * sl);
                    emin = std::min(emin, erows[m.off + i].x);
                }
                // lower window edge: the first row a pair can reach thermally (U - U_min < 30 kT); without a thermostat: row 0
                int w0 = 0;
                if (lang_on && kT > 0) { while (w0 < m.n - 1 && erows[m.off + w0].x - emin > 30.0 * kT) ++w0; w0 = std::max(0, w0 - 2); }
                clb_engine::T3Slot sl3 = {m.off, m.n, w0, m.n, 0.0, -1, (int)rint((m.x0 - x0min) / m.dx)};
                t3_slots.push_back(sl3);
            }
            if (same_grid) {
                t3_pair_slot.assign(pd.size(), -1); t3_pair_rc2.assign(pd.size(), -1.0);
                std::vector<double> rcmax(tm.size(), 0.0);
                for (int a = 0; a < nt; ++a) for (int b = 0; b < nt; ++b) {
                    const size_t k = (size_t)a * nt + b;
                    if (!pd[k].kind) continue;
                    // pd[k].tab: row offset when ugrid_on, slot index otherwise
                    int slot_i = -1;
                    if (ugrid_on) { for (size_t z = 0; z < tm.size(); ++z) if (tm[z].off == pd[k].tab) slot_i = (int)z; }
                    else slot_i = pd[k].tab;
                    t3_pair_slot[k] = slot_i; t3_pair_rc2[k] = pd[k].rc2;
                    rcmax[slot_i] = std::max(rcmax[slot_i], pp[a][b].rc);
                }
                for (size_t z = 0; z < tm.size(); ++z) {
                    int w1 = (int)floor((rcmax[z] - tm[z].x0) / tm[z].dx) + 2;      // in rows of THIS table
                    t3_slots[z].w1 = std::max(t3_slots[z].w0 + 1, std::min(tm[z].n, w1));
                }
                CK(d_rows2.ensure(t3_rows.size()));
                CK(cudaMemcpyAsync(d_rows2.p, t3_rows.data(), t3_rows.size() * sizeof(double2), cudaMemcpyHostToDevice, stream));
                CK(cudaStreamSynchronize(stream));
                tab2_invdx = q / tm[0].dx; tab2_cmagic = 6755399441055744.0 - rint(k0);
                tab3_ok = 1; t3_dirty = true;
            }
        }
    }
    CK(d_plj.ensure(plj.size()));
    CK(cudaMemcpyAsync(d_plj.p, plj.data(), plj.size() * sizeof(double2), cudaMemcpyHostToDevice, stream));
    nt_dev = nt; ntabs_dev = (int)tm.size(); nrows_dev = (int)frows.size();
    CK(d_pd.ensure(pd.size())); CK(d_pe.ensure(pe.size()));
    CK(d_tm.ensure(std::max<size_t>(tm.size(), 1))); CK(d_frows.ensure(std::max<size_t>(frows.size(), 1))); CK(d_erows.ensure(std::max<size_t>(erows.size(), 1)));
    CK(cudaMemcpyAsync(d_pd.p, pd.data(), pd.size() * sizeof(ClbPairDesc), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(d_pe.p, pe.data(), pe.size() * sizeof(ClbPairDescE), cudaMemcpyHostToDevice, stream));
    if (!tm.empty()) {
        CK(cudaMemcpyAsync(d_tm.p, tm.data(), tm.size() * sizeof(ClbTabMeta), cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(d_frows.p, frows.data(), frows.size() * sizeof(double2), cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(d_erows.p, erows.data(), erows.size() * sizeof(double2), cudaMemcpyHostToDevice, stream));
    }
    // bonded descriptors, potentials and tables (cubic coefficients per interval)
    std::vector<ClbBondedDesc> bd(std::max<size_t>(lists.size(), 1));
    std::vector<ClbBPot> bp; std::vector<ClbBTabMeta> btm; std::vector<double> cf, cec;
    std::vector<int> tabmap(tables.size(), -1);
    for (size_t l = 0; l < lists.size(); ++l) {
        ClbBondedDesc d; memset(&d, 0, sizeof(d));
        d.arity = lists[l].arity; d.inter = -1;
        int bi = lists[l].bonded;
        if (bi >= 0) {
            const HostBonded& hb = bondeds[bi];
            d.typed = hb.typed; d.inter = hb.inter; d.npot = (int)hb.pots.size(); d.pot_off = (int)bp.size(); d.active = d.npot > 0;
            for (ClbBPot p : hb.pots) {
                if (p.kind == 2 || p.kind == 4 || p.kind == 5) {
                    int ht = p.table;
                    if (tabmap[ht] < 0) {
                        const HostTable& T = tables[ht];
                        ClbBTabMeta m; m.x0 = T.x0; m.dx = T.dx; m.invdx = 1.0 / T.dx; m.n = T.n; m.off = (int)(cf.size() / 4);
                        std::vector<double> c1, c2;
                        if (T.interp == 1) { linear_coeffs(T.n, T.dx, T.f.data(), c1); linear_coeffs(T.n, T.dx, T.e.data(), c2); }
                        else if (T.interp == 2) { akima_coeffs(T.n, T.dx, T.f.data(), c1); akima_coeffs(T.n, T.dx, T.e.data(), c2); }
                        else { spline_coeffs(T.n, T.dx, T.f.data(), c1); spline_coeffs(T.n, T.dx, T.e.data(), c2); }
                        cf.insert(cf.end(), c1.begin(), c1.end()); cec.insert(cec.end(), c2.begin(), c2.end());
                        tabmap[ht] = (int)btm.size(); btm.push_back(m);
                    }
                    p.table = tabmap[ht];
                }
                bp.push_back(p);
            }
        }
        bd[l] = d;
    }
    CK(d_bdesc.ensure(bd.size())); CK(d_bpots.ensure(std::max<size_t>(bp.size(), 1))); CK(d_btm.ensure(std::max<size_t>(btm.size(), 1)));
    CK(d_bcf.ensure(std::max<size_t>(cf.size(), 4))); CK(d_bce.ensure(std::max<size_t>(cec.size(), 4)));
    CK(cudaMemcpyAsync(d_bdesc.p, bd.data(), bd.size() * sizeof(ClbBondedDesc), cudaMemcpyHostToDevice, stream));
    if (!bp.empty()) CK(cudaMemcpyAsync(d_bpots.p, bp.data(), bp.size() * sizeof(ClbBPot), cudaMemcpyHostToDevice, stream));
    if (!btm.empty()) {
        CK(cudaMemcpyAsync(d_btm.p, btm.data(), btm.size() * sizeof(ClbBTabMeta), cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(d_bcf.p, cf.data(), cf.size() * 8, cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(d_bce.p, cec.data(), cec.size() * 8, cudaMemcpyHostToDevice, stream));
    }
    CK(cudaStreamSynchronize(stream));
    pots_dirty = false;
    forces_valid = false;
    if (lists_valid) TRY(configure_pair_launch());
    return CLB_OK;
}

// ------------------------------------------------------------------------------------------ tuple lists
extern "C" int clb_add_list(clb_engine* e, int arity, int* out) {
    if (!e || !out || arity < 2 || arity > 4) return e ? e->fail(CLB_ERR_ARG, "arity must be 2, 3 or 4") : CLB_ERR_ARG;
    if (e->lists.size() >= CLB_MAX_LISTS) return e->fail(CLB_ERR_UNSUPPORTED, "too many tuple lists");
    HostList l; l.arity = arity;
    e->lists.push_back(std::move(l));
    *out = (int)e->lists.size() - 1;
    e->pots_dirty = true; e->terms_dirty = true; e->react_dirty = true;
    return CLB_OK;
}
int clb_engine::list_reserve(int li, long long need) {
    clb_engine* e = this;
    HostList& l = lists[li];
    if ((size_t)need * l.arity <= l.d.n) return CLB_OK;
    // generous: cudaMalloc/cudaFree inside a run can stall for 100s of ms, so regrowth must be rare
    size_t newcap = std::max<size_t>((size_t)need * l.arity * 2 + 1024, (size_t)std::max(n, 4096) * l.arity);
    DevBuf<int> nb;
    CK(nb.ensure(newcap));
    if (l.n) CK(cudaMemcpyAsync(nb.p, l.d.p, (size_t)l.n * l.arity * 4, cudaMemcpyDeviceToDevice, stream));
    CK(cudaStreamSynchronize(stream));
    l.d.release(); l.d = nb; nb.p = nullptr; nb.n = 0;
    lists_ptr_dirty = true;
    return CLB_OK;
}
extern "C" int clb_list_add(clb_engine* e, int list, int64_t n, const int64_t* ids) {
    if (!e || list < 0 || list >= (int)e->lists.size() || n < 0 || (n && !ids)) return e ? e->fail(CLB_ERR_ARG, "clb_list_add: bad argument") : CLB_ERR_ARG;
    cudaSetDevice(e->device);
    HostList& l = e->lists[list];
    TRY(e->list_reserve(list, l.n + n));
    // ids -> slots straight into the tuple array (on the device for dense ids)
    TRY(slots_to_device(e, (long long)n * l.arity, ids, e->lists[list].d.p + (size_t)l.n * l.arity, "tuple list"));
    l.n += n;
    e->terms_dirty = true; e->forces_valid = false; e->cont_ok = false;
    if (l.arity == 2 && l.tm_observed) e->topo_dirty = true;
    return CLB_OK;
}
extern "C" int64_t clb_list_size(clb_engine* e, int list) {
    if (!e || list < 0 || list >= (int)e->lists.size()) return -1;
    return e->lists[list].n;
}
extern "C" int clb_list_get(clb_engine* e, int list, int64_t cap, int64_t* ids, int64_t* n_out) {
    if (!e || list < 0 || list >= (int)e->lists.size()) return e ? e->fail(CLB_ERR_ARG, "bad list handle") : CLB_ERR_ARG;
    cudaSetDevice(e->device);
    HostList& l = e->lists[list];
    if (n_out) *n_out = l.n;
    long long m = std::min<long long>(l.n, cap);
    if (m > 0 && ids) {
        std::vector<int> h((size_t)m * l.arity);
        CK(cudaMemcpy(h.data(), l.d.p, h.size() * 4, cudaMemcpyDeviceToHost));
        for (size_t k = 0; k < h.size(); ++k) ids[k] = e->ids[h[k]];
    }
    return CLB_OK;
}
extern "C" int clb_add_bonded(clb_engine* e, int list, int typed, int* out) {
    if (!e || !out || list < 0 || list >= (int)e->lists.size()) return e ? e->fail(CLB_ERR_ARG, "bad list handle") : CLB_ERR_ARG;
    if (e->lists[list].bonded >= 0) return e->fail(CLB_ERR_UNSUPPORTED, "list %d already carries a bonded interaction", list);
    HostBonded b; b.list = list; b.typed = typed ? 1 : 0;
    HostInter it; it.kind = 10 + e->lists[list].arity; it.bonded = (int)e->bondeds.size();
    e->inters.push_back(it);
    b.inter = (int)e->inters.size() - 1;
    e->lists[list].bonded = (int)e->bondeds.size();
    e->bondeds.push_back(b);
    *out = b.inter;
    e->pots_dirty = true; e->terms_dirty = true;
    return CLB_OK;
}
extern "C" int clb_bonded_set_potential(clb_engine* e, int inter, int t1, int t2, int t3, int t4, int kind, const double* params,
                                        int np, int table) {
    if (!e || inter < 0 || inter >= (int)e->inters.size() || e->inters[inter].bonded < 0) return e ? e->fail(CLB_ERR_ARG, "bad bonded interaction handle") : CLB_ERR_ARG;
    HostBonded& b = e->bondeds[e->inters[inter].bonded];
    int ar = e->lists[b.list].arity;
    bool ok = (ar == 2 && (kind == CLB_POT_HARMONIC || kind == CLB_POT_TABULATED || kind == CLB_POT_FENE || kind == CLB_POT_FENE_LJ || kind == CLB_POT_LENNARD_JONES)) ||
              (ar == 3 && (kind == CLB_POT_ANGULAR_HARMONIC || kind == CLB_POT_TABULATED_ANGULAR || kind == CLB_POT_COSINE)) ||
              (ar == 4 && (kind == CLB_POT_TABULATED_DIHEDRAL || kind == CLB_POT_DIHEDRAL_HARMONIC));
    if (!ok) return e->fail(CLB_ERR_ARG, "potential kind %d does not fit a list of arity %d", kind, ar);
    ClbBPot p; memset(&p, 0, sizeof(p));
    p.kind = kind; p.table = table; p.t[0] = t1; p.t[1] = t2; p.t[2] = t3; p.t[3] = t4;
    for (int i = 0; i < np && i < 6; ++i) p.p[i] = params[i];
    if (kind == 2 || kind == 4 || kind == 5) { if (table < 0 || table >= (int)e->tables.size()) return e->fail(CLB_ERR_ARG, "bad table handle"); }
    if (!b.typed) { b.pots.clear(); b.pots.push_back(p); }
    else {
        bool replaced = false;
        for (auto& o : b.pots) {
            bool fwd = true, rev = true;
            for (int m = 0; m < ar; ++m) { fwd &= o.t[m] == p.t[m]; rev &= o.t[m] == p.t[ar - 1 - m]; }
            if (fwd || rev) { o = p; replaced = true; break; }
        }
        if (!replaced) b.pots.push_back(p);
    }
    e->pots_dirty = true; e->forces_valid = false; e->cont_ok = false;
    return CLB_OK;
}

// CSR by slot of all (tuple, role) memberships of lists that carry a bonded interaction.
int clb_engine::build_term_csr() {
    clb_engine* e = this;
    long long total = 0;
    for (auto& l : lists) if (l.bonded >= 0) total += l.n * l.arity;
    nterms = total;
    CK(term_off.ensure((size_t)n + 2));
    CK(term_meta.ensure((size_t)total + 16)); CK(term_tuple.ensure((size_t)total + 16));
    if (total == 0) { CK(cudaMemsetAsync(term_off.p, 0, ((size_t)n + 2) * 4, stream)); terms_dirty = false; return CLB_OK; }
    ClbTrace tr(stream, "term_csr");
    { size_t want = tkey.p ? (size_t)total : (size_t)total * 2 + n;   // headroom: reactions append terms
      CK(tkey.ensure(want)); CK(tkey2.ensure(want)); CK(tval.ensure(want)); CK(tval2.ensure(want)); CK(term_meta.ensure(want + 16)); CK(term_tuple.ensure(want + 16)); }
    tr.mark("alloc");
    long long base = 0;
    for (size_t li = 0; li < lists.size(); ++li) {
        HostList& l = lists[li];
        if (l.bonded < 0 || l.n == 0) continue;
        k_term_expand<<<ceil_div(l.n, 256), 256, 0, stream>>>((int)l.n, l.arity, (int)li, l.d.p, (int)base, tkey.p, tval.p);
        base += l.n * l.arity;
    }
    tr.mark("expand");
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, tkey.p, tkey2.p, tval.p, tval2.p, (int)tkey.n, 0, 32, stream);
    CK(cubtmp2.ensure(tb + 256));
    tr.mark("tmpalloc");
    cub::DeviceRadixSort::SortPairs(cubtmp2.p, tb, tkey.p, tkey2.p, tval.p, tval2.p, (int)total, 0, 32, stream);
    tr.mark("sort");
    k_term_unpack<<<ceil_div(total, 256), 256, 0, stream>>>((int)total, tval2.p, term_meta.p, term_tuple.p);
    tr.mark("unpack");
    k_lower_bounds<<<ceil_div(n + 1, 256), 256, 0, stream>>>((int)total, tkey2.p, n, term_off.p);
    tr.mark("bounds");
    CK(cudaGetLastError());
    terms_dirty = false;
    lists_ptr_dirty = true;
    rt_valid = false;
    return CLB_OK;
}
int clb_engine::upload_list_ptrs() {
    clb_engine* e = this;
    std::vector<const int*> h(CLB_MAX_LISTS, nullptr);
    for (size_t l = 0; l < lists.size(); ++l) h[l] = lists[l].d.p;
    CK(d_list_ptrs.ensure(CLB_MAX_LISTS));
    CK(cudaMemcpyAsync(d_list_ptrs.p, h.data(), CLB_MAX_LISTS * sizeof(int*), cudaMemcpyHostToDevice, stream));
    CK(cudaStreamSynchronize(stream));
    lists_ptr_dirty = false;
    return CLB_OK;
}
// resolve memberships to sorted indices (after every re-sort or topology change)
int clb_engine::resolve_terms() {
    clb_engine* e = this;
    int no = own1 - own0;
    CK(rt_off.ensure((size_t)no + 2)); CK(rt_cnt.ensure((size_t)no + 2));
    CK(rt_mem.ensure((size_t)nterms + 16)); CK(rt_meta.ensure((size_t)nterms + 16));
    if (lists_ptr_dirty) TRY(upload_list_ptrs());
    if (no > 0) {
        CK(cudaMemsetAsync(rt_cnt.p, 0, ((size_t)no + 2) * 4, stream));
        k_term_counts<<<ceil_div(no, 256), 256, 0, stream>>>(own0, own1, slot.p, term_off.p, rt_cnt.p);
        size_t tb = cubtmp.n;
        cub::DeviceScan::ExclusiveSum(cubtmp.p, tb, rt_cnt.p, rt_off.p, no + 1, stream);
        if (nterms) k_term_resolve<<<ceil_div(no, 256), 256, 0, stream>>>(own0, own1, slot.p, term_off.p, term_meta.p, term_tuple.p,
                                                                           (const int* const*)d_list_ptrs.p, d_bdesc.p, id2idx.p, rt_off.p, rt_mem.p, rt_meta.p, d_ctl);
    }
    CK(cudaGetLastError());
    rt_valid = true;
    return CLB_OK;
}

// ------------------------------------------------------------------------------------------ rebuild
int clb_engine::read_ctl() {
    clb_engine* e = this;
    CK(cudaMemcpyAsync(h_ctl, d_ctl, sizeof(ClbCtl), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    return CLB_OK;
}
__global__ void k_ctl_reset_stats(ClbCtl* c) { c->tile_max = 0; c->home_max = 0; c->cell_max = 0; c->nl_max = 0; c->nl_total = 0; c->err &= ~(CLB_EF_LIST_OVERFLOW | CLB_EF_TILE_OVERFLOW); }
__global__ void k_ctl_reset_after_overflow(ClbCtl* c) { c->nl_max = 0; c->nl_total = 0; c->err &= ~(CLB_EF_LIST_OVERFLOW | CLB_EF_TILE_OVERFLOW); }
__global__ void k_ctl_after_rebuild(ClbCtl* c) { c->stall = 0; c->accum_maxdist = 0.0; c->maxdisp2_bits = 0u; c->force_rebuild = 0; c->maybe = 0; }

int clb_engine::setup_sync() {
    if (n <= 0) return fail(CLB_ERR_STATE, "no particles");
    ClbTrace tr(stream, "setup_sync", pots_dirty || excl_dirty || terms_dirty || topo_dirty || react_dirty);
    if (pots_dirty) { TRY(upload_potentials()); tr.mark("pots"); }
    if (excl_dirty) { TRY(build_excl_csr()); lists_valid = false; tr.mark("excl"); }
    if (terms_dirty) { TRY(build_term_csr()); tr.mark("terms"); }
    if (topo_dirty) { TRY(build_topology()); tr.mark("topo"); }
    if (react_dirty) { TRY(upload_reactions()); tr.mark("react"); }
    return CLB_OK;
}

// launch geometry and shared-memory carve-up of the pair-force kernel; depends on the tile size of the
// last rebuild AND on the potentials, so it is refreshed after either changes
PairKernel clb_engine_pair_fn(const clb_engine* e, int in_smem, int split) {
    return (e->all_tab && e->branchfree_user) ? pair_kernel_tab(e->geo.cubic, in_smem, e->ugrid_on, split) : pair_kernel(e->geo.cubic, in_smem, e->ugrid_on, split);
}

// type populations (replicated per-slot type|state words) -> residency weights of the table windows
__global__ void k_type_hist(int n, const int* __restrict__ wslot, unsigned long long* __restrict__ hist) {
    __shared__ unsigned int s_h[CLB_MAX_TYPES];
    if (threadIdx.x < CLB_MAX_TYPES) s_h[threadIdx.x] = 0;
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) atomicAdd(&s_h[pw_type(wslot[i])], 1u);
    __syncthreads();
    if (threadIdx.x < CLB_MAX_TYPES && s_h[threadIdx.x]) atomicAdd(hist + threadIdx.x, (unsigned long long)s_h[threadIdx.x]);
}
// Choose which table windows live in shared memory (hottest type pairs first, by population product) and upload the
// descriptors + the shared-memory image.  budget_bytes: shared memory available for the windows; rlog: log2 of the replication.
int clb_engine::configure_tables(size_t budget_bytes, int rlog) {
    clb_engine* e = this;
    const int nt = nt_dev, ntp = nt * nt;
    CK(d_hist.ensure(CLB_MAX_TYPES));
    CK(cudaMemsetAsync(d_hist.p, 0, CLB_MAX_TYPES * 8, stream));
    k_type_hist<<<std::min(ceil_div(n, 256), 4 * nsm), 256, 0, stream>>>(n, wslot.p, d_hist.p);
    unsigned long long hist[CLB_MAX_TYPES];
    CK(cudaMemcpyAsync(hist, d_hist.p, sizeof(hist), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    for (auto& sl : t3_slots) { sl.weight = 0.0; sl.srow = -1; }
    for (int a = 0; a < nt; ++a) for (int b = 0; b < nt; ++b) {
        const int z = t3_pair_slot[(size_t)a * nt + b];
        if (z >= 0) t3_slots[z].weight += (double)hist[a] * (double)hist[b];
    }
    std::vector<int> order(t3_slots.size());
    for (size_t z = 0; z < order.size(); ++z) order[z] = (int)z;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return t3_slots[x].weight > t3_slots[y].weight; });
    const size_t budget_rows = budget_bytes / (sizeof(double2) << rlog);
    size_t used = 0; double wsum = 0, wres = 0; int nres = 0;
    for (int z : order) wsum += t3_slots[z].weight;
    for (int z : order) {
        T3Slot& sl = t3_slots[z];
        const size_t rows = (size_t)(sl.w1 - sl.w0);
        if (sl.weight <= 0 && nres > 0) continue;             // unpopulated type pairs stay in global memory
        if (used + rows > budget_rows) continue;
        sl.srow = (int)used; used += rows; wres += sl.weight; ++nres;
    }
    std::vector<double2> img(std::max<size_t>(used << rlog, 1));
    const int R = 1 << rlog;
    for (const T3Slot& sl : t3_slots) if (sl.srow >= 0)
        for (int i = sl.w0; i < sl.w1; ++i) for (int c = 0; c < R; ++c) img[(((size_t)sl.srow + (i - sl.w0)) << rlog) + c] = t3_rows[sl.off + i];
    std::vector<ClbPairDesc3> pd3(ntp); std::vector<int2> gm(ntp);
    bool one = true;
    for (int k = 0; k < ntp; ++k) {
        ClbPairDesc3 d; d.rc2 = -1.0; d.soff = 0; d.w0 = 0; d.wn = 0;
        gm[k] = make_int2(0, 0);
        const int z = t3_pair_slot[k];
        if (z >= 0) {
            const T3Slot& sl = t3_slots[z];
            d.rc2 = t3_pair_rc2[k];
            // the kernel computes ONE row index ix on the common grid; this table's row is ix - shift
            gm[k] = make_int2(sl.off - sl.shift, (sl.shift << 20) | (sl.n - 1));
            if (sl.srow >= 0) { d.soff = (sl.srow - sl.w0 - sl.shift) * R; d.w0 = (unsigned short)(sl.w0 + sl.shift); d.wn = (unsigned short)(sl.w1 - sl.w0); }
        }
        pd3[k] = d;
        one = one && z >= 0 && z == t3_pair_slot[0] && t3_pair_rc2[k] == t3_pair_rc2[0];
    }
    CK(d_pd3.ensure(ntp)); CK(d_gmeta.ensure(ntp)); CK(d_swin.ensure(img.size()));
    CK(cudaMemcpyAsync(d_pd3.p, pd3.data(), ntp * sizeof(ClbPairDesc3), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(d_gmeta.p, gm.data(), ntp * sizeof(int2), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(d_swin.p, img.data(), img.size() * sizeof(double2), cudaMemcpyHostToDevice, stream));
    CK(cudaStreamSynchronize(stream));
    tab3_nsrows = (int)used; tab3_rlog = rlog; tab3_onepd = one ? 1 : 0; tab3_resident = nres;
    tab3_resident_weight = wsum > 0 ? wres / wsum : 1.0;
    tab3_one = pd3[0]; tab3_one_g = gm[0];
    t3_dirty = false;
    return CLB_OK;
}

int clb_engine::configure_pair_launch() {
    pair_kernel_active = (tab2_ok && branchfree_user && pair_kernel_user != 1) ? 2 : 1;
    if (tab3_ok && branchfree_user && (pair_kernel_user == 0 || pair_kernel_user == 3)) pair_kernel_active = 3;
    if (pair_kernel_active == 3) {
        const double mean_home = (double)(own1 - own0) / std::max(1, grid.nblocks);
        int npw = std::max(1, std::min((home_max + 31) / 32, (int)ceil((mean_home + 3.0 * sqrt(mean_home)) / 32.0)));
        if (pair_warps_user > 0) npw = pair_warps_user;
        npw = std::min(npw, 16);
        const int tile_cap = (tile_max + 3) & ~3;
        pair_pipe = pair_pipe_user > 0 ? 1 : 0;     // opt-in: measured slower on B200 (two buffers halve the resident tiles; profiles/r2k_pipe_c2.log)
        pair_tile_cap = tile_cap;
        const size_t vcb = 64 + (size_t)(pair_pipe ? 2 : 1) * tile_cap * sizeof(int4);   // [mbarriers 32 B][TileMeta x 2][tile buffer(s)]
        size_t want_rows = 0;
        bool single = true;
        for (const T3Slot& sl : t3_slots) want_rows += (size_t)(sl.w1 - sl.w0);
        for (size_t k = 0; k < t3_pair_slot.size(); ++k) single = single && t3_pair_slot[k] >= 0 && t3_pair_slot[k] == t3_pair_slot[0] && t3_pair_rc2[k] == t3_pair_rc2[0];
        const size_t fixed = single ? 0 : (size_t)nt_dev * nt_dev * sizeof(ClbPairDesc3);
        // replication (8 conflict-free copies of the windows) is opt-in: measured on B200 (1M-bead melt, one table) the copies
        // cost two of five resident tiles and the kernel, which is latency-bound rather than wavefront-bound, gets 20 % slower
        // (profiles/r2b_sweep_c2.log: 0.334 -> 0.403 ms)
        int rlog = 0;
        if (pair_rep_user >= 8 && single) rlog = 3;
        // table budget: everything that is wanted, but at least 3 tiles (virtual CTAs) must stay resident
        const int nv_keep = pair_pipe ? 2 : 3;
        const size_t per_vc = vcb + (size_t)npw * 32 * 32;          // tile buffer(s) + the two 16-byte entry slots of its threads
        size_t budget = (size_t)smem_optin > fixed + nv_keep * per_vc ? (size_t)smem_optin - fixed - nv_keep * per_vc : 0;
        if (pair_table_kb_user >= 0) budget = std::min(budget, (size_t)pair_table_kb_user * 1024);
        budget = std::min(budget, (want_rows << rlog) * sizeof(double2));
        if (t3_dirty || rlog != tab3_rlog || ((size_t)tab3_nsrows << tab3_rlog) * sizeof(double2) > budget) { int r = configure_tables(budget, rlog); if (r != CLB_OK) return r; }
        const size_t shared_part = fixed + ((size_t)tab3_nsrows << tab3_rlog) * sizeof(double2);
        tab3_fb = pair_fb_user >= 0 ? pair_fb_user : (tab3_resident_weight < 0.999 ? 1 : 0);
        if (tab3_onepd) tab3_fb = 0;
        if (tab3_fb && !pair_ni_user) pair_ni = 2;          // the inline global gather needs the registers of two interleaved pairs
        else if (!pair_ni_user) pair_ni = 4;
        int best_nv = 1, best_nb = 1; double best_w = -1;
        const int nv_max = pair_nv_user > 0 ? pair_nv_user : 15;
        for (int nv = (pair_nv_user > 0 ? pair_nv_user : 1); nv <= nv_max; ++nv) {
            if (nv * npw * 32 > 1024) break;
            size_t smem = shared_part + nv * vcb + (size_t)nv * npw * 32 * 32;      // + two 16-byte entry slots per thread
            if ((int)smem > smem_optin) break;
            int nb = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, pair_kernel_tab3(tab3_onepd, tab3_rlog, pair_ni, tab3_fb), nv * npw * 32, smem);
            double w = (double)nb * nv * npw;
            if (w > best_w * 1.001) { best_w = w; best_nv = nv; best_nb = std::max(nb, 1); }
        }
        pair_nv = best_nv; pair_vc_bytes = (int)vcb;
        pair_split = 1; pair_npw = npw; pair_threads = best_nv * npw * 32;
        pair_smem = (int)(shared_part + best_nv * vcb + (size_t)best_nv * npw * 32 * 32);
        tabs_smem = tab3_nsrows > 0;
        if (pair_smem > smem_optin) return fail(CLB_ERR_UNSUPPORTED, "pair kernel needs %d B of shared memory: lower block_cells", pair_smem);
        pair_grid = std::min(ceil_div(grid.nblocks, best_nv), best_nb * nsm);
        return CLB_OK;
    }
    if (pair_kernel_active == 2) {
        const double mean_home = (double)(own1 - own0) / std::max(1, grid.nblocks);
        int npw = std::max(1, std::min((home_max + 31) / 32, (int)ceil((mean_home + 3.0 * sqrt(mean_home)) / 32.0)));
        if (pair_warps_user > 0) npw = pair_warps_user;
        npw = std::min(npw, 16);
        // one CTA hosts `nv` virtual CTAs (npw warps, one tile each) that share the table rows: pick the nv that keeps
        // the most warps resident per SM (the kernel is latency-bound: DESIGN.md 3.1)
        const size_t fixed = tab2_onepd ? 0 : (size_t)nt_dev * nt_dev * sizeof(double2);
        const size_t rows = (size_t)nrows_dev * sizeof(double2);
        const int tile_cap = (tile_max + 3) & ~3;
        const size_t vcb = (size_t)(CLB_TILE_CELLS + 4 + CLB_TILE_CELLS) * sizeof(int) + (size_t)tile_cap * sizeof(int4);
        tabs_smem = 1;
        if (fixed + rows + vcb > (size_t)smem_optin || tabs_smem_user == 0) tabs_smem = 0;
        const size_t shared_part = fixed + (tabs_smem ? rows : 0);
        int best_nv = 1, best_nb = 1; double best_w = -1;
        const int nv_max = pair_nv_user > 0 ? pair_nv_user : 15;
        for (int nv = (pair_nv_user > 0 ? pair_nv_user : 1); nv <= nv_max; ++nv) {
            if (nv * npw * 32 > 1024) break;
            size_t smem = shared_part + nv * vcb;
            if ((int)smem > smem_optin) break;
            int nb = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, pair_kernel_tab2(tabs_smem, tab2_onepd, pair_ni), nv * npw * 32, smem);
            double w = (double)nb * nv * npw;
            if (w > best_w * 1.001) { best_w = w; best_nv = nv; best_nb = std::max(nb, 1); }
        }
        pair_nv = best_nv; pair_vc_bytes = (int)vcb;
        pair_split = 1; pair_npw = npw; pair_threads = best_nv * npw * 32;
        pair_smem = (int)(shared_part + best_nv * vcb);
        if (pair_smem > smem_optin) return fail(CLB_ERR_UNSUPPORTED, "pair kernel needs %d B of shared memory: lower block_cells", pair_smem);
        pair_grid = std::min(ceil_div(grid.nblocks, best_nv), best_nb * nsm);
        return CLB_OK;
    }
    // warps that cover the home particles of a block: mean + 3 sigma (Poisson); rarer, fuller blocks take a second pass
    const double mean_home = (double)(own1 - own0) / std::max(1, grid.nblocks);
    // warps that cover the home particles of a block in one pass: mean + 3 sigma (a 1 x 1 x bx row of cells has a large
    // surface, its occupancy fluctuates almost like a Poisson variable); measured on B200: sizing for the mean (5 warps
    // at 148 beads/block) forces a second pass on ~1/4 of the blocks and costs 22 %
    int npw = std::max(1, std::min((home_max + 31) / 32, (int)ceil((mean_home + 3.0 * sqrt(mean_home)) / 32.0)));
    if (pair_warps_user > 0) npw = pair_warps_user;
    size_t fixed = (size_t)nt_dev * nt_dev * (sizeof(ClbPairDesc) + sizeof(double2)) + (size_t)ntabs_dev * sizeof(ClbTabMeta);
    size_t rows = (size_t)nrows_dev * sizeof(double2);
    size_t tile = (size_t)tile_max * sizeof(int4) + 16;
    bool in_smem = tabs_smem_user != 0 && fixed + rows + tile <= (size_t)std::min(smem_optin, 110 * 1024);
    if (tabs_smem_user == 2 && fixed + rows + tile <= (size_t)smem_optin) in_smem = true;
    tabs_smem = in_smem ? 1 : 0;
    // split factor: as many working warps per SM as registers allow (~32) for the CTAs that fit in shared memory
    int best_split = 1;
    double best_warps = 0;
    for (int sp = 1; sp <= 4; sp *= 2) {
        if (npw * sp * 32 > 512) break;
        size_t smem = fixed + (in_smem ? rows : 0) + tile + (size_t)(sp - 1) * 3 * npw * 32 * sizeof(double);
        if ((int)smem > smem_optin) break;
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, clb_engine_pair_fn(this, in_smem, sp), npw * sp * 32, smem);
        double w = (double)nb * npw * sp;
        if (w > best_warps * 1.05) { best_warps = w; best_split = sp; }
    }
    // measured on B200 (1M-bead melt): the kernel is shared-memory-pipe bound, extra split warps do not pay -> default 1
    (void)best_split;
    pair_split = pair_split_user > 0 ? pair_split_user : 1;
    if (npw * pair_split * 32 > 512) pair_split = std::max(1, 512 / (npw * 32));
    if (pair_split == 3) pair_split = 2;
    pair_npw = npw;
    pair_threads = npw * pair_split * 32;
    pair_smem = (int)(fixed + (in_smem ? rows : 0) + tile + (size_t)(pair_split - 1) * 3 * npw * 32 * sizeof(double));
    if (pair_smem > smem_optin) return fail(CLB_ERR_UNSUPPORTED, "pair kernel needs %d B of shared memory: lower block_cells", pair_smem);
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, clb_engine_pair_fn(this, tabs_smem, pair_split), pair_threads, pair_smem);
    pair_grid = std::min(grid.nblocks, std::max(1, nb) * nsm);
    return CLB_OK;
}

int clb_engine::rebuild() {
    clb_engine* e = this;
    ClbTrace tr(stream, "rebuild");
    bucket_begin(CLB_B_NEIGH);
    if (nranks > 1) {
        // ghosts are dropped, owned particles that left the slab move to the neighbour ranks (engine_comm.inl)
        CK(cudaMemsetAsync(id2idx.p, 0xff, (size_t)n * sizeof(int), stream));
        TRY(peer_active() ? comm_migrate_peer() : comm_migrate());
    }
    if (!bx_user && !bx_auto_done) {
        // blocks are cut by home-particle count (make_blocks); the cell cap only bounds the shared-memory offset tables and
        // keeps sparse rows from producing very long tiles
        // at most ncx - 2 cells: a tile row then never wraps onto itself (the prefilter build kernel needs that)
        set_block_cells(block_target_user < 0 ? 8 : std::max(1, std::min(CLB_MAX_BX, grid.ncx - 2)));
        bx_auto_done = true;
    }
    int ns = own1;
    // 1. sort the owned particles by cell
    k_cell_keys<<<ceil_div(std::max(ns, 1), 256), 256, 0, stream>>>(ns, pos.p, grid, key.p, val.p);
    int bits = 1; while ((1ll << bits) < grid.ncell) ++bits;
    size_t tb = cubtmp.n;
    cub::DeviceRadixSort::SortPairs(cubtmp.p, tb, key.p, key2.p, val.p, val2.p, ns, 0, bits, stream);
    k_gather<<<ceil_div(std::max(ns, 1), 256), 256, 0, stream>>>(ns, val2.p, pos.p, vel.p, slot.p, pos2.p, vel2.p, slot2.p, xref.p, id2idx.p);
    std::swap(pos, pos2); std::swap(vel, vel2); std::swap(slot, slot2);      // whole handles (pointer, capacity and block size)
    if (nranks > 1) {
        // boundary planes of the sorted owned range -> neighbours' ghost planes, appended behind the owned range
        k_cell_start<<<ceil_div(ns + 1, 256), 256, 0, stream>>>(ns, key2.p, grid.ncell, cell_start.p);
        TRY(peer_active() ? comm_exchange_ghosts_peer() : comm_exchange_ghosts());
        ns = nstored;
    }
    k_cell_start<<<ceil_div(ns + 1, 256), 256, 0, stream>>>(ns, key2.p, grid.ncell, cell_start.p);
    tr.mark("sort");
    // 2. tile statistics (read back together with the build result: one host check per rebuild instead of two)
    k_ctl_reset_stats<<<1, 1, 0, stream>>>(d_ctl);
    TRY(make_blocks());
    ClbGrid gdev = grid; gdev.nblocks = -1;                 // the block count of this rebuild is only on the device so far
    const int nblk_bound = grid.ncy * grid.nczl * grid.ncx;
    k_block_stats<<<ceil_div(nblk_bound, 128), 128, 0, stream>>>(gdev, cell_start.p, d_ctl);
    int tile_guess = tile_max > 0 ? (int)(tile_max * 1.08) + 32 : 0;
    if (tile_guess == 0) {                       // first rebuild: nothing to extrapolate from
        TRY(read_ctl());
        tile_guess = h_ctl->tile_max;
    }
    tr.mark("stats");
    // 3. neighbour lists (retry with a larger capacity on overflow)
    if (nl_cap == 0 || nl_cap_user != nl_cap_user_seen) {
        double rho = (double)n / (box[0] * box[1] * box[2]);
        double rl = rc + skin;
        int expect = (int)(4.18879 * rl * rl * rl * rho);
        nl_cap = nl_cap_user > 0 ? nl_cap_user : ((expect * 3 / 2 + 32 + 7) / 8) * 8;
        nl_cap_user_seen = nl_cap_user;
    }
    const int threads = build_threads;
    const unsigned long long rl2_lat = (unsigned long long)floor(geo.rl2 / geo.q2);
    // prefilter kernel: cubic boxes whose tile rows never wrap around x
    build_kernel_active = (geo.cubic && grid.ncx >= grid.bx + 2 && build_kernel_user != 1) ? 2 : 1;
    int rq2 = 0;
    if (build_kernel_active == 2) {
        CK(qsub.ensure(ncap));
        k_qsub<<<ceil_div(std::max(ns, 1), 256), 256, 0, stream>>>(ns, pos.p, grid, qsub.p);
        const double rq = (rc + skin) * CLB_QCELL * grid.ncx / box[0] + 1.7320508075688772 + 0.02;
        rq2 = (int)floor(rq * rq);
        ++launches;
    }
    for (int attempt = 0;; ++attempt) {
        CK(nl_entries.ensure((size_t)ncap * nl_cap));
        const int tile_cap = (tile_guess + 3) & ~3;     // keeps the per-warp staging rows 16-byte aligned
        size_t smem = (size_t)tile_cap * (sizeof(int4) + sizeof(int) + (build_kernel_active == 2 ? sizeof(unsigned) : 0)) +
                      (size_t)(threads / 32) * CLB_BUILD_G * nl_cap * sizeof(unsigned short) + 16;
        if ((int)smem > smem_optin) return fail(CLB_ERR_UNSUPPORTED, "tile needs %zu B of shared memory: lower block_cells", smem);
        int nb = 0;
        if (build_kernel_active == 2) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_build_lists2, threads, smem);
        else if (geo.cubic) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_build_lists<true>, threads, smem);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_build_lists<false>, threads, smem);
        int gridsz = std::min(grid.nblocks > 0 ? grid.nblocks + grid.nblocks / 8 + 1 : nblk_bound, std::max(1, nb) * nsm);
        if (build_kernel_active == 2) k_build_lists2<<<gridsz, threads, smem, stream>>>(gdev, rl2_lat, rq2, cell_start.p, pos.p, slot.p, qsub.p, excl_off.p, excl_ids.p, nl_entries.p, nl_count.p, nl_perm.p, nl_cap, tile_cap, d_ctl);
        else if (geo.cubic) k_build_lists<true><<<gridsz, threads, smem, stream>>>(gdev, geo, rl2_lat, cell_start.p, pos.p, slot.p, excl_off.p, excl_ids.p, nl_entries.p, nl_count.p, nl_perm.p, nl_cap, tile_cap, d_ctl);
        else k_build_lists<false><<<gridsz, threads, smem, stream>>>(gdev, geo, rl2_lat, cell_start.p, pos.p, slot.p, excl_off.p, excl_ids.p, nl_entries.p, nl_count.p, nl_perm.p, nl_cap, tile_cap, d_ctl);
        ++launches;
        TRY(read_ctl());
        tile_max = h_ctl->tile_max; home_max = h_ctl->home_max;
        grid.nblocks = h_ctl->nblocks; blk_p1 = h_ctl->blk_p1; blk_pl = h_ctl->blk_pl;
        if (tile_max > 65535) {
            if (getenv("CLB_TRACE")) {        // debugging aid: where do the sorted keys break?
                std::vector<int> hk(ns), hc(grid.ncell + 1);
                cudaMemcpy(hk.data(), key2.p, (size_t)ns * 4, cudaMemcpyDeviceToHost);
                cudaMemcpy(hc.data(), cell_start.p, ((size_t)grid.ncell + 1) * 4, cudaMemcpyDeviceToHost);
                int inv = -1; for (int i = 1; i < ns && inv < 0; ++i) if (hk[i] < hk[i - 1]) inv = i;
                int big = -1; for (int c = 0; c < grid.ncell && big < 0; ++c) if (hc[c + 1] - hc[c] > 4096) big = c;
                fprintf(stderr, "[clb rebuild] rank %d: tile_max %d ns %d own1 %d ncell %d first key inversion at %d (keys %d -> %d) first big cell %d (%d..%d) key[0] %d key[ns-1] %d nblocks %d\n",
                        rank, tile_max, ns, own1, grid.ncell, inv, inv > 0 ? hk[inv - 1] : -1, inv > 0 ? hk[inv] : -1, big, big >= 0 ? hc[big] : -1, big >= 0 ? hc[big + 1] : -1,
                        ns > 0 ? hk[0] : -1, ns > 0 ? hk[ns - 1] : -1, h_ctl->nblocks);
            }
            return fail(CLB_ERR_UNSUPPORTED, "tile of %d particles exceeds 16-bit list entries: lower block_cells", tile_max);
        }
        const bool tile_over = (h_ctl->err & CLB_EF_TILE_OVERFLOW) || tile_max > tile_cap;
        const bool list_over = (h_ctl->err & CLB_EF_LIST_OVERFLOW) != 0;
        if (!tile_over && !list_over) break;
        if (attempt > 6) return fail(CLB_ERR_RANGE, "neighbour list keeps overflowing (max %d entries, tile %d)", h_ctl->nl_max, tile_max);
        if (tile_over) tile_guess = tile_max;
        if (list_over) nl_cap = ((h_ctl->nl_max * 5 / 4 + 8 + 7) / 8) * 8;   // multiple of 8: rows are read as uint4
        k_ctl_reset_after_overflow<<<1, 1, 0, stream>>>(d_ctl);
    }
    nl_max = h_ctl->nl_max; nl_total = h_ctl->nl_total;
    tr.mark("build");
    // 4. pair-force launch configuration
    TRY(configure_pair_launch());
    // 5. bonded memberships -> sorted indices
    TRY(resolve_terms());
    k_ctl_after_rebuild<<<1, 1, 0, stream>>>(d_ctl);
    CK(cudaGetLastError());
    launches += 8;
    lists_valid = true; forces_valid = false;
    ++nrebuild;
    tr.mark("terms");
    bucket_end(CLB_B_NEIGH);
    return CLB_OK;
}

extern "C" int clb_decompose(clb_engine* e) {
    if (!e) return CLB_ERR_ARG;
    cudaSetDevice(e->device);
    TRY(e->setup_sync());
    TRY(e->rebuild());
    CK(cudaStreamSynchronize(e->stream));
    return CLB_OK;
}

// ------------------------------------------------------------------------------------------ forces
void clb_engine::launch_pair(int b0, int seg0, int b1, int nidx) {
    const int gridsz = std::max(1, std::min(pair_grid, nidx));
    if (pair_kernel_active == 3) {
        ClbPairArgs3 A;
        A.cell_start = cell_start.p; A.pos = pos.p; A.entries = nl_entries.p; A.nl_count = nl_count.p;
        A.perm = pair_perm_user ? nl_perm.p : nullptr;
        A.pd3 = d_pd3.p; A.gmeta = d_gmeta.p; A.trows = d_rows2.p; A.swin = d_swin.p; A.force = force.p; A.ctl = d_ctl;
        A.cap = nl_cap; A.ntypes = nt_dev; A.nsrows = tab3_nsrows; A.fstride = ncap; A.npw = pair_npw;
        A.invdx = tab2_invdx; A.cmagic = tab2_cmagic; A.one = tab3_one; A.one_g = tab3_one_g;
        A.b0 = b0; A.seg0 = seg0; A.b1 = b1; A.nidx = nidx; A.nv = pair_nv; A.vc_bytes = pair_vc_bytes;
        A.pipe = pair_pipe; A.tile_cap = pair_tile_cap;
        pair_kernel_tab3(tab3_onepd, tab3_rlog, pair_ni, tab3_fb)<<<std::max(1, std::min(pair_grid, ceil_div(nidx, pair_nv))), pair_threads, pair_smem, stream>>>(grid, A);
    } else if (pair_kernel_active == 2) {
        ClbPairArgs2 A;
        A.cell_start = cell_start.p; A.pos = pos.p; A.entries = nl_entries.p; A.nl_count = nl_count.p;
        A.pd2 = d_pd2.p; A.trows = d_rows2.p; A.force = force.p; A.ctl = d_ctl;
        A.cap = nl_cap; A.ntypes = nt_dev; A.nrows_total = nrows_dev; A.fstride = ncap; A.npw = pair_npw;
        A.invdx = tab2_invdx; A.cmagic = tab2_cmagic; A.nm1 = tab2_nm1; A.one_rc2 = tab2_one_rc2; A.one_off = tab2_one_off;
        A.b0 = b0; A.seg0 = seg0; A.b1 = b1; A.nidx = nidx; A.nv = pair_nv; A.vc_bytes = pair_vc_bytes;
        pair_kernel_tab2(tabs_smem, tab2_onepd, pair_ni)<<<std::max(1, std::min(pair_grid, ceil_div(nidx, pair_nv))), pair_threads, pair_smem, stream>>>(grid, A);
    } else {
        ClbPairArgs A;
        A.cell_start = cell_start.p; A.pos = pos.p; A.entries = nl_entries.p; A.nl_count = nl_count.p;
        A.pdesc = d_pd.p; A.plj = d_plj.p; A.tmeta = d_tm.p; A.trows = d_frows.p; A.force = force.p; A.ctl = d_ctl;
        A.cap = nl_cap; A.ntypes = nt_dev; A.ntabs = ntabs_dev; A.nrows_total = nrows_dev; A.fstride = ncap; A.npw = pair_npw;
        A.ugrid = ugrid_meta;
        A.b0 = b0; A.seg0 = seg0; A.b1 = b1; A.nidx = nidx;
        clb_engine_pair_fn(this, tabs_smem, pair_split)<<<gridsz, pair_threads, pair_smem, stream>>>(grid, geo, A);
    }
    ++launches;
}

// Pair + bonded forces of the owned particles.  overlap_halo (multi-GPU steps): the row blocks of the interior planes
// need no ghost, so they are launched first while the halo of this step is still in flight on the comm stream; the
// caller then makes the main stream wait for the halo (k_check_resort in between) and the two boundary planes follow.
void clb_engine::enqueue_forces(bool overlap_halo, bool resort_checked) {
    bucket_begin(CLB_B_PAIR);
    const int nb = grid.nblocks;
    const int p1 = blk_p1, pl = blk_pl;                  // first block of owned plane 1 / of the last owned plane
    const bool split = overlap_halo && nranks > 1 && grid.nczl >= 3;
    const size_t slot_i = pair_event_used / 2;
    if (pair_event_timing) {
        pair_event_valid.resize(slot_i + 1, 1); pair_event_valid[slot_i] = 1;
        pair_event_has2.resize(slot_i + 1, 0); pair_event_has2[slot_i] = 0;
        cudaEventRecord(next_pair_event(), stream);
    }
    if (split) launch_pair(p1, pl - p1, 0, pl - p1);                  // interior planes
    else if (!overlap_halo || nranks == 1) launch_pair(0, nb, 0, nb);
    if (pair_event_timing) cudaEventRecord(next_pair_event(), stream);
    if (overlap_halo && nranks > 1) {
        cudaStreamWaitEvent(stream, ev_comm, 0);                      // halo + global max displacement have arrived
        if (!resort_checked) {                                        // (the peer-mailbox path checks on the comm stream: k_check_resort_peer)
            k_check_resort<<<1, 1, 0, stream>>>(d_ctl, criterion, 0.5 * skin, pending_step_index);
            ++launches;
        }
        if (pair_event_timing) {
            while (pair_events2.size() < 2 * (slot_i + 1)) { cudaEvent_t ev; cudaEventCreate(&ev); pair_events2.push_back(ev); }
            pair_event_has2[slot_i] = 1;
            cudaEventRecord(pair_events2[2 * slot_i], stream);
        }
        if (split) launch_pair(0, p1, pl, p1 + (nb - pl));            // bottom and top owned planes
        else launch_pair(0, nb, 0, nb);
        if (pair_event_timing) cudaEventRecord(pair_events2[2 * slot_i + 1], stream);
    }
    bucket_end(CLB_B_PAIR);
    ++pair_launches_total;
    if (nterms > 0) {
        bucket_begin(CLB_B_BONDED);
        int no = own1 - own0;
        k_bonded<false><<<ceil_div(no, 256), 256, 0, stream>>>(own0, own1, pos.p, geo, rt_off.p, rt_mem.p, rt_meta.p, d_bdesc.p, d_bpots.p, d_btm.p,
                                                                 (const double4*)d_bcf.p, (const double4*)d_bce.p, force.p, ncap, -1, nullptr, d_ctl);
        bucket_end(CLB_B_BONDED);
        ++launches;
    }
    if (cap_force > 0) {
        int no = own1 - own0;
        k_cap_force<<<ceil_div(no, 256), 256, 0, stream>>>(own0, own1, force.p, ncap, cap_force, d_ctl);
        ++launches;
    }
}

int clb_engine::check_device_errors(const char* where) {
    unsigned er = h_ctl->err;
    if (er & CLB_EF_TABLE_RANGE) return fail(CLB_ERR_RANGE, "%s: tabulated potential index out of range (fatal in the reference as well)", where);
    if (er & CLB_EF_PARTNER_LOST) return fail(CLB_ERR_RANGE, "%s: a bonded partner is outside the ghost layer", where);
    if (er & CLB_EF_DEGREE) return fail(CLB_ERR_RANGE, "%s: more than %d bonds on one particle", where, CLB_MAXDEG);
    if (er & CLB_EF_TUPLE_OVERFLOW) return fail(CLB_ERR_RANGE, "%s: tuple list overflow", where);
    if (er & CLB_EF_BFS_OVERFLOW) return fail(CLB_ERR_RANGE, "%s: neighbour-property search exceeded its buffers (more than 64 particles in one bond shell)", where);
    return CLB_OK;
}

extern "C" int clb_compute_forces(clb_engine* e) {
    if (!e) return CLB_ERR_ARG;
    cudaSetDevice(e->device);
    TRY(e->setup_sync());
    if (!e->lists_valid) TRY(e->rebuild());
    e->enqueue_forces();
    TRY(e->read_ctl());
    CK(cudaGetLastError());
    e->forces_valid = true;
    return e->check_device_errors("clb_compute_forces");
}

extern "C" int clb_energy(clb_engine* e, int inter, double* out) {
    if (!e || !out || inter < 0 || inter >= (int)e->inters.size()) return e ? e->fail(CLB_ERR_ARG, "bad interaction handle") : CLB_ERR_ARG;
    cudaSetDevice(e->device);
    TRY(e->setup_sync());
    if (!e->lists_valid) TRY(e->rebuild());
    const HostInter& it = e->inters[inter];
    if (it.bonded < 0) {
        size_t smem = (size_t)e->tile_max * sizeof(int4) + 16;
        int threads = 256;
        int gridsz = std::min(e->grid.nblocks, 4 * e->nsm);
        gridsz = std::min(gridsz, 65536);
        if (e->geo.cubic) k_pair_energy<true><<<gridsz, threads, smem, e->stream>>>(e->grid, e->geo, e->cell_start.p, e->pos.p, e->nl_entries.p, e->nl_count.p, e->nl_cap, e->d_pd_e.p, e->d_pe.p, e->nt_dev, e->d_tm_e.p, e->d_erows.p, inter, e->partial.p, e->partial_u64.p);
        else k_pair_energy<false><<<gridsz, threads, smem, e->stream>>>(e->grid, e->geo, e->cell_start.p, e->pos.p, e->nl_entries.p, e->nl_count.p, e->nl_cap, e->d_pd_e.p, e->d_pe.p, e->nt_dev, e->d_tm_e.p, e->d_erows.p, inter, e->partial.p, e->partial_u64.p);
        k_sum_partials<<<1, 256, 0, e->stream>>>(gridsz, e->partial.p, (double*)e->d_scalar);
        k_sum_partials_u64<<<1, 256, 0, e->stream>>>(gridsz, e->partial_u64.p, (unsigned long long*)e->d_scalar + 1);
        CK(cudaMemcpyAsync(e->h_scalar, e->d_scalar, 16, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        *out = ((double*)e->h_scalar)[0];
        // full lists visit every interacting pair twice (once per owner, possibly on two ranks)
        double both[2] = {*out, (double)((unsigned long long*)e->h_scalar)[1]};
        if (e->nranks > 1) TRY(e->comm_allreduce_sum(both, 2));
        *out = both[0];
        e->last_interacting = (unsigned long long)(both[1] + 0.5) / 2;
        CK(cudaGetLastError());
        return CLB_OK;
    } else {
        int no = e->own1 - e->own0;
        int nb = ceil_div(no, 256);
        if (e->nterms == 0 || nb == 0) { *out = 0.0; return CLB_OK; }
        CK(e->partial.ensure(nb));
        k_bonded<true><<<nb, 256, 0, e->stream>>>(e->own0, e->own1, e->pos.p, e->geo, e->rt_off.p, e->rt_mem.p, e->rt_meta.p, e->d_bdesc.p, e->d_bpots.p, e->d_btm.p,
                                                  (const double4*)e->d_bcf.p, (const double4*)e->d_bce.p, e->force.p, e->ncap, inter, e->partial.p, e->d_ctl);
        k_sum_partials<<<1, 256, 0, e->stream>>>(nb, e->partial.p, (double*)e->d_scalar);
        CK(cudaMemcpyAsync(e->h_scalar, e->d_scalar, 8, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        *out = ((double*)e->h_scalar)[0];
    }
    CK(cudaGetLastError());
    if (e->nranks > 1) TRY(e->comm_allreduce_sum(out, 1));
    return CLB_OK;
}

extern "C" int clb_kinetics(clb_engine* e, double out[3]) {
    if (!e || !out) return CLB_ERR_ARG;
    cudaSetDevice(e->device);
    int no = e->own1 - e->own0, nb = ceil_div(no, 256);
    CK(e->partial.ensure(std::max(nb, 1)));
    k_kinetic<<<nb, 256, 0, e->stream>>>(e->own0, e->own1, e->vel.p, e->partial.p);
    k_sum_partials<<<1, 256, 0, e->stream>>>(nb, e->partial.p, (double*)e->d_scalar);
    CK(cudaMemcpyAsync(e->h_scalar, e->d_scalar, 8, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    double ek = ((double*)e->h_scalar)[0];
    if (e->nranks > 1) TRY(e->comm_allreduce_sum(&ek, 1));
    out[0] = ek; out[1] = 2.0 * ek / (3.0 * e->n); out[2] = (double)e->n;
    return CLB_OK;
}
extern "C" int clb_count_type(clb_engine* e, int type, int state, int64_t* out) {
    if (!e || !out) return CLB_ERR_ARG;
    cudaSetDevice(e->device);
    CK(cudaMemsetAsync(e->d_scalar, 0, 8, e->stream));
    int no = e->own1 - e->own0;
    k_count_type<<<ceil_div(no, 256), 256, 0, e->stream>>>(e->own0, e->own1, e->pos.p, type, state, (unsigned long long*)e->d_scalar);
    CK(cudaMemcpyAsync(e->h_scalar, e->d_scalar, 8, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    double c = (double)((unsigned long long*)e->h_scalar)[0];
    if (e->nranks > 1) TRY(e->comm_allreduce_sum(&c, 1));
    *out = (int64_t)c;
    return CLB_OK;
}

// ------------------------------------------------------------------------------------------ integrator
extern "C" int clb_set_dt(clb_engine* e, double dt) {
    if (!e || dt <= 0) return CLB_ERR_ARG;
    if (dt != e->dt) e->react_dirty = true;      // acceptance probability p = rate * dt * interval (U5)
    e->dt = dt;
    return CLB_OK;
}
extern "C" int clb_set_langevin(clb_engine* e, int enabled, double kT, double gamma, int ntypes, const int32_t* types) {
    if (!e) return CLB_ERR_ARG;
    if (enabled != e->lang_on || kT != e->kT) e->pots_dirty = true;     // the table windows of the pair kernel are thermal (30 kT)
    e->lang_on = enabled; e->kT = kT; e->gamma = gamma;
    if (ntypes <= 0) e->lang_mask = ~0ull;
    else { e->lang_mask = 0; for (int i = 0; i < ntypes; ++i) if (types[i] >= 0 && types[i] < 64) e->lang_mask |= 1ull << types[i]; }
    return CLB_OK;
}
extern "C" int clb_set_cap_force(clb_engine* e, double cap) {
    if (!e) return CLB_ERR_ARG;
    if (cap != e->cap_force) { e->forces_valid = false; e->cont_ok = false; }
    e->cap_force = cap;
    return CLB_OK;
}
extern "C" int64_t clb_step(const clb_engine* e) { return e ? e->step : 0; }

ClbIntegParams clb_engine::integ_params(uint64_t key_step) const {
    ClbIntegParams P;
    P.dt = dt;
    for (int d = 0; d < 3; ++d) { P.q[d] = geo.q[d]; P.invq[d] = 1.0 / geo.q[d]; }
    P.langevin = lang_on; P.pref1 = -gamma; P.pref2 = lang_on ? sqrt(24.0 * kT * gamma / dt) : 0.0;
    P.type_mask_lo = lang_mask; P.seed = seed; P.step = key_step; P.criterion = criterion; P.i0 = own0; P.i1 = own1;
    return P;
}
void clb_engine::enqueue_integrate(int mode, uint64_t key_step) {
    bucket_begin(CLB_B_INTEG);
    ClbIntegParams P = integ_params(key_step);
    int no = own1 - own0, nb = ceil_div(no, 256);
    if (mode == (CLB_INT_SECOND | CLB_INT_FIRST)) k_integrate<CLB_INT_SECOND | CLB_INT_FIRST><<<nb, 256, 0, stream>>>(P, pos.p, vel.p, force.p, ncap, xref.p, slot.p, image.p, d_ctl);
    else if (mode == CLB_INT_SECOND) k_integrate<CLB_INT_SECOND><<<nb, 256, 0, stream>>>(P, pos.p, vel.p, force.p, ncap, xref.p, slot.p, image.p, d_ctl);
    else k_integrate<CLB_INT_FIRST><<<nb, 256, 0, stream>>>(P, pos.p, vel.p, force.p, ncap, xref.p, slot.p, image.p, d_ctl);
    ++launches;
    bucket_end(CLB_B_INTEG);
}

// integrator.run(n) -- SURVEY 3.2.  Steps are enqueued in chunks without host synchronisation; the
// skin/2 test runs on the device and stalls the rest of the chunk when a rebuild is due.
// cont: continuation of the previous clb_run inside ONE integrator.run of the reference (the Python surface splits a
// run at ExtAnalyze / ATRPActivator intervals): the force array still holds the thermostatted forces of the last
// step, so neither the run-entry recalculation nor the heat-up kick is repeated (VelocityVerlet::run does them once).
static int run_impl(clb_engine* e, int64_t nsteps, bool cont) {
    if (!e || nsteps < 0) return CLB_ERR_ARG;
    cudaSetDevice(e->device);
    cont = cont && e->cont_ok && e->lists_valid;
    e->criterion = e->resort_criterion();          // 2 (per-cell bound) on one GPU, 1 (global maximum) across ranks, unless set by the caller
    if (e->criterion == 2) { if (e->nranks > 1) e->criterion = 1; else CK(e->cell_disp.ensure((size_t)e->grid.ncell + 1)); }
    ClbTrace trr(e->stream, "run");
    if (cont) {
        // list / term / exclusion updates left by a reaction pass are picked up at the forced rebuild of the first step
        if (e->pots_dirty) TRY(e->upload_potentials());
        if (e->topo_dirty) TRY(e->build_topology());
        if (e->react_dirty) TRY(e->upload_reactions());
        if (!e->pending_rebuild && (e->excl_dirty || e->terms_dirty)) cont = false;
    }
    if (!cont) {
        TRY(e->setup_sync());
        if (e->pending_rebuild) { e->lists_valid = false; e->pending_rebuild = false; }
        trr.mark("setup_sync");
        if (!e->lists_valid) TRY(e->rebuild());
        trr.mark("rebuild");
    }
    cudaEventRecord(e->ev_a[CLB_B_TOTAL], e->stream);
    if (!cont) {
        // run entry: recalc forces (+ thermostat heat-up), SURVEY 3.2
        e->enqueue_forces();
        if (e->lang_on) {
            ClbIntegParams P = e->integ_params((uint64_t)e->step);
            int no = e->own1 - e->own0;
            k_thermalize<<<ceil_div(no, 256), 256, 0, e->stream>>>(P, sqrt(3.0), CLB_STREAM_HEATUP, e->pos.p, e->vel.p, e->slot.p, e->force.p, e->ncap);
            ++e->launches;
        }
    }
    const double half_skin = 0.5 * e->skin;
    int64_t i = 0;
    bool pend = false;
    const bool react = e->react_on && !e->reactions.empty();
    auto boundary_after = [&](int64_t s) { return react && ((e->step + s + 1) % e->react_interval) == 0; };
    while (i < nsteps) {
        int chunk = e->chunk_user > 0 ? e->chunk_user : std::max(1, std::min(64, e->last_interval > 0 ? e->last_interval : 4));
        int64_t j = std::min<int64_t>(nsteps, i + chunk);
        const size_t chunk_first_pair = e->pair_event_used / 2;
        for (int64_t s = i; s < j; ++s) {
            if (pend && e->fuse) e->enqueue_integrate(CLB_INT_SECOND | CLB_INT_FIRST, (uint64_t)(e->step + s - 1));
            else {
                if (pend) e->enqueue_integrate(CLB_INT_SECOND, (uint64_t)(e->step + s - 1));
                e->enqueue_integrate(CLB_INT_FIRST, 0);
            }
            if (e->peer_active()) {
                // peer mailboxes: push boundary planes + displacement maximum over NVLink, wait for the neighbours, resort check.
                // Option overlap_halo=1: with >= 3 owned planes the exchange runs on the comm stream while the interior planes (whose
                // tiles hold no ghost) are evaluated; the two boundary planes follow once the neighbours' planes have landed.
                // Off by default: measured at 4 ranks the split launch costs more than the exchange it hides (3252 vs 3443 steps/s,
                // profiles/r2z_n4_ov{1,0}.json).
                if (e->overlap_user > 0 && e->grid.nczl >= 3) {
                    cudaEventRecord(e->ev_int, e->stream);
                    cudaStreamWaitEvent(e->comm_stream, e->ev_int, 0);
                    TRY(e->comm_step_peer(e->comm_stream, (int)(s - i)));
                    cudaEventRecord(e->ev_comm, e->comm_stream);
                    e->pending_step_index = (int)(s - i);
                    e->enqueue_forces(true, true);
                } else {
                    TRY(e->comm_step_peer(e->stream, (int)(s - i)));
                    e->enqueue_forces();
                }
            } else if (e->nranks > 1 && (e->overlap_user > 0 || (e->overlap_user < 0 && e->nranks >= 4))) {
                // global max displacement + position halo travel on the comm stream while the interior planes compute
                cudaEventRecord(e->ev_int, e->stream);
                cudaStreamWaitEvent(e->comm_stream, e->ev_int, 0);
                if (e->comm_group_user) TRY(e->comm_step(e->comm_stream));
                else { TRY(e->comm_max_displacement(e->comm_stream)); TRY(e->comm_halo_positions(e->comm_stream)); }
                cudaEventRecord(e->ev_comm, e->comm_stream);
                e->pending_step_index = (int)(s - i);
                e->enqueue_forces(true);            // interior blocks, wait, k_check_resort, boundary blocks, bonded
            } else {
                // single stream: the halo may travel before the resort check (a stalled step re-sends it after the rebuild)
                if (e->nranks > 1) {
                    if (e->comm_group_user) TRY(e->comm_step(e->stream));
                    else { TRY(e->comm_max_displacement(e->stream)); TRY(e->comm_halo_positions(e->stream)); }
                }
                k_check_resort<<<1, 1, 0, e->stream>>>(e->d_ctl, e->criterion, half_skin, (int)(s - i));
                ++e->launches;
                if (e->criterion == 2) {
                    // the per-cell bound (clb_kernels.cuh): both kernels return at once unless the global maximum is past skin/2
                    const int nc = e->grid.ncell;
                    k_cell_disp<<<ceil_div(nc, 128), 128, 0, e->stream>>>(e->d_ctl, nc, e->cell_start.p, e->pos.p, e->xref.p, e->geo.q[0], e->geo.q[1], e->geo.q[2], e->cell_disp.p);
                    k_cell_pairs<<<ceil_div(nc, 128), 128, 0, e->stream>>>(e->d_ctl, e->grid.ncx, e->grid.ncy, e->grid.ncz, e->cell_disp.p, (float)e->skin);
                    e->launches += 2;
                }
                e->enqueue_forces();
            }
            pend = true;
            if (boundary_after(s)) { j = s + 1; break; }
        }
        TRY(e->read_ctl());
        int64_t done;
        if (e->h_ctl->stall) {
            int64_t s = i + e->h_ctl->stall_step;
            // force launches of the stalled step and of every later step of this chunk were no-ops
            if (e->pair_event_timing) for (int64_t q = s; q < j; ++q) { size_t pi = chunk_first_pair + (size_t)(q - i); if (pi < e->pair_event_valid.size()) e->pair_event_valid[pi] = 0; }
            e->last_interval = (int)std::max<int64_t>(1, (e->step + s) - e->last_rebuild_step);
            e->last_rebuild_step = e->step + s;
            TRY(e->setup_sync());
            TRY(e->rebuild());
            e->pending_rebuild = false;
            e->enqueue_forces();
            done = s;
        } else done = j - 1;
        if (e->h_ctl->err & ~(CLB_EF_LIST_OVERFLOW | CLB_EF_TILE_OVERFLOW)) { TRY(e->check_device_errors("clb_run")); }
        pend = true;
        if (boundary_after(done)) {
            e->enqueue_integrate(CLB_INT_SECOND, (uint64_t)(e->step + done));
            pend = false;
            int64_t keep = e->step;
            e->step = keep + done + 1;
            int64_t nev = 0;
            int rr = e->react_pass(&nev);
            e->step = keep;
            if (rr != CLB_OK) return rr;
        }
        i = done + 1;
    }
    if (pend) e->enqueue_integrate(CLB_INT_SECOND, (uint64_t)(e->step + nsteps - 1));
    e->step += nsteps;
    trr.mark("steps");
    cudaEventRecord(e->ev_b[CLB_B_TOTAL], e->stream);
    TRY(e->read_ctl());
    CK(cudaGetLastError());
    e->collect_timers();
    e->nsteps_total += nsteps;
    e->forces_valid = true; // force array holds the thermostatted force of the last step
    e->cont_ok = true;
    return e->check_device_errors("clb_run");
}
extern "C" int clb_run(clb_engine* e, int64_t nsteps) { return run_impl(e, nsteps, false); }
extern "C" int clb_run_continue(clb_engine* e, int64_t nsteps) { return run_impl(e, nsteps, true); }

// ------------------------------------------------------------------------------------------ parity / timers
extern "C" int clb_get_pairs(clb_engine* e, int64_t cap, int64_t* pairs, int64_t* n_out) {
    if (!e) return CLB_ERR_ARG;
    cudaSetDevice(e->device);
    TRY(e->setup_sync());
    if (!e->lists_valid) TRY(e->rebuild());
    // single GPU: every pair sits in two rows; with ghosts a row entry may be the only local copy of its pair
    size_t outcap = (e->nranks > 1 ? (size_t)e->nl_total : (size_t)e->nl_total / 2) + 1024;
    DevBuf<int2> out;
    CK(out.ensure(outcap));
    CK(cudaMemsetAsync(&e->d_ctl->npairs_out, 0, 8, e->stream));
    size_t smem = (size_t)e->tile_max * sizeof(int) + 16;
    k_decode_pairs<<<std::min(e->grid.nblocks, 4 * e->nsm), 256, smem, e->stream>>>(e->grid, e->cell_start.p, e->slot.p, e->nl_entries.p, e->nl_count.p, e->nl_cap, out.p, outcap, e->d_ctl);
    TRY(e->read_ctl());
    CK(cudaGetLastError());
    size_t m = (size_t)e->h_ctl->npairs_out;
    if (m > outcap) { out.release(); return e->fail(CLB_ERR_STATE, "pair list asymmetric: %zu decoded pairs for %llu entries", m, e->nl_total); }
    std::vector<int2> h(m);
    if (e->nranks > 1) {
        // every unordered pair is decoded by the rank that owns its lower slot (U16): the union is the global pair set
        void* all = nullptr; size_t tot = 0;
        int rr = e->comm_allgatherv(out.p, m * sizeof(int2), &all, &tot);
        if (rr != CLB_OK) { out.release(); return rr; }
        m = tot / sizeof(int2);
        h.resize(m);
        if (m) CK(cudaMemcpy(h.data(), all, tot, cudaMemcpyDeviceToHost));
    } else if (m) CK(cudaMemcpy(h.data(), out.p, m * sizeof(int2), cudaMemcpyDeviceToHost));
    out.release();
    std::sort(h.begin(), h.end(), [](const int2& a, const int2& b) { return a.x != b.x ? a.x < b.x : a.y < b.y; });
    if (n_out) *n_out = (int64_t)m;
    for (size_t k = 0; k < m && (int64_t)k < cap; ++k) { pairs[2 * k] = e->ids[h[k].x]; pairs[2 * k + 1] = e->ids[h[k].y]; }
    return CLB_OK;
}

void clb_engine::bucket_begin(int b) { if (timers_on) { if (!bucket_open[b]) { cudaEventRecord(ev_a[b], stream); bucket_open[b] = 1; } } }
void clb_engine::bucket_end(int b) {
    if (timers_on && bucket_open[b]) {
        cudaEventRecord(ev_b[b], stream); cudaEventSynchronize(ev_b[b]);
        float ms = 0; cudaEventElapsedTime(&ms, ev_a[b], ev_b[b]); bucket_s[b] += ms * 1e-3; bucket_open[b] = 0;
    }
}
cudaEvent_t clb_engine::next_pair_event() {
    if (pair_events.size() <= pair_event_used) { cudaEvent_t ev; cudaEventCreate(&ev); pair_events.push_back(ev); }
    return pair_events[pair_event_used++];
}
void clb_engine::collect_timers() {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ev_a[CLB_B_TOTAL], ev_b[CLB_B_TOTAL]) == cudaSuccess) bucket_s[CLB_B_TOTAL] += ms * 1e-3;
    for (size_t k = 0; k + 1 < pair_event_used; k += 2) {
        float t = 0;
        if (k / 2 < pair_event_valid.size() && !pair_event_valid[k / 2]) continue;
        if (cudaEventElapsedTime(&t, pair_events[k], pair_events[k + 1]) == cudaSuccess) {
            pair_ms += t; ++pair_launches;
            float t2 = 0;   // boundary-plane launch of an overlapped multi-GPU step
            if (k / 2 < pair_event_has2.size() && pair_event_has2[k / 2] && cudaEventElapsedTime(&t2, pair_events2[k], pair_events2[k + 1]) == cudaSuccess) pair_ms += t2;
        }
    }
    pair_event_used = 0;
    pair_event_valid.clear();
    pair_event_has2.clear();
}
extern "C" int clb_timers(clb_engine* e, double out[8], int64_t counters[8]) {
    if (!e) return CLB_ERR_ARG;
    if (out) for (int k = 0; k < 8; ++k) out[k] = e->bucket_s[k];
    if (counters) {
        counters[0] = e->nsteps_total; counters[1] = e->nrebuild; counters[2] = e->launches; counters[3] = (int64_t)e->nl_total;
        counters[4] = e->nreact_pass; counters[5] = e->nreact_events; counters[6] = e->nstored - (e->own1 - e->own0); counters[7] = (int64_t)e->last_interacting;
    }
    return CLB_OK;
}
extern "C" int clb_reset_timers(clb_engine* e) {
    if (!e) return CLB_ERR_ARG;
    for (int k = 0; k < 8; ++k) e->bucket_s[k] = 0;
    e->nsteps_total = 0; e->nrebuild = 0; e->launches = 0; e->nreact_pass = 0; e->nreact_events = 0; e->pair_ms = 0; e->pair_launches = 0;
    return CLB_OK;
}
extern "C" int clb_device_ptr(clb_engine* e, int which, void** ptr, int64_t* n_out) {
    if (!e || !ptr) return CLB_ERR_ARG;
    switch (which) {
        case 0: *ptr = e->pos.p; break;
        case 1: *ptr = e->vel.p; break;
        case 2: *ptr = e->force.p; break;
        default: return e->fail(CLB_ERR_ARG, "unknown buffer");
    }
    if (n_out) *n_out = which == 2 ? e->ncap : e->nstored;
    return CLB_OK;
}
extern "C" int clb_stream(clb_engine* e, void** s) { if (!e || !s) return CLB_ERR_ARG; *s = (void*)e->stream; return CLB_OK; }

void clb_engine::free_all() {
    for (auto& l : lists) l.d.release();
    // device buffers are plain handles (DevBuf has no destructor): release them here so that a process creating many
    // engines (tests, the e2e leg of bench.py) does not accumulate device memory
    pos.release(); pos2.release(); xref.release(); vel.release(); vel2.release(); slot.release(); slot2.release(); id2idx.release();
    image.release(); resid.release(); mol.release(); wslot.release(); force.release(); charge.release(); key.release(); key2.release();
    val.release(); val2.release(); cell_start.release(); cubtmp.release(); cubtmp2.release(); stage.release(); partial.release();
    partial_u64.release(); nl_entries.release(); nl_count.release(); excl_pairs.release(); excl_off.release(); excl_ids.release();
    ekey.release(); ekey2.release(); eval.release(); d_pd.release(); d_pd_e.release(); d_plj.release(); d_pe.release(); d_tm.release();
    d_tm_e.release(); d_frows.release(); d_erows.release(); d_rows2.release(); d_pd2.release(); d_bdesc.release(); d_bpots.release();
    d_btm.release(); d_bcf.release(); d_bce.release(); d_list_ptrs.release(); term_off.release(); term_meta.release(); term_tuple.release();
    tkey.release(); tkey2.release(); tval.release(); tval2.release(); rt_off.release(); rt_cnt.release(); rt_meta.release(); rt_mem.release();
    react_free();
    for (auto ev : pair_events) cudaEventDestroy(ev);
    for (auto ev : pair_events2) cudaEventDestroy(ev);
    for (int i = 0; i < CLB_NBUCKET; ++i) { cudaEventDestroy(ev_a[i]); cudaEventDestroy(ev_b[i]); }
    if (d_ctl) cudaFree(d_ctl);
    if (h_ctl) cudaFreeHost(h_ctl);
    if (d_scalar) cudaFree(d_scalar);
    if (h_scalar) cudaFreeHost(h_scalar);
    comm_destroy();
    if (stream) cudaStreamDestroy(stream);
}

#include "engine_react.inl"
#include "engine_comm.inl"
