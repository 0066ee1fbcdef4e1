// engine_comm.inl -- multi-GPU slab decomposition over NCCL (included by engine.cu).
//
// Replaces storage.DomainDecomposition's MPI node grid + ghost exchange (src/start_simulation.py:152-163;
// [EXT] storage/DomainDecomposition.cpp): one engine per GPU/process, the box is cut into slabs of whole cell
// planes along z.  Rank r owns planes [cz0, cz0+nczl) and stores one ghost plane on either side:
//
//      sorted particle arrays:   [ owned (cell-sorted) | upper ghost plane | lower ghost plane ]
//                                  0 ............ own1   own1 ...... +n_hi   ........ nstored
//
//   * positions are lattice integers modulo 2^32, so ghosts need no coordinate shift;
//   * a neighbour's boundary plane is ONE contiguous range of its sorted arrays, and it arrives in the
//     sender's cell order, so the per-step halo is two contiguous ncclSend/ncclRecv pairs per rank and
//     the ghost planes never need sorting;
//   * forces are owner-computes over full lists -> there is no reverse (force) halo;
//   * the resort criterion is the global maximum displacement: one 4-byte ncclAllReduce(max) per step,
//     so every rank stalls at the same step and the host-side chunk loop stays in lock-step;
//   * topology (tuple lists, bond graph, exclusions, molecule ids, type|state words) is replicated; the
//     reaction candidates of all ranks are all-gathered and every rank runs the same deterministic
//     conflict pass (replaces the reference's three neighbour-rank multimap exchanges).
//
// NCCL is resolved with dlopen at clb_comm_init time (inside a torch process this is torch's own
// libnccl.so.2), so the library loads on machines without NCCL and single-GPU use never touches it.
#include <dlfcn.h>
#include <nccl.h>

struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
    bool load() {
        if (h) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) { h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
        if (!h) { err = std::string("cannot load libnccl: ") + dlerror(); return false; }
#define CLB_NCCL_SYM(field, sym) *(void**)(&field) = dlsym(h, sym); if (!field) { err = std::string("libnccl lacks ") + sym; h = nullptr; return false; }
        CLB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId") CLB_NCCL_SYM(CommInitRank, "ncclCommInitRank") CLB_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        CLB_NCCL_SYM(Send, "ncclSend") CLB_NCCL_SYM(Recv, "ncclRecv") CLB_NCCL_SYM(AllReduce, "ncclAllReduce") CLB_NCCL_SYM(AllGather, "ncclAllGather")
        CLB_NCCL_SYM(Broadcast, "ncclBroadcast") CLB_NCCL_SYM(GroupStart, "ncclGroupStart") CLB_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        CLB_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef CLB_NCCL_SYM
        return true;
    }
};
static NcclApi g_nccl;

#define NC(call)                                                                                              \
    do {                                                                                                      \
        ncclResult_t _r = (call);                                                                             \
        if (_r != ncclSuccess) return e->fail(CLB_ERR_COMM, "%s failed: %s", #call, g_nccl.GetErrorString(_r)); \
    } while (0)

// migrating particle: everything that is stored per sorted index plus the owner-only per-slot image counters
struct ClbMig { int4 p; ClbVel v; int slot, ix, iy, iz; };

// ---- peer mailbox layout (all offsets in bytes from the start of the block) ---------------------------------------------
#define CLB_MAX_RANKS 16
struct ClbMailHdr {                       // written by peers
    unsigned long long halo_flag[2];      // [0]: from the rank above (its bottom plane = my upper ghost), [1]: from the rank below
    unsigned long long ghost_flag[2];
    unsigned long long mig_flag[2];
    unsigned long long disp[CLB_MAX_RANKS][2];   // per source rank and parity: (epoch << 32) | float bits of its max |dx|^2
    int ghost_cnt[2];
    int mig_cnt[2];
    int pad[4];
};
struct ClbHalo {                          // device-resident description of what this rank sends / stores
    int send_lo0, send_lo1, send_hi0, send_hi1;   // bottom / top owned plane as index ranges of the sorted arrays
    int n_hi, n_lo, own1, pad;
};
struct ClbMailLayout { size_t halo, gpos, gslot, mig, total; int plane_cap, mig_cap; };
static ClbMailLayout mail_layout(int plane_cap, int mig_cap) {
    ClbMailLayout L; L.plane_cap = plane_cap; L.mig_cap = mig_cap;
    size_t o = (sizeof(ClbMailHdr) + 255) & ~(size_t)255;
    L.halo = o;  o += (size_t)4 * plane_cap * sizeof(int4);          // [parity][dir][plane_cap]
    L.gpos = o;  o += (size_t)2 * plane_cap * sizeof(int4);          // [dir][plane_cap]
    L.gslot = o; o += (size_t)2 * plane_cap * sizeof(int);
    o = (o + 255) & ~(size_t)255;
    L.mig = o;   o += (size_t)2 * mig_cap * sizeof(ClbMig);          // [dir][mig_cap]
    L.total = (o + 255) & ~(size_t)255;
    return L;
}

struct clb_engine::CommDev {
    ncclComm_t comm = nullptr;
    int up = 0, dn = 0;                        // ranks owning the planes above / below
    int n_hi = 0, n_lo = 0;                    // ghost counts: upper plane (from `up`), lower plane (from `dn`)
    int send_lo0 = 0, send_lo1 = 0;            // my bottom plane [send_lo0, send_lo1) -> dn's upper ghost
    int send_hi0 = 0, send_hi1 = 0;            // my top plane                         -> up's lower ghost
    DevBuf<int> cnt;                           // small device scratch for count exchanges
    int* h_cnt = nullptr;                      // pinned mirror
    DevBuf<ClbMig> mig_send, mig_recv;
    DevBuf<int> mkey, mkey2, mval, mval2, moff;
    DevBuf<unsigned char> gat;                 // all-gather staging
    DevBuf<double> red;                        // host-value reductions
    DevBuf<long long> sizes;
    std::vector<int> plane0, planes;           // cz0 and nczl of every rank
    // peer-memory path (round 2): every rank exports ONE mailbox block with cudaIpc; neighbours write boundary planes, migrants
    // and the per-step displacement maximum straight into it over NVLink and raise an epoch flag -- no NCCL on the step path
    int peer_ok = 0;
    unsigned char* mb = nullptr;               // my mailbox (device memory, exported)
    size_t mb_bytes = 0;
    int plane_cap = 0, mig_cap = 0;
    std::vector<void*> mb_peer;                // mailbox of every rank as mapped into this process (own rank: mb)
    DevBuf<unsigned char*> d_mb_all;           // the same table on the device
    ClbHalo* d_halo = nullptr;                 // send ranges and ghost counts of the current decomposition (device)
    ClbHalo* h_halo = nullptr;                 // pinned mirror
    unsigned long long mig_epoch = 0, ghost_epoch = 0;
};

extern "C" int clb_nccl_unique_id(void* id128_out) {
    if (!id128_out) return CLB_ERR_ARG;
    if (!g_nccl.load()) { g_create_error = g_nccl.err; return CLB_ERR_COMM; }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) { g_create_error = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r); return CLB_ERR_COMM; }
    memcpy(id128_out, &id, 128);
    return CLB_OK;
}

// planes owned by rank r of nr: ncz/nr each, the first ncz%nr ranks one more (host and tests share this rule:
// chemlab_b200/engine.py::slab_planes)
static void slab_planes(int ncz, int nr, int r, int* cz0, int* nczl) {
    int base = ncz / nr, rem = ncz % nr;
    *nczl = base + (r < rem ? 1 : 0);
    *cz0 = r * base + std::min(r, rem);
}

extern "C" int clb_comm_init(clb_engine* e, int rank, int nranks, const void* nccl_id128) {
    if (!e || nranks < 1 || rank < 0 || rank >= nranks) return e ? e->fail(CLB_ERR_ARG, "clb_comm_init: bad rank/nranks") : CLB_ERR_ARG;
    if (nranks == 1) return CLB_OK;
    if (!nccl_id128) return e->fail(CLB_ERR_ARG, "clb_comm_init: NULL nccl id");
    if (e->n > 0) return e->fail(CLB_ERR_STATE, "clb_comm_init must precede clb_set_particles");
    if (e->cd) return e->fail(CLB_ERR_STATE, "clb_comm_init called twice");
    cudaSetDevice(e->device);
    if (!g_nccl.load()) return e->fail(CLB_ERR_COMM, "%s", g_nccl.err.c_str());
    ClbGrid& g = e->grid;
    auto* cd = new clb_engine::CommDev();
    cd->plane0.resize(nranks); cd->planes.resize(nranks);
    for (int r = 0; r < nranks; ++r) {
        slab_planes(g.ncz, nranks, r, &cd->plane0[r], &cd->planes[r]);
        // a rank must not see the same foreign plane as both its upper and its lower ghost
        if (cd->planes[r] < 1 || g.ncz - cd->planes[r] < 2) {
            delete cd;
            return e->fail(CLB_ERR_UNSUPPORTED, "%d cell planes along z cannot be split over %d ranks (every rank needs >= 1 owned and >= 2 foreign planes)", g.ncz, nranks);
        }
    }
    ncclUniqueId id; memcpy(&id, nccl_id128, 128);
    ncclResult_t r = g_nccl.CommInitRank(&cd->comm, nranks, id, rank);
    if (r != ncclSuccess) { delete cd; return e->fail(CLB_ERR_COMM, "ncclCommInitRank: %s", g_nccl.GetErrorString(r)); }
    cd->up = (rank + 1) % nranks; cd->dn = (rank + nranks - 1) % nranks;
    e->cd = cd; e->rank = rank; e->nranks = nranks;
    g.cz0 = cd->plane0[rank]; g.nczl = cd->planes[rank]; g.zoff = g.cz0; g.nplanes = g.nczl + 2; g.ghost = 1;
    e->set_block_cells(g.bx);
    CK(cd->cnt.ensure(64));
    CK(cudaMallocHost(&cd->h_cnt, 64 * sizeof(int)));
    // per-step communication (global max displacement + position halo) runs beside the interior-plane force kernel
    CK(cudaStreamCreateWithFlags(&e->comm_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&e->ev_int, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&e->ev_comm, cudaEventDisableTiming));
    // NCCL builds its peer connections lazily at the first use of each pattern (100s of ms each): do that here, once, so
    // that the first rebuild / reaction pass of a run does not pay for it
    CK(cudaMemsetAsync(cd->cnt.p, 0, 64 * sizeof(int), e->stream));
    CK(cd->sizes.ensure(2 * (size_t)nranks));
    for (int pass = 0; pass < 2; ++pass) {
        cudaStream_t st = pass == 0 ? e->stream : e->comm_stream;
        NC(g_nccl.AllReduce(cd->cnt.p, cd->cnt.p, 1, ncclUint32, ncclMax, cd->comm, st));
        NC(g_nccl.GroupStart());
        NC(g_nccl.Send(cd->cnt.p + 1, 1, ncclInt32, cd->up, cd->comm, st));
        NC(g_nccl.Send(cd->cnt.p + 2, 1, ncclInt32, cd->dn, cd->comm, st));
        NC(g_nccl.Recv(cd->cnt.p + 4, 1, ncclInt32, cd->dn, cd->comm, st));
        NC(g_nccl.Recv(cd->cnt.p + 5, 1, ncclInt32, cd->up, cd->comm, st));
        NC(g_nccl.GroupEnd());
        CK(cudaStreamSynchronize(st));
    }
    NC(g_nccl.AllGather(cd->sizes.p + nranks + rank, cd->sizes.p, 1, ncclInt64, cd->comm, e->stream));
    NC(g_nccl.GroupStart());
    for (int r2 = 0; r2 < nranks; ++r2) NC(g_nccl.Broadcast(cd->cnt.p + 16 + (r2 == rank ? 0 : 1), cd->cnt.p + 18, 4, ncclChar, r2, cd->comm, e->stream));
    NC(g_nccl.GroupEnd());
    NC(g_nccl.AllReduce(cd->cnt.p + 24, cd->cnt.p + 24, 2, ncclDouble, ncclSum, cd->comm, e->stream));
    // ... and the large-message protocol of the collective read-back (get_particles sums an [n x 15] matrix over the ranks): its
    // channels are set up at the first large all-reduce (110 ms inside the first download at 8 ranks, profiles/r2q_n8_default.json)
    CK(cd->red.ensure((size_t)1 << 21));
    CK(cudaMemsetAsync(cd->red.p, 0, ((size_t)1 << 21) * sizeof(double), e->stream));
    NC(g_nccl.AllReduce(cd->red.p, cd->red.p, (size_t)1 << 21, ncclDouble, ncclSum, cd->comm, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return CLB_OK;
}

void clb_engine::comm_destroy() {
    if (!cd) return;
    comm_peer_teardown();
    if (cd->d_halo) cudaFree(cd->d_halo);
    if (cd->h_halo) cudaFreeHost(cd->h_halo);
    if (cd->comm) g_nccl.CommDestroy(cd->comm);
    if (cd->h_cnt) cudaFreeHost(cd->h_cnt);
    if (comm_stream) { cudaStreamSynchronize(comm_stream); cudaStreamDestroy(comm_stream); comm_stream = nullptr; }
    if (ev_int) cudaEventDestroy(ev_int);
    if (ev_comm) cudaEventDestroy(ev_comm);
    delete cd; cd = nullptr;
}

// ---- migration (before the owned sort) ------------------------------------------------------------------
// destination of every owned particle after the drift: 0 stay, 1 -> up, 2 -> dn
__global__ void k_mig_classify(int n, const int4* __restrict__ pos, ClbGrid g, int* __restrict__ key, int* __restrict__ val, ClbCtl* ctl) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cz = __umulhi((unsigned)pos[i].z, (unsigned)g.ncz);
    int l = local_plane(g, cz);
    if (l < 0) { atomicOr(&ctl->err, CLB_EF_PARTNER_LOST); l = 0; }   // moved more than one plane: cannot happen with skin/2
    key[i] = l < g.nczl ? 0 : (l == g.nczl ? 1 : 2);
    val[i] = i;
}
__global__ void k_mig_pack(int n, const int* __restrict__ perm, const int4* __restrict__ pos, const ClbVel* __restrict__ vel,
                           const int* __restrict__ slot, const int* __restrict__ image, int* __restrict__ id2idx, ClbMig* __restrict__ out) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int i = perm[k];
    ClbMig m; m.p = pos[i]; m.v = vel[i]; m.slot = slot[i];
    m.ix = image[3 * m.slot]; m.iy = image[3 * m.slot + 1]; m.iz = image[3 * m.slot + 2];
    out[k] = m;
}
// new owned set = stayers (in their previous order) followed by the immigrants
__global__ void k_mig_compact(int nstay, const int* __restrict__ perm, const int4* __restrict__ pos_in, const ClbVel* __restrict__ vel_in,
                              const int* __restrict__ slot_in, int4* __restrict__ pos, ClbVel* __restrict__ vel, int* __restrict__ slot) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nstay) return;
    int i = perm[k];
    pos[k] = pos_in[i]; vel[k] = vel_in[i]; slot[k] = slot_in[i];
}
__global__ void k_mig_unpack(int n, const ClbMig* __restrict__ in, int base, int4* __restrict__ pos, ClbVel* __restrict__ vel,
                             int* __restrict__ slot, int* __restrict__ image) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    ClbMig m = in[k];
    pos[base + k] = m.p; vel[base + k] = m.v; slot[base + k] = m.slot;
    image[3 * m.slot] = m.ix; image[3 * m.slot + 1] = m.iy; image[3 * m.slot + 2] = m.iz;
}
__global__ void k_store4(int* dst, int a, int b, int c, int d) { dst[0] = a; dst[1] = b; dst[2] = c; dst[3] = d; }
__global__ void k_mig_counts(const int* __restrict__ off, int* __restrict__ cnt) {   // off = lower bounds of keys 0,1,2,3
    cnt[0] = off[1];            // stayers
    cnt[1] = off[2] - off[1];   // to up
    cnt[2] = off[3] - off[2];   // to dn
}

int clb_engine::comm_migrate() {
    clb_engine* e = this;
    CommDev& c = *cd;
    const int no = own1;
    CK(c.mkey.ensure(ncap)); CK(c.mkey2.ensure(ncap)); CK(c.mval.ensure(ncap)); CK(c.mval2.ensure(ncap)); CK(c.moff.ensure(8));
    k_mig_classify<<<ceil_div(std::max(no, 1), 256), 256, 0, stream>>>(no, pos.p, grid, c.mkey.p, c.mval.p, d_ctl);
    size_t tb = cubtmp.n;
    cub::DeviceRadixSort::SortPairs(cubtmp.p, tb, c.mkey.p, c.mkey2.p, c.mval.p, c.mval2.p, no, 0, 2, stream);
    k_lower_bounds<<<1, 32, 0, stream>>>(no, c.mkey2.p, 3, c.moff.p);
    k_mig_counts<<<1, 1, 0, stream>>>(c.moff.p, c.cnt.p);
    // exchange the counts: cnt[1] (to up) -> up's cnt[4] (from dn); cnt[2] (to dn) -> dn's cnt[5] (from up)
    NC(g_nccl.GroupStart());
    NC(g_nccl.Send(c.cnt.p + 1, 1, ncclInt32, c.up, c.comm, stream));
    NC(g_nccl.Send(c.cnt.p + 2, 1, ncclInt32, c.dn, c.comm, stream));
    NC(g_nccl.Recv(c.cnt.p + 4, 1, ncclInt32, c.dn, c.comm, stream));
    NC(g_nccl.Recv(c.cnt.p + 5, 1, ncclInt32, c.up, c.comm, stream));
    NC(g_nccl.GroupEnd());
    CK(cudaMemcpyAsync(c.h_cnt, c.cnt.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    const int nstay = c.h_cnt[0], to_up = c.h_cnt[1], to_dn = c.h_cnt[2], from_dn = c.h_cnt[4], from_up = c.h_cnt[5];
    const int nnew = nstay + from_dn + from_up;
    if (nnew + 64 > ncap) return fail(CLB_ERR_RANGE, "rank %d: %d owned particles exceed the local capacity %d", rank, nnew, ncap);
    CK(c.mig_send.ensure((size_t)to_up + to_dn + 1)); CK(c.mig_recv.ensure((size_t)from_dn + from_up + 1));
    if (to_up + to_dn > 0)
        k_mig_pack<<<ceil_div(to_up + to_dn, 256), 256, 0, stream>>>(to_up + to_dn, c.mval2.p + nstay, pos.p, vel.p, slot.p, image.p, id2idx.p, c.mig_send.p);
    NC(g_nccl.GroupStart());
    if (to_up) NC(g_nccl.Send(c.mig_send.p, (size_t)to_up * sizeof(ClbMig), ncclChar, c.up, c.comm, stream));
    if (to_dn) NC(g_nccl.Send(c.mig_send.p + to_up, (size_t)to_dn * sizeof(ClbMig), ncclChar, c.dn, c.comm, stream));
    if (from_dn) NC(g_nccl.Recv(c.mig_recv.p, (size_t)from_dn * sizeof(ClbMig), ncclChar, c.dn, c.comm, stream));
    if (from_up) NC(g_nccl.Recv(c.mig_recv.p + from_dn, (size_t)from_up * sizeof(ClbMig), ncclChar, c.up, c.comm, stream));
    NC(g_nccl.GroupEnd());
    if (nstay > 0) k_mig_compact<<<ceil_div(nstay, 256), 256, 0, stream>>>(nstay, c.mval2.p, pos.p, vel.p, slot.p, pos2.p, vel2.p, slot2.p);
    if (from_dn + from_up > 0)
        k_mig_unpack<<<ceil_div(from_dn + from_up, 256), 256, 0, stream>>>(from_dn + from_up, c.mig_recv.p, nstay, pos2.p, vel2.p, slot2.p, image.p);
    std::swap(pos, pos2); std::swap(vel, vel2); std::swap(slot, slot2);      // whole handles (pointer, capacity and block size)
    own0 = 0; own1 = nnew; nstored = nnew;
    CK(cudaGetLastError());
    launches += 6;
    return CLB_OK;
}

// ---- ghost planes (after the owned sort) ----------------------------------------------------------------
__global__ void k_ghost_finish(int n0, int n1, const int4* __restrict__ pos, const int* __restrict__ slot, ClbGrid g,
                               int* __restrict__ key, int* __restrict__ id2idx) {
    int i = n0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1) return;
    int4 p = pos[i];
    int cx = __umulhi((unsigned)p.x, (unsigned)g.ncx), cy = __umulhi((unsigned)p.y, (unsigned)g.ncy), cz = __umulhi((unsigned)p.z, (unsigned)g.ncz);
    int lz = local_plane(g, cz);
    key[i] = lz < 0 ? g.ncell - 1 : (lz * g.ncy + cy) * g.ncx + cx;
    id2idx[slot[i]] = i;
}
__global__ void k_plane_bounds(const int* __restrict__ cell_start, int plane_cells, int nczl, int nown, int* __restrict__ cnt) {
    cnt[0] = 0; cnt[1] = cell_start[plane_cells];                 // bottom owned plane
    cnt[2] = cell_start[(nczl - 1) * plane_cells]; cnt[3] = nown; // top owned plane
    cnt[8] = cnt[1] - cnt[0]; cnt[9] = cnt[3] - cnt[2];
}

// called by rebuild() after the owned particles are cell-sorted (keys in key2, cell_start over the owned part)
int clb_engine::comm_exchange_ghosts() {
    clb_engine* e = this;
    CommDev& c = *cd;
    const int no = own1;
    k_plane_bounds<<<1, 1, 0, stream>>>(cell_start.p, grid.ncx * grid.ncy, grid.nczl, no, c.cnt.p);
    // my bottom-plane count -> dn (its n_hi); my top-plane count -> up (its n_lo)
    NC(g_nccl.GroupStart());
    NC(g_nccl.Send(c.cnt.p + 8, 1, ncclInt32, c.dn, c.comm, stream));
    NC(g_nccl.Send(c.cnt.p + 9, 1, ncclInt32, c.up, c.comm, stream));
    NC(g_nccl.Recv(c.cnt.p + 10, 1, ncclInt32, c.up, c.comm, stream));
    NC(g_nccl.Recv(c.cnt.p + 11, 1, ncclInt32, c.dn, c.comm, stream));
    NC(g_nccl.GroupEnd());
    CK(cudaMemcpyAsync(c.h_cnt, c.cnt.p, 12 * sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    c.send_lo0 = c.h_cnt[0]; c.send_lo1 = c.h_cnt[1]; c.send_hi0 = c.h_cnt[2]; c.send_hi1 = c.h_cnt[3];
    c.n_hi = c.h_cnt[10]; c.n_lo = c.h_cnt[11];
    const int ns = no + c.n_hi + c.n_lo;
    if (ns + 64 > ncap) return fail(CLB_ERR_RANGE, "rank %d: %d stored particles (owned + ghosts) exceed the local capacity %d", rank, ns, ncap);
    NC(g_nccl.GroupStart());
    NC(g_nccl.Send(pos.p + c.send_lo0, (size_t)(c.send_lo1 - c.send_lo0) * sizeof(int4), ncclChar, c.dn, c.comm, stream));
    NC(g_nccl.Send(pos.p + c.send_hi0, (size_t)(c.send_hi1 - c.send_hi0) * sizeof(int4), ncclChar, c.up, c.comm, stream));
    NC(g_nccl.Recv(pos.p + no, (size_t)c.n_hi * sizeof(int4), ncclChar, c.up, c.comm, stream));
    NC(g_nccl.Recv(pos.p + no + c.n_hi, (size_t)c.n_lo * sizeof(int4), ncclChar, c.dn, c.comm, stream));
    NC(g_nccl.Send(slot.p + c.send_lo0, (size_t)(c.send_lo1 - c.send_lo0), ncclInt32, c.dn, c.comm, stream));
    NC(g_nccl.Send(slot.p + c.send_hi0, (size_t)(c.send_hi1 - c.send_hi0), ncclInt32, c.up, c.comm, stream));
    NC(g_nccl.Recv(slot.p + no, (size_t)c.n_hi, ncclInt32, c.up, c.comm, stream));
    NC(g_nccl.Recv(slot.p + no + c.n_hi, (size_t)c.n_lo, ncclInt32, c.dn, c.comm, stream));
    NC(g_nccl.GroupEnd());
    nstored = ns;
    if (ns > no) k_ghost_finish<<<ceil_div(ns - no, 256), 256, 0, stream>>>(no, ns, pos.p, slot.p, grid, key2.p, id2idx.p);
    CK(cudaGetLastError());
    launches += 4;
    return CLB_OK;
}

// per-step forward halo: boundary-plane positions (with their type|state word) -> the neighbours' ghost planes
int clb_engine::comm_halo_positions(cudaStream_t st) {
    clb_engine* e = this;
    CommDev& c = *cd;
    NC(g_nccl.GroupStart());
    NC(g_nccl.Send(pos.p + c.send_lo0, (size_t)(c.send_lo1 - c.send_lo0) * sizeof(int4), ncclChar, c.dn, c.comm, st));
    NC(g_nccl.Send(pos.p + c.send_hi0, (size_t)(c.send_hi1 - c.send_hi0) * sizeof(int4), ncclChar, c.up, c.comm, st));
    NC(g_nccl.Recv(pos.p + own1, (size_t)c.n_hi * sizeof(int4), ncclChar, c.up, c.comm, st));
    NC(g_nccl.Recv(pos.p + own1 + c.n_hi, (size_t)c.n_lo * sizeof(int4), ncclChar, c.dn, c.comm, st));
    NC(g_nccl.GroupEnd());
    ++launches;
    return CLB_OK;
}
// both per-step operations in ONE NCCL group (one host-side launch sequence, no gap between the two kernels)
int clb_engine::comm_step(cudaStream_t st) {
    clb_engine* e = this;
    CommDev& c = *cd;
    NC(g_nccl.GroupStart());
    NC(g_nccl.AllReduce(&d_ctl->maxdisp2_bits, &d_ctl->maxdisp2_bits, 1, ncclUint32, ncclMax, c.comm, st));
    NC(g_nccl.Send(pos.p + c.send_lo0, (size_t)(c.send_lo1 - c.send_lo0) * sizeof(int4), ncclChar, c.dn, c.comm, st));
    NC(g_nccl.Send(pos.p + c.send_hi0, (size_t)(c.send_hi1 - c.send_hi0) * sizeof(int4), ncclChar, c.up, c.comm, st));
    NC(g_nccl.Recv(pos.p + own1, (size_t)c.n_hi * sizeof(int4), ncclChar, c.up, c.comm, st));
    NC(g_nccl.Recv(pos.p + own1 + c.n_hi, (size_t)c.n_lo * sizeof(int4), ncclChar, c.dn, c.comm, st));
    NC(g_nccl.GroupEnd());
    launches += 2;
    return CLB_OK;
}
// global maximum displacement of this step (float bits of a non-negative number order like unsigned integers)
int clb_engine::comm_max_displacement(cudaStream_t st) {
    clb_engine* e = this;
    NC(g_nccl.AllReduce(&d_ctl->maxdisp2_bits, &d_ctl->maxdisp2_bits, 1, ncclUint32, ncclMax, cd->comm, st));
    ++launches;
    return CLB_OK;
}

// =====================================================================================================================
// Peer-memory path: boundary planes, migrants and the displacement maximum travel as plain NVLink stores into the
// neighbour's mailbox, followed by an epoch flag (system-scope fence in between).  One push kernel and one receive kernel
// per step replace the NCCL group (all-reduce + 2 send/recv pairs, ~40 us of launch/protocol latency at 8 ranks).
// Replaces [EXT] storage.updateGhosts() / DomainDecomposition::doGhostCommunication for the per-step halo.
__device__ __forceinline__ unsigned long long ld_flag(const unsigned long long* p) { return *(const volatile unsigned long long*)p; }
__device__ __forceinline__ unsigned long long clb_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
// wait until *flag >= epoch (written by a peer over NVLink); gives up after ~20 s so that a dead neighbour cannot hang the GPU
__device__ __forceinline__ bool wait_flag(const unsigned long long* flag, unsigned long long epoch) {
    const unsigned long long t0 = clb_ns();
    while (ld_flag(flag) < epoch) {
        __nanosleep(64);
        if (clb_ns() - t0 > 20000000000ull) return false;
    }
    return true;
}
struct ClbPeers { unsigned char* self; unsigned char* up; unsigned char* dn; unsigned char* const* all; int rank, nranks; ClbMailLayout L; };

// last block of a push kernel: everything written by the grid is visible system-wide -> raise the flags
__device__ __forceinline__ bool grid_last_block(ClbCtl* ctl) {
    __shared__ int s_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) { unsigned d = atomicAdd(&ctl->comm_done, 1u); s_last = (d == gridDim.x - 1); if (s_last) ctl->comm_done = 0; }
    __syncthreads();
    return s_last != 0;
}

__global__ void __launch_bounds__(256) k_halo_push(ClbCtl* ctl, const ClbHalo* __restrict__ H, const int4* __restrict__ pos, ClbPeers P) {
    if (*(volatile int*)&ctl->stall) return;
    const unsigned long long epoch = ctl->halo_epoch + 1;
    const int par = (int)(epoch & 1ull);
    const ClbHalo h = *H;
    const int nlo = h.send_lo1 - h.send_lo0, nhi = h.send_hi1 - h.send_hi0;
    // my bottom plane -> upper ghost of the rank below (its slot 0 "from above"); my top plane -> lower ghost of the rank above (slot 1)
    int4* dst_dn = reinterpret_cast<int4*>(P.dn + P.L.halo) + ((size_t)par * 2 + 0) * P.L.plane_cap;
    int4* dst_up = reinterpret_cast<int4*>(P.up + P.L.halo) + ((size_t)par * 2 + 1) * P.L.plane_cap;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nlo + nhi; i += gridDim.x * blockDim.x) {
        if (i < nlo) dst_dn[i] = pos[h.send_lo0 + i];
        else dst_up[i - nlo] = pos[h.send_hi0 + (i - nlo)];
    }
    if (grid_last_block(ctl)) {
        if (threadIdx.x < P.nranks) {
            ClbMailHdr* hd = reinterpret_cast<ClbMailHdr*>(P.all[threadIdx.x]);
            *(volatile unsigned long long*)&hd->disp[P.rank][par] = (epoch << 32) | (unsigned long long)ctl->maxdisp2_bits;
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            *(volatile unsigned long long*)&reinterpret_cast<ClbMailHdr*>(P.dn)->halo_flag[0] = epoch;
            *(volatile unsigned long long*)&reinterpret_cast<ClbMailHdr*>(P.up)->halo_flag[1] = epoch;
        }
    }
}
// wait for both neighbours, copy the landed planes behind the owned range, reduce the displacement maximum over all ranks
__global__ void __launch_bounds__(256) k_halo_recv(ClbCtl* ctl, const ClbHalo* __restrict__ H, int4* __restrict__ pos, ClbPeers P) {
    if (*(volatile int*)&ctl->stall) return;
    const unsigned long long epoch = ctl->halo_epoch + 1;
    const int par = (int)(epoch & 1ull);
    const ClbMailHdr* hd = reinterpret_cast<const ClbMailHdr*>(P.self);
    __shared__ int s_ok;
    if (threadIdx.x == 0) s_ok = wait_flag(&hd->halo_flag[0], epoch) && wait_flag(&hd->halo_flag[1], epoch);
    __syncthreads();
    if (!s_ok) { if (threadIdx.x == 0) atomicOr(&ctl->err, CLB_EF_COMM_TIMEOUT); return; }
    __threadfence_system();
    const ClbHalo h = *H;
    const int4* src_hi = reinterpret_cast<const int4*>(P.self + P.L.halo) + ((size_t)par * 2 + 0) * P.L.plane_cap;
    const int4* src_lo = reinterpret_cast<const int4*>(P.self + P.L.halo) + ((size_t)par * 2 + 1) * P.L.plane_cap;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < h.n_hi + h.n_lo; i += gridDim.x * blockDim.x)
        pos[h.own1 + i] = i < h.n_hi ? __ldcg(src_hi + i) : __ldcg(src_lo + (i - h.n_hi));
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        unsigned m = 0; bool ok = true;
        for (int r = threadIdx.x; r < P.nranks; r += 32) {
            const unsigned long long* w = &hd->disp[r][par];
            const unsigned long long t0 = clb_ns();
            unsigned long long v;
            while (((v = ld_flag(w)) >> 32) < (epoch & 0xffffffffull)) { __nanosleep(64); if (clb_ns() - t0 > 20000000000ull) { ok = false; break; } }
            m = max(m, (unsigned)(v & 0xffffffffull));
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
        if (!__all_sync(0xffffffffu, ok)) { if (threadIdx.x == 0) atomicOr(&ctl->err, CLB_EF_COMM_TIMEOUT); }
        if (threadIdx.x == 0) ctl->maxdisp2_bits = m;       // global maximum: every rank takes the same resort decision
    }
}
// resort check of the peer path: the exchange of this step has happened on every rank, so the epoch advances even when the
// step then stalls
__global__ void k_check_resort_peer(ClbCtl* ctl, int criterion, double half_skin, int step_index) {
    if (ctl->stall) return;
    ctl->halo_epoch += 1;
    float m2 = __uint_as_float(ctl->maxdisp2_bits);
    ctl->maxdisp2_bits = 0u;
    bool need;
    if (criterion == 0) { ctl->accum_maxdist += sqrt((double)m2); need = ctl->accum_maxdist > half_skin; }
    else need = sqrt((double)m2) > half_skin;
    if (need || ctl->force_rebuild) { ctl->stall = 1; ctl->stall_step = step_index; }
    else ctl->steps_ok += 1;
}

// ---- rebuild: migrants and ghost planes through the mailboxes --------------------------------------------------------
__global__ void __launch_bounds__(256) k_mig_push(ClbCtl* ctl, const int* __restrict__ cnt, const int* __restrict__ perm, const int4* __restrict__ pos,
                                                  const ClbVel* __restrict__ vel, const int* __restrict__ slot, const int* __restrict__ image,
                                                  ClbPeers P, unsigned long long epoch) {
    const int nstay = cnt[0], to_up = cnt[1], to_dn = cnt[2];
    if (to_up > P.L.mig_cap || to_dn > P.L.mig_cap) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&ctl->err, CLB_EF_COMM_OVERFLOW); }
    const int nu = min(to_up, P.L.mig_cap), nd = min(to_dn, P.L.mig_cap);
    ClbMig* dst_up = reinterpret_cast<ClbMig*>(P.up + P.L.mig) + (size_t)1 * P.L.mig_cap;     // slot 1 of the rank above: "from below"
    ClbMig* dst_dn = reinterpret_cast<ClbMig*>(P.dn + P.L.mig) + (size_t)0 * P.L.mig_cap;     // slot 0 of the rank below: "from above"
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nu + nd; k += gridDim.x * blockDim.x) {
        const bool up = k < nu;
        const int i = perm[nstay + (up ? k : to_up + (k - nu))];
        ClbMig m; m.p = pos[i]; m.v = vel[i]; m.slot = slot[i];
        m.ix = image[3 * m.slot]; m.iy = image[3 * m.slot + 1]; m.iz = image[3 * m.slot + 2];
        if (up) dst_up[k] = m; else dst_dn[k - nu] = m;
    }
    if (grid_last_block(ctl) && threadIdx.x == 0) {
        ClbMailHdr* hu = reinterpret_cast<ClbMailHdr*>(P.up); ClbMailHdr* hdn = reinterpret_cast<ClbMailHdr*>(P.dn);
        *(volatile int*)&hu->mig_cnt[1] = nu; *(volatile int*)&hdn->mig_cnt[0] = nd;
        __threadfence_system();
        *(volatile unsigned long long*)&hu->mig_flag[1] = epoch; *(volatile unsigned long long*)&hdn->mig_flag[0] = epoch;
    }
}
__global__ void k_mig_wait(ClbCtl* ctl, ClbPeers P, unsigned long long epoch, int* __restrict__ cnt) {
    const ClbMailHdr* hd = reinterpret_cast<const ClbMailHdr*>(P.self);
    if (!(wait_flag(&hd->mig_flag[0], epoch) && wait_flag(&hd->mig_flag[1], epoch))) { atomicOr(&ctl->err, CLB_EF_COMM_TIMEOUT); cnt[4] = 0; cnt[5] = 0; return; }
    __threadfence_system();
    cnt[4] = *(const volatile int*)&hd->mig_cnt[1];     // from below
    cnt[5] = *(const volatile int*)&hd->mig_cnt[0];     // from above
}
__global__ void k_mig_unpack_peer(int from_dn, int from_up, ClbPeers P, int base, int4* __restrict__ pos, ClbVel* __restrict__ vel,
                                  int* __restrict__ slot, int* __restrict__ image) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= from_dn + from_up) return;
    const ClbMig* src = reinterpret_cast<const ClbMig*>(P.self + P.L.mig) + (k < from_dn ? (size_t)1 * P.L.mig_cap + k : (size_t)0 * P.L.mig_cap + (k - from_dn));
    ClbMig m;
    const int4* s4 = reinterpret_cast<const int4*>(src); int4* d4 = reinterpret_cast<int4*>(&m);
#pragma unroll
    for (int q = 0; q < (int)(sizeof(ClbMig) / 16); ++q) d4[q] = __ldcg(s4 + q);
    pos[base + k] = m.p; vel[base + k] = m.v; slot[base + k] = m.slot;
    image[3 * m.slot] = m.ix; image[3 * m.slot + 1] = m.iy; image[3 * m.slot + 2] = m.iz;
}
__global__ void k_halo_ranges(const int* __restrict__ cell_start, int plane_cells, int nczl, int nown, ClbHalo* H) {
    H->send_lo0 = 0; H->send_lo1 = cell_start[plane_cells];                   // bottom owned plane
    H->send_hi0 = cell_start[(nczl - 1) * plane_cells]; H->send_hi1 = nown;   // top owned plane
    H->own1 = nown;
}
__global__ void __launch_bounds__(256) k_ghost_push(ClbCtl* ctl, const ClbHalo* __restrict__ H, const int4* __restrict__ pos, const int* __restrict__ slot,
                                                    ClbPeers P, unsigned long long epoch) {
    const ClbHalo h = *H;
    int nlo = h.send_lo1 - h.send_lo0, nhi = h.send_hi1 - h.send_hi0;
    if (nlo > P.L.plane_cap || nhi > P.L.plane_cap) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&ctl->err, CLB_EF_COMM_OVERFLOW); nlo = min(nlo, P.L.plane_cap); nhi = min(nhi, P.L.plane_cap); }
    int4* p_dn = reinterpret_cast<int4*>(P.dn + P.L.gpos) + (size_t)0 * P.L.plane_cap; int* s_dn = reinterpret_cast<int*>(P.dn + P.L.gslot) + (size_t)0 * P.L.plane_cap;
    int4* p_up = reinterpret_cast<int4*>(P.up + P.L.gpos) + (size_t)1 * P.L.plane_cap; int* s_up = reinterpret_cast<int*>(P.up + P.L.gslot) + (size_t)1 * P.L.plane_cap;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nlo + nhi; i += gridDim.x * blockDim.x) {
        if (i < nlo) { p_dn[i] = pos[h.send_lo0 + i]; s_dn[i] = slot[h.send_lo0 + i]; }
        else { p_up[i - nlo] = pos[h.send_hi0 + (i - nlo)]; s_up[i - nlo] = slot[h.send_hi0 + (i - nlo)]; }
    }
    if (grid_last_block(ctl) && threadIdx.x == 0) {
        ClbMailHdr* hu = reinterpret_cast<ClbMailHdr*>(P.up); ClbMailHdr* hdn = reinterpret_cast<ClbMailHdr*>(P.dn);
        *(volatile int*)&hdn->ghost_cnt[0] = nlo; *(volatile int*)&hu->ghost_cnt[1] = nhi;
        __threadfence_system();
        *(volatile unsigned long long*)&hdn->ghost_flag[0] = epoch; *(volatile unsigned long long*)&hu->ghost_flag[1] = epoch;
    }
}
__global__ void __launch_bounds__(256) k_ghost_recv(ClbCtl* ctl, ClbHalo* H, int4* __restrict__ pos, int* __restrict__ slot, ClbPeers P,
                                                    unsigned long long epoch, int room) {
    const ClbMailHdr* hd = reinterpret_cast<const ClbMailHdr*>(P.self);
    __shared__ int s_ok;
    if (threadIdx.x == 0) s_ok = wait_flag(&hd->ghost_flag[0], epoch) && wait_flag(&hd->ghost_flag[1], epoch);
    __syncthreads();
    if (!s_ok) { if (threadIdx.x == 0) atomicOr(&ctl->err, CLB_EF_COMM_TIMEOUT); return; }
    __threadfence_system();
    int n_hi = *(const volatile int*)&hd->ghost_cnt[0], n_lo = *(const volatile int*)&hd->ghost_cnt[1];
    const int own1 = H->own1;
    if (n_hi + n_lo > room) { if (blockIdx.x == 0 && threadIdx.x == 0) { atomicOr(&ctl->err, CLB_EF_COMM_OVERFLOW); H->n_hi = 0; H->n_lo = 0; } return; }
    const int4* p_hi = reinterpret_cast<const int4*>(P.self + P.L.gpos); const int* s_hi = reinterpret_cast<const int*>(P.self + P.L.gslot);
    const int4* p_lo = p_hi + P.L.plane_cap; const int* s_lo = s_hi + P.L.plane_cap;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_hi + n_lo; i += gridDim.x * blockDim.x) {
        if (i < n_hi) { pos[own1 + i] = __ldcg(p_hi + i); slot[own1 + i] = __ldcg(s_hi + i); }
        else { pos[own1 + i] = __ldcg(p_lo + (i - n_hi)); slot[own1 + i] = __ldcg(s_lo + (i - n_hi)); }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { H->n_hi = n_hi; H->n_lo = n_lo; }
}

bool clb_engine::peer_active() const { return nranks > 1 && cd && cd->peer_ok; }
static ClbPeers make_peers(clb_engine* e) {
    clb_engine::CommDev& c = *e->cd;
    ClbPeers P; P.self = c.mb; P.up = (unsigned char*)c.mb_peer[c.up]; P.dn = (unsigned char*)c.mb_peer[c.dn];
    P.all = (unsigned char* const*)c.d_mb_all.p; P.rank = e->rank; P.nranks = e->nranks; P.L = mail_layout(c.plane_cap, c.mig_cap);
    return P;
}

// Allocate and export the mailbox, import the mailboxes of all ranks (cudaIpc); called whenever the particle capacity changes.
// Any failure (IPC not permitted in this container, peer access unavailable) leaves peer_ok = 0 on EVERY rank: NCCL path.
int clb_engine::comm_peer_setup() {
    clb_engine* e = this;
    CommDev& c = *cd;
    // auto: mailboxes from 3 ranks on.  Measured on the 1M-bead melt (profiles/r2q_*, r2o_*): 2 ranks 1918 (mailboxes) vs 2161 steps/s
    // (NCCL: both planes of a rank go to the same peer, one grouped NCCL kernel does it all), 4 ranks 2987 vs 2740, 8 ranks 4060 vs 3870
    if (peer_user == 0 || (peer_user < 0 && nranks < 3)) { c.peer_ok = 0; return CLB_OK; }
    const int plane_cap = (int)(2.0 * n / std::max(1, grid.ncz)) + 8192;
    // the layout must be the SAME on every rank (a sender addresses slot 1 of its neighbour's block with its own copy of the
    // layout): sizes come from global quantities only.  (ncap differs between ranks with 19 and 18 planes: at 1M beads the migrants
    // of one rank landed 0.3 MB beside the slot the receiver read.)
    const int mig_cap = std::max(16384, plane_cap / 2);
    if (c.mb && plane_cap <= c.plane_cap && mig_cap <= c.mig_cap) return CLB_OK;
    comm_peer_teardown();
    if (nranks > CLB_MAX_RANKS) { c.peer_ok = 0; return CLB_OK; }
    ClbMailLayout L = mail_layout(plane_cap, mig_cap);
    int ok = 1;
    if (cudaMalloc((void**)&c.mb, L.total) != cudaSuccess) { (void)cudaGetLastError(); c.mb = nullptr; ok = 0; }
    cudaIpcMemHandle_t mine; memset(&mine, 0, sizeof(mine));
    if (ok) {
        CK(cudaMemsetAsync(c.mb, 0, L.total, stream));
        if (cudaIpcGetMemHandle(&mine, c.mb) != cudaSuccess) { (void)cudaGetLastError(); ok = 0; }
    }
    // all-gather of the 64-byte handles (+ an ok byte) over NCCL
    const size_t rec = sizeof(cudaIpcMemHandle_t) + 8;
    DevBuf<unsigned char> dsend, drecv;
    CK(dsend.ensure(rec)); CK(drecv.ensure(rec * nranks));
    std::vector<unsigned char> hs(rec, 0), hr(rec * nranks, 0);
    memcpy(hs.data(), &mine, sizeof(mine)); hs[sizeof(mine)] = (unsigned char)ok;
    CK(cudaMemcpyAsync(dsend.p, hs.data(), rec, cudaMemcpyHostToDevice, stream));
    NC(g_nccl.AllGather(dsend.p, drecv.p, rec, ncclChar, c.comm, stream));
    CK(cudaMemcpyAsync(hr.data(), drecv.p, rec * nranks, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    dsend.release(); drecv.release();
    for (int r = 0; r < nranks; ++r) ok = ok && hr[r * rec + sizeof(mine)];
    c.mb_peer.assign(nranks, nullptr);
    if (ok) {
        for (int r = 0; r < nranks && ok; ++r) {
            if (r == rank) { c.mb_peer[r] = c.mb; continue; }
            cudaIpcMemHandle_t h; memcpy(&h, &hr[r * rec], sizeof(h));
            void* ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { (void)cudaGetLastError(); ok = 0; break; }
            c.mb_peer[r] = ptr;
        }
    }
    // every rank must take the same decision
    double v = ok ? 0.0 : 1.0;
    TRY(comm_allreduce_sum(&v, 1));
    ok = v == 0.0;
    if (!ok) { comm_peer_teardown(); c.peer_ok = 0; if (getenv("CLB_TRACE")) fprintf(stderr, "[clb comm] rank %d: peer mailboxes unavailable, using NCCL for the step path\n", rank); return CLB_OK; }
    c.plane_cap = plane_cap; c.mig_cap = mig_cap; c.mb_bytes = L.total;
    std::vector<unsigned char*> tbl(nranks);
    for (int r = 0; r < nranks; ++r) tbl[r] = (unsigned char*)c.mb_peer[r];
    CK(c.d_mb_all.ensure(nranks));
    CK(cudaMemcpyAsync(c.d_mb_all.p, tbl.data(), nranks * sizeof(unsigned char*), cudaMemcpyHostToDevice, stream));
    if (!c.d_halo) { CK(cudaMalloc((void**)&c.d_halo, sizeof(ClbHalo))); CK(cudaMallocHost((void**)&c.h_halo, sizeof(ClbHalo))); }
    CK(cudaMemsetAsync(c.d_halo, 0, sizeof(ClbHalo), stream));
    CK(cudaMemsetAsync(&d_ctl->halo_epoch, 0, 16, stream));
    c.mig_epoch = 0; c.ghost_epoch = 0;
    CK(cudaStreamSynchronize(stream));
    // nobody may write into a mailbox before its owner has cleared it
    double z = 0.0; TRY(comm_allreduce_sum(&z, 1));
    c.peer_ok = 1;
    return CLB_OK;
}
void clb_engine::comm_peer_teardown() {
    if (!cd) return;
    CommDev& c = *cd;
    for (int r = 0; r < (int)c.mb_peer.size(); ++r) if (c.mb_peer[r] && r != rank) cudaIpcCloseMemHandle(c.mb_peer[r]);
    c.mb_peer.clear();
    if (c.mb) { cudaFree(c.mb); c.mb = nullptr; }
    c.plane_cap = 0; c.mig_cap = 0; c.peer_ok = 0;
    (void)cudaGetLastError();
}

int clb_engine::comm_step_peer(cudaStream_t st, int step_index) {
    CommDev& c = *cd;
    ClbPeers P = make_peers(this);
    const int nb = std::max(1, std::min(64, ceil_div(2 * c.plane_cap / 4, 256)));
    k_halo_push<<<nb, 256, 0, st>>>(d_ctl, c.d_halo, pos.p, P);
    k_halo_recv<<<nb, 256, 0, st>>>(d_ctl, c.d_halo, pos.p, P);
    k_check_resort_peer<<<1, 1, 0, st>>>(d_ctl, criterion, 0.5 * skin, step_index);
    launches += 3;
    return CLB_OK;
}

int clb_engine::comm_migrate_peer() {
    clb_engine* e = this;
    CommDev& c = *cd;
    const int no = own1;
    CK(c.mkey.ensure(ncap)); CK(c.mkey2.ensure(ncap)); CK(c.mval.ensure(ncap)); CK(c.mval2.ensure(ncap)); CK(c.moff.ensure(8));
    k_mig_classify<<<ceil_div(std::max(no, 1), 256), 256, 0, stream>>>(no, pos.p, grid, c.mkey.p, c.mval.p, d_ctl);
    size_t tb = cubtmp.n;
    cub::DeviceRadixSort::SortPairs(cubtmp.p, tb, c.mkey.p, c.mkey2.p, c.mval.p, c.mval2.p, no, 0, 2, stream);
    k_lower_bounds<<<1, 32, 0, stream>>>(no, c.mkey2.p, 3, c.moff.p);
    k_mig_counts<<<1, 1, 0, stream>>>(c.moff.p, c.cnt.p);
    ClbPeers P = make_peers(this);
    const unsigned long long ep = ++c.mig_epoch;
    k_mig_push<<<std::max(1, std::min(64, ceil_div(c.mig_cap, 256))), 256, 0, stream>>>(d_ctl, c.cnt.p, c.mval2.p, pos.p, vel.p, slot.p, image.p, P, ep);
    k_mig_wait<<<1, 1, 0, stream>>>(d_ctl, P, ep, c.cnt.p);
    CK(cudaMemcpyAsync(c.h_cnt, c.cnt.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(h_ctl, d_ctl, sizeof(ClbCtl), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (h_ctl->err & CLB_EF_COMM_TIMEOUT) return fail(CLB_ERR_COMM, "rank %d: a neighbour rank did not deliver its migrants (peer flag timeout)", rank);
    if (h_ctl->err & CLB_EF_COMM_OVERFLOW) return fail(CLB_ERR_RANGE, "rank %d: more migrants than the peer mailbox holds (%d)", rank, c.mig_cap);
    const int nstay = c.h_cnt[0], from_dn = c.h_cnt[4], from_up = c.h_cnt[5];
    const int nnew = nstay + from_dn + from_up;
    if (nnew + 64 > ncap) return fail(CLB_ERR_RANGE, "rank %d: %d owned particles exceed the local capacity %d", rank, nnew, ncap);
    if (nstay > 0) k_mig_compact<<<ceil_div(nstay, 256), 256, 0, stream>>>(nstay, c.mval2.p, pos.p, vel.p, slot.p, pos2.p, vel2.p, slot2.p);
    if (from_dn + from_up > 0)
        k_mig_unpack_peer<<<ceil_div(from_dn + from_up, 256), 256, 0, stream>>>(from_dn, from_up, P, nstay, pos2.p, vel2.p, slot2.p, image.p);
    std::swap(pos, pos2); std::swap(vel, vel2); std::swap(slot, slot2);      // whole handles (pointer, capacity and block size)
    own0 = 0; own1 = nnew; nstored = nnew;
    CK(cudaGetLastError());
    launches += 7;
    return CLB_OK;
}

int clb_engine::comm_exchange_ghosts_peer() {
    clb_engine* e = this;
    CommDev& c = *cd;
    const int no = own1;
    ClbPeers P = make_peers(this);
    const unsigned long long ep = ++c.ghost_epoch;
    k_halo_ranges<<<1, 1, 0, stream>>>(cell_start.p, grid.ncx * grid.ncy, grid.nczl, no, c.d_halo);
    const int nb = std::max(1, std::min(64, ceil_div(2 * c.plane_cap / 4, 256)));
    k_ghost_push<<<nb, 256, 0, stream>>>(d_ctl, c.d_halo, pos.p, slot.p, P, ep);
    k_ghost_recv<<<nb, 256, 0, stream>>>(d_ctl, c.d_halo, pos.p, slot.p, P, ep, ncap - 64 - no);
    CK(cudaMemcpyAsync(c.h_halo, c.d_halo, sizeof(ClbHalo), cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(h_ctl, d_ctl, sizeof(ClbCtl), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (h_ctl->err & CLB_EF_COMM_TIMEOUT) return fail(CLB_ERR_COMM, "rank %d: a neighbour rank did not deliver its boundary plane (peer flag timeout)", rank);
    if (h_ctl->err & CLB_EF_COMM_OVERFLOW) return fail(CLB_ERR_RANGE, "rank %d: boundary planes exceed the peer mailbox (%d beads) or the local capacity", rank, c.plane_cap);
    c.send_lo0 = c.h_halo->send_lo0; c.send_lo1 = c.h_halo->send_lo1; c.send_hi0 = c.h_halo->send_hi0; c.send_hi1 = c.h_halo->send_hi1;
    c.n_hi = c.h_halo->n_hi; c.n_lo = c.h_halo->n_lo;
    const int ns = no + c.n_hi + c.n_lo;
    nstored = ns;
    if (ns > no) k_ghost_finish<<<ceil_div(ns - no, 256), 256, 0, stream>>>(no, ns, pos.p, slot.p, grid, key2.p, id2idx.p);
    CK(cudaGetLastError());
    launches += 4;
    return CLB_OK;
}

// sum of host doubles over the ranks (observables)
int clb_engine::comm_allreduce_sum(double* v, int nv) {
    clb_engine* e = this;
    if (nranks == 1) return CLB_OK;
    CK(cd->red.ensure(nv));
    CK(cudaMemcpyAsync(cd->red.p, v, nv * sizeof(double), cudaMemcpyHostToDevice, stream));
    NC(g_nccl.AllReduce(cd->red.p, cd->red.p, nv, ncclDouble, ncclSum, cd->comm, stream));
    CK(cudaMemcpyAsync(v, cd->red.p, nv * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    return CLB_OK;
}
// in-place sum of a device array of doubles over the ranks
int clb_engine::comm_allreduce_sum_dev(double* d, size_t nv) {
    clb_engine* e = this;
    if (nranks == 1) return CLB_OK;
    NC(g_nccl.AllReduce(d, d, nv, ncclDouble, ncclSum, cd->comm, stream));
    return CLB_OK;
}

// variable-size all-gather of device bytes: every rank ends with the concatenation (rank order) in cd->gat
int clb_engine::comm_allgatherv(const void* dsend, size_t bytes, void** dout, size_t* total) {
    clb_engine* e = this;
    CommDev& c = *cd;
    CK(c.sizes.ensure(2 * (size_t)nranks));
    long long mine = (long long)bytes;
    CK(cudaMemcpyAsync(c.sizes.p + nranks + rank, &mine, 8, cudaMemcpyHostToDevice, stream));
    NC(g_nccl.AllGather(c.sizes.p + nranks + rank, c.sizes.p, 1, ncclInt64, c.comm, stream));
    std::vector<long long> hs(nranks);
    CK(cudaMemcpyAsync(hs.data(), c.sizes.p, nranks * 8, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    size_t tot = 0;
    for (int r = 0; r < nranks; ++r) tot += (size_t)hs[r];
    CK(c.gat.ensure(tot + 16));
    size_t off = 0;
    NC(g_nccl.GroupStart());
    for (int r = 0; r < nranks; ++r) {
        if (hs[r] > 0) NC(g_nccl.Broadcast(r == rank ? dsend : (const void*)(c.gat.p + off), c.gat.p + off, (size_t)hs[r], ncclChar, r, c.comm, stream));
        off += (size_t)hs[r];
    }
    NC(g_nccl.GroupEnd());
    CK(cudaStreamSynchronize(stream));   // callers read the result with plain (legacy-stream) copies
    *dout = c.gat.p; *total = tot;
    return CLB_OK;
}

// reaction handshake: concatenate the candidates of all ranks into R.cands (every rank then sorts them canonically)
int clb_engine::comm_gather_candidates(long long* nc) {
    clb_engine* e = this;
    ReactDev& R = *rd;
    void* all = nullptr; size_t tot = 0;
    TRY(comm_allgatherv(R.cands.p, (size_t)*nc * sizeof(ClbCand), &all, &tot));
    size_t m = tot / sizeof(ClbCand);
    if (m > R.candcap) { R.candcap = m * 5 / 4 + 1024; CK(R.cands.ensure(R.candcap)); }
    if (m) CK(cudaMemcpyAsync(R.cands.p, all, tot, cudaMemcpyDeviceToDevice, stream));
    *nc = (long long)m;
    return CLB_OK;
}
