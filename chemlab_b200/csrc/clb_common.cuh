// clb_common.cuh -- shared device/host definitions of the B200 engine (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define CLB_MAX_TYPES 32
#define CLB_MAXDEG 12          // max bonds per particle in the topology graph
#define CLB_MAX_REACTIONS 32
#define CLB_MAX_LISTS 64
#define CLB_TILE_ROWS 9        // (dy,dz) in {-1,0,1}^2
#define CLB_MAX_BX 16          // max home cells per row block

// error flags raised by kernels (ctl->err)
#define CLB_EF_TABLE_RANGE 1u   // table index out of range (fatal in the reference, U12)
#define CLB_EF_LIST_OVERFLOW 2u // neighbour list capacity exceeded -> host regrows and rebuilds
#define CLB_EF_DEGREE 4u        // topology graph degree overflow
#define CLB_EF_CAND_OVERFLOW 8u // reaction candidate buffer overflow -> host regrows and rescans
#define CLB_EF_TUPLE_OVERFLOW 16u
#define CLB_EF_PARTNER_LOST 32u // bonded partner not resolvable (outside ghost layer)
#define CLB_EF_BFS_OVERFLOW 128u // neighbour-property BFS frontier exceeded its fixed buffers, or molecule-id hooking did not converge
#define CLB_EF_COMM_TIMEOUT 256u // a peer flag did not arrive within the time limit (a neighbour rank died or fell out of step)
#define CLB_EF_COMM_OVERFLOW 512u // more boundary-plane particles / migrants than the peer mailbox holds
#define CLB_EF_TILE_OVERFLOW 64u // a tile holds more particles than the shared-memory carve-up assumed -> host retries

// velocity + mass per sorted particle: fp64 throughout ({vx, vy, vz, mass}); the reference's `real` is double
typedef double4 ClbVel;
__host__ __device__ inline ClbVel clb_make_vel(double x, double y, double z, double m) { ClbVel v; v.x = x; v.y = y; v.z = z; v.w = m; return v; }

// pos.w packing: type in bits 0..7, chemical state in bits 8..23 (signed 16 bit)
__host__ __device__ inline int pw_type(int w) { return w & 0xff; }
__host__ __device__ inline int pw_state(int w) { return (int)((int16_t)((w >> 8) & 0xffff)); }
__host__ __device__ inline int pw_pack(int type, int state) { return (type & 0xff) | ((state & 0xffff) << 8); }

// device-resident control block: everything the step kernels need to decide without the host
struct ClbCtl {
    int stall;              // set by k_check_resort when a rebuild is needed; all step kernels early-exit
    int stall_step;         // index (within the current clb_run) of the step that stalled
    unsigned err;           // CLB_EF_*
    unsigned maxdisp2_bits; // float bits: max |dx|^2 of this step (criterion 0) or since rebuild (criterion 1)
    double accum_maxdist;   // criterion 0 accumulator (sum over steps of per-step max displacement)
    int steps_ok;           // steps whose phase A completed without stalling in this chunk
    int force_rebuild;      // host/reaction request: rebuild at the next resort check
    unsigned long long ncand;   // reaction candidates found
    unsigned long long npairs_out;
    int nev;                // reaction events applied in the last pass
    int rounds;
    int tile_max;           // max tile particle count over blocks (rebuild statistics)
    int home_max;           // max home particle count over blocks
    int cell_max;
    int nl_max;             // max neighbour count
    unsigned long long nl_total;
    unsigned long long halo_epoch;  // multi-GPU peer path: number of completed per-step halo exchanges (identical on every rank)
    unsigned comm_done;             // block-completion counter of the push kernels
    unsigned comm_pad;
    int nblocks;            // row blocks of the current block table (k_blocks_scan)
    int blk_p1, blk_pl;     // first block of owned plane 1 and of the last owned plane (multi-GPU interior / boundary split)
    int maybe;              // resort criterion 2: the global maximum passed skin/2 -> the per-cell bound decides (k_cell_*)
    int maybe_step;         // step index of that check
    int blk_pad[3];
};

struct ClbGrid {
    int ncx, ncy, ncz, ncell;
    int bx;      // max home cells per row block
    int target;  // max home particles per row block (a single fuller cell still makes a block); <= 0: cells only (fixed-length blocks)
    int nblocks; // number of row blocks in the table (host copy of ctl->nblocks; < 0: read ctl->nblocks on the device)
    const int4* blk;     // block table {first home cell x, home cells, row y, local plane}, rebuilt with the cell lists (k_blocks_*)
    const int* nblk_d;   // &ctl->nblocks
    int cz0, nczl; // owned z-plane range [cz0, cz0+nczl) of this rank (single GPU: 0, ncz)
    int zoff;      // = cz0.  Local plane of global plane cz: l = (cz - zoff) mod ncz; owned planes are l in [0,nczl),
                   // the upper ghost plane (cz0+nczl) is local plane nczl, the lower ghost plane (cz0-1, l = ncz-1) is
                   // stored as local plane nczl+1 -> plane arithmetic is periodic modulo nplanes on every rank
    int nplanes;   // number of locally stored planes: ncz (single GPU) or nczl + 2
    int ghost;     // 1 when ghost planes are present (multi-GPU)
};

// geometry constants
struct ClbGeom {
    double q[3];     // lattice spacing per dimension: L_d / 2^32
    double box[3];
    double rl2;      // (rc_max + skin)^2
    int cut[3];      // integer prefilter: ceil((rc+skin)/q_d)+1
    int cubic;
    double q2;       // q[0]^2 when cubic
};

struct ClbPairDesc {   // per type pair, 16 bytes: one LDS.128 per listed pair
    double rc2;        // cutoff^2 (real units); negative when the pair has no potential
    int kind;          // 0 none, 1 table, 2 LJ
    int tab;           // table slot (general) or first row of the table (uniform-grid fast path)
};                     // LJ coefficients {48 eps sigma^12, 24 eps sigma^6} live in a parallel double2 array
struct ClbPairDescE {  // energy-side extras
    double e12, e6, shift; // LJ: 4 eps sigma^12, 4 eps sigma^6, shift
    int inter;             // owning interaction handle
    int pad;
};
struct ClbTabMeta {    // per pair table
    double x0, invdx, c_t, c_idx; // c_t = -x0*invdx - 0.5 (so that u = r*invdx + c_t = (r-x0)/dx - 1/2); c_idx unused
    double dx;
    int n, off;                   // rows, offset into the row arrays
};

// ---- Philox4x32-10 (the CPU checker restates the same function; Random123 known answers are tested on both) ----
__host__ __device__ inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
#define CLB_STREAM_LANGEVIN 0x4c414e47u
#define CLB_STREAM_HEATUP 0x48454154u
#define CLB_STREAM_REACT 0x52454143u
#define CLB_STREAM_PARTNER 0x50415254u
#define CLB_STREAM_ATRP 0x41545250u

__host__ __device__ inline void clb_draw3(uint64_t seed, uint32_t stream, uint64_t step, uint32_t idx, double u[3]) {
    uint32_t c[4] = {idx, 0u, (uint32_t)step, (uint32_t)(step >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32) ^ stream);
#pragma unroll
    for (int k = 0; k < 3; ++k) u[k] = ((double)c[k] + 0.5) * (1.0 / 4294967296.0);
}
__host__ __device__ inline void clb_draw_pair(uint64_t seed, uint32_t stream, uint64_t step, uint32_t a, uint32_t b,
                                              uint32_t r, uint32_t out[4]) {
    uint32_t c[4] = {a, b, (uint32_t)step, ((uint32_t)(step >> 32) << 8) ^ r};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32) ^ stream);
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

#ifdef __CUDACC__
// exact int32 -> double without the conversion unit: low word = (d + 2^31) mod 2^32 under exponent 2^52
__device__ __forceinline__ double lat2d(int d) {
    return __hiloint2double(0x43300000, (int)((unsigned)d ^ 0x80000000u)) - 4503601774854144.0; // 2^52 + 2^31
}
// lattice coordinates wrap modulo 2^32: do the arithmetic unsigned (signed overflow is undefined behaviour)
__host__ __device__ __forceinline__ int wsub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }
__host__ __device__ __forceinline__ int wadd(int a, int b) { return (int)((unsigned)a + (unsigned)b); }
__device__ __forceinline__ int wrapi(int c, int n) { return c < 0 ? c + n : (c >= n ? c - n : c); }
// number of row blocks: the host's copy, or (kernels enqueued between k_blocks_scan and the host check of a rebuild) the device's
__device__ __forceinline__ int grid_nblocks(const ClbGrid& g) { return g.nblocks >= 0 ? g.nblocks : *(volatile const int*)g.nblk_d; }
// local plane index of global plane cz (see ClbGrid::zoff); planes that are neither owned nor ghost map to -1
__device__ __forceinline__ int local_plane(const ClbGrid& g, int cz) {
    if (!g.ghost) return cz;
    int l = wrapi(cz - g.zoff, g.ncz);
    if (l <= g.nczl) return l;
    return l == g.ncz - 1 ? g.nczl + 1 : -1;
}
#endif
