// clb_tile.cuh -- row-block tiles: neighbour-list build, tabulated/LJ pair forces, pair energies,
// pair-set decode and the reaction candidate scan.  All of them share one tile geometry:
//
//   cells are ordered x-fastest:  c = (lz*ncy + cy)*ncx + cx, particles are sorted by cell, so a
//   row block of `bx` consecutive home cells owns ONE contiguous particle range, and its 27-cell
//   neighbourhood is 9 rows (dy,dz) of at most bx+2 consecutive cells: 9 contiguous runs that are
//   staged once into shared memory as int4 {x,y,z lattice, type|state}.  A neighbour-list entry is
//   the 16-bit position of the partner inside that tile, so the force kernel gathers partners from
//   shared memory only (no global gather, no index->address arithmetic).
//
// Replaces [EXT] VerletList::rebuild + CellListAllPairsIterator and
// VerletListInteractionTemplate<Tabulated|LennardJones>::addForces.
#pragma once
#include "clb_common.cuh"

#define CLB_TILE_CELLS (CLB_TILE_ROWS * (CLB_MAX_BX + 2))

struct TileCtx {
    int cx0, cy, lz, bxe;   // first home cell x, row y, local plane, effective home cells
    int W;                  // tile cells per row
    int whole;              // tile row = whole x row (small ncx)
    int hs, nh;             // first home particle (global sorted index), number of home particles
    int T;                  // tile particle count
};

// decode block -> geometry.  Must be called by all threads (uniform).
__device__ __forceinline__ void tile_geometry(const ClbGrid& g, int b, TileCtx& t) {
    const int4 q = __ldg(g.blk + b);        // block table (k_blocks_rows): {first home cell, home cells, row y, owned plane}
    t.cx0 = q.x;
    t.bxe = q.y;
    t.cy = q.z;
    t.lz = q.w;
    t.whole = (t.bxe + 2 > g.ncx);
    t.W = t.whole ? g.ncx : t.bxe + 2;
}
// global (local-grid) cell index of tile cell (row k, column m)
__device__ __forceinline__ int tile_cell(const ClbGrid& g, const TileCtx& t, int k, int m) {
    int dy = k % 3 - 1, dz = k / 3 - 1;
    int cy = wrapi(t.cy + dy, g.ncy);
    int lz = wrapi(t.lz + dz, g.nplanes);   // owned planes first, then upper ghost, then lower ghost (ClbGrid::zoff)
    int cx = t.whole ? m : wrapi(t.cx0 - 1 + m, g.ncx);
    return (lz * g.ncy + cy) * g.ncx + cx;
}
// A CTA may host several independent "virtual CTAs" (groups of whole warps with their own tile and their own named
// barrier) that share read-only shared-memory data such as the force table: vc_sync is their private barrier.
__device__ __forceinline__ void vc_sync(int bar_id, int nthreads) {
    asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "r"(nthreads) : "memory");
}
// Builds s_off[0..9W] (exclusive prefix of cell counts in tile order) and s_src[tc] (global start of
// each tile cell).  Needs blockDim.x >= 32.  Ends with __syncthreads().
__device__ __forceinline__ void tile_offsets(const ClbGrid& g, TileCtx& t, const int* __restrict__ cell_start,
                                             int* s_off, int* s_src) {
    const int nct = CLB_TILE_ROWS * t.W;
    for (int tc = threadIdx.x; tc < nct; tc += blockDim.x) {
        int k = tc / t.W, m = tc - k * t.W;
        int gc = tile_cell(g, t, k, m);
        int s = __ldg(cell_start + gc), e = __ldg(cell_start + gc + 1);
        s_src[tc] = s;
        s_off[tc + 1] = e - s;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        int lane = threadIdx.x;
        int per = (nct + 31) >> 5;
        int lo = lane * per, hi = min(lo + per, nct);
        int sum = 0;
        for (int i = lo; i < hi; ++i) sum += s_off[i + 1];
        int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
        int run = incl - sum;
        for (int i = lo; i < hi; ++i) { int c = s_off[i + 1]; s_off[i + 1] = run + c; run += c; }
        if (lane == 0) s_off[0] = 0;
    }
    __syncthreads();
    t.T = s_off[nct];
    // home range: tile row 4 (dy=dz=0), columns of the home cells
    int mh0 = t.whole ? t.cx0 : 1;
    t.hs = s_src[4 * t.W + mh0];
    t.nh = s_off[4 * t.W + mh0 + t.bxe] - s_off[4 * t.W + mh0];
}
// stage positions (and optionally slots / global indices) of the whole tile into shared memory
__device__ __forceinline__ void tile_stage(const TileCtx& t, const int* s_off, const int* s_src,
                                           const int4* __restrict__ pos, int4* s_pos,
                                           const int* __restrict__ slot, int* s_slot, int* s_gidx) {
    const int nct = CLB_TILE_ROWS * t.W;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int tc = warp; tc < nct; tc += nw) {
        int o = s_off[tc], c = s_off[tc + 1] - o, s = s_src[tc];
        for (int i = lane; i < c; i += 32) {
            s_pos[o + i] = __ldg(pos + s + i);
            if (s_slot) s_slot[o + i] = __ldg(slot + s + i);
            if (s_gidx) s_gidx[o + i] = s + i;
        }
    }
}
// tile_offsets / tile_stage for a group of `nth` threads (tid in [0,nth), nth % 32 == 0) synchronised by named barrier `bar`
__device__ __forceinline__ void tile_offsets_vc(const ClbGrid& g, TileCtx& t, const int* __restrict__ cell_start,
                                                int* s_off, int* s_src, int tid, int nth, int bar) {
    const int nct = CLB_TILE_ROWS * t.W;
    for (int tc = tid; tc < nct; tc += nth) {
        int k = tc / t.W, m = tc - k * t.W;
        int gc = tile_cell(g, t, k, m);
        int s = __ldg(cell_start + gc), e = __ldg(cell_start + gc + 1);
        s_src[tc] = s;
        s_off[tc + 1] = e - s;
    }
    vc_sync(bar, nth);
    if (tid < 32) {
        int lane = tid;
        int per = (nct + 31) >> 5;
        int lo = lane * per, hi = min(lo + per, nct);
        int sum = 0;
        for (int i = lo; i < hi; ++i) sum += s_off[i + 1];
        int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
        int run = incl - sum;
        for (int i = lo; i < hi; ++i) { int c = s_off[i + 1]; s_off[i + 1] = run + c; run += c; }
        if (lane == 0) s_off[0] = 0;
    }
    vc_sync(bar, nth);
    t.T = s_off[nct];
    int mh0 = t.whole ? t.cx0 : 1;
    t.hs = s_src[4 * t.W + mh0];
    t.nh = s_off[4 * t.W + mh0 + t.bxe] - s_off[4 * t.W + mh0];
}
__device__ __forceinline__ void tile_stage_vc(const TileCtx& t, const int* s_off, const int* s_src,
                                              const int4* __restrict__ pos, int4* s_pos, int tid, int nth) {
    const int nct = CLB_TILE_ROWS * t.W;
    const int warp = tid >> 5, lane = tid & 31, nw = nth >> 5;
    for (int tc = warp; tc < nct; tc += nw) {
        int o = s_off[tc], c = s_off[tc + 1] - o, s = s_src[tc];
        for (int i = lane; i < c; i += 32) s_pos[o + i] = __ldg(pos + s + i);
    }
}
// tile column of the home cell that holds home particle p (local index in [0,nh))
__device__ __forceinline__ int home_column(const TileCtx& t, const int* s_off, int p) {
    int mh0 = t.whole ? t.cx0 : 1;
    int base = s_off[4 * t.W + mh0];
    int m = mh0;
    while (m < mh0 + t.bxe - 1 && s_off[4 * t.W + m + 1] - base <= p) ++m;
    return m;
}

// Work order of a block's home particles for the pair kernel: longest neighbour row first.  Lane l of warp w takes the home
// particle perm[32 w + l], so the rows of one warp have nearly the same length and few lanes idle while the longest row of
// the warp finishes (cell order mixes rows of 55..95 entries in every warp: 12 % of the lane slots).  Stable counting rank,
// called by the whole CTA after all rows of the block are written; s_cnt[p] = row length of home particle p.
#define CLB_PERM_MAX 1024
__device__ __forceinline__ void block_perm(int hs, int nh, const int* s_cnt, unsigned short* __restrict__ perm) {
    if (nh > CLB_PERM_MAX) { for (int p = threadIdx.x; p < nh; p += blockDim.x) perm[hs + p] = (unsigned short)(p & 0xffff); return; }
    for (int p = threadIdx.x; p < nh; p += blockDim.x) {
        const int c = s_cnt[p];
        int r = 0;
        for (int q = 0; q < nh; ++q) { const int d = s_cnt[q]; r += (d > c || (d == c && q < p)) ? 1 : 0; }
        perm[hs + r] = (unsigned short)p;
    }
}

// ------------------------------------------------------------------------------------------
// Neighbour-list build.  One WARP per home particle, lanes = candidates: the 27 cells around the
// particle's cell are 9 contiguous tile ranges, read 32 at a time (conflict-free LDS.128), tested
// exactly, compacted with a ballot and appended to the particle's row -> coalesced stores and no
// divergence.  Inclusion test (U1): r^2 <= (rc+skin)^2.  Cubic boxes: exact unsigned 64-bit integer
// arithmetic on lattice differences (3 IMAD.WIDE + one compare); otherwise fp64 per-dimension scaling.
// Exclusions: the particle's partner-slot row is held across the lanes and matched by shuffles.
// Entry layout: entries[gi*cap + k] (cap % 8 == 0 so the force kernel reads 8 entries per LDG.128).
#define CLB_BUILD_G 4   // home particles of one cell that share each candidate load
template <bool CUBIC>
__global__ void __launch_bounds__(512) k_build_lists(ClbGrid g, ClbGeom geo, unsigned long long rl2_lat,
                                                     const int* __restrict__ cell_start,
                                                     const int4* __restrict__ pos, const int* __restrict__ slot,
                                                     const int* __restrict__ excl_off, const int* __restrict__ excl_ids,
                                                     unsigned short* __restrict__ entries, int* __restrict__ nl_count,
                                                     unsigned short* __restrict__ nl_perm, int cap, int tile_cap, ClbCtl* ctl) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_off[CLB_TILE_CELLS + 1];
    __shared__ int s_src[CLB_TILE_CELLS];
    __shared__ int s_item[CLB_MAX_BX + 1];      // prefix of (cell, group-of-G) work items over the home cells
    __shared__ int s_cnt[CLB_PERM_MAX];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    // dynamic smem: tile positions | tile slots | G staging rows (cap entries each) per warp
    int4* s_pos = reinterpret_cast<int4*>(smem);
    int* s_slot = reinterpret_cast<int*>(s_pos + tile_cap);
    unsigned short* s_rows = reinterpret_cast<unsigned short*>(s_slot + tile_cap) + (size_t)warp * CLB_BUILD_G * cap;
    int lmax = 0;
    unsigned long long ltot = 0;
    const int nblk = grid_nblocks(g);
    for (int b = blockIdx.x; b < nblk; b += gridDim.x) {
        TileCtx t;
        tile_geometry(g, b, t);
        __syncthreads();
        tile_offsets(g, t, cell_start, s_off, s_src);
        // the host sizes the tile from the previous rebuild (no extra host check per rebuild); a fuller tile is reported
        if (t.T > tile_cap) { if (threadIdx.x == 0) atomicOr(&ctl->err, CLB_EF_TILE_OVERFLOW); continue; }
        tile_stage(t, s_off, s_src, pos, s_pos, slot, s_slot, nullptr);
        const int mh0 = t.whole ? t.cx0 : 1;
        if (threadIdx.x == 0) {
            int acc = 0;
            for (int c = 0; c < t.bxe; ++c) { s_item[c] = acc; acc += (s_off[4 * t.W + mh0 + c + 1] - s_off[4 * t.W + mh0 + c] + CLB_BUILD_G - 1) / CLB_BUILD_G; }
            s_item[t.bxe] = acc;
        }
        __syncthreads();
        const int tbase = s_off[4 * t.W + mh0];
        const int nitems = s_item[t.bxe];
        const int nsegs = t.whole ? 27 : 9;       // whole-row tiles: the 3 x-neighbours may wrap -> 27 single cells
#pragma unroll 1
        for (int item = warp; item < nitems; item += nw) {
            int c = 0;
            while (c + 1 < t.bxe && s_item[c + 1] <= item) ++c;
            const int mh = mh0 + c;                                         // tile column of the home cell
            const int cell_lo = s_off[4 * t.W + mh], cell_hi = s_off[4 * t.W + mh + 1];
            const int t0 = cell_lo + (item - s_item[c]) * CLB_BUILD_G;      // tile position of the first particle
            const int np = min(CLB_BUILD_G, cell_hi - t0);
            int px[CLB_BUILD_G], py[CLB_BUILD_G], pz[CLB_BUILD_G], cnt[CLB_BUILD_G];
#pragma unroll
            for (int q = 0; q < CLB_BUILD_G; ++q) { const int4 v = s_pos[t0 + min(q, np - 1)]; px[q] = v.x; py[q] = v.y; pz[q] = v.z; cnt[q] = 0; }
            // 1. every candidate is loaded once and tested against the G home particles
#pragma unroll 1
            for (int sgm = 0; sgm < nsegs; ++sgm) {
                int lo, hi;
                if (!t.whole) { lo = s_off[sgm * t.W + mh - 1]; hi = s_off[sgm * t.W + mh + 2]; }
                else { int k = sgm / 3, m = wrapi(mh + (sgm - 3 * k) - 1, g.ncx); lo = s_off[k * t.W + m]; hi = s_off[k * t.W + m + 1]; }
#pragma unroll 1
                for (int j0 = lo; j0 < hi; j0 += 32) {
                    const int j = j0 + lane;
                    const bool valid = j < hi;
                    const int4 pj = s_pos[valid ? j : lo];
#pragma unroll
                    for (int q = 0; q < CLB_BUILD_G; ++q) {
                        const int dx = wsub(px[q], pj.x), dy = wsub(py[q], pj.y), dz = wsub(pz[q], pj.z);
                        bool pass;
                        if (CUBIC) {
                            unsigned long long r2 = (unsigned long long)((long long)dx * dx) + (unsigned long long)((long long)dy * dy) +
                                                    (unsigned long long)((long long)dz * dz);
                            pass = valid && r2 <= rl2_lat;
                        } else {
                            double fx = lat2d(dx) * geo.q[0], fy = lat2d(dy) * geo.q[1], fz = lat2d(dz) * geo.q[2];
                            pass = valid && (fx * fx + fy * fy + fz * fz) <= geo.rl2;
                        }
                        const unsigned bal = __ballot_sync(0xffffffffu, pass);
                        // ballot compaction with a PREDICATED store: some lane passes in ~95 % of the chunks, so a branch
                        // around the store is almost always taken and only adds BSSY/BRA/BSYNC overhead
                        const int o = cnt[q] + __popc(bal & lt_mask);
                        const unsigned st = (pass && o < cap) ? 1u : 0u;
                        const unsigned addr = (unsigned)__cvta_generic_to_shared(s_rows + q * cap + o);
                        asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p st.shared.u16 [%0], %1; }"
                                     :: "r"(addr), "h"((unsigned short)j), "r"(st) : "memory");
                        cnt[q] += __popc(bal);
                    }
                }
            }
            __syncwarp();
            // 2./3. per particle: drop itself and its excluded partners (swap-remove), write the row out
#pragma unroll 1
            for (int q = 0; q < np; ++q) {
                unsigned short* s_row = s_rows + q * cap;
                const int ti = t0 + q;
                const int gi = t.hs + (ti - tbase);
                int found = 0;
#pragma unroll
                for (int r = 0; r < CLB_BUILD_G; ++r) if (r == q) found = cnt[r];
                int n = min(found, cap);
                const int myslot = s_slot[ti];
                const int e0 = __ldg(excl_off + myslot), nex = __ldg(excl_off + myslot + 1) - e0;
#pragma unroll 1
                for (int e = -1; e < nex; ++e) {
                    const int x = e < 0 ? myslot : __ldg(excl_ids + e0 + e);
                    int hit = -1;
                    for (int k = lane; k < n; k += 32) if (s_slot[s_row[k]] == x) hit = k;
                    const unsigned bal = __ballot_sync(0xffffffffu, hit >= 0);
                    if (bal) {                                    // a slot appears at most once in a row
                        const int k = __shfl_sync(0xffffffffu, hit, __ffs(bal) - 1);
                        if (lane == 0) s_row[k] = s_row[n - 1];
                        --n;
                        __syncwarp();
                    }
                }
                const uint4* srow4 = reinterpret_cast<const uint4*>(s_row);
                uint4* grow4 = reinterpret_cast<uint4*>(entries + (size_t)gi * cap);
                for (int k = lane; k * 8 < n; k += 32) grow4[k] = srow4[k];
                if (lane == 0) {
                    nl_count[gi] = n;
                    if (ti - tbase < CLB_PERM_MAX) s_cnt[ti - tbase] = n;
                    if (found > cap) atomicOr(&ctl->err, CLB_EF_LIST_OVERFLOW);
                    lmax = max(lmax, found); ltot += (unsigned long long)n;
                }
            }
            __syncwarp();
        }
        __syncthreads();
        block_perm(t.hs, t.nh, s_cnt, nl_perm);
    }
    if (lane == 0 && ltot) { atomicMax(&ctl->nl_max, lmax); atomicAdd(&ctl->nl_total, ltot); }
}

// ------------------------------------------------------------------------------------------
// Neighbour-list build, second generation (round 2): same result as k_build_lists (the pair SET is bit-identical, entries
// keep the candidate order), about 2.5x fewer instructions.  The 5*10^8 candidate tests per rebuild of the 1M-bead melt hit
// only 15 % of the time, so they are split in two phases:
//   phase 1  an 8-bit PREFILTER on one DP4A per test.  Every particle carries its position inside its cell quantised to
//            84 units per cell edge (k_qsub: q = floor(frac * 84), exactly consistent with the cell assignment), the tile
//            stage adds the cell offsets, so tile-relative coordinates are integers X' in [0, 84*W), Y', Z' in [2, 254).
//            For a home cell the candidates of its 27 cells lie within +-126 units of the cell centre: signed bytes.  With
//            a = home bead, b = candidate (bytes x, 0, z, y):   |b - a|^2 - |a|^2 = DP4A(-2a, b, DP4A(b, b, 0))  and the
//            test is  <= Rq2 - |a|^2.  Floor quantisation moves each coordinate difference by less than one unit, so with
//            Rq = (rc+skin)/u + sqrt(3) the prefilter keeps a strict SUPERSET (+6 % at the melt's geometry).
//   phase 2  the survivors (~80 per bead instead of 540 candidates) get the exact 64-bit integer test, lose the bead itself
//            and its excluded partners (slot compare against the bead's exclusion row held across the lanes) and are
//            compacted in place, order preserved.
// Cubic boxes with tiles that do not wrap around x only (everything else takes k_build_lists).
#define CLB_QCELL 84
__global__ void k_qsub(int n, const int4* __restrict__ pos, ClbGrid g, unsigned* __restrict__ qsub) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 p = pos[i];
    // low word of x * nc = position inside the cell as a 2^-32 fraction (the cell index is the high word: k_cell_keys)
    const unsigned qx = __umulhi((unsigned)p.x * (unsigned)g.ncx, CLB_QCELL);
    const unsigned qy = __umulhi((unsigned)p.y * (unsigned)g.ncy, CLB_QCELL);
    const unsigned qz = __umulhi((unsigned)p.z * (unsigned)g.ncz, CLB_QCELL);
    qsub[i] = (qy << 24) | (qz << 16) | qx;
}
__global__ void __launch_bounds__(512) k_build_lists2(ClbGrid g, unsigned long long rl2_lat, int rq2,
                                                      const int* __restrict__ cell_start,
                                                      const int4* __restrict__ pos, const int* __restrict__ slot, const unsigned* __restrict__ qsub,
                                                      const int* __restrict__ excl_off, const int* __restrict__ excl_ids,
                                                      unsigned short* __restrict__ entries, int* __restrict__ nl_count,
                                                      unsigned short* __restrict__ nl_perm, int cap, int tile_cap, ClbCtl* ctl) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_off[CLB_TILE_CELLS + 1];
    __shared__ int s_src[CLB_TILE_CELLS];
    __shared__ int s_item[CLB_MAX_BX + 1];
    __shared__ int s_cnt[CLB_PERM_MAX];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    // dynamic smem: tile positions | tile slots | tile quantised words | G staging rows (cap entries each) per warp
    int4* s_pos = reinterpret_cast<int4*>(smem);
    int* s_slot = reinterpret_cast<int*>(s_pos + tile_cap);
    unsigned* s_q = reinterpret_cast<unsigned*>(s_slot + tile_cap);
    unsigned short* s_rows = reinterpret_cast<unsigned short*>(s_q + tile_cap) + (size_t)warp * CLB_BUILD_G * cap;
    int lmax = 0;
    unsigned long long ltot = 0;
    const int nblk = grid_nblocks(g);
    for (int b = blockIdx.x; b < nblk; b += gridDim.x) {
        TileCtx t;
        tile_geometry(g, b, t);
        __syncthreads();
        tile_offsets(g, t, cell_start, s_off, s_src);
        if (t.T > tile_cap) { if (threadIdx.x == 0) atomicOr(&ctl->err, CLB_EF_TILE_OVERFLOW); continue; }
        {   // stage positions, slots and the tile-relative quantised words
            const int nct = CLB_TILE_ROWS * t.W;
            for (int tc = warp; tc < nct; tc += nw) {
                const int k = tc / t.W, m = tc - k * t.W;
                const unsigned offc = ((unsigned)((k % 3) * CLB_QCELL + 2) << 24) | ((unsigned)((k / 3) * CLB_QCELL + 2) << 16) | (unsigned)(m * CLB_QCELL);
                const int o = s_off[tc], c = s_off[tc + 1] - o, sgl = s_src[tc];
                for (int i = lane; i < c; i += 32) {
                    s_pos[o + i] = __ldg(pos + sgl + i);
                    s_slot[o + i] = __ldg(slot + sgl + i);
                    s_q[o + i] = (__ldg(qsub + sgl + i) + offc) ^ 0x80800000u;
                }
            }
        }
        const int mh0 = t.whole ? t.cx0 : 1;
        if (threadIdx.x == 0) {
            int acc = 0;
            for (int c = 0; c < t.bxe; ++c) { s_item[c] = acc; acc += (s_off[4 * t.W + mh0 + c + 1] - s_off[4 * t.W + mh0 + c] + CLB_BUILD_G - 1) / CLB_BUILD_G; }
            s_item[t.bxe] = acc;
        }
        __syncthreads();
        const int tbase = s_off[4 * t.W + mh0];
        const int nitems = s_item[t.bxe];
#pragma unroll 1
        for (int item = warp; item < nitems; item += nw) {
            int c = 0;
            while (c + 1 < t.bxe && s_item[c + 1] <= item) ++c;
            const int mh = mh0 + c;
            const int cell_lo = s_off[4 * t.W + mh], cell_hi = s_off[4 * t.W + mh + 1];
            const int t0 = cell_lo + (item - s_item[c]) * CLB_BUILD_G;
            const int np = min(CLB_BUILD_G, cell_hi - t0);
            int an2[CLB_BUILD_G], cq[CLB_BUILD_G], cnt[CLB_BUILD_G];
            const int xbias = 128 - (mh * CLB_QCELL + CLB_QCELL / 2);
#pragma unroll
            for (int q = 0; q < CLB_BUILD_G; ++q) {
                const unsigned w = s_q[t0 + min(q, np - 1)];
                const int ax = (int)(w & 0xffffu) + xbias - 128, az = (int)(signed char)(w >> 16), ay = (int)(signed char)(w >> 24);
                an2[q] = (int)(((unsigned)(-2 * ax) & 0xffu) | (((unsigned)(-2 * az) & 0xffu) << 16) | (((unsigned)(-2 * ay) & 0xffu) << 24));
                cq[q] = rq2 - (ax * ax + ay * ay + az * az);
                cnt[q] = 0;
            }
            // phase 1: prefilter, ballot compaction into the staging rows
#pragma unroll 1
            for (int sgm = 0; sgm < CLB_TILE_ROWS; ++sgm) {
                const int lo = s_off[sgm * t.W + mh - 1], hi = s_off[sgm * t.W + mh + 2];
#pragma unroll 1
                for (int j0 = lo; j0 < hi; j0 += 32) {
                    const int j = j0 + lane;
                    const bool valid = j < hi;
                    const unsigned bw = (s_q[valid ? j : lo] + (unsigned)xbias) ^ 0x80u;     // bytes {x, 0, z, y} relative to the home cell centre
                    const int bb = valid ? __dp4a((int)bw, (int)bw, 0) : 0x3fffffff;
#pragma unroll
                    for (int q = 0; q < CLB_BUILD_G; ++q) {
                        const bool pass = __dp4a(an2[q], (int)bw, bb) <= cq[q];
                        const unsigned bal = __ballot_sync(0xffffffffu, pass);
                        const int o = cnt[q] + __popc(bal & lt_mask);
                        const unsigned st = (pass && o < cap) ? 1u : 0u;
                        const unsigned addr = (unsigned)__cvta_generic_to_shared(s_rows + q * cap + o);
                        asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p st.shared.u16 [%0], %1; }"
                                     :: "r"(addr), "h"((unsigned short)j), "r"(st) : "memory");
                        cnt[q] += __popc(bal);
                    }
                }
            }
            __syncwarp();
            // phase 2: exact test, self / exclusions, in-place compaction, row output
#pragma unroll 1
            for (int q = 0; q < np; ++q) {
                unsigned short* s_row = s_rows + q * cap;
                const int ti = t0 + q;
                const int gi = t.hs + (ti - tbase);
                int found = 0;
#pragma unroll
                for (int r = 0; r < CLB_BUILD_G; ++r) if (r == q) found = cnt[r];
                const int n1 = min(found, cap);
                const int4 pi = s_pos[ti];
                const int myslot = s_slot[ti];
                const int e0 = __ldg(excl_off + myslot), nex = __ldg(excl_off + myslot + 1) - e0;
                const int exid = lane < nex ? __ldg(excl_ids + e0 + lane) : -1;      // first 32 excluded partners across the lanes
                int n2 = 0;
#pragma unroll 1
                for (int k0 = 0; k0 < n1; k0 += 32) {
                    const int k = k0 + lane;
                    const bool valid = k < n1;
                    const unsigned e = s_row[valid ? k : 0];
                    const int4 pj = s_pos[e];
                    const int dx = wsub(pi.x, pj.x), dy = wsub(pi.y, pj.y), dz = wsub(pi.z, pj.z);
                    const unsigned long long r2 = (unsigned long long)((long long)dx * dx) + (unsigned long long)((long long)dy * dy) +
                                                  (unsigned long long)((long long)dz * dz);
                    const int sj = s_slot[e];
                    bool keep = valid && r2 <= rl2_lat && (int)e != ti;
                    const int nx32 = min(nex, 32);
                    for (int x = 0; x < nx32; ++x) { const int xid = __shfl_sync(0xffffffffu, exid, x); keep = keep && (sj != xid); }   // all lanes shuffle
                    for (int x = 32; x < nex; ++x) { const int xid = __ldg(excl_ids + e0 + x); keep = keep && (sj != xid); }          // more than 32 exclusions: rare
                    const unsigned bal = __ballot_sync(0xffffffffu, keep);
                    if (keep) s_row[n2 + __popc(bal & lt_mask)] = (unsigned short)e;      // position <= k: never ahead of an unread entry
                    n2 += __popc(bal);
                    __syncwarp();
                }
                const uint4* srow4 = reinterpret_cast<const uint4*>(s_row);
                uint4* grow4 = reinterpret_cast<uint4*>(entries + (size_t)gi * cap);
                for (int k = lane; k * 8 < n2; k += 32) grow4[k] = srow4[k];
                if (lane == 0) {
                    nl_count[gi] = n2;
                    if (ti - tbase < CLB_PERM_MAX) s_cnt[ti - tbase] = n2;
                    if (found > cap) atomicOr(&ctl->err, CLB_EF_LIST_OVERFLOW);
                    lmax = max(lmax, found > cap ? found : n2); ltot += (unsigned long long)n2;
                }
            }
            __syncwarp();
        }
        __syncthreads();
        block_perm(t.hs, t.nh, s_cnt, nl_perm);
    }
    if (lane == 0 && ltot) { atomicMax(&ctl->nl_max, lmax); atomicAdd(&ctl->nl_total, ltot); }
}

// ------------------------------------------------------------------------------------------
// Pair forces.  Full (two-sided) list: no scatter, no atomics, deterministic.  Everything after the
// integer subtraction is fp64 (B200 has a full-rate fp64 pipe that issues beside the fp32/int pipes);
// no float<->double conversion instructions:
//   d       exact lattice difference -> double by the 2^52 trick (lat2d)
//   1/r     MUFU.RSQ seed on the truncated high word + one fp64 Newton step
//   index   u = (r-x0)/dx - 1/2 ; idx = round(u) and fraction b' = u - idx by the 1.5*2^52 trick
//   F(r)    (f[i] + df[i]/2) + b'*df[i]  (reference: linear interpolation itype=1)
// Table rows {f_i + df_i/2, df_i} as double2 live in shared memory (persistent CTAs load them once).
// SPLIT warps share one particle: warp group s takes the 8-entry batches b with b % SPLIT == s, partial
// sums are combined through shared memory in fixed order (bit-reproducible) -> more resident working
// warps per SM for the same shared-memory footprint.
__device__ __forceinline__ double rsqrt_seed(double x) {
    // float with the same value as x truncated to 24 bits, via integer ops only (no F2F conversions)
    int hi = __double2hiint(x), lo = __double2loint(x);
    unsigned fb = ((unsigned)(hi - 0x38000000) << 3) | ((unsigned)lo >> 29);
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(__uint_as_float(fb)));
    unsigned yb = __float_as_uint(y);
    return __hiloint2double((int)((yb >> 3) + 0x38000000u), (int)(yb << 29));
}

struct ClbPairArgs {
    const int* cell_start; const int4* pos; const unsigned short* entries; const int* nl_count;
    const ClbPairDesc* pdesc; const double2* plj; const ClbTabMeta* tmeta; const double2* trows;
    double* force; ClbCtl* ctl;
    int cap, ntypes, ntabs, nrows_total, fstride, npw;   // npw = warps per split group
    ClbTabMeta ugrid;                                      // common grid when UGRID
    int b0, seg0, b1, nidx;   // row blocks of this launch: idx < seg0 -> b0 + idx, else b1 + (idx - seg0); idx in [0, nidx)
};
// CUBIC boxes work in lattice units end to end: r2 and the cutoffs are lattice^2, the table index uses
// invdx*q, and F(r)*(1/r_lat)*d_lat is already the real force vector -- no per-pair unit conversion.
// The host pre-scales ClbPairDesc.rc2, ClbTabMeta.invdx and the LJ coefficients accordingly.
template <bool CUBIC, bool TABS_SMEM, bool UGRID, int SPLIT>
__global__ void __launch_bounds__(512) k_pair_forces(ClbGrid g, ClbGeom geo, ClbPairArgs A) {
    if (*(volatile int*)&A.ctl->stall) return;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_off[CLB_TILE_CELLS + 1];
    __shared__ int s_src[CLB_TILE_CELLS];
    const int ntp = A.ntypes * A.ntypes;
    ClbPairDesc* s_pd = reinterpret_cast<ClbPairDesc*>(smem);
    double2* s_lj = reinterpret_cast<double2*>(s_pd + ntp);
    ClbTabMeta* s_tm = reinterpret_cast<ClbTabMeta*>(s_lj + ntp);
    double2* s_rows = reinterpret_cast<double2*>(s_tm + A.ntabs);
    double* s_red = reinterpret_cast<double*>(s_rows + (TABS_SMEM ? A.nrows_total : 0));
    int4* s_pos = reinterpret_cast<int4*>(s_red + (SPLIT - 1) * 3 * A.npw * 32);
    for (int i = threadIdx.x; i < ntp; i += blockDim.x) { s_pd[i] = A.pdesc[i]; s_lj[i] = A.plj[i]; }
    for (int i = threadIdx.x; i < A.ntabs; i += blockDim.x) s_tm[i] = A.tmeta[i];
    if (TABS_SMEM) for (int i = threadIdx.x; i < A.nrows_total; i += blockDim.x) s_rows[i] = __ldg(A.trows + i);
    const double2* rows = TABS_SMEM ? s_rows : A.trows;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sg = SPLIT > 1 ? warp / A.npw : 0;          // split group of this warp
    const int wp = SPLIT > 1 ? warp - sg * A.npw : warp;  // particle warp inside the group
    const int nhpass = A.npw * 32;
    const double u_invdx = A.ugrid.invdx, u_ct = A.ugrid.c_t;
    const unsigned u_n = (unsigned)A.ugrid.n;
    unsigned err = 0;
    for (int idx = blockIdx.x; idx < A.nidx; idx += gridDim.x) {
        const int b = idx < A.seg0 ? A.b0 + idx : A.b1 + (idx - A.seg0);
        TileCtx t;
        tile_geometry(g, b, t);
        __syncthreads();
        tile_offsets(g, t, A.cell_start, s_off, s_src);
        tile_stage(t, s_off, s_src, A.pos, s_pos, nullptr, nullptr, nullptr);
        __syncthreads();
        for (int p0 = 0; p0 < t.nh; p0 += nhpass) {
            const int pl = wp * 32 + lane;                  // particle index inside this pass
            const int p = p0 + pl;
            const bool act = p < t.nh;
            const int gi = t.hs + (act ? p : 0);
            const int4 pi = __ldg(A.pos + gi);
            // bias by 2^31 once so that (pi - pj) is directly the low word of the 2^52 trick
            const unsigned pix = (unsigned)pi.x + 0x80000000u, piy = (unsigned)pi.y + 0x80000000u, piz = (unsigned)pi.z + 0x80000000u;
            const int cnt = act ? __ldg(A.nl_count + gi) : 0;
            const int trow = pw_type(pi.w) * A.ntypes;
            const uint4* row = reinterpret_cast<const uint4*>(A.entries + (size_t)gi * A.cap);
            const int nb = (cnt + 7) >> 3;                   // 8-entry batches of this particle
            double ax = 0.0, ay = 0.0, az = 0.0;
            uint4 ev = make_uint4(0, 0, 0, 0);
            if (sg < nb) ev = __ldg(row + sg);
            for (int bi = sg; bi < nb; bi += SPLIT) {
                uint4 cur = ev;
                if (bi + SPLIT < nb) ev = __ldg(row + bi + SPLIT);     // prefetch the next batch
                const int ne = min(8, cnt - bi * 8);
#pragma unroll 1
                for (int u = 0; u < ne; ++u) {
                    const unsigned e = cur.x & 0xffffu;
                    cur.x = __funnelshift_r(cur.x, cur.y, 16); cur.y = __funnelshift_r(cur.y, cur.z, 16);
                    cur.z = __funnelshift_r(cur.z, cur.w, 16); cur.w >>= 16;
                    const int4 pj = s_pos[e];
                    double dx = __hiloint2double(0x43300000, (int)(pix - (unsigned)pj.x)) - 4503601774854144.0;
                    double dy = __hiloint2double(0x43300000, (int)(piy - (unsigned)pj.y)) - 4503601774854144.0;
                    double dz = __hiloint2double(0x43300000, (int)(piz - (unsigned)pj.z)) - 4503601774854144.0;
                    if (!CUBIC) { dx *= geo.q[0]; dy *= geo.q[1]; dz *= geo.q[2]; }
                    const double r2 = dx * dx + dy * dy + dz * dz;
                    const int tp = trow + pw_type(pj.w);
                    const ClbPairDesc pd = s_pd[tp];
                    if (r2 <= pd.rc2) {                         // rc2 < 0 for pairs without a potential (U2)
                        double y = rsqrt_seed(r2);
                        double h = r2 * y;
                        double ee = fma(-h, y, 1.0);
                        y = fma(0.5 * y, ee, y);              // 1/r to ~1e-14
                        double fr;
                        if (pd.kind == 1) {
                            double r = r2 * y;
                            double invdx, c_t; unsigned n; int off;
                            if (UGRID) { invdx = u_invdx; c_t = u_ct; n = u_n; off = pd.tab; }
                            else { const ClbTabMeta tm = s_tm[pd.tab]; invdx = tm.invdx; c_t = tm.c_t; n = (unsigned)tm.n; off = tm.off; }
                            double uu = fma(r, invdx, c_t);
                            double ti = uu + 6755399441055744.0;            // 1.5 * 2^52: integer in the low word
                            int idx = __double2loint(ti);
                            double bfrac = uu - (ti - 6755399441055744.0);
                            // row n-1 exists ({f[n-1], 0}) so r == table end is exact; anything else is fatal (U12)
                            if ((unsigned)idx >= n) { err |= CLB_EF_TABLE_RANGE; idx = 0; }
                            const double2 rw = rows[off + idx];                // {f_i + df_i/2, df_i}
                            fr = fma(bfrac, rw.y, rw.x) * y;
                        } else {
                            const double2 lj = s_lj[tp];
                            double y2 = y * y;
                            double y6 = y2 * y2 * y2;
                            fr = y6 * fma(lj.x, y6, -lj.y) * y2;
                        }
                        ax = fma(fr, dx, ax); ay = fma(fr, dy, ay); az = fma(fr, dz, az);
                    }
                }
            }
            if (SPLIT > 1) {
                if (sg > 0) { double* r = s_red + ((sg - 1) * nhpass + pl) * 3; r[0] = ax; r[1] = ay; r[2] = az; }
                __syncthreads();
                if (sg == 0) {
#pragma unroll
                    for (int s2 = 1; s2 < SPLIT; ++s2) { const double* r = s_red + ((s2 - 1) * nhpass + pl) * 3; ax += r[0]; ay += r[1]; az += r[2]; }
                }
            }
            if (act && sg == 0) { A.force[gi] = ax; A.force[gi + A.fstride] = ay; A.force[gi + 2 * A.fstride] = az; }
            if (SPLIT > 1) __syncthreads();
        }
    }
    if (err) atomicOr(&A.ctl->err, err);
}

// ------------------------------------------------------------------------------------------
// Branch-free variant for systems whose pair potentials are ALL tabulated (every shipped chemlab
// example but atrp_lj).  The 8 entries of a batch are evaluated as straight-line code without any
// divergent branch (cutoff and list-tail handled by selects), so the compiler interleaves the eight
// independent fp64 dependency chains -> instruction-level parallelism instead of more resident warps.
template <bool CUBIC, bool TABS_SMEM, bool UGRID, int SPLIT>
__global__ void __launch_bounds__(512) k_pair_forces_tab(ClbGrid g, ClbGeom geo, ClbPairArgs A) {
    if (*(volatile int*)&A.ctl->stall) return;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_off[CLB_TILE_CELLS + 1];
    __shared__ int s_src[CLB_TILE_CELLS];
    const int ntp = A.ntypes * A.ntypes;
    ClbPairDesc* s_pd = reinterpret_cast<ClbPairDesc*>(smem);
    double2* s_lj = reinterpret_cast<double2*>(s_pd + ntp);
    ClbTabMeta* s_tm = reinterpret_cast<ClbTabMeta*>(s_lj + ntp);
    double2* s_rows = reinterpret_cast<double2*>(s_tm + A.ntabs);
    double* s_red = reinterpret_cast<double*>(s_rows + (TABS_SMEM ? A.nrows_total : 0));
    int4* s_pos = reinterpret_cast<int4*>(s_red + (SPLIT - 1) * 3 * A.npw * 32);
    for (int i = threadIdx.x; i < ntp; i += blockDim.x) s_pd[i] = A.pdesc[i];
    for (int i = threadIdx.x; i < A.ntabs; i += blockDim.x) s_tm[i] = A.tmeta[i];
    if (TABS_SMEM) for (int i = threadIdx.x; i < A.nrows_total; i += blockDim.x) s_rows[i] = __ldg(A.trows + i);
    const double2* rows = TABS_SMEM ? s_rows : A.trows;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sg = SPLIT > 1 ? warp / A.npw : 0;
    const int wp = SPLIT > 1 ? warp - sg * A.npw : warp;
    const int nhpass = A.npw * 32;
    const double u_invdx = A.ugrid.invdx, u_ct = A.ugrid.c_t;
    const unsigned u_nm1 = (unsigned)A.ugrid.n - 1u;
    unsigned err = 0;
    for (int idx = blockIdx.x; idx < A.nidx; idx += gridDim.x) {
        const int b = idx < A.seg0 ? A.b0 + idx : A.b1 + (idx - A.seg0);
        TileCtx t;
        tile_geometry(g, b, t);
        __syncthreads();
        tile_offsets(g, t, A.cell_start, s_off, s_src);
        tile_stage(t, s_off, s_src, A.pos, s_pos, nullptr, nullptr, nullptr);
        __syncthreads();
        for (int p0 = 0; p0 < t.nh; p0 += nhpass) {
            const int pl = wp * 32 + lane;
            const int p = p0 + pl;
            const bool act = p < t.nh;
            const int gi = t.hs + (act ? p : 0);
            const int4 pi = __ldg(A.pos + gi);
            const unsigned pix = (unsigned)pi.x + 0x80000000u, piy = (unsigned)pi.y + 0x80000000u, piz = (unsigned)pi.z + 0x80000000u;
            const int cnt = act ? __ldg(A.nl_count + gi) : 0;
            const int trow = pw_type(pi.w) * A.ntypes;
            const uint4* row = reinterpret_cast<const uint4*>(A.entries + (size_t)gi * A.cap);
            const int nb = (cnt + 7) >> 3;
            double ax = 0.0, ay = 0.0, az = 0.0;
            uint4 ev = make_uint4(0, 0, 0, 0);
            if (sg < nb) ev = __ldg(row + sg);
#pragma unroll 1
            for (int bi = sg; bi < nb; bi += SPLIT) {
                const uint4 cur = ev;
                if (bi + SPLIT < nb) ev = __ldg(row + bi + SPLIT);
                const int ne = cnt - bi * 8;                      // >= 1; entries beyond ne are stale
                const unsigned wds[4] = {cur.x, cur.y, cur.z, cur.w};
                // two entries per stage, written stage-major so that the two fp64 chains interleave
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    bool live[2], in[2];
                    unsigned e[2], idc[2];
                    int4 pj[2];
                    ClbPairDesc pd[2];
                    double dx[2], dy[2], dz[2], r2[2], y[2], uu[2], ti[2], bfrac[2], fr[2];
                    double2 rw[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        live[q] = (2 * w + q) < ne;
                        e[q] = q ? (wds[w] >> 16) : (wds[w] & 0xffffu);
                        e[q] = live[q] ? e[q] : 0u;
                        pj[q] = s_pos[e[q]];
                    }
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        dx[q] = __hiloint2double(0x43300000, (int)(pix - (unsigned)pj[q].x)) - 4503601774854144.0;
                        dy[q] = __hiloint2double(0x43300000, (int)(piy - (unsigned)pj[q].y)) - 4503601774854144.0;
                        dz[q] = __hiloint2double(0x43300000, (int)(piz - (unsigned)pj[q].z)) - 4503601774854144.0;
                        if (!CUBIC) { dx[q] *= geo.q[0]; dy[q] *= geo.q[1]; dz[q] *= geo.q[2]; }
                        pd[q] = s_pd[trow + pw_type(pj[q].w)];
                    }
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        r2[q] = dx[q] * dx[q] + dy[q] * dy[q] + dz[q] * dz[q];
                        in[q] = live[q] && (r2[q] <= pd[q].rc2);   // rc2 < 0: no potential for this type pair
                        r2[q] = in[q] ? r2[q] : 1.0;                // keeps the masked lanes' arithmetic finite
                        y[q] = rsqrt_seed(r2[q]);
                    }
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const double h = r2[q] * y[q];
                        const double ee = fma(-h, y[q], 1.0);
                        y[q] = fma(0.5 * y[q], ee, y[q]);
                    }
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const double r = r2[q] * y[q];
                        double invdx, c_t; unsigned nm1; int off;
                        if (UGRID) { invdx = u_invdx; c_t = u_ct; nm1 = u_nm1; off = pd[q].tab; }
                        else { const ClbTabMeta tm = s_tm[pd[q].tab]; invdx = tm.invdx; c_t = tm.c_t; nm1 = (unsigned)tm.n - 1u; off = tm.off; }
                        uu[q] = fma(r, invdx, c_t);
                        ti[q] = uu[q] + 6755399441055744.0;
                        const unsigned idx = (unsigned)__double2loint(ti[q]);
                        idc[q] = min(idx, nm1);                     // row n-1 = {f[n-1], 0}: r == table end is exact
                        if (in[q] && idx > nm1) err |= CLB_EF_TABLE_RANGE;   // predicated OR, no divergence (U12)
                        rw[q] = rows[off + (in[q] ? idc[q] : 0u)];
                    }
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        bfrac[q] = uu[q] - (ti[q] - 6755399441055744.0);
                        fr[q] = fma(bfrac[q], rw[q].y, rw[q].x) * y[q];
                        fr[q] = in[q] ? fr[q] : 0.0;
                    }
#pragma unroll
                    for (int q = 0; q < 2; ++q) { ax = fma(fr[q], dx[q], ax); ay = fma(fr[q], dy[q], ay); az = fma(fr[q], dz[q], az); }
                }
            }
            if (SPLIT > 1) {
                if (sg > 0) { double* r = s_red + ((sg - 1) * nhpass + pl) * 3; r[0] = ax; r[1] = ay; r[2] = az; }
                __syncthreads();
                if (sg == 0) {
#pragma unroll
                    for (int s2 = 1; s2 < SPLIT; ++s2) { const double* r = s_red + ((s2 - 1) * nhpass + pl) * 3; ax += r[0]; ay += r[1]; az += r[2]; }
                }
            }
            if (act && sg == 0) { A.force[gi] = ax; A.force[gi + A.fstride] = ay; A.force[gi + 2 * A.fstride] = az; }
            if (SPLIT > 1) __syncthreads();
        }
    }
    if (err) atomicOr(&A.ctl->err, err);
}

// ------------------------------------------------------------------------------------------
// Third generation of the all-tabulated kernel (cubic box + one common table grid; the benchmark melt and every
// shipped chemlab example with a cubic box).  Same algorithm and results as k_pair_forces_tab, fewer instructions:
//   * 16 fp64 operations per listed pair instead of 20:
//       1/r   : y0 = MUFU.RSQ seed, yh = y0/2 by an exponent decrement (free), h = r2*y0, e = fma(-h, yh, 1/2),
//               y = fma(y0, e, y0), r = fma(h, e, h)                                  (4 instead of 5)
//       table : rows are stored as {A_i, B_i} with F(r) = A_i + r*B_i on [r_i, r_i+1) -- the SAME straight line as
//               the reference's (1-b) f_i + b f_i+1 (U12) -- and the index comes from the low word of
//               fma.rd(r, 1/dx, 1.5*2^52 - x0/dx) = floor((r-x0)/dx): no fraction, no back-subtraction (2 instead of 5)
//   * one 64-bit select per pair (on the final force factor; the row index of a masked pair is merely clamped) instead of two;
//   * ONEPD: when every type pair shares one table and one cutoff the descriptor lives in registers -> one
//     shared-memory gather less per pair (the kernel is shared-memory-pipe bound);
//   * NI entries are evaluated in lock step (stage-major source) so that their dependency chains interleave.
struct ClbPairArgs2 {
    const int* cell_start; const int4* pos; const unsigned short* entries; const int* nl_count;
    const double2* pd2;      // per type pair {rc2 (lattice^2; < 0: no potential), 1.5*2^52 + first row (row = low word)}
    const double2* trows;    // {A_i, B_i} rows of all tables, lattice units
    double* force; ClbCtl* ctl;
    int cap, ntypes, nrows_total, fstride, npw;
    double invdx, cmagic;    // 1/dx (lattice units) and 1.5*2^52 - x0/dx (x0/dx integral: checked by the host)
    unsigned nm1;            // rows - 1 of every table
    double one_rc2; int one_off;   // ONEPD descriptor
    int b0, seg0, b1, nidx;        // row blocks of this launch (see ClbPairArgs)
    int nv, vc_bytes;              // virtual CTAs per CTA and their shared-memory stride (offset tables + tile)
};
__device__ __forceinline__ void rsqrt_seed2(double x, double& y0, double& yh) {
    int hi = __double2hiint(x), lo = __double2loint(x);
    unsigned fb = ((unsigned)(hi - 0x38000000) << 3) | ((unsigned)lo >> 29);
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(__uint_as_float(fb)));
    unsigned yb = __float_as_uint(y);
    int h = (int)((yb >> 3) + 0x38000000u), l = (int)(yb << 29);
    y0 = __hiloint2double(h, l);
    yh = __hiloint2double(h - 0x00100000, l);      // y0 / 2
}
template <bool TABS_SMEM, bool ONEPD, int NI>
__global__ void __launch_bounds__(1024) k_pair_forces_tab2(ClbGrid g, ClbPairArgs2 A) {
    if (*(volatile int*)&A.ctl->stall) return;
    extern __shared__ __align__(16) unsigned char smem[];
    const int ntp = A.ntypes * A.ntypes;
    // shared by the whole CTA: descriptors + table rows; per virtual CTA: offset tables + tile positions
    double2* s_pd = reinterpret_cast<double2*>(smem);
    double2* s_rows = s_pd + (ONEPD ? 0 : ntp);
    unsigned char* vc_base = reinterpret_cast<unsigned char*>(s_rows + (TABS_SMEM ? A.nrows_total : 0));
    const int nth = A.npw * 32;                               // threads of one virtual CTA
    const int vc = threadIdx.x / nth, tid = threadIdx.x - vc * nth;
    const int bar = 1 + vc;
    int* s_off = reinterpret_cast<int*>(vc_base + (size_t)vc * A.vc_bytes);
    int* s_src = s_off + (CLB_TILE_CELLS + 4);
    int4* s_pos = reinterpret_cast<int4*>(s_src + CLB_TILE_CELLS);
    if (!ONEPD) for (int i = threadIdx.x; i < ntp; i += blockDim.x) s_pd[i] = A.pd2[i];
    if (TABS_SMEM) for (int i = threadIdx.x; i < A.nrows_total; i += blockDim.x) s_rows[i] = __ldg(A.trows + i);
    __syncthreads();
    const double2* rows = TABS_SMEM ? s_rows : A.trows;
    const int lane = tid & 31, warp = tid >> 5;
    const int nhpass = nth;
    const double invdx = A.invdx, cmagic = A.cmagic;
    const unsigned nm1 = A.nm1;
    unsigned err = 0;
    // virtual CTA v of CTA c takes tiles v*gridDim + c, + nv*gridDim, ...: the tiles of the last, partial round land on
    // DIFFERENT SMs (one each) instead of filling a few SMs completely
    for (int idx = vc * gridDim.x + blockIdx.x; idx < A.nidx; idx += gridDim.x * A.nv) {
        const int b = idx < A.seg0 ? A.b0 + idx : A.b1 + (idx - A.seg0);
        TileCtx t;
        tile_geometry(g, b, t);
        vc_sync(bar, nth);
        tile_offsets_vc(g, t, A.cell_start, s_off, s_src, tid, nth, bar);
        tile_stage_vc(t, s_off, s_src, A.pos, s_pos, tid, nth);
        vc_sync(bar, nth);
        for (int p0 = 0; p0 < t.nh; p0 += nhpass) {
            const int p = p0 + warp * 32 + lane;
            const bool act = p < t.nh;
            const int gi = t.hs + (act ? p : 0);
            const int4 pi = __ldg(A.pos + gi);
            const unsigned pix = (unsigned)pi.x + 0x80000000u, piy = (unsigned)pi.y + 0x80000000u, piz = (unsigned)pi.z + 0x80000000u;
            const int cnt = act ? __ldg(A.nl_count + gi) : 0;
            const int trow = pw_type(pi.w) * A.ntypes;
            const uint4* row = reinterpret_cast<const uint4*>(A.entries + (size_t)gi * A.cap);
            const int nb = (cnt + 7) >> 3;
            double ax = 0.0, ay = 0.0, az = 0.0;
            uint4 ev = make_uint4(0, 0, 0, 0);
            if (nb > 0) ev = __ldg(row);
#pragma unroll 1
            for (int bi = 0; bi < nb; ++bi) {
                const uint4 cur = ev;
                if (bi + 1 < nb) ev = __ldg(row + bi + 1);
                const int ne = cnt - bi * 8;                      // >= 1; entries beyond ne are stale
                const unsigned wds[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
                for (int s0 = 0; s0 < 8; s0 += NI) {
                    bool in[NI];
                    int4 pj[NI];
                    double dx[NI], dy[NI], dz[NI], r2[NI], y0[NI], yh[NI], y[NI], r[NI], rc2[NI], ti[NI], F[NI];
                    int off[NI];
                    double2 rw[NI];
#pragma unroll
                    for (int q = 0; q < NI; ++q) {
                        const int k = s0 + q;
                        unsigned e = (k & 1) ? (wds[k >> 1] >> 16) : (wds[k >> 1] & 0xffffu);
                        in[q] = k < ne;
                        e = in[q] ? e : 0u;
                        pj[q] = s_pos[e];
                    }
#pragma unroll
                    for (int q = 0; q < NI; ++q) {
                        dx[q] = __hiloint2double(0x43300000, (int)(pix - (unsigned)pj[q].x)) - 4503601774854144.0;
                        dy[q] = __hiloint2double(0x43300000, (int)(piy - (unsigned)pj[q].y)) - 4503601774854144.0;
                        dz[q] = __hiloint2double(0x43300000, (int)(piz - (unsigned)pj[q].z)) - 4503601774854144.0;
                        if (ONEPD) { rc2[q] = A.one_rc2; off[q] = A.one_off; }
                        else { const double2 d = s_pd[trow + pw_type(pj[q].w)]; rc2[q] = d.x; off[q] = __double2loint(d.y); }
                    }
#pragma unroll
                    for (int q = 0; q < NI; ++q) r2[q] = fma(dz[q], dz[q], fma(dy[q], dy[q], dx[q] * dx[q]));
#pragma unroll
                    for (int q = 0; q < NI; ++q) { in[q] = in[q] && (r2[q] <= rc2[q]); rsqrt_seed2(r2[q], y0[q], yh[q]); }
#pragma unroll
                    for (int q = 0; q < NI; ++q) {
                        const double h = r2[q] * y0[q];
                        const double ee = fma(-h, yh[q], 0.5);
                        y[q] = fma(y0[q], ee, y0[q]);             // 1/r
                        r[q] = fma(h, ee, h);                     // r
                    }
#pragma unroll
                    for (int q = 0; q < NI; ++q) {
                        ti[q] = __fma_rd(r[q], invdx, cmagic);    // round down: the low word is floor((r - x0)/dx) exactly
                        const unsigned idx = (unsigned)__double2loint(ti[q]);
                        if (in[q] && idx > nm1) err |= CLB_EF_TABLE_RANGE;       // fatal in the reference (U12)
                        rw[q] = rows[off[q] + (int)min(idx, nm1)];
                    }
#pragma unroll
                    for (int q = 0; q < NI; ++q) {
                        // outside the cutoff / list tail: whatever was computed from the (clamped) row is discarded here
                        F[q] = fma(r[q], rw[q].y, rw[q].x) * y[q];
                        F[q] = in[q] ? F[q] : 0.0;
                    }
#pragma unroll
                    for (int q = 0; q < NI; ++q) { ax = fma(F[q], dx[q], ax); ay = fma(F[q], dy[q], ay); az = fma(F[q], dz[q], az); }
                }
            }
            if (act) { A.force[gi] = ax; A.force[gi + A.fstride] = ay; A.force[gi + 2 * A.fstride] = az; }
        }
    }
    if (err) atomicOr(&A.ctl->err, err);
}


// ---- TMA bulk staging of a tile (round 2) -------------------------------------------------------------------------
// With x-fastest cells a tile ROW (the W cells of one (dy,dz)) is ONE contiguous range of the sorted particle arrays
// (two when the row wraps around the periodic x boundary), so a whole tile is the concatenation of at most 18 contiguous
// global ranges.  The pair kernel needs no per-cell offsets: warp 0 reads the 18 range bounds from cell_start, a warp scan
// gives the tile offsets, and lanes 0..17 each issue ONE cp.async.bulk (global -> shared, completion on an mbarrier).  The
// other warps never touch the staging: they wait on the mbarrier.  (Round 1 staged cell by cell through registers: 15
// dependent global-load latencies per warp and a serial prefix over 90 cells, 20 % of the kernel in barrier stalls.)
#define CLB_TILE_RANGES 18
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* mbar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* mbar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* mbar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "CLB_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra CLB_MBAR_DONE;\n"
        "bra CLB_MBAR_WAIT;\n"
        "CLB_MBAR_DONE:\n"
        "}\n" :: "r"(smem_u32(mbar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(mbar)) : "memory");
}
struct TileMeta { int hs, nh, T, pad; };   // per virtual CTA, in shared memory behind the mbarrier
// Warp 0 of a virtual CTA (all 32 lanes) computes the ranges of tile `b` and launches the bulk copies; returns nothing: the
// caller synchronises the group, reads *meta and waits on the mbarrier.  tile order = rows k = 0..8, cells m = 0..W-1 (tile_cell).
__device__ __forceinline__ void tile_stage_bulk(const ClbGrid& g, const TileCtx& t, const int* __restrict__ cell_start,
                                                const int4* __restrict__ pos, int4* s_pos, TileMeta* meta,
                                                unsigned long long* mbar, int lane) {
    int cnt = 0, gs = 0;
    if (lane < CLB_TILE_RANGES) {
        const int k = lane >> 1, half = lane & 1;
        const int dy = k % 3 - 1, dz = k / 3 - 1;
        const int rowbase = (wrapi(t.lz + dz, g.nplanes) * g.ncy + wrapi(t.cy + dy, g.ncy)) * g.ncx;
        int c0, c1;
        if (t.whole) { c0 = 0; c1 = half ? 0 : g.ncx; }
        else {
            const int a0 = wrapi(t.cx0 - 1, g.ncx), lenA = min(t.W, g.ncx - a0);
            if (!half) { c0 = a0; c1 = a0 + lenA; } else { c0 = 0; c1 = t.W - lenA; }
        }
        if (c1 > c0) { gs = __ldg(cell_start + rowbase + c0); cnt = __ldg(cell_start + rowbase + c1) - gs; }
    } else if (lane == CLB_TILE_RANGES) {
        const int c = (t.lz * g.ncy + t.cy) * g.ncx + t.cx0;
        gs = __ldg(cell_start + c); cnt = __ldg(cell_start + c + t.bxe) - gs;     // home range (not part of the scan)
    }
    int incl = lane < CLB_TILE_RANGES ? cnt : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
    const int T = __shfl_sync(0xffffffffu, incl, CLB_TILE_RANGES - 1);
    const int ts = incl - cnt;
    const int hs_ = __shfl_sync(0xffffffffu, gs, CLB_TILE_RANGES), nh_ = __shfl_sync(0xffffffffu, cnt, CLB_TILE_RANGES);
    if (lane == 0) { meta->hs = hs_; meta->nh = nh_; meta->T = T; }   // written by the arriving thread: the mbarrier release publishes it
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // earlier generic reads of the tile are ordered before the async writes
    if (lane == 0) mbar_expect_tx(mbar, (unsigned)T * 16u);
    __syncwarp();
    if (lane < CLB_TILE_RANGES && cnt > 0) bulk_g2s(s_pos + ts, pos + gs, (unsigned)cnt * 16u, mbar);
}

// ------------------------------------------------------------------------------------------
// Fourth generation (round 2): the all-tabulated kernel for MANY tables.  Same arithmetic as k_pair_forces_tab2 (bit-identical
// forces), different table storage:
//   * WINDOWS: a pair only ever reads the rows between the repulsive wall it cannot climb (U - U_min > 30 kT, chosen by the host
//     from the energy column) and its cutoff -- about 600 of 1750 rows for the shipped chemlab tables.  Only that window of each
//     type pair's table lives in shared memory, hottest pairs (by type population) first, until the budget is used; rows outside
//     a window and tables without a window are read from global memory (L2) on a rare, divergent path.  28 tables x 1750 rows
//     (784 KB) do not fit an SM; their hot windows (74 % of the pair work in 70 KB for rim135) do.
//   * REPLICATION (RLOG = 3): 8 interleaved copies of the windows, lane l reads copy l & 7, so the 8 lanes of a quarter warp
//     hit 8 different 16-byte bank groups: the random row gather becomes conflict-free (single-table systems, where it fits).
//   * the row gather is PREDICATED on the cutoff test: lanes outside the cutoff do not take part in the gather (fewer bank
//     conflicts), instead of gathering a clamped row and discarding the result.
//   * tables may differ in length (common x0 and dx only): the range check uses the pair's own row count on the global path.
struct ClbPairDesc3 {          // 16 bytes per type pair
    double rc2;                // cutoff^2 in lattice^2; < 0: no potential
    int soff;                  // shared-memory row of this pair's table row 0, (window start row - w0) << RLOG; unused when wn == 0
    unsigned short w0, wn;     // first row and number of rows of the shared-memory window (wn == 0: global memory only)
};
struct ClbPairArgs3 {
    const int* cell_start; const int4* pos; const unsigned short* entries; const int* nl_count;
    const unsigned short* perm;   // work order of the home particles of every block (block_perm); NULL: cell order
    const ClbPairDesc3* pd3;   // [ntypes^2]
    const int2* gmeta;         // [ntypes^2] {first row of the pair's table in trows - shift, (shift << 20) | (rows - 1)}; shift = rows between
                               // the common grid origin and the table's first abscissa (tables may start at different r)
    const double2* trows;      // {A_i, B_i} rows of all tables, lattice units (global memory)
    const double2* swin;       // shared-memory image of the windows: nsrows << RLOG rows, copied in at kernel start
    double* force; ClbCtl* ctl;
    int cap, ntypes, nsrows, fstride, npw;
    double invdx, cmagic;
    ClbPairDesc3 one; int2 one_g;  // ONEPD: the only descriptor
    int b0, seg0, b1, nidx;
    int nv, vc_bytes;
    int pipe, tile_cap;            // pipe: two tile buffers per virtual CTA, the next tile is staged while the current one is evaluated
};
// Rare path of k_pair_forces_tab3: the listed pairs of one 8-entry batch whose table row is NOT in a shared-memory window
// (bit k of `defer`) are evaluated again from scratch with the row read from global memory.  Kept out of line so that its
// registers do not burden the hot loop.
__device__ __noinline__ void pair_tab3_slow(unsigned defer, uint4 cur, unsigned pix, unsigned piy, unsigned piz, const int4* s_pos,
                                            const int2* __restrict__ gmeta_row, int2 one_g, bool onepd, const double2* __restrict__ trows,
                                            double invdx, double cmagic, double* acc, unsigned* err) {
    const unsigned wds[4] = {cur.x, cur.y, cur.z, cur.w};
    while (defer) {
        const int k = __ffs(defer) - 1;
        defer &= defer - 1;
        const unsigned w = wds[k >> 1];
        const unsigned e = (k & 1) ? (w >> 16) : (w & 0xffffu);
        const int4 pj = s_pos[e];
        const double dx = __hiloint2double(0x43300000, (int)(pix - (unsigned)pj.x)) - 4503601774854144.0;
        const double dy = __hiloint2double(0x43300000, (int)(piy - (unsigned)pj.y)) - 4503601774854144.0;
        const double dz = __hiloint2double(0x43300000, (int)(piz - (unsigned)pj.z)) - 4503601774854144.0;
        const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
        double y0, yh;
        rsqrt_seed2(r2, y0, yh);
        const double h = r2 * y0, ee = fma(-h, yh, 0.5);
        const double y = fma(y0, ee, y0), r = fma(h, ee, h);
        const unsigned ix = (unsigned)__double2loint(__fma_rd(r, invdx, cmagic));
        const int2 gm = onepd ? one_g : __ldg(gmeta_row + pw_type(pj.w));
        const unsigned sh = (unsigned)gm.y >> 20, nm1 = (unsigned)gm.y & 0xfffffu;      // this table covers common rows [sh, sh + nm1]
        if (ix - sh > nm1) *err |= CLB_EF_TABLE_RANGE;                // fatal in the reference (U12)
        const double2 rw = __ldg(trows + gm.x + (int)min(max(ix, sh), sh + nm1));
        const double F = fma(r, rw.y, rw.x) * y;
        acc[0] = fma(F, dx, acc[0]); acc[1] = fma(F, dy, acc[1]); acc[2] = fma(F, dz, acc[2]);
    }
}
// INLINE_FB: type pairs whose table has no shared-memory window are common (many-table systems: rim135 keeps 13 of 28 windows
// resident = 90 % of the pair work) -> their rows are gathered from global memory by a predicated load inside the hot loop
// instead of the out-of-line rare path (which re-evaluates the pair and serialises a warp whenever ANY lane needs it).
template <bool ONEPD, int RLOG, int NI, bool INLINE_FB>
__global__ void __launch_bounds__(1024) k_pair_forces_tab3(ClbGrid g, ClbPairArgs3 A) {
    if (*(volatile int*)&A.ctl->stall) return;
    extern __shared__ __align__(16) unsigned char smem[];
    const int ntp = A.ntypes * A.ntypes;
    ClbPairDesc3* s_pd = reinterpret_cast<ClbPairDesc3*>(smem);
    double2* s_rows = reinterpret_cast<double2*>(s_pd + (ONEPD ? 0 : ntp));
    unsigned char* vc_base = reinterpret_cast<unsigned char*>(s_rows + ((size_t)A.nsrows << RLOG));
    const int nth = A.npw * 32;
    const int vc = threadIdx.x / nth, tid = threadIdx.x - vc * nth;
    const int bar = 1 + vc;
    // per virtual CTA: [full mbarriers 2 x 8 B][empty mbarriers 2 x 8 B][TileMeta x 2][tile buffer 0][tile buffer 1 (pipe)]
    unsigned long long* mb_full = reinterpret_cast<unsigned long long*>(vc_base + (size_t)vc * A.vc_bytes);
    unsigned long long* mb_empty = mb_full + 2;
    TileMeta* metas = reinterpret_cast<TileMeta*>(mb_empty + 2);
    int4* s_pos0 = reinterpret_cast<int4*>(metas + 2);
    const bool pipe = A.pipe != 0;
    if (!ONEPD) for (int i = threadIdx.x; i < ntp; i += blockDim.x) s_pd[i] = A.pd3[i];
    for (int i = threadIdx.x; i < (A.nsrows << RLOG); i += blockDim.x) s_rows[i] = __ldg(A.swin + i);
    if (tid == 0) {
        mbar_init(mb_full, 1); mbar_init(mb_full + 1, 1); mbar_init(mb_empty, (unsigned)A.npw); mbar_init(mb_empty + 1, (unsigned)A.npw);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5;
    // neighbour-row batches (8 entries = 16 B per lane) are prefetched ONE BATCH AHEAD with cp.async into two private 16-byte
    // slots per thread.  (Round 1 / first round-2 version prefetched into registers: ptxas sank the load to the end of the loop
    // body, right in front of its use, and the warps spent 20 % of the kernel waiting for DRAM at that one instruction --
    // profiles/r2c_ncu_pair_sass.csv.)
    uint4* s_ev = reinterpret_cast<uint4*>(vc_base + (size_t)A.nv * A.vc_bytes) + threadIdx.x;
    const unsigned s_ev_u32 = smem_u32(s_ev);
    const unsigned ev_stride = blockDim.x * 16u;
    const double2* s_rows_rep = s_rows + (lane & ((1 << RLOG) - 1));     // this lane's copy of the replicated windows
    const int nhpass = nth;
    const double invdx = A.invdx, cmagic = A.cmagic;
    unsigned err = 0;
    // PIPE (round 2): the profile of the single-buffer version showed 21 % of the warp time at the group barrier between two
    // tiles (waiting for the slowest warp of the group, then for the bulk copies of the next tile).  With two buffers the warps
    // of a group are decoupled: warp 0 stages tile k+1 as soon as it starts tile k (after the `empty` mbarrier of that buffer
    // has collected one arrival per warp for tile k-1), every warp waits only for the `full` mbarrier of its own next tile,
    // and warp 0 takes the shortest rows of the block (block_perm) so that it is the first to move on.
    const int first = vc * gridDim.x + blockIdx.x, stride = gridDim.x * A.nv;
    const int wv = pipe ? (A.npw - 1 - warp) : warp;            // chunk of the (length-sorted) home particles this warp takes
    if (pipe && warp == 0 && first < A.nidx) {
        TileCtx t0; tile_geometry(g, first < A.seg0 ? A.b0 + first : A.b1 + (first - A.seg0), t0);
        tile_stage_bulk(g, t0, A.cell_start, A.pos, s_pos0, metas, mb_full, lane);
    }
    int k = 0;
    for (int idx = first; idx < A.nidx; idx += stride, ++k) {
        const int buf = pipe ? (k & 1) : 0;
        const int4* s_pos = s_pos0 + (size_t)buf * A.tile_cap;
        TileCtx t;
        if (!pipe) {
            const int b = idx < A.seg0 ? A.b0 + idx : A.b1 + (idx - A.seg0);
            tile_geometry(g, b, t);
            vc_sync(bar, nth);                                     // every warp is done with the previous tile
            if (warp == 0) tile_stage_bulk(g, t, A.cell_start, A.pos, s_pos0, metas, mb_full, lane);
            vc_sync(bar, nth);                                     // meta visible
            t.hs = metas->hs; t.nh = metas->nh;
            mbar_wait(mb_full, (unsigned)(k & 1));                 // tile landed
        } else {
            if (warp == 0 && idx + stride < A.nidx) {
                const int nx = idx + stride;
                if (k >= 1) mbar_wait(mb_empty + (buf ^ 1), (unsigned)(((k - 1) >> 1) & 1));       // tile k-1 has left that buffer
                TileCtx tn; tile_geometry(g, nx < A.seg0 ? A.b0 + nx : A.b1 + (nx - A.seg0), tn);
                tile_stage_bulk(g, tn, A.cell_start, A.pos, s_pos0 + (size_t)(buf ^ 1) * A.tile_cap, metas + (buf ^ 1), mb_full + (buf ^ 1), lane);
            }
            mbar_wait(mb_full + buf, (unsigned)((k >> 1) & 1));
            t.hs = metas[buf].hs; t.nh = metas[buf].nh;
        }
        for (int p0 = 0; p0 < t.nh; p0 += nhpass) {
            const int p = p0 + wv * 32 + lane;
            const bool act = p < t.nh;
            const int gi = t.hs + (act ? (A.perm ? (int)__ldg(A.perm + t.hs + p) : p) : 0);
            const int4 pi = __ldg(A.pos + gi);
            const unsigned pix = (unsigned)pi.x + 0x80000000u, piy = (unsigned)pi.y + 0x80000000u, piz = (unsigned)pi.z + 0x80000000u;
            const int cnt = act ? __ldg(A.nl_count + gi) : 0;
            const int trow = pw_type(pi.w) * A.ntypes;
            const uint4* row = reinterpret_cast<const uint4*>(A.entries + (size_t)gi * A.cap);
            const int nb = (cnt + 7) >> 3;
            double ax = 0.0, ay = 0.0, az = 0.0;
            if (nb > 0) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s_ev_u32), "l"(row) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
#pragma unroll 1
            for (int bi = 0; bi < nb; ++bi) {
                if (bi + 1 < nb) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s_ev_u32 + ((bi + 1) & 1) * ev_stride), "l"(row + bi + 1) : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 1;" ::: "memory");          // batch bi has landed (only the newest group may be pending)
                uint4 cur;
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(cur.x), "=r"(cur.y), "=r"(cur.z), "=r"(cur.w) : "r"(s_ev_u32 + (bi & 1) * ev_stride) : "memory");
                const int ne = cnt - bi * 8;
                const unsigned wds[4] = {cur.x, cur.y, cur.z, cur.w};
                unsigned defer = 0;
#pragma unroll
                for (int s0 = 0; s0 < 8; s0 += NI) {
                    bool in[NI];
                    int4 pj[NI];
                    double dx[NI], dy[NI], dz[NI], r2[NI], y0[NI], yh[NI], y[NI], r[NI], F[NI];
                    ClbPairDesc3 d[NI];
                    int tpj[NI];
                    double2 rw[NI];
#pragma unroll
                    for (int q = 0; q < NI; ++q) {
                        const int k = s0 + q;
                        unsigned e = (k & 1) ? (wds[k >> 1] >> 16) : (wds[k >> 1] & 0xffffu);
                        in[q] = k < ne;
                        e = in[q] ? e : 0u;
                        pj[q] = s_pos[e];
                    }
#pragma unroll
                    for (int q = 0; q < NI; ++q) {
                        dx[q] = __hiloint2double(0x43300000, (int)(pix - (unsigned)pj[q].x)) - 4503601774854144.0;
                        dy[q] = __hiloint2double(0x43300000, (int)(piy - (unsigned)pj[q].y)) - 4503601774854144.0;
                        dz[q] = __hiloint2double(0x43300000, (int)(piz - (unsigned)pj[q].z)) - 4503601774854144.0;
                        if (ONEPD) d[q] = A.one;
                        else { d[q] = s_pd[trow + pw_type(pj[q].w)]; if (INLINE_FB) tpj[q] = pw_type(pj[q].w); }
                    }
#pragma unroll
                    for (int q = 0; q < NI; ++q) r2[q] = fma(dz[q], dz[q], fma(dy[q], dy[q], dx[q] * dx[q]));
#pragma unroll
                    for (int q = 0; q < NI; ++q) { in[q] = in[q] && (r2[q] <= d[q].rc2); rsqrt_seed2(r2[q], y0[q], yh[q]); }
#pragma unroll
                    for (int q = 0; q < NI; ++q) {
                        const double h = r2[q] * y0[q];
                        const double ee = fma(-h, yh[q], 0.5);
                        y[q] = fma(y0[q], ee, y0[q]);             // 1/r
                        r[q] = fma(h, ee, h);                     // r
                    }
#pragma unroll
                    for (int q = 0; q < NI; ++q) {
                        const double ti = __fma_rd(r[q], invdx, cmagic);    // low word = floor((r - x0)/dx) exactly
                        const unsigned ix = (unsigned)__double2loint(ti);
                        const bool inw = in[q] && (ix - (unsigned)d[q].w0) < (unsigned)d[q].wn;
                        rw[q] = make_double2(0.0, 0.0);
                        if (inw) rw[q] = s_rows_rep[d[q].soff + (int)(ix << RLOG)];     // predicated gather: lanes beyond the cutoff stay out
                        else if (in[q]) {
                            if (INLINE_FB) {                                                  // cold type pair / wall row: global memory (L2), same arithmetic
                                const int2 gm = ONEPD ? A.one_g : __ldg(A.gmeta + trow + tpj[q]);
                                const unsigned sh = (unsigned)gm.y >> 20, nm1 = (unsigned)gm.y & 0xfffffu;
                                if (ix - sh > nm1) err |= CLB_EF_TABLE_RANGE;
                                rw[q] = __ldg(A.trows + gm.x + (int)min(max(ix, sh), sh + nm1));
                            } else defer |= 1u << (s0 + q);                                   // rare: out-of-line path below
                        }
                    }
#pragma unroll
                    for (int q = 0; q < NI; ++q) {
                        F[q] = fma(r[q], rw[q].y, rw[q].x) * y[q];
                        F[q] = in[q] ? F[q] : 0.0;
                    }
#pragma unroll
                    for (int q = 0; q < NI; ++q) { ax = fma(F[q], dx[q], ax); ay = fma(F[q], dy[q], ay); az = fma(F[q], dz[q], az); }
                }
                if (!INLINE_FB && defer) {
                    double acc[3] = {ax, ay, az};
                    pair_tab3_slow(defer, cur, pix, piy, piz, s_pos, A.gmeta + trow, A.one_g, ONEPD, A.trows, invdx, cmagic, acc, &err);
                    ax = acc[0]; ay = acc[1]; az = acc[2];
                }
            }
            if (act) { A.force[gi] = ax; A.force[gi + A.fstride] = ay; A.force[gi + 2 * A.fstride] = az; }
        }
        if (pipe) { __syncwarp(); if (lane == 0) mbar_arrive(mb_empty + buf); }      // this warp has left the buffer
    }
    if (err) atomicOr(&A.ctl->err, err);
}

// ------------------------------------------------------------------------------------------
// Pair energy of one interaction handle (analysis.PotentialEnergy): fp64 throughout, each pair
// visited twice (full list) -> factor 1/2.  Per-block partial sums are reduced in a fixed order
// by k_sum_partials, so the result is bit-reproducible.
template <bool CUBIC>
__global__ void __launch_bounds__(512) k_pair_energy(ClbGrid g, ClbGeom geo, const int* __restrict__ cell_start,
                                                     const int4* __restrict__ pos,
                                                     const unsigned short* __restrict__ entries,
                                                     const int* __restrict__ nl_count, int cap,
                                                     const ClbPairDesc* __restrict__ pdesc,
                                                     const ClbPairDescE* __restrict__ pdesce, int ntypes,
                                                     const ClbTabMeta* __restrict__ tmeta,
                                                     const double2* __restrict__ erows, int inter,
                                                     double* __restrict__ partial, unsigned long long* __restrict__ pcount) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_off[CLB_TILE_CELLS + 1];
    __shared__ int s_src[CLB_TILE_CELLS];
    __shared__ double s_red[32];
    __shared__ unsigned long long s_cnt[32];
    int4* s_pos = reinterpret_cast<int4*>(smem);
    double esum = 0.0;
    unsigned long long ninter = 0;
    for (int b = blockIdx.x; b < g.nblocks; b += gridDim.x) {
        TileCtx t;
        tile_geometry(g, b, t);
        __syncthreads();
        tile_offsets(g, t, cell_start, s_off, s_src);
        tile_stage(t, s_off, s_src, pos, s_pos, nullptr, nullptr, nullptr);
        __syncthreads();
        for (int p = threadIdx.x; p < t.nh; p += blockDim.x) {
            const int gi = t.hs + p;
            const int4 pi = __ldg(pos + gi);
            const int cnt = __ldg(nl_count + gi);
            const unsigned short* ent = entries + (size_t)gi * cap;
            for (int k = 0; k < cnt; ++k) {
                const int4 pj = s_pos[ent[k]];
                double dx = lat2d(wsub(pi.x, pj.x)) * geo.q[0], dy = lat2d(wsub(pi.y, pj.y)) * geo.q[1], dz = lat2d(wsub(pi.z, pj.z)) * geo.q[2];
                double r2 = dx * dx + dy * dy + dz * dz;
                int tp = pw_type(pi.w) * ntypes + pw_type(pj.w);
                const ClbPairDesc pd = pdesc[tp];
                const ClbPairDescE pe = pdesce[tp];
                if (pd.kind == 0 || r2 > pd.rc2) continue;   // kind 0 <=> rc2 < 0
                ++ninter;
                if (pe.inter != inter) continue;
                if (pd.kind == 1) {
                    const ClbTabMeta tm = tmeta[pd.tab];
                    double r = sqrt(r2);
                    double s = (r - tm.x0) * tm.invdx;
                    int idx = (int)floor(s);
                    idx = max(0, min(idx, tm.n - 2));
                    double bfrac = s - (double)idx;
                    double2 row = __ldg(erows + tm.off + idx);
                    esum += fma(bfrac, row.y, row.x);
                } else {
                    double f2 = 1.0 / r2, f6 = f2 * f2 * f2;
                    esum += (pe.e12 * f6 - pe.e6) * f6 - pe.shift;
                }
            }
        }
    }
    // block reduction in fixed order
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { esum += __shfl_down_sync(0xffffffffu, esum, d); ninter += __shfl_down_sync(0xffffffffu, ninter, d); }
    if (lane == 0) { s_red[warp] = esum; s_cnt[warp] = ninter; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0; unsigned long long c = 0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) { a += s_red[w]; c += s_cnt[w]; }
        partial[blockIdx.x] = 0.5 * a; pcount[blockIdx.x] = c;
    }
}

// ------------------------------------------------------------------------------------------
// Decode the tile lists into (slot_i, slot_j) rows with slot_i < slot_j: the Verlet pair SET
// for parity (clb_get_pairs).  Order of rows is arbitrary; the host sorts.
__global__ void __launch_bounds__(512) k_decode_pairs(ClbGrid g, const int* __restrict__ cell_start,
                                                      const int* __restrict__ slot,
                                                      const unsigned short* __restrict__ entries,
                                                      const int* __restrict__ nl_count, int cap, int2* __restrict__ out,
                                                      unsigned long long outcap, ClbCtl* ctl) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_off[CLB_TILE_CELLS + 1];
    __shared__ int s_src[CLB_TILE_CELLS];
    int* s_slot = reinterpret_cast<int*>(smem);
    for (int b = blockIdx.x; b < g.nblocks; b += gridDim.x) {
        TileCtx t;
        tile_geometry(g, b, t);
        __syncthreads();
        tile_offsets(g, t, cell_start, s_off, s_src);
        const int nct = CLB_TILE_ROWS * t.W;
        for (int tc = threadIdx.x >> 5; tc < nct; tc += blockDim.x >> 5)
            for (int i = threadIdx.x & 31; i < s_off[tc + 1] - s_off[tc]; i += 32) s_slot[s_off[tc] + i] = __ldg(slot + s_src[tc] + i);
        __syncthreads();
        for (int p = threadIdx.x; p < t.nh; p += blockDim.x) {
            const int gi = t.hs + p;
            const int si = __ldg(slot + gi);
            const int cnt = __ldg(nl_count + gi);
            const unsigned short* ent = entries + (size_t)gi * cap;
            for (int k = 0; k < cnt; ++k) {
                int sj = s_slot[ent[k]];
                if (si < sj) {
                    unsigned long long o = atomicAdd(&ctl->npairs_out, 1ull);
                    if (o < outcap) out[o] = make_int2(si, sj);
                }
            }
        }
    }
}
