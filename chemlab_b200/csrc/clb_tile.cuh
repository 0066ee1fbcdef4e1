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
    int row = b / g.nbx, bxi = b - row * g.nbx;
    t.cx0 = bxi * g.bx;
    t.bxe = min(g.bx, g.ncx - t.cx0);
    t.cy = row % g.ncy;
    int zrow = row / g.ncy;                 // 0..nczl-1 : owned plane index
    t.lz = g.ghost ? zrow + 1 : zrow;
    t.whole = (t.bxe + 2 > g.ncx);
    t.W = t.whole ? g.ncx : t.bxe + 2;
}
// global (local-grid) cell index of tile cell (row k, column m)
__device__ __forceinline__ int tile_cell(const ClbGrid& g, const TileCtx& t, int k, int m) {
    int dy = k % 3 - 1, dz = k / 3 - 1;
    int cy = wrapi(t.cy + dy, g.ncy);
    int lz = g.ghost ? t.lz + dz : wrapi(t.lz + dz, g.ncz);
    int cx = t.whole ? m : wrapi(t.cx0 - 1 + m, g.ncx);
    return (lz * g.ncy + cy) * g.ncx + cx;
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
// tile column of the home cell that holds home particle p (local index in [0,nh))
__device__ __forceinline__ int home_column(const TileCtx& t, const int* s_off, int p) {
    int mh0 = t.whole ? t.cx0 : 1;
    int base = s_off[4 * t.W + mh0];
    int m = mh0;
    while (m < mh0 + t.bxe - 1 && s_off[4 * t.W + m + 1] - base <= p) ++m;
    return m;
}

// ------------------------------------------------------------------------------------------
// Neighbour-list build.  One thread per home particle; candidates are the 27 cells around the
// particle's own cell, read from the shared tile (all lanes of a cell read the same candidate:
// broadcast).  Inclusion test (U1): r^2 <= (rc+skin)^2 in fp64 on exact lattice differences,
// after an integer box prefilter.  Exclusions: per-particle sorted partner-slot rows.
// Entry layout: entries[hs*cap + k*nh + p] (coalesced over p).
template <bool CUBIC>
__global__ void __launch_bounds__(512) k_build_lists(ClbGrid g, ClbGeom geo, const int* __restrict__ cell_start,
                                                     const int4* __restrict__ pos, const int* __restrict__ slot,
                                                     const int* __restrict__ excl_off, const int* __restrict__ excl_ids,
                                                     unsigned short* __restrict__ entries, int* __restrict__ nl_count,
                                                     int cap, ClbCtl* ctl) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_off[CLB_TILE_CELLS + 1];
    __shared__ int s_src[CLB_TILE_CELLS];
    __shared__ int s_stat[2];
    for (int b = blockIdx.x; b < g.nblocks; b += gridDim.x) {
        TileCtx t;
        tile_geometry(g, b, t);
        __syncthreads();
        tile_offsets(g, t, cell_start, s_off, s_src);
        int4* s_pos = reinterpret_cast<int4*>(smem);
        int* s_slot = reinterpret_cast<int*>(s_pos + t.T);
        tile_stage(t, s_off, s_src, pos, s_pos, slot, s_slot, nullptr);
        if (threadIdx.x == 0) { s_stat[0] = 0; s_stat[1] = 0; }
        __syncthreads();
        int lmax = 0, ltot = 0;
        for (int p = threadIdx.x; p < t.nh; p += blockDim.x) {
            const int gi = t.hs + p;
            const int mh = home_column(t, s_off, p);
            const int ti = s_off[4 * t.W + (t.whole ? t.cx0 : 1)] + p;   // home cells are consecutive tile cells
            const int4 pi = s_pos[ti];
            const int myslot = s_slot[ti];
            const int e0 = __ldg(excl_off + myslot), e1 = __ldg(excl_off + myslot + 1);
            int cnt = 0;
            unsigned short* out = entries + (size_t)t.hs * cap + p;
            for (int k = 0; k < CLB_TILE_ROWS; ++k) {
#pragma unroll
                for (int dm = -1; dm <= 1; ++dm) {
                    int m = t.whole ? wrapi(mh + dm, g.ncx) : mh + dm;
                    int tc = k * t.W + m;
                    int lo = s_off[tc], hi = s_off[tc + 1];
                    for (int j = lo; j < hi; ++j) {
                        int4 pj = s_pos[j];
                        int dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
                        if (abs(dx) > geo.cut[0] || abs(dy) > geo.cut[1] || abs(dz) > geo.cut[2]) continue;
                        double fx = lat2d(dx), fy = lat2d(dy), fz = lat2d(dz), r2;
                        if (CUBIC) r2 = (fx * fx + fy * fy + fz * fz) * geo.q2;
                        else { fx *= geo.q[0]; fy *= geo.q[1]; fz *= geo.q[2]; r2 = fx * fx + fy * fy + fz * fz; }
                        if (r2 > geo.rl2 || j == ti) continue;
                        int sj = s_slot[j];
                        bool ex = false;
                        for (int e = e0; e < e1; ++e) ex |= (__ldg(excl_ids + e) == sj);
                        if (ex) continue;
                        if (cnt < cap) out[(size_t)cnt * t.nh] = (unsigned short)j;
                        ++cnt;
                    }
                }
            }
            nl_count[gi] = min(cnt, cap);
            if (cnt > cap) atomicOr(&ctl->err, CLB_EF_LIST_OVERFLOW);
            lmax = max(lmax, cnt);
            ltot += min(cnt, cap);
        }
        // statistics (block reduce through shared atomics: order-independent integers)
        atomicMax(&s_stat[0], lmax);
        atomicAdd(&s_stat[1], ltot);
        __syncthreads();
        if (threadIdx.x == 0) { atomicMax(&ctl->nl_max, s_stat[0]); atomicAdd(&ctl->nl_total, (unsigned long long)s_stat[1]); }
    }
}

// ------------------------------------------------------------------------------------------
// Pair forces.  One thread per home particle, full (two-sided) list: no scatter, no atomics,
// deterministic.  Everything after the integer subtraction is fp64 (B200 has a full-rate fp64
// pipe that issues beside the fp32/int pipes); no float<->double conversion instructions:
//   d       exact lattice difference -> double by the 2^52 trick (lat2d)
//   1/r     MUFU.RSQ seed on the truncated high word + one fp64 Newton step
//   index   t = (r-x0)/dx ; floor and fraction by the 2^52 rounding trick
//   F(r)    f[i] + b*(f[i+1]-f[i])  (reference: linear interpolation itype=1)
// Table rows {f_i + df_i/2, df_i = f_{i+1}-f_i} as double2 live in shared memory (persistent CTAs load them once).
__device__ __forceinline__ double rsqrt_seed(double x) {
    // float with the same value as x truncated to 24 bits, via integer ops only
    int hi = __double2hiint(x), lo = __double2loint(x);
    unsigned fb = ((unsigned)(hi - 0x38000000) << 3) | ((unsigned)lo >> 29);
    float y = rsqrtf(__uint_as_float(fb));
    unsigned yb = __float_as_uint(y);
    return __hiloint2double((int)((yb >> 3) + 0x38000000u), (int)(yb << 29));
}

template <bool CUBIC, bool TABS_SMEM>
__global__ void __launch_bounds__(512) k_pair_forces(ClbGrid g, ClbGeom geo, const int* __restrict__ cell_start,
                                                     const int4* __restrict__ pos,
                                                     const unsigned short* __restrict__ entries,
                                                     const int* __restrict__ nl_count, int cap,
                                                     const ClbPairDesc* __restrict__ pdesc, int ntypes,
                                                     const ClbTabMeta* __restrict__ tmeta, int ntabs,
                                                     const double2* __restrict__ trows, int nrows_total,
                                                     double* __restrict__ force, int fstride, ClbCtl* ctl) {
    if (*(volatile int*)&ctl->stall) return;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_off[CLB_TILE_CELLS + 1];
    __shared__ int s_src[CLB_TILE_CELLS];
    // static part of dynamic smem: pair descriptors, table meta, (tables), then the tile
    ClbPairDesc* s_pd = reinterpret_cast<ClbPairDesc*>(smem);
    ClbTabMeta* s_tm = reinterpret_cast<ClbTabMeta*>(s_pd + ntypes * ntypes);
    double2* s_rows = reinterpret_cast<double2*>(s_tm + ntabs);
    int4* s_pos = reinterpret_cast<int4*>(s_rows + (TABS_SMEM ? nrows_total : 0));
    for (int i = threadIdx.x; i < ntypes * ntypes; i += blockDim.x) s_pd[i] = pdesc[i];
    for (int i = threadIdx.x; i < ntabs; i += blockDim.x) s_tm[i] = tmeta[i];
    if (TABS_SMEM) for (int i = threadIdx.x; i < nrows_total; i += blockDim.x) s_rows[i] = __ldg(trows + i);
    const double2* rows = TABS_SMEM ? s_rows : trows;
    unsigned err = 0;
    for (int b = blockIdx.x; b < g.nblocks; b += gridDim.x) {
        TileCtx t;
        tile_geometry(g, b, t);
        __syncthreads();
        tile_offsets(g, t, cell_start, s_off, s_src);
        tile_stage(t, s_off, s_src, pos, s_pos, nullptr, nullptr, nullptr);
        __syncthreads();
        for (int p0 = 0; p0 < t.nh; p0 += blockDim.x) {
            const int p = p0 + threadIdx.x;
            const bool act = p < t.nh;
            const int gi = t.hs + (act ? p : 0);
            const int4 pi = __ldg(pos + gi);
            const int cnt = act ? __ldg(nl_count + gi) : 0;
            const ClbPairDesc* pdrow = s_pd + pw_type(pi.w) * ntypes;
            const unsigned short* ent = entries + (size_t)t.hs * cap + p;
            double ax = 0.0, ay = 0.0, az = 0.0;
            int kmax = cnt;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, d));
            unsigned e_next = (0 < cnt) ? ent[0] : 0;
            for (int k = 0; k < kmax; ++k) {
                unsigned e = e_next;
                if (k + 1 < cnt) e_next = ent[(size_t)(k + 1) * t.nh];
                if (k < cnt) {
                    const int4 pj = s_pos[e];
                    double dx = lat2d(pi.x - pj.x), dy = lat2d(pi.y - pj.y), dz = lat2d(pi.z - pj.z), r2;
                    if (CUBIC) r2 = (dx * dx + dy * dy + dz * dz) * geo.q2;
                    else { dx *= geo.q[0]; dy *= geo.q[1]; dz *= geo.q[2]; r2 = dx * dx + dy * dy + dz * dz; }
                    const ClbPairDesc pd = pdrow[pw_type(pj.w)];
                    if (pd.kind != 0 && r2 <= pd.rc2) {
                        double y = rsqrt_seed(r2);
                        double h = r2 * y;
                        double ee = fma(-h, y, 1.0);
                        y = fma(0.5 * y, ee, y);              // 1/r to ~1e-14
                        double fr;
                        if (pd.kind == 1) {
                            const ClbTabMeta tm = s_tm[pd.tab];
                            double r = r2 * y;
                            // u = (r-x0)/dx - 1/2 ; idx = round(u) = floor((r-x0)/dx) ; b' = u - idx in [-1/2,1/2]
                            double u = fma(r, tm.invdx, tm.c_t);
                            double ti = u + 6755399441055744.0;            // 1.5 * 2^52: integer in the low word
                            int idx = __double2loint(ti);
                            double bfrac = u - (ti - 6755399441055744.0);
                            if (idx < 0 || idx > tm.n - 2) {
                                if (idx < 0 || r > tm.x0 + tm.dx * (tm.n - 1) * (1.0 + 1e-12)) err |= CLB_EF_TABLE_RANGE;
                                int ic = idx < 0 ? 0 : tm.n - 2;
                                bfrac += (double)(idx - ic); idx = ic;
                            }
                            double2 row = rows[tm.off + idx];                // {f_i + df_i/2, df_i}
                            fr = fma(bfrac, row.y, row.x) * y;
                        } else {
                            double y2 = y * y;
                            double y6 = y2 * y2 * y2;
                            fr = y6 * fma(pd.c12, y6, -pd.c6) * y2;
                        }
                        ax = fma(fr, dx, ax); ay = fma(fr, dy, ay); az = fma(fr, dz, az);
                    }
                }
            }
            if (act) {
                if (CUBIC) { ax *= geo.q[0]; ay *= geo.q[0]; az *= geo.q[0]; }
                force[gi] = ax; force[gi + fstride] = ay; force[gi + 2 * fstride] = az;
            }
        }
    }
    if (err) atomicOr(&ctl->err, err);
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
            const unsigned short* ent = entries + (size_t)t.hs * cap + p;
            for (int k = 0; k < cnt; ++k) {
                const int4 pj = s_pos[ent[(size_t)k * t.nh]];
                double dx = lat2d(pi.x - pj.x) * geo.q[0], dy = lat2d(pi.y - pj.y) * geo.q[1], dz = lat2d(pi.z - pj.z) * geo.q[2];
                double r2 = dx * dx + dy * dy + dz * dz;
                int tp = pw_type(pi.w) * ntypes + pw_type(pj.w);
                const ClbPairDesc pd = pdesc[tp];
                const ClbPairDescE pe = pdesce[tp];
                if (pd.kind == 0 || r2 > pd.rc2) continue;
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
            const unsigned short* ent = entries + (size_t)t.hs * cap + p;
            for (int k = 0; k < cnt; ++k) {
                int sj = s_slot[ent[(size_t)k * t.nh]];
                if (si < sj) {
                    unsigned long long o = atomicAdd(&ctl->npairs_out, 1ull);
                    if (o < outcap) out[o] = make_int2(si, sj);
                }
            }
        }
    }
}
