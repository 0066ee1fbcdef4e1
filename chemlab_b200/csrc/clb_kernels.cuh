// clb_kernels.cuh -- O(N) kernels: cell keys / reorder, Velocity-Verlet + Langevin, resort check,
// bonded forces (owner-computes CSR), observables.
#pragma once
#include "clb_common.cuh"

// ---------------------------------------------------------------- rebuild helpers -----------
// cell key of every stored particle (storage.decompose(): src/start_simulation.py:171,205,295)
__global__ void k_cell_keys(int n, const int4* __restrict__ pos, ClbGrid g, int* __restrict__ key, int* __restrict__ val) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int4 p = pos[i];
    int cx = __umulhi((unsigned)p.x, (unsigned)g.ncx);
    int cy = __umulhi((unsigned)p.y, (unsigned)g.ncy);
    int cz = __umulhi((unsigned)p.z, (unsigned)g.ncz);
    int lz = local_plane(g, cz);
    // a particle outside the owned+ghost planes cannot occur between rebuilds (displacement < skin/2); park it in
    // the last cell so that the sort stays well defined
    key[i] = lz < 0 ? g.ncell - 1 : (lz * g.ncy + cy) * g.ncx + cx;
    val[i] = i;
}
__global__ void k_gather(int n, const int* __restrict__ perm, const int4* __restrict__ pos_in, const ClbVel* __restrict__ vel_in,
                         const int* __restrict__ slot_in, int4* __restrict__ pos, ClbVel* __restrict__ vel,
                         int* __restrict__ slot, int4* __restrict__ xref, int* __restrict__ id2idx) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int s = perm[i];
    int4 p = pos_in[s];
    pos[i] = p; xref[i] = p;
    vel[i] = vel_in[s];
    int sl = slot_in[s];
    slot[i] = sl;
    id2idx[sl] = i;
}
__global__ void k_cell_start(int n, const int* __restrict__ key, int ncell, int* __restrict__ cell_start) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (n == 0) { if (i == 0) for (int c = 0; c <= ncell; ++c) cell_start[c] = 0; return; }
    int prev = i == 0 ? -1 : key[i - 1];
    int cur = i == n ? ncell : key[i];
    for (int c = prev + 1; c <= cur; ++c) cell_start[c] = i;
}
// ---- row-block table ---------------------------------------------------------------------------------------------------
// A row block = consecutive cells of one x-row whose particles are the HOME particles of one tile (clb_tile.cuh).  Blocks are
// cut greedily along each row: a block closes when the next cell would push it over `target` home particles (so that the
// home particles fill the lanes of the warps that work on the tile: fixed-length blocks of a fluctuating melt left a quarter
// of the lanes idle) or when it holds `bx` cells (shared-memory offset tables).  Empty blocks are dropped.  Three tiny
// kernels: blocks per row, exclusive scan over the rows (one CTA), table fill.
template <bool WRITE>
__global__ void k_blocks_rows(ClbGrid g, const int* __restrict__ cell_start, int* __restrict__ row_n, const int* __restrict__ row_off, int4* __restrict__ blk) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= g.ncy * g.nczl) return;
    const int cy = row % g.ncy, lz = row / g.ncy;
    const int* cs = cell_start + (size_t)row * g.ncx;
    const int target = g.target > 0 ? g.target : 0x7fffffff;
    int k = 0, c = 0;
    int o = WRITE ? row_off[row] : 0;
    while (c < g.ncx) {
        int c1 = c + 1, cnt = cs[c1] - cs[c];
        while (c1 < g.ncx && c1 - c < g.bx) { const int nx = cs[c1 + 1] - cs[c1]; if (cnt + nx > target) break; cnt += nx; ++c1; }
        if (cnt > 0) { if (WRITE) blk[o + k] = make_int4(c, c1 - c, cy, lz); ++k; }
        c = c1;
    }
    if (!WRITE) row_n[row] = k;
}
__global__ void __launch_bounds__(1024) k_blocks_scan(int nrows, int ncy, int nczl, const int* __restrict__ row_n, int* __restrict__ row_off, ClbCtl* ctl) {
    __shared__ int s_w[32];
    __shared__ int s_run;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (int base = 0; base < nrows; base += 1024) {
        const int r = base + threadIdx.x;
        const int v = r < nrows ? row_n[r] : 0;
        int incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += u; }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = s_w[lane], wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += u; }
            s_w[lane] = wi - w;
        }
        __syncthreads();
        const int run = s_run;
        const int excl = run + s_w[warp] + incl - v;
        if (r < nrows) {
            row_off[r] = excl;
            if (r == ncy) ctl->blk_p1 = excl;
            if (r == (nczl - 1) * ncy) ctl->blk_pl = excl;
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_run = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) { ctl->nblocks = s_run; if (nczl < 2) { ctl->blk_p1 = s_run; ctl->blk_pl = 0; } }
}
// per-block tile statistics -> dynamic shared memory size and block size of the tile kernels
__global__ void k_block_stats(ClbGrid g, const int* __restrict__ cell_start, ClbCtl* ctl) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= grid_nblocks(g)) return;
    const int4 q = g.blk[b];
    int cx0 = q.x, bxe = q.y, cy = q.z, lz = q.w;
    bool whole = bxe + 2 > g.ncx;
    int W = whole ? g.ncx : bxe + 2;
    int T = 0, cmax = 0;
    for (int k = 0; k < 9; ++k) {
        int dy = k % 3 - 1, dz = k / 3 - 1;
        int yy = wrapi(cy + dy, g.ncy);
        int zz = wrapi(lz + dz, g.nplanes);
        for (int m = 0; m < W; ++m) {
            int cx = whole ? m : wrapi(cx0 - 1 + m, g.ncx);
            int c = (zz * g.ncy + yy) * g.ncx + cx;
            int cnt = cell_start[c + 1] - cell_start[c];
            T += cnt; cmax = max(cmax, cnt);
        }
    }
    int c0 = (lz * g.ncy + cy) * g.ncx + cx0;
    int nh = cell_start[c0 + bxe] - cell_start[c0];
    atomicMax(&ctl->tile_max, T);
    atomicMax(&ctl->home_max, nh);
    atomicMax(&ctl->cell_max, cmax);
}

// ---------------------------------------------------------------- integrator ----------------
struct ClbIntegParams {
    double dt;
    double invq[3], q[3];
    int langevin;          // thermostat enabled
    double pref1, pref2;   // -gamma, sqrt(24 kT gamma / dt)
    unsigned long long type_mask_lo; // bit t set -> type t thermalised (all ones when add_valid_types unused)
    uint64_t seed;
    uint64_t step;         // RNG step key of the half-kick being closed
    int criterion;
    int i0, i1;            // owned particle range
};
#define CLB_INT_SECOND 1
#define CLB_INT_FIRST 2

// VelocityVerlet::integrate2 (+ LangevinThermostat::thermalize) and/or integrate1.
// MODE = SECOND|FIRST fuses the closing half-kick of step i-1 with the opening half-kick and drift
// of step i (one pass over pos/vel/force instead of two).
template <int MODE>
__global__ void __launch_bounds__(256) k_integrate(ClbIntegParams P, int4* __restrict__ pos, ClbVel* __restrict__ vel,
                                                   double* __restrict__ force, int fstride,
                                                   const int4* __restrict__ xref, const int* __restrict__ slot,
                                                   int* __restrict__ image, ClbCtl* ctl) {
    if (*(volatile int*)&ctl->stall) return;
    int i = P.i0 + blockIdx.x * blockDim.x + threadIdx.x;
    float d2 = 0.f;
    if (i < P.i1) {
        const ClbVel v4 = vel[i];
        const double m = v4.w;
        double vx = v4.x, vy = v4.y, vz = v4.z;
        double fx = force[i], fy = force[i + fstride], fz = force[i + 2 * fstride];
        int4 p = pos[i];
        if (MODE & CLB_INT_SECOND) {
            if (P.langevin && ((P.type_mask_lo >> pw_type(p.w)) & 1ull)) {
                double u[3];
                clb_draw3(P.seed, CLB_STREAM_LANGEVIN, P.step, (uint32_t)slot[i], u);
                double sm = sqrt(m);
                fx += P.pref1 * m * vx + P.pref2 * sm * (u[0] - 0.5);
                fy += P.pref1 * m * vy + P.pref2 * sm * (u[1] - 0.5);
                fz += P.pref1 * m * vz + P.pref2 * sm * (u[2] - 0.5);
            }
            if (!(MODE & CLB_INT_FIRST)) { force[i] = fx; force[i + fstride] = fy; force[i + 2 * fstride] = fz; }
        }
        double k = ((MODE & CLB_INT_SECOND) && (MODE & CLB_INT_FIRST)) ? P.dt / m : 0.5 * P.dt / m;
        vx += k * fx; vy += k * fy; vz += k * fz;
        vel[i] = clb_make_vel(vx, vy, vz, m);
        if (MODE & CLB_INT_FIRST) {
            double dpx = P.dt * vx, dpy = P.dt * vy, dpz = P.dt * vz;
            int dx = __double2int_rn(dpx * P.invq[0]), dy = __double2int_rn(dpy * P.invq[1]), dz = __double2int_rn(dpz * P.invq[2]);
            int nx = wadd(p.x, dx), ny = wadd(p.y, dy), nz = wadd(p.z, dz);
            int wx = (dx > 0 && (unsigned)nx < (unsigned)p.x) ? 1 : ((dx < 0 && (unsigned)nx > (unsigned)p.x) ? -1 : 0);
            int wy = (dy > 0 && (unsigned)ny < (unsigned)p.y) ? 1 : ((dy < 0 && (unsigned)ny > (unsigned)p.y) ? -1 : 0);
            int wz = (dz > 0 && (unsigned)nz < (unsigned)p.z) ? 1 : ((dz < 0 && (unsigned)nz > (unsigned)p.z) ? -1 : 0);
            if (wx | wy | wz) { int s = slot[i]; image[3 * s] += wx; image[3 * s + 1] += wy; image[3 * s + 2] += wz; }
            pos[i] = make_int4(nx, ny, nz, p.w);
            if (P.criterion == 0) d2 = (float)(dpx * dpx + dpy * dpy + dpz * dpz);
            else {
                int4 r = xref[i];
                double ex = (double)wsub(nx, r.x) * P.q[0], ey = (double)wsub(ny, r.y) * P.q[1], ez = (double)wsub(nz, r.z) * P.q[2];
                d2 = __double2float_ru(ex * ex + ey * ey + ez * ez);
            }
        }
    }
    if (MODE & CLB_INT_FIRST) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) d2 = fmaxf(d2, __shfl_xor_sync(0xffffffffu, d2, d));
        __shared__ float s_m[8];
        if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = d2;
        __syncthreads();
        if (threadIdx.x == 0) {
            float mm = 0.f;
            for (int w = 0; w < (blockDim.x >> 5); ++w) mm = fmaxf(mm, s_m[w]);
            atomicMax(&ctl->maxdisp2_bits, __float_as_uint(mm));
        }
    }
}
// integrator.CapForce (src/start_simulation.py:320-324) [EXT, U26]: force vectors longer than `cap` are scaled back to that length
__global__ void k_cap_force(int i0, int i1, double* __restrict__ force, int fstride, double cap, const ClbCtl* ctl) {
    if (*(volatile const int*)&ctl->stall) return;
    int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= i1) return;
    const double fx = force[i], fy = force[i + fstride], fz = force[i + 2 * fstride], f2 = fx * fx + fy * fy + fz * fz;
    if (f2 > cap * cap) { const double k = cap / sqrt(f2); force[i] = fx * k; force[i + fstride] = fy * k; force[i + 2 * fstride] = fz * k; }
}
// LangevinThermostat at run entry (recalc1/updateForces/recalc2 with heatUp: pref2 *= sqrt(3))
__global__ void k_thermalize(ClbIntegParams P, double scale, uint32_t stream, const int4* __restrict__ pos,
                             const ClbVel* __restrict__ vel, const int* __restrict__ slot, double* __restrict__ force,
                             int fstride) {
    int i = P.i0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.i1) return;
    int4 p = pos[i];
    if (!((P.type_mask_lo >> pw_type(p.w)) & 1ull)) return;
    const ClbVel v4 = vel[i];
    double m = v4.w, sm = sqrt(m), u[3];
    clb_draw3(P.seed, stream, P.step, (uint32_t)slot[i], u);
    force[i] += P.pref1 * m * v4.x + P.pref2 * scale * sm * (u[0] - 0.5);
    force[i + fstride] += P.pref1 * m * v4.y + P.pref2 * scale * sm * (u[1] - 0.5);
    force[i + 2 * fstride] += P.pref1 * m * v4.z + P.pref2 * scale * sm * (u[2] - 0.5);
}
// skin/2 criterion evaluated on the device; sets ctl->stall so that every later kernel of the
// enqueued chunk becomes a no-op until the host has rebuilt the lists (no per-step host sync).
//   criterion 0  the reference's rule: sum over steps of the per-step maximum displacement > skin/2 ([EXT] VelocityVerlet)
//   criterion 1  true maximum displacement since the rebuild > skin/2 (never later than 0, same results)
//   criterion 2  (single GPU, option resort_criterion=2) like 1, but when the global maximum passes skin/2 the decision goes to a per-cell bound
//                (k_cell_disp / k_cell_pairs below): the list stays valid while D(c) + D(c') <= skin for every pair of cells
//                within two cells of each other, D(c) = largest displacement of the particles sorted into cell c at the rebuild.
//                Why that is exact: a pair that is NOT listed was farther apart than rc+skin at the rebuild; to come within rc
//                its two particles must together move more than skin (cells within two of each other: r0 > rc+skin), more than
//                edge - rc >= skin (|dc| = 2 in some dimension), or more than 2 edge - rc (farther cells: excluded by the cap
//                D_max <= skin).  Opt-in: measured on the 1M-bead melt it saves only 3 % of the rebuilds (66 instead of 68 per 400
//                steps), because among the 2500 beads around the fastest one there is always another fast one
//                (tests/test_gpu_parity.py::test_cell_pair_resort_criterion checks that the lists stay complete).
__global__ void k_check_resort(ClbCtl* ctl, int criterion, double half_skin, int step_index) {
    if (ctl->stall) return;
    float m2 = __uint_as_float(ctl->maxdisp2_bits);
    ctl->maxdisp2_bits = 0u;
    bool need;
    ctl->maybe = 0;
    if (criterion == 0) { ctl->accum_maxdist += sqrt((double)m2); need = ctl->accum_maxdist > half_skin; }
    else if (criterion == 2) {
        const double m = sqrt((double)m2);
        need = m > 2.0 * half_skin;                                   // cap: beyond it cells farther than two apart would matter
        if (!need && m > half_skin) { ctl->maybe = 1; ctl->maybe_step = step_index; }
    } else need = sqrt((double)m2) > half_skin;
    if (need || ctl->force_rebuild) { ctl->stall = 1; ctl->stall_step = step_index; }
    else ctl->steps_ok += 1;
}
// The two largest displacements of every cell (rounded up), only on the steps where the global maximum is between skin/2 and skin
__global__ void k_cell_disp(ClbCtl* ctl, int ncell, const int* __restrict__ cell_start, const int4* __restrict__ pos, const int4* __restrict__ xref,
                            double q0, double q1, double q2, float2* __restrict__ Dc) {
    if (!*(volatile int*)&ctl->maybe || *(volatile int*)&ctl->stall) return;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    float m1 = 0.f, m2 = 0.f;
    for (int i = cell_start[c]; i < cell_start[c + 1]; ++i) {
        const int4 p = pos[i], r = xref[i];
        const double ex = (double)wsub(p.x, r.x) * q0, ey = (double)wsub(p.y, r.y) * q1, ez = (double)wsub(p.z, r.z) * q2;
        const float d = __double2float_ru(ex * ex + ey * ey + ez * ez);
        if (d > m1) { m2 = m1; m1 = d; } else m2 = fmaxf(m2, d);
    }
    Dc[c] = make_float2(__fsqrt_ru(m1), __fsqrt_ru(m2));
}
// pairs inside one cell are bounded by its two largest displacements, pairs of different cells by the largest of each
__global__ void k_cell_pairs(ClbCtl* ctl, int ncx, int ncy, int ncz, const float2* __restrict__ Dc, float skin) {
    if (!*(volatile int*)&ctl->maybe || *(volatile int*)&ctl->stall) return;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncx * ncy * ncz) return;
    const float2 me = Dc[c];
    const float a = me.x;
    if (a <= 0.4999f * skin) return;        // D(c) + D(c') > skin needs one of the two above skin/2: that cell's thread finds the pair
    const int cx = c % ncx, cy = (c / ncx) % ncy, cz = c / (ncx * ncy);
    bool viol = __fadd_ru(a, me.y) > skin;
    for (int dz = -2; dz <= 2; ++dz) for (int dy = -2; dy <= 2; ++dy) for (int dx = -2; dx <= 2; ++dx) {
        int x = cx + dx, y = cy + dy, z = cz + dz;
        x += x < 0 ? ncx : 0; x -= x >= ncx ? ncx : 0; x += x < 0 ? ncx : 0; x -= x >= ncx ? ncx : 0;
        y += y < 0 ? ncy : 0; y -= y >= ncy ? ncy : 0; y += y < 0 ? ncy : 0; y -= y >= ncy ? ncy : 0;
        z += z < 0 ? ncz : 0; z -= z >= ncz ? ncz : 0; z += z < 0 ? ncz : 0; z -= z >= ncz ? ncz : 0;
        const int c2 = (z * ncy + y) * ncx + x;
        if (c2 != c) viol |= __fadd_ru(a, Dc[c2].x) > skin;
    }
    if (viol) { ctl->stall = 1; ctl->stall_step = ctl->maybe_step; }
}

// ---------------------------------------------------------------- bonded --------------------
struct ClbBondedDesc { int arity, typed, inter, npot, pot_off, active; };
struct ClbBPot { int t[4]; int kind, table; double p[6]; };
struct ClbBTabMeta { double x0, invdx, dx; int n, off; };

__device__ __forceinline__ const ClbBPot* bpot_lookup(const ClbBondedDesc& d, const ClbBPot* pots, const int* ty) {
    if (!d.typed) return d.npot ? pots + d.pot_off : nullptr;
    for (int k = 0; k < d.npot; ++k) {
        const ClbBPot* p = pots + d.pot_off + k;
        bool fwd = true, rev = true;
        for (int m = 0; m < d.arity; ++m) { fwd &= (p->t[m] == ty[m]); rev &= (p->t[m] == ty[d.arity - 1 - m]); }
        if (fwd || rev) return p;
    }
    return nullptr;
}
// 1/sqrt(x) in fp64 without the slow-path library call: MUFU.RSQ seed (24 bits) + two Newton steps (the second is
// nearly free and leaves ~1 ulp); x > 0 and finite.  Divisions and square roots of the bonded terms are built from it.
__device__ __forceinline__ double clb_rsqrt(double x) {
    int hi = __double2hiint(x), lo = __double2loint(x);
    unsigned fb = ((unsigned)(hi - 0x38000000) << 3) | ((unsigned)lo >> 29);
    float yf;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(yf) : "f"(__uint_as_float(fb)));
    unsigned yb = __float_as_uint(yf);
    double y = __hiloint2double((int)((yb >> 3) + 0x38000000u), (int)(yb << 29));
    double e = fma(-x * y, y, 1.0);
    y = fma(0.5 * y, e, y);
    e = fma(-x * y, y, 1.0);
    return fma(0.5 * y, e, y);
}
// returns F = -dU/dx and U for the scalar coordinate x (r, theta or phi)
__device__ __forceinline__ void bpot_eval(const ClbBPot* p, double x, const ClbBTabMeta* tm, const double4* cf, const double4* ce,
                                          double& F, double& E, bool want_e, unsigned& err) {
    switch (p->kind) {
        case 1: { double d = x - p->p[1]; F = -2.0 * p->p[0] * d; E = p->p[0] * d * d; break; }            // Harmonic
        case 3: { double d = x - p->p[1]; F = -2.0 * p->p[0] * d; E = p->p[0] * d * d; break; }            // AngularHarmonic
        case 6: { F = p->p[0] * sin(x - p->p[1]); E = p->p[0] * (1.0 + cos(x - p->p[1])); break; }         // Cosine
        case 7: { double d = x - p->p[1], xx = d / p->p[2]; F = -p->p[0] * d / (1.0 - xx * xx);
                  E = -0.5 * p->p[0] * p->p[2] * p->p[2] * log(1.0 - xx * xx); break; }                      // FENE
        case 9: { double d = x - p->p[1], xx = d / p->p[2], sr2 = p->p[3] * p->p[3] / (x * x), sr6 = sr2 * sr2 * sr2;          // FENE + LJ (func 9)
                  F = -p->p[0] * d / (1.0 - xx * xx) + 24.0 * p->p[4] * (2.0 * sr6 * sr6 - sr6) / x;
                  E = -0.5 * p->p[0] * p->p[2] * p->p[2] * log(1.0 - xx * xx) + 4.0 * p->p[4] * (sr6 * sr6 - sr6); break; }
        case 10: { if (x > p->p[2]) { F = 0.0; E = 0.0; break; }                                               // LennardJones on a pair list (1-4 pairs)
                   double sr2 = p->p[1] * p->p[1] / (x * x), sr6 = sr2 * sr2 * sr2;
                   F = 24.0 * p->p[0] * (2.0 * sr6 * sr6 - sr6) / x; E = 4.0 * p->p[0] * (sr6 * sr6 - sr6) - p->p[3]; break; }
        case 8: { double d = x - p->p[1]; d -= 6.283185307179586 * rint(d / 6.283185307179586);
                  F = -2.0 * p->p[0] * d; E = p->p[0] * d * d; break; }                                      // DihedralHarmonic
        case 2: case 4: case 5: {                                                                            // tables
            const ClbBTabMeta t = tm[p->table];
            double s = (x - t.x0) * t.invdx;
            int idx = (int)floor(s);
            if (idx < 0) { idx = 0; err |= CLB_EF_TABLE_RANGE; }
            if (idx > t.n - 2) { if (x > t.x0 + t.dx * (t.n - 1) * (1.0 + 1e-12)) err |= CLB_EF_TABLE_RANGE; idx = t.n - 2; }
            double d = x - (t.x0 + idx * t.dx);
            double4 c = cf[t.off + idx];
            F = c.x + d * (c.y + d * (c.z + d * c.w));
            if (want_e) { double4 e = ce[t.off + idx]; E = e.x + d * (e.y + d * (e.z + d * e.w)); }
            break;
        }
        default: F = 0.0; E = 0.0;
    }
}
// Owner-computes bonded forces: thread i evaluates every tuple it is a member of and keeps only its
// own force -> deterministic, no atomics (each tuple is evaluated arity times; O(N) work).
// FixedPairList/TripleList/QuadrupleList interaction templates [EXT], SURVEY 8a7-a9, U15.
template <bool ENERGY>
__global__ void __launch_bounds__(256) k_bonded(int i0, int i1, const int4* __restrict__ pos, ClbGeom geo,
                                                const int* __restrict__ rt_off, const int4* __restrict__ rt_mem,
                                                const int* __restrict__ rt_meta, const ClbBondedDesc* __restrict__ bdesc,
                                                const ClbBPot* __restrict__ pots, const ClbBTabMeta* __restrict__ btm,
                                                const double4* __restrict__ cf, const double4* __restrict__ ce,
                                                double* __restrict__ force, int fstride, int inter,
                                                double* __restrict__ partial, ClbCtl* ctl) {
    if (!ENERGY && *(volatile int*)&ctl->stall) return;
    int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    double ax = 0, ay = 0, az = 0, esum = 0;
    unsigned err = 0;
    if (i < i1) {
        int t0 = rt_off[i - i0], t1 = rt_off[i - i0 + 1];
        for (int t = t0; t < t1; ++t) {
            int4 mem = rt_mem[t];
            int meta = rt_meta[t];
            int role = meta & 3, li = meta >> 2;
            const ClbBondedDesc d = bdesc[li];
            if (ENERGY && (d.inter != inter || role != 0)) continue;
            int idx[4] = {mem.x, mem.y, mem.z, mem.w};
            int4 P[4]; int ty[4];
            for (int m = 0; m < d.arity; ++m) { P[m] = __ldg(pos + idx[m]); ty[m] = pw_type(P[m].w); }
            const ClbBPot* pp = bpot_lookup(d, pots, ty);
            if (!pp) continue;
            double F, E = 0;
            if (d.arity == 2) {
                int o = 1 - role;
                double dx = lat2d(wsub(P[role].x, P[o].x)) * geo.q[0], dy = lat2d(wsub(P[role].y, P[o].y)) * geo.q[1], dz = lat2d(wsub(P[role].z, P[o].z)) * geo.q[2];
                const double r2 = dx * dx + dy * dy + dz * dz, ir = clb_rsqrt(r2), r = r2 * ir;
                bpot_eval(pp, r, btm, cf, ce, F, E, ENERGY, err);
                double fr = F * ir;
                ax += fr * dx; ay += fr * dy; az += fr * dz;
            } else if (d.arity == 3) {
                double d1[3] = {lat2d(wsub(P[0].x, P[1].x)) * geo.q[0], lat2d(wsub(P[0].y, P[1].y)) * geo.q[1], lat2d(wsub(P[0].z, P[1].z)) * geo.q[2]};
                double d2[3] = {lat2d(wsub(P[2].x, P[1].x)) * geo.q[0], lat2d(wsub(P[2].y, P[1].y)) * geo.q[1], lat2d(wsub(P[2].z, P[1].z)) * geo.q[2]};
                const double i1 = clb_rsqrt(d1[0] * d1[0] + d1[1] * d1[1] + d1[2] * d1[2]);      // 1/r1
                const double i2 = clb_rsqrt(d2[0] * d2[0] + d2[1] * d2[1] + d2[2] * d2[2]);      // 1/r2
                const double i12 = i1 * i2;
                double c = (d1[0] * d2[0] + d1[1] * d2[1] + d1[2] * d2[2]) * i12;
                c = fmin(1.0, fmax(-1.0, c));
                double th = acos(c);
                const double s2 = fmax(1.0 - c * c, 1e-18);                                          // sin^2, floor = (1e-9)^2
                const double isn = clb_rsqrt(s2);                                                    // 1/sin(theta)
                bpot_eval(pp, th, btm, cf, ce, F, E, ENERGY, err);
                double g[3];
                const double c11 = c * i1 * i1, c22 = c * i2 * i2;
                for (int k = 0; k < 3; ++k) {
                    double g1 = -(d2[k] * i12 - c11 * d1[k]) * isn;
                    double g3 = -(d1[k] * i12 - c22 * d2[k]) * isn;
                    g[k] = role == 0 ? F * g1 : (role == 2 ? F * g3 : -F * (g1 + g3));
                }
                ax += g[0]; ay += g[1]; az += g[2];
            } else {
                double rij[3] = {lat2d(wsub(P[0].x, P[1].x)) * geo.q[0], lat2d(wsub(P[0].y, P[1].y)) * geo.q[1], lat2d(wsub(P[0].z, P[1].z)) * geo.q[2]};
                double rkj[3] = {lat2d(wsub(P[2].x, P[1].x)) * geo.q[0], lat2d(wsub(P[2].y, P[1].y)) * geo.q[1], lat2d(wsub(P[2].z, P[1].z)) * geo.q[2]};
                double rkl[3] = {lat2d(wsub(P[2].x, P[3].x)) * geo.q[0], lat2d(wsub(P[2].y, P[3].y)) * geo.q[1], lat2d(wsub(P[2].z, P[3].z)) * geo.q[2]};
                double mv[3] = {rij[1] * rkj[2] - rij[2] * rkj[1], rij[2] * rkj[0] - rij[0] * rkj[2], rij[0] * rkj[1] - rij[1] * rkj[0]};
                double nv[3] = {rkj[1] * rkl[2] - rkj[2] * rkl[1], rkj[2] * rkl[0] - rkj[0] * rkl[2], rkj[0] * rkl[1] - rkj[1] * rkl[0]};
                double m2 = mv[0] * mv[0] + mv[1] * mv[1] + mv[2] * mv[2];
                double n2 = nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2];
                double rkj2 = rkj[0] * rkj[0] + rkj[1] * rkj[1] + rkj[2] * rkj[2], nrkj = sqrt(rkj2);
                double cphi = (mv[0] * nv[0] + mv[1] * nv[1] + mv[2] * nv[2]) / sqrt(m2 * n2);
                cphi = fmin(1.0, fmax(-1.0, cphi));
                double phi = acos(cphi);
                if (rij[0] * nv[0] + rij[1] * nv[1] + rij[2] * nv[2] < 0) phi = -phi;
                bpot_eval(pp, phi, btm, cf, ce, F, E, ENERGY, err);
                double pq = (rij[0] * rkj[0] + rij[1] * rkj[1] + rij[2] * rkj[2]) / rkj2;
                double qq = (rkl[0] * rkj[0] + rkl[1] * rkj[1] + rkl[2] * rkj[2]) / rkj2;
                for (int k = 0; k < 3; ++k) {
                    double fi = F * (nrkj / m2) * mv[k], fl = -F * (nrkj / n2) * nv[k];
                    double sv = pq * fi - qq * fl;
                    double fk = role == 0 ? fi : (role == 1 ? -fi + sv : (role == 2 ? -fl - sv : fl));
                    if (k == 0) ax += fk; else if (k == 1) ay += fk; else az += fk;
                }
            }
            esum += E;
        }
        if (!ENERGY) { force[i] += ax; force[i + fstride] += ay; force[i + 2 * fstride] += az; }
    }
    if (err) atomicOr(&ctl->err, err);
    if (ENERGY) {
        __shared__ double s_red[8];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) esum += __shfl_down_sync(0xffffffffu, esum, d);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = esum;
        __syncthreads();
        if (threadIdx.x == 0) { double a = 0; for (int w = 0; w < (blockDim.x >> 5); ++w) a += s_red[w]; partial[blockIdx.x] = a; }
    }
}
// count of bonded terms per owned sorted particle (CSR by slot -> CSR by sorted index)
__global__ void k_term_counts(int i0, int i1, const int* __restrict__ slot, const int* __restrict__ term_off, int* __restrict__ cnt) {
    int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= i1) return;
    int s = slot[i];
    cnt[i - i0] = term_off[s + 1] - term_off[s];
}
// resolve tuple members (slots) to sorted indices after every re-sort
__global__ void k_term_resolve(int i0, int i1, const int* __restrict__ slot, const int* __restrict__ term_off,
                               const int* __restrict__ term_meta, const int* __restrict__ term_tuple,
                               const int* const* __restrict__ list_tuples, const ClbBondedDesc* __restrict__ bdesc,
                               const int* __restrict__ id2idx, const int* __restrict__ rt_off, int4* __restrict__ rt_mem,
                               int* __restrict__ rt_meta, ClbCtl* ctl) {
    int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= i1) return;
    int s = slot[i];
    int t0 = term_off[s], n = term_off[s + 1] - t0, o = rt_off[i - i0];
    for (int t = 0; t < n; ++t) {
        int meta = term_meta[t0 + t], li = meta >> 2, tu = term_tuple[t0 + t];
        int ar = bdesc[li].arity;
        const int* ids = list_tuples[li] + (size_t)tu * ar;
        int m[4] = {0, 0, 0, 0};
        for (int k = 0; k < ar; ++k) { m[k] = id2idx[ids[k]]; if (m[k] < 0) atomicOr(&ctl->err, CLB_EF_PARTNER_LOST); }
        rt_mem[o + t] = make_int4(m[0], m[1], m[2], m[3]);
        rt_meta[o + t] = meta;
    }
}
// expand tuple lists into (member slot, meta, tuple) rows for the CSR-by-slot sort
__global__ void k_term_expand(int n, int arity, int li, const int* __restrict__ tuples, int base, int* __restrict__ key,
                              unsigned long long* __restrict__ val) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    for (int r = 0; r < arity; ++r) {
        key[base + t * arity + r] = tuples[(size_t)t * arity + r];
        val[base + t * arity + r] = ((unsigned long long)(unsigned)((li << 2) | r) << 32) | (unsigned)t;
    }
}
__global__ void k_term_unpack(int n, const unsigned long long* __restrict__ val, int* __restrict__ meta, int* __restrict__ tuple) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    meta[t] = (int)(val[t] >> 32); tuple[t] = (int)(val[t] & 0xffffffffu);
}
// row offsets of a sorted key array: off[s] = lower_bound(keys, s)
__global__ void k_lower_bounds(int nkeys, const int* __restrict__ keys, int nslots, int* __restrict__ off) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > nslots) return;
    int lo = 0, hi = nkeys;
    while (lo < hi) { int m = (lo + hi) >> 1; if (keys[m] < s) lo = m + 1; else hi = m; }
    off[s] = lo;
}

// ---------------------------------------------------------------- observables ---------------
__global__ void __launch_bounds__(256) k_kinetic(int i0, int i1, const ClbVel* __restrict__ vel, double* __restrict__ partial) {
    int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0;
    if (i < i1) { const ClbVel v = vel[i]; e = 0.5 * v.w * (v.x * v.x + v.y * v.y + v.z * v.z); }
    __shared__ double s_red[8];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) e += __shfl_down_sync(0xffffffffu, e, d);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = e;
    __syncthreads();
    if (threadIdx.x == 0) { double a = 0; for (int w = 0; w < (blockDim.x >> 5); ++w) a += s_red[w]; partial[blockIdx.x] = a; }
}
// fixed-order final reduction of per-block partial sums (bit-reproducible observables)
__global__ void k_sum_partials(int n, const double* __restrict__ partial, double* __restrict__ out) {
    __shared__ double s[256];
    double a = 0;
    for (int i = threadIdx.x; i < n; i += 256) a += partial[i];
    s[threadIdx.x] = a;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) { if (threadIdx.x < d) s[threadIdx.x] += s[threadIdx.x + d]; __syncthreads(); }
    if (threadIdx.x == 0) *out = s[0];
}
__global__ void k_sum_partials_u64(int n, const unsigned long long* __restrict__ partial, unsigned long long* __restrict__ out) {
    __shared__ unsigned long long s[256];
    unsigned long long a = 0;
    for (int i = threadIdx.x; i < n; i += 256) a += partial[i];
    s[threadIdx.x] = a;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) { if (threadIdx.x < d) s[threadIdx.x] += s[threadIdx.x + d]; __syncthreads(); }
    if (threadIdx.x == 0) *out = s[0];
}
__global__ void k_count_type(int i0, int i1, const int4* __restrict__ pos, int type, int state, unsigned long long* __restrict__ out) {
    int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    bool hit = false;
    if (i < i1) { int w = pos[i].w; hit = pw_type(w) == type && (state < 0 || pw_state(w) == state); }
    unsigned b = __ballot_sync(0xffffffffu, hit);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(out, (unsigned long long)__popc(b));
}
__global__ void k_zero_force(int n, double* __restrict__ f, int fstride) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { f[i] = 0; f[i + fstride] = 0; f[i + 2 * fstride] = 0; }
}
