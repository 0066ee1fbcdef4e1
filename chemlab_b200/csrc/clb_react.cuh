// clb_react.cuh -- ChemicalReaction::React on the device:
//   k_react_scan      candidate search over the tile neighbour lists + acceptance draw, warp-aggregated
//                     compaction into a candidate buffer
//   k_uniq_*          UniqueA / UniqueB by order-independent 64-bit atomicMin keys (deterministic)
//   k_resolve         U8 greedy in canonical order, solved by dependency rounds inside one CTA
//   k_apply_*         reactant property changes, state deltas, bond append, graph update,
//                     neighbour-property BFS with deterministic claims, TopologyManager tuple generation
// [EXT] integrator/ChemicalReaction.cpp, ChemicalReactionPostProcess.cpp, TopologyManager.cpp;
// chemlab call sites: src/chemlab/reaction_setup.py:81-163,417-427,506; reaction_post_process.py:76-115.
#pragma once
#include <cooperative_groups.h>

#include "clb_common.cuh"
#include "clb_tile.cuh"
namespace cg = cooperative_groups;

struct ClbReactSpec {
    int type_1, type_2, delta_1, delta_2, min1, max1, min2, max2;
    double cutoff2, min_cutoff2, p;
    int list, intramolecular, intraresidual, is_virtual, active;
    int conn_n, conn_off, pad;     // RestrictReaction: conn_n >= 0 -> only pairs among conn[conn_off, conn_off + conn_n) may react; -1 unrestricted
};
struct ClbCand { int a, b, r, accepted; double d2; unsigned long long rnd; };
struct ClbChange { int reaction, side, nb_level, old_type, new_type, state_mode, state_value, pad; double new_mass, new_q; };
struct ClbTmReg { int list, arity; int t[4]; };
struct ClbListDev { int* tuples; int arity, tm_observed, excl_observed, pad; };
template <int N> struct ClbTup { int v[N]; };
template <int N> struct ClbTupLess {
    __host__ __device__ bool operator()(const ClbTup<N>& a, const ClbTup<N>& b) const {
#pragma unroll
        for (int k = 0; k < N; ++k) { if (a.v[k] != b.v[k]) return a.v[k] < b.v[k]; }
        return false;
    }
};

// RestrictReaction.define_connection (reaction_setup.py:115-126): sorted keys (lower slot << 32 | higher slot) per reaction
__device__ __forceinline__ bool conn_has(const unsigned long long* __restrict__ conn, int off, int n, int lo, int hi) {
    const unsigned long long key = ((unsigned long long)(unsigned)lo << 32) | (unsigned)hi;
    int a = 0, b = n;
    while (a < b) { const int m = (a + b) >> 1; if (__ldg(conn + off + m) < key) a = m + 1; else b = m; }
    return a < n && __ldg(conn + off + a) == key;
}

__device__ __forceinline__ bool side_ok(const ClbReactSpec& r, int wa, int wb) {
    int sa = pw_state(wa), sb = pw_state(wb);
    return pw_type(wa) == r.type_1 && pw_type(wb) == r.type_2 && sa >= r.min1 && sa < r.max1 && sb >= r.min2 && sb < r.max2;
}

__global__ void __launch_bounds__(512) k_react_scan(ClbGrid g, ClbGeom geo, const int* __restrict__ cell_start,
                                                    const int4* __restrict__ pos, const int* __restrict__ slot,
                                                    const unsigned short* __restrict__ entries,
                                                    const int* __restrict__ nl_count, int cap,
                                                    const ClbReactSpec* __restrict__ specs, int nspec,
                                                    const int* __restrict__ resid, const int* __restrict__ mol,
                                                    const unsigned long long* __restrict__ conn,
                                                    uint64_t seed, uint64_t step, ClbCand* __restrict__ cands,
                                                    unsigned long long candcap, ClbCtl* ctl) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_off[CLB_TILE_CELLS + 1];
    __shared__ int s_src[CLB_TILE_CELLS];
    __shared__ ClbReactSpec s_spec[CLB_MAX_REACTIONS];
    for (int i = threadIdx.x; i < nspec; i += blockDim.x) s_spec[i] = specs[i];
    for (int b = blockIdx.x; b < g.nblocks; b += gridDim.x) {
        TileCtx t;
        tile_geometry(g, b, t);
        __syncthreads();
        tile_offsets(g, t, cell_start, s_off, s_src);
        int4* s_pos = reinterpret_cast<int4*>(smem);
        int* s_slot = reinterpret_cast<int*>(s_pos + t.T);
        tile_stage(t, s_off, s_src, pos, s_pos, slot, s_slot, nullptr);
        __syncthreads();
        for (int p = threadIdx.x; p < t.nh; p += blockDim.x) {
            const int gi = t.hs + p;
            const int4 pi = __ldg(pos + gi);
            // a home particle that fits neither side of any active reaction has no candidate pair: skip its row
            bool may = false;
            for (int ri = 0; ri < nspec; ++ri) {
                const ClbReactSpec& r = s_spec[ri];
                if (!r.active) continue;
                const int ty = pw_type(pi.w), st = pw_state(pi.w);
                may |= (ty == r.type_1 && st >= r.min1 && st < r.max1) || (ty == r.type_2 && st >= r.min2 && st < r.max2);
            }
            if (!may) continue;
            const int si = __ldg(slot + gi);
            const int cnt = __ldg(nl_count + gi);
            const uint4* row = reinterpret_cast<const uint4*>(entries + (size_t)gi * cap);   // cap % 8 == 0: 8 entries per load
            for (int k0 = 0; k0 < cnt; k0 += 8) {
                const uint4 q = __ldg(row + (k0 >> 3));
                const unsigned wds[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (k0 + u >= cnt) break;
                    const unsigned e = (u & 1) ? (wds[u >> 1] >> 16) : (wds[u >> 1] & 0xffffu);
                    const int sj = s_slot[e];
                    if (si >= sj) continue;           // every unordered pair once, lower slot first
                    const int4 pj = s_pos[e];
                    double dx = lat2d(wsub(pi.x, pj.x)) * geo.q[0], dy = lat2d(wsub(pi.y, pj.y)) * geo.q[1], dz = lat2d(wsub(pi.z, pj.z)) * geo.q[2];
                    double d2 = dx * dx + dy * dy + dz * dz;
                    for (int ri = 0; ri < nspec; ++ri) {
                        const ClbReactSpec& r = s_spec[ri];
                        if (!r.active) continue;
                        int A, B;
                        if (side_ok(r, pi.w, pj.w)) { A = si; B = sj; }
                        else if (side_ok(r, pj.w, pi.w)) { A = sj; B = si; }
                        else continue;
                        if (!(d2 >= r.min_cutoff2 && d2 < r.cutoff2)) continue;               // U3
                        if (!r.intraresidual && __ldg(resid + A) == __ldg(resid + B)) continue; // U10
                        if (!r.intramolecular && __ldg(mol + A) == __ldg(mol + B)) continue;
                        if (r.conn_n >= 0 && !conn_has(conn, r.conn_off, r.conn_n, si, sj)) continue;   // si < sj
                        uint32_t w[4], h[4];
                        clb_draw_pair(seed, CLB_STREAM_REACT, step, (uint32_t)si, (uint32_t)sj, (uint32_t)ri, w);
                        clb_draw_pair(seed, CLB_STREAM_PARTNER, step, (uint32_t)si, (uint32_t)sj, (uint32_t)ri, h);
                        ClbCand c;
                        c.a = A; c.b = B; c.r = ri; c.d2 = d2;
                        c.accepted = ((double)w[0] * (1.0 / 4294967296.0) < r.p) ? 1 : 0;        // U5
                        c.rnd = ((unsigned long long)h[0] << 32) | h[1];
                        cg::coalesced_group grp = cg::coalesced_threads();
                        unsigned long long base = 0;
                        if (grp.thread_rank() == 0) base = atomicAdd(&ctl->ncand, (unsigned long long)grp.size());
                        base = grp.shfl(base, 0);
                        unsigned long long o = base + grp.thread_rank();
                        if (o < candcap) cands[o] = c;
                    }
                }
            }
        }
    }
}

__global__ void k_cand_keys(int n, const ClbCand* __restrict__ c, unsigned long long* __restrict__ key, int* __restrict__ val) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    key[k] = ((unsigned long long)(unsigned)c[k].a << 36) | ((unsigned long long)(unsigned)c[k].b << 8) | (unsigned)c[k].r;
    val[k] = k;
}
__global__ void k_cand_gather(int n, const int* __restrict__ perm, const ClbCand* __restrict__ in, ClbCand* __restrict__ out, int* __restrict__ alive) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    out[k] = in[perm[k]];
    alive[k] = out[k].accepted;
}
// order inside a partner group (U7): nearest -> (d2, partner, reaction); random -> (hash, partner, reaction)
__device__ __forceinline__ unsigned long long uniq_key1(const ClbCand& c, int nearest) {
    return nearest ? (unsigned long long)__double_as_longlong(c.d2) : c.rnd;
}
__global__ void k_uniq1(int n, const ClbCand* __restrict__ c, const int* __restrict__ alive, int role, int nearest, unsigned long long* __restrict__ best1) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n || !alive[k]) return;
    atomicMin(best1 + (role == 0 ? c[k].a : c[k].b), uniq_key1(c[k], nearest));
}
__global__ void k_uniq2(int n, const ClbCand* __restrict__ c, const int* __restrict__ alive, int role, int nearest,
                        const unsigned long long* __restrict__ best1, unsigned long long* __restrict__ best2) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n || !alive[k]) return;
    int me = role == 0 ? c[k].a : c[k].b, partner = role == 0 ? c[k].b : c[k].a;
    if (uniq_key1(c[k], nearest) != best1[me]) return;
    atomicMin(best2 + me, ((unsigned long long)(unsigned)partner << 8) | (unsigned)c[k].r);
}
__global__ void k_uniq3(int n, const ClbCand* __restrict__ c, int* __restrict__ alive, int role, int nearest,
                        const unsigned long long* __restrict__ best1, const unsigned long long* __restrict__ best2) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n || !alive[k]) return;
    int me = role == 0 ? c[k].a : c[k].b, partner = role == 0 ? c[k].b : c[k].a;
    bool win = uniq_key1(c[k], nearest) == best1[me] && ((((unsigned long long)(unsigned)partner << 8) | (unsigned)c[k].r) == best2[me]);
    if (!win) alive[k] = 0;
}
// reset only the per-slot entries the candidates touch (cheaper than clearing n entries)
__global__ void k_uniq_reset(int n, const ClbCand* __restrict__ c, unsigned long long* __restrict__ b1, unsigned long long* __restrict__ b2,
                             int* __restrict__ asA, int* __restrict__ inB) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int a = c[k].a, b = c[k].b;
    b1[a] = ~0ull; b1[b] = ~0ull; b2[a] = ~0ull; b2[b] = ~0ull;
    asA[a] = -1; asA[b] = -1; inB[a] = -1; inB[b] = -1;
}
// U8: one reaction per particle per interval, greedy in canonical (A,B,r) order.  After UniqueA and
// UniqueB every particle is A in at most one surviving pair and B in at most one, so pair k only
// conflicts with the pair where its A is a B and the pair where its B is an A; k is decided once
// those with a smaller index are decided.  Rounds run inside ONE CTA until nothing is undecided.
// surv = indices of surviving candidates in canonical order.
__global__ void __launch_bounds__(1024) k_resolve(int ns, const int* __restrict__ surv, const ClbCand* __restrict__ c,
                                                  int* __restrict__ asA, int* __restrict__ inB, int* __restrict__ status,
                                                  int max_per_interval, int* __restrict__ ev, ClbCtl* ctl) {
    __shared__ int s_undecided;
    __shared__ int s_scan[1024];
    __shared__ int s_base;
    for (int k = threadIdx.x; k < ns; k += blockDim.x) { const ClbCand& x = c[surv[k]]; asA[x.a] = k; inB[x.b] = k; status[k] = 0; }
    __syncthreads();
    int rounds = 0;
    for (;;) {
        if (threadIdx.x == 0) s_undecided = 0;
        __syncthreads();
        for (int k = threadIdx.x; k < ns; k += blockDim.x) {
            if (status[k] != 0) continue;
            const ClbCand& x = c[surv[k]];
            int c1 = inB[x.a], c2 = asA[x.b];     // pairs sharing A (as their B) / sharing B (as their A)
            bool wait = false, blocked = false;
            if (c1 >= 0 && c1 < k) { int s = status[c1]; wait |= (s == 0); blocked |= (s == 1); }
            if (c2 >= 0 && c2 < k) { int s = status[c2]; wait |= (s == 0); blocked |= (s == 1); }
            if (blocked) status[k] = 2;
            else if (!wait) status[k] = 1;
            else atomicAdd(&s_undecided, 1);
        }
        __syncthreads();
        ++rounds;
        if (s_undecided == 0) break;
        __syncthreads();
    }
    // ordered compaction of applied pairs (+ max_per_interval cap)
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int k0 = 0; k0 < ns; k0 += blockDim.x) {
        int k = k0 + threadIdx.x;
        int f = (k < ns && status[k] == 1) ? 1 : 0;
        s_scan[threadIdx.x] = f;
        __syncthreads();
        for (int d = 1; d < blockDim.x; d <<= 1) {
            int v = threadIdx.x >= d ? s_scan[threadIdx.x - d] : 0;
            __syncthreads();
            s_scan[threadIdx.x] += v;
            __syncthreads();
        }
        int rank = s_base + s_scan[threadIdx.x] - f;
        if (f) {
            if (max_per_interval > 0 && rank >= max_per_interval) status[k] = 2;
            else ev[rank] = surv[k];
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_base += s_scan[threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x == 0) { ctl->nev = max_per_interval > 0 ? min(s_base, max_per_interval) : s_base; ctl->rounds = rounds; }
}

// Particle properties during a reaction pass: the type|state word of EVERY particle lives in the replicated
// per-slot array `wslot` (identical on all ranks, so every rank takes identical decisions); the copy inside
// pos[].w, the mass (vel[].w) and the charge are updated when the particle is stored locally (idx >= 0).
__device__ __forceinline__ void apply_props(const ClbChange& g, int idx, int s, int* wslot, int4* pos, ClbVel* vel, double* charge) {
    int w = wslot[s];
    if (pw_type(w) != g.old_type) return;
    int st = pw_state(w);
    if (g.state_mode == 1) st = g.state_value; else if (g.state_mode == 2) st += g.state_value;
    w = pw_pack(g.new_type, st);
    wslot[s] = w;
    if (idx >= 0) { pos[idx].w = w; if (g.new_mass > 0) vel[idx].w = g.new_mass; }
    if (g.new_q == g.new_q) charge[s] = g.new_q;
}
// phase 5: reactant changes (nb_level 0) in rule order, then state deltas; per-list bond ranks
__global__ void k_apply_reactants(int nev, const int* __restrict__ ev, const ClbCand* __restrict__ c, const ClbReactSpec* __restrict__ specs,
                                  const ClbChange* __restrict__ chg, int nchg, const int* __restrict__ id2idx, int* wslot, int4* pos, ClbVel* vel,
                                  double* charge, unsigned long long* __restrict__ counters) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nev) return;
    const ClbCand x = c[ev[e]];
    const ClbReactSpec r = specs[x.r];
    int ia = id2idx[x.a], ib = id2idx[x.b];
    for (int q = 0; q < nchg; ++q) {
        const ClbChange g = chg[q];
        if (g.reaction != x.r || g.nb_level != 0) continue;
        if (g.side & 1) apply_props(g, ia, x.a, wslot, pos, vel, charge);
        if (g.side & 2) apply_props(g, ib, x.b, wslot, pos, vel, charge);
    }
    int wa = wslot[x.a], wb = wslot[x.b];
    wa = pw_pack(pw_type(wa), pw_state(wa) + r.delta_1);
    wb = pw_pack(pw_type(wb), pw_state(wb) + r.delta_2);
    wslot[x.a] = wa; wslot[x.b] = wb;
    if (ia >= 0) pos[ia].w = wa;
    if (ib >= 0) pos[ib].w = wb;
    atomicAdd(counters + x.r, 1ull);
}
// phase 6: bonds -> tuple lists (slot at list_n[list] + rank in event order), graph, exclusions
__global__ void __launch_bounds__(1024) k_event_ranks(int nev, const int* __restrict__ ev, const ClbCand* __restrict__ c,
                                                      const ClbReactSpec* __restrict__ specs, int nlists, int* __restrict__ erank,
                                                      int* __restrict__ list_n) {
    // one CTA; for every list a block-wide ordered scan over the events that append to it
    __shared__ int s_scan[1024];
    __shared__ int s_base;
    for (int l = 0; l < nlists; ++l) {
        if (threadIdx.x == 0) s_base = list_n[l];
        __syncthreads();
        for (int k0 = 0; k0 < nev; k0 += blockDim.x) {
            int k = k0 + threadIdx.x;
            int f = 0;
            if (k < nev) { const ClbReactSpec& r = specs[c[ev[k]].r]; f = (!r.is_virtual && r.list == l) ? 1 : 0; }
            s_scan[threadIdx.x] = f;
            __syncthreads();
            for (int d = 1; d < blockDim.x; d <<= 1) {
                int v = threadIdx.x >= d ? s_scan[threadIdx.x - d] : 0;
                __syncthreads();
                s_scan[threadIdx.x] += v;
                __syncthreads();
            }
            if (f) erank[k] = s_base + s_scan[threadIdx.x] - 1;
            __syncthreads();
            if (threadIdx.x == blockDim.x - 1) s_base += s_scan[threadIdx.x];
            __syncthreads();
        }
        if (threadIdx.x == 0) list_n[l] = s_base;
        __syncthreads();
    }
}
__device__ __forceinline__ bool graph_has(const int* adj, const int* deg, int a, int b) {
    for (int k = 0; k < deg[a]; ++k) if (adj[a * CLB_MAXDEG + k] == b) return true;
    return false;
}
__global__ void k_apply_bonds(int nev, const int* __restrict__ ev, const ClbCand* __restrict__ c, const ClbReactSpec* __restrict__ specs,
                              const ClbListDev* __restrict__ lists, const int* __restrict__ erank, int* __restrict__ adj, int* __restrict__ deg,
                              int2* __restrict__ excl_pairs, unsigned long long* __restrict__ nexcl, ClbCtl* ctl) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nev) return;
    const ClbCand x = c[ev[e]];
    const ClbReactSpec r = specs[x.r];
    if (r.is_virtual) return;
    const ClbListDev L = lists[r.list];
    L.tuples[2 * (size_t)erank[e]] = x.a; L.tuples[2 * (size_t)erank[e] + 1] = x.b;
    if (L.tm_observed && !graph_has(adj, deg, x.a, x.b)) {
        // a particle takes part in at most one event per pass (U8): plain appends are race-free
        if (deg[x.a] < CLB_MAXDEG && deg[x.b] < CLB_MAXDEG) {
            adj[x.a * CLB_MAXDEG + deg[x.a]] = x.b; deg[x.a] += 1;
            adj[x.b * CLB_MAXDEG + deg[x.b]] = x.a; deg[x.b] += 1;
        } else atomicOr(&ctl->err, CLB_EF_DEGREE);
    }
    if (L.excl_observed) {
        unsigned long long o = atomicAdd(nexcl, 1ull);
        excl_pairs[o] = make_int2(min(x.a, x.b), max(x.a, x.b));
    }
}
// phase 7: PostProcessChangeNeighboursProperty -- BFS to exactly nb_level bonds; every reached particle
// whose CURRENT type equals the rule's old type files a claim (event, side, level, rule); the smallest
// claim per particle wins (deterministic stand-in for the sequential event order of the reference).
__global__ void k_nb_claims(int nev, const int* __restrict__ ev, const ClbCand* __restrict__ c, const ClbChange* __restrict__ chg, int nchg,
                            const int* __restrict__ adj, const int* __restrict__ deg, const int* __restrict__ wslot,
                            unsigned long long* __restrict__ claim, int* __restrict__ touched,
                            unsigned long long* __restrict__ ntouched, unsigned long long touchcap, ClbCtl* ctl) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * nev) return;
    int e = t >> 1, side = (t & 1) + 1;
    const ClbCand x = c[ev[e]];
    int root = side == 1 ? x.a : x.b;
    int maxlev = 0;
    for (int q = 0; q < nchg; ++q) if (chg[q].reaction == x.r && (chg[q].side & side) && chg[q].nb_level > maxlev) maxlev = chg[q].nb_level;
    if (!maxlev) return;
    int front[64], seen[192], nf = 1, ns = 1;
    front[0] = root; seen[0] = root;
    for (int lev = 1; lev <= maxlev; ++lev) {
        int nxt[64], nn = 0;
        for (int a = 0; a < nf; ++a) for (int k = 0; k < deg[front[a]]; ++k) {
            int y = adj[front[a] * CLB_MAXDEG + k];
            bool dup = false;
            for (int z = 0; z < ns; ++z) dup |= (seen[z] == y);
            if (!dup) {
                if (nn < 64 && ns < 192) { nxt[nn++] = y; seen[ns++] = y; }
                else atomicOr(&ctl->err, CLB_EF_BFS_OVERFLOW);   // a silently dropped neighbour would miss its type change
            }
        }
        for (int a = 0; a < nn; ++a) {
            int ty = pw_type(wslot[nxt[a]]);
            for (int q = 0; q < nchg; ++q) {
                const ClbChange& g = chg[q];
                if (g.reaction == x.r && (g.side & side) && g.nb_level == lev && g.old_type == ty) {
                    unsigned long long key = ((unsigned long long)e << 24) | ((unsigned long long)(side - 1) << 20) | ((unsigned long long)lev << 12) | (unsigned)q;
                    unsigned long long old = atomicMin(claim + nxt[a], key);
                    if (old == ~0ull) { unsigned long long o = atomicAdd(ntouched, 1ull); if (o < touchcap) touched[o] = nxt[a]; }
                    break;   // first matching rule of this (event, side, level)
                }
            }
        }
        for (int a = 0; a < nn; ++a) front[a] = nxt[a];
        nf = nn;
    }
}
__global__ void k_nb_apply(int ntouched, const int* __restrict__ touched, unsigned long long* __restrict__ claim, const ClbChange* __restrict__ chg,
                           const int* __restrict__ id2idx, int* wslot, int4* pos, ClbVel* vel, double* charge) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntouched) return;
    int s = touched[t];
    unsigned long long key = claim[s];
    claim[s] = ~0ull;
    apply_props(chg[key & 0xfff], id2idx[s], s, wslot, pos, vel, charge);
}
// phase 8: TopologyManager -- angles/dihedrals through each new bond whose (final) type tuple is
// registered.  A tuple containing several bonds of this pass is emitted by the LAST of them in event
// order.  MODE 0 counts per target list, MODE 1 writes (cursor by atomicAdd; the host sorts the new
// segment afterwards so that list order is deterministic).
__device__ __forceinline__ int find_event_of_bond(int nev, const int* ev, const ClbCand* c, const ClbReactSpec* specs, const int* ev_of_slot, int a, int b) {
    // event index that created bond (a,b) in this pass, or -1
    int e = ev_of_slot[a];
    if (e < 0 || e != ev_of_slot[b]) return -1;
    const ClbCand& x = c[ev[e]];
    if (specs[x.r].is_virtual) return -1;
    return ((x.a == a && x.b == b) || (x.a == b && x.b == a)) ? e : -1;
}
__device__ __forceinline__ void tm_emit(int mode, int ar, const int* ids, const int* wslot, const ClbTmReg* regs, int nreg,
                                        const ClbListDev* lists, int* list_cursor, int* list_count, int2* excl_pairs,
                                        unsigned long long* nexcl, unsigned long long* nexcl_count) {
    int ty[4];
    for (int m = 0; m < ar; ++m) ty[m] = pw_type(wslot[ids[m]]);
    for (int k = 0; k < nreg; ++k) {
        const ClbTmReg& g = regs[k];
        if (g.arity != ar) continue;
        bool fwd = true, rev = true;
        for (int m = 0; m < ar; ++m) { fwd &= g.t[m] == ty[m]; rev &= g.t[m] == ty[ar - 1 - m]; }
        if (!(fwd || rev)) continue;
        if (mode == 0) {
            atomicAdd(list_count + g.list, 1);
            if (lists[g.list].excl_observed) atomicAdd(nexcl_count, 1ull);
        } else {
            int o = atomicAdd(list_cursor + g.list, 1);
            for (int m = 0; m < ar; ++m) lists[g.list].tuples[(size_t)o * ar + m] = ids[m];
            if (lists[g.list].excl_observed) { unsigned long long q = atomicAdd(nexcl, 1ull); excl_pairs[q] = make_int2(min(ids[0], ids[ar - 1]), max(ids[0], ids[ar - 1])); }
        }
        return;  // first matching registration wins
    }
}
__global__ void k_ev_of_slot(int nev, const int* __restrict__ ev, const ClbCand* __restrict__ c, int* __restrict__ ev_of_slot, int set) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nev) return;
    const ClbCand& x = c[ev[e]];
    ev_of_slot[x.a] = set ? e : -1; ev_of_slot[x.b] = set ? e : -1;
}
template <int MODE>
__global__ void k_topo_tuples(int nev, const int* __restrict__ ev, const ClbCand* __restrict__ c, const ClbReactSpec* __restrict__ specs,
                              const ClbListDev* __restrict__ lists, const ClbTmReg* __restrict__ regs, int nreg, const int* __restrict__ adj,
                              const int* __restrict__ deg, const int* __restrict__ ev_of_slot, const int* __restrict__ wslot,
                              int* list_cursor, int* list_count, int2* excl_pairs, unsigned long long* nexcl,
                              unsigned long long* nexcl_count) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nev) return;
    const ClbCand x = c[ev[e]];
    const ClbReactSpec r = specs[x.r];
    if (r.is_virtual || !lists[r.list].tm_observed) return;
    const int a = x.a, b = x.b;
    for (int side = 0; side < 2; ++side) {
        int p = side ? b : a, o = side ? a : b;
        for (int k = 0; k < deg[p]; ++k) {
            int xx = adj[p * CLB_MAXDEG + k];
            if (xx == o) continue;
            if (find_event_of_bond(nev, ev, c, specs, ev_of_slot, xx, p) > e) continue;
            int t3[3] = {xx, p, o};
            tm_emit(MODE, 3, t3, wslot, regs, nreg, lists, list_cursor, list_count, excl_pairs, nexcl, nexcl_count);
            for (int k2 = 0; k2 < deg[xx]; ++k2) {
                int y = adj[xx * CLB_MAXDEG + k2];
                if (y == p || y == o) continue;
                if (find_event_of_bond(nev, ev, c, specs, ev_of_slot, y, xx) > e) continue;
                int t4[4] = {y, xx, p, o};
                tm_emit(MODE, 4, t4, wslot, regs, nreg, lists, list_cursor, list_count, excl_pairs, nexcl, nexcl_count);
            }
        }
    }
    for (int k = 0; k < deg[a]; ++k) {
        int xx = adj[a * CLB_MAXDEG + k];
        if (xx == b) continue;
        for (int k2 = 0; k2 < deg[b]; ++k2) {
            int y = adj[b * CLB_MAXDEG + k2];
            if (y == a || y == xx) continue;
            if (find_event_of_bond(nev, ev, c, specs, ev_of_slot, xx, a) > e || find_event_of_bond(nev, ev, c, specs, ev_of_slot, b, y) > e) continue;
            int t4[4] = {xx, a, b, y};
            tm_emit(MODE, 4, t4, wslot, regs, nreg, lists, list_cursor, list_count, excl_pairs, nexcl, nexcl_count);
        }
    }
}
// topology graph from the observed pair lists (TopologyManager.initialize_topology, :395-444)
__global__ void k_graph_add(long long nb, const int* __restrict__ tuples, int* __restrict__ adj, int* __restrict__ deg, ClbCtl* ctl) {
    long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= nb) return;
    int a = tuples[2 * k], b = tuples[2 * k + 1];
    int ia = atomicAdd(deg + a, 1), ib = atomicAdd(deg + b, 1);
    if (ia < CLB_MAXDEG && ib < CLB_MAXDEG) { adj[a * CLB_MAXDEG + ia] = b; adj[b * CLB_MAXDEG + ib] = a; }
    else atomicOr(&ctl->err, CLB_EF_DEGREE);
}
// molecule ids = smallest slot of the connected component: min-hooking + pointer jumping
__global__ void k_mol_init(int n, int* __restrict__ mol) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) mol[i] = i; }
__device__ __forceinline__ int mol_find(const int* mol, int i) { int p = mol[i]; while (p != i) { i = p; p = mol[i]; } return i; }
__global__ void k_mol_hook(long long nb, const int* __restrict__ pairs, int stride, int* __restrict__ mol, int* __restrict__ changed) {
    long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= nb) return;
    int ra = mol_find(mol, pairs[stride * k]), rb = mol_find(mol, pairs[stride * k + 1]);
    if (ra == rb) return;
    atomicMin(mol + max(ra, rb), min(ra, rb));
    *changed = 1;
}
__global__ void k_mol_compress(int n, int* __restrict__ mol) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) mol[i] = mol_find(mol, i); }
__global__ void k_event_pairs(int nev, const int* __restrict__ ev, const ClbCand* __restrict__ c, const ClbReactSpec* __restrict__ specs,
                              const ClbListDev* __restrict__ lists, int* __restrict__ out) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nev) return;
    const ClbCand& x = c[ev[e]];
    bool link = !specs[x.r].is_virtual && lists[specs[x.r].list].tm_observed;
    out[2 * e] = x.a; out[2 * e + 1] = link ? x.b : x.a;
}
__global__ void k_iota(int n, int* __restrict__ v) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) v[i] = i; }
__global__ void k_fill_u64(long long n, unsigned long long* __restrict__ v, unsigned long long x) { long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; if (i < n) v[i] = x; }
__global__ void k_fill_i32(long long n, int* __restrict__ v, int x) { long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; if (i < n) v[i] = x; }
__global__ void k_set_force_rebuild(ClbCtl* c) { c->force_rebuild = 1; }

// ------------------------------------------------------------------------------------------------------------------
// ATRPActivator (src/chemlab/reaction_post_process.py:393-424; [EXT] integrator/ATRPActivator.cpp, U23): every `interval`
// steps up to `num_particles` particles sitting on a reactive centre (type, state) are picked at random and switch between
// the dormant and the active state with the probability of meeting the right catalyst complex.
//   candidate : first centre k (registration order) with type == type_k and state == state_k
//   selection : counter-based key  u = Philox(seed, "ATRP", step, slot).x ; the num_particles smallest (u, slot) are taken
//   reaction  : w = Philox(...).y ;  w < k_deactivate*ratio_deactivator (flag "DA") or k_activate*ratio_activator (flag "A")
//   effect    : state += delta_state ; type / mass / charge from new_property
// The ratios used are those at the START of the pass; the host moves them by delta_catalyst*(n_act - n_deact)/num_particles.
struct ClbAtrpCenter { int type, state, deactivator, new_type, delta_state, pad; double new_mass, new_q, p; };
__global__ void k_atrp_scan(int i0, int i1, const int4* __restrict__ pos, const int* __restrict__ slot, const ClbAtrpCenter* __restrict__ cen, int ncen,
                            uint64_t seed, uint64_t step, unsigned long long* __restrict__ keys, unsigned long long* __restrict__ count, unsigned long long cap) {
    int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    int hit = -1, s = 0;
    if (i < i1) {
        const int w = pos[i].w;
        for (int k = 0; k < ncen && hit < 0; ++k) if (cen[k].type == pw_type(w) && cen[k].state == pw_state(w)) hit = k;
        s = slot[i];
    }
    const unsigned bal = __ballot_sync(0xffffffffu, hit >= 0);
    if (!bal) return;
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(count, (unsigned long long)__popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (hit >= 0) {
        uint32_t c[4] = {(uint32_t)s, 0u, (uint32_t)step, (uint32_t)(step >> 32)};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32) ^ CLB_STREAM_ATRP);
        const unsigned long long o = base + __popc(bal & ((1u << lane) - 1u));
        if (o < cap) keys[o] = ((unsigned long long)c[0] << 32) | (unsigned)s;
    }
}
// the first nsel keys (ascending) are the selected particles
__global__ void k_atrp_apply(int nsel, const unsigned long long* __restrict__ keys, const ClbAtrpCenter* __restrict__ cen, int ncen, uint64_t seed, uint64_t step,
                             const int* __restrict__ id2idx, int* wslot, int4* pos, ClbVel* vel, double* charge, unsigned long long* __restrict__ counts) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nsel) return;
    const int s = (int)(keys[k] & 0xffffffffull);
    int w = wslot[s];
    int hit = -1;
    for (int q = 0; q < ncen && hit < 0; ++q) if (cen[q].type == pw_type(w) && cen[q].state == pw_state(w)) hit = q;
    if (hit < 0) return;
    const ClbAtrpCenter C = cen[hit];
    uint32_t c[4] = {(uint32_t)s, 0u, (uint32_t)step, (uint32_t)(step >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32) ^ CLB_STREAM_ATRP);
    const double u = ((double)c[1] + 0.5) * (1.0 / 4294967296.0);
    if (!(u < C.p)) return;
    w = pw_pack(C.new_type >= 0 ? C.new_type : pw_type(w), pw_state(w) + C.delta_state);
    wslot[s] = w;
    const int idx = id2idx[s];
    if (idx >= 0) { pos[idx].w = w; if (C.new_mass > 0) vel[idx].w = C.new_mass; }
    if (C.new_q == C.new_q) charge[s] = C.new_q;
    atomicAdd(counts + (C.deactivator ? 1 : 0), 1ull);
}
