// engine_react.inl -- host orchestration of the reaction pass and the topology manager (included by engine.cu)

struct clb_engine::ReactDev {
    DevBuf<ClbReactSpec> specs;
    DevBuf<ClbChange> chg;
    DevBuf<ClbTmReg> regs;
    DevBuf<ClbListDev> lists;
    DevBuf<ClbCand> cands, cands_sorted;
    DevBuf<unsigned long long> ckey, ckey2, best1, best2, claim, counters, scalars, conn;
    DevBuf<int> cval, cval2, alive, surv, status, ev, erank, asA, inB, adj, deg, ev_of_slot, list_n, list_cnt, touched, evpairs, flag, iota;
    size_t candcap = 0;
    long long ncand_last = 0;
    bool slot_arrays_init = false;
};

void clb_engine::react_free() {
    if (!rd) return;
    ReactDev& R = *rd;
    R.specs.release(); R.chg.release(); R.regs.release(); R.lists.release(); R.cands.release(); R.cands_sorted.release();
    R.ckey.release(); R.ckey2.release(); R.best1.release(); R.best2.release(); R.claim.release(); R.counters.release(); R.scalars.release(); R.conn.release();
    R.cval.release(); R.cval2.release(); R.alive.release(); R.surv.release(); R.status.release(); R.ev.release(); R.erank.release();
    R.asA.release(); R.inB.release(); R.adj.release(); R.deg.release(); R.ev_of_slot.release(); R.list_n.release(); R.list_cnt.release();
    R.touched.release(); R.evpairs.release(); R.flag.release(); R.iota.release();
    delete rd; rd = nullptr;
}

extern "C" int clb_reaction_general(clb_engine* e, int enabled, int interval, int nearest, int max_per_interval) {
    if (!e || interval < 1) return e ? e->fail(CLB_ERR_ARG, "reaction interval must be >= 1") : CLB_ERR_ARG;
    if (interval != e->react_interval) e->react_dirty = true;       // p = rate * dt * interval is cached in upload_reactions
    e->react_on = enabled; e->react_interval = interval; e->react_nearest = nearest; e->react_max_per_interval = max_per_interval;
    return CLB_OK;
}
extern "C" int clb_add_reaction(clb_engine* e, const clb_reaction_spec* s, int* out) {
    if (!e || !s || !out) return CLB_ERR_ARG;
    if (e->reactions.size() >= CLB_MAX_REACTIONS) return e->fail(CLB_ERR_UNSUPPORTED, "too many reactions");
    if (s->list < 0 || s->list >= (int)e->lists.size() || e->lists[s->list].arity != 2) return e->fail(CLB_ERR_ARG, "reaction needs a pair list (fpl=)");
    // between rebuilds the Verlet rows are only complete up to rc: a larger reaction cutoff would make the candidate set depend
    // on the rebuild schedule (and on the rank count)
    if (s->cutoff > e->rc * (1 + 1e-12)) return e->fail(CLB_ERR_ARG, "reaction cutoff %g exceeds the Verlet cutoff %g", s->cutoff, e->rc);
    e->reactions.push_back(*s);
    e->react_counters.push_back(0);
    e->react_conn.emplace_back();
    e->react_restricted.push_back(0);
    e->ntypes = std::max(e->ntypes, std::max(s->type_1, s->type_2) + 1);
    *out = (int)e->reactions.size() - 1;
    e->react_dirty = true;
    return CLB_OK;
}
extern "C" int clb_reaction_set_rate(clb_engine* e, int r, double rate) {
    if (!e || r < 0 || r >= (int)e->reactions.size()) return CLB_ERR_ARG;
    e->reactions[r].rate = rate; e->react_dirty = true; return CLB_OK;
}
extern "C" int clb_reaction_set_active(clb_engine* e, int r, int a) {
    if (!e || r < 0 || r >= (int)e->reactions.size()) return CLB_ERR_ARG;
    e->reactions[r].active = a; e->react_dirty = true; return CLB_OK;
}
extern "C" int clb_reaction_define_connections(clb_engine* e, int r, int64_t n, const int64_t* pairs) {
    if (!e || r < 0 || r >= (int)e->reactions.size() || n < 0 || (n && !pairs)) return e ? e->fail(CLB_ERR_ARG, "clb_reaction_define_connections: bad argument") : CLB_ERR_ARG;
    for (int64_t k = 0; k < 2 * n; ++k)
        if (e->slot_of(pairs[k]) < 0) return e->fail(CLB_ERR_ARG, "connectivity map names unknown particle id %lld", (long long)pairs[k]);
    e->react_conn[r].assign(pairs, pairs + 2 * n);     // kept as ids: slots are resolved at upload (set_particles may follow)
    e->react_restricted[r] = 1;
    e->react_dirty = true;
    return CLB_OK;
}
extern "C" int clb_reaction_add_change(clb_engine* e, int reaction, int side, int nb_level, int old_type, int new_type, double new_mass,
                                       double new_q, int state_mode, int state_value) {
    if (!e || reaction < 0 || reaction >= (int)e->reactions.size() || side < 1 || side > 3 || nb_level < 0 || nb_level > 15) return e ? e->fail(CLB_ERR_ARG, "clb_reaction_add_change: bad argument") : CLB_ERR_ARG;
    if (new_type < 0 || new_type >= CLB_MAX_TYPES) return e->fail(CLB_ERR_ARG, "type out of range");
    if (e->changes.size() >= 4095) return e->fail(CLB_ERR_UNSUPPORTED, "too many change rules");
    HostChange c = {reaction, side, nb_level, old_type, new_type, state_mode, state_value, new_mass, new_q};
    e->changes.push_back(c);
    if (new_type + 1 > e->ntypes) { e->ntypes = new_type + 1; e->pots_dirty = true; }
    e->react_dirty = true;
    return CLB_OK;
}
extern "C" int clb_topology_observe(clb_engine* e, int list) {
    if (!e || list < 0 || list >= (int)e->lists.size() || e->lists[list].arity != 2) return e ? e->fail(CLB_ERR_ARG, "observe_tuple needs a pair list") : CLB_ERR_ARG;
    e->lists[list].tm_observed = 1; e->topo_dirty = true; e->react_dirty = true;
    return CLB_OK;
}
extern "C" int clb_topology_register_triplet(clb_engine* e, int list, int t1, int t2, int t3) {
    if (!e || list < 0 || list >= (int)e->lists.size() || e->lists[list].arity != 3) return e ? e->fail(CLB_ERR_ARG, "register_triplet needs a triple list") : CLB_ERR_ARG;
    HostTmReg r = {list, {t1, t2, t3, -1}}; e->tmregs.push_back(r); e->react_dirty = true; return CLB_OK;
}
extern "C" int clb_topology_register_quadruplet(clb_engine* e, int list, int t1, int t2, int t3, int t4) {
    if (!e || list < 0 || list >= (int)e->lists.size() || e->lists[list].arity != 4) return e ? e->fail(CLB_ERR_ARG, "register_quadruplet needs a quadruple list") : CLB_ERR_ARG;
    HostTmReg r = {list, {t1, t2, t3, t4}}; e->tmregs.push_back(r); e->react_dirty = true; return CLB_OK;
}
extern "C" int clb_topology_initialize(clb_engine* e) {
    if (!e) return CLB_ERR_ARG;
    e->topo_initialized = true; e->topo_dirty = true;
    return CLB_OK;
}

int clb_engine::upload_reactions() {
    clb_engine* e = this;
    if (!rd) rd = new ReactDev();
    std::vector<ClbReactSpec> hs(std::max<size_t>(reactions.size(), 1));
    std::vector<unsigned long long> hconn;       // connectivity maps of the restricted reactions, one sorted range each
    for (size_t k = 0; k < reactions.size(); ++k) {
        const clb_reaction_spec& s = reactions[k];
        ClbReactSpec d; memset(&d, 0, sizeof(d));
        d.type_1 = s.type_1; d.type_2 = s.type_2; d.delta_1 = s.delta_1; d.delta_2 = s.delta_2;
        d.min1 = s.min_state_1; d.max1 = s.max_state_1; d.min2 = s.min_state_2; d.max2 = s.max_state_2;
        d.cutoff2 = s.cutoff * s.cutoff; d.min_cutoff2 = s.min_cutoff * s.min_cutoff;
        d.p = s.rate * dt * react_interval;                      // U5
        d.list = s.list; d.intramolecular = s.intramolecular; d.intraresidual = s.intraresidual; d.is_virtual = s.is_virtual; d.active = s.active;
        d.conn_n = -1;
        if (react_restricted[k]) {
            const std::vector<int64_t>& cp = react_conn[k];
            std::vector<unsigned long long> keys(cp.size() / 2);
            for (size_t q = 0; q < keys.size(); ++q) {
                const int a = slot_of(cp[2 * q]), b = slot_of(cp[2 * q + 1]);
                if (a < 0 || b < 0) return fail(CLB_ERR_ARG, "connectivity map of reaction %d names an unknown particle id", (int)k);
                keys[q] = ((unsigned long long)(unsigned)std::min(a, b) << 32) | (unsigned)std::max(a, b);
            }
            std::sort(keys.begin(), keys.end());
            keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
            d.conn_off = (int)hconn.size(); d.conn_n = (int)keys.size();
            hconn.insert(hconn.end(), keys.begin(), keys.end());
        }
        hs[k] = d;
    }
    CK(rd->conn.ensure(std::max<size_t>(hconn.size(), 1)));
    if (!hconn.empty()) CK(cudaMemcpyAsync(rd->conn.p, hconn.data(), hconn.size() * 8, cudaMemcpyHostToDevice, stream));
    std::vector<ClbChange> hc(std::max<size_t>(changes.size(), 1));
    for (size_t k = 0; k < changes.size(); ++k) {
        const HostChange& c = changes[k];
        ClbChange d; memset(&d, 0, sizeof(d));
        d.reaction = c.reaction; d.side = c.side; d.nb_level = c.nb_level; d.old_type = c.old_type; d.new_type = c.new_type;
        d.state_mode = c.state_mode; d.state_value = c.state_value; d.new_mass = c.new_mass; d.new_q = c.new_q;
        hc[k] = d;
    }
    std::vector<ClbTmReg> hr(std::max<size_t>(tmregs.size(), 1));
    for (size_t k = 0; k < tmregs.size(); ++k) { hr[k].list = tmregs[k].list; hr[k].arity = lists[tmregs[k].list].arity; for (int m = 0; m < 4; ++m) hr[k].t[m] = tmregs[k].t[m]; }
    CK(rd->specs.ensure(hs.size())); CK(rd->chg.ensure(hc.size())); CK(rd->regs.ensure(hr.size()));
    CK(cudaMemcpyAsync(rd->specs.p, hs.data(), hs.size() * sizeof(ClbReactSpec), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(rd->chg.p, hc.data(), hc.size() * sizeof(ClbChange), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(rd->regs.p, hr.data(), hr.size() * sizeof(ClbTmReg), cudaMemcpyHostToDevice, stream));
    CK(rd->counters.ensure(CLB_MAX_REACTIONS)); CK(rd->scalars.ensure(16));
    if (!rd->slot_arrays_init) {
        CK(rd->best1.ensure(n)); CK(rd->best2.ensure(n)); CK(rd->claim.ensure(n)); CK(rd->asA.ensure(n)); CK(rd->inB.ensure(n)); CK(rd->ev_of_slot.ensure(n));
        k_fill_u64<<<ceil_div(n, 256), 256, 0, stream>>>(n, rd->claim.p, ~0ull);
        k_fill_i32<<<ceil_div(n, 256), 256, 0, stream>>>(n, rd->ev_of_slot.p, -1);
        CK(cudaMemsetAsync(rd->counters.p, 0, CLB_MAX_REACTIONS * 8, stream));
        rd->slot_arrays_init = true;
    }
    // every buffer of a reaction pass is allocated here, at set-up time: device allocation inside the first pass cost up to
    // 0.4 s on the B200 box (bench.py times one pass on its own)
    if (!reactions.empty()) {
        if (R_prealloc_n != n) {
            ReactDev& R = *rd;
            if (R.candcap == 0) R.candcap = std::max<size_t>(65536, (size_t)n / 4);
            const size_t cc = R.candcap;
            CK(R.cands.ensure(cc)); CK(R.cands_sorted.ensure(cc)); CK(R.ckey.ensure(cc)); CK(R.ckey2.ensure(cc)); CK(R.cval.ensure(cc)); CK(R.cval2.ensure(cc));
            CK(R.alive.ensure(cc)); CK(R.surv.ensure(cc)); CK(R.status.ensure(cc)); CK(R.ev.ensure(cc)); CK(R.erank.ensure(cc)); CK(R.iota.ensure(cc));
            CK(R.evpairs.ensure(2 * cc)); CK(R.flag.ensure(4)); CK(R.touched.ensure(cc * 8));
            CK(R.lists.ensure(CLB_MAX_LISTS)); CK(R.list_n.ensure(CLB_MAX_LISTS)); CK(R.list_cnt.ensure(CLB_MAX_LISTS));
            if (!R.adj.p) { CK(R.adj.ensure((size_t)n * CLB_MAXDEG)); CK(R.deg.ensure(n)); CK(cudaMemsetAsync(R.deg.p, 0, (size_t)n * 4, stream)); }
            size_t tb = 0, tb2 = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, tb, R.ckey.p, R.ckey2.p, R.cval.p, R.cval2.p, (int)cc, 0, 64, stream);
            cub::DeviceSelect::Flagged(nullptr, tb2, R.iota.p, R.alive.p, R.surv.p, (int*)R.scalars.p, (int)cc, stream);
            CK(cubtmp2.ensure(std::max(tb, tb2) + 256));
            CK(excl_pairs.ensure_keep((size_t)nexcl + cc + 1024, (size_t)nexcl, stream));
            R_prealloc_n = n;
        }
    }
    CK(cudaStreamSynchronize(stream));
    // lists that reactions append to (bond targets, registered angle/dihedral lists): capacity up front
    for (auto& r : reactions) TRY(list_reserve(r.list, lists[r.list].n + std::max<long long>(1, n / 8)));
    for (auto& g : tmregs) TRY(list_reserve(g.list, lists[g.list].n + std::max<long long>(1, n / 8)));
    react_dirty = false;
    return CLB_OK;
}
int clb_engine::upload_list_descs() {
    clb_engine* e = this;
    std::vector<ClbListDev> hl(CLB_MAX_LISTS);
    std::vector<int> hn(CLB_MAX_LISTS, 0);
    for (size_t l = 0; l < lists.size(); ++l) {
        hl[l].tuples = lists[l].d.p; hl[l].arity = lists[l].arity; hl[l].tm_observed = lists[l].tm_observed; hl[l].excl_observed = lists[l].excl_observed;
        hn[l] = (int)lists[l].n;
    }
    CK(rd->lists.ensure(CLB_MAX_LISTS)); CK(rd->list_n.ensure(CLB_MAX_LISTS)); CK(rd->list_cnt.ensure(CLB_MAX_LISTS));
    CK(cudaMemcpyAsync(rd->lists.p, hl.data(), hl.size() * sizeof(ClbListDev), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(rd->list_n.p, hn.data(), hn.size() * 4, cudaMemcpyHostToDevice, stream));
    CK(cudaStreamSynchronize(stream));
    return CLB_OK;
}

// bond graph + molecule ids from the observed pair lists
int clb_engine::build_topology() {
    clb_engine* e = this;
    if (!rd) rd = new ReactDev();
    CK(rd->adj.ensure((size_t)n * CLB_MAXDEG)); CK(rd->deg.ensure(n)); CK(rd->flag.ensure(4));
    CK(cudaMemsetAsync(rd->deg.p, 0, (size_t)n * 4, stream));
    k_mol_init<<<ceil_div(n, 256), 256, 0, stream>>>(n, mol.p);
    bool any = false;
    for (auto& l : lists) if (l.arity == 2 && l.tm_observed && l.n > 0) {
        k_graph_add<<<ceil_div(l.n, 256), 256, 0, stream>>>(l.n, l.d.p, rd->adj.p, rd->deg.p, d_ctl);
        any = true;
    }
    if (any) {
        for (int it = 0; it < 64; ++it) {
            CK(cudaMemsetAsync(rd->flag.p, 0, 4, stream));
            for (auto& l : lists) if (l.arity == 2 && l.tm_observed && l.n > 0)
                k_mol_hook<<<ceil_div(l.n, 256), 256, 0, stream>>>(l.n, l.d.p, 2, mol.p, rd->flag.p);
            int changed = 0;
            CK(cudaMemcpyAsync(&changed, rd->flag.p, 4, cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            if (!changed) break;
            if (it == 63) return fail(CLB_ERR_RANGE, "molecule ids did not converge in 64 hooking rounds");
        }
        k_mol_compress<<<ceil_div(n, 256), 256, 0, stream>>>(n, mol.p);
    }
    TRY(read_ctl());
    CK(cudaGetLastError());
    topo_dirty = false;
    return check_device_errors("topology");
}

int clb_engine::update_mixing() {
    bool changed = false;
    for (int a = 0; a < CLB_MAX_TYPES; ++a) for (int b = a; b < CLB_MAX_TYPES; ++b) {
        HostPairPot& p = pp[a][b];
        if (p.kind == 3 && p.conv_type >= 0) {
            int64_t c = 0;
            TRY(clb_count_type(this, p.conv_type, -1, &c));
            double x = (double)c / p.conv_total;
            if (x != p.mix) { p.mix = x; pp[b][a].mix = x; changed = true; }
        }
    }
    if (changed) pots_dirty = true;
    return CLB_OK;
}

// One ChemicalReaction::React pass at the current state; the caller has made `step` the number of
// completed steps (RNG key).
int clb_engine::react_pass(int64_t* events_out) {
    clb_engine* e = this;
    ClbTrace tr(stream, "react");
    if (events_out) *events_out = 0;
    if (reactions.empty()) return CLB_OK;
    bucket_begin(CLB_B_REACT);
    TRY(setup_sync());
    if (!lists_valid) TRY(rebuild());
    ReactDev& R = *rd;
    ++nreact_pass;
    // 1. candidate scan (retry on buffer overflow)
    // candidates are a small fraction of the Verlet pairs (reactive types, reaction cutoff < rc): start at n/4, regrow on overflow
    if (R.candcap == 0) { R.candcap = std::max<size_t>(65536, (size_t)n / 4); CK(R.cands.ensure(R.candcap)); }
    size_t smem = (size_t)tile_max * (sizeof(int4) + sizeof(int)) + 16;
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_react_scan, 256, smem);
    int gridsz = std::min(grid.nblocks, std::max(1, nb) * nsm);
    long long nc = 0;
    for (;;) {
        CK(cudaMemsetAsync(&d_ctl->ncand, 0, 8, stream));
        k_react_scan<<<gridsz, 256, smem, stream>>>(grid, geo, cell_start.p, pos.p, slot.p, nl_entries.p, nl_count.p, nl_cap, R.specs.p, (int)reactions.size(),
                                                            resid.p, mol.p, R.conn.p, seed, (uint64_t)step, R.cands.p, (unsigned long long)R.candcap, d_ctl);
        ++launches;
        TRY(read_ctl());
        nc = (long long)h_ctl->ncand;
        if ((size_t)nc <= R.candcap) break;
        R.candcap = (size_t)nc * 5 / 4 + 1024;
        CK(R.cands.ensure(R.candcap));
    }
    CK(cudaGetLastError());
    tr.mark("scan");
    if (nranks > 1) TRY(comm_gather_candidates(&nc));
    R.ncand_last = nc;
    int nev = 0;
    if (nc > 0) {
        // 2. canonical order (A, B, r)
        { size_t cc = R.candcap;
          CK(R.cands_sorted.ensure(cc)); CK(R.ckey.ensure(cc)); CK(R.ckey2.ensure(cc)); CK(R.cval.ensure(cc)); CK(R.cval2.ensure(cc));
          CK(R.alive.ensure(cc)); CK(R.surv.ensure(cc)); CK(R.status.ensure(cc)); CK(R.ev.ensure(cc)); CK(R.erank.ensure(cc)); CK(R.iota.ensure(cc)); }
        int g1 = ceil_div(nc, 256);
        k_cand_keys<<<g1, 256, 0, stream>>>((int)nc, R.cands.p, R.ckey.p, R.cval.p);
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, R.ckey.p, R.ckey2.p, R.cval.p, R.cval2.p, (int)R.candcap, 0, 64, stream);
        CK(cubtmp2.ensure(tb + 256));
        cub::DeviceRadixSort::SortPairs(cubtmp2.p, tb, R.ckey.p, R.ckey2.p, R.cval.p, R.cval2.p, (int)nc, 0, 64, stream);
        k_cand_gather<<<g1, 256, 0, stream>>>((int)nc, R.cval2.p, R.cands.p, R.cands_sorted.p, R.alive.p);
        tr.mark("sort");
        // 3. UniqueA then UniqueB (U7)
        for (int role = 0; role < 2; ++role) {
            k_uniq_reset<<<g1, 256, 0, stream>>>((int)nc, R.cands_sorted.p, R.best1.p, R.best2.p, R.asA.p, R.inB.p);
            k_uniq1<<<g1, 256, 0, stream>>>((int)nc, R.cands_sorted.p, R.alive.p, role, react_nearest, R.best1.p);
            k_uniq2<<<g1, 256, 0, stream>>>((int)nc, R.cands_sorted.p, R.alive.p, role, react_nearest, R.best1.p, R.best2.p);
            k_uniq3<<<g1, 256, 0, stream>>>((int)nc, R.cands_sorted.p, R.alive.p, role, react_nearest, R.best1.p, R.best2.p);
        }
        tr.mark("uniq");
        // 4. survivors in canonical order, U8 resolution, event list
        k_iota<<<g1, 256, 0, stream>>>((int)nc, R.iota.p);
        tb = 0;
        cub::DeviceSelect::Flagged(nullptr, tb, R.iota.p, R.alive.p, R.surv.p, (int*)R.scalars.p, (int)nc, stream);
        CK(cubtmp2.ensure(tb + 256));
        cub::DeviceSelect::Flagged(cubtmp2.p, tb, R.iota.p, R.alive.p, R.surv.p, (int*)R.scalars.p, (int)nc, stream);
        int nsurv = 0;
        CK(cudaMemcpyAsync(&nsurv, R.scalars.p, 4, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        if (nsurv > 0) {
            k_resolve<<<1, 1024, 0, stream>>>(nsurv, R.surv.p, R.cands_sorted.p, R.asA.p, R.inB.p, R.status.p, react_max_per_interval, R.ev.p, d_ctl);
            TRY(read_ctl());
            nev = h_ctl->nev;
        }
        launches += 14;
    }
    tr.mark("resolve");
    // 5..8 apply
    if (nev > 0) {
        TRY(upload_list_descs());
        int ge = ceil_div(nev, 128);
        // capacity for the new bonds
        std::vector<long long> add(lists.size(), 0);
        for (size_t l = 0; l < lists.size(); ++l) add[l] = 0;
        for (auto& s : reactions) if (!s.is_virtual) add[s.list] = nev;   // upper bound per bond list
        for (size_t l = 0; l < lists.size(); ++l) if (add[l]) TRY(list_reserve((int)l, lists[l].n + add[l]));
        CK(excl_pairs.ensure_keep((size_t)nexcl + nev + 1024, (size_t)nexcl, stream));
        TRY(upload_list_descs());
        k_apply_reactants<<<ge, 128, 0, stream>>>(nev, R.ev.p, R.cands_sorted.p, R.specs.p, R.chg.p, (int)changes.size(), id2idx.p, wslot.p, pos.p, vel.p, charge.p, R.counters.p);
        k_event_ranks<<<1, 1024, 0, stream>>>(nev, R.ev.p, R.cands_sorted.p, R.specs.p, (int)lists.size(), R.erank.p, R.list_n.p);
        unsigned long long hx = (unsigned long long)nexcl;
        CK(cudaMemcpyAsync(R.scalars.p + 2, &hx, 8, cudaMemcpyHostToDevice, stream));
        if (!topo_initialized && !R.adj.p) { CK(R.adj.ensure((size_t)n * CLB_MAXDEG)); CK(R.deg.ensure(n)); CK(cudaMemsetAsync(R.deg.p, 0, (size_t)n * 4, stream)); }
        k_apply_bonds<<<ge, 128, 0, stream>>>(nev, R.ev.p, R.cands_sorted.p, R.specs.p, R.lists.p, R.erank.p, R.adj.p, R.deg.p, excl_pairs.p, R.scalars.p + 2, d_ctl);
        tr.mark("apply_bonds");
        // molecule ids (only needed when some reaction forbids intramolecular bonds)
        bool need_mol = false;
        for (auto& s : reactions) need_mol |= !s.intramolecular;
        if (need_mol) {
            CK(R.evpairs.ensure(2 * (size_t)nev)); CK(R.flag.ensure(4));
            k_event_pairs<<<ge, 128, 0, stream>>>(nev, R.ev.p, R.cands_sorted.p, R.specs.p, R.lists.p, R.evpairs.p);
            for (int it = 0; it < 64; ++it) {
                CK(cudaMemsetAsync(R.flag.p, 0, 4, stream));
                k_mol_hook<<<ge, 128, 0, stream>>>(nev, R.evpairs.p, 2, mol.p, R.flag.p);
                int changed = 0;
                CK(cudaMemcpyAsync(&changed, R.flag.p, 4, cudaMemcpyDeviceToHost, stream));
                CK(cudaStreamSynchronize(stream));
                if (!changed) break;
                if (it == 63) return fail(CLB_ERR_RANGE, "molecule ids did not converge in 64 hooking rounds");
            }
            k_mol_compress<<<ceil_div(n, 256), 256, 0, stream>>>(n, mol.p);
        }
        // neighbour property changes
        bool any_nb = false;
        for (auto& c : changes) any_nb |= c.nb_level > 0;
        if (any_nb) {
            size_t touchcap = (size_t)nev * 2 * 64;
            CK(R.touched.ensure(touchcap));
            CK(cudaMemsetAsync(R.scalars.p + 4, 0, 8, stream));
            k_nb_claims<<<ceil_div(2 * nev, 128), 128, 0, stream>>>(nev, R.ev.p, R.cands_sorted.p, R.chg.p, (int)changes.size(), R.adj.p, R.deg.p, wslot.p,
                                                                   R.claim.p, R.touched.p, R.scalars.p + 4, (unsigned long long)touchcap, d_ctl);
            unsigned long long nt = 0;
            CK(cudaMemcpyAsync(&nt, R.scalars.p + 4, 8, cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            if (nt > touchcap) return fail(CLB_ERR_RANGE, "neighbour-change buffer overflow");
            if (nt) k_nb_apply<<<ceil_div((long long)nt, 128), 128, 0, stream>>>((int)nt, R.touched.p, R.claim.p, R.chg.p, id2idx.p, wslot.p, pos.p, vel.p, charge.p);
        }
        tr.mark("mol_nb");
        // read back the new bond counts
        std::vector<int> hn(CLB_MAX_LISTS);
        CK(cudaMemcpyAsync(hn.data(), R.list_n.p, CLB_MAX_LISTS * 4, cudaMemcpyDeviceToHost, stream));
        unsigned long long hx2 = 0;
        CK(cudaMemcpyAsync(&hx2, R.scalars.p + 2, 8, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        for (size_t l = 0; l < lists.size(); ++l) lists[l].n = hn[l];
        nexcl = (long long)hx2;
        // topology tuples: count, reserve, emit, sort the new segments
        if (!tmregs.empty()) {
            k_ev_of_slot<<<ge, 128, 0, stream>>>(nev, R.ev.p, R.cands_sorted.p, R.ev_of_slot.p, 1);
            CK(cudaMemsetAsync(R.list_cnt.p, 0, CLB_MAX_LISTS * 4, stream));
            CK(cudaMemsetAsync(R.scalars.p + 6, 0, 8, stream));
            k_topo_tuples<0><<<ge, 128, 0, stream>>>(nev, R.ev.p, R.cands_sorted.p, R.specs.p, R.lists.p, R.regs.p, (int)tmregs.size(), R.adj.p, R.deg.p, R.ev_of_slot.p,
                                                     wslot.p, R.list_n.p, R.list_cnt.p, excl_pairs.p, R.scalars.p + 2, R.scalars.p + 6);
            std::vector<int> hc(CLB_MAX_LISTS);
            unsigned long long nx = 0;
            CK(cudaMemcpyAsync(hc.data(), R.list_cnt.p, CLB_MAX_LISTS * 4, cudaMemcpyDeviceToHost, stream));
            CK(cudaMemcpyAsync(&nx, R.scalars.p + 6, 8, cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            bool anyt = false;
            for (size_t l = 0; l < lists.size(); ++l) if (hc[l]) { TRY(list_reserve((int)l, lists[l].n + hc[l])); anyt = true; }
            if (nx) CK(excl_pairs.ensure_keep((size_t)nexcl + nx + 1024, (size_t)nexcl, stream));
            if (anyt) {
                TRY(upload_list_descs());
                hx = (unsigned long long)nexcl;
                CK(cudaMemcpyAsync(R.scalars.p + 2, &hx, 8, cudaMemcpyHostToDevice, stream));
                k_topo_tuples<1><<<ge, 128, 0, stream>>>(nev, R.ev.p, R.cands_sorted.p, R.specs.p, R.lists.p, R.regs.p, (int)tmregs.size(), R.adj.p, R.deg.p, R.ev_of_slot.p,
                                                         wslot.p, R.list_n.p, R.list_cnt.p, excl_pairs.p, R.scalars.p + 2, R.scalars.p + 6);
                CK(cudaStreamSynchronize(stream));
                // deterministic order of the appended tuples: sort each new segment (device merge sort, lexicographic)
                for (size_t l = 0; l < lists.size(); ++l) if (hc[l]) {
                    int* seg = lists[l].d.p + (size_t)lists[l].n * lists[l].arity;
                    size_t tb2 = 0;
                    if (lists[l].arity == 3) {
                        cub::DeviceMergeSort::SortKeys(nullptr, tb2, (ClbTup<3>*)seg, hc[l], ClbTupLess<3>(), stream);
                        CK(cubtmp2.ensure(tb2 + 256));
                        cub::DeviceMergeSort::SortKeys(cubtmp2.p, tb2, (ClbTup<3>*)seg, hc[l], ClbTupLess<3>(), stream);
                    } else if (lists[l].arity == 4) {
                        cub::DeviceMergeSort::SortKeys(nullptr, tb2, (ClbTup<4>*)seg, hc[l], ClbTupLess<4>(), stream);
                        CK(cubtmp2.ensure(tb2 + 256));
                        cub::DeviceMergeSort::SortKeys(cubtmp2.p, tb2, (ClbTup<4>*)seg, hc[l], ClbTupLess<4>(), stream);
                    }
                    lists[l].n += hc[l];
                }
                nexcl += (long long)nx;
            }
            k_ev_of_slot<<<ge, 128, 0, stream>>>(nev, R.ev.p, R.cands_sorted.p, R.ev_of_slot.p, 0);
        }
        tr.mark("topo");
        // counters
        std::vector<unsigned long long> hcnt(CLB_MAX_REACTIONS);
        CK(cudaMemcpyAsync(hcnt.data(), R.counters.p, CLB_MAX_REACTIONS * 8, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        for (size_t k = 0; k < reactions.size(); ++k) react_counters[k] = (int64_t)hcnt[k];
        launches += 10;
        nreact_events += nev;
        terms_dirty = true; excl_dirty = true; lists_ptr_dirty = true;
        if (has_mixed) TRY(update_mixing());
        t3_dirty = true;      // type populations moved: the set of table windows kept in shared memory is re-ranked at the next rebuild
        tr.mark("tail");
        // U9: new exclusions / bonds take effect through a forced rebuild at the next resort check
        k_set_force_rebuild<<<1, 1, 0, stream>>>(d_ctl);
        pending_rebuild = true;
        forces_valid = false;
    }
    TRY(read_ctl());
    CK(cudaGetLastError());
    if (events_out) *events_out = nev;
    bucket_end(CLB_B_REACT);
    return check_device_errors("reaction pass");
}

extern "C" int clb_react_now(clb_engine* e, int64_t* events_out) {
    if (!e) return CLB_ERR_ARG;
    cudaSetDevice(e->device);
    int r = e->react_pass(events_out);
    if (r != CLB_OK) return r;
    if (e->pending_rebuild) { e->lists_valid = false; e->pending_rebuild = false; }
    return CLB_OK;
}
extern "C" int clb_reaction_counters(clb_engine* e, int cap, int64_t* out) {
    if (!e || !out) return CLB_ERR_ARG;
    for (int k = 0; k < cap && k < (int)e->react_counters.size(); ++k) out[k] = e->react_counters[k];
    return CLB_OK;
}
extern "C" int clb_get_last_candidates(clb_engine* e, int64_t cap, int64_t* rows, double* d2, int64_t* n_out) {
    if (!e) return CLB_ERR_ARG;
    cudaSetDevice(e->device);
    long long nc = e->rd ? e->rd->ncand_last : 0;
    if (n_out) *n_out = nc;
    if (nc == 0 || !rows) return CLB_OK;
    std::vector<ClbCand> h(nc);
    CK(cudaMemcpy(h.data(), e->rd->cands_sorted.p, nc * sizeof(ClbCand), cudaMemcpyDeviceToHost));
    for (long long k = 0; k < nc && k < cap; ++k) {
        rows[4 * k] = e->ids[h[k].a]; rows[4 * k + 1] = e->ids[h[k].b]; rows[4 * k + 2] = h[k].r; rows[4 * k + 3] = h[k].accepted;
        if (d2) d2[k] = h[k].d2;
    }
    return CLB_OK;
}

// ---- ATRPActivator (clb_react.cuh) ---------------------------------------------------------------------------------------
extern "C" int clb_atrp_configure(clb_engine* e, int num_particles, double ratio_activator, double ratio_deactivator, double delta_catalyst,
                                  double k_activate, double k_deactivate) {
    if (!e || num_particles < 0) return e ? e->fail(CLB_ERR_ARG, "clb_atrp_configure: bad argument") : CLB_ERR_ARG;
    e->atrp_num = num_particles; e->atrp_ratio_act = ratio_activator; e->atrp_ratio_deact = ratio_deactivator;
    e->atrp_delta = delta_catalyst; e->atrp_k_act = k_activate; e->atrp_k_deact = k_deactivate;
    e->atrp_centers.clear();
    return CLB_OK;
}
extern "C" int clb_atrp_add_center(clb_engine* e, int type, int state, int needs_deactivator, int new_type, double new_mass, double new_q, int delta_state) {
    if (!e || type < 0 || type >= CLB_MAX_TYPES || new_type >= CLB_MAX_TYPES) return e ? e->fail(CLB_ERR_ARG, "clb_atrp_add_center: type out of range") : CLB_ERR_ARG;
    clb_engine::HostAtrpCenter c = {type, state, needs_deactivator ? 1 : 0, new_type, delta_state, new_mass, new_q};
    e->atrp_centers.push_back(c);
    if (new_type + 1 > e->ntypes) { e->ntypes = new_type + 1; e->pots_dirty = true; }
    return CLB_OK;
}
// one pass at the current step; counts = {activated, deactivated} of this pass, ratios = {activator, deactivator} after it
extern "C" int clb_atrp_now(clb_engine* e, int64_t counts[2], double ratios[2]) {
    if (!e) return CLB_ERR_ARG;
    cudaSetDevice(e->device);
    if (counts) { counts[0] = 0; counts[1] = 0; }
    const int ncen = (int)e->atrp_centers.size();
    if (ncen == 0 || e->atrp_num <= 0 || e->n <= 0) { if (ratios) { ratios[0] = e->atrp_ratio_act; ratios[1] = e->atrp_ratio_deact; } return CLB_OK; }
    std::vector<ClbAtrpCenter> hc(ncen);
    for (int k = 0; k < ncen; ++k) {
        const auto& c = e->atrp_centers[k];
        ClbAtrpCenter d; memset(&d, 0, sizeof(d));
        d.type = c.type; d.state = c.state; d.deactivator = c.deactivator; d.new_type = c.new_type; d.delta_state = c.delta_state;
        d.new_mass = c.new_mass; d.new_q = c.new_q;
        d.p = c.deactivator ? e->atrp_k_deact * e->atrp_ratio_deact : e->atrp_k_act * e->atrp_ratio_act;
        hc[k] = d;
    }
    const size_t cap = (size_t)e->n;
    CK(e->atrp_cen.ensure(ncen * sizeof(ClbAtrpCenter))); CK(e->atrp_keys.ensure(cap)); CK(e->atrp_keys2.ensure(cap)); CK(e->atrp_scal.ensure(8));
    CK(cudaMemcpyAsync(e->atrp_cen.p, hc.data(), ncen * sizeof(ClbAtrpCenter), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemsetAsync(e->atrp_scal.p, 0, 64, e->stream));
    const int no = e->own1 - e->own0;
    if (no > 0) k_atrp_scan<<<ceil_div(no, 256), 256, 0, e->stream>>>(e->own0, e->own1, e->pos.p, e->slot.p, (const ClbAtrpCenter*)e->atrp_cen.p, ncen, e->seed,
                                                                       (uint64_t)e->step, e->atrp_keys.p, e->atrp_scal.p, (unsigned long long)cap);
    unsigned long long nc = 0;
    CK(cudaMemcpyAsync(&nc, e->atrp_scal.p, 8, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    unsigned long long* keys = e->atrp_keys.p;
    if (e->nranks > 1) {
        // every rank selects from the candidates of ALL ranks (same canonical order everywhere)
        void* all = nullptr; size_t tot = 0;
        TRY(e->comm_allgatherv(e->atrp_keys.p, (size_t)nc * 8, &all, &tot));
        nc = tot / 8;
        CK(e->atrp_keys.ensure(std::max<size_t>(nc, cap))); CK(e->atrp_keys2.ensure(std::max<size_t>(nc, cap)));
        if (nc) CK(cudaMemcpyAsync(e->atrp_keys.p, all, tot, cudaMemcpyDeviceToDevice, e->stream));
        keys = e->atrp_keys.p;
    }
    if (nc > 0) {
        size_t tb = 0;
        cub::DeviceRadixSort::SortKeys(nullptr, tb, keys, e->atrp_keys2.p, (int)nc, 0, 64, e->stream);
        CK(e->cubtmp2.ensure(tb + 256));
        cub::DeviceRadixSort::SortKeys(e->cubtmp2.p, tb, keys, e->atrp_keys2.p, (int)nc, 0, 64, e->stream);
        const int nsel = (int)std::min<unsigned long long>(nc, (unsigned long long)e->atrp_num);
        k_atrp_apply<<<ceil_div(nsel, 128), 128, 0, e->stream>>>(nsel, e->atrp_keys2.p, (const ClbAtrpCenter*)e->atrp_cen.p, ncen, e->seed, (uint64_t)e->step,
                                                                 e->id2idx.p, e->wslot.p, e->pos.p, e->vel.p, e->charge.p, e->atrp_scal.p + 2);
        unsigned long long hcnt[2] = {0, 0};
        CK(cudaMemcpyAsync(hcnt, e->atrp_scal.p + 2, 16, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        CK(cudaGetLastError());
        const long long n_act = (long long)hcnt[0], n_deact = (long long)hcnt[1];
        if (counts) { counts[0] = n_act; counts[1] = n_deact; }
        // an activation turns one activator complex into a deactivator complex and vice versa
        const double d = e->atrp_delta * (double)(n_act - n_deact) / (double)std::max(1, e->atrp_num);
        e->atrp_ratio_act = std::min(1.0, std::max(0.0, e->atrp_ratio_act - d));
        e->atrp_ratio_deact = std::min(1.0, std::max(0.0, e->atrp_ratio_deact + d));
        if (n_act + n_deact > 0) { e->t3_dirty = true; e->forces_valid = false; }      // type populations moved (table windows)
    }
    if (ratios) { ratios[0] = e->atrp_ratio_act; ratios[1] = e->atrp_ratio_deact; }
    return CLB_OK;
}
