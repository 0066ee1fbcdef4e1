/*
 * chemlab_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * fp64 CPU restatement of the reactive-MD hot path that chemlab drives inside the external
 * (un-vendored, un-pinned) modified ESPResSo++ `cgchemlab/espressopp`.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * PARITY UNPINNED: the reference ships no golden vector for forces, energies, pair sets or bond
 * lists and espressopp cannot be built here.  What IS pinned by
 * reference artefacts (table conversion, exclusion lists, parsers) is tested in tests/.  Every
 * [EXT] semantic below is our restatement of upstream ESPResSo++ and is listed as a named
 * choice in REFERENCE_UNVERIFIED.md (U1..U17).
 *
 * Citations `file:line` are relative to /root/reference; [EXT] names are the upstream
 * ESPResSo++ source files that hold the arithmetic.
 *
 * Particle indices are dense 0..n-1 (ascending caller id).  Everything is double precision,
 * single algorithmic path, optional OpenMP on the pair-force loop and list build.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_TYPES 64
#define ORC_MAXDEG 12

/* ------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11) -- counter-based RNG shared by oracle and engine so that
 * "identical per-pair uniform draws" (north_star) hold by construction.  Replaces esutil.RNG
 * (src/start_simulation.py:149) [EXT src/esutil/RNG.cpp].                                      */
static inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
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
void orc_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    philox4x32_10(c, key[0], key[1]);
    memcpy(out, c, sizeof(c));
}
#define STREAM_LANGEVIN 0x4c414e47u /* "LANG" */
#define STREAM_HEATUP 0x48454154u   /* "HEAT" */
#define STREAM_REACT 0x52454143u    /* "REAC" */
#define STREAM_PARTNER 0x50415254u  /* "PART" */
/* three uniforms in (0,1) for particle `idx` at step `step` */
static inline void draw3(uint64_t seed, uint32_t stream, uint64_t step, uint32_t idx, double u[3]) {
    uint32_t c[4] = {idx, 0u, (uint32_t)step, (uint32_t)(step >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32) ^ stream);
    for (int k = 0; k < 3; ++k) u[k] = ((double)c[k] + 0.5) * (1.0 / 4294967296.0);
}
/* per-pair uniform in [0,1) for (min idx, max idx, reaction) at step */
static inline void draw_pair(uint64_t seed, uint32_t stream, uint64_t step, uint32_t a, uint32_t b,
                             uint32_t r, uint32_t out[4]) {
    uint32_t c[4] = {a, b, (uint32_t)step, ((uint32_t)(step >> 32) << 8) ^ r};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32) ^ stream);
    memcpy(out, c, 4 * sizeof(uint32_t));
}

/* ------------------------------------------------------------------------------------------ */
/* Tabulated functions: linear (itype=1) [EXT interaction/InterpolationLinear.cpp],
 * Akima (itype=2) [EXT InterpolationAkima.cpp], cubic (itype=3, natural spline).
 * Table rows come from `.pot` files (tools/convert_gromacs2espp.py:84,107).                   */
typedef struct {
    int n, interp;
    double x0, dx;
    double *e, *f;
    double *ce, *cf; /* per-interval cubic coefficients [n-1][4] for interp 2/3 */
} orc_table;

static void akima_coeffs(int n, double h, const double *y, double *c) {
    /* standard Akima (1970) with linear end extrapolation of the slopes */
    double *m = (double *)malloc(sizeof(double) * (n + 3));
    double *mm = m + 2; /* mm[-2..n] */
    for (int i = 0; i < n - 1; ++i) mm[i] = (y[i + 1] - y[i]) / h;
    mm[-1] = 2 * mm[0] - mm[1];
    mm[-2] = 2 * mm[-1] - mm[0];
    mm[n - 1] = 2 * mm[n - 2] - mm[n - 3];
    mm[n] = 2 * mm[n - 1] - mm[n - 2];
    double *t = (double *)malloc(sizeof(double) * n);
    for (int i = 0; i < n; ++i) {
        double w1 = fabs(mm[i + 1] - mm[i]), w2 = fabs(mm[i - 1] - mm[i - 2]);
        if (w1 + w2 == 0.0) t[i] = 0.5 * (mm[i - 1] + mm[i]);
        else t[i] = (w1 * mm[i - 1] + w2 * mm[i]) / (w1 + w2);
    }
    for (int i = 0; i < n - 1; ++i) {
        c[4 * i + 0] = y[i];
        c[4 * i + 1] = t[i];
        c[4 * i + 2] = (3 * mm[i] - 2 * t[i] - t[i + 1]) / h;
        c[4 * i + 3] = (t[i] + t[i + 1] - 2 * mm[i]) / (h * h);
    }
    free(m); free(t);
}
static void cubic_coeffs(int n, double h, const double *y, double *c) {
    /* natural cubic spline, tridiagonal solve for second derivatives */
    double *y2 = (double *)calloc(n, sizeof(double)), *u = (double *)calloc(n, sizeof(double));
    for (int i = 1; i < n - 1; ++i) {
        double p = 0.5 * y2[i - 1] + 2.0;
        y2[i] = -0.5 / p;
        u[i] = (y[i + 1] - 2 * y[i] + y[i - 1]) / h;
        u[i] = (3.0 * u[i] / h - 0.5 * u[i - 1]) / p;
    }
    y2[n - 1] = 0;
    for (int k = n - 2; k >= 0; --k) y2[k] = y2[k] * y2[k + 1] + u[k];
    for (int i = 0; i < n - 1; ++i) {
        c[4 * i + 0] = y[i];
        c[4 * i + 1] = (y[i + 1] - y[i]) / h - h * (2 * y2[i] + y2[i + 1]) / 6.0;
        c[4 * i + 2] = 0.5 * y2[i];
        c[4 * i + 3] = (y2[i + 1] - y2[i]) / (6.0 * h);
    }
    free(y2); free(u);
}
/* returns 0 ok, 1 out of range (fatal in the reference) */
static inline int table_eval(const orc_table *t, double x, double *e, double *f) {
    double s = (x - t->x0) / t->dx;
    int idx = (int)floor(s);
    int bad = 0;
    if (idx < 0) { idx = 0; bad = 1; }
    if (idx > t->n - 2) { bad = (x > t->x0 + t->dx * (t->n - 1) * (1 + 1e-12)); idx = t->n - 2; }
    double xi = t->x0 + idx * t->dx;
    if (t->interp == 1) {
        double b = (x - xi) / t->dx, a = 1.0 - b;
        if (e) *e = a * t->e[idx] + b * t->e[idx + 1];
        if (f) *f = a * t->f[idx] + b * t->f[idx + 1];
    } else {
        double d = x - xi;
        const double *ce = t->ce + 4 * idx, *cf = t->cf + 4 * idx;
        if (e) *e = ce[0] + d * (ce[1] + d * (ce[2] + d * ce[3]));
        if (f) *f = cf[0] + d * (cf[1] + d * (cf[2] + d * cf[3]));
    }
    return bad;
}

/* ------------------------------------------------------------------------------------------ */
enum { NB_NONE = 0, NB_TAB = 1, NB_LJ = 2, NB_MIX = 3 };
typedef struct {
    int kind, inter; /* owning interaction handle */
    int tab1, tab2;
    double rc2, eps, sig, shift, mix;
    int conv_type; double conv_total;
} orc_pairpot;

enum { POT_HARMONIC = 1, POT_TAB = 2, POT_ANG_HARM = 3, POT_TAB_ANG = 4, POT_TAB_DIH = 5,
       POT_COSINE = 6, POT_FENE = 7, POT_DIH_HARM = 8, POT_FENE_LJ = 9, POT_LJ = 10 };
typedef struct { int kind, table; double p[6]; } orc_bpot;
typedef struct { int t[4]; orc_bpot pot; } orc_typed_pot;

typedef struct { int arity; int64_t n, cap; int *ids; } orc_list;
typedef struct {
    int list, typed, inter;
    orc_bpot pot;            /* typed == 0 */
    int ntp; orc_typed_pot *tp; /* typed == 1 */
} orc_bonded;

typedef struct {
    int type_1, type_2, delta_1, delta_2, min1, max1, min2, max2;
    double rate, cutoff, min_cutoff;
    int list, intramolecular, intraresidual, is_virtual, active;
    int64_t counter;
    /* RestrictReaction.define_connection (reaction_setup.py:74-75,115-126): nconn >= 0 -> only the pairs of the
     * connectivity map (sorted keys lower index << 32 | higher index) may react; -1 = plain Reaction */
    int64_t nconn; uint64_t *conn;
} orc_reaction;
typedef struct {
    int reaction, side, nb_level, old_type, new_type, state_mode, state_value;
    double new_mass, new_q;
} orc_change;
typedef struct { int list; int t[4]; } orc_tmreg;
typedef struct { int a, b, r; double d2; uint64_t rnd; int accepted; } orc_cand;

typedef struct orc_sim {
    int n, ntypes;
    double box[3], rc, skin, dt;
    double *x, *v, *f, *mass, *q;
    int *image, *type, *state, *resid, *mol;
    /* exclusions: sorted keys (min<<32|max) */
    uint64_t *excl; int64_t nexcl, capexcl;
    int nobs_excl, obs_excl[32];
    /* verlet list (half) */
    int *pairs; int64_t npairs, cappairs;
    double maxdist; double *xref; int criterion; int64_t nrebuild;
    /* tables, potentials */
    orc_table *tables; int ntables;
    orc_pairpot pp[ORC_MAX_TYPES][ORC_MAX_TYPES];
    int ninter; int inter_kind[256]; double inter_energy[256];
    orc_list *lists; int nlists;
    orc_bonded *bonded; int nbonded;
    /* integrator */
    int64_t step; int forces_valid; int lists_valid; int cont_ok;
    int lang_on; double kT, gamma; uint64_t seed; int lang_all; unsigned char lang_type[ORC_MAX_TYPES];
    double cap_force;   /* CapForce: <= 0 off */
    /* ATRPActivator */
    int atrp_num, atrp_ncen; double atrp_ratio[2], atrp_delta, atrp_k[2];
    struct { int type, state, deact, new_type, delta; double new_mass, new_q; } atrp_cen[16];
    /* reactions */
    int react_on, interval, nearest, max_per_interval;
    orc_reaction *reac; int nreac;
    orc_change *chg; int nchg;
    int ntm_obs, tm_obs[32]; orc_tmreg *tmreg; int ntmreg; int tm_init;
    int *deg, *adj; /* bond graph of observed lists, ORC_MAXDEG per particle */
    orc_cand *cands; int64_t ncands, capcands;
    int64_t last_events;
    int range_error;
    int nthreads;
    double *fbuf; size_t fbuf_cap;   /* per-thread force buffers (force_buffers) */
} orc_sim;

static inline double minimg(double d, double L) { return d - L * nearbyint(d / L); }

orc_sim *orc_create(int n, const double box[3], double rc, double skin, uint64_t seed) {
    orc_sim *s = (orc_sim *)calloc(1, sizeof(orc_sim));
    s->n = n; memcpy(s->box, box, 3 * sizeof(double)); s->rc = rc; s->skin = skin; s->seed = seed;
    s->x = calloc(3 * (size_t)n, 8); s->v = calloc(3 * (size_t)n, 8); s->f = calloc(3 * (size_t)n, 8);
    s->xref = calloc(3 * (size_t)n, 8);
    s->mass = calloc(n, 8); s->q = calloc(n, 8);
    s->image = calloc(3 * (size_t)n, 4); s->type = calloc(n, 4); s->state = calloc(n, 4);
    s->resid = calloc(n, 4); s->mol = calloc(n, 4);
    s->deg = calloc(n, 4); s->adj = calloc((size_t)n * ORC_MAXDEG, 4);
    s->dt = 0.001; s->criterion = 1; s->lang_all = 1; s->interval = 1; s->nearest = 1;
    s->nthreads = 1;
    return s;
}
void orc_destroy(orc_sim *s) {
    if (!s) return;
    free(s->x); free(s->v); free(s->f); free(s->xref); free(s->mass); free(s->q); free(s->image);
    free(s->type); free(s->state); free(s->resid); free(s->mol); free(s->deg); free(s->adj);
    free(s->excl); free(s->pairs); free(s->cands); free(s->fbuf);
    for (int i = 0; i < s->ntables; ++i) { free(s->tables[i].e); free(s->tables[i].f); free(s->tables[i].ce); free(s->tables[i].cf); }
    free(s->tables);
    for (int i = 0; i < s->nlists; ++i) free(s->lists[i].ids);
    free(s->lists);
    for (int i = 0; i < s->nbonded; ++i) free(s->bonded[i].tp);
    for (int i = 0; i < s->nreac; ++i) free(s->reac[i].conn);
    free(s->bonded); free(s->reac); free(s->chg); free(s->tmreg);
    free(s);
}
void orc_set_threads(orc_sim *s, int nt) {
    s->nthreads = nt < 1 ? 1 : nt;
#ifdef _OPENMP
    omp_set_num_threads(s->nthreads);
#endif
}
int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_option(orc_sim *s, const char *name, double v) {
    if (!strcmp(name, "resort_criterion")) s->criterion = (int)v;
    else if (!strcmp(name, "step")) s->step = (int64_t)v;   /* integrator.step: keys the thermostat and reaction draws */
}
/* storage.addParticles + decompose: fold into the box, keep image counters
 * (src/start_simulation.py:169-171) [EXT bc/OrthorhombicBC.cpp foldPosition] */
void orc_set_particles(orc_sim *s, const double *x, const double *v, const double *mass, const double *q,
                       const int *type, const int *state, const int *resid) {
    int nt = 0;
    for (int i = 0; i < s->n; ++i) {
        for (int d = 0; d < 3; ++d) {
            double L = s->box[d], xx = x[3 * i + d];
            double im = floor(xx / L);
            s->image[3 * i + d] = (int)im;
            xx -= im * L;
            if (xx >= L) { xx -= L; s->image[3 * i + d] += 1; }
            s->x[3 * i + d] = xx;
            s->v[3 * i + d] = v ? v[3 * i + d] : 0.0;
        }
        s->mass[i] = mass[i]; s->q[i] = q ? q[i] : 0.0; s->type[i] = type[i];
        s->state[i] = state ? state[i] : 0; s->resid[i] = resid ? resid[i] : 0;
        s->mol[i] = i;
        if (type[i] + 1 > nt) nt = type[i] + 1;
    }
    if (nt > s->ntypes) s->ntypes = nt;
    s->forces_valid = 0; s->lists_valid = 0; s->cont_ok = 0;
}
void orc_set_positions(orc_sim *s, const double *x) {
    /* folded into the box (the cell binning of orc_rebuild expects [0, L)); a coordinate outside the box moves the image counter
     * like fold() does, a folded coordinate leaves it alone */
    for (int i = 0; i < s->n; ++i) for (int d = 0; d < 3; ++d) {
        double L = s->box[d], xx = x[3 * i + d], im = floor(xx / L);
        xx -= im * L;
        if (xx >= L) { xx -= L; im += 1.0; }
        s->x[3 * i + d] = xx; s->image[3 * i + d] += (int)im;
    }
    s->forces_valid = 0; s->lists_valid = 0; s->cont_ok = 0;
}
void orc_set_velocities(orc_sim *s, const double *v) { memcpy(s->v, v, 3 * (size_t)s->n * 8); }
void orc_get(orc_sim *s, double *x, double *v, double *f, int *type, int *state, double *mass, int *image) {
    size_t n = s->n;
    if (x) memcpy(x, s->x, 24 * n);
    if (v) memcpy(v, s->v, 24 * n);
    if (f) memcpy(f, s->f, 24 * n);
    if (type) memcpy(type, s->type, 4 * n);
    if (state) memcpy(state, s->state, 4 * n);
    if (mass) memcpy(mass, s->mass, 8 * n);
    if (image) memcpy(image, s->image, 12 * n);
}
void orc_modify(orc_sim *s, int i, int field, const double *val) {
    switch (field) {
        case 0: s->type[i] = (int)val[0]; if (s->type[i] + 1 > s->ntypes) s->ntypes = s->type[i] + 1; break;
        case 1: s->state[i] = (int)val[0]; break;
        case 2: s->mass[i] = val[0]; break;
        case 3: s->q[i] = val[0]; break;
        case 4: s->resid[i] = (int)val[0]; break;
        case 5: for (int d = 0; d < 3; ++d) s->x[3 * i + d] = val[d]; s->lists_valid = 0; s->cont_ok = 0; break;
        case 6: for (int d = 0; d < 3; ++d) s->v[3 * i + d] = val[d]; break;
    }
    s->forces_valid = 0;
}

/* ---- exclusions: DynamicExcludeList (src/start_simulation.py:189) [EXT VerletList.cpp] ---- */
static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : (x > y);
}
static inline uint64_t pkey(int a, int b) {
    uint32_t lo = a < b ? a : b, hi = a < b ? b : a;
    return ((uint64_t)lo << 32) | hi;
}
static void excl_finalize(orc_sim *s) {
    qsort(s->excl, s->nexcl, 8, cmp_u64);
    int64_t m = 0;
    for (int64_t i = 0; i < s->nexcl; ++i)
        if (m == 0 || s->excl[i] != s->excl[m - 1]) s->excl[m++] = s->excl[i];
    s->nexcl = m;
}
static void excl_push(orc_sim *s, int a, int b) {
    if (s->nexcl == s->capexcl) {
        s->capexcl = s->capexcl ? 2 * s->capexcl : 1024;
        s->excl = realloc(s->excl, s->capexcl * 8);
    }
    s->excl[s->nexcl++] = pkey(a, b);
}
void orc_set_exclusions(orc_sim *s, int64_t n, const int *pairs) {
    s->nexcl = 0;
    for (int64_t i = 0; i < n; ++i) if (pairs[2 * i] != pairs[2 * i + 1]) excl_push(s, pairs[2 * i], pairs[2 * i + 1]);
    excl_finalize(s); s->lists_valid = 0; s->cont_ok = 0;
}
int64_t orc_get_exclusions(orc_sim *s, int64_t cap, int *out) {
    for (int64_t i = 0; i < s->nexcl && i < cap; ++i) { out[2 * i] = (int)(s->excl[i] >> 32); out[2 * i + 1] = (int)(s->excl[i] & 0xffffffffu); }
    return s->nexcl;
}
static inline int is_excluded(const orc_sim *s, int a, int b) {
    uint64_t k = pkey(a, b);
    int64_t lo = 0, hi = s->nexcl;
    while (lo < hi) { int64_t m = (lo + hi) >> 1; if (s->excl[m] < k) lo = m + 1; else hi = m; }
    return lo < s->nexcl && s->excl[lo] == k;
}
void orc_excl_observe(orc_sim *s, int list) { s->obs_excl[s->nobs_excl++] = list; }

/* ---- tables / potentials ---------------------------------------------------------------- */
int orc_add_table(orc_sim *s, int n, const double *x, const double *e, const double *f, int interp) {
    s->tables = realloc(s->tables, sizeof(orc_table) * (s->ntables + 1));
    orc_table *t = &s->tables[s->ntables];
    t->n = n; t->interp = interp; t->x0 = x[0]; t->dx = (x[n - 1] - x[0]) / (n - 1);
    t->e = malloc(8 * n); t->f = malloc(8 * n);
    memcpy(t->e, e, 8 * n); memcpy(t->f, f, 8 * n);
    t->ce = t->cf = NULL;
    if (interp != 1) {
        t->ce = malloc(32 * (n - 1)); t->cf = malloc(32 * (n - 1));
        if (interp == 2) { akima_coeffs(n, t->dx, e, t->ce); akima_coeffs(n, t->dx, f, t->cf); }
        else { cubic_coeffs(n, t->dx, e, t->ce); cubic_coeffs(n, t->dx, f, t->cf); }
    }
    return s->ntables++;
}
int orc_table_eval(orc_sim *s, int tab, double x, double *e, double *f) { return table_eval(&s->tables[tab], x, e, f); }

int orc_add_interaction(orc_sim *s, int kind) { s->inter_kind[s->ninter] = kind; return s->ninter++; }
static void grow_types(orc_sim *s, int t) { if (t + 1 > s->ntypes) s->ntypes = t + 1; }
void orc_nb_set_tab(orc_sim *s, int inter, int t1, int t2, int tab, double rc) {
    s->cont_ok = 0;
    orc_pairpot p; memset(&p, 0, sizeof(p)); p.kind = NB_TAB; p.inter = inter; p.tab1 = tab; p.rc2 = rc * rc;
    s->pp[t1][t2] = s->pp[t2][t1] = p; grow_types(s, t1); grow_types(s, t2); s->forces_valid = 0;
}
/* LennardJones(epsilon, sigma, cutoff, shift='auto'): gromacs_topology.py:715-721 [EXT LennardJones.hpp] */
void orc_nb_set_lj(orc_sim *s, int inter, int t1, int t2, double eps, double sig, double rc, int shift_auto) {
    s->cont_ok = 0;
    orc_pairpot p; memset(&p, 0, sizeof(p)); p.kind = NB_LJ; p.inter = inter; p.eps = eps; p.sig = sig; p.rc2 = rc * rc;
    double sr6 = pow(sig / rc, 6);
    p.shift = shift_auto ? 4 * eps * (sr6 * sr6 - sr6) : 0.0;
    s->pp[t1][t2] = s->pp[t2][t1] = p; grow_types(s, t1); grow_types(s, t2); s->forces_valid = 0;
}
void orc_nb_set_mixed(orc_sim *s, int inter, int t1, int t2, int tab1, int tab2, double mix, int conv_type,
                      double conv_total, double rc) {
    s->cont_ok = 0;
    orc_pairpot p; memset(&p, 0, sizeof(p)); p.kind = NB_MIX; p.inter = inter; p.tab1 = tab1; p.tab2 = tab2;
    p.mix = mix; p.conv_type = conv_type; p.conv_total = conv_total; p.rc2 = rc * rc;
    s->pp[t1][t2] = s->pp[t2][t1] = p; grow_types(s, t1); grow_types(s, t2); s->forces_valid = 0;
}
static void update_mixing(orc_sim *s) {
    /* analysis.ChemicalConversion(system, type, total) = N(type)/total (gromacs_topology.py:574-583) */
    for (int a = 0; a < s->ntypes; ++a) for (int b = 0; b < s->ntypes; ++b) {
        orc_pairpot *p = &s->pp[a][b];
        if (p->kind == NB_MIX && p->conv_type >= 0) {
            int64_t c = 0; for (int i = 0; i < s->n; ++i) c += (s->type[i] == p->conv_type);
            p->mix = (double)c / p->conv_total;
        }
    }
}

/* ---- tuple lists ------------------------------------------------------------------------- */
int orc_add_list(orc_sim *s, int arity) {
    s->lists = realloc(s->lists, sizeof(orc_list) * (s->nlists + 1));
    orc_list *l = &s->lists[s->nlists]; memset(l, 0, sizeof(*l)); l->arity = arity;
    return s->nlists++;
}
static void list_push(orc_sim *s, int list, const int *ids) {
    orc_list *l = &s->lists[list];
    if (l->n == l->cap) { l->cap = l->cap ? 2 * l->cap : 256; l->ids = realloc(l->ids, l->cap * l->arity * 4); }
    memcpy(l->ids + l->n * l->arity, ids, l->arity * 4); l->n++;
}
static int tm_observed(const orc_sim *s, int list) { for (int i = 0; i < s->ntm_obs; ++i) if (s->tm_obs[i] == list) return 1; return 0; }
static void graph_add(orc_sim *s, int a, int b) {
    for (int k = 0; k < s->deg[a]; ++k) if (s->adj[a * ORC_MAXDEG + k] == b) return;
    if (s->deg[a] < ORC_MAXDEG) s->adj[a * ORC_MAXDEG + s->deg[a]++] = b;
    if (s->deg[b] < ORC_MAXDEG) s->adj[b * ORC_MAXDEG + s->deg[b]++] = a;
}
static int mol_find(orc_sim *s, int i) { while (s->mol[i] != i) { s->mol[i] = s->mol[s->mol[i]]; i = s->mol[i]; } return i; }
static void mol_union(orc_sim *s, int a, int b) {
    int ra = mol_find(s, a), rb = mol_find(s, b);
    if (ra == rb) return;
    if (ra < rb) s->mol[rb] = ra; else s->mol[ra] = rb; /* representative = smallest index */
}
void orc_list_add(orc_sim *s, int list, int64_t n, const int *ids) {
    s->cont_ok = 0;
    int ar = s->lists[list].arity;
    for (int64_t i = 0; i < n; ++i) {
        list_push(s, list, ids + i * ar);
        if (ar == 2 && s->tm_init && tm_observed(s, list)) { graph_add(s, ids[2 * i], ids[2 * i + 1]); mol_union(s, ids[2 * i], ids[2 * i + 1]); }
    }
    s->forces_valid = 0;
}
int64_t orc_list_size(orc_sim *s, int list) { return s->lists[list].n; }
int64_t orc_list_get(orc_sim *s, int list, int64_t cap, int *out) {
    orc_list *l = &s->lists[list];
    int64_t m = l->n < cap ? l->n : cap;
    memcpy(out, l->ids, m * l->arity * 4);
    return l->n;
}
int orc_add_bonded(orc_sim *s, int list, int typed) {
    s->bonded = realloc(s->bonded, sizeof(orc_bonded) * (s->nbonded + 1));
    orc_bonded *b = &s->bonded[s->nbonded]; memset(b, 0, sizeof(*b));
    b->list = list; b->typed = typed; b->inter = orc_add_interaction(s, 10 + s->lists[list].arity);
    s->nbonded++;
    return b->inter;
}
static orc_bonded *bonded_by_inter(orc_sim *s, int inter) { for (int i = 0; i < s->nbonded; ++i) if (s->bonded[i].inter == inter) return &s->bonded[i]; return NULL; }
void orc_bonded_set_potential(orc_sim *s, int inter, int t1, int t2, int t3, int t4, int kind, const double *params,
                              int np, int table) {
    s->cont_ok = 0;
    orc_bonded *b = bonded_by_inter(s, inter);
    orc_bpot p; memset(&p, 0, sizeof(p)); p.kind = kind; p.table = table;
    for (int i = 0; i < np && i < 6; ++i) p.p[i] = params[i];
    if (!b->typed) { b->pot = p; }
    else {
        b->tp = realloc(b->tp, sizeof(orc_typed_pot) * (b->ntp + 1));
        b->tp[b->ntp].t[0] = t1; b->tp[b->ntp].t[1] = t2; b->tp[b->ntp].t[2] = t3; b->tp[b->ntp].t[3] = t4;
        b->tp[b->ntp].pot = p; b->ntp++;
    }
    s->forces_valid = 0;
}
/* FixedXListTypes lookup: the type tuple or its reverse (key canonicalisation gromacs_topology.py:405-408,426-429) */
static const orc_bpot *typed_lookup(const orc_bonded *b, int ar, const int *ty) {
    for (int k = 0; k < b->ntp; ++k) {
        const int *t = b->tp[k].t;
        int fwd = 1, rev = 1;
        for (int m = 0; m < ar; ++m) { if (t[m] != ty[m]) fwd = 0; if (t[m] != ty[ar - 1 - m]) rev = 0; }
        if (fwd || rev) return &b->tp[k].pot;
    }
    return NULL;
}

/* ---- Verlet list: VerletList(system, cutoff, exclusionlist) rebuild (src/start_simulation.py:193-197)
 * [EXT VerletList.cpp checkPair: include iff r^2 <= (rc+skin)^2 and pair not excluded (U1);
 *  iterator/CellListAllPairsIterator.cpp: in-cell i<j plus 13 half-shell neighbour cells]. ---- */
static void pairs_reserve(orc_sim *s, int64_t n) {
    if (n > s->cappairs) { s->cappairs = n + n / 2 + 1024; s->pairs = realloc(s->pairs, s->cappairs * 8); }
}
static inline double dist2(const orc_sim *s, int i, int j, double d[3]) {
    double r2 = 0;
    for (int k = 0; k < 3; ++k) { d[k] = minimg(s->x[3 * i + k] - s->x[3 * j + k], s->box[k]); r2 += d[k] * d[k]; }
    return r2;
}
static int cmp_pair(const void *a, const void *b) {
    const int *x = a, *y = b;
    if (x[0] != y[0]) return x[0] < y[0] ? -1 : 1;
    return x[1] < y[1] ? -1 : (x[1] > y[1]);
}
int64_t orc_pairs_brute(orc_sim *s, int64_t cap, int *out) {
    double rl2 = (s->rc + s->skin) * (s->rc + s->skin), d[3];
    int64_t m = 0;
    for (int i = 0; i < s->n; ++i) for (int j = i + 1; j < s->n; ++j)
        if (dist2(s, i, j, d) <= rl2 && !is_excluded(s, i, j)) { if (m < cap) { out[2 * m] = i; out[2 * m + 1] = j; } ++m; }
    return m;
}
void orc_rebuild(orc_sim *s) {
    double rl = s->rc + s->skin, rl2 = rl * rl;
    int nc[3]; double cs[3];
    for (int d = 0; d < 3; ++d) { nc[d] = (int)floor(s->box[d] / rl); if (nc[d] < 1) nc[d] = 1; cs[d] = s->box[d] / nc[d]; }
    s->npairs = 0;
    if (nc[0] < 3 || nc[1] < 3 || nc[2] < 3) {
        int64_t m = orc_pairs_brute(s, 0, NULL);
        pairs_reserve(s, m); s->npairs = orc_pairs_brute(s, m, s->pairs);
    } else {
        int ncell = nc[0] * nc[1] * nc[2];
        int *head = malloc(4 * (ncell + 1)), *cell = malloc(4 * s->n), *order = malloc(4 * s->n);
        memset(head, 0, 4 * (ncell + 1));
        for (int i = 0; i < s->n; ++i) {
            int c[3];
            for (int d = 0; d < 3; ++d) { c[d] = (int)(s->x[3 * i + d] / cs[d]); if (c[d] >= nc[d]) c[d] = nc[d] - 1; if (c[d] < 0) c[d] = 0; }
            cell[i] = (c[2] * nc[1] + c[1]) * nc[0] + c[0];
            head[cell[i] + 1]++;
        }
        for (int c = 0; c < ncell; ++c) head[c + 1] += head[c];
        int *fill = malloc(4 * ncell); memcpy(fill, head, 4 * ncell);
        for (int i = 0; i < s->n; ++i) order[fill[cell[i]]++] = i;
        free(fill);
        /* 13 forward neighbours */
        int off[13][3], no = 0;
        for (int dz = -1; dz <= 1; ++dz) for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) {
            if (dz > 0 || (dz == 0 && dy > 0) || (dz == 0 && dy == 0 && dx > 0)) { off[no][0] = dx; off[no][1] = dy; off[no][2] = dz; ++no; }
        }
        int nth = s->nthreads;
        int64_t *cnt = calloc(nth + 1, 8); int **buf = calloc(nth, sizeof(int *)); int64_t *bcap = calloc(nth, 8);
#pragma omp parallel num_threads(nth)
        {
            int tid = 0;
#ifdef _OPENMP
            tid = omp_get_thread_num();
#endif
            int64_t m = 0, cp = 1 << 16; int *b = malloc(cp * 8); double d[3];
#pragma omp for schedule(dynamic, 64)
            for (int c = 0; c < ncell; ++c) {
                int cx = c % nc[0], cy = (c / nc[0]) % nc[1], cz = c / (nc[0] * nc[1]);
                for (int a = head[c]; a < head[c + 1]; ++a) {
                    int i = order[a];
                    for (int bb = a + 1; bb < head[c + 1]; ++bb) {
                        int j = order[bb];
                        if (dist2(s, i, j, d) <= rl2 && !is_excluded(s, i, j)) {
                            if (m == cp) { cp *= 2; b = realloc(b, cp * 8); }
                            b[2 * m] = i < j ? i : j; b[2 * m + 1] = i < j ? j : i; ++m;
                        }
                    }
                    for (int o = 0; o < 13; ++o) {
                        int ex = (cx + off[o][0] + nc[0]) % nc[0], ey = (cy + off[o][1] + nc[1]) % nc[1], ez = (cz + off[o][2] + nc[2]) % nc[2];
                        int c2 = (ez * nc[1] + ey) * nc[0] + ex;
                        for (int bb = head[c2]; bb < head[c2 + 1]; ++bb) {
                            int j = order[bb];
                            if (dist2(s, i, j, d) <= rl2 && !is_excluded(s, i, j)) {
                                if (m == cp) { cp *= 2; b = realloc(b, cp * 8); }
                                b[2 * m] = i < j ? i : j; b[2 * m + 1] = i < j ? j : i; ++m;
                            }
                        }
                    }
                }
            }
            cnt[tid + 1] = m; buf[tid] = b; bcap[tid] = cp;
        }
        for (int t = 0; t < nth; ++t) cnt[t + 1] += cnt[t];
        pairs_reserve(s, cnt[nth]);
        for (int t = 0; t < nth; ++t) { if (buf[t]) { memcpy(s->pairs + 2 * cnt[t], buf[t], (cnt[t + 1] - cnt[t]) * 8); free(buf[t]); } }
        s->npairs = cnt[nth];
        free(cnt); free(buf); free(bcap); free(head); free(cell); free(order);
    }
    s->maxdist = 0; memcpy(s->xref, s->x, 24 * (size_t)s->n);
    s->lists_valid = 1; s->nrebuild++;
}
/* canonical (sorted) pair set for parity */
int64_t orc_get_pairs(orc_sim *s, int64_t cap, int *out) {
    if (!s->lists_valid) orc_rebuild(s);
    int64_t m = s->npairs < cap ? s->npairs : cap;
    memcpy(out, s->pairs, m * 8);
    qsort(out, m, 8, cmp_pair);
    return s->npairs;
}

/* the same pair set in list order (unsorted; rows are (min id, max id)): large systems are sorted by the caller (numpy) */
int64_t orc_get_pairs_raw(orc_sim *s, int64_t cap, int *out) {
    if (!s->lists_valid) orc_rebuild(s);
    int64_t m = s->npairs < cap ? s->npairs : cap;
    if (out) memcpy(out, s->pairs, m * 8);
    return s->npairs;
}

/* ---- forces ------------------------------------------------------------------------------ */
/* VerletListInteractionTemplate<Potential>::addForces [EXT]:
 * Tabulated (gromacs_topology.py:696-707), LennardJones (:715-721), MixedTabulated (:757-790). */
static inline int pair_eval(orc_sim *s, const orc_pairpot *p, double r2, double *fr, double *e) {
    /* returns F/r (so that f_i += fr * d) and energy */
    int bad = 0;
    if (p->kind == NB_LJ) {
        double f2 = 1.0 / r2, f6 = f2 * f2 * f2;
        double s6 = pow(p->sig, 6), s12 = s6 * s6;
        *fr = f6 * (48 * p->eps * s12 * f6 - 24 * p->eps * s6) * f2;
        *e = 4 * p->eps * (s12 * f6 * f6 - s6 * f6) - p->shift;
    } else if (p->kind == NB_TAB) {
        double r = sqrt(r2), F, E;
        bad = table_eval(&s->tables[p->tab1], r, &E, &F);
        *fr = F / r; *e = E;
    } else {
        double r = sqrt(r2), F1, E1, F2, E2;
        bad = table_eval(&s->tables[p->tab1], r, &E1, &F1);
        bad |= table_eval(&s->tables[p->tab2], r, &E2, &F2);
        *fr = (p->mix * F1 + (1 - p->mix) * F2) / r; *e = p->mix * E1 + (1 - p->mix) * E2;
    }
    return bad;
}
/* Threading of the force evaluation (the reference parallelises by MPI domains; this restatement by OpenMP threads):
 * every thread accumulates into its own force buffer, the buffers are summed in thread order afterwards, so the result is
 * reproducible for a given thread count.  One thread writes straight into s->f (the order of the serial code). */
static double *force_buffers(orc_sim *s) {
    int nth = s->nthreads;
    if (nth <= 1) return NULL;
    size_t n3 = 3 * (size_t)s->n, need = n3 * (size_t)nth;
    if (s->fbuf_cap < need) { free(s->fbuf); s->fbuf = malloc(need * 8); s->fbuf_cap = need; }
#pragma omp parallel num_threads(nth)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        memset(s->fbuf + n3 * tid, 0, n3 * 8);          /* first touch by the owning thread */
    }
    return s->fbuf;
}
static void reduce_force_buffers(orc_sim *s, const double *fbuf) {
    int nth = s->nthreads;
    if (nth <= 1 || !fbuf) return;
    size_t n3 = 3 * (size_t)s->n;
#pragma omp parallel for num_threads(nth) schedule(static)
    for (int64_t i = 0; i < (int64_t)n3; ++i) { double a = 0; for (int t = 0; t < nth; ++t) a += fbuf[n3 * t + i]; s->f[i] += a; }
}
static void nonbonded_forces(orc_sim *s, double *fbuf) {
    int nth = s->nthreads;
    size_t n3 = 3 * (size_t)s->n;
    double *ebuf = calloc((size_t)nth * 256, 8);
    int bad = 0;
#pragma omp parallel num_threads(nth) reduction(| : bad)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        double *f = nth > 1 ? fbuf + n3 * tid : s->f;
        double *en = ebuf + 256 * tid;
#pragma omp for schedule(static)
        for (int64_t k = 0; k < s->npairs; ++k) {
            int i = s->pairs[2 * k], j = s->pairs[2 * k + 1];
            const orc_pairpot *p = &s->pp[s->type[i]][s->type[j]];
            if (p->kind == NB_NONE) continue;
            double d[3], r2 = dist2(s, i, j, d);
            if (r2 > p->rc2) continue; /* U2 */
            double fr, e;
            bad |= pair_eval(s, p, r2, &fr, &e);
            for (int c = 0; c < 3; ++c) { f[3 * i + c] += fr * d[c]; f[3 * j + c] -= fr * d[c]; }
            en[p->inter] += e;
        }
    }
    for (int t = 0; t < nth; ++t) for (int k = 0; k < s->ninter; ++k) s->inter_energy[k] += ebuf[256 * t + k];
    free(ebuf);
    if (bad) s->range_error = 1;
}
/* bonded: FixedPairListInteractionTemplate / FixedTripleList... / FixedQuadrupleList... [EXT]
 * Harmonic U=K(r-r0)^2 (gromacs_topology.py:918), AngularHarmonic U=K(th-th0)^2 (:1073),
 * tables in r / theta / phi (:919-925,1074-1080,1192-1198). U15 conventions. */
static int bond_eval(orc_sim *s, const orc_bpot *p, double r, double *F, double *E) {
    switch (p->kind) {
        case POT_HARMONIC: *E = p->p[0] * (r - p->p[1]) * (r - p->p[1]); *F = -2 * p->p[0] * (r - p->p[1]); return 0;
        case POT_FENE: { double x = (r - p->p[1]) / p->p[2]; *E = -0.5 * p->p[0] * p->p[2] * p->p[2] * log(1 - x * x);
                         *F = -p->p[0] * (r - p->p[1]) / (1 - x * x); return 0; }
        /* FENE + LJ bond, [ bondtypes ] func 9 (doc/topology.rst:72-79, gromacs_topology.py:935-944): p = {K, r0, rMax, sigma, epsilon};
         * U = -K rMax^2/2 ln(1 - ((r-r0)/rMax)^2) + 4 eps [(sigma/r)^12 - (sigma/r)^6], no cutoff on the LJ part (formula as documented) */
        /* 1-4 [ pairs ]: FixedPairList[Types]LennardJones with LennardJones(epsilon, sigma, cutoff), shift 'auto'
         * (gromacs_topology.py:1314-1411); p = {epsilon, sigma, cutoff, shift}; no force and no energy beyond the cutoff */
        case POT_LJ: { if (r > p->p[2]) { *E = 0; *F = 0; return 0; }
                       double sr2 = p->p[1] * p->p[1] / (r * r), sr6 = sr2 * sr2 * sr2;
                       *E = 4 * p->p[0] * (sr6 * sr6 - sr6) - p->p[3]; *F = 24 * p->p[0] * (2 * sr6 * sr6 - sr6) / r; return 0; }
        case POT_FENE_LJ: { double x = (r - p->p[1]) / p->p[2], sr2 = p->p[3] * p->p[3] / (r * r), sr6 = sr2 * sr2 * sr2;
                            *E = -0.5 * p->p[0] * p->p[2] * p->p[2] * log(1 - x * x) + 4 * p->p[4] * (sr6 * sr6 - sr6);
                            *F = -p->p[0] * (r - p->p[1]) / (1 - x * x) + 24 * p->p[4] * (2 * sr6 * sr6 - sr6) / r; return 0; }
        case POT_TAB: return table_eval(&s->tables[p->table], r, E, F);
    }
    *E = *F = 0; return 0;
}
static int angle_eval(orc_sim *s, const orc_bpot *p, double th, double *F, double *E) {
    switch (p->kind) {
        case POT_ANG_HARM: *E = p->p[0] * (th - p->p[1]) * (th - p->p[1]); *F = -2 * p->p[0] * (th - p->p[1]); return 0;
        case POT_COSINE: *E = p->p[0] * (1 + cos(th - p->p[1])); *F = p->p[0] * sin(th - p->p[1]); return 0;
        case POT_TAB_ANG: return table_eval(&s->tables[p->table], th, E, F);
    }
    *E = *F = 0; return 0;
}
static int dih_eval(orc_sim *s, const orc_bpot *p, double phi, double *F, double *E) {
    switch (p->kind) {
        case POT_DIH_HARM: { double d = phi - p->p[1]; d -= 2 * M_PI * nearbyint(d / (2 * M_PI)); *E = p->p[0] * d * d; *F = -2 * p->p[0] * d; return 0; }
        case POT_TAB_DIH: return table_eval(&s->tables[p->table], phi, E, F);
    }
    *E = *F = 0; return 0;
}
static void cross(const double a[3], const double b[3], double c[3]) {
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
static double dot(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void bonded_forces(orc_sim *s, double *fbuf) {
    const int nth = s->nthreads;
    const size_t n3 = 3 * (size_t)s->n;
    for (int bi = 0; bi < s->nbonded; ++bi) {
        orc_bonded *b = &s->bonded[bi];
        orc_list *l = &s->lists[b->list];
        double etot = 0;
        int bad = 0;
        double *et = calloc((size_t)nth, 8);
#pragma omp parallel num_threads(nth) reduction(| : bad)
        {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        double *f = nth > 1 ? fbuf + n3 * tid : s->f;
        double esum = 0;
#pragma omp for schedule(static)
        for (int64_t k = 0; k < l->n; ++k) {
            const int *id = l->ids + k * l->arity;
            const orc_bpot *p = &b->pot;
            if (b->typed) {
                int ty[4]; for (int m = 0; m < l->arity; ++m) ty[m] = s->type[id[m]];
                p = typed_lookup(b, l->arity, ty);
                if (!p) continue; /* no potential registered for this type tuple */
            }
            double F, E;
            if (l->arity == 2) {
                double d[3], r2 = dist2(s, id[0], id[1], d), r = sqrt(r2);
                if (bond_eval(s, p, r, &F, &E)) bad = 1;
                for (int c = 0; c < 3; ++c) { f[3 * id[0] + c] += F / r * d[c]; f[3 * id[1] + c] -= F / r * d[c]; }
            } else if (l->arity == 3) {
                double d1[3], d2[3];
                double r1 = sqrt(dist2(s, id[0], id[1], d1)), r2 = sqrt(dist2(s, id[2], id[1], d2));
                double c = dot(d1, d2) / (r1 * r2); if (c > 1) c = 1; if (c < -1) c = -1;
                double th = acos(c), sn = sqrt(1 - c * c); if (sn < 1e-9) sn = 1e-9;
                if (angle_eval(s, p, th, &F, &E)) bad = 1;
                /* F = -dU/dtheta ; dtheta/dx_i = -(1/sin) * (d2/(r1 r2) - c d1/r1^2) */
                for (int k3 = 0; k3 < 3; ++k3) {
                    double g1 = -(d2[k3] / (r1 * r2) - c * d1[k3] / (r1 * r1)) / sn;
                    double g3 = -(d1[k3] / (r1 * r2) - c * d2[k3] / (r2 * r2)) / sn;
                    f[3 * id[0] + k3] += F * g1; f[3 * id[2] + k3] += F * g3; f[3 * id[1] + k3] -= F * (g1 + g3);
                }
            } else {
                /* GROMACS/IUPAC convention (U15): r_ij = xi-xj, r_kj = xk-xj, r_kl = xk-xl */
                double rij[3], rkj[3], rkl[3], m[3], nn[3];
                dist2(s, id[0], id[1], rij); dist2(s, id[2], id[1], rkj); dist2(s, id[2], id[3], rkl);
                cross(rij, rkj, m); cross(rkj, rkl, nn);
                double m2 = dot(m, m), n2 = dot(nn, nn), rkj2 = dot(rkj, rkj), nrkj = sqrt(rkj2);
                double cphi = dot(m, nn) / sqrt(m2 * n2); if (cphi > 1) cphi = 1; if (cphi < -1) cphi = -1;
                double phi = acos(cphi); if (dot(rij, nn) < 0) phi = -phi;
                if (dih_eval(s, p, phi, &F, &E)) bad = 1;
                /* GROMACS do_dih_fup with ddphi = dU/dphi = -F (Bekker et al. 1995):
                 * f_i = -ddphi nrkj/|m|^2 m ; f_l = +ddphi nrkj/|n|^2 n */
                double fi[3], fl[3], pp = dot(rij, rkj) / rkj2, qq = dot(rkl, rkj) / rkj2;
                for (int c = 0; c < 3; ++c) { fi[c] = F * (nrkj / m2) * m[c]; fl[c] = -F * (nrkj / n2) * nn[c]; }
                for (int c = 0; c < 3; ++c) {
                    double sv = pp * fi[c] - qq * fl[c];
                    double fj = -fi[c] + sv, fk = -fl[c] - sv;
                    f[3 * id[0] + c] += fi[c]; f[3 * id[1] + c] += fj; f[3 * id[2] + c] += fk; f[3 * id[3] + c] += fl[c];
                }
            }
            esum += E;
        }
        et[tid] = esum;
        }
        for (int t = 0; t < nth; ++t) etot += et[t];
        free(et);
        if (bad) s->range_error = 1;
        s->inter_energy[b->inter] += etot;
    }
}
/* updateForces(): initForces + every interaction in addInteraction order */
void orc_compute_forces(orc_sim *s) {
    if (!s->lists_valid) orc_rebuild(s);
    memset(s->f, 0, 24 * (size_t)s->n);
    for (int k = 0; k < s->ninter; ++k) s->inter_energy[k] = 0;
    double *fbuf = force_buffers(s);
    nonbonded_forces(s, fbuf);
    bonded_forces(s, fbuf);
    reduce_force_buffers(s, fbuf);
    /* integrator.CapForce(system, capForce) (src/start_simulation.py:320-324) [EXT, U26]: after the force calculation every
     * particle's force vector longer than capForce is scaled back to that length; the thermostat (added to the integrator
     * after CapForce) acts on the capped forces */
    if (s->cap_force > 0) {
#pragma omp parallel for num_threads(s->nthreads) schedule(static)
        for (int i = 0; i < s->n; ++i) {
            double *f = s->f + 3 * (size_t)i, f2 = f[0] * f[0] + f[1] * f[1] + f[2] * f[2];
            if (f2 > s->cap_force * s->cap_force) { double k = s->cap_force / sqrt(f2); f[0] *= k; f[1] *= k; f[2] *= k; }
        }
    }
    s->forces_valid = 1;
}
void orc_set_cap_force(orc_sim *s, double cap) { s->cap_force = cap; s->forces_valid = 0; }
double orc_energy(orc_sim *s, int inter) {
    if (!s->forces_valid) orc_compute_forces(s);
    return s->inter_energy[inter];
}
void orc_kinetics(orc_sim *s, double out[3]) {
    double ek = 0;
    for (int i = 0; i < s->n; ++i) ek += 0.5 * s->mass[i] * (s->v[3 * i] * s->v[3 * i] + s->v[3 * i + 1] * s->v[3 * i + 1] + s->v[3 * i + 2] * s->v[3 * i + 2]);
    out[0] = ek; out[1] = 2 * ek / (3.0 * s->n); out[2] = s->n;
}
int orc_range_error(orc_sim *s) { return s->range_error; }

/* ---- integrator: VelocityVerlet + LangevinThermostat [EXT VelocityVerlet.cpp,
 * LangevinThermostat.cpp]: f += -gamma m v + sqrt(24 kT gamma/dt) sqrt(m) (u - 1/2)  (U11) ---- */
void orc_set_dt(orc_sim *s, double dt) { s->dt = dt; }
void orc_set_langevin(orc_sim *s, int on, double kT, double gamma, int ntypes, const int *types) {
    s->lang_on = on; s->kT = kT; s->gamma = gamma; s->lang_all = (ntypes == 0);
    memset(s->lang_type, 0, sizeof(s->lang_type));
    for (int i = 0; i < ntypes; ++i) s->lang_type[types[i]] = 1;
}
static void thermalize(orc_sim *s, uint32_t stream, uint64_t step, double scale) {
    if (!s->lang_on) return;
    double pref1 = -s->gamma, pref2 = sqrt(24.0 * s->kT * s->gamma / s->dt) * scale;
#pragma omp parallel for num_threads(s->nthreads) schedule(static)
    for (int i = 0; i < s->n; ++i) {
        if (!s->lang_all && !s->lang_type[s->type[i]]) continue;
        double u[3]; draw3(s->seed, stream, step, (uint32_t)i, u);
        double m = s->mass[i], sm = sqrt(m);
        for (int c = 0; c < 3; ++c) s->f[3 * i + c] += pref1 * m * s->v[3 * i + c] + pref2 * sm * (u[c] - 0.5);
    }
}
static void fold(orc_sim *s) {
#pragma omp parallel for num_threads(s->nthreads) schedule(static)
    for (int i = 0; i < s->n; ++i) for (int d = 0; d < 3; ++d) {
        double L = s->box[d], xx = s->x[3 * i + d];
        if (xx < 0 || xx >= L) { double im = floor(xx / L); s->image[3 * i + d] += (int)im; xx -= im * L; if (xx >= L) { xx -= L; s->image[3 * i + d]++; } s->x[3 * i + d] = xx; }
    }
}
void orc_react(orc_sim *s);
static void run_impl(orc_sim *s, int64_t nsteps, int cont) {
    /* run entry: resort if flagged, recompute forces with the thermostat heat-up factor sqrt(3).
     * cont: continuation inside one integrator.run of the reference (ExtAnalyze / ATRPActivator fire from signals inside
     * VelocityVerlet::run [EXT]): the forces of the last step are still in place, runInit/recalc are not repeated. */
    if (!cont) {
        if (!s->lists_valid) { fold(s); orc_rebuild(s); }
        orc_compute_forces(s);
        thermalize(s, STREAM_HEATUP, (uint64_t)s->step, sqrt(3.0));
    }
    double dt = s->dt;
    for (int64_t it = 0; it < nsteps; ++it) {
        double maxsq = 0;
#pragma omp parallel for num_threads(s->nthreads) schedule(static) reduction(max : maxsq)
        for (int i = 0; i < s->n; ++i) {
            double dtfm = 0.5 * dt / s->mass[i], sq = 0;
            for (int c = 0; c < 3; ++c) {
                s->v[3 * i + c] += dtfm * s->f[3 * i + c];
                double dp = dt * s->v[3 * i + c];
                s->x[3 * i + c] += dp; sq += dp * dp;
            }
            if (sq > maxsq) maxsq = sq;
        }
        int resort;
        if (s->criterion == 0) { s->maxdist += sqrt(maxsq); resort = s->maxdist > 0.5 * s->skin; }
        else {
            double m2 = 0;
#pragma omp parallel for num_threads(s->nthreads) schedule(static) reduction(max : m2)
            for (int i = 0; i < s->n; ++i) { double q = 0; for (int c = 0; c < 3; ++c) { double d = minimg(s->x[3 * i + c] - s->xref[3 * i + c], s->box[c]); q += d * d; } if (q > m2) m2 = q; }
            resort = sqrt(m2) > 0.5 * s->skin;
        }
        if (resort || !s->lists_valid) { fold(s); orc_rebuild(s); }
        orc_compute_forces(s);
        thermalize(s, STREAM_LANGEVIN, (uint64_t)(s->step + it), 1.0);
#pragma omp parallel for num_threads(s->nthreads) schedule(static)
        for (int i = 0; i < s->n; ++i) {
            double dtfm = 0.5 * dt / s->mass[i];
            for (int c = 0; c < 3; ++c) s->v[3 * i + c] += dtfm * s->f[3 * i + c];
        }
        /* aftIntV: ChemicalReaction::React every `interval` completed steps (U17) */
        if (s->react_on && ((s->step + it + 1) % s->interval) == 0) {
            int64_t keep = s->step; s->step = keep + it + 1; orc_react(s); s->step = keep;
        }
    }
    s->step += nsteps;
    s->cont_ok = 1;
}
void orc_run(orc_sim *s, int64_t nsteps) { run_impl(s, nsteps, 0); }
void orc_run_continue(orc_sim *s, int64_t nsteps) { run_impl(s, nsteps, s->cont_ok); }
int64_t orc_step(orc_sim *s) { return s->step; }
int64_t orc_nrebuild(orc_sim *s) { return s->nrebuild; }
int64_t orc_npairs(orc_sim *s) { return s->npairs; }
/* number of pairs within their force cutoff at the current positions (pair-interactions per step) */
int64_t orc_count_interacting(orc_sim *s) {
    int64_t c = 0; double d[3];
    for (int64_t k = 0; k < s->npairs; ++k) {
        int i = s->pairs[2 * k], j = s->pairs[2 * k + 1];
        const orc_pairpot *p = &s->pp[s->type[i]][s->type[j]];
        if (p->kind != NB_NONE && dist2(s, i, j, d) <= p->rc2) ++c;
    }
    return c;
}

/* ---- reactions: ChemicalReaction::React (reaction_setup.py:417-427,81-113,506)
 * [EXT integrator/ChemicalReaction.cpp, ChemicalReactionPostProcess.cpp, TopologyManager.cpp] ---- */
void orc_reaction_general(orc_sim *s, int on, int interval, int nearest, int max_per_interval) {
    s->react_on = on; s->interval = interval > 0 ? interval : 1; s->nearest = nearest; s->max_per_interval = max_per_interval;
}
int orc_add_reaction(orc_sim *s, int type_1, int type_2, int delta_1, int delta_2, int min1, int max1, int min2, int max2,
                     double rate, double cutoff, double min_cutoff, int list, int intramolecular, int intraresidual,
                     int is_virtual, int active) {
    s->reac = realloc(s->reac, sizeof(orc_reaction) * (s->nreac + 1));
    orc_reaction r = {type_1, type_2, delta_1, delta_2, min1, max1, min2, max2, rate, cutoff, min_cutoff, list,
                      intramolecular, intraresidual, is_virtual, active, 0, -1, NULL};
    s->reac[s->nreac] = r;
    return s->nreac++;
}
/* the call replaces the map; pairs = [n][2] dense indices, either order */
void orc_reaction_define_connections(orc_sim *s, int r, int64_t n, const int *pairs) {
    orc_reaction *R = &s->reac[r];
    free(R->conn);
    R->conn = malloc(sizeof(uint64_t) * (size_t)(n > 0 ? n : 1));
    for (int64_t k = 0; k < n; ++k) {
        int a = pairs[2 * k], b = pairs[2 * k + 1];
        R->conn[k] = ((uint64_t)(uint32_t)(a < b ? a : b) << 32) | (uint32_t)(a < b ? b : a);
    }
    qsort(R->conn, (size_t)n, sizeof(uint64_t), cmp_u64);
    R->nconn = n;
}
static int conn_allows(const orc_reaction *r, int i, int j) {
    if (r->nconn < 0) return 1;
    uint64_t key = ((uint64_t)(uint32_t)(i < j ? i : j) << 32) | (uint32_t)(i < j ? j : i);
    return bsearch(&key, r->conn, (size_t)r->nconn, sizeof(uint64_t), cmp_u64) != NULL;
}
void orc_reaction_set_rate(orc_sim *s, int r, double rate) { s->reac[r].rate = rate; }
void orc_reaction_set_active(orc_sim *s, int r, int a) { s->reac[r].active = a; }
void orc_reaction_add_change(orc_sim *s, int reaction, int side, int nb_level, int old_type, int new_type,
                             double new_mass, double new_q, int state_mode, int state_value) {
    s->chg = realloc(s->chg, sizeof(orc_change) * (s->nchg + 1));
    orc_change c = {reaction, side, nb_level, old_type, new_type, state_mode, state_value, new_mass, new_q};
    s->chg[s->nchg++] = c;
    if (new_type + 1 > s->ntypes) s->ntypes = new_type + 1;
}
void orc_tm_observe(orc_sim *s, int list) { s->tm_obs[s->ntm_obs++] = list; }
void orc_tm_register(orc_sim *s, int list, int t1, int t2, int t3, int t4) {
    s->tmreg = realloc(s->tmreg, sizeof(orc_tmreg) * (s->ntmreg + 1));
    orc_tmreg r = {list, {t1, t2, t3, t4}}; s->tmreg[s->ntmreg++] = r;
}
/* TopologyManager.initialize_topology(): bond graph + molecule ids from the observed pair lists */
void orc_tm_initialize(orc_sim *s) {
    memset(s->deg, 0, 4 * (size_t)s->n);
    for (int i = 0; i < s->n; ++i) s->mol[i] = i;
    for (int k = 0; k < s->ntm_obs; ++k) {
        orc_list *l = &s->lists[s->tm_obs[k]];
        if (l->arity != 2) continue;
        for (int64_t b = 0; b < l->n; ++b) { graph_add(s, l->ids[2 * b], l->ids[2 * b + 1]); mol_union(s, l->ids[2 * b], l->ids[2 * b + 1]); }
    }
    s->tm_init = 1;
}
int64_t orc_reaction_counter(orc_sim *s, int r) { return s->reac[r].counter; }
int64_t orc_last_events(orc_sim *s) { return s->last_events; }

static void cand_push(orc_sim *s, orc_cand c) {
    if (s->ncands == s->capcands) { s->capcands = s->capcands ? 2 * s->capcands : 1024; s->cands = realloc(s->cands, s->capcands * sizeof(orc_cand)); }
    s->cands[s->ncands++] = c;
}
static int cmp_cand_rows(const void *a, const void *b) {
    const orc_cand *x = a, *y = b;
    if (x->a != y->a) return x->a < y->a ? -1 : 1;
    if (x->b != y->b) return x->b < y->b ? -1 : 1;
    return x->r < y->r ? -1 : (x->r > y->r);
}
/* order inside a partner group: nearest -> (d2, partner, reaction); random -> (hash, partner, reaction) (U7) */
static inline int better(const orc_sim *s, const orc_cand *x, const orc_cand *y, int partner_is_b) {
    if (s->nearest) { if (x->d2 != y->d2) return x->d2 < y->d2; }
    else { if (x->rnd != y->rnd) return x->rnd < y->rnd; }
    int px = partner_is_b ? x->b : x->a, py = partner_is_b ? y->b : y->a;
    if (px != py) return px < py;
    return x->r < y->r;
}
static int side_ok(const orc_sim *s, const orc_reaction *r, int A, int B) {
    return s->type[A] == r->type_1 && s->type[B] == r->type_2 && s->state[A] >= r->min1 && s->state[A] < r->max1 &&
           s->state[B] >= r->min2 && s->state[B] < r->max2; /* U6 */
}
static void apply_props(orc_sim *s, const orc_change *c, int p) {
    if (s->type[p] != c->old_type) return;
    s->type[p] = c->new_type;
    if (c->new_mass > 0) s->mass[p] = c->new_mass;
    if (c->new_q == c->new_q) s->q[p] = c->new_q;
    if (c->state_mode == 1) s->state[p] = c->state_value;
    else if (c->state_mode == 2) s->state[p] += c->state_value;
}
static void tm_emit(orc_sim *s, int ar, const int *ids) {
    int ty[4]; for (int m = 0; m < ar; ++m) ty[m] = s->type[ids[m]];
    for (int k = 0; k < s->ntmreg; ++k) {
        orc_tmreg *g = &s->tmreg[k];
        if (s->lists[g->list].arity != ar) continue;
        int fwd = 1, rev = 1;
        for (int m = 0; m < ar; ++m) { if (g->t[m] != ty[m]) fwd = 0; if (g->t[m] != ty[ar - 1 - m]) rev = 0; }
        if (fwd || rev) {
            list_push(s, g->list, ids);
            for (int o = 0; o < s->nobs_excl; ++o) if (s->obs_excl[o] == g->list) excl_push(s, ids[0], ids[ar - 1]);
            return; /* first matching registration wins */
        }
    }
}
void orc_react(orc_sim *s) {
    if (!s->lists_valid) orc_rebuild(s);
    uint64_t step = (uint64_t)s->step;
    s->ncands = 0; s->last_events = 0;
    /* 1. candidate search over the Verlet pairs (same list as the force loop, reaction_setup.py:419) */
    for (int64_t k = 0; k < s->npairs; ++k) {
        int i = s->pairs[2 * k], j = s->pairs[2 * k + 1]; /* i < j */
        double d[3], d2 = -1;
        for (int ri = 0; ri < s->nreac; ++ri) {
            const orc_reaction *r = &s->reac[ri];
            if (!r->active) continue;
            int A, B;
            if (side_ok(s, r, i, j)) { A = i; B = j; }       /* orientation 1: lower index as type_1 */
            else if (side_ok(s, r, j, i)) { A = j; B = i; }
            else continue;
            if (!r->intraresidual && s->resid[A] == s->resid[B]) continue; /* U10 */
            if (!r->intramolecular && mol_find(s, A) == mol_find(s, B)) continue;
            if (!conn_allows(r, i, j)) continue;
            if (d2 < 0) d2 = dist2(s, i, j, d);
            if (!(d2 >= r->min_cutoff * r->min_cutoff && d2 < r->cutoff * r->cutoff)) continue; /* U3 */
            uint32_t w[4]; draw_pair(s->seed, STREAM_REACT, step, (uint32_t)i, (uint32_t)j, (uint32_t)ri, w);
            double W = (double)w[0] * (1.0 / 4294967296.0);
            double p = r->rate * s->dt * s->interval; /* U5 */
            orc_cand c; c.a = A; c.b = B; c.r = ri; c.d2 = d2; c.accepted = (W < p);
            uint32_t h[4]; draw_pair(s->seed, STREAM_PARTNER, step, (uint32_t)i, (uint32_t)j, (uint32_t)ri, h);
            c.rnd = ((uint64_t)h[0] << 32) | h[1];
            cand_push(s, c);
        }
    }
    qsort(s->cands, s->ncands, sizeof(orc_cand), cmp_cand_rows);
    int64_t nc = s->ncands;
    if (nc == 0) return;
    /* 2. UniqueA: every A keeps one partner among its ACCEPTED candidates */
    unsigned char *alive = calloc(nc, 1);
    for (int64_t k = 0; k < nc; ++k) alive[k] = (unsigned char)s->cands[k].accepted;
    {
        int64_t k = 0;
        while (k < nc) {
            int64_t e = k, best = -1;
            while (e < nc && s->cands[e].a == s->cands[k].a) {
                if (alive[e] && (best < 0 || better(s, &s->cands[e], &s->cands[best], 1))) best = e;
                ++e;
            }
            for (int64_t m = k; m < e; ++m) if (m != best) alive[m] = 0;
            k = e;
        }
    }
    /* 3. UniqueB among the survivors */
    {
        int64_t *bestb = malloc(8 * (size_t)s->n);
        for (int i = 0; i < s->n; ++i) bestb[i] = -1;
        for (int64_t k = 0; k < nc; ++k) if (alive[k]) {
            int b = s->cands[k].b;
            if (bestb[b] < 0 || better(s, &s->cands[k], &s->cands[bestb[b]], 0)) bestb[b] = k;
        }
        for (int64_t k = 0; k < nc; ++k) if (alive[k] && bestb[s->cands[k].b] != k) alive[k] = 0;
        free(bestb);
    }
    /* 4. U8: one reaction per particle per interval; greedy in canonical (A, B, r) order */
    unsigned char *used = calloc(s->n, 1);
    int64_t nev = 0; int64_t *ev = malloc(8 * (size_t)nc);
    for (int64_t k = 0; k < nc; ++k) if (alive[k]) {
        orc_cand *c = &s->cands[k];
        if (used[c->a] || used[c->b]) { alive[k] = 0; continue; }
        if (s->max_per_interval > 0 && nev >= s->max_per_interval) { alive[k] = 0; continue; }
        used[c->a] = used[c->b] = 1; ev[nev++] = k;
    }
    free(used);
    /* 5. apply: reactant property changes + state deltas */
    for (int64_t e = 0; e < nev; ++e) {
        orc_cand *c = &s->cands[ev[e]]; orc_reaction *r = &s->reac[c->r];
        for (int q = 0; q < s->nchg; ++q) {
            orc_change *g = &s->chg[q];
            if (g->reaction != c->r || g->nb_level != 0) continue;
            if (g->side & 1) apply_props(s, g, c->a);
            if (g->side & 2) apply_props(s, g, c->b);
        }
        s->state[c->a] += r->delta_1; s->state[c->b] += r->delta_2;
        r->counter++;
    }
    /* 6. new bonds -> lists, graph, molecule ids, (A,B) exclusions */
    for (int64_t e = 0; e < nev; ++e) {
        orc_cand *c = &s->cands[ev[e]]; orc_reaction *r = &s->reac[c->r];
        if (r->is_virtual) continue;
        int ids[2] = {c->a, c->b};
        list_push(s, r->list, ids);
        if (tm_observed(s, r->list)) { graph_add(s, c->a, c->b); mol_union(s, c->a, c->b); }
        for (int o = 0; o < s->nobs_excl; ++o) if (s->obs_excl[o] == r->list) excl_push(s, c->a, c->b);
    }
    /* 7. neighbour property changes: particles exactly nb_level bonds from the reactant (BFS on the
     * updated graph), PostProcessChangeNeighboursProperty (reaction_post_process.py:76-115).
     * Deterministic, order-independent formulation (the engine runs the events in parallel): every
     * reached particle whose type -- as it stands after step 6 -- equals a rule's old type files the
     * claim (event, side, level, rule) with the FIRST matching rule; the smallest claim per particle
     * is applied. */
    {
        uint64_t *claim = malloc(8 * (size_t)s->n);
        int *touched = malloc(4 * (size_t)s->n); int nt = 0;
        for (int i = 0; i < s->n; ++i) claim[i] = ~0ull;
        for (int64_t e = 0; e < nev; ++e) {
            orc_cand *c = &s->cands[ev[e]];
            for (int side = 1; side <= 2; ++side) {
                int root = side == 1 ? c->a : c->b;
                int maxlev = 0;
                for (int q = 0; q < s->nchg; ++q) if (s->chg[q].reaction == c->r && (s->chg[q].side & side) && s->chg[q].nb_level > maxlev) maxlev = s->chg[q].nb_level;
                if (!maxlev) continue;
                int front[64], nf = 1, seen[192], ns = 1; front[0] = root; seen[0] = root;
                for (int lev = 1; lev <= maxlev; ++lev) {
                    int nxt[64], nn = 0;
                    for (int a = 0; a < nf; ++a) for (int k = 0; k < s->deg[front[a]]; ++k) {
                        int y = s->adj[front[a] * ORC_MAXDEG + k], dup = 0;
                        for (int z = 0; z < ns; ++z) if (seen[z] == y) { dup = 1; break; }
                        if (!dup && nn < 64 && ns < 192) { nxt[nn++] = y; seen[ns++] = y; }
                    }
                    for (int a = 0; a < nn; ++a) for (int q = 0; q < s->nchg; ++q) {
                        orc_change *g = &s->chg[q];
                        if (g->reaction == c->r && (g->side & side) && g->nb_level == lev && g->old_type == s->type[nxt[a]]) {
                            uint64_t key = ((uint64_t)e << 24) | ((uint64_t)(side - 1) << 20) | ((uint64_t)lev << 12) | (uint64_t)q;
                            if (claim[nxt[a]] == ~0ull) touched[nt++] = nxt[a];
                            if (key < claim[nxt[a]]) claim[nxt[a]] = key;
                            break;
                        }
                    }
                    memcpy(front, nxt, 4 * nn); nf = nn;
                }
            }
        }
        for (int t = 0; t < nt; ++t) apply_props(s, &s->chg[claim[touched[t]] & 0xfff], touched[t]);
        free(claim); free(touched);
    }
    /* 8. TopologyManager: new angles / dihedrals through the new bond, final types */
    for (int64_t e = 0; e < nev; ++e) {
        orc_cand *c = &s->cands[ev[e]]; orc_reaction *r = &s->reac[c->r];
        if (r->is_virtual || !tm_observed(s, r->list)) continue;
        int a = c->a, b = c->b;
        /* a tuple that contains several bonds created in this pass is emitted by the LAST of them in
         * event order (sequential semantics: it only exists once all of its bonds exist) */
        /* triples x-a-b and a-b-y */
        for (int side = 0; side < 2; ++side) {
            int p = side ? b : a, o = side ? a : b;
            for (int k = 0; k < s->deg[p]; ++k) {
                int x = s->adj[p * ORC_MAXDEG + k]; if (x == o) continue;
                int later = 0;
                for (int64_t e2 = e + 1; e2 < nev; ++e2) { orc_cand *c2 = &s->cands[ev[e2]]; if (!s->reac[c2->r].is_virtual && pkey(c2->a, c2->b) == pkey(x, p)) later = 1; }
                if (later) continue;
                int t3[3] = {x, p, o}; tm_emit(s, 3, t3);
                /* quads y-x-p-o */
                for (int k2 = 0; k2 < s->deg[x]; ++k2) {
                    int y = s->adj[x * ORC_MAXDEG + k2]; if (y == p || y == o) continue;
                    int later2 = 0;
                    for (int64_t e2 = e + 1; e2 < nev; ++e2) { orc_cand *c2 = &s->cands[ev[e2]]; if (!s->reac[c2->r].is_virtual && pkey(c2->a, c2->b) == pkey(y, x)) later2 = 1; }
                    if (later2) continue;
                    int t4[4] = {y, x, p, o}; tm_emit(s, 4, t4);
                }
            }
        }
        /* quads x-a-b-y */
        for (int k = 0; k < s->deg[a]; ++k) {
            int x = s->adj[a * ORC_MAXDEG + k]; if (x == b) continue;
            for (int k2 = 0; k2 < s->deg[b]; ++k2) {
                int y = s->adj[b * ORC_MAXDEG + k2]; if (y == a || y == x) continue;
                int later = 0;
                for (int64_t e2 = e + 1; e2 < nev; ++e2) { orc_cand *c2 = &s->cands[ev[e2]]; if (s->reac[c2->r].is_virtual) continue; uint64_t kk = pkey(c2->a, c2->b); if (kk == pkey(x, a) || kk == pkey(b, y)) later = 1; }
                if (later) continue;
                int t4[4] = {x, a, b, y}; tm_emit(s, 4, t4);
            }
        }
    }
    free(ev); free(alive);
    s->last_events = nev;
    if (nev) {
        excl_finalize(s);
        update_mixing(s);
        s->lists_valid = 0; /* U9: new exclusions take effect through a forced rebuild */
        s->forces_valid = 0;
    }
}
int64_t orc_get_candidates(orc_sim *s, int64_t cap, int *rows, double *d2) {
    for (int64_t k = 0; k < s->ncands && k < cap; ++k) {
        rows[4 * k] = s->cands[k].a; rows[4 * k + 1] = s->cands[k].b; rows[4 * k + 2] = s->cands[k].r; rows[4 * k + 3] = s->cands[k].accepted;
        if (d2) d2[k] = s->cands[k].d2;
    }
    return s->ncands;
}
int64_t orc_count_type(orc_sim *s, int type, int state) {
    int64_t c = 0; for (int i = 0; i < s->n; ++i) if (s->type[i] == type && (state < 0 || s->state[i] == state)) ++c; return c;
}
void orc_update_mixing(orc_sim *s) { update_mixing(s); }

/* ---- ATRPActivator (src/chemlab/reaction_post_process.py:393-424; [EXT] integrator/ATRPActivator.cpp, U22) ----------------
 * Sequential restatement: list every particle that sits on a reactive centre (the first registered centre with its type and
 * state), give it the key (u, index) with u = first word of Philox(seed ^ "ATRP", step, index), take the num_particles
 * smallest keys, and let each of them react when the second Philox word (as a uniform in (0,1)) is below
 * k_deactivate*ratio_deactivator (flag "DA") or k_activate*ratio_activator (flag "A") -- the ratios of the START of the pass.
 * An activation moves one catalyst complex from the activator to the deactivator pool and vice versa. */
#define ORC_STREAM_ATRP 0x41545250u
void orc_atrp_configure(orc_sim *s, int num_particles, double ratio_activator, double ratio_deactivator, double delta_catalyst,
                        double k_activate, double k_deactivate) {
    s->atrp_num = num_particles; s->atrp_ratio[0] = ratio_activator; s->atrp_ratio[1] = ratio_deactivator;
    s->atrp_delta = delta_catalyst; s->atrp_k[0] = k_activate; s->atrp_k[1] = k_deactivate; s->atrp_ncen = 0;
}
void orc_atrp_add_center(orc_sim *s, int type, int state, int needs_deactivator, int new_type, double new_mass, double new_q, int delta_state) {
    int k = s->atrp_ncen++;
    s->atrp_cen[k].type = type; s->atrp_cen[k].state = state; s->atrp_cen[k].deact = needs_deactivator ? 1 : 0;
    s->atrp_cen[k].new_type = new_type; s->atrp_cen[k].delta = delta_state; s->atrp_cen[k].new_mass = new_mass; s->atrp_cen[k].new_q = new_q;
    if (new_type + 1 > s->ntypes) s->ntypes = new_type + 1;
}
static int atrp_center_of(const orc_sim *s, int i) {
    for (int k = 0; k < s->atrp_ncen; ++k) if (s->atrp_cen[k].type == s->type[i] && s->atrp_cen[k].state == s->state[i]) return k;
    return -1;
}
void orc_atrp_now(orc_sim *s, int64_t counts[2], double ratios[2]) {
    int64_t nact = 0, ndeact = 0;
    if (s->atrp_ncen > 0 && s->atrp_num > 0) {
        uint64_t *keys = malloc((size_t)s->n * 8 + 8);
        int64_t nc = 0;
        for (int i = 0; i < s->n; ++i) {
            if (atrp_center_of(s, i) < 0) continue;
            uint32_t c[4] = {(uint32_t)i, 0u, (uint32_t)s->step, (uint32_t)((uint64_t)s->step >> 32)};
            philox4x32_10(c, (uint32_t)s->seed, (uint32_t)(s->seed >> 32) ^ ORC_STREAM_ATRP);
            keys[nc++] = ((uint64_t)c[0] << 32) | (uint32_t)i;
        }
        qsort(keys, (size_t)nc, 8, cmp_u64);
        const double p_act = s->atrp_k[0] * s->atrp_ratio[0], p_deact = s->atrp_k[1] * s->atrp_ratio[1];
        for (int64_t k = 0; k < nc && k < s->atrp_num; ++k) {
            const int i = (int)(keys[k] & 0xffffffffu);
            const int ce = atrp_center_of(s, i);
            uint32_t c[4] = {(uint32_t)i, 0u, (uint32_t)s->step, (uint32_t)((uint64_t)s->step >> 32)};
            philox4x32_10(c, (uint32_t)s->seed, (uint32_t)(s->seed >> 32) ^ ORC_STREAM_ATRP);
            const double w = ((double)c[1] + 0.5) * (1.0 / 4294967296.0);
            const int de = s->atrp_cen[ce].deact;
            if (!(w < (de ? p_deact : p_act))) continue;
            s->state[i] += s->atrp_cen[ce].delta;
            if (s->atrp_cen[ce].new_type >= 0) s->type[i] = s->atrp_cen[ce].new_type;
            if (s->atrp_cen[ce].new_mass > 0) s->mass[i] = s->atrp_cen[ce].new_mass;
            if (s->atrp_cen[ce].new_q == s->atrp_cen[ce].new_q) s->q[i] = s->atrp_cen[ce].new_q;
            if (de) ++ndeact; else ++nact;
        }
        free(keys);
        const double d = s->atrp_delta * (double)(nact - ndeact) / (double)(s->atrp_num > 1 ? s->atrp_num : 1);
        double a = s->atrp_ratio[0] - d, b = s->atrp_ratio[1] + d;
        s->atrp_ratio[0] = a < 0 ? 0 : (a > 1 ? 1 : a); s->atrp_ratio[1] = b < 0 ? 0 : (b > 1 ? 1 : b);
    }
    if (counts) { counts[0] = nact; counts[1] = ndeact; }
    if (ratios) { ratios[0] = s->atrp_ratio[0]; ratios[1] = s->atrp_ratio[1]; }
}
