/* TEST INFRASTRUCTURE -- scalar CPU mirror of the NATIVE-mode kernel (Philox4x32-7 draws, FP32 state).
 *
 * The native kernel cannot be compared with the reference draw by draw (different RNG), only
 * statistically.  To still check its *logic* exactly, this file restates the native algorithm the plain
 * way -- one car at a time, full sorts, no warp tricks -- following the reference's structure
 * (src/simulation.py:102-560, cited per function) with the native arithmetic:
 *   - draws: Philox4x32-R (R = MIRROR_PHILOX_ROUNDS = 7, as the kernel), key = seed, counter = (sim, lap<<8 | lane, stream)
 *       lap 0  word0 of lane p          -> grid position p's uniform
 *       lap 1  word0 lap-1 DNF; words 1,2 -> Box-Muller pair (cos: pace noise, sin: start delta);
 *              word3 (top 23 bits) -> the retirement lap >= 2, geometric: 2 + floor(ln u / ln(1 - rate))
 *       laps (l, l+1), l even: ONE call per lane: words 0,1 (top 24 bits) -> Box-Muller pair (cos: lap l,
 *              sin: lap l+1); word2 lo/hi 16 -> lap l overtake pass 1/2; word3 -> the same for lap l+1;
 *              pass 3 of laps l / l+1 = lo / hi 16 bits of a borrowed word: n <= 20: word0 of lane 20+d (d < 10)
 *              or word1 of lane 10+d (d >= 10); n > 20: word0 of lane 32+d.
 *              events: red, else SC, else VSC decided by ONE word against cumulative thresholds -- word0 (lap l) /
 *              word1 (lap l+1) of lane 31, VSC tyre roll-back: lo / hi 16 bits of word2 (n <= 20); n > 20: words 1 / 2 / 3 of
 *              lane 32's call (driver 0's second call, whose word0 is its pass-3 draw)
 *   - FP32, every fused op explicit (fmaf), times re-based on the leader after every lap,
 *     overtake chains in closed form  base - 0.1*(k - 2*sn), ordering by (time, driver index).
 * With the "exact" normal generator (IEEE-only arithmetic) kernel and mirror agree bit for bit; the
 * mirror itself is checked statistically against the FP64 oracle (tests/test_native_mirror.py).
 * Build flags: -ffp-contract=off (oracle/Makefile). */
#include "native_mirror.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint32_t x, y, z, w; } u4;

/* must equal MCGP_PHILOX_ROUNDS of monte-carlo-gp_b200/csrc/native_math.cuh (checked by tests/test_native_mirror.py) */
#ifndef MIRROR_PHILOX_ROUNDS
#define MIRROR_PHILOX_ROUNDS 7
#endif

static u4 philox_r(int rounds, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    for (int r = 0; r < rounds; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    u4 o = {c0, c1, c2, c3};
    return o;
}
static u4 philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    return philox_r(MIRROR_PHILOX_ROUNDS, c0, c1, c2, c3, k0, k1);
}
int orc_native_philox_rounds(void) { return MIRROR_PHILOX_ROUNDS; }
/* Philox4x32-`rounds` on one counter block: exported for the Random123 known-answer vectors */
void orc_philox4x32(int rounds, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    const u4 o = philox_r(rounds, ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    out[0] = o.x; out[1] = o.y; out[2] = o.z; out[3] = o.w;
}

static float as_float(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }
static uint32_t as_uint(float f) { uint32_t b; memcpy(&b, &f, 4); return b; }

static float exact_log(float x) { /* Cephes logf on [2^-24, 1] */
    uint32_t b = as_uint(x);
    int e = (int)(b >> 23) - 126;
    float m = as_float((b & 0x007fffffu) | 0x3f000000u);
    if (m < 0.70710678118654752440f) { e -= 1; m = (m + m) + -1.0f; } else m = m + -1.0f;
    float z = m * m;
    float y = 7.0376836292e-2f;
    y = fmaf(y, m, -1.1514610310e-1f);
    y = fmaf(y, m, 1.1676998740e-1f);
    y = fmaf(y, m, -1.2420140846e-1f);
    y = fmaf(y, m, 1.4249322787e-1f);
    y = fmaf(y, m, -1.6668057665e-1f);
    y = fmaf(y, m, 2.0000714765e-1f);
    y = fmaf(y, m, -2.4999993993e-1f);
    y = fmaf(y, m, 3.3333331174e-1f);
    y = (y * m) * z;
    float fe = (float)e;
    y = fmaf(-2.12194440e-4f, fe, y);
    y = fmaf(-0.5f, z, y);
    float r = m + y;
    return fmaf(0.693359375f, fe, r);
}

static void normal2(int exact, uint32_t w1, uint32_t w2, float* za, float* zb) {
    if (!exact) { /* libm stand-in for the MUFU path: same formula, not bit-identical */
        float lg = log2f((float)((w1 >> 8) + 1u));
        float r2 = fmaxf(fmaf(lg, -1.3862943611198906f, 33.27106466687737f), 0.0f);
        float r = sqrtf(r2), a = (float)(w2 >> 8) * 3.7450703e-07f;
        *za = r * cosf(a);
        *zb = r * sinf(a);
        return;
    }
    float u = (float)((w1 >> 8) + 1u) * 5.9604644775390625e-08f;
    float rad = sqrtf(-2.0f * exact_log(u));
    uint32_t k = w2 >> 8;
    uint32_t q = (k + (1u << 21)) >> 22;
    int r = (int)k - (int)(q << 22);
    float x = (float)r * 3.7450703e-07f;
    float x2 = x * x;
    float pc = fmaf(2.443315711809948e-5f, x2, -1.388731625493765e-3f);
    pc = fmaf(pc, x2, 4.166664568298827e-2f);
    float cx = fmaf(x2 * x2, pc, fmaf(-0.5f, x2, 1.0f));
    float ps = fmaf(-1.9515295891e-4f, x2, 8.3321608736e-3f);
    ps = fmaf(ps, x2, -1.6666654611e-1f);
    float sx = fmaf(x * x2, ps, x);
    float c, s;
    switch (q & 3u) {
        case 0: c = cx; s = sx; break;
        case 1: c = -sx; s = cx; break;
        case 2: c = -cx; s = -sx; break;
        default: c = sx; s = -cx; break;
    }
    *za = rad * c;
    *zb = rad * s;
}

static uint32_t prob_threshold(double p) {
    if (!(p > 0.0)) return 0u;
    if (p >= 1.0) return 0xffffffffu;
    return (uint32_t)floor(p * 4294967296.0);
}

static double clamp01(double p) { return !(p > 0.0) ? 0.0 : p > 1.0 ? 1.0 : p; }
#define DNF_NEVER (-1e30f)
static float dnf_scale(double rate) { /* 1 / ln(1 - rate) on the 2^-32 probability grid */
    const double thr = (double)prob_threshold(rate) / 4294967296.0;
    if (!(thr > 0.0)) return DNF_NEVER;
    if (thr >= 1.0 || rate >= 1.0) return -0.0f;
    const float s = (float)(1.0 / log1p(-thr));
    return s < -1e29f ? -1e29f : s;
}

#define N32 32
typedef struct {
    int n, L, track;
    float pace[N32], deg_ovt[N32], sigma[N32], eff[5][N32], opt[5][N32], pc[5][N32], dnf_scale[N32], G[N32][N32];
    uint32_t lap1_thr[N32], red_thr, sc_thr, vsc_thr;
    float pit_loss, ovt_delta, drs_delta, dirty_thr, dirty_pen;
    int grid_fixed, fixed_slot[N32];
} nat_t;

/* Overtake paces for the PROBABILITY (x 2^15, FP32): all reachable paces base_pace + age * tire_deg (FP64, :514-515),
 * sorted, mapped to a strictly increasing float sequence (two paces that round to the same float: the later one is
 * nudged up by one ulp) -- the same construction as the kernel's host-built pace table, where the strict
 * monotonicity is what lets ONE float compare reproduce the FP64 pair decision. */
static int cmp_double(const void* a, const void* b) {
    const double x = *(const double*)a, y = *(const double*)b;
    return x < y ? -1 : x > y;
}
static float* build_op32(const orc_params* p) {
    const int n = p->n_drivers, rows = p->total_laps + 5;
    double* P = (double*)malloc(sizeof(double) * (size_t)rows * n);
    double* u = (double*)malloc(sizeof(double) * (size_t)rows * n);
    float* f = (float*)malloc(sizeof(float) * (size_t)rows * n);
    float* tab = (float*)malloc(sizeof(float) * (size_t)rows * n);
    int m = 0;
    for (int a = 0; a < rows; a++)
        for (int d = 0; d < n; d++) {
            const double wear = (double)a * p->tire_deg[d];
            const double v = p->base_pace[d] + wear;
            P[(size_t)a * n + d] = v;
            if (v == v) u[m++] = v;
        }
    qsort(u, (size_t)m, sizeof(double), cmp_double);
    int k = 0;
    for (int i = 0; i < m; i++) if (i == 0 || u[i] != u[k - 1]) u[k++] = u[i];
    m = k;
    for (int i = 0; i < m; i++) {
        f[i] = (float)(u[i] * 32768.0);
        if (i > 0 && !(f[i] > f[i - 1])) f[i] = nextafterf(f[i - 1], INFINITY);
    }
    for (size_t i = 0; i < (size_t)rows * n; i++) {
        const double v = P[i];
        if (v != v) { tab[i] = NAN; continue; }
        int lo = 0, hi = m;  /* first index with u[idx] >= v */
        while (lo < hi) { const int mid = (lo + hi) / 2; if (u[mid] < v) lo = mid + 1; else hi = mid; }
        tab[i] = f[lo];
    }
    free(P); free(u); free(f);
    return tab;
}

int orc_native_op32_table(const orc_params* p, float* out) {
    if (p->n_drivers < 1 || p->n_drivers > N32 || !out) return -1;
    float* tab = build_op32(p);
    memcpy(out, tab, sizeof(float) * (size_t)(p->total_laps + 5) * p->n_drivers);
    free(tab);
    return 0;
}

static void derive(const orc_params* p, nat_t* o) {
    memset(o, 0, sizeof(*o));
    int n = p->n_drivers;
    o->n = n; o->L = p->total_laps; o->track = p->track_condition;
    o->pit_loss = (float)p->pit_loss; o->ovt_delta = (float)p->overtake_delta; o->drs_delta = (float)p->drs_delta;
    o->dirty_thr = (float)p->dirty_thr; o->dirty_pen = (float)p->dirty_pen;
    { /* red flag, else SC, else VSC: one draw against the cumulative probabilities */
        const double pr = clamp01(p->red_p), ps = clamp01(p->sc_p), pv = clamp01(p->vsc_p);
        o->red_thr = prob_threshold(pr);
        o->sc_thr = prob_threshold(pr + (1.0 - pr) * ps);
        o->vsc_thr = prob_threshold(pr + (1.0 - pr) * (ps + (1.0 - ps) * pv));
        if (o->sc_thr < o->red_thr) o->sc_thr = o->red_thr;
        if (o->vsc_thr < o->sc_thr) o->vsc_thr = o->sc_thr;
    }
    for (int d = 0; d < n; d++) {
        o->pace[d] = (float)p->base_pace[d]; o->deg_ovt[d] = (float)p->tire_deg[d]; o->sigma[d] = (float)p->variance[d];
        o->dnf_scale[d] = dnf_scale(p->dnf_rate[d]);
        o->lap1_thr[d] = prob_threshold(p->team_rate[d] * 4.0);
        double deg = p->tire_deg[d], factor = deg > 0 ? deg / 0.05 : 1.0;
        for (int c = 0; c < 5; c++) {
            o->eff[c][d] = (float)(p->compound_deg_rate[c] * factor);
            double optimal = p->compound_optimal[c], dd = p->tire_deg_pit[d];
            if (dd > 0.05) optimal = trunc(optimal * 0.85); else if (dd < 0.02) optimal = trunc(optimal * 1.1);
            o->opt[c][d] = (float)optimal;
            o->pc[c][d] = (float)p->base_pace[d] + (float)p->compound_pace_delta[c];
        }
        for (int pos = 0; pos < n; pos++)
            o->G[pos][d] = p->grid_kind[d][pos] != ORC_ITEM_INT0 ? (float)p->grid_probs[d][pos] : 0.0f;
    }
    /* a permutation-pattern grid (one-hot rows) is deterministic: no sampling */
    int fixed = 1, owner[N32], seen[N32] = {0};
    for (int pos = 0; pos < n && fixed; pos++) {
        int cnt = 0;
        for (int d = 0; d < n; d++) if (o->G[pos][d] > 0.0f) { cnt++; owner[pos] = d; }
        if (cnt != 1) fixed = 0;
    }
    for (int pos = 0; pos < n && fixed; pos++) if (seen[owner[pos]]++) fixed = 0;
    if (fixed) for (int pos = 0; pos < n; pos++) o->fixed_slot[owner[pos]] = pos;
    o->grid_fixed = fixed;
}

typedef struct {
    int slot, comp, dnf, dnf_lap, drs, pos_live;
    uint32_t used;
    float age, eff, opt, pc, t, last, ahead_last;
} ncar;

static void load_tables(const nat_t* R, ncar* c, int d) {
    c->eff = R->eff[c->comp][d]; c->opt = R->opt[c->comp][d]; c->pc = R->pc[c->comp][d];
}

/* all cars ordered by (time, driver index) -- what every sorted() of the reference becomes */
static void order_all(const ncar* cars, int n, int* ord) {
    for (int i = 0; i < n; i++) ord[i] = i;
    for (int a = 1; a < n; a++) {
        int v = ord[a];
        float tv = cars[v].t;
        int b = a - 1;
        while (b >= 0 && (cars[ord[b]].t > tv || (cars[ord[b]].t == tv && ord[b] > v))) { ord[b + 1] = ord[b]; b--; }
        ord[b + 1] = v;
    }
}

/* _update_positions :538-560, then re-base every time on the leader */
static void update_positions(ncar* cars, int n, int lap, int drs_until) {
    int ord[N32];
    order_all(cars, n, ord);
    int pred = -1, i = 0;
    float tl = 0.0f;
    int have = 0;
    const int drs_on = lap > 2 && lap > drs_until;
    for (int r = 0; r < n; r++) {
        ncar* c = &cars[ord[r]];
        if (c->dnf) { c->drs = 0; continue; }
        if (!have) { tl = c->t; have = 1; }
        c->pos_live = i++;
        c->drs = pred >= 0 && drs_on && (c->t + -cars[pred].t < 1.0f);
        c->ahead_last = pred >= 0 ? cars[pred].last : 0.0f;
        pred = ord[r];
    }
    if (have) for (int d = 0; d < n; d++) cars[d].t = cars[d].t + -tl;
}

typedef struct { uint8_t position, compound, tire_age, flags; float gap; } trace_rec;

static void emit_trace(trace_rec* trace, int64_t s, int L, int n, int lap, const ncar* cars, const int* pitted, int ev) {
    if (!trace) return;
    for (int d = 0; d < n; d++) {
        trace_rec* r = &trace[((s * L) + (lap - 1)) * n + d];
        r->position = cars[d].dnf ? 0 : (uint8_t)(cars[d].pos_live + 1);
        r->compound = (uint8_t)cars[d].comp;
        r->tire_age = (uint8_t)((int)cars[d].age > 255 ? 255 : (int)cars[d].age);
        r->flags = (uint8_t)((cars[d].dnf ? 1 : 0) | (cars[d].drs ? 2 : 0) | (pitted[d] ? 4 : 0) | ((ev > 3 ? 3 : ev) << 4));
        r->gap = cars[d].t;
    }
}

int orc_run_native(const orc_params* p, uint64_t seed, uint32_t stream, uint64_t sim_begin, int64_t n_sims, int exact,
                   int64_t* hist, uint8_t* finish, float* times, void* trace_out) {
    trace_rec* trace = (trace_rec*)trace_out;
    int pitted[N32];
    if (p->n_drivers < 1 || p->n_drivers > N32) return -1;
    nat_t R;
    derive(p, &R);
    float* op32_tab = build_op32(p);
    const int n = R.n, L = R.L;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    ncar cars[N32];
    u4 w[N32];
    for (int64_t s = 0; s < n_sims; s++) {
        const uint64_t sim = sim_begin + (uint64_t)s;
        const uint32_t s0 = (uint32_t)sim, s1 = (uint32_t)(sim >> 32);
        /* ---- _sample_grid :102-145 ---- */
        if (R.grid_fixed) {
            for (int d = 0; d < n; d++) cars[d].slot = R.fixed_slot[d];
        } else {
            int remaining[N32];
            for (int d = 0; d < N32; d++) remaining[d] = d < n;
            for (int pos = 0; pos < n; pos++) {
                float pr[N32], c[N32], nx[N32];
                for (int i = 0; i < N32; i++) { pr[i] = remaining[i] ? R.G[pos][i] : 0.0f; c[i] = pr[i]; }
                for (int d = 1; d < 32; d <<= 1) { /* the kernel's Hillis-Steele warp scan, same addition tree */
                    for (int i = 0; i < N32; i++) nx[i] = i >= d ? c[i] + c[i - d] : c[i];
                    memcpy(c, nx, sizeof(c));
                }
                const float total = c[31];
                const float u = (float)(philox(s0, s1, (uint32_t)pos, stream, k0, k1).x >> 8) * 5.9604644775390625e-08f;
                int sel = -1;
                if (total > 0.0f) {
                    const float target = u * total;
                    for (int i = 0; i < n && sel < 0; i++) if (remaining[i] && pr[i] > 0.0f && c[i] > target) sel = i;
                    /* u * total rounded up to total: the last remaining driver that has probability mass */
                    if (sel < 0) for (int i = 0; i < n; i++) if (remaining[i] && pr[i] > 0.0f) sel = i;
                } else {
                    int nrem = 0;
                    for (int i = 0; i < n; i++) nrem += remaining[i];
                    int k = (int)(u * (float)nrem);
                    if (k > nrem - 1) k = nrem - 1;
                    for (int i = 0; i < n; i++) if (remaining[i] && k-- == 0) { sel = i; break; }
                }
                cars[sel].slot = pos;
                remaining[sel] = 0;
            }
        }
        /* ---- _initialize_cars :244-273, _simulate_lap_1 :275-311 ---- */
        for (int d = 0; d < n; d++) {
            ncar* c = &cars[d];
            if (R.track == 2) { c->comp = 4; c->age = 0.0f; }
            else if (R.track == 1) { c->comp = 3; c->age = 0.0f; }
            else { c->comp = c->slot < 10 ? 0 : 1; c->age = c->slot < 10 ? 4.0f : 0.0f; }
            c->used = 1u << c->comp;
            load_tables(&R, c, d);
            c->last = 0.0f; c->ahead_last = 0.0f; c->drs = 0; c->pos_live = 0;
            u4 ww = philox(s0, s1, (1u << 8) | (uint32_t)d, stream, k0, k1);
            {
                const float ug = (float)(2u * (ww.w >> 9) + 1u) * 5.9604644775390625e-08f;
                const float ln_u = exact ? exact_log(ug) : logf(ug);
                const float xl = R.dnf_scale[d] <= DNF_NEVER ? 70000.0f : fminf(ln_u * R.dnf_scale[d], 70000.0f);
                c->dnf_lap = 2 + (int)xl;
            }
            c->dnf = ww.x < R.lap1_thr[d];
            if (c->dnf) c->dnf_lap = 1;
            float z1, z2;
            normal2(exact, ww.y, ww.z, &z1, &z2);
            float x = fmaf(c->age, c->eff, c->pc);
            x = fmaf(R.sigma[d], z1, x);
            const float pf = fminf(1.5f, fmaf(0.1f, (float)(c->slot + 1), 0.5f));
            float sd = pf * z2;
            if (c->slot < 3) sd = fminf(sd, 1.0f);
            const float lt = fmaf(-0.5f, sd, x);
            c->t = c->dnf ? -(float)(d + 1) : lt;
            c->age = c->age + 1.0f;
        }
        int drs_until = 0;
        update_positions(cars, n, 1, drs_until);
        memset(pitted, 0, sizeof(pitted));
        emit_trace(trace, s, L, n, 1, cars, pitted, 0);

        float fuel = 0.0f;
        float zn[N32];   /* the second Box-Muller variate, kept for the odd lap of the pair */
        uint32_t u12[N32], u3[N32];
        for (int lap = 2; lap <= L; lap++) {
            const uint32_t pair = (uint32_t)(lap & ~1), odd = (uint32_t)(lap & 1);
            /* ---- events :168-176 ---- */
            /* n <= 20: lane 31's call, words x / y / z; more cars: the spare words y / z / w of driver 0's second call */
            u4 we = philox(s0, s1, (pair << 8) | (n <= 20 ? 31u : 32u), stream, k0, k1);
            if (n > 20) { we.x = we.y; we.y = we.z; we.z = we.w; }
            const uint32_t evw = odd ? we.y : we.x, roll = odd ? we.z >> 16 : we.z & 0xffffu;
            const int ev = evw < R.red_thr ? 1 : evw < R.sc_thr ? 2 : evw < R.vsc_thr ? (roll < 19660u ? 4 : 3) : 0;
            const int rem = L - lap;
            const int nc_rule = R.track == 2 ? 4 : R.track == 1 ? 3 : rem > 30 ? 2 : rem > 15 ? 1 : 0;
            if (ev) {
                for (int d = 0; d < n; d++) {
                    ncar* c = &cars[d];
                    if (c->dnf) continue;
                    if (ev == 1) { /* _handle_red_flag :397-431 */
                        c->t = 0.1f * (float)c->pos_live; c->age = 0.0f; c->comp = nc_rule; c->used |= 1u << c->comp;
                        load_tables(&R, c, d);
                    } else if (ev == 2) { /* _handle_safety_car :334-376 */
                        c->t = 0.5f * (float)c->pos_live; c->age = fmaxf(0.0f, c->age + -1.0f);
                    } else { /* _handle_vsc :378-395 */
                        c->t = c->t * 0.8f;
                        if (ev == 4) c->age = fmaxf(0.0f, c->age + -1.0f);
                    }
                }
                drs_until = ev >= 3 ? lap + 1 : lap + 2;
            }
            /* ---- per-car lap :186-223 ---- */
            fuel = fminf(3.3f, fuel + 0.045f); /* (110 - fuel_load) * 0.03, every runner burns 1.5 kg per lap */
            for (int d = 0; d < n; d++) {
                ncar* c = &cars[d];
                const uint32_t extra = n <= 20 ? (d < 10 ? philox(s0, s1, (pair << 8) | (uint32_t)(20 + d), stream, k0, k1).x
                                                         : philox(s0, s1, (pair << 8) | (uint32_t)(10 + d), stream, k0, k1).y)
                                               : philox(s0, s1, (pair << 8) | (uint32_t)(32 + d), stream, k0, k1).x;
                float z;
                if (!odd) {
                    w[d] = philox(s0, s1, (pair << 8) | (uint32_t)d, stream, k0, k1);
                    normal2(exact, w[d].x, w[d].y, &z, &zn[d]);
                    u12[d] = w[d].z;
                    u3[d] = extra & 0xffffu;
                } else {
                    z = zn[d];
                    u12[d] = w[d].w;
                    u3[d] = extra >> 16;
                }
                if (lap >= c->dnf_lap) c->dnf = 1;
                if (c->dnf) { c->age = c->age + 1.0f; continue; } /* (the kernel lets retired cars age on; never used) */
                float x = fmaf(c->age, c->eff, c->pc);
                x = x + -fuel;
                x = x + -(c->drs ? R.drs_delta : 0.0f);
                const float clean = fmaf(R.sigma[d], z, x);
                float lt = clean;
                /* ahead_last is 0 for the leader: the reference's `time_behind_leader > 0` is implied */
                if (c->ahead_last > 0.0f && c->t < R.dirty_thr) lt = fmaxf(clean + R.dirty_pen, c->ahead_last);
                c->t = c->t + lt;
                c->last = lt;
                c->age = c->age + 1.0f;
            }
            /* ---- _handle_pit_stops :433-494 ---- */
            memset(pitted, 0, sizeof(pitted));
            for (int d = 0; d < n; d++) {
                ncar* c = &cars[d];
                if (c->dnf || !(rem > 5) || !(c->age > c->opt)) continue;
                pitted[d] = 1;
                c->t = c->t + R.pit_loss;
                int nc = nc_rule;
                const uint32_t ud = c->used & 7u;
                if (R.track == 0 && __builtin_popcount(ud) == 1 && ((ud >> nc) & 1u)) {
                    const uint32_t avail = 7u & ~ud;
                    if (rem > 20) nc = (avail & 2u) ? 1 : p->pop_no_medium;
                    else nc = (avail & 1u) ? 0 : p->pop_no_soft;
                }
                c->comp = nc; c->used |= 1u << nc; c->age = 0.0f;
                load_tables(&R, c, d);
            }
            /* ---- _simulate_overtakes :496-536 ---- */
            float op[N32];
            double P64[N32];
            /* The pair test `pace_delta > overtake_delta` (:514-521) is decided in FP64, op for op as upstream (with
             * round-number inputs it lands exactly on the threshold for some (driver, age) pairs and rounding decides):
             * the kernel reads the same decision from a host-built rank table.  The x 2^15 FP32 paces only feed the
             * (continuous) probability: the 16-bit uniform compares against min(32768, delta) directly. */
            for (int d = 0; d < n; d++) {
                const double wear = (double)(int)cars[d].age * p->tire_deg[d];
                P64[d] = p->base_pace[d] + wear;
                op[d] = cars[d].dnf ? NAN : op32_tab[(size_t)(int)cars[d].age * n + d];
            }
            for (int pass = 0; pass < 3; pass++) {
                int ord[N32], succ[N32 + 1], any = 0;
                float T[N32];
                order_all(cars, n, ord);
                memset(succ, 0, sizeof(succ));
                for (int r = 0; r < n; r++) T[r] = cars[ord[r]].t;
                for (int r = 1; r < n; r++) {
                    const int b = ord[r], a = ord[r - 1];
                    const float opb = cars[b].drs ? op[b] + -(R.drs_delta * 32768.0f) : op[b];
                    const float delta = op[a] + -opb;
                    const uint32_t u16 = pass == 0 ? (u12[b] & 0xffffu) : pass == 1 ? (u12[b] >> 16) : u3[b];
                    /* u16 * 2^-16 < min(0.5, delta / 2), both sides x 2^16 */
                    double pace_delta = P64[a] - P64[b];
                    if (cars[b].drs) pace_delta = pace_delta + p->drs_delta;
                    succ[r] = !cars[a].dnf && !cars[b].dnf && pace_delta > p->overtake_delta && (float)u16 < fminf(32768.0f, delta);
                    any |= succ[r];
                }
                if (!any) break;
                for (int r = 0; r < n; r++) { /* :528-530 over runs of consecutive successes, closed form */
                    int k = 0;
                    while (succ[r - k]) k++; /* succ[0] == 0 stops the scan */
                    const int sn = succ[r + 1];
                    if (k + sn > 0) cars[ord[r]].t = fmaf(-0.1f, (float)(k - 2 * sn), T[r - k]);
                }
            }
            update_positions(cars, n, lap, drs_until);
            emit_trace(trace, s, L, n, lap, cars, pitted, ev);
        }
        /* ---- final classification :231-242 ---- */
        int n_live = 0;
        for (int d = 0; d < n; d++) n_live += !cars[d].dnf;
        for (int d = 0; d < n; d++) {
            const ncar* c = &cars[d];
            int pos;
            if (!c->dnf) pos = c->pos_live;
            else {
                const float tc = c->dnf_lap == 1 ? 0.0f : c->t;
                int worse = 0;
                for (int j = 0; j < n; j++) {
                    const ncar* o = &cars[j];
                    if (!o->dnf || j == d) continue;
                    const float tj = o->dnf_lap == 1 ? 0.0f : o->t;
                    if (o->dnf_lap > c->dnf_lap || (o->dnf_lap == c->dnf_lap && (tj > tc || (tj == tc && o->slot < c->slot)))) worse++;
                }
                pos = n_live + worse;
            }
            if (hist) hist[(int64_t)d * n + pos] += 1;
            if (finish) finish[s * n + pos] = (uint8_t)d;
            if (times) times[s * n + d] = c->t;
        }
    }
    free(op32_tab);
    return 0;
}
