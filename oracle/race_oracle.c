/* TEST INFRASTRUCTURE -- see race_oracle.h.  Scalar FP64 restatement of the reference hot path
 * (/root/reference/src/simulation.py; every function cites the lines it follows).
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared (oracle/Makefile).
 * No FMA contraction, strict left-to-right evaluation: results are bit-identical to CPython's. */
#include "race_oracle.h"

#include <math.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * MT19937 (Matsumoto & Nishimura 1998) -- shared by CPython `_random` and NumPy legacy RandomState
 * ------------------------------------------------------------------------------------------ */
static void mt_init_genrand(orc_mt* s, uint32_t seed) {
    s->mt[0] = seed;
    for (int i = 1; i < 624; i++)
        s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
    s->idx = 624;
}

static void mt_init_by_array(orc_mt* s, const uint32_t* key, int len) {
    mt_init_genrand(s, 19650218u);
    int i = 1, j = 0;
    int k = 624 > len ? 624 : len;
    for (; k; k--) {
        s->mt[i] = (s->mt[i] ^ ((s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
        i++; j++;
        if (i >= 624) { s->mt[0] = s->mt[623]; i = 1; }
        if (j >= len) j = 0;
    }
    for (k = 623; k; k--) {
        s->mt[i] = (s->mt[i] ^ ((s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
        i++;
        if (i >= 624) { s->mt[0] = s->mt[623]; i = 1; }
    }
    s->mt[0] = 0x80000000u;
    s->idx = 624;
}

static uint32_t mt_next(orc_mt* s) {
    if (s->idx >= 624) {
        uint32_t* mt = s->mt;
        for (int kk = 0; kk < 624; kk++) {
            uint32_t y = (mt[kk] & 0x80000000u) | (mt[(kk + 1) % 624] & 0x7fffffffu);
            mt[kk] = mt[(kk + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        s->idx = 0;
    }
    uint32_t y = s->mt[s->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* genrand_res53: 53-bit double in [0,1) -- random.random() and NumPy's legacy double alike */
static double mt_res53(orc_mt* s) {
    uint32_t a = mt_next(s) >> 5, b = mt_next(s) >> 6;
    return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
}

void orc_rng_seed(orc_rng* rng, uint64_t seed) {
    /* random.seed(int): init_by_array over the 32-bit little-endian words of abs(seed) */
    uint32_t key[2] = {(uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32)};
    mt_init_by_array(&rng->py, key, key[1] ? 2 : 1);
    /* np.random.seed(int): init_genrand(seed), Gaussian cache cleared (_legacy_seeding) */
    mt_init_genrand(&rng->np, (uint32_t)seed);
    rng->has_gauss = 0;
    rng->gauss = 0.0;
}

double orc_py_random(orc_rng* rng) { return mt_res53(&rng->py); }
double orc_np_random_sample(orc_rng* rng) { return mt_res53(&rng->np); }

/* legacy_gauss (numpy/random/src/legacy/legacy-distributions.c): Marsaglia polar, caches f*x1 */
double orc_np_standard_normal(orc_rng* rng) {
    if (rng->has_gauss) {
        double t = rng->gauss;
        rng->has_gauss = 0;
        rng->gauss = 0.0;
        return t;
    }
    double f, x1, x2, r2;
    do {
        x1 = 2.0 * mt_res53(&rng->np) - 1.0;
        x2 = 2.0 * mt_res53(&rng->np) - 1.0;
        r2 = x1 * x1 + x2 * x2;
    } while (r2 >= 1.0 || r2 == 0.0);
    f = sqrt(-2.0 * log(r2) / r2);
    rng->gauss = f * x1;
    rng->has_gauss = 1;
    return f * x2;
}

/* ------------------------------------------------------------------------------------------
 * draw source: MT streams (optionally logged to tapes) or explicit tapes
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    orc_rng* rng; /* NULL => tape mode */
    const double *upy, *z, *unp;
    int64_t i_py, i_z, i_np;       /* tape cursors */
    int64_t e_py, e_z, e_np;       /* tape ends for the current sim */
    int64_t n_py, n_z, n_np;       /* cumulative draw counts */
    orc_outputs* out;
    int err;
} draw_src;

static double draw_py(draw_src* s) {
    double v;
    if (s->rng) {
        v = mt_res53(&s->rng->py);
        if (s->out && s->out->log_upy) {
            if (s->n_py < s->out->cap_upy) s->out->log_upy[s->n_py] = v; else s->err = -2;
        }
    } else {
        if (s->i_py >= s->e_py) { s->err = -3; v = 0.5; } else v = s->upy[s->i_py++];
    }
    s->n_py++;
    return v;
}

static double draw_z(draw_src* s) {
    double v;
    if (s->rng) {
        v = orc_np_standard_normal(s->rng);
        if (s->out && s->out->log_z) {
            if (s->n_z < s->out->cap_z) s->out->log_z[s->n_z] = v; else s->err = -2;
        }
    } else {
        if (s->i_z >= s->e_z) { s->err = -3; v = 0.0; } else v = s->z[s->i_z++];
    }
    s->n_z++;
    return v;
}

static double draw_np(draw_src* s) {
    double v;
    if (s->rng) {
        v = mt_res53(&s->rng->np);
        if (s->out && s->out->log_unp) {
            if (s->n_np < s->out->cap_unp) s->out->log_unp[s->n_np] = v; else s->err = -2;
        }
    } else {
        if (s->i_np >= s->e_np) { s->err = -3; v = 0.5; } else v = s->unp[s->i_np++];
    }
    s->n_np++;
    return v;
}

/* np.random.normal(0, scale) == loc + scale * legacy_gauss (legacy_normal) */
static double normal0(draw_src* s, double scale) { return 0.0 + scale * draw_z(s); }

/* ------------------------------------------------------------------------------------------
 * CPython 3.12 builtin sum() (Python/bltinmodule.c builtin_sum_impl), SURVEY Q12
 * ------------------------------------------------------------------------------------------ */
double orc_py_sum(const double* v, const uint8_t* kind, int n, int* result_kind) {
    int i = 0;
    while (i < n && kind[i] == ORC_ITEM_INT0) i++; /* exact-int fast path: 0 + 0 + ... */
    if (i == n) { *result_kind = ORC_ITEM_INT0; return 0.0; }
    double r = 0.0 + v[i]; /* PyNumber_Add(int 0, item) */
    int k = kind[i];
    i++;
    if (k == ORC_ITEM_FLOAT) {
        double c = 0.0; /* Neumaier compensation */
        int fell_out = 0;
        for (; i < n; i++) {
            if (kind[i] == ORC_ITEM_FLOAT) {
                double x = v[i];
                double t = r + x;
                if (fabs(r) >= fabs(x)) c += (r - t) + x; else c += (x - t) + r;
                r = t;
            } else if (kind[i] == ORC_ITEM_INT0) {
                r += 0.0; /* PyLong item inside the float loop: f_result += (double)value */
            } else {
                if (c != 0.0 && isfinite(c)) r += c;
                r = r + v[i]; /* float + np.float64 -> np.float64: leaves the float loop */
                i++;
                fell_out = 1;
                break;
            }
        }
        if (!fell_out) {
            if (c != 0.0 && isfinite(c)) r += c;
            *result_kind = ORC_ITEM_FLOAT;
            return r;
        }
    }
    for (; i < n; i++) r = r + (kind[i] == ORC_ITEM_INT0 ? 0.0 : v[i]); /* generic PyNumber_Add loop */
    *result_kind = ORC_ITEM_NPFLOAT;
    return r;
}

/* ------------------------------------------------------------------------------------------
 * race state
 * ------------------------------------------------------------------------------------------ */
typedef struct { /* CarState, src/simulation.py:9-34 */
    int drv;              /* index into the per-driver parameter arrays */
    int position, lap, compound, tire_age, pit_stops, drs, dnf, laps_completed;
    unsigned used;        /* bit mask of compounds used */
    double fuel, tbl, cum, last_lap;
} car_t;

/* stable ascending sort of car indices by cumulative_time (Python sorted()/list.sort are stable) */
static void sort_by_time(const car_t* cars, int* idx, int m) {
    for (int a = 1; a < m; a++) {
        int v = idx[a];
        double key = cars[v].cum;
        int b = a - 1;
        while (b >= 0 && cars[idx[b]].cum > key) { idx[b + 1] = idx[b]; b--; }
        idx[b + 1] = v;
    }
}

static int active_sorted(const car_t* cars, int n, int* idx) {
    int m = 0;
    for (int i = 0; i < n; i++) if (!cars[i].dnf) idx[m++] = i;
    sort_by_time(cars, idx, m);
    return m;
}

/* _sample_grid, src/simulation.py:102-145 (+ RandomState.choice restated) */
static void sample_grid(const orc_params* p, draw_src* ds, int* grid) {
    int n = p->n_drivers;
    int remaining[ORC_MAX_DRIVERS];
    int n_rem = n;
    for (int i = 0; i < n; i++) remaining[i] = 1;
    double probs[ORC_MAX_DRIVERS];
    uint8_t kind[ORC_MAX_DRIVERS];
    for (int pos = 0; pos < n; pos++) {
        for (int d = 0; d < n; d++) { /* :119-122 */
            if (remaining[d] && p->grid_kind[d][pos] != ORC_ITEM_INT0) {
                probs[d] = p->grid_probs[d][pos]; kind[d] = p->grid_kind[d][pos];
            } else { probs[d] = 0.0; kind[d] = ORC_ITEM_INT0; }
        }
        int tk;
        double total = orc_py_sum(probs, kind, n, &tk); /* :123 */
        if (total > 0) { /* :125-126 */
            for (int d = 0; d < n; d++) {
                probs[d] = probs[d] / total;
                kind[d] = (tk == ORC_ITEM_NPFLOAT || kind[d] == ORC_ITEM_NPFLOAT) ? ORC_ITEM_NPFLOAT : ORC_ITEM_FLOAT;
            }
        } else { /* :127-130 uniform over the remaining drivers */
            for (int d = 0; d < n; d++) {
                if (remaining[d]) { probs[d] = 1.0 / n_rem; kind[d] = ORC_ITEM_FLOAT; }
                else { probs[d] = 0.0; kind[d] = ORC_ITEM_INT0; }
            }
        }
        double prob_sum = orc_py_sum(probs, kind, n, &tk); /* :133 */
        if (prob_sum > 0 && fabs(prob_sum - 1.0) > 1e-9) /* :134-135 */
            for (int d = 0; d < n; d++) probs[d] = probs[d] / prob_sum;
        /* np.random.choice(drivers, p=probs) :137 */
        double cdf[ORC_MAX_DRIVERS];
        double acc = probs[0];
        cdf[0] = acc;
        for (int d = 1; d < n; d++) { acc = acc + probs[d]; cdf[d] = acc; } /* p.cumsum() */
        double last = cdf[n - 1];
        for (int d = 0; d < n; d++) cdf[d] = cdf[d] / last;
        double u = draw_np(ds);
        int sel = 0;
        while (sel < n && cdf[sel] <= u) sel++; /* searchsorted(u, side='right') */
        if (sel >= n) sel = n - 1; /* cannot happen for u < 1 = cdf[-1] */
        grid[pos] = sel;
        if (remaining[sel]) { remaining[sel] = 0; n_rem--; } /* :139 */
    }
}

/* _calculate_lap_time, src/simulation.py:313-332 */
static double lap_time(const orc_params* p, draw_src* ds, const car_t* car) {
    int d = car->drv;
    double base = p->base_pace[d], deg = p->tire_deg[d], variance = p->variance[d];
    double compound_deg = p->compound_deg_rate[car->compound];
    double driver_factor = deg > 0 ? deg / 0.05 : 1.0;
    double effective_deg = compound_deg * driver_factor;
    double tire_effect = (double)car->tire_age * effective_deg;
    double fuel_effect = (110.0 - car->fuel) * 0.03;
    double compound_delta = p->compound_pace_delta[car->compound];
    double drs_gain = car->drs ? p->drs_delta : 0.0;
    double noise = normal0(ds, variance);
    return base + tire_effect - fuel_effect + compound_delta - drs_gain + noise; /* :332 left to right */
}

/* _update_positions, src/simulation.py:538-560 */
static void update_positions(car_t* cars, int n, int lap, int drs_disabled) {
    int idx[ORC_MAX_DRIVERS];
    int m = active_sorted(cars, n, idx);
    for (int i = 0; i < m; i++) {
        car_t* c = &cars[idx[i]];
        c->position = i + 1;
        c->tbl = c->cum - cars[idx[0]].cum;
        if (lap <= 2 || drs_disabled || i == 0) c->drs = 0;
        else c->drs = (c->cum - cars[idx[i - 1]].cum) < 1.0;
    }
}

static double fuel_burn(double fuel) { /* max(0, fuel - 1.5) :221,:309 */
    double f = fuel - 1.5;
    return f > 0 ? f : 0.0;
}

/* _handle_safety_car, src/simulation.py:334-376 */
static void handle_safety_car(car_t* cars, int n) {
    int idx[ORC_MAX_DRIVERS];
    int m = active_sorted(cars, n, idx);
    if (!m) return;
    double leader_time = cars[idx[0]].cum;
    int leader_laps = cars[idx[0]].laps_completed;
    for (int i = 0; i < m; i++) {
        car_t* c = &cars[idx[i]];
        int laps_down = leader_laps - c->laps_completed;
        if (laps_down <= 0) c->cum = leader_time + (double)i * 0.5;
        else c->cum = leader_time + ((double)laps_down * 90.0) + (double)i * 0.5; /* dead code in practice (Q6) */
        c->tbl = c->cum - leader_time;
        c->tire_age = c->tire_age - 1 > 0 ? c->tire_age - 1 : 0;
    }
}

/* _handle_vsc, src/simulation.py:378-395 */
static void handle_vsc(car_t* cars, int n, draw_src* ds) {
    int idx[ORC_MAX_DRIVERS];
    int m = active_sorted(cars, n, idx);
    if (!m) return; /* early return BEFORE the extra draw :381-382 */
    double leader_time = cars[idx[0]].cum;
    for (int i = 0; i < m; i++) {
        car_t* c = &cars[idx[i]];
        double gap = c->cum - leader_time;
        c->cum = leader_time + gap * 0.8;
        c->tbl = c->cum - leader_time;
    }
    if (draw_py(ds) < 0.3) /* :392 */
        for (int i = 0; i < m; i++) {
            car_t* c = &cars[idx[i]];
            c->tire_age = c->tire_age - 1 > 0 ? c->tire_age - 1 : 0;
        }
}

/* _handle_red_flag, src/simulation.py:397-431 */
static void handle_red_flag(const orc_params* p, car_t* cars, int n, int lap) {
    int idx[ORC_MAX_DRIVERS];
    int m = active_sorted(cars, n, idx);
    if (!m) return;
    double leader_time = cars[idx[0]].cum;
    int remaining = p->total_laps - lap;
    for (int i = 0; i < m; i++) {
        car_t* c = &cars[idx[i]];
        c->cum = leader_time + (double)i * 0.1;
        c->tbl = c->cum - leader_time;
        c->tire_age = 0;
        if (p->track_condition == ORC_WETTRACK) c->compound = ORC_WET;
        else if (p->track_condition == ORC_DAMP) c->compound = ORC_INTER;
        else if (remaining > 30) c->compound = ORC_HARD;
        else if (remaining > 15) c->compound = ORC_MEDIUM;
        else c->compound = ORC_SOFT;
        c->used |= 1u << c->compound;
    }
}

/* _handle_pit_stops, src/simulation.py:433-494 */
static void handle_pit_stops(const orc_params* p, car_t* cars, int n, int lap) {
    int remaining = p->total_laps - lap;
    const unsigned dry = (1u << ORC_SOFT) | (1u << ORC_MEDIUM) | (1u << ORC_HARD);
    int is_wet = p->track_condition != ORC_DRY;
    for (int i = 0; i < n; i++) {
        car_t* c = &cars[i];
        if (c->dnf) continue;
        double optimal = p->compound_optimal[c->compound];
        double driver_deg = p->tire_deg_pit[c->drv];
        if (driver_deg > 0.05) optimal = trunc(optimal * 0.85);      /* int(optimal_laps * 0.85) */
        else if (driver_deg < 0.02) optimal = trunc(optimal * 1.1);  /* int(optimal_laps * 1.1)  */
        if ((double)c->tire_age > optimal && remaining > 5) {
            c->cum += p->pit_loss;
            int nc;
            if (p->track_condition == ORC_WETTRACK) nc = ORC_WET;
            else if (p->track_condition == ORC_DAMP) nc = ORC_INTER;
            else if (remaining > 30) nc = ORC_HARD;
            else if (remaining > 15) nc = ORC_MEDIUM;
            else nc = ORC_SOFT;
            unsigned used_dry = c->used & dry;
            if (__builtin_popcount(used_dry) == 1 && (used_dry & (1u << nc)) && !is_wet) {
                unsigned avail = dry & ~used_dry;
                if (remaining > 20) {
                    if (avail & (1u << ORC_MEDIUM)) nc = ORC_MEDIUM;
                    else nc = p->pop_no_medium;          /* available.pop() :486 (Q1) */
                } else {
                    if (avail & (1u << ORC_SOFT)) nc = ORC_SOFT;
                    else nc = p->pop_no_soft;            /* available.pop() :488 (Q1) */
                }
            }
            c->compound = nc;
            c->used |= 1u << nc;
            c->tire_age = 0;
            c->pit_stops++;
        }
    }
}

/* _simulate_overtakes, src/simulation.py:496-536 */
static void simulate_overtakes(const orc_params* p, car_t* cars, int n, draw_src* ds) {
    for (int pass = 0; pass < 3; pass++) {
        int occurred = 0;
        int idx[ORC_MAX_DRIVERS];
        for (int i = 0; i < n; i++) idx[i] = i;
        sort_by_time(cars, idx, n); /* ALL cars, DNF included (Q5) */
        for (int i = 1; i < n; i++) {
            car_t* behind = &cars[idx[i]];
            car_t* ahead = &cars[idx[i - 1]];
            if (behind->dnf || ahead->dnf) continue;
            double pace_behind = p->base_pace[behind->drv] + (double)behind->tire_age * p->tire_deg[behind->drv];
            double pace_ahead = p->base_pace[ahead->drv] + (double)ahead->tire_age * p->tire_deg[ahead->drv];
            double pace_delta = pace_ahead - pace_behind;
            if (behind->drs) pace_delta += p->drs_delta;
            if (pace_delta > p->overtake_delta) {
                double prob = pace_delta / 2.0;
                if (prob > 0.5) prob = 0.5; /* min(0.5, pace_delta / 2.0) */
                if (draw_py(ds) < prob) {
                    double nb = ahead->cum - 0.1;
                    if (!(nb > 0.1)) nb = 0.1; /* max(0.1, ahead - 0.1) */
                    behind->cum = nb;
                    ahead->cum = nb + 0.3;
                    occurred = 1;
                }
            }
        }
        if (!occurred) break;
    }
}

/* ---- optional per-lap trace of the FP64 race (debug aid for the statistical parity tests; single-threaded) ---- */
typedef struct { uint8_t position, compound, tire_age, flags; float gap; } orc_trace_rec;
static orc_trace_rec* g_trace = NULL;
static int64_t g_trace_cap = 0, g_trace_sim = 0;
void orc_debug_trace(void* buf, int64_t cap_sims) { g_trace = (orc_trace_rec*)buf; g_trace_cap = cap_sims; g_trace_sim = 0; }
static void trace_lap(const orc_params* p, const car_t* cars, int n, int lap) {
    if (!g_trace || g_trace_sim >= g_trace_cap) return;
    double lead = 0.0; int have = 0;
    for (int i = 0; i < n; i++) if (!cars[i].dnf && (!have || cars[i].cum < lead)) { lead = cars[i].cum; have = 1; }
    for (int i = 0; i < n; i++) {
        orc_trace_rec* r = &g_trace[((g_trace_sim * p->total_laps) + (lap - 1)) * n + cars[i].drv];
        r->position = cars[i].dnf ? 0 : (uint8_t)cars[i].position;
        r->compound = (uint8_t)cars[i].compound;
        r->tire_age = (uint8_t)cars[i].tire_age;
        r->flags = (uint8_t)((cars[i].dnf ? 1 : 0) | (cars[i].drs ? 2 : 0));
        r->gap = (float)(cars[i].cum - lead);
    }
}

/* simulate_race, src/simulation.py:147-242 (+ _initialize_cars :244-273, _simulate_lap_1 :275-311) */
static void simulate_race(const orc_params* p, draw_src* ds, const int* grid, car_t* cars, int* finish) {
    int n = p->n_drivers;
    for (int pos = 0; pos < n; pos++) { /* _initialize_cars */
        car_t* c = &cars[pos];
        memset(c, 0, sizeof(*c));
        c->drv = grid[pos];
        c->position = pos + 1;
        if (p->track_condition == ORC_WETTRACK) { c->compound = ORC_WET; c->tire_age = 0; }
        else if (p->track_condition == ORC_DAMP) { c->compound = ORC_INTER; c->tire_age = 0; }
        else { c->compound = pos < 10 ? ORC_SOFT : ORC_MEDIUM; c->tire_age = pos < 10 ? 4 : 0; }
        c->fuel = 110.0;
        c->used = 1u << c->compound; /* __post_init__ :31-34 */
    }
    for (int i = 0; i < n; i++) { /* _simulate_lap_1 */
        car_t* c = &cars[i];
        double base_dnf_rate = p->team_rate[c->drv];
        if (draw_py(ds) < base_dnf_rate * 4.0) { c->dnf = 1; c->lap = 1; continue; }
        double base_lap = lap_time(p, ds, c);
        double pf = 0.5 + (double)c->position * 0.1;
        if (pf > 1.5) pf = 1.5; /* min(1.5, ...) */
        double start_delta = normal0(ds, pf);
        if (c->position <= 3 && start_delta > 1.0) start_delta = 1.0; /* min(start_delta, 1.0) */
        double lt = base_lap - start_delta * 0.5;
        c->cum += lt;
        c->tire_age += 1;
        c->fuel = fuel_burn(c->fuel);
        c->lap = 1;
    }
    update_positions(cars, n, 1, 1);
    trace_lap(p, cars, n, 1);
    int drs_disabled_until = 0;

    for (int lap = 2; lap <= p->total_laps; lap++) { /* :166 */
        if (draw_py(ds) < p->red_p) { handle_red_flag(p, cars, n, lap); drs_disabled_until = lap + 2; }
        else if (draw_py(ds) < p->sc_p) { handle_safety_car(cars, n); drs_disabled_until = lap + 2; }
        else if (draw_py(ds) < p->vsc_p) { handle_vsc(cars, n, ds); drs_disabled_until = lap + 1; }

        int idx[ORC_MAX_DRIVERS];
        int m = active_sorted(cars, n, idx); /* :179 */
        double ahead_lap[ORC_MAX_DRIVERS];
        int has_ahead[ORC_MAX_DRIVERS];
        for (int i = 0; i < n; i++) { has_ahead[i] = 0; ahead_lap[i] = 0.0; }
        for (int i = 1; i < m; i++) { ahead_lap[idx[i]] = cars[idx[i - 1]].last_lap; has_ahead[idx[i]] = 1; }

        for (int i = 0; i < n; i++) { /* :186-223, grid order */
            car_t* c = &cars[i];
            if (c->dnf) continue;
            if (draw_py(ds) < p->dnf_rate[c->drv]) { c->dnf = 1; c->lap = lap; continue; }
            double clean = lap_time(p, ds, c);
            double lt = clean;
            if (c->tbl > 0) {
                double car_ahead_lap = has_ahead[i] ? ahead_lap[i] : 0.0;
                if (car_ahead_lap > 0 && c->tbl < p->dirty_thr) {
                    double dirty = clean + p->dirty_pen;
                    lt = dirty >= car_ahead_lap ? dirty : car_ahead_lap; /* max(dirty, car_ahead_lap) */
                }
            }
            c->cum += lt;
            c->last_lap = lt;
            c->tire_age += 1;
            c->fuel = fuel_burn(c->fuel);
            c->lap = lap;
            c->laps_completed += 1;
        }
        handle_pit_stops(p, cars, n, lap);
        simulate_overtakes(p, cars, n, ds);
        update_positions(cars, n, lap, lap <= drs_disabled_until);
        trace_lap(p, cars, n, lap);
    }
    if (g_trace) g_trace_sim++;

    /* final classification :231-242 */
    int idx[ORC_MAX_DRIVERS];
    int m = active_sorted(cars, n, idx);
    int k = 0;
    for (int i = 0; i < m; i++) finish[k++] = idx[i];
    int dn[ORC_MAX_DRIVERS], nd = 0;
    for (int i = 0; i < n; i++) if (cars[i].dnf) dn[nd++] = i;
    /* sorted(key=(lap, cumulative_time), reverse=True): stable, descending */
    for (int a = 1; a < nd; a++) {
        int v = dn[a];
        int b = a - 1;
        while (b >= 0) {
            const car_t *x = &cars[dn[b]], *y = &cars[v];
            int less = x->lap < y->lap || (x->lap == y->lap && x->cum < y->cum);
            if (!less) break;
            dn[b + 1] = dn[b];
            b--;
        }
        dn[b + 1] = v;
    }
    for (int i = 0; i < nd; i++) finish[k++] = dn[i];
}

static void emit(const orc_params* p, orc_outputs* out, int64_t s, const int* grid, const car_t* cars,
                 const int* finish, const draw_src* ds) {
    int n = p->n_drivers;
    if (!out) return;
    for (int pos = 0; pos < n; pos++) {
        int drv = cars[finish[pos]].drv;
        if (out->hist) out->hist[(int64_t)drv * n + pos] += 1;
        if (out->finish) out->finish[s * n + pos] = (uint8_t)drv;
    }
    for (int i = 0; i < n; i++) {
        if (out->grid) out->grid[s * n + i] = (uint8_t)grid[i];
        if (out->times) out->times[s * n + cars[i].drv] = cars[i].cum;
        if (out->dnf_lap) out->dnf_lap[s * n + cars[i].drv] = (int16_t)(cars[i].dnf ? cars[i].lap : 0);
    }
    if (out->draws) {
        out->draws[s * 3 + 0] = ds->n_py;
        out->draws[s * 3 + 1] = ds->n_z;
        out->draws[s * 3 + 2] = ds->n_np;
    }
}

int orc_run_streams(const orc_params* p, orc_rng* rng, int64_t n_sims, orc_outputs* out) {
    if (p->n_drivers < 1 || p->n_drivers > ORC_MAX_DRIVERS) return -1;
    draw_src ds;
    memset(&ds, 0, sizeof(ds));
    ds.rng = rng;
    ds.out = out;
    int grid[ORC_MAX_DRIVERS], finish[ORC_MAX_DRIVERS];
    car_t cars[ORC_MAX_DRIVERS];
    for (int64_t s = 0; s < n_sims; s++) {
        sample_grid(p, &ds, grid);
        simulate_race(p, &ds, grid, cars, finish);
        emit(p, out, s, grid, cars, finish, &ds);
        if (ds.err) return ds.err;
    }
    return 0;
}

int orc_run_tapes(const orc_params* p, int64_t n_sims, const double* upy, const double* z,
                  const double* unp, const int64_t* off, orc_outputs* out) {
    if (p->n_drivers < 1 || p->n_drivers > ORC_MAX_DRIVERS) return -1;
    draw_src ds;
    memset(&ds, 0, sizeof(ds));
    ds.upy = upy; ds.z = z; ds.unp = unp;
    ds.out = out;
    int grid[ORC_MAX_DRIVERS], finish[ORC_MAX_DRIVERS];
    car_t cars[ORC_MAX_DRIVERS];
    for (int64_t s = 0; s < n_sims; s++) {
        ds.i_py = off[3 * s]; ds.i_z = off[3 * s + 1]; ds.i_np = off[3 * s + 2];
        ds.e_py = off[3 * s + 3]; ds.e_z = off[3 * s + 4]; ds.e_np = off[3 * s + 5];
        sample_grid(p, &ds, grid);
        simulate_race(p, &ds, grid, cars, finish);
        emit(p, out, s, grid, cars, finish, &ds);
        if (ds.err) return ds.err;
    }
    return 0;
}
