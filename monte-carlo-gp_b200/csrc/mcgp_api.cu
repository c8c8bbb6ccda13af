// C ABI of libmcgp.so (declared in include/mcgp.h): context management, host-side derivation of the
// device parameter blocks from mcgp_race_params, and the launches.  No torch types, no CPU fallback.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "../../include/mcgp.h"
#include "device_params.h"

namespace mcgp {
cudaError_t launch_native(const NativeRace* races_dev, const PacePair* pace_dev, int pace_rows, int pace_stride, int n_races,
                          int max_n, unsigned long long n_sims,
                          unsigned long long sim_begin, unsigned long long seed, bool exact,
                          unsigned long long* hist, uint8_t* finish, float* times, TraceRecord* trace,
                          unsigned long long trace_first, unsigned long long trace_count, unsigned long long* laphist,
                          unsigned long long* work_counter, int sm_count, cudaStream_t st, uint8_t* grid_out = nullptr);
cudaError_t launch_replay(const ReplayRace* race_dev, int n_drivers, unsigned long long n_sims, const double* u_py, const double* z,
                          const double* u_np, const long long* off, unsigned long long* hist, uint8_t* finish,
                          double* times, int16_t* dnf_lap, uint8_t* grid, long long* used, int* status,
                          unsigned long long* work_counter, int sm_count, cudaStream_t st, bool serial_grid);
cudaError_t launch_season_grid(NativeRace* race_dev, const double* quali_dev, const int32_t* penalty_dev, const double* teammate_dev,
                               const double* form_dev, const double* circuit_dev, int n, double* rows_dev, cudaStream_t st);
cudaError_t launch_season_elo(const uint8_t* grid_order, const uint8_t* finish_order, const double* q_before, const double* r_before,
                              double* q_after, double* r_after, int n, double k, cudaStream_t st);
cudaError_t launch_season_actual(const uint8_t* finish_order, int32_t* winner, int32_t* podium, int n, cudaStream_t st);
cudaError_t launch_score_counts(const unsigned long long* hist, int n_races, int n, unsigned long long n_sims, const int32_t* winner,
                                const int32_t* podium, unsigned long long* tallies, double* brier, int32_t* podium_hits,
                                double* calib, int32_t* calib_bins, cudaStream_t st);
int native_philox_rounds();
size_t pace_pairs_per_race(int rows, int stride);  // device form of one race's overtake pace tables (device_params.h: PacePair)
}  // namespace mcgp

struct mcgp_context {
    int device = 0;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    std::string err;
    NativeRace* native_dev = nullptr;
    PacePair* pace_dev = nullptr;      // [race][drs][pace_rows][pace_stride] overtake pace tables (+ padding per race)
    int pace_rows = 0, pace_stride = 0;
    int cap_races = 0;                 // allocated capacity of native_dev / replay_dev / work_counter (races)
    size_t cap_pace = 0;               // allocated capacity of pace_dev (entries)
    ReplayRace* replay_dev = nullptr;
    unsigned long long* work_counter = nullptr;  // one claim counter per race of the batch (dynamic sim distribution)
    int n_races = 0, n_drivers = 0;    // the uploaded batch; n_races == 0: nothing (valid) is resident
    bool replay_ready = false;         // replay_dev holds the blocks of the resident batch (derived lazily, see ensure_replay)
    bool uniform_laps = true;          // every race of the resident batch has the same total_laps
    bool replay_serial_grid = false;   // mcgp_replay_serial_grid: every _sample_grid position on the serial path
    int launches = 0;
    uint64_t upload_bytes = 0;
    std::vector<mcgp_race_params> resident;  // host copy of the resident batch: a repeated call with identical
                                             // parameters (the product calls the same race again and again) skips
                                             // the derivation and the upload
    // Launches on one handle share the claim counters and the parameter blocks, so they are ordered against each other
    // (and re-uploads against them) through this event, whatever stream the caller passes.
    cudaEvent_t last_launch = nullptr;
    bool launched = false;
    cudaStream_t own_stream = nullptr;  // host-buffer entry points: async copies + kernel + ONE synchronisation
    void* pinned = nullptr;             // grow-only pinned staging block of the host-buffer entry points
    size_t pinned_sz = 0;
    // grow-only scratch for the host-buffer entry points
    void* scratch[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t scratch_sz[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

static thread_local std::string g_create_error;  // message of the calling thread's last failed mcgp_create

// Every entry point runs on the context's device and leaves the caller's current device as it found it.
struct DeviceGuard {
    int prev = -1;
    cudaError_t status;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        status = prev == device ? cudaSuccess : cudaSetDevice(device);
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static int fail(mcgp_handle h, int code, const std::string& msg) {
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}
#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(h, MCGP_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));             \
    } while (0)

static int scratch_get(mcgp_handle h, int slot, size_t bytes, void** out) {
    if (bytes == 0) bytes = 16;
    if (h->scratch_sz[slot] < bytes) {
        if (h->scratch[slot]) cudaFree(h->scratch[slot]);
        h->scratch[slot] = nullptr;
        h->scratch_sz[slot] = 0;
        cudaError_t e = cudaMalloc(&h->scratch[slot], bytes);
        if (e != cudaSuccess) return fail(h, MCGP_ENOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
        h->scratch_sz[slot] = bytes;
    }
    *out = h->scratch[slot];
    return MCGP_OK;
}

// ---- host-side parameter derivation ---------------------------------------------------------
static uint32_t prob_threshold(double p) {  // event iff 32-bit word < threshold
    if (!(p > 0.0)) return 0u;
    if (p >= 1.0) return 0xffffffffu;
    return (uint32_t)floor(p * 4294967296.0);
}

// The per-lap test `u < rate` of src/simulation.py:194 makes a car's retirement lap geometric: lap = 2 + floor(ln u / ln(1 - rate)).
static float dnf_scale(double rate) {
    const double thr = (double)prob_threshold(rate) / 4294967296.0;  // same 2^-32 grid as every other probability
    if (!(thr > 0.0)) return MCGP_DNF_NEVER;
    if (thr >= 1.0 || rate >= 1.0) return -0.0f;
    const float s = (float)(1.0 / log1p(-thr));
    return s < -1e29f ? -1e29f : s;
}

static double clamp01(double p) { return !(p > 0.0) ? 0.0 : p > 1.0 ? 1.0 : p; }

// laps the overtake pace table (rows = laps + 5, `stride` entries of 16 B per row, + one padding row) can hold in the
// 160 KB of shared memory it is staged into
static int max_laps_for(int n_drivers) {
    const int stride = n_drivers <= 20 ? 20 : MCGP_LANES;
    return (int)((160u * 1024u / sizeof(PaceEntry) - MCGP_LANES) / stride) - 5;   // 505 (<= 20 drivers), 314 (more)
}

static int validate(mcgp_handle h, const mcgp_race_params* r) {
    if (r->n_drivers < 1 || r->n_drivers > MCGP_MAX_DRIVERS) return fail(h, MCGP_EINVAL, "n_drivers must be in 1..32");
    if (r->total_laps < 1) return fail(h, MCGP_EINVAL, "total_laps must be positive");
    if (r->total_laps > max_laps_for(r->n_drivers))
        return fail(h, MCGP_EINVAL, "total_laps too large: the overtake pace table (laps + 5 rows) must fit 160 KB of shared memory "
                                    "(505 laps for <= 20 drivers, 314 for more)");
    if (r->track_condition < 0 || r->track_condition > 2) return fail(h, MCGP_EINVAL, "bad track_condition");
    if (!(r->pop_no_medium == MCGP_SOFT || r->pop_no_medium == MCGP_HARD)) return fail(h, MCGP_EINVAL, "pop_no_medium must be SOFT or HARD");
    if (!(r->pop_no_soft == MCGP_MEDIUM || r->pop_no_soft == MCGP_HARD)) return fail(h, MCGP_EINVAL, "pop_no_soft must be MEDIUM or HARD");
    for (int d = 0; d < r->n_drivers; d++)
        for (int p = 0; p < r->n_drivers; p++) {
            const double v = r->grid_probs[d][p];
            if (v != v) return fail(h, MCGP_EINVAL, "probabilities contain NaN");              // np.random.choice's message
            if (v < 0) return fail(h, MCGP_EINVAL, "probabilities are not non-negative");      // idem
            if (r->grid_kind[d][p] > 2) return fail(h, MCGP_EINVAL, "bad grid_kind");
        }
    return MCGP_OK;
}

static double pit_window(const mcgp_race_params* r, int c, int d) {  // src/simulation.py:455-462
    double optimal = r->compound_optimal_laps[c];
    const double driver_deg = r->tire_deg_pit[d];
    if (driver_deg > 0.05) optimal = trunc(optimal * 0.85);
    else if (driver_deg < 0.02) optimal = trunc(optimal * 1.1);
    return optimal;
}

static void derive_native(const mcgp_race_params* r, NativeRace* o) {
    memset(o, 0, sizeof(*o));
    const int n = r->n_drivers;
    o->n = n; o->total_laps = r->total_laps; o->track = r->track_condition;
    o->pop_no_medium = r->pop_no_medium; o->pop_no_soft = r->pop_no_soft; o->stream = r->stream;
    o->pit_loss = (float)r->pit_loss; o->drs_delta = (float)r->drs_delta;
    o->drs32 = (float)r->drs_delta * 32768.0f;
    o->dirty_thr = (float)r->dirty_air_threshold; o->dirty_pen = (float)r->dirty_air_penalty;
    {   // red flag, else SC, else VSC (:168-176): one draw against the cumulative probabilities
        const double pr = clamp01(r->red_flag_probability), ps = clamp01(r->sc_probability), pv = clamp01(r->vsc_probability);
        o->red_thr = prob_threshold(pr);
        o->sc_thr = prob_threshold(pr + (1.0 - pr) * ps);
        o->vsc_thr = prob_threshold(pr + (1.0 - pr) * (ps + (1.0 - ps) * pv));
        if (o->sc_thr < o->red_thr) o->sc_thr = o->red_thr;
        if (o->vsc_thr < o->sc_thr) o->vsc_thr = o->sc_thr;
    }
    for (int d = 0; d < MCGP_LANES; d++) {
        const bool car = d < n;
        o->sigma[d] = car ? (float)r->driver_variance[d] : 0.0f;
        o->dnf_scale[d] = car ? dnf_scale(r->dnf_rate[d]) : MCGP_DNF_NEVER;
        o->lap1_thr[d] = car ? prob_threshold(r->team_dnf_rate[d] * 4.0) : 0u;  // LAP_1_DNF_MULTIPLIER :282
        const double deg = car ? r->tire_deg[d] : 0.05;
        const double driver_factor = deg > 0 ? deg / 0.05 : 1.0;  // :321
        for (int c = 0; c < MCGP_NC; c++) {
            o->eff_deg[c][d] = car ? (float)(r->compound_deg_rate[c] * driver_factor) : 0.0f;
            o->opt[c][d] = car ? (float)pit_window(r, c, d) : 1e30f;
            o->pc[c][d] = car ? (float)r->base_pace[d] + (float)r->compound_pace_delta[c] : 0.0f;
        }
    }
    for (int p = 0; p < MCGP_LANES; p++)
        for (int d = 0; d < MCGP_LANES; d++)
            o->grid[p][d] = (p < n && d < n && r->grid_kind[d][p] != MCGP_ITEM_INT0) ? (float)r->grid_probs[d][p] : 0.0f;
    // deterministic grid (one-hot rows forming a permutation, src/predictor.py:189-205): skip the sampler
    bool fixed = true;
    int col_owner[MCGP_LANES];
    for (int p = 0; p < n && fixed; p++) {
        int cnt = 0;
        for (int d = 0; d < n; d++) if (o->grid[p][d] > 0.0f) { cnt++; col_owner[p] = d; }
        if (cnt != 1) fixed = false;
    }
    if (fixed) {
        int seen[MCGP_LANES] = {0};
        for (int p = 0; p < n; p++) {
            if (seen[col_owner[p]]++) { fixed = false; break; }
        }
        if (fixed) for (int p = 0; p < n; p++) o->fixed_slot[col_owner[p]] = (uint8_t)p;
    }
    o->grid_fixed = fixed ? 1 : 0;
}

// ---- overtake pace table (device_params.h: PaceEntry) ----------------------------------------
// FP64, op for op as src/simulation.py:514-521 (this file is compiled with -ffp-contract=off):
//   pace = base_pace + tire_age * tire_deg;  pace_delta = pace_ahead - pace_behind;  if drs: pace_delta += drs_delta;
//   eligible = pace_delta > overtake_delta
static inline bool attack_allowed(double pace_ahead, double pace_behind, bool drs, double drs_delta, double overtake_delta) {
    volatile double pace_delta = pace_ahead - pace_behind;
    if (drs) pace_delta = pace_delta + drs_delta;
    return pace_delta > overtake_delta;
}

static inline int pace_rows(int total_laps) { return total_laps + 5; }  // tyre age <= 4 (used set at the start) + laps

static void build_pace_table(const mcgp_race_params* r, int rows_total, int stride, PaceEntry* out) {
    const int n = r->n_drivers, rows = pace_rows(r->total_laps);
    const float inf = __builtin_inff();
    for (size_t i = 0; i < (size_t)rows_total * stride; i++) out[i] = PaceEntry{__builtin_nanf(""), inf, inf, 0u};
    std::vector<double> P((size_t)rows * n), uniq;
    for (int a = 0; a < rows; a++)
        for (int d = 0; d < n; d++) {
            volatile double wear = (double)a * r->tire_deg[d];
            P[(size_t)a * n + d] = r->base_pace[d] + wear;
        }
    for (double v : P) if (v == v) uniq.push_back(v);
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    const int m = (int)uniq.size();
    std::vector<float> f(m);  // strictly increasing float image of the sorted paces
    for (int i = 0; i < m; i++) {
        f[i] = (float)(uniq[i] * 32768.0);
        if (i > 0 && !(f[i] > f[i - 1])) f[i] = nextafterf(f[i - 1], inf);
    }
    for (int a = 0; a < rows; a++)
        for (int d = 0; d < n; d++) {
            const double pb = P[(size_t)a * n + d];
            if (pb != pb) continue;
            PaceEntry e{f[std::lower_bound(uniq.begin(), uniq.end(), pb) - uniq.begin()], inf, inf, 0u};
            for (int k = 0; k < 2; k++) {  // the slowest ahead car this one may attack: the decision is monotone in its pace
                int lo = 0, hi = m;
                while (lo < hi) {
                    const int mid = (lo + hi) / 2;
                    if (attack_allowed(uniq[mid], pb, k == 1, r->drs_delta, r->overtake_delta)) hi = mid; else lo = mid + 1;
                }
                if (lo < m) (k ? e.thr1 : e.thr0) = f[lo];
            }
            out[(size_t)a * stride + d] = e;
        }
}

static void derive_replay(const mcgp_race_params* r, ReplayRace* o) {
    memset(o, 0, sizeof(*o));
    const int n = r->n_drivers;
    o->n = n; o->total_laps = r->total_laps; o->track = r->track_condition;
    o->pop_no_medium = r->pop_no_medium; o->pop_no_soft = r->pop_no_soft;
    o->pit_loss = r->pit_loss; o->ovt_delta = r->overtake_delta; o->sc_p = r->sc_probability;
    o->vsc_p = r->vsc_probability; o->red_p = r->red_flag_probability; o->drs_delta = r->drs_delta;
    o->dirty_thr = r->dirty_air_threshold; o->dirty_pen = r->dirty_air_penalty;
    for (int c = 0; c < MCGP_NC; c++) { o->cdelta[c] = r->compound_pace_delta[c]; o->cdeg[c] = r->compound_deg_rate[c]; }
    for (int d = 0; d < n; d++) {
        o->pace[d] = r->base_pace[d]; o->deg[d] = r->tire_deg[d]; o->sigma[d] = r->driver_variance[d];
        o->dnf_rate[d] = r->dnf_rate[d];
        o->lap1_rate[d] = r->team_dnf_rate[d] * 4.0;  // base_dnf_rate * LAP_1_DNF_MULTIPLIER :287
        for (int c = 0; c < MCGP_NC; c++) o->opt[c][d] = pit_window(r, c, d);
        for (int p = 0; p < n; p++) { o->grid[d][p] = r->grid_probs[d][p]; o->kind[d][p] = r->grid_kind[d][p]; }
    }
}

// ---- ABI ------------------------------------------------------------------------------------
// ---- resident batch management ------------------------------------------------------------------
static void drop_resident(mcgp_handle h) {  // after a failed upload nothing may look valid
    h->n_races = 0; h->n_drivers = 0; h->replay_ready = false; h->resident.clear();
}

static void free_blocks(mcgp_handle h) {
    if (h->native_dev) { cudaFree(h->native_dev); h->native_dev = nullptr; }
    if (h->replay_dev) { cudaFree(h->replay_dev); h->replay_dev = nullptr; }
    if (h->work_counter) { cudaFree(h->work_counter); h->work_counter = nullptr; }
    if (h->pace_dev) { cudaFree(h->pace_dev); h->pace_dev = nullptr; }
    h->cap_races = 0; h->cap_pace = 0;
    drop_resident(h);
}

static int pinned_get(mcgp_handle h, size_t bytes, void** out) {
    if (h->pinned_sz < bytes) {
        if (h->pinned) cudaFreeHost(h->pinned);
        h->pinned = nullptr; h->pinned_sz = 0;
        const size_t want = std::max(bytes, (size_t)1 << 16);
        cudaError_t e = cudaMallocHost(&h->pinned, want);
        if (e != cudaSuccess) return fail(h, MCGP_ENOMEM, std::string("cudaMallocHost: ") + cudaGetErrorString(e));
        h->pinned_sz = want;
    }
    *out = h->pinned;
    return MCGP_OK;
}

// Wait (on the host) for the last launch this handle made before its blocks or counters are overwritten.
static void wait_last_launch(mcgp_handle h) {
    if (h->launched) { cudaEventSynchronize(h->last_launch); h->launched = false; }
}

// Derives and uploads the NATIVE blocks (parameter blocks + overtake pace tables) of a batch on `st`, staged through the
// pinned block at `stage` (NULL: pageable temporaries + synchronous copies).  A batch identical to the resident one is
// not derived or copied again.  The replay blocks are derived lazily (ensure_replay): the product path never replays.
static int upload_native(mcgp_handle h, const mcgp_race_params* races, int n_races, cudaStream_t st, bool async) {
    if (!races || n_races < 1) return fail(h, MCGP_EINVAL, "races is NULL or n_races < 1");
    if (h->n_races == n_races && h->resident.size() == (size_t)n_races &&
        memcmp(h->resident.data(), races, sizeof(mcgp_race_params) * (size_t)n_races) == 0) {
        h->upload_bytes = 0;  // resident already: nothing crosses the bus
        return MCGP_OK;
    }
    for (int r = 0; r < n_races; r++) {
        int rc = validate(h, &races[r]);
        if (rc) return rc;
        if (races[r].n_drivers != races[0].n_drivers) return fail(h, MCGP_EINVAL, "all races of a batch must have the same n_drivers");
    }
    int rows = 0;
    bool uniform = true;
    for (int r = 0; r < n_races; r++) {
        rows = std::max(rows, pace_rows(races[r].total_laps));
        uniform = uniform && races[r].total_laps == races[0].total_laps;
    }
    const int stride = races[0].n_drivers <= 20 ? 20 : MCGP_LANES;
    const size_t per_race = mcgp::pace_pairs_per_race(rows, stride);
    const size_t b_nat = sizeof(NativeRace) * (size_t)n_races, b_pace = sizeof(PacePair) * per_race * n_races;
    wait_last_launch(h);  // in-flight kernels still read the old blocks
    drop_resident(h);
    // host staging: pinned (async path) or a plain temporary
    std::vector<char> tmp;
    char* stage = nullptr;
    if (async) {
        void* pin = nullptr;
        // (room behind the staging area for the count tables mcgp_run_native sends through the same block)
        int rc = pinned_get(h, ((b_nat + b_pace + 4095) & ~(size_t)4095) + (size_t)n_races * MCGP_LANES * MCGP_LANES * 8, &pin);
        if (rc) return rc;
        stage = (char*)pin;
    } else {
        try { tmp.resize(b_nat + b_pace); } catch (...) { return fail(h, MCGP_ENOMEM, "out of host memory"); }
        stage = tmp.data();
    }
    NativeRace* nat = reinterpret_cast<NativeRace*>(stage);
    PacePair* pace = reinterpret_cast<PacePair*>(stage + b_nat);
    std::vector<PaceEntry> entries;
    try { entries.resize((size_t)rows * stride); } catch (...) { return fail(h, MCGP_ENOMEM, "out of host memory"); }
    memset(pace, 0, b_pace);
    for (int r = 0; r < n_races; r++) {
        derive_native(&races[r], &nat[r]);
        build_pace_table(&races[r], rows, stride, entries.data());
        PacePair* t0 = pace + per_race * r, *t1 = t0 + entries.size();  // without / with DRS
        for (size_t i = 0; i < entries.size(); i++) {
            t0[i] = PacePair{entries[i].op32, entries[i].thr0};
            t1[i] = PacePair{entries[i].op32, entries[i].thr1};
        }
    }
    // device blocks are grow-only: a product-sized call (10 000 sims = 0.13 ms of kernel) must not pay for cudaMalloc / cudaFree
    cudaError_t e = cudaSuccess;
    if (n_races > h->cap_races) {
        if (h->native_dev) { cudaFree(h->native_dev); h->native_dev = nullptr; }
        if (h->replay_dev) { cudaFree(h->replay_dev); h->replay_dev = nullptr; }
        if (h->work_counter) { cudaFree(h->work_counter); h->work_counter = nullptr; }
        h->cap_races = 0;
        e = cudaMalloc(&h->native_dev, sizeof(NativeRace) * n_races);
        if (e == cudaSuccess) e = cudaMalloc(&h->work_counter, sizeof(unsigned long long) * n_races);
        if (e == cudaSuccess) e = cudaMalloc(&h->replay_dev, sizeof(ReplayRace) * n_races);
        if (e == cudaSuccess) h->cap_races = n_races;
    }
    if (e == cudaSuccess && per_race * n_races > h->cap_pace) {
        if (h->pace_dev) { cudaFree(h->pace_dev); h->pace_dev = nullptr; }
        h->cap_pace = 0;
        e = cudaMalloc(&h->pace_dev, b_pace);
        if (e == cudaSuccess) h->cap_pace = per_race * n_races;
    }
    if (e == cudaSuccess) e = async ? cudaMemcpyAsync(h->native_dev, nat, b_nat, cudaMemcpyHostToDevice, st)
                                    : cudaMemcpy(h->native_dev, nat, b_nat, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = async ? cudaMemcpyAsync(h->pace_dev, pace, b_pace, cudaMemcpyHostToDevice, st)
                                    : cudaMemcpy(h->pace_dev, pace, b_pace, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {  // leave nothing half-initialised behind: a later launch must fail cleanly, not run on garbage
        free_blocks(h);
        return fail(h, MCGP_ECUDA, std::string("upload: ") + cudaGetErrorString(e));
    }
    try { h->resident.assign(races, races + n_races); } catch (...) { drop_resident(h); return fail(h, MCGP_ENOMEM, "out of host memory"); }
    h->n_races = n_races; h->n_drivers = races[0].n_drivers;
    h->pace_rows = rows; h->pace_stride = stride; h->uniform_laps = uniform;
    h->upload_bytes = b_nat + b_pace;
    return MCGP_OK;
}

// The FP64 blocks of the resident batch, derived and uploaded the first time a replay needs them.
static int ensure_replay(mcgp_handle h) {
    if (h->replay_ready) return MCGP_OK;
    if (h->n_races < 1) return fail(h, MCGP_EINVAL, "mcgp_upload_races has not been called");
    std::vector<ReplayRace> rep;
    try { rep.resize(h->n_races); } catch (...) { return fail(h, MCGP_ENOMEM, "out of host memory"); }
    for (int r = 0; r < h->n_races; r++) derive_replay(&h->resident[r], &rep[r]);
    wait_last_launch(h);
    cudaError_t e = cudaMemcpy(h->replay_dev, rep.data(), sizeof(ReplayRace) * h->n_races, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return fail(h, MCGP_ECUDA, std::string("upload (replay blocks): ") + cudaGetErrorString(e));
    h->replay_ready = true;
    h->upload_bytes += sizeof(ReplayRace) * (uint64_t)h->n_races;
    return MCGP_OK;
}

// Launches on one handle are ordered against each other whatever stream they are given: they share the claim counters.
static cudaError_t order_before(mcgp_handle h, cudaStream_t st) {
    return h->launched ? cudaStreamWaitEvent(st, h->last_launch, 0) : cudaSuccess;
}
static cudaError_t order_after(mcgp_handle h, cudaStream_t st) {
    cudaError_t e = cudaEventRecord(h->last_launch, st);
    if (e == cudaSuccess) h->launched = true;
    return e;
}

// ---- ABI ------------------------------------------------------------------------------------
extern "C" {

int mcgp_abi_version(void) { return MCGP_ABI_VERSION; }
int mcgp_native_philox_rounds(void) { return mcgp::native_philox_rounds(); }

int mcgp_create(mcgp_handle* out, int device) {
    mcgp_handle h = nullptr;
    if (!out) return fail(nullptr, MCGP_EINVAL, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, MCGP_ENODEVICE, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                                                 " (libmcgp has no CPU fallback)");
    if (device < 0 || device >= count) return fail(nullptr, MCGP_EINVAL, "device index out of range");
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(nullptr, MCGP_ENODEVICE, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, MCGP_ENODEVICE, "device is sm_" + std::to_string(prop.major * 10 + prop.minor) +
                                                 "; libmcgp is built for sm_100a (B200) only");
    h = new (std::nothrow) mcgp_context();
    if (!h) return fail(nullptr, MCGP_ENOMEM, "out of host memory");
    h->device = device; h->sm_count = prop.multiProcessorCount; h->cc_major = prop.major; h->cc_minor = prop.minor;
    DeviceGuard g(device);
    e = g.status;
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->last_launch, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        if (h->last_launch) cudaEventDestroy(h->last_launch);
        delete h;
        return fail(nullptr, MCGP_ENODEVICE, std::string("context setup: ") + cudaGetErrorString(e));
    }
    *out = h;
    return MCGP_OK;
}

int mcgp_destroy(mcgp_handle h) {
    if (!h) return MCGP_OK;
    DeviceGuard g(h->device);
    wait_last_launch(h);
    free_blocks(h);
    for (int i = 0; i < 8; i++) if (h->scratch[i]) cudaFree(h->scratch[i]);
    if (h->pinned) cudaFreeHost(h->pinned);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->last_launch) cudaEventDestroy(h->last_launch);
    delete h;
    return MCGP_OK;
}

const char* mcgp_last_error(mcgp_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int mcgp_device_info(mcgp_handle h, int* sm_count, int* sm_clock_khz, int* cc_major, int* cc_minor) {
    if (!h) return MCGP_EINVAL;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, h->device);
    if (sm_count) *sm_count = h->sm_count;
    if (sm_clock_khz) *sm_clock_khz = khz;
    if (cc_major) *cc_major = h->cc_major;
    if (cc_minor) *cc_minor = h->cc_minor;
    return MCGP_OK;
}

int mcgp_last_launch_count(mcgp_handle h) { return h ? h->launches : 0; }
uint64_t mcgp_last_upload_bytes(mcgp_handle h) { return h ? h->upload_bytes : 0; }

int mcgp_pace_table(const mcgp_race_params* race, int32_t* rows, int32_t* stride, float* out) {
    if (!race || race->n_drivers < 1 || race->n_drivers > MCGP_MAX_DRIVERS || race->total_laps < 1) return MCGP_EINVAL;
    if (race->total_laps > max_laps_for(race->n_drivers)) return MCGP_EINVAL;  // same limit as mcgp_upload_races
    const int r = pace_rows(race->total_laps), st = race->n_drivers <= 20 ? 20 : MCGP_LANES;
    if (rows) *rows = r;
    if (stride) *stride = st;
    if (out) {
        static_assert(sizeof(PaceEntry) == 4 * sizeof(float), "PaceEntry is four 32-bit words");
        build_pace_table(race, r, st, reinterpret_cast<PaceEntry*>(out));
    }
    return MCGP_OK;
}

int mcgp_upload_races(mcgp_handle h, const mcgp_race_params* races, int n_races) {
    if (!h) return MCGP_EINVAL;
    DeviceGuard g(h->device);
    if (g.status != cudaSuccess) return fail(h, MCGP_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(g.status));
    return upload_native(h, races, n_races, nullptr, false);
}

static int launch_native_common(mcgp_handle h, uint64_t n_sims, uint64_t sim_begin, uint64_t seed, uint32_t flags,
                                uint64_t* hist_dev, uint8_t* finish_dev, float* times_dev, mcgp_trace_record* trace_dev,
                                uint64_t trace_first, uint64_t trace_count, void* cuda_stream, uint64_t* laphist_dev = nullptr) {
    if (!h) return MCGP_EINVAL;
    if (!h->native_dev || h->n_races < 1) return fail(h, MCGP_EINVAL, "mcgp_upload_races has not been called");
    if (!hist_dev) return fail(h, MCGP_EINVAL, "hist_dev is NULL");
    h->launches = 0;
    if (n_sims == 0) return MCGP_OK;
    // the per-block count tables are 32-bit (shared memory), flushed once per block: a block cannot see 2^32 sims
    if (n_sims > 0xffffffffull) return fail(h, MCGP_EINVAL, "at most 2^32 - 1 sims per launch: split the range over several launches");
    if (trace_dev && (trace_first > n_sims || trace_count > n_sims - trace_first))
        return fail(h, MCGP_EINVAL, "trace window exceeds the launched sim range");
    if (trace_dev && trace_count && !h->uniform_laps)  // the trace is laid out [race][sim][lap][driver] with ONE lap count
        return fail(h, MCGP_EINVAL, "traced batches need equal total_laps");
    DeviceGuard g(h->device);
    if (g.status != cudaSuccess) return fail(h, MCGP_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(g.status));
    static_assert(sizeof(mcgp_trace_record) == sizeof(TraceRecord) && sizeof(TraceRecord) == 8, "trace record layout");
    const cudaStream_t st = (cudaStream_t)cuda_stream;
    CU(order_before(h, st));
    CU(mcgp::launch_native(h->native_dev, h->pace_dev, h->pace_rows, h->pace_stride, h->n_races, h->n_drivers, n_sims, sim_begin, seed, (flags & MCGP_F_EXACT_NORMAL) != 0,
                           (unsigned long long*)hist_dev, finish_dev, times_dev, (TraceRecord*)(trace_count ? trace_dev : nullptr),
                           trace_first, trace_count, (unsigned long long*)laphist_dev, h->work_counter, h->sm_count, st));
    CU(order_after(h, st));
    h->launches = 2;  // the claim-counter reset + the race kernel
    return MCGP_OK;
}

int mcgp_launch_native(mcgp_handle h, uint64_t n_sims, uint64_t sim_begin, uint64_t seed, uint32_t flags,
                       uint64_t* hist_dev, uint8_t* finish_dev, float* times_dev, void* cuda_stream) {
    return launch_native_common(h, n_sims, sim_begin, seed, flags, hist_dev, finish_dev, times_dev, nullptr, 0, 0, cuda_stream);
}

int mcgp_launch_native_traced(mcgp_handle h, uint64_t n_sims, uint64_t sim_begin, uint64_t seed, uint32_t flags,
                              uint64_t* hist_dev, mcgp_trace_record* trace_dev, uint64_t trace_first, uint64_t trace_count,
                              void* cuda_stream) {
    if (h && !trace_dev) return fail(h, MCGP_EINVAL, "trace_dev is NULL");
    return launch_native_common(h, n_sims, sim_begin, seed, flags, hist_dev, nullptr, nullptr, trace_dev, trace_first, trace_count, cuda_stream);
}

int mcgp_lap_histogram_laps(mcgp_handle h) { return h && h->native_dev && h->n_races > 0 ? h->pace_rows - 5 : 0; }

int mcgp_launch_native_laphist(mcgp_handle h, uint64_t n_sims, uint64_t sim_begin, uint64_t seed, uint32_t flags,
                               uint64_t* hist_dev, uint64_t* laphist_dev, void* cuda_stream) {
    if (h && !laphist_dev) return fail(h, MCGP_EINVAL, "laphist_dev is NULL");
    if (h && h->native_dev && h->n_races > 0) {
        const size_t cells = (size_t)(h->pace_rows - 5) * h->n_drivers * h->n_drivers;
        const size_t smem = mcgp::pace_pairs_per_race(h->pace_rows, h->pace_stride) * sizeof(PacePair) + cells * 4;
        if (smem > 180u * 1024u)
            return fail(h, MCGP_EINVAL, "laps x drivers^2 too large: the lap histogram (4 B per cell) and the pace table must fit 180 KB of shared memory");
    }
    return launch_native_common(h, n_sims, sim_begin, seed, flags, hist_dev, nullptr, nullptr, nullptr, 0, 0, cuda_stream, laphist_dev);
}

int mcgp_run_native_laphist(mcgp_handle h, const mcgp_race_params* races, int n_races, uint64_t n_sims, uint64_t sim_begin,
                            uint64_t seed, uint32_t flags, uint64_t* hist_host, uint64_t* laphist_host) {
    if (!h) return MCGP_EINVAL;
    if (!hist_host || !laphist_host) return fail(h, MCGP_EINVAL, "NULL output pointer");
    int rc = mcgp_upload_races(h, races, n_races);
    if (rc) return rc;
    DeviceGuard g(h->device);
    const size_t n = (size_t)h->n_drivers;
    const size_t hist_bytes = (size_t)n_races * n * n * sizeof(uint64_t);
    const size_t lh_bytes = (size_t)n_races * (size_t)(h->pace_rows - 5) * n * n * sizeof(uint64_t);
    void *hist_dev = nullptr, *lh_dev = nullptr;
    if ((rc = scratch_get(h, 0, hist_bytes, &hist_dev))) return rc;
    if ((rc = scratch_get(h, 6, lh_bytes, &lh_dev))) return rc;
    CU(cudaMemcpy(hist_dev, hist_host, hist_bytes, cudaMemcpyHostToDevice));      // counts accumulate (+=)
    CU(cudaMemcpy(lh_dev, laphist_host, lh_bytes, cudaMemcpyHostToDevice));
    rc = mcgp_launch_native_laphist(h, n_sims, sim_begin, seed, flags, (uint64_t*)hist_dev, (uint64_t*)lh_dev, nullptr);
    if (rc) return rc;
    CU(cudaMemcpy(hist_host, hist_dev, hist_bytes, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(laphist_host, lh_dev, lh_bytes, cudaMemcpyDeviceToHost));
    CU(cudaDeviceSynchronize());
    return MCGP_OK;
}

int mcgp_run_native_traced(mcgp_handle h, const mcgp_race_params* races, int n_races, uint64_t n_sims, uint64_t sim_begin,
                           uint64_t seed, uint32_t flags, uint64_t* hist_host, mcgp_trace_record* trace_host,
                           uint64_t trace_first, uint64_t trace_count) {
    if (!h) return MCGP_EINVAL;
    if (!hist_host || !trace_host) return fail(h, MCGP_EINVAL, "NULL output pointer");
    if (!races || n_races < 1) return fail(h, MCGP_EINVAL, "races is NULL or n_races < 1");
    for (int r = 1; r < n_races; r++)
        if (races[r].total_laps != races[0].total_laps) return fail(h, MCGP_EINVAL, "traced batches need equal total_laps");
    int rc = mcgp_upload_races(h, races, n_races);
    if (rc) return rc;
    DeviceGuard g(h->device);
    const size_t n = (size_t)h->n_drivers;
    const size_t hist_bytes = (size_t)n_races * n * n * sizeof(uint64_t);
    const size_t tr_bytes = (size_t)n_races * trace_count * (size_t)races[0].total_laps * n * sizeof(mcgp_trace_record);
    void *hist_dev = nullptr, *tr_dev = nullptr;
    if ((rc = scratch_get(h, 0, hist_bytes, &hist_dev))) return rc;
    if ((rc = scratch_get(h, 6, tr_bytes, &tr_dev))) return rc;
    CU(cudaMemcpy(hist_dev, hist_host, hist_bytes, cudaMemcpyHostToDevice));
    rc = mcgp_launch_native_traced(h, n_sims, sim_begin, seed, flags, (uint64_t*)hist_dev, (mcgp_trace_record*)tr_dev, trace_first,
                                   trace_count, nullptr);
    if (rc) return rc;
    CU(cudaMemcpy(hist_host, hist_dev, hist_bytes, cudaMemcpyDeviceToHost));
    if (tr_bytes) CU(cudaMemcpy(trace_host, tr_dev, tr_bytes, cudaMemcpyDeviceToHost));
    CU(cudaDeviceSynchronize());
    return MCGP_OK;
}

// The product call (src/predictor.py:283-291 makes it with 10 000 sims: 0.13 ms of kernel): parameters derived only
// when they differ from the resident batch, everything staged through ONE pinned block, async copies and the two
// launches on the handle's own stream, ONE synchronisation.
int mcgp_run_native(mcgp_handle h, const mcgp_race_params* races, int n_races, uint64_t n_sims, uint64_t sim_begin,
                    uint64_t seed, uint32_t flags, uint64_t* hist_host, uint8_t* finish_host, float* times_host) {
    if (!h) return MCGP_EINVAL;
    if (!hist_host) return fail(h, MCGP_EINVAL, "hist_host is NULL");
    DeviceGuard g(h->device);
    if (g.status != cudaSuccess) return fail(h, MCGP_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(g.status));
    const cudaStream_t st = h->own_stream;
    int rc = upload_native(h, races, n_races, st, true);   // (stages through the front of the pinned block)
    if (rc) return rc;
    const size_t n = (size_t)h->n_drivers;
    const size_t hist_bytes = (size_t)n_races * n * n * sizeof(uint64_t);
    const size_t fin_bytes = finish_host ? (size_t)n_races * n_sims * n : 0;
    const size_t tim_bytes = times_host ? (size_t)n_races * n_sims * n * sizeof(float) : 0;
    void *hist_dev = nullptr, *fin_dev = nullptr, *tim_dev = nullptr;
    if ((rc = scratch_get(h, 0, hist_bytes, &hist_dev))) return rc;
    if (finish_host && (rc = scratch_get(h, 1, fin_bytes, &fin_dev))) return rc;
    if (times_host && (rc = scratch_get(h, 6, tim_bytes, &tim_dev))) return rc;
    // the count table travels through the pinned block BEHIND the parameter staging area (which the async upload of
    // this very call may still be reading)
    const size_t stage_off = (sizeof(NativeRace) * (size_t)n_races + sizeof(PacePair) * mcgp::pace_pairs_per_race(h->pace_rows, h->pace_stride) * n_races + 4095) & ~(size_t)4095;
    void* pin = nullptr;
    if (h->pinned_sz < stage_off + hist_bytes) {
        // growing the block would free memory an in-flight copy reads: finish the upload first
        CU(cudaStreamSynchronize(st));
        if ((rc = pinned_get(h, stage_off + hist_bytes, &pin))) return rc;
    }
    pin = h->pinned;
    uint64_t* hist_pin = reinterpret_cast<uint64_t*>((char*)pin + stage_off);
    memcpy(hist_pin, hist_host, hist_bytes);
    CU(cudaMemcpyAsync(hist_dev, hist_pin, hist_bytes, cudaMemcpyHostToDevice, st));  // counts accumulate (+=)
    rc = launch_native_common(h, n_sims, sim_begin, seed, flags, (uint64_t*)hist_dev, (uint8_t*)fin_dev, (float*)tim_dev, nullptr, 0, 0, st);
    if (rc) return rc;
    CU(cudaMemcpyAsync(hist_pin, hist_dev, hist_bytes, cudaMemcpyDeviceToHost, st));
    if (finish_host && fin_bytes) CU(cudaMemcpyAsync(finish_host, fin_dev, fin_bytes, cudaMemcpyDeviceToHost, st));
    if (times_host && tim_bytes) CU(cudaMemcpyAsync(times_host, tim_dev, tim_bytes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    memcpy(hist_host, hist_pin, hist_bytes);
    return MCGP_OK;
}

// ---- scoring and the device-resident season (kernels in season_kernels.cu) -------------------
int mcgp_score_counts(mcgp_handle h, const uint64_t* hist_dev, int n_races, int n_drivers, uint64_t n_sims, const int32_t* winner,
                      const int32_t* podium, uint64_t* tallies, double* brier, int32_t* podium_hits, double* calib,
                      int32_t* calib_bins, void* cuda_stream) {
    if (!h) return MCGP_EINVAL;
    if (!hist_dev || !winner || n_races < 1 || n_drivers < 1 || n_drivers > MCGP_MAX_DRIVERS || n_sims == 0)
        return fail(h, MCGP_EINVAL, "mcgp_score_counts: bad argument");
    DeviceGuard g(h->device);
    if (g.status != cudaSuccess) return fail(h, MCGP_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(g.status));
    const size_t R = (size_t)n_races, n = (size_t)n_drivers;
    // one scratch block: winner | podium | tallies | brier | hits | calib | bins
    const size_t o_win = 0, o_pod = o_win + R * 4, o_tal = (o_pod + R * 12 + 7) & ~(size_t)7, o_bri = o_tal + R * 3 * n * 8,
                 o_cal = o_bri + R * 8, o_hit = o_cal + 30 * 8, o_bin = o_hit + R * 4, total = o_bin + 8;
    void* blk = nullptr;
    int rc = scratch_get(h, 7, total, &blk);
    if (rc) return rc;
    char* b = (char*)blk;
    const cudaStream_t st = (cudaStream_t)cuda_stream;
    CU(cudaMemcpyAsync(b + o_win, winner, R * 4, cudaMemcpyHostToDevice, st));
    if (podium) CU(cudaMemcpyAsync(b + o_pod, podium, R * 12, cudaMemcpyHostToDevice, st));
    CU(mcgp::launch_score_counts((const unsigned long long*)hist_dev, n_races, n_drivers, n_sims, (const int32_t*)(b + o_win),
                                 podium ? (const int32_t*)(b + o_pod) : nullptr, (unsigned long long*)(b + o_tal), (double*)(b + o_bri),
                                 (int32_t*)(b + o_hit), (double*)(b + o_cal), (int32_t*)(b + o_bin), st));
    if (tallies) CU(cudaMemcpyAsync(tallies, b + o_tal, R * 3 * n * 8, cudaMemcpyDeviceToHost, st));
    if (brier) CU(cudaMemcpyAsync(brier, b + o_bri, R * 8, cudaMemcpyDeviceToHost, st));
    if (podium_hits) CU(cudaMemcpyAsync(podium_hits, b + o_hit, R * 4, cudaMemcpyDeviceToHost, st));
    if (calib) CU(cudaMemcpyAsync(calib, b + o_cal, 30 * 8, cudaMemcpyDeviceToHost, st));
    if (calib_bins) CU(cudaMemcpyAsync(calib_bins, b + o_bin, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    h->launches = 1;
    return MCGP_OK;
}

int mcgp_run_season(mcgp_handle h, const mcgp_race_params* races, int n_races, uint64_t n_sims, uint64_t seed, uint32_t flags,
                    double k_factor, const double* quali0, const double* race0, const int32_t* penalties, uint64_t* hist,
                    double* quali_hist, double* race_hist, double* grid_rows, uint8_t* actual_grid, uint8_t* actual_finish,
                    uint64_t* tallies, double* brier, int32_t* podium_hits, double* calib, int32_t* calib_bins) {
    if (!h) return MCGP_EINVAL;
    if (!races || n_races < 1 || !quali0 || !race0 || n_sims == 0) return fail(h, MCGP_EINVAL, "mcgp_run_season: bad argument");
    if (n_sims >= 0xffffffffull) return fail(h, MCGP_EINVAL, "at most 2^32 - 2 sims per race");
    DeviceGuard g(h->device);
    if (g.status != cudaSuccess) return fail(h, MCGP_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(g.status));
    const cudaStream_t st = h->own_stream;
    int rc = upload_native(h, races, n_races, st, true);
    if (rc) return rc;
    drop_resident(h);   // the grid blocks are rewritten on the device below: the resident copy no longer mirrors `races`
    h->n_races = n_races; h->n_drivers = races[0].n_drivers;
    const size_t R = (size_t)n_races, n = (size_t)h->n_drivers, nn = n * n;
    const size_t per_race_pairs = mcgp::pace_pairs_per_race(h->pace_rows, h->pace_stride);
    // device scratch: hist | scratch hist of the "actual" sims | q_hist | r_hist | rows | tallies | brier | calib | penalties |
    //                 winner | podium | hits | bins | a_grid | a_finish
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o = (o + bytes + 15) & ~(size_t)15; return at; };
    const size_t o_hist = take(R * nn * 8), o_hscr = take(nn * 8), o_q = take((R + 1) * n * 8), o_r = take((R + 1) * n * 8),
                 o_rows = take(R * nn * 8), o_tal = take(R * 3 * n * 8), o_bri = take(R * 8), o_cal = take(30 * 8),
                 o_pen = take(R * n * 4), o_win = take(R * 4), o_pod = take(R * 12), o_hit = take(R * 4), o_bin = take(16),
                 o_ag = take(R * n), o_af = take(R * n);
    void* blk = nullptr;
    if ((rc = scratch_get(h, 6, o, &blk))) return rc;
    char* b = (char*)blk;
    CU(cudaMemsetAsync(b + o_hist, 0, o_q - o_hist, st));   // the count tables and the scratch table of the actual sims
    CU(cudaMemcpyAsync(b + o_q, quali0, n * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(b + o_r, race0, n * 8, cudaMemcpyHostToDevice, st));
    if (penalties) CU(cudaMemcpyAsync(b + o_pen, penalties, R * n * 4, cudaMemcpyHostToDevice, st));
    int launches = 0;
    for (int r = 0; r < n_races; r++) {
        NativeRace* race_dev = h->native_dev + r;
        const double* q_r = (const double*)(b + o_q) + (size_t)r * n;
        const double* r_r = (const double*)(b + o_r) + (size_t)r * n;
        uint8_t* ag = (uint8_t*)(b + o_ag) + (size_t)r * n;
        uint8_t* af = (uint8_t*)(b + o_af) + (size_t)r * n;
        // quali ratings -> this race's grid rows, written into its parameter block in place
        CU(mcgp::launch_season_grid(race_dev, q_r, penalties ? (const int32_t*)(b + o_pen) + (size_t)r * n : nullptr, nullptr, nullptr,
                                    nullptr, (int)n, (double*)(b + o_rows) + (size_t)r * nn, st));
        // the race: n_sims sims into this race's count table ...
        CU(mcgp::launch_native(race_dev, h->pace_dev + per_race_pairs * r, h->pace_rows, h->pace_stride, 1, (int)n, n_sims, 0, seed,
                               (flags & MCGP_F_EXACT_NORMAL) != 0, (unsigned long long*)(b + o_hist) + (size_t)r * nn, nullptr, nullptr,
                               nullptr, 0, 0, nullptr, h->work_counter + r, h->sm_count, st));
        // ... and one more (global sim index n_sims) whose grid and finishing order stand in for what really happened
        CU(mcgp::launch_native(race_dev, h->pace_dev + per_race_pairs * r, h->pace_rows, h->pace_stride, 1, (int)n, 1, n_sims, seed,
                               (flags & MCGP_F_EXACT_NORMAL) != 0, (unsigned long long*)(b + o_hscr), af, nullptr, nullptr, 0, 0, nullptr,
                               h->work_counter + r, h->sm_count, st, ag));
        CU(mcgp::launch_season_actual(af, (int32_t*)(b + o_win) + r, (int32_t*)(b + o_pod) + 3 * r, (int)n, st));
        // pairwise Elo: quali ratings from the actual grid, race ratings from the actual finishing order -> next race
        CU(mcgp::launch_season_elo(ag, af, q_r, r_r, (double*)(b + o_q) + (size_t)(r + 1) * n, (double*)(b + o_r) + (size_t)(r + 1) * n,
                                   (int)n, k_factor, st));
        launches += 7;
    }
    CU(mcgp::launch_score_counts((const unsigned long long*)(b + o_hist), n_races, (int)n, n_sims, (const int32_t*)(b + o_win),
                                 (const int32_t*)(b + o_pod), (unsigned long long*)(b + o_tal), (double*)(b + o_bri),
                                 (int32_t*)(b + o_hit), (double*)(b + o_cal), (int32_t*)(b + o_bin), st));
    CU(order_after(h, st));
    if (hist) CU(cudaMemcpyAsync(hist, b + o_hist, R * nn * 8, cudaMemcpyDeviceToHost, st));
    if (quali_hist) CU(cudaMemcpyAsync(quali_hist, b + o_q, (R + 1) * n * 8, cudaMemcpyDeviceToHost, st));
    if (race_hist) CU(cudaMemcpyAsync(race_hist, b + o_r, (R + 1) * n * 8, cudaMemcpyDeviceToHost, st));
    if (grid_rows) CU(cudaMemcpyAsync(grid_rows, b + o_rows, R * nn * 8, cudaMemcpyDeviceToHost, st));
    if (actual_grid) CU(cudaMemcpyAsync(actual_grid, b + o_ag, R * n, cudaMemcpyDeviceToHost, st));
    if (actual_finish) CU(cudaMemcpyAsync(actual_finish, b + o_af, R * n, cudaMemcpyDeviceToHost, st));
    if (tallies) CU(cudaMemcpyAsync(tallies, b + o_tal, R * 3 * n * 8, cudaMemcpyDeviceToHost, st));
    if (brier) CU(cudaMemcpyAsync(brier, b + o_bri, R * 8, cudaMemcpyDeviceToHost, st));
    if (podium_hits) CU(cudaMemcpyAsync(podium_hits, b + o_hit, R * 4, cudaMemcpyDeviceToHost, st));
    if (calib) CU(cudaMemcpyAsync(calib, b + o_cal, 30 * 8, cudaMemcpyDeviceToHost, st));
    if (calib_bins) CU(cudaMemcpyAsync(calib_bins, b + o_bin, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    h->launches = launches + 1;
    // nothing valid stays resident: the parameter blocks now hold device-derived grids
    h->n_races = 0; h->n_drivers = 0;
    return MCGP_OK;
}

int mcgp_replay_serial_grid(mcgp_handle h, int on) {
    if (!h) return MCGP_EINVAL;
    h->replay_serial_grid = on != 0;
    return MCGP_OK;
}

int mcgp_launch_replay(mcgp_handle h, uint64_t n_sims, const double* u_py_dev, const double* z_dev, const double* u_np_dev,
                       const int64_t* off_dev, uint64_t* hist_dev, uint8_t* finish_dev, double* times_dev,
                       int16_t* dnf_lap_dev, uint8_t* grid_dev, int64_t* used_dev, int32_t* status_dev, void* cuda_stream) {
    if (!h) return MCGP_EINVAL;
    if (!h->replay_dev || h->n_races < 1) return fail(h, MCGP_EINVAL, "mcgp_upload_races has not been called");
    if (h->n_races != 1) return fail(h, MCGP_EINVAL, "replay mode takes exactly one race");
    if (!hist_dev || !off_dev || !u_py_dev || !z_dev || !u_np_dev) return fail(h, MCGP_EINVAL, "NULL tape/hist pointer");
    h->launches = 0;
    if (n_sims == 0) return MCGP_OK;
    DeviceGuard g(h->device);
    if (g.status != cudaSuccess) return fail(h, MCGP_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(g.status));
    int rc = ensure_replay(h);
    if (rc) return rc;
    const cudaStream_t st = (cudaStream_t)cuda_stream;
    CU(order_before(h, st));
    CU(mcgp::launch_replay(h->replay_dev, h->n_drivers, n_sims, u_py_dev, z_dev, u_np_dev, (const long long*)off_dev,
                           (unsigned long long*)hist_dev, finish_dev, times_dev, dnf_lap_dev, grid_dev,
                           (long long*)used_dev, status_dev, h->work_counter, h->sm_count, st, h->replay_serial_grid));
    CU(order_after(h, st));
    h->launches = 2;  // the claim-counter reset + the replay kernel
    return MCGP_OK;
}

int mcgp_run_replay(mcgp_handle h, const mcgp_race_params* race, uint64_t n_sims, const double* u_py, const double* z,
                    const double* u_np, const int64_t* off, uint64_t* hist_host, uint8_t* finish_host, double* times_host,
                    int16_t* dnf_lap_host, uint8_t* grid_host, int64_t* used_host) {
    if (!h) return MCGP_EINVAL;
    if (!race || !off || !hist_host) return fail(h, MCGP_EINVAL, "NULL argument");
    int rc = mcgp_upload_races(h, race, 1);
    if (rc) return rc;
    DeviceGuard g(h->device);
    const size_t n = (size_t)h->n_drivers;
    const int64_t* end = off + 3 * n_sims;
    const size_t b_py = (size_t)end[0] * 8, b_z = (size_t)end[1] * 8, b_np = (size_t)end[2] * 8;
    if ((b_py && !u_py) || (b_z && !z) || (b_np && !u_np)) return fail(h, MCGP_EINVAL, "NULL tape");
    const size_t b_off = (size_t)(3 * (n_sims + 1)) * 8, b_hist = n * n * 8;
    void *d_py, *d_z, *d_np, *d_off, *d_hist, *d_out, *d_status;
    if ((rc = scratch_get(h, 2, b_py, &d_py)) || (rc = scratch_get(h, 3, b_z, &d_z)) || (rc = scratch_get(h, 4, b_np, &d_np)) ||
        (rc = scratch_get(h, 5, b_off, &d_off)) || (rc = scratch_get(h, 0, b_hist, &d_hist)) || (rc = scratch_get(h, 7, 16, &d_status)))
        return rc;
    // per-sim outputs packed into one scratch buffer: times | used | dnf_lap | finish | grid
    const size_t o_times = 0, o_used = o_times + n_sims * n * 8, o_dnf = o_used + n_sims * 3 * 8;
    const size_t o_fin = o_dnf + ((n_sims * n * 2 + 7) & ~(size_t)7), o_grid = o_fin + ((n_sims * n + 7) & ~(size_t)7);
    const size_t b_out = o_grid + n_sims * n;
    if ((rc = scratch_get(h, 6, b_out, &d_out))) return rc;
    char* ob = (char*)d_out;
    if (b_py) CU(cudaMemcpy(d_py, u_py, b_py, cudaMemcpyHostToDevice));
    if (b_z) CU(cudaMemcpy(d_z, z, b_z, cudaMemcpyHostToDevice));
    if (b_np) CU(cudaMemcpy(d_np, u_np, b_np, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_off, off, b_off, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_hist, hist_host, b_hist, cudaMemcpyHostToDevice));
    CU(cudaMemset(d_status, 0, 16));
    rc = mcgp_launch_replay(h, n_sims, (const double*)d_py, (const double*)d_z, (const double*)d_np, (const int64_t*)d_off,
                            (uint64_t*)d_hist, finish_host ? (uint8_t*)(ob + o_fin) : nullptr,
                            times_host ? (double*)(ob + o_times) : nullptr, dnf_lap_host ? (int16_t*)(ob + o_dnf) : nullptr,
                            grid_host ? (uint8_t*)(ob + o_grid) : nullptr, used_host ? (int64_t*)(ob + o_used) : nullptr,
                            (int32_t*)d_status, nullptr);
    if (rc) return rc;
    int32_t status = 0;
    CU(cudaMemcpy(&status, d_status, 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(hist_host, d_hist, b_hist, cudaMemcpyDeviceToHost));
    if (finish_host) CU(cudaMemcpy(finish_host, ob + o_fin, n_sims * n, cudaMemcpyDeviceToHost));
    if (times_host) CU(cudaMemcpy(times_host, ob + o_times, n_sims * n * 8, cudaMemcpyDeviceToHost));
    if (dnf_lap_host) CU(cudaMemcpy(dnf_lap_host, ob + o_dnf, n_sims * n * 2, cudaMemcpyDeviceToHost));
    if (grid_host) CU(cudaMemcpy(grid_host, ob + o_grid, n_sims * n, cudaMemcpyDeviceToHost));
    if (used_host) CU(cudaMemcpy(used_host, ob + o_used, n_sims * 3 * 8, cudaMemcpyDeviceToHost));
    CU(cudaDeviceSynchronize());
    if (status) return fail(h, MCGP_ETAPE, "a simulated race ran past the end of its tape");
    return MCGP_OK;
}

}  // extern "C"
