/* mcgp.h -- C ABI of the B200-native Monte Carlo race engine (libmcgp.so, sm_100a).
 *
 * Drop-in boundary for ONE hot path of dan-lee-gh/monte-carlo-gp: the lap-by-lap race simulation
 * `RaceSimulator.run_monte_carlo` -> `simulate_race` (reference src/simulation.py:59-560).  The
 * reference is pure Python and has no FFI of its own; these entry points are what a ctypes binding
 * placed behind `src/simulation.py`'s `RaceSimulator` would call (see INTEGRATION.md for the stub).
 * Each declaration cites the reference interface it replaces.
 *
 * Conventions: plain C, no torch/CUDA types in signatures (streams and device buffers travel as
 * raw void / uint64_t pointers).  Every function returns 0 on success, a negative MCGP_E* code on
 * failure; mcgp_last_error() gives the message.  No exceptions cross the ABI.  The caller owns all
 * buffers.  Host calls on one handle must be serialised by the caller (several handles per device are
 * fine: each owns its parameter blocks and claim counters).  GPU work of one handle is ordered by the
 * library itself: launches made through one handle run one after the other whatever streams they are
 * given (they share the handle's claim counters), and mcgp_upload_races / the host-buffer calls wait
 * for the handle's last launch before they overwrite the resident parameter blocks.  Every entry point
 * restores the caller's current CUDA device before it returns.
 * There is NO CPU fallback: every entry point fails with MCGP_ENODEVICE when no sm_100 GPU is present.
 */
#ifndef MCGP_H
#define MCGP_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCGP_ABI_VERSION 1
#define MCGP_MAX_DRIVERS 32 /* one lane per driver, one warp per simulated race */
#define MCGP_N_COMPOUNDS 5

/* tyre compounds in the order of reference src/config.py:45-51 */
enum { MCGP_SOFT = 0, MCGP_MEDIUM = 1, MCGP_HARD = 2, MCGP_INTERMEDIATE = 3, MCGP_WET = 4 };
/* track_condition of run_monte_carlo / simulate_race (src/simulation.py:68,154) */
enum { MCGP_DRY = 0, MCGP_DAMP = 1, MCGP_WETTRACK = 2 };
/* how CPython's builtin sum() treats one grid_probs item (src/simulation.py:123,133) */
enum { MCGP_ITEM_INT0 = 0, MCGP_ITEM_FLOAT = 1, MCGP_ITEM_NPFLOAT = 2 };

enum {
    MCGP_OK = 0,
    MCGP_EINVAL = -1,    /* bad argument (n_drivers out of range, NULL pointer, NaN/negative probability ...) */
    MCGP_ENODEVICE = -2, /* no CUDA device / not an sm_100 part / driver failure at create */
    MCGP_ECUDA = -3,     /* a CUDA runtime call or the kernel failed */
    MCGP_ETAPE = -4,     /* replay: a simulated race ran past the end of its tape */
    MCGP_ENOMEM = -5
};

/* flags for mcgp_run_native* */
enum {
    MCGP_F_EXACT_NORMAL = 1u << 0 /* use the bit-reproducible (IEEE-only) normal generator instead of the
                                     MUFU one; exists so tests can compare against the scalar CPU mirror */
};

/* One race = RaceConfig (src/simulation.py:37-52) + the per-driver inputs of run_monte_carlo
 * (src/simulation.py:59-69), flattened to dense arrays.  Driver index = position of the driver in
 * grid_probs' key order (src/simulation.py:107).  The reference's `.get(key, default)` fallbacks
 *  are applied by the caller when filling this block; the library applies the arithmetic
 * (x4 lap-1 multiplier, 0.85/1.1 stint scaling, deg/0.05 driver factor) itself. */
typedef struct mcgp_race_params {
    int32_t n_drivers;       /* 1..MCGP_MAX_DRIVERS                                              */
    int32_t total_laps;      /* RaceConfig.total_laps, 1..505 (n_drivers <= 20) / 1..314 (more):   *
                              * the per-race overtake pace table must fit in shared memory;     *
                              * anything larger is rejected with MCGP_EINVAL at upload          */
    int32_t track_condition; /* MCGP_DRY / MCGP_DAMP / MCGP_WETTRACK                              */
    int32_t pop_no_medium;   /* `({S,M,H}-{M}).pop()`: MCGP_SOFT or MCGP_HARD  (src/simulation.py:486)  */
    int32_t pop_no_soft;     /* `({S,M,H}-{S}).pop()`: MCGP_MEDIUM or MCGP_HARD (src/simulation.py:488) */
    uint32_t stream;         /* native RNG: extra stream id so races of one batch draw independently */
    double pit_loss, overtake_delta;                       /* RaceConfig :40-41 */
    double sc_probability, vsc_probability, red_flag_probability; /* :42-44 */
    double drs_delta;                                      /* :47 */
    double dirty_air_threshold, dirty_air_penalty;         /* :51-52 */
    double compound_pace_delta[MCGP_N_COMPOUNDS];   /* tire_compounds[c].get('pace_delta', 0)    :325 */
    double compound_deg_rate[MCGP_N_COMPOUNDS];     /* tire_compounds[c].get('deg_rate', 0.05)   :320 */
    double compound_optimal_laps[MCGP_N_COMPOUNDS]; /* tire_compounds[c].get('optimal_laps', 30) :455 */
    double base_pace[MCGP_MAX_DRIVERS];       /* base_pace.get(d, 90.0)                      :202,:514 */
    double tire_deg[MCGP_MAX_DRIVERS];        /* tire_deg.get(d, 0.05)  lap time + overtakes :203,:514 */
    double tire_deg_pit[MCGP_MAX_DRIVERS];    /* tire_deg.get(d, 0.0)   pit-window scaling   :458      */
    double driver_variance[MCGP_MAX_DRIVERS]; /* driver_variance.get(d, 0.15)                :204      */
    double dnf_rate[MCGP_MAX_DRIVERS];        /* driver_dnf_rates.get(d, team rate), laps>=2 :190-193  */
    double team_dnf_rate[MCGP_MAX_DRIVERS];   /* config.dnf_rates.get(team, 0.002), lap 1 x4 :286-287  */
    double grid_probs[MCGP_MAX_DRIVERS][MCGP_MAX_DRIVERS]; /* [driver][grid position]        :119-122  */
    uint8_t grid_kind[MCGP_MAX_DRIVERS][MCGP_MAX_DRIVERS]; /* MCGP_ITEM_* (INT0 also marks "pos >= len(row)") */
} mcgp_race_params;

typedef struct mcgp_context* mcgp_handle;

/* library / device --------------------------------------------------------------------------- */
int mcgp_abi_version(void);
/* Creates a context on CUDA device `device`.  Fails with MCGP_ENODEVICE if there is none. */
int mcgp_create(mcgp_handle* out, int device);
int mcgp_destroy(mcgp_handle h);
/* Message of the last failure on this handle (or of the last failed mcgp_create when h == NULL). */
const char* mcgp_last_error(mcgp_handle h);
/* multiProcessorCount and current SM clock (kHz) of the context's device, for roofline arithmetic. */
int mcgp_device_info(mcgp_handle h, int* sm_count, int* sm_clock_khz, int* cc_major, int* cc_minor);

/* Philox rounds the native kernels were built with (7: the fastest Crush-resistant Philox4x32 of the Random123 paper;
 * -DMCGP_PHILOX_ROUNDS=10 at build time restores the paper's default). */
int mcgp_native_philox_rounds(void);

/* native mode: counter-based Philox4x32-7 keyed (seed ; sim, lap pair, lane, stream), FP32 ------ *
 * Replaces the loop of run_monte_carlo (src/simulation.py:83-94) for sims
 * [sim_begin, sim_begin + n_sims) of each of the n_races races; results do not depend on how a sim
 * range is split over calls or GPUs.
 *   hist    [n_races][n][n] uint64 counts, hist[(r*n + driver)*n + pos] (pos 0 = P1), accumulated (+=)
 *           -- the reference's results[driver][position] (:93-94) before the division by n (:97-100).
 *   finish  optional, [n_races][n_sims][n] driver index per finishing position (:236-242).
 *   times   optional, [n_races][n_sims][n] float: final time behind the winner, per driver index. */

/* Host-buffer form: parameters are copied in, counts copied out, the call is synchronous.  A batch identical
 * (byte for byte) to the one already resident on the handle is not derived or uploaded again.  At most 2^32 - 1
 * sims per launch (all native entry points). */
int mcgp_run_native(mcgp_handle h, const mcgp_race_params* races, int n_races, uint64_t n_sims,
                    uint64_t sim_begin, uint64_t seed, uint32_t flags, uint64_t* hist_host,
                    uint8_t* finish_host /* nullable */, float* times_host /* nullable */);

/* Device-resident form: upload the race blocks once ... */
int mcgp_upload_races(mcgp_handle h, const mcgp_race_params* races, int n_races);
/* ... then launch asynchronously on `cuda_stream` (a cudaStream_t, NULL = default stream); `hist_dev`
 * (and `finish_dev`, `times_dev` if non-NULL) are device pointers, e.g. torch tensors' data_ptr().
 * times_dev: optional [n_races][n_sims][n] float, final time behind the winner per driver index. */
int mcgp_launch_native(mcgp_handle h, uint64_t n_sims, uint64_t sim_begin, uint64_t seed, uint32_t flags,
                       uint64_t* hist_dev, uint8_t* finish_dev, float* times_dev, void* cuda_stream);
/* Optional per-lap trace (BASELINE config 5; an output the reference does not have): one 8-byte record per
 * (sim, lap, driver) for the sims [trace_first, trace_first + trace_count) of the launched range (indices relative
 * to sim_begin), laid out trace[race][sim - trace_first][lap - 1][driver].  9.1 KB per 20-driver x 57-lap race. */
typedef struct mcgp_trace_record {
    uint8_t position; /* running position after the lap (1 = leader), 0 = retired                               */
    uint8_t compound; /* MCGP_SOFT .. MCGP_WET                                                                    */
    uint8_t tire_age; /* laps on the current set (saturates at 255)                                               */
    uint8_t flags;    /* bit0 retired, bit1 DRS armed for the next lap, bit2 pitted this lap, bits4-5 event this lap
                         (1 red flag, 2 safety car, 3 VSC)                                                         */
    float gap;        /* seconds behind the leader                                                                 */
} mcgp_trace_record;
int mcgp_launch_native_traced(mcgp_handle h, uint64_t n_sims, uint64_t sim_begin, uint64_t seed, uint32_t flags,
                              uint64_t* hist_dev, mcgp_trace_record* trace_dev, uint64_t trace_first,
                              uint64_t trace_count, void* cuda_stream);
int mcgp_run_native_traced(mcgp_handle h, const mcgp_race_params* races, int n_races, uint64_t n_sims,
                           uint64_t sim_begin, uint64_t seed, uint32_t flags, uint64_t* hist_host,
                           mcgp_trace_record* trace_host, uint64_t trace_first, uint64_t trace_count);
/* Per-lap position histogram (the on-chip reduction of the trace; absent upstream): besides hist, the launch
 * accumulates (+=)  laphist[((r * laps + lap-1) * n + driver) * n + pos]  = number of sims in which `driver` RUNS in
 * position pos (0 = leading) after lap `lap`; retired cars are not counted, so a row sums to the number of sims in
 * which the driver is still running.  laps = mcgp_lap_histogram_laps(h) = the longest race of the uploaded batch.
 * laps * n * n * 4 B + the pace table must fit 180 KB of shared memory (57 laps x 20 cars: 111 KB). */
int mcgp_lap_histogram_laps(mcgp_handle h);
int mcgp_launch_native_laphist(mcgp_handle h, uint64_t n_sims, uint64_t sim_begin, uint64_t seed, uint32_t flags,
                               uint64_t* hist_dev, uint64_t* laphist_dev, void* cuda_stream);
int mcgp_run_native_laphist(mcgp_handle h, const mcgp_race_params* races, int n_races, uint64_t n_sims,
                            uint64_t sim_begin, uint64_t seed, uint32_t flags, uint64_t* hist_host,
                            uint64_t* laphist_host);
/* Number of kernel launches the last mcgp_launch_native / mcgp_run_* call on this handle made. */
int mcgp_last_launch_count(mcgp_handle h);
/* Bytes the last mcgp_upload_races on this handle copied host -> device (the derived parameter blocks and the
 * overtake pace tables of all races); the host-buffer calls upload on every call. */
uint64_t mcgp_last_upload_bytes(mcgp_handle h);
/* Host-only (no device, no handle): the overtake pace table mcgp_upload_races derives for one race, so that a
 * binding can inspect or test it.  The reference decides `pace_delta > overtake_delta` (src/simulation.py:514-521)
 * in FP64 on pace = base_pace + tire_age * tire_deg; the table holds, per tyre age a (rows) and driver d (`stride`
 * entries per row), four floats {f(P[d][a]), thr_no_drs, thr_drs, 0}: driver b on tyres of age A_b, chasing driver a
 * on tyres of age A_a, may attack iff  f(P[a][A_a]) >= thr[b][A_b]  -- by construction the same truth value as
 * fl(fl(P_a - P_b) [+ drs_delta]) > overtake_delta.  out == NULL: only the sizes are returned.
 * out must hold rows * stride * 4 floats. */
int mcgp_pace_table(const mcgp_race_params* race, int32_t* rows, int32_t* stride, float* out);

/* scoring on the device (reference src/validation.py:82-158) ------------------------------------- *
 * Win / podium / points tallies, the per-race Brier term of the win probabilities, podium hits and the calibration
 * bins computed from count tables that STAY on the GPU (hist_dev: device pointer, [n_races][n][n] as the native
 * launches leave it); only the small results come back.  winner[r] = driver index of the actual winner of race r or
 * -1 (race skipped, brier[r] = NaN: brier_score's `actual is None`); podium[r][3] = actual podium or -1 (skipped,
 * podium_hits[r] = -1); podium may be NULL.
 *   tallies     [n_races][3][n]  counts of P1 / top-3 / top-10 finishes per driver
 *   brier       [n_races]        mean_d (count[d][P1] / n_sims - [d == winner])^2   (:82-106; the season score is their mean)
 *   podium_hits [n_races]        |three highest podium probabilities  ∩  actual podium|  (:109-130)
 *   calib       [3][10]          per bin of np.linspace(0, 1, bins + 1): pairs, sum of outcomes, sum of probabilities (:133-158)
 *   calib_bins  [1]              bins = min(10, max(2, pairs / 10))
 * All outputs are host pointers and may be NULL; the call synchronises `cuda_stream`. */
int mcgp_score_counts(mcgp_handle h, const uint64_t* hist_dev, int n_races, int n_drivers, uint64_t n_sims,
                      const int32_t* winner, const int32_t* podium, uint64_t* tallies, double* brier,
                      int32_t* podium_hits, double* calib, int32_t* calib_bins, void* cuda_stream);

/* A device-resident season: simulate race r -> update the pairwise Elo ratings with its result -> derive race r+1's
 * grid probabilities from the new qualifying ratings -> next launch, all on one stream with no host round trip
 * (reference: backtest_model's loop src/validation.py:176-198, F1EloSystem src/elo.py:45-141, _predict_quali /
 * _adjust_for_penalties src/predictor.py:321-407).  grid_probs of `races` is ignored: race r's grid block is derived
 * on the device from the qualifying ratings before it.  The "actual" result of race r is its sim number n_sims (the
 * one after the counted range): its sampled grid is the qualifying result, its finishing order the race result.
 *   quali0 / race0 [n]            ratings before the first race;  k_factor: F1EloSystem.k
 *   penalties      [n_races][n]   grid places lost per driver (0 = none), may be NULL
 *   hist           [n_races][n][n]  count tables (overwritten, not accumulated)
 *   quali_hist / race_hist [n_races + 1][n]  ratings before each race and after the last
 *   grid_rows      [n_races][n][n]  the derived grid_probs [driver][position] (FP64, before the float conversion)
 *   actual_grid / actual_finish [n_races][n]  driver index per grid slot / finishing position of the actual sims
 *   tallies, brier, podium_hits, calib, calib_bins: as mcgp_score_counts, scored against the actual results
 * All outputs are host pointers and may be NULL. */
int mcgp_run_season(mcgp_handle h, const mcgp_race_params* races, int n_races, uint64_t n_sims, uint64_t seed,
                    uint32_t flags, double k_factor, const double* quali0, const double* race0,
                    const int32_t* penalties, uint64_t* hist, double* quali_hist, double* race_hist,
                    double* grid_rows, uint8_t* actual_grid, uint8_t* actual_finish, uint64_t* tallies,
                    double* brier, int32_t* podium_hits, double* calib, int32_t* calib_bins);

/* replay mode: FP64, consumes the reference's own draws, bit-exact ------------------------------ *
 * Sim s reads u_py[off[3s]..off[3s+3]) (random.random() values, :168-194,:287,:392,:524),
 * z[off[3s+1]..) (standard normals; np.random.normal(0,s) = 0 + s*z, :302,:330) and
 * u_np[off[3s+2]..) (the one random_sample() per np.random.choice, :137), in the draw order of
 * SURVEY.md §8 "Draw-order specification".  off has 3*(n_sims+1) entries.
 *   hist    [n][n] uint64, accumulated (+=)
 *   finish  optional [n_sims][n] driver index per finishing position
 *   times   optional [n_sims][n] final cumulative_time per driver index (double, bit-exact)
 *   dnf_lap optional [n_sims][n] int16 lap of retirement per driver index (0 = classified finisher)
 *   grid    optional [n_sims][n] driver index per grid slot (_sample_grid, :102-145)
 *   used    optional [n_sims][3] int64 number of u_py / z / u_np draws each sim consumed */
int mcgp_run_replay(mcgp_handle h, const mcgp_race_params* race, uint64_t n_sims, const double* u_py,
                    const double* z, const double* u_np, const int64_t* off, uint64_t* hist_host,
                    uint8_t* finish_host, double* times_host, int16_t* dnf_lap_host, uint8_t* grid_host,
                    int64_t* used_host);
/* _sample_grid in replay mode (src/simulation.py:102-145): the kernel evaluates each grid position's selection with a
 * warp scan and accepts it only where that provably equals the reference's serial evaluation (CPython sum(), division,
 * NumPy cumsum, searchsorted: every intermediate within (n + 4) ulp of the scan, the draw farther than 1e-12 from every
 * boundary); everything else -- about one position in 1e11, rows with negative or non-finite items, all-zero rows --
 * takes the serial evaluation in the reference's operation order.  on != 0 sends EVERY position down the serial path
 * (verification of the above: both settings must give identical grids); default 0.  Applies to later launches. */
int mcgp_replay_serial_grid(mcgp_handle h, int on);
/* Device-resident form (tapes and outputs are device pointers; race block from mcgp_upload_races). */
int mcgp_launch_replay(mcgp_handle h, uint64_t n_sims, const double* u_py_dev, const double* z_dev,
                       const double* u_np_dev, const int64_t* off_dev, uint64_t* hist_dev,
                       uint8_t* finish_dev, double* times_dev, int16_t* dnf_lap_dev, uint8_t* grid_dev,
                       int64_t* used_dev, int32_t* status_dev /* nullable: set to MCGP_ETAPE on overrun */,
                       void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* MCGP_H */
