// Device-resident season loop and scoring: the callers either side of the race kernel.
//
// Reference behaviour restated here (all FP64, operation order as upstream so that everything but exp / pow is bit-identical
// to the host ports monte-carlo-gp_b200/{grid_model,ratings,scoring}.py, which are bit-exact to the reference):
//   season_grid_kernel   src/elo.py:124-141 (softmax of rating / 100) -> src/predictor.py:321-375 (teammate boost, form /
//                        circuit adjustment, Gaussian position spread) -> :377-407 (penalty shift) -> the `grid` block of the
//                        NEXT race's parameter block, in place, in device memory
//   season_elo_kernel    src/elo.py:45-122 (pairwise update of the quali ratings from the actual grid, of the race ratings
//                        from the actual finishing order; deltas against the ratings BEFORE the event)
//   score_counts_kernel  src/validation.py:82-130 (+ :133-158 binning): win / podium / points tallies, per-race Brier term,
//                        podium hits and the calibration bins straight from the [R][n][n] count tables
// A season is then one stream of launches -- grid rows(r) -> race kernel(r) -> "actual" race (one extra sim whose grid and
// finishing order play the real result) -> Elo update -> grid rows(r + 1) ... -- with no host round trip in between
// (mcgp_api.cu: mcgp_run_season).  exp() and pow() are CUDA's (<= 2 ulp), numpy's / libm's on the host differ in the
// last bits: tests hold ratings and grid rows to 1e-9 / 1e-12, everything integer (tallies, podium hits) exactly.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "device_params.h"

namespace mcgp {

// ---- ratings -> grid rows of one race ---------------------------------------------------------------------------------
// One warp, lane = driver.  Sequential sums run in driver order on every lane (they are Python sum()s upstream).
__global__ void season_grid_kernel(NativeRace* __restrict__ race, const double* __restrict__ quali, const int32_t* __restrict__ penalty,
                                   const double* __restrict__ teammate_delta, const double* __restrict__ form_score,
                                   const double* __restrict__ circuit_affinity, int n, double* __restrict__ rows_out) {
    __shared__ double S[MCGP_LANES];
    const int d = threadIdx.x;
    const bool car = d < n;
    // pole probabilities, src/elo.py:131-141
    const double scaled = car ? quali[d] / 100.0 : -INFINITY;
    double top = scaled;
    for (int o = 16; o > 0; o >>= 1) top = fmax(top, __shfl_xor_sync(0xffffffffu, top, o));
    const double w = car ? exp(scaled - top) : 0.0;
    S[d] = w;
    __syncwarp();
    double total = 0.0;                       // Python sum(): 0 + v0 + v1 + ... left to right
    for (int i = 0; i < n; i++) total += S[i];
    __syncwarp();
    double p = total > 0 ? w / total : 1.0 / (double)n;
    // teammate comparison, src/predictor.py:333-343
    const double delta = (car && teammate_delta) ? teammate_delta[d] : 0.0;
    if (delta != 0) p = p * fmax(0.5, fmin(1.5, 1 + (delta * 0.25)));
    S[d] = car ? p : 0.0;
    __syncwarp();
    total = 0.0;
    for (int i = 0; i < n; i++) total += S[i];
    __syncwarp();
    if (total > 0) p = p / total;
    // position spread, src/predictor.py:347-375
    const double form = ((car && form_score) ? form_score[d] : 0.0) * 0.15;
    const double circuit = ((car && circuit_affinity) ? circuit_affinity[d] : 0.0) * 0.10;
    double adjusted = p * (1 + form + circuit);
    adjusted = fmax(0.001, fmin(0.999, adjusted));
    const double sigma = fmax(1.0, (double)n / 4);
    const double centre = (1 - adjusted) * (double)n;
    double row[MCGP_LANES];
    double norm = 0.0;
    for (int pos = 0; pos < n; pos++) {
        const double x = (double)pos - centre;
        const double b = exp(-(x * x) / (2 * (sigma * sigma)));
        row[pos] = b;
        norm += b;
    }
    for (int pos = 0; pos < n; pos++) row[pos] = norm > 0 ? row[pos] / norm : 1.0 / (double)n;
    // grid penalties, src/predictor.py:377-407
    const int places = (car && penalty) ? penalty[d] : 0;
    if (places > 0 && n > 0) {
        double moved[MCGP_LANES];
        for (int i = 0; i < n; i++) moved[i] = 0.0;
        if (places >= n) {
            moved[n - 1] = 1.0;
        } else {
            for (int i = 0; i < n; i++) { const int to = i + places < n - 1 ? i + places : n - 1; moved[to] += row[i]; }
        }
        for (int i = 0; i < n; i++) row[i] = moved[i];
    }
    if (car) {
        for (int pos = 0; pos < n; pos++) {
            if (rows_out) rows_out[d * n + pos] = row[pos];
            race->grid[pos][d] = (float)row[pos];  // the same double -> float conversion as derive_native (mcgp_api.cu)
        }
    }
    if (d == 0) race->grid_fixed = 0;
}

// ---- pairwise Elo update of one event (src/elo.py:45-122) ---------------------------------------------------------------
// order[i] = driver listed i-th (finishing order / grid order); listed earlier = better (no ties by construction).
__device__ __forceinline__ double elo_expected(double ra, double rb) {  // src/elo.py:40-43
    const double e = fmax(-10.0, fmin(10.0, (rb - ra) / 400));
    return 1 / (1 + pow(10.0, e));
}
__device__ void elo_update(const uint8_t* order, const double* before, double* after, int n, double k, int i) {
    if (i < n) {
        const double ra = before[order[i]];
        double delta = 0.0;
        for (int j = 0; j < n; j++) {
            if (j == i) continue;
            const double actual = i < j ? 1.0 : 0.0;
            delta += k * (actual - elo_expected(ra, before[order[j]])) / (double)(n - 1);
        }
        after[order[i]] = before[order[i]] + delta;
    }
}
__global__ void season_elo_kernel(const uint8_t* __restrict__ grid_order, const uint8_t* __restrict__ finish_order,
                                  const double* __restrict__ q_before, const double* __restrict__ r_before,
                                  double* __restrict__ q_after, double* __restrict__ r_after, int n, double k) {
    const int i = threadIdx.x;
    if (n < 2) {
        if (i < n) { q_after[i] = q_before[i]; r_after[i] = r_before[i]; }
        return;
    }
    elo_update(grid_order, q_before, q_after, n, k, i);      // update_quali_ratings: lower "lap time" = earlier grid slot
    elo_update(finish_order, r_before, r_after, n, k, i);    // update_race_ratings: finishing position
}

// ---- scoring of count tables (src/validation.py:82-158) ---------------------------------------------------------------
// One block per race for the tallies; thread 0 of block 0 then walks the races in order for the floating-point parts
// (their sums are sequential upstream: bit-identical results need the same order).
struct ScoreOutputs {
    unsigned long long* tallies;   // [R][3][n]: win, podium (top 3), points (top 10) counts per driver
    double* brier;                 // [R] mean_d (p_d - [d == winner])^2 over the win probabilities; NaN = race skipped
    int32_t* podium_hits;          // [R] |predicted top 3 by podium probability  ∩  actual podium|, -1 = race skipped
    double* calib;                 // [3][10]: per calibration bin the number of (race, driver) pairs, sum of outcomes, sum of p
    int32_t* calib_bins;           // [1] number of bins used (min(10, max(2, pairs / 10)))
};

__global__ void score_counts_kernel(const unsigned long long* __restrict__ hist, int n_races, int n, unsigned long long n_sims,
                                    const int32_t* __restrict__ winner /* [R] driver or -1 */,
                                    const int32_t* __restrict__ podium /* [R][3] drivers or -1, may be NULL */, ScoreOutputs out) {
    for (int r = blockIdx.x; r < n_races; r += gridDim.x) {
        const unsigned long long* h = hist + (size_t)r * n * n;
        for (int d = threadIdx.x; d < n; d += blockDim.x) {
            unsigned long long win = h[d * n], pod = 0, pts = 0;
            for (int pos = 0; pos < n && pos < 10; pos++) { if (pos < 3) pod += h[d * n + pos]; pts += h[d * n + pos]; }
            out.tallies[((size_t)r * 3 + 0) * n + d] = win;
            out.tallies[((size_t)r * 3 + 1) * n + d] = pod;
            out.tallies[((size_t)r * 3 + 2) * n + d] = pts;
        }
    }
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    // (thread 0 of block 0 recomputes the few tallies it needs: no grid-wide barrier)
    int pairs = 0;
    for (int r = 0; r < n_races; r++) if (winner[r] >= 0) pairs += n;
    const int bins = pairs / 10 < 2 ? 2 : pairs / 10 > 10 ? 10 : pairs / 10;
    *out.calib_bins = bins;
    for (int b = 0; b < 30; b++) out.calib[b] = 0.0;
    const double step = 1.0 / (double)bins;   // np.linspace(0, 1, bins + 1): edge k = k * step
    for (int r = 0; r < n_races; r++) {
        const unsigned long long* h = hist + (size_t)r * n * n;
        // brier_score :82-106 on the win probabilities count / n_sims
        if (winner[r] < 0 || n == 0) {
            out.brier[r] = NAN;
        } else {
            double score = 0.0;
            for (int d = 0; d < n; d++) {
                const double pr = (double)h[d * n] / (double)n_sims;
                const double diff = pr - (d == winner[r] ? 1.0 : 0.0);
                score += diff * diff;
            }
            out.brier[r] = score / (double)n;
            // calibration_analysis :133-158: np.searchsorted(edges[1:-1], p) = number of interior edges < p
            for (int d = 0; d < n; d++) {
                const double pr = (double)h[d * n] / (double)n_sims;
                int b = 0;
                for (int e = 1; e < bins; e++) if ((double)e * step < pr) b = e;
                out.calib[b] += 1.0;
                out.calib[10 + b] += d == winner[r] ? 1.0 : 0.0;
                out.calib[20 + b] += pr;
            }
        }
        // podium_accuracy :109-130: the three highest podium probabilities (stable descending sort: ties keep driver order)
        if (!podium || podium[3 * r] < 0) {
            out.podium_hits[r] = -1;
        } else {
            int top[3] = {-1, -1, -1};
            double best[3] = {0.0, 0.0, 0.0};
            for (int d = 0; d < n; d++) {
                double c = 0.0;   // sum(race_probs[d].get(p, 0) for p in [1, 2, 3]), src/predictor.py:310-313
                for (int pos = 0; pos < 3 && pos < n; pos++) c += (double)h[d * n + pos] / (double)n_sims;
                for (int s = 0; s < 3; s++) {
                    if (top[s] < 0 || c > best[s]) {
                        for (int q = 2; q > s; q--) { top[q] = top[q - 1]; best[q] = best[q - 1]; }
                        top[s] = d; best[s] = c;
                        break;
                    }
                }
            }
            int hits = 0;
            for (int s = 0; s < 3; s++)
                for (int q = 0; q < 3; q++) if (top[s] >= 0 && top[s] == podium[3 * r + q]) { hits++; break; }
            out.podium_hits[r] = hits;
        }
    }
}

// The "actual" result of a race of the synthetic season: the drivers' finishing positions of its extra sim, in the form the
// scorer wants (winner, podium).
__global__ void season_actual_kernel(const uint8_t* __restrict__ finish_order, int32_t* __restrict__ winner, int32_t* __restrict__ podium, int n) {
    if (threadIdx.x == 0) {
        *winner = n > 0 ? finish_order[0] : -1;
        for (int q = 0; q < 3; q++) podium[q] = q < n ? finish_order[q] : -1;
    }
}

// ---- host-side launchers ------------------------------------------------------------------------------------------------
cudaError_t launch_season_grid(NativeRace* race_dev, const double* quali_dev, const int32_t* penalty_dev, const double* teammate_dev,
                               const double* form_dev, const double* circuit_dev, int n, double* rows_dev, cudaStream_t st) {
    season_grid_kernel<<<1, 32, 0, st>>>(race_dev, quali_dev, penalty_dev, teammate_dev, form_dev, circuit_dev, n, rows_dev);
    return cudaGetLastError();
}
cudaError_t launch_season_elo(const uint8_t* grid_order, const uint8_t* finish_order, const double* q_before, const double* r_before,
                              double* q_after, double* r_after, int n, double k, cudaStream_t st) {
    season_elo_kernel<<<1, 32, 0, st>>>(grid_order, finish_order, q_before, r_before, q_after, r_after, n, k);
    return cudaGetLastError();
}
cudaError_t launch_season_actual(const uint8_t* finish_order, int32_t* winner, int32_t* podium, int n, cudaStream_t st) {
    season_actual_kernel<<<1, 32, 0, st>>>(finish_order, winner, podium, n);
    return cudaGetLastError();
}
cudaError_t launch_score_counts(const unsigned long long* hist, int n_races, int n, unsigned long long n_sims, const int32_t* winner,
                                const int32_t* podium, unsigned long long* tallies, double* brier, int32_t* podium_hits,
                                double* calib, int32_t* calib_bins, cudaStream_t st) {
    ScoreOutputs out{tallies, brier, podium_hits, calib, calib_bins};
    const int blocks = n_races < 148 ? n_races : 148;
    score_counts_kernel<<<blocks, 32, 0, st>>>(hist, n_races, n, n_sims, winner, podium, out);
    return cudaGetLastError();
}

}  // namespace mcgp
