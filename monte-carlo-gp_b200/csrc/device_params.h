// Device-side race parameter blocks, derived on the host from mcgp_race_params (include/mcgp.h).
#pragma once
#include <stdint.h>

#define MCGP_LANES 32
#define MCGP_NC 5
#define MCGP_DNF_NEVER (-1e30f)

// ---- native mode (FP32 / integer thresholds) -------------------------------------------------
// Everything a warp needs for one race.  ~6.6 KB, staged once per block into shared memory; the
// per-driver scalars then live in registers (lane == driver index) for the whole kernel.
struct NativeRace {
    int32_t n;            // drivers
    int32_t total_laps;
    int32_t track;        // 0 dry, 1 damp, 2 wet
    int32_t pop_no_medium, pop_no_soft;
    uint32_t stream;
    int32_t grid_fixed;   // 1: grid_probs is a deterministic permutation, fixed_slot[] is the grid
    int32_t _pad;
    float pit_loss, drs_delta, dirty_thr, dirty_pen;
    float drs32;                             // drs_delta x 2^15: the overtake probability runs on x 2^15 paces
    float _pad2;
    uint32_t red_thr, sc_thr, vsc_thr;       // CUMULATIVE floor(P * 2^32): red if w < red_thr, else SC if w < sc_thr, else VSC if w < vsc_thr
    float sigma[MCGP_LANES];                 // driver_variance
    float dnf_scale[MCGP_LANES];             // 1 / ln(1 - dnf_rate) <= 0: retirement lap = 2 + floor(ln(u) * dnf_scale);
                                             // MCGP_DNF_NEVER for rate <= 0 (laps >= 2, src/simulation.py:190-197)
    uint32_t lap1_thr[MCGP_LANES];           // 4 x team rate, lap 1
    float eff_deg[MCGP_NC][MCGP_LANES];      // compound_deg * (deg/0.05 if deg>0 else 1)   (:320-322)
    float opt[MCGP_NC][MCGP_LANES];          // pit window per compound, 0.85/1.1 scaled+truncated (:455-462)
    float pc[MCGP_NC][MCGP_LANES];           // base_pace + compound pace delta (:325), rounded to FP32 once
    float grid[MCGP_LANES][MCGP_LANES];      // [pos][driver] qualifying probabilities
    uint8_t fixed_slot[MCGP_LANES];          // grid_fixed: grid slot of each driver
};

// ---- overtake pace table (native mode) ---------------------------------------------------------
// The reference decides `pace_delta > overtake_delta` (src/simulation.py:514-521) in FP64 on
// pace = base_pace + tire_age * tire_deg -- integers times per-driver constants, so with round-number inputs the
// comparison lands EXACTLY on the threshold for some (driver, age) pairs and FP64 rounding decides it.  FP32
// arithmetic decides those ties differently (measured: a 3-5 % shift of single position probabilities), so the
// decision is tabulated on the host in FP64, op for op as upstream.  All reachable paces P[d][age] are sorted and
// mapped to a STRICTLY increasing sequence of floats f(P) ~ P * 2^15 (equal after rounding only if two paces are
// closer than 8e-6 s: then the later one is nudged up by one ulp), and for every (chasing driver, age, DRS) the
// smallest pace an ahead car must have for fl(fl(P_ahead - P_chasing) [+ drs_delta]) > overtake_delta is stored
// as its float: the decision is monotone in P_ahead, so on the GPU it is ONE float compare, f(P_ahead) >= thr,
// bit-identical to the FP64 decision; the same floats feed the (continuous) overtake probability.
// Layout: entry[age][lane], `stride` lanes per row, rows = total_laps + 5 (a tyre set is at most 4 + laps old).
struct PaceEntry {  // host-side / mcgp_pace_table form
    float op32;     // f(P[d][age])
    float thr0;     // this car chasing WITHOUT DRS may attack iff op32_ahead >= thr0 (+inf: never)
    float thr1;     // ... with DRS
    uint32_t _pad;
};
// Device form: pairs[drs][age][lane] -- two dense tables of 8-byte entries {op32, thr}, the first for a car without
// DRS, the second with, followed by MCGP_LANES entries of padding (lanes without a car read past their row).  A lane
// reads ONE 8-byte entry per lap at `age row + lane` of the table its DRS state selects: a contiguous 160 bytes per
// warp when the cars' tyres are equally old, i.e. 2 shared-memory wavefronts where the 16-byte entries of round 1
// needed 4-5 (the LSU data pipe was as busy as the issue port).
struct PacePair {
    float op32, thr;
};

// ---- replay mode (FP64, bit-exact) -----------------------------------------------------------
struct ReplayRace {
    int32_t n, total_laps, track, pop_no_medium, pop_no_soft, _pad;
    double pit_loss, ovt_delta, sc_p, vsc_p, red_p, drs_delta, dirty_thr, dirty_pen;
    double cdelta[MCGP_NC], cdeg[MCGP_NC];
    double opt[MCGP_NC][MCGP_LANES];         // already scaled/truncated per driver
    double pace[MCGP_LANES], deg[MCGP_LANES], sigma[MCGP_LANES];
    double dnf_rate[MCGP_LANES], lap1_rate[MCGP_LANES];   // lap1_rate = team_rate * 4.0
    double grid[MCGP_LANES][MCGP_LANES];     // [driver][pos]
    uint8_t kind[MCGP_LANES][MCGP_LANES];    // MCGP_ITEM_*
};

// ---- optional per-lap trace (BASELINE config 5; absent upstream) -------------------------------
// One record per (sim, lap, driver).  Same layout as mcgp_trace_record in include/mcgp.h.
struct TraceRecord {
    uint8_t position;  // 1-based running position after the lap, 0 = retired
    uint8_t compound;  // MCGP_SOFT .. MCGP_WET
    uint8_t tire_age;  // laps on the current set
    uint8_t flags;     // bit0 retired, bit1 DRS enabled for the next lap, bit2 pitted this lap, bits4-5 event (1 red, 2 SC, 3 VSC)
    float gap;         // seconds behind the leader (retired cars: frozen time relative to the leader)
};
