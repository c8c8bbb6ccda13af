// Native-mode race kernel for sm_100a: one warp per simulated race, one lane per driver.
//
// Behavioural source: reference src/simulation.py:59-560 (run_monte_carlo / simulate_race and the
// handlers; each step below cites its lines).  What is B200-native here and absent upstream:
//   * lane == driver index, so every per-driver parameter sits in a register for the whole launch;
//   * draws come from Philox4x32-7 (native_math.cuh) keyed by (seed; sim, lap pair, lane, stream): any sim range can be
//     launched on any GPU in any order and gives the same counts.  One call per lane serves TWO laps
//     (a Box-Muller pair + 2 x 3 overtake uniforms); the otherwise idle lanes 20..31 supply the extra
//     words and the race-event draws, so no lane computes Philox for nothing;
//   * a car's retirement lap is drawn once per race from the geometric law that the reference's per-lap
//     test `u < dnf_rate` (:194) induces, instead of one test per driver-lap;
//   * race times are FP32 *relative to the current leader* (re-based every lap in
//     update_positions), which keeps ~1e-5 s resolution where absolute FP32 time would have 5e-4 s;
//   * ordering: every car keeps its rank; a lap moves few cars far, so the new rank is the old one plus
//     the crossings counted against the two old neighbours on each side, then VERIFIED (strictly sorted
//     + a permutation) through a rank-indexed record array in shared memory; the rare misses fall back
//     to rank-by-counting over all keys (LDS.128 broadcast reads).  The same records give each car its
//     neighbour's time / overtake pace / last lap with one LDS.128 -- no inverse permutation;
//   * overtakes: pair conditions and draws are evaluated in parallel in rank space, the sequential
//     time re-write chain of :522-531 collapses to a closed form over runs of consecutive successes
//     (one REDUX.OR + bit scans), and the order after a pass is the old one with each run reversed
//     (verified with one neighbour compare);
//   * the one comparison of the loop whose operands carry no noise, `pace_delta > overtake_delta` (:514-521), is the
//     reference's FP64 decision bit for bit: the host tabulates it per (driver, tyre age, DRS) as strictly increasing
//     float images of the FP64 paces (device_params.h: PaceEntry), staged in shared memory; one LDS.128 per lap and one
//     float compare per pair (FP32 paces decided exact ties of the round-number BASELINE inputs differently);
//   * the lap loop is unrolled by PAIRS (the unit of the draw schedule), so which half of a pair's draws a lap uses is
//     a compile-time fact;
//   * the finish-position histogram accumulates in shared memory (uint32) and is flushed once per
//     block with 64-bit global atomics.
// The scalar CPU mirror of exactly this algorithm is oracle/native_mirror.c (test infrastructure).
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "device_params.h"
#include "native_math.cuh"

namespace mcgp {

// Block shape: 32 resident warps per SM (64 registers each) as ONE block of 32 warps.  A/B r2z on the headline launch:
// 4 blocks of 8 warps 89.5 M, 2 blocks of 16 warps 90.0 M, 1 block of 32 warps 91.2 M races/s (one copy of the parameter
// block and the pace tables per SM instead of four; a 505-lap race's 163 KB of pace tables still leaves all 32 warps
// resident where four 8-warp blocks would have been cut to one).
#ifndef MCGP_MIN_BLOCKS
#define MCGP_MIN_BLOCKS 1  // resident blocks per SM the register budget is tuned for
#endif
#ifndef MCGP_WARPS_PER_BLOCK
#define MCGP_WARPS_PER_BLOCK 32
#endif

constexpr int kWarpsPerBlock = MCGP_WARPS_PER_BLOCK;
constexpr int kThreads = kWarpsPerBlock * 32;
// the lap-histogram variant runs ONE block per SM holding all resident warps (at most 32: 1024 threads per block)
constexpr int kLapHistWarps = kWarpsPerBlock * MCGP_MIN_BLOCKS < 32 ? kWarpsPerBlock * MCGP_MIN_BLOCKS : 32;
constexpr unsigned FULL = 0xffffffffu;

// VSC tyre-age rollback probability 0.3 (src/simulation.py:392) as a 16-bit threshold
constexpr uint32_t kVscRoll16 = 19660u;  // floor(0.3 * 2^16)

// Ordering point between a warp's shared-memory writes and the reads of other lanes.  Every use below sits in
// warp-convergent code (all loop bounds and branches are provably uniform), so it compiles to a scheduling fence
// (a NOP), not a WARPSYNC.  An empty asm with a memory clobber is NOT enough: ptxas reorders a thread's LDS above
// its own STS to a different address.
#define WARP_FENCE() __syncwarp()
// Branches that are rarely taken (events, pit laps, ordering misses, ties): the hint makes ptxas lay their bodies out of the
// hot straight-line path, which is fetched through a 6 KB L0 instruction cache (A/B r2w: +0.27 %; -DMCGP_NO_COLD_HINTS to compare)
#ifndef MCGP_NO_COLD_HINTS
#define MCGP_UNLIKELY(x) __builtin_expect(!!(x), 0)
#else
#define MCGP_UNLIKELY(x) (x)
#endif
// XCHG_FENCE: the ordering point between a lane's store into a rank-indexed exchange array (records, window) and the
// loads of OTHER lanes' slots that follow, and between such loads and the next rewrite of the array.  Round 1 relied
// on the volatile accessors alone (a convergent warp issues its LDS / STS in program order); the CUDA memory model
// does not promise that, so the fence is spelled out.  Measured cost: 0.7 % (83.3 -> 82.7 M races/s, r2a A/B);
// -DMCGP_RELAXED_FENCES rebuilds the round-1 behaviour for comparison.
#ifdef MCGP_RELAXED_FENCES
#define XCHG_FENCE() ((void)0)
#else
#define XCHG_FENCE() __syncwarp()
#endif

// Shared-memory accessors on 32-bit shared addresses.  `volatile` keeps ptxas from reordering a lane's LDS above
// its own STS to another address (which it otherwise does: the two never alias for ONE thread), so inside
// warp-convergent code -- where a warp's LDS/STS issue in program order -- they need no __syncwarp() between
// them (that would compile to a NOP, but a NOP still costs an issue slot).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int OFF>
__device__ __forceinline__ float lds_f(uint32_t a) {
    float v;
    asm volatile("ld.volatile.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(OFF) : "memory");
    return v;
}
template <int OFF>
__device__ __forceinline__ void sts_f(uint32_t a, float v) {
    asm volatile("st.volatile.shared.f32 [%0+%1], %2;" ::"r"(a), "n"(OFF), "f"(v) : "memory");
}
template <int OFF>
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
    float4 v;
    asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a), "n"(OFF) : "memory");
    return v;
}
template <int OFF>
__device__ __forceinline__ void sts_f4(uint32_t a, float x, float y, float z, float w) {
    asm volatile("st.volatile.shared.v4.f32 [%0+%1], {%2, %3, %4, %5};" ::"r"(a), "n"(OFF), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

template <int OFF>
__device__ __forceinline__ float2 lds_f2(uint32_t a) {
    float2 v;
    asm volatile("ld.volatile.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(v.x), "=f"(v.y) : "r"(a), "n"(OFF) : "memory");
    return v;
}
template <int OFF>
__device__ __forceinline__ void sts_f2(uint32_t a, float x, float y) {
    asm volatile("st.volatile.shared.v2.f32 [%0+%1], {%2, %3};" ::"r"(a), "n"(OFF), "f"(x), "f"(y) : "memory");
}

__device__ __forceinline__ float2 lds_pair(uint32_t a) {  // read-only data: no ordering needed
    float2 v;
    asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}

// index of the highest / lowest set bit (x != 0): FLO, resp. BREV + FLO.SH, without the compiler's 31 - clz detour
__device__ __forceinline__ int msb(uint32_t x) {
    int r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
    return r;
}
__device__ __forceinline__ int lsb(uint32_t x) {
    int r;
    asm("bfind.shiftamt.u32 %0, %1;" : "=r"(r) : "r"(__brev(x)));
    return r;
}

// 1.0f iff a < b: a single FSET.BF on sm_100 (the integer-mask form costs FSETP + SEL)
__device__ __forceinline__ float lt_one(float a, float b) {
    float m;
    asm("set.lt.f32.f32 %0, %1, %2;" : "=f"(m) : "f"(a), "f"(b));
    return m;
}

// Rank of this lane's time among all 32 lanes (ascending, ties broken by lane).  S_t: 32 floats of warp scratch;
// every lane reads the keys back with broadcast LDS.128s and counts with FSET.BF + FADD (2 instr per key, one on
// each math pipe, four independent accumulation chains).  Lanes without a car hold huge, increasing "parked" times,
// so they rank behind every car in lane order; with NV4 == 5 only the first 20 keys are read and `park` (lane - 20
// on lanes >= 20, else 0) completes their count.
template <int NV4>
__device__ __noinline__ int rank_by_count(float t, float* S_t, int lane, int park) {
    S_t[lane] = t;
    WARP_FENCE();
    const float4* v4 = reinterpret_cast<const float4*>(S_t);
    const float4 v0 = v4[0];
    float c0 = lt_one(v0.x, t), c1 = lt_one(v0.y, t), c2 = lt_one(v0.z, t), c3 = lt_one(v0.w, t);
#pragma unroll
    for (int q = 1; q < NV4; q++) {
        const float4 v = v4[q];
        c0 += lt_one(v.x, t);
        c1 += lt_one(v.y, t);
        c2 += lt_one(v.z, t);
        c3 += lt_one(v.w, t);
    }
    int cnt = (int)((c0 + c1) + (c2 + c3)) + park;
    // exact ties are measure-zero events; detect them by a hole in the rank set and fix up
    const uint32_t seen = __reduce_or_sync(FULL, 1u << (cnt & 31));
    if (MCGP_UNLIKELY(seen != FULL)) {
#pragma unroll 1
        for (int j = 0; j < lane; j++) cnt += (S_t[j] == t) ? 1 : 0;
    }
    // (no trailing fence: the only caller, full_rank, fences right after publishing the records, before S_t can be rewritten)
    return cnt;
}

struct Tables {  // per-lane view of the compound tables in shared memory
    const NativeRace* R;
    int lane;
    __device__ __forceinline__ void load(int comp, float& eff, float& opt, float& pc) const {
        eff = R->eff_deg[comp][lane];
        opt = R->opt[comp][lane];
        pc = R->pc[comp][lane];
    }
};

// per-sim outputs beyond the count table (all optional)
struct NativeOutputs {
    uint8_t* finish;            // [race][sim][pos]  driver index
    float* times;               // [race][sim][driver] final gap to the winner
    uint8_t* grid;              // [race][sim][slot]  driver index on each grid slot (_sample_grid's result)
    TraceRecord* trace;         // [race][sim - trace_first][lap][driver]
    unsigned long long trace_first, trace_count;  // window of sims (indices within the launch) that are traced
    unsigned long long* laphist;  // [race][lap][driver][pos] running-position counts after every lap (kOut == 3)
    int lh_cells;                 // laps x n x n of one race's lap histogram (laps = the longest race of the batch)
};

// kOut: 0 = count table only, 1 = + finish/times, 2 = + per-lap trace, 3 = + per-lap position histogram (the on-chip
// reduction of the trace: BASELINE config 5's alternative output).  kWarps: warps per block -- the lap histogram lives
// in shared memory (laps x n x n counters: 91 KB for 57 laps x 20 cars), so that variant always runs ONE block of 32
// warps per SM (since round 2 the shape of every variant; a build with smaller blocks keeps the lap histogram at 32).
template <int NV4, bool kExact, int kOut, int kWarps>
__global__ void __launch_bounds__(kWarps * 32, (MCGP_MIN_BLOCKS * MCGP_WARPS_PER_BLOCK) / kWarps > 0 ? (MCGP_MIN_BLOCKS * MCGP_WARPS_PER_BLOCK) / kWarps : 1)
native_race_kernel(const NativeRace* __restrict__ races, const uint4* __restrict__ ptab, const int pt_rows, const int pt_stride,  // ptab: PacePair tables, two pairs per uint4
                   unsigned long long n_sims, unsigned long long sim_begin,
                   const __grid_constant__ PhiloxKeys key, unsigned long long* __restrict__ hist,
                   const __grid_constant__ NativeOutputs out, unsigned long long* __restrict__ work_counter) {
    constexpr bool kDetail = kOut == 1 || kOut == 2, kTrace = kOut == 2, kLapHist = kOut == 3;
    constexpr int kThr = kWarps * 32;
    constexpr bool kSmall = NV4 == 5;  // n <= 20: lanes 20..31 carry no car and lend their Philox words
    uint8_t* __restrict__ finish = out.finish;
    float* __restrict__ times = out.times;
    __shared__ NativeRace R;
    extern __shared__ __align__(16) uint4 PT[];  // overtake pace tables [drs][age][lane] of 8-byte pairs (device_params.h: PacePair) + padding
    __shared__ uint32_t hist_s[MCGP_LANES * MCGP_LANES];
    __shared__ __align__(16) float S_t_all[kWarps][32];
    __shared__ float S_w_all[kWarps][48];              // window scratch: times by OLD rank, -inf / +inf pads
    __shared__ float S_l_all[kWarps][34];              // last lap times by rank (update_positions), [0] = -inf pad
    // Records by rank: {time, overtake pace (NaN = retired)}, 8 bytes.  A rank-indexed access is a random permutation over the
    // banks; as 16-byte records (round 1: + last lap, padding) each exchange cost 8-10 shared-memory wavefronts instead of
    // the 4 a contiguous access needs, and the LSU data pipe ran at 77 % of its peak -- as busy as the issue port (ncu r2a).
    // The last lap of the car ahead is only needed once per lap, so it travels through the window scratch instead.
    __shared__ __align__(16) float2 S_rec_all[kWarps][36];

    // A block starts on race blockIdx.y of the batch and, when that race's sims are all claimed, hops to the next
    // race that still has work (a season batch mixes 44- and 78-lap races: without hopping the blocks of the short
    // races idle while the long ones finish).  One hop = reload the 7 KB parameter block, flush the count table.
    __shared__ int more_work;
    for (int hop = 0; hop < (int)gridDim.y; hop++) {
    const int race = (int)((blockIdx.y + hop) % gridDim.y);
    if (hop > 0) {
        __syncthreads();  // everybody is done with R and hist_s of the previous race
        if (threadIdx.x == 0) more_work = *reinterpret_cast<volatile unsigned long long*>(work_counter + race) < n_sims;
        __syncthreads();
        if (!more_work) continue;
    }
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(races + race);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&R);
        for (int i = threadIdx.x; i < (int)(sizeof(NativeRace) / 4); i += kThr) dst[i] = src[i];
        for (int i = threadIdx.x; i < MCGP_LANES * MCGP_LANES; i += kThr) hist_s[i] = 0;
        const int pt_n = pt_rows * pt_stride + MCGP_LANES / 2;  // uint4 words = pairs of entries: 2 tables + padding
        const uint4* psrc = ptab + (size_t)race * pt_n;
        for (int i = threadIdx.x; i < pt_n; i += kThr) PT[i] = psrc[i];
        if (kLapHist) {
            uint32_t* lh = reinterpret_cast<uint32_t*>(PT + pt_n);
            for (int i = threadIdx.x; i < out.lh_cells; i += kThr) lh[i] = 0u;
        }
    }
    __syncthreads();

    // warp-uniform values are broadcast from lane 0 so that the compiler can PROVE them uniform: every loop bound
    // and branch below is then convergent and the warp collectives need no divergence guards (BRA.DIV/WARPSYNC)
    int lane = threadIdx.x & 31;
    asm volatile("" : "+r"(lane));  // pin it to a register: under pressure the compiler re-reads SR_TID + masks per use
    const int warp = __shfl_sync(FULL, (int)(threadIdx.x >> 5), 0);
    const float kInf = __int_as_float(0x7f800000), kNaN = __int_as_float(0x7fc00000);
    float* S_t = S_t_all[warp];
    float* W = S_w_all[warp] + 8;   // W[-8..-1] = -inf, W[0..32) = times by rank, W[32..40) = +inf: no index guards needed
    float2* REC = S_rec_all[warp] + 2;  // REC[-1] = {-inf, NaN}: "no car ahead" blocks the pair and ends the order check
    W[lane - 8] = lane < 8 ? -kInf : kInf;
    if (lane < 16) W[lane + 24] = kInf;
    if (lane < 2) REC[lane - 2] = make_float2(-kInf, kNaN);
    __syncwarp();
    float* LST = S_l_all[warp] + 1;  // LST[-1] = -inf: the leader has no car ahead
    if (lane == 0) LST[-1] = -kInf;
    __syncwarp();
    const uint32_t w_sh = smem_u32(W), rec_sh = smem_u32(REC), lst_sh = smem_u32(LST);
    uint32_t grid_sh = smem_u32(&R.grid[0][lane]);
    asm volatile("" : "+r"(grid_sh));  // (kept in a register: recomputing a shared-window address costs 6 instructions)
    const int n = __shfl_sync(FULL, R.n, 0), L = __shfl_sync(FULL, R.total_laps, 0), track = __shfl_sync(FULL, R.track, 0);
    const bool grid_fixed = __shfl_sync(FULL, R.grid_fixed, 0) != 0;
    const bool is_car = lane < n;
    // Lanes without a car behave like cars parked behind the field for good: a huge time that grows with the lane,
    // retired from lap 0 (NaN overtake pace).  No per-lap code then needs an `is_car` guard.
    const float park_t = __fmul_rn(1e30f, (float)(lane + 1));
    const int park = (NV4 == 5 && lane >= 20) ? lane - 20 : 0;
    // overtake paces are carried pre-scaled by 2^15 (exact) so that the 16-bit uniform compares against them directly
    const float sigma = R.sigma[lane];
    // this lane's column of the pace table: entry [age][lane] sits at tb0 + age * rowb (lanes without a car read
    // a neighbouring entry or the padding row; they are retired from lap 0, so nothing of it is used)
    uint32_t tb0 = smem_u32(PT) + 8u * (uint32_t)lane;
    asm volatile("" : "+r"(tb0));  // (kept in a register, like grid_sh)
    const uint32_t rowb = 8u * (uint32_t)__shfl_sync(FULL, pt_stride, 0);
    const uint32_t tsel_on = rowb * (uint32_t)__shfl_sync(FULL, pt_rows, 0);  // byte offset of the with-DRS table
    // (values needed once per race or only on rare paths -- retirement law, pit loss, red / SC thresholds -- are read
    // from the shared parameter block where they are used: the hot loop has no register to spare for them)
    const float ndrs_delta = -R.drs_delta;
    const float ndrs32 = -R.drs32;
    const float dirty_thr = R.dirty_thr, dirty_pen = R.dirty_pen;
    // cumulative event thresholds (red | SC | VSC share one draw); only the event lane ever sees a non-zero ev_any
    const int ev_lane = kSmall ? 31 : 0;
    uint32_t ev_any = lane == ev_lane ? R.vsc_thr : 0u;
    asm volatile("" : "+r"(ev_any));  // keep it a per-lane register: one compare per lap instead of compare + lane test
    const uint32_t stream = __shfl_sync(FULL, R.stream, 0);
    const Tables tab{&R, lane};

    // Sims are handed out dynamically: a warp's first sim is its global warp index, every further one is claimed from
    // the race's counter (host-initialised to the number of warps) -- one atomic per race, issued a whole race ahead
    // of its use.  The warp scheduler favours some warps over others, so an even static split left 20 % of the
    // warp-slots idle at the end; which warp runs which sim does not matter (draws are keyed by the sim index).
    unsigned long long* const claim = work_counter + race;
    unsigned long long s = (unsigned long long)blockIdx.x * kWarps + (unsigned)warp;  // static first sim on the home race
    if (hop > 0) {
        if (lane == 0) s = atomicAdd(claim, 1ull);
        s = __shfl_sync(FULL, s, 0);
    }
    while (s < n_sims) {
        unsigned long long s_next = 0;
        if (lane == 0) s_next = atomicAdd(claim, 1ull);
        const unsigned long long sim = sim_begin + s;
        const uint32_t sim_lo = (uint32_t)sim, sim_hi = (uint32_t)(sim >> 32);

        // ---- _sample_grid (src/simulation.py:102-145): sequential draw without replacement -------
        int slot = 0;
        if (grid_fixed) {
            slot = R.fixed_slot[lane];
        } else {
            const uint4 wg = philox4x32_10(sim_lo, sim_hi, (uint32_t)lane, stream, key);
            const float ug = __fmul_rn((float)(wg.x >> 8), 5.9604644775390625e-08f);  // lane p holds position p's uniform
            bool remaining = is_car;
            uint32_t ga = grid_sh;  // &R.grid[pos][lane] as a 32-bit shared address: one add per position
            for (int pos = 0; pos < n; pos++, ga += 4u * MCGP_LANES) {
                float p = remaining ? lds_f<0>(ga) : 0.0f;
                float c = p;  // inclusive Hillis-Steele scan over lanes (the mirror replays this exact tree)
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
#ifdef MCGP_SCAN_PLAIN
                    float v = __shfl_up_sync(FULL, c, d);
                    if (lane >= d) c = __fadd_rn(c, v);
#else
                    // the shuffle's own "source lane in range" predicate guards the add: no lane compare per step
                    asm volatile("{\n\t.reg .pred q;\n\t.reg .f32 v;\n\tshfl.sync.up.b32 v|q, %0, %1, 0, 0xffffffff;\n\t@q add.rn.f32 %0, %0, v;\n\t}"
                                 : "+f"(c) : "r"(d));
#endif
                }
                const float total = __shfl_sync(FULL, c, 31);
                const float u = __shfl_sync(FULL, ug, pos);
                // np.random.choice (:137): the first lane whose inclusive sum exceeds u * total -- that lane is
                // necessarily a remaining driver with p > 0 (a lane that adds nothing cannot be the first to exceed)
                const uint32_t m = __ballot_sync(FULL, c > __fmul_rn(u, total));
                int sel = __ffs(m) - 1;
                if (MCGP_UNLIKELY(m == 0u)) {  // rare: total == 0 (:127-130, uniform over the remaining drivers) or u * total rounded up to total
                    const uint32_t rem_mask = __ballot_sync(FULL, remaining);
                    if (total > 0.0f) {  // the last remaining driver that has probability mass
                        sel = 31 - __clz(__ballot_sync(FULL, p > 0.0f));
                    } else {
                        const int nrem = __popc(rem_mask);
                        int k = (int)__fmul_rn(u, (float)nrem);
                        k = k < nrem - 1 ? k : nrem - 1;
                        uint32_t mm = rem_mask;
                        for (int i = 0; i < k; i++) mm &= mm - 1;
                        sel = __ffs(mm) - 1;
                    }
                }
                if (lane == sel) { slot = pos; remaining = false; }
            }
        }

        // ---- _initialize_cars (:244-273) -------------------------------------------------------
        int comp;
        float age;
        if (track == 2) { comp = 4; age = 0.0f; }
        else if (track == 1) { comp = 3; age = 0.0f; }
        else { comp = slot < 10 ? 0 : 1; age = slot < 10 ? 4.0f : 0.0f; }
        uint32_t used = 1u << comp;
        float eff, opt, pc;
        tab.load(comp, eff, opt, pc);

        // ---- _simulate_lap_1 (:275-311) --------------------------------------------------------
        // dnf_lap: the lap on which this car retires (0 on lanes without a car, > L: never).  Lap 1 uses 4 x the team
        // rate (:286-287); from lap 2 on the per-lap test u < rate (:194) makes the retirement lap geometric.
        int dnf_lap;
        float t, last = 0.0f, ahead_last = 0.0f;
        {
            const uint4 w = philox4x32_10(sim_lo, sim_hi, (1u << 8) | (uint32_t)lane, stream, key);
            const float ug = __fmul_rn((float)(2u * (w.w >> 9) + 1u), 5.9604644775390625e-08f);  // (0,1), exact
            float ln_u;
            if (kExact) ln_u = exact_log(ug);
            else { asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(ln_u) : "f"(ug)); ln_u = __fmul_rn(ln_u, 0.6931471805599453f); }
            const float dnf_scale = R.dnf_scale[lane];
            dnf_lap = 2 + (int)(dnf_scale <= MCGP_DNF_NEVER ? 70000.0f : fminf(__fmul_rn(ln_u, dnf_scale), 70000.0f));
            if (w.x < R.lap1_thr[lane]) dnf_lap = 1;
            if (!is_car) dnf_lap = 0;
            float z1, z2;
            if (kExact) exact_normal2(w.y, w.z, z1, z2); else fast_normal2(w.y, w.z, z1, z2);
            float x = __fmaf_rn(age, eff, pc);
            x = __fmaf_rn(sigma, z1, x);
            const float pf = fminf(1.5f, __fmaf_rn(0.1f, (float)(slot + 1), 0.5f));
            float sd = __fmul_rn(pf, z2);
            if (slot < 3) sd = fminf(sd, 1.0f);
            const float lt = __fmaf_rn(-0.5f, sd, x);
            // retired on lap 1: distinct sentinel times below every runner; lanes without a car: +inf for good
            t = !is_car ? park_t : dnf_lap == 1 ? -(float)(lane + 1) : lt;
            age = __fadd_rn(age, 1.0f);
        }
        uint32_t tba = tb0 + (uint32_t)(int)age * rowb;  // shared address of this car's pace-table entry at its current tyre age

        int drs_until = 2;  // DRS is off through this lap: laps 1-2 (:551), then through the laps an event disables it
        int rank;
        uint32_t bit;          // 1 << rank
        uint32_t ra;           // shared address of REC[rank]
        float2 prev;           // record of the car one rank ahead (REC[rank - 1]), valid whenever have_rank
        bool have_rank;        // warp-uniform: rank / bit / wa / ra / prev / REC describe the current times
        float drsf = 0.0f;                 // 1.0f while DRS is enabled for this car, else 0.0f: fma(drsf, -delta, x) == x - delta / x
        uint32_t tsel = 0u;                // byte offset of the pace table this car reads: 0 without DRS, tsel_on with
        float dthr = -kInf;                // dirty_thr while a car with a positive last lap runs ahead, else -inf (:208-216)
        float fuel = 0.0f;     // (110 - fuel_load) * 0.03 of the current lap: every runner burns 1.5 kg per lap (:221, Q11)
        const bool traced = kTrace && s >= out.trace_first && s - out.trace_first < out.trace_count;
        int tr_event = 0;
        bool tr_pit = false, tr_drs = false;

        auto set_rank = [&](int r) {
            rank = r;
            bit = 1u << (r & 31);
            ra = rec_sh + 8u * (uint32_t)r;
        };
        // all-cars rank by counting, then publish the records
        auto full_rank = [&](float op32) {
            set_rank(rank_by_count<NV4>(t, S_t, lane, park));
            sts_f2<0>(ra, t, op32);
            XCHG_FENCE();
            prev = lds_f2<-8>(ra);
            have_rank = true;
        };
        // live cars' rank among the runners, derived on demand from the all-cars rank (events, classification)
        auto live_position = [&](bool live) -> int {
            const uint32_t LM = __reduce_or_sync(FULL, live ? bit : 0u);
            return __popc(LM & (bit - 1u));
        };
        // _update_positions (:538-560), plus re-basing on the leader.  `prev` tells whether the car one rank ahead
        // runs (its overtake pace is not NaN); the runner whose car ahead does not run is the leader, and when there
        // is exactly one such runner every other runner's predecessor is simply the record already in `prev`.
        auto update_positions = [&](const int lap, const bool dnf_now) {
            const bool live = !dnf_now;
            const bool pd = prev.y != prev.y;
            const uint32_t B = __ballot_sync(FULL, live && pd);
            bool has_pred = live && !pd;
            // last lap times by rank through their own scratch array: the next write is a lap (and several fences) away,
            // so only the store -> load direction needs ordering here
            const uint32_t wl = lst_sh + 4u * (uint32_t)rank;
            sts_f<0>(wl, last);
            XCHG_FENCE();
            float tl, t_pred = prev.x, last_pred = lds_f<-4>(wl);  // (rank 0 reads the -inf pad: no car ahead)
            if (__popc(B) == 1) {
                tl = __shfl_sync(FULL, t, msb(B));
            } else if (B) {  // retired cars sit between runners (the lap of a retirement): search the live mask
                const uint32_t LM = __reduce_or_sync(FULL, live ? bit : 0u);
                tl = lds_f<0>(rec_sh + 8u * (uint32_t)(__ffs(LM) - 1));
                const uint32_t below = live ? (LM & (bit - 1u)) : 0u;
                has_pred = below != 0u;
                const int rp = has_pred ? 31 - __clz(below) : 0;
                t_pred = lds_f<0>(rec_sh + 8u * (uint32_t)rp);
                last_pred = lds_f<0>(lst_sh + 4u * (uint32_t)rp);
            } else {  // nobody left running: times stay as they are
                return;
            }
            const bool drs_now = has_pred && lap > drs_until && (__fadd_rn(t, -t_pred) < 1.0f);
            if (kTrace) tr_drs = drs_now;
            drsf = drs_now ? 1.0f : 0.0f;
            tsel = drs_now ? tsel_on : 0u;
            // dirty air needs a running car ahead whose previous lap time is positive (0.0 on lap 2, Q3): both folded
            // into the threshold the gap to the leader is compared with, so the lap itself tests `t < dthr` only
            ahead_last = last_pred;
            dthr = (has_pred && last_pred > 0.0f) ? dirty_thr : -kInf;
            t = __fadd_rn(t, -tl);  // (+inf stays +inf on lanes without a car)
        };
        // per-lap trace (BASELINE config 5): one 8-byte record per driver per lap, 8 n contiguous bytes per warp and lap;
        // the record pointer walks the sim's [lap][driver] block, the first word is packed with integer ops
        uint2* tr_ptr = nullptr;
        if (kTrace && traced)
            tr_ptr = reinterpret_cast<uint2*>(out.trace) +
                     ((unsigned long long)race * out.trace_count + (s - out.trace_first)) * (unsigned)L * (unsigned)n + lane;
        auto emit_trace = [&](const int lap, const bool dnf_now) {
            if (kTrace) {
                if (traced) {
                    const int pl = live_position(!dnf_now);
                    // byte 0 position (0 = retired), 1 compound, 2 tyre age, 3 flags (bit0 retired, bit1 DRS, bit2 pitted, bits4-5 event)
                    uint32_t w0 = dnf_now ? 0x01000000u : (uint32_t)(pl + 1);
                    w0 |= (uint32_t)comp << 8;
                    w0 |= min((uint32_t)(int)age, 255u) << 16;  // (a set can be older than 255 laps in a 300+-lap race: saturates)
                    w0 |= (tr_drs ? 0x02000000u : 0u) | (tr_pit ? 0x04000000u : 0u) | ((uint32_t)tr_event << 28);
                    if (is_car) *tr_ptr = make_uint2(w0, __float_as_uint(t));
                    tr_ptr += n;
                }
                tr_event = 0;
                tr_pit = false;
            }
        };

        // per-lap position histogram: every running car's position after the lap, one shared-memory atomic per car
        auto count_lap = [&](const int lap, const bool dnf_now) {
            if (kLapHist) {
                const int pl = live_position(!dnf_now);
                uint32_t* lh = reinterpret_cast<uint32_t*>(PT + pt_rows * pt_stride + MCGP_LANES / 2);
                if (is_car && !dnf_now) atomicAdd(&lh[((lap - 1) * n + lane) * n + pl], 1u);
            }
        };
        full_rank(dnf_lap <= 1 ? kNaN : 0.0f);
        update_positions(1, dnf_lap <= 1);
        emit_trace(1, dnf_lap <= 1);
        count_lap(1, dnf_lap <= 1);

        // One lap >= 2.  z: this lap's pace noise; u12: overtake uniforms of passes 1 / 2 in the low / high half;
        // ext: pass 3 of the pair's even / odd lap in its low / high half; ev: the word that decides the race event
        // (compared on the event lane only: the threshold is 0 on every other lane); evz: the two 16-bit VSC
        // tyre roll-back draws (even / odd lap).
        auto run_lap = [&](auto odd_c, const int lap, const float z, const uint32_t u12, const uint32_t ext, const uint32_t ev, const uint32_t evz) {
            constexpr bool odd = decltype(odd_c)::value;  // second lap of its pair (compile-time: the loop below is unrolled by pairs)
            const int rem = L - lap;
            // ---- race-interrupting events (:168-176): one draw on the cumulative thresholds ---------
            if (MCGP_UNLIKELY(__any_sync(FULL, ev < ev_any))) {  // rare (2.7 % of laps with the product probabilities)
                const uint32_t roll = odd ? evz >> 16 : evz & 0xffffu;
                const int code = ev < R.red_thr ? 1 : ev < R.sc_thr ? 2 : (roll < kVscRoll16 ? 4 : 3);
                const int e = __shfl_sync(FULL, code, ev_lane);
                const bool out_before = lap > dnf_lap;  // retired on an earlier lap (this lap's retirements still run here)
                const int pos_live = live_position(!out_before);
                if (kTrace) tr_event = e > 3 ? 3 : e;
                if (e == 1) {  // _handle_red_flag :397-431
                    if (!out_before) {
                        t = __fmul_rn(0.1f, (float)pos_live);
                        age = 0.0f;
                        comp = track == 2 ? 4 : track == 1 ? 3 : rem > 30 ? 2 : rem > 15 ? 1 : 0;
                        used |= 1u << comp;
                        tab.load(comp, eff, opt, pc);
                    }
                    drs_until = lap + 2;
                } else if (e == 2) {  // _handle_safety_car :334-376
                    if (!out_before) {
                        t = __fmul_rn(0.5f, (float)pos_live);
                        age = fmaxf(0.0f, __fadd_rn(age, -1.0f));
                    }
                    drs_until = lap + 2;
                } else {  // _handle_vsc :378-395
                    if (!out_before) {
                        t = __fmul_rn(t, 0.8f);
                        if (e == 4) age = fmaxf(0.0f, __fadd_rn(age, -1.0f));
                    }
                    drs_until = lap + 1;
                }
                tba = tb0 + (uint32_t)(int)age * rowb;
            }

            // ---- per-car lap (:186-223) --------------------------------------------------------
            const bool dnf = lap >= dnf_lap;
            // _calculate_lap_time :313-332
            fuel = fminf(3.3f, __fadd_rn(fuel, 0.045f));
            float x = __fmaf_rn(age, eff, pc);
            x = __fadd_rn(x, -fuel);
            x = __fmaf_rn(drsf, ndrs_delta, x);  // == x - drs_delta with DRS, x without (one rounding either way)
            const float clean = __fmaf_rn(sigma, z, x);
            // dirty air :208-216 (gap to the LEADER, Q3); ahead_last is 0 for the leader and on lap 2
            const float held = fmaxf(__fadd_rn(clean, dirty_pen), ahead_last);
            last = t < dthr ? held : clean;  // (a retired car's `last` is never read)
            if (!dnf) t = __fadd_rn(t, last);
            age = __fadd_rn(age, 1.0f);  // (retired cars age on: harmless, and one predicate less)
            tba += rowb;

            // ---- _handle_pit_stops (:433-494) ----------------------------------------------
            const bool pit = !dnf && age > opt && rem > 5;
            if (kTrace) tr_pit = pit;
            if (MCGP_UNLIKELY(__any_sync(FULL, pit))) {
                if (pit) {
                    t = __fadd_rn(t, R.pit_loss);
                    int nc = track == 2 ? 4 : track == 1 ? 3 : rem > 30 ? 2 : rem > 15 ? 1 : 0;
                    const uint32_t ud = used & 7u;
                    if (track == 0 && __popc(ud) == 1 && ((ud >> nc) & 1u)) {  // two-compound rule :481-488
                        const uint32_t avail = 7u & ~ud;
                        if (rem > 20) nc = (avail & 2u) ? 1 : R.pop_no_medium;
                        else nc = (avail & 1u) ? 0 : R.pop_no_soft;
                    }
                    comp = nc;
                    used |= 1u << comp;
                    age = 0.0f;
                    tba = tb0;
                    tab.load(comp, eff, opt, pc);
                }
            }

            // ---- _simulate_overtakes (:496-536): <= 3 passes in rank space ------------------
            // overtake pace (x 2^15); NaN for a retired car blocks both pairs it sits in (Q5)
            // The pair test `pace_delta > overtake_delta` (:514-521) is decided in FP64 on the host for every reachable
            // (driver, tyre age, DRS) and tabulated (device_params.h: PaceEntry): this car, chasing, may attack the car
            // ahead iff op32_ahead >= thr -- one float compare that reproduces the FP64 decision bit for bit.
            const float2 pe = lds_pair(tba + tsel);
            const float op32 = dnf ? kNaN : pe.x;
            const float thr = pe.y;
            const float opb = __fmaf_rn(drsf, ndrs32, op32);  // as the chasing car: DRS helps (:517-518)
            // Re-ordering from a good guess.  `rank` holds an order in which few cars are off by more than two places
            // (last lap's order after the lap times were added: true on 4 laps of 5; a run reversal that leapfrogged a
            // neighbour): count crossings against the two neighbours on each side only, then verify (a permutation +
            // strictly sorted); the caller falls back to the full count otherwise.
            auto window_place = [&]() {
                const uint32_t wa = w_sh + 4u * (uint32_t)rank;
                sts_f<0>(wa, t);
                XCHG_FENCE();
                const float a1 = lds_f<-4>(wa), a2 = lds_f<-8>(wa), b1 = lds_f<4>(wa), b2 = lds_f<8>(wa);
                // (tried: integer masks added with IADD3 instead of FSET.BF + FADD + F2I -- `set.lt.s32.f32` compiles to
                //  FSETP + SEL on sm_100, one instruction more per neighbour)
                const float moved = (lt_one(b1, t) + lt_one(b2, t)) - (lt_one(t, a1) + lt_one(t, a2));
                set_rank(rank + (int)moved);
                // (no fence needed here: REC was last READ before the fence above, W is not written again in this call)
                sts_f2<0>(ra, t, op32);
                const uint32_t cover = __reduce_or_sync(FULL, bit);
                XCHG_FENCE();
                prev = lds_f2<-8>(ra);
                have_rank = cover == FULL && !__any_sync(FULL, !(prev.x < t));
            };
            window_place();  // first ordering of the lap
            if (MCGP_UNLIKELY(!have_rank)) full_rank(op32);  // (from here on rank / prev always describe the current times)
            // one pass; returns true when another pass may follow
            auto one_pass = [&](const uint32_t u16) -> bool {
                const float delta = __fadd_rn(prev.y, -opb);  // pace_ahead - pace_behind (+ drs_delta), x 2^15
                // u16 * 2^-16 < min(0.5, delta / 2)   <=>   u16 < min(32768, delta * 32768)   (exact scaling); a retired
                // car on either side makes delta NaN, the NaN-propagating min keeps it and the compare fails (Q5)
                uint32_t mine;  // bit if this car overtakes the one ahead: two chained compares and ONE select
                asm("{\n\t.reg .pred s, q;\n\t.reg .f32 m;\n\tmin.NaN.f32 m, %1, 0f47000000;\n\tsetp.lt.f32 s, %2, m;\n\t"
                    "setp.ge.and.f32 q, %3, %4, s;\n\tselp.u32 %0, %5, 0, q;\n\t}"
                    : "=r"(mine) : "f"(delta), "f"((float)u16), "f"(prev.y), "f"(thr), "r"(bit));
                const uint32_t M = __reduce_or_sync(FULL, mine);
                if (M == 0u) return false;
                // closed form of the sequential re-write chain :522-531 over runs of consecutive successes: with j the
                // run start, k = rank - j and sn = "the car behind me succeeds too", the car ends on
                // T[j] - 0.1 (k + sn) + 0.3 sn = T[j] - 0.1 (k - 2 sn); a car outside every run has k = sn = 0.
                const uint32_t clear_below = ~M & ((bit << 1) - 1u);
                const int j = msb(clear_below);  // run start (bit 0 of M is never set)
                const uint32_t above = ~((M >> rank) >> 1);  // bit i clear <=> pair (rank+i, rank+i+1) swapped
                const int sn = (int)(~above & 1u);
                const float base = lds_f<0>(rec_sh + 8u * (uint32_t)j);
                t = __fmaf_rn(-0.1f, (float)(rank - j - 2 * sn), base);  // (k = sn = 0: base is the car's own time)
                // The new order is almost always the old one with every run [j, e] reversed (the re-written times
                // descend by 0.1 s inside a run); only a run that leapfrogs a neighbour outside it breaks that.
                // Verify the presumed order with one neighbour compare instead of re-counting all ranks.
                set_rank(j + lsb(above));  // j + e - rank with e = rank + lsb(above) the run end
                XCHG_FENCE();  // (every lane has read its run's base time before REC is rewritten)
                sts_f2<0>(ra, t, op32);
                XCHG_FENCE();
                prev = lds_f2<-8>(ra);
                have_rank = !__any_sync(FULL, !(prev.x < t));
                if (MCGP_UNLIKELY(!have_rank)) {
                    window_place();  // second chance before counting all ranks
                    if (!have_rank) full_rank(op32);
                }
                return true;
            };
            if (one_pass(u12 & 0xffffu))
                if (one_pass(u12 >> 16)) one_pass(odd ? ext >> 16 : ext & 0xffffu);
            update_positions(lap, dnf);
            emit_trace(lap, dnf);
            count_lap(lap, dnf);
        };

        // One Philox call per lane per lap PAIR (made on the even lap): x, y -> Box-Muller pair (cos: this lap, sin:
        // the next); z -> this lap's overtake passes 1, 2; w -> the next lap's; passes 3 come from a borrowed word.
        // Event lane (31; with more than 20 cars: words y, z, w of lane 0's second call): the event draws of the
        // two laps and the two 16-bit VSC roll-back draws.
        // The loop runs over lap PAIRS with both lap bodies spelled out: which half of the pair's draws a lap uses is
        // then a compile-time fact (no per-lap parity tests, no hand-over moves: -8 instructions per lap), and the
        // scheduler can start the pair's Philox call under the previous lap's tail.  The price is a 40 KB kernel
        // (the 32 KB L1.5 instruction cache: 0.7 instead of 0.5 stall cycles per issue on instruction fetch); measured
        // net +2.8 % (79.9 -> 82.1 M races/s).
#pragma unroll 1
        for (int lap = 2; lap <= L; lap += 2) {
            const uint4 w = philox4x32_10(sim_lo, sim_hi, ((uint32_t)lap << 8) | (uint32_t)lane, stream, key);
            uint4 ev = w;
            uint32_t ext;  // pass 3 of the pair's even / odd lap in its low / high half
            if (kSmall) {  // lanes 20..29 lend x (to drivers 0..9) and y (to drivers 10..19)
                const uint32_t e1 = __shfl_down_sync(FULL, w.x, 20), e2 = __shfl_down_sync(FULL, w.y, 10);
                ext = lane < 10 ? e1 : e2;
            } else {  // up to 32 cars: a second call per lane; lane 0's spare words are the event draws
                const uint4 xw = philox4x32_10(sim_lo, sim_hi, ((uint32_t)lap << 8) | (uint32_t)(lane + 32), stream, key);
                ext = xw.x;
                ev = make_uint4(xw.y, xw.z, xw.w, 0u);
            }
            float z, z_nx;  // this lap's / the next lap's pace noise
            if (kExact) exact_normal2(w.x, w.y, z, z_nx); else fast_normal2(w.x, w.y, z, z_nx);
            run_lap(std::false_type{}, lap, z, w.z, ext, ev.x, ev.z);
            if (lap < L) run_lap(std::true_type{}, lap + 1, z_nx, w.w, ext, ev.y, ev.z);
        }
        const bool dnf = L >= dnf_lap;
        const int pos_live = live_position(!dnf);

        // ---- final classification (:231-242) -----------------------------------------------------
        {
            const bool live = !dnf;
            const int n_live = __popc(__ballot_sync(FULL, live));
            // retired cars: later lap first, then larger time, then grid order (stable sort, reverse=True)
            const float tc = dnf_lap == 1 ? 0.0f : t;
            int worse = 0;
            for (uint32_t dm = __ballot_sync(FULL, is_car && dnf); dm; dm &= dm - 1u) {  // retired cars only (~1 per race)
                const int j = __ffs(dm) - 1;
                const int jl = __shfl_sync(FULL, dnf_lap, j);
                const float jt = __shfl_sync(FULL, tc, j);
                const int js = __shfl_sync(FULL, slot, j);
                const bool ahead = j != lane && (jl > dnf_lap || (jl == dnf_lap && (jt > tc || (jt == tc && js < slot))));
                worse += ahead ? 1 : 0;
            }
            const int pos = live ? pos_live : n_live + worse;
            if (is_car) {
                atomicAdd(&hist_s[lane * n + pos], 1u);
                if (kDetail) {
                    const unsigned long long o = ((unsigned long long)race * n_sims + s) * (unsigned long long)n;
                    if (finish) finish[o + pos] = (uint8_t)lane;
                    if (times) times[o + lane] = t;
                    if (out.grid) out.grid[o + slot] = (uint8_t)lane;
                }
            }
        }
        s = __shfl_sync(FULL, s_next, 0);
    }

    __syncthreads();
    for (int i = threadIdx.x; i < n * n; i += kThr) {
        const uint32_t v = hist_s[i];
        if (v) atomicAdd(&hist[(unsigned long long)race * n * n + i], (unsigned long long)v);
    }
    if (kLapHist) {
        const uint32_t* lh = reinterpret_cast<const uint32_t*>(PT + pt_rows * pt_stride + MCGP_LANES / 2);
        for (int i = threadIdx.x; i < out.lh_cells; i += kThr) {
            const uint32_t v = lh[i];
            if (v) atomicAdd(&out.laphist[(unsigned long long)race * out.lh_cells + i], (unsigned long long)v);
        }
    }
    }  // hop
}

// every race's claim counter starts behind the statically assigned first sims
__global__ void init_work_counters(unsigned long long* wc, int n_races, unsigned long long first) {
    for (int i = threadIdx.x; i < n_races; i += blockDim.x) wc[i] = first;
}

// ---- host-side launcher ------------------------------------------------------------------------
struct LaunchArgs {
    const NativeRace* races;
    const uint4* ptab;
    int pt_rows, pt_stride;
    unsigned long long n_sims, sim_begin;
    PhiloxKeys key;
    unsigned long long* hist;
    NativeOutputs out;
    unsigned long long* wc;
    int n_races, sm_count;
    size_t dyn_smem;
    cudaStream_t st;
};

template <int NV4, bool kExact, int kOut, int kWarps>
static cudaError_t launch_one(const LaunchArgs& a) {
    auto kern = native_race_kernel<NV4, kExact, kOut, kWarps>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.dyn_smem);
    if (e != cudaSuccess) return e;
    // persistent-style grid: as many blocks as are resident at once (the register budget allows MCGP_MIN_BLOCKS blocks of
    // MCGP_WARPS_PER_BLOCK warps per SM, a long race's pace table may allow fewer), split evenly over the races of the batch
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kWarps * 32, a.dyn_smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    const long long resident = (long long)a.sm_count * per_sm;
#ifdef MCGP_FLOOR_BLOCKS_PER_RACE
    long long bpr = resident / a.n_races;  // (never more blocks than fit at once: a waiting block could only start late)
#else
    // Blocks per race, rounded UP: with one block per SM, 148 SMs over a 24-race season would otherwise leave 4 SMs idle.
    // The few blocks beyond what is resident start when the first ones retire -- by then the hopping blocks have claimed
    // nearly everything, and what is left for a late block is its statically assigned first sim per warp.
    long long bpr = (resident + a.n_races - 1) / a.n_races;
#endif
    const long long need = (long long)((a.n_sims + kWarps - 1) / kWarps);
    if (bpr > need) bpr = need;
    if (bpr < 1) bpr = 1;
    const dim3 grid((unsigned)bpr, a.n_races), block(kWarps * 32);
    init_work_counters<<<1, 32, 0, a.st>>>(a.wc, a.n_races, (unsigned long long)bpr * kWarps);
    kern<<<grid, block, a.dyn_smem, a.st>>>(a.races, a.ptab, a.pt_rows, a.pt_stride, a.n_sims, a.sim_begin, a.key, a.hist, a.out, a.wc);
    return cudaGetLastError();
}

template <int NV4, bool kExact>
static cudaError_t launch_out(int kout, const LaunchArgs& a) {
    if (kout == 0) return launch_one<NV4, kExact, 0, kWarpsPerBlock>(a);
    if (kout == 1) return launch_one<NV4, kExact, 1, kWarpsPerBlock>(a);
    if (kout == 2) return launch_one<NV4, kExact, 2, kWarpsPerBlock>(a);
    return launch_one<NV4, kExact, 3, kLapHistWarps>(a);
}

int native_philox_rounds() { return MCGP_PHILOX_ROUNDS; }
size_t pace_pairs_per_race(int rows, int stride) { return 2 * (size_t)rows * stride + MCGP_LANES; }

cudaError_t launch_native(const NativeRace* races_dev, const PacePair* pace_dev, int pace_rows, int pace_stride, int n_races,
                          int max_n, unsigned long long n_sims, unsigned long long sim_begin, unsigned long long seed, bool exact,
                          unsigned long long* hist, uint8_t* finish, float* times, TraceRecord* trace,
                          unsigned long long trace_first, unsigned long long trace_count, unsigned long long* laphist,
                          unsigned long long* work_counter, int sm_count, cudaStream_t st, uint8_t* grid_out) {
    static_assert(2 * sizeof(PacePair) == sizeof(uint4), "pace table entries are staged two to a 16-byte word");
    const int kout = laphist ? 3 : trace ? 2 : (finish != nullptr || times != nullptr || grid_out != nullptr) ? 1 : 0;
    LaunchArgs a;
    a.races = races_dev; a.ptab = reinterpret_cast<const uint4*>(pace_dev); a.pt_rows = pace_rows; a.pt_stride = pace_stride;
    a.n_sims = n_sims; a.sim_begin = sim_begin;
    a.key = philox_expand_key((uint32_t)seed, (uint32_t)(seed >> 32));
    a.hist = hist;
    const int lh_cells = laphist ? (pace_rows - 5) * max_n * max_n : 0;  // laps of the longest race x n x n
    a.out = NativeOutputs{finish, times, grid_out, trace, trace_first, trace ? trace_count : 0ull, laphist, lh_cells};
    a.wc = work_counter; a.n_races = n_races; a.sm_count = sm_count; a.st = st;
    a.dyn_smem = pace_pairs_per_race(pace_rows, pace_stride) * sizeof(PacePair)  // both tables + padding (lanes without a car)
                 + (size_t)lh_cells * sizeof(uint32_t);
    if (max_n <= 20) return exact ? launch_out<5, true>(kout, a) : launch_out<5, false>(kout, a);
    return exact ? launch_out<8, true>(kout, a) : launch_out<8, false>(kout, a);
}

}  // namespace mcgp
