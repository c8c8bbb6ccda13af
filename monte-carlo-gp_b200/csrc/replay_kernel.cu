// Replay-mode race kernel for sm_100a: FP64, one warp per simulated race, one lane per GRID SLOT.
//
// Consumes the reference's own random draws (three per-sim tapes: U_py = random.random(), Z = standard
// normals behind np.random.normal, U_np = the uniform behind np.random.choice) in exactly the order
// reference src/simulation.py consumes them (SURVEY.md §8 "Draw-order specification"), and evaluates
// every expression in IEEE double with the reference's operation order.  This translation unit is
// compiled with -fmad=false: no FMA contraction anywhere, so finishing orders AND race times are
// bit-identical to CPython's.  lane == grid slot makes every "for car in cars" loop of the reference
// (grid order) a lane-ordered prefix count (ballot + popc) over the tape cursors, and makes
// the reference's stable-sort tie-break (list order) a plain (time, lane) comparison.
// What is B200-native here and absent upstream (each verified bit-exact on the reference fixtures):
//   * _sample_grid's per-position selection is a DISCRETE result: a 5-step warp scan of the raw items decides it, certified
//     against the reference's serial arithmetic by an error bound (the serial evaluation, kept out of line, runs only where
//     the draw is within 1e-12 of a boundary or the row is unusual);
//   * every car keeps its rank; a lap's first ordering is the old rank plus the crossings against two neighbours on each
//     side (float keys), an overtake pass's order the old one with each run of successes reversed -- guesses, VERIFIED in
//     FP64 ((time, slot) strictly increasing along a permutation: one reduction) and recounted over all keys on a miss;
//   * what a car needs from the car ahead (time, pace, last lap) sits in rank-indexed arrays with -inf / NaN pads in
//     front of rank 0; lanes without a car are parked cars (time +inf, rank == lane): no guards in the exchanges;
//   * the lap-loop tapes arrive through per-warp rings in shared memory, filled with cp.async a lap or more ahead;
//   * sims are claimed dynamically; the warp index is broadcast from lane 0 so that every loop bound and branch of the sim /
//     lap loops is provably uniform (no BRA.DIV guards around the warp collectives).
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "device_params.h"

namespace mcgp {

__global__ void init_work_counters(unsigned long long* wc, int n_races, unsigned long long first);  // native_kernel.cu

constexpr int kRWarps = 4;
constexpr int kRThreads = kRWarps * 32;
constexpr unsigned RFULL = 0xffffffffu;

struct Tape {
    const double* p;
    long long i, e;
    // lanes that do not consume a draw pass active=false: they neither read nor flag an overrun
    __device__ __forceinline__ double at(long long k, bool active, int& err) const {
        if (!active) return 0.5;
        const long long q = i + k;
        if (q < e) return p[q];
        err = 1;
        return 0.5;
    }
};

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(RFULL, v, src); }

// rank of (v, lane) among the lanes selected by `in_set` (warp-uniform mask), ascending, ties by lane.
// Keys go through 32 doubles of warp scratch: every lane reads them back with broadcast LDS.128s and counts the
// strictly smaller ones (2 instructions per key instead of a 2-shuffle round trip per key); exact ties -- possible
// here: VSC scaling, the 0.1 s floor of the overtake re-write -- show up as a hole in the rank set and are fixed up.
// NP = key pairs read (10 for fields of <= 20 cars, else 16: lanes outside the set hold +inf), fully unrolled.
template <int NP>
__device__ __forceinline__ int rank_set(double v, uint32_t in_set, int lane, double* S_key) {
    const bool member = (in_set >> lane) & 1u;
    S_key[lane] = member ? v : __longlong_as_double(0x7ff0000000000000ll);
    __syncwarp();
    int c0 = 0, c1 = 0;
    const double2* k2 = reinterpret_cast<const double2*>(S_key);
#pragma unroll
    for (int j = 0; j < NP; j++) {  // DSETP + predicated add per key (the plain C form compiles to three instructions)
        const double2 k = k2[j];
        asm("{\n\t.reg .pred p;\n\tsetp.lt.f64 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(c0) : "d"(k.x), "d"(v));
        asm("{\n\t.reg .pred p;\n\tsetp.lt.f64 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(c1) : "d"(k.y), "d"(v));
    }
    int cnt = c0 + c1;
    const int m = __popc(in_set);
    const uint32_t want = m >= 32 ? RFULL : ((1u << m) - 1u);
    const uint32_t seen = __reduce_or_sync(RFULL, member ? (1u << (cnt & 31)) : 0u);
    if (seen != want) {
#pragma unroll 1
        for (int j = 0; j < lane; j++) cnt += (S_key[j] == v) ? 1 : 0;
    }
    __syncwarp();
    return member ? cnt : lane;  // (lanes outside the set keep their own index: the all-cars order stays a permutation of 0..31)
}

// CPython 3.12 builtin sum() over items P[0..n) with kinds K[0..n)  (executed uniformly by all lanes)
__device__ double py_sum(const double* P, const uint8_t* K, int n, int& result_kind) {
    int i = 0;
    while (i < n && K[i] == 0) i++;
    if (i == n) { result_kind = 0; return 0.0; }
    double r = 0.0 + P[i];
    int k = K[i];
    i++;
    if (k == 1) {
        double c = 0.0;
        bool fell_out = false;
        for (; i < n; i++) {
            if (K[i] == 1) {  // Neumaier
                const double x = P[i];
                const double t = r + x;
                if (fabs(r) >= fabs(x)) c += (r - t) + x; else c += (x - t) + r;
                r = t;
            } else if (K[i] == 0) {
                r += 0.0;
            } else {
                if (c != 0.0 && isfinite(c)) r += c;
                r = r + P[i];
                i++;
                fell_out = true;
                break;
            }
        }
        if (!fell_out) {
            if (c != 0.0 && isfinite(c)) r += c;
            result_kind = 1;
            return r;
        }
    }
    for (; i < n; i++) r = r + (K[i] == 0 ? 0.0 : P[i]);
    result_kind = 2;
    return r;
}

// The same sum() when every item is an exact float or an int 0 (`items`: lanes whose item is a float; warp-uniform):
// int zeros add 0.0 to a non-negative partial sum (a no-op), so only the float items are visited, in index order, and
// with non-negative terms Neumaier's branch `abs(r) >= abs(x)` is max / min.  Bit-identical to py_sum on such lists.
// D: the float items compacted in index order (D[k] = k-th float item), cnt of them.
__device__ __forceinline__ double py_sum_floats(const double* D, int cnt, int& result_kind) {
    if (cnt == 0) { result_kind = 0; return 0.0; }
    double r = 0.0 + D[0], c = 0.0;
#pragma unroll 4
    for (int i = 1; i < cnt; i++) {
        const double x = D[i];
        const double t = r + x;
        // every lane walks the same list, so this branch is warp-uniform (and almost always taken: r is a running sum)
        if (r >= x) c += (r - t) + x; else c += (x - t) + r;
        r = t;
    }
    if (c != 0.0 && isfinite(c)) r += c;
    result_kind = 1;
    return r;
}

// One grid position of _sample_grid (:119-139) evaluated serially, in the reference's operation order: CPython's sum()
// (py_sum / py_sum_floats), the division, the 1e-9 renormalisation test, NumPy's left-to-right cumsum, the division by
// its last element, searchsorted(u, side='right').  lane d holds driver d's item (pd, kind kd).  Out of line: the
// kernel takes this path only where the parallel evaluation cannot certify the selection (see the call site).
__device__ __noinline__ int sample_slot_serial(double* S_p, double* S_c, double pd, uint8_t kd, uint32_t remaining, int n, int lane, double u) {
    const bool is_car = lane < n;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t m_float = __ballot_sync(RFULL, kd == 1), m_np = __ballot_sync(RFULL, kd == 2);
    int tk;
    double total;  // :123
    if (m_np == 0u) {  // floats and int zeros only: the float items, compacted in driver order
        if (kd == 1) S_c[__popc(m_float & lt_mask)] = pd;
        __syncwarp();
        total = py_sum_floats(S_c, __popc(m_float), tk);
    } else {  // np.float64 items: CPython's sum() leaves its compensated loop (Q12) -- the general state machine
        S_p[lane] = pd;
        __syncwarp();
        uint8_t* S_k = reinterpret_cast<uint8_t*>(S_c);
        S_k[lane] = kd;
        __syncwarp();
        total = py_sum(S_p, S_k, n, tk);
    }
    __syncwarp();
    if (total > 0) {  // :125-126
        pd = pd / total;
        kd = (tk == 2 || kd == 2) ? 2 : 1;
    } else {  // :127-130
        const int n_rem = __popc(remaining);
        if ((remaining >> lane) & 1u) { pd = 1.0 / (double)n_rem; kd = 1; } else { pd = 0.0; kd = 0; }
    }
    // :133-135 renormalise if abs(sum(probs) - 1) > 1e-9.  The exact sum() only matters when that fires:
    // a butterfly sum is within ~1e-15 of it, so unless it lands near the threshold the (usual) answer
    // "no renormalisation" is certain without replaying CPython's serial summation.
    double approx = pd;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) approx += shfl_d(approx, lane ^ d);
    if (!__all_sync(RFULL, fabs(approx - 1.0) < 1e-9 - 1e-12)) {
        uint8_t* S_k = reinterpret_cast<uint8_t*>(S_c);
        S_p[lane] = pd; S_k[lane] = kd;
        __syncwarp();
        const double prob_sum = py_sum(S_p, S_k, n, tk);  // :133
        __syncwarp();
        if (prob_sum > 0 && fabs(prob_sum - 1.0) > 1e-9) pd = pd / prob_sum;  // :134-135
    }
    S_p[lane] = pd;
    __syncwarp();
    // p.cumsum(); cdf /= cdf[-1]; searchsorted(u, side='right').  cumsum is a serial left-to-right loop:
    // every lane runs it, the running sums go through shared memory and each lane picks up its own.
    double acc = S_p[0];
    S_c[0] = acc;
#pragma unroll 4
    for (int d = 1; d < n; d++) { acc = acc + S_p[d]; S_c[d] = acc; }
    __syncwarp();
    const double cdf = (is_car ? S_c[lane] : acc) / acc;
    __syncwarp();
    int sel = __popc(__ballot_sync(RFULL, is_car && cdf <= u));
    if (sel > n - 1) sel = n - 1;
    return sel;
}

// 8-byte asynchronous copy global -> shared (LDGSTS: no register, no scoreboard wait at the issue site)
__device__ __forceinline__ void cp_async8(uint32_t dst_sh, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_sh), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

#ifndef MCGP_REPLAY_MIN_BLOCKS
#define MCGP_REPLAY_MIN_BLOCKS 7   // resident 128-thread blocks per SM the register budget is tuned for (r2p A/B: 5 blocks / 96 registers 27.6 M, 6 / 80 29.7 M, 7 / 72 30.7 M races/s)
#endif
template <int NP>
__global__ void __launch_bounds__(kRThreads, MCGP_REPLAY_MIN_BLOCKS)
replay_race_kernel(const ReplayRace* __restrict__ race, unsigned long long n_sims, const double* __restrict__ u_py,
                   const double* __restrict__ zt, const double* __restrict__ u_np, const long long* __restrict__ off,
                   unsigned long long* __restrict__ hist, uint8_t* __restrict__ finish, double* __restrict__ times,
                   int16_t* __restrict__ dnf_lap_out, uint8_t* __restrict__ grid_out, long long* __restrict__ used_out,
                   int* __restrict__ status, unsigned long long* __restrict__ work_counter, const int serial_grid) {
    // draws one lap can consume at most: 4 event draws + n retirement tests + 3 passes x (n - 1) pairs (U_py), n normals (Z)
    // The two lap-loop tapes reach the kernel through per-warp RINGS in shared memory, filled with cp.async a lap or more
    // ahead of their use: ring slot = (draw index within the sim) & (size - 1).  kPyAhead / kZAhead: how far beyond the
    // cursor the fill is kept (one 32-draw group per lap tops it up).
    constexpr int kPyRing = NP == 10 ? 128 : 256, kPyAhead = kPyRing - 32, kZRing = 128, kZAhead = 64;
    __shared__ ReplayRace R;
    __shared__ uint32_t hist_s[MCGP_LANES * MCGP_LANES];
    // Everything a warp exchanges through shared memory sits in ONE struct per warp, addressed from one pinned base
    // pointer with compile-time offsets (as eight separate arrays ptxas re-derived each array's address from the warp
    // index at every use to save registers: 8 % of all executed instructions, ncu r2m).
    struct WarpScratch {
        double p[32];              // grid sampling items, then the rank keys
        double c[32];              // running cumsum of the grid probabilities (first bytes: item kinds on the general sum() path)
        // rank-indexed copies of what a car needs from its neighbours in the order (the car ahead's pace, time and last
        // lap; the time at a run's start): one LDS each instead of a rank -> lane lookup plus two shuffles per double
        // ([-1] pads: "the car ahead of the leader" has time -inf and no pace, so rank 0 needs no special case)
        double cum_pad, cum[32], op_pad, op[32], last[32];
        // first ordering of a lap: the new times by OLD rank between -inf / +inf pads (win[2 + rank]) -- as FLOATS relative to
        // the previous leader: they only feed a guess that is verified in FP64, and a 4-byte rank-permuted exchange is one
        // conflict-free wavefront where an 8-byte one is two to four (ncu r2s: the LSU data pipe at 81 %, busier than the issue port)
        float win[36];
        // Per-warp rings of the two lap-loop tapes.  The draw sites of a lap (events, retirement tests, noise, <= 3
        // overtake passes) depend on each other's outcome, so reading the tapes from global memory where they are consumed
        // cost 5-6 dependent DRAM/L2 round trips per lap (ncu r2b: 2.3 stall cycles per issue on the long scoreboard);
        // one coalesced refill per lap still exposed one (r2n: 2.4 per issue once the rest of the lap got shorter).  Now
        // cp.async fills the rings a lap or more ahead and the lap reads shared memory only.
        double py[kPyRing];
        double z[kZRing];
        uint32_t inv[32];          // rank -> lane of the current all-cars order
    };
    __shared__ __align__(16) WarpScratch scratch[kRWarps];
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(race);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&R);
        for (int i = threadIdx.x; i < (int)(sizeof(ReplayRace) / 4); i += kRThreads) dst[i] = src[i];
        for (int i = threadIdx.x; i < MCGP_LANES * MCGP_LANES; i += kRThreads) hist_s[i] = 0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = __shfl_sync(RFULL, (int)(threadIdx.x >> 5), 0);
    // one live base register (the 32-bit shared-window address, pinned); every array below is a constant offset from it,
    // and converting it back with cvta keeps the accesses LDS / STS (a pinned generic pointer made them generic LD / ST)
    uint32_t ws_sh = (uint32_t)__cvta_generic_to_shared(&scratch[warp]);
    asm volatile("" : "+r"(ws_sh));
    WarpScratch* ws = reinterpret_cast<WarpScratch*>(__cvta_shared_to_generic(ws_sh));
    double* const S_p = ws->p;
    double* const S_c = ws->c;
    uint32_t* const S_inv = ws->inv;
    double* const S_cum = ws->cum;
    double* const S_op = ws->op;
    double* const S_last = ws->last;
    double* const S_py = ws->py;
    double* const S_z = ws->z;
    float* const W = ws->win + 2;
    for (int i = lane; i < 36; i += 32) ws->win[i] = __int_as_float(i < 2 ? 0xff800000 : 0x7f800000);
    if (lane == 0) { ws->cum_pad = __longlong_as_double(0xfff0000000000000ll); ws->op_pad = __longlong_as_double(0x7ff8000000000000ll); }
    __syncwarp();
    // warp-uniform values are broadcast from lane 0 so that the compiler can PROVE them uniform: loops and branches on them
    // are then convergent and need no divergence guards (BSSY / BSYNC around every block, BRA.DIV around every __syncwarp)
    const int n = __shfl_sync(RFULL, R.n, 0), L = __shfl_sync(RFULL, R.total_laps, 0), track = __shfl_sync(RFULL, R.track, 0);
    const bool is_car = lane < n;
    const uint32_t nmask = n >= 32 ? RFULL : ((1u << n) - 1u);
    uint32_t lt_mask = (1u << lane) - 1u;
    asm volatile("" : "+r"(lt_mask));  // (kept in a register: ptxas otherwise rebuilds it from SR_TID at each of its 14 uses per lap)
    int err = 0;

    // sims are claimed dynamically (counter host-initialised to the number of warps), one race ahead of their use:
    // races differ in length (retirements, overtake passes) and the warp scheduler is priority based
    for (unsigned long long s = (unsigned long long)blockIdx.x * kRWarps + warp; s < n_sims;) {
        unsigned long long s_next = 0;
        if (lane == 0) s_next = atomicAdd(work_counter, 1ull);
        // U_py and Z: this sim's stretch of the tape as a base pointer + 32-bit cursors (py_rel / z_rel: draws consumed,
        // py_len / z_len: draws the tape holds for this sim)
        const long long py0 = off[3 * s], z0 = off[3 * s + 1];
        const double* const py_base = u_py + py0;
        const double* const z_base = zt + z0;
        const long long py_len64 = off[3 * s + 3] - py0, z_len64 = off[3 * s + 4] - z0;
        const int py_len = py_len64 > 0x40000000 ? 0x40000000 : (int)py_len64, z_len = z_len64 > 0x40000000 ? 0x40000000 : (int)z_len64;
        int py_rel = 0, z_rel = 0;
        Tape np{u_np, off[3 * s + 2], off[3 * s + 5]};
        const long long np0 = np.i;
        // Ring fill.  *_st: draws of this sim (counted from its first) staged or in flight; *_ok: of those, landed and
        // visible to the whole warp.  A group = 32 consecutive draws, one per lane; a read past the end of the tape is
        // clamped to its last draw (the overrun itself is detected where the draws are consumed).
        int py_st = 0, py_ok = 0, z_st = 0, z_ok = 0;
        auto fill_py = [&]() {
            int q = py_st + lane;
            q = q < py_len ? q : py_len - 1;
            if (q >= 0) cp_async8(ws_sh + (uint32_t)offsetof(WarpScratch, py) + 8u * (uint32_t)((py_st + lane) & (kPyRing - 1)), py_base + q);
            py_st += 32;
        };
        auto fill_z = [&]() {
            int q = z_st + lane;
            q = q < z_len ? q : z_len - 1;
            if (q >= 0) cp_async8(ws_sh + (uint32_t)offsetof(WarpScratch, z) + 8u * (uint32_t)((z_st + lane) & (kZRing - 1)), z_base + q);
            z_st += 32;
        };
        auto landed = [&]() {  // everything issued so far has arrived and is visible to every lane
            cp_async_wait_all();
            __syncwarp();
            py_ok = py_st;
            z_ok = z_st;
        };
        cp_async_wait_all();  // (prefetches of the previous sim still in flight must not land on top of this sim's)
        __syncwarp();
#pragma unroll
        for (int g = 0; g < kPyAhead / 32; g++) fill_py();
#pragma unroll
        for (int g = 0; g < (kZAhead + 32) / 32; g++) fill_z();   // lap 1 takes up to 2 n normals

        // ---- _sample_grid (src/simulation.py:102-145) + RandomState.choice restated -------------
        int drv = 0;
        {
            // the n uniforms behind the n np.random.choice calls (:137): one coalesced read
            const double u_mine = np.at(lane, is_car, err);
            np.i += n;
            uint32_t remaining = nmask;  // by driver index
            for (int pos = 0; pos < n; pos++) {
                // lane d prepares driver d's item :119-122
                uint8_t kd = 0;
                double pd = 0.0;
                if (is_car && ((remaining >> lane) & 1u) && R.kind[lane][pos] != 0) { pd = R.grid[lane][pos]; kd = R.kind[lane][pos]; }
                const double u = shfl_d(u_mine, pos);
                // What :123-137 compute is  sel = #{i : cdf_i <= u}  with cdf = cumsum(p / sum(p)) / cumsum(...)[-1]: a DISCRETE
                // result.  With non-negative items every intermediate of the reference's serial evaluation (CPython's
                // compensated sum(), the division, NumPy's left-to-right cumsum, the final division) is within (n + 4) ulp
                // < 1e-14 of the true F_i = sum_{j<=i} p_j / sum_j p_j, and so is a warp scan of the raw items divided by its
                // last element.  So: scan the items (5 shuffle steps), compare prefix_i with u x total, and accept the count
                // when no prefix is within 1e-12 x total of the threshold -- the serial evaluation then provably selects the
                // same driver.  Otherwise (once in ~1e11 draws; negative / non-finite items; an all-zero row, :127-130)
                // the position is evaluated serially in the reference's operation order (sample_slot_serial).
                double pre = pd;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const double v = __shfl_up_sync(RFULL, pre, d);
                    if (lane >= d) pre += v;
                }
                const double tot = shfl_d(pre, 31);
                const double diff = pre - u * tot;
                const bool certain = pd >= 0.0 && (!is_car || fabs(diff) > 1e-12 * tot);   // (NaN anywhere: not certain)
                int sel;
                if (!serial_grid && tot > 1e-280 && tot < 1e280 && __all_sync(RFULL, certain))
                    sel = __popc(__ballot_sync(RFULL, is_car && diff <= 0.0));
                else
                    sel = sample_slot_serial(S_p, S_c, pd, kd, remaining, n, lane, u);
                if (lane == pos) drv = sel;
                remaining &= ~(1u << sel);  // :139
            }
        }

        // per-sim driver parameters (lane == grid slot, so they are looked up by the sampled driver)
        const double pace = R.pace[drv], deg = R.deg[drv], sigma = R.sigma[drv];
        const double dnf_rate = R.dnf_rate[drv], lap1_rate = R.lap1_rate[drv];
        const double driver_factor = deg > 0 ? deg / 0.05 : 1.0;  // :321

        // ---- _initialize_cars (:244-273) -------------------------------------------------------
        int comp, age;
        if (track == 2) { comp = 4; age = 0; }
        else if (track == 1) { comp = 3; age = 0; }
        else { comp = lane < 10 ? 0 : 1; age = lane < 10 ? 4 : 0; }
        uint32_t used = 1u << comp;
        bool dnf = !is_car, drs = false;
        int dnf_lap = 0, pos_live = 0;
        double t_lead = 0.0;  // time of the leading runner as of the last update_positions (warp-uniform)
        // Lanes without a car behave like cars parked behind the field for good: time +inf, retired from the start, rank ==
        // lane.  The rank-indexed exchanges below then need no `is_car` guard.
        double cum = is_car ? 0.0 : __longlong_as_double(0x7ff0000000000000ll), last = 0.0, ahead_last = 0.0;

        // _calculate_lap_time :313-332, strictly left to right
        auto lap_time = [&](int lap, double z) -> double {
            const double effective_deg = R.cdeg[comp] * driver_factor;
            const double tire_effect = (double)age * effective_deg;
            const double fuel0 = 110.0 - 1.5 * (double)(lap - 1);  // every runner burns 1.5 kg per lap (exact in binary)
            const double fuel = fuel0 > 0 ? fuel0 : 0.0;          // max(0, ...) :221 (Q11)
            const double fuel_effect = (110.0 - fuel) * 0.03;
            const double drs_gain = drs ? R.drs_delta : 0.0;
            const double noise = 0.0 + sigma * z;
            return pace + tire_effect - fuel_effect + R.cdelta[comp] - drs_gain + noise;
        };
        // _update_positions :538-560.  r_all: this car's rank among ALL cars by (time, grid slot) with S_inv the matching
        // rank -> lane map -- the order the last overtake pass worked on when it changed nothing (have_r), else derived
        // here.  The live order is its restriction to the runners (same tie-break), so no second sort is needed.
        auto update_positions = [&](int lap, bool drs_disabled, bool have_r, int& r_all) {
            if (!have_r) {
                r_all = rank_set<NP>(cum, nmask, lane, S_p);
                S_inv[r_all] = lane;
                S_cum[r_all] = cum;
            }
            S_last[r_all] = last;
            __syncwarp();
            const bool runner = !dnf;
            const uint32_t LM = __reduce_or_sync(RFULL, runner ? (1u << r_all) : 0u);  // ranks held by runners
            if (LM) {
                t_lead = S_cum[__ffs(LM) - 1];
                const uint32_t below = runner ? (LM & ((1u << r_all) - 1u)) : 0u;
                const int pr = below ? 31 - __clz(below) : (r_all & 31);   // rank of the runner ahead (own rank: unused)
                const double tp = S_cum[pr], lp = S_last[pr];
                if (!dnf) {
                    pos_live = __popc(below);
                    if (lap <= 2 || drs_disabled || pos_live == 0) drs = false;
                    else drs = (cum - tp) < 1.0;
                    ahead_last = pos_live > 0 ? lp : 0.0;
                }
            }
            __syncwarp();
        };

        // r: this car's rank among ALL cars by (time, grid slot) as of the last update_positions -- the guess the next
        // lap's first ordering starts from
        int r = lane;
        // ---- _simulate_lap_1 (:275-311) --------------------------------------------------------
        {
            landed();  // (issued before the grid was sampled)
            const double u = S_py[lane & (kPyRing - 1)];
            if (is_car && lane >= py_len) err = 1;
            py_rel = n;
            if (is_car && u < lap1_rate) { dnf = true; dnf_lap = 1; }
            const uint32_t surv = __ballot_sync(RFULL, !dnf);
            const int k = 2 * __popc(surv & lt_mask);
            const double z_noise = S_z[k & (kZRing - 1)], z_start = S_z[(k + 1) & (kZRing - 1)];
            if (!dnf && k + 1 >= z_len) err = 1;
            z_rel = 2 * __popc(surv);
            if (!dnf) {
                const double base_lap = lap_time(1, z_noise);
                double pf = 0.5 + (double)(lane + 1) * 0.1;
                if (pf > 1.5) pf = 1.5;
                double sd = 0.0 + pf * z_start;
                if (lane + 1 <= 3 && sd > 1.0) sd = 1.0;
                const double lt = base_lap - sd * 0.5;
                cum += lt;
                age += 1;
            }
            update_positions(1, true, false, r);
        }

        int drs_until = 0;
        for (int lap = 2; lap <= L; lap++) {
            // ---- this lap's draws: already in the rings ---------------------------------------------
            // pc / zc: draws consumed so far this lap.  A read past the end of the tape sets err (MCGP_ETAPE).
            // What was issued during the previous lap has had a lap to arrive; then top the rings up (typically one
            // group each) -- those land under this lap's work and are not waited for before the next lap.
            landed();
            while (py_st - py_rel < kPyAhead) fill_py();
            while (z_st - z_rel < kZAhead) fill_z();
            if (py_ok - py_rel < 4 + n) landed();  // (only after a lap that consumed nearly a whole ring)
            int pc = 0, zc = 0;
            // the k-th unread U_py draw: a ring read.  Whether the lap consumed more than the tape had left is checked once, below.
            auto py_draw = [&](int k, bool) -> double { return S_py[(py_rel + pc + k) & (kPyRing - 1)]; };

            // ---- events :168-176 (short-circuit draws) ------------------------------------------
            int ev = 0;
            {
                const double r1 = py_draw(0, true); pc++;
                if (r1 < R.red_p) ev = 1;
                else {
                    const double r2 = py_draw(0, true); pc++;
                    if (r2 < R.sc_p) ev = 2;
                    else {
                        const double r3 = py_draw(0, true); pc++;
                        if (r3 < R.vsc_p) ev = 3;
                    }
                }
            }
            ev = __shfl_sync(RFULL, ev, 0);  // (every lane read the same draws: make the uniformity provable)
            if (ev) {
                const uint32_t live_m = __ballot_sync(RFULL, !dnf);
                if (live_m) {
                    const double t0 = t_lead;  // the leader's time as of the last update_positions (nothing moved since)
                    const int rem = L - lap;
                    if (ev == 1) {  // _handle_red_flag :397-431
                        if (!dnf) {
                            cum = t0 + (double)pos_live * 0.1;
                            age = 0;
                            comp = track == 2 ? 4 : track == 1 ? 3 : rem > 30 ? 2 : rem > 15 ? 1 : 0;
                            used |= 1u << comp;
                        }
                    } else if (ev == 2) {  // _handle_safety_car :334-376 (lapped branch is dead code, Q6)
                        if (!dnf) {
                            cum = t0 + (double)pos_live * 0.5;
                            age = age - 1 > 0 ? age - 1 : 0;
                        }
                    } else {  // _handle_vsc :378-395
                        if (!dnf) {
                            const double gap = cum - t0;
                            cum = t0 + gap * 0.8;
                        }
                        const double r4 = py_draw(0, true); pc++;  // drawn only when somebody is still running :381-392
                        if (r4 < 0.3 && !dnf) age = age - 1 > 0 ? age - 1 : 0;
                    }
                    // :179 re-sorts after the handler; a VSC can create exact ties, so re-derive the car ahead
                    const int r = rank_set<NP>(cum, live_m, lane, S_p);
                    __syncwarp();
                    if (!dnf) S_inv[r] = lane;
                    __syncwarp();
                    const int pl = (!dnf && r > 0) ? (int)S_inv[r - 1] : lane;
                    const double lp = shfl_d(last, pl);
                    if (!dnf) { pos_live = r; ahead_last = r > 0 ? lp : 0.0; }
                    __syncwarp();
                }
                drs_until = ev == 3 ? lap + 1 : lap + 2;
            }

            // ---- per-car lap :186-223 (grid order == lane order) -------------------------------
            {
                const uint32_t live_m = __ballot_sync(RFULL, !dnf);
                const double u = py_draw(__popc(live_m & lt_mask), !dnf);
                pc += __popc(live_m);
                const bool was_live = !dnf;
                if (was_live && u < dnf_rate) { dnf = true; dnf_lap = lap; }
                const uint32_t surv = __ballot_sync(RFULL, !dnf);
                const int zk = __popc(surv & lt_mask);
                const double z = S_z[(z_rel + zk) & (kZRing - 1)];
                zc = __popc(surv);
                {   // (evaluated by every lane, kept by the runners: no divergent region around ~25 FP64 instructions)
                    const double clean = lap_time(lap, z);
                    const double dirty = clean + R.dirty_pen;
                    // time_behind_leader (:552, :371, :389, :417) is always cumulative_time minus the leader's time as of the last
                    // re-sort: recomputed here instead of carried across the lap (two registers)
                    const double tbl = cum - t_lead;
                    const bool in_dirty_air = tbl > 0 && ahead_last > 0 && tbl < R.dirty_thr;
                    const double lt = in_dirty_air ? (dirty >= ahead_last ? dirty : ahead_last) : clean;
                    const double cum_new = cum + lt;
                    cum = dnf ? cum : cum_new;
                    last = dnf ? last : lt;
                    age += dnf ? 0 : 1;
                }
            }

            // ---- _handle_pit_stops :433-494 ----------------------------------------------------
            {
                const int rem = L - lap;
                if (!dnf && (double)age > R.opt[comp][drv] && rem > 5) {
                    cum += R.pit_loss;
                    int nc = track == 2 ? 4 : track == 1 ? 3 : rem > 30 ? 2 : rem > 15 ? 1 : 0;
                    const uint32_t ud = used & 7u;
                    if (track == 0 && __popc(ud) == 1 && ((ud >> nc) & 1u)) {
                        const uint32_t avail = 7u & ~ud;
                        if (rem > 20) nc = (avail & 2u) ? 1 : R.pop_no_medium;
                        else nc = (avail & 1u) ? 0 : R.pop_no_soft;
                    }
                    comp = nc;
                    used |= 1u << nc;
                    age = 0;
                }
            }

            // ---- _simulate_overtakes :496-536 --------------------------------------------------
            bool have_r = false;  // r / S_inv / S_cum below describe the current times (the last pass changed nothing)
            {
                const double op = R.pace[drv] + (double)age * deg;  // :514-515 (raw driver deg)
                const double op_pub = dnf ? __longlong_as_double(0x7ff8000000000000ll) : op;  // NaN: a retired car blocks its pairs (Q5)
                // First ordering of the lap (:506), from a guess: a lap moves few cars far, so the new rank is the old one
                // plus the crossings counted against the two old neighbours on each side (times by OLD rank through W);
                // like the run-reversal guess of the later passes it is VERIFIED below and recounted on a miss.
                const float key = (float)(cum - t_lead);
                W[r] = key;
                __syncwarp();
                {
                    const float a1 = W[r - 1], a2 = W[r - 2], b1 = W[r + 1], b2 = W[r + 2];
                    r += (int)(b1 < key) + (int)(b2 < key) - (int)(key < a1) - (int)(key < a2);
                }
                for (int pass = 0; pass < 3; pass++) {
                    S_inv[r] = lane;
                    S_cum[r] = cum;
                    S_op[r] = op_pub;
                    __syncwarp();
                    {
                        // r is a guess (window placement / every run of the last pass reversed: the re-written times descend
                        // by 0.1 s inside a run); it is THE order iff it is a permutation along which (time, grid slot)
                        // increases strictly.  ONE reduction checks both: every rank must be claimed by a lane that sits
                        // strictly behind the car in the slot before it (two lanes on one rank leave another rank unclaimed).
                        const double c_prev = S_cum[r - 1];  // (rank 0 reads the -inf pad)
                        const int l_prev = (int)S_inv[(r - 1) & 31];
                        const bool ok = (c_prev < cum) | ((c_prev == cum) & (l_prev < lane));
                        if (__reduce_or_sync(RFULL, ok ? (1u << r) : 0u) != RFULL) {
                            __syncwarp();
                            r = rank_set<NP>(cum, nmask, lane, S_p);  // ALL cars, retired ones included (Q5)
                            S_inv[r] = lane;
                            S_cum[r] = cum;
                            S_op[r] = op_pub;
                            __syncwarp();
                        }
                    }
                    const double op_a = S_op[r - 1];  // (rank 0 reads the NaN pad)
                    double delta = op_a - op;   // NaN if the car ahead is retired (or there is none): every comparison below is false
                    if (drs) delta += R.drs_delta;
                    const bool cond = !dnf && delta > R.ovt_delta;
                    const uint32_t CM = __reduce_or_sync(RFULL, cond ? (1u << r) : 0u);
                    if (CM == 0u) { have_r = true; break; }  // nobody may attack: no draw is taken (:522-524), the pass changes nothing
                    if (py_rel + pc + __popc(CM) > py_ok) landed();  // (warp-uniform, rare: the lap outran what had landed at its start)
                    const double u = py_draw(__popc(CM & ((1u << r) - 1u)), cond);  // draws in sorted order :524
                    pc += __popc(CM);
                    double prob = delta / 2.0;
                    if (prob > 0.5) prob = 0.5;
                    const bool succ = cond && u < prob;
                    const uint32_t M = __reduce_or_sync(RFULL, succ ? (1u << r) : 0u);
                    if (M == 0u) { have_r = true; break; }
                    // sequential re-write chain :528-530, replayed op by op for bit-exactness
                    const uint32_t clear_below = ~M & ((2u << r) - 1u);
                    const int j = 31 - __clz(clear_below);
                    const int k = r - j;   // (lanes without a car: no bit of M at or above their rank, so j == r and sn == 0)
                    double a = S_cum[j & 31];
                    const int sn = (int)(((M >> r) >> 1) & 1u);
                    const int steps = k + sn;
                    const int max_steps = __reduce_max_sync(RFULL, steps);
                    for (int q = 0; q < max_steps; q++)
                        if (q < steps) { a = a - 0.1; if (!(a > 0.1)) a = 0.1; }  // max(0.1, ahead - 0.1)
                    if (steps > 0) cum = sn ? a + 0.3 : a;
                    // presumed order for the next pass: every run [j, e] reversed
                    r = j + __ffs(~((M >> r) >> 1)) - 1;   // j + e - r with e = r + (successes right behind this car)
                    __syncwarp();
                }
            }
            py_rel += pc;
            z_rel += zc;
            if (py_rel > py_len || z_rel > z_len) err = 1;  // the lap read past the end of a tape (MCGP_ETAPE)
            update_positions(lap, lap <= drs_until, have_r, r);
        }

        // ---- final classification :231-242 -------------------------------------------------------
        {
            const uint32_t live_m = __ballot_sync(RFULL, !dnf);
            const int n_live = __popc(live_m & nmask);
            int worse = 0;
            // retired cars only: sorted(key=(lap, cumulative_time), reverse=True) is stable -- equal keys keep grid order
            for (uint32_t dm = ~live_m & nmask; dm; dm &= dm - 1u) {
                const int j = __ffs(dm) - 1;
                const int jl = __shfl_sync(RFULL, dnf_lap, j);
                const double jt = shfl_d(cum, j);
                const bool ahead = j != lane && (jl > dnf_lap || (jl == dnf_lap && (jt > cum || (jt == cum && j < lane))));
                worse += ahead ? 1 : 0;
            }
            if (is_car) {
                const int pos = !dnf ? pos_live : n_live + worse;
                atomicAdd(&hist_s[drv * n + pos], 1u);
                const unsigned long long o = s * (unsigned long long)n;
                if (finish) finish[o + pos] = (uint8_t)drv;
                if (times) times[o + drv] = cum;
                if (dnf_lap_out) dnf_lap_out[o + drv] = (int16_t)dnf_lap;
                if (grid_out) grid_out[o + lane] = (uint8_t)drv;
            }
            if (used_out && lane == 0) {
                used_out[3 * s] = py_rel;
                used_out[3 * s + 1] = z_rel;
                used_out[3 * s + 2] = np.i - np0;
            }
        }
        s = __shfl_sync(RFULL, s_next, 0);
    }
    if (err && status) atomicExch(status, -4);

    __syncthreads();
    for (int i = threadIdx.x; i < n * n; i += kRThreads) {
        const uint32_t v = hist_s[i];
        if (v) atomicAdd(&hist[i], (unsigned long long)v);
    }
}

cudaError_t launch_replay(const ReplayRace* race_dev, int n_drivers, unsigned long long n_sims, const double* u_py, const double* z,
                          const double* u_np, const long long* off, unsigned long long* hist, uint8_t* finish,
                          double* times, int16_t* dnf_lap, uint8_t* grid, long long* used, int* status,
                          unsigned long long* work_counter, int sm_count, cudaStream_t st, bool serial_grid) {
    auto kern = n_drivers <= 20 ? replay_race_kernel<10> : replay_race_kernel<16>;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kRThreads, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    long long blocks = (long long)sm_count * per_sm;  // as many blocks as are resident at once; sims are claimed dynamically
    const long long need = (long long)((n_sims + kRWarps - 1) / kRWarps);
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    init_work_counters<<<1, 32, 0, st>>>(work_counter, 1, (unsigned long long)blocks * kRWarps);
    kern<<<(unsigned)blocks, kRThreads, 0, st>>>(race_dev, n_sims, u_py, z, u_np, off, hist, finish, times,
                                                 dnf_lap, grid, used, status, work_counter, serial_grid ? 1 : 0);
    return cudaGetLastError();
}

}  // namespace mcgp
