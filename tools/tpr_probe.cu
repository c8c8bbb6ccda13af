// Throughput probe of an ALTERNATIVE mapping: one THREAD per simulated race (32 races per warp), the cars of a race in
// shared-memory columns [item][thread] (conflict-free for any per-thread car index), the running order kept as a
// per-thread permutation and repaired by insertion sort.  It implements the "floor" model of DESIGN.md ("nothing but
// lap times, ordering and a first-pass test on a fixed grid": the shipped warp-per-race kernel does 116 M races/s on
// it) with the native kernel's arithmetic and draw schedule (Philox4x32-7, one call per car and lap pair, Box-Muller
// pair, FP32 times relative to the leader, dirty air, DRS), so that the two mappings can be compared on the same work.
// A measurement tool, not a product path: no pit stops, events, overtake re-writes or grid sampling.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -o tools/tpr_probe tools/tpr_probe.cu
//   tools/tpr_probe [n_sims] [threads_per_block]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../monte-carlo-gp_b200/csrc/native_math.cuh"

using namespace mcgp;

constexpr int N = 20, LAPS = 57;

struct ProbeParams {
    float pc[N], eff[N], sigma[N], pace[N], deg[N];
    float drs_delta, dirty_thr, dirty_pen, ovt_thr;
};
__constant__ ProbeParams P;

template <int TPB>
__global__ void __launch_bounds__(TPB) tpr_probe_kernel(unsigned long long n_sims, const __grid_constant__ PhiloxKeys key,
                                                        unsigned long long* __restrict__ hist, unsigned long long* __restrict__ eligible_total) {
    // columns: item index major, thread minor -> bank = thread % 32 whatever car a thread touches
    __shared__ float T[N][TPB], LAST[N][TPB], AL[N][TPB], ZN[N][TPB], OP[N][TPB];   // OP: overtake pace of the car at its current tyre age
    __shared__ uint32_t U1[N][TPB], U2[N][TPB];   // overtake draws of the pair's even / odd lap
    __shared__ uint8_t AGE[N][TPB], FLG[N][TPB], ORD[N][TPB];
    __shared__ unsigned int hist_s[N * N];
    const int tid = threadIdx.x;
    for (int i = tid; i < N * N; i += TPB) hist_s[i] = 0;
    __syncthreads();
    unsigned long long elig = 0;
    for (unsigned long long sim = (unsigned long long)blockIdx.x * TPB + tid; sim < n_sims; sim += (unsigned long long)gridDim.x * TPB) {
        const uint32_t sim_lo = (uint32_t)sim, sim_hi = (uint32_t)(sim >> 32);
        // lap 1 on a fixed grid (slot == driver index)
#pragma unroll 4
        for (int d = 0; d < N; d++) {
            const uint4 w = philox4x32_10(sim_lo, sim_hi, (1u << 8) | (uint32_t)d, 0u, key);
            float z1, z2;
            fast_normal2(w.y, w.z, z1, z2);
            const float age = d < 10 ? 4.0f : 0.0f;
            float x = __fmaf_rn(age, P.eff[d], P.pc[d]);
            x = __fmaf_rn(P.sigma[d], z1, x);
            const float pf = fminf(1.5f, __fmaf_rn(0.1f, (float)(d + 1), 0.5f));
            T[d][tid] = __fmaf_rn(-0.5f, __fmul_rn(pf, z2), x);
            LAST[d][tid] = 0.0f;
            AL[d][tid] = 0.0f;
            AGE[d][tid] = (uint8_t)(d < 10 ? 5 : 1);
            OP[d][tid] = __fmaf_rn(age + 1.0f, P.deg[d], P.pace[d]);
            FLG[d][tid] = 0;
            ORD[d][tid] = (uint8_t)d;
        }
        float fuel = 0.0f;
        for (int lap = 1; lap <= LAPS; lap++) {
            if (lap >= 2) {
                // ---- per-car lap ----
                fuel = fminf(3.3f, __fadd_rn(fuel, 0.045f));
                const bool even = (lap & 1) == 0;
#pragma unroll 4
                for (int d = 0; d < N; d++) {
                    float z;
                    if (even) {
                        const uint4 w = philox4x32_10(sim_lo, sim_hi, ((uint32_t)lap << 8) | (uint32_t)d, 0u, key);
                        float zn;
                        fast_normal2(w.x, w.y, z, zn);
                        ZN[d][tid] = zn;
                        U1[d][tid] = w.z;
                        U2[d][tid] = w.w;
                    } else {
                        z = ZN[d][tid];
                    }
                    const float t = T[d][tid];
                    const uint32_t f = FLG[d][tid];
                    const float age = (float)AGE[d][tid];
                    float x = __fmaf_rn(age, P.eff[d], P.pc[d]);
                    x = __fadd_rn(x, -fuel);
                    x = __fmaf_rn((f & 1u) ? 1.0f : 0.0f, -P.drs_delta, x);
                    const float clean = __fmaf_rn(P.sigma[d], z, x);
                    const float held = fmaxf(__fadd_rn(clean, P.dirty_pen), AL[d][tid]);
                    const float last = ((f & 2u) && t < P.dirty_thr) ? held : clean;
                    T[d][tid] = __fadd_rn(t, last);
                    LAST[d][tid] = last;
                    AGE[d][tid] = (uint8_t)(AGE[d][tid] + 1);
                    OP[d][tid] = __fmaf_rn(age + 1.0f, P.deg[d], P.pace[d]);   // (per-driver constants are only ever read with the uniform loop index)
                }
            }
            // ---- ordering: insertion sort of the permutation by time ----
            {
                float prev_key = T[ORD[0][tid]][tid];
                for (int r = 1; r < N; r++) {
                    const int d = ORD[r][tid];
                    const float k = T[d][tid];
                    if (k < prev_key) {
                        int j = r - 1;
                        do {
                            ORD[j + 1][tid] = ORD[j][tid];
                            j--;
                        } while (j >= 0 && T[ORD[j][tid]][tid] > k);
                        ORD[j + 1][tid] = (uint8_t)d;
                        // prev_key stays: the element now at r is the old r-1
                    } else {
                        prev_key = k;
                    }
                }
            }
            // ---- first-pass overtake test (eligibility + draw, no re-write) and positions / DRS / dirty air ----
            {
                int a = ORD[0][tid];
                const float tl = T[a][tid];
                float t_prev = tl, last_prev = LAST[a][tid];
                float op_prev = OP[a][tid];
                T[a][tid] = 0.0f;
                FLG[a][tid] = 0;
                AL[a][tid] = 0.0f;
                for (int r = 1; r < N; r++) {
                    const int b = ORD[r][tid];
                    const float t = T[b][tid];
                    const float op = OP[b][tid];
                    const uint32_t f = FLG[b][tid];
                    const float delta = __fadd_rn(__fadd_rn(op_prev, -op), (f & 1u) ? P.drs_delta : 0.0f);
                    const uint32_t u16 = ((lap & 1) ? U2[b][tid] : U1[b][tid]) & 0xffffu;
                    if (delta > P.ovt_thr && (float)u16 < fminf(32768.0f, __fmul_rn(delta, 32768.0f))) elig++;
                    const bool drs = lap > 2 && __fadd_rn(t, -t_prev) < 1.0f;
                    FLG[b][tid] = (uint8_t)((drs ? 1u : 0u) | (last_prev > 0.0f ? 2u : 0u));
                    AL[b][tid] = last_prev;
                    T[b][tid] = __fadd_rn(t, -tl);
                    t_prev = t;
                    last_prev = LAST[b][tid];
                    op_prev = op;
                }
            }
        }
        for (int r = 0; r < N; r++) atomicAdd(&hist_s[ORD[r][tid] * N + r], 1u);
    }
    __syncthreads();
    for (int i = tid; i < N * N; i += TPB)
        if (hist_s[i]) atomicAdd(&hist[i], (unsigned long long)hist_s[i]);
    if (elig) atomicAdd(eligible_total, elig);
}

template <int TPB>
static void run(unsigned long long n_sims, int sm_count) {
    unsigned long long *hist, *elig;
    cudaMalloc(&hist, N * N * 8);
    cudaMalloc(&elig, 8);
    const PhiloxKeys key = philox_expand_key(42u, 0u);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tpr_probe_kernel<TPB>, TPB, 0);
    const int blocks = sm_count * per_sm;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaMemset(hist, 0, N * N * 8);
        cudaMemset(elig, 0, 8);
        cudaEventRecord(a);
        tpr_probe_kernel<TPB><<<blocks, TPB>>>(n_sims, key, hist, elig);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
    }
    unsigned long long h[N * N], e;
    cudaMemcpy(h, hist, sizeof(h), cudaMemcpyDeviceToHost);
    cudaMemcpy(&e, elig, 8, cudaMemcpyDeviceToHost);
    unsigned long long tot = 0;
    for (int i = 0; i < N * N; i++) tot += h[i];
    printf("{\"probe\": \"thread-per-race floor model\", \"threads_per_block\": %d, \"blocks_per_sm\": %d, \"warps_per_sm\": %d, \"sims\": %llu, "
           "\"ms\": %.3f, \"races_per_s\": %.4g, \"p_win_driver0\": %.4f, \"first_pass_successes_per_race\": %.3f, \"table_ok\": %s, \"cuda\": \"%s\"}\n",
           TPB, per_sm, per_sm * TPB / 32, n_sims, best, n_sims / (best * 1e-3), (double)h[0] / n_sims, (double)e / n_sims,
           tot == n_sims * N ? "true" : "false", cudaGetErrorString(cudaGetLastError()));
    cudaFree(hist);
    cudaFree(elig);
}

int main(int argc, char** argv) {
    const unsigned long long n_sims = argc > 1 ? strtoull(argv[1], nullptr, 10) : 4000000ull;
    const int tpb = argc > 2 ? atoi(argv[2]) : 64;
    ProbeParams p;
    for (int d = 0; d < N; d++) {   // the Bahrain-like synthetic inputs of SURVEY 8(d)
        p.pace[d] = 92.0f + 0.07f * d;
        p.deg[d] = 0.015f + 0.003f * (d % 10);
        p.pc[d] = p.pace[d] - 0.6f;                       // SOFT compound delta
        p.eff[d] = 0.08f * (p.deg[d] / 0.05f);
        p.sigma[d] = 0.15f + 0.01f * (d % 5);
    }
    p.drs_delta = 0.3f; p.dirty_thr = 2.0f; p.dirty_pen = 0.5f; p.ovt_thr = 0.6f;
    cudaMemcpyToSymbol(P, &p, sizeof(p));
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    if (tpb == 32) run<32>(n_sims, prop.multiProcessorCount);
    else run<64>(n_sims, prop.multiProcessorCount);   // (static shared memory: 540 B per thread, 48 KB per block at most)
    return 0;
}
