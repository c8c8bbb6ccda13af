/* TEST INFRASTRUCTURE -- scalar CPU mirror of the native-mode (Philox / FP32) kernel; see native_mirror.c. */
#ifndef NATIVE_MIRROR_H
#define NATIVE_MIRROR_H
#include "race_oracle.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Sims [sim_begin, sim_begin+n_sims) of one race with Philox key `seed` and stream id `stream`.
 * exact != 0 selects the IEEE-only normal generator (bit-identical to the kernel's MCGP_F_EXACT_NORMAL).
 * hist[n][n] is accumulated (+=); finish [n_sims][n] / times [n_sims][n] (float, time behind the winner,
 * by driver index) may be NULL. */
int orc_run_native(const orc_params* p, uint64_t seed, uint32_t stream, uint64_t sim_begin, int64_t n_sims, int exact,
                   int64_t* hist, uint8_t* finish, float* times, void* trace /* [n_sims][laps][n] 8-byte records, or NULL */);

/* The mirror's overtake paces x 2^15 (FP32), out[(age * n + driver)] for age < total_laps + 5: the strictly increasing
 * float images of the FP64 paces.  Exported so a CPU test can hold this construction against the library's own
 * (mcgp_pace_table): the two are written independently and must agree bit for bit. */
int orc_native_op32_table(const orc_params* p, float* out);

/* Philox rounds the mirror was built with (must equal the kernel's MCGP_PHILOX_ROUNDS), and the bare block function
 * for the Random123 known-answer vectors. */
int orc_native_philox_rounds(void);
void orc_philox4x32(int rounds, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif
