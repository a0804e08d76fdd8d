/* TEST INFRASTRUCTURE (oracle) -- not part of the product path.
 * CPU restatement of the reference's simulation over flat tables; see rxm_oracle.c. */
#ifndef RXM_ORACLE_H
#define RXM_ORACLE_H
#include <stdint.h>
#include "../include/rxm.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Returns 0/1 = the reference's match bit, or a negative value:
 *   -1 recursion deeper than RXM_ORACLE_MAX_DEPTH (the reference would overflow
 *      its stack on an epsilon cycle, automata.cpp:108-110 / mfa.cpp:143-147)
 *   -2 malformed table */
int rxm_oracle_match(const rxm_tables *t, const uint8_t *s, uint64_t n);

/* out_bits[i] = 0/1, or 2 where rxm_oracle_match returned < 0. */
void rxm_oracle_match_batch(const rxm_tables *t, const uint8_t *chars, const uint64_t *offsets,
                            uint64_t n, uint8_t *out_bits);

/* Per-string statistics of the last rxm_oracle_match call on this thread (MFA). */
typedef struct rxm_oracle_stats {
    uint64_t steps;          /* steps actually run (<= n + 1)                  */
    uint64_t evals;          /* configurations expanded                        */
    uint64_t max_live;       /* max occupied node slots after a step           */
    uint64_t max_depth;      /* max eps / absent-cell recursion depth          */
    uint64_t block_reads;    /* backreference block compares attempted         */
} rxm_oracle_stats;
void rxm_oracle_last_stats(rxm_oracle_stats *out);

#define RXM_ORACLE_MAX_DEPTH 256

#ifdef __cplusplus
}
#endif
#endif
