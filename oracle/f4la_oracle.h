/* f4la_oracle.h -- CPU restatement of the reference's GF(p) row reduction.
 * TEST INFRASTRUCTURE ONLY: used by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg as the checker; never linked into the product.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement
 *  (a) against the unmodified reference compiled into oracle/_ref/
 *      libneogb_ref.so on every matrix of recorded F4 runs, and
 *  (b) against the committed fixtures tests/golden/ (*.npz), which are matrices
 *      and echelon forms produced by the reference itself
 *      (tests/golden/make_golden.py).
 */
#ifndef F4LA_ORACLE_H
#define F4LA_ORACLE_H
#include "f4la_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct f4la_oracle_counters {
    uint64_t units;               /* coefficient applications (1 mul + 1 add) */
    uint64_t pivot_applications;  /* sparse AXPYs                             */
} f4la_oracle_counters;

/* Reduced row echelon form of the lower rows modulo the known pivots,
 * following exact_sparse_reduced_echelon_form_ff_{8,16,32}
 * (reference la_ff_32.c:2595-2770, la_ff_16.c:1046-1221, la_ff_8.c:1319-1494)
 * with one thread.  Honours F4LA_FLAG_NORMALFORM / F4LA_FLAG_PIVOT_BY_LEAD. */
int f4la_oracle_reduce(const f4la_flat_matrix *in, f4la_flat_result *out,
                       f4la_oracle_counters *cnt);

/* Modular inverse by extended Euclid (reference tools.h:95-122). */
uint32_t f4la_oracle_mod_inverse(uint32_t val, uint32_t p);

void f4la_oracle_free_result(f4la_flat_result *res);

#ifdef __cplusplus
}
#endif
#endif
