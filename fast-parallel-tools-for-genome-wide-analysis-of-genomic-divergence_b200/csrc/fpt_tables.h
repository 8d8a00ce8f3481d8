/*
 * fpt_tables.h — host-side construction of the two lookup tables the FET score kernel reads:
 * exact binomials C(n,k), n <= 67 (the domain in which the reference's u64 `binomial`,
 * fisher/cFisher.c:256-284, is exact) and log-factorials lgamma(k+1) up to the largest table total.
 */
#ifndef FPT_TABLES_H
#define FPT_TABLES_H

#include <math.h>
#include <stdint.h>
#include <vector>

#include "fpt_fet.cuh"

static inline std::vector<unsigned long long> fpt_build_binom_table() {
    std::vector<unsigned long long> t(FPT_BINOM_ENTRIES, 0ULL);
    /* Pascal's triangle in u64: every C(n,k) with n <= 67 fits (C(67,33) ~ 1.42e19 < 2^64) */
    unsigned long long row[FPT_FET_EXACT_MAX_N + 2] = { 1ULL };
    for (int n = 0; n <= FPT_FET_EXACT_MAX_N; n++) {
        if (n > 0) {
            for (int k = n; k >= 1; k--) row[k] = (k == n ? 0ULL : row[k]) + row[k - 1];
        }
        for (int k = 0; k <= n / 2; k++) t[fpt_binom_index(n, k)] = row[k];
    }
    return t;
}

static inline std::vector<double> fpt_build_lfact_table(int maxn) {
    std::vector<double> t((size_t)maxn + 1);
    for (int k = 0; k <= maxn; k++) t[k] = lgamma((double)k + 1.0);
    return t;
}

#endif
