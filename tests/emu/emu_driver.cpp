/*
 * emu_driver.cpp — runs the product's kernels (csrc/fpt_*.cuh) on the CPU through tests/emu/cuda_emu.h.
 * TEST INFRASTRUCTURE ONLY: built by tests/emu/build.sh into tests/emu/libfpt_emu.so and loaded by
 * tests/test_emu_*.py so that kernel logic is exercised in the CPU-only test suite. The product never
 * loads this library.
 */
#include "cuda_emu.h"
FPT_EMU_DEFINE_GLOBALS

#include "fpt_rt.cuh"
#include "fpt_fet.cuh"
#include "fpt_css.cuh"
#include "fpt_css_eig.cuh"
#include "fpt_css_eig_reg.cuh"
#include "fpt_css_lanczos.cuh"
#include "fpt_css_k4.cuh"
#include "fpt_css_perm.cuh"
#include "fpt_css_perm3.cuh"
#include "fpt_css_perm_large.cuh"
#include "fpt_css_observed.cuh"
#include "fpt_tables.h"

template <class F>
static void run_grid(unsigned nb, unsigned nt, size_t smem, F f) {
    emu::launch(nb, nt, smem, [](void *p) { (*(F *)p)(); }, &f);
}

/* detector self-test for profiles/emu_sanitizers.sh: neighbouring threads exchange a value through dynamic shared memory, with
   or without the barrier between the write and the read. Under ThreadSanitizer the barrier-less variant MUST be reported (it shows
   that the instrumentation sees the kernels' shared-memory traffic), and overrun != 0 reads one word past the allocation for ASan. */
static void emu_selftest_kernel(int with_barrier, int overrun, int *out) {
    FPT_DYN_SMEM(smem);
    int *buf = reinterpret_cast<int *>(smem);
    buf[threadIdx.x] = (int)threadIdx.x;
    if (with_barrier) __syncthreads();
    out[threadIdx.x] = buf[(threadIdx.x + 1) % blockDim.x + (overrun ? blockDim.x : 0)];
}

extern "C" {

void emu_selftest(int with_barrier, int overrun, int *out) {
    run_grid(1, 64, 64 * sizeof(int), [=]() { emu_selftest_kernel(with_barrier, overrun, out); });
}

uint64_t emu_window_state(uint64_t seed, long long w, int stream) { return fpt_stream_state(seed, w, stream); }
uint64_t emu_lcg_skip(uint64_t s, uint64_t n) { return fpt_lcg_skip(s, n); }

void emu_fet_count_f64(const double *a, const double *b, long long nsnp, int asize, int bsize, int tile, int grid,
                       int *tables) {
    size_t smem = (((size_t)tile * asize + 15) & ~(size_t)15) + (size_t)tile * bsize + 16;
    run_grid(grid, 64, smem, [=]() { fpt_fet_count_kernel<double>(a, b, nsnp, asize, bsize, tile, (int4 *)tables); });
}

void emu_fet_count_i8(const signed char *a, const signed char *b, long long nsnp, int asize, int bsize, int tile,
                      int grid, int *tables) {
    size_t smem = (((size_t)tile * asize + 15) & ~(size_t)15) + (size_t)tile * bsize + 16;
    run_grid(grid, 64, smem, [=]() { fpt_fet_count_kernel<signed char>(a, b, nsnp, asize, bsize, tile, (int4 *)tables); });
}

void emu_fet_score(const int *tables, long long n, int maxn, int lf_in_smem, int force_log, int grid, double *scores) {
    std::vector<unsigned long long> binom = fpt_build_binom_table();
    std::vector<double> lf = fpt_build_lfact_table(maxn);
    size_t smem = FPT_BINOM_ENTRIES * 8 + (lf_in_smem ? ((size_t)maxn + 1) * 8 : 0);
    const unsigned long long *bp = binom.data();
    const double *lp = lf.data();
    run_grid(grid, 64, smem, [=]() {
        fpt_fet_score_kernel((const int4 *)tables, n, bp, lp, maxn, lf_in_smem, force_log, scores);
    });
}

/* the tile-sorting form of the score kernel (coverage-scale tables): needs 256 threads per CTA */
void emu_fet_score_sorted(const int *tables, long long n, int maxn, int lf_in_smem, int force_log, int grid, double *scores) {
    std::vector<unsigned long long> binom = fpt_build_binom_table();
    std::vector<double> lf = fpt_build_lfact_table(maxn);
    size_t smem = FPT_BINOM_ENTRIES * 8 + (lf_in_smem ? ((size_t)maxn + 1) * 8 : 0) + fpt_fet_sorted_smem_extra();
    const unsigned long long *bp = binom.data();
    const double *lp = lf.data();
    run_grid(grid, FPT_FET_SORT_THREADS, smem, [=]() {
        fpt_fet_score_sorted_kernel((const int4 *)tables, n, bp, lp, maxn, lf_in_smem, force_log, scores);
    });
}

int emu_fet_maxn(const int *tables, long long n) {
    int out = 0;
    int *po = &out;
    run_grid(2, 64, 0, [=]() { fpt_fet_maxn_kernel((const int4 *)tables, n, po); });
    return out;
}

int emu_window_table(const int *pos, long long nsnp, long long wbase, long long nwin, int regend, int wsize, int wstep, int threaded,
                     int *wleft, int *wright) {
    int mx = 0;
    int *pm = &mx;
    unsigned nb = (unsigned)((nwin + 63) / 64);
    run_grid(nb, 64, 0, [=]() { fpt_window_table_kernel(pos, nsnp, wbase, nwin, regend, wsize, wstep, threaded, wleft, wright, pm); });
    return mx;
}

void emu_fet_window(const double *snp_scores, const int *wleft, const int *wright, long long wbase, long long nwin, double perc,
                    uint64_t seed, const uint64_t *state_override, int max_npos, int use_hist, int grid,
                    double *out_score, double *out_std, unsigned char *out_flag) {
    int npad = 2;
    while (npad < max_npos) npad <<= 1;
    size_t smem = (size_t)npad * 8 + (use_hist ? (size_t)FPT_FET_HIST_LD * max_npos * 2 : 0);
    run_grid(grid, 128, smem, [=]() {
        fpt_fet_window_kernel(snp_scores, wleft, wright, wbase, nwin, perc, seed, state_override, npad, use_hist, out_score,
                              out_std, out_flag);
    });
}

void emu_css_pack_f64(const double *a, const double *b, long long nsnp, int asize, int bsize, int wpt, int grid,
                      unsigned *planes) {
    size_t smem = (((size_t)wpt * 32 * asize + 15) & ~(size_t)15) + (size_t)wpt * 32 * bsize + 16;
    run_grid(grid, 64, smem, [=]() { fpt_css_pack_kernel<double>(a, b, nsnp, asize, bsize, wpt, planes); });
}

void emu_css_pack_i8(const signed char *a, const signed char *b, long long nsnp, int asize, int bsize, int wpt,
                     int grid, unsigned *planes) {
    size_t smem = (((size_t)wpt * 32 * asize + 15) & ~(size_t)15) + (size_t)wpt * 32 * bsize + 16;
    run_grid(grid, 64, smem, [=]() { fpt_css_pack_kernel<signed char>(a, b, nsnp, asize, bsize, wpt, planes); });
}

void emu_css_absdiff(const double *a, const double *b, long long n, double *out) {
    run_grid(2, 64, 0, [=]() { fpt_css_absdiff_kernel(a, b, n, out); });
}

void emu_css_mds_large(const unsigned *planes, const double *absdiff, int m, const int *wleft, const int *wright,
                       long long nwin, int wch, int threads, int grid, double *X, double *evals, unsigned char *status, int *steps) {
    std::vector<double> gs((size_t)grid * fpt_css_mats_doubles(m));
    double *gp = gs.data();
    run_grid(grid, threads, fpt_lanczos_smem_bytes(m, wch), [=]() {
        fpt_css_mds_large_kernel(planes, absdiff, m, wleft, wright, nwin, wch, gp, X, evals, status, steps);
    });
}

/* large-cohort code route: popcount count codes (fpt_css_k4.cuh), then Lanczos on the codes */
void emu_css_mds_codes(const unsigned *planes, int m, const int *wleft, const int *wright, long long nwin, int threads, int grid,
                       double *X, double *evals, unsigned char *status, int *steps, int arithmetic) {
    const size_t stride = fpt_k4_window_stride(m);
    std::vector<unsigned char> codes((size_t)nwin * stride + 64, 0xAB);
    unsigned char *pc = codes.data();
    run_grid(grid, 64, fpt_k4_popc_smem(m), [=]() { fpt_css_k4_popc_kernel(planes, m, wleft, wright, nwin, pc, stride); });
    std::vector<double> basis((size_t)grid * fpt_lanczos_cta_scratch_bytes(m) / 8 + 1);
    double *pb = basis.data();
    run_grid(grid, threads, fpt_lanczos_smem_bytes(m, 0), [=]() {
        if (arithmetic) fpt_css_mds_codes_kernel<512, true>(pc, stride, m, wleft, wright, nwin, pb, X, evals, status, steps, 0);
    });
    run_grid(grid, threads, fpt_lanczos_smem_bytes(m, 0), [=]() {
        fpt_css_mds_codes_kernel<512, false>(pc, stride, m, wleft, wright, nwin, pb, X, evals, status, steps, arithmetic);
    });
}

void emu_css_mds_warp(const unsigned *planes, const double *absdiff, int m, const int *wleft, const int *wright,
                      long long nwin, int wch, int warps, int grid, double *X, double *evals, unsigned char *status) {
    /* phase A (tridiagonalisation) then phase B (eigenvectors), as the library launches them */
    std::vector<double> tri((size_t)nwin * 3 * m + 1), refl((size_t)nwin * ((size_t)m * (m - 1) / 2) + 1);
    double *pt = tri.data(), *pr = refl.data();
    run_grid(grid, 32 * warps, fpt_tridiag_work_bytes(m, wch) * warps, [=]() {
        fpt_css_tridiag_kernel(planes, absdiff, m, wleft, wright, nwin, wch, pt, pr, status);
    });
    run_grid(grid, 32 * warps, fpt_eigvec_work_bytes(m) * warps, [=]() {
        fpt_css_eigvec_kernel(m, nwin, pt, pr, status, X, evals);
    });
}

/* the same with phase A in registers (fpt_css_eig_reg.cuh), 3 <= m <= 48 */
void emu_css_mds_warp_reg(const unsigned *planes, int m, const int *wleft, const int *wright, long long nwin, int grid,
                          double *X, double *evals, unsigned char *status) {
    std::vector<double> tri((size_t)nwin * 3 * m + 1), refl((size_t)nwin * ((size_t)m * (m - 1) / 2) + 1);
    double *pt = tri.data(), *pr = refl.data();
    const int pad = fpt_tridiag_reg_pad(m);
    run_grid(grid, 32 * FPT_TREG_WARPS, fpt_tridiag_reg_work_bytes(m) * FPT_TREG_WARPS, [=]() {
        if (pad == 32) fpt_css_tridiag_reg_kernel<4, 8>(planes, m, wleft, wright, nwin, pt, pr, status);
        else if (pad == 40) fpt_css_tridiag_reg_kernel<5, 10>(planes, m, wleft, wright, nwin, pt, pr, status);
        else fpt_css_tridiag_reg_kernel<6, 12>(planes, m, wleft, wright, nwin, pt, pr, status);
    });
    run_grid(grid, 64, fpt_eigvec_work_bytes(m) * 2, [=]() {
        fpt_css_eigvec_kernel(m, nwin, pt, pr, status, X, evals);
    });
}

void emu_css_smacof(const unsigned *planes, const double *absdiff, int m, const int *wleft, const int *wright,
                    long long wbase, long long nwin, int wch, int mats_in_smem, int grid, int nruns, int random_start, uint64_t seed,
                    const uint64_t *state_override, int max_iters, double eps, const double *Xin, double *Xruns,
                    double *sigma_runs, int *iters_runs, unsigned char *status) {
    size_t smem = fpt_css_smem_bytes(m, wch, mats_in_smem);
    std::vector<double> gs(mats_in_smem ? 1 : (size_t)grid * fpt_css_mats_doubles(m));
    double *gp = mats_in_smem ? 0 : gs.data();
    run_grid(grid, 128, smem, [=]() {
        if (mats_in_smem) fpt_css_smacof_kernel<true>(planes, absdiff, m, wleft, wright, wbase, nwin, wch, 1, gp, nruns, random_start, seed,
                              state_override, max_iters, eps, Xin, Xruns, sigma_runs, iters_runs, status);
        else fpt_css_smacof_kernel<false>(planes, absdiff, m, wleft, wright, wbase, nwin, wch, 0, gp, nruns, random_start, seed,
                              state_override, max_iters, eps, Xin, Xruns, sigma_runs, iters_runs, status);
    });
}

void emu_css_pick(const double *Xruns, const double *sigma_runs, int m, int nruns, long long nwin,
                  const unsigned char *status, double *Xout) {
    run_grid(2, 64, 0, [=]() { fpt_css_pick_kernel(Xruns, sigma_runs, m, nruns, nwin, status, Xout); });
}

static unsigned long long emu_css_perm_impl(const double *Xall, int m, int asize, int bsize, long long wbase, long long nwin, const unsigned char *status,
                  int treshold, int runs, uint64_t seed, const uint64_t *state_override, int dist_in_smem,
                  int tracks_in_smem, int nthreads, int grid, int wide_tracks, int chain, int qbits, double *out_score, double *out_p,
                  int *out_hits, int *out_n) {
    int tb = wide_tracks ? 2 : 1;
    size_t smem = fpt_css_perm_smem_bytes(m, nthreads, tb, dist_in_smem, tracks_in_smem, qbits > 0);
    size_t per_cta = (dist_in_smem ? 0 : (size_t)m * m * 8) + (tracks_in_smem ? 0 : (((size_t)2 * nthreads * m * tb + 15) & ~(size_t)15));
    if (qbits > 0) per_cta += fpt_perm_large_sur_scratch(m);
    per_cta = (per_cta + 15) & ~(size_t)15;
    std::vector<double> gs(per_cta ? (size_t)grid * per_cta / 8 + 2 : 1);
    double *gp = per_cta ? gs.data() : 0;
    unsigned long long rechecks = 0, *pr = &rechecks;
    if (wide_tracks)
        run_grid(grid, nthreads, smem, [=]() {
            fpt_css_perm_kernel<unsigned short>(Xall, m, asize, bsize, wbase, nwin, status, treshold, runs, seed, state_override, chain,
                                                dist_in_smem, tracks_in_smem, gp, per_cta, qbits, out_score, out_p, out_hits, out_n, pr);
        });
    else
        run_grid(grid, nthreads, smem, [=]() {
            fpt_css_perm_kernel<unsigned char>(Xall, m, asize, bsize, wbase, nwin, status, treshold, runs, seed, state_override, chain,
                                               dist_in_smem, tracks_in_smem, gp, per_cta, qbits, out_score, out_p, out_hits, out_n, pr);
        });
    return rechecks;
}

void emu_css_perm(const double *Xall, int m, int asize, int bsize, long long wbase, long long nwin, const unsigned char *status,
                  int treshold, int runs, uint64_t seed, const uint64_t *state_override, int dist_in_smem,
                  int tracks_in_smem, int nthreads, int grid, int wide_tracks, int chain, double *out_score, double *out_p,
                  int *out_hits, int *out_n) {
    emu_css_perm_impl(Xall, m, asize, bsize, wbase, nwin, status, treshold, runs, seed, state_override, dist_in_smem, tracks_in_smem,
                      nthreads, grid, wide_tracks, chain, 0, out_score, out_p, out_hits, out_n);
}

/* the general kernel with the integer surrogate on (q and digit matrices in global scratch) */
unsigned long long emu_css_perm_sur(const double *Xall, int m, int asize, int bsize, long long wbase, long long nwin, const unsigned char *status,
                  int treshold, int runs, uint64_t seed, const uint64_t *state_override, int dist_in_smem,
                  int tracks_in_smem, int nthreads, int grid, int wide_tracks, int chain, int qbits, double *out_score, double *out_p,
                  int *out_hits, int *out_n) {
    return emu_css_perm_impl(Xall, m, asize, bsize, wbase, nwin, status, treshold, runs, seed, state_override, dist_in_smem, tracks_in_smem,
                             nthreads, grid, wide_tracks, chain, qbits, out_score, out_p, out_hits, out_n);
}

unsigned long long emu_css_perm3(const double *Xall, int m, int asize, int bsize, long long wbase, long long nwin,
                                 const unsigned char *status, int treshold, int runs, uint64_t seed, const uint64_t *state_override,
                                 int qbits, int grid, double *out_score, double *out_p, int *out_hits, int *out_n) {
    unsigned long long rechecks = 0, *pr = &rechecks;
    /* observed scores first, as fpt_api.cu does */
    run_grid(grid, FPT_OBS_WARPS * 32, fpt_css_observed_smem_bytes(m), [=]() {
        fpt_css_observed_kernel(Xall, m, asize, bsize, nwin, status, out_score);
    });
    if (fpt_css_perm3_ksteps(m) == 2)
        run_grid(grid, FPT_P3_T, fpt_css_perm3_smem_bytes(m), [=]() {
            fpt_css_perm3_kernel<2>(Xall, m, asize, bsize, wbase, nwin, status, treshold, runs, seed, state_override, qbits, out_score, out_p,
                                    out_hits, out_n, pr);
        });
    else
        run_grid(grid, FPT_P3_T, fpt_css_perm3_smem_bytes(m), [=]() {
            fpt_css_perm3_kernel<1>(Xall, m, asize, bsize, wbase, nwin, status, treshold, runs, seed, state_override, qbits, out_score, out_p,
                                    out_hits, out_n, pr);
        });
    return rechecks;
}

/* exact-quotient magic of the headline permutation kernel: returns the number of (n, r) pairs with a wrong remainder */
long long emu_p3_magic_check(const unsigned *rs, long long nr) {
    long long bad = 0;
    for (unsigned n = 2; n <= 64; n++) {
        const uint2 mg = fpt_p3_magic(n);
        for (long long k = 0; k < nr; k++) {
            const unsigned r = rs[k] & 0x7fffffffu;
            const unsigned rem = r - (unsigned)((((unsigned long long)r * mg.x) >> 32) >> mg.y) * n;
            if (rem != r % n) bad++;
        }
    }
    return bad;
}

/* Sturm counts of fpt_css_eig.cuh (scaled determinant recurrence) for nx probes of one tridiagonal (d, e2 = e^2, unit norm) */
void emu_sturm_counts(const double *d, const double *e2, int n, const double *xs, int nx, int *counts) {
    for (int k = 0; k < nx; k++) counts[k] = fpt_sturm_count(d, e2, n, xs[k], 1e-30);
}

/* the same for the large-cohort shuffle's table (fpt_magic31, shift from clz): n = 2 .. nmax */
long long emu_magic31_check(const unsigned *rs, long long nr, unsigned nmax) {
    long long bad = 0;
    for (unsigned n = 2; n <= nmax; n++) {
        const uint2 lm = fpt_umma_rtab_entry((int)n);
        const int sh = 31 - __clz((int)(n - 1));
        for (long long k = 0; k < nr; k++) {
            const unsigned r = rs[k] & 0x7fffffffu;
            const unsigned rem = r - (__umulhi(r, lm.y) >> sh) * n;
            if (rem != r % n) bad++;
            if ((r > lm.x) != (r > 2147483647u - (2147483648u % n))) bad++;
        }
    }
    return bad;
}

unsigned long long emu_css_perm2(const double *Xall, int m, int asize, int bsize, long long wbase, long long nwin,
                                 const unsigned char *status, int treshold, int runs, uint64_t seed, const uint64_t *state_override,
                                 int chain, int qbits, int nthreads, int grid, double *out_score, double *out_p, int *out_hits,
                                 int *out_n) {
    size_t smem = fpt_css_perm2_smem_bytes(m, nthreads, chain);
    unsigned long long rechecks = 0, *pr = &rechecks;
    run_grid(grid, nthreads, smem, [=]() {
        fpt_css_perm2_kernel(Xall, m, asize, bsize, wbase, nwin, status, treshold, runs, seed, state_override, chain, qbits,
                             out_score, out_p, out_hits, out_n, pr);
    });
    return rechecks;
}

/* fpt_css_observed.cuh: the observed-score kernel (a warp per window) */
void emu_css_observed(const double *Xall, int m, int asize, int bsize, long long nwin, const unsigned char *status, int grid,
                      double *out_score) {
    run_grid(grid, FPT_OBS_WARPS * 32, fpt_css_observed_smem_bytes(m), [=]() {
        fpt_css_observed_kernel(Xall, m, asize, bsize, nwin, status, out_score);
    });
}

/* the pipelined shuffle next to fpt_generate_labels from the same stream positions, and the warp-wide css() of the shuffled
   labels: one warp, permutation p on lane p % 32; labels_fast / labels_ref: nperm x m, scores: nperm */
void emu_umma_shuffle_and_score(const double *X, int m, int asize, int bsize, uint64_t state, int nperm,
                                unsigned short *labels_fast, unsigned short *labels_ref, double *scores) {
    std::vector<uint2> rtab(m + 1);
    for (int n = 0; n <= m; n++) { rtab[n].x = n > 0 ? fpt_randint_limit((uint32_t)n) : 0u; rtab[n].y = n > 0 ? fpt_randint_magic((uint32_t)n) : 0u; }
    const uint2 *rt = rtab.data();
    std::vector<uint2> rtab_fast(m + 1);
    for (int n = 0; n <= m; n++) rtab_fast[n] = fpt_umma_rtab_entry(n);
    const uint2 *rtf = rtab_fast.data();
    run_grid(1, 32, 32 * 8, [=]() {
        const int lane = threadIdx.x;
        double *stage = (double *)emu::g_dyn_smem;
        for (int p0 = 0; p0 < nperm; p0 += 32) {
            const int p = p0 + lane;
            if (p < nperm) {
                const uint64_t st = fpt_lcg_skip(state, (uint64_t)p * (uint64_t)(m - 1));
                long long sa = -1, sb = -1;
                const double S = 1048576.0 / 64.0;
                fpt_umma_shuffle(labels_fast + (size_t)p * m, m, rtf, st, true, X, S, asize, &sa, &sb);
                fpt_generate_labels<unsigned short>(labels_ref + (size_t)p * m, m, rt, st);
                /* the adjacent-pair sums read off the swaps against a sweep over the finished labels */
                long long ra = 0, rb = 0;
                const unsigned short *lf = labels_fast + (size_t)p * m;
                for (int col = 1; col < m; col++) {
                    const long long qv = (long long)fpt_umma_q(X, lf[col], lf[col - 1], S);
                    if (col < asize) ra += qv; else if (col > asize) rb += qv;
                }
                if (ra != sa || rb != sb) labels_fast[(size_t)p * m] = 0xffff;          /* poison: the test's comparison fails */
            }
            __syncwarp();
            for (int q = p0; q < p0 + 32 && q < nperm; q++) {
                const double sc = fpt_warp_css_score(X, stage, labels_fast + (size_t)q * m, asize, bsize, lane);
                if (lane == 0) scores[q] = sc;
            }
        }
    });
}

/* surrogate distance of fpt_css_observed.cuh for n point pairs, both orders */
void emu_umma_dist(const double *xi, const double *yi, const double *xj, const double *yj, int n, double *d_ij, double *d_ji) {
    run_grid(1, 32, 0, [=]() {
        for (int e = threadIdx.x; e < n; e += 32) {
            d_ij[e] = fpt_umma_dist(xi[e], yi[e], xj[e], yj[e]);
            d_ji[e] = fpt_umma_dist(xj[e], yj[e], xi[e], yi[e]);
        }
    });
}

}  // extern "C"
