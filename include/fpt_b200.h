/*
 * fpt_b200.h — C ABI of libfpt_b200.so, the B200 (sm_100a) implementation of the reference's two
 * genome-wide divergence scans: the per-SNP Fisher's exact test with its per-window percentile and
 * bootstrap sigma (FET) and the per-window Cluster Separation Score with MDS and permutation test (CSS).
 *
 * Plain C: pointers, sizes and scalars only. Reference paths below are relative to
 * /root/reference/statistics/.
 *
 * Three layers, all in one library:
 *   1. drop-in entry points with the reference's exact argument lists (host pointers);
 *   2. extended host entry points (compact inputs, window ranges for sharding, explicit RNG control,
 *      parity probes);
 *   3. a device-resident API (device pointers + cudaStream_t) used by pipelines that keep data in HBM.
 *
 * Every function returns FPT_OK (0) or a negative FPT_ERR_* code; fpt_last_error() gives the text.
 * There is no CPU fallback: without a CUDA device every compute entry point fails with
 * FPT_ERR_NO_DEVICE.
 */
#ifndef FPT_B200_H
#define FPT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FPT_OK 0
#define FPT_ERR_CUDA (-1)              /* a CUDA runtime call or kernel failed */
#define FPT_ERR_ARG (-2)               /* invalid argument */
#define FPT_ERR_POSITIONS (-3)         /* populations A and B do not share the same SNP positions (the
                                          reference silently mis-pairs SNPs in that case) */
#define FPT_ERR_WINDOW_TOO_LARGE (-4)  /* a window holds more SNPs than one CTA can sort in shared memory */
#define FPT_ERR_NO_DEVICE (-5)         /* no usable CUDA device */

/* scan semantics: which windows a scan visits */
#define FPT_SCAN_SERIAL 0              /* `compute`:       fisher/cFisher.c:81, css/css.c:117 */
#define FPT_SCAN_THREADED 1            /* `threadcompute`: fisher/threadfisher.c:55-58,191-218 (tasks of 100
                                          windows; nothing at all when regend/wstep < 103) */

/* per-window status reported by the CSS probes */
#define FPT_WIN_EMPTY 0
#define FPT_WIN_DISCARDED 1
#define FPT_WIN_SCORED 2

/* ------------------------------------------------------------------------------------------------
 * library state
 */
const char *fpt_last_error(void);
int fpt_device_count(void);                 /* number of CUDA devices, 0 if none / no driver */
int fpt_set_device(int device);             /* device used by the host entry points of this process */
/* The settings below are process-wide. Each is atomic, and every entry point takes one snapshot of them when it starts: a
   setter racing a scan in another thread affects that thread's NEXT call, never the kernel route of a call in flight.
   Host entry points that share a device's staging buffers take turns (one per device at a time). */
void fpt_set_seed(uint64_t seed);           /* seed of the window-keyed random streams (default 20261018) */
uint64_t fpt_get_seed(void);
/* CSS label shuffles. 0 (default): permutation k of a window is the reference's Fisher-Yates shuffle (css/css.c:700-706)
   of fresh identity labels, its nrand48 stream positioned k*(m-1) draws into the window's stream (counter-based).
   1: the reference's exact chain — one persistent label array shuffled again and again (css/css.c:727-752), bit-identical
   to significance_treshold() run on identity labels from the same 48-bit state. */
void fpt_set_perm_mode(int chain);
int fpt_get_perm_mode(void);
/* Cohorts too large for the all-in-shared-memory permutation kernel (m > 250), independent shuffles: 1 (default) scores a
   batch of 128 permutations as one u8 contraction on tcgen05 / tensor memory (csrc/fpt_css_perm_umma.cuh, m <= 1024),
   0 keeps the general kernel (csrc/fpt_css_perm_large.cuh), 2 is the tensor-memory kernel with a 10-bit surrogate, which sends
   a large share of the permutations through the exact re-scoring. Same decisions in every mode; the switch exists for the
   parity tests. */
void fpt_set_perm_large_kernel(int tensor_memory);
/* Cohorts of 8..64 individuals (the genome scans), independent shuffles: 1 (default) fpt_css_perm3_kernel (two permutations in
   flight per thread, conflict-free label layout, no fp64 distance matrix; csrc/fpt_css_perm3.cuh), 0 the round-1 kernel
   fpt_css_perm2_kernel. Same hits, permutations drawn, p and scores from both; the switch exists for the parity tests. */
void fpt_set_perm_small_kernel(int v);
/* Cohorts of 3..48 individuals, classical MDS (mds 0 and the start of mds 2): 1 (default) the Householder tridiagonalisation with the
   matrix in the registers of one warp (csrc/fpt_css_eig_reg.cuh), 0 the shared-memory kernel that serves every cohort up to the
   one-warp limit (csrc/fpt_css_eig.cuh). Results agree to rounding; the switch exists for the parity tests. */
void fpt_set_mds_small_kernel(int v);
/* Large-cohort MDS on count codes: threads per CTA — 256 (default: up to 128 registers per thread, the pipelined loads of the
   Lanczos product stay in registers), 384 (80 registers) or 512 (64). Tuning aid; results are identical up to the summation order inside a warp, which does not depend on it. */
void fpt_set_lanczos_threads(int threads);
/* diagnostic: SM cycles per phase of the tensor-memory permutation kernel since the last call, summed over CTAs and windows
   (0 distance pass, 1 hand-over of the observed score, 2 shuffles, 3 membership rows, 4 contraction, 5 decisions, 6 label copy-out
   and adjacent-pair sums); synchronises the device */
int fpt_debug_umma_phases(unsigned long long *out8);
/* the same for the large-cohort MDS kernel (0 dissimilarity = compare_all + fill_averages, 1 row means + code conversion, 2 products, 3 Gram-Schmidt, 4 tridiagonal
   solves, 5 norms / next vector, 6 coordinates; slot 7 is not a cycle count: it is the number of Lanczos steps taken, summed over windows) */
int fpt_debug_lanczos_phases(unsigned long long *out8);
/* Large-cohort classical MDS (csrc/fpt_css_lanczos.cuh): the highest form of the matrix the Lanczos product may stream —
   3 (default) 8-bit count codes squared arithmetically with the blank entries kept as a list, 2 8-bit count codes through a
   shared-memory table of squares, 1 16-bit count codes, 0 the fp64 matrix B. Every window takes the highest form it qualifies
   for; lower settings exist for the parity tests. */
void fpt_set_lanczos_form(int max_form);
/* Large cohorts, the genotype-distance matrix D = [P|M][M|P]' (compare_all, css/css.c:277-327; csrc/fpt_css_k4.cuh): 2 (default) one
   u8 GEMM per window on tcgen05 with the accumulators in tensor memory, 1 bit-plane popcounts — both write count codes that the
   Lanczos kernel streams — 0 the round-1 route (fp64 matrix in global memory, converted to codes in place). Identical results in
   modes 1 and 2 (integers); mode 0 differs by rounding only. A lanczos_form below 2 implies mode 0. */
void fpt_set_k4_mode(int mode);
/* diagnostic: SM cycles per phase of the tcgen05 GEMM kernel since the last call (0 operand expansion, 1 waiting for MMAs,
   2 accumulator drain + code stores) */
int fpt_debug_k4_phases(unsigned long long *out4);
/* parity probe: opposite-homozygote counts of every pair of individuals in window `window` (global index inside r), as the chosen
   kernel (2 tcgen05 GEMM, 1 popcounts) writes them: counts[m * m] int32 on the host; -1 everywhere for an empty window */
struct fpt_genotypes;
struct fpt_scan_range;
int fpt_debug_k4_counts(const struct fpt_genotypes *g, const struct fpt_scan_range *r, int mode, int64_t window, int32_t *counts);
/* number of permutations since the last call whose integer surrogate score could not decide `permuted >= observed` and
   were re-scored in the reference's summation order (diagnostic; synchronises the device); -1 on error */
long long fpt_css_perm_rechecks(void);
void fpt_release(void);                     /* free cached device/pinned buffers */
/* 48-bit LCG state of (seed, global window index, stream); stream 0 = bootstrap / label shuffles
   (nrand48 draws), 1 = SMACOF starts (drand48 draws). Pure host arithmetic. */
uint64_t fpt_window_state(uint64_t seed, int64_t window, int stream);
/* Per-kernel device timing. While enabled, every kernel this library launches is bracketed by CUDA events on
   its stream; fpt_profile_summary() waits for them, writes a JSON object {"kernel": {"launches": n, "ms": t}, ...}
   into buf and clears the record. Not thread-safe; meant for benchmarks. */
int fpt_profile_enable(int on);
int fpt_profile_summary(char *buf, size_t buflen);

/* ------------------------------------------------------------------------------------------------
 * 1. drop-in entry points (host pointers, caller-owned, outputs pre-zeroed by the caller; only scored
 *    windows are written, at index window_start / wstep; nothing is written at or beyond index
 *    regend / wstep).
 *
 * fpt_fet_threadcompute == `threadcompute` of fisher/threadfisher.h:33-34 (fisher/threadfisher.c:47-100),
 *                          bound by fisher/fisher_cython_parallel.pyx:4-5,14-15
 * fpt_fet_compute       == `compute` of fisher/cFisher.h:11 (fisher/cFisher.c:38-115),
 *                          bound by fisher/fisher_cython.pyx:4-5,10-11
 * fpt_css_threadcompute == `threadcompute` of css/threadcss.h:36-37 (css/threadcss.c:52-109),
 *                          bound by css/css_cython_parallel.pyx:4-5,14-15
 * fpt_css_compute       == `compute` of css/css.h:10 (css/css.c:49-156), bound by css/css_cython.pyx
 *
 * vals: float64 genotype codes, position-major / individual-minor. DOMAIN: {3, -3, 0, -10000}, the codes tools/VCFConvert.py:8-17
 *       emits. On that domain the CSS pair test is the reference's `((int)a)*b == -9` (css/css.c:291) and the FET count its
 *       `== 3` / `== -3` (fisher/cFisher.c:212-216). Other doubles are classified as "neither homozygote" here, whereas css.c
 *       truncates only its first operand (3.9 paired with -3, or 1 with -9, would count there): outside the domain the two differ.
 * pos: int32 position of every value (each SNP position repeated once per individual).
 */
int fpt_fet_threadcompute(double *avals, double *bvals, int *apos, int *bpos, int regstart, int regend, int wsize,
                          int wstep, int alen, int blen, double perc, double *scores, double *stddev);
int fpt_fet_compute(double *avals, double *bvals, int *apos, int *bpos, int regstart, int regend, int wsize,
                    int wstep, int alen, int blen, double perc, double *scores, double *stddev);
int fpt_css_threadcompute(double *avals, double *bvals, int *apos, int *bpos, int regstart, int regend, int wsize,
                          int wstep, int alen, int blen, int treshold, int runs, int drosophila, int mds,
                          double *scores, double *p);
int fpt_css_compute(double *avals, double *bvals, int *apos, int *bpos, int regstart, int regend, int wsize,
                    int wstep, int alen, int blen, int treshold, int runs, int drosophila, int mds, double *scores,
                    double *p);

/* ------------------------------------------------------------------------------------------------
 * 2. extended host entry points
 */

/* Input description shared by the extended scans. Exactly one of {avals/bvals} (float64, reference
   layout) or {acodes/bcodes} (int8: 3, -3, anything else = other) must be set; pos holds ONE int32
   position per SNP (nsnp entries, sorted ascending). */
typedef struct fpt_genotypes {
    const double *avals, *bvals;
    const int8_t *acodes, *bcodes;
    const int32_t *pos;
    int64_t nsnp;
    int asize, bsize;
} fpt_genotypes;

/* Window range and RNG control. Windows window_begin <= w < window_end (global indices, w = start/wstep)
   are processed; outputs are indexed by w - window_begin and must hold window_end - window_begin entries.
   states_* (optional, same indexing) override the window-keyed streams with explicit 48-bit LCG states. */
typedef struct fpt_scan_range {
    int regend, wsize, wstep;
    int semantics;                      /* FPT_SCAN_SERIAL or FPT_SCAN_THREADED */
    int64_t window_begin, window_end;
    uint64_t seed;
    const uint64_t *states_resample;    /* bootstrap (FET) / permutation (CSS) streams */
    const uint64_t *states_init;        /* SMACOF random starts (CSS, mds = 1) */
} fpt_scan_range;

/* optional per-window probes of the CSS scan (any pointer may be NULL) */
typedef struct fpt_css_probes {
    uint8_t *status;                    /* FPT_WIN_* */
    double *X;                          /* final embedding, m x 2 per window */
    double *evals;                      /* three largest eigenvalues of the classical-MDS matrix */
    int32_t *hits, *nperm;              /* permutation test: hits and permutations actually drawn */
    int32_t *smacof_iters;              /* nruns entries per window (mds 1: 4, mds 2: 1) */
    double *smacof_sigma;
} fpt_css_probes;

int fpt_fet_scan(const fpt_genotypes *g, const fpt_scan_range *r, double perc, double *scores, double *stddev,
                 uint8_t *written);
int fpt_css_scan(const fpt_genotypes *g, const fpt_scan_range *r, int treshold, int runs, int drosophila, int mds,
                 double *scores, double *p, uint8_t *written, const fpt_css_probes *probes);

/* per-SNP stage on its own: tables[4*nsnp] = {A major, A minor, B major, B minor} (fetcount,
   fisher/cFisher.c:208-238) and/or neglog10p[nsnp] = -log10 P (fet, cFisher.c:405-455); either output
   may be NULL */
int fpt_fet_per_snp(const fpt_genotypes *g, int32_t *tables, double *neglog10p);
/* direct 2x2 tables (BASELINE config "genome-scale FET"): tables[4*n] -> neglog10p[n];
   force_log != 0 evaluates every table in log mode (cross-check of the two arithmetic modes) */
int fpt_fet_tables(const int32_t *tables, int64_t n, int force_log, double *neglog10p);

/* ------------------------------------------------------------------------------------------------
 * 3. device-resident API. All pointers are DEVICE pointers of the current device; `stream` is a
 *    cudaStream_t passed as void* (NULL = default stream). Calls only enqueue work unless noted.
 */
int fpt_dev_fet_count_f64(const double *avals, const double *bvals, int64_t nsnp, int asize, int bsize,
                          int32_t *tables, void *stream);
int fpt_dev_fet_count_i8(const int8_t *acodes, const int8_t *bcodes, int64_t nsnp, int asize, int bsize,
                         int32_t *tables, void *stream);
/* max_n = largest a+b+c+d among the tables (sizes the log-factorial table); <= 0: computed (synchronises) */
int fpt_dev_fet_score(const int32_t *tables, int64_t n, int max_n, int force_log, double *neglog10p, void *stream);
/* wleft/wright: nwin int32 each; max_npos: one int32 (device), must be zeroed by the caller */
int fpt_dev_window_table(const int32_t *pos, int64_t nsnp, const fpt_scan_range *r, int32_t *wleft, int32_t *wright,
                         int32_t *max_npos, void *stream);
/* max_npos: host value read back from fpt_dev_window_table's counter */
int fpt_dev_fet_windows(const double *snp_scores, const int32_t *wleft, const int32_t *wright,
                        const fpt_scan_range *r, int max_npos, double perc, const uint64_t *states, double *scores,
                        double *stddev, uint8_t *written, void *stream);
/* planes: 2 * m * ceil(nsnp/32) uint32 */
size_t fpt_dev_css_planes_bytes(int64_t nsnp, int m);
int fpt_dev_css_pack_f64(const double *avals, const double *bvals, int64_t nsnp, int asize, int bsize,
                         uint32_t *planes, void *stream);
int fpt_dev_css_pack_i8(const int8_t *acodes, const int8_t *bcodes, int64_t nsnp, int asize, int bsize,
                        uint32_t *planes, void *stream);
int fpt_dev_css_absdiff(const double *afreq, const double *bfreq, int64_t nsnp, double *absdiff, void *stream);
/* workspace for fpt_dev_css_windows (embeddings of every window and start, per-CTA scratch for large m). Cohorts beyond the
   one-warp MDS path (m >~ 165) also hold the genotype-distance count codes of one pass here: 2 m^2 bytes per window for at most
   4096 windows (8 GB at m = 1000); longer ranges run in several passes inside the call. */
size_t fpt_dev_css_workspace_bytes(int m, int64_t nwin, int mds);
/* planes (or absdiff when drosophila != 0) -> scores, p, status (nwin each). states_* and probes hold
   DEVICE pointers here. */
int fpt_dev_css_windows(const uint32_t *planes, const double *absdiff, int asize, int bsize, const int32_t *wleft,
                        const int32_t *wright, const fpt_scan_range *r, int treshold, int runs, int mds,
                        void *workspace, size_t workspace_bytes, double *scores, double *p, uint8_t *status,
                        const fpt_css_probes *probes, void *stream);

/* ================================================================================================
 * Text ingest (host only). VCF and GTrack "valued points" text -> positions plus one int8 code per
 * individual and SNP (3 / 0 / -3, -128 for the reference's -10000), i.e. the `fpt_genotypes` compact layout.
 * Replaces tools/VCFConvert.py:5-89 (GT table, "#CHROM" header, GT slot taken from the first record's FORMAT, one point
 * per record and selected individual) and the reading loop of statistics/fisher/testFisher.c:193-227.
 * Chromosomes come back as runs of consecutive records that share a CHROM / seqid string.
 */
typedef struct fpt_chrom_run {
    int64_t name_off;       /* byte offset of the name inside the text buffer */
    int32_t name_len;
    int32_t reserved;
    int64_t first_record;   /* index of the run's first record */
} fpt_chrom_run;

int fpt_vcf_scan(const char *buf, size_t len, int64_t *header_off, int64_t *body_off, int64_t *nrecords);
/* codes: nrecords * nsamples int8, record-major, sample order = sample_cols (0-based VCF column numbers).
   A genotype outside the reference's table (a KeyError there) gives FPT_ERR_ARG naming record and column. */
int fpt_vcf_parse(const char *buf, size_t len, int64_t body_off, int chrom_col, int pos_col, int format_col,
                  const int32_t *sample_cols, int nsamples, int64_t nrecords, int8_t *codes, int32_t *pos,
                  fpt_chrom_run *runs, int64_t max_runs, int64_t *nruns);
int fpt_gtrack_scan(const char *buf, size_t len, int64_t *nrecords);
int fpt_gtrack_parse(const char *buf, size_t len, int seqid_col, int start_col, int value_col, int64_t nrecords,
                     int32_t *pos, double *vals, fpt_chrom_run *runs, int64_t max_runs, int64_t *nruns);
/* reference-layout float64 values -> compact codes */
int fpt_compact_codes(const double *vals, int64_t n, int8_t *codes);

#ifdef __cplusplus
}
#endif
#endif
