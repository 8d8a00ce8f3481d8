/*
 * fpt_fet.cuh — Fisher's-exact-test hot path on the device.
 *
 *   fpt_fet_count_kernel   genotype codes of two populations -> 2x2 allele-count table per SNP
 *                          (reference: fetcount, fisher/cFisher.c:208-238)
 *   fpt_fet_score_kernel   2x2 table -> -log10 of the reference's two-tailed P
 *                          (reference: fet/shift_table/create_table/fet_p/binomial,
 *                           fisher/cFisher.c:245-483, and the -1.0*log10 of cFisher.c:183)
 *   fpt_window_table_kernel  window index -> [left,right) SNP range + "is this window visited"
 *                          (reference: slide_right comparative.c:49-71; loops cFisher.c:81-99 and
 *                           threadfisher.c:191-238)
 *   fpt_fet_window_kernel  per window: percentile of the SNP scores and sigma of 100 bootstrap
 *                          percentiles (reference: percentile/calc_std/bootstrap_sample/std/mean,
 *                           fisher/cFisher.c:136-144, 492-518, 547-597)
 *
 * Paths are relative to /root/reference/statistics/.
 */
#ifndef FPT_FET_CUH
#define FPT_FET_CUH

#include "fpt_rt.cuh"

/* ============================================================================================
 * K1: counting. Codes: 1 = homozygous major (value 3), 2 = homozygous minor (value -3), 0 = anything
 * else (heterozygous 0, missing -10000). A tile of SNPs is staged through shared memory as one byte
 * per genotype so that the global reads are full-width and coalesced whatever the population size;
 * then one thread owns one SNP and writes its table as a single 16-byte store.
 */
FPT_D unsigned char fpt_code_of(double v) { return v == 3.0 ? 1 : (v == -3.0 ? 2 : 0); }
FPT_D unsigned char fpt_code_of(signed char v) { return v == 3 ? 1 : (v == -3 ? 2 : 0); }

template <typename T>
FPT_D void fpt_stage_codes(const T *__restrict__ src, long long n, unsigned char *dst) {
    for (long long e = threadIdx.x; e < n; e += blockDim.x) dst[e] = fpt_code_of(src[e]);
}

#ifndef FPT_EMU
/* 16-byte loads when the tile start is 16-byte aligned (it is for every tile when the base pointer
   is and tile*size is even); scalar head/tail otherwise */
template <>
FPT_D void fpt_stage_codes<double>(const double *__restrict__ src, long long n, unsigned char *dst) {
    if ((((uintptr_t)src) & 15) == 0) {
        const double2 *s2 = reinterpret_cast<const double2 *>(src);
        long long n2 = n >> 1;
#pragma unroll 4
        for (long long e = threadIdx.x; e < n2; e += blockDim.x) {
            double2 v = __ldg(s2 + e);
            unsigned short packed = (unsigned short)(fpt_code_of(v.x) | (fpt_code_of(v.y) << 8));
            *reinterpret_cast<unsigned short *>(dst + 2 * e) = packed;
        }
        if ((n & 1) && threadIdx.x == 0) dst[n - 1] = fpt_code_of(src[n - 1]);
    } else {
#pragma unroll 4
        for (long long e = threadIdx.x; e < n; e += blockDim.x) dst[e] = fpt_code_of(__ldg(src + e));
    }
}

template <>
FPT_D void fpt_stage_codes<signed char>(const signed char *__restrict__ src, long long n, unsigned char *dst) {
    if ((((uintptr_t)src) & 15) == 0) {
        const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
        long long n16 = n >> 4;
        for (long long e = threadIdx.x; e < n16; e += blockDim.x) {
            uint4 v = __ldg(s4 + e);
            unsigned w[4] = { v.x, v.y, v.z, v.w };
            unsigned o[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                unsigned r = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) r |= (unsigned)fpt_code_of((signed char)((w[q] >> (8 * b)) & 0xff)) << (8 * b);
                o[q] = r;
            }
            *reinterpret_cast<uint4 *>(dst + 16 * e) = make_uint4(o[0], o[1], o[2], o[3]);
        }
        for (long long e = (n16 << 4) + threadIdx.x; e < n; e += blockDim.x) dst[e] = fpt_code_of(src[e]);
    } else {
        for (long long e = threadIdx.x; e < n; e += blockDim.x) dst[e] = fpt_code_of(src[e]);
    }
}
#endif

template <typename T>
__global__ void __launch_bounds__(256)
fpt_fet_count_kernel(const T *__restrict__ avals, const T *__restrict__ bvals, long long nsnp, int asize, int bsize,
                     int tile, int4 *__restrict__ tables) {
    FPT_DYN_SMEM(smem);
    unsigned char *sa = smem;                                  /* tile * asize codes, 16-byte aligned */
    unsigned char *sb = smem + (((size_t)tile * asize + 15) & ~(size_t)15);
    for (long long t0 = (long long)blockIdx.x * tile; t0 < nsnp; t0 += (long long)gridDim.x * tile) {
        int nt = (int)min((long long)tile, nsnp - t0);
        fpt_stage_codes<T>(avals + t0 * asize, (long long)nt * asize, sa);
        fpt_stage_codes<T>(bvals + t0 * bsize, (long long)nt * bsize, sb);
        __syncthreads();
        for (int t = threadIdx.x; t < nt; t += blockDim.x) {
            int a3 = 0, am = 0, b3 = 0, bm = 0;
            const unsigned char *pa = sa + (size_t)t * asize, *pb = sb + (size_t)t * bsize;
            for (int i = 0; i < asize; i++) { unsigned c = pa[i]; a3 += (c == 1); am += (c == 2); }
            for (int i = 0; i < bsize; i++) { unsigned c = pb[i]; b3 += (c == 1); bm += (c == 2); }
            tables[t0 + t] = make_int4(a3, am, b3, bm);
        }
        __syncthreads();
    }
}

/* ============================================================================================
 * K2: the two-tailed P of cFisher.c:405-455 (Feldman-Klinger / Zar walk), per table.
 *
 * Exact mode (N <= 67 and no u64 overflow in the numerator) repeats the reference's arithmetic
 * operation for operation — u64 binomials (here: a shared-memory triangle of exact C(n,k)), one
 * integer product, two u64->f64 conversions, one divide, then ratio-times-term recurrences with
 * separately rounded divide/multiply/add — so that its rounding-dependent `P2 < P0` decisions come
 * out the same. Everything else runs the same walk in log space relative to the
 * observed table (log mode), from a shared-memory log-factorial table sized to the largest N.
 */
#define FPT_FET_EXACT_MAX_N 67
#define FPT_BINOM_ENTRIES 1190                         /* sum_{n<=67} (n/2 + 1) */
#define FPT_FET_LOG_SKIP 41.58883083359672            /* 60 ln 2 */
#define FPT_FET_TIE_GUARD 1e-10
#define FPT_FET_TINY 8.6736173798840355e-19           /* 2^-60 */

FPT_HD int fpt_binom_index(int n, int k) {
    int h = n >> 1;
    if (k > n - k) k = n - k;
    return ((n & 1) ? (h + 1) * (h + 1) : h * (h + 1)) + k;
}

struct FptTable { int a, b, c, d; };

/* cFisher.c:327-346: clockwise order a,b,d,c; rotate so that the first minimum leads */
FPT_D void fpt_rotate_min_first(FptTable &f) {
    int cw0 = f.a, cw1 = f.b, cw2 = f.d, cw3 = f.c;
    int at = 0, mn = cw0;
    if (cw1 < mn) { mn = cw1; at = 1; }
    if (cw2 < mn) { mn = cw2; at = 2; }
    if (cw3 < mn) { mn = cw3; at = 3; }
    if (at == 1)      { f.a = cw1; f.b = cw2; f.d = cw3; f.c = cw0; }
    else if (at == 2) { f.a = cw2; f.b = cw3; f.d = cw0; f.c = cw1; }
    else if (at == 3) { f.a = cw3; f.b = cw0; f.d = cw1; f.c = cw2; }
}

/* cFisher.c:357-390: most extreme table on the other side, margins kept */
FPT_D void fpt_opposite_extreme(FptTable &f) {
    int R1 = f.a + f.b, R2 = f.c + f.d, C1 = f.a + f.c, C2 = f.b + f.d;
    int m1 = min(min(R1, R2), min(C1, C2));
    if (R1 <= R2 && C1 <= C2)      { f.a = m1 - f.a; f.b = R1 - f.a; f.c = C1 - f.a; f.d = C2 - f.b; }
    else if (R1 <= R2 && C2 <= C1) { f.b = m1 - f.b; f.a = R1 - f.b; f.d = C2 - f.b; f.c = C1 - f.a; }
    else if (R1 >= R2 && C1 <= C2) { f.c = m1 - f.c; f.a = C1 - f.c; f.d = R2 - f.c; f.b = R1 - f.a; }
    else                           { f.d = m1 - f.d; f.b = C2 - f.d; f.c = R2 - f.d; f.a = R1 - f.b; }
}

/* cFisher.c:473-483 with the table of exact binomials; ok = false on u64 overflow of the numerator */
FPT_D double fpt_point_prob_exact(const FptTable &f, const unsigned long long *binom, bool &ok) {
    unsigned long long x = binom[fpt_binom_index(f.a + f.b, f.a)];
    unsigned long long y = binom[fpt_binom_index(f.c + f.d, f.c)];
    unsigned long long z = binom[fpt_binom_index(f.a + f.b + f.c + f.d, f.a + f.c)];
    if (__umul64hi(x, y) != 0ULL) { ok = false; return 0.0; }
    return __ddiv_rn(__ull2double_rn(x * y), __ull2double_rn(z));
}

FPT_D double fpt_ratio(int n1, int n2, int d1, int d2) {
    /* (1.0*n1*n2)/(d1*d2): both products are exact in fp64 */
    return __ddiv_rn(__dmul_rn((double)n1, (double)n2), __dmul_rn((double)d1, (double)d2));
}

/* returns true and sets P when the table lies in the exact domain */
FPT_D bool fpt_fet_exact(FptTable f, const unsigned long long *binom, double &Pout) {
    if (f.a + f.b + f.c + f.d > FPT_FET_EXACT_MAX_N) return false;
    const int R1 = f.a + f.b, R2 = f.c + f.d, C1 = f.a + f.c, C2 = f.b + f.d;
    bool ok = true;
    fpt_rotate_min_first(f);
    const double P0 = fpt_point_prob_exact(f, binom, ok);
    if (!ok) return false;
    double P = P0, P1 = P0;
    while (f.a > 0) {
        f.b++; f.c++;
        P1 = __dmul_rn(fpt_ratio(f.a, f.d, f.b, f.c), P1);
        P = __dadd_rn(P, P1);
        f.a--; f.d--;
    }
    if (R1 == R2 || C1 == C2) {
        P = __dmul_rn(2.0, P);
    } else {
        fpt_opposite_extreme(f);
        fpt_rotate_min_first(f);
        double P2 = fpt_point_prob_exact(f, binom, ok);
        if (!ok) return false;
        while (P2 < P0) {
            P = __dadd_rn(P, P2);
            if (f.b == 0 || f.c == 0) break;
            f.a++; f.d++;
            P2 = __dmul_rn(fpt_ratio(f.b, f.c, f.a, f.d), P2);
            f.b--; f.c--;
        }
    }
    if (P > 1.0) P = 1.0;
    Pout = P;
    return true;
}

FPT_D double fpt_log_point_prob(const FptTable &f, const double *lf) {
    return __dsub_rn(
        __dsub_rn(__dadd_rn(__dadd_rn(__dadd_rn(lf[f.a + f.b], lf[f.c + f.d]), lf[f.a + f.c]), lf[f.b + f.d]),
                  lf[f.a + f.b + f.c + f.d]),
        __dadd_rn(__dadd_rn(__dadd_rn(lf[f.a], lf[f.b]), lf[f.c]), lf[f.d]));
}

/* 1/q for q = a product of two cell counts (1 .. 2^40: inside single-precision range): single-precision seed and ONE Newton step,
   relative error ~2^-44. The log-mode walk is not an operation-for-operation replay of anything (the reference overflows where it
   runs): it is validated against exact rationals to 1e-9, its `<` decisions carry a 1e-10 guard, and a walk has at most
   a few hundred terms, so 2^-44 per term is three orders of magnitude inside both. The fp64 pipe is what bounds this kernel
   (58 DFMA per clock and SM): every fp64 instruction taken out of the step counts. */
FPT_D double fpt_fet_rcp(double q) {
#ifndef FPT_EMU
    float r0;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"((float)q));
    const double r = (double)r0;
#else
    const double r = (double)(1.0f / (float)q);
#endif
    return fma(r, fma(-q, r, 1.0), r);
}

FPT_D double fpt_fet_neglog10_logmode(FptTable f, const double *lf) {
    const int R1 = f.a + f.b, R2 = f.c + f.d, C1 = f.a + f.c, C2 = f.b + f.d;
    fpt_rotate_min_first(f);
    const double lp0 = fpt_log_point_prob(f, lf);
    double S = 1.0, u = 1.0;
    {   /* first tail, towards a = 0: u_k+1 = u_k (a-k)(d-k) / ((b+1+k)(c+1+k)). Numerator and denominator are quadratics in k: they
           advance by their first differences (exact small integers in fp64), two adds each instead of two adds and a multiply;
           the division is a multiplication by the Newton reciprocal; the cut-off is looked at every fourth term */
        double num = (double)f.a * (double)f.d, den = (double)(f.b + 1) * (double)(f.c + 1);
        double dn = (double)(f.a + f.d - 1), dd = (double)(f.b + f.c + 3);
        int left = f.a;
        while (left > 0) {
            u *= num * fpt_fet_rcp(den);
            S += u;
            num -= dn; dn -= 2.0; den += dd; dd += 2.0; left--;
            if ((left & 3) == 0 && u < S * FPT_FET_TINY) break;     /* nothing further along this tail can change S */
        }
        /* the walk ends at a = 0 either way (the skipped terms are below 2^-60 of S) */
        f.b += f.a; f.c += f.a; f.d -= f.a; f.a = 0;
    }
    if (R1 == R2 || C1 == C2) {
        S = __dmul_rn(2.0, S);
    } else {
        fpt_opposite_extreme(f);
        fpt_rotate_min_first(f);
        double lu = __dsub_rn(fpt_log_point_prob(f, lf), lp0);
        if (lu < -FPT_FET_LOG_SKIP) {
            /* jump to the first inward table that still matters: bisection on the concave log-pmf,
               bounded by the mode of cell a, floor((a+c+1)(a+b+1)/(n+2)) */
            int K = min(f.b, f.c);
            int n = f.a + f.b + f.c + f.d;
            long long mode = ((long long)(f.a + f.c + 1) * (long long)(f.a + f.b + 1)) / (long long)(n + 2);
            int hi = (int)(mode - f.a);
            hi = max(0, min(hi, K));
            int lo = 0;
            while (lo < hi) {
                int mid = (lo + hi) / 2;
                FptTable g = { f.a + mid, f.b - mid, f.c - mid, f.d + mid };
                double l = __dsub_rn(fpt_log_point_prob(g, lf), lp0);
                if (l < -FPT_FET_LOG_SKIP) lo = mid + 1; else hi = mid;
            }
            f.a += lo; f.b -= lo; f.c -= lo; f.d += lo;
            lu = __dsub_rn(fpt_log_point_prob(f, lf), lp0);
        }
        double u2 = exp(lu);
        /* second tail, inwards from the far extreme: u2_k+1 = u2_k (b-k)(c-k) / ((a+1+k)(d+1+k)), the same way */
        double num = (double)f.b * (double)f.c, den = (double)(f.a + 1) * (double)(f.d + 1);
        double dn = (double)(f.b + f.c - 1), dd = (double)(f.a + f.d + 3);
        int left = min(f.b, f.c);
        while (u2 < 1.0 - FPT_FET_TIE_GUARD) {
            S += u2;
            if (left == 0) break;
            u2 *= num * fpt_fet_rcp(den);
            num -= dn; dn -= 2.0; den += dd; dd += 2.0; left--;
        }
    }
    double lp = __dadd_rn(lp0, log(S));
    if (lp >= 0.0) return -0.0;                          /* clamp at P = 1; -1.0*log10(1) = -0.0 (Q4) */
    return -__dmul_rn(lp, 0.43429448190325182765);
}

/* tables: int4 {A major, A minor, B major, B minor}; lf_global: lgamma(k+1), k = 0..maxn;
   binom_global: FPT_BINOM_ENTRIES exact binomials. Dynamic shared memory: binomials then lf. */
__global__ void __launch_bounds__(256)
fpt_fet_score_kernel(const int4 *__restrict__ tables, long long n, const unsigned long long *__restrict__ binom_global,
                     const double *__restrict__ lf_global, int maxn, int lf_in_smem, int force_log,
                     double *__restrict__ scores) {
    FPT_DYN_SMEM(smem);
    unsigned long long *binom = reinterpret_cast<unsigned long long *>(smem);
    double *lf_s = reinterpret_cast<double *>(smem + FPT_BINOM_ENTRIES * sizeof(unsigned long long));
    for (int i = threadIdx.x; i < FPT_BINOM_ENTRIES; i += blockDim.x) binom[i] = binom_global[i];
    if (lf_in_smem)
        for (int i = threadIdx.x; i <= maxn; i += blockDim.x) lf_s[i] = lf_global[i];
    __syncthreads();
    const double *lf = lf_in_smem ? lf_s : lf_global;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int4 t = tables[i];
        FptTable f = { t.x, t.y, t.z, t.w };
        double P, s;
        if (!force_log && fpt_fet_exact(f, binom, P)) s = __dmul_rn(-1.0, log10(P));
        else s = fpt_fet_neglog10_logmode(f, lf);
        scores[i] = s;
    }
}

/* Tables at sequencing coverage (row sums up to 500 and beyond, BASELINE configs[3]): the walk takes 10 .. 400 terms per table, and at
   one table per thread a warp waits for its longest one (host model on the benchmark tables: the longest of 32 is 1.9x the mean).
   This kernel sorts before it walks. A CTA stages a tile of FPT_FET_TILE tables in shared memory, estimates each table's walk
   length from the normal approximation of its hypergeometric law — both tails end where the term drops below 2^-60 of the
   observed one, i.e. L = sigma (sqrt(z0^2 + 120 ln 2) - z0) terms away, capped by the cells — counting-sorts the tile by that
   estimate (64 buckets of 8 terms) and hands consecutive runs of 32 to its warps: the lanes of a warp then walk tables of nearly
   equal length (model: 1.2x), every warp gets one run from each eighth of the distribution. Scores go back through shared memory,
   so loads and stores stay coalesced; the order inside a bucket does not matter (each score lands at its table's index). */
#define FPT_FET_TILE 1024
#define FPT_FET_SORT_THREADS 256
#define FPT_FET_BUCKETS 64
FPT_HD size_t fpt_fet_sorted_smem_extra() {
    return (size_t)FPT_FET_TILE * 16 + (size_t)FPT_FET_TILE * 8 + (size_t)FPT_FET_TILE * 2 + (size_t)(FPT_FET_BUCKETS + 1) * 4 + 12;
}

FPT_D int fpt_fet_walk_bucket(int4 t) {
    const float a = (float)t.x, b = (float)t.y, c = (float)t.z, d = (float)t.w;
    const float R1 = a + b, R2 = c + d, C1 = a + c, C2 = b + d, N = R1 + R2;
    if (!(N > 1.0f)) return 0;
    const int a0i = min(min(t.x, t.y), min(t.z, t.w));
    const float mu = R1 * C1 / N;
    const float var = fmaxf(R1 * R2 / N * (C1 / N) * (C2 / (N - 1.0f)), 1e-6f);
    const float sg = sqrtf(var), z0 = fabsf(a - mu) / sg;
    const float L = sg * (sqrtf(z0 * z0 + 83.18f) - z0);
    const bool towards_low = (t.x == a0i) || (t.w == a0i);          /* the first tail shrinks cell a (or d) */
    const float lo = fmaxf(0.0f, R1 + C1 - N), hi = fminf(R1, C1);
    const float far_ = towards_low ? hi - a : a - lo;
    const bool sym = (R1 == R2) || (C1 == C2);
    const float est = fminf((float)a0i, L) + (sym ? 0.0f : fminf(far_, L));
    return min(FPT_FET_BUCKETS - 1, (int)(est * 0.125f));
}

__global__ void __launch_bounds__(FPT_FET_SORT_THREADS)
fpt_fet_score_sorted_kernel(const int4 *__restrict__ tables, long long n, const unsigned long long *__restrict__ binom_global,
                            const double *__restrict__ lf_global, int maxn, int lf_in_smem, int force_log,
                            double *__restrict__ scores) {
    FPT_DYN_SMEM(smem);
    const int T = FPT_FET_SORT_THREADS, tid = threadIdx.x;
    constexpr int PER = FPT_FET_TILE / FPT_FET_SORT_THREADS;
    size_t off = 0;
    int4 *tt = reinterpret_cast<int4 *>(smem + off); off += (size_t)FPT_FET_TILE * 16;
    double *ts = reinterpret_cast<double *>(smem + off); off += (size_t)FPT_FET_TILE * 8;
    unsigned long long *binom = reinterpret_cast<unsigned long long *>(smem + off); off += FPT_BINOM_ENTRIES * sizeof(unsigned long long);
    double *lf_s = reinterpret_cast<double *>(smem + off); off += lf_in_smem ? ((size_t)maxn + 1) * 8 : 0;
    off = (off + 3) & ~(size_t)3;
    int *hist = reinterpret_cast<int *>(smem + off); off += (size_t)(FPT_FET_BUCKETS + 1) * 4;
    unsigned short *perm = reinterpret_cast<unsigned short *>(smem + off);
    for (int i = tid; i < FPT_BINOM_ENTRIES; i += T) binom[i] = binom_global[i];
    if (lf_in_smem)
        for (int i = tid; i <= maxn; i += T) lf_s[i] = lf_global[i];
    const double *lf = lf_in_smem ? lf_s : lf_global;
    const long long ntiles = (n + FPT_FET_TILE - 1) / FPT_FET_TILE;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long base = tile * FPT_FET_TILE;
        const int cnt = (int)min((long long)FPT_FET_TILE, n - base);
        for (int b = tid; b <= FPT_FET_BUCKETS; b += T) hist[b] = 0;
        __syncthreads();                                           /* also: the previous tile's scores have been written out */
        int bucket[PER], rank[PER];
#pragma unroll
        for (int k = 0; k < PER; k++) {
            const int e = tid + k * T;
            bucket[k] = 0; rank[k] = 0;
            if (e < cnt) {
                const int4 t = tables[base + e];
                tt[e] = t;
                bucket[k] = fpt_fet_walk_bucket(t);
                rank[k] = atomicAdd(&hist[bucket[k]], 1);
            }
        }
        __syncthreads();
        if (tid < 32) {                                            /* exclusive scan of the 64 bucket counts */
            const int c0 = hist[2 * tid], c1 = hist[2 * tid + 1];
            int inc = c0 + c1;
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FPT_FULL_MASK, inc, o); if (tid >= o) inc += y; }
            const int ex = inc - (c0 + c1);
            __syncwarp();
            hist[2 * tid] = ex; hist[2 * tid + 1] = ex + c0;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PER; k++) {
            const int e = tid + k * T;
            if (e < cnt) perm[hist[bucket[k]] + rank[k]] = (unsigned short)e;
        }
        __syncthreads();
#pragma unroll 1
        for (int k = 0; k < PER; k++) {
            const int pos = tid + k * T;
            if (pos < cnt) {
                const int e = perm[pos];
                const int4 t = tt[e];
                FptTable f = { t.x, t.y, t.z, t.w };
                double P, sc;
                if (!force_log && fpt_fet_exact(f, binom, P)) sc = __dmul_rn(-1.0, log10(P));
                else sc = fpt_fet_neglog10_logmode(f, lf);
                ts[e] = sc;
            }
        }
        __syncthreads();
        for (int e = tid; e < cnt; e += T) scores[base + e] = ts[e];
    }
}

/* max over all tables of N = a+b+c+d (sizes the log-factorial table for direct-table input) */
__global__ void fpt_fet_maxn_kernel(const int4 *__restrict__ tables, long long n, int *__restrict__ out) {
    int m = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int4 t = tables[i];
        m = max(m, t.x + t.y + t.z + t.w);
    }
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(FPT_FULL_MASK, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

/* ============================================================================================
 * Window table. Window w covers positions [w*wstep, w*wstep + wsize], both ends inclusive (Q6);
 * left = first SNP with pos >= start, right = first SNP with pos > stop (comparative.c:58-65).
 * mode 0 = serial scan (cFisher.c:81), mode 1 = the pthreads scan (threadfisher.c:55-58,191-218):
 * tasks 0..num_tasks of 100 windows, num_tasks = (regend/wstep - 3)/100, nothing when that is 0,
 * a stop >= regend is moved to regend + wstep. Windows at or beyond index regend/wstep (the length
 * of the caller's output arrays, Q17) are never produced.
 */
FPT_HD bool fpt_window_scheduled(long long w, int regend, int wsize, int wstep, int threaded) {
    long long start = w * (long long)wstep;
    if (w < 0 || start + wsize > (long long)regend + wstep) return false;
    if (w >= (long long)(regend / wstep)) return false;
    if (!threaded) return true;
    int num_tasks = (regend / wstep - 3) / 100;
    if (num_tasks <= 0) return false;
    if (w / 100 <= num_tasks) return true;
    long long stop = ((long long)num_tasks + 1) * 100 * (long long)wstep + (wsize - wstep);
    return stop >= regend;
}

__global__ void fpt_window_table_kernel(const int *__restrict__ pos, long long nsnp, long long wbase, long long nwin,
                                        int regend, int wsize, int wstep, int threaded, int *__restrict__ wleft,
                                        int *__restrict__ wright, int *__restrict__ max_npos) {
    /* local window w is global window wbase + w (a rank of a sharded scan owns a contiguous range) */
    long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int np = 0;
    if (w < nwin) {
        int l = 0, r = 0;
        if (fpt_window_scheduled(wbase + w, regend, wsize, wstep, threaded)) {
            long long start = (wbase + w) * (long long)wstep, stop = start + wsize;
            long long lo = 0, hi = nsnp;
            while (lo < hi) { long long mid = (lo + hi) >> 1; if ((long long)pos[mid] < start) lo = mid + 1; else hi = mid; }
            l = (int)lo;
            hi = nsnp;
            while (lo < hi) { long long mid = (lo + hi) >> 1; if ((long long)pos[mid] <= stop) lo = mid + 1; else hi = mid; }
            r = (int)lo;
        }
        wleft[w] = l; wright[w] = r;
        np = r - l;
    }
    for (int o = 16; o > 0; o >>= 1) np = max(np, __shfl_xor_sync(FPT_FULL_MASK, np, o));
    if ((threadIdx.x & 31) == 0 && np > 0) atomicMax(max_npos, np);
}

/* ============================================================================================
 * K3: per-window percentile and bootstrap sigma. One CTA per window.
 *
 * The reference sorts the window's scores, takes (1-d)*x[idx] + d*x[idx+1] with idx = (int)((n-1)q),
 * then 100 times draws n indices with replacement FROM THE SORTED ARRAY, sorts the resample and takes
 * the same percentile; sigma is the population standard deviation of those 100 values, summed from
 * the last index down. Because the resample is drawn from sorted data, its order statistics are the
 * sorted scores at the order statistics of the drawn indices — so each replicate is a counting pass
 * over its n draws, no second sort. Replicate s starts s*n draws into the window's LCG stream
 * (skip-ahead); the rare rejected draw (probability < n/2^31 each) shifts the later replicates, which
 * is detected from the per-replicate draw counts and repaired by re-running with corrected offsets.
 */
#define FPT_FET_NSAMPLES 100
#define FPT_FET_HIST_LD 128              /* u16 counters per value: one per replicate, padded to 2 x 64 (see the kernel) */
#define FPT_FET_HIST_MAX_NPOS 768        /* counting pass while 128 x npos u16 counters fit shared memory */

FPT_D double fpt_percentile_sorted(const double *x, int n, double q) {
    double h = __dmul_rn((double)(n - 1), q);
    int idx = (int)h;
    double delta = __dsub_rn(h, (double)idx);
    double lo = __dmul_rn(__dsub_rn(1.0, delta), x[idx]);
    if (idx + 1 >= n) return lo;                         /* the reference reads one past the data here (Q8) */
    return __dadd_rn(lo, __dmul_rn(delta, x[idx + 1]));
}

FPT_D void fpt_bitonic_sort(double *x, int npad) {
    for (int k = 2; k <= npad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < npad; i += blockDim.x) {
                int p = i ^ j;
                if (p > i) {
                    double a = x[i], b = x[p];
                    bool up = (i & k) == 0;
                    if ((a > b) == up) { x[i] = b; x[p] = a; }
                }
            }
            __syncthreads();
        }
    }
}

/* order statistics k and k+1 (0-based, k+1 optional) of replicate draws, by bisection on the index
   value with the draws regenerated on every pass: O(1) memory, used for very large windows */
FPT_D void fpt_select_by_regeneration(uint64_t st0, int n, int k, bool need_next, int &v0, int &v1, int &used) {
    const uint32_t lim = fpt_randint_limit((uint32_t)n), mag = fpt_randint_magic((uint32_t)n);
    int lo = 0, hi = n - 1;
    while (lo < hi) {                                    /* smallest v with #(draw <= v) >= k+1 */
        int mid = (lo + hi) >> 1, cnt = 0;
        uint64_t st = st0; used = 0;
        for (int i = 0; i < n; i++) cnt += ((int)fpt_randint_fast((uint32_t)n, lim, mag, st, used) <= mid);
        if (cnt >= k + 1) hi = mid; else lo = mid + 1;
    }
    v0 = lo; v1 = lo;
    uint64_t st = st0; used = 0;
    int cnt = 0, nxt = n;
    for (int i = 0; i < n; i++) {
        int r = (int)fpt_randint_fast((uint32_t)n, lim, mag, st, used);
        cnt += (r <= lo);
        if (r > lo && r < nxt) nxt = r;
    }
    if (need_next && cnt < k + 2) v1 = nxt;
}

__global__ void __launch_bounds__(128)
fpt_fet_window_kernel(const double *__restrict__ snp_scores, const int *__restrict__ wleft,
                      const int *__restrict__ wright, long long wbase, long long nwin, double perc, uint64_t seed,
                      const uint64_t *__restrict__ state_override, int npad_max, int use_hist,
                      double *__restrict__ out_score, double *__restrict__ out_std,
                      unsigned char *__restrict__ out_flag) {
    FPT_DYN_SMEM(smem);
    __shared__ double reps[FPT_FET_NSAMPLES];
    __shared__ int offs[FPT_FET_NSAMPLES + 1];
    __shared__ int cons[FPT_FET_NSAMPLES];
    __shared__ int changed;
    double *sorted = reinterpret_cast<double *>(smem);
    unsigned short *hist = reinterpret_cast<unsigned short *>(smem + (size_t)npad_max * sizeof(double));

    for (long long w = blockIdx.x; w < nwin; w += gridDim.x) {
        const int l = wleft[w], n = wright[w] - l;
        if (n <= 0) continue;                            /* uniform across the CTA */
        int npad = 2;
        while (npad < n) npad <<= 1;
        for (int i = threadIdx.x; i < npad; i += blockDim.x)
            sorted[i] = i < n ? snp_scores[l + i] : __longlong_as_double(0x7ff0000000000000LL);
        __syncthreads();
        fpt_bitonic_sort(sorted, npad);
        const uint64_t st_win = state_override ? state_override[w] : fpt_stream_state(seed, wbase + w, FPT_STREAM_RESAMPLE);
        const int k = (int)__dmul_rn((double)(n - 1), perc);
        const double h = __dmul_rn((double)(n - 1), perc);
        const double delta = __dsub_rn(h, (double)k);
        const bool need_next = k + 1 < n;
        const uint32_t lim = fpt_randint_limit((uint32_t)n), mag = fpt_randint_magic((uint32_t)n);
        const int s = threadIdx.x;
        if (s < FPT_FET_NSAMPLES) offs[s] = s * n;
        __syncthreads();
        for (int pass = 0;; pass++) {
            int used = 0;
            if (s < FPT_FET_NSAMPLES) {
                uint64_t st = fpt_lcg_skip(st_win, (uint64_t)offs[s]);
                int v0 = 0, v1 = 0;
                if (use_hist) {
                    /* counter of (replicate s, value v) = hcol[v * FPT_FET_HIST_LD]: replicates s and s + 64 share a 32-bit word,
                       so the 32 lanes of a warp always touch 32 different banks whatever they drew (no replay) */
                    unsigned short *hcol = hist + (((s & 63) << 1) | (s >> 6));
                    for (int i = 0; i < n; i++) hcol[(size_t)i * FPT_FET_HIST_LD] = 0;
                    for (int i = 0; i < n; i++) hcol[(size_t)fpt_randint_fast((uint32_t)n, lim, mag, st, used) * FPT_FET_HIST_LD]++;
                    int cum = 0, v = 0;
                    for (; v < n; v++) { cum += hcol[(size_t)v * FPT_FET_HIST_LD]; if (cum > k) break; }
                    v0 = v; v1 = v;
                    if (need_next && cum <= k + 1) { for (v++; v < n; v++) if (hcol[(size_t)v * FPT_FET_HIST_LD]) break; v1 = v; }
                } else {
                    fpt_select_by_regeneration(st, n, k, need_next, v0, v1, used);
                }
                double val = __dmul_rn(__dsub_rn(1.0, delta), sorted[v0]);
                if (need_next) val = __dadd_rn(val, __dmul_rn(delta, sorted[v1]));
                reps[s] = val;
                cons[s] = used;
            }
            /* common case: no replicate had a rejected draw, so the assumed offsets s*n were right */
            if (pass == 0 && !__syncthreads_or(s < FPT_FET_NSAMPLES && used != n)) break;
            if (threadIdx.x == 0) changed = 0;
            __syncthreads();
            if (threadIdx.x == 0) {                      /* replicate s starts after the draws of 0..s-1 */
                int run = 0;
                for (int t = 0; t < FPT_FET_NSAMPLES; t++) {
                    if (offs[t] != run) { offs[t] = run; changed = 1; }
                    run += cons[t];
                }
            }
            __syncthreads();
            const int again = changed;                   /* read by everyone BEFORE thread 0 may reset it in the next pass */
            __syncthreads();
            if (!again) break;
        }
        if (threadIdx.x == 0) {
            double mu = 0.0;
            for (int i = FPT_FET_NSAMPLES; i--;) mu = __dadd_rn(mu, reps[i]);
            mu = __ddiv_rn(mu, (double)FPT_FET_NSAMPLES);
            double var = 0.0;
            for (int i = FPT_FET_NSAMPLES; i--;) {
                double e = __dsub_rn(reps[i], mu);
                var = __dadd_rn(var, __dmul_rn(e, e));
            }
            var = __ddiv_rn(var, (double)FPT_FET_NSAMPLES);
            out_score[w] = fpt_percentile_sorted(sorted, n, perc);
            out_std[w] = __dsqrt_rn(var);
            out_flag[w] = 1;
        }
        __syncthreads();
    }
}

#endif
