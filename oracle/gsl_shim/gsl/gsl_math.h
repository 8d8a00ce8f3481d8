/*
 * Minimal stand-in for the dozen GNU GSL 1.16 symbols that the reference's css.c touches
 * (/root/reference/statistics/css/css.c:421-430 dgemm, :533-555 symmetric eigensolver).
 *
 * TEST INFRASTRUCTURE ONLY: it lets oracle/Makefile compile the *unmodified* reference css.c
 * into oracle/_ref/libref_css.so where real GSL is not installed. Original code, not GSL's.
 *
 * Numerical contract:
 *   - gsl_blas_dgemm: row-major, NoTrans/NoTrans only, i-k-j accumulation; when beta == 0 the
 *     destination is overwritten without being read (the reference hands in uninitialised
 *     malloc memory as C).
 *   - gsl_eigen_symmv: Householder tridiagonalisation + implicit QL in fp64 (the algorithm family GSL itself
 *     uses), eigenvectors in the COLUMNS of evec, input matrix is destroyed (as in GSL).
 *   - gsl_eigen_symmv_sort: descending by eigenvalue, permuting the columns along.
 */
#ifndef FPT_GSL_SHIM_H
#define FPT_GSL_SHIM_H

#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef struct { size_t size; size_t stride; double *data; } gsl_vector;
typedef struct { size_t size1; size_t size2; size_t tda; double *data; } gsl_matrix;
typedef struct { gsl_matrix matrix; } gsl_matrix_view;
typedef struct { gsl_vector vector; } gsl_vector_view;
typedef struct { size_t size; } gsl_eigen_symmv_workspace;

enum { CblasNoTrans = 111 };
enum { GSL_EIGEN_SORT_VAL_ASC = 0, GSL_EIGEN_SORT_VAL_DESC = 1 };

static inline gsl_matrix_view gsl_matrix_view_array(double *base, size_t n1, size_t n2) {
    gsl_matrix_view v;
    v.matrix.size1 = n1; v.matrix.size2 = n2; v.matrix.tda = n2; v.matrix.data = base;
    return v;
}

static inline gsl_vector *gsl_vector_alloc(size_t n) {
    gsl_vector *v = (gsl_vector *)malloc(sizeof *v);
    v->size = n; v->stride = 1; v->data = (double *)calloc(n ? n : 1, sizeof(double));
    return v;
}
static inline void gsl_vector_free(gsl_vector *v) { if (v) { free(v->data); free(v); } }
static inline double gsl_vector_get(const gsl_vector *v, size_t i) { return v->data[i * v->stride]; }

static inline gsl_matrix *gsl_matrix_alloc(size_t n1, size_t n2) {
    gsl_matrix *m = (gsl_matrix *)malloc(sizeof *m);
    m->size1 = n1; m->size2 = n2; m->tda = n2;
    m->data = (double *)calloc(n1 * n2 ? n1 * n2 : 1, sizeof(double));
    return m;
}
static inline void gsl_matrix_free(gsl_matrix *m) { if (m) { free(m->data); free(m); } }

static inline gsl_vector_view gsl_matrix_column(gsl_matrix *m, size_t j) {
    gsl_vector_view v;
    v.vector.size = m->size1; v.vector.stride = m->tda; v.vector.data = m->data + j;
    return v;
}

static inline int gsl_blas_dgemm(int ta, int tb, double alpha, const gsl_matrix *A, const gsl_matrix *B,
                                 double beta, gsl_matrix *C) {
    (void)ta; (void)tb;
    size_t M = A->size1, K = A->size2, N = B->size2;
    for (size_t i = 0; i < M; i++) {
        double *c = C->data + i * C->tda;
        if (beta == 0.0) { for (size_t j = 0; j < N; j++) c[j] = 0.0; }
        else if (beta != 1.0) { for (size_t j = 0; j < N; j++) c[j] *= beta; }
        for (size_t k = 0; k < K; k++) {
            double a = alpha * A->data[i * A->tda + k];
            const double *b = B->data + k * B->tda;
            for (size_t j = 0; j < N; j++) c[j] += a * b[j];
        }
    }
    return 0;
}

static inline gsl_eigen_symmv_workspace *gsl_eigen_symmv_alloc(size_t n) {
    gsl_eigen_symmv_workspace *w = (gsl_eigen_symmv_workspace *)malloc(sizeof *w);
    w->size = n;
    return w;
}
static inline void gsl_eigen_symmv_free(gsl_eigen_symmv_workspace *w) { free(w); }

/* Symmetric eigensolver of the same family as GSL's gsl_eigen_symmv: Householder reduction to tridiagonal form
   followed by implicit-shift QL iterations with the transformations accumulated (the classic tred2 / tql2 pair, written
   here from the textbook algorithm). A is destroyed, eigenvalues go to eval (unsorted), eigenvectors to the COLUMNS of
   evec. O(n^3) with a small constant, so the CPU baseline is not handicapped by the stand-in. */
static inline int gsl_eigen_symmv(gsl_matrix *A, gsl_vector *eval, gsl_matrix *evec, gsl_eigen_symmv_workspace *w) {
    (void)w;
    const int n = (int)A->size1;
    const size_t lda = A->tda, ldv = evec->tda;
    double *V = evec->data;
    double *d = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    double *e = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
#define VV(i, j) V[(size_t)(i) * ldv + (size_t)(j)]
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) VV(i, j) = A->data[(size_t)i * lda + j];
    /* ---- Householder tridiagonalisation */
    for (int j = 0; j < n; j++) d[j] = VV(n - 1, j);
    for (int i = n - 1; i > 0; i--) {
        double scale = 0.0, h = 0.0;
        for (int k = 0; k < i; k++) scale += fabs(d[k]);
        if (scale == 0.0) {
            e[i] = d[i - 1];
            for (int j = 0; j < i; j++) { d[j] = VV(i - 1, j); VV(i, j) = 0.0; VV(j, i) = 0.0; }
        } else {
            for (int k = 0; k < i; k++) { d[k] /= scale; h += d[k] * d[k]; }
            double f = d[i - 1], g = sqrt(h);
            if (f > 0) g = -g;
            e[i] = scale * g; h -= f * g; d[i - 1] = f - g;
            for (int j = 0; j < i; j++) e[j] = 0.0;
            for (int j = 0; j < i; j++) {
                f = d[j]; VV(j, i) = f; g = e[j] + VV(j, j) * f;
                for (int k = j + 1; k <= i - 1; k++) { g += VV(k, j) * d[k]; e[k] += VV(k, j) * f; }
                e[j] = g;
            }
            f = 0.0;
            for (int j = 0; j < i; j++) { e[j] /= h; f += e[j] * d[j]; }
            const double hh = f / (h + h);
            for (int j = 0; j < i; j++) e[j] -= hh * d[j];
            for (int j = 0; j < i; j++) {
                f = d[j]; g = e[j];
                for (int k = j; k <= i - 1; k++) VV(k, j) -= (f * e[k] + g * d[k]);
                d[j] = VV(i - 1, j); VV(i, j) = 0.0;
            }
        }
        d[i] = h;
    }
    for (int i = 0; i < n - 1; i++) {
        VV(n - 1, i) = VV(i, i); VV(i, i) = 1.0;
        const double h = d[i + 1];
        if (h != 0.0) {
            for (int k = 0; k <= i; k++) d[k] = VV(k, i + 1) / h;
            for (int j = 0; j <= i; j++) {
                double g = 0.0;
                for (int k = 0; k <= i; k++) g += VV(k, i + 1) * VV(k, j);
                for (int k = 0; k <= i; k++) VV(k, j) -= g * d[k];
            }
        }
        for (int k = 0; k <= i; k++) VV(k, i + 1) = 0.0;
    }
    for (int j = 0; j < n; j++) { d[j] = VV(n - 1, j); VV(n - 1, j) = 0.0; }
    if (n > 0) { VV(n - 1, n - 1) = 1.0; e[0] = 0.0; }
    /* ---- implicit QL on the tridiagonal, rotations applied to V */
    for (int i = 1; i < n; i++) e[i - 1] = e[i];
    if (n > 0) e[n - 1] = 0.0;
    double f = 0.0, tst1 = 0.0;
    const double eps = 2.220446049250313e-16;
    for (int l = 0; l < n; l++) {
        tst1 = fmax(tst1, fabs(d[l]) + fabs(e[l]));
        int m = l;
        while (m < n) { if (fabs(e[m]) <= eps * tst1) break; m++; }
        if (m > l) {
            int iter = 0;
            do {
                iter++;
                double g = d[l], p = (d[l + 1] - g) / (2.0 * e[l]), r = hypot(p, 1.0);
                if (p < 0) r = -r;
                d[l] = e[l] / (p + r); d[l + 1] = e[l] * (p + r);
                const double dl1 = d[l + 1];
                double h = g - d[l];
                for (int i = l + 2; i < n; i++) d[i] -= h;
                f += h;
                p = d[m];
                double c = 1.0, c2 = c, c3 = c, s = 0.0, s2 = 0.0;
                const double el1 = e[l + 1];
                for (int i = m - 1; i >= l; i--) {
                    c3 = c2; c2 = c; s2 = s;
                    g = c * e[i]; h = c * p; r = hypot(p, e[i]);
                    e[i + 1] = s * r; s = e[i] / r; c = p / r;
                    p = c * d[i] - s * g; d[i + 1] = h + s * (c * g + s * d[i]);
                    for (int k = 0; k < n; k++) {
                        h = VV(k, i + 1);
                        VV(k, i + 1) = s * VV(k, i) + c * h;
                        VV(k, i) = c * VV(k, i) - s * h;
                    }
                }
                p = -s * s2 * c3 * el1 * e[l] / dl1;
                e[l] = s * p; d[l] = c * p;
            } while (fabs(e[l]) > eps * tst1 && iter < 200);
        }
        d[l] = d[l] + f;
        e[l] = 0.0;
    }
#undef VV
    for (int i = 0; i < n; i++) eval->data[(size_t)i * eval->stride] = d[i];
    free(d); free(e);
    return 0;
}

static inline int gsl_eigen_symmv_sort(gsl_vector *eval, gsl_matrix *evec, int order) {
    size_t n = eval->size, ldv = evec->tda;
    for (size_t i = 0; i + 1 < n; i++) {
        size_t best = i;
        for (size_t j = i + 1; j < n; j++) {
            double ej = eval->data[j * eval->stride], eb = eval->data[best * eval->stride];
            if (order == GSL_EIGEN_SORT_VAL_DESC ? (ej > eb) : (ej < eb)) best = j;
        }
        if (best != i) {
            double t = eval->data[i * eval->stride];
            eval->data[i * eval->stride] = eval->data[best * eval->stride];
            eval->data[best * eval->stride] = t;
            for (size_t k = 0; k < evec->size1; k++) {
                double u = evec->data[k * ldv + i];
                evec->data[k * ldv + i] = evec->data[k * ldv + best];
                evec->data[k * ldv + best] = u;
            }
        }
    }
    return 0;
}

#endif
