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
 *   - gsl_eigen_symmv: cyclic two-sided Jacobi in fp64, eigenvectors in the COLUMNS of evec,
 *     input matrix is destroyed (as in GSL).
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

/* cyclic Jacobi; A (n x n, symmetric) is overwritten, eval gets the diagonal, evec the rotations */
static inline int gsl_eigen_symmv(gsl_matrix *A, gsl_vector *eval, gsl_matrix *evec, gsl_eigen_symmv_workspace *w) {
    (void)w;
    size_t n = A->size1, lda = A->tda, ldv = evec->tda;
    double *a = A->data, *v = evec->data;
    for (size_t i = 0; i < n; i++)
        for (size_t j = 0; j < n; j++) v[i * ldv + j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 100; sweep++) {
        double off = 0.0, diag = 0.0;
        for (size_t i = 0; i < n; i++) {
            diag += a[i * lda + i] * a[i * lda + i];
            for (size_t j = i + 1; j < n; j++) off += a[i * lda + j] * a[i * lda + j];
        }
        if (off <= 1e-300 || off <= 1e-34 * diag) break;
        for (size_t p = 0; p + 1 < n; p++) {
            for (size_t q = p + 1; q < n; q++) {
                double apq = a[p * lda + q];
                if (apq == 0.0) continue;
                double app = a[p * lda + p], aqq = a[q * lda + q];
                double theta = (aqq - app) / (2.0 * apq);
                double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (size_t k = 0; k < n; k++) {           /* columns p,q */
                    double akp = a[k * lda + p], akq = a[k * lda + q];
                    a[k * lda + p] = c * akp - s * akq;
                    a[k * lda + q] = s * akp + c * akq;
                }
                for (size_t k = 0; k < n; k++) {           /* rows p,q */
                    double apk = a[p * lda + k], aqk = a[q * lda + k];
                    a[p * lda + k] = c * apk - s * aqk;
                    a[q * lda + k] = s * apk + c * aqk;
                }
                for (size_t k = 0; k < n; k++) {
                    double vkp = v[k * ldv + p], vkq = v[k * ldv + q];
                    v[k * ldv + p] = c * vkp - s * vkq;
                    v[k * ldv + q] = s * vkp + c * vkq;
                }
            }
        }
    }
    for (size_t i = 0; i < n; i++) eval->data[i * eval->stride] = a[i * lda + i];
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
