/* oracle/cpu_update.c -- TEST / BASELINE INFRASTRUCTURE, not product code (nothing under eqvio_b200/ may use it).
 *
 * CPU restatement, without an interpreter in the timed region, of the dense linear algebra of one EqF propagate + correct
 * step of pvangoor/eqvio -- the part that holds > 99 % of the reference's flops at N >= 64:
 *
 *   cpu_dense_propagate / cpu_dense_correct      the REFERENCE's evaluation order
 *       Sigma <- (I + dt A) Sigma (I + dt A)^T + dt (B Q B^T + P)                     src/mathematical/VIO_eqf.cpp:62-72
 *       S = C Sigma C^T + R,  S^-1 by LU inverse,  K = Sigma C^T S^-1 (an Eigen expression: evaluated for Gamma = K ytilde
 *       and AGAIN for Sigma <- Sigma - K C Sigma),  dense dim x dim A and m x dim C          VIO_eqf.cpp:105-135
 *   cpu_structured_propagate / cpu_structured_correct   the minimal-flop form of the same update (block-sparse A and C, one
 *       Cholesky of S, Sigma -= Y^T Y) -- what separates the algorithmic part of the GPU speed-up from the hardware part.
 *
 * The O(N) glue (Lie-group actions, per-landmark Jacobian blocks, bookkeeping) is NOT restated here: the callers
 * (bench.py cpu_baseline / --impl reference, tests/test_cpu_update.py) take A, B, C, ytilde from the numpy oracle outside
 * the timed region and check every result of this file against the oracle's.  Materialising the dense A and C (what the
 * reference's stateMatrixA / outputMatrixC do) is charged here as a copy inside the timed region.
 *
 * GEMM / LU / Cholesky come from the OpenBLAS that numpy bundles (ILP64 build, scipy_*64_ symbols), opened with dlopen:
 * the reference's Eigen products are single-threaded (its build never enables OpenMP) -- cpu_set_threads(1) is the
 * faithful setting, all threads the generous one.  All matrices are row-major (numpy C order).
 *
 * Build: gcc -O3 -march=native -shared -fPIC -o oracle/_build/libcpu_update.so oracle/cpu_update.c -ldl   (oracle/build.py)
 */
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef long long bint; /* ILP64 */
enum { RowMajor = 101, NoTrans = 111, Trans = 112, Lower = 122, NonUnit = 131, Left = 141 };

static void (*p_dgemm)(int, int, int, bint, bint, bint, double, const double*, bint, const double*, bint, double, double*, bint);
static void (*p_dsyrk)(int, int, int, bint, bint, double, const double*, bint, double, double*, bint);
static void (*p_dtrsm)(int, int, int, int, int, bint, bint, double, const double*, bint, double*, bint);
static void (*p_dgetrf)(bint*, bint*, double*, bint*, bint*, bint*);
static void (*p_dgetri)(bint*, double*, bint*, bint*, double*, bint*, bint*);
static void (*p_dpotrf)(char*, bint*, double*, bint*, bint*);
static void (*p_set_threads)(int);
static int (*p_get_threads)(void);

static double now(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

int cpu_init(const char* blas_path) {
    void* h = dlopen(blas_path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) return -1;
    p_dgemm = dlsym(h, "scipy_cblas_dgemm64_");
    p_dsyrk = dlsym(h, "scipy_cblas_dsyrk64_");
    p_dtrsm = dlsym(h, "scipy_cblas_dtrsm64_");
    p_dgetrf = dlsym(h, "scipy_dgetrf_64_");
    p_dgetri = dlsym(h, "scipy_dgetri_64_");
    p_dpotrf = dlsym(h, "scipy_dpotrf_64_");
    p_set_threads = dlsym(h, "scipy_openblas_set_num_threads64_");
    p_get_threads = dlsym(h, "scipy_openblas_get_num_threads64_");
    return (p_dgemm && p_dsyrk && p_dtrsm && p_dgetrf && p_dgetri && p_dpotrf && p_set_threads && p_get_threads) ? 0 : -2;
}
void cpu_set_threads(int n) { p_set_threads(n); }
int cpu_get_threads(void) { return p_get_threads(); }

static void gemm(int ta, int tb, bint M, bint N, bint K, double alpha, const double* A, bint lda, const double* B, bint ldb, double beta,
                 double* C, bint ldc) {
    p_dgemm(RowMajor, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
}

/* Sigma <- (I + dt A) Sigma (I + dt A)^T + dt (B diag(q) B^T + diag(p))   -- VIO_eqf.cpp:62-72.  Returns seconds. */
double cpu_dense_propagate(int dim, double* Sigma, const double* A, const double* B, const double* q, const double* p, double dt) {
    const size_t n2 = (size_t)dim * dim;
    double* F = malloc(n2 * sizeof(double));
    double* T = malloc(n2 * sizeof(double));
    double* BQ = malloc((size_t)dim * 12 * sizeof(double));
    const double t0 = now();
    for (size_t i = 0; i < n2; ++i) F[i] = dt * A[i]; /* the dense state matrix is materialised (stateMatrixA), then I + dt A */
    for (int i = 0; i < dim; ++i) F[(size_t)i * dim + i] += 1.0;
    gemm(NoTrans, NoTrans, dim, dim, dim, 1.0, F, dim, Sigma, dim, 0.0, T, dim);
    gemm(NoTrans, Trans, dim, dim, dim, 1.0, T, dim, F, dim, 0.0, Sigma, dim);
    for (int i = 0; i < dim; ++i)
        for (int k = 0; k < 12; ++k) BQ[(size_t)i * 12 + k] = B[(size_t)i * 12 + k] * q[k];
    gemm(NoTrans, Trans, dim, dim, 12, dt, BQ, 12, B, 12, 1.0, Sigma, dim);
    for (int i = 0; i < dim; ++i) Sigma[(size_t)i * dim + i] += dt * p[i];
    const double t1 = now();
    free(F);
    free(T);
    free(BQ);
    return t1 - t0;
}

/* K = Sigma C^T (C Sigma C^T + r2 I)^-1, dim x m, the way the Eigen expression evaluates it */
static int dense_gain(int dim, int m, const double* Sigma, const double* C, double r2, double* K, double* CS, double* S, double* SCt) {
    gemm(NoTrans, NoTrans, m, dim, dim, 1.0, C, dim, Sigma, dim, 0.0, CS, dim);
    gemm(NoTrans, Trans, m, m, dim, 1.0, CS, dim, C, dim, 0.0, S, m);
    for (int i = 0; i < m; ++i) S[(size_t)i * m + i] += r2;
    bint n = m, info = 0, lwork = (bint)m * 64;
    bint* ipiv = malloc((size_t)m * sizeof(bint));
    double* work = malloc((size_t)lwork * sizeof(double));
    p_dgetrf(&n, &n, S, &n, ipiv, &info); /* S is symmetric: its row-major image is its column-major image */
    if (info == 0) p_dgetri(&n, S, &n, ipiv, work, &lwork, &info);
    free(ipiv);
    free(work);
    if (info != 0) return (int)info;
    gemm(NoTrans, Trans, dim, m, dim, 1.0, Sigma, dim, C, dim, 0.0, SCt, m);
    gemm(NoTrans, NoTrans, dim, m, m, 1.0, SCt, m, S, m, 0.0, K, m);
    return 0;
}

/* Gamma = K ytilde;  Sigma <- Sigma - (K C) Sigma with K evaluated a second time   -- VIO_eqf.cpp:116-131.
 * Cin is copied inside the timed region (the reference materialises the dense m x dim output matrix).  Returns seconds (< 0: LU failed). */
double cpu_dense_correct(int dim, int m, double* Sigma, const double* Cin, double r2, const double* y, double* Gamma) {
    const size_t n2 = (size_t)dim * dim, md = (size_t)m * dim;
    double* C = malloc(md * sizeof(double));
    double* K = malloc(md * sizeof(double));
    double* CS = malloc(md * sizeof(double));
    double* SCt = malloc(md * sizeof(double));
    double* S = malloc((size_t)m * m * sizeof(double));
    double* KC = malloc(n2 * sizeof(double));
    double* T = malloc(n2 * sizeof(double));
    const double t0 = now();
    memcpy(C, Cin, md * sizeof(double));
    int rc = dense_gain(dim, m, Sigma, C, r2, K, CS, S, SCt);
    if (rc == 0) {
        for (int i = 0; i < dim; ++i) {
            double s = 0.0;
            for (int k = 0; k < m; ++k) s += K[(size_t)i * m + k] * y[k];
            Gamma[i] = s;
        }
        rc = dense_gain(dim, m, Sigma, C, r2, K, CS, S, SCt);
    }
    if (rc == 0) {
        gemm(NoTrans, NoTrans, dim, dim, m, 1.0, K, m, C, dim, 0.0, KC, dim);
        memcpy(T, Sigma, n2 * sizeof(double));
        gemm(NoTrans, NoTrans, dim, dim, dim, -1.0, KC, dim, T, dim, 1.0, Sigma, dim);
    }
    const double t1 = now();
    free(C);
    free(K);
    free(CS);
    free(SCt);
    free(S);
    free(KC);
    free(T);
    return rc == 0 ? t1 - t0 : -1.0;
}

/* Block structure of A: As = A[:, 0:21] (dim x 21, dense sensor columns), D = the 3x3 diagonal blocks of the
 * landmark part (N x 3 x 3).  F = I + dt A;  Sigma <- F Sigma F^T + dt (B Q B^T + P) in O(21 dim^2). */
double cpu_structured_propagate(int dim, int N, double* Sigma, const double* As, const double* D, const double* B, const double* q,
                                const double* p, double dt) {
    const size_t n2 = (size_t)dim * dim;
    double* FS = malloc(n2 * sizeof(double));
    double* BQ = malloc((size_t)dim * 12 * sizeof(double));
    const double t0 = now();
    /* FS = F Sigma = Sigma + dt As Sigma[0:21, :] + dt blockdiag(0, D) Sigma */
    memcpy(FS, Sigma, n2 * sizeof(double));
    gemm(NoTrans, NoTrans, dim, dim, 21, dt, As, 21, Sigma, dim, 1.0, FS, dim);
    for (int i = 0; i < N; ++i) {
        const double* Di = D + 9 * (size_t)i;
        const double* s0 = Sigma + (size_t)(21 + 3 * i) * dim;
        double* f0 = FS + (size_t)(21 + 3 * i) * dim;
        for (int c = 0; c < dim; ++c) {
            const double x0 = s0[c], x1 = s0[dim + c], x2 = s0[2 * (size_t)dim + c];
            f0[c] += dt * (Di[0] * x0 + Di[1] * x1 + Di[2] * x2);
            f0[dim + c] += dt * (Di[3] * x0 + Di[4] * x1 + Di[5] * x2);
            f0[2 * (size_t)dim + c] += dt * (Di[6] * x0 + Di[7] * x1 + Di[8] * x2);
        }
    }
    /* Sigma' = FS F^T = FS + dt FS[:, 0:21] As^T + dt FS blockdiag(0, D)^T */
    memcpy(Sigma, FS, n2 * sizeof(double));
    gemm(NoTrans, Trans, dim, dim, 21, dt, FS, dim, As, 21, 1.0, Sigma, dim);
    for (int r = 0; r < dim; ++r) {
        const double* fr = FS + (size_t)r * dim + 21;
        double* sr = Sigma + (size_t)r * dim + 21;
        for (int i = 0; i < N; ++i) {
            const double* Di = D + 9 * (size_t)i;
            const double x0 = fr[3 * i], x1 = fr[3 * i + 1], x2 = fr[3 * i + 2];
            sr[3 * i] += dt * (Di[0] * x0 + Di[1] * x1 + Di[2] * x2);
            sr[3 * i + 1] += dt * (Di[3] * x0 + Di[4] * x1 + Di[5] * x2);
            sr[3 * i + 2] += dt * (Di[6] * x0 + Di[7] * x1 + Di[8] * x2);
        }
    }
    for (int i = 0; i < dim; ++i)
        for (int k = 0; k < 12; ++k) BQ[(size_t)i * 12 + k] = B[(size_t)i * 12 + k] * q[k];
    gemm(NoTrans, Trans, dim, dim, 12, dt, BQ, 12, B, 12, 1.0, Sigma, dim);
    for (int i = 0; i < dim; ++i) Sigma[(size_t)i * dim + i] += dt * p[i];
    const double t1 = now();
    free(FS);
    free(BQ);
    return t1 - t0;
}

/* Cb = the 2x3 output blocks (n x 2 x 3), lm[j] = state landmark of measured landmark j:
 * W = C Sigma (gather), S = W[:, L] C^T + r2 I = L L^T, Y = L^-1 W, z = L^-1 ytilde, Gamma = Y^T z, Sigma -= Y^T Y. */
double cpu_structured_correct(int dim, int n, double* Sigma, const double* Cb, const int* lm, double r2, const double* y, double* Gamma) {
    const int m = 2 * n;
    double* W = malloc((size_t)m * dim * sizeof(double));
    double* S = malloc((size_t)m * m * sizeof(double));
    double* z = malloc((size_t)m * sizeof(double));
    const double t0 = now();
    for (int j = 0; j < n; ++j) {
        const double* s0 = Sigma + (size_t)(21 + 3 * lm[j]) * dim;
        const double* c = Cb + 6 * (size_t)j;
        double* w0 = W + (size_t)(2 * j) * dim;
        for (int col = 0; col < dim; ++col) {
            const double x0 = s0[col], x1 = s0[dim + col], x2 = s0[2 * (size_t)dim + col];
            w0[col] = c[0] * x0 + c[1] * x1 + c[2] * x2;
            w0[dim + col] = c[3] * x0 + c[4] * x1 + c[5] * x2;
        }
    }
    for (int r = 0; r < m; ++r) {
        const double* wr = W + (size_t)r * dim + 21;
        for (int j = 0; j < n; ++j) {
            const double* c = Cb + 6 * (size_t)j;
            const double x0 = wr[3 * lm[j]], x1 = wr[3 * lm[j] + 1], x2 = wr[3 * lm[j] + 2];
            S[(size_t)r * m + 2 * j] = c[0] * x0 + c[1] * x1 + c[2] * x2;
            S[(size_t)r * m + 2 * j + 1] = c[3] * x0 + c[4] * x1 + c[5] * x2;
        }
        S[(size_t)r * m + r] += r2;
    }
    /* S symmetric: LAPACK's column-major "U" factor of the row-major image is the row-major lower factor L (S = L L^T) */
    bint nn = m, info = 0;
    char up = 'U';
    p_dpotrf(&up, &nn, S, &nn, &info);
    if (info == 0) {
        p_dtrsm(RowMajor, Left, Lower, NoTrans, NonUnit, m, dim, 1.0, S, m, W, dim); /* W <- Y */
        memcpy(z, y, (size_t)m * sizeof(double));
        p_dtrsm(RowMajor, Left, Lower, NoTrans, NonUnit, m, 1, 1.0, S, m, z, 1);
        for (int i = 0; i < dim; ++i) Gamma[i] = 0.0;
        for (int k = 0; k < m; ++k) {
            const double zk = z[k];
            const double* yk = W + (size_t)k * dim;
            for (int i = 0; i < dim; ++i) Gamma[i] += yk[i] * zk;
        }
        p_dsyrk(RowMajor, Lower, Trans, dim, m, -1.0, W, dim, 1.0, Sigma, dim); /* lower triangle of Sigma - Y^T Y */
        for (int r = 0; r < dim; ++r)
            for (int c = r + 1; c < dim; ++c) Sigma[(size_t)r * dim + c] = Sigma[(size_t)c * dim + r];
    }
    const double t1 = now();
    free(W);
    free(S);
    free(z);
    return info == 0 ? t1 - t0 : -1.0;
}
