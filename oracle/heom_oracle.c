/*
 * heom_oracle.c - plain C (C99 + optional OpenMP) restatement of the reference's
 * DEOM RK4 path.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): it is the
 * checker for mid-size parity cases and the multi-threaded CPU baseline of
 * bench.py; nothing under pyqed_b200/ links or calls it.
 *
 * Parity status: pinned - tests/test_oracle_golden.py runs it against the
 * fixtures in tests/golden (outputs of the unmodified reference).
 *
 * What is restated (paths relative to /root/reference):
 *   pascal table, nmax          DEOMSolver.init_           pyqed/heom/deom.py:1048-1060
 *   ado id of a multi-index     gen_hash_value             pyqed/heom/deom.py:555-565
 *   right-hand side per ADO     generate_dot_element       pyqed/heom/deom.py:641-664
 *   H(t), Q_m(t)                generate_time              pyqed/heom/deom.py:676-687
 *   classical RK4               rk4                        pyqed/heom/deom.py:725-766
 *   trajectory of rho_sys       DEOMSolver.run             pyqed/heom/deom.py:1094-1113
 * Like the reference it multiplies dense N x N matrices for every term (no use
 * of sparsity); unlike the reference the loop over ADOs is an OpenMP loop.
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex cplx;

static long long binom(int a, int b) {
    if (b < 0 || b > a) return 0;
    long long r = 1;
    for (int i = 1; i <= b; ++i) r = r * (a - b + i) / i;
    return r;
}

/* id = sum_i C(s_i + i, i + 1), s_i = running sum of the key */
static long long ado_id(const int* key, int K) {
    long long id = 0;
    int run = 0;
    for (int i = 0; i < K; ++i) {
        run += key[i];
        id += binom(run + i, i + 1);
    }
    return id;
}

long long oracle_c_nmax(int K, int L) { return binom(L + K, L); }

/* keys[id][K] for every |n| <= L, by enumerating multi-indices and hashing them */
static void fill_keys(int* keys, int K, int L) {
    int* key = (int*)calloc((size_t)K, sizeof(int));
    for (;;) {
        memcpy(keys + (size_t)ado_id(key, K) * K, key, sizeof(int) * (size_t)K);
        int pos = K - 1, total = 0;
        for (int i = 0; i < K; ++i) total += key[i];
        /* odometer step over all tuples with sum <= L */
        while (pos >= 0) {
            if (total < L) {
                key[pos]++;
                break;
            }
            total -= key[pos];
            key[pos] = 0;
            pos--;
        }
        if (pos < 0) break;
    }
    free(key);
}

int oracle_c_keys(int K, int L, int* keys_out) {
    fill_keys(keys_out, K, L);
    return 0;
}

static void matmul(cplx* restrict c, const cplx* restrict a, const cplx* restrict b, int n) {
    /* i-l-j order: the inner loop runs over contiguous elements of b and c */
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) c[i * n + j] = 0;
        for (int l = 0; l < n; ++l) {
            const cplx ail = a[i * n + l];
            for (int j = 0; j < n; ++j) c[i * n + j] += ail * b[l * n + j];
        }
    }
}

/* k[n] = F(y)[n] for every ADO; H, Q are the operators at this stage time */
static void rhs_all(cplx* k, const cplx* y, const int* keys, const long long* minus,
                    const long long* plus, int N, int K, int L, long long nmax, const cplx* H,
                    const cplx* Q, const cplx* expn, const cplx* etal, const cplx* etar,
                    const cplx* etaa, const int64_t* mode) {
    const int NN = N * N;
#pragma omp parallel
    {
        cplx* t1 = (cplx*)malloc(sizeof(cplx) * (size_t)NN);
        cplx* t2 = (cplx*)malloc(sizeof(cplx) * (size_t)NN);
#pragma omp for schedule(dynamic, 16)
        for (long long n = 0; n < nmax; ++n) {
            const int* key = keys + n * K;
            const cplx* rho = y + n * NN;
            cplx* out = k + n * NN;
            cplx damp = 0;
            int tier = 0;
            for (int q = 0; q < K; ++q) {
                damp += (double)key[q] * expn[q];
                tier += key[q];
            }
            matmul(t1, H, rho, N);
            matmul(t2, rho, H, N);
            for (int e = 0; e < NN; ++e) out[e] = -damp * rho[e] - I * (t1[e] - t2[e]);
            for (int q = 0; q < K; ++q) {
                const cplx* Qm = Q + (size_t)mode[q] * NN;
                if (key[q] > 0) {
                    const cplx* src = y + minus[n * K + q] * NN;
                    const cplx c = I * sqrt((double)key[q]) / csqrt(etaa[q]);
                    matmul(t1, Qm, src, N);
                    matmul(t2, src, Qm, N);
                    for (int e = 0; e < NN; ++e) out[e] -= c * (etal[q] * t1[e] - etar[q] * t2[e]);
                }
                if (tier < L) {
                    const cplx* src = y + plus[n * K + q] * NN;
                    const cplx c = I * sqrt((double)key[q] + 1.0) * csqrt(etaa[q]);
                    matmul(t1, Qm, src, N);
                    matmul(t2, src, Qm, N);
                    for (int e = 0; e < NN; ++e) out[e] -= c * (t1[e] - t2[e]);
                }
            }
        }
        free(t1);
        free(t2);
    }
}

/*
 * Propagate nt RK4 steps.  All complex arrays are interleaved (re, im) doubles.
 *   H, mu [N][N]; Q, Qd [M][N][N]; bath arrays [K]; mode [K] (int64)
 *   ados  [nmax][N][N] in/out (reference id order)
 *   fs, fc [nt][3] pulse samples at i*dt, i*dt+dt/2, i*dt+dt, or NULL
 *   traj  [nt+1][N][N] out (rho_sys before the first step and after every step), or NULL
 * Returns 0, or -1 on allocation failure.
 */
int oracle_c_deom_rk4(int N, int K, int M, int L, const double* H_, const double* mu_, const double* Q_,
                      const double* Qd_, const double* expn_, const double* etal_, const double* etar_,
                      const double* etaa_, const int64_t* mode, double* ados_, double dt, long long nt,
                      const double* fs, const double* fc, double* traj_, int nthreads) {
    const int NN = N * N;
    const long long nmax = binom(L + K, L);
    const cplx *H0 = (const cplx*)H_, *mu = (const cplx*)mu_, *Q0 = (const cplx*)Q_, *Qd = (const cplx*)Qd_;
    const cplx *expn = (const cplx*)expn_, *etal = (const cplx*)etal_, *etar = (const cplx*)etar_,
               *etaa = (const cplx*)etaa_;
    cplx* y = (cplx*)ados_;
    cplx* traj = (cplx*)traj_;
#ifdef _OPENMP
    omp_set_num_threads(nthreads > 0 ? nthreads : omp_get_num_procs());
#else
    (void)nthreads;
#endif
    int* keys = (int*)malloc(sizeof(int) * (size_t)nmax * K);
    long long* minus = (long long*)malloc(sizeof(long long) * (size_t)nmax * K);
    long long* plus = (long long*)malloc(sizeof(long long) * (size_t)nmax * K);
    cplx* kbuf = (cplx*)malloc(sizeof(cplx) * (size_t)nmax * NN);
    cplx* acc = (cplx*)malloc(sizeof(cplx) * (size_t)nmax * NN);
    cplx* ys = (cplx*)malloc(sizeof(cplx) * (size_t)nmax * NN);
    cplx* Ht = (cplx*)malloc(sizeof(cplx) * (size_t)NN);
    cplx* Qt = (cplx*)malloc(sizeof(cplx) * (size_t)M * NN);
    if (!keys || !minus || !plus || !kbuf || !acc || !ys || !Ht || !Qt) return -1;
    fill_keys(keys, K, L);
    /* neighbour ids (hash_minus / hash_plus, deom.py:588-605) */
    for (long long n = 0; n < nmax; ++n) {
        int* key = keys + n * K;
        int tier = 0;
        for (int q = 0; q < K; ++q) tier += key[q];
        for (int q = 0; q < K; ++q) {
            minus[n * K + q] = plus[n * K + q] = -1;
            if (key[q] > 0) {
                key[q]--;
                minus[n * K + q] = ado_id(key, K);
                key[q]++;
            }
            if (tier < L) {
                key[q]++;
                plus[n * K + q] = ado_id(key, K);
                key[q]--;
            }
        }
    }
    const long long total = nmax * NN;
    if (traj) memcpy(traj, y, sizeof(cplx) * (size_t)NN);
    for (long long step = 0; step < nt; ++step) {
        const double w[4] = {1.0, 2.0, 2.0, 1.0};
        const double a[3] = {dt / 2, dt / 2, dt};
        const int tix[4] = {0, 1, 1, 2};
        const cplx* yin = y;
        for (int st = 0; st < 4; ++st) {
            const double f = fs ? fs[step * 3 + tix[st]] : 0.0, g = fc ? fc[step * 3 + tix[st]] : 0.0;
            for (int e = 0; e < NN; ++e) Ht[e] = H0[e] + mu[e] * f;
            for (int e = 0; e < M * NN; ++e) Qt[e] = Q0[e] + Qd[e] * g;
            rhs_all(kbuf, yin, keys, minus, plus, N, K, L, nmax, Ht, Qt, expn, etal, etar, etaa, mode);
#pragma omp parallel for
            for (long long e = 0; e < total; ++e) {
                acc[e] = (st == 0 ? 0 : acc[e]) + w[st] * kbuf[e];
                if (st < 3) ys[e] = y[e] + kbuf[e] * a[st];
            }
            yin = ys;
        }
#pragma omp parallel for
        for (long long e = 0; e < total; ++e) y[e] += acc[e] * dt / 6;
        if (traj) memcpy(traj + (step + 1) * NN, y, sizeof(cplx) * (size_t)NN);
    }
    free(keys);
    free(minus);
    free(plus);
    free(kbuf);
    free(acc);
    free(ys);
    free(Ht);
    free(Qt);
    return 0;
}

int oracle_c_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
