// heom_stage_rows.cuh - kernel 1, the plain-load row kernel (stage_rows_kernel).
// Included by heom_kernels.cu only (one translation unit); split out for readability.
#pragma once
#include "heom_core.cuh"
#include "heom_device.cuh"

// build-time tuning knobs (see pyqed_b200/build.py)
#ifndef HEOM_U
#define HEOM_U 4          // links fetched per batch in the diagonal-Q path
#endif
#ifndef HEOM_MINBLOCKS
#define HEOM_MINBLOCKS 2  // __launch_bounds__ min blocks per SM for the row kernel
#endif

// ---------------------------------------------------------------------------
// stage kernels
// ---------------------------------------------------------------------------
// Kernel 1 (N <= 8): a warp owns 32/N consecutive ADOs; lane (sub,row) owns one
// matrix row in registers.  -i[H,rho] uses H from the constant bank (kernel
// parameter) or, when H depends on time/trajectory, from shared memory.
//
// Neighbour terms, two variants:
//  QDIAG (every Q_m diagonal: projectors, sigma_z, occupation numbers): the
//    coupling is element-wise, (Q rho' - rho' Q)_ij = (q_i - q_j) rho'_ij, so a
//    link only needs the rows r of rho' with q_r != 0 (plus, for non-Hermitian
//    ADOs, the matching column entries).  The N lanes of an ADO fetch such a row
//    with one coalesced 16N-byte request and accumulate into the shared k tile;
//    U links are fetched per batch to keep U independent loads in flight per lane.
//  general Q: per-lane sparse row/column products in registers.
template <int N, bool TDEP, bool QDIAG>
__global__ void __launch_bounds__(256, HEOM_MINBLOCKS) stage_rows_kernel(const StageArgs a,
                                                         const __grid_constant__ HParam<N> hp) {
    constexpr int NN = N * N, APW = 32 / N, LD = (N % 2 == 0) ? N + 1 : N, TILE = APW * N * LD;
    constexpr int U = HEOM_U;
    extern __shared__ double2 smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int b = blockIdx.y;
    const double2* __restrict__ ops = a.ops + (long long)b * a.ops_bstride;
    double2* Hs = smem;
    double2* coef_s = Hs + (TDEP ? NN : 0);
    double2* qd_s = coef_s + (QDIAG ? 2 * a.ncoef : 0);
    double2* tiles = qd_s + (QDIAG ? a.nmod * N : 0);
    double2* rho_s = tiles + wid * 2 * TILE;
    double2* k_s = rho_s + TILE;
    unsigned char* supp_s = (unsigned char*)(tiles + nwarps * 2 * TILE);
    if (TDEP) {
        for (int e = threadIdx.x; e < NN; e += blockDim.x) Hs[e] = ops[e];
    }
    if (QDIAG) {
        for (int e = threadIdx.x; e < 2 * a.ncoef; e += blockDim.x) coef_s[e] = a.coef[e];
        for (int e = threadIdx.x; e < a.nmod * N; e += blockDim.x) {
            const int m = e / N, j = e - m * N;
            qd_s[e] = ops[(1 + m) * NN + j * N + j];
        }
        for (int e = threadIdx.x; e < a.nmod * (2 * N + 1); e += blockDim.x) supp_s[e] = a.supp[e];
    }
    if (TDEP || QDIAG) __syncthreads();
#define HEL(r_, c_) (TDEP ? Hs[(r_) * N + (c_)] : hp.v[(r_) * N + (c_)])
    const long long boff = (long long)b * a.nmax * NN;
    const double2* __restrict__ yin = a.yin + boff;
    const int sub = lane / N, row = lane - sub * N;
    const bool lane_ok = lane < APW * N;
    const unsigned submask = lane_ok ? (((1u << N) - 1u) << (sub * N)) : 0u;
    const unsigned char* insupp_s = supp_s + a.nmod * (N + 1);
    const long long step = a.traj ? (*a.step_base + a.local_step) : 0;

    for (long long g = (long long)blockIdx.x * nwarps + wid; g < a.ngroups;
         g += (long long)gridDim.x * nwarps) {
        const long long gm = (a.scramble && g < (a.ngroups & ~15ll))
                                 ? ((g & ~15ll) | ((g + (((unsigned)(g >> 4) * 2654435761u) >> 28)) & 15ll))
                                 : g;
        const long long base = a.slot_lo + gm * APW;
        const int cnt = (int)min((long long)APW, a.slot_hi - base);
        const int nelem = cnt * NN;
        const double2* src = yin + base * NN;
        for (int e = lane; e < nelem; e += 32) {
            const int s = e / NN, r = e - s * NN, i = r / N, j = r - i * N;
            rho_s[(s * N + i) * LD + j] = ldg2(src + e);
        }
        __syncwarp();
        const bool on = lane_ok && sub < cnt;
        if (on) {  // column pass: (H rho)[:, row]
            double2 col[N];
#pragma unroll
            for (int l = 0; l < N; ++l) col[l] = rho_s[(sub * N + l) * LD + row];
#pragma unroll
            for (int rr = 0; rr < N; ++rr) {
                double2 c = make_double2(0.0, 0.0);
#pragma unroll
                for (int l = 0; l < N; ++l) cfma(c, HEL(rr, l), col[l]);
                k_s[(sub * N + rr) * LD + row] = c;
            }
        }
        __syncwarp();
        if (on) {
            const long long slot = base + sub;
            double2 r[N];
            {
                double2 rv[N];
#pragma unroll
                for (int l = 0; l < N; ++l) rv[l] = rho_s[(sub * N + row) * LD + l];
                const double2 d = a.damp[slot];
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    double2 t = k_s[(sub * N + row) * LD + j];
#pragma unroll
                    for (int l = 0; l < N; ++l) cfms(t, rv[l], HEL(l, j));
                    // -i t - damp * rho
                    r[j] = make_double2(t.y - (d.x * rv[j].x - d.y * rv[j].y),
                                        -t.x - (d.x * rv[j].y + d.y * rv[j].x));
                }
            }
            const int lbeg = a.link_ptr[slot], lend = a.link_ptr[slot + 1];
            if (QDIAG) {
#pragma unroll
                for (int j = 0; j < N; ++j) k_s[(sub * N + row) * LD + j] = r[j];
                __syncwarp(submask);
                for (int lp = lbeg; lp < lend; lp += U) {
                    int2 lk[U];
                    double2 A[U];
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        lk[u] = (lp + u < lend) ? __ldg(a.links + lp + u) : make_int2((int)slot, 0);
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int m = heom::meta_mode(lk[u].y);
                        const int r0 = supp_s[m * (N + 1) + 1];
                        A[u] = ldg2(yin + (long long)lk[u].x * NN + r0 * N + row);
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int ci = heom::meta_ci(lk[u].y, a.nind, a.lmax), m = heom::meta_mode(lk[u].y);
                        const double2 aL = coef_s[2 * ci], aR = coef_s[2 * ci + 1];
                        const double2* __restrict__ pn = yin + (long long)lk[u].x * NN;
                        const int ns = supp_s[m * (N + 1)];
                        const double2 qj = qd_s[m * N + row];
                        const bool outside = insupp_s[m * N + row] == 0;
                        for (int t = 0; t < ns; ++t) {
                            const int rr = supp_s[m * (N + 1) + 1 + t];
                            const double2 Aj = (t == 0) ? A[u] : ldg2(pn + rr * N + row);
                            const double2 qr = qd_s[m * N + rr];
                            double2 c = cmul(aL, qr);
                            cfma(c, aR, qj);
                            double2* d1 = &k_s[(sub * N + rr) * LD + row];
                            double2 v1 = *d1;
                            cfma(v1, c, Aj);
                            *d1 = v1;
                            if (outside) {  // element (row, rr): only the right product survives
                                const double2 Bj = a.herm ? make_double2(Aj.x, -Aj.y)
                                                          : ldg2(pn + row * N + rr);
                                double2* d2 = &k_s[(sub * N + row) * LD + rr];
                                double2 v2 = *d2;
                                cfma(v2, cmul(aR, qr), Bj);
                                *d2 = v2;
                            }
                        }
                        __syncwarp(submask);
                    }
                }
            } else {
                for (int lp = lbeg; lp < lend; ++lp) {
                    const int2 lk = a.links[lp];
                    const double2* __restrict__ pn = yin + (long long)lk.x * NN;
                    const int ci = heom::meta_ci(lk.y, a.nind, a.lmax), m1 = 1 + heom::meta_mode(lk.y);
                    const double2 aL = a.coef[2 * ci], aR = a.coef[2 * ci + 1];
                    const double2* __restrict__ Qm = ops + m1 * NN;
                    const short* rp = a.row_ptr + m1 * (N + 1);
                    const short* ri = a.row_idx + m1 * NN;
                    for (int t = rp[row]; t < rp[row + 1]; ++t) {
                        const int l = ri[t];
                        const double2 q = cmul(aL, Qm[row * N + l]);
#pragma unroll
                        for (int j = 0; j < N; ++j) cfma(r[j], q, ldg2(pn + l * N + j));
                    }
                    const short* cp = a.col_ptr + m1 * (N + 1);
                    const short* cidx = a.col_idx + m1 * NN;
#pragma unroll
                    for (int j = 0; j < N; ++j) {
                        for (int t = cp[j]; t < cp[j + 1]; ++t) {
                            const int l = cidx[t];
                            const double2 q = cmul(aR, Qm[l * N + j]);
                            cfma(r[j], q, ldg2(pn + row * N + l));
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < N; ++j) k_s[(sub * N + row) * LD + j] = r[j];
            }
        }
        __syncwarp();
        // flat epilogue: coalesced 128-bit streaming traffic on y / acc / outputs.
        // All loads of the group are issued before the first store so that they
        // overlap (the compiler cannot prove the outputs do not alias the inputs).
        const long long gbase = boff + base * NN;
        constexpr int EIT = (APW * NN + 31) / 32;
        double2 yv[EIT], bs[EIT];
#pragma unroll
        for (int it = 0; it < EIT; ++it) {
            const int e = lane + 32 * it;
            if (e < nelem) {
                const long long gi = gbase + e;
                if (!a.first) {
                    bs[it] = ld_stream(a.acc + gi);
                    if (!a.last) yv[it] = ld_stream(a.y + gi);
                }
            }
        }
#pragma unroll
        for (int it = 0; it < EIT; ++it) {
            const int e = lane + 32 * it;
            if (e < nelem) {
                const int s = e / NN, rr = e - s * NN, i = rr / N, j = rr - i * N;
                const int si = (s * N + i) * LD + j;
                const double2 k = k_s[si];
                const long long gi = gbase + e;
                if (a.first) {
                    yv[it] = rho_s[si];
                    bs[it] = yv[it];
                }
                const double2 res = make_double2(fma(a.w, k.x, bs[it].x), fma(a.w, k.y, bs[it].y));
                if (a.last) {
                    st_stream(a.ydst + gi, res);
                    if (a.traj && base + s == a.slot0)
                        a.traj[b * a.traj_bstride + (step + 1) * NN + rr] = res;
                } else {
                    st_stream(a.acc + gi, res);
                    st_stream(a.yout + gi,
                              make_double2(fma(a.a, k.x, yv[it].x), fma(a.a, k.y, yv[it].y)));
                }
            }
        }
        __syncwarp();
    }
#undef HEL
}

