// heom_stage_generic.cuh - kernel 2, the generic one-CTA-per-ADO stage kernel (any N).
// Included by heom_kernels.cu only (one translation unit); split out for readability.
#pragma once
#include "heom_core.cuh"
#include "heom_device.cuh"

// Kernel 2 (any N): one CTA per ADO, one thread per matrix element (strided);
// all operators (H and Q_m) go through their sparsity lists, so cost scales
// with nnz.  rho_n is staged in shared memory; neighbours are read through L2.
__global__ void __launch_bounds__(1024) stage_generic_kernel(const StageArgs a) {
    extern __shared__ double2 smem[];
    const int N = a.N, NN = N * N, M1 = 1 + a.nmod;
    // shared memory: rho_n | operator values [1+M][NN] | links of this ADO (slot, alphaL, alphaR)
    //                | sparsity lists (row_ptr, row_idx, col_ptr, col_idx as shorts)
    double2* rho_s = smem;
    double2* ops_s = rho_s + NN;
    double2* lcf_s = ops_s + (size_t)M1 * NN;          // [maxl][2]
    const int maxl = 2 * a.nind;
    int2* lk_s = (int2*)(lcf_s + 2 * maxl);            // [maxl]
    short* rp_s = (short*)(lk_s + maxl);
    short* ri_s = rp_s + M1 * (N + 1);
    short* cp_s = ri_s + M1 * NN;
    short* ci_s = cp_s + M1 * (N + 1);
    const int b = blockIdx.y;
    const double2* __restrict__ ops = a.ops + (long long)b * a.ops_bstride;
    for (int e = threadIdx.x; e < M1 * NN; e += blockDim.x) {
        ops_s[e] = ops[e];
        ri_s[e] = a.row_idx[e];
        ci_s[e] = a.col_idx[e];
    }
    for (int e = threadIdx.x; e < M1 * (N + 1); e += blockDim.x) {
        rp_s[e] = a.row_ptr[e];
        cp_s[e] = a.col_ptr[e];
    }
    const long long boff = (long long)b * a.nmax * NN;
    const double2* __restrict__ yin = a.yin + boff;
    const long long step = a.traj ? (*a.step_base + a.local_step) : 0;
    for (long long slot = a.slot_lo + blockIdx.x; slot < a.slot_hi; slot += gridDim.x) {
        const int lbeg = a.link_ptr[slot], nl = a.link_ptr[slot + 1] - lbeg;
        __syncthreads();   // previous ADO fully consumed (and tables loaded on the first pass)
        for (int e = threadIdx.x; e < NN; e += blockDim.x) rho_s[e] = ldg2(yin + slot * NN + e);
        for (int t = threadIdx.x; t < nl; t += blockDim.x) {
            const int2 lk = __ldg(a.links + lbeg + t);
            const int ci = heom::meta_ci(lk.y, a.nind, a.lmax);
            lk_s[t] = lk;
            lcf_s[2 * t] = a.coef[2 * ci];
            lcf_s[2 * t + 1] = a.coef[2 * ci + 1];
        }
        const double2 d = a.damp[slot];
        __syncthreads();
        for (int e = threadIdx.x; e < NN; e += blockDim.x) {
            const int i = e / N, j = e - i * N;
            const long long gi = boff + slot * NN + e;
            // the epilogue's operands are fetched now so that they are in flight
            // while the right-hand side is evaluated
            double2 yv = make_double2(0.0, 0.0), bs = make_double2(0.0, 0.0);
            if (!a.first) {
                bs = ld_stream(a.acc + gi);
                if (!a.last) yv = ld_stream(a.y + gi);
            }
            const double2 own = rho_s[e];
            double2 v = make_double2(-(d.x * own.x - d.y * own.y), -(d.x * own.y + d.y * own.x));
            for (int t = rp_s[i]; t < rp_s[i + 1]; ++t) {   // -i H rho
                const int l = ri_s[t];
                const double2 h = ops_s[i * N + l];
                cfma(v, make_double2(h.y, -h.x), rho_s[l * N + j]);
            }
            for (int t = cp_s[j]; t < cp_s[j + 1]; ++t) {   // +i rho H
                const int l = ci_s[t];
                const double2 h = ops_s[l * N + j];
                cfma(v, make_double2(-h.y, h.x), rho_s[i * N + l]);
            }
            for (int lp = 0; lp < nl; ++lp) {
                const int2 lk = lk_s[lp];
                const double2* __restrict__ pn = yin + (long long)lk.x * NN;
                const int m1 = 1 + heom::meta_mode(lk.y);
                const double2* Qm = ops_s + m1 * NN;
                const short* rp = rp_s + m1 * (N + 1);
                const short* ri = ri_s + m1 * NN;
                const short* cp = cp_s + m1 * (N + 1);
                const short* cx = ci_s + m1 * NN;
                const int r0 = rp[i], r1 = rp[i + 1], c0 = cp[j], c1 = cp[j + 1];
                double2 sl = make_double2(0.0, 0.0), sr = make_double2(0.0, 0.0);
                if (r1 - r0 <= 2 && c1 - c0 <= 2) {
                    // sparse coupling operator: issue the (at most four) neighbour loads together
                    double2 q[4], x[4];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const bool lv = r0 + u < r1, rv = c0 + u < c1;
                        const int ll = lv ? ri[r0 + u] : 0, lr = rv ? cx[c0 + u] : 0;
                        q[u] = lv ? Qm[i * N + ll] : make_double2(0.0, 0.0);
                        q[2 + u] = rv ? Qm[lr * N + j] : make_double2(0.0, 0.0);
                        x[u] = lv ? ldg2(pn + ll * N + j) : make_double2(0.0, 0.0);
                        x[2 + u] = rv ? ldg2(pn + i * N + lr) : make_double2(0.0, 0.0);
                    }
                    cfma(sl, q[0], x[0]);
                    cfma(sl, q[1], x[1]);
                    cfma(sr, q[2], x[2]);
                    cfma(sr, q[3], x[3]);
                } else {
                    for (int t = r0; t < r1; ++t) {
                        const int l = ri[t];
                        cfma(sl, Qm[i * N + l], ldg2(pn + l * N + j));
                    }
                    for (int t = c0; t < c1; ++t) {
                        const int l = cx[t];
                        cfma(sr, Qm[l * N + j], ldg2(pn + i * N + l));
                    }
                }
                cfma(v, lcf_s[2 * lp], sl);
                cfma(v, lcf_s[2 * lp + 1], sr);
            }
            if (a.last) {
                if (a.first) bs = own;
                const double2 res = make_double2(fma(a.w, v.x, bs.x), fma(a.w, v.y, bs.y));
                a.ydst[gi] = res;
                if (a.traj && slot == a.slot0)
                    a.traj[b * a.traj_bstride + (step + 1) * NN + e] = res;
            } else {
                if (a.first) {
                    yv = own;
                    bs = own;
                }
                a.acc[gi] = make_double2(fma(a.w, v.x, bs.x), fma(a.w, v.y, bs.y));
                a.yout[gi] = make_double2(fma(a.a, v.x, yv.x), fma(a.a, v.y, yv.y));
            }
        }
    }
}

