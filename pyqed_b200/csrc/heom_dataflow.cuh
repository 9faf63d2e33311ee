// heom_dataflow.cuh - kernel 8: all nt RK4 steps of a small hierarchy with a large system
// matrix (8 < N <= 32; BASELINE configs[3]: cavity-molecule polariton, N = 32, 210 ADOs) in ONE
// persistent launch, synchronised ADO to ADO instead of launch to launch.
//
// Same arithmetic as the generic stage kernel (kernel 2; generate_dot_element,
// pyqed/heom/deom.py:641-664, and the accumulator form of rk4, deom.py:725-766): every
// operator goes through its sparsity lists, one thread per matrix element.  What changes is the
// control: the CTAs of a cooperative launch (all co-resident) each own a fixed set of ADOs and
// walk through the stages of all steps on their own; before evaluating stage g of an ADO a CTA
// waits until the ADOs it reads (its n+-e_k neighbours) have published stage g-1 - a release /
// acquire flag per ADO in global memory.  No grid-wide barrier, no kernel boundary, no table
// reload: a 210-ADO hierarchy is bound by launch latency and by two waves of CTAs otherwise
// (4 launches x ~25 us per step).  The state (4 arrays x 3.4 MB) lives in L2 for the whole run;
// cross-CTA reads bypass L1 (ld.global.cg).
//
// Why the two stage buffers can be reused: a CTA overwrites SA / SB / Y of its ADO only in a
// stage that it may enter after all its neighbours have finished the stage in which they read the
// old contents (they read a buffer exactly one stage after it was written).
#pragma once
#include "heom_core.cuh"
#include "heom_device.cuh"

struct DataflowArgs {
    StageArgs s;          // tables (links, coef, ops, sparsity lists, damp), slot0, traj, step_base
    double2* Y;           // state
    double2* SA;
    double2* SB;
    double2* ACC;
    unsigned* flags;      // [B * nmax][DATAFLOW_FLAG_STRIDE] stage counters, zero on entry
    double dt;
    long long nt;
    int B;                // trajectories
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

constexpr int DATAFLOW_THREADS = 512;
constexpr int DATAFLOW_FLAG_STRIDE = 32;   // one flag per 128-byte line: a flag is polled by up to 2K CTAs while its owner
                                           // writes it (dense flags cost kernel 9 a fifth of its time)
constexpr int DATAFLOW_EPT = 2;   // matrix elements per thread: N <= 32

// Per ADO and stage: (A) wait for the neighbours' flags, (B) own tile -> shared memory, (C) per
// coupling mode m: every thread sums ITS element of the mode's neighbours - S^L = sum c^L rho',
// S^R = sum c^R rho', coalesced loads that are all in flight together - and the products
// Q_m S^L + S^R Q_m are taken once per mode (element-wise for a diagonal Q_m, through shared
// memory and the sparsity lists otherwise) instead of once per link, (D) -i[H, rho] from shared
// memory, (E) the stage update.  The reference sums Q rho' per link (deom.py:656-664); by
// linearity the per-mode form is the same number up to rounding.
__global__ void __launch_bounds__(DATAFLOW_THREADS, 2) stage_dataflow_kernel(const DataflowArgs da) {
    extern __shared__ double2 smem[];
    const StageArgs& a = da.s;
    const int N = a.N, NN = N * N, M1 = 1 + a.nmod;
    constexpr int EPT = DATAFLOW_EPT;
    double2* rho_s = smem;
    double2* sl_s = rho_s + NN;                         // S^L of the current mode
    double2* sr_s = sl_s + NN;                          // S^R
    double2* ops_s = sr_s + NN;
    const int maxl = 2 * a.nind;
    double2* lcf_s = ops_s + (size_t)M1 * NN;          // [maxl][2]
    int2* lk_s = (int2*)(lcf_s + 2 * maxl);            // [maxl]
    short* rp_s = (short*)(lk_s + maxl);
    short* ri_s = rp_s + M1 * (N + 1);
    short* cp_s = ri_s + M1 * NN;
    short* ci_s = cp_s + M1 * (N + 1);
    int* qdiag_s = (int*)(ci_s + M1 * NN);   // [M1]: operator o is diagonal (the four lists hold an even number of shorts)
    for (int e = threadIdx.x; e < M1 * NN; e += blockDim.x) {
        ops_s[e] = a.ops[e];
        ri_s[e] = a.row_idx[e];
        ci_s[e] = a.col_idx[e];
    }
    for (int e = threadIdx.x; e < M1 * (N + 1); e += blockDim.x) {
        rp_s[e] = a.row_ptr[e];
        cp_s[e] = a.col_ptr[e];
    }
    __syncthreads();
    for (int o = threadIdx.x; o < M1; o += blockDim.x) {
        int diag = 1;
        for (int i = 0; i < N; ++i)
            for (int t = rp_s[o * (N + 1) + i]; t < rp_s[o * (N + 1) + i + 1]; ++t)
                if (ri_s[o * NN + t] != i) diag = 0;
        qdiag_s[o] = diag;
    }
    const long long total = a.nmax * (long long)da.B;     // work items: (trajectory, ADO)
    const double dt = da.dt;
    for (long long step = 0; step < da.nt; ++step) {
        for (int stage = 0; stage < 4; ++stage) {
            const unsigned g = (unsigned)(4 * step + stage) + 1u;   // this stage's counter value
            const double2* yin = stage == 0 ? da.Y : (stage == 1 ? da.SA : (stage == 2 ? da.SB : da.SA));
            double2* yout = stage == 0 ? da.SA : (stage == 1 ? da.SB : da.SA);
            const double ca = stage == 2 ? dt : 0.5 * dt;
            const double cw = (stage == 0 || stage == 3) ? dt / 6.0 : dt / 3.0;
            for (long long item = blockIdx.x; item < total; item += gridDim.x) {
                const int b = (int)(item / a.nmax);
                const long long slot = item - (long long)b * a.nmax;
                const long long boff = (long long)b * a.nmax * NN;
                const int lbeg = a.link_ptr[slot], nl = a.link_ptr[slot + 1] - lbeg;
                __syncthreads();   // the previous item is fully consumed (tables loaded on the first pass)
                for (int t = threadIdx.x; t < nl; t += blockDim.x) {
                    const int2 lk = __ldg(a.links + lbeg + t);
                    const int ci = heom::meta_ci(lk.y, a.nind, a.lmax);
                    lk_s[t] = lk;
                    lcf_s[2 * t] = a.coef[2 * ci];
                    lcf_s[2 * t + 1] = a.coef[2 * ci + 1];
                    // (A) the neighbours must have published the stage input this stage reads
                    const unsigned* f = da.flags + ((long long)b * a.nmax + lk.x) * DATAFLOW_FLAG_STRIDE;
                    while (ld_acquire_u32(f) < g - 1u) __nanosleep(20);
                }
                __syncthreads();
                // (B) own elements; the epilogue's operands are requested now as well
                double2 own[EPT], yv[EPT], bs[EPT], v[EPT];
                const double2 d = a.damp[slot];
#pragma unroll
                for (int u = 0; u < EPT; ++u) {
                    const int e = threadIdx.x + u * DATAFLOW_THREADS;
                    own[u] = yv[u] = bs[u] = v[u] = make_double2(0.0, 0.0);
                    if (e < NN) {
                        const long long gi = boff + slot * NN + e;
                        own[u] = __ldcg(yin + gi);
                        yv[u] = bs[u] = own[u];
                        if (stage != 0) {
                            bs[u] = __ldcg(da.ACC + gi);
                            if (stage != 3) yv[u] = __ldcg(da.Y + gi);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < EPT; ++u) {
                    const int e = threadIdx.x + u * DATAFLOW_THREADS;
                    if (e < NN) {
                        rho_s[e] = own[u];
                        v[u] = make_double2(-(d.x * own[u].x - d.y * own[u].y), -(d.x * own[u].y + d.y * own[u].x));
                    }
                }
                // (C) coupling terms, mode by mode
                for (int m = 0; m < a.nmod; ++m) {
                    double2 SL[EPT], SR[EPT];
#pragma unroll
                    for (int u = 0; u < EPT; ++u) SL[u] = SR[u] = make_double2(0.0, 0.0);
                    for (int lp = 0; lp < nl; ++lp) {
                        const int2 lk = lk_s[lp];
                        if (heom::meta_mode(lk.y) != m) continue;
                        const double2* __restrict__ pn = yin + boff + (long long)lk.x * NN;
                        const double2 cl = lcf_s[2 * lp], cr = lcf_s[2 * lp + 1];
#pragma unroll
                        for (int u = 0; u < EPT; ++u) {
                            const int e = threadIdx.x + u * DATAFLOW_THREADS;
                            if (e < NN) {
                                const double2 x = __ldcg(pn + e);
                                cfma(SL[u], cl, x);
                                cfma(SR[u], cr, x);
                            }
                        }
                    }
                    const int m1 = 1 + m;
                    const double2* Qm = ops_s + m1 * NN;
                    if (qdiag_s[m1]) {
#pragma unroll
                        for (int u = 0; u < EPT; ++u) {
                            const int e = threadIdx.x + u * DATAFLOW_THREADS;
                            if (e < NN) {
                                const int i = e / N, j = e - i * N;
                                cfma(v[u], Qm[i * N + i], SL[u]);
                                cfma(v[u], Qm[j * N + j], SR[u]);
                            }
                        }
                    } else {
                        __syncthreads();   // the previous mode's products have read S^L, S^R
#pragma unroll
                        for (int u = 0; u < EPT; ++u) {
                            const int e = threadIdx.x + u * DATAFLOW_THREADS;
                            if (e < NN) {
                                sl_s[e] = SL[u];
                                sr_s[e] = SR[u];
                            }
                        }
                        __syncthreads();
                        const short* rp = rp_s + m1 * (N + 1);
                        const short* ri = ri_s + m1 * NN;
                        const short* cp = cp_s + m1 * (N + 1);
                        const short* cx = ci_s + m1 * NN;
#pragma unroll
                        for (int u = 0; u < EPT; ++u) {
                            const int e = threadIdx.x + u * DATAFLOW_THREADS;
                            if (e < NN) {
                                const int i = e / N, j = e - i * N;
                                for (int t = rp[i]; t < rp[i + 1]; ++t) {
                                    const int l = ri[t];
                                    cfma(v[u], Qm[i * N + l], sl_s[l * N + j]);
                                }
                                for (int t = cp[j]; t < cp[j + 1]; ++t) {
                                    const int l = cx[t];
                                    cfma(v[u], Qm[l * N + j], sr_s[i * N + l]);
                                }
                            }
                        }
                    }
                }
                __syncthreads();   // rho_s is complete
                // (D) -i [H, rho] and (E) the stage update
#pragma unroll
                for (int u = 0; u < EPT; ++u) {
                    const int e = threadIdx.x + u * DATAFLOW_THREADS;
                    if (e >= NN) continue;
                    const int i = e / N, j = e - i * N;
                    const long long gi = boff + slot * NN + e;
                    double2 w = v[u];
                    for (int t = rp_s[i]; t < rp_s[i + 1]; ++t) {   // -i H rho
                        const int l = ri_s[t];
                        const double2 h = ops_s[i * N + l];
                        cfma(w, make_double2(h.y, -h.x), rho_s[l * N + j]);
                    }
                    for (int t = cp_s[j]; t < cp_s[j + 1]; ++t) {   // +i rho H
                        const int l = ci_s[t];
                        const double2 h = ops_s[l * N + j];
                        cfma(w, make_double2(-h.y, h.x), rho_s[i * N + l]);
                    }
                    if (stage == 3) {
                        const double2 res = make_double2(fma(cw, w.x, bs[u].x), fma(cw, w.y, bs[u].y));
                        da.Y[gi] = res;
                        if (a.traj && slot == a.slot0) a.traj[b * a.traj_bstride + (step + 1) * NN + e] = res;
                    } else {
                        da.ACC[gi] = make_double2(fma(cw, w.x, bs[u].x), fma(cw, w.y, bs[u].y));
                        yout[gi] = make_double2(fma(ca, w.x, yv[u].x), fma(ca, w.y, yv[u].y));
                    }
                }
                __syncthreads();   // every element of this ADO's stage output is written ...
                if (threadIdx.x == 0) {
                    __threadfence();
                    st_release_u32(da.flags + ((long long)b * a.nmax + slot) * DATAFLOW_FLAG_STRIDE, g);   // ... and published
                }
            }
        }
    }
}
