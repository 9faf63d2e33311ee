// heom_resident.cuh - kernels 4 and 5, cluster-resident propagation of small hierarchies.
// Included by heom_kernels.cu only (one translation unit); split out for readability.
#pragma once
#include "heom_core.cuh"
#include "heom_device.cuh"
#include <cooperative_groups.h>

// ---------------------------------------------------------------------------
// Kernel 4: cluster-resident propagation for small hierarchies (N <= 8, diagonal
// Q_m).  One thread-block cluster per trajectory keeps the whole hierarchy - y,
// acc and both stage buffers - in distributed shared memory for all nt steps;
// neighbour rows are read from the owning CTA's shared memory (DSMEM) and the
// only synchronisation per RK stage is a hardware cluster barrier.  Hierarchies
// of a few hundred ADOs (BASELINE configs 1, 2, 5) are otherwise bound by
// launch and L2 latency, not bandwidth.
// ---------------------------------------------------------------------------
struct ResidentArgs {
    StageArgs s;          // tables, traj, herm, ...; array pointers: s.y = state (global)
    const double* fsys;   // [B][nt][3] or null
    const double* fcoup;
    const double2* ops_base;  // [1+M][NN]
    const double2* ops_dip;
    double dt;
    long long nt;
    int apc;              // ADOs per CTA (multiple of 32/N)
    int tdep;
    int maxlinks;         // links per ADO, upper bound (sizes the per-warp link cache)
};

template <int N, bool HREAL>
__global__ void __launch_bounds__(512, 1)
resident_cluster_kernel(const ResidentArgs ra, const __grid_constant__ HParam<N> hp) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const StageArgs& a = ra.s;
    constexpr int NN = N * N, APW = 32 / N, LD = (N % 2 == 0) ? N + 1 : N, ADO = N * LD;
    constexpr int FLAT = APW * NN, EIT = (FLAT + 31) / 32;
    extern __shared__ double2 smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int crank = (int)cluster.block_rank(), csize = (int)cluster.num_blocks();
    const int b = blockIdx.x / csize;       // trajectory
    const int apc = ra.apc;
    const AsyncTables T = async_tables(N, a.nind, a.nmod, a.lmax, true);
    double2* Hs = smem + T.H;
    double2* cb_s = smem + T.cb;
    double2* cq_s = smem + T.cq;
    double2* qd_s = smem + T.qd;
    double* sq_s = (double*)(smem + T.sq);
    double2* arr0 = smem + T.warp0;          // 4 arrays of apc ADOs each
    double2* Yb = arr0;
    double2* ACCb = Yb + (size_t)apc * ADO;
    double2* SAb = ACCb + (size_t)apc * ADO;
    double2* SBb = SAb + (size_t)apc * ADO;
    // per-ADO link cache: (generic pointer to the neighbour's row in its CTA's Y array, meta)
    struct LinkEnt { const double2* rowp; int meta; int pad; };
    LinkEnt* lk_s = (LinkEnt*)(SBb + (size_t)apc * ADO);
    unsigned char* supp_s = (unsigned char*)(lk_s + (size_t)apc * ra.maxlinks);
    const unsigned char* insupp_s = supp_s + a.nmod * (N + 1);

    // ---- static tables
    for (int e = threadIdx.x; e < 4 * a.nind; e += blockDim.x) cb_s[e] = a.cbase[e];
    for (int e = threadIdx.x; e <= a.lmax; e += blockDim.x) sq_s[e] = sqrt((double)e);
    for (int e = threadIdx.x; e < a.nmod * (2 * N + 1); e += blockDim.x) supp_s[e] = a.supp[e];
    auto load_ops = [&](long long step, int tidx) {
        // H(t), diag Q_m(t) and the single-row coefficient table (generate_time, deom.py:676-687)
        const double fs = (ra.tdep && ra.fsys) ? ra.fsys[((long long)b * ra.nt + step) * 3 + tidx] : 0.0;
        const double fc = (ra.tdep && ra.fcoup) ? ra.fcoup[((long long)b * ra.nt + step) * 3 + tidx] : 0.0;
        for (int e = threadIdx.x; e < NN; e += blockDim.x) {
            const double2 v = ra.ops_base[e], d = ra.ops_dip[e];
            Hs[e] = make_double2(fma(d.x, fs, v.x), fma(d.y, fs, v.y));
        }
        for (int e = threadIdx.x; e < a.nmod * N; e += blockDim.x) {
            const int m = e / N, j = e - m * N, o = (1 + m) * NN + j * N + j;
            const double2 v = ra.ops_base[o], d = ra.ops_dip[o];
            qd_s[e] = make_double2(fma(d.x, fc, v.x), fma(d.y, fc, v.y));
        }
        __syncthreads();
        for (int e = threadIdx.x; e < 2 * a.nind; e += blockDim.x) {
            const int k = e >> 1, dir = e & 1;
            const int m = a.kmode[k] & 0xff, r0 = a.kmode[k] >> 8;
            const double2 q = qd_s[m * N + r0];
            const double2 bL = cb_s[4 * k + 2 * dir], bR = cb_s[4 * k + 2 * dir + 1];
            cq_s[3 * e + 0] = cmul(bL, q);
            cq_s[3 * e + 1] = cmul(make_double2(bL.x + bR.x, bL.y + bR.y), q);
            cq_s[3 * e + 2] = cmul(bR, q);
        }
        __syncthreads();
    };
    __syncthreads();
    load_ops(0, 0);

    // ---- this warp's ADOs
    const int sub = lane / N, row = lane - sub * N;
    const bool lane_ok = lane < APW * N;
    const unsigned submask = lane_ok ? (((1u << N) - 1u) << (sub * N)) : 0u;
    const int li0 = wid * APW;                                   // first local ADO of the warp
    // groups of APW consecutive slots are dealt round-robin to the CTAs of the
    // cluster, so the link-heavy low tiers (which are contiguous in the reference
    // order) do not all land in one CTA: group g lives in CTA g % csize
    const long long slot_w = ((long long)wid * csize + crank) * APW;   // its global slot
    const long long slot = slot_w + sub;
    const int cnt = (int)max(0ll, min((long long)APW, a.nmax - slot_w));
    const bool on = lane_ok && sub < cnt;
    const int nelem = cnt * NN;
    const long long boff = (long long)b * a.nmax * NN;
    int pofs[EIT];   // flat element -> offset inside the warp's (padded) tile
#pragma unroll
    for (int it = 0; it < EIT; ++it) {
        const int e = lane + 32 * it;
        const int s2 = e / NN, r = e - s2 * NN, i = r / N, j = r - i * N;
        pofs[it] = (s2 * N + i) * LD + j;
    }
    const int woff = li0 * ADO;                                   // warp tile offset in each array
    double2 damp = make_double2(0.0, 0.0);
    int lbeg = 0, lend = 0;
    if (on) {
        damp = a.damp[slot];
        lbeg = a.link_ptr[slot];
        lend = a.link_ptr[slot + 1];
    }
    const int nl = lend - lbeg;
    LinkEnt* const mylk = lk_s + (size_t)(li0 + sub) * ra.maxlinks;
    if (on) {
        for (int t = row; t < nl; t += N) {
            const int2 lk = __ldg(a.links + lbeg + t);
            const int og = lk.x / APW;                       // owner group of the neighbour
            const int orank = og % csize, oli = (og / csize) * APW + (lk.x - og * APW);
            LinkEnt en;
            en.rowp = cluster.map_shared_rank(Yb, orank) + (size_t)oli * ADO + heom::meta_r0(lk.y) * LD;
            en.meta = lk.y;
            en.pad = 0;
            mylk[t] = en;
        }
    }
    // initial state from global memory; the other arrays start at zero
#pragma unroll
    for (int it = 0; it < EIT; ++it) {
        const int e = lane + 32 * it;
        if (e < FLAT) {
            const double2 z = make_double2(0.0, 0.0);
            Yb[woff + pofs[it]] = e < nelem ? a.y[boff + slot_w * NN + e] : z;
            ACCb[woff + pofs[it]] = z;
            SAb[woff + pofs[it]] = z;
            SBb[woff + pofs[it]] = z;
        }
    }
    cluster.sync();

#define HEL(r_, c_) (Hs[(r_) * N + (c_)])
    for (long long step = 0; step < ra.nt; ++step) {
        for (int st = 0; st < 4; ++st) {
            double2* inb = st == 0 ? Yb : (st == 2 ? SBb : SAb);
            double2* outb = st == 0 ? SAb : (st == 1 ? SBb : (st == 2 ? SAb : Yb));
            const double ac = st == 2 ? ra.dt : ra.dt * 0.5;
            const double wc = (st == 0 || st == 3) ? ra.dt / 6.0 : ra.dt / 3.0;
            if (ra.tdep && st != 2 && !(step == 0 && st == 0)) load_ops(step, st == 0 ? 0 : (st == 3 ? 2 : 1));
            // stage 3 writes y in place: its k tile lives in SB (free at that point)
            double2* kt = (st == 3 ? SBb : outb) + woff;
            const double2* rsub = inb + woff + sub * ADO;
            double2* ksub = kt + sub * ADO;
            if (on) {
                double2 col[N];
#pragma unroll
                for (int l = 0; l < N; ++l) col[l] = rsub[l * LD + row];
#pragma unroll
                for (int rr = 0; rr < N; ++rr) {
                    double2 c = make_double2(0.0, 0.0);
#pragma unroll
                    for (int l = 0; l < N; ++l) {
                        if (HREAL) {
                            const double h = HEL(rr, l).x;
                            c.x = fma(h, col[l].x, c.x);
                            c.y = fma(h, col[l].y, c.y);
                        } else {
                            cfma(c, HEL(rr, l), col[l]);
                        }
                    }
                    ksub[rr * LD + row] = c;
                }
            }
            __syncwarp();
            if (on) {
                double2 rv[N];
#pragma unroll
                for (int l = 0; l < N; ++l) rv[l] = rsub[row * LD + l];
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    double2 t = ksub[row * LD + j];
#pragma unroll
                    for (int l = 0; l < N; ++l) {
                        if (HREAL) {
                            const double h = HEL(l, j).x;
                            t.x = fma(-h, rv[l].x, t.x);
                            t.y = fma(-h, rv[l].y, t.y);
                        } else {
                            cfms(t, rv[l], HEL(l, j));
                        }
                    }
                    ksub[row * LD + j] = make_double2(t.y - (damp.x * rv[j].x - damp.y * rv[j].y),
                                                      -t.x - (damp.x * rv[j].y + damp.y * rv[j].x));
                }
            }
            __syncwarp();
            // ---- neighbour terms: rows read through distributed shared memory
            if (on) {
                double2 X = make_double2(0.0, 0.0), Y = make_double2(0.0, 0.0);
                int cur_rr = -1;
                bool yused = false;
                auto flush = [&]() {
                    double2* d1 = ksub + cur_rr * LD + row;
                    double2 v1 = *d1;
                    v1.x += X.x;
                    v1.y += X.y;
                    *d1 = v1;
                    if (yused) {
                        double2* d2 = ksub + row * LD + cur_rr;
                        double2 v2 = *d2;
                        v2.x += Y.x;
                        v2.y += Y.y;
                        *d2 = v2;
                    }
                };
                const ptrdiff_t boffs = inb - Yb;   // same layout in every CTA of the cluster
                constexpr int U = 4;
                for (int c0 = 0; c0 < nl; c0 += U) {
                    LinkEnt en[U];
                    double2 A[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        en[u] = mylk[min(c0 + u, nl - 1)];
                        A[u] = en[u].rowp[boffs + row];   // first support row, element `row`
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (c0 + u < nl) {
                            const int meta = en[u].meta;
                            const int m = heom::meta_mode(meta);
                            const int r0 = heom::meta_r0(meta);
                            const double2* rin = en[u].rowp + boffs - r0 * LD;   // neighbour ADO base
                            const double sq = sq_s[heom::meta_neff(meta)];
                            const int ns = supp_s[m * (N + 1)];
                            const int kd = heom::meta_kdir(meta);
                            const double2 qj = qd_s[m * N + row];
                            const bool outside = insupp_s[m * N + row] == 0;
                            for (int t2 = 0; t2 < ns; ++t2) {
                                const int rr = supp_s[m * (N + 1) + 1 + t2];
                                const double2 Aj = (t2 == 0) ? A[u] : rin[rr * LD + row];
                                if (rr != cur_rr) {
                                    if (cur_rr >= 0) {
                                        flush();
                                        __syncwarp(submask);
                                    }
                                    cur_rr = rr;
                                    X = make_double2(0.0, 0.0);
                                    Y = make_double2(0.0, 0.0);
                                    yused = false;
                                }
                                double2 c;
                                if (ns == 1) {
                                    const double2 c1 = cq_s[3 * kd + (row == rr ? 1 : 0)];
                                    c = make_double2(c1.x * sq, c1.y * sq);
                                } else {
                                    const double2 bL = cb_s[2 * kd], bR = cb_s[2 * kd + 1];
                                    c = cmul(make_double2(bL.x * sq, bL.y * sq), qd_s[m * N + rr]);
                                    cfma(c, make_double2(bR.x * sq, bR.y * sq), qj);
                                }
                                cfma(X, c, Aj);
                                if (outside) {
                                    const double2 bR = cb_s[2 * kd + 1];
                                    const double2 cr =
                                        cmul(make_double2(bR.x * sq, bR.y * sq), qd_s[m * N + rr]);
                                    const double2 Bj = a.herm ? make_double2(Aj.x, -Aj.y) : rin[row * LD + rr];
                                    cfma(Y, cr, Bj);
                                    yused = true;
                                }
                            }
                        }
                    }
                }
                if (cur_rr >= 0) flush();
            }
            __syncwarp();
            // ---- stage update in shared memory
#pragma unroll
            for (int it = 0; it < EIT; ++it) {
                const int e = lane + 32 * it;
                if (e < nelem) {
                    const int o = woff + pofs[it];
                    const double2 k = kt[pofs[it]];
                    if (st == 3) {
                        const double2 bs = ACCb[o];
                        const double2 res = make_double2(fma(wc, k.x, bs.x), fma(wc, k.y, bs.y));
                        Yb[o] = res;
                        if (a.traj && slot_w + e / NN == a.slot0)
                            a.traj[b * a.traj_bstride + (step + 1) * NN + e % NN] = res;
                    } else {
                        const double2 yv = Yb[o];
                        const double2 bs = st == 0 ? yv : ACCb[o];
                        ACCb[o] = make_double2(fma(wc, k.x, bs.x), fma(wc, k.y, bs.y));
                        outb[o] = make_double2(fma(ac, k.x, yv.x), fma(ac, k.y, yv.y));
                    }
                }
            }
            cluster.sync();
        }
    }
#undef HEL
    // ---- final state back to global memory
#pragma unroll
    for (int it = 0; it < EIT; ++it) {
        const int e = lane + 32 * it;
        if (e < nelem) const_cast<double2*>(a.y)[boff + slot_w * NN + e] = Yb[woff + pofs[it]];
    }
}

// ---------------------------------------------------------------------------
// Kernel 5: cluster-resident propagation, element-parallel.  Same residency as
// kernel 4 (whole hierarchy in distributed shared memory, one cluster per
// trajectory, one hardware cluster barrier per RK stage) but a whole warp works
// on one ADO, one or two matrix elements per lane: the dependent chain per lane
// is 2N complex FMAs instead of 2N^2, which is what bounds tiny hierarchies.
// For diagonal Q_m the coupling is element-wise, so a lane reads exactly its own
// element of each neighbour through DSMEM - no Hermiticity assumption needed.
// ADOs are dealt round-robin to the CTAs of the cluster (slot s -> CTA s % C).
// ---------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(512, 1)
resident_elem_kernel(const ResidentArgs ra) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const StageArgs& a = ra.s;
    constexpr int NN = N * N, EPL = (NN + 31) / 32;
    extern __shared__ double2 smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int crank = (int)cluster.block_rank(), csize = (int)cluster.num_blocks();
    const int b = blockIdx.x / csize;
    const int apc = ra.apc;
    struct LinkEnt { const double2* tile; int meta; int pad; };
    struct AdoMeta { double2 damp; int nl; int pad; };
    // shared-memory carve-up
    double2* Hs = smem;
    double2* qd_s = Hs + NN;
    double2* cb_s = qd_s + a.nmod * N;
    double* sq_s = (double*)(cb_s + 4 * a.nind);
    double2* Yb = (double2*)(sq_s + ((a.lmax + 2) & ~1));
    double2* ACCb = Yb + (size_t)apc * NN;
    double2* SAb = ACCb + (size_t)apc * NN;
    double2* SBb = SAb + (size_t)apc * NN;
    LinkEnt* lk_s = (LinkEnt*)(SBb + (size_t)apc * NN);
    AdoMeta* am_s = (AdoMeta*)(lk_s + (size_t)apc * ra.maxlinks);

    for (int e = threadIdx.x; e < 4 * a.nind; e += blockDim.x) cb_s[e] = a.cbase[e];
    for (int e = threadIdx.x; e <= a.lmax; e += blockDim.x) sq_s[e] = sqrt((double)e);
    auto load_ops = [&](long long step, int tidx) {
        const double fs = (ra.tdep && ra.fsys) ? ra.fsys[((long long)b * ra.nt + step) * 3 + tidx] : 0.0;
        const double fc = (ra.tdep && ra.fcoup) ? ra.fcoup[((long long)b * ra.nt + step) * 3 + tidx] : 0.0;
        __syncthreads();
        for (int e = threadIdx.x; e < NN; e += blockDim.x) {
            const double2 v = ra.ops_base[e], d = ra.ops_dip[e];
            Hs[e] = make_double2(fma(d.x, fs, v.x), fma(d.y, fs, v.y));
        }
        for (int e = threadIdx.x; e < a.nmod * N; e += blockDim.x) {
            const int m = e / N, j = e - m * N, o = (1 + m) * NN + j * N + j;
            const double2 v = ra.ops_base[o], d = ra.ops_dip[o];
            qd_s[e] = make_double2(fma(d.x, fc, v.x), fma(d.y, fc, v.y));
        }
        __syncthreads();
    };
    load_ops(0, 0);

    int ei[EPL], ej[EPL];
    bool ev[EPL];
#pragma unroll
    for (int t = 0; t < EPL; ++t) {
        const int e = lane + 32 * t;
        ev[t] = e < NN;
        ei[t] = ev[t] ? e / N : 0;
        ej[t] = ev[t] ? e - (e / N) * N : 0;
    }
    const long long boff = (long long)b * a.nmax * NN;
    // ---- per-ADO setup: state from global memory, link cache, damping
    for (int li = wid; li < apc; li += nwarps) {
        const long long slot = (long long)li * csize + crank;
        const bool on = slot < a.nmax;
        int lbeg = 0, nl = 0;
        double2 damp = make_double2(0.0, 0.0);
        if (on) {
            lbeg = a.link_ptr[slot];
            nl = a.link_ptr[slot + 1] - lbeg;
            damp = a.damp[slot];
        }
        if (lane == 0) {
            AdoMeta m;
            m.damp = damp;
            m.nl = nl;
            m.pad = 0;
            am_s[li] = m;
        }
        for (int t = lane; t < nl; t += 32) {
            const int2 lk = __ldg(a.links + lbeg + t);
            LinkEnt en;
            en.tile = cluster.map_shared_rank(Yb, lk.x % csize) + (size_t)(lk.x / csize) * NN;
            en.meta = lk.y;
            en.pad = 0;
            lk_s[(size_t)li * ra.maxlinks + t] = en;
        }
#pragma unroll
        for (int t = 0; t < EPL; ++t)
            if (ev[t]) {
                const int e = lane + 32 * t;
                const double2 z = make_double2(0.0, 0.0);
                Yb[li * NN + e] = on ? a.y[boff + slot * NN + e] : z;
                ACCb[li * NN + e] = z;
                SAb[li * NN + e] = z;
                SBb[li * NN + e] = z;
            }
    }
    cluster.sync();

    for (long long step = 0; step < ra.nt; ++step) {
        for (int st = 0; st < 4; ++st) {
            double2* inb = st == 0 ? Yb : (st == 2 ? SBb : SAb);
            double2* outb = st == 0 ? SAb : (st == 1 ? SBb : (st == 2 ? SAb : Yb));
            const double ac = st == 2 ? ra.dt : ra.dt * 0.5;
            const double wc = (st == 0 || st == 3) ? ra.dt / 6.0 : ra.dt / 3.0;
            if (ra.tdep && st != 2 && !(step == 0 && st == 0)) load_ops(step, st == 0 ? 0 : (st == 3 ? 2 : 1));
            const ptrdiff_t boffs = inb - Yb;
            for (int li = wid; li < apc; li += nwarps) {
                const long long slot = (long long)li * csize + crank;
                if (slot >= a.nmax) continue;   // warp-uniform
                const double2* in = inb + (size_t)li * NN;
                const AdoMeta am = am_s[li];
                const LinkEnt* mylk = lk_s + (size_t)li * ra.maxlinks;
                double2 k[EPL];
#pragma unroll
                for (int t = 0; t < EPL; ++t) {
                    k[t] = make_double2(0.0, 0.0);
                    if (ev[t]) {
                        double2 c = make_double2(0.0, 0.0);
#pragma unroll
                        for (int l = 0; l < N; ++l) {
                            cfma(c, Hs[ei[t] * N + l], in[l * N + ej[t]]);
                            cfms(c, in[ei[t] * N + l], Hs[l * N + ej[t]]);
                        }
                        const double2 own = in[lane + 32 * t];
                        k[t] = make_double2(c.y - (am.damp.x * own.x - am.damp.y * own.y),
                                            -c.x - (am.damp.x * own.y + am.damp.y * own.x));
                    }
                }
                constexpr int U = 4;
                for (int c0 = 0; c0 < am.nl; c0 += U) {
                    LinkEnt en[U];
                    double2 A[U][EPL];
                    double2 cf[U][EPL];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const bool live = c0 + u < am.nl;
                        en[u] = mylk[min(c0 + u, am.nl - 1)];
                        const int meta = en[u].meta;
                        const int m = heom::meta_mode(meta), kd = heom::meta_kdir(meta);
                        const double sq = live ? sq_s[heom::meta_neff(meta)] : 0.0;
                        const double2 bL = cb_s[2 * kd], bR = cb_s[2 * kd + 1];
#pragma unroll
                        for (int t = 0; t < EPL; ++t) {
                            const double2 qi = qd_s[m * N + ei[t]], qj = qd_s[m * N + ej[t]];
                            double2 c = cmul(make_double2(bL.x * sq, bL.y * sq), qi);
                            cfma(c, make_double2(bR.x * sq, bR.y * sq), qj);
                            cf[u][t] = c;
                            const bool need = ev[t] && live && (c.x != 0.0 || c.y != 0.0);
                            A[u][t] = need ? en[u].tile[boffs + lane + 32 * t] : make_double2(0.0, 0.0);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int t = 0; t < EPL; ++t) cfma(k[t], cf[u][t], A[u][t]);
                }
                // stage update in shared memory (each lane owns its elements)
#pragma unroll
                for (int t = 0; t < EPL; ++t)
                    if (ev[t]) {
                        const int o = li * NN + lane + 32 * t;
                        if (st == 3) {
                            const double2 bs = ACCb[o];
                            const double2 res = make_double2(fma(wc, k[t].x, bs.x), fma(wc, k[t].y, bs.y));
                            Yb[o] = res;
                            if (a.traj && slot == a.slot0)
                                a.traj[b * a.traj_bstride + (step + 1) * NN + lane + 32 * t] = res;
                        } else {
                            const double2 yv = Yb[o];
                            const double2 bs = st == 0 ? yv : ACCb[o];
                            ACCb[o] = make_double2(fma(wc, k[t].x, bs.x), fma(wc, k[t].y, bs.y));
                            outb[o] = make_double2(fma(ac, k[t].x, yv.x), fma(ac, k[t].y, yv.y));
                        }
                    }
            }
            cluster.sync();
        }
    }
    for (int li = wid; li < apc; li += nwarps) {
        const long long slot = (long long)li * csize + crank;
        if (slot >= a.nmax) continue;
#pragma unroll
        for (int t = 0; t < EPL; ++t)
            if (ev[t]) const_cast<double2*>(a.y)[boff + slot * NN + lane + 32 * t] = Yb[li * NN + lane + 32 * t];
    }
}

