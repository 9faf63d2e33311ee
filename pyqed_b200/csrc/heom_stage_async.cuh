// heom_stage_async.cuh - kernel 3, the async row kernel (stage_rows_async_kernel).
//
// Kept in its own header so that the CPU tests can compile the kernel source
// unchanged against tests/_shim/cuda_emu.h (tests/test_async_kernel_emu.py); the
// launcher and everything else stay in heom_kernels.cu.
#pragma once
#include "heom_core.cuh"
#include "heom_device.cuh"

// build-time tuning knobs (see pyqed_b200/build.py)
#ifndef HEOM_TMA_BULK
#define HEOM_TMA_BULK 1   // async kernel: 1 = contiguous tiles via cp.async.bulk + mbarrier, 2 = neighbour rows too
#endif
#ifndef HEOM_L2_HINTS
#define HEOM_L2_HINTS 0   // async kernel: L2 eviction-priority hints on the cp.async loads
#endif

// ---------------------------------------------------------------------------
// Kernel 3 (N <= 8, every Q_m diagonal): same lane mapping as kernel 1, but every
// global read of a group goes through cp.async into shared memory so that no
// registers are tied up by loads in flight:
//   group A: the group's own y_in tile + one row of each of the first N
//            neighbours of every ADO (link records are prefetched one group ahead)
//   group B: the y and acc tiles the epilogue will need
// The commutator runs while B (and later link chunks) are still in flight.
// Neighbour contributions are accumulated in registers per target row and
// flushed to the k tile once per row instead of once per link.
// RK4 is done in difference form: the three stage inputs S1 = y + dt/2 k1,
// S2 = y + dt/2 k2, S3 = y + dt k3 are all kept and the last stage writes
// y' = -y/3 + S1/3 + 2 S2/3 + S3/3 + dt/6 k4, so no accumulator array is read or
// written: 13 array passes per step instead of 16.
// ---------------------------------------------------------------------------
#ifndef HEOM_ASYNC_THREADS
#define HEOM_ASYNC_THREADS 512
#endif
constexpr int ASYNC_MAX_THREADS = HEOM_ASYNC_THREADS;

// Column entry (row, rr) of a neighbour, needed only when the ADOs are not
// Hermitian.  Kept out of line so that the address arithmetic is not hoisted
// into the common (Hermitian) path of the link loop.
static __device__ __noinline__ double2 load_neighbour_entry(const double2* yin, int nbr, int NN, int off) {
    return __ldg(yin + ((long long)nbr * NN + off));
}

// shared-memory tables of the async kernel (sizes in double2 units unless noted)
struct AsyncTables {
    int H, cb, cq, qd, sq, warp0;   // offsets in double2 units
    int bytes_tail;                 // supp (cnt/rows, membership) bytes after the warp buffers
};
__host__ __device__ inline AsyncTables async_tables(int N, int K, int M, int L, bool tdep) {
    AsyncTables t;
    int o = 0;
    t.H = o;  o += tdep ? N * N : 0;
    t.cb = o; o += 4 * K;           // general path: (mL, mR, pL, pR) per k
    t.cq = o; o += 3 * 2 * K;       // single-row path: per 2k+dir (bL q, (bL+bR) q, bR q)
    t.qd = o; o += M * N;           // diagonal entries of Q_m
    t.sq = o; o += (L + 2) / 2;     // sqrt(n) as doubles
    t.warp0 = o;
    t.bytes_tail = (M * (2 * N + 1) + 15) / 16 * 16;
    return t;
}

// SYM: every ADO is Hermitian and every Q_m has exactly one non-zero diagonal
// entry.  Then k is Hermitian too, so (i) rho H = (H rho)^dagger and one product
// gives the commutator, and (ii) a link's column update is the conjugate of its
// row update; the multi-row branch of the link loop is not compiled at all.
template <int N, bool TDEP, bool HREAL, bool PUSH, bool SYM>
__global__ void __launch_bounds__(ASYNC_MAX_THREADS, 1)
stage_rows_async_kernel(const StageArgs a, const __grid_constant__ HParam<N> hp) {
    constexpr int NN = N * N, APW = 32 / N, LD = (N % 2 == 0) ? N + 1 : N, TILE = APW * N * LD;
    constexpr int FLAT = APW * NN, PERWARP = 2 * TILE + 3 * FLAT + 2;   // +2: four mbarriers
    constexpr bool BULK_TILE = HEOM_TMA_BULK && (LD == N);   // padded tiles cannot be one bulk copy
    constexpr bool BULK_FLAT = HEOM_TMA_BULK != 0;
    constexpr bool BULK_ROWS = HEOM_TMA_BULK >= 2 && BULK_TILE;   // neighbour rows as 16N-byte bulk copies
    constexpr int EIT = (FLAT + 31) / 32;
    HEOM_DYN_SMEM(double2, smem);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    // one launch per trajectory: the host passes array, operator and trajectory
    // pointers of this batch entry, so no batch offset is carried (or rebuilt) here
    const double2* __restrict__ ops = a.ops;
    const AsyncTables T = async_tables(N, a.nind, a.nmod, a.lmax, TDEP);
    double2* Hs = smem + T.H;
    double2* cb_s = smem + T.cb;
    double2* cq_s = smem + T.cq;
    double2* qd_s = smem + T.qd;
    double* sq_s = (double*)(smem + T.sq);
    double2* warp0 = smem + T.warp0;
    // per-warp buffers; the first-stage-buffer tile (acc_s) exists only in the last
    // stage, so the other three stages fit more warps per SM
    const int perwarp = a.last ? PERWARP : PERWARP - FLAT;
    double2* rho_s = warp0 + wid * perwarp;
    double2* k_s = rho_s + TILE;
    double2* y_s = k_s + TILE;
    double2* nb_s = y_s + FLAT;
    double2* acc_s = nb_s + FLAT;   // valid only when a.last
    unsigned long long* barA = (unsigned long long*)(rho_s + perwarp - 2);   // own tile (+ rows)
    unsigned long long* barB = barA + 1;                              // y / 1st stage buffer
    unsigned long long* barC = barA + 2;                              // neighbour rows of later chunks
    unsigned long long* barD = barA + 3;                              // 2nd stage buffer (last stage)
    unsigned phA = 0, phB = 0, phC = 0, phD = 0;
    if (BULK_FLAT && lane == 0) {
        mbar_init(barA, 1);
        mbar_init(barB, 1);
        mbar_init(barC, 1);
        mbar_init(barD, 1);
        fence_proxy_async();
    }
    unsigned char* supp_s = (unsigned char*)(warp0 + nwarps * perwarp);
    if (TDEP) {
        for (int e = threadIdx.x; e < NN; e += blockDim.x) Hs[e] = ops[e];
    }
    for (int e = threadIdx.x; e < 4 * a.nind; e += blockDim.x) cb_s[e] = a.cbase[e];
    for (int e = threadIdx.x; e < a.nmod * N; e += blockDim.x) {
        const int m = e / N, j = e - m * N;
        qd_s[e] = ops[(1 + m) * NN + j * N + j];
    }
    for (int e = threadIdx.x; e <= a.lmax; e += blockDim.x) sq_s[e] = sqrt((double)e);
    for (int e = threadIdx.x; e < a.nmod * (2 * N + 1); e += blockDim.x) supp_s[e] = a.supp[e];
    for (int e = threadIdx.x; e < 2 * a.nind; e += blockDim.x) {
        // e = 2k + dir; mode and first support row of dissipaton k
        const int k = e >> 1, dir = e & 1;
        const int m = a.kmode[k] & 0xff, r0 = a.kmode[k] >> 8;
        const double2 q = ops[(1 + m) * NN + r0 * N + r0];
        const double2 bL = a.cbase[4 * k + 2 * dir], bR = a.cbase[4 * k + 2 * dir + 1];
        cq_s[3 * e + 0] = cmul(bL, q);
        cq_s[3 * e + 1] = cmul(make_double2(bL.x + bR.x, bL.y + bR.y), q);
        cq_s[3 * e + 2] = cmul(bR, q);
    }
    __syncthreads();
    const unsigned char* insupp_s = supp_s + a.nmod * (N + 1);
    const double2* __restrict__ yin = a.yin;
    const int sub = lane / N, row = lane - sub * N;
    // neighbour rows are addressed with 32-bit element offsets (host checks nmax N^2 < 2^32)
    const bool lane_ok = lane < APW * N;
    const unsigned submask = lane_ok ? (((1u << N) - 1u) << (sub * N)) : 0u;
    const long long step = a.traj ? (*a.step_base + a.local_step) : 0;
    const long long gstride = (long long)gridDim.x * nwarps;
#if HEOM_L2_HINTS
    const unsigned long long pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
#define CP_KEEP(d_, s_) cp_async16_hint(d_, s_, pol_keep)
#define CP_STREAM(d_, s_) cp_async16_hint(d_, s_, pol_stream)
#define CP_ROW(t_, s_) cp_async16_hint(nbrow + (t_) * N, s_, pol_keep)
#else
#define CP_KEEP(d_, s_) cp_async16(d_, s_)
#define CP_STREAM(d_, s_) cp_async16(d_, s_)
#define CP_ROW(t_, s_) cp_async16_s(nbrow_u32 + (t_) * (N * 16), s_)
#endif
    // flat element e = lane + 32 it  ->  offset in the (possibly padded) tile
    int pofs[EIT];
#pragma unroll
    for (int it = 0; it < EIT; ++it) {
        const int e = lane + 32 * it;
        if (LD == N) pofs[it] = e;
        else {
            const int s = e / NN, r = e - s * NN, i = r / N, j = r - i * N;
            pofs[it] = (s * N + i) * LD + j;
        }
    }
    double2* const ksub = k_s + sub * N * LD;     // this ADO's k tile
    double2* const rsub = rho_s + sub * N * LD;
    double2* const nbrow = nb_s + sub * NN + row; // + t*N: row element of staged link t
    const unsigned nbrow_u32 = smem_u32(nbrow);   // the same as a shared-window address for cp.async

    // Bookkeeping pipeline, carried in registers across iterations so that no
    // dependent global load sits on the critical path of a group:
    //   two groups ahead : link offsets (link_ptr)
    //   one group ahead  : damping rate and the first NCH*N link records
    constexpr int NCH = (32 + N - 1) / N > 4 ? 4 : (32 + N - 1) / N;  // record chunks prefetched
    long long g = (long long)blockIdx.x * nwarps + wid;
    int nx_lbeg = 0, nx_lend = 0, nn_lbeg = 0, nn_lend = 0;
    int nx_pb = 0, nx_pe = 0, nn_pb = 0, nn_pe = 0, nx_pent = 0;   // fused halo push bookkeeping
    constexpr bool pushing = PUSH;
    int2 nx_rec[NCH];
    double2 nx_damp = make_double2(0.0, 0.0);
#pragma unroll
    for (int c = 0; c < NCH; ++c) nx_rec[c] = make_int2(0, 0);
    // Groups are visited in a scrambled order: the position inside every run of
    // 16 groups is rotated by the run index, so a warp (whose stride is a
    // multiple of 16) does not see the same position of the 64-slot blocks of
    // storage order 2 every time (their head holds the link-heavy ADOs).
    const long long gfull = a.scramble ? (a.ngroups & ~15ll) : 0;
    auto gmap = [&](long long gg) {
        // rotation amount = top bits of a multiplicative hash of the run index, so that
        // every warp sees all 16 positions whatever its stride is
        const unsigned rot = ((unsigned)(gg >> 4) * 2654435761u) >> 28;
        return gg < gfull ? ((gg & ~15ll) | ((gg + rot) & 15ll)) : gg;
    };
    auto fetch_ptr = [&](long long gg, int& lb, int& le, int& pb, int& pe) {
        const long long slot = a.slot_lo + gmap(gg) * APW + sub;
        lb = le = pb = pe = 0;
        if (gg < a.ngroups && lane_ok && slot < a.slot_hi) {
            lb = a.link_ptr[slot];
            le = a.link_ptr[slot + 1];
            if (pushing) {
                pb = a.push_ptr[slot - a.slot_lo];
                pe = a.push_ptr[slot - a.slot_lo + 1];
            }
        }
    };
    auto fetch_rec = [&](long long gg, int lb, int le) {
        const long long slot = a.slot_lo + gmap(gg) * APW + sub;
        if (gg < a.ngroups && lane_ok && slot < a.slot_hi) nx_damp = a.damp[slot];
        nx_pent = 0;
        if (pushing && nx_pb + row < nx_pe) nx_pent = a.push_ent[nx_pb + row];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            nx_rec[c] = make_int2(0, 0);
            if (lb + c * N + row < le) nx_rec[c] = __ldg(a.links + lb + c * N + row);
        }
    };
    fetch_ptr(g, nx_lbeg, nx_lend, nx_pb, nx_pe);
    fetch_rec(g, nx_lbeg, nx_lend);
    fetch_ptr(g + gstride, nn_lbeg, nn_lend, nn_pb, nn_pe);

    for (; g < a.ngroups; g += gstride) {
        const long long base = a.slot_lo + gmap(g) * APW;
        const int cnt = (int)min((long long)APW, a.slot_hi - base);
        const int nelem = cnt * NN;
        const bool on = lane_ok && sub < cnt;
        const int lbeg = nx_lbeg, lend = nx_lend;
        const int nl = on ? (lend - lbeg) : 0;
        int2 recs[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) recs[c] = nx_rec[c];
        int2 rec = recs[0];
        const double2 d = nx_damp;
        const int pb = nx_pb, npush = on ? (nx_pe - nx_pb) : 0, pent = nx_pent;
        int2 rts[N];
        const long long gbase = base * NN;
        // next group's records / damping (offsets arrived during the previous
        // iteration), and the offsets of the group after that
        nx_lbeg = nn_lbeg;
        nx_lend = nn_lend;
        nx_pb = nn_pb;
        nx_pe = nn_pe;
        fetch_rec(g + gstride, nx_lbeg, nx_lend);
        fetch_ptr(g + 2 * gstride, nn_lbeg, nn_lend, nn_pb, nn_pe);

        // ---- issue: own tile + first chunk of neighbour rows (group A), y/acc (group B)
        if (BULK_TILE) {
            unsigned rowbytes = 0;
            if (BULK_ROWS) {
                const int mine = (on && row == 0) ? min(nl, N) : 0;
                rowbytes = (unsigned)__reduce_add_sync(0xffffffffu, mine) * (unsigned)(N * 16);
            }
            if (lane == 0) {
                fence_proxy_async();   // earlier generic-proxy reads of these buffers are done (warp sync)
                mbar_expect_tx(barA, nelem * 16u + rowbytes);
                bulk_g2s(rho_s, yin + base * NN, nelem * 16u, barA);
            }
            if (BULK_ROWS) {
                __syncwarp();          // the transaction count is posted before any row copy can complete
                if (on && row < nl)    // lane `row` of an ADO fetches the row its link number `row` needs
                    bulk_g2s(nb_s + (sub * N + row) * N,
                             yin + ((long long)rec.x * NN + heom::meta_r0(rec.y) * N), N * 16u, barA);
            }
        } else {
            const double2* src = yin + base * NN + lane;
#pragma unroll
            for (int it = 0; it < EIT; ++it)
                if (lane + 32 * it < nelem) CP_KEEP(&rho_s[pofs[it]], src + 32 * it);
        }
#pragma unroll
        for (int t = 0; t < N; ++t) {
            const int srcl = (sub * N + t) & 31;
            rts[t].x = __shfl_sync(0xffffffffu, rec.x, srcl);
            rts[t].y = __shfl_sync(0xffffffffu, rec.y, srcl);
            if (!BULK_ROWS && t < nl)
                CP_ROW(t,
                        yin + (((unsigned)rts[t].x * (unsigned)N + (unsigned)heom::meta_r0(rts[t].y)) * (unsigned)N + (unsigned)row));
        }
        cp_async_commit();
        if (!a.first) {
            // y always; in the last stage also the first stage buffer (passed in a.acc) -
            // the second one follows into rho_s once the commutator has consumed it
            if (BULK_FLAT) {
                if (lane == 0) {
                    if (!BULK_TILE) fence_proxy_async();
                    mbar_expect_tx(barB, nelem * 16u * (a.last ? 2u : 1u));
                    bulk_g2s(y_s, a.y + gbase, nelem * 16u, barB);
                    if (a.last) bulk_g2s(acc_s, a.acc + gbase, nelem * 16u, barB);
                }
            } else {
                const double2* sa = a.acc + gbase + lane;
                const double2* sy = a.y + gbase + lane;
#pragma unroll
                for (int it = 0; it < EIT; ++it)
                    if (lane + 32 * it < nelem) {
                        if (a.last) CP_STREAM(&acc_s[lane + 32 * it], sa + 32 * it);
                        CP_STREAM(&y_s[lane + 32 * it], sy + 32 * it);
                    }
            }
        }
        cp_async_commit();

        cp_async_wait<1>();
        if (BULK_TILE) {
            mbar_wait(barA, phA);
            phA ^= 1u;
        }
        __syncwarp();

        // ---- -i[H, rho] - damp rho
#define HEL(r_, c_) (TDEP ? Hs[(r_) * N + (c_)] : hp.v[(r_) * N + (c_)])
        double2 ccol[SYM ? N : 1];   // SYM: (H rho)[rr][row], this lane's column
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
                if constexpr (SYM) ccol[rr] = c;
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
                if constexpr (SYM) {
                    // (rho H)[row][j] = conj((H rho)[j][row])
                    t.x -= ccol[j].x;
                    t.y += ccol[j].y;
                } else {
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
                }
                double2 kv = make_double2(t.y - (d.x * rv[j].x - d.y * rv[j].y),
                                          -t.x - (d.x * rv[j].y + d.y * rv[j].x));
                if (a.last) {
                    // fold the stage input's own weight into k: w (k + (2/dt) y_in) = w k + y_in / 3
                    kv.x = fma(a.a, rv[j].x, kv.x);
                    kv.y = fma(a.a, rv[j].y, kv.y);
                }
                ksub[row * LD + j] = kv;
            }
        }
#undef HEL
        __syncwarp();
        if (a.last) {
            // rho_s is free now: fetch the second stage buffer (a.yout) into it for the epilogue
            if (BULK_TILE) {
                if (lane == 0) {
                    fence_proxy_async();
                    mbar_expect_tx(barD, nelem * 16u);
                    bulk_g2s(rho_s, a.yout + gbase, nelem * 16u, barD);
                }
            } else {
                const double2* sb = a.yout + gbase + lane;
#pragma unroll
                for (int it = 0; it < EIT; ++it)
                    if (lane + 32 * it < nelem) CP_STREAM(&rho_s[pofs[it]], sb + 32 * it);
                cp_async_commit();
            }
        }

        // ---- neighbour terms, N links per chunk; contributions to one target row
        //      are summed in registers (X: element (cur_rr,row), Y: element (row,cur_rr))
        const int maxl = __reduce_max_sync(0xffffffffu, nl);
        double2 X = make_double2(0.0, 0.0), Y = make_double2(0.0, 0.0);
        int cur_rr = -1;
        bool yused = false;
        auto flush = [&]() {
            double2* d1 = ksub + cur_rr * LD + row;
            double2 v1 = *d1;
            v1.x += X.x;
            v1.y += X.y;
            *d1 = v1;
            if constexpr (SYM) {
                if (row != cur_rr) {   // column update = conjugate of the row update
                    double2* d2 = ksub + row * LD + cur_rr;
                    double2 v2 = *d2;
                    v2.x += X.x;
                    v2.y -= X.y;
                    *d2 = v2;
                }
            } else if (yused) {
                double2* d2 = ksub + row * LD + cur_rr;
                double2 v2 = *d2;
                v2.x += Y.x;
                v2.y += Y.y;
                *d2 = v2;
            }
        };
        for (int c0 = 0; c0 < maxl; c0 += N) {
            if (c0 > 0) {
                rec = make_int2(0, 0);
                {
                    const int c = c0 / N;
                    bool have = false;
#pragma unroll
                    for (int cc = 1; cc < NCH; ++cc)
                        if (c == cc) {
                            rec = recs[cc];
                            have = true;
                        }
                    if (!have && lbeg + c0 + row < lend) rec = __ldg(a.links + lbeg + c0 + row);
                }
                __syncwarp();  // every lane is done with the previous chunk's rows
#pragma unroll
                for (int t = 0; t < N; ++t) {
                    const int srcl = (sub * N + t) & 31;
                    rts[t].x = __shfl_sync(0xffffffffu, rec.x, srcl);
                    rts[t].y = __shfl_sync(0xffffffffu, rec.y, srcl);
                    if (!BULK_ROWS && c0 + t < nl)
                        CP_ROW(t,
                                yin + (((unsigned)rts[t].x * (unsigned)N + (unsigned)heom::meta_r0(rts[t].y)) * (unsigned)N + (unsigned)row));
                }
                if (BULK_ROWS) {
                    const int mine = (on && row == 0) ? max(0, min(nl - c0, N)) : 0;
                    const unsigned rowbytes = (unsigned)__reduce_add_sync(0xffffffffu, mine) * (unsigned)(N * 16);
                    if (lane == 0) {
                        fence_proxy_async();
                        mbar_expect_tx(barC, rowbytes);
                    }
                    __syncwarp();
                    if (on && c0 + row < nl)
                        bulk_g2s(nb_s + (sub * N + row) * N,
                                 yin + ((long long)rec.x * NN + heom::meta_r0(rec.y) * N), N * 16u, barC);
                    mbar_wait(barC, phC);
                    phC ^= 1u;
                } else {
                    cp_async_commit();
                    cp_async_wait<0>();
                }
                __syncwarp();
            }
            if (c0 < nl) {
#pragma unroll
                for (int t = 0; t < N; ++t) {
                    if (c0 + t < nl) {
                        const int meta = rts[t].y;
                        const int m = heom::meta_mode(meta);
                        const int rr = heom::meta_r0(meta);
                        const double2 Aj = nbrow[t * N];
                        const double sq = sq_s[heom::meta_neff(meta)];
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
                        if constexpr (SYM) {
                            const double2 c1 = cq_s[3 * heom::meta_kdir(meta) + (row == rr ? 1 : 0)];
                            cfma(X, make_double2(c1.x * sq, c1.y * sq), Aj);
                            continue;
                        }
                        const int ns = supp_s[m * (N + 1)];
                        if (ns == 1) {
                            const int cid = heom::meta_kdir(meta);
                            const double2 c1 = cq_s[3 * cid + (row == rr ? 1 : 0)];
                            cfma(X, make_double2(c1.x * sq, c1.y * sq), Aj);
                            if (row != rr) {
                                const double2 c2 = cq_s[3 * cid + 2];
                                double2 Bj = make_double2(Aj.x, -Aj.y);
                                if (!a.herm) Bj = load_neighbour_entry(yin, rts[t].x, NN, row * N + rr);
                                cfma(Y, make_double2(c2.x * sq, c2.y * sq), Bj);
                                yused = true;
                            }
                        } else {
                            // several non-zero diagonal entries: further rows straight from global
                            const int kd = heom::meta_kdir(meta);
                            const double2 bL = cb_s[2 * kd], bR = cb_s[2 * kd + 1];
                            const double2 aL = make_double2(bL.x * sq, bL.y * sq);
                            const double2 aR = make_double2(bR.x * sq, bR.y * sq);
                            const double2* __restrict__ pn = yin + (long long)rts[t].x * NN;
                            const double2 qj = qd_s[m * N + row];
                            const bool outside = insupp_s[m * N + row] == 0;
                            for (int t2 = 0; t2 < ns; ++t2) {
                                const int r2 = supp_s[m * (N + 1) + 1 + t2];
                                const double2 A2 = (t2 == 0) ? Aj : ldg2(pn + r2 * N + row);
                                if (r2 != cur_rr) {
                                    flush();
                                    __syncwarp(submask);
                                    cur_rr = r2;
                                    X = make_double2(0.0, 0.0);
                                    Y = make_double2(0.0, 0.0);
                                    yused = false;
                                }
                                const double2 qr = qd_s[m * N + r2];
                                double2 c = cmul(aL, qr);
                                cfma(c, aR, qj);
                                cfma(X, c, A2);
                                if (outside) {
                                    const double2 B2 = a.herm ? make_double2(A2.x, -A2.y)
                                                              : ldg2(pn + row * N + r2);
                                    cfma(Y, cmul(aR, qr), B2);
                                    yused = true;
                                }
                            }
                        }
                    }
                }
            }
        }
        if (cur_rr >= 0) flush();
        cp_async_wait<0>();
        if (BULK_FLAT && !a.first) {
            mbar_wait(barB, phB);
            phB ^= 1u;
        }
        if (BULK_TILE && a.last) {
            mbar_wait(barD, phD);
            phD ^= 1u;
        }
        __syncwarp();

        // ---- epilogue from shared memory, streaming stores; rows that other ranks
        //      need go straight into their arrays (peer memory over NVLink)
        const int maxpush = PUSH ? __reduce_max_sync(0xffffffffu, npush) : 0;
#pragma unroll
        for (int it = 0; it < EIT; ++it) {
            const int e = lane + 32 * it;
            const bool live = e < nelem;
            double2 outv = make_double2(0.0, 0.0);
            long long gi = 0;
            if (live) {
                const double2 k = k_s[pofs[it]];
                gi = gbase + e;
                if (a.last) {
                    // y' = -y/3 + S1/3 + 2 S2/3 + w (k4 + (2/dt) S3)   (S3's share is already in k)
                    const double2 y0 = y_s[e], s1 = acc_s[e], s2 = rho_s[pofs[it]];
                    const double third = 1.0 / 3.0;
                    double2 res = make_double2(fma(a.w, k.x, third * (s1.x - y0.x)),
                                               fma(a.w, k.y, third * (s1.y - y0.y)));
                    res.x = fma(2.0 * third, s2.x, res.x);
                    res.y = fma(2.0 * third, s2.y, res.y);
                    outv = res;
                    st_stream(a.ydst + gi, res);
                    if (a.traj && base + e / NN == a.slot0)
                        a.traj[(step + 1) * NN + e % NN] = res;
                } else {
                    const double2 yv = a.first ? rho_s[pofs[it]] : y_s[e];
                    outv = make_double2(fma(a.a, k.x, yv.x), fma(a.a, k.y, yv.y));
                    st_stream(a.yout + gi, outv);
                }
            }
            if (PUSH && maxpush > 0) {   // warp-uniform
                const int s2 = min(e / NN, APW - 1), erow = (e - (e / NN) * NN) / N;
                const int cnt2 = __shfl_sync(0xffffffffu, npush, s2 * N);
                const int pb2 = __shfl_sync(0xffffffffu, pb, s2 * N);
                for (int t = 0; t < maxpush; ++t) {
                    int ent = __shfl_sync(0xffffffffu, pent, (s2 * N + min(t, N - 1)) & 31);
                    if (live && t < cnt2) {
                        if (t >= N) ent = a.push_ent[pb2 + t];
                        const int r = ent & 15;
                        if (r == 15 || r == erow) {
                            double2* dst = reinterpret_cast<double2*>(__ldg(a.peer + (ent >> 4))) + a.out_elem_off + gi;
                            *dst = outv;
                        }
                    }
                }
            }
        }
        __syncwarp();
    }
    // remote rows must have landed before the stream-ordered barrier that follows the kernel
    if (pushing) __threadfence_system();
}

