// heom_dataflow_tma.cuh - kernel 9: kernel 8's ADO-to-ADO synchronised propagation (all nt RK4
// steps of a small hierarchy with a large system matrix in ONE cooperative launch; BASELINE
// configs[3]: N = 32, 210 ADOs) for Hermitian problems, rebuilt around what bounded kernel 8
// (profiles/r02_kernel8_polariton32_ncu.txt: ~450 issued instructions per matrix element and
// stage, neighbour loads serialised link by link, 2 CTAs on 105 of the 148 SMs).
//
// Same equations (generate_dot_element, pyqed/heom/deom.py:641-664; accumulator form of rk4,
// deom.py:725-766).  What is different:
//  * one CTA per (trajectory, ADO) for the whole run: y stays in a register, the RK4 accumulator
//    and the own stage input in shared memory; global memory only carries the stage outputs that
//    the neighbours read.
//  * Hermitian form: with Hermitian operators, a real-exponent bath (eta_r = conj eta_l) and a
//    Hermitian state the link coefficients obey alphaR = conj(alphaL), so
//        d rho/dt = W + W^dagger,  W = (-iH - gamma/2) rho + sum_m Q_m S_m,  S_m = sum_{links of m} alphaL rho'
//    - half the products.
//  * a thread owns a UNIT = the element pair (i,j), (j,i) with i < j, or two diagonal elements:
//    W_ij and W_ji are formed by the same thread, so k_ij = W_ij + conj W_ji needs no transpose,
//    and one double2 per unit (rho_ij, or the two real diagonal entries) is the whole state.  The
//    stage outputs are published in this packed unit order - N(N-1)/2 + ceil(N/2) values, 8 KB
//    instead of 16 KB for N = 32 - and a reader's thread u needs exactly value u of each neighbour.
//  * neighbours: after its share of the own term, warp w (< 8) owns link w: its lane 0 waits for that
//    neighbour's flag (relaxed polls, one acquire), then the warp copies the neighbour's packed matrix
//    with cp.async (LDGSTS.128) into staging slot w - every link is fetched as soon as ITS flag is
//    there, and only one CTA barrier stands between the last arrival and the link sums, in which
//    thread u reads value u of every slot.  (Measured on the way: one cp.async.bulk per link cost
//    ~800 cycles per copy - UBLKCP takes uniform operands, so per-lane copies become a serial loop,
//    each with its generic->async proxy fence; every thread fetching its own values after a
//    poll-all-flags + barrier was 5 % slower than the per-warp scheme.)
//  * S_m of both elements of a unit from four real sums over the links (4 DFMA per link, the same
//    loop for pair and diagonal units); operators as zero-padded, entry-major sparse rows in shared
//    memory (values pre-multiplied, uniform trip count, no row pointers); links sorted by coupling
//    mode once; one flag per 128-byte line.
//  * placement: CTAs are numbered per SM after a one-off grid barrier; SMs that host one CTA take
//    the ADOs with the most links, shared SMs the lighter ones - with 210 ADOs on 148 SMs the
//    8-link ADOs run alone on their SM.
// The kernel name keeps "tma" from its first version; the neighbour copies are plain cp.async now.
// Every wait is bounded (globaltimer); a run that times out poisons Y with NaN instead of
// hanging the GPU.
#pragma once
#include "heom_dataflow.cuh"

constexpr int DF9_THREADS = 512;   // one unit per thread: N(N-1)/2 + ceil(N/2) <= 512 for N <= 32
constexpr int DF9_SLOTS = 8;       // neighbour matrices (packed) in flight
constexpr int DF9_CTRL = 4;        // control words in front of the per-SM counters
constexpr int DF9_MAXSM = 2048;
constexpr int DF9_FLAG_STRIDE = 32;   // words between two ADOs' flags: one 128-byte line each (many pollers per flag)
constexpr int DF9_NMAX = 32, DF9_NP = DF9_NMAX + 1;   // padded row stride of the full-matrix tiles
constexpr int DF9_MAXNNZ = 256;    // operator entries (H and all Q_m together)
constexpr int DF9_MAXL = 32;       // links per ADO
constexpr int DF9_MAXOPS = 8;      // 1 + coupling modes

// Shared memory of a CTA: every section at a compile-time offset (the kernel is register bound -
// with run-time section offsets ptxas re-derived them inside every loop)
struct __align__(128) Df9Smem {
    double2 slot[DF9_SLOTS][DF9_THREADS];   // packed neighbour matrices: slot w filled by warp w, value u read by thread u
    double2 rho[DF9_NMAX * DF9_NP];         // own stage input, full matrix
    double2 buf[DF9_NMAX * DF9_NP];         // S_m of a non-diagonal mode
    double2 acc[DF9_THREADS];               // RK4 accumulator (thread-private)
    double2 val[DF9_MAXNNZ];                // operator entries (-iH, Q_1, ...), padded rows: [obase[o] + k N + row] = k-th entry of the row
    double2 lcf[DF9_MAXL];                  // alphaL per link (links sorted by mode)
    int lnb[DF9_MAXL];                      // neighbour slot per link
    int mord[DF9_MAXOPS];                   // modes in processing order: non-diagonal Q_m first
    int mend[DF9_MAXOPS];                   // end of the k-th processed mode's links in the sorted list
    int obase[DF9_MAXOPS];                  // first entry of operator o in val / ri
    int mrow[DF9_MAXOPS];                   // entries per (padded) row of operator o
    int qdiag[DF9_MAXOPS];
    int misc[8];
    short ri[DF9_MAXNNZ];                   // column of an entry (padding: value 0, column = the row itself)
};

struct Dataflow9Args {
    StageArgs s;
    double2* Y;           // full matrices: read once at the start, written once at the end
    double2* P0;          // packed stage outputs [B * nmax][units]: the state (stage 3 -> stage 0)
    double2* P1;          // ... stage 0 -> 1 and stage 2 -> 3
    double2* P2;          // ... stage 1 -> 2
    unsigned* flags;      // [B * nmax][DF9_FLAG_STRIDE] number of published stage outputs, zero on entry
    unsigned* ctrl;       // [0] arrivals, [1] abort, [DF9_CTRL + smid] CTAs per SM; zero on entry
    const int* order;     // [B * nmax] work items by descending link count
    double dt;
    long long nt;
    unsigned long long timeout_ns;
    int B, units;
};

__device__ __forceinline__ unsigned long long df9_timer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
    return t;
}

// spin until *f >= need; gives up (and raises the abort word) after timeout_ns
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// (relaxed polls - an acquire load invalidates L1 on every iteration - and ONE acquire fence by the caller)
__device__ __forceinline__ void df9_wait_flag(const unsigned* f, unsigned need, unsigned* ctrl,
                                              unsigned long long timeout_ns) {
    unsigned long long t0 = 0;
    unsigned spins = 0;
    while (ld_relaxed_u32(f) < need) {
        if ((++spins & 255u) == 0) {
            if (__ldcg(ctrl + 1)) break;
            const unsigned long long t = df9_timer();
            if (!t0) t0 = t;
            else if (t - t0 > timeout_ns) {
                atomicExch(ctrl + 1, 1u);
                break;
            }
        }
    }
}

// wa += sum_k val[k][ia] src[col[k][ia]][ja] over the (padded) row ia of a sparse operator, the same for
// (wb, ib, jb).  Rows are padded to the operator's longest row with zero entries, stored entry-major
// ([k][row]): the trip count is uniform, lanes with consecutive rows read consecutive words, and there
// are no row pointers to load - the stage loop is bound by shared-memory wavefronts and issue slots.
__device__ __forceinline__ void df9_rows2(double2& wa, double2& wb, int n, int nrow, const double2* val, const short* ri,
                                          const double2* src, int ia, int ib, int ja, int jb) {
#pragma unroll 1
    for (int k = 0; k < n; ++k) {
        const int ea = k * nrow + ia, eb = k * nrow + ib;
        cfma(wa, val[ea], src[ri[ea] * DF9_NP + ja]);
        cfma(wb, val[eb], src[ri[eb] * DF9_NP + jb]);
    }
}

// items by descending number of links (ties by index): T <= a few hundred
__global__ void dataflow_order_kernel(const int* link_ptr, long long nmax, long long total, int* order) {
    for (long long it = threadIdx.x; it < total; it += blockDim.x) {
        const long long s = it % nmax;
        const int w = link_ptr[s + 1] - link_ptr[s];
        int rank = 0;
        for (long long o = 0; o < total; ++o) {
            const long long so = o % nmax;
            const int wo = link_ptr[so + 1] - link_ptr[so];
            rank += (wo > w) || (wo == w && o < it);
        }
        order[rank] = (int)it;
    }
}

// DENSE_H: H has (nearly) full rows - the N = 32 stress variant of config 4.  Its rows do not
// fit the operator table, and a row-per-unit product would read a different row in every lane.  Instead the
// own term W_own = -i H rho is formed as a plain matrix product: warp w owns rows w and w + 16, lane = column,
// so H[row][l] is one broadcast read for the whole warp and rho[l][:] one conflict-free row read; the product
// goes through the (still unused) S tile to the units' owners.  H (16 KB) is copied with cp.async into the
// last two staging slots at the end of every stage - they are free until the next stage's fetch - so the
// copy hides behind the publication.  32768 complex FMAs per ADO and stage.  (A first version kept -iH in
// the kernel's parameter space: 16 warps reading 32 different rows thrash the constant cache - 157 us/step.)
template <bool DENSE_H>
__global__ void __launch_bounds__(DF9_THREADS, 2) stage_dataflow_tma_kernel(const Dataflow9Args da) {
    extern __shared__ Df9Smem df9_smem[];   // typed: every access is a shared-space access at a constant offset
#define sm df9_smem[0]
    const StageArgs& a = da.s;
    const int N = a.N, NN = N * N, M1 = 1 + a.nmod, U = da.units;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- which ADO: per-SM numbering after a one-off grid barrier ----
    const long long total = a.nmax * (long long)da.B;
    if (tid == 0) {
        unsigned smid, nsm;
        asm volatile("mov.u32 %0, %%smid;\n" : "=r"(smid));
        asm volatile("mov.u32 %0, %%nsmid;\n" : "=r"(nsm));
        if (smid >= (unsigned)DF9_MAXSM) smid = DF9_MAXSM - 1;
        const unsigned r = atomicAdd(da.ctrl + DF9_CTRL + smid, 1u);
        __threadfence();
        atomicAdd(da.ctrl, 1u);
        df9_wait_flag(da.ctrl, gridDim.x, da.ctrl, da.timeout_ns);
        __threadfence();
        sm.misc[0] = (int)r;
        sm.misc[1] = (int)smid;
        sm.misc[2] = (int)min(nsm, (unsigned)DF9_MAXSM);
        sm.misc[3] = sm.misc[4] = sm.misc[5] = sm.misc[6] = 0;
    }
    __syncthreads();
    {
        // SMs that host one CTA take the heaviest items, first CTAs of shared SMs the next ones
        // (lightest of them on the lowest SM), second CTAs the rest: the ADOs with the most links
        // have an SM to themselves
        const int smid = sm.misc[1], nsm = sm.misc[2];
        int nfirst = 0, lowsingle = 0, lowshared = 0, lowextra = 0;
        for (int s = tid; s < nsm; s += DF9_THREADS) {
            const int c = (int)__ldcg(da.ctrl + DF9_CTRL + s);
            nfirst += c > 0;
            if (s < smid) {
                lowsingle += c == 1;
                lowshared += c > 1;
                lowextra += c > 1 ? c - 1 : 0;
            }
        }
        if (nfirst) atomicAdd(sm.misc + 3, nfirst);
        if (lowsingle) atomicAdd(sm.misc + 4, lowsingle);
        if (lowextra) atomicAdd(sm.misc + 5, lowextra);
        if (lowshared) atomicAdd(sm.misc + 6, lowshared);
    }
    __syncthreads();
    const int mycount = (int)__ldcg(da.ctrl + DF9_CTRL + sm.misc[1]);
    const long long widx = sm.misc[0] > 0 ? (long long)sm.misc[3] + sm.misc[5] + (sm.misc[0] - 1)
                                          : (mycount == 1 ? (long long)sm.misc[4] : (long long)sm.misc[3] - 1 - sm.misc[6]);
    __syncthreads();   // misc[6] is reused below
    if (widx >= total || __ldcg(da.ctrl + 1)) return;   // spare CTA (or the barrier timed out)
    const long long item = da.order[widx];
    const int b = (int)(item / a.nmax);
    const long long slot = item - (long long)b * a.nmax;
    if (tid == 0) {   // rarely used values stay out of the registers
        sm.misc[6] = (int)item;
        sm.misc[7] = (a.traj && slot == a.slot0) ? 1 : 0;
    }

    // ---- operators: padded sparse rows, entry-major, H pre-multiplied by -i ----
    if (tid < M1) {
        int longest = 0, diag = 1;
        for (int i = 0; i < N; ++i) {
            const int t0 = a.row_ptr[tid * (N + 1) + i], t1 = a.row_ptr[tid * (N + 1) + i + 1];
            longest = max(longest, t1 - t0);
            for (int t = t0; t < t1; ++t)
                if (a.row_idx[tid * NN + t] != i) diag = 0;
        }
        sm.mrow[tid] = (DENSE_H && tid == 0) ? 0 : longest;   // (dense H: not in the table)
        sm.qdiag[tid] = diag;
    }
    __syncthreads();
    if (tid == 0) {
        int base = 0;
        for (int o = 0; o < M1; ++o) {
            sm.obase[o] = base;
            base += sm.mrow[o] * N;
        }
    }
    __syncthreads();
    for (int r = tid; r < M1 * N; r += DF9_THREADS) {
        const int o = r / N, i = r - o * N;
        const int t0 = a.row_ptr[o * (N + 1) + i], len = a.row_ptr[o * (N + 1) + i + 1] - t0;
        for (int k = 0; k < sm.mrow[o]; ++k) {
            const int e = sm.obase[o] + k * N + i;
            if (k < len) {
                const int l = a.row_idx[o * NN + t0 + k];
                const double2 x = a.ops[o * NN + i * N + l];
                sm.ri[e] = (short)l;
                sm.val[e] = o == 0 ? make_double2(x.y, -x.x) : x;
            } else {
                sm.ri[e] = (short)i;
                sm.val[e] = make_double2(0.0, 0.0);
            }
        }
    }
    // ---- links of this ADO, sorted by coupling mode ----
    const int lbeg = a.link_ptr[slot], nl = a.link_ptr[slot + 1] - lbeg;
    int* ltmp = reinterpret_cast<int*>(&sm.slot[0][0]);   // unsorted (slot, meta); the slots are not in use yet
    for (int t = tid; t < nl; t += DF9_THREADS) {
        const int2 lk = __ldg(a.links + lbeg + t);
        ltmp[2 * t] = lk.x;
        ltmp[2 * t + 1] = lk.y;
    }
    __syncthreads();
    // modes with a non-diagonal Q_m first: their S_m goes through shared memory, and the barrier of
    // that exchange is covered by the link sums of the diagonal modes that follow
    if (tid == 0) {
        int k = 0;
        for (int pass = 0; pass < 2; ++pass)
            for (int m = 0; m < a.nmod; ++m)
                if ((sm.qdiag[1 + m] != 0) == (pass == 1)) sm.mord[k++] = m;
    }
    __syncthreads();
    auto mrank = [&](int m) {
        int r = 0;
        for (int k = 0; k < a.nmod; ++k)
            if (sm.mord[k] == m) r = k;
        return r;
    };
    for (int t = tid; t < nl; t += DF9_THREADS) {
        const int meta = ltmp[2 * t + 1], m = mrank(heom::meta_mode(meta));
        int pos = 0;
        for (int q = 0; q < nl; ++q) {
            const int mq = mrank(heom::meta_mode(ltmp[2 * q + 1]));
            pos += (mq < m) || (mq == m && q < t);
        }
        sm.lnb[pos] = ltmp[2 * t];
        sm.lcf[pos] = a.coef[2 * heom::meta_ci(meta, a.nind, a.lmax)];
    }
    for (int k = tid; k < a.nmod; k += DF9_THREADS) {
        int e = 0;
        for (int q = 0; q < nl; ++q) e += mrank(heom::meta_mode(ltmp[2 * q + 1])) <= k;
        sm.mend[k] = e;
    }
    // ---- this thread's unit: elements A = (ia, ja) and B = (ib, jb) ----
    //   u <  N(N-1)/2 : the pair i < j (row-major over the strict upper triangle): A = (i, j), B = (j, i),
    //                   value rho_ij
    //   then          : two diagonal elements d = 2 (u - N(N-1)/2), d + 1: value (rho_dd, rho_d+1,d+1)
    const int uoff = N * (N - 1) / 2;
    const bool valid = tid < U, isdiag = tid >= uoff;
    int ia = 0, ja = 0, ib = 0, jb = 0;
    bool hasb = false;
    if (valid) {
        if (!isdiag) {
            int i = 0, rem = tid;
            while (rem >= N - 1 - i) {
                rem -= N - 1 - i;
                ++i;
            }
            ia = i;
            ja = i + 1 + rem;
            ib = ja;
            jb = ia;
            hasb = true;
        } else {
            ia = ja = 2 * (tid - uoff);
            hasb = ia + 1 < N;
            ib = jb = hasb ? ia + 1 : ia;
        }
    }
    // element values of a unit value x
    auto elem_a = [&](const double2 x) { return isdiag ? make_double2(x.x, 0.0) : x; };
    auto elem_b = [&](const double2 x) { return isdiag ? make_double2(x.y, 0.0) : make_double2(x.x, -x.y); };
    const double hg = -0.5 * a.damp[slot].x;   // W carries half of the (real) damping
    // y lives in a register for the whole run, the accumulator in a thread-private shared-memory word
    double2 y = make_double2(0.0, 0.0);
    const long long pbase = (long long)b * a.nmax * U;   // this trajectory's packed matrices
    const long long myoff = pbase + slot * U + tid;      // this thread's value in a packed array
    __syncthreads();   // ltmp (in the slots) is consumed
    if (valid) {
        const double2* Yf = da.Y + item * NN;
        const double2 xa = Yf[ia * N + ja];
        y = isdiag ? make_double2(xa.x, hasb ? Yf[ib * N + jb].x : 0.0) : xa;
        sm.rho[ia * DF9_NP + ja] = elem_a(y);
        if (hasb) sm.rho[ib * DF9_NP + jb] = elem_b(y);
        da.P0[myoff] = y;
    }
    __syncthreads();
    // Publication: CTA barrier, then ONE release store by thread 0.  (Measured alternative: every
    // warp releasing its own part with red.release - 16 fences per CTA - is 5 % slower.)
    unsigned* const myflag = da.flags + (long long)sm.misc[6] * DF9_FLAG_STRIDE;
    if (tid == 0) st_release_u32(myflag, 1u);   // the packed initial state is published

    // DENSE_H: H lives in the last two staging slots between the end of a stage and the next stage's fetch
    double2* const Hs = &sm.slot[DF9_SLOTS - 2][0];
    auto prefetch_h = [&]() {
        for (int e = tid; e < NN; e += DF9_THREADS) cp_async16(Hs + e, a.ops + e);
        cp_async_commit();
    };
    if (DENSE_H) prefetch_h();
#ifdef DF9_PROFILE
    long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pt = clock64();
#define DF9_MARK(k) do { if (tid == 0) { const long long t_ = clock64(); pf[k] += t_ - pt; pt = t_; } } while (0)
#else
#define DF9_MARK(k)
#endif
    const int nt = (int)da.nt;
#pragma unroll 1
    for (int step = 0; step < nt; ++step) {
#pragma unroll 1
        for (int stage = 0; stage < 4; ++stage) {
            const unsigned need = 4u * (unsigned)step + (unsigned)stage + 1u;   // outputs the neighbours must have published
            const double2* yin = (stage == 0 ? da.P0 : (stage == 2 ? da.P2 : da.P1)) + pbase;
            // warp w (< 8) owns link first + w of a batch: its lane 0 waits for that neighbour's flag, then the
            // warp copies the neighbour's packed matrix into staging slot w - every link is fetched as soon as
            // ITS flag is there, and only the CTA barrier stands between the last arrival and the sums
            auto fetch_batch = [&](int first) {
                const int lp = first + warp;
                if (warp < DF9_SLOTS && lp < nl) {
                    const int nb = sm.lnb[lp];
                    if (lane == 0) {
                        const unsigned* f = da.flags + ((sm.misc[6] / a.nmax) * a.nmax + nb) * DF9_FLAG_STRIDE;
                        df9_wait_flag(f, need, da.ctrl, da.timeout_ns);
                        (void)ld_acquire_u32(f);   // the polls were relaxed: one acquire load of the final value
                    }
                    __syncwarp();
                    const double2* src = yin + (long long)nb * U;
#pragma unroll 4
                    for (int c = lane; c < U; c += 32) cp_async16(&sm.slot[warp][c], src + c);
                    cp_async_commit();
                    cp_async_wait<0>();
                }
            };
            // (A) own term: W = (-iH - gamma/2) rho for both elements of the unit (needs no neighbour)
            double2 wa = make_double2(0.0, 0.0), wb = wa;
            if (DENSE_H) {
                cp_async_wait<0>();
                __syncthreads();   // H has arrived in the last two slots
                // rows `warp` and `warp + 16` of H rho, column `lane`; -i H rho goes into the S tile
                const int r0 = warp, r1 = warp + DF9_THREADS / 32;
                if (lane < N && r0 < N) {
                    double2 t0 = make_double2(0.0, 0.0), t1 = t0;
                    const bool two = r1 < N;
#pragma unroll 4
                    for (int l = 0; l < N; ++l) {
                        const double2 x = sm.rho[l * DF9_NP + lane];
                        cfma(t0, Hs[r0 * N + l], x);
                        if (two) cfma(t1, Hs[r1 * N + l], x);
                    }
                    sm.buf[r0 * DF9_NP + lane] = make_double2(t0.y, -t0.x);
                    if (two) sm.buf[r1 * DF9_NP + lane] = make_double2(t1.y, -t1.x);
                }
                __syncthreads();   // every warp is done with H before links 6 and 7 are fetched over it
            } else if (valid) {
                const double2 oa = sm.rho[ia * DF9_NP + ja], ob = sm.rho[ib * DF9_NP + jb];
                wa = make_double2(hg * oa.x, hg * oa.y);
                wb = make_double2(hg * ob.x, hg * ob.y);
                df9_rows2(wa, wb, sm.mrow[0], N, sm.val, sm.ri, sm.rho, ia, ib, ja, jb);
            }
            DF9_MARK(0);
            // (B) the neighbours' stage outputs (first batch), after this warp's share of the own term
            fetch_batch(0);
            __syncthreads();   // the staging slots are filled; rho may be overwritten by the epilogue from here on
            if (DENSE_H) {
                if (valid) {   // the units pick up their two elements of A rho
                    const double2 oa = sm.rho[ia * DF9_NP + ja], ob = sm.rho[ib * DF9_NP + jb];
                    const double2 pa = sm.buf[ia * DF9_NP + ja], pb = sm.buf[ib * DF9_NP + jb];
                    wa = make_double2(fma(hg, oa.x, pa.x), fma(hg, oa.y, pa.y));
                    wb = make_double2(fma(hg, ob.x, pb.x), fma(hg, ob.y, pb.y));
                }
                __syncthreads();   // the S tile is free for the coupling modes; rho for the epilogue
            }
            DF9_MARK(1);
            // (C) coupling terms, mode by mode: S_m summed link by link from the staging slots, then Q_m S_m
            int lp = 0;
            int pend = -1;   // non-diagonal mode whose S is in buf, products not taken yet (the readers of the
                             // previous stage are behind the end-of-stage barrier)
            // S_A = sum alphaL x_A, S_B = sum alphaL x_B from four real sums (x = xr + i xi, alphaL = cr + i ci):
            // A1 = sum cr xr, A2 = sum ci xi, A3 = sum cr xi, A4 = sum ci xr; pair units (x_A = x, x_B = conj x):
            // S_A = (A1 - A2, A3 + A4), S_B = (A1 + A2, A4 - A3); diagonal units (x_A = xr, x_B = xi real):
            // S_A = (A1, A4), S_B = (A3, A2) - the loop is the same for both, four independent chains
            double A1 = 0.0, A2 = 0.0, A3 = 0.0, A4 = 0.0;
            int k = 0;
#pragma unroll 1
            while (k < a.nmod) {
                const int lend = sm.mend[k];
                if (lp < lend) {
                    if ((lp & (DF9_SLOTS - 1)) == 0 && lp > 0) {   // more than eight links: next batch
                        __syncthreads();   // the slots are consumed
                        fetch_batch(lp);
                        __syncthreads();
                        DF9_MARK(2);
                    }
                    const int stop = min(lend, (lp | (DF9_SLOTS - 1)) + 1);   // this mode's links inside the batch
#pragma unroll 1
                    for (; lp + 2 <= stop; lp += 2) {   // two links at a time: the loads of both before the sums
                        const int q = lp & (DF9_SLOTS - 1);
                        const double2 c0 = sm.lcf[lp], c1 = sm.lcf[lp + 1];
                        const double2 x0 = sm.slot[q][tid], x1 = sm.slot[q + 1][tid];
                        A1 = fma(c0.x, x0.x, A1);
                        A2 = fma(c0.y, x0.y, A2);
                        A3 = fma(c0.x, x0.y, A3);
                        A4 = fma(c0.y, x0.x, A4);
                        A1 = fma(c1.x, x1.x, A1);
                        A2 = fma(c1.y, x1.y, A2);
                        A3 = fma(c1.x, x1.y, A3);
                        A4 = fma(c1.y, x1.x, A4);
                    }
                    if (lp < stop) {
                        const double2 c0 = sm.lcf[lp];
                        const double2 x0 = sm.slot[lp & (DF9_SLOTS - 1)][tid];
                        A1 = fma(c0.x, x0.x, A1);
                        A2 = fma(c0.y, x0.y, A2);
                        A3 = fma(c0.x, x0.y, A3);
                        A4 = fma(c0.y, x0.x, A4);
                        ++lp;
                    }
                    if (lp < lend) continue;   // the mode goes on in the next batch
                    DF9_MARK(3);
                    // the mode is complete: Q_m S_m
                    const int m = sm.mord[k];
                    const double2 Sa = isdiag ? make_double2(A1, A4) : make_double2(A1 - A2, A3 + A4);
                    const double2 Sb = isdiag ? make_double2(A3, A2) : make_double2(A1 + A2, A4 - A3);
                    A1 = A2 = A3 = A4 = 0.0;
                    if (sm.qdiag[1 + m]) {   // at most one entry per row, on the diagonal
                        if (valid && sm.mrow[1 + m]) {
                            const int ob = sm.obase[1 + m];
                            cfma(wa, sm.val[ob + ia], Sa);
                            cfma(wb, sm.val[ob + ib], Sb);
                        }
                    } else {
                        if (pend >= 0) {   // finish the previous exchange before buf is overwritten
                            __syncthreads();
                            if (valid)
                                df9_rows2(wa, wb, sm.mrow[1 + pend], N, sm.val + sm.obase[1 + pend], sm.ri + sm.obase[1 + pend],
                                          sm.buf, ia, ib, ja, jb);
                            __syncthreads();
                        }
                        if (valid) {
                            sm.buf[ia * DF9_NP + ja] = Sa;
                            if (hasb) sm.buf[ib * DF9_NP + jb] = Sb;
                        }
                        pend = m;
                    }
                    DF9_MARK(4);
                }
                ++k;
            }
            if (pend >= 0) {
                __syncthreads();
                if (valid)
                    df9_rows2(wa, wb, sm.mrow[1 + pend], N, sm.val + sm.obase[1 + pend], sm.ri + sm.obase[1 + pend], sm.buf, ia,
                              ib, ja, jb);
            }
            DF9_MARK(4);
            // (D) k = W + W^dagger inside the thread, (E) stage update; the output goes to shared
            // memory (own next input, full matrix) and packed to global memory (the neighbours')
            if (valid) {
                const double2 k = isdiag ? make_double2(2.0 * wa.x, 2.0 * wb.x) : make_double2(wa.x + wb.x, wa.y - wb.y);
                const double ca = stage == 2 ? da.dt : 0.5 * da.dt;
                const double cw = (stage == 0 || stage == 3) ? da.dt / 6.0 : da.dt / 3.0;
                const double2 base = stage == 0 ? y : sm.acc[tid];
                const double2 nacc = make_double2(fma(cw, k.x, base.x), fma(cw, k.y, base.y));
                double2 out;
                if (stage < 3) {
                    sm.acc[tid] = nacc;
                    out = make_double2(fma(ca, k.x, y.x), fma(ca, k.y, y.y));
                } else {
                    out = y = nacc;
                    if (sm.misc[7]) {
                        double2* tr = a.traj + (sm.misc[6] / a.nmax) * a.traj_bstride + (long long)(step + 1) * NN;
                        tr[ia * N + ja] = elem_a(out);
                        if (hasb) tr[ib * N + jb] = elem_b(out);
                    }
                }
                double2* yout = stage == 1 ? da.P2 : (stage == 3 ? da.P0 : da.P1);
                yout[myoff] = out;
                sm.rho[ia * DF9_NP + ja] = elem_a(out);
                if (hasb) sm.rho[ib * DF9_NP + jb] = elem_b(out);
            }
            DF9_MARK(5);
            __syncthreads();   // every value of this stage's output is written ...
            DF9_MARK(6);
            if (DENSE_H) prefetch_h();   // the staging slots are free until the next stage's fetch
            if (tid == 0) st_release_u32(myflag, need + 1u);   // ... and published (release: bar.sync + cumulativity)
            DF9_MARK(7);
        }
    }
    // the evolved state as full matrices
    if (valid) {
        const bool bad = __ldcg(da.ctrl + 1) != 0;   // a wait timed out somewhere: make the failure loud
        const double qnan = __longlong_as_double(0x7ff8000000000000ll);
        double2* Yf = da.Y + (long long)sm.misc[6] * NN;
        Yf[ia * N + ja] = bad ? make_double2(qnan, qnan) : elem_a(y);
        if (hasb) Yf[ib * N + jb] = bad ? make_double2(qnan, qnan) : elem_b(y);
    }
#ifdef DF9_PROFILE
    if (tid == 0 && (widx == 0 || widx == 100 || widx == 147 || widx == 148 || widx == total - 1)) {
        const double sc = 1.0 / (4.0 * da.nt);
        printf("df9 widx %lld slot %lld nl %d smid %d r %d cycles/stage: poll %.0f own+sync %.0f fetch %.0f consume %.0f "
               "Q.S %.0f epilogue %.0f sync %.0f publish %.0f\n",
               widx, slot, nl, sm.misc[1], sm.misc[0], pf[0] * sc, pf[1] * sc, pf[2] * sc, pf[3] * sc, pf[4] * sc, pf[5] * sc,
               pf[6] * sc, pf[7] * sc);
    }
#endif
}
#undef sm
