// heom_stage_sym.cu - kernel 6: Hermitian-symmetric RK4 stage kernel for N <= 8
// with projector-like (diagonal, one non-zero entry) coupling operators.
//
// Evaluates, for every owned ADO n (generate_dot_element, pyqed/heom/deom.py:641-664),
//     k_n = -(sum_k n_k gamma_k) rho_n - i[H, rho_n]
//           - i sum_k sqrt(n_k)/sqrt(a_k) (eta_k Q_m rho_{n-e_k} - conj(eta_k) rho_{n-e_k} Q_m)
//           - i sum_k sqrt(n_k+1) sqrt(a_k) [Q_m, rho_{n+e_k}]
// and applies the stage update of the difference-form RK4 (rk4, deom.py:725-766;
// DESIGN.md section 4) in the same pass.  heom_stage_sym.cuh says what differs
// from kernel 3.
//
// Work split: persistent CTAs (one per SM); a warp owns APW = 32/N consecutive
// ADOs ("group"), lane = (ADO sub, matrix row).  Per group:
//   1. the group's link records (prefetched one group ahead into registers by a
//      coalesced load) are published in the warp's shared-memory strip;
//   2. own tile by one bulk copy (odd N) or cp.async, first chunk of neighbour
//      rows by cp.async (16N bytes per link), y / stage buffers by bulk copies;
//   3. column `row` of P = -i H rho - (damp/2) rho into the k tile;
//   4. neighbour terms: (Q rho' - rho' Q) touches row r0 and column r0 only; the
//      lane's element of the row update is summed in registers per target row and
//      added into the lane's own column of the tile;
//   5. k = P' + P'^dagger (each lane finishes (N-1)/2 element pairs and its
//      diagonal element) - this supplies both the rho H half of the commutator and
//      the links' column updates;
//   6. epilogue from shared memory with streaming stores.
#ifndef HEOM_HOST_EMU
#include <cuda_runtime.h>
#endif

#include <algorithm>
#include <type_traits>

#include "heom_core.cuh"
#include "heom_stage_sym.cuh"

#ifndef HEOM_SYM_THREADS
#define HEOM_SYM_THREADS 512       // first / middle stage: 16 warps, 128 registers
#endif
#ifndef HEOM_SYM_LAST_THREADS
#define HEOM_SYM_LAST_THREADS 448  // last stage stages one more tile per warp: at most 14 warps, 144 registers
#endif

namespace {

// L2 eviction-priority hints (build-time, pyqed_b200/build.py defines=...): 1 = the arrays that are
// only streamed (y, stage buffers of the last stage) are fetched evict_first, so that L2 is left
// to the stage input, the only array with reuse (own tile + neighbour rows); 2 = the stage input
// is fetched evict_last as well
#ifndef HEOM_SYM_L2HINT
#define HEOM_SYM_L2HINT 0
#endif

constexpr int SYM_NCH = 4;   // link chunks (of N links) whose records are prefetched per ADO

constexpr int SYM_TOFS = 20;  // double2 units reserved for the packed-row offset table ((N+1)*N ints)

// per-warp shared memory in double2 units.  `packed`: the ADO arrays hold the upper
// triangle only (N(N+1)/2 elements per ADO, row-major), see stage_rows_sym_kernel.
// `db`: the streamed tiles (own tile, y, first stage buffer) are double-buffered and fetched
// one group ahead.
// Shared-memory layouts chosen for the banks (a 128-bit access is served per quarter warp, eight
// 16-byte slots wide): with lane = sub N + row,
//   k tile:         element (i, j) of ADO `sub` at sub KSUB + KLD i + j, KLD = 8, KSUB = 9N - 8 (= N mod 8):
//                   column accesses (fixed i), the diagonal, the (row, row+d) diagonals and their
//                   transposes all land in slot lane + const (mod 8) - no conflicts;
//   neighbour rows: element `row` of link t at sub NBSUB + t N + row, NBSUB = N mod 8.
// (The staged own tile keeps the layout of the global array, it arrives by bulk copy.)
// Packed storage (kernel 7) forms k = P' + P'^dagger in the epilogue, which reads (and, with the
// fused push, writes) element (j, i) next to (i, j) with consecutive lanes on consecutive j: there
// the row stride is 9, so that the transposed accesses spread over the slots as well.
__host__ __device__ constexpr int sym_kld(bool packed) { return packed ? 9 : 8; }
__host__ __device__ constexpr int sym_ksub(int N, bool packed) {
    const int base = (N - 1) * sym_kld(packed) + N;
    return base + ((N - base) % 8 + 8) % 8;
}
__host__ __device__ constexpr int sym_nbsub(int N) { return N * N + ((N - N * N) % 8 + 8) % 8; }

// fused push: per-warp staging area for the rows a group sends to other ranks (N elements each)
#ifndef HEOM_SYM_PUSH_SLOTS
#define HEOM_SYM_PUSH_SLOTS 12
#endif
constexpr int SYM_PUSH_SLOTS = HEOM_SYM_PUSH_SLOTS;

__host__ __device__ constexpr int sym_perwarp(int N, int stage, bool packed, bool db = false, bool push = false) {
    const int APW = 32 / N, LD = (N % 2 == 0) ? N + 1 : N, TILE_R = APW * N * LD;
    const int TILE = APW * sym_ksub(N, packed), FLAT = APW * sym_nbsub(N);
    const int FE = APW * (packed ? N * (N + 1) / 2 : N * N);   // a group's elements in the global arrays
    const int RT = packed ? FE : TILE_R;                     // own tile as staged
    const int nb = db ? 2 : 1;
    // own tile, k tile, neighbour rows, [y], [first stage buffer], record strip, mbarriers
    return nb * RT + TILE + FLAT + (stage == 0 ? 0 : nb * FE) + (stage == 2 ? nb * FE : 0) +
           SYM_NCH * APW * N / 2 + 3 + (push ? SYM_PUSH_SLOTS * sym_pool_stride(N) : 0);
}
__host__ __device__ constexpr int sym_max_threads(int stage) {
    return stage == 2 ? HEOM_SYM_LAST_THREADS : HEOM_SYM_THREADS;
}

// PACKED (kernel 7): every ADO is Hermitian, so the global arrays hold only the upper
// triangle, N(N+1)/2 elements per ADO in row-major order (element (i,j), i <= j, at
// i N - i(i-1)/2 + j - i) - 448 instead of 784 bytes for N = 7.  The k tile in shared memory
// stays a full N x N matrix; the own tile is read through the triangle (conjugating below
// the diagonal), a neighbour-row element (r0, j) comes from (min, max) of the pair, and the
// epilogue writes the upper triangle of the stage output.
// H as a kernel parameter (constant bank): complex entries, or the real parts only when H
// is real - two of those come with one 128-bit constant load
template <int N>
struct HParamReal {
    double v[N * N + (N * N & 1)];
};
template <int N, bool HREAL>
using HParamOf = typename std::conditional<HREAL, HParamReal<N>, HParam<N>>::type;

// DB: the bulk-copied tiles of a group (own tile, y, first stage buffer) live in two buffer
// sets; the copies for the next group are issued before the current group is processed, so
// they are in flight during the whole group instead of being waited for right after issue.
// PUSH: fused multi-GPU halo.  The epilogue also leaves the stage output in the k tile, and
// the rows other ranks need (push tables, as for kernel 3's fused push) go from there into
// the peers' arrays as bulk shared->global stores over NVLink - one instruction per row, in
// flight while the warp loads its next group.
template <int N, bool HREAL, int STAGE, bool PACKED, bool DB, bool PUSH>
__global__ void __launch_bounds__(sym_max_threads(STAGE), 1)
stage_rows_sym_kernel(const SymArgs a, const __grid_constant__ HParamOf<N, HREAL> hp) {
    static_assert(!PUSH || !DB, "the fused push has no double-buffered variant");
    constexpr bool FIRST = STAGE == 0, LAST = STAGE == 2;
    constexpr int NN = N * N, APW = 32 / N, LD = (N % 2 == 0) ? N + 1 : N, TILE_R = APW * N * LD;
    constexpr int KLD = sym_kld(PACKED), KSUB = sym_ksub(N, PACKED), NBSUB = sym_nbsub(N);
    constexpr int TILE = APW * KSUB, FLAT = APW * NBSUB;   // k tile, neighbour rows
    constexpr int PERWARP = sym_perwarp(N, STAGE, PACKED, DB, PUSH), NCH = SYM_NCH;
    constexpr int PSLOTS = SYM_PUSH_SLOTS, PS = sym_pool_stride(N);   // staging slots, pool row stride
    constexpr int NBUF = DB ? 2 : 1;
    constexpr int PK = N * (N + 1) / 2, EL = PACKED ? PK : NN;   // elements per ADO in the global arrays
    constexpr int FE = APW * EL, RT = PACKED ? FE : TILE_R;
    constexpr bool BULK_TILE = PACKED || (LD == N);   // padded tiles cannot be one bulk copy
    static_assert(!DB || BULK_TILE, "double buffering needs bulk-copied tiles");
    constexpr int EIT = (FE + 31) / 32;
    HEOM_DYN_SMEM(double2, smem);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int L1 = a.lmax + 1, ncq = 2 * a.nind * L1;
    // coefficient pairs per (2k+dir, n_eff): [0] element off the diagonal, [1] the diagonal
    // element (row == r0), both already times sqrt(n_eff)
    double2* cq_s = smem;
    int* tofs_s = (int*)(smem + 2 * ncq);   // PACKED: [r0][j] -> offset of element (r0, j) in the triangle
    // buffer set b of the streamed tiles: rho0 + b*RT, y0 + b*FE, acc0 + b*FE
    double2* const rho0 = smem + 2 * ncq + SYM_TOFS + wid * PERWARP;
    double2* k_s = rho0 + NBUF * RT;
    double2* nb_s = k_s + TILE;
    double2* const y0 = nb_s + FLAT;                            // !FIRST
    double2* const acc0 = y0 + (FIRST ? 0 : NBUF * FE);         // LAST
    int2* strip = (int2*)(acc0 + (LAST ? NBUF * FE : 0));
    unsigned long long* barA0 = (unsigned long long*)(strip + NCH * APW * N);   // own tile, per buffer set
    unsigned long long* barB0 = barA0 + NBUF;                                   // y / first stage buffer
    unsigned long long* barD = barA0 + 2 * NBUF;                                // second stage buffer
    double2* const push_s = rho0 + PERWARP - PSLOTS * PS;                       // PUSH: rows on their way to peers
    unsigned phA = 0, phB = 0, phD = 0;   // phA / phB: one phase bit per buffer set
    if (lane == 0) {
#pragma unroll
        for (int b = 0; b < NBUF; ++b) {
            mbar_init(barA0 + b, 1);
            mbar_init(barB0 + b, 1);
        }
        mbar_init(barD, 1);
        fence_proxy_async();
    }
    for (int e = threadIdx.x; e < ncq; e += blockDim.x) {
        const int kd = e / L1, n = e - kd * L1;
        const int k = kd >> 1, dir = kd & 1;
        const int m = a.kmode[k] & 0xff, r0 = a.kmode[k] >> 8;
        const double2 q = a.ops[(1 + m) * NN + r0 * N + r0];
        const double2 bL = a.cbase[4 * k + 2 * dir], bR = a.cbase[4 * k + 2 * dir + 1];
        const double2 c0 = cmul(bL, q);
        const double2 c1 = cmul(make_double2(bL.x + bR.x, bL.y + bR.y), q);
        const double sq = sqrt((double)n);
        cq_s[2 * e] = make_double2(c0.x * sq, c0.y * sq);
        // the diagonal element is counted twice by the P' + P'^dagger pass: halve it here
        cq_s[2 * e + 1] = make_double2(0.5 * (c1.x * sq), 0.5 * (c1.y * sq));
    }
    if (PACKED) {
        // [r0][j] -> offset of element (min(r0, j), max(r0, j)) in the triangle; row N: identity (pool rows)
        for (int e = threadIdx.x; e < NN + N; e += blockDim.x) {
            const int r0 = e / N, j = e - r0 * N, lo = min(r0, j), hi = max(r0, j);
            tofs_s[e] = r0 < N ? lo * N - lo * (lo - 1) / 2 + (hi - lo) : j;
        }
    }
    __syncthreads();

    const int sub = lane / N, row = lane - sub * N;
    const bool lane_ok = lane < APW * N;
    constexpr int L2H = HEOM_SYM_L2HINT;
    const unsigned long long pol_stream = L2H >= 1 ? l2_policy_evict_first() : 0ull;
    const unsigned long long pol_keep = L2H >= 2 ? l2_policy_evict_last() : 0ull;
    auto g2s_stream = [&](void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
        if (L2H >= 1) bulk_g2s_hint(dst, src, bytes, bar, pol_stream);
        else bulk_g2s(dst, src, bytes, bar);
    };
    auto g2s_yin = [&](void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
        if (L2H >= 2) bulk_g2s_hint(dst, src, bytes, bar, pol_keep);
        else bulk_g2s(dst, src, bytes, bar);
    };
    auto row_copy = [&](unsigned dst_u32, const void* src, bool pred) {
        if (L2H >= 2) cp_async16_s_if_hint(dst_u32, src, pred, pol_keep);
        else cp_async16_s_if(dst_u32, src, pred);
    };
    const long long step = (LAST && a.traj) ? (*a.step_base + a.local_step) : 0;
    // slots, groups and element offsets are 32-bit here (the host checks nmax N^2 < 2^32)
    const int ngroups = (int)a.ngroups, gstride = (int)gridDim.x * nwarps;
    const int slot_lo = (int)a.slot_lo, slot_hi = (int)a.slot_hi;
    // flat element e = lane + 32 it of the group in the global arrays  ->  offset in the
    // staged own tile (rofs) and in the k tile (kofs)
    int kofs[EIT], rofs_[PACKED || BULK_TILE ? 1 : EIT];
#pragma unroll
    for (int it = 0; it < EIT; ++it) {
        const int e = lane + 32 * it;
        const int s = e / EL;
        int i = 0, j = e - s * EL;
        if (PACKED) {
            while (i < N - 1 && j >= N - i) {   // row i of the triangle holds N - i elements
                j -= N - i;
                ++i;
            }
            j += i;
        } else {
            i = j / N;
            j -= i * N;
        }
        // bits 0-9: offset of (i, j) in the k tile, 10-19: offset of (j, i), 20-22: i, 23-25: j, 26-30: ADO
        kofs[it] = (s * KSUB + i * KLD + j) | ((s * KSUB + j * KLD + i) << 10) | (i << 20) | (j << 23) | (s << 26);
        if (!(PACKED || BULK_TILE)) rofs_[it] = (s * N + i) * LD + j;   // padded rows (even N)
    }
    // bulk-copied tiles (packed, or odd N) keep the flat layout of the global array
    auto rofs = [&](int it) { return (PACKED || BULK_TILE) ? lane + 32 * it : rofs_[it]; };
    double2* const ksub = k_s + sub * KSUB;     // this ADO's k tile
    const int frow = row * N - row * (row - 1) / 2 - row;   // PACKED: element (row, l), l >= row, at frow + l
    const int* const trow = tofs_s + row;                   // PACKED: + r0*N: offset of element (r0, row)
    const double2* const nbrow = nb_s + sub * NBSUB + row;   // + t*N: row element of staged link t
    const unsigned nbrow_u32 = smem_u32(nbrow);
    // + (c*APW*N + t): record t of chunk c (idle lanes stay inside the strip)
    const int2* const strip_sub = strip + (lane_ok ? sub * N : 0);
    // this lane's element of the neighbour row of link record r (links2): x is the element offset
    // of the row's storage - full matrices: the row itself; packed: the neighbour's triangle, the
    // lane's element sits at trow[table row * N] behind it (table row = r0, or N = identity for the
    // halo rows of a sharded run, which lie as N consecutive elements in the pool)
    auto row_src = [&](const int2 r) -> const double2* {
        const unsigned off = PACKED ? (unsigned)r.x + (unsigned)trow[((unsigned)r.y >> 28) * N]
                                    : (unsigned)r.x + (unsigned)row;
        return a.yin + off;
    };
    const char* const cq_row = (const char*)cq_s;

    // Bookkeeping pipeline carried in registers: first slot and link offsets of the
    // group two iterations ahead, damping rate and the first NCH*N link records of
    // the next group.
    // Visiting order: see stage_rows_async_kernel (rotation inside runs of 16 groups).
    // Dynamic schedule (a.sched): groups are handed out in storage order by a global counter, so
    // the groups in flight on the whole chip always form one short window of consecutive slots
    // whatever the speed of the individual warps.  With a static stride the CTAs drift apart by
    // several sweeps (a sweep of the whole grid moves ~40 MB through L2), and the temporal
    // locality of the lexicographic storage order - half of all links point less than 1024
    // slots away - never reaches L2.  The counter is not reset between launches: the host passes
    // its value at launch (a.sched_base); every warp draws exactly one index past the end.
    const bool dyn = a.sched != nullptr;
    const int gfull = (a.scramble && !dyn) ? (ngroups & ~15) : 0;
    auto group_base = [&](int gg) {   // first slot of group gg, -1 past the end
        const unsigned rot = ((unsigned)(gg >> 4) * 2654435761u) >> 28;
        const int gm = gg < gfull ? ((gg & ~15) | ((gg + (int)rot) & 15)) : gg;
        return gg < ngroups ? slot_lo + gm * APW : -1;
    };
    int glast = dyn ? -1 : (int)blockIdx.x * nwarps + wid - gstride;
    auto next_group = [&]() {
        if (dyn) {
            if (glast < ngroups) {
                unsigned v = 0;
                if (lane == 0) v = atomicAdd(a.sched, 1u) - a.sched_base;
                v = (unsigned)__shfl_sync(0xffffffffu, (int)v, 0);
                glast = (int)(v < (unsigned)ngroups ? v : (unsigned)ngroups);
            }
        } else {
            glast += gstride;
        }
        return glast;
    };
    int nx_lbeg = 0, nx_lend = 0, nn_lbeg = 0, nn_lend = 0;
    int nx_pb = 0, nx_pe = 0, nn_pb = 0, nn_pe = 0;   // PUSH: entry range of this lane's ADO
    int2 nx_rec[NCH];
    double nx_damp = 0.0;
    auto fetch_ptr = [&](int b0, int& lb, int& le, int& pb, int& pe) {
        const int slot = b0 + sub;
        lb = le = pb = pe = 0;
        if (b0 >= 0 && lane_ok && slot < slot_hi) {
            lb = a.link_ptr[slot];
            le = a.link_ptr[slot + 1];
            if (PUSH) {
                pb = a.push_ptr[slot - slot_lo];
                pe = a.push_ptr[slot - slot_lo + 1];
            }
        }
    };
    auto fetch_rec = [&](int b0, int lb, int le) {
        const int slot = b0 + sub;
        if (b0 >= 0 && lane_ok && slot < slot_hi) nx_damp = a.damp[slot].x;   // real: Hermitian problem
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            nx_rec[c] = make_int2(0, 0);
            if (lb + c * N + row < le) nx_rec[c] = __ldg(a.links2 + lb + c * N + row);
        }
    };
    int cur_base = group_base(next_group());
    int nx_base = group_base(next_group());
    fetch_ptr(cur_base, nx_lbeg, nx_lend, nx_pb, nx_pe);
    fetch_rec(cur_base, nx_lbeg, nx_lend);
    fetch_ptr(nx_base, nn_lbeg, nn_lend, nn_pb, nn_pe);
    // DB: bulk copies of the streamed tiles of the group that starts at slot b0 into buffer set b
    // (one lane; the warp has synchronised after its last generic-proxy access to that set)
    auto issue_streamed = [&](int b0, int b) {
        const unsigned ne = (unsigned)(min(APW, slot_hi - b0) * EL) * 16u;
        const unsigned gb = (unsigned)b0 * (unsigned)EL;
        fence_proxy_async();
        mbar_expect_tx(barA0 + b, ne);
        g2s_yin(rho0 + b * RT, a.yin + gb, ne, barA0 + b);
        if (!FIRST) {
            mbar_expect_tx(barB0 + b, ne * (LAST ? 2u : 1u));
            g2s_stream(y0 + b * FE, a.y + gb, ne, barB0 + b);
            if (LAST) g2s_stream(acc0 + b * FE, a.s1 + gb, ne, barB0 + b);
        }
    };
    int buf = 0;
    if (DB && lane == 0 && cur_base >= 0) issue_streamed(cur_base, 0);

    while (cur_base >= 0) {
        const int base = cur_base;
        double2* const rho_s = rho0 + buf * RT;
        double2* const y_s = y0 + buf * FE;
        double2* const acc_s = acc0 + buf * FE;
        unsigned long long* const barA = barA0 + buf;
        unsigned long long* const barB = barB0 + buf;
        const double2* const rsub = rho_s + sub * (PACKED ? PK : N * LD);
        const int cnt = min(APW, slot_hi - base);
        const int nelem = cnt * EL;
        const bool on = lane_ok && sub < cnt;
        const int lbeg = nx_lbeg, lend = nx_lend;
        const int nl = on ? (lend - lbeg) : 0;
        const int pb = nx_pb, pe = on ? nx_pe : nx_pb;
        // PUSH: this lane's first push entry is requested now and used after the epilogue
        int2 pent = make_int2(0, 0);
        if (PUSH && pb + row < pe) pent = __ldg(a.push_ent + pb + row);
        const double dh = 0.5 * nx_damp;
        const unsigned gbase = (unsigned)base * (unsigned)EL;
        // publish this group's records, then start the prefetch of the next group's
        if (lane_ok) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) strip[(c * APW + sub) * N + row] = nx_rec[c];
        }
        cur_base = nx_base;
        nx_lbeg = nn_lbeg;
        nx_lend = nn_lend;
        nx_pb = nn_pb;
        nx_pe = nn_pe;
        fetch_rec(cur_base, nx_lbeg, nx_lend);
        nx_base = group_base(next_group());
        fetch_ptr(nx_base, nn_lbeg, nn_lend, nn_pb, nn_pe);

        // ---- issue: own tile + first chunk of neighbour rows, y / first stage buffer
        if (DB) {
            // this group's tiles were issued one iteration ago; now the next group's
            if (lane == 0 && cur_base >= 0) issue_streamed(cur_base, buf ^ 1);
        } else if (BULK_TILE) {
            if (lane == 0) {
                fence_proxy_async();   // earlier generic-proxy reads of these buffers are done (warp sync)
                mbar_expect_tx(barA, nelem * 16u);
                g2s_yin(rho_s, a.yin + gbase, nelem * 16u, barA);
            }
        } else {
            const double2* src = a.yin + gbase + lane;
#pragma unroll
            for (int it = 0; it < EIT; ++it)
                if (lane + 32 * it < nelem) cp_async16(&rho_s[rofs(it)], src + 32 * it);
        }
        __syncwarp();   // the strip is visible to the whole warp
        int ry[N];      // coefficient offset | target row of the chunk's links
#pragma unroll
        for (int t = 0; t < N; ++t) {
            const int2 r = strip_sub[t];
            ry[t] = r.y;
            row_copy(nbrow_u32 + t * (N * 16), row_src(r), t < nl);
        }
        cp_async_commit();
        if (!FIRST && !DB) {
            // y always; in the last stage also the first stage buffer - the second one
            // follows into rho_s once the commutator has consumed the own tile
            if (lane == 0) {
                if (!BULK_TILE) fence_proxy_async();
                mbar_expect_tx(barB, nelem * 16u * (LAST ? 2u : 1u));
                g2s_stream(y_s, a.y + gbase, nelem * 16u, barB);
                if (LAST) g2s_stream(acc_s, a.s1 + gbase, nelem * 16u, barB);
            }
        }
        cp_async_wait<0>();
        if (BULK_TILE) {
            mbar_wait(barA, (phA >> buf) & 1u);
            phA ^= 1u << buf;
        }
        __syncwarp();

        // ---- P = -i H rho - (damp/2) rho, column `row` of this ADO.  The full
        //      k = P' + P'^dagger (P' = P + row updates of the links) is formed after the
        //      link loop: -i[H,rho] = -i H rho + (-i H rho)^dagger for Hermitian rho.
        //      Last stage: the stage input's own weight, w (k + (2/dt) y_in), is folded in too.
#define HEL(r_, c_) (hp.v[(r_) * N + (c_)])
        if (on) {
            double2 col[N];
#pragma unroll
            for (int l = 0; l < N; ++l) {
                if (PACKED) {
                    // element (l, row) of the Hermitian tile from its upper triangle
                    const int cl = l * N - l * (l - 1) / 2 - l;   // (l, j), j >= l, at cl + j
                    const double2 v = rsub[l <= row ? cl + row : frow + l];
                    col[l] = make_double2(v.x, l <= row ? v.y : -v.y);
                } else {
                    col[l] = rsub[l * LD + row];
                }
            }
            const double sh = LAST ? (0.5 * a.a - dh) : -dh;
#pragma unroll
            for (int rr = 0; rr < N; ++rr) {
                double2 c = make_double2(0.0, 0.0);
#pragma unroll
                for (int l = 0; l < N; ++l) {
                    if constexpr (HREAL) {
                        const double h = HEL(rr, l);
                        c.x = fma(h, col[l].x, c.x);
                        c.y = fma(h, col[l].y, c.y);
                    } else {
                        cfma(c, HEL(rr, l), col[l]);
                    }
                }
                ksub[rr * KLD + row] = make_double2(fma(sh, col[rr].x, c.y), fma(sh, col[rr].y, -c.x));
            }
        }
#undef HEL
        if (LAST) {
            __syncwarp();   // every lane has read its column of the own tile
            // rho_s is free now: fetch the second stage buffer into it for the epilogue
            if (BULK_TILE) {
                if (lane == 0) {
                    fence_proxy_async();
                    mbar_expect_tx(barD, nelem * 16u);
                    g2s_stream(rho_s, a.s2 + gbase, nelem * 16u, barD);
                }
            } else {
                const double2* sb = a.s2 + gbase + lane;
#pragma unroll
                for (int it = 0; it < EIT; ++it)
                    if (lane + 32 * it < nelem) cp_async16(&rho_s[rofs(it)], sb + 32 * it);
                cp_async_commit();
            }
        }

        // ---- neighbour terms, N links per chunk.  (Q rho' - rho' Q) touches row r0 and
        //      column r0 only; a lane adds its element of the row update into its own column
        //      of the tile (no other lane touches that column here) - the column update is
        //      the conjugate and comes from the P' + P'^dagger pass below.  Contributions to
        //      one target row are summed in registers (X: element (cur_rr, row)).
        const int maxl = __reduce_max_sync(0xffffffffu, nl);
        double2 X = make_double2(0.0, 0.0);
        int cur_rr = -1;
        auto flush = [&]() {
            double2* d1 = ksub + cur_rr * KLD + row;
            double2 v1 = *d1;
            v1.x += X.x;
            v1.y += X.y;
            *d1 = v1;
        };
        for (int c0 = 0, c = 0; c0 < maxl; c0 += N, ++c) {
            if (c0 > 0) {
                __syncwarp();  // every lane is done with the previous chunk's rows (and records)
                if ((c & (NCH - 1)) == 0) {
                    // more than NCH chunks (rare): refill the strip with the next NCH chunks
                    if (lane_ok) {
#pragma unroll
                        for (int cc = 0; cc < NCH; ++cc) {
                            int2 r = make_int2(0, 0);
                            if (lbeg + c0 + cc * N + row < lend) r = __ldg(a.links2 + lbeg + c0 + cc * N + row);
                            strip[(cc * APW + sub) * N + row] = r;
                        }
                    }
                    __syncwarp();
                }
                const int2* const recs = strip_sub + (c & (NCH - 1)) * (APW * N);
#pragma unroll
                for (int t = 0; t < N; ++t) {
                    const int2 r = recs[t];
                    ry[t] = r.y;
                    row_copy(nbrow_u32 + t * (N * 16), row_src(r), c0 + t < nl);
                }
                cp_async_commit();
                cp_async_wait<0>();
                __syncwarp();
            }
            if (c0 < nl) {
#pragma unroll
                for (int t = 0; t < N; ++t) {
                    if (c0 + t < nl) {
                        const int rr = ry[t] & 15;
                        double2 Aj = nbrow[t * N];
                        // read through the triangle as (row, r0): conjugate (packed pool rows are stored
                        // the way the triangle would deliver them, so the same rule holds)
                        if (PACKED && row < rr) Aj.y = -Aj.y;
                        if (rr != cur_rr) {
                            if (cur_rr >= 0) flush();
                            cur_rr = rr;
                            X = make_double2(0.0, 0.0);
                        }
                        const double2 cf =
                            *(const double2*)(cq_row + ((ry[t] & 0x0fffffe0) + (row == rr ? 16 : 0)));
                        cfma(X, cf, Aj);
                    }
                }
            }
        }
        if (cur_rr >= 0) flush();
        __syncwarp();
        // ---- k = P' + P'^dagger.  FUSE_HERM (packed storage): formed by the epilogue, which reads
        //      (i, j) and (j, i) of P' for each of the N(N+1)/2 elements it writes - no separate pass
        //      over the tile.  Full storage keeps the pass: this lane finishes its diagonal
        //      element and the pairs (row, row+d), d = 1..(N-1)/2 (and d = N/2 from the lower half
        //      when N is even).
        constexpr bool FUSE_HERM = PACKED;   // (full storage: the transposed reads of all N*N elements cost more than the pass)
        if (!FUSE_HERM && on) {
            {
                double2* pd = ksub + row * KLD + row;
                const double2 v = *pd;
                *pd = make_double2(v.x + v.x, 0.0);
            }
#pragma unroll
            for (int dd = 1; dd <= N / 2; ++dd) {
                if (2 * dd == N && row >= N / 2) continue;
                int j = row + dd;
                if (j >= N) j -= N;
                double2* pa = ksub + row * KLD + j;
                double2* pb = ksub + j * KLD + row;
                const double2 va = *pa, vb = *pb;
                const double2 sum = make_double2(va.x + vb.x, va.y - vb.y);
                *pa = sum;
                *pb = make_double2(sum.x, -sum.y);
            }
        }
        cp_async_wait<0>();
        if (!FIRST) {
            mbar_wait(barB, (phB >> buf) & 1u);
            phB ^= 1u << buf;
        }
        if (BULK_TILE && LAST) {
            mbar_wait(barD, phD);
            phD ^= 1u;
        }
        __syncwarp();

        // ---- epilogue from shared memory, streaming stores
        // PUSH: the rows of this group's output that other ranks read are collected in the warp's
        // staging area while the epilogue computes them, and leave from there as bulk stores.  The
        // push table gives every distinct (ADO, row) of the group a staging slot (bits 8-15 of an
        // entry's y; 255 = no slot left, element-wise path below); rowmap[ADO][row] = slot or 255.
        const bool any_push = PUSH && __reduce_max_sync(0xffffffffu, pe - pb) > 0;
        unsigned char* const rowmap = reinterpret_cast<unsigned char*>(strip);   // (the strip is free after the link loop)
        if (PUSH && any_push) {
            bulk_wait_read();   // the previous group's rows have left the staging area
            if (lane_ok) rowmap[sub * 8 + row] = 255;
            __syncwarp();
            if (pb + row < pe) rowmap[sub * 8 + (pent.y & 15)] = (unsigned char)(pent.y >> 8);
            for (int q = pb + row + N; q < pe; q += N) {   // (more than N entries per ADO: rare)
                const int y_ = a.push_ent[q].y;
                rowmap[sub * 8 + (y_ & 15)] = (unsigned char)(y_ >> 8);
            }
            __syncwarp();
        }
        auto stage_rows = [&](int kk, const double2 v) {
            const int i = (kk >> 20) & 7, j = (kk >> 23) & 7;
            const unsigned char* const rm = rowmap + ((kk >> 26) & 31) * 8;
            const int mi = rm[i];
            if (mi < PSLOTS) push_s[mi * PS + j] = v;
            if (PACKED && i != j) {
                const int mj = rm[j];
                if (mj < PSLOTS) push_s[mj * PS + i] = v;   // row j, position i < j: as read through the triangle
            }
        };
        auto get_k = [&](int kk) {
            double2 v = k_s[kk & 0x3ff];
            if (FUSE_HERM) {
                const double2 m = k_s[(kk >> 10) & 0x3ff];
                v = make_double2(v.x + m.x, v.y - m.y);
            }
            return v;
        };
        int e0 = -1;   // LAST: flat offset of ADO 0 (rho_sys) inside this group, if it is here
        if (LAST && a.traj) {
            const long long d0 = a.slot0 - base;
            if (d0 >= 0 && d0 < cnt) e0 = (int)d0 * EL;
        }
#pragma unroll
        for (int it = 0; it < EIT; ++it) {
            const int e = lane + 32 * it;
            if (e < nelem) {
                const double2 k = get_k(kofs[it]);
                if (LAST) {
                    // y' = -y/3 + S1/3 + 2 S2/3 + w (k4 + (2/dt) S3)   (S3's share is already in k)
                    const double2 y0 = y_s[e], s1 = acc_s[e], s2 = rho_s[rofs(it)];
                    const double third = 1.0 / 3.0;
                    double2 res = make_double2(fma(a.w, k.x, third * (s1.x - y0.x)),
                                               fma(a.w, k.y, third * (s1.y - y0.y)));
                    res.x = fma(2.0 * third, s2.x, res.x);
                    res.y = fma(2.0 * third, s2.y, res.y);
                    st_stream(a.out + (gbase + e), res);
                    if (PUSH && any_push) stage_rows(kofs[it], res);
                    if (e0 >= 0 && (unsigned)(e - e0) < (unsigned)EL) {
                        if (PACKED) {   // the trajectory holds full matrices
                            const int i = (kofs[it] >> 20) & 7, j = (kofs[it] >> 23) & 7;
                            a.traj[(step + 1) * NN + i * N + j] = res;
                            if (i != j) a.traj[(step + 1) * NN + j * N + i] = make_double2(res.x, -res.y);
                        } else {
                            a.traj[(step + 1) * NN + (e - e0)] = res;
                        }
                    }
                } else {
                    const double2 yv = FIRST ? rho_s[rofs(it)] : y_s[e];
                    const double2 res = make_double2(fma(a.a, k.x, yv.x), fma(a.a, k.y, yv.y));
                    st_stream(a.out + (gbase + e), res);
                    if (PUSH && any_push) stage_rows(kofs[it], res);
                }
            }
        }
        if (PUSH && any_push) {
            fence_proxy_async();   // the staging writes are ordered before the bulk stores
            __syncwarp();
            bool overflow = false;
            for (int q = pb + row; q < pe; q += N) {
                const int2 ent = q == pb + row ? pent : a.push_ent[q];
                const int slot = (ent.y >> 8) & 255;
                if (slot < PSLOTS) {
                    double2* dst = reinterpret_cast<double2*>(a.peer[(ent.y >> 4) & 15]) + a.out_elem_off +
                                   (size_t)(unsigned)ent.x * PS;
                    bulk_s2g(dst, push_s + slot * PS, PS * 16u);   // whole sectors (the pad element is never read)
                } else {
                    overflow = true;
                }
            }
            bulk_commit();
            // rare: rows that got no staging slot - the N lanes of the ADO store such a row element by
            // element, read back from the output this warp has just written
            if (__reduce_max_sync(0xffffffffu, overflow ? 1 : 0) > 0) {
                for (int q = pb; q < pe; ++q) {
                    const int2 ent = a.push_ent[q];
                    if (((ent.y >> 8) & 255) < PSLOTS) continue;
                    const int r = ent.y & 15, lo_ = min(r, row), hi_ = max(r, row);
                    const int off = PACKED ? lo_ * N - lo_ * (lo_ - 1) / 2 + (hi_ - lo_) : r * N + row;
                    const double2 v = __ldcg(a.out + (gbase + (unsigned)(sub * EL + off)));   // (packed: as read through the triangle)
                    reinterpret_cast<double2*>(a.peer[(ent.y >> 4) & 15])[a.out_elem_off + (size_t)(unsigned)ent.x * PS + row] = v;
                }
            }
        }
        __syncwarp();
        if (DB) buf ^= 1;
    }
    if (PUSH) {
        // remote rows must have landed before the stream-ordered barrier that follows the kernel
        bulk_wait_all();
        __threadfence_system();
    }
}

// links -> links2 (one thread per link) for full or packed storage
__global__ void sym_convert_links_kernel(const int2* links, int2* links2, long long nlinks, int N, int L, int packed) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= nlinks) return;
    const int2 r = links[i];
    const int r0 = heom::meta_r0(r.y);
    links2[i] = make_int2((int)sym_link_x((unsigned)r.x, r0, N, packed != 0),
                          sym_link_y(heom::meta_kdir(r.y), heom::meta_neff(r.y), L, r0, r0));
}

thread_local const char* g_sym_err = "";

size_t sym_table_bytes(int K, int L) { return sizeof(double2) * (2 * 2 * (size_t)K * (L + 1) + SYM_TOFS); }

// full [nmax][N][N] <-> upper triangle [nmax][N(N+1)/2] (one thread per triangle element)
__global__ void sym_pack_kernel(double2* tri, const double2* full, long long nmax, int N) {
    const int PK = N * (N + 1) / 2;
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= nmax * PK) return;
    const long long slot = e / PK;
    int i = 0, j = (int)(e - slot * PK);
    while (i < N - 1 && j >= N - i) {
        j -= N - i;
        ++i;
    }
    j += i;
    tri[e] = full[slot * N * N + i * N + j];
}
__global__ void sym_unpack_kernel(double2* full, const double2* tri, long long nmax, int N) {
    const int PK = N * (N + 1) / 2;
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= nmax * PK) return;
    const long long slot = e / PK;
    int i = 0, j = (int)(e - slot * PK);
    while (i < N - 1 && j >= N - i) {
        j -= N - i;
        ++i;
    }
    j += i;
    const double2 v = tri[e];
    full[slot * N * N + i * N + j] = v;
    if (i != j) full[slot * N * N + j * N + i] = make_double2(v.x, -v.y);
}
constexpr size_t SYM_SMEM_BUDGET = 227 * 1024;

template <int N, bool HREAL, int STAGE, bool PACKED, bool DB, bool PUSH>
int sym_launch_t(const SymLaunch& s) {
    constexpr int APW = 32 / N;
    SymArgs args = s.a;
    args.ngroups = (s.part_hi - s.part_lo + APW - 1) / APW;
    args.slot_lo = s.part_lo;
    args.slot_hi = s.part_hi;
    const size_t table_bytes = sym_table_bytes(s.K, s.L);
    const size_t per_warp = sizeof(double2) * sym_perwarp(N, STAGE, PACKED, DB, PUSH);
    if (table_bytes + per_warp > SYM_SMEM_BUDGET) {
        g_sym_err = "shared-memory tables too large for kernel 6";
        return 1;
    }
    const int maxw = (int)std::min<size_t>(sym_max_threads(STAGE) / 32, (SYM_SMEM_BUDGET - table_bytes) / per_warp);
    int warps = s.warps > 0 ? std::min(s.warps, maxw) : maxw;
    if (s.warps <= 0) {
        // small hierarchies: spread the groups over all SMs first
        const long long per_sm = (args.ngroups + s.sm_count - 1) / s.sm_count;
        warps = (int)std::max<long long>(1, std::min<long long>(maxw, per_sm));
    }
    const size_t smem = table_bytes + per_warp * warps;
    const long long ctas = (args.ngroups + warps - 1) / warps;
    const unsigned grid = (unsigned)std::min<long long>(ctas, s.sm_count);
    HParamOf<N, HREAL> hp{};
    for (int e = 0; e < N * N; ++e) {
        if constexpr (HREAL) hp.v[e] = s.H[2 * e];
        else hp.v[e] = make_double2(s.H[2 * e], s.H[2 * e + 1]);
    }
    auto kern = stage_rows_sym_kernel<N, HREAL, STAGE, PACKED, DB, PUSH>;
#ifndef HEOM_HOST_EMU
    // the opt-in to > 48 KB of dynamic shared memory is per device
    static unsigned long long attr_done = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(attr_done >> (dev & 63) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)SYM_SMEM_BUDGET);
        if (e != cudaSuccess) {
            g_sym_err = cudaGetErrorString(e);
            return 1;
        }
        attr_done |= 1ull << (dev & 63);
    }
#endif
    // dynamic schedule only where a warp processes several groups (otherwise the order is moot)
    const bool dyn = s.sched && s.sched_total && args.ngroups > 2ll * grid * warps;
    args.sched = dyn ? s.sched : nullptr;
    // one launch per trajectory of the batch, each with its own array / trajectory pointers
    for (int b = 0; b < s.B; ++b) {
        if (dyn) {
            args.sched_base = *s.sched_total;
            *s.sched_total += (unsigned)args.ngroups + grid * (unsigned)warps;
        }
        HEOM_LAUNCH(kern, grid, warps * 32, smem, s.stream, args, hp);
#ifndef HEOM_HOST_EMU
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            g_sym_err = cudaGetErrorString(e);
            return 1;
        }
#endif
        args.yin += s.batch_elems;
        if (args.y) args.y += s.batch_elems;
        if (args.s1) args.s1 += s.batch_elems;
        if (args.s2) args.s2 += s.batch_elems;
        args.out += s.batch_elems;
        args.out_elem_off += s.batch_elems;
        if (args.traj) args.traj += s.traj_bstride;
    }
    return 0;
}

template <int N, bool PACKED, bool DB, bool PUSH = false>
int sym_launch_p(const SymLaunch& s) {
    if (s.hreal) {
        switch (s.stage) {
            case 0: return sym_launch_t<N, true, 0, PACKED, DB, PUSH>(s);
            case 1: return sym_launch_t<N, true, 1, PACKED, DB, PUSH>(s);
            default: return sym_launch_t<N, true, 2, PACKED, DB, PUSH>(s);
        }
    }
    switch (s.stage) {
        case 0: return sym_launch_t<N, false, 0, PACKED, DB, PUSH>(s);
        case 1: return sym_launch_t<N, false, 1, PACKED, DB, PUSH>(s);
        default: return sym_launch_t<N, false, 2, PACKED, DB, PUSH>(s);
    }
}
template <int N>
int sym_launch_n(const SymLaunch& s) {
    // the double-buffered variant exists for packed storage only (its tiles are small enough
    // to keep nearly all warps)
    // sharded runs: the epilogue stores the halo rows into the peers' row pools
    if (s.push) return s.packed ? sym_launch_p<N, true, false, true>(s) : sym_launch_p<N, false, false, true>(s);
    if (s.packed) return s.prefetch ? sym_launch_p<N, true, true>(s) : sym_launch_p<N, true, false>(s);
    return sym_launch_p<N, false, false>(s);
}

}  // namespace

// One object file per system size (build.py: -DHEOM_SYM_INST_N=n): heom_sym_launch_n<n> is the
// only thing such an object defines.  Without the macro (the common object, and the
// single-translation-unit build of the CPU emulation tests) everything else is compiled.
#if defined(HEOM_SYM_INST_N)
#define HEOM_SYM_CAT2(a, b) a##b
#define HEOM_SYM_CAT(a, b) HEOM_SYM_CAT2(a, b)
int HEOM_SYM_CAT(heom_sym_launch_n, HEOM_SYM_INST_N)(const SymLaunch& s, const char** err) {
    g_sym_err = "";
    const int rc = sym_launch_n<HEOM_SYM_INST_N>(s);
    if (rc && err) *err = g_sym_err;
    return rc;
}
#else
int heom_sym_push_slots(void) { return SYM_PUSH_SLOTS; }

int heom_sym_supported(int N, int K, int M, int L, const char** err) {
    (void)M;
    const char* why = nullptr;
    if (N < 2 || N > 8) why = "kernel 6 needs 2 <= N <= 8";
    else if (sym_table_bytes(K, L) + sizeof(double2) * sym_perwarp(N, 2, false) > SYM_SMEM_BUDGET)
        why = "shared-memory tables too large for kernel 6";
    else if ((((long long)2 * K) * (L + 1) + L) >= (1ll << 26))
        why = "coefficient index does not fit the link record of kernel 6";
    if (why && err) *err = why;
    return why ? 1 : 0;
}

int heom_sym_convert_links(const int2* links, int2* links2, long long nlinks, int N, int L, int packed, void* stream,
                           const char** err) {
    if (nlinks <= 0) return 0;
    const int threads = 256;
    const unsigned blocks = (unsigned)((nlinks + threads - 1) / threads);
    HEOM_LAUNCH(sym_convert_links_kernel, blocks, threads, 0, stream, links, links2, nlinks, N, L, packed);
#ifndef HEOM_HOST_EMU
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        if (err) *err = cudaGetErrorString(e);
        return 1;
    }
#endif
    (void)err;
    return 0;
}

#if defined(HEOM_SYM_SPLIT)
#define HEOM_SYM_DECL(n) int heom_sym_launch_n##n(const SymLaunch& s, const char** err);
HEOM_SYM_DECL(2) HEOM_SYM_DECL(3) HEOM_SYM_DECL(4) HEOM_SYM_DECL(5) HEOM_SYM_DECL(6) HEOM_SYM_DECL(7) HEOM_SYM_DECL(8)
#undef HEOM_SYM_DECL
#define HEOM_SYM_CALL(n) heom_sym_launch_n##n(s, err)
#else
#define HEOM_SYM_CALL(n) sym_launch_one<n>(s, err)
namespace {
template <int N>
int sym_launch_one(const SymLaunch& s, const char** err) {
    const int rc = sym_launch_n<N>(s);
    if (rc && err) *err = g_sym_err;
    return rc;
}
}  // namespace
#endif

int heom_sym_launch(const SymLaunch& s, const char** err) {
    switch (s.N) {
#ifndef HEOM_EMU_FEW_N   // (the sanitizer builds of the CPU tests instantiate N = 4 and 7 only)
        case 2: return HEOM_SYM_CALL(2);
        case 3: return HEOM_SYM_CALL(3);
        case 5: return HEOM_SYM_CALL(5);
        case 6: return HEOM_SYM_CALL(6);
        case 8: return HEOM_SYM_CALL(8);
#endif
        case 4: return HEOM_SYM_CALL(4);
        case 7: return HEOM_SYM_CALL(7);
        default: break;
    }
    if (err) *err = "kernel 6 needs 2 <= N <= 8";
    return 1;
}

// ---------------------------------------------------------------------------
// kernel 7: nt RK4 steps on upper-triangle arrays.  Y (full matrices) is packed into the
// work region, propagated there by the PACKED instantiations - the four triangle arrays play
// the roles of Y, SA, SB, ACC in run_stage's difference-form plan (heom_kernels.cu) - and
// unpacked back into Y at the end.
// ---------------------------------------------------------------------------
// full [n][N][N] -> upper triangles (unpack = 0) or back (unpack = 1), n ADOs
int heom_sym_pack(double2* tri, double2* full, long long n, int N, int unpack, void* stream) {
    if (n <= 0) return 0;
    const long long total = n * (N * (N + 1) / 2);
    const int threads = 256;
    const unsigned blocks = (unsigned)((total + threads - 1) / threads);
    if (unpack) {
        HEOM_LAUNCH(sym_unpack_kernel, blocks, threads, 0, stream, full, (const double2*)tri, n, N);
    } else {
        HEOM_LAUNCH(sym_pack_kernel, blocks, threads, 0, stream, tri, (const double2*)full, n, N);
    }
#ifndef HEOM_HOST_EMU
    if (cudaGetLastError() != cudaSuccess) return 1;
#endif
    return 0;
}

int heom_packed_propagate(const PackedRun& r, const char** err) {
    const int N = r.N, PK = N * (N + 1) / 2;
    const long long tri = r.nmax * PK;
    const size_t arr = ((size_t)tri * sizeof(double2) + 255) / 256 * 256 / sizeof(double2);   // aligned stride
    if (4 * arr * sizeof(double2) > r.work_bytes) {
        if (err) *err = "kernel 7: work region too small for four triangle arrays";
        return 1;
    }
    double2 *P0 = r.work, *P1 = P0 + arr, *P2 = P1 + arr, *P3 = P2 + arr;
    const int threads = 256;
    const unsigned blocks = (unsigned)((tri + threads - 1) / threads);
    HEOM_LAUNCH(sym_pack_kernel, blocks, threads, 0, r.stream, P0, (const double2*)r.Y, r.nmax, N);
    for (long long step = 0; step < r.nt; ++step) {
        for (int stage = 0; stage < 4; ++stage) {
            SymLaunch s{};
            s.a = r.tables;   // damp, link_ptr, links2 (packed form), cbase, kmode, ops, step_base, slot0, ...
            s.a.local_step = (int)step;
            s.a.y = P0;
            switch (stage) {
                case 0: s.a.yin = P0; s.a.out = P1; s.a.a = r.dt / 2; s.stage = 0; break;
                case 1: s.a.yin = P1; s.a.out = P2; s.a.a = r.dt / 2; s.stage = 1; break;
                case 2: s.a.yin = P2; s.a.out = P3; s.a.a = r.dt;     s.stage = 1; break;
                default:
                    s.a.yin = P3; s.a.s1 = P1; s.a.s2 = P2; s.a.out = P0;
                    s.a.a = 2.0 / r.dt; s.a.w = r.dt / 6; s.stage = 2;
                    break;
            }
            if (stage != 3) s.a.traj = nullptr;
            s.H = r.H;
            s.N = N; s.K = r.K; s.M = r.M; s.L = r.L; s.B = 1;
            s.hreal = r.hreal;
            s.packed = 1;
            s.prefetch = r.prefetch;
            s.warps = r.warps;
            s.sm_count = r.sm_count;
            s.part_lo = 0;
            s.part_hi = r.nmax;
            s.batch_elems = tri;
            s.traj_bstride = 0;
            s.stream = r.stream;
            s.sched = r.sched;
            s.sched_total = r.sched_total;
            if (heom_sym_launch(s, err)) return 1;
        }
    }
    HEOM_LAUNCH(sym_unpack_kernel, blocks, threads, 0, r.stream, r.Y, (const double2*)P0, r.nmax, N);
#ifndef HEOM_HOST_EMU
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        if (err) *err = cudaGetErrorString(e);
        return 1;
    }
#endif
    return 0;
}
#endif  // !HEOM_SYM_INST_N
