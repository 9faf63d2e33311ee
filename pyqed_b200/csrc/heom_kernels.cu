// heom_kernels.cu - sm_100a kernels and C ABI for HEOM/DEOM RK4 propagation.
//
// What the reference does per RK stage (pyqed/heom/deom.py:641-673, 725-766):
// for every ADO n, dρ_n/dt = -(Σ n_k γ_k) ρ_n - i[H, ρ_n]
//        - i Σ_k √n_k/√a_k (η_k Q_m ρ_{n-e_k} - η̄_k ρ_{n-e_k} Q_m)
//        - i Σ_k √(n_k+1) √a_k [Q_m, ρ_{n+e_k}]
// followed by separate axpy sweeps.  Here one kernel per stage evaluates the
// right-hand side and applies the stage update in the same pass:
//     k = F(y_in);  acc' = (first ? y : acc) + w k;  y_out = y + a k
// so an RK4 step moves 16 array passes of N*N*16 B per ADO.
//
// Layout: ADO arrays are [batch][slot][N][N] complex128, interleaved re/im, so
// one matrix element is one 128-bit access.  "slot" is the storage order
// (heom_core.cuh); links hold neighbour slots.
#include "heom_plan.cuh"
#include <cub/device/device_scan.cuh>

thread_local std::string g_heom_err;


// Structure of the problem (host): diagonal / single-entry coupling operators, Hermiticity of
// operators and bath, real H.  `supp`: [M][N+1] (count, rows) then [M][N] membership of the
// diagonal supports.
static void analyse_structure(pyqed_heom_plan* p, std::vector<unsigned char>& supp) {
    const int N = p->N, NN = N * N, K = p->K;
    const std::complex<double> Z(0, 0);
    bool diag = true, herm = true;
    for (int m = 0; m < p->M && diag; ++m)
        for (int i = 0; i < N && diag; ++i)
            for (int j = 0; j < N; ++j)
                if (i != j && (p->Q[(size_t)m * NN + i * N + j] != Z || p->Qd[(size_t)m * NN + i * N + j] != Z)) {
                    diag = false;
                    break;
                }
    auto is_herm = [&](const std::complex<double>* A) {
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j)
                if (A[i * N + j] != std::conj(A[j * N + i])) return false;
        return true;
    };
    herm = is_herm(p->H.data()) && is_herm(p->mu.data());
    for (int m = 0; m < p->M && herm; ++m)
        herm = is_herm(p->Q.data() + (size_t)m * NN) && is_herm(p->Qd.data() + (size_t)m * NN);
    for (int k = 0; k < K && herm; ++k)
        herm = p->expn[k].imag() == 0.0 && p->etaa[k].imag() == 0.0 && p->etaa[k].real() > 0.0 &&
               p->etar[k] == std::conj(p->etal[k]);
    p->q_diagonal = diag;
    p->herm_inputs = herm;
    p->use_qdiag = diag && p->opt_qdiag != 0 && N <= 8;
    supp.assign((size_t)p->M * (2 * N + 1), 0);
    unsigned char* ins = supp.data() + (size_t)p->M * (N + 1);
    for (int m = 0; m < p->M; ++m) {
        int c = 0;
        for (int j = 0; j < N; ++j) {
            const bool nzd = p->Q[(size_t)m * NN + j * N + j] != Z || p->Qd[(size_t)m * NN + j * N + j] != Z;
            if (nzd) {
                supp[(size_t)m * (N + 1) + 1 + c++] = (unsigned char)j;
                ins[(size_t)m * N + j] = 1;
            }
        }
        supp[(size_t)m * (N + 1)] = (unsigned char)c;
    }
    p->single_support = diag;
    for (int m = 0; m < p->M; ++m)
        if (supp[(size_t)m * (N + 1)] != 1) p->single_support = false;
    p->r0mode.assign(p->M, 0);
    for (int m = 0; m < p->M; ++m) p->r0mode[m] = supp[(size_t)m * (N + 1)] ? supp[(size_t)m * (N + 1) + 1] : 0;
    p->h_real = true;
    for (int e = 0; e < NN; ++e)
        if (p->H[e].imag() != 0.0 || p->mu[e].imag() != 0.0) p->h_real = false;
}

// kernels 6 / 7 (heom_stage_sym.cu) can take this problem: chosen automatically (kernel 0) or
// asked for (6, 7); the state must be Hermitian too, which is checked per propagation
static bool sym_wanted(const pyqed_heom_plan* p) {
    if (!(p->kernel == 0 || p->kernel == 6 || p->kernel == 7)) return false;
    if (p->N < 2 || p->N > 8 || !p->use_qdiag || !p->single_support || !p->herm_inputs) return false;
    if (p->opt_herm == 0 || p->opt_sym == 0 || p->opt_rk13 == 0) return false;
    const char* err = "";
    return heom_sym_supported(p->N, p->K, p->M, p->L, &err) == 0 &&
           (unsigned long long)p->nmax * p->N * p->N < (1ull << 32);
}

static int compute_layout(pyqed_heom_plan* p) {
    const size_t NN = (size_t)p->N * p->N, M1 = 1 + p->M;
    TableLayout& t = p->tl;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes);
        return o;
    };
    t.pascal = take(sizeof(long long) * p->side * p->side);
    t.keys = take((size_t)p->nmax * p->K);
    t.id_of_slot = take(sizeof(int) * p->nmax);
    t.slot_of_id = take(sizeof(int) * p->nmax);
    t.damp = take(sizeof(double2) * p->nmax);
    t.link_ptr = take(sizeof(int) * (p->nmax + 1));
    t.links = take(sizeof(int2) * (size_t)std::max(1ll, p->nlinks));
    t.coef = take(sizeof(double2) * 2 * 2 * p->K * (p->L + 1));
    t.ops_base = take(sizeof(double2) * M1 * NN);
    t.ops_dip = take(sizeof(double2) * M1 * NN);
    t.ops_t = take(sizeof(double2) * (size_t)p->B * M1 * NN);
    t.row_ptr = take(sizeof(short) * M1 * (p->N + 1));
    t.row_idx = take(sizeof(short) * M1 * NN);
    t.col_ptr = take(sizeof(short) * M1 * (p->N + 1));
    t.col_idx = take(sizeof(short) * M1 * NN);
    t.supp = take((size_t)p->M * (2 * p->N + 1));
    t.cbase = take(sizeof(double2) * 4 * p->K);
    t.kmode = take(sizeof(int) * p->K);
    t.lex2slot = take(sizeof(int) * (p->order == 2 ? p->nmax : 1));
    t.slot2lex = take(sizeof(int) * (p->order == 2 ? p->nmax : 1));
    t.step_base = take(sizeof(long long));
    t.sched = take(sizeof(unsigned) * 4);
    // second link table of kernels 6 / 7 (heom_stage_sym.cuh), only when they can take the problem
    bool sym = false;
    if (p->have_sys && p->have_coup && p->have_bath) {
        std::vector<unsigned char> supp;
        analyse_structure(p, supp);
        sym = sym_wanted(p);
    }
    t.links2 = take(sym ? sizeof(int2) * (size_t)std::max(1ll, p->nlinks) : 0);
    t.total = off;
    p->array_bytes = align_up(sizeof(double2) * (size_t)p->B * p->nmax * NN);
    return 0;
}

#include "heom_hierarchy.cuh"
#include "heom_stage_async.cuh"   // AsyncTables (shared-memory layout, also used by the resident kernels)
#include "heom_resident.cuh"      // ResidentArgs; the kernels are instantiated in heom_inst.cu
#include "heom_stage_generic.cuh"
#include "heom_dataflow.cuh"
#include "heom_dataflow_tma.cuh"

extern "C" {
static void fill_stage_args(pyqed_heom_plan* p, StageArgs& a);
}

// returns 0 = done, -1 = not applicable (caller falls back to per-stage launches), 1 = error
static int try_resident(pyqed_heom_plan* p) {
    const bool whole = p->part_lo == 0 && p->part_hi == p->nmax;
    const bool want = p->kernel == 4 || (p->kernel == 0 && p->opt_resident != 0);
    if (!want || !p->use_qdiag || p->N < 2 || p->N > 8 || !whole || p->ctx_nt <= 0) return -1;
    ResidentConfig rc;
    bool fits = false;
    switch (p->N) {
#define FITS_CASE(n) case n: fits = heom_resident_fits_##n(p, rc); break;
        FITS_CASE(2) FITS_CASE(3) FITS_CASE(4) FITS_CASE(5) FITS_CASE(6) FITS_CASE(7) FITS_CASE(8)
#undef FITS_CASE
    }
    if (!fits && p->opt_resident == 4) return -1;
    ResidentArgs ra;
    fill_stage_args(p, ra.s);
    ra.s.y = p->arr(ARR_Y);
    ra.fsys = p->ctx_use_fs ? p->d_fsys : nullptr;
    ra.fcoup = p->ctx_use_fc ? p->d_fcoup : nullptr;
    ra.ops_base = p->tab<double2>(p->tl.ops_base);
    ra.ops_dip = p->tab<double2>(p->tl.ops_dip);
    ra.dt = p->ctx_dt;
    ra.nt = p->ctx_nt;
    ra.apc = 0;
    ra.tdep = p->ctx_tdep ? 1 : 0;
    ra.maxlinks = std::max(1, std::min(p->L, p->K) + p->K);
    if (p->timing) {
        if (p->ev_used == p->ev.size()) {
            cudaEvent_t e0, e1;
            CU_TRY(cudaEventCreate(&e0));
            CU_TRY(cudaEventCreate(&e1));
            p->ev.emplace_back(e0, e1);
        }
        CU_TRY(cudaEventRecord(p->ev[p->ev_used].first, p->stream));
    }
    const bool hr = p->h_real && p->opt_hreal != 0;
    int rcode = -1;
    if (p->opt_resident != 4) {   // default: element-parallel kernel 5; "resident" = 4 forces kernel 4
#define RESE_CASE(n) \
    case n: rcode = heom_launch_resident_elem_##n(p, ra); break;
        switch (p->N) { RESE_CASE(2) RESE_CASE(3) RESE_CASE(4) RESE_CASE(5) RESE_CASE(6) RESE_CASE(7) RESE_CASE(8) }
#undef RESE_CASE
        if (rcode > 0) return rcode;
        if (rcode == 0) p->resident_kind = 5;
    }
    if (rcode != 0 && fits) {
#define RES_CASE(n) \
    case n: rcode = heom_launch_resident_##n(p, ra, rc, hr); break;
        switch (p->N) { RES_CASE(2) RES_CASE(3) RES_CASE(4) RES_CASE(5) RES_CASE(6) RES_CASE(7) RES_CASE(8) }
#undef RES_CASE
        if (rcode == 0) p->resident_kind = 4;
    }
    if (rcode == 0 && p->timing) {
        CU_TRY(cudaEventRecord(p->ev[p->ev_used].second, p->stream));
        p->ev_used++;
    }
    if (rcode == 0) p->resident_launches++;
    return rcode;
}

// ---- kernel 9: kernel 8's scheme for Hermitian problems with one CTA per ADO (heom_dataflow_tma.cuh) ----
// returns 0 = done, -1 = not applicable (kernel 8 or per-stage launches take it), 1 = error
extern "C" {
static void sparsity(const pyqed_heom_plan* p, int o, std::vector<short>& row_ptr, std::vector<short>& row_idx,
                     std::vector<short>& col_ptr, std::vector<short>& col_idx);
}
static int try_dataflow_tma(pyqed_heom_plan* p) {
    const bool whole = p->part_lo == 0 && p->part_hi == p->nmax;
    if (!(p->kernel == 0 || p->kernel == 9) || stage_kernel_of(p) != 2 || p->ctx_tdep || !whole || p->ctx_nt <= 0 ||
        p->N > 32 || p->N < 2 || p->opt_resident == 0 || p->opt_dataflow_tma == 0)
        return -1;
    if (!(p->herm_inputs && p->herm_state && p->opt_herm != 0) || p->ctx_nt >= (1ll << 29)) return -1;
    const int N = p->N, M1 = 1 + p->M, maxl = p->K + std::min(p->K, p->L);
    const long long total = p->nmax * (long long)p->B;
    const int nsm = sm_count_of(p->device);
    if (total > 2ll * nsm || M1 > DF9_MAXOPS || maxl > DF9_MAXL) return -1;
    int nnz = 0, nnz_h = 0;   // entries of the padded rows (every row of an operator as long as its longest)
    for (int o = 0; o < M1; ++o) {
        std::vector<short> a, b, c, d;
        sparsity(p, o, a, b, c, d);
        int longest = 0;
        for (int i = 0; i < N; ++i) longest = std::max(longest, (int)a[i + 1] - (int)a[i]);
        (o == 0 ? nnz_h : nnz) += longest * N;
    }
    // H with long rows: as a dense matrix in parameter space instead of the operator table
    const bool dense_h = nnz_h + nnz > DF9_MAXNNZ || nnz_h > 8 * N;
    if ((dense_h ? nnz : nnz + nnz_h) > DF9_MAXNNZ) return -1;
    const void* kern = dense_h ? (const void*)stage_dataflow_tma_kernel<true> : (const void*)stage_dataflow_tma_kernel<false>;
    const int units = N * (N - 1) / 2 + (N + 1) / 2;   // element pairs + pairs of diagonal elements
    const size_t smem = sizeof(Df9Smem);               // > 227 KB / 3: at most two CTAs per SM (the placement counts on it)
    static_assert(sizeof(Df9Smem) <= 113 * 1024 && sizeof(Df9Smem) > 78 * 1024, "two CTAs per SM");
    static PerDeviceOnce attr[2];
    if (attr[dense_h].need(p->device))
        CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
    int coop = 0, per_sm = 0;
    CU_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, p->device));
    if (dense_h)
        CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stage_dataflow_tma_kernel<true>, DF9_THREADS, smem));
    else
        CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stage_dataflow_tma_kernel<false>, DF9_THREADS, smem));
    const long long grid = total <= nsm ? total : 2ll * nsm;
    if (!coop || (long long)per_sm * nsm < grid) return -1;
    const size_t nflag = (size_t)total * DF9_FLAG_STRIDE;
    const size_t words = nflag + DF9_CTRL + DF9_MAXSM + (size_t)total;   // flags, control block, work order
    if (words > p->df9_cap) {
        if (p->d_df9) cudaFree(p->d_df9);
        p->d_df9 = nullptr;
        p->df9_cap = 0;
        CU_TRY(cudaMalloc(&p->d_df9, sizeof(unsigned) * words));
        p->df9_cap = words;
    }
    CU_TRY(cudaMemsetAsync(p->d_df9, 0, sizeof(unsigned) * (nflag + DF9_CTRL + DF9_MAXSM), p->stream));
    Dataflow9Args da;
    fill_stage_args(p, da.s);
    if (!da.s.herm) return -1;
    da.Y = p->arr(ARR_Y);
    da.P0 = p->arr(ARR_ACC);   // packed stage outputs live in the three work arrays
    da.P1 = p->arr(ARR_SA);
    da.P2 = p->arr(ARR_SB);
    da.units = units;
    da.flags = p->d_df9;
    da.ctrl = p->d_df9 + nflag;
    int* order = (int*)(p->d_df9 + nflag + DF9_CTRL + DF9_MAXSM);
    da.order = order;
    da.dt = p->ctx_dt;
    da.nt = p->ctx_nt;
    da.timeout_ns = 2000000000ull;   // a flag that does not arrive in 2 s poisons the result instead of hanging the GPU
    if (const char* t = std::getenv("PYQED_HEOM_DATAFLOW_TIMEOUT_MS"))   // (compute-sanitizer runs are 100x slower)
        da.timeout_ns = 1000000ull * std::strtoull(t, nullptr, 10);
    da.B = p->B;
    if (p->timing) {
        if (p->ev_used == p->ev.size()) {
            cudaEvent_t e0, e1;
            CU_TRY(cudaEventCreate(&e0));
            CU_TRY(cudaEventCreate(&e1));
            p->ev.emplace_back(e0, e1);
        }
        CU_TRY(cudaEventRecord(p->ev[p->ev_used].first, p->stream));
    }
    dataflow_order_kernel<<<1, 512, 0, p->stream>>>(da.s.link_ptr, p->nmax, total, order);
    if (post_launch(p, "dataflow_order_kernel")) return 1;
    void* kargs[] = {&da};
    CU_TRY(cudaLaunchCooperativeKernel(kern, dim3((unsigned)grid), dim3(DF9_THREADS), kargs, smem, p->stream));
    if (dense_h) p->dataflow_dense_launches++;
    if (post_launch(p, "stage_dataflow_tma_kernel")) return 1;
    if (p->timing) {
        CU_TRY(cudaEventRecord(p->ev[p->ev_used].second, p->stream));
        p->ev_used++;
    }
    p->dataflow_launches++;
    p->dataflow_tma_launches++;
    return 0;
}

// ---- kernel 8: persistent, ADO-to-ADO synchronised propagation of small hierarchies with N > 8 ----
// returns 0 = done, -1 = not applicable (caller falls back to per-stage launches), 1 = error
static int try_dataflow(pyqed_heom_plan* p) {
    const bool whole = p->part_lo == 0 && p->part_hi == p->nmax;
    if (!(p->kernel == 0 || p->kernel == 8 || p->kernel == 9) || stage_kernel_of(p) != 2 || p->ctx_tdep || !whole || p->ctx_nt <= 0 ||
        p->N > 32 || p->opt_resident == 0)
        return -1;
    const int N = p->N, NN = N * N;
    const size_t M1 = 1 + (size_t)p->M;
    const size_t smem = sizeof(double2) * (3 * NN + M1 * NN + 2 * 2 * (size_t)p->K) + sizeof(int2) * 2 * (size_t)p->K +
                        sizeof(short) * 2 * (M1 * NN + M1 * (N + 1)) + sizeof(int) * (M1 + 2) + 16;
    if (smem > 200 * 1024) return -1;
    static PerDeviceOnce attr;
    if (attr.need(p->device))
        CU_TRY(cudaFuncSetAttribute(stage_dataflow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int coop = 0, per_sm = 0;
    CU_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, p->device));
    CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stage_dataflow_kernel, DATAFLOW_THREADS, smem));
    const long long cap = (long long)per_sm * sm_count_of(p->device);
    const long long total = p->nmax * (long long)p->B;
    // worth it while a stage is latency bound: a few ADOs per CTA at most
    if (!coop || cap < 1 || (p->kernel != 8 && total > 8 * cap)) return -1;
    const size_t nflag = (size_t)total * DATAFLOW_FLAG_STRIDE;
    if (nflag > p->flags_cap) {
        if (p->d_flags) cudaFree(p->d_flags);
        p->d_flags = nullptr;
        p->flags_cap = 0;
        CU_TRY(cudaMalloc(&p->d_flags, sizeof(unsigned) * nflag));
        p->flags_cap = nflag;
    }
    CU_TRY(cudaMemsetAsync(p->d_flags, 0, sizeof(unsigned) * nflag, p->stream));
    DataflowArgs da;
    fill_stage_args(p, da.s);
    da.Y = p->arr(ARR_Y);
    da.SA = p->arr(ARR_SA);
    da.SB = p->arr(ARR_SB);
    da.ACC = p->arr(ARR_ACC);
    da.flags = p->d_flags;
    da.dt = p->ctx_dt;
    da.nt = p->ctx_nt;
    da.B = p->B;
    if (p->timing) {
        if (p->ev_used == p->ev.size()) {
            cudaEvent_t e0, e1;
            CU_TRY(cudaEventCreate(&e0));
            CU_TRY(cudaEventCreate(&e1));
            p->ev.emplace_back(e0, e1);
        }
        CU_TRY(cudaEventRecord(p->ev[p->ev_used].first, p->stream));
    }
    void* kargs[] = {&da};
    const unsigned grid = (unsigned)std::min<long long>(total, cap);
    CU_TRY(cudaLaunchCooperativeKernel((void*)stage_dataflow_kernel, dim3(grid), dim3(DATAFLOW_THREADS), kargs, smem,
                                       p->stream));
    if (post_launch(p, "stage_dataflow_kernel")) return 1;
    if (p->timing) {
        CU_TRY(cudaEventRecord(p->ev[p->ev_used].second, p->stream));
        p->ev_used++;
    }
    p->dataflow_launches++;
    return 0;
}

// kernel 6 (heom_stage_sym.cu) takes the difference-form RK4 stages of kernel 3 when every
// ADO is Hermitian, every Q_m has one non-zero diagonal entry and H does not depend on time
// (with push tables its PUSH instantiation stores the halo rows into the peers' arrays);
// everything else stays with kernel 3
static bool sym_eligible(const pyqed_heom_plan* p, const StageArgs& a, bool tdep) {
    return (p->kernel == 0 || p->kernel == 6 || p->kernel == 7) && p->links2_built && !tdep && a.herm &&
           p->single_support && p->opt_sym != 0 &&
           !a.push_ptr && a.scheme == 1 && !(a.first && a.last);   // (legacy fused push: kernel 3)
}
// kernels 6 / 7 read link records resolved for their storage (heom_stage_sym.cuh): (re)write the
// table from the general link table when the storage changes - one pass over 8 bytes per link
static int ensure_links2(pyqed_heom_plan* p, int packed) {
    if (p->links2_mode == packed) return 0;
    REQUIRE(!p->shard.on, "the link table of a sharded plan is resolved for its rank-local layout");
    const char* err = "";
    if (heom_sym_convert_links(p->tab<int2>(p->tl.links), p->tab<int2>(p->tl.links2), p->nlinks, p->N, p->L, packed,
                               p->stream, &err))
        return fail(std::string("sym_convert_links_kernel launch: ") + err);
    p->launches++;
    p->links2_mode = packed;
    return 0;
}

static int launch_sym(pyqed_heom_plan* p, const StageArgs& a, int sm_count) {
    if (ensure_links2(p, 0)) return 1;
    SymLaunch s{};
    s.a = sym_args_from_stage(a, p->tab<int2>(p->tl.links2));
    s.H = reinterpret_cast<const double*>(p->H.data());
    s.N = p->N;
    s.K = p->K;
    s.M = p->M;
    s.L = p->L;
    s.B = p->B;
    s.stage = sym_stage_kind(a);
    s.hreal = (p->h_real && p->opt_hreal != 0) ? 1 : 0;
    s.warps = p->warps;
    s.sm_count = sm_count;
    s.part_lo = p->part_lo;
    s.part_hi = p->part_hi;
    s.batch_elems = p->nmax * p->N * p->N;
    s.traj_bstride = a.traj_bstride;
    s.stream = p->stream;
    s.sched = p->opt_dynsched != 0 ? p->tab<unsigned>(p->tl.sched) : nullptr;
    s.sched_total = &p->sched_total;
    const char* err = "";
    if (heom_sym_launch(s, &err)) return fail(std::string("stage_rows_sym_kernel launch: ") + err);
    p->launches += p->B;
    p->sym_launches++;
    if (p->debug_sync) {
        cudaError_t e = cudaStreamSynchronize(p->stream);
        if (e != cudaSuccess) return fail(std::string("stage_rows_sym_kernel exec: ") + cudaGetErrorString(e));
    }
    return 0;
}

static int launch_stage(pyqed_heom_plan* p, const StageArgs& a, bool tdep) {
    if (p->part_hi <= p->part_lo) return 0;  // this rank owns nothing (tiny hierarchy, many ranks)
    const int sm_count = sm_count_of(p->device);
    if (p->timing) {
        if (p->ev_used == p->ev.size()) {
            cudaEvent_t e0, e1;
            CU_TRY(cudaEventCreate(&e0));
            CU_TRY(cudaEventCreate(&e1));
            p->ev.emplace_back(e0, e1);
        }
        CU_TRY(cudaEventRecord(p->ev[p->ev_used].first, p->stream));
    }
    int rc = 0;
    const int kern = stage_kernel_of(p);
    if (kern == 3 && sym_eligible(p, a, tdep)) {
        rc = launch_sym(p, a, sm_count);
    } else if (kern == 3 && a.first && a.last) {
        // single-stage (Euler) update: the async kernel only implements the difference-form
        // RK4 stages, so the plain-load row kernel takes it
        switch (p->N) {
#define ROWS_EULER_CASE(n) case n: rc = heom_launch_rows_##n(p, a, sm_count, tdep, p->use_qdiag); break;
            ROWS_EULER_CASE(2) ROWS_EULER_CASE(3) ROWS_EULER_CASE(4) ROWS_EULER_CASE(5) ROWS_EULER_CASE(6)
            ROWS_EULER_CASE(7) ROWS_EULER_CASE(8)
#undef ROWS_EULER_CASE
        }
    } else if (kern == 3) {
        REQUIRE(p->N >= 2 && p->N <= 8 && p->use_qdiag,
                "kernel 3 needs 2 <= N <= 8 and diagonal coupling operators");
#define ASYNC_CASE(n)                                                                  \
    case n:                                                                            \
        rc = heom_launch_async_##n(p, a, sm_count, tdep, p->h_real && p->opt_hreal != 0); \
        break;
        switch (p->N) {
            ASYNC_CASE(2) ASYNC_CASE(3) ASYNC_CASE(4) ASYNC_CASE(5) ASYNC_CASE(6) ASYNC_CASE(7) ASYNC_CASE(8)
        }
#undef ASYNC_CASE
    } else if (kern == 1) {
        REQUIRE(p->N >= 2 && p->N <= 8, "kernel 1 needs 2 <= N <= 8");
#define ROWS_CASE(n)                                              \
    case n:                                                       \
        rc = heom_launch_rows_##n(p, a, sm_count, tdep, p->use_qdiag); \
        break;
        switch (p->N) {
            ROWS_CASE(2) ROWS_CASE(3) ROWS_CASE(4) ROWS_CASE(5) ROWS_CASE(6) ROWS_CASE(7) ROWS_CASE(8)
        }
#undef ROWS_CASE
    } else {
        const int NN = p->N * p->N;
        // one thread per matrix element where possible: the per-thread chain of dependent
        // neighbour loads is what bounds small hierarchies
        int threads = std::min(1024, (NN + 31) / 32 * 32);
        const size_t M1 = 1 + (size_t)p->M;
        const size_t smem = sizeof(double2) * (NN + M1 * NN + 2 * 2 * (size_t)p->K) + sizeof(int2) * 2 * (size_t)p->K +
                            sizeof(short) * 2 * (M1 * NN + M1 * (p->N + 1)) + 16;
        REQUIRE(smem <= 200 * 1024, "operators too large for the generic kernel's shared memory");
        static PerDeviceOnce gen_attr;
        if (gen_attr.need(p->device))
            CU_TRY(cudaFuncSetAttribute(stage_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        dim3 grid((unsigned)std::max(1ll, std::min(p->part_hi - p->part_lo, (long long)sm_count * 32)), p->B);
        stage_generic_kernel<<<grid, threads, smem, p->stream>>>(a);
        rc = post_launch(p, "stage_generic_kernel");
    }
    if (rc) return rc;
    if (p->timing) {
        CU_TRY(cudaEventRecord(p->ev[p->ev_used].second, p->stream));
        p->ev_used++;
    }
    return 0;
}

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

int pyqed_heom_version(void) { return 1; }
const char* pyqed_heom_last_error(void) { return g_heom_err.c_str(); }

static bool build_pascal(int side, std::vector<long long>& tab) {
    tab.assign((size_t)side * side, 0);
    const long long SAT = (1ll << 62);
    for (int a = 0; a < side; ++a) {
        tab[(size_t)a * side] = 1;
        for (int b = 1; b <= a; ++b) {
            const long long u = tab[(size_t)(a - 1) * side + b - 1];
            const long long v = b <= a - 1 ? tab[(size_t)(a - 1) * side + b] : 0;
            tab[(size_t)a * side + b] = (u >= SAT || v >= SAT || u + v > SAT) ? SAT : u + v;
        }
    }
    return true;
}

int64_t pyqed_heom_hierarchy_size(int nind, int lmax) {
    if (nind < 1 || lmax < 0 || nind > heom::MAX_NIND || nind + lmax + 1 > heom::MAX_SIDE) return -1;
    std::vector<long long> tab;
    build_pascal(nind + lmax + 1, tab);
    const long long v = tab[(size_t)(lmax + nind) * (nind + lmax + 1) + lmax];
    return v >= (1ll << 31) ? -1 : v;
}

int pyqed_heom_plan_create(pyqed_heom_plan** out, int device, int nsys, int nind, int nmod,
                           int lmax, int batch) {
    REQUIRE(out, "plan pointer is null");
    REQUIRE(nsys >= 1 && nsys <= 55, "nsys must be in [1, 55]");
    REQUIRE(nind >= 1 && nind <= heom::MAX_NIND, "nind must be in [1, 64]");
    REQUIRE(nmod >= 1 && nmod <= 127, "nmod must be in [1, 127]");
    REQUIRE(lmax >= 0 && lmax <= 255 && nind + lmax + 1 <= heom::MAX_SIDE, "lmax out of range");
    REQUIRE(batch >= 1 && batch <= 65535, "batch must be in [1, 65535]");
    const int64_t nmax = pyqed_heom_hierarchy_size(nind, lmax);
    REQUIRE(nmax > 0, "hierarchy too large (nmax must be < 2^31)");
    int ndev = 0;
    CU_TRY(cudaGetDeviceCount(&ndev));
    REQUIRE(device >= 0 && device < ndev, "no such CUDA device");
    CU_TRY(cudaSetDevice(device));
    auto* p = new pyqed_heom_plan();
    p->device = device;
    p->N = nsys;
    p->K = nind;
    p->M = nmod;
    p->L = lmax;
    p->B = batch;
    p->nmax = nmax;
    p->side = nind + lmax + 1;
    build_pascal(p->side, p->pascal);
    // links: every n+e_k edge appears once as a "plus" and once as a "minus" link
    const long long below = lmax >= 1 ? p->pascal[(size_t)(lmax - 1 + nind) * p->side + nind] : 0;
    p->nlinks = 2ll * nind * below;
    if (p->nlinks >= (1ll << 31)) {
        delete p;
        return fail("too many links for 32-bit offsets");
    }
    const char* dbg = getenv("PYQED_HEOM_DEBUG");
    p->debug_sync = dbg && dbg[0] == '1';
    *out = p;
    return 0;
}

void pyqed_heom_plan_destroy(pyqed_heom_plan* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    for (auto& e : p->ev) {
        cudaEventDestroy(e.first);
        cudaEventDestroy(e.second);
    }
    if (p->d_fsys) cudaFree(p->d_fsys);
    if (p->d_fcoup) cudaFree(p->d_fcoup);
    if (p->d_peer) cudaFree(p->d_peer);
    if (p->shard.d_peer) cudaFree(p->shard.d_peer);
    if (p->d_flags) cudaFree(p->d_flags);
    if (p->d_df9) cudaFree(p->d_df9);
    delete p;
}

static void to_complex(std::vector<std::complex<double>>& dst, const double* src, size_t n) {
    dst.resize(n);
    for (size_t i = 0; i < n; ++i)
        dst[i] = src ? std::complex<double>(src[2 * i], src[2 * i + 1]) : std::complex<double>(0, 0);
}
static bool any_nonzero(const std::vector<std::complex<double>>& v) {
    for (auto& x : v)
        if (x != std::complex<double>(0, 0)) return true;
    return false;
}

int pyqed_heom_set_system(pyqed_heom_plan* p, const double* H, const double* mu) {
    REQUIRE(p && H, "set_system: null argument");
    const size_t NN = (size_t)p->N * p->N;
    to_complex(p->H, H, NN);
    to_complex(p->mu, mu, NN);
    p->mu_nonzero = any_nonzero(p->mu);
    p->have_sys = true;
    p->built = false;
    return 0;
}

int pyqed_heom_set_coupling(pyqed_heom_plan* p, const double* Q, const double* Qdip) {
    REQUIRE(p && Q, "set_coupling: null argument");
    const size_t n = (size_t)p->M * p->N * p->N;
    to_complex(p->Q, Q, n);
    to_complex(p->Qd, Qdip, n);
    p->qd_nonzero = any_nonzero(p->Qd);
    p->have_coup = true;
    p->built = false;
    return 0;
}

int pyqed_heom_set_bath(pyqed_heom_plan* p, const double* expn, const double* etal,
                        const double* etar, const double* etaa, const int64_t* mode) {
    REQUIRE(p && expn && etal && etar && etaa && mode, "set_bath: null argument");
    to_complex(p->expn, expn, p->K);
    to_complex(p->etal, etal, p->K);
    to_complex(p->etar, etar, p->K);
    to_complex(p->etaa, etaa, p->K);
    p->mode.assign(mode, mode + p->K);
    for (int k = 0; k < p->K; ++k)
        REQUIRE(p->mode[k] >= 0 && p->mode[k] < p->M, "set_bath: mode index out of range");
    p->have_bath = true;
    p->built = false;
    return 0;
}

int pyqed_heom_set_order(pyqed_heom_plan* p, int order) {
    REQUIRE(p, "null plan");
    REQUIRE(order >= 0 && order <= 2, "order must be 0 (reference), 1 (lexicographic) or 2 (blocked lexicographic)");
    REQUIRE(!p->bound, "set_order must precede bind");
    p->order = order;
    return 0;
}

int pyqed_heom_set_tuning(pyqed_heom_plan* p, int kernel, int warps, int use_graph) {
    REQUIRE(p, "null plan");
    REQUIRE((kernel >= 0 && kernel <= 4) || (kernel >= 6 && kernel <= 9), "kernel must be 0..4 or 6..9");
    REQUIRE(warps >= 0 && warps <= 16, "warps_per_cta must be in [0, 16]");
    REQUIRE(use_graph == 0, "use_graph is reserved and must be 0");
    p->kernel = kernel;
    p->warps = warps;
    p->use_graph = use_graph;
    return 0;
}

int pyqed_heom_set_option(pyqed_heom_plan* p, const char* name, int value) {
    REQUIRE(p && name, "null argument");
    const std::string n(name);
    if (n == "qdiag") p->opt_qdiag = value;
    else if (n == "hermitian") p->opt_herm = value;
    else if (n == "sym") p->opt_sym = value;
    else if (n == "real_h") p->opt_hreal = value;
    else if (n == "resident") p->opt_resident = value;
    else if (n == "rk13") p->opt_rk13 = value;
    else if (n == "prefetch") p->opt_prefetch = value;
    else if (n == "dynsched") p->opt_dynsched = value;
    else if (n == "packed") p->opt_packed = value;
    else if (n == "dataflow_tma") p->opt_dataflow_tma = value;
    else if (n == "debug_sync") p->debug_sync = value != 0;
    else return fail("unknown option '" + n + "'");
    return 0;
}

int64_t pyqed_heom_get_info(pyqed_heom_plan* p, const char* name) {
    if (!p || !name) return -1;
    const std::string n(name);
    if (n == "qdiag") return p->use_qdiag;
    if (n == "q_diagonal") return p->q_diagonal;
    if (n == "hermitian") return p->herm_inputs && p->herm_state && p->opt_herm != 0;
    if (n == "hermitian_inputs") return p->herm_inputs && p->opt_herm != 0;
    if (n == "sym") return p->herm_inputs && p->herm_state && p->opt_herm != 0 && p->single_support && p->opt_sym != 0;
    if (n == "real_h") return p->h_real && p->opt_hreal != 0;
    if (n == "resident_launches") return p->resident_launches;
    if (n == "resident_kind") return p->resident_kind;
    if (n == "sym_launches") return p->sym_launches;
    if (n == "packed_steps") return p->packed_steps;
    if (n == "dataflow_launches") return p->dataflow_launches;
    if (n == "dataflow_tma_launches") return p->dataflow_tma_launches;
    if (n == "dataflow_dense_launches") return p->dataflow_dense_launches;
    if (n == "rk_scheme") return rk_scheme(p) ? 1 : 0;
    if (n == "stage_kernel") return stage_kernel_of(p);
    if (n == "nlinks") return p->nlinks;
    if (n == "nmax") return p->nmax;
    if (n == "slot0") return p->slot0;
    if (n == "table_bytes") return (int64_t)p->tl.total;
    if (n == "off_id_of_slot") return (int64_t)p->tl.id_of_slot;
    if (n == "off_link_ptr") return (int64_t)p->tl.link_ptr;
    if (n == "off_links") return (int64_t)p->tl.links;
    if (n == "array_bytes") return (int64_t)p->array_bytes;
    if (n == "part_lo") return p->part_lo;
    if (n == "part_hi") return p->part_hi;
    if (n == "sym_inputs") return sym_wanted(p) ? 1 : 0;
    if (n == "push_slots") return heom_sym_push_slots();
    if (n == "off_links2") return p->links2_built ? (int64_t)p->tl.links2 : -1;
    if (n == "shard_packed") return p->shard.on ? (p->shard.packed ? 1 : 0) : -1;
    if (n == "shard_epoch") return p->shard.epoch;
    return -1;
}

int pyqed_heom_table_bytes(pyqed_heom_plan* p, size_t* bytes) {
    REQUIRE(p && bytes, "null argument");
    compute_layout(p);
    *bytes = p->tl.total;
    return 0;
}
int pyqed_heom_state_bytes(pyqed_heom_plan* p, size_t* bytes) {
    REQUIRE(p && bytes, "null argument");
    compute_layout(p);
    *bytes = 4 * p->array_bytes;
    return 0;
}

int pyqed_heom_bind(pyqed_heom_plan* p, void* d_tables, size_t table_bytes, void* d_state,
                    size_t state_bytes, void* stream) {
    REQUIRE(p && d_tables, "bind: null argument");
    compute_layout(p);
    REQUIRE(table_bytes >= p->tl.total, "bind: table buffer too small");
    // d_state may be NULL for a sharded run: pyqed_heom_shard_setup binds the rank-local state buffer
    REQUIRE(!d_state || state_bytes >= 4 * p->array_bytes, "bind: state buffer too small");
    REQUIRE(((uintptr_t)d_tables % 256) == 0 && ((uintptr_t)d_state % 256) == 0,
            "bind: buffers must be 256-byte aligned");
    p->d_tables = (char*)d_tables;
    p->bound_table_bytes = table_bytes;
    p->d_state = (char*)d_state;
    p->stream = (cudaStream_t)stream;
    p->bound = true;
    p->built = false;
    return 0;
}

// sparsity lists of operator o (union of the static and the dipole pattern)
static void sparsity(const pyqed_heom_plan* p, int o, std::vector<short>& row_ptr,
                     std::vector<short>& row_idx, std::vector<short>& col_ptr,
                     std::vector<short>& col_idx) {
    const int N = p->N, NN = N * N;
    const std::complex<double>* base = o == 0 ? p->H.data() : p->Q.data() + (size_t)(o - 1) * NN;
    const std::complex<double>* dip = o == 0 ? p->mu.data() : p->Qd.data() + (size_t)(o - 1) * NN;
    auto nz = [&](int i, int j) {
        return base[i * N + j] != std::complex<double>(0, 0) ||
               dip[i * N + j] != std::complex<double>(0, 0);
    };
    row_ptr.assign(N + 1, 0);
    col_ptr.assign(N + 1, 0);
    row_idx.assign(NN, 0);
    col_idx.assign(NN, 0);
    int w = 0;
    for (int i = 0; i < N; ++i) {
        row_ptr[i] = (short)w;
        for (int l = 0; l < N; ++l)
            if (nz(i, l)) row_idx[w++] = (short)l;
    }
    row_ptr[N] = (short)w;
    w = 0;
    for (int j = 0; j < N; ++j) {
        col_ptr[j] = (short)w;
        for (int l = 0; l < N; ++l)
            if (nz(l, j)) col_idx[w++] = (short)l;
    }
    col_ptr[N] = (short)w;
}

int pyqed_heom_build_hierarchy(pyqed_heom_plan* p) {
    REQUIRE(p && p->bound, "build_hierarchy: bind the buffers first");
    REQUIRE(p->have_sys && p->have_coup && p->have_bath,
            "build_hierarchy: system, coupling and bath must be set");
    CU_TRY(cudaSetDevice(p->device));
    const int N = p->N, NN = N * N, K = p->K, L = p->L, M1 = 1 + p->M;
    const TableLayout& t = p->tl;
    cudaStream_t s = p->stream;
    // small host-built tables -------------------------------------------------
    CU_TRY(cudaMemcpyAsync(p->d_tables + t.pascal, p->pascal.data(),
                           sizeof(long long) * p->pascal.size(), cudaMemcpyHostToDevice, s));
    // coefficient pairs (alphaL, alphaR) per (dir, k, n_eff); generate_dot_element,
    // deom.py:656-664: minus: -i sqrt(n)/sqrt(a) (eta_l Q rho - eta_r rho Q)
    //                  plus : -i sqrt(n+1) sqrt(a) (Q rho - rho Q)
    std::vector<double> coef((size_t)2 * 2 * 2 * K * (L + 1), 0.0);
    const std::complex<double> I(0, 1);
    for (int k = 0; k < K; ++k) {
        const std::complex<double> sa = std::sqrt(p->etaa[k]);
        for (int n = 1; n <= L; ++n) {
            const std::complex<double> cm = I * std::sqrt((double)n) / sa;
            const std::complex<double> cp = I * std::sqrt((double)n) * sa;
            const std::complex<double> vals[2][2] = {{-cm * p->etal[k], cm * p->etar[k]}, {-cp, cp}};
            for (int dir = 0; dir < 2; ++dir) {
                const size_t ci = ((size_t)dir * K + k) * (L + 1) + n;
                coef[4 * ci + 0] = vals[dir][0].real();
                coef[4 * ci + 1] = vals[dir][0].imag();
                coef[4 * ci + 2] = vals[dir][1].real();
                coef[4 * ci + 3] = vals[dir][1].imag();
            }
        }
    }
    CU_TRY(cudaMemcpyAsync(p->d_tables + t.coef, coef.data(), sizeof(double) * coef.size(),
                           cudaMemcpyHostToDevice, s));
    {   // per-dissipaton base coefficients (n_eff = 1); the async kernel scales by sqrt(n_eff)
        std::vector<double> cb((size_t)8 * K);
        for (int k = 0; k < K; ++k) {
            const std::complex<double> sa = std::sqrt(p->etaa[k]);
            const std::complex<double> v[4] = {-(I / sa) * p->etal[k], (I / sa) * p->etar[k], -I * sa, I * sa};
            for (int q = 0; q < 4; ++q) {
                cb[(size_t)8 * k + 2 * q] = v[q].real();
                cb[(size_t)8 * k + 2 * q + 1] = v[q].imag();
            }
        }
        CU_TRY(cudaMemcpyAsync(p->d_tables + t.cbase, cb.data(), sizeof(double) * cb.size(),
                               cudaMemcpyHostToDevice, s));
        CU_TRY(cudaStreamSynchronize(s));
    }
    // operators and their sparsity lists
    std::vector<double> base((size_t)M1 * NN * 2), dip((size_t)M1 * NN * 2);
    std::vector<short> rp((size_t)M1 * (N + 1)), ri((size_t)M1 * NN), cp((size_t)M1 * (N + 1)),
        cidx((size_t)M1 * NN);
    for (int o = 0; o < M1; ++o) {
        const std::complex<double>* bsrc = o == 0 ? p->H.data() : p->Q.data() + (size_t)(o - 1) * NN;
        const std::complex<double>* dsrc = o == 0 ? p->mu.data() : p->Qd.data() + (size_t)(o - 1) * NN;
        for (int e = 0; e < NN; ++e) {
            base[((size_t)o * NN + e) * 2] = bsrc[e].real();
            base[((size_t)o * NN + e) * 2 + 1] = bsrc[e].imag();
            dip[((size_t)o * NN + e) * 2] = dsrc[e].real();
            dip[((size_t)o * NN + e) * 2 + 1] = dsrc[e].imag();
        }
        std::vector<short> a, b, c, d;
        sparsity(p, o, a, b, c, d);
        std::copy(a.begin(), a.end(), rp.begin() + (size_t)o * (N + 1));
        std::copy(b.begin(), b.end(), ri.begin() + (size_t)o * NN);
        std::copy(c.begin(), c.end(), cp.begin() + (size_t)o * (N + 1));
        std::copy(d.begin(), d.end(), cidx.begin() + (size_t)o * NN);
    }
    CU_TRY(cudaMemcpyAsync(p->d_tables + t.ops_base, base.data(), sizeof(double) * base.size(),
                           cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(p->d_tables + t.ops_dip, dip.data(), sizeof(double) * dip.size(),
                           cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(p->d_tables + t.row_ptr, rp.data(), sizeof(short) * rp.size(),
                           cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(p->d_tables + t.row_idx, ri.data(), sizeof(short) * ri.size(),
                           cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(p->d_tables + t.col_ptr, cp.data(), sizeof(short) * cp.size(),
                           cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(p->d_tables + t.col_idx, cidx.data(), sizeof(short) * cidx.size(),
                           cudaMemcpyHostToDevice, s));
    {
        std::vector<unsigned char> supp;
        analyse_structure(p, supp);
        CU_TRY(cudaMemcpyAsync(p->d_tables + t.supp, supp.data(), supp.size(), cudaMemcpyHostToDevice, s));
        CU_TRY(cudaStreamSynchronize(s));
    }
    CU_TRY(cudaMemsetAsync(p->d_tables + t.step_base, 0, sizeof(long long), s));
    // expn and mode for the builder kernels: staged in ops_t (rebuilt before use)
    std::vector<double> ex(2 * K);
    std::vector<int> md(K);
    for (int k = 0; k < K; ++k) {
        ex[2 * k] = p->expn[k].real();
        ex[2 * k + 1] = p->expn[k].imag();
        md[k] = (int)p->mode[k] | (p->r0mode[p->mode[k]] << 8);
    }
    double2* d_expn = nullptr;
    int* d_mode = nullptr;
    CU_TRY(cudaMalloc(&d_expn, sizeof(double2) * K));
    CU_TRY(cudaMalloc(&d_mode, sizeof(int) * K));
    CU_TRY(cudaMemcpyAsync(d_expn, ex.data(), sizeof(double) * ex.size(), cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(d_mode, md.data(), sizeof(int) * K, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(p->d_tables + t.kmode, md.data(), sizeof(int) * K, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaStreamSynchronize(s));  // host vectors above go out of scope
    // device-built tables -------------------------------------------------------
    HierArgs h;
    h.pascal = p->tab<long long>(t.pascal);
    h.side = p->side;
    h.K = K;
    h.L = L;
    h.order = p->order;
    h.nmax = p->nmax;
    h.keys = p->tab<uint8_t>(t.keys);
    h.id_of_slot = p->tab<int>(t.id_of_slot);
    h.slot_of_id = p->tab<int>(t.slot_of_id);
    h.damp = p->tab<double2>(t.damp);
    h.link_ptr = p->tab<int>(t.link_ptr);
    h.links = p->tab<int2>(t.links);
    h.expn = d_expn;
    h.mode = d_mode;
    h.lex2slot = p->tab<int>(t.lex2slot);
    h.slot2lex = p->tab<int>(t.slot2lex);
    const int threads = 128;
    const unsigned blocks = (unsigned)((p->nmax + threads - 1) / threads);
    if (p->order == 2) {
        const long long nblk = (p->nmax + ORDER2_BLOCK - 1) / ORDER2_BLOCK;
        hier_blockperm_kernel<<<(unsigned)((nblk + 63) / 64), 64, 0, s>>>(h);
        if (post_launch(p, "hier_blockperm_kernel")) return 1;
    }
    hier_keys_kernel<<<blocks, threads, 0, s>>>(h);
    if (post_launch(p, "hier_keys_kernel")) return 1;
    size_t tmp_bytes = 0;
    CU_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, h.link_ptr, h.link_ptr,
                                         (int)(p->nmax + 1), s));
    void* d_tmp = nullptr;
    CU_TRY(cudaMalloc(&d_tmp, std::max<size_t>(tmp_bytes, 16)));
    CU_TRY(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, h.link_ptr, h.link_ptr,
                                         (int)(p->nmax + 1), s));
    p->launches++;
    hier_links_kernel<<<blocks, threads, 0, s>>>(h);
    if (post_launch(p, "hier_links_kernel")) return 1;
    int total_links = 0, slot0 = 0;
    CU_TRY(cudaMemcpyAsync(&total_links, h.link_ptr + p->nmax, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(&slot0, h.slot_of_id, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    cudaFree(d_tmp);
    cudaFree(d_expn);
    cudaFree(d_mode);
    REQUIRE(total_links == p->nlinks, "hierarchy builder: link count mismatch (" +
                                          std::to_string(total_links) + " vs " +
                                          std::to_string(p->nlinks) + ")");
    p->slot0 = slot0;
    p->links2_built = false;
    if (sym_wanted(p)) {
        if (t.links2 + sizeof(int2) * (size_t)p->nlinks <= p->bound_table_bytes) {   // (options changed after bind)
            p->links2_built = true;   // (filled for full or packed storage when a propagation needs it)
            p->links2_mode = -1;
        }
    }
    if (p->part_hi <= p->part_lo) {
        p->part_lo = 0;
        p->part_hi = p->nmax;
    }
    p->built = true;
    return 0;
}

int pyqed_heom_get_keys(pyqed_heom_plan* p, uint8_t* keys_host) {
    REQUIRE(p && p->built && keys_host, "get_keys: build the hierarchy first");
    CU_TRY(cudaSetDevice(p->device));
    const size_t bytes = (size_t)p->nmax * p->K;
    std::vector<uint8_t> slot_keys(bytes);
    std::vector<int> ids(p->nmax);
    CU_TRY(cudaMemcpyAsync(slot_keys.data(), p->d_tables + p->tl.keys, bytes, cudaMemcpyDeviceToHost,
                           p->stream));
    CU_TRY(cudaMemcpyAsync(ids.data(), p->d_tables + p->tl.id_of_slot, sizeof(int) * p->nmax,
                           cudaMemcpyDeviceToHost, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    for (long long sidx = 0; sidx < p->nmax; ++sidx)
        memcpy(keys_host + (size_t)ids[sidx] * p->K, slot_keys.data() + (size_t)sidx * p->K, p->K);
    return 0;
}

int pyqed_heom_set_state(pyqed_heom_plan* p, const double* rho0_host) {
    REQUIRE(p && p->built && rho0_host, "set_state: build the hierarchy first");
    REQUIRE(p->d_state && !p->shard.on, "set_state: no full-size state buffer bound (sharded plans use shard_set_state)");
    CU_TRY(cudaSetDevice(p->device));
    const size_t NN = (size_t)p->N * p->N;
    {
        bool h = true;
        for (int b = 0; b < p->B && h; ++b)
            for (int i = 0; i < p->N && h; ++i)
                for (int j = 0; j < p->N; ++j) {
                    const double* x = rho0_host + 2 * ((size_t)b * NN + i * p->N + j);
                    const double* y = rho0_host + 2 * ((size_t)b * NN + j * p->N + i);
                    if (x[0] != y[0] || x[1] != -y[1]) {
                        h = false;
                        break;
                    }
                }
        p->herm_state = h;
    }
    CU_TRY(cudaMemsetAsync(p->d_state, 0, 4 * p->array_bytes, p->stream));
    CU_TRY(cudaMemcpy2DAsync(p->arr(ARR_Y) + (size_t)p->slot0 * NN, sizeof(double2) * p->nmax * NN,
                             rho0_host, sizeof(double2) * NN, sizeof(double2) * NN, p->B,
                             cudaMemcpyHostToDevice, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    return 0;
}

static int permute(pyqed_heom_plan* p, double2* dst, const double2* src, bool to_id_order) {
    const int NN = p->N * p->N;
    const long long total = p->nmax * NN;
    dim3 grid((unsigned)std::min<long long>((total + 255) / 256, 148 * 32), p->B);
    // to id order:   dst[id] = src[slot_of_id[id]]  (gather through slot_of_id)
    // to slot order: dst[slot] = src[id_of_slot[slot]]
    const int* map = p->tab<int>(to_id_order ? p->tl.slot_of_id : p->tl.id_of_slot);
    permute_kernel<<<grid, 256, 0, p->stream>>>(dst, src, map, p->nmax, NN, 0);
    return post_launch(p, "permute_kernel");
}

int pyqed_heom_load_ados(pyqed_heom_plan* p, const double* ados_host) {
    REQUIRE(p && p->built && ados_host, "load_ados: build the hierarchy first");
    REQUIRE(p->d_state && !p->shard.on, "load_ados: no full-size state buffer bound");
    CU_TRY(cudaSetDevice(p->device));
    p->herm_state = false;  // arbitrary ADOs: do not assume Hermiticity
    const size_t bytes = sizeof(double2) * (size_t)p->B * p->nmax * p->N * p->N;
    if (p->order == 0) {
        CU_TRY(cudaMemcpyAsync(p->arr(ARR_Y), ados_host, bytes, cudaMemcpyHostToDevice, p->stream));
    } else {
        CU_TRY(cudaMemcpyAsync(p->arr(ARR_ACC), ados_host, bytes, cudaMemcpyHostToDevice, p->stream));
        if (permute(p, p->arr(ARR_Y), p->arr(ARR_ACC), false)) return 1;
    }
    CU_TRY(cudaStreamSynchronize(p->stream));
    return 0;
}

int pyqed_heom_get_ados(pyqed_heom_plan* p, double* ados_host) {
    REQUIRE(p && p->built && ados_host, "get_ados: build the hierarchy first");
    REQUIRE(p->d_state && !p->shard.on, "get_ados: no full-size state buffer bound (sharded plans use shard_get_owned)");
    CU_TRY(cudaSetDevice(p->device));
    const size_t bytes = sizeof(double2) * (size_t)p->B * p->nmax * p->N * p->N;
    const double2* src = p->arr(ARR_Y);
    if (p->order != 0) {
        if (permute(p, p->arr(ARR_ACC), p->arr(ARR_Y), true)) return 1;
        src = p->arr(ARR_ACC);
    }
    CU_TRY(cudaMemcpyAsync(ados_host, src, bytes, cudaMemcpyDeviceToHost, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    return 0;
}

static int upload_fields(pyqed_heom_plan* p, const double* fsys, const double* fcoup, long long nt) {
    const size_t n = (size_t)p->B * nt * 3;
    if (n > p->field_cap) {
        if (p->d_fsys) cudaFree(p->d_fsys);
        if (p->d_fcoup) cudaFree(p->d_fcoup);
        p->d_fsys = p->d_fcoup = nullptr;
        CU_TRY(cudaMalloc(&p->d_fsys, sizeof(double) * n));
        CU_TRY(cudaMalloc(&p->d_fcoup, sizeof(double) * n));
        p->field_cap = n;
    }
    if (fsys)
        CU_TRY(cudaMemcpyAsync(p->d_fsys, fsys, sizeof(double) * n, cudaMemcpyHostToDevice, p->stream));
    if (fcoup)
        CU_TRY(cudaMemcpyAsync(p->d_fcoup, fcoup, sizeof(double) * n, cudaMemcpyHostToDevice, p->stream));
    return 0;
}

// ---- propagation: begin (context) / single stage / whole run -----------------
static void fill_stage_args(pyqed_heom_plan* p, StageArgs& a) {
    const int N = p->N, NN = N * N, M1 = 1 + p->M;
    const TableLayout& t = p->tl;
    memset(&a, 0, sizeof(a));
    a.damp = p->tab<double2>(t.damp);
    a.link_ptr = p->tab<int>(t.link_ptr);
    a.links = p->tab<int2>(t.links);
    a.coef = p->tab<double2>(t.coef);
    a.ops = p->tab<double2>(p->ctx_tdep ? t.ops_t : t.ops_base);
    a.ops_bstride = p->ctx_tdep ? (long long)M1 * NN : 0;
    a.row_ptr = p->tab<short>(t.row_ptr);
    a.row_idx = p->tab<short>(t.row_idx);
    a.col_ptr = p->tab<short>(t.col_ptr);
    a.col_idx = p->tab<short>(t.col_idx);
    a.supp = p->tab<unsigned char>(t.supp);
    a.herm = (p->herm_inputs && p->herm_state && p->opt_herm != 0) ? 1 : 0;
    a.ncoef = 2 * p->K * (p->L + 1);
    a.nmod = p->M;
    a.nind = p->K;
    a.lmax = p->L;
    a.cbase = p->tab<double2>(t.cbase);
    a.kmode = p->tab<int>(t.kmode);
    a.traj = p->ctx_traj;
    a.step_base = p->tab<long long>(t.step_base);
    a.traj_bstride = (long long)(p->ctx_nt + 1) * NN;
    a.nmax = p->nmax;
    a.slot_lo = p->part_lo;
    a.slot_hi = p->part_hi;
    a.slot0 = p->slot0;
    a.N = N;
    a.scramble = p->order == 2;
}

static int run_prep(pyqed_heom_plan* p, long long step, int tidx) {
    if (!p->ctx_tdep) return 0;
    const int NN = p->N * p->N, M1 = 1 + p->M;
    const TableLayout& t = p->tl;
    prep_ops_kernel<<<1, 256, 0, p->stream>>>(
        p->tab<double2>(t.ops_t), p->tab<double2>(t.ops_base), p->tab<double2>(t.ops_dip),
        p->ctx_use_fs ? p->d_fsys : nullptr, p->ctx_use_fc ? p->d_fcoup : nullptr,
        p->tab<long long>(t.step_base), (int)step, tidx, p->ctx_nt, p->B, M1, NN);
    return post_launch(p, "prep_ops_kernel");
}

// one RK4 stage (0..3) of step `step`; method 1 (Euler) uses stage -1
static int run_stage(pyqed_heom_plan* p, long long step, int stage) {
    StageArgs s;
    fill_stage_args(p, s);
    double2 *Y = p->arr(ARR_Y), *SA = p->arr(ARR_SA), *SB = p->arr(ARR_SB), *ACC = p->arr(ARR_ACC);
    const double dt = p->ctx_dt;
    s.y = Y;
    s.acc = ACC;
    s.local_step = (int)step;
    const int scheme = (stage >= 0 && rk_scheme(p)) ? 1 : 0;
    s.scheme = scheme;
    if (scheme == 1) {
        // difference form: the stage buffers SA, SB, SC (= the accumulator array) are all
        // kept and the last stage combines them, y' = -y/3 + SA/3 + 2SB/3 + SC/3 + dt/6 k4
        switch (stage) {
            case 0: s.yin = Y;  s.yout = SA;  s.a = dt / 2; s.first = 1; break;
            case 1: s.yin = SA; s.yout = SB;  s.a = dt / 2; break;
            case 2: s.yin = SB; s.yout = ACC; s.a = dt; break;
            default: s.yin = ACC; s.acc = SA; s.yout = SB; s.ydst = Y; s.a = 2.0 / dt; s.w = dt / 6; s.last = 1; break;
        }
    } else
    switch (stage) {
        case 0: s.yin = Y;  s.yout = SA; s.a = dt / 2; s.w = dt / 6; s.first = 1; break;
        case 1: s.yin = SA; s.yout = SB; s.a = dt / 2; s.w = dt / 3; break;
        case 2: s.yin = SB; s.yout = SA; s.a = dt;     s.w = dt / 3; break;
        case 3: s.yin = SA; s.ydst = Y;  s.w = dt / 6; s.last = 1; break;
        default: s.yin = Y; s.ydst = SA; s.w = dt; s.first = 1; s.last = 1; break;  // Euler
    }
    if (p->push_ptr && stage >= 0) {
        const int out_arr = scheme == 1 ? (stage == 0 ? ARR_SA : (stage == 1 ? ARR_SB : (stage == 2 ? ARR_ACC : ARR_Y)))
                                        : (stage == 0 ? ARR_SA : (stage == 1 ? ARR_SB : (stage == 2 ? ARR_SA : ARR_Y)));
        s.push_ptr = p->push_ptr;
        s.push_ent = p->push_ent;
        s.peer = p->d_peer;
        s.out_elem_off = (long long)((size_t)out_arr * p->array_bytes / sizeof(double2));
    }
    const int tidx = stage <= 0 ? 0 : (stage == 3 ? 2 : 1);
    if (stage != 2 && run_prep(p, step, tidx)) return 1;  // stages 1 and 2 share t + dt/2
    return launch_stage(p, s, p->ctx_tdep);
}

int pyqed_heom_propagate_begin(pyqed_heom_plan* p, double dt, int64_t nt, const double* fsys,
                               const double* fcoup, double* d_traj) {
    REQUIRE(p && p->built, "propagate: build the hierarchy first");
    REQUIRE(p->d_state && !p->shard.on, "propagate: no full-size state buffer bound (sharded plans use shard_propagate)");
    REQUIRE(nt >= 0, "propagate: nt must be >= 0");
    CU_TRY(cudaSetDevice(p->device));
    const int NN = p->N * p->N;
    p->ctx_use_fs = fsys && p->mu_nonzero;
    p->ctx_use_fc = fcoup && p->qd_nonzero;
    p->ctx_tdep = p->ctx_use_fs || p->ctx_use_fc;
    p->ctx_dt = dt;
    p->ctx_nt = nt;
    p->ctx_traj = (double2*)d_traj;
    if (p->ctx_tdep) {
        if (upload_fields(p, p->ctx_use_fs ? fsys : nullptr, p->ctx_use_fc ? fcoup : nullptr, nt)) return 1;
    }
    CU_TRY(cudaMemsetAsync(p->d_tables + p->tl.step_base, 0, sizeof(long long), p->stream));
    CU_TRY(cudaMemsetAsync(p->d_tables + p->tl.sched, 0, sizeof(unsigned) * 4, p->stream));
    p->sched_total = 0;
    if (p->ctx_traj) {
        record_kernel<<<p->B, 64, 0, p->stream>>>(p->ctx_traj, p->arr(ARR_Y), p->nmax, p->slot0, NN,
                                                  (long long)(nt + 1) * NN, 0);
        if (post_launch(p, "record_kernel")) return 1;
    }
    p->ctx_valid = true;
    return 0;
}

// kernel 7 (heom_stage_sym.cu, heom_packed_propagate): the whole propagation on packed
// Hermitian storage.  Same eligibility as kernel 6, plus one trajectory and the whole
// hierarchy on this GPU (the halo exchange works on full matrices).
static bool packed_eligible(const pyqed_heom_plan* p) {
    // four triangle arrays (256-byte aligned) must fit into the three stage arrays
    const size_t tri = align_up(sizeof(double2) * (size_t)p->nmax * (p->N * (p->N + 1) / 2));
    return (p->kernel == 0 || p->kernel == 7) && p->opt_packed != 0 && p->links2_built && !p->ctx_tdep &&
           p->herm_inputs && p->herm_state &&
           p->opt_herm != 0 && p->single_support && p->opt_sym != 0 && !p->push_ptr && p->B == 1 &&
           p->part_lo == 0 && p->part_hi == p->nmax && rk_scheme(p) && 4 * tri <= 3 * p->array_bytes;
}
static int run_packed(pyqed_heom_plan* p, double dt, int64_t nt) {
    if (ensure_links2(p, 1)) return 1;
    const int sm_count = sm_count_of(p->device);
    const TableLayout& t = p->tl;
    PackedRun r{};
    r.Y = p->arr(ARR_Y);
    r.work = p->arr(ARR_SA);   // the three stage arrays hold the four triangle arrays
    r.work_bytes = 3 * p->array_bytes;
    r.tables.damp = p->tab<double2>(t.damp);
    r.tables.link_ptr = p->tab<int>(t.link_ptr);
    r.tables.links2 = p->tab<int2>(t.links2);
    r.tables.cbase = p->tab<double2>(t.cbase);
    r.tables.kmode = p->tab<int>(t.kmode);
    r.tables.ops = p->tab<double2>(t.ops_base);
    r.tables.traj = p->ctx_traj;
    r.tables.step_base = p->tab<long long>(t.step_base);
    r.tables.slot0 = p->slot0;
    r.tables.scramble = p->order == 2;
    r.tables.nind = p->K;
    r.tables.nmod = p->M;
    r.tables.lmax = p->L;
    r.H = reinterpret_cast<const double*>(p->H.data());
    r.N = p->N;
    r.K = p->K;
    r.M = p->M;
    r.L = p->L;
    r.nmax = p->nmax;
    r.nt = nt;
    r.dt = dt;
    r.hreal = (p->h_real && p->opt_hreal != 0) ? 1 : 0;
    r.warps = p->warps;
    r.sm_count = sm_count;
    r.prefetch = p->opt_prefetch > 0 ? 1 : 0;
    r.stream = p->stream;
    r.sched = p->opt_dynsched != 0 ? p->tab<unsigned>(p->tl.sched) : nullptr;
    r.sched_total = &p->sched_total;
    const char* err = "";
    if (heom_packed_propagate(r, &err)) return fail(std::string("packed propagation (kernel 7): ") + err);
    p->launches += 4 * nt + 2;
    p->packed_steps += nt;
    if (p->debug_sync) {
        cudaError_t e = cudaStreamSynchronize(p->stream);
        if (e != cudaSuccess) return fail(std::string("packed propagation (kernel 7) exec: ") + cudaGetErrorString(e));
    }
    return 0;
}

int pyqed_heom_propagate_stage(pyqed_heom_plan* p, int64_t step, int stage) {
    REQUIRE(p && p->built && p->ctx_valid, "propagate_stage: call propagate_begin first");
    REQUIRE(stage >= 0 && stage <= 3 && step >= 0 && step < p->ctx_nt, "propagate_stage: bad step/stage");
    CU_TRY(cudaSetDevice(p->device));
    return run_stage(p, step, stage);
}

int pyqed_heom_propagate(pyqed_heom_plan* p, double dt, int64_t nt, const double* fsys,
                         const double* fcoup, double* d_traj, int method) {
    REQUIRE(method == 0 || method == 1, "propagate: method must be 0 (rk4) or 1 (euler)");
    if (pyqed_heom_propagate_begin(p, dt, nt, fsys, fcoup, d_traj)) return 1;
    if (method == 0) {
        const int rr = try_resident(p);
        if (rr == 0) return 0;
        if (rr > 0) return 1;
        REQUIRE(p->kernel != 4, "kernel 4 (cluster-resident) is not applicable to this problem");
        const int d9 = try_dataflow_tma(p);
        if (d9 == 0) return 0;
        if (d9 > 0) return 1;
        const int df = try_dataflow(p);
        if (df == 0) return 0;
        if (df > 0) return 1;
        REQUIRE(p->kernel != 8 && p->kernel != 9, "kernels 8 / 9 (persistent dataflow) are not applicable to this problem");
        if (packed_eligible(p)) return run_packed(p, dt, nt);
        for (int64_t i = 0; i < nt; ++i)
            for (int st = 0; st < 4; ++st)
                if (run_stage(p, i, st)) return 1;
    } else {
        // explicit Euler: y' = y + dt F(y) through SA so that no CTA reads a
        // neighbour that another CTA has already advanced
        const size_t bytes = sizeof(double2) * (size_t)p->B * p->nmax * p->N * p->N;
        for (int64_t i = 0; i < nt; ++i) {
            if (run_stage(p, i, -1)) return 1;
            CU_TRY(cudaMemcpyAsync(p->arr(ARR_Y), p->arr(ARR_SA), bytes, cudaMemcpyDeviceToDevice, p->stream));
        }
    }
    return 0;
}

// ---- multi-GPU: owned range and halo rows --------------------------------------
int pyqed_heom_set_partition(pyqed_heom_plan* p, int64_t slot_lo, int64_t slot_hi) {
    REQUIRE(p && p->built, "set_partition: build the hierarchy first");
    REQUIRE(slot_lo >= 0 && slot_lo <= slot_hi && slot_hi <= p->nmax, "set_partition: bad range");
    p->part_lo = slot_lo;
    p->part_hi = slot_hi;
    return 0;
}

// items: slot * 8 + row  (rows == 1, diagonal-Q row items) or slot (rows == 0, whole ADOs)
__global__ void halo_pack_kernel(double2* buf, const double2* arr, const int* items, long long n,
                                 int N, int rows, long long nmax, int unpack) {
    const int NN = N * N, per = rows ? N : NN;
    const long long total = n * per;
    const long long boff = (long long)blockIdx.y * nmax * NN;
    double2* out = buf + (long long)blockIdx.y * total;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long i = e / per;
        const int j = (int)(e - i * per);
        const int it = items[i];
        const long long src = rows ? ((long long)(it >> 3) * NN + (it & 7) * N + j) : ((long long)it * NN + j);
        if (unpack) const_cast<double2*>(arr)[boff + src] = out[e];
        else out[e] = arr[boff + src];
    }
}

// Peer-memory halo: every rank maps the other ranks' state buffers (symmetric
// memory / CUDA IPC) and stores the requested rows straight into their arrays
// over NVLink - no pack buffer, no collective, no unpack on the receiver.
struct PushArgs {
    long long off[17];             // items [off[q], off[q+1]) go to rank q
    unsigned long long peer[16];   // base address of rank q's state buffer
    int world;
};
__global__ void halo_push_kernel(const double2* arr, long long arr_elem_off, const int* items,
                                 long long n, int N, int rows, PushArgs pa) {
    const int NN = N * N, per = rows ? N : NN;
    const long long total = n * per;
    const long long stride = (long long)gridDim.x * blockDim.x;
    constexpr int U = 4;   // independent element copies in flight per thread
    for (long long e0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; e0 < total; e0 += U * stride) {
        long long src[U];
        int q[U];
        double2 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long e = e0 + u * stride;
            src[u] = -1;
            if (e < total) {
                const long long i = e / per;
                const int j = (int)(e - i * per);
                const int it = __ldg(items + i);
                src[u] = rows ? ((long long)(it >> 3) * NN + (it & 7) * N + j) : ((long long)it * NN + j);
                int qq = 0;
                while (qq + 1 < pa.world && i >= pa.off[qq + 1]) ++qq;
                q[u] = qq;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (src[u] >= 0) v[u] = arr[src[u]];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (src[u] >= 0) reinterpret_cast<double2*>(pa.peer[q[u]])[arr_elem_off + src[u]] = v[u];
    }
    // the rows must have landed in the peers' memory before this rank signals the
    // barrier that follows on the stream
    __threadfence_system();
}

int pyqed_heom_halo_push(pyqed_heom_plan* p, int array_id, const int32_t* d_items, int64_t n_items,
                         int row_items, const int64_t* dest_offsets, const uint64_t* peer_state_ptrs,
                         int world) {
    REQUIRE(p && p->built && array_id >= 0 && array_id <= 3, "halo_push: bad argument");
    REQUIRE(world >= 1 && world <= 16 && dest_offsets && peer_state_ptrs, "halo_push: bad peer table");
    REQUIRE(p->B == 1, "halo_push: batch must be 1");
    if (n_items == 0) return 0;
    REQUIRE(d_items, "halo_push: null item list");
    CU_TRY(cudaSetDevice(p->device));
    PushArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.world = world;
    for (int q = 0; q <= world; ++q) pa.off[q] = dest_offsets[q];
    for (int q = 0; q < world; ++q) pa.peer[q] = peer_state_ptrs[q];
    const long long total = n_items * (row_items ? p->N : p->N * p->N);
    const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148 * 16);
    const long long arr_elem_off = (long long)((size_t)array_id * p->array_bytes / sizeof(double2));
    halo_push_kernel<<<grid, 256, 0, p->stream>>>(p->arr(array_id), arr_elem_off, d_items, n_items, p->N,
                                                  row_items, pa);
    return post_launch(p, "halo_push_kernel");
}

int pyqed_heom_set_push_table(pyqed_heom_plan* p, const int32_t* d_push_ptr, const uint8_t* d_push_ent,
                              const uint64_t* peer_state_ptrs, int world) {
    REQUIRE(p && p->built, "set_push_table: build the hierarchy first");
    CU_TRY(cudaSetDevice(p->device));
    if (!d_push_ptr) {   // switch the fused push off
        p->push_ptr = nullptr;
        p->push_ent = nullptr;
        return 0;
    }
    REQUIRE(d_push_ent && peer_state_ptrs && world >= 1 && world <= 16, "set_push_table: bad argument");
    REQUIRE(p->B == 1, "set_push_table: batch must be 1");
    REQUIRE(rk_scheme(p), "set_push_table: the fused push needs the async row kernel (kernel 3)");
    if (!p->d_peer) CU_TRY(cudaMalloc(&p->d_peer, sizeof(unsigned long long) * 16));
    CU_TRY(cudaMemcpyAsync(p->d_peer, peer_state_ptrs, sizeof(unsigned long long) * world,
                           cudaMemcpyHostToDevice, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    p->push_ptr = d_push_ptr;
    p->push_ent = d_push_ent;
    return 0;
}

int pyqed_heom_halo_pack(pyqed_heom_plan* p, int array_id, const int32_t* d_items, int64_t n_items,
                         int row_items, double* d_buf, int unpack) {
    REQUIRE(p && p->built && array_id >= 0 && array_id <= 3, "halo_pack: bad argument");
    if (n_items == 0) return 0;
    REQUIRE(d_items && d_buf, "halo_pack: null buffer");
    CU_TRY(cudaSetDevice(p->device));
    const long long total = n_items * (row_items ? p->N : p->N * p->N);
    dim3 grid((unsigned)std::min<long long>((total + 255) / 256, 148 * 16), p->B);
    halo_pack_kernel<<<grid, 256, 0, p->stream>>>((double2*)d_buf, p->arr(array_id), d_items, n_items,
                                                  p->N, row_items, p->nmax, unpack);
    return post_launch(p, unpack ? "halo_unpack_kernel" : "halo_pack_kernel");
}

// ---- operator action on every ADO (operator_action_ddos, deom.py:945-950) ---------
// rho_n <- A rho_n (side 0) or rho_n A (side 1), for all n and all trajectories;
// the building block of HEOM-space correlation functions
// <A(t) B(0)> = Tr[A G(t) (B rho)] (intent of pyqed/deom.py:921-952).
__global__ void apply_operator_kernel(double2* y, const double2* A, long long nado_total, int N, int side) {
    extern __shared__ double2 smem[];
    const int NN = N * N;
    double2* As = smem;
    double2* rs = smem + NN;
    for (int e = threadIdx.x; e < NN; e += blockDim.x) As[e] = A[e];
    for (long long n = blockIdx.x; n < nado_total; n += gridDim.x) {
        double2* r = y + n * NN;
        __syncthreads();
        for (int e = threadIdx.x; e < NN; e += blockDim.x) rs[e] = r[e];
        __syncthreads();
        for (int e = threadIdx.x; e < NN; e += blockDim.x) {
            const int i = e / N, j = e - i * N;
            double2 v = make_double2(0.0, 0.0);
            if (side == 0)
                for (int l = 0; l < N; ++l) cfma(v, As[i * N + l], rs[l * N + j]);
            else
                for (int l = 0; l < N; ++l) cfma(v, rs[i * N + l], As[l * N + j]);
            r[e] = v;
        }
    }
}

int pyqed_heom_apply_operator(pyqed_heom_plan* p, const double* op_host, int side) {
    REQUIRE(p && p->built && op_host && (side == 0 || side == 1), "apply_operator: bad argument");
    REQUIRE(p->d_state && !p->shard.on, "apply_operator: no full-size state buffer bound");
    CU_TRY(cudaSetDevice(p->device));
    const size_t NN = (size_t)p->N * p->N;
    double2* d_op = nullptr;
    CU_TRY(cudaMalloc(&d_op, sizeof(double2) * NN));
    CU_TRY(cudaMemcpyAsync(d_op, op_host, sizeof(double2) * NN, cudaMemcpyHostToDevice, p->stream));
    const long long total = (long long)p->B * p->nmax;
    const int threads = (int)std::min<size_t>(256, (NN + 31) / 32 * 32);
    const unsigned grid = (unsigned)std::min<long long>(total, 148 * 16);
    apply_operator_kernel<<<grid, threads, sizeof(double2) * 2 * NN, p->stream>>>(p->arr(ARR_Y), d_op, total,
                                                                             p->N, side);
    int rc = post_launch(p, "apply_operator_kernel");
    CU_TRY(cudaStreamSynchronize(p->stream));
    cudaFree(d_op);
    p->herm_state = false;  // A rho is not Hermitian in general
    return rc;
}

// ---- single-exponential chain, explicit Euler with in-place sequential sweep ------
// Restates the Euler `_heom` of pyqed/oqs.py:1808-1875 and the Liouville-space
// `_heom_propagator` (pyqed/HEOM/heom.py:349-413, pyqed/oqs.py:1877-1941):
//   ado[0] += dt (-i[H,ado0] - [S,ado1])
//   ado[n] += dt (-i[H,adon] - [S,ado_{n+1}] - n gamma adon
//                 + n (c_re [S,ado_{n-1}] + i c_im {S,ado_{n-1}})),  n = 1..nado-2
// where ado_{n-1} has already been advanced (Gauss-Seidel order); the last ADO
// never changes.  One CTA per trajectory (the propagator is the batch of the N^2
// unit matrices), one thread per matrix element.  double0 reproduces the second
// update of ado[0] per step that the loop bounds of oqs.py:1930 cause.
__global__ void chain_euler_kernel(double2* ado, const double2* Hg, const double2* Sg, int N, int nado,
                                   double gamma, double c_re, double c_im, double dt, long long nt,
                                   int double0, const double2* eops, int n_e, double2* obs) {
    extern __shared__ double2 smem[];
    const int NN = N * N;
    double2* H = smem;
    double2* S = smem + NN;
    double2* red = smem + 2 * NN;   // [n_e] partial traces
    for (int e = threadIdx.x; e < NN; e += blockDim.x) {
        H[e] = Hg[e];
        S[e] = Sg[e];
    }
    double2* a = ado + (long long)blockIdx.x * nado * NN;
    __syncthreads();
    const int e = threadIdx.x, i = e / N, j = e - i * N;
    const bool act = e < NN;
    for (long long step = 0; step < nt; ++step) {
        const int first = double0 ? -1 : 0;
        for (int nn = first; nn < nado - 1; ++nn) {
            const int n = nn < 0 ? 0 : nn;
            double2 v = make_double2(0.0, 0.0);
            if (act) {
                const double2* A = a + (long long)n * NN;
                const double2* Ap = a + (long long)(n + 1) * NN;
                double2 cH = make_double2(0.0, 0.0), cS = make_double2(0.0, 0.0);
                for (int l = 0; l < N; ++l) {
                    cfma(cH, H[i * N + l], A[l * N + j]);
                    cfms(cH, A[i * N + l], H[l * N + j]);
                    cfma(cS, S[i * N + l], Ap[l * N + j]);
                    cfms(cS, Ap[i * N + l], S[l * N + j]);
                }
                // -i [H, A] - [S, A+]
                v = make_double2(cH.y - cS.x, -cH.x - cS.y);
                if (n >= 1) {
                    const double2* Am = a + (long long)(n - 1) * NN;
                    double2 cm = make_double2(0.0, 0.0), am = make_double2(0.0, 0.0);
                    for (int l = 0; l < N; ++l) {
                        const double2 sa = cmul(S[i * N + l], Am[l * N + j]);
                        const double2 as = cmul(Am[i * N + l], S[l * N + j]);
                        cm.x += sa.x - as.x;
                        cm.y += sa.y - as.y;
                        am.x += sa.x + as.x;
                        am.y += sa.y + as.y;
                    }
                    const double2 own = A[e];
                    // - n gamma A + n (c_re [S,A-] + i c_im {S,A-})
                    v.x += -n * gamma * own.x + n * (c_re * cm.x - c_im * am.y);
                    v.y += -n * gamma * own.y + n * (c_re * cm.y + c_im * am.x);
                }
            }
            __syncthreads();   // everyone has read ado[n] before it is overwritten
            if (act) {
                double2* A = a + (long long)n * NN;
                const double2 own = A[e];
                A[e] = make_double2(own.x + dt * v.x, own.y + dt * v.y);
            }
            __syncthreads();
        }
        if (obs) {   // Tr(e rho_0) after the step (obs, superoperator.py:313)
            for (int o = 0; o < n_e; ++o) {
                if (threadIdx.x == 0) red[o] = make_double2(0.0, 0.0);
            }
            __syncthreads();
            if (threadIdx.x < n_e) {
                double2 s = make_double2(0.0, 0.0);
                const double2* op = eops + (long long)threadIdx.x * NN;
                for (int r = 0; r < N; ++r)
                    for (int c = 0; c < N; ++c) cfma(s, op[r * N + c], a[c * N + r]);
                obs[((long long)blockIdx.x * n_e + threadIdx.x) * nt + step] = s;
            }
            __syncthreads();
        }
    }
}

int pyqed_heom_chain_euler(int device, void* stream, int N, int nado, int batch, const double* H,
                           const double* S, double gamma, double c_re, double c_im, double dt,
                           int64_t nt, int double_update0, double* d_ado, const double* e_ops_host,
                           int n_e, double* d_obs) {
    REQUIRE(N >= 1 && N <= 32 && nado >= 2 && batch >= 1 && nt >= 0 && H && S && d_ado,
            "chain_euler: bad argument (1 <= N <= 32, nado >= 2)");
    REQUIRE(n_e >= 0 && n_e <= 32 && (n_e == 0 || (e_ops_host && d_obs)), "chain_euler: bad observables");
    CU_TRY(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t NN = (size_t)N * N;
    double2* d_ops = nullptr;
    CU_TRY(cudaMalloc(&d_ops, sizeof(double2) * NN * (2 + (size_t)std::max(n_e, 1))));
    CU_TRY(cudaMemcpyAsync(d_ops, H, sizeof(double2) * NN, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(d_ops + NN, S, sizeof(double2) * NN, cudaMemcpyHostToDevice, st));
    if (n_e)
        CU_TRY(cudaMemcpyAsync(d_ops + 2 * NN, e_ops_host, sizeof(double2) * NN * n_e,
                               cudaMemcpyHostToDevice, st));
    const int threads = (int)std::max<size_t>(32, (NN + 31) / 32 * 32);
    const size_t smem = sizeof(double2) * (2 * NN + 32);
    chain_euler_kernel<<<batch, threads, smem, st>>>((double2*)d_ado, d_ops, d_ops + NN, N, nado, gamma,
                                                     c_re, c_im, dt, nt, double_update0,
                                                     n_e ? d_ops + 2 * NN : nullptr, n_e, (double2*)d_obs);
    cudaError_t e = cudaGetLastError();
    cudaError_t e2 = cudaStreamSynchronize(st);
    cudaFree(d_ops);
    if (e != cudaSuccess) return fail(std::string("chain_euler_kernel launch: ") + cudaGetErrorString(e));
    if (e2 != cudaSuccess) return fail(std::string("chain_euler_kernel: ") + cudaGetErrorString(e2));
    return 0;
}

int pyqed_heom_expectation(pyqed_heom_plan* p, const double* d_rho, int64_t npts,
                           const double* ops_host, int n_ops, double* d_out) {
    REQUIRE(p && d_rho && ops_host && d_out && n_ops >= 1 && npts >= 1, "expectation: bad argument");
    CU_TRY(cudaSetDevice(p->device));
    const size_t NN = (size_t)p->N * p->N;
    double2* d_ops = nullptr;
    CU_TRY(cudaMalloc(&d_ops, sizeof(double2) * n_ops * NN));
    CU_TRY(cudaMemcpyAsync(d_ops, ops_host, sizeof(double2) * n_ops * NN, cudaMemcpyHostToDevice,
                           p->stream));
    dim3 grid((unsigned)std::min<long long>(((long long)n_ops * npts + 127) / 128, 4096), p->B);
    expectation_kernel<<<grid, 128, 0, p->stream>>>((double2*)d_out, (const double2*)d_rho, d_ops,
                                                    npts, n_ops, p->N);
    int rc = post_launch(p, "expectation_kernel");
    CU_TRY(cudaStreamSynchronize(p->stream));
    cudaFree(d_ops);
    return rc;
}

int pyqed_heom_memcpy_d2h(pyqed_heom_plan* p, void* host, const void* dev, size_t bytes) {
    REQUIRE(p && host && dev, "memcpy_d2h: null argument");
    CU_TRY(cudaSetDevice(p->device));
    CU_TRY(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    return 0;
}
int pyqed_heom_memcpy_h2d(pyqed_heom_plan* p, void* dev, const void* host, size_t bytes) {
    REQUIRE(p && host && dev, "memcpy_h2d: null argument");
    CU_TRY(cudaSetDevice(p->device));
    CU_TRY(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    return 0;
}
int pyqed_heom_synchronize(pyqed_heom_plan* p) {
    REQUIRE(p, "null plan");
    CU_TRY(cudaSetDevice(p->device));
    CU_TRY(cudaStreamSynchronize(p->stream));
    return 0;
}

int64_t pyqed_heom_launch_count(pyqed_heom_plan* p) { return p ? p->launches : -1; }

int pyqed_heom_stage_timing(pyqed_heom_plan* p, int enable, double* total_ms, int64_t* launches) {
    REQUIRE(p, "null plan");
    CU_TRY(cudaSetDevice(p->device));
    if (total_ms || launches) {
        CU_TRY(cudaStreamSynchronize(p->stream));
        double tot = 0.0;
        for (size_t i = 0; i < p->ev_used; ++i) {
            float ms = 0.f;
            CU_TRY(cudaEventElapsedTime(&ms, p->ev[i].first, p->ev[i].second));
            tot += ms;
        }
        if (total_ms) *total_ms = tot;
        if (launches) *launches = (int64_t)p->ev_used;
    }
    p->ev_used = 0;
    p->timing = enable != 0;
    return 0;
}

}  // extern "C"
