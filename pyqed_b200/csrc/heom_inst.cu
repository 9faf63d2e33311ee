// heom_inst.cu - launchers of the N-templated stage kernels (kernels 1, 3, 4, 5), compiled
// once per system size: build.py passes -DHEOM_INST_N=n for n = 2..8, so the ~110 kernel
// instantiations are spread over seven object files that compile side by side.
#include "heom_plan.cuh"
#include "heom_stage_async.cuh"
#include "heom_stage_rows.cuh"
#include "heom_resident.cuh"

#ifndef HEOM_INST_N
#error "compile with -DHEOM_INST_N=<2..8>"
#endif

namespace {
template <int N, bool TDEP, bool QDIAG>
static int launch_rows(pyqed_heom_plan* p, const StageArgs& a, int sm_count) {
    constexpr int APW = 32 / N, LD = (N % 2 == 0) ? N + 1 : N, TILE = APW * N * LD;
    StageArgs args = a;
    args.ngroups = (p->part_hi - p->part_lo + APW - 1) / APW;
    int warps = p->warps > 0 ? std::min(p->warps, 8) : 8;
    if (p->warps <= 0) {
        // small hierarchies: prefer more CTAs over fuller CTAs
        while (warps > 1 && args.ngroups * p->B < (long long)warps * sm_count * 2) warps >>= 1;
    }
    size_t smem = sizeof(double2) * ((TDEP ? N * N : 0) + (size_t)warps * 2 * TILE);
    if (QDIAG) smem += sizeof(double2) * (2 * (size_t)args.ncoef + (size_t)p->M * N) +
                       align_up((size_t)p->M * (2 * N + 1), 16);
    REQUIRE(smem <= 200 * 1024, "shared-memory tables too large for the row kernel");
    static PerDeviceOnce attr;
    if (attr.need(p->device))
        CU_TRY(cudaFuncSetAttribute(stage_rows_kernel<N, TDEP, QDIAG>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    long long ctas = (args.ngroups + warps - 1) / warps;
    const long long cap = (long long)sm_count * 16;
    dim3 grid((unsigned)std::min(ctas, cap), p->B);
    HParam<N> hp;
    for (int e = 0; e < N * N; ++e) hp.v[e] = make_double2(p->H[e].real(), p->H[e].imag());
    stage_rows_kernel<N, TDEP, QDIAG><<<grid, warps * 32, smem, p->stream>>>(args, hp);
    return post_launch(p, "stage_rows_kernel");
}

template <int N, bool TDEP, bool HREAL, bool PUSH, bool SYM>
static int launch_async(pyqed_heom_plan* p, const StageArgs& a, int sm_count) {
    constexpr int NN = N * N, APW = 32 / N, LD = (N % 2 == 0) ? N + 1 : N, TILE = APW * N * LD;
    constexpr int FLAT = APW * NN, PERWARP = 2 * TILE + 3 * FLAT + 2;
    StageArgs args = a;
    args.ngroups = (p->part_hi - p->part_lo + APW - 1) / APW;
    const AsyncTables T = async_tables(N, p->K, p->M, p->L, TDEP);
    const size_t table_bytes = sizeof(double2) * T.warp0 + T.bytes_tail;
    const size_t per_warp = sizeof(double2) * (a.last ? PERWARP : PERWARP - FLAT);
    const size_t budget = 227 * 1024;
    REQUIRE(table_bytes + per_warp <= budget, "shared-memory tables too large for the async row kernel");
    int maxw = (int)std::min<size_t>(ASYNC_MAX_THREADS / 32, (budget - table_bytes) / per_warp);
    int warps = p->warps > 0 ? std::min(p->warps, maxw) : maxw;
    if (p->warps <= 0) {
        // small hierarchies: spread the groups over all SMs first
        const long long per_sm = (args.ngroups + sm_count - 1) / sm_count;
        warps = (int)std::max<long long>(1, std::min<long long>(maxw, per_sm));
    }
    REQUIRE((unsigned long long)p->nmax * NN < (1ull << 32),
            "hierarchy too large for the async row kernel's 32-bit element offsets (use kernel 1)");
    const size_t smem = table_bytes + per_warp * warps;
    static PerDeviceOnce attr;
    if (attr.need(p->device))
        CU_TRY(cudaFuncSetAttribute(stage_rows_async_kernel<N, TDEP, HREAL, PUSH, SYM>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    const long long ctas = (args.ngroups + warps - 1) / warps;
    dim3 grid((unsigned)std::min<long long>(ctas, sm_count), 1);
    HParam<N> hp;
    for (int e = 0; e < NN; ++e) hp.v[e] = make_double2(p->H[e].real(), p->H[e].imag());
    // one launch per trajectory of the batch, each with its own array / operator /
    // trajectory pointers: the kernel then carries no batch offset
    const long long boff = p->nmax * NN;
    for (int b = 0; b < p->B; ++b) {
        stage_rows_async_kernel<N, TDEP, HREAL, PUSH, SYM><<<grid, warps * 32, smem, p->stream>>>(args, hp);
        int rc = post_launch(p, "stage_rows_async_kernel");
        if (rc) return rc;
        args.yin += boff;
        args.y += boff;
        args.acc += boff;
        args.yout += boff;
        args.ydst += boff;
        args.ops += args.ops_bstride;
        if (args.traj) args.traj += args.traj_bstride;
        args.out_elem_off += boff;
    }
    return 0;
}

template <int N, bool PUSH, bool SYM>
static int launch_async_p(pyqed_heom_plan* p, const StageArgs& a, int sm_count, bool tdep, bool hreal) {
    if (tdep) return hreal ? launch_async<N, true, true, PUSH, SYM>(p, a, sm_count) : launch_async<N, true, false, PUSH, SYM>(p, a, sm_count);
    return hreal ? launch_async<N, false, true, PUSH, SYM>(p, a, sm_count) : launch_async<N, false, false, PUSH, SYM>(p, a, sm_count);
}
template <int N>
static int launch_async_n(pyqed_heom_plan* p, const StageArgs& a, int sm_count, bool tdep, bool hreal) {
    const bool sym = a.herm && p->single_support && p->opt_sym != 0;
    if (a.push_ptr)
        return sym ? launch_async_p<N, true, true>(p, a, sm_count, tdep, hreal)
                   : launch_async_p<N, true, false>(p, a, sm_count, tdep, hreal);
    return sym ? launch_async_p<N, false, true>(p, a, sm_count, tdep, hreal)
               : launch_async_p<N, false, false>(p, a, sm_count, tdep, hreal);
}

template <int N>
static int launch_rows_n(pyqed_heom_plan* p, const StageArgs& a, int sm_count, bool tdep, bool qdiag) {
    if (tdep) return qdiag ? launch_rows<N, true, true>(p, a, sm_count) : launch_rows<N, true, false>(p, a, sm_count);
    return qdiag ? launch_rows<N, false, true>(p, a, sm_count) : launch_rows<N, false, false>(p, a, sm_count);
}


// ---- cluster-resident propagation (kernel 4) -----------------------------------

template <int N>
static bool resident_fits(const pyqed_heom_plan* p, ResidentConfig& rc) {
    constexpr int APW = 32 / N, LD = (N % 2 == 0) ? N + 1 : N;
    const AsyncTables T = async_tables(N, p->K, p->M, p->L, true);
    const size_t table_bytes = sizeof(double2) * T.warp0 + T.bytes_tail;
    const int maxlinks = std::max(1, std::min(p->L, p->K) + p->K);
    const size_t per_warp = sizeof(double2) * 4 * APW * N * LD + (size_t)16 * APW * maxlinks;
    const size_t budget = 227 * 1024;
    if (table_bytes + per_warp > budget) return false;
    const int maxw = (int)std::min<size_t>(16, (budget - table_bytes) / per_warp);
    const long long groups = (p->nmax + APW - 1) / APW;
    int cs_min = 1;
    while (cs_min <= 16 && (groups + cs_min - 1) / cs_min > maxw) cs_min <<= 1;
    if (cs_min > 16) return false;
    int cs = cs_min;
    if (p->B <= 8)  // few trajectories: spread one hierarchy over as many SMs as a cluster allows
        while (cs < 16 && cs < groups) cs <<= 1;
    rc.cluster = cs;
    rc.warps = (int)((groups + cs - 1) / cs);
    rc.apc = rc.warps * APW;
    rc.smem = table_bytes + per_warp * rc.warps;
    return true;
}

template <int N, bool HREAL>
static int launch_resident_t(pyqed_heom_plan* p, const ResidentArgs& ra_in, ResidentConfig rc) {
    auto kern = resident_cluster_kernel<N, HREAL>;
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    HParam<N> hp;
    for (int e = 0; e < N * N; ++e) hp.v[e] = make_double2(p->H[e].real(), p->H[e].imag());
    for (;;) {
        ResidentArgs ra = ra_in;
        ra.apc = rc.apc;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(rc.cluster * p->B));
        cfg.blockDim = dim3((unsigned)(rc.warps * 32));
        cfg.dynamicSmemBytes = rc.smem;
        cfg.stream = p->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)rc.cluster;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int nclusters = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg);
        if (e == cudaSuccess && nclusters >= 1) {
            CU_TRY(cudaLaunchKernelEx(&cfg, kern, ra, hp));
            return post_launch(p, "resident_cluster_kernel");
        }
        cudaGetLastError();
        // this cluster shape cannot be co-scheduled: halve the cluster if the
        // hierarchy still fits, otherwise report that kernel 4 is unavailable
        constexpr int APW = 32 / N, LD = (N % 2 == 0) ? N + 1 : N;
        const long long groups = (p->nmax + APW - 1) / APW;
        const int cs = rc.cluster / 2;
        if (cs < 1) return -1;
        const int warps = (int)((groups + cs - 1) / cs);
        const int maxlinks = std::max(1, std::min(p->L, p->K) + p->K);
        const size_t per_warp = sizeof(double2) * 4 * APW * N * LD + (size_t)16 * APW * maxlinks;
        const size_t smem = rc.smem - per_warp * rc.warps + per_warp * warps;
        if (warps > 16 || smem > 227 * 1024) return -1;
        rc.cluster = cs;
        rc.warps = warps;
        rc.apc = warps * APW;
        rc.smem = smem;
    }
}

template <int N>
static int launch_resident_elem(pyqed_heom_plan* p, const ResidentArgs& ra_in) {
    constexpr int NN = N * N;
    const int maxlinks = ra_in.maxlinks;
    const size_t table_bytes = sizeof(double2) * (NN + (size_t)p->M * N + 4 * (size_t)p->K) +
                               sizeof(double) * ((p->L + 2) & ~1);
    const size_t per_ado = sizeof(double2) * 4 * NN + (size_t)16 * maxlinks + 32;
    const size_t budget = 227 * 1024;
    if (table_bytes + per_ado > budget) return -1;
    const long long cap = (long long)((budget - table_bytes) / per_ado);   // ADOs per CTA
    int cs = 1;
    while (cs <= 16 && (p->nmax + cs - 1) / cs > cap) cs <<= 1;
    if (cs > 16) return -1;
    if (p->B <= 8)
        while (cs < 16 && cs < p->nmax) cs <<= 1;
    auto kern = resident_elem_kernel<N>;
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    for (; cs >= 1; cs >>= 1) {
        const long long apc = (p->nmax + cs - 1) / cs;
        if (apc > cap) return -1;
        ResidentArgs ra = ra_in;
        ra.apc = (int)apc;
        const int warps = (int)std::min<long long>(16, apc);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(cs * p->B));
        cfg.blockDim = dim3((unsigned)(warps * 32));
        cfg.dynamicSmemBytes = table_bytes + per_ado * apc;
        cfg.stream = p->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)cs;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int nclusters = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg);
        if (e == cudaSuccess && nclusters >= 1) {
            CU_TRY(cudaLaunchKernelEx(&cfg, kern, ra));
            return post_launch(p, "resident_elem_kernel");
        }
        cudaGetLastError();
    }
    return -1;
}

}  // namespace

#define HEOM_CAT2(a, b) a##b
#define HEOM_CAT(a, b) HEOM_CAT2(a, b)
int HEOM_CAT(heom_launch_async_, HEOM_INST_N)(pyqed_heom_plan* p, const StageArgs& a, int sm_count, bool tdep, bool hreal) {
    return launch_async_n<HEOM_INST_N>(p, a, sm_count, tdep, hreal);
}
int HEOM_CAT(heom_launch_rows_, HEOM_INST_N)(pyqed_heom_plan* p, const StageArgs& a, int sm_count, bool tdep, bool qdiag) {
    return launch_rows_n<HEOM_INST_N>(p, a, sm_count, tdep, qdiag);
}
bool HEOM_CAT(heom_resident_fits_, HEOM_INST_N)(const pyqed_heom_plan* p, ResidentConfig& rc) {
    return resident_fits<HEOM_INST_N>(p, rc);
}
int HEOM_CAT(heom_launch_resident_, HEOM_INST_N)(pyqed_heom_plan* p, const ResidentArgs& ra, ResidentConfig rc, bool hreal) {
    return hreal ? launch_resident_t<HEOM_INST_N, true>(p, ra, rc) : launch_resident_t<HEOM_INST_N, false>(p, ra, rc);
}
int HEOM_CAT(heom_launch_resident_elem_, HEOM_INST_N)(pyqed_heom_plan* p, const ResidentArgs& ra) {
    return launch_resident_elem<HEOM_INST_N>(p, ra);
}
