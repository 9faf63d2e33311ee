// heom_plan.cuh - the plan object behind the C ABI (include/pyqed_heom.h) and the plumbing
// shared by the translation units of the library: heom_kernels.cu (C ABI, hierarchy builder,
// stage dispatch), heom_inst.cu (compiled once per system size N: the launchers of the
// N-templated stage kernels) and heom_stage_sym.cu (kernels 6 / 7).
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>

#include <algorithm>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pyqed_heom.h"
#include "heom_core.cuh"
#include "heom_device.cuh"
#include "heom_stage_sym.cuh"

using heom::Pascal;


// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
// one error string per host thread, shared by every translation unit of the library
extern thread_local std::string g_heom_err;
inline int fail(const std::string& msg) {
    g_heom_err = msg;
    return 1;
}
#define CU_TRY(expr)                                                                    \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess)                                                          \
            return fail(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +     \
                        __FILE__ + ":" + std::to_string(__LINE__) + ")");               \
    } while (0)
#define REQUIRE(cond, msg)               \
    do {                                 \
        if (!(cond)) return fail(msg);   \
    } while (0)

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------
struct TableLayout {
    size_t pascal, keys, id_of_slot, slot_of_id, damp, link_ptr, links, coef, ops_base, ops_dip,
        ops_t, row_ptr, row_idx, col_ptr, col_idx, supp, cbase, kmode, lex2slot, slot2lex, step_base, sched, links2, total;
};

struct pyqed_heom_plan {
    int device = 0, N = 0, K = 0, M = 0, L = 0, B = 1, order = 0;
    long long nmax = 0, nlinks = 0;
    int side = 0;
    std::vector<long long> pascal;
    std::vector<std::complex<double>> H, mu, Q, Qd, expn, etal, etar, etaa;
    std::vector<long long> mode;
    bool have_sys = false, have_coup = false, have_bath = false, bound = false, built = false;
    bool mu_nonzero = false, qd_nonzero = false;
    bool q_diagonal = false;     // every Q_m (and its dipole) is diagonal
    bool herm_inputs = false;    // operators/bath keep every ADO Hermitian
    bool herm_state = false;     // ... and so is the state that was loaded
    bool use_qdiag = false;      // resolved at build time from the options below
    bool h_real = false;         // H and mu have no imaginary part
    std::vector<int> r0mode;     // first row with a non-zero diagonal entry, per mode
    int opt_qdiag = -1, opt_herm = -1, opt_hreal = -1, opt_resident = -1;  // -1 auto, 0 off, 1 on
    int opt_sym = -1;            // async kernel: Hermitian-symmetric shortcuts (0 off)
    bool single_support = false; // every Q_m has exactly one non-zero diagonal entry
    int opt_rk13 = -1;  // difference-form RK4 in the async kernel (-1/1 on, 0 off)
    int opt_prefetch = 0;  // kernel 7: double-buffered streamed tiles, fetched one group ahead (1 on)
    int opt_packed = -1;   // packed Hermitian storage (kernel 7) where eligible (-1/1 on, 0 off)
    int opt_dynsched = -1;  // kernels 6 / 7: groups handed out by a global counter (-1/1 on, 0 static stride)
    unsigned sched_total = 0;  // host copy of the schedule counter (tl.sched), see SymArgs::sched_base
    long long resident_launches = 0;
    long long sym_launches = 0;  // stage launches that went to kernel 6
    long long packed_steps = 0;  // RK4 steps done by kernel 7 (packed Hermitian storage)
    long long dataflow_launches = 0;  // propagations done by kernel 8 (one persistent launch each)
    long long dataflow_tma_launches = 0;  // ... of which by kernel 9 (Hermitian, one CTA per ADO, TMA staging)
    long long dataflow_dense_launches = 0;  // ... of which with H as a dense matrix in parameter space
    int opt_dataflow_tma = -1;        // kernel 9 where eligible (-1/1 on, 0 off: kernel 8)
    unsigned* d_df9 = nullptr;        // kernel 9: stage counters, control block, work order
    size_t df9_cap = 0;
    unsigned* d_flags = nullptr;      // kernel 8: per-ADO stage counters
    size_t flags_cap = 0;
    bool links2_built = false;
    int links2_mode = -1;        // what links2 is resolved for: 0 full matrices, 1 packed storage, -1 neither
    size_t bound_table_bytes = 0;
    int resident_kind = 0;  // 4 or 5: which resident kernel ran last
    TableLayout tl{};
    char* d_tables = nullptr;
    char* d_state = nullptr;
    size_t array_bytes = 0;  // one [B][nmax][N][N] array, aligned
    cudaStream_t stream = nullptr;
    long long slot0 = 0;  // storage slot of ADO id 0
    long long part_lo = 0, part_hi = 0;  // owned slot range (multi-GPU); [0, nmax) by default
    // tuning
    int kernel = 0, warps = 0, use_graph = 0;
    // accounting
    long long launches = 0;
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
    size_t ev_used = 0;
    // internal small device buffers for the field tables of one propagate call
    double* d_fsys = nullptr;
    double* d_fcoup = nullptr;
    size_t field_cap = 0;
    bool debug_sync = false;
    // fused peer push (multi-GPU)
    const int* push_ptr = nullptr;
    const unsigned char* push_ent = nullptr;
    unsigned long long* d_peer = nullptr;
    // sharded run with rank-local arrays (heom_shard.cu): own ADOs [lo, hi) followed by a pool of
    // halo rows; only kernels 6 / 7
    struct Shard {
        bool on = false, packed = false, device_barrier = false;
        int rank = 0, world = 1;
        long long lo = 0, hi = 0, n_own_max = 0, pool_max = 0;
        size_t arr_full = 0, arr_packed = 0;     // array strides (double2 elements), identical on all ranks
        const int* push_ptr = nullptr;           // caller-owned device tables (CSR over owned slots)
        const int2* push_ent = nullptr;
        unsigned long long* d_peer = nullptr;    // device copy of the peers' state buffer addresses
        unsigned long long peer_state[16] = {0}; // host copy of the peers' state buffer addresses
        unsigned long long peer_flags[16] = {0}; // address of every rank's flag block (world uint32 + error word)
        unsigned epoch = 0;                      // barriers done so far
        long long pushed_rows = 0;               // rows per stage this rank stores into its peers
    } shard;
    // context of the propagation in progress (propagate_begin)
    bool ctx_valid = false, ctx_tdep = false, ctx_use_fs = false, ctx_use_fc = false;
    double ctx_dt = 0.0;
    long long ctx_nt = 0;
    double2* ctx_traj = nullptr;

    double2* arr(int which) const { return (double2*)(d_state + (size_t)which * array_bytes); }
    template <typename T> T* tab(size_t off) const { return (T*)(d_tables + off); }
};
enum { ARR_Y = 0, ARR_SA = 1, ARR_SB = 2, ARR_ACC = 3 };

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
// 13-pass difference form of RK4: only the async row kernel implements it
// Stage kernel of this plan: the explicit choice, else the async row kernel for
// diagonal coupling (its neighbour rows use 32-bit element offsets, so only while
// nmax N^2 < 2^32), the plain row kernel for other N <= 8, the generic kernel above.
inline int stage_kernel_of(const pyqed_heom_plan* p) {
    // 6 = kernel 3's scheme and buffers; launch_stage hands the eligible stages to kernel 6.
    // 7 = whole propagations on packed Hermitian storage where eligible (pyqed_heom_propagate),
    //     kernel 3 otherwise
    // 8 = persistent dataflow propagation (heom_dataflow.cuh) where eligible, the generic kernel otherwise
    if (p->kernel && p->kernel != 4 && p->kernel != 6 && p->kernel != 7 && p->kernel != 8 && p->kernel != 9) return p->kernel;
    if (p->N > 8 || p->N < 2 || p->kernel == 8 || p->kernel == 9) return 2;   // (the row kernels need 2 <= N <= 8)
    const bool fits32 = (unsigned long long)p->nmax * p->N * p->N < (1ull << 32);
    return (p->use_qdiag && p->opt_rk13 != 0 && fits32) ? 3 : 1;
}
inline bool rk_scheme(const pyqed_heom_plan* p) { return stage_kernel_of(p) == 3; }

inline int post_launch(pyqed_heom_plan* p, const char* what) {
    p->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(std::string(what) + " launch: " + cudaGetErrorString(e));
    if (p->debug_sync) {
        e = cudaStreamSynchronize(p->stream);
        if (e != cudaSuccess) return fail(std::string(what) + " exec: " + cudaGetErrorString(e));
    }
    return 0;
}

// ---------------------------------------------------------------------------
// launchers of the N-templated kernels, one object file per N (heom_inst.cu)
// ---------------------------------------------------------------------------
struct ResidentArgs;
struct ResidentConfig {
    int cluster = 0, warps = 0, apc = 0;
    size_t smem = 0;
};
#define HEOM_DECLARE_N(n)                                                                              \
    int heom_launch_async_##n(pyqed_heom_plan* p, const StageArgs& a, int sm_count, bool tdep, bool hreal); \
    int heom_launch_rows_##n(pyqed_heom_plan* p, const StageArgs& a, int sm_count, bool tdep, bool qdiag);  \
    bool heom_resident_fits_##n(const pyqed_heom_plan* p, ResidentConfig& rc);                         \
    int heom_launch_resident_##n(pyqed_heom_plan* p, const ResidentArgs& ra, ResidentConfig rc, bool hreal); \
    int heom_launch_resident_elem_##n(pyqed_heom_plan* p, const ResidentArgs& ra);
HEOM_DECLARE_N(2) HEOM_DECLARE_N(3) HEOM_DECLARE_N(4) HEOM_DECLARE_N(5) HEOM_DECLARE_N(6) HEOM_DECLARE_N(7) HEOM_DECLARE_N(8)
#undef HEOM_DECLARE_N

// cudaFuncSetAttribute is per device: remember per (kernel instantiation, device) that the
// opt-in to > 48 KB of dynamic shared memory was made
struct PerDeviceOnce {
    unsigned long long done = 0;
    bool need(int device) {
        const unsigned long long bit = 1ull << (device & 63);
        if (done & bit) return false;
        done |= bit;
        return true;
    }
};
inline int sm_count_of(int device) {
    static int cache[64] = {0};
    int& c = cache[device & 63];
    if (!c) cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, device);
    return c;
}
