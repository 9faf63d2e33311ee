// sym_emu.cpp - runs kernel 6 (pyqed_b200/csrc/heom_stage_sym.cu) on the CPU
// through tests/_shim/cuda_emu.h.  TEST INFRASTRUCTURE ONLY: built and loaded by
// tests/test_sym_kernel_emu.py, never by the product.
#include "cuda_emu.h"

#include "../../pyqed_b200/csrc/heom_stage_sym.cu"

extern "C" {

// 0: asynchronous copies land when issued, 1: only when waited for (cuda_emu.h)
#ifndef EMU_ONE_SETTER
void emu_set_async_late(int late) { emu::g_async_late = late; }
#endif

// dynamic group schedule (SymLaunch::sched): the counter is never reset between launches
static int g_dyn = 0;
static unsigned g_sched_ctr = 0, g_sched_total = 0;
void emu_set_dynsched(int on) { g_dyn = on; }
long long emu_sched_counter(void) { return g_sched_ctr == g_sched_total ? (long long)g_sched_ctr : -1; }

// One call = nt difference-form RK4 steps (the stage plan of run_stage in
// heom_kernels.cu) on host arrays.  `state` holds the four ADO arrays Y, SA, SB,
// ACC ([nmax][N][N] complex128 each, contiguous); `links` are the (slot, meta)
// records of hier_links_kernel and are converted with the product's converter.
// `parts` = [lo0, hi0, lo1, hi1, ...]: every stage is launched once per owned
// range, like the ranks of a sharded run (all ranges read the same stage input).
int emu_sym_run(int N, int K, int M, int L, long long nmax, const double* H, const double* ops,
                const double* cbase, const int* kmode, const double* damp, const int* link_ptr,
                const int* links, long long nlinks, double* state, double dt, int nt, int hreal,
                int sm_count, int warps, const long long* parts, int nparts, long long slot0,
                int scramble, double* traj, const char** err) {
    static const char* none = "";
    *err = none;
    if (heom_sym_supported(N, K, M, L, err)) return 1;
    std::vector<int2> links2((size_t)std::max(1ll, nlinks));
    if (heom_sym_convert_links(reinterpret_cast<const int2*>(links), links2.data(), nlinks, N, L, 0, nullptr, err))
        return 1;
    const long long NN = (long long)N * N, asz = nmax * NN;
    double2* Y = reinterpret_cast<double2*>(state);
    double2 *SA = Y + asz, *SB = SA + asz, *ACC = SB + asz;
    long long step_base = 0;
    if (traj) std::memcpy(traj, Y + slot0 * NN, sizeof(double2) * NN);
    for (int step = 0; step < nt; ++step) {
        for (int stage = 0; stage < 4; ++stage) {
            // the stage as run_stage (heom_kernels.cu) describes it ...
            StageArgs st{};
            st.damp = reinterpret_cast<const double2*>(damp);
            st.link_ptr = link_ptr;
            st.cbase = reinterpret_cast<const double2*>(cbase);
            st.kmode = kmode;
            st.ops = reinterpret_cast<const double2*>(ops);
            st.step_base = &step_base;
            st.local_step = step;
            st.slot0 = slot0;
            st.scramble = scramble;
            st.nind = K;
            st.nmod = M;
            st.lmax = L;
            st.traj = reinterpret_cast<double2*>(traj);
            st.y = Y;
            st.acc = ACC;
            st.scheme = 1;
            switch (stage) {
                case 0: st.yin = Y;  st.yout = SA;  st.a = dt / 2; st.first = 1; break;
                case 1: st.yin = SA; st.yout = SB;  st.a = dt / 2; break;
                case 2: st.yin = SB; st.yout = ACC; st.a = dt; break;
                default:
                    st.yin = ACC; st.acc = SA; st.yout = SB; st.ydst = Y;
                    st.a = 2.0 / dt; st.w = dt / 6; st.last = 1;
                    break;
            }
            // ... and the product's mapping of it onto kernel 6
            SymLaunch s{};
            s.a = sym_args_from_stage(st, links2.data());
            s.stage = sym_stage_kind(st);
            s.H = H;
            s.N = N; s.K = K; s.M = M; s.L = L; s.B = 1;
            s.hreal = hreal;
            s.warps = warps;
            s.sm_count = sm_count;
            s.batch_elems = asz;
            s.traj_bstride = (long long)(nt + 1) * NN;
            if (g_dyn) {
                s.sched = &g_sched_ctr;
                s.sched_total = &g_sched_total;
            }
            for (int p = 0; p < nparts; ++p) {
                s.part_lo = parts[2 * p];
                s.part_hi = parts[2 * p + 1];
                if (s.part_hi <= s.part_lo) continue;
                if (heom_sym_launch(s, err)) return 1;
            }
        }
    }
    return 0;
}

// links (slot, meta) -> links2 with the product's converter
int emu_convert_links(const int* links, int* links2, long long nlinks, int N, int L, int packed, const char** err) {
    static const char* none = "";
    *err = none;
    return heom_sym_convert_links(reinterpret_cast<const int2*>(links), reinterpret_cast<int2*>(links2), nlinks, N, L,
                                  packed, nullptr, err);
}

// One stage of one rank of a sharded run with rank-local arrays (heom_shard.cu in miniature): the
// arrays hold the rank's n_own ADOs (full matrices or upper triangles) followed by its pool of
// halo rows; `links2` is the rank's localized link table (local slots / pool rows), `link_ptr`
// the CSR offsets of its owned slots.  With push tables the PUSH instantiation stores the rows
// other ranks read into their pools (peer[q] = base address of rank q's state).
int emu_sym_stage(int N, int K, int M, int L, const double* H, const double* ops, const double* cbase,
                  const int* kmode, const double* damp, const int* link_ptr, const int* links2,
                  const double* yin, const double* y, const double* s1, const double* s2, double* out, double a,
                  double w, int stage_kind, int hreal, int sm_count, int warps, long long n_own, int packed,
                  long long pool_off, const int* push_ptr, const int* push_ent,
                  const unsigned long long* peer, long long out_elem_off, const char** err) {
    static const char* none = "";
    *err = none;
    long long step_base = 0;
    SymLaunch s{};
    s.a.yin = reinterpret_cast<const double2*>(yin);
    s.a.y = reinterpret_cast<const double2*>(y);
    s.a.s1 = reinterpret_cast<const double2*>(s1);
    s.a.s2 = reinterpret_cast<const double2*>(s2);
    s.a.out = reinterpret_cast<double2*>(out);
    s.a.damp = reinterpret_cast<const double2*>(damp);
    s.a.link_ptr = link_ptr;
    s.a.links2 = reinterpret_cast<const int2*>(links2);
    s.a.cbase = reinterpret_cast<const double2*>(cbase);
    s.a.kmode = kmode;
    s.a.ops = reinterpret_cast<const double2*>(ops);
    s.a.step_base = &step_base;
    s.a.slot0 = -1;
    s.a.a = a;
    s.a.w = w;
    s.a.nind = K;
    s.a.nmod = M;
    s.a.lmax = L;
    s.a.pool_off = (unsigned)pool_off;
    s.a.push_ptr = push_ptr;     // null: no fused push
    s.a.push_ent = reinterpret_cast<const int2*>(push_ent);
    for (int q = 0; q < 16; ++q) s.a.peer[q] = 0;
    if (peer) s.a.peer[0] = peer[0], s.a.peer[1] = peer[1];   // (the tests use two ranks)
    s.a.out_elem_off = out_elem_off;
    s.push = push_ptr ? 1 : 0;
    s.packed = packed;
    s.H = H;
    s.N = N; s.K = K; s.M = M; s.L = L; s.B = 1;
    s.stage = stage_kind;
    s.hreal = hreal;
    s.warps = warps;
    s.sm_count = sm_count;
    s.part_lo = 0;
    s.part_hi = n_own;
    s.batch_elems = 0;
    if (g_dyn) {
        s.sched = &g_sched_ctr;
        s.sched_total = &g_sched_total;
    }
    return heom_sym_launch(s, err);
}

// Kernel 7: the product's packed-storage driver (heom_packed_propagate) on host arrays.
// `Y` is the full state [nmax][N][N]; the work region is filled with NaNs first.
int emu_packed_run(int N, int K, int M, int L, long long nmax, const double* H, const double* ops,
                   const double* cbase, const int* kmode, const double* damp, const int* link_ptr,
                   const int* links, long long nlinks, double* Y, double dt, int nt, int hreal,
                   int sm_count, int warps, long long slot0, int scramble, double* traj, int prefetch,
                   const char** err) {
    static const char* none = "";
    *err = none;
    if (heom_sym_supported(N, K, M, L, err)) return 1;
    std::vector<int2> links2((size_t)std::max(1ll, nlinks));
    if (heom_sym_convert_links(reinterpret_cast<const int2*>(links), links2.data(), nlinks, N, L, 1, nullptr, err))
        return 1;
    const long long NN = (long long)N * N, PK = (long long)N * (N + 1) / 2;
    const size_t arr = ((size_t)(nmax * PK) * sizeof(double2) + 255) / 256 * 256;
    std::vector<char> work(4 * arr, (char)0xff);
    long long step_base = 0;
    if (traj) std::memcpy(traj, reinterpret_cast<double2*>(Y) + slot0 * NN, sizeof(double2) * NN);
    PackedRun r{};
    r.Y = reinterpret_cast<double2*>(Y);
    r.work = reinterpret_cast<double2*>(work.data());
    r.work_bytes = work.size();
    r.tables.damp = reinterpret_cast<const double2*>(damp);
    r.tables.link_ptr = link_ptr;
    r.tables.links2 = links2.data();
    r.tables.cbase = reinterpret_cast<const double2*>(cbase);
    r.tables.kmode = kmode;
    r.tables.ops = reinterpret_cast<const double2*>(ops);
    r.tables.traj = reinterpret_cast<double2*>(traj);
    r.tables.step_base = &step_base;
    r.tables.slot0 = slot0;
    r.tables.scramble = scramble;
    r.tables.nind = K;
    r.tables.nmod = M;
    r.tables.lmax = L;
    r.H = H;
    r.N = N; r.K = K; r.M = M; r.L = L;
    r.nmax = nmax;
    r.nt = nt;
    r.dt = dt;
    r.hreal = hreal;
    r.warps = warps;
    r.sm_count = sm_count;
    r.prefetch = prefetch;
    r.stream = nullptr;
    if (g_dyn) {
        r.sched = &g_sched_ctr;
        r.sched_total = &g_sched_total;
    }
    return heom_packed_propagate(r, err);
}

}  // extern "C"
