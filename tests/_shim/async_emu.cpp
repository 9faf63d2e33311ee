// async_emu.cpp - runs kernel 3 (pyqed_b200/csrc/heom_stage_async.cuh, the default stage
// kernel for N <= 8 with diagonal coupling operators) on the CPU through
// tests/_shim/cuda_emu.h.  TEST INFRASTRUCTURE ONLY: built and loaded by
// tests/test_async_kernel_emu.py, never by the product.
#include "cuda_emu.h"

#include "../../pyqed_b200/csrc/heom_stage_async.cuh"

namespace {

// launch_async of heom_kernels.cu, without the batch loop
template <int N, bool HREAL, bool SYM>
void launch(StageArgs a, const double* H, int K, int M, int L, int warps_req, int sm_count) {
    constexpr int NN = N * N, APW = 32 / N, LD = (N % 2 == 0) ? N + 1 : N, TILE = APW * N * LD;
    constexpr int FLAT = APW * NN, PERWARP = 2 * TILE + 3 * FLAT + 2;
    a.ngroups = (a.slot_hi - a.slot_lo + APW - 1) / APW;
    const AsyncTables T = async_tables(N, K, M, L, false);
    const size_t table_bytes = sizeof(double2) * T.warp0 + T.bytes_tail;
    const size_t per_warp = sizeof(double2) * (a.last ? PERWARP : PERWARP - FLAT);
    const int warps = warps_req;
    const size_t smem = table_bytes + per_warp * warps;
    const long long ctas = (a.ngroups + warps - 1) / warps;
    const unsigned grid = (unsigned)std::min<long long>(ctas, sm_count);
    HParam<N> hp;
    for (int e = 0; e < NN; ++e) hp.v[e] = make_double2(H[2 * e], H[2 * e + 1]);
    auto kern = stage_rows_async_kernel<N, false, HREAL, false, SYM>;
    HEOM_LAUNCH(kern, grid, warps * 32, smem, nullptr, a, hp);
}

template <int N>
void launch_n(const StageArgs& a, const double* H, int K, int M, int L, int warps, int sm_count, int hreal,
              int sym) {
    if (hreal) {
        if (sym) launch<N, true, true>(a, H, K, M, L, warps, sm_count);
        else launch<N, true, false>(a, H, K, M, L, warps, sm_count);
    } else {
        if (sym) launch<N, false, true>(a, H, K, M, L, warps, sm_count);
        else launch<N, false, false>(a, H, K, M, L, warps, sm_count);
    }
}

}  // namespace

extern "C" {

#ifndef EMU_ONE_SETTER
void emu_set_async_late(int late) { emu::g_async_late = late; }
#endif

// nt difference-form RK4 steps with the stage plan of run_stage (heom_kernels.cu, scheme 1)
// on host arrays; `state` = Y, SA, SB, ACC.  `parts`: owned slot ranges, one launch each.
int emu_async_run(int N, int K, int M, int L, long long nmax, const double* H, const double* ops,
                  const double* cbase, const int* kmode, const unsigned char* supp, const double* damp,
                  const int* link_ptr, const int* links, double* state, double dt, int nt, int hreal,
                  int sym, int herm, int sm_count, int warps, const long long* parts, int nparts,
                  long long slot0, int scramble, double* traj) {
    const long long NN = (long long)N * N, asz = nmax * NN;
    double2* Y = reinterpret_cast<double2*>(state);
    double2 *SA = Y + asz, *SB = SA + asz, *ACC = SB + asz;
    long long step_base = 0;
    if (traj) std::memcpy(traj, Y + slot0 * NN, sizeof(double2) * NN);
    for (int step = 0; step < nt; ++step) {
        for (int stage = 0; stage < 4; ++stage) {
            StageArgs s{};
            s.damp = reinterpret_cast<const double2*>(damp);
            s.link_ptr = link_ptr;
            s.links = reinterpret_cast<const int2*>(links);
            s.ops = reinterpret_cast<const double2*>(ops);
            s.supp = supp;
            s.cbase = reinterpret_cast<const double2*>(cbase);
            s.kmode = kmode;
            s.traj = reinterpret_cast<double2*>(traj);
            s.step_base = &step_base;
            s.traj_bstride = (long long)(nt + 1) * NN;
            s.nmax = nmax;
            s.slot0 = slot0;
            s.N = N;
            s.scramble = scramble;
            s.scheme = 1;
            s.herm = herm;
            s.nmod = M;
            s.nind = K;
            s.lmax = L;
            s.local_step = step;
            s.y = Y;
            s.acc = ACC;
            switch (stage) {
                case 0: s.yin = Y;  s.yout = SA;  s.a = dt / 2; s.first = 1; break;
                case 1: s.yin = SA; s.yout = SB;  s.a = dt / 2; break;
                case 2: s.yin = SB; s.yout = ACC; s.a = dt; break;
                default: s.yin = ACC; s.acc = SA; s.yout = SB; s.ydst = Y; s.a = 2.0 / dt; s.w = dt / 6; s.last = 1; break;
            }
            for (int p = 0; p < nparts; ++p) {
                s.slot_lo = parts[2 * p];
                s.slot_hi = parts[2 * p + 1];
                if (s.slot_hi <= s.slot_lo) continue;
                switch (N) {
#ifndef HEOM_EMU_FEW_N
                    case 2: launch_n<2>(s, H, K, M, L, warps, sm_count, hreal, sym); break;
                    case 3: launch_n<3>(s, H, K, M, L, warps, sm_count, hreal, sym); break;
                    case 5: launch_n<5>(s, H, K, M, L, warps, sm_count, hreal, sym); break;
                    case 6: launch_n<6>(s, H, K, M, L, warps, sm_count, hreal, sym); break;
                    case 8: launch_n<8>(s, H, K, M, L, warps, sm_count, hreal, sym); break;
#endif
                    case 4: launch_n<4>(s, H, K, M, L, warps, sm_count, hreal, sym); break;
                    case 7: launch_n<7>(s, H, K, M, L, warps, sm_count, hreal, sym); break;
                    default: return 1;
                }
            }
        }
    }
    return 0;
}

}  // extern "C"
