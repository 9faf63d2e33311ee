// emu_sanitize_main.cpp - stand-alone driver that runs the emulated kernels 3, 6 and 7 on a
// problem dumped by tests/test_emu_sanitizers.py.  Built with -fsanitize=address (out-of-bounds
// accesses to global or shared memory) and with -fsanitize=thread (data races between lanes
// that a missing __syncwarp / __syncthreads / wait would cause).  TEST INFRASTRUCTURE ONLY.
#include "cuda_emu.h"

#define EMU_ONE_SETTER 1
#include "sym_emu.cpp"
#include "async_emu.cpp"

#include <fstream>
#include <iostream>
#include <string>

template <typename T>
static std::vector<T> load(const std::string& dir, const char* name) {
    std::ifstream f(dir + "/" + name + ".bin", std::ios::binary | std::ios::ate);
    if (!f) {
        std::cerr << "cannot open " << name << "\n";
        std::exit(2);
    }
    const std::streamsize bytes = f.tellg();
    f.seekg(0);
    std::vector<T> v((size_t)bytes / sizeof(T));   // exactly sized: the sanitizer guards both ends
    f.read(reinterpret_cast<char*>(v.data()), bytes);
    return v;
}

static double max_abs_diff(const std::vector<double>& a, const std::vector<double>& b) {
    double m = 0.0;
    for (size_t i = 0; i < a.size(); ++i) {
        const double d = std::fabs(a[i] - b[i]);
        if (!(d <= m)) m = d;   // NaN propagates
    }
    return m;
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    const std::string dir = argv[1];
    const auto meta = load<long long>(dir, "meta");   // N K M L nmax nlinks nt hreal
    const int N = (int)meta[0], K = (int)meta[1], M = (int)meta[2], L = (int)meta[3];
    const long long nmax = meta[4], nlinks = meta[5];
    const int nt = (int)meta[6], hreal = (int)meta[7];
    const auto dtv = load<double>(dir, "dt");
    const auto H = load<double>(dir, "H"), ops = load<double>(dir, "ops"), cbase = load<double>(dir, "cbase");
    const auto damp = load<double>(dir, "damp"), rho0 = load<double>(dir, "rho0");
    const auto kmode = load<int>(dir, "kmode"), link_ptr = load<int>(dir, "link_ptr"), links = load<int>(dir, "links");
    const auto supp = load<unsigned char>(dir, "supp");
    const size_t NN = (size_t)N * N, asz = 2 * (size_t)nmax * NN;
    const long long parts[4] = {0, nmax / 2 + 1, nmax / 2 + 1, nmax};
    int bad = 0;
    for (int late = 0; late < 2; ++late) {
        emu::g_async_late = late;
        std::vector<double> s6(4 * asz, 0.0), s3(4 * asz, 0.0), s3g(4 * asz, 0.0), y7(asz, 0.0), y7p(asz, 0.0);
        for (auto* v : {&s6, &s3, &s3g})
            for (size_t i = asz; i < 4 * asz; ++i) (*v)[i] = std::nan("");
        for (size_t i = 0; i < 2 * NN; ++i) s6[i] = s3[i] = s3g[i] = y7[i] = y7p[i] = rho0[i];
        std::vector<double> t6(2 * NN * (nt + 1)), t7(t6.size()), t3(t6.size());
        const char* err = "";
        if (emu_sym_run(N, K, M, L, nmax, H.data(), ops.data(), cbase.data(), kmode.data(), damp.data(),
                        link_ptr.data(), links.data(), nlinks, s6.data(), dtv[0], nt, hreal, 3, 2, parts, 2, 0, 1,
                        t6.data(), &err) ||
            emu_packed_run(N, K, M, L, nmax, H.data(), ops.data(), cbase.data(), kmode.data(), damp.data(),
                           link_ptr.data(), links.data(), nlinks, y7.data(), dtv[0], nt, hreal, 3, 2, 0, 1,
                           t7.data(), 0, &err) ||
            emu_packed_run(N, K, M, L, nmax, H.data(), ops.data(), cbase.data(), kmode.data(), damp.data(),
                           link_ptr.data(), links.data(), nlinks, y7p.data(), dtv[0], nt, hreal, 3, 2, 0, 1,
                           nullptr, 1, &err)) {
            std::cerr << "kernel 6/7 failed: " << err << "\n";
            return 1;
        }
        emu_async_run(N, K, M, L, nmax, H.data(), ops.data(), cbase.data(), kmode.data(), supp.data(), damp.data(),
                      link_ptr.data(), links.data(), s3.data(), dtv[0], nt, hreal, 1, 1, 3, 2, parts, 2, 0, 1,
                      t3.data());
        emu_async_run(N, K, M, L, nmax, H.data(), ops.data(), cbase.data(), kmode.data(), supp.data(), damp.data(),
                      link_ptr.data(), links.data(), s3g.data(), dtv[0], nt, hreal, 0, 0, 3, 2, parts, 2, 0, 0,
                      nullptr);
        // two ranks with rank-local arrays (own ADOs + pool of halo rows, heom_shard.cu in miniature)
        // and the fused push, on full matrices (kernel 6) and on upper triangles (kernel 7): every
        // rank must end up with kernel 6's single-rank result on the slots it owns.  The arrays are
        // sized exactly, so any access outside them is an AddressSanitizer error.
        for (int packed = 0; packed < 2; ++packed) {
            const long long lo[2] = {0, nmax / 2 + 1}, hi[2] = {nmax / 2 + 1, nmax};
            const int PK = N * (N + 1) / 2, EL = packed ? PK : (int)NN;
            std::vector<int> links2(2 * (size_t)nlinks);
            if (emu_convert_links(links.data(), links2.data(), nlinks, N, L, packed, &err)) return 1;
            std::vector<std::vector<long long>> need(2);
            for (int r = 0; r < 2; ++r) {
                for (int l = link_ptr[lo[r]]; l < link_ptr[hi[r]]; ++l) {
                    const long long nb = links[2 * l];
                    if (nb < lo[r] || nb >= hi[r]) need[r].push_back(nb * 8 + ((links[2 * l + 1] >> 16) & 15));
                }
                std::sort(need[r].begin(), need[r].end());
                need[r].erase(std::unique(need[r].begin(), need[r].end()), need[r].end());
            }
            const int PS = (N + 1) & ~1;                                 // sym_pool_stride
            const long long n_own_max = std::max(hi[0] - lo[0], hi[1] - lo[1]);
            const long long pool_max = (long long)std::max(need[0].size(), need[1].size());
            const long long pool_off = (n_own_max * EL + 7) & ~7ll, arr = pool_off + pool_max * PS;
            // rank-local link tables (shard_localize_links_kernel restated)
            std::vector<std::vector<int>> loc(2, links2);
            for (int r = 0; r < 2; ++r)
                for (int l = link_ptr[lo[r]]; l < link_ptr[hi[r]]; ++l) {
                    const long long nb = links[2 * l];
                    const int r0 = (links[2 * l + 1] >> 16) & 15;
                    unsigned x, tr;
                    if (nb >= lo[r] && nb < hi[r]) {
                        x = packed ? (unsigned)(nb - lo[r]) * PK : ((unsigned)(nb - lo[r]) * N + r0) * N;
                        tr = r0;
                    } else {
                        const long long item = nb * 8 + r0;
                        x = (unsigned)pool_off +
                            (unsigned)(std::lower_bound(need[r].begin(), need[r].end(), item) - need[r].begin()) * PS;
                        tr = N;
                    }
                    loc[r][2 * l] = (int)x;
                    loc[r][2 * l + 1] = (int)(((unsigned)links2[2 * l + 1] & 0x0fffffffu) | (tr << 28));
                }
            std::vector<std::vector<double>> st(2, std::vector<double>(2 * 4 * (size_t)arr, std::nan("")));
            auto elem = [&](long long slot_local, int i, int j) {   // offset (double2) of element (i, j), i <= j if packed
                return packed ? slot_local * EL + i * N - i * (i - 1) / 2 + (j - i) : slot_local * EL + i * N + j;
            };
            for (int r = 0; r < 2; ++r) {   // Y: zeros, ADO 0 = rho0 on its owner
                for (long long e = 0; e < (hi[r] - lo[r]) * EL; ++e) st[r][2 * e] = st[r][2 * e + 1] = 0.0;
                if (lo[r] == 0)
                    for (int i = 0; i < N; ++i)
                        for (int j = packed ? i : 0; j < N; ++j) {
                            st[r][2 * elem(0, i, j)] = rho0[2 * (i * N + j)];
                            st[r][2 * elem(0, i, j) + 1] = rho0[2 * (i * N + j) + 1];
                        }
            }
            auto row_value = [&](int r, long long slot, int row, int j, int arrid, int part) {   // element (row, j) of an owned ADO
                const int i0 = packed ? std::min(row, j) : row, j0 = packed ? std::max(row, j) : j;
                const double v = st[r][2 * ((size_t)arrid * arr + elem(slot - lo[r], i0, j0)) + part];
                return (packed && part == 1 && j < row) ? -v : v;
            };
            for (int r = 0; r < 2; ++r)   // halo rows of the stage-0 input (what shard_begin pushes)
                for (size_t i = 0; i < need[r].size(); ++i)
                    for (int j = 0; j < N; ++j)
                        for (int part = 0; part < 2; ++part)
                            st[r][2 * (pool_off + (long long)i * PS + j) + part] =   // packed: as read through the triangle
                                (packed && part == 1 && j < (int)(need[r][i] & 7) ? -1.0 : 1.0) *
                                row_value(1 - r, need[r][i] >> 3, (int)(need[r][i] & 7), j, 0, part);
            std::vector<std::vector<int>> pptr(2), pent(2);
            for (int r = 0; r < 2; ++r) {   // rows of rank r's slots that the other rank's links read
                std::vector<std::vector<std::pair<int, int>>> per((size_t)(hi[r] - lo[r]));   // (row, pool index)
                const int q = 1 - r, apw = 32 / N;
                for (size_t i = 0; i < need[q].size(); ++i)
                    per[(size_t)((need[q][i] >> 3) - lo[r])].push_back({(int)(need[q][i] & 7), (int)i});
                pptr[r].push_back(0);
                for (size_t g0 = 0; g0 < per.size(); g0 += apw) {   // staging slots: distinct (ADO, row) per group
                    int next_id = 0;
                    for (size_t sl = g0; sl < std::min(per.size(), g0 + apw); ++sl) {
                        std::sort(per[sl].begin(), per[sl].end());
                        int last_row = -1, sid = 255;
                        for (auto& e : per[sl]) {
                            if (e.first != last_row) {
                                sid = next_id < 12 ? next_id : 255;
                                ++next_id;
                                last_row = e.first;
                            }
                            pent[r].push_back(e.second);
                            pent[r].push_back((sid << 8) | (q << 4) | e.first);
                        }
                        pptr[r].push_back((int)pent[r].size() / 2);
                    }
                }
                if (pent[r].empty()) pent[r].assign(2, 0);
            }
            const unsigned long long peers[2] = {(unsigned long long)(uintptr_t)st[0].data(),
                                                 (unsigned long long)(uintptr_t)st[1].data()};
            const double dt = dtv[0];
            const struct { int yin, out, kind; double a, w; } plan[4] = {
                {0, 1, 0, dt / 2, 0.0}, {1, 2, 1, dt / 2, 0.0}, {2, 3, 1, dt, 0.0}, {3, 0, 2, 2.0 / dt, dt / 6}};
            for (int step = 0; step < nt; ++step)
                for (const auto& p : plan)
                    for (int r = 0; r < 2; ++r) {
                        double* b = st[r].data();
                        if (emu_sym_stage(N, K, M, L, H.data(), ops.data(), cbase.data(), kmode.data(),
                                          damp.data() + 2 * lo[r], link_ptr.data() + lo[r], loc[r].data(),
                                          b + 2 * p.yin * arr, b, b + 2 * arr, b + 4 * arr, b + 2 * p.out * arr, p.a, p.w,
                                          p.kind, hreal, 2, 2, hi[r] - lo[r], packed, pool_off, pptr[r].data(),
                                          pent[r].data(), peers, (long long)p.out * arr + pool_off, &err)) {
                            std::cerr << "fused push failed: " << err << "\n";
                            return 1;
                        }
                    }
            double dpush = 0.0;
            for (int r = 0; r < 2; ++r)
                for (long long slot = lo[r]; slot < hi[r]; ++slot)
                    for (int i = 0; i < N; ++i)
                        for (int j = 0; j < N; ++j)
                            for (int part = 0; part < 2; ++part) {
                                const double d = std::fabs(row_value(r, slot, i, j, 0, part) - s6[2 * ((size_t)slot * NN + i * N + j) + part]);
                                if (!(d <= dpush)) dpush = d;
                            }
            std::cout << "late=" << late << " packed=" << packed << " |fused-push ranks - k6|=" << dpush << "\n";
            if (!(dpush == 0.0)) bad = 1;
        }
        s6.resize(asz);
        s3.resize(asz);
        s3g.resize(asz);
        const double d67 = std::max(max_abs_diff(s6, y7), max_abs_diff(y7, y7p));   // incl. the prefetching variant
        const double d36 = max_abs_diff(s3, s6), d33 = max_abs_diff(s3, s3g);
        std::cout << "late=" << late << " |k6-k7|=" << d67 << " |k3-k6|=" << d36 << " |k3sym-k3general|=" << d33
                  << " |traj6-traj7|=" << max_abs_diff(t6, t7) << "\n";
        if (!(d67 == 0.0) || !(d36 < 1e-12) || !(d33 < 1e-12) || !(max_abs_diff(t6, t3) < 1e-12)) bad = 1;
    }
    std::cout << (bad ? "MISMATCH" : "emulated kernels agree") << "\n";
    return bad;
}
