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
        // two ranks with their own state buffers and kernel 6's fused push (no exchange by the
        // host): every rank must end up with kernel 6's single-rank result on the slots it owns
        {
            const long long lo[2] = {0, nmax / 2 + 1}, hi[2] = {nmax / 2 + 1, nmax};
            std::vector<std::vector<double>> st(2, std::vector<double>(4 * asz, std::nan("")));
            for (int r = 0; r < 2; ++r) {
                for (size_t i = 0; i < asz; ++i) st[r][i] = 0.0;           // Y: every rank starts from the full state
                for (size_t i = 0; i < 2 * NN; ++i) st[r][i] = rho0[i];
            }
            std::vector<std::vector<int>> pptr(2);
            std::vector<std::vector<unsigned char>> pent(2);
            for (int r = 0; r < 2; ++r) {   // rows of rank r's slots that the other rank's links read
                std::vector<std::vector<unsigned char>> per((size_t)(hi[r] - lo[r]));
                const int q = 1 - r;
                for (long long n = lo[q]; n < hi[q]; ++n)
                    for (int l = link_ptr[n]; l < link_ptr[n + 1]; ++l) {
                        const long long nb = links[2 * l];
                        const unsigned char ent = (unsigned char)((q << 4) | ((links[2 * l + 1] >> 16) & 0xf));
                        if (nb >= lo[r] && nb < hi[r]) {
                            auto& v = per[(size_t)(nb - lo[r])];
                            if (std::find(v.begin(), v.end(), ent) == v.end()) v.push_back(ent);
                        }
                    }
                pptr[r].push_back(0);
                for (auto& v : per) {
                    pent[r].insert(pent[r].end(), v.begin(), v.end());
                    pptr[r].push_back((int)pent[r].size());
                }
                if (pent[r].empty()) pent[r].push_back(0);
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
                        if (emu_sym_stage(N, K, M, L, H.data(), ops.data(), cbase.data(), kmode.data(), damp.data(),
                                          link_ptr.data(), links.data(), nlinks, b + p.yin * asz, b, b + asz,
                                          b + 2 * asz, b + p.out * asz, p.a, p.w, p.kind, hreal, 2, 2, lo[r], hi[r],
                                          nmax, pptr[r].data(), pent[r].data(), peers, (long long)p.out * nmax * NN,
                                          &err)) {
                            std::cerr << "fused push failed: " << err << "\n";
                            return 1;
                        }
                    }
            double dpush = 0.0;
            for (int r = 0; r < 2; ++r)
                for (size_t i = 2 * NN * (size_t)lo[r]; i < 2 * NN * (size_t)hi[r]; ++i) {
                    const double d = std::fabs(st[r][i] - s6[i]);
                    if (!(d <= dpush)) dpush = d;
                }
            std::cout << "late=" << late << " |fused-push ranks - k6|=" << dpush << "\n";
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
