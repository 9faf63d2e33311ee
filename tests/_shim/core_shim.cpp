// Test-only shim: exposes the host/device indexing functions of
// pyqed_b200/csrc/heom_core.cuh to the CPU test-suite (compiled with g++).
#include <vector>
#include "../../pyqed_b200/csrc/heom_core.cuh"

static std::vector<long long> pascal(int side) {
    std::vector<long long> t((size_t)side * side, 0);
    for (int a = 0; a < side; ++a) {
        t[(size_t)a * side] = 1;
        for (int b = 1; b <= a; ++b)
            t[(size_t)a * side + b] = t[(size_t)(a - 1) * side + b - 1] + (b <= a - 1 ? t[(size_t)(a - 1) * side + b] : 0);
    }
    return t;
}

extern "C" {
long long shim_rank(int order, const unsigned char* key, int K, int L) {
    auto t = pascal(K + L + 1);
    heom::Pascal P{t.data(), K + L + 1};
    return heom::rank_slot(order, key, K, L, P);
}
void shim_unrank(int order, long long slot, int K, int L, unsigned char* key) {
    auto t = pascal(K + L + 1);
    heom::Pascal P{t.data(), K + L + 1};
    heom::unrank_slot(order, slot, K, L, P, key);
}
void shim_unrank_all(int order, long long nmax, int K, int L, unsigned char* keys) {
    auto t = pascal(K + L + 1);
    heom::Pascal P{t.data(), K + L + 1};
    for (long long s = 0; s < nmax; ++s) heom::unrank_slot(order, s, K, L, P, keys + s * K);
}
int shim_link_meta(int dir, int k, int neff, int mode, int K, int L) {
    (void)K; (void)L; return heom::link_meta(dir, k, neff, mode, 0);
}
}
