// heom_core.cuh - hierarchy indexing shared by host code and device kernels.
//
// The reference indexes ADOs by a combinatorial-number-system hash of the
// multi-index n (gen_hash_value, pyqed/heom/deom.py:555-565):
//     id(n) = sum_i C(s_i + i, i + 1),  s_i = n_0 + ... + n_i
// which is a bijection onto [0, C(L+K, L)), tier-major.  The device may store
// ADOs in a different ("storage slot") order; both orders are ranked/unranked
// in closed form here so no table of keys is needed to build the neighbour
// lists.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define HEOM_HD __host__ __device__ __forceinline__
#else
#define HEOM_HD inline
#endif

namespace heom {

constexpr int MAX_NIND = 64;   // K
constexpr int MAX_SIDE = 96;   // K + L + 1

enum Order { ORDER_REF = 0, ORDER_LEX = 1 };

struct Pascal {
    const long long* tab;  // [side][side], tab[a*side+b] = C(a,b)
    int side;
    HEOM_HD long long C(int a, int b) const {
        return (b < 0 || a < b || a < 0) ? 0ll : tab[a * side + b];
    }
};

// ---- reference (tier-major) order ------------------------------------------
HEOM_HD long long rank_ref(const uint8_t* key, int K, const Pascal& P) {
    int run = 0;
    long long id = 0;
    for (int i = 0; i < K; ++i) {
        run += key[i];
        id += P.C(run + i, i + 1);
    }
    return id;
}

HEOM_HD void unrank_ref(long long id, int K, int L, const Pascal& P, uint8_t* key) {
    // greedy inverse of the combinatorial number system: c_i = s_i + i strictly
    // increasing, c_i = max{c : C(c, i+1) <= remainder}
    long long rem = id;
    int hi = K + L;  // exclusive upper bound for c_{K-1}
    int prev_s = 0;
    uint8_t s[MAX_NIND];
    for (int i = K - 1; i >= 0; --i) {
        int c = hi - 1;
        while (c > i && P.C(c, i + 1) > rem) --c;
        rem -= P.C(c, i + 1);
        s[i] = (uint8_t)(c - i);
        hi = c;
    }
    for (int i = 0; i < K; ++i) {
        key[i] = (uint8_t)(s[i] - prev_s);
        prev_s = s[i];
    }
}

// ---- lexicographic order (dimension 0 most significant, all tiers mixed) ----
HEOM_HD long long rank_lex(const uint8_t* key, int K, int L, const Pascal& P) {
    long long r = 0;
    int b = L;
    for (int i = 0; i < K; ++i) {
        int d = K - 1 - i;
        r += P.C(b + d + 1, d + 1) - P.C(b - key[i] + d + 1, d + 1);
        b -= key[i];
    }
    return r;
}

HEOM_HD void unrank_lex(long long r, int K, int L, const Pascal& P, uint8_t* key) {
    int b = L;
    for (int i = 0; i < K; ++i) {
        int d = K - 1 - i;
        int v = 0;
        while (v < b) {
            long long cnt = P.C(b - v + d, d);
            if (r < cnt) break;
            r -= cnt;
            ++v;
        }
        key[i] = (uint8_t)v;
        b -= v;
    }
}

HEOM_HD long long rank_slot(int order, const uint8_t* key, int K, int L, const Pascal& P) {
    return order == ORDER_LEX ? rank_lex(key, K, L, P) : rank_ref(key, K, P);
}
HEOM_HD void unrank_slot(int order, long long slot, int K, int L, const Pascal& P, uint8_t* key) {
    if (order == ORDER_LEX) unrank_lex(slot, K, L, P, key);
    else unrank_ref(slot, K, L, P, key);
}

// ---- link metadata ----------------------------------------------------------
// One link = (neighbour slot, meta).  meta packs
//   bits 0-7   n_eff (n_k for the n-e_k link, n_k+1 for the n+e_k link)
//   bit  8     direction (0 = minus, 1 = plus)
//   bits 9-15  dissipaton k          -> bits 8-15 together are 2k+dir
//   bits 16-19 first row r with (Q_mode)_rr != 0 (used by the diagonal-Q kernels)
//   bits 24-31 coupling mode
HEOM_HD int link_meta(int dir, int k, int neff, int mode, int r0) {
    return neff | (dir << 8) | (k << 9) | ((r0 & 0xf) << 16) | (mode << 24);
}
HEOM_HD int meta_neff(int meta) { return meta & 0xff; }
HEOM_HD int meta_dir(int meta) { return (meta >> 8) & 1; }
HEOM_HD int meta_k(int meta) { return (meta >> 9) & 0x7f; }
HEOM_HD int meta_kdir(int meta) { return (meta >> 8) & 0xff; }  // 2k + dir
HEOM_HD int meta_r0(int meta) { return (meta >> 16) & 0xf; }
HEOM_HD int meta_mode(int meta) { return (meta >> 24) & 0xff; }
// index into the full coefficient table [dir][k][n_eff]
HEOM_HD int meta_ci(int meta, int K, int L) {
    return (meta_dir(meta) * K + meta_k(meta)) * (L + 1) + meta_neff(meta);
}

}  // namespace heom
