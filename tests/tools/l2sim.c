// l2sim.c - sector-granular L2 model of one RK stage of the row-gather stage kernels.
//
// Development tooling (not part of the product path): replays the global-memory access
// stream of a middle RK4 stage - own tile, y tile, neighbour rows (16N bytes each), output
// write - in a given storage order through a set-associative LRU cache with 128-byte lines
// and 32-byte sectors, and reports the DRAM read traffic per class.  Used to choose storage /
// visiting orders and cache policies before spending GPU time; calibrated against the ncu
// DRAM byte counts under profiles/.
//
//   gcc -O2 -o /tmp/l2sim tests/tools/l2sim.c -lm && /tmp/l2sim K L order cap_MB [packed] [evict_first_streams]
//   order: 0 reference (tier-major hash, deom.py:555-565), 1 lexicographic, 2 blocked lex (64),
//          3 lexicographic on the reversed index
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAXK 64
static long long Ctab[128][128];
static int K, L;

static long long C(int a, int b) { return (b < 0 || a < b || a < 0) ? 0 : Ctab[a][b]; }

static long long rank_ref(const uint8_t* key) {
    int run = 0;
    long long id = 0;
    for (int i = 0; i < K; ++i) {
        run += key[i];
        id += C(run + i, i + 1);
    }
    return id;
}
static long long rank_lex(const uint8_t* key) {
    long long r = 0;
    int b = L;
    for (int i = 0; i < K; ++i) {
        int d = K - 1 - i;
        r += C(b + d + 1, d + 1) - C(b - key[i] + d + 1, d + 1);
        b -= key[i];
    }
    return r;
}
static void unrank_lex(long long r, uint8_t* key) {
    int b = L;
    for (int i = 0; i < K; ++i) {
        int d = K - 1 - i, v = 0;
        while (v < b) {
            long long cnt = C(b - v + d, d);
            if (r < cnt) break;
            r -= cnt;
            ++v;
        }
        key[i] = (uint8_t)v;
        b -= v;
    }
}

// ---- cache ------------------------------------------------------------------
typedef struct {
    uint64_t tag;      // line address + 1 (0 = empty)
    uint32_t stamp;    // LRU time
    uint8_t valid;     // sector mask
} Way;
static Way* cache;
static long long nsets;
static int ways = 16;
static uint32_t now = 1;
static double dram[8];   // bytes read per class
static double accs[8];   // sectors accessed per class

// access [addr, addr+bytes) ; cls = traffic class; is_write: sectors become valid without a fill;
// low_prio: insert / touch at LRU position (evict-first)
static void touch(uint64_t addr, int bytes, int cls, int is_write, int low_prio) {
    uint64_t s0 = addr >> 5, s1 = (addr + bytes - 1) >> 5;
    for (uint64_t s = s0; s <= s1; ++s) {
        uint64_t line = s >> 2;
        int sec = (int)(s & 3);
        uint64_t h = line * 0x9E3779B97F4A7C15ull;
        long long set = (long long)((h >> 20) % (uint64_t)nsets);
        Way* w = cache + set * ways;
        int hit = -1, victim = 0;
        uint32_t oldest = 0xffffffffu;
        for (int i = 0; i < ways; ++i) {
            if (w[i].tag == line + 1) { hit = i; break; }
            if (w[i].tag == 0) { victim = i; oldest = 0; }
            else if (w[i].stamp < oldest) { oldest = w[i].stamp; victim = i; }
        }
        accs[cls] += 1;
        ++now;
        if (hit >= 0) {
            if (!(w[hit].valid & (1 << sec))) {
                if (!is_write) dram[cls] += 32;
                w[hit].valid |= (uint8_t)(1 << sec);
            }
            if (!low_prio) w[hit].stamp = now;
        } else {
            if (!is_write) dram[cls] += 32;
            w[victim].tag = line + 1;
            w[victim].valid = (uint8_t)(1 << sec);
            w[victim].stamp = low_prio ? (now > 40000000u ? now - 40000000u : 1) : now;
        }
    }
}

int main(int argc, char** argv) {
    if (argc < 5) {
        fprintf(stderr, "usage: l2sim K L order cap_MB [packed=0] [evict_first=0] [N=7] [nmode_per=3]\n");
        return 1;
    }
    K = atoi(argv[1]);
    L = atoi(argv[2]);
    const int order = atoi(argv[3]);
    const double cap_mb = atof(argv[4]);
    const int packed = argc > 5 ? atoi(argv[5]) : 0;
    const int evict_first = argc > 6 ? atoi(argv[6]) : 0;
    const int N = argc > 7 ? atoi(argv[7]) : 7;
    const int per_mode = argc > 8 ? atoi(argv[8]) : 3;
    for (int a = 0; a < 128; ++a) {
        Ctab[a][0] = 1;
        for (int b = 1; b <= a; ++b) Ctab[a][b] = Ctab[a - 1][b - 1] + (b <= a - 1 ? Ctab[a - 1][b] : 0);
    }
    const long long nmax = C(L + K, L);
    nsets = (long long)(cap_mb * 1e6 / 128 / ways);
    cache = (Way*)calloc((size_t)nsets * ways, sizeof(Way));
    // storage order: slot -> key ; key -> slot
    uint8_t* keys = (uint8_t*)malloc((size_t)nmax * K);
    int* slot_of_lex = NULL;
    for (long long r = 0; r < nmax; ++r) unrank_lex(r, keys + (size_t)r * K);   // lex rank -> key
    int* slot_of = (int*)malloc(sizeof(int) * nmax);    // lex rank -> storage slot
    int* lex_of = (int*)malloc(sizeof(int) * nmax);     // storage slot -> lex rank
    if (order == 1) {
        for (long long r = 0; r < nmax; ++r) slot_of[r] = lex_of[r] = (int)r;
    } else if (order == 0) {
        for (long long r = 0; r < nmax; ++r) {
            long long id = rank_ref(keys + (size_t)r * K);
            slot_of[r] = (int)id;
            lex_of[id] = (int)r;
        }
    } else if (order == 2) {
        for (long long r0 = 0; r0 < nmax; r0 += 64) {
            int cnt = (int)((nmax - r0) < 64 ? (nmax - r0) : 64), a = 0, b = 0;
            for (int i = 0; i < cnt; ++i) {
                int t = 0;
                for (int k = 0; k < K; ++k) t += keys[(size_t)(r0 + i) * K + k];
                if (t < L) ++b;
            }
            for (int i = 0; i < cnt; ++i) {
                int t = 0;
                for (int k = 0; k < K; ++k) t += keys[(size_t)(r0 + i) * K + k];
                int pos = (t == L) ? b++ : a++;
                slot_of[r0 + i] = (int)(r0 + pos);
                lex_of[r0 + pos] = (int)(r0 + i);
            }
        }
    } else {   // 3: lexicographic on the reversed multi-index (dimension K-1 most significant)
        uint8_t rev[MAXK];
        for (long long r = 0; r < nmax; ++r) {
            for (int k = 0; k < K; ++k) rev[k] = keys[(size_t)r * K + (K - 1 - k)];
            long long s = rank_lex(rev);
            slot_of[r] = (int)s;
            lex_of[s] = (int)r;
        }
    }
    (void)slot_of_lex;
    const int EL = packed ? N * (N + 1) / 2 : N * N;
    const uint64_t ado_bytes = (uint64_t)EL * 16, arr = (uint64_t)nmax * ado_bytes + (1 << 20);
    const uint64_t YIN = 0, Y = arr, OUT = 2 * arr;
    enum { OWN = 0, YT = 1, NBR = 2, WR = 3, TAB = 4 };
    const uint64_t TABB = 3 * arr;
    long long nlinks = 0;
    uint8_t key[MAXK];
    for (long long slot = 0; slot < nmax; ++slot) {
        const long long r = lex_of[slot];
        memcpy(key, keys + (size_t)r * K, K);
        int tier = 0;
        for (int k = 0; k < K; ++k) tier += key[k];
        touch(YIN + slot * ado_bytes, (int)ado_bytes, OWN, 0, 0);
        touch(Y + slot * ado_bytes, (int)ado_bytes, YT, 0, evict_first);
        for (int k = 0; k < K; ++k) {
            const int m = k / per_mode, r0 = m % N;
            for (int dir = 0; dir < 2; ++dir) {
                if (dir == 0 && key[k] == 0) continue;
                if (dir == 1 && tier >= L) continue;
                key[k] += dir ? 1 : -1;
                const long long nb = slot_of[rank_lex(key)];
                key[k] -= dir ? 1 : -1;
                touch(TABB + (uint64_t)nlinks * 8, 8, TAB, 0, evict_first);
                ++nlinks;
                if (!packed) {
                    touch(YIN + nb * ado_bytes + (uint64_t)r0 * N * 16, N * 16, NBR, 0, 0);
                } else {
                    // row r0 of the upper triangle: (i, r0) for i < r0, then (r0, j) for j >= r0
                    for (int i = 0; i < r0; ++i)
                        touch(YIN + nb * ado_bytes + (uint64_t)(i * N - i * (i - 1) / 2 + (r0 - i)) * 16, 16, NBR, 0, 0);
                    touch(YIN + nb * ado_bytes + (uint64_t)(r0 * N - r0 * (r0 - 1) / 2) * 16, (N - r0) * 16, NBR, 0, 0);
                }
            }
        }
        touch(OUT + slot * ado_bytes, (int)ado_bytes, WR, 1, evict_first);
    }
    const double gb = 1e-9;
    printf("K=%d L=%d N=%d order=%d cap=%.0fMB packed=%d evict_first=%d nmax=%lld links=%lld\n", K, L, N, order,
           cap_mb, packed, evict_first, nmax, nlinks);
    printf("  own  read %.3f GB (ideal %.3f)\n", dram[OWN] * gb, nmax * (double)ado_bytes * gb);
    printf("  y    read %.3f GB\n", dram[YT] * gb);
    printf("  nbr  read %.3f GB (sectors accessed %.3f GB, miss %.3f)\n", dram[NBR] * gb, accs[NBR] * 32 * gb,
           dram[NBR] / (accs[NBR] * 32));
    printf("  tab  read %.3f GB\n", dram[TAB] * gb);
    printf("  total read %.3f GB, write %.3f GB\n", (dram[OWN] + dram[YT] + dram[NBR] + dram[TAB]) * gb,
           nmax * (double)ado_bytes * gb);
    return 0;
}
