// heom_hierarchy.cuh - hierarchy builder kernels (keys, links, block permutation) and the small per-step kernels (H(t)/Q(t), rho_sys record, expectation values).
// Included by heom_kernels.cu only (one translation unit); split out for readability.
#pragma once
#include "heom_core.cuh"
#include "heom_device.cuh"

using heom::Pascal;

// ---------------------------------------------------------------------------
// hierarchy builder kernels
// ---------------------------------------------------------------------------
struct HierArgs {
    const long long* pascal;
    int side, K, L, order;
    long long nmax;
    uint8_t* keys;
    int* id_of_slot;
    int* slot_of_id;
    double2* damp;
    int* link_ptr;
    int2* links;
    const double2* expn;  // [K] device copy (stored at the head of coef scratch)
    const int* mode;      // [K]: mode | first support row << 8
    int* lex2slot;        // order 2 only: lexicographic rank <-> storage slot
    int* slot2lex;
};

// Storage order 2 = lexicographic order with a stable partition inside every
// aligned block of ORDER2_BLOCK ranks: ADOs below the top tier (which carry the
// K extra n+e_k links) first, top-tier ADOs after them.  Consecutive slots then
// have similar link counts (warps stay balanced) while the locality and the
// small rank-boundary halos of the lexicographic order are kept.
constexpr int ORDER2_BLOCK = 64;
__global__ void hier_blockperm_kernel(HierArgs h) {
    const long long blk = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long r0 = blk * ORDER2_BLOCK;
    if (r0 >= h.nmax) return;
    Pascal P{h.pascal, h.side};
    const int cnt = (int)min((long long)ORDER2_BLOCK, h.nmax - r0);
    unsigned long long top = 0ull;
    uint8_t key[heom::MAX_NIND];
    for (int i = 0; i < cnt; ++i) {
        heom::unrank_lex(r0 + i, h.K, h.L, P, key);
        int tier = 0;
        for (int k = 0; k < h.K; ++k) tier += key[k];
        if (tier == h.L) top |= 1ull << i;
    }
    const int nlow = cnt - __popcll(top);
    int a = 0, b = nlow;
    for (int i = 0; i < cnt; ++i) {
        const int pos = ((top >> i) & 1ull) ? b++ : a++;
        h.lex2slot[r0 + i] = (int)(r0 + pos);
        h.slot2lex[r0 + pos] = (int)(r0 + i);
    }
}
__device__ __forceinline__ void unrank_any(const HierArgs& h, long long slot, const Pascal& P, uint8_t* key) {
    if (h.order == 2) heom::unrank_lex(h.slot2lex[slot], h.K, h.L, P, key);
    else heom::unrank_slot(h.order, slot, h.K, h.L, P, key);
}
__device__ __forceinline__ long long rank_any(const HierArgs& h, const uint8_t* key, const Pascal& P) {
    if (h.order == 2) return h.lex2slot[heom::rank_lex(key, h.K, h.L, P)];
    return heom::rank_slot(h.order, key, h.K, h.L, P);
}

// pass 1: one thread per storage slot - multi-index, damping rate, link count
__global__ void hier_keys_kernel(HierArgs h) {
    const long long slot = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (slot >= h.nmax) return;
    Pascal P{h.pascal, h.side};
    uint8_t key[heom::MAX_NIND];
    unrank_any(h, slot, P, key);
    int tier = 0, nz = 0;
    double dr = 0.0, di = 0.0;
    for (int k = 0; k < h.K; ++k) {
        h.keys[slot * h.K + k] = key[k];
        tier += key[k];
        nz += key[k] > 0;
        const double2 g = h.expn[k];
        dr += key[k] * g.x;
        di += key[k] * g.y;
    }
    h.damp[slot] = make_double2(dr, di);
    const long long id = heom::rank_ref(key, h.K, P);
    h.id_of_slot[slot] = (int)id;
    h.slot_of_id[id] = (int)slot;
    h.link_ptr[slot] = nz + (tier < h.L ? h.K : 0);
    if (slot == 0) h.link_ptr[h.nmax] = 0;
}

// pass 2: fill links in the reference's summation order (k ascending; for each
// k the n-e_k term, then the n+e_k term; deom.py:651-664)
__global__ void hier_links_kernel(HierArgs h) {
    const long long slot = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (slot >= h.nmax) return;
    Pascal P{h.pascal, h.side};
    uint8_t key[heom::MAX_NIND];
    int tier = 0;
    for (int k = 0; k < h.K; ++k) {
        key[k] = h.keys[slot * h.K + k];
        tier += key[k];
    }
    int w = h.link_ptr[slot];
    for (int k = 0; k < h.K; ++k) {
        const int nk = key[k];
        if (nk > 0) {
            key[k] = (uint8_t)(nk - 1);
            const long long nb = rank_any(h, key, P);
            key[k] = (uint8_t)nk;
            h.links[w++] = make_int2((int)nb, heom::link_meta(0, k, nk, h.mode[k] & 0xff, h.mode[k] >> 8));
        }
        if (tier < h.L) {
            key[k] = (uint8_t)(nk + 1);
            const long long nb = rank_any(h, key, P);
            key[k] = (uint8_t)nk;
            h.links[w++] = make_int2((int)nb, heom::link_meta(1, k, nk + 1, h.mode[k] & 0xff, h.mode[k] >> 8));
        }
    }
}

// gather / scatter between storage-slot order and reference id order
__global__ void permute_kernel(double2* dst, const double2* src, const int* map, long long nmax,
                               int NN, int dst_is_mapped) {
    // dst_is_mapped: dst[map[i]] = src[i]   else   dst[i] = src[map[i]]
    const long long total = nmax * NN;
    const long long boff = blockIdx.y * total;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long i = e / NN;
        const int r = (int)(e - i * NN);
        const long long j = map[i];
        if (dst_is_mapped) dst[boff + j * NN + r] = src[boff + e];
        else dst[boff + e] = src[boff + j * NN + r];
    }
}

// ---------------------------------------------------------------------------
// time-dependent operators: ops_t[b][o] = base[o] + dip[o] * field_o[b][step][tidx]
// (generate_time, deom.py:676-687).  Single block; the tables are tiny.
// ---------------------------------------------------------------------------
__global__ void prep_ops_kernel(double2* ops_t, const double2* base, const double2* dip,
                                const double* fsys, const double* fcoup, const long long* step_base,
                                int local_step, int tidx, long long nt, int B, int M1, int NN) {
    const long long step = *step_base + local_step;
    const int per = M1 * NN;
    for (int e = threadIdx.x; e < B * per; e += blockDim.x) {
        const int b = e / per, r = e - b * per, o = r / NN;
        const double* f = (o == 0) ? fsys : fcoup;
        const double s = f ? f[((long long)b * nt + step) * 3 + tidx] : 0.0;
        const double2 v = base[r], d = dip[r];
        ops_t[e] = make_double2(fma(d.x, s, v.x), fma(d.y, s, v.y));
    }
}

__global__ void advance_kernel(long long* step_base, long long by) { *step_base += by; }

// rho_sys of every trajectory -> traj[b][index]
__global__ void record_kernel(double2* traj, const double2* y, long long nmax, long long slot0,
                              int NN, long long traj_bstride, long long index) {
    const int b = blockIdx.x;
    for (int e = threadIdx.x; e < NN; e += blockDim.x)
        traj[b * traj_bstride + index * NN + e] = y[(b * nmax + slot0) * NN + e];
}

// out[b][o][p] = Tr(op_o rho[b][p])
__global__ void expectation_kernel(double2* out, const double2* rho, const double2* ops,
                                   long long npts, int n_ops, int N) {
    const long long total = (long long)gridDim.y * n_ops * npts;
    const int NN = N * N;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n_ops * npts;
         t += (long long)gridDim.x * blockDim.x) {
        (void)total;
        const int b = blockIdx.y;
        const int o = (int)(t / npts);
        const long long pt = t - (long long)o * npts;
        const double2* r = rho + ((long long)b * npts + pt) * NN;
        const double2* a = ops + (long long)o * NN;
        double2 s = make_double2(0.0, 0.0);
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) cfma(s, a[i * N + j], r[j * N + i]);
        out[((long long)b * n_ops + o) * npts + pt] = s;
    }
}

