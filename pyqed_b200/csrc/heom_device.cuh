// heom_device.cuh - device helpers shared by the translation units of the
// HEOM kernels: complex arithmetic on double2, asynchronous copies (cp.async,
// cp.async.bulk + mbarrier) and the argument block of a stage launch.
//
// Every hardware-specific operation goes through one of the small wrappers
// below.  With HEOM_HOST_EMU defined (tests/_shim/cuda_emu.h, CPU tests only)
// the wrappers become plain C++ so that a kernel's indexing and arithmetic can
// be exercised without a GPU; the product build never defines it.
#pragma once
#ifndef HEOM_HOST_EMU
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#ifndef HEOM_HOST_EMU
// dynamic shared memory of a kernel / a kernel launch (the emulation header has its own)
#define HEOM_DYN_SMEM(T, name) extern __shared__ T name[]
#define HEOM_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)
#endif

// ---------------------------------------------------------------------------
// complex helpers (double2 = re, im)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cfma(double2& acc, const double2 a, const double2 b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}
__device__ __forceinline__ void cfms(double2& acc, const double2 a, const double2 b) {  // acc -= a*b
    acc.x = fma(-a.x, b.x, acc.x);
    acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(-a.x, b.y, acc.y);
    acc.y = fma(-a.y, b.x, acc.y);
}
__device__ __forceinline__ double2 cmul(const double2 a, const double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 ldg2(const double2* p) { return __ldg(p); }

// streaming (evict-first) accesses for the arrays that are touched once per launch
__device__ __forceinline__ double2 ld_stream(const double2* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(double2* p, const double2 v) { __stcs(p, v); }

// ---------------------------------------------------------------------------
// stage launch arguments (kernels 1-3 and the generic kernel)
// ---------------------------------------------------------------------------
struct StageArgs {
    const double2* yin;   // stage input (read-only in this launch)
    const double2* y;     // state at the start of the step
    double2* acc;         // running combination
    double2* yout;        // next stage input (unused when last)
    double2* ydst;        // end-of-step state (used when last)
    const double2* damp;
    const int* link_ptr;
    const int2* links;
    const double2* coef;  // [ci] -> (alphaL, alphaR)
    const double2* ops;   // [b][1+M][N*N] operators at this stage time
    long long ops_bstride;
    const short* row_ptr;
    const short* row_idx;
    const short* col_ptr;
    const short* col_idx;
    const unsigned char* supp;  // diagonal-Q tables: [M][N+1] (count, rows) then [M][N] membership
    double2* traj;        // may be null
    const long long* step_base;
    long long traj_bstride;
    long long nmax, slot0, ngroups;
    long long slot_lo, slot_hi;  // owned slot range of this rank (whole hierarchy on one GPU)
    double a, w;
    int local_step, first, last, N;
    int scramble;  // rotate the visiting order inside runs of 16 groups (storage order 2)
    int scheme;  // 0: running accumulator (16 passes/step); 1: difference form (13 passes/step, async kernel)
    int herm, ncoef, nmod, nind, lmax;
    const double2* cbase;  // [K][4]: minus (L,R) and plus (L,R) coefficients for n_eff = 1
    const int* kmode;      // [K]: mode | first support row << 8
    // fused multi-GPU halo (async kernel): rows of the stage output that other
    // ranks need are stored into their arrays by the epilogue
    const int* push_ptr;            // [owned+1] CSR over the owned slots, or null
    const unsigned char* push_ent;  // entries: peer << 4 | row (15 = every row)
    const unsigned long long* peer; // [world] base address of every rank's state buffer (device array)
    long long out_elem_off;         // offset (double2) of this stage's output array in the state buffer
};

template <int N>
struct HParam {
    double2 v[N * N];
};

// ---------------------------------------------------------------------------
// asynchronous copies
// ---------------------------------------------------------------------------
#ifndef HEOM_HOST_EMU
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16_s(unsigned smem_dst_u32, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_dst_u32), "l"(gsrc) : "memory");
}
// predicated form: one instruction under a predicate instead of a branch around the copy
__device__ __forceinline__ void cp_async16_s_if(unsigned smem_dst_u32, const void* gsrc, bool pred) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %2, 0;\n"
        "@p cp.async.cg.shared.global [%0], [%1], 16;\n"
        "}\n" ::"r"(smem_dst_u32),
        "l"(gsrc), "r"((int)pred)
        : "memory");
}
// same with an L2 eviction-priority hint (createpolicy): the stage input and the
// neighbour rows are the only data with reuse (evict_last), y/acc are read once
// per launch (evict_first)
__device__ __forceinline__ void cp_async16_hint(void* smem_dst, const void* gsrc, unsigned long long pol) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}
// ---- TMA bulk copies (cp.async.bulk, 1-D) completing on an mbarrier: the
// contiguous tiles of a group (own y_in, y, stage buffers) are fetched with one
// instruction each by one lane instead of 16 bytes per lane per LDGSTS
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// the same with an L2 eviction-priority hint (createpolicy value)
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, unsigned bytes, unsigned long long* bar,
                                              unsigned long long pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void cp_async16_s_if_hint(unsigned smem_dst_u32, const void* gsrc, bool pred,
                                                     unsigned long long pol) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %2, 0;\n"
        "@p cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %3;\n"
        "}\n" ::"r"(smem_dst_u32),
        "l"(gsrc), "r"((int)pred), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// shared -> global bulk stores (TMA, 1-D), tracked per thread in bulk async-groups.  Used for
// the fused halo push: a row staged in shared memory goes to a peer's array with one
// instruction.  wait_read: the sources may be overwritten; wait_all: the stores are complete.
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int NWAIT>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(NWAIT) : "memory");
}
#endif  // !HEOM_HOST_EMU (the emulation header provides the same names)
