// heom_stage_sym.cuh - interface of the Hermitian-symmetric stage kernel
// (kernel 6, heom_stage_sym.cu) towards the plan code in heom_kernels.cu.
//
// Kernel 6 is the instruction-diet successor of the async row kernel's SYM path
// (kernel 3): same work split (a warp owns 32/N consecutive ADOs, lane = (ADO,
// row)), same difference-form RK4 and the same staging of tiles and neighbour
// rows, but
//   * Hermiticity is used one step further: a lane computes column `row` of
//     P = -i H rho - (damp/2) rho, adds the links' row updates into that same
//     column (no other lane touches it, so the link loop needs no partial-warp
//     barrier and no column read-modify-write), and one pass forms
//     k = P' + P'^dagger at the end.  The second commutator pass of kernel 3
//     (transposed reads of the tile) does not exist;
//   * the RK stage kind (first / middle / last) is a template parameter;
//   * links come from a second table ("links2") that already holds the element
//     offset of the neighbour row and the byte offset of the link's coefficient
//     pair in a shared-memory table pre-multiplied by sqrt(n_eff);
//   * link records reach the lanes through a per-warp shared-memory strip (one
//     coalesced load + broadcast reads) instead of warp shuffles;
//   * slots, groups and element offsets are 32-bit, and only the pointers the
//     kernel needs are passed (fewer constant-bank reloads).
// Every stage output is Hermitian bit for bit.  Results differ from kernel 3's
// by summation order only (~1e-16); both are held to 1e-12 of the reference.
// It applies when every Q_m is diagonal with one non-zero entry, every ADO is
// Hermitian (so the damping rates are real) and H is time independent.  With push
// tables (sharded runs, ShardedDEOM(fused_push=True)) the PUSH instantiation also
// performs the halo exchange: the epilogue leaves the stage output in shared
// memory and the rows other ranks read go from there into their arrays as bulk
// shared->global stores (cp.async.bulk, one instruction per 16N-byte row) over
// NVLink, in flight while the warp loads its next group.
#pragma once
#include "heom_device.cuh"

struct SymArgs {
    const double2* yin;   // stage input
    const double2* y;     // state at the start of the step (middle / last)
    const double2* s1;    // first stage buffer (last stage)
    const double2* s2;    // second stage buffer (last stage)
    double2* out;         // next stage input, or the end-of-step state in the last stage
    const double2* damp;
    const int* link_ptr;
    const int2* links2;   // x: element offset of the neighbour row's storage, y: coefficient byte offset | row | table row << 28
    const double2* cbase; // [K][4]
    const int* kmode;     // [K]: mode | first support row << 8
    const double2* ops;   // [1+M][N*N] (diagonal entries of Q_m are read)
    double2* traj;        // may be null (last stage only)
    const long long* step_base;
    long long ngroups, slot_lo, slot_hi, slot0;
    double a, w;
    int local_step, scramble, nind, nmod, lmax;
    // Sharded runs (heom_shard.cu).  A rank's arrays hold its own ADOs followed by a pool of halo
    // rows that the owning ranks store there (sym_pool_stride(N) elements apart, starting pool_off
    // elements behind the array base); the link table of the owned range points into it.
    // PUSH instantiations: the epilogue stores the rows of the stage output that other ranks read
    // into their pools - push_ptr[owned+1] is a CSR over the owned slots, an entry is
    // (x = row index in the destination's pool, y = staging slot << 8 | peer << 4 | matrix row;
    // the staging slot numbers the distinct (ADO, row) pairs inside a group of 32/N consecutive
    // ADOs, 255 = beyond heom_sym_push_slots()), peer[q] the base
    // address of rank q's state buffer and out_elem_off the offset (double2) of the pool of this
    // stage's output array inside a state buffer (the same on every rank).
    unsigned pool_off;
    const int* push_ptr;
    const int2* push_ent;
    unsigned long long peer[16];    // (kernel parameter space: no load from global memory on the push path)
    long long out_elem_off;
    // dynamic group schedule: global counter (never reset; sched_base = its value at launch), or null
    unsigned* sched;
    unsigned sched_base;
};

// links2 record, resolved for the storage the kernel runs on (a plan converts the table when it
// switches between kernel 6 and kernel 7):
//   x = element offset (double2 units from the array base) of the neighbour row's storage -
//       full matrices: (slot N + r0) N, the row itself; packed: slot N(N+1)/2, the neighbour's
//       triangle; sharded runs, row received from a peer: its place in the halo-row pool;
//   y = ((2k+dir) (L+1) + n_eff) << 5 | r0, plus in bits 28-31 the row of the kernel's offset table
//       that places the lane's element behind x (packed storage): r0, or N for pool rows.
__host__ __device__ inline unsigned sym_link_x(unsigned slot, int r0, int N, bool packed) {
    return packed ? slot * (unsigned)(N * (N + 1) / 2) : (slot * (unsigned)N + (unsigned)r0) * (unsigned)N;
}
__host__ __device__ inline int sym_link_y(int kdir, int neff, int L, int r0, int table_row) {
    return (int)((((unsigned)(kdir * (L + 1) + neff)) << 5) | (unsigned)(r0 & 0xf) | ((unsigned)table_row << 28));
}
// A pool row occupies an even number of elements (whole 32-byte sectors; 128 bytes = one line for
// N = 7, 8) and the pool starts on a 128-byte boundary: the rows arrive as peer writes over NVLink,
// and a write that covers sectors only partly would make the receiving L2 merge it with memory.
__host__ __device__ constexpr int sym_pool_stride(int N) { return (N + 1) & ~1; }
__host__ __device__ constexpr long long sym_pool_offset(long long n_own_max, int elems_per_ado) {
    return (n_own_max * elems_per_ado + 7) & ~7ll;
}
__host__ __device__ inline int sym_link_y(int kdir, int neff, int L, int r0) {
    return ((kdir * (L + 1) + neff) << 5) | (r0 & 0xf);
}

// The kernel-6 view of a difference-form RK4 stage described by StageArgs (run_stage in
// heom_kernels.cu): the last stage reads the first stage buffer from `acc` and the second
// one from `yout` and writes the end-of-step state to `ydst`.
inline SymArgs sym_args_from_stage(const StageArgs& a, const int2* links2) {
    SymArgs s{};
    s.yin = a.yin;
    s.y = a.y;
    s.s1 = a.acc;
    s.s2 = a.yout;
    s.out = a.last ? a.ydst : a.yout;
    s.damp = a.damp;
    s.link_ptr = a.link_ptr;
    s.links2 = links2;
    s.cbase = a.cbase;
    s.kmode = a.kmode;
    s.ops = a.ops;
    s.traj = a.last ? a.traj : nullptr;
    s.step_base = a.step_base;
    s.slot0 = a.slot0;
    s.a = a.a;
    s.w = a.w;
    s.local_step = a.local_step;
    s.scramble = a.scramble;
    s.nind = a.nind;
    s.nmod = a.nmod;
    s.lmax = a.lmax;
    return s;   // (the fused push of sharded runs is set up by heom_shard.cu, not through StageArgs)
}
inline int sym_stage_kind(const StageArgs& a) { return a.first ? 0 : (a.last ? 2 : 1); }

struct SymLaunch {
    SymArgs a;
    const double* H;      // host, N*N interleaved complex (time-independent Hamiltonian)
    int N, K, M, L, B;
    int stage;            // 0 first, 1 middle, 2 last
    int hreal;            // H has no imaginary part
    int packed;           // kernel 7: the ADO arrays hold upper triangles (N(N+1)/2 elements per ADO)
    int push;             // sharded run: PUSH instantiation (a.push_ptr, a.push_ent, a.peer set)
    int prefetch;         // packed only: double-buffered tiles, fetched one group ahead
    int warps;            // 0 = automatic
    int sm_count;
    long long part_lo, part_hi;   // owned slot range
    long long batch_elems;        // nmax * N * N: distance between trajectories in the ADO arrays
    long long traj_bstride;
    void* stream;
    // dynamic group schedule (large hierarchies): device counter and the host's running copy of its
    // value (advanced by groups + warps per launch); null = static stride
    unsigned* sched;
    unsigned* sched_total;
};

// All return 0 on success; on failure *err points to a static message.
int heom_sym_supported(int N, int K, int M, int L, const char** err);
int heom_sym_push_slots(void);   // staging slots per group of the fused push
int heom_sym_convert_links(const int2* links, int2* links2, long long nlinks, int N, int L, int packed, void* stream,
                           const char** err);
int heom_sym_launch(const SymLaunch& L, const char** err);

// Kernel 7 = kernel 6 on packed Hermitian storage.  Every ADO is Hermitian, so only the
// upper triangle (N(N+1)/2 of N^2 elements, row-major) is kept in the arrays the stage
// kernel streams and gathers from: 448 instead of 784 bytes per ADO for N = 7.
struct PackedRun {
    double2* Y;           // full state [nmax][N][N] (packed on entry, unpacked on exit)
    double2* work;        // room for four triangle arrays
    size_t work_bytes;
    SymArgs tables;       // damp, link_ptr, links2 (packed form), cbase, kmode, ops, traj, step_base,
                          // slot0, scramble, nind, nmod, lmax; the array pointers are filled per stage
    const double* H;      // host, N*N interleaved complex
    int N, K, M, L;
    long long nmax, nt;
    double dt;
    int hreal, warps, sm_count;
    int prefetch;         // 1: double-buffered streamed tiles (stage_rows_sym_kernel<..., DB>)
    void* stream;
    unsigned* sched;        // as in SymLaunch
    unsigned* sched_total;
};
int heom_packed_propagate(const PackedRun& r, const char** err);
// full [n][N][N] -> upper triangles [n][N(N+1)/2] (unpack = 0) or back (unpack = 1)
int heom_sym_pack(double2* tri, double2* full, long long n, int N, int unpack, void* stream);
