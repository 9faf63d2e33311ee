/*
 * pyqed_heom.h - C ABI of the B200-native HEOM/DEOM RK4 propagation path.
 *
 * The reference (binggu56/pyqed) has no FFI layer for this path: its boundary is
 * the Python class API of pyqed/heom/deom.py (DEOMSolver, :953-1114) and
 * pyqed/HEOM/heom.py (HEOMSolver, :161-205).  The entry points below are what a
 * ctypes binding inside those classes would call; each one names the reference
 * code it replaces.  INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure;
 *     pyqed_heom_last_error() then describes the failure (thread-local).
 *   - complex128 is passed as interleaved (re, im) doubles, matrices row-major.
 *   - "host" pointers are ordinary (or pinned) host memory; "d_" pointers are
 *     CUDA device memory owned by the caller (e.g. a torch tensor).  The
 *     library never allocates the large arrays itself.
 *   - ADO order on the ABI is always the reference's flat id
 *     (gen_hash_value, deom.py:555-565), whatever the device layout is.
 *   - no exceptions, no callbacks into the host language; one plan per host
 *     thread at a time (the reference solver is not re-entrant either).
 */
#ifndef PYQED_HEOM_H
#define PYQED_HEOM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pyqed_heom_plan pyqed_heom_plan;

/* ABI version of this header (bumped on incompatible change). */
int pyqed_heom_version(void);

/* Message for the last failure on this thread ("" if none). */
const char* pyqed_heom_last_error(void);

/* nmax = C(lmax + nind, lmax): number of ADOs.  Replaces the Pascal-table
 * lookup in DEOMSolver.init_ (deom.py:1048-1060).  Returns -1 on overflow or
 * invalid arguments. */
int64_t pyqed_heom_hierarchy_size(int nind, int lmax);

/* Create / destroy a plan for an N-level system with K = nind dissipatons
 * grouped into M = nmod coupling operators, hierarchy depth lmax and `batch`
 * independent trajectories (batch >= 1; trajectories differ only in their
 * initial state and field tables).  Replaces DEOMSolver.__init__ / check_
 * (deom.py:958-1046). */
int pyqed_heom_plan_create(pyqed_heom_plan** plan, int device, int nsys, int nind,
                           int nmod, int lmax, int batch);
void pyqed_heom_plan_destroy(pyqed_heom_plan* plan);

/* System Hamiltonian H and its dipole mu, both N x N complex128 on the host;
 * H(t) = H + mu * f(t) (generate_time, deom.py:681).  mu may be NULL (= 0).
 * Replaces set_system / set_system_dipole (deom.py:998-1008). */
int pyqed_heom_set_system(pyqed_heom_plan* plan, const double* H, const double* mu);

/* Coupling operators Q[M][N][N] and their dipoles; Q_m(t) = Q_m + Qdip_m g(t)
 * (deom.py:684-686).  Qdip may be NULL.  Replaces set_coupling /
 * set_coupling_dipole (deom.py:1010-1021). */
int pyqed_heom_set_coupling(pyqed_heom_plan* plan, const double* Q, const double* Qdip);

/* Bath exponents: expn, etal, etar, etaa complex128[K], mode int64[K] with
 * values in [0, M).  These are the five attributes the reference solver reads
 * from its Bath object (deom.py:1040-1041, 1070). */
int pyqed_heom_set_bath(pyqed_heom_plan* plan, const double* expn, const double* etal,
                        const double* etar, const double* etaa, const int64_t* mode);

/* ADO storage order on the device: 0 = reference id order (tier-major),
 * 1 = lexicographic in the multi-index (gather locality, small halos),
 * 2 = lexicographic with top-tier ADOs grouped inside aligned 64-slot blocks
 *     (same locality, balanced link counts per warp).  Must be called before
 * pyqed_heom_table_bytes.  Results on the ABI are unaffected. */
int pyqed_heom_set_order(pyqed_heom_plan* plan, int order);

/* Sizes of the two caller-owned device buffers: the index/coefficient tables
 * and the ADO state (4 arrays of [batch, nmax, N, N] complex128). */
int pyqed_heom_table_bytes(pyqed_heom_plan* plan, size_t* bytes);
int pyqed_heom_state_bytes(pyqed_heom_plan* plan, size_t* bytes);

/* Hand the buffers (256-byte aligned) and a cudaStream_t (as void*, NULL =
 * default stream) to the plan.  All later calls are enqueued on that stream. */
int pyqed_heom_bind(pyqed_heom_plan* plan, void* d_tables, size_t table_bytes,
                    void* d_state, size_t state_bytes, void* stream);

/* Build keys, damping rates and the n+-1 neighbour/coefficient tables on the
 * device.  Replaces init_ / gen_keys (deom.py:608-638, 1048-1064) and the
 * per-call hash_plus / hash_minus of generate_dot_element (deom.py:653-661). */
int pyqed_heom_build_hierarchy(pyqed_heom_plan* plan);

/* keys[nmax][K] (uint8, reference id order) -> host.  DEOMSolver.keys. */
int pyqed_heom_get_keys(pyqed_heom_plan* plan, uint8_t* keys_host);

/* Zero every ADO and set ADO 0 of trajectory b to rho0[b] (host, [batch][N][N]).
 * Replaces the allocation block of DEOMSolver.run (deom.py:1084-1092). */
int pyqed_heom_set_state(pyqed_heom_plan* plan, const double* rho0_host);

/* Whole ADO array [batch][nmax][N][N] host <-> device, reference id order.
 * get replaces reading DEOMSolver.ddos after run(). */
int pyqed_heom_load_ados(pyqed_heom_plan* plan, const double* ados_host);
int pyqed_heom_get_ados(pyqed_heom_plan* plan, double* ados_host);

/* Propagate nt steps of size dt from the current state.
 *   method 0: classical RK4 with stage times t, t+dt/2, t+dt/2, t+dt
 *             (rk4, deom.py:725-766; rem_cal / generate_dot_element :641-673)
 *   method 1: explicit Euler (one stage).
 * fsys / fcoup: host tables [batch][nt][3] of the pulse functions sampled at
 * i*dt, i*dt+dt/2, i*dt+dt for every step (NULL = identically zero); they are
 * what generate_time (deom.py:676-687) would evaluate.
 * d_traj: device [batch][nt+1][N][N] complex128 or NULL; entry 0 receives the
 * system density matrix before the first step, entry i+1 after step i
 * (ddos_save of DEOMSolver.run, deom.py:1094-1113).
 * Asynchronous on the plan's stream. */
int pyqed_heom_propagate(pyqed_heom_plan* plan, double dt, int64_t nt, const double* fsys,
                         const double* fcoup, double* d_traj, int method);

/* The same propagation split up so that a multi-GPU driver can interleave the
 * halo exchange: propagate_begin takes the arguments of pyqed_heom_propagate
 * (RK4 only), propagate_stage launches stage 0..3 of step `step`.  The stage
 * outputs live in arrays 1 (stages 0, 2), 2 (stage 1) and 0 (stage 3). */
int pyqed_heom_propagate_begin(pyqed_heom_plan* plan, double dt, int64_t nt, const double* fsys,
                               const double* fcoup, double* d_traj);
int pyqed_heom_propagate_stage(pyqed_heom_plan* plan, int64_t step, int stage);

/* Multi-GPU (one process per GPU): this rank owns storage slots [lo, hi) and
 * only advances those; all four ADO arrays stay full-size so that neighbour
 * reads use global slot numbers.  halo_pack copies the listed items of array
 * `array_id` (0 state, 1/2 stage buffers, 3 accumulator) into a contiguous
 * buffer [batch][n_items][N or N*N] (unpack = 0) or back (unpack = 1).
 * An item is slot*8+row when row_items = 1 (one matrix row, the only part of a
 * neighbour a diagonal Q_m needs) or a slot when row_items = 0 (whole ADO). */
int pyqed_heom_set_partition(pyqed_heom_plan* plan, int64_t slot_lo, int64_t slot_hi);
int pyqed_heom_halo_pack(pyqed_heom_plan* plan, int array_id, const int32_t* d_items,
                         int64_t n_items, int row_items, double* d_buf, int unpack);

/* Peer-memory variant of the halo exchange: the ranks' state buffers are mapped
 * into each other's address space (e.g. torch symmetric memory); this rank
 * stores the listed items of its array `array_id` directly into the same
 * positions of the destination ranks' arrays.  d_items holds the items grouped
 * by destination, dest_offsets[world+1] (host) the group boundaries and
 * peer_state_ptrs[world] (host) the device address of every rank's state
 * buffer.  The caller synchronises the ranks before the next stage reads. */
int pyqed_heom_halo_push(pyqed_heom_plan* plan, int array_id, const int32_t* d_items,
                         int64_t n_items, int row_items, const int64_t* dest_offsets,
                         const uint64_t* peer_state_ptrs, int world);

/* rho_n <- A rho_n (side 0) or rho_n A (side 1) for every ADO of every
 * trajectory, A an N x N complex128 host matrix.  Replaces operator_action_ddos
 * (deom.py:945-950); with propagate it gives HEOM-space correlation functions
 * by time propagation (pyqed/deom.py:921-952). */
int pyqed_heom_apply_operator(pyqed_heom_plan* plan, const double* op_host, int side);

/* Single-exponential chain HEOM by explicit Euler with the reference's in-place
 * sequential sweep: the Euler `_heom` of pyqed/oqs.py:1808-1875 (what
 * examples/heom.py imports) and, with the N*N unit matrices as a batch, the
 * Liouville-space `_heom_propagator` (pyqed/HEOM/heom.py:349-413;
 * double_update0 = 1 gives the pyqed/oqs.py:1877-1941 variant, whose loop
 * advances ADO 0 twice per step).  d_ado [batch][nado][N][N] is read and
 * updated in place; d_obs [batch][n_e][nt] receives Tr(e rho_0) after each step
 * (n_e may be 0).  Plan-free; synchronous on `stream`. */
int pyqed_heom_chain_euler(int device, void* stream, int N, int nado, int batch, const double* H,
                           const double* S, double gamma, double c_re, double c_im, double dt,
                           int64_t nt, int double_update0, double* d_ado, const double* e_ops_host,
                           int n_e, double* d_obs);

/* Fused form of the peer-memory halo for the async row kernel (kernel 3): the
 * stage kernel's epilogue stores every output row another rank needs straight
 * into that rank's array, so the exchange overlaps the stage itself.
 * d_push_ptr[owned+1] / d_push_ent: CSR over this rank's owned slots, entry =
 * peer << 4 | row (row 15 = every row of the ADO).  Pass d_push_ptr = NULL to
 * switch it off.  The caller still synchronises the ranks between stages. */
int pyqed_heom_set_push_table(pyqed_heom_plan* plan, const int32_t* d_push_ptr,
                              const uint8_t* d_push_ent, const uint64_t* peer_state_ptrs, int world);

/* ---- Sharded propagation with rank-local arrays (one process per GPU; csrc/heom_shard.cu) ----
 * Specification: SURVEY.md section 8e (the reference is a single Python loop, deom.py:1072-1114).
 * Each rank owns the storage slots [lo, hi) of the blocked-lexicographic order (set_order 2).
 * Its four ADO arrays hold n_own_max ADOs followed by a pool of pool_max halo rows (N elements
 * each): one row for every (foreign ADO, matrix row) a link of an owned ADO reads.  The owners
 * store those rows there themselves, from the stage kernel's epilogue, with bulk stores on
 * peer addresses (NVLink); a flag barrier in peer memory closes every stage.  Only for the
 * problems kernels 6 / 7 take (Hermitian ADOs, one-entry diagonal Q_m, time-independent H),
 * batch = 1.  Everything else shards through set_partition / halo_pack / halo_push.
 *
 * shared_alloc / open / close / free: device memory that other processes on this box can map
 * (cudaMalloc + CUDA IPC; handle = 64 bytes to be sent to the peers by the caller's transport).
 * shard_state_bytes: size of a rank's state buffer and the offset of its flag block, given
 * the largest owned range and the largest pool over all ranks (identical layouts everywhere).
 * shard_setup: bind that buffer (allocated with shared_alloc), the peers' mapped addresses,
 * this rank's sorted unique need list d_need[n_need] (items slot*8+row, device, int64) and
 * its push table (CSR over the owned slots: d_push_ptr[n_own+1]; entry i = two int32,
 * d_push_ent[2i] = row index in the destination's pool, d_push_ent[2i+1] = peer << 4 | row),
 * and rewrite the links of the owned range to local slots / pool rows.  Collective in spirit:
 * every rank calls it, then the caller synchronises the ranks once on the host.
 * device_barrier = 1: every rank has its own GPU and shard_propagate may spin on peer flags;
 * 0 (ranks share a GPU): drive the run with shard_begin / shard_stage / shard_end and a host
 * barrier (stream synchronise + transport barrier) after begin and after every stage.
 * shard_set_state: zero the arrays, ADO 0 = rho0 on its owner (callers barrier afterwards).
 * shard_get_owned: this rank's ADOs and their reference ids (DEOMSolver.ddos is their union).
 * shard_error: 0, or q+1 if a device barrier gave up waiting for rank q (20 s). */
int pyqed_heom_shared_alloc(int device, size_t bytes, void** d_ptr, uint8_t* handle64);
int pyqed_heom_shared_open(int device, const uint8_t* handle64, void** d_ptr);
int pyqed_heom_shared_close(int device, void* d_ptr);
int pyqed_heom_shared_free(int device, void* d_ptr);
int pyqed_heom_shard_state_bytes(pyqed_heom_plan* plan, int64_t n_own_max, int64_t pool_max,
                                 size_t* state_bytes, size_t* flag_offset);
int pyqed_heom_shard_setup(pyqed_heom_plan* plan, int rank, int world, int64_t lo, int64_t hi,
                           int64_t n_own_max, int64_t pool_max, const int64_t* d_need, int64_t n_need,
                           const int32_t* d_push_ptr, const int32_t* d_push_ent, int64_t n_push,
                           void* d_state, size_t state_bytes, const uint64_t* peer_state_ptrs,
                           int device_barrier);
int pyqed_heom_shard_set_state(pyqed_heom_plan* plan, const double* rho0_host);
int pyqed_heom_shard_get_owned(pyqed_heom_plan* plan, double* ados_host, int32_t* ids_host);
int pyqed_heom_shard_begin(pyqed_heom_plan* plan, double dt, int64_t nt, double* d_traj);
int pyqed_heom_shard_stage(pyqed_heom_plan* plan, int64_t step, int stage);
int pyqed_heom_shard_end(pyqed_heom_plan* plan);
int pyqed_heom_shard_barrier(pyqed_heom_plan* plan);
int pyqed_heom_shard_propagate(pyqed_heom_plan* plan, double dt, int64_t nt, double* d_traj);
int pyqed_heom_shard_error(pyqed_heom_plan* plan, int* code);

/* Tr(op_e rho) for npts density matrices per trajectory:
 * d_rho [batch][npts][N][N] (device), ops_host [n_ops][N][N] (host),
 * d_out [batch][n_ops][npts] complex128 (device).  Replaces
 * (p1 @ ddos[0]).trace() (deom.py:1104,1113) and obs (superoperator.py:313). */
int pyqed_heom_expectation(pyqed_heom_plan* plan, const double* d_rho, int64_t npts,
                           const double* ops_host, int n_ops, double* d_out);

/* Copy `bytes` device<->host on the plan's stream and wait (so a ctypes caller
 * needs no CUDA runtime binding of its own). */
int pyqed_heom_memcpy_d2h(pyqed_heom_plan* plan, void* host, const void* dev, size_t bytes);
int pyqed_heom_memcpy_h2d(pyqed_heom_plan* plan, void* dev, const void* host, size_t bytes);
int pyqed_heom_synchronize(pyqed_heom_plan* plan);

/* Number of kernels this plan has launched so far (bench.py's gpu_launches),
 * and CUDA-event timing of the stage kernel accumulated since the last reset:
 * total milliseconds and launch count (roofline.achieved in bench.py). */
int64_t pyqed_heom_launch_count(pyqed_heom_plan* plan);
int pyqed_heom_stage_timing(pyqed_heom_plan* plan, int enable, double* total_ms,
                            int64_t* launches);

/* Tuning knobs (0 = library default): kernel 0 auto, 1 row-per-lane kernel
 * (N <= 8), 2 generic one-CTA-per-ADO kernel, 3 row-per-lane kernel with
 * cp.async staging (N <= 8, diagonal Q_m), 4 cluster-resident propagation
 * (whole hierarchy in distributed shared memory, small hierarchies only;
 * chosen automatically when it fits), 6 Hermitian-symmetric stage kernel and
 * 7 the same on packed (upper-triangle) storage for whole propagate calls.
 * Kernel 0 picks 7, then 6, then 3 / 1 / 2 by applicability: 6 and 7 need
 * Hermitian ADOs, one-entry diagonal Q_m and a time-independent H; 7 also one
 * trajectory and the whole hierarchy on this GPU; both fall back to kernel 3
 * where they do not apply.  8 and 9: ONE cooperative launch for all steps of a
 * small hierarchy with 8 < N <= 32 (or any N when asked for), the ADOs
 * synchronised by per-ADO release/acquire flags instead of kernel boundaries;
 * 9 (Hermitian problems, at most two ADOs per SM, sparse operators) keeps one
 * CTA per ADO with the state in registers / shared memory and exchanges packed
 * Hermitian stage outputs; kernel 0 picks 9, then 8, for N > 8.  Choose the kernel and set system, coupling and bath
 * before pyqed_heom_table_bytes: kernels 6 / 7 add a second link table to the
 * table buffer.  warps per CTA for kernels 1, 3, 6 and 7;
 * use_graph: reserved, must be 0 (small hierarchies are propagated by a single
 * cluster-resident launch instead of a graph). */
int pyqed_heom_set_tuning(pyqed_heom_plan* plan, int kernel, int warps_per_cta,
                          int use_graph);

/* Named integer options (-1 = automatic, 0 = off, 1 = on):
 *   "qdiag"      use the element-wise neighbour path when every Q_m is diagonal
 *   "hermitian"  when H, Q, the bath (real expn, etar = conj(etal), etaa > 0)
 *                and the loaded state keep every ADO Hermitian, fetch a
 *                neighbour's column entries as the conjugate of its row
 *   "sym"        async row kernel: when "hermitian" holds and every Q_m has a
 *                single non-zero diagonal entry, k is Hermitian as well; form
 *                -i[H, rho] from one product (rho H = (H rho)^dagger) and take a
 *                link's column update as the conjugate of its row update
 *   "real_h"     use real arithmetic for the H products when H and mu are real
 *   "rk13"       0 = do not use the async row kernel (kernel 3), whose RK4 is in
 *                difference form - the three stage buffers are kept and combined
 *                in the last stage instead of a running accumulator, 13 instead
 *                of 16 array passes per step (stage outputs then live in arrays
 *                1, 2, 3, 0) - and fall back to the accumulator form of kernel 1
 *   "resident"   allow the cluster-resident kernels for small hierarchies
 *                (4 = prefer the row-per-lane variant, kernel 4, over the
 *                element-parallel kernel 5)
 *   "prefetch"   1 = kernel 7 keeps the streamed tiles in two buffer sets and fetches
 *                them one group ahead (default 0)
 *   "packed"     0 = never run on packed Hermitian storage (kernel 7)
 *   "dynsched"   0 = kernels 6 / 7 visit their groups with a static stride instead of
 *                drawing them from a global work counter (the counter keeps the groups
 *                in flight on the whole chip inside one short window of the storage
 *                order, which is what lets neighbour rows hit in L2)
 *   "dataflow_tma" 0 = never use kernel 9 (kernel 8 takes its problems)
 *   "debug_sync" synchronise and check after every launch
 * get_info reports resolved properties ("qdiag", "q_diagonal", "hermitian", "sym", "real_h", "off_link_ptr", "off_links" (byte offsets into the table buffer),
 * "array_bytes", "part_lo", "part_hi", "nlinks", "nmax", "slot0", "table_bytes", "sym_launches" (stage launches done by
 * kernel 6), "packed_steps" (RK4 steps done by kernel 7), "dataflow_launches" (propagations done by kernels 8 / 9),
 * "dataflow_tma_launches" (... by kernel 9), "dataflow_dense_launches" (... by its dense-H instantiation)); -1 for an
 * unknown name. */
int pyqed_heom_set_option(pyqed_heom_plan* plan, const char* name, int value);
int64_t pyqed_heom_get_info(pyqed_heom_plan* plan, const char* name);

#ifdef __cplusplus
}
#endif
#endif /* PYQED_HEOM_H */
