// heom_shard.cu - multi-GPU propagation with rank-local arrays (one process per GPU).
//
// The reference has nothing here (a single Python loop, pyqed/heom/deom.py:1072-1114);
// SURVEY.md section 8e is the specification: owner-computes over contiguous ranges of the
// blocked-lexicographic storage order, one exchange of boundary rows per RK stage.
//
// Layout of a rank: every ADO array holds the rank's own ADOs [lo, hi) (full matrices, or
// upper triangles while kernel 7 runs) followed by a POOL of halo rows - for each (foreign
// ADO, row) that a link of an owned ADO reads, N elements.  The link table of the owned range
// is written for that layout: a local neighbour is addressed by its local slot, a foreign one
// by its pool row.  The owners store those rows themselves: the stage kernel's epilogue
// (PUSH instantiations of kernels 6 / 7) sends each row of its output that a peer reads from
// shared memory straight into that peer's pool with one bulk store (cp.async.bulk
// shared -> global on a peer address, i.e. over NVLink), in flight while the warp works on
// its next group.  A flag barrier in peer memory closes the stage.  No pack buffers, no
// collective, no host round trip inside a step.
//
// The peers' state buffers are mapped with CUDA IPC (pyqed_heom_shared_*): that works between
// processes on different GPUs of a box (NVLink peer access) and between processes sharing one
// GPU (the CPU-side tests and the one-GPU parity test).
#include "heom_plan.cuh"

namespace {

// ---- flag barrier ------------------------------------------------------------------------
// flags[q] of rank r = last epoch rank q has announced to r.  One warp: lane q announces this
// rank's arrival to rank q (after a system fence: this rank's stores to q's pool come first)
// and waits for q's.  A rank that waits longer than `timeout_ns` sets its error word and
// leaves, so that a lost peer cannot hang the GPU.
__global__ void shard_barrier_kernel(unsigned* my_flags, const unsigned long long* peer_flags, int rank, int world,
                                     unsigned epoch, unsigned long long timeout_ns) {
    const int q = threadIdx.x;
    if (q >= world || q == rank) return;
    __threadfence_system();
    volatile unsigned* theirs = reinterpret_cast<volatile unsigned*>(peer_flags[q]) + rank;
    *theirs = epoch;
    __threadfence_system();
    volatile unsigned* mine = my_flags + q;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((int)(*mine - epoch) < 0) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > timeout_ns) {
            my_flags[16] = 1u + (unsigned)q;   // error word: which peer never arrived
            break;
        }
        __nanosleep(100);
    }
    __threadfence_system();
}

// ---- rows of an array into the peers' pools (initial exchange, and the unfused fallback) -----
// one thread per (entry, element): entry e of the push table belongs to owned slot s(e)
__global__ void shard_push_rows_kernel(const double2* arr, const int* push_ptr, const int2* push_ent, long long n_own,
                                       const unsigned long long* peer, long long dst_pool_off, int N, int packed) {
    const int EL = packed ? N * (N + 1) / 2 : N * N;
    for (long long slot = blockIdx.x * (long long)(blockDim.x / 32) + (threadIdx.x >> 5); slot < n_own;
         slot += (long long)gridDim.x * (blockDim.x / 32)) {
        const int pb = push_ptr[slot], pe = push_ptr[slot + 1];
        const int lane = threadIdx.x & 31;
        for (int q = pb + lane / N; q < pe && lane < (32 / N) * N; q += 32 / N) {
            const int2 ent = push_ent[q];
            const int r = ent.y & 15, j = lane % N;
            double2 v;
            if (packed) {   // as a gather through the triangle delivers the row: (min, max), not conjugated
                const int lo = min(r, j), hi = max(r, j);
                v = arr[slot * EL + lo * N - lo * (lo - 1) / 2 + (hi - lo)];
            } else {
                v = arr[slot * EL + r * N + j];
            }
            reinterpret_cast<double2*>(peer[(ent.y >> 4) & 15])[dst_pool_off + (long long)(unsigned)ent.x * sym_pool_stride(N) + j] = v;
        }
    }
    __threadfence_system();
}

// ---- link table of the owned range: local rows and pool rows -----------------------------------
// need[] = sorted unique items (slot * 8 + row) this rank reads from other ranks.  Written from the
// plan's general link table (slot, meta), so it can be redone for other ranges.
__global__ void shard_localize_links_kernel(int2* links2, const int2* links, const int* link_ptr, long long lo,
                                            long long hi, const long long* need, long long n_need, int N, int L,
                                            int packed, unsigned pool_off, int* bad) {
    const long long slot = lo + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (slot >= hi) return;
    for (int l = link_ptr[slot]; l < link_ptr[slot + 1]; ++l) {
        const int2 r = links[l];
        const long long nb = r.x;
        const int r0 = heom::meta_r0(r.y);
        unsigned x;
        int table_row = r0;
        if (nb >= lo && nb < hi) {
            x = sym_link_x((unsigned)(nb - lo), r0, N, packed != 0);
        } else {
            const long long item = nb * 8 + r0;
            long long a = 0, b = n_need;
            while (a < b) {
                const long long m = (a + b) >> 1;
                if (need[m] < item) a = m + 1;
                else b = m;
            }
            if (a >= n_need || need[a] != item) {
                atomicExch(bad, 1);
                continue;
            }
            x = pool_off + (unsigned)a * (unsigned)sym_pool_stride(N);
            table_row = N;   // pool rows: N consecutive elements
        }
        links2[l] = make_int2((int)x, sym_link_y(heom::meta_kdir(r.y), heom::meta_neff(r.y), L, r0, table_row));
    }
}

__global__ void shard_gather_ids_kernel(int* out, const int* id_of_slot, long long lo, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = id_of_slot[lo + i];
}

int shard_barrier(pyqed_heom_plan* p) {
    auto& sh = p->shard;
    if (sh.world == 1) return 0;
    REQUIRE(sh.device_barrier, "sharded run without a device barrier: drive it stage by stage");
    sh.epoch++;
    unsigned* mine = reinterpret_cast<unsigned*>(sh.peer_flags[sh.rank]);
    shard_barrier_kernel<<<1, 32, 0, p->stream>>>(mine, sh.d_peer + 16, sh.rank, sh.world, sh.epoch,
                                                  20ull * 1000 * 1000 * 1000);
    return post_launch(p, "shard_barrier_kernel");
}

// pool of array `which` (0..3 full arrays; packed: 0..3 = P0..P3 inside the stage arrays):
// offset (double2) from the state buffer base
long long pool_off_in_state(const pyqed_heom_plan* p, int which, bool packed) {
    const auto& sh = p->shard;
    const int N = p->N, EL = packed ? N * (N + 1) / 2 : N * N;
    const long long arr0 = packed ? (long long)(p->array_bytes / sizeof(double2)) + (long long)which * (long long)sh.arr_packed
                                  : (long long)which * (long long)(p->array_bytes / sizeof(double2));
    return arr0 + sym_pool_offset(sh.n_own_max, EL);
}

int push_rows(pyqed_heom_plan* p, const double2* arr, int which, bool packed) {
    auto& sh = p->shard;
    if (sh.world == 1 || sh.pushed_rows == 0) return 0;
    const long long n_own = sh.hi - sh.lo;
    const unsigned grid = (unsigned)std::min<long long>((n_own + 7) / 8, 148 * 8);
    shard_push_rows_kernel<<<grid, 256, 0, p->stream>>>(arr, sh.push_ptr, sh.push_ent, n_own, sh.d_peer,
                                                        pool_off_in_state(p, which, packed), p->N, packed ? 1 : 0);
    return post_launch(p, "shard_push_rows_kernel");
}

}  // namespace

extern "C" {

// ---- peer-visible device memory (CUDA IPC) -------------------------------------------------
int pyqed_heom_shared_alloc(int device, size_t bytes, void** d_ptr, uint8_t* handle64) {
    REQUIRE(d_ptr && handle64 && bytes > 0, "shared_alloc: bad argument");
    CU_TRY(cudaSetDevice(device));
    void* ptr = nullptr;
    CU_TRY(cudaMalloc(&ptr, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
    if (e != cudaSuccess) {
        cudaFree(ptr);
        return fail(std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
    }
    static_assert(sizeof(h) == 64, "CUDA IPC handles are 64 bytes");
    memcpy(handle64, &h, 64);
    *d_ptr = ptr;
    return 0;
}
int pyqed_heom_shared_open(int device, const uint8_t* handle64, void** d_ptr) {
    REQUIRE(d_ptr && handle64, "shared_open: bad argument");
    CU_TRY(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CU_TRY(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
int pyqed_heom_shared_close(int device, void* d_ptr) {
    CU_TRY(cudaSetDevice(device));
    CU_TRY(cudaIpcCloseMemHandle(d_ptr));
    return 0;
}
int pyqed_heom_shared_free(int device, void* d_ptr) {
    CU_TRY(cudaSetDevice(device));
    CU_TRY(cudaFree(d_ptr));
    return 0;
}

// ---- sizes ------------------------------------------------------------------------------------
// State buffer of a rank whose largest peer owns n_own_max ADOs and whose largest pool has
// pool_max rows: four arrays of (n_own_max N^2 + pool_max sym_pool_stride(N)) elements, then the flag block.
int pyqed_heom_shard_state_bytes(pyqed_heom_plan* p, int64_t n_own_max, int64_t pool_max, size_t* state_bytes,
                                 size_t* flag_offset) {
    REQUIRE(p && state_bytes && flag_offset && n_own_max >= 0 && pool_max >= 0, "shard_state_bytes: bad argument");
    const size_t arr = align_up(sizeof(double2) * ((size_t)sym_pool_offset(n_own_max, p->N * p->N) +
                                                   (size_t)pool_max * sym_pool_stride(p->N)));
    *flag_offset = 4 * arr;
    *state_bytes = 4 * arr + 256;
    return 0;
}

// ---- setup --------------------------------------------------------------------------------------
int pyqed_heom_shard_setup(pyqed_heom_plan* p, int rank, int world, int64_t lo, int64_t hi, int64_t n_own_max,
                           int64_t pool_max, const int64_t* d_need, int64_t n_need, const int32_t* d_push_ptr,
                           const int32_t* d_push_ent, int64_t n_push, void* d_state, size_t state_bytes,
                           const uint64_t* peer_state_ptrs, int device_barrier) {
    REQUIRE(p && p->built, "shard_setup: build the hierarchy first");
    REQUIRE(world >= 1 && world <= 16 && rank >= 0 && rank < world, "shard_setup: bad rank / world");
    REQUIRE(lo >= 0 && lo <= hi && hi <= p->nmax && hi - lo <= n_own_max, "shard_setup: bad range");
    REQUIRE(p->B == 1, "shard_setup: batch must be 1");
    REQUIRE(p->links2_built, "shard_setup: the rank-local layout needs kernels 6 / 7 (Hermitian problem, one-entry "
                             "diagonal coupling operators)");
    REQUIRE(n_need <= pool_max && (n_need == 0 || d_need), "shard_setup: bad need list");
    REQUIRE(d_state && peer_state_ptrs && ((uintptr_t)d_state % 256) == 0, "shard_setup: bad state buffer");
    CU_TRY(cudaSetDevice(p->device));
    auto& sh = p->shard;
    const int N = p->N, NN = N * N, PK = N * (N + 1) / 2;
    size_t need_bytes = 0, flag_off = 0;
    if (pyqed_heom_shard_state_bytes(p, n_own_max, pool_max, &need_bytes, &flag_off)) return 1;
    REQUIRE(state_bytes >= need_bytes, "shard_setup: state buffer too small");
    // storage of this run (identical decision on every rank: it depends on the maxima only)
    const size_t arr_full_bytes = flag_off / 4;
    const size_t arr_packed = align_up(sizeof(double2) * ((size_t)sym_pool_offset(n_own_max, PK) +
                                                          (size_t)pool_max * sym_pool_stride(N))) / sizeof(double2);
    const bool packed = (p->kernel == 0 || p->kernel == 7) && p->opt_packed != 0 &&
                        4 * arr_packed * sizeof(double2) <= 3 * arr_full_bytes;
    // link table of the owned range -> local rows / pool rows
    int* d_bad = nullptr;
    CU_TRY(cudaMalloc(&d_bad, sizeof(int)));
    CU_TRY(cudaMemsetAsync(d_bad, 0, sizeof(int), p->stream));
    if (hi > lo) {
        const unsigned blocks = (unsigned)((hi - lo + 127) / 128);
        shard_localize_links_kernel<<<blocks, 128, 0, p->stream>>>(
            p->tab<int2>(p->tl.links2), p->tab<int2>(p->tl.links), p->tab<int>(p->tl.link_ptr), lo, hi,
            (const long long*)d_need, n_need, N, p->L, packed ? 1 : 0,
            (unsigned)sym_pool_offset(n_own_max, packed ? PK : NN), d_bad);
        if (post_launch(p, "shard_localize_links_kernel")) return 1;
    }
    p->links2_mode = -1;   // the table no longer describes the whole hierarchy
    int bad = 0;
    CU_TRY(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    cudaFree(d_bad);
    REQUIRE(!bad, "shard_setup: a link of the owned range reads a row that is not in the need list");
    sh.on = true;
    sh.rank = rank;
    sh.world = world;
    sh.lo = lo;
    sh.hi = hi;
    sh.n_own_max = n_own_max;
    sh.pool_max = pool_max;
    sh.push_ptr = d_push_ptr;
    sh.push_ent = reinterpret_cast<const int2*>(d_push_ent);
    sh.pushed_rows = n_push;
    sh.device_barrier = device_barrier != 0;
    sh.epoch = 0;
    p->d_state = (char*)d_state;
    p->array_bytes = arr_full_bytes;
    sh.arr_full = p->array_bytes / sizeof(double2);
    sh.arr_packed = arr_packed;
    sh.packed = packed;
    // peers: [0, 16) state buffers, [16, 32) flag blocks
    unsigned long long tab[32] = {0};
    for (int q = 0; q < 16; ++q) sh.peer_state[q] = 0;
    for (int q = 0; q < world; ++q) {
        sh.peer_state[q] = peer_state_ptrs[q];
        tab[q] = peer_state_ptrs[q];
        tab[16 + q] = peer_state_ptrs[q] + flag_off;
        sh.peer_flags[q] = tab[16 + q];
    }
    if (!sh.d_peer) CU_TRY(cudaMalloc(&sh.d_peer, sizeof(tab)));
    CU_TRY(cudaMemcpyAsync(sh.d_peer, tab, sizeof(tab), cudaMemcpyHostToDevice, p->stream));
    CU_TRY(cudaMemsetAsync(p->d_state + flag_off, 0, 256, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    p->part_lo = 0;
    p->part_hi = hi - lo;
    return 0;
}

// ---- state in and out ------------------------------------------------------------------------------
int pyqed_heom_shard_set_state(pyqed_heom_plan* p, const double* rho0_host) {
    REQUIRE(p && p->shard.on && rho0_host, "shard_set_state: call shard_setup first");
    CU_TRY(cudaSetDevice(p->device));
    const auto& sh = p->shard;
    const size_t NN = (size_t)p->N * p->N;
    bool h = true;
    for (int i = 0; i < p->N && h; ++i)
        for (int j = 0; j < p->N; ++j) {
            const double* x = rho0_host + 2 * (i * p->N + j);
            const double* y = rho0_host + 2 * (j * p->N + i);
            if (x[0] != y[0] || x[1] != -y[1]) {
                h = false;
                break;
            }
        }
    REQUIRE(h, "shard_set_state: the rank-local layout needs a Hermitian initial state");
    p->herm_state = true;
    CU_TRY(cudaMemsetAsync(p->d_state, 0, 4 * p->array_bytes, p->stream));
    if (p->slot0 >= sh.lo && p->slot0 < sh.hi)
        CU_TRY(cudaMemcpyAsync(p->arr(ARR_Y) + (size_t)(p->slot0 - sh.lo) * NN, rho0_host, sizeof(double2) * NN,
                               cudaMemcpyHostToDevice, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    return 0;
}

// owned ADOs (storage order of the range) and their reference ids -> host
int pyqed_heom_shard_get_owned(pyqed_heom_plan* p, double* ados_host, int32_t* ids_host) {
    REQUIRE(p && p->shard.on && ados_host && ids_host, "shard_get_owned: call shard_setup first");
    CU_TRY(cudaSetDevice(p->device));
    const auto& sh = p->shard;
    const long long n = sh.hi - sh.lo;
    if (n == 0) return 0;
    const size_t NN = (size_t)p->N * p->N;
    CU_TRY(cudaMemcpyAsync(ados_host, p->arr(ARR_Y), sizeof(double2) * n * NN, cudaMemcpyDeviceToHost, p->stream));
    CU_TRY(cudaMemcpyAsync(ids_host, p->tab<int>(p->tl.id_of_slot) + sh.lo, sizeof(int) * n, cudaMemcpyDeviceToHost,
                           p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    return 0;
}

// ---- propagation ------------------------------------------------------------------------------------
// One RK4 stage of the sharded run (difference form, kernels 6 / 7 with the fused push), without
// the barrier that must follow it.  which arrays: see run_stage (heom_kernels.cu) / heom_packed_propagate.
static int shard_stage(pyqed_heom_plan* p, int64_t step, int stage) {
    auto& sh = p->shard;
    const int N = p->N, EL = sh.packed ? N * (N + 1) / 2 : N * N;
    const long long n_own = sh.hi - sh.lo;
    double2* A[4];
    if (sh.packed)
        for (int q = 0; q < 4; ++q) A[q] = p->arr(ARR_SA) + (size_t)q * sh.arr_packed;
    else {
        A[0] = p->arr(ARR_Y); A[1] = p->arr(ARR_SA); A[2] = p->arr(ARR_SB); A[3] = p->arr(ARR_ACC);
    }
    const TableLayout& t = p->tl;
    SymLaunch s{};
    SymArgs& a = s.a;
    const double dt = p->ctx_dt;
    a.y = A[0];
    int out = 0;
    switch (stage) {
        case 0: a.yin = A[0]; out = 1; a.a = dt / 2; s.stage = 0; break;
        case 1: a.yin = A[1]; out = 2; a.a = dt / 2; s.stage = 1; break;
        case 2: a.yin = A[2]; out = 3; a.a = dt; s.stage = 1; break;
        default: a.yin = A[3]; a.s1 = A[1]; a.s2 = A[2]; out = 0; a.a = 2.0 / dt; a.w = dt / 6; s.stage = 2; break;
    }
    a.out = A[out];
    a.damp = p->tab<double2>(t.damp) + sh.lo;
    a.link_ptr = p->tab<int>(t.link_ptr) + sh.lo;
    a.links2 = p->tab<int2>(t.links2);
    a.cbase = p->tab<double2>(t.cbase);
    a.kmode = p->tab<int>(t.kmode);
    a.ops = p->tab<double2>(t.ops_base);
    a.traj = stage == 3 ? p->ctx_traj : nullptr;
    a.step_base = p->tab<long long>(t.step_base);
    a.slot0 = (p->slot0 >= sh.lo && p->slot0 < sh.hi) ? p->slot0 - sh.lo : -1;
    if (a.slot0 < 0) a.traj = nullptr;
    a.local_step = (int)step;
    a.scramble = 0;
    a.nind = p->K;
    a.nmod = p->M;
    a.lmax = p->L;
    a.pool_off = (unsigned)sym_pool_offset(sh.n_own_max, EL);
    a.push_ptr = sh.push_ptr;
    a.push_ent = sh.push_ent;
    for (int q = 0; q < 16; ++q) a.peer[q] = sh.peer_state[q];
    a.out_elem_off = pool_off_in_state(p, out, sh.packed);
    s.push = (sh.world > 1 && sh.pushed_rows > 0) ? 1 : 0;
    s.H = reinterpret_cast<const double*>(p->H.data());
    s.N = N; s.K = p->K; s.M = p->M; s.L = p->L; s.B = 1;
    s.hreal = (p->h_real && p->opt_hreal != 0) ? 1 : 0;
    s.packed = sh.packed ? 1 : 0;
    s.warps = p->warps;
    s.sm_count = sm_count_of(p->device);
    s.part_lo = 0;
    s.part_hi = n_own;
    s.batch_elems = 0;
    s.traj_bstride = 0;
    s.stream = p->stream;
    s.sched = p->opt_dynsched != 0 ? p->tab<unsigned>(t.sched) : nullptr;
    s.sched_total = &p->sched_total;
    if (n_own <= 0) return 0;
    if (p->timing) {
        if (p->ev_used == p->ev.size()) {
            cudaEvent_t e0, e1;
            CU_TRY(cudaEventCreate(&e0));
            CU_TRY(cudaEventCreate(&e1));
            p->ev.emplace_back(e0, e1);
        }
        CU_TRY(cudaEventRecord(p->ev[p->ev_used].first, p->stream));
    }
    const char* err = "";
    if (heom_sym_launch(s, &err)) return fail(std::string("stage_rows_sym_kernel (sharded) launch: ") + err);
    p->launches++;
    p->sym_launches++;
    if (p->timing) {
        CU_TRY(cudaEventRecord(p->ev[p->ev_used].second, p->stream));
        p->ev_used++;
    }
    if (p->debug_sync) {
        cudaError_t e = cudaStreamSynchronize(p->stream);
        if (e != cudaSuccess) return fail(std::string("stage_rows_sym_kernel (sharded) exec: ") + cudaGetErrorString(e));
    }
    return 0;
}

// begin: context, rho_sys(0), (packed) Y -> P0, then this rank's rows of the stage-0 input into
// the peers' pools.  The caller follows with a barrier (device or host).
int pyqed_heom_shard_begin(pyqed_heom_plan* p, double dt, int64_t nt, double* d_traj) {
    REQUIRE(p && p->shard.on, "shard_begin: call shard_setup first");
    REQUIRE(p->herm_state && nt >= 0, "shard_begin: needs a Hermitian state (shard_set_state)");
    CU_TRY(cudaSetDevice(p->device));
    auto& sh = p->shard;
    const int N = p->N, NN = N * N;
    p->ctx_tdep = p->ctx_use_fs = p->ctx_use_fc = false;
    p->ctx_dt = dt;
    p->ctx_nt = nt;
    p->ctx_traj = (double2*)d_traj;
    CU_TRY(cudaMemsetAsync(p->d_tables + p->tl.step_base, 0, sizeof(long long), p->stream));
    CU_TRY(cudaMemsetAsync(p->d_tables + p->tl.sched, 0, sizeof(unsigned) * 4, p->stream));
    p->sched_total = 0;
    const long long n_own = sh.hi - sh.lo;
    if (p->ctx_traj && p->slot0 >= sh.lo && p->slot0 < sh.hi)
        CU_TRY(cudaMemcpyAsync(p->ctx_traj, p->arr(ARR_Y) + (size_t)(p->slot0 - sh.lo) * NN, sizeof(double2) * NN,
                               cudaMemcpyDeviceToDevice, p->stream));
    if (sh.packed) {
        double2* P0 = p->arr(ARR_SA);
        if (n_own > 0 && heom_sym_pack(P0, p->arr(ARR_Y), n_own, N, 0, p->stream)) return fail("sym_pack_kernel launch failed");
        p->launches++;
        if (push_rows(p, P0, 0, true)) return 1;
    } else {
        if (push_rows(p, p->arr(ARR_Y), 0, false)) return 1;
    }
    p->ctx_valid = true;
    return 0;
}

int pyqed_heom_shard_stage(pyqed_heom_plan* p, int64_t step, int stage) {
    REQUIRE(p && p->shard.on && p->ctx_valid, "shard_stage: call shard_begin first");
    REQUIRE(stage >= 0 && stage <= 3 && step >= 0 && step < p->ctx_nt, "shard_stage: bad step/stage");
    CU_TRY(cudaSetDevice(p->device));
    return shard_stage(p, step, stage);
}

// end: (packed) P0 -> Y
int pyqed_heom_shard_end(pyqed_heom_plan* p) {
    REQUIRE(p && p->shard.on && p->ctx_valid, "shard_end: call shard_begin first");
    CU_TRY(cudaSetDevice(p->device));
    auto& sh = p->shard;
    const long long n_own = sh.hi - sh.lo;
    if (sh.packed && n_own > 0) {
        if (heom_sym_pack(p->arr(ARR_SA), p->arr(ARR_Y), n_own, p->N, 1, p->stream)) return fail("sym_unpack_kernel launch failed");
        p->launches++;
    }
    p->packed_steps += sh.packed ? p->ctx_nt : 0;
    p->ctx_valid = false;
    return 0;
}

int pyqed_heom_shard_barrier(pyqed_heom_plan* p) {
    REQUIRE(p && p->shard.on, "shard_barrier: call shard_setup first");
    CU_TRY(cudaSetDevice(p->device));
    return shard_barrier(p);
}

// Whole run on the device: begin, barrier, nt x 4 x (stage with fused push, barrier), end.
// Needs the device barrier (every rank on its own GPU).
int pyqed_heom_shard_propagate(pyqed_heom_plan* p, double dt, int64_t nt, double* d_traj) {
    if (pyqed_heom_shard_begin(p, dt, nt, d_traj)) return 1;
    if (shard_barrier(p)) return 1;
    for (int64_t i = 0; i < nt; ++i)
        for (int st = 0; st < 4; ++st) {
            if (shard_stage(p, i, st)) return 1;
            if (shard_barrier(p)) return 1;
        }
    return pyqed_heom_shard_end(p);
}

// 0 = every barrier so far completed; q + 1 = rank q never arrived (barrier timed out)
int pyqed_heom_shard_error(pyqed_heom_plan* p, int* code) {
    REQUIRE(p && p->shard.on && code, "shard_error: call shard_setup first");
    CU_TRY(cudaSetDevice(p->device));
    unsigned v = 0;
    CU_TRY(cudaMemcpyAsync(&v, reinterpret_cast<unsigned*>(p->shard.peer_flags[p->shard.rank]) + 16, sizeof(unsigned),
                           cudaMemcpyDeviceToHost, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    *code = (int)v;
    return 0;
}

}  // extern "C"
