// hier_emu.cpp - runs the hierarchy builder kernels (pyqed_b200/csrc/heom_hierarchy.cuh: storage
// order, keys, damping rates, CSR links) on the CPU through tests/_shim/cuda_emu.h, the way
// pyqed_heom_build_hierarchy (heom_kernels.cu) sequences them.  TEST INFRASTRUCTURE ONLY.
#include "cuda_emu.h"

#include "../../pyqed_b200/csrc/heom_hierarchy.cuh"

extern "C" int emu_build_hierarchy(int K, int L, int order, const long long* pascal, int side, long long nmax,
                                   const double* expn, const int* mode, unsigned char* keys, int* id_of_slot,
                                   int* slot_of_id, double* damp, int* link_ptr, int* links, long long nlinks,
                                   int* lex2slot, int* slot2lex) {
    HierArgs h;
    h.pascal = pascal;
    h.side = side;
    h.K = K;
    h.L = L;
    h.order = order;
    h.nmax = nmax;
    h.keys = keys;
    h.id_of_slot = id_of_slot;
    h.slot_of_id = slot_of_id;
    h.damp = reinterpret_cast<double2*>(damp);
    h.link_ptr = link_ptr;
    h.links = reinterpret_cast<int2*>(links);
    h.expn = reinterpret_cast<const double2*>(expn);
    h.mode = mode;
    h.lex2slot = lex2slot;
    h.slot2lex = slot2lex;
    const int threads = 128;
    const unsigned blocks = (unsigned)((nmax + threads - 1) / threads);
    if (order == 2) {
        const long long nblk = (nmax + ORDER2_BLOCK - 1) / ORDER2_BLOCK;
        emu::launch(hier_blockperm_kernel, (unsigned)((nblk + 63) / 64), 64u, 0, h);
    }
    emu::launch(hier_keys_kernel, blocks, (unsigned)threads, 0, h);
    // cub::DeviceScan::ExclusiveSum over nmax + 1 entries in the product
    int run = 0;
    for (long long i = 0; i <= nmax; ++i) {
        const int c = link_ptr[i];
        link_ptr[i] = run;
        run += c;
    }
    if (run != nlinks) return 1;
    emu::launch(hier_links_kernel, blocks, (unsigned)threads, 0, h);
    return 0;
}
