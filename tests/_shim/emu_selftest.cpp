// emu_selftest.cpp - checks that tests/_shim/cuda_emu.h rejects the copy operands the hardware
// rejects (misaligned cp.async / cp.async.bulk addresses, bulk sizes that are not a multiple of
// 16, shared-memory operands outside the CTA's allocation).  argv[1] selects the case; case 0 is
// legal and must exit 0, every other case must abort.  TEST INFRASTRUCTURE ONLY.
#include "cuda_emu.h"

static void probe(int which, char* g) {
    HEOM_DYN_SMEM(char, s);
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(s + 256);
    if (threadIdx.x != 0) return;
    mbar_init(bar, 1);
    switch (which) {
    case 0:   // legal: every primitive once
        cp_async16(s, g);
        cp_async_commit();
        cp_async_wait<0>();
        mbar_expect_tx(bar, 112);
        bulk_g2s(s + 16, g + 16, 112, bar);
        mbar_wait(bar, 0);
        bulk_s2g(g + 128, s + 16, 112);
        bulk_commit();
        bulk_wait_all();
        break;
    case 1: cp_async16(s + 8, g); break;              // shared destination off by 8
    case 2: cp_async16(s, g + 8); break;              // global source off by 8
    case 3: bulk_g2s(s, g, 104, bar); break;          // size not a multiple of 16
    case 4: bulk_g2s(s + 8, g, 112, bar); break;      // shared destination misaligned
    case 5: bulk_g2s(s, g + 8, 112, bar); break;      // global source misaligned
    case 6: bulk_s2g(g + 8, s, 112); break;           // global destination misaligned
    case 7: bulk_s2g(g, s + 512 - 16, 32); break;     // shared source runs past the allocation
    case 8: mbar_init(reinterpret_cast<unsigned long long*>(s + 4), 1); break;
    case 9: bulk_g2s(s, g, 0, bar); break;            // empty bulk copy
    }
}

int main(int argc, char** argv) {
    const int which = argc > 1 ? std::atoi(argv[1]) : 0;
    alignas(64) static char g[512];
    if (which == 10) emu::launch(probe, 1u, 32u, (size_t)(228 * 1024), which, g);   // above 227 KB per CTA
    else emu::launch(probe, 1u, 32u, (size_t)512, which, g);
    std::puts("ok");
    return 0;
}
