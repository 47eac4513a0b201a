// TEST INFRASTRUCTURE ONLY -- stand-in for CANN's `tikicpulib.h` (the cpu-mode kernel runner the
// reference's src/main.cpp:12,27-44 uses).  See kernel_operator.h in this directory for the why.
//
// ICPU_RUN_KF(fn, blockDim, args...) runs the kernel function once per block index, as the real
// library does (it forks one process per block); here the blocks run on std::threads, PT_REF_THREADS
// at a time (default 1 = sequential).  Blocks write disjoint slices, so the result does not depend
// on the thread count.
#pragma once
#include "kernel_operator.h"
#include <algorithm>
#include <atomic>
#include <thread>

namespace AscendC {
inline void *GmAlloc(size_t bytes) {
    void *p = std::aligned_alloc(4096, (bytes + 4095) & ~size_t(4095));
    if (p)
        std::memset(p, 0, bytes);
    return p;
}
inline void GmFree(void *p) { std::free(p); }
inline void SetKernelMode(KernelMode) {}

template <typename F, typename... Args> inline void shim_run_blocks(F fn, uint32_t blockDim, Args... args) {
    const char *e = std::getenv("PT_REF_THREADS");
    uint32_t nthr = e ? static_cast<uint32_t>(std::max(1, std::atoi(e))) : 1u;
    nthr = std::min(nthr, blockDim);
    if (nthr <= 1) {
        for (uint32_t b = 0; b < blockDim; b++) {
            shim_block_idx_ref() = b;
            fn(args...);
        }
        shim_block_idx_ref() = 0;
        return;
    }
    std::atomic<uint32_t> next{0};
    std::vector<std::thread> pool;
    for (uint32_t t = 0; t < nthr; t++)
        pool.emplace_back([&]() {
            for (uint32_t b = next.fetch_add(1); b < blockDim; b = next.fetch_add(1)) {
                shim_block_idx_ref() = b;
                fn(args...);
            }
        });
    for (auto &t : pool)
        t.join();
}
} // namespace AscendC

#define ICPU_RUN_KF(func, blockDim, ...) AscendC::shim_run_blocks(func, blockDim, __VA_ARGS__)
