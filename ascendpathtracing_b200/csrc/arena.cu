// Device arena: the B200 counterpart of the reference's on-chip MemoryPool (src/allocator.h).
//
// The reference carves a 16 KB unified-buffer tensor into per-op scratch with a first-fit free list that
// splits on Alloc and coalesces neighbours on Free (allocator.h:68-151,204-241), behind an RAII handle
// (AllocDecorator, :253-289).  On a GPU those per-op temporaries are registers, so what remains worth
// pooling is the large device buffers (ray tiles, colour tiles, framebuffers, per-GPU stripes): one
// cudaMalloc up front, then the same Init/Alloc/Free discipline at 256-byte granularity, O(#blocks)
// first-fit, neighbour coalescing, and the reference's debug checks (foreign pointer / double free are
// reported instead of corrupting the list).  The C++ RAII handle lives in host/pt_arena.hpp.
#include <map>

#include "pt_host.h"

struct PtArena {
    char *base = nullptr;
    size_t capacity = 0;
    int device = 0;
    bool owns = true;
    std::mutex mu;
    // offset -> (size, free?) ; blocks tile [0, capacity) exactly, like the reference's linked list
    struct Block {
        size_t size;
        bool is_free;
    };
    std::map<size_t, Block> blocks;
    size_t in_use = 0;
};

namespace {
constexpr size_t kGranule = 256;
size_t round_up(size_t v) { return (v + kGranule - 1) / kGranule * kGranule; }
}  // namespace

extern "C" {

int ptb200_arena_create(size_t bytes, PtArena **out) {
    if (out == nullptr || bytes == 0)
        return ptb200::fail(PTB200_EINVAL, "ptb200_arena_create: null out pointer or zero size");
    *out = nullptr;
    PtArena *a = new PtArena();
    a->capacity = round_up(bytes);
    cudaError_t e = cudaGetDevice(&a->device);
    if (e == cudaSuccess)
        e = cudaMalloc(reinterpret_cast<void **>(&a->base), a->capacity);
    if (e != cudaSuccess) {
        delete a;
        if (e == cudaErrorMemoryAllocation) {
            cudaGetLastError();
            return ptb200::fail(PTB200_ENOMEM, "ptb200_arena_create: cudaMalloc of %zu bytes failed", bytes);
        }
        return ptb200::fail_cuda(e, "ptb200_arena_create");
    }
    a->blocks[0] = {a->capacity, true};  // Allocator::Init: one free node spanning everything
    *out = a;
    return PTB200_OK;
}

int ptb200_arena_wrap(void *base, size_t bytes, PtArena **out) {
    if (out == nullptr || base == nullptr || bytes < kGranule)
        return ptb200::fail(PTB200_EINVAL, "ptb200_arena_wrap: null pointer or region smaller than %zu bytes", kGranule);
    PtArena *a = new PtArena();
    a->base = static_cast<char *>(base);
    a->capacity = bytes / kGranule * kGranule;
    a->owns = false;
    a->blocks[0] = {a->capacity, true};
    *out = a;
    return PTB200_OK;
}

int ptb200_arena_destroy(PtArena *a) {
    if (a == nullptr)
        return PTB200_OK;
    cudaError_t e = a->owns ? cudaFree(a->base) : cudaSuccess;
    delete a;
    return e == cudaSuccess ? PTB200_OK : ptb200::fail_cuda(e, "ptb200_arena_destroy");
}

void *ptb200_arena_alloc(PtArena *a, size_t bytes) {
    if (a == nullptr || bytes == 0) {
        ptb200::fail(PTB200_EINVAL, "ptb200_arena_alloc: null arena or zero size");
        return nullptr;
    }
    const size_t need = round_up(bytes);
    std::lock_guard<std::mutex> lock(a->mu);
    for (auto it = a->blocks.begin(); it != a->blocks.end(); ++it) {  // first fit, allocator.h:69-70
        if (!it->second.is_free || it->second.size < need)
            continue;
        const size_t off = it->first, rest = it->second.size - need;
        it->second = {need, false};
        if (rest > 0)
            a->blocks[off + need] = {rest, true};  // split, allocator.h:76-83
        a->in_use += need;
        return a->base + off;
    }
    ptb200::fail(PTB200_ENOMEM, "ptb200_arena_alloc: no free block of %zu bytes (capacity %zu, in use %zu)", need, a->capacity, a->in_use);
    return nullptr;
}

int ptb200_arena_free(PtArena *a, void *ptr) {
    if (a == nullptr)
        return ptb200::fail(PTB200_EINVAL, "ptb200_arena_free: null arena");
    if (ptr == nullptr)
        return PTB200_OK;
    std::lock_guard<std::mutex> lock(a->mu);
    const char *p = static_cast<const char *>(ptr);
    if (p < a->base || p >= a->base + a->capacity)
        return ptb200::fail(PTB200_EINVAL, "ptb200_arena_free: pointer not owned by this arena");
    auto it = a->blocks.find(static_cast<size_t>(p - a->base));
    if (it == a->blocks.end())
        return ptb200::fail(PTB200_EINVAL, "ptb200_arena_free: pointer is not the start of an allocation");
    if (it->second.is_free)
        return ptb200::fail(PTB200_EINVAL, "ptb200_arena_free: double free");  // allocator.h:262-266
    it->second.is_free = true;
    a->in_use -= it->second.size;
    // coalesce with the next, then the previous neighbour (MergeNode, allocator.h:204-241)
    auto next = std::next(it);
    if (next != a->blocks.end() && next->second.is_free) {
        it->second.size += next->second.size;
        a->blocks.erase(next);
    }
    if (it != a->blocks.begin()) {
        auto prev = std::prev(it);
        if (prev->second.is_free) {
            prev->second.size += it->second.size;
            a->blocks.erase(it);
        }
    }
    return PTB200_OK;
}

size_t ptb200_arena_capacity(const PtArena *a) { return a ? a->capacity : 0; }

size_t ptb200_arena_in_use(const PtArena *a) {
    if (a == nullptr)
        return 0;
    std::lock_guard<std::mutex> lock(const_cast<PtArena *>(a)->mu);
    return a->in_use;
}

size_t ptb200_arena_largest_free(const PtArena *a) {
    if (a == nullptr)
        return 0;
    std::lock_guard<std::mutex> lock(const_cast<PtArena *>(a)->mu);
    size_t best = 0;
    for (const auto &kv : a->blocks)
        if (kv.second.is_free && kv.second.size > best)
            best = kv.second.size;
    return best;
}

}  // extern "C"
