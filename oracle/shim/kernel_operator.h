// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Stand-in for the un-vendored third-party header `kernel_operator.h` of Huawei CANN 8.0RC2
// (AscendC), which the reference's cpu run mode needs (reference cmake/cpu/CMakeLists.txt:21-24,
// problem.md:2) and which cannot be installed here.  It lets the reference's OWN sources
// (src/render.cpp, src/rt_helper.h, src/allocator.h, src/main.cpp) compile unmodified with g++ so
// that they -- not a restatement -- are the known-answer source for the parity tests.
//
// Only the symbols the reference calls are provided.  Every vector op is
// restated from the published AscendC semantics as a plain element-wise IEEE-754 binary32 loop;
// build with -ffp-contract=off and without -ffast-math so each op rounds once, as the AI-core
// vector unit (and the reference's own NumPy golden, scripts/gen_data.py:190-243) does.
#pragma once
#include <cassert>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <vector>

#define __aicore__
#define __global__
#define __gm__
#define GM_ADDR uint8_t *
#ifndef ASSERT
#define ASSERT(x) assert(x)
#endif

enum class KernelMode { MIX_MODE = 0, AIC_MODE, AIV_MODE };

namespace AscendC {

// ---- block index: one "AI core" per call of the kernel function --------------------------------
inline int64_t &shim_block_idx_ref() {
    static thread_local int64_t idx = 0;
    return idx;
}
inline int64_t GetBlockIdx() { return shim_block_idx_ref(); }

// ---- debug output: tikicpulib prints; the stand-in stays silent unless PT_REF_VERBOSE is set ----
inline bool shim_verbose() {
    static const bool v = std::getenv("PT_REF_VERBOSE") != nullptr;
    return v;
}
inline int printf(const char *fmt, ...) {
    if (!shim_verbose())
        return 0;
    va_list ap;
    va_start(ap, fmt);
    int r = std::vprintf(fmt, ap);
    va_end(ap);
    return r;
}
inline int PRINTF(const char *fmt, ...) {
    if (!shim_verbose())
        return 0;
    va_list ap;
    va_start(ap, fmt);
    int r = std::vprintf(fmt, ap);
    va_end(ap);
    return r;
}

// ---- tensors: non-owning typed views ------------------------------------------------------------
template <typename T> class LocalTensor {
  public:
    LocalTensor() : p_(nullptr) {}
    explicit LocalTensor(T *p) : p_(p) {}
    LocalTensor operator[](int64_t off) const { return LocalTensor(p_ + off); }
    T GetValue(int64_t i) const { return p_[i]; }
    template <typename V> void SetValue(int64_t i, V v) const { p_[i] = static_cast<T>(v); }
    template <typename U> LocalTensor<U> ReinterpretCast() const { return LocalTensor<U>(reinterpret_cast<U *>(p_)); }
    T *shim_ptr() const { return p_; }

  private:
    T *p_;
};

template <typename T> class GlobalTensor {
  public:
    GlobalTensor() : p_(nullptr), n_(0) {}
    void SetGlobalBuffer(T *p, uint64_t n = 0) {
        p_ = p;
        n_ = n;
    }
    GlobalTensor operator[](int64_t off) const {
        GlobalTensor g;
        g.p_ = p_ + off;
        g.n_ = n_;
        return g;
    }
    T *shim_ptr() const { return p_; }

  private:
    T *p_;
    uint64_t n_;
};

// ---- pipe / queues / buffers: plain host allocations -------------------------------------------
enum class QuePosition { GM = 0, VECIN, VECOUT, VECCALC, A1, A2, B1, B2, CO1, CO2 };
using TPosition = QuePosition;

template <QuePosition P> class TBuf {
  public:
    template <typename T> LocalTensor<T> Get() const { return LocalTensor<T>(reinterpret_cast<T *>(base_)); }
    uint8_t *base_ = nullptr;
};

template <QuePosition P, int DEPTH> class TQue {
  public:
    template <typename T> LocalTensor<T> AllocTensor() {
        uint8_t *b = base_ + static_cast<size_t>(next_ % num_) * bytes_;
        next_++;
        return LocalTensor<T>(reinterpret_cast<T *>(b));
    }
    template <typename T> void EnQue(const LocalTensor<T> &t) { fifo_.push_back(reinterpret_cast<uint8_t *>(t.shim_ptr())); }
    template <typename T> LocalTensor<T> DeQue() {
        assert(!fifo_.empty());
        uint8_t *b = fifo_.front();
        fifo_.erase(fifo_.begin());
        return LocalTensor<T>(reinterpret_cast<T *>(b));
    }
    template <typename T> void FreeTensor(const LocalTensor<T> &) {}
    uint8_t *base_ = nullptr;
    uint32_t num_ = 1, bytes_ = 0, next_ = 0;
    std::vector<uint8_t *> fifo_;
};

class TPipe {
  public:
    TPipe() {}
    ~TPipe() {
        for (void *p : owned_)
            std::free(p);
    }
    template <QuePosition P, int D> bool InitBuffer(TQue<P, D> &q, uint8_t num, uint32_t bytes) {
        q.base_ = grab(static_cast<size_t>(num) * bytes);
        q.num_ = num;
        q.bytes_ = bytes;
        return true;
    }
    template <QuePosition P> bool InitBuffer(TBuf<P> &b, uint32_t bytes) {
        b.base_ = grab(bytes);
        return true;
    }

  private:
    uint8_t *grab(size_t bytes) {
        void *p = std::aligned_alloc(256, (bytes + 255) & ~size_t(255));
        std::memset(p, 0, (bytes + 255) & ~size_t(255));
        owned_.push_back(p);
        return static_cast<uint8_t *>(p);
    }
    std::vector<void *> owned_;
};

// ---- data movement ------------------------------------------------------------------------------
template <typename T> inline void DataCopy(const LocalTensor<T> &dst, const GlobalTensor<T> &src, uint32_t n) {
    std::memcpy(dst.shim_ptr(), src.shim_ptr(), sizeof(T) * n);
}
template <typename T> inline void DataCopy(const GlobalTensor<T> &dst, const LocalTensor<T> &src, uint32_t n) {
    std::memcpy(dst.shim_ptr(), src.shim_ptr(), sizeof(T) * n);
}
template <typename T> inline void DataCopy(const LocalTensor<T> &dst, const LocalTensor<T> &src, uint32_t n) {
    std::memmove(dst.shim_ptr(), src.shim_ptr(), sizeof(T) * n);
}

// ---- element-wise vector ops (each rounds once, binary32) ---------------------------------------
template <typename T, typename S> inline void Duplicate(const LocalTensor<T> &dst, S v, int32_t n) {
    T *d = dst.shim_ptr();
    for (int32_t i = 0; i < n; i++)
        d[i] = static_cast<T>(v);
}
template <typename T> inline void Adds(const LocalTensor<T> &dst, const LocalTensor<T> &a, T s, int32_t n) {
    T *d = dst.shim_ptr();
    const T *x = a.shim_ptr();
    for (int32_t i = 0; i < n; i++)
        d[i] = x[i] + s;
}
template <typename T> inline void Muls(const LocalTensor<T> &dst, const LocalTensor<T> &a, T s, int32_t n) {
    T *d = dst.shim_ptr();
    const T *x = a.shim_ptr();
    for (int32_t i = 0; i < n; i++)
        d[i] = x[i] * s;
}
template <typename T> inline void Add(const LocalTensor<T> &dst, const LocalTensor<T> &a, const LocalTensor<T> &b, int32_t n) {
    T *d = dst.shim_ptr();
    const T *x = a.shim_ptr(), *y = b.shim_ptr();
    for (int32_t i = 0; i < n; i++)
        d[i] = x[i] + y[i];
}
template <typename T> inline void Sub(const LocalTensor<T> &dst, const LocalTensor<T> &a, const LocalTensor<T> &b, int32_t n) {
    T *d = dst.shim_ptr();
    const T *x = a.shim_ptr(), *y = b.shim_ptr();
    for (int32_t i = 0; i < n; i++)
        d[i] = x[i] - y[i];
}
template <typename T> inline void Mul(const LocalTensor<T> &dst, const LocalTensor<T> &a, const LocalTensor<T> &b, int32_t n) {
    T *d = dst.shim_ptr();
    const T *x = a.shim_ptr(), *y = b.shim_ptr();
    for (int32_t i = 0; i < n; i++)
        d[i] = x[i] * y[i];
}
template <typename T> inline void Div(const LocalTensor<T> &dst, const LocalTensor<T> &a, const LocalTensor<T> &b, int32_t n) {
    T *d = dst.shim_ptr();
    const T *x = a.shim_ptr(), *y = b.shim_ptr();
    for (int32_t i = 0; i < n; i++)
        d[i] = x[i] / y[i];
}
inline void Sqrt(const LocalTensor<float> &dst, const LocalTensor<float> &a, int32_t n) {
    float *d = dst.shim_ptr();
    const float *x = a.shim_ptr();
    for (int32_t i = 0; i < n; i++)
        d[i] = std::sqrt(x[i]);
}
template <typename T> inline void And(const LocalTensor<T> &dst, const LocalTensor<T> &a, const LocalTensor<T> &b, int32_t n) {
    T *d = dst.shim_ptr();
    const T *x = a.shim_ptr(), *y = b.shim_ptr();
    for (int32_t i = 0; i < n; i++)
        d[i] = static_cast<T>(x[i] & y[i]);
}

enum class RoundMode { CAST_NONE = 0, CAST_RINT, CAST_FLOOR, CAST_CEIL, CAST_ROUND, CAST_TRUNC, CAST_ODD };
template <typename D, typename S> inline void Cast(const LocalTensor<D> &dst, const LocalTensor<S> &src, RoundMode, int32_t n) {
    D *d = dst.shim_ptr();
    const S *x = src.shim_ptr();
    for (int32_t i = 0; i < n; i++)
        d[i] = static_cast<D>(x[i]);
}

// bit j of mask byte i  <->  element 8*i + j  (little-endian bit order, cf. reference rt_helper.h:165-180)
enum class SELMODE { VSEL_CMPMASK_SPR = 0, VSEL_TENSOR_SCALAR_MODE, VSEL_TENSOR_TENSOR_MODE };
template <typename T>
inline void Select(const LocalTensor<T> &dst, const LocalTensor<uint8_t> &mask, const LocalTensor<T> &src, T scalar, SELMODE, int32_t n) {
    T *d = dst.shim_ptr();
    const T *x = src.shim_ptr();
    const uint8_t *m = mask.shim_ptr();
    for (int32_t i = 0; i < n; i++)
        d[i] = ((m[i >> 3] >> (i & 7)) & 1) ? x[i] : scalar;
}

enum class CMPMODE { LT = 0, GT, EQ, LE, GE, NE };
template <typename T>
inline void Compare(const LocalTensor<uint8_t> &dst, const LocalTensor<T> &a, const LocalTensor<T> &b, CMPMODE mode, int32_t n) {
    uint8_t *d = dst.shim_ptr();
    const T *x = a.shim_ptr(), *y = b.shim_ptr();
    for (int32_t i = 0; i < n; i += 8) {
        uint8_t byte = 0;
        for (int32_t j = 0; j < 8 && i + j < n; j++) {
            bool r = false;
            switch (mode) {
            case CMPMODE::LT: r = x[i + j] < y[i + j]; break;
            case CMPMODE::GT: r = x[i + j] > y[i + j]; break;
            case CMPMODE::EQ: r = x[i + j] == y[i + j]; break;
            case CMPMODE::LE: r = x[i + j] <= y[i + j]; break;
            case CMPMODE::GE: r = x[i + j] >= y[i + j]; break;
            case CMPMODE::NE: r = x[i + j] != y[i + j]; break;
            }
            byte = static_cast<uint8_t>(byte | (r ? (1u << j) : 0u));
        }
        d[i >> 3] = byte;
    }
}

// One repeat covers 8 datablocks of 32 bytes; each datablock reduces to one value.
template <typename T>
inline void BlockReduceMin(const LocalTensor<T> &dst, const LocalTensor<T> &src, int32_t repeat, int32_t mask, int32_t dstRepStride,
                           int32_t srcBlkStride, int32_t srcRepStride) {
    constexpr int32_t perBlk = 32 / sizeof(T);
    T *d = dst.shim_ptr();
    const T *x = src.shim_ptr();
    for (int32_t r = 0; r < repeat; r++) {
        for (int32_t b = 0; b < 8; b++) {
            if (b * perBlk >= mask)
                break;
            const T *blk = x + (static_cast<int64_t>(r) * srcRepStride + static_cast<int64_t>(b) * srcBlkStride) * perBlk;
            T m = blk[0];
            for (int32_t e = 1; e < perBlk; e++)
                m = (blk[e] < m) ? blk[e] : m;
            d[static_cast<int64_t>(r) * dstRepStride * 8 + b] = m;
        }
    }
}

struct BrcbRepeatParams {
    uint16_t dstBlkStride;
    uint16_t dstRepStride;
};
// Each repeat reads 8 scalars and broadcasts each across one 32-byte datablock.
template <typename T> inline void Brcb(const LocalTensor<T> &dst, const LocalTensor<T> &src, uint8_t repeat, BrcbRepeatParams p) {
    constexpr int32_t perBlk = 32 / sizeof(T);
    T *d = dst.shim_ptr();
    const T *x = src.shim_ptr();
    for (int32_t r = 0; r < repeat; r++)
        for (int32_t j = 0; j < 8; j++) {
            T *blk = d + (static_cast<int64_t>(r) * p.dstRepStride + static_cast<int64_t>(j) * p.dstBlkStride) * perBlk;
            for (int32_t e = 0; e < perBlk; e++)
                blk[e] = x[r * 8 + j];
        }
}

template <typename T> inline void DumpTensor(const LocalTensor<T> &t, uint32_t desc, uint32_t n) {
    if (!shim_verbose())
        return;
    std::printf("DumpTensor desc=%u:", desc);
    for (uint32_t i = 0; i < n; i++)
        std::printf(" %g", static_cast<double>(t.GetValue(i)));
    std::printf("\n");
}

} // namespace AscendC

inline int64_t get_block_idx() { return AscendC::GetBlockIdx(); }
