// C++ face of the device arena (include/ptb200.h: ptb200_arena_*): the roles of the reference's Allocator
// and its RAII handle AllocDecorator (src/allocator.h:54-66,253-289) for device buffers.
#pragma once
#include <cstddef>
#include <stdexcept>
#include <utility>

#include "../../include/ptb200.h"

namespace ptb200 {

class DeviceArena {
  public:
    explicit DeviceArena(size_t bytes) {  // Allocator::Init
        if (ptb200_arena_create(bytes, &a_) != PTB200_OK)
            throw std::runtime_error(ptb200_last_error());
    }
    ~DeviceArena() { ptb200_arena_destroy(a_); }
    DeviceArena(const DeviceArena &) = delete;
    DeviceArena &operator=(const DeviceArena &) = delete;

    // RAII block: frees itself on scope exit like AllocDecorator; Release() frees early, Get() refuses a freed block.
    class Block {
      public:
        Block(PtArena *a, void *p, size_t n) : a_(a), p_(p), n_(n) {}
        Block(Block &&o) noexcept : a_(o.a_), p_(std::exchange(o.p_, nullptr)), n_(o.n_) {}
        Block(const Block &) = delete;
        Block &operator=(const Block &) = delete;
        ~Block() {
            if (p_ != nullptr)
                ptb200_arena_free(a_, p_);
        }
        template <typename T = uint8_t> T *Get() const {
            if (p_ == nullptr)
                throw std::logic_error("try to access a free memory");  // allocator.h:267-269
            return static_cast<T *>(p_);
        }
        size_t size() const { return n_; }
        void Release() {
            if (p_ == nullptr)
                throw std::logic_error("double free manually");  // allocator.h:280-282
            ptb200_arena_free(a_, p_);
            p_ = nullptr;
        }

      private:
        PtArena *a_;
        void *p_;
        size_t n_;
    };

    Block Alloc(size_t bytes) {
        void *p = ptb200_arena_alloc(a_, bytes);
        if (p == nullptr)
            throw std::runtime_error("no enough memory");  // allocator.h:103-105
        return Block(a_, p, bytes);
    }
    size_t capacity() const { return ptb200_arena_capacity(a_); }
    size_t in_use() const { return ptb200_arena_in_use(a_); }

  private:
    PtArena *a_ = nullptr;
};

}  // namespace ptb200
