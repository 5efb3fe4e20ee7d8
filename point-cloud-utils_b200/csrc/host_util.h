// host_util.h -- small host-side helpers shared by the launch-sequencing code.
#pragma once
#include <cstddef>

namespace pcu {

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

// Walks a scratch arena handing out aligned sub-buffers.  With a null base it only measures.
struct Carver {
    unsigned char* base;
    size_t off = 0;
    explicit Carver(unsigned char* b) : base(b) {}
    template <typename U> U* take(size_t count) {
        U* p = base ? reinterpret_cast<U*>(base + off) : nullptr;
        off += align_up(count * sizeof(U));
        return p;
    }
};

}  // namespace pcu
