// nm_vmm.hpp — device buffers that grow IN PLACE (CUDA virtual memory management).
//
// The mirror of a store that keeps receiving `store_embedding` calls grows without bound
// (vector_engine/src/lib.rs:1840-1868).  A realloc-and-copy growth needs old + new buffer at the
// same time, i.e. a 100 GB mirror could never grow on a 180 GB part.  GrowBuf reserves a virtual
// address range and backs it with physical chunks (cuMemCreate + cuMemMap) as the contents grow:
// existing rows are never copied.  When the reservation itself runs out, a larger range is
// reserved and the SAME physical chunks are re-mapped into it — still no copy, no extra HBM.
// Precedent in the reference: the 16 MB-chunk EmbeddingSlab (tensor_store/src/embedding_slab.rs:27,
// 92-124).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <utility>
#include <vector>

namespace nmi {

class GrowBuf {
  public:
    GrowBuf() = default;
    GrowBuf(const GrowBuf &) = delete;
    GrowBuf &operator=(const GrowBuf &) = delete;
    ~GrowBuf() { release(); }

    // Make at least `bytes` usable at ptr().  Contents below the old size are preserved and keep
    // their physical location; ptr() may change (re-mapped into a larger reservation) — the caller
    // re-derives pointers / tensor maps and must have no work in flight on the buffer.
    // `exact` maps just what is asked for (bulk loads); otherwise growth is geometric (appends).
    // Returns an nm_status; nothing is lost on failure.
    int ensure(int device, size_t bytes, bool exact);
    void release();
    void swap(GrowBuf &o) {
        std::swap(device_, o.device_);
        std::swap(base_, o.base_);
        std::swap(va_size_, o.va_size_);
        std::swap(mapped_, o.mapped_);
        std::swap(gran_, o.gran_);
        chunks_.swap(o.chunks_);
        std::swap(remaps_, o.remaps_);
        std::swap(plain_, o.plain_);
    }

    void *ptr() const { return reinterpret_cast<void *>(base_); }
    size_t mapped() const { return mapped_; }      // usable bytes
    size_t reserved() const { return va_size_; }   // virtual bytes
    size_t chunks() const { return chunks_.size(); }
    uint64_t remaps() const { return remaps_; }    // how often the VA range was replaced
    bool vmm() const { return !plain_; }

  private:
    struct Chunk {
        CUmemGenericAllocationHandle h;
        size_t size;
    };
    int grow_plain(size_t bytes, bool exact);
    int device_ = -1;
    CUdeviceptr base_ = 0;
    size_t va_size_ = 0;
    size_t mapped_ = 0;
    size_t gran_ = 0;
    std::vector<Chunk> chunks_;
    uint64_t remaps_ = 0;
    bool plain_ = false;  // fallback: cudaMalloc + copy (driver without VMM, or NM_NO_VMM=1)
};

}  // namespace nmi
