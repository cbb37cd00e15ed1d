// nm_vmm.cu — GrowBuf: device buffers that grow in place (see nm_vmm.hpp).
#include "nm_vmm.hpp"

#include "nm_internal.hpp"

#include <cstdlib>

namespace nmi {
namespace {

// Driver entry points through the runtime (no link-time libcuda dependency, like the tensor-map
// encoder in nm_core.cu).
struct VmmApi {
    CUresult (*GetGranularity)(size_t *, const CUmemAllocationProp *,
                               CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*AddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*AddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*Create)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *,
                       unsigned long long) = nullptr;
    CUresult (*Release)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*Map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*Unmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*SetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
    bool ok = false;
};

template <typename F>
bool entry(const char *name, F *out) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess || !p) {
        cudaGetLastError();
        return false;
    }
    *out = reinterpret_cast<F>(p);
    return true;
}

const VmmApi &vmm_api() {
    static const VmmApi api = [] {
        VmmApi a;
        const char *off = getenv("NM_NO_VMM");
        if (off && off[0] == '1') return a;
        a.ok = entry("cuMemGetAllocationGranularity", &a.GetGranularity) &&
               entry("cuMemAddressReserve", &a.AddressReserve) &&
               entry("cuMemAddressFree", &a.AddressFree) && entry("cuMemCreate", &a.Create) &&
               entry("cuMemRelease", &a.Release) && entry("cuMemMap", &a.Map) &&
               entry("cuMemUnmap", &a.Unmap) && entry("cuMemSetAccess", &a.SetAccess);
        return a;
    }();
    return api;
}

constexpr size_t kMaxPiece = 1ull << 30;      // physical allocations of at most 1 GiB each
constexpr size_t kMinReserve = 256ull << 20;  // virtual: costs nothing until it is mapped
constexpr size_t kMaxSlack = 4ull << 30;      // geometric growth never maps more than this ahead

size_t round_up(size_t v, size_t g) { return (v + g - 1) / g * g; }

CUmemAllocationProp device_prop(int device) {
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof(prop));
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    return prop;
}

}  // namespace

int GrowBuf::grow_plain(size_t bytes, bool exact) {
    size_t cap = exact ? bytes : std::max(bytes, std::min(mapped_ + mapped_ / 2, bytes + kMaxSlack));
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, cap);
    if (e != cudaSuccess && cap > bytes) {
        cudaGetLastError();
        cap = bytes;
        e = cudaMalloc(&p, cap);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(NM_ERR_STORAGE, "out of device memory growing a buffer to %zu bytes: %s", cap,
                    cudaGetErrorString(e));
    }
    if (base_ && mapped_) CUDA_TRY(cudaMemcpy(p, ptr(), mapped_, cudaMemcpyDeviceToDevice));
    if (base_) CUDA_TRY(cudaFree(ptr()));
    base_ = reinterpret_cast<CUdeviceptr>(p);
    mapped_ = va_size_ = cap;
    return NM_OK;
}

int GrowBuf::ensure(int device, size_t bytes, bool exact) {
    if (bytes <= mapped_) return NM_OK;
    if (device_ >= 0 && device_ != device)
        return fail(NM_ERR_INVALID_ARGUMENT, "GrowBuf belongs to device %d, not %d", device_, device);
    const VmmApi &api = vmm_api();
    if (device_ < 0) {
        device_ = device;
        plain_ = !api.ok;
        if (!plain_) {
            CUDA_TRY(cudaFree(nullptr));  // primary context current on this thread
            CUmemAllocationProp prop = device_prop(device);
            if (api.GetGranularity(&gran_, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS ||
                gran_ == 0)
                plain_ = true;
        }
    }
    if (plain_) return grow_plain(bytes, exact);

    const size_t need = round_up(bytes, gran_);
    size_t target = need;
    if (!exact) target = std::max(need, std::min(round_up(mapped_ + mapped_ / 2, gran_), need + kMaxSlack));

    // ---- virtual range: reserve a larger one and re-map the existing chunks (no copy) ----
    if (target > va_size_) {
        const size_t want = round_up(std::max(target * 2, kMinReserve), gran_);
        CUdeviceptr nb = 0;
        CUresult r = api.AddressReserve(&nb, want, 0, 0, 0);
        if (r != CUDA_SUCCESS)
            return fail(NM_ERR_STORAGE, "cuMemAddressReserve(%zu) failed with CUresult %d", want, (int)r);
        if (base_) {
            // Map every chunk into the new range FIRST (one physical allocation may be mapped at
            // several addresses), switch over only when all of it worked: a failure half way
            // leaves the old mapping — and the mirror — untouched.
            size_t off = 0, done = 0;
            for (const Chunk &c : chunks_) {
                r = api.Map(nb + off, c.size, 0, c.h, 0);
                if (r != CUDA_SUCCESS) break;
                off += c.size;
                ++done;
            }
            if (r == CUDA_SUCCESS && mapped_) {
                CUmemAccessDesc acc;
                acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
                acc.location.id = device_;
                acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
                r = api.SetAccess(nb, mapped_, &acc, 1);
            }
            if (r != CUDA_SUCCESS) {
                off = 0;
                for (size_t i = 0; i < done; ++i) {
                    api.Unmap(nb + off, chunks_[i].size);
                    off += chunks_[i].size;
                }
                api.AddressFree(nb, want);
                return fail(NM_ERR_STORAGE, "re-mapping the mirror into a larger range failed with CUresult %d",
                            (int)r);
            }
            off = 0;
            for (const Chunk &c : chunks_) {
                api.Unmap(base_ + off, c.size);
                off += c.size;
            }
            api.AddressFree(base_, va_size_);
            ++remaps_;
        }
        base_ = nb;
        va_size_ = want;
    }

    // ---- physical chunks for [mapped_, target) ----
    CUmemAllocationProp prop = device_prop(device_);
    while (mapped_ < target) {
        const size_t piece = std::min(target - mapped_, kMaxPiece);
        CUmemGenericAllocationHandle h;
        CUresult r = api.Create(&h, piece, &prop, 0);
        if (r != CUDA_SUCCESS) {
            if (mapped_ >= need) break;  // the geometric slack did not fit: what is needed is there
            if (target > need) {         // retry without the slack
                target = need;
                continue;
            }
            return fail(NM_ERR_STORAGE,
                        "out of device memory growing the mirror to %zu bytes (cuMemCreate: CUresult %d)",
                        need, (int)r);
        }
        r = api.Map(base_ + mapped_, piece, 0, h, 0);
        if (r == CUDA_SUCCESS) {
            CUmemAccessDesc acc;
            acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
            acc.location.id = device_;
            acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
            r = api.SetAccess(base_ + mapped_, piece, &acc, 1);
            if (r != CUDA_SUCCESS) api.Unmap(base_ + mapped_, piece);
        }
        if (r != CUDA_SUCCESS) {
            api.Release(h);
            return fail(NM_ERR_STORAGE, "mapping a mirror chunk failed with CUresult %d", (int)r);
        }
        chunks_.push_back(Chunk{h, piece});
        mapped_ += piece;
    }
    return NM_OK;
}

void GrowBuf::release() {
    if (!base_) return;
    if (device_ >= 0) cudaSetDevice(device_);
    if (plain_) {
        cudaFree(ptr());
    } else {
        const VmmApi &api = vmm_api();
        size_t off = 0;
        for (const Chunk &c : chunks_) {
            api.Unmap(base_ + off, c.size);
            api.Release(c.h);
            off += c.size;
        }
        api.AddressFree(base_, va_size_);
    }
    chunks_.clear();
    base_ = 0;
    va_size_ = mapped_ = 0;
}

}  // namespace nmi
