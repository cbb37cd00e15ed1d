// nm_comm.cu — row-range sharding across processes: lazily loaded NCCL, the communicator, and
// the CUDA-IPC mailboxes of the fused peer-memory exchange (see exchange_and_merge in
// scan_kernels.cuh).
#include "nm_internal.hpp"

#include <dlfcn.h>

#include <cstdlib>

namespace nmi {

NcclApi &nccl() {
    static NcclApi api = [] {
        NcclApi a;
        a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!a.handle) a.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!a.handle) return a;
        a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.handle, "ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.handle, "ncclCommInitRank");
        a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.handle, "ncclCommDestroy");
        a.AllGather = (decltype(a.AllGather))dlsym(a.handle, "ncclAllGather");
        a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.handle, "ncclGetErrorString");
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather &&
               a.GetErrorString;
        return a;
    }();
    return api;
}

constexpr size_t kXchgFlagBytes = 256;
size_t xchg_bytes(int n_ranks) {
    return kXchgFlagBytes + (size_t)2 * n_ranks * nm::kMaxFastK * sizeof(nm::ShardHit);
}

nm::PeerXchg make_xchg(const nm_index *idx, uint32_t seq) {
    nm::PeerXchg x;
    memset(&x, 0, sizeof(x));
    x.n_ranks = (uint32_t)idx->n_ranks;
    x.rank = (uint32_t)idx->rank;
    x.seq = seq;
    x.kcap = nm::kMaxFastK;
    for (int r = 0; r < idx->n_ranks; ++r) {
        uint8_t *base = static_cast<uint8_t *>(idx->xchg_peer[r]);
        x.flags[r] = reinterpret_cast<uint32_t *>(base);
        x.mailbox[r] = reinterpret_cast<nm::ShardHit *>(base + kXchgFlagBytes);
    }
    return x;
}

// Map every rank's exchange buffer into this process.  Failure is not fatal: the index then
// keeps using ncclAllGather + merge_shards_kernel.
void setup_peer_exchange(nm_index *idx, cudaStream_t stream) {
    idx->xchg_ok = false;
    const char *off = getenv("NM_DISABLE_PEER_EXCHANGE");
    if (off && off[0] == '1') return;
    if (idx->n_ranks < 2 || idx->n_ranks > nm::kMaxRanks) return;
    const int n = idx->n_ranks;
    bool ok = true;
    cudaIpcMemHandle_t mine;
    cudaIpcMemHandle_t *d_handles = nullptr;
    std::vector<cudaIpcMemHandle_t> all((size_t)n);
    // every rank must reach the all-gather below, so failures only clear `ok`
    if (cudaMalloc(&idx->xchg_mem, xchg_bytes(n)) != cudaSuccess) ok = false;
    if (ok && cudaMemset(idx->xchg_mem, 0, xchg_bytes(n)) != cudaSuccess) ok = false;
    memset(&mine, 0, sizeof(mine));
    if (ok && cudaIpcGetMemHandle(&mine, idx->xchg_mem) != cudaSuccess) ok = false;
    if (!ok) memset(&mine, 0, sizeof(mine));
    if (cudaMalloc(&d_handles, sizeof(mine) * n) != cudaSuccess) {
        cudaGetLastError();
        return;  // cannot even exchange: peers time out in NCCL, nothing we can do here
    }
    cudaMemcpyAsync(d_handles + idx->rank, &mine, sizeof(mine), cudaMemcpyHostToDevice, stream);
    ncclResult_t nr = nccl().AllGather(d_handles + idx->rank, d_handles, sizeof(mine), ncclChar,
                                       idx->comm, stream);
    cudaMemcpyAsync(all.data(), d_handles, sizeof(mine) * n, cudaMemcpyDeviceToHost, stream);
    if (cudaStreamSynchronize(stream) != cudaSuccess || nr != ncclSuccess) ok = false;
    cudaFree(d_handles);
    cudaIpcMemHandle_t zero;
    memset(&zero, 0, sizeof(zero));
    for (int r = 0; r < n && ok; ++r)
        if (memcmp(&all[r], &zero, sizeof(zero)) == 0) ok = false;  // some rank failed
    for (int r = 0; r < n && ok; ++r) {
        if (r == idx->rank) {
            idx->xchg_peer[r] = idx->xchg_mem;
        } else if (cudaIpcOpenMemHandle(&idx->xchg_peer[r], all[r],
                                        cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            idx->xchg_peer[r] = nullptr;
            ok = false;
        }
    }
    // agree on the outcome: everyone uses the exchange or nobody does
    int *d_flag = nullptr;
    std::vector<int> flags((size_t)n, 0);
    int my = ok ? 1 : 0;
    if (cudaMalloc(&d_flag, sizeof(int) * n) == cudaSuccess) {
        cudaMemcpyAsync(d_flag + idx->rank, &my, sizeof(int), cudaMemcpyHostToDevice, stream);
        nr = nccl().AllGather(d_flag + idx->rank, d_flag, sizeof(int), ncclChar, idx->comm, stream);
        cudaMemcpyAsync(flags.data(), d_flag, sizeof(int) * n, cudaMemcpyDeviceToHost, stream);
        if (cudaStreamSynchronize(stream) != cudaSuccess || nr != ncclSuccess) ok = false;
        cudaFree(d_flag);
        for (int r = 0; r < n; ++r) ok = ok && flags[r] == 1;
    } else {
        ok = false;
    }
    cudaGetLastError();
    idx->xchg_ok = ok;
    idx->xchg_seq = 0;
}

void teardown_peer_exchange(nm_index *idx) {
    for (int r = 0; r < nm::kMaxRanks; ++r) {
        if (idx->xchg_peer[r] && idx->xchg_peer[r] != idx->xchg_mem)
            cudaIpcCloseMemHandle(idx->xchg_peer[r]);
        idx->xchg_peer[r] = nullptr;
    }
    idx->xchg_ok = false;
}

}  // namespace nmi

using namespace nmi;

extern "C" {

int nm_comm_create_id(void *out_id) {
    if (!out_id) return fail(NM_ERR_INVALID_ARGUMENT, "null id buffer");
    if (!nccl().ok) return fail(NM_ERR_STORAGE, "libnccl.so.2 could not be loaded");
    static_assert(sizeof(ncclUniqueId) == NM_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    NCCL_TRY(nccl().GetUniqueId(&id));
    memcpy(out_id, &id, sizeof(id));
    return NM_OK;
}

int nm_index_attach_comm(nm_index *idx, const void *id, int n_ranks, int rank, uint64_t row_base) {
    if (!idx || !id) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks)
        return fail(NM_ERR_INVALID_ARGUMENT, "bad rank %d of %d", rank, n_ranks);
    if (!nccl().ok) return fail(NM_ERR_STORAGE, "libnccl.so.2 could not be loaded");
    std::unique_lock<std::shared_mutex> g(idx->mu);
    if (idx->shards.size() != 1)
        return fail(NM_ERR_CONFIGURATION, "a communicator needs a single-device index per rank");
    if (idx->comm) return fail(NM_ERR_CONFIGURATION, "communicator already attached");
    CUDA_TRY(cudaSetDevice(idx->shards[0]->device));
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    NCCL_TRY(nccl().CommInitRank(&idx->comm, n_ranks, uid, rank));
    idx->n_ranks = n_ranks;
    idx->rank = rank;
    idx->comm_row_base = row_base;
    setup_peer_exchange(idx, idx->shards[0]->copy_stream);
    return NM_OK;
}

int nm_index_detach_comm(nm_index *idx) {
    if (!idx) return fail(NM_ERR_INVALID_ARGUMENT, "null index");
    std::unique_lock<std::shared_mutex> g(idx->mu);
    if (idx->comm) {
        CUDA_TRY(cudaSetDevice(idx->shards[0]->device));
        CUDA_TRY(cudaDeviceSynchronize());
        teardown_peer_exchange(idx);
        NCCL_TRY(nccl().CommDestroy(idx->comm));  // collective: every rank has unmapped by now
        idx->comm = nullptr;
        if (idx->xchg_mem) cudaFree(idx->xchg_mem);
        idx->xchg_mem = nullptr;
    }
    idx->n_ranks = 1;
    idx->rank = 0;
    idx->comm_row_base = 0;
    return NM_OK;
}

}  // extern "C"
