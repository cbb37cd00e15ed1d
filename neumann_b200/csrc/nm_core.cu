// nm_core.cu — the device mirror behind include/neumann_b200.h: lifecycle, pinned
// double-buffered staging, mutations (append / update / swap-remove), upkeep of the optional
// int8 copy, statistics.  Row-major f32, pitch = dim rounded up to 4 floats so every row is
// 16-byte aligned for TMA.  There is deliberately no CPU code path for the scan.
#include "nm_internal.hpp"
#include "nm_trace.hpp"

#include <cstdarg>
#include <cstdio>
#include <thread>

using namespace nmi;

namespace nmi {

static thread_local std::string g_last_error;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}
const char *last_error() { return g_last_error.c_str(); }

}  // namespace nmi

namespace {

// ---- cuTensorMapEncodeTiled through the runtime (no link-time libcuda dependency) --------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) !=
                cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

constexpr size_t kStagingBytes = 64u << 20;  // per pinned staging buffer (two per shard)

// Pageable -> pinned staging copy, split over a few host threads (one core tops out near
// 11 GB/s, well below what the DMA engine takes from pinned memory).
void parallel_memcpy(void *dst, const void *src, size_t bytes) {
    unsigned hw = std::thread::hardware_concurrency();
    unsigned n = std::min<unsigned>(8u, hw ? hw : 1u);
    if (bytes < (4u << 20) || n < 2) {
        memcpy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> th;
    const size_t per = ((bytes / n) + 4095) & ~size_t(4095);
    for (unsigned i = 1; i < n; ++i) {
        size_t off = per * i;
        if (off >= bytes) break;
        size_t len = std::min(per, bytes - off);
        th.emplace_back([=] { memcpy((char *)dst + off, (const char *)src + off, len); });
    }
    memcpy(dst, src, std::min(per, bytes));
    for (auto &t : th) t.join();
}

}  // namespace

namespace nmi {

int build_tmap(nm_index *idx, Shard &sh) {
    sh.tmap_valid = false;
    if (sh.rows == 0) return NM_OK;
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return fail(NM_ERR_STORAGE, "cuTensorMapEncodeTiled unavailable (driver too old?)");
    cuuint64_t gdim[2] = {idx->dim, sh.rows};
    cuuint64_t gstride[1] = {(cuuint64_t)idx->pitch * 4};
    cuuint32_t box[2] = {nm::kChunkFloats, nm::kRowsPerBlock};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&sh.tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, sh.d_rows, gdim, gstride, box,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(NM_ERR_STORAGE, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    sh.tmap_valid = true;
    return NM_OK;
}


// ---- int8 pre-filter copy upkeep ---------------------------------------------------------
uint32_t q8_pitch(uint32_t dim) { return (dim + 15u) & ~15u; }

int build_tmap8(nm_index *idx, Shard &sh) {
    sh.tmap8_valid = false;
    if (sh.rows == 0 || !sh.d_q8) return NM_OK;
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return fail(NM_ERR_STORAGE, "cuTensorMapEncodeTiled unavailable (driver too old?)");
    cuuint64_t gdim[2] = {idx->dim, sh.rows};
    cuuint64_t gstride[1] = {(cuuint64_t)q8_pitch(idx->dim)};
    cuuint32_t box[2] = {128, nm::kRowsPerBlock};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&sh.tmap8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, sh.d_q8, gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(NM_ERR_STORAGE, "cuTensorMapEncodeTiled (int8) failed with CUresult %d", (int)r);
    // the tensor-core batch pre-filter reads the same copy in [128 rows x 128 B] MMA tiles
    int rc = encode_tmap_u8(&sh.tmap8_tc, sh.d_q8, idx->dim, sh.rows, q8_pitch(idx->dim), 128,
                            nm::kTcTileRows);
    if (rc) return rc;
    sh.tmap8_valid = true;
    return NM_OK;
}

// 2-D uint8 tensor map [rows, inner] (row pitch in bytes), SWIZZLE_128B, zero fill out of bounds.
int encode_tmap_u8(CUtensorMap *out, void *base, uint64_t inner, uint64_t rows, uint64_t pitch_bytes,
                   uint32_t box_inner, uint32_t box_rows) {
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return fail(NM_ERR_STORAGE, "cuTensorMapEncodeTiled unavailable (driver too old?)");
    cuuint64_t gdim[2] = {inner, rows};
    cuuint64_t gstride[1] = {pitch_bytes};
    cuuint32_t box[2] = {box_inner, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(NM_ERR_STORAGE, "cuTensorMapEncodeTiled (u8) failed with CUresult %d", (int)r);
    return NM_OK;
}

// Bring the int8 copy of rows [first, first+n) up to date (call with the index write lock held,
// after the f32 mirror holds the new data).  No-op while the pre-filter is off.
int q8_refresh(nm_index *idx, Shard &sh, uint64_t first, uint64_t n, bool build) {
    const int mode = idx->prefilter.load();
    if (mode == 0) return NM_OK;
    NM_TRACE("q8_refresh");
    if (mode == 0) return NM_OK;
    if (mode == 2 && !build && !sh.d_q8) return NM_OK;  // auto: nothing to keep up to date yet
    const uint32_t pitch8 = q8_pitch(idx->dim);
    if (sh.rows > sh.q8_capacity) {
        // grows in place like the f32 mirror: rows already quantised keep their bytes
        const bool exact = sh.q8_rows == 0;
        int rc = sh.q8_buf.ensure(sh.device, sh.rows * (size_t)pitch8, exact);
        if (!rc) rc = sh.meta_buf.ensure(sh.device, sh.rows * sizeof(nm::RowMeta), exact);
        if (!rc) rc = sh.norms_buf.ensure(sh.device, sh.rows * sizeof(float2), exact);
        if (rc) {
            sh.q8_buf.release();
            sh.meta_buf.release();
            sh.norms_buf.release();
            sh.d_q8 = nullptr;
            sh.d_meta = nullptr;
            sh.d_norms = nullptr;
            sh.q8_capacity = sh.q8_rows = 0;
            sh.tmap8_valid = false;
            return rc;
        }
        sh.d_q8 = static_cast<int8_t *>(sh.q8_buf.ptr());
        sh.d_meta = static_cast<nm::RowMeta *>(sh.meta_buf.ptr());
        sh.d_norms = static_cast<float2 *>(sh.norms_buf.ptr());
        sh.q8_capacity = std::min<uint64_t>({sh.q8_buf.mapped() / pitch8,
                                             sh.meta_buf.mapped() / sizeof(nm::RowMeta),
                                             sh.norms_buf.mapped() / sizeof(float2)});
    }
    if (!sh.d_q8_flag) {
        CUDA_TRY(cudaMalloc(&sh.d_q8_flag, sizeof(uint32_t)));
        CUDA_TRY(cudaMemsetAsync(sh.d_q8_flag, 0, sizeof(uint32_t), sh.copy_stream));
    }
    // anything the copy has never seen is (re)quantised together with the requested range
    if (sh.q8_rows < first) {
        n += first - sh.q8_rows;
        first = sh.q8_rows;
    }
    if (first + n > sh.rows) n = sh.rows > first ? sh.rows - first : 0;
    if (n) {
        int rc = launch_quantize(sh, sh.d_rows, idx->pitch, idx->dim, first, n, sh.d_q8, pitch8,
                                 sh.d_meta, sh.d_norms, sh.d_q8_flag, sh.copy_stream);
        if (rc) return rc;
        CUDA_TRY(cudaStreamSynchronize(sh.copy_stream));
    }
    sh.q8_rows = sh.rows;
    return build_tmap8(idx, sh);
}

// Auto mode (nm_index_set_prefilter(idx, 2), the default): the first batch that the
// tensor-core pre-filter can serve builds the int8 copy, provided it leaves a comfortable
// margin of free HBM (results are bit-identical either way, so this is purely a memory
// decision).  Called WITHOUT the index lock; takes the write lock for the build.
int q8_auto_prepare(nm_index *idx, uint32_t nq, uint32_t k) {
    if (idx->prefilter.load() != 2 || !idx->tensor_core.load() || !idx->batching.load()) return NM_OK;
    {
        std::shared_lock<std::shared_mutex> g(idx->mu);
        bool wanted = false;
        for (auto &sh : idx->shards)
            if (tc_shape_ok(idx, sh->rows, nq, k) && !(sh->tmap8_valid && sh->q8_rows == sh->rows))
                wanted = true;
        if (!wanted || idx->q8_auto_declined_rows.load() == idx->total_rows()) return NM_OK;
    }
    std::unique_lock<std::shared_mutex> g(idx->mu);
    if (idx->prefilter.load() != 2) return NM_OK;
    const uint32_t pitch8 = q8_pitch(idx->dim);
    for (auto &shp : idx->shards) {
        Shard &sh = *shp;
        if (!tc_shape_ok(idx, sh.rows, nq, k) || (sh.tmap8_valid && sh.q8_rows == sh.rows)) continue;
        CUDA_TRY(cudaSetDevice(sh.device));
        // async searches on caller streams never read the copy, but they may be reading memory
        // the allocator is about to hand out: nothing to wait for here (fresh allocations only)
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        const uint64_t cap = std::max<uint64_t>(sh.rows, sh.capacity);
        const uint64_t need = sh.q8_capacity >= sh.rows
                                  ? 0
                                  : cap * (pitch8 + sizeof(nm::RowMeta) + sizeof(float2));
        const uint64_t margin = std::max<uint64_t>(4ull << 30, total_b / 16);
        if (need + margin > free_b) {
            idx->q8_auto_declined_rows = idx->total_rows();
            return NM_OK;  // not an error: batches stay on the exact kernels
        }
        int rc = q8_refresh(idx, sh, 0, sh.rows, true);
        if (rc) return rc;
    }
    return NM_OK;
}

}  // namespace nmi

namespace {

int shard_reserve(nm_index *idx, Shard &sh, uint64_t rows, bool keep) {
    if (rows <= sh.capacity) return NM_OK;
    if (rows > nm::kMaxLocalRows)
        return fail(NM_ERR_INVALID_ARGUMENT, "shard would hold %llu rows; limit is %u per device",
                    (unsigned long long)rows, nm::kMaxLocalRows);
    // keep == appending: grow geometrically, IN PLACE (more physical chunks mapped behind the
    // existing rows; see nm_vmm.hpp).  !keep == bulk load: map exactly what the load needs.
    const size_t row_bytes = (size_t)idx->pitch * sizeof(float);
    int rc = sh.rows_buf.ensure(sh.device, rows * row_bytes, !keep);
    if (rc) return rc;
    sh.d_rows = static_cast<float *>(sh.rows_buf.ptr());
    sh.capacity = sh.rows_buf.mapped() / row_bytes;
    return NM_OK;
}

// Host rows [n, dim] -> device rows [first, first+n) of the shard.  Pinned sources are DMA'd
// directly; pageable sources go through two pinned staging buffers so the host memcpy of
// chunk i+1 overlaps the DMA of chunk i.
int shard_upload(nm_index *idx, Shard &sh, uint64_t first, const float *src, uint64_t n) {
    if (n == 0) return NM_OK;
    const size_t row_bytes = (size_t)idx->dim * 4, pitch_bytes = (size_t)idx->pitch * 4;
    float *dst = sh.d_rows + first * idx->pitch;
    cudaPointerAttributes attr;
    bool pinned = cudaPointerGetAttributes(&attr, src) == cudaSuccess &&
                  attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (idx->pitch != idx->dim)
        CUDA_TRY(cudaMemsetAsync(dst, 0, n * pitch_bytes, sh.copy_stream));
    if (pinned) {
        CUDA_TRY(cudaMemcpy2DAsync(dst, pitch_bytes, src, row_bytes, row_bytes, n,
                                   cudaMemcpyHostToDevice, sh.copy_stream));
    } else {
        if (!sh.staging[0]) {
            for (int b = 0; b < 2; ++b) {
                CUDA_TRY(cudaMallocHost(&sh.staging[b], kStagingBytes));
                CUDA_TRY(cudaEventCreateWithFlags(&sh.staging_done[b], cudaEventDisableTiming));
            }
        }
        const uint64_t rows_per_chunk = std::max<uint64_t>(1, kStagingBytes / row_bytes);
        if (row_bytes > kStagingBytes)
            return fail(NM_ERR_DIMENSION_MISMATCH, "dimension %u too large for staging", idx->dim);
        int b = 0;
        for (uint64_t r = 0; r < n; r += rows_per_chunk, b ^= 1) {
            uint64_t m = std::min(rows_per_chunk, n - r);
            CUDA_TRY(cudaEventSynchronize(sh.staging_done[b]));
            parallel_memcpy(sh.staging[b], src + r * idx->dim, m * row_bytes);
            CUDA_TRY(cudaMemcpy2DAsync(dst + r * idx->pitch, pitch_bytes, sh.staging[b], row_bytes,
                                       row_bytes, m, cudaMemcpyHostToDevice, sh.copy_stream));
            CUDA_TRY(cudaEventRecord(sh.staging_done[b], sh.copy_stream));
        }
    }
    CUDA_TRY(cudaStreamSynchronize(sh.copy_stream));
    idx->h2d_bytes += n * row_bytes;
    return NM_OK;
}

}  // namespace

// ======================================================================================
// C ABI: library, mirror lifecycle and mutations
// ======================================================================================
extern "C" {

int nm_abi_version(void) { return NM_ABI_VERSION; }

const char *nm_last_error(void) { return nmi::last_error(); }

int nm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int nm_index_create(uint32_t dim, const int *devices, int n_dev, nm_index **out) {
    if (!out) return fail(NM_ERR_INVALID_ARGUMENT, "null out pointer");
    *out = nullptr;
    if (dim == 0) return fail(NM_ERR_EMPTY_VECTOR, "dimension must be >= 1");
    int avail = nm_device_count();
    if (avail == 0)
        return fail(NM_ERR_STORAGE, "no CUDA device visible: the SIMILAR scan has no CPU path");
    std::vector<int> devs;
    if (!devices || n_dev <= 0) {
        int cur = 0;
        CUDA_TRY(cudaGetDevice(&cur));
        devs.push_back(cur);
    } else {
        for (int i = 0; i < n_dev; ++i) {
            if (devices[i] < 0 || devices[i] >= avail)
                return fail(NM_ERR_INVALID_ARGUMENT, "device %d out of range (0..%d)", devices[i],
                            avail - 1);
            devs.push_back(devices[i]);
        }
    }
    std::unique_ptr<nm_index> idx(new nm_index());
    idx->dim = dim;
    idx->pitch = (dim + 3u) & ~3u;
    for (int d : devs) {
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, d));
        if (prop.major < 10)
            return fail(NM_ERR_STORAGE, "device %d is sm_%d%d; this library is built for sm_100a", d,
                        prop.major, prop.minor);
        std::unique_ptr<Shard> sh(new Shard());
        sh->device = d;
        sh->sm_count = prop.multiProcessorCount;
        CUDA_TRY(cudaSetDevice(d));
        CUDA_TRY(cudaStreamCreateWithFlags(&sh->copy_stream, cudaStreamNonBlocking));
        idx->shards.push_back(std::move(sh));
    }
    *out = idx.release();
    return NM_OK;
}

void nm_index_destroy(nm_index *idx) {
    if (!idx) return;
    if (idx->comm && nccl().ok) {
        cudaSetDevice(idx->shards[0]->device);
        cudaDeviceSynchronize();
        teardown_peer_exchange(idx);
        nccl().CommDestroy(idx->comm);
        if (idx->xchg_mem) cudaFree(idx->xchg_mem);
    }
    for (auto &sh : idx->shards) {
        cudaSetDevice(sh->device);
        cudaDeviceSynchronize();  // asynchronous nm_search_device work may still be in flight
        sh->pool.clear();
        sh->stream_ws.clear();
        sh->rows_buf.release();
        sh->q8_buf.release();
        sh->meta_buf.release();
        sh->norms_buf.release();
        if (sh->d_q8_flag) cudaFree(sh->d_q8_flag);
        for (int b = 0; b < 2; ++b) {
            if (sh->staging[b]) cudaFreeHost(sh->staging[b]);
            if (sh->staging_done[b]) cudaEventDestroy(sh->staging_done[b]);
        }
        if (sh->copy_stream) cudaStreamDestroy(sh->copy_stream);
    }
    delete idx;
}

int nm_index_clear(nm_index *idx) {
    if (!idx) return fail(NM_ERR_INVALID_ARGUMENT, "null index");
    std::unique_lock<std::shared_mutex> g(idx->mu);
    if (int wrc = wait_async_searches(idx)) return wrc;  // never tear an in-flight async scan
    for (auto &sh : idx->shards) {
        sh->rows = 0;
        sh->row_base = 0;
        sh->tmap_valid = false;
        sh->q8_rows = 0;
        sh->tmap8_valid = false;
        columns_drop(*sh);
    }
    idx->mutation_epoch++;
    return NM_OK;
}

int nm_index_load(nm_index *idx, const float *rows, uint64_t n) {
    NM_TRACE("nm_index_load");
    if (!idx) return fail(NM_ERR_INVALID_ARGUMENT, "null index");
    if (n && !rows) return fail(NM_ERR_INVALID_ARGUMENT, "null rows");
    std::unique_lock<std::shared_mutex> g(idx->mu);
    if (int wrc = wait_async_searches(idx)) return wrc;  // never tear an in-flight async scan
    const uint64_t G = idx->shards.size();
    for (uint64_t s = 0; s < G; ++s) {
        Shard &sh = *idx->shards[s];
        uint64_t lo = n * s / G, hi = n * (s + 1) / G;
        CUDA_TRY(cudaSetDevice(sh.device));
        sh.rows = 0;
        columns_drop(sh);
        idx->mutation_epoch++;
        int rc = shard_reserve(idx, sh, hi - lo, false);
        if (rc) return rc;
        rc = shard_upload(idx, sh, 0, rows + lo * idx->dim, hi - lo);
        if (rc) return rc;
        sh.rows = hi - lo;
        sh.row_base = lo;
        rc = build_tmap(idx, sh);
        if (rc) return rc;
        sh.q8_rows = 0;
        rc = q8_refresh(idx, sh, 0, sh.rows);
        if (rc) return rc;
    }
    return NM_OK;
}

// Global row ids are positions in the concatenation of the shards: keep every shard's base the
// prefix sum of the shards before it (appends and swap-removes change the last shards' sizes).
static void recompute_row_bases(nm_index *idx) {
    uint64_t base = 0;
    for (auto &sh : idx->shards) {
        sh->row_base = base;
        base += sh->rows;
    }
}

// Device-to-device copy of `n` mirror rows between two shards; different devices go through the
// source shard's pinned staging buffer (works with or without peer access).
static int copy_rows_between(nm_index *idx, Shard &src, uint64_t src_row, Shard &dst, float *dst_base,
                             uint64_t dst_row, uint64_t n) {
    const size_t row_bytes = (size_t)idx->pitch * 4;
    if (n == 0) return NM_OK;
    if (src.device == dst.device) {
        CUDA_TRY(cudaSetDevice(src.device));
        CUDA_TRY(cudaMemcpyAsync(dst_base + dst_row * idx->pitch, src.d_rows + src_row * idx->pitch,
                                 n * row_bytes, cudaMemcpyDeviceToDevice, src.copy_stream));
        CUDA_TRY(cudaStreamSynchronize(src.copy_stream));
        return NM_OK;
    }
    CUDA_TRY(cudaSetDevice(src.device));
    if (!src.staging[0]) {
        for (int b = 0; b < 2; ++b) {
            CUDA_TRY(cudaMallocHost(&src.staging[b], kStagingBytes));
            CUDA_TRY(cudaEventCreateWithFlags(&src.staging_done[b], cudaEventDisableTiming));
        }
    }
    const uint64_t per = std::max<uint64_t>(1, kStagingBytes / row_bytes);
    if (row_bytes > kStagingBytes)
        return fail(NM_ERR_DIMENSION_MISMATCH, "dimension %u too large for staging", idx->dim);
    for (uint64_t r = 0; r < n; r += per) {
        const uint64_t m = std::min(per, n - r);
        CUDA_TRY(cudaSetDevice(src.device));
        CUDA_TRY(cudaMemcpyAsync(src.staging[0], src.d_rows + (src_row + r) * idx->pitch,
                                 m * row_bytes, cudaMemcpyDeviceToHost, src.copy_stream));
        CUDA_TRY(cudaStreamSynchronize(src.copy_stream));
        CUDA_TRY(cudaSetDevice(dst.device));
        CUDA_TRY(cudaMemcpyAsync(dst_base + (dst_row + r) * idx->pitch, src.staging[0], m * row_bytes,
                                 cudaMemcpyHostToDevice, dst.copy_stream));
        CUDA_TRY(cudaStreamSynchronize(dst.copy_stream));
    }
    return NM_OK;
}

// In-process multi-device indexes: appends land on the last shard (the end of the global row
// order).  Once a shard holds more than 1.5x its share, the rows are re-split into equal contiguous
// ranges [s*N/G, (s+1)*N/G) — global row ids do not change, so the caller's key table stays
// valid.  Each shard is rebuilt in a fresh buffer (old + new coexist per device while it runs).
static int rebalance_shards(nm_index *idx, bool force) {
    const uint64_t G = idx->shards.size();
    if (G < 2) return NM_OK;
    const uint64_t N = idx->total_rows();
    uint64_t mx = 0;
    for (auto &sh : idx->shards) mx = std::max(mx, sh->rows);
    const uint64_t share = (N + G - 1) / G;
    if (!force && mx <= share + share / 2 + 4096) return NM_OK;
    std::shared_ptr<ColumnsSnapshot> cols;
    {
        int rc = columns_gather(idx, &cols);  // metadata columns follow the rows
        if (rc) return rc;
    }
    idx->mutation_epoch++;
    std::vector<GrowBuf> nb(G);  // (default-constructed in place: GrowBuf is not copyable)
    const size_t row_bytes = (size_t)idx->pitch * 4;
    for (uint64_t s = 0; s < G; ++s) {
        Shard &dst = *idx->shards[s];
        const uint64_t lo = N * s / G, hi = N * (s + 1) / G;
        if (hi == lo) continue;
        CUDA_TRY(cudaSetDevice(dst.device));
        int rc = nb[s].ensure(dst.device, (hi - lo) * row_bytes, true);
        if (rc) return rc;
        float *base = static_cast<float *>(nb[s].ptr());
        for (auto &tp : idx->shards) {
            Shard &src = *tp;
            const uint64_t a = std::max(lo, src.row_base), b = std::min(hi, src.row_base + src.rows);
            if (a >= b) continue;
            rc = copy_rows_between(idx, src, a - src.row_base, dst, base, a - lo, b - a);
            if (rc) return rc;
        }
    }
    for (uint64_t s = 0; s < G; ++s) {
        Shard &sh = *idx->shards[s];
        const uint64_t lo = N * s / G, hi = N * (s + 1) / G;
        CUDA_TRY(cudaSetDevice(sh.device));
        sh.rows_buf.swap(nb[s]);
        nb[s].release();
        sh.d_rows = static_cast<float *>(sh.rows_buf.ptr());
        sh.capacity = sh.rows_buf.mapped() / row_bytes;
        sh.rows = hi - lo;
        sh.row_base = lo;
        int rc = build_tmap(idx, sh);
        if (rc) return rc;
        sh.q8_rows = 0;
        sh.tmap8_valid = false;
        rc = q8_refresh(idx, sh, 0, sh.rows);
        if (rc) return rc;
    }
    return columns_scatter(idx, *cols);
}

int nm_index_append(nm_index *idx, const float *rows, uint64_t n) {
    NM_TRACE("nm_index_append");
    if (!idx) return fail(NM_ERR_INVALID_ARGUMENT, "null index");
    if (n == 0) return NM_OK;
    if (!rows) return fail(NM_ERR_INVALID_ARGUMENT, "null rows");
    std::unique_lock<std::shared_mutex> g(idx->mu);
    if (int wrc = wait_async_searches(idx)) return wrc;  // never tear an in-flight async scan
    Shard &sh = *idx->shards.back();
    CUDA_TRY(cudaSetDevice(sh.device));
    int rc = shard_reserve(idx, sh, sh.rows + n, true);
    if (rc) return rc;
    rc = shard_upload(idx, sh, sh.rows, rows, n);
    if (rc) return rc;
    sh.rows += n;
    idx->mutation_epoch++;
    recompute_row_bases(idx);
    rc = build_tmap(idx, sh);
    if (rc) return rc;
    rc = columns_after_resize(idx, sh);  // the new rows read as "missing" in every column
    if (rc) return rc;
    rc = q8_refresh(idx, sh, sh.rows - n, n);
    if (rc) return rc;
    return rebalance_shards(idx, false);
}

static int locate_row(nm_index *idx, uint64_t row, Shard **out, uint64_t *local) {
    for (auto &sh : idx->shards) {
        if (row >= sh->row_base && row < sh->row_base + sh->rows) {
            *out = sh.get();
            *local = row - sh->row_base;
            return NM_OK;
        }
    }
    return fail(NM_ERR_INVALID_ARGUMENT, "row %llu out of range", (unsigned long long)row);
}

int nm_index_update(nm_index *idx, uint64_t row, const float *vec) {
    NM_TRACE("nm_index_update");
    if (!idx || !vec) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
    std::unique_lock<std::shared_mutex> g(idx->mu);
    if (int wrc = wait_async_searches(idx)) return wrc;  // never tear an in-flight async scan
    Shard *sh = nullptr;
    uint64_t local = 0;
    int rc = locate_row(idx, row, &sh, &local);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(sh->device));
    rc = shard_upload(idx, *sh, local, vec, 1);
    if (rc) return rc;
    return q8_refresh(idx, *sh, local, 1);
}

int nm_index_swap_remove(nm_index *idx, uint64_t row, uint64_t *moved_from) {
    NM_TRACE("nm_index_swap_remove");
    if (!idx) return fail(NM_ERR_INVALID_ARGUMENT, "null index");
    std::unique_lock<std::shared_mutex> g(idx->mu);
    if (int wrc = wait_async_searches(idx)) return wrc;  // never tear an in-flight async scan
    Shard *sh = nullptr;
    uint64_t local = 0;
    int rc = locate_row(idx, row, &sh, &local);
    if (rc) return rc;
    // the globally last row lives in the last non-empty shard
    Shard *last = nullptr;
    for (auto it = idx->shards.rbegin(); it != idx->shards.rend(); ++it)
        if ((*it)->rows) {
            last = it->get();
            break;
        }
    uint64_t last_global = last->row_base + last->rows - 1;
    if (moved_from) *moved_from = last_global;
    if (last_global != row) {
        rc = copy_rows_between(idx, *last, last->rows - 1, *sh, sh->d_rows, local, 1);
        if (rc) return rc;
        rc = columns_swap_remove(idx, *sh, local, *last, last->rows - 1);
        if (rc) return rc;
    }
    last->rows -= 1;
    idx->mutation_epoch++;
    rc = columns_after_resize(idx, *last);
    if (rc) return rc;
    recompute_row_bases(idx);
    CUDA_TRY(cudaSetDevice(last->device));
    rc = build_tmap(idx, *last);
    if (rc) return rc;
    last->q8_rows = std::min(last->q8_rows, last->rows);
    rc = q8_refresh(idx, *last, last->rows, 0);
    if (rc) return rc;
    if (last_global != row) {
        CUDA_TRY(cudaSetDevice(sh->device));
        rc = q8_refresh(idx, *sh, local, 1);  // the moved row took this slot
    }
    return rc;
}

int nm_index_get_row(nm_index *idx, uint64_t row, float *out_vec) {
    if (!idx || !out_vec) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
    std::shared_lock<std::shared_mutex> g(idx->mu);
    Shard *sh = nullptr;
    uint64_t local = 0;
    int rc = locate_row(idx, row, &sh, &local);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(sh->device));
    CUDA_TRY(cudaMemcpy(out_vec, sh->d_rows + local * idx->pitch, (size_t)idx->dim * 4,
                        cudaMemcpyDeviceToHost));
    return NM_OK;
}

int nm_index_get_rows(nm_index *idx, uint64_t first, uint64_t n, float *out_rows) {
    if (!idx || (n && !out_rows)) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
    std::shared_lock<std::shared_mutex> g(idx->mu);
    if (first + n < first || first + n > idx->total_rows())
        return fail(NM_ERR_INVALID_ARGUMENT, "rows [%llu, %llu) out of range",
                    (unsigned long long)first, (unsigned long long)(first + n));
    const size_t row_bytes = (size_t)idx->dim * 4, pitch_bytes = (size_t)idx->pitch * 4;
    for (auto &sh : idx->shards) {
        const uint64_t lo = std::max(first, sh->row_base);
        const uint64_t hi = std::min(first + n, sh->row_base + sh->rows);
        if (lo >= hi) continue;
        CUDA_TRY(cudaSetDevice(sh->device));
        CUDA_TRY(cudaMemcpy2D(out_rows + (lo - first) * idx->dim, row_bytes,
                              sh->d_rows + (lo - sh->row_base) * idx->pitch, pitch_bytes, row_bytes,
                              hi - lo, cudaMemcpyDeviceToHost));
    }
    return NM_OK;
}

uint64_t nm_index_rows(const nm_index *idx) {
    if (!idx) return 0;
    std::shared_lock<std::shared_mutex> g(idx->mu);
    return idx->total_rows();
}
uint32_t nm_index_dim(const nm_index *idx) { return idx ? idx->dim : 0; }
int nm_index_device_count(const nm_index *idx) { return idx ? (int)idx->shards.size() : 0; }

int nm_index_shard_info(nm_index *idx, int shard, nm_shard_info *out) {
    if (!idx || !out) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
    std::shared_lock<std::shared_mutex> g(idx->mu);
    if (shard < 0 || shard >= (int)idx->shards.size())
        return fail(NM_ERR_INVALID_ARGUMENT, "shard %d out of range", shard);
    const Shard &sh = *idx->shards[shard];
    out->device = sh.device;
    out->rows = sh.rows;
    out->row_base = sh.row_base;
    out->capacity_rows = sh.capacity;
    out->mapped_bytes = sh.rows_buf.mapped();
    out->reserved_bytes = sh.rows_buf.reserved();
    out->chunks = sh.rows_buf.chunks();
    out->remaps = sh.rows_buf.remaps();
    out->q8_rows = sh.d_q8 ? sh.q8_rows : 0;
    out->grows_in_place = sh.rows_buf.vmm() ? 1 : 0;
    return NM_OK;
}

int nm_index_fill_synthetic(nm_index *idx, uint64_t n, uint64_t seed, uint64_t row_offset) {
    NM_TRACE("nm_index_fill_synthetic");
    if (!idx) return fail(NM_ERR_INVALID_ARGUMENT, "null index");
    std::unique_lock<std::shared_mutex> g(idx->mu);
    if (int wrc = wait_async_searches(idx)) return wrc;  // never tear an in-flight async scan
    const uint64_t G = idx->shards.size();
    for (uint64_t s = 0; s < G; ++s) {
        Shard &sh = *idx->shards[s];
        uint64_t lo = n * s / G, hi = n * (s + 1) / G;
        CUDA_TRY(cudaSetDevice(sh.device));
        sh.rows = 0;
        columns_drop(sh);
        idx->mutation_epoch++;
        int rc = shard_reserve(idx, sh, hi - lo, false);
        if (rc) return rc;
        if (hi > lo) {
            rc = launch_fill_synthetic(sh, sh.d_rows, hi - lo, idx->dim, idx->pitch, seed,
                                       row_offset + lo, sh.copy_stream);
            if (rc) return rc;
            CUDA_TRY(cudaStreamSynchronize(sh.copy_stream));
        }
        sh.rows = hi - lo;
        sh.row_base = lo;
        rc = build_tmap(idx, sh);
        if (rc) return rc;
        sh.q8_rows = 0;
        rc = q8_refresh(idx, sh, 0, sh.rows);
        if (rc) return rc;
    }
    return NM_OK;
}

int nm_index_set_prefilter(nm_index *idx, int mode) {
    if (!idx) return fail(NM_ERR_INVALID_ARGUMENT, "null index");
    if (mode < 0 || mode > 2) return fail(NM_ERR_INVALID_ARGUMENT, "unknown pre-filter mode %d", mode);
    std::unique_lock<std::shared_mutex> g(idx->mu);
    if (int wrc = wait_async_searches(idx)) return wrc;  // never tear an in-flight async scan
    idx->prefilter = mode;
    idx->q8_auto_declined_rows = ~0ull;
    for (auto &shp : idx->shards) {
        Shard &sh = *shp;
        CUDA_TRY(cudaSetDevice(sh.device));
        if (mode == 0) {
            sh.q8_buf.release();
            sh.meta_buf.release();
            sh.norms_buf.release();
            sh.d_q8 = nullptr;
            sh.d_meta = nullptr;
            sh.d_norms = nullptr;
            sh.q8_capacity = sh.q8_rows = 0;
            sh.tmap8_valid = false;
        } else if (mode == 1) {
            sh.q8_rows = 0;
            int rc = q8_refresh(idx, sh, 0, sh.rows);
            if (rc) return rc;
        }  // mode 2 keeps an existing copy and otherwise builds it at the first eligible batch
    }
    return NM_OK;
}

int nm_index_set_tensor_core(nm_index *idx, int enable) {
    if (!idx) return fail(NM_ERR_INVALID_ARGUMENT, "null index");
    idx->tensor_core = enable ? 1 : 0;
    return NM_OK;
}

int nm_debug_q8_row(nm_index *idx, uint64_t row, int8_t *out_q8, float *out_scale) {
    if (!idx || !out_q8 || !out_scale) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
    std::shared_lock<std::shared_mutex> g(idx->mu);
    if (idx->shards.size() != 1 || !idx->shards[0]->d_q8 || row >= idx->shards[0]->q8_rows)
        return fail(NM_ERR_CONFIGURATION, "needs a single-device index with the pre-filter on");
    Shard &sh = *idx->shards[0];
    CUDA_TRY(cudaSetDevice(sh.device));
    const uint32_t pitch8 = q8_pitch(idx->dim);
    CUDA_TRY(cudaMemcpy(out_q8, sh.d_q8 + row * pitch8, idx->dim, cudaMemcpyDeviceToHost));
    nm::RowMeta m;
    CUDA_TRY(cudaMemcpy(&m, sh.d_meta + row, sizeof(m), cudaMemcpyDeviceToHost));
    *out_scale = m.scale;
    return NM_OK;
}

int nm_index_set_batching(nm_index *idx, int enable) {
    if (!idx) return fail(NM_ERR_INVALID_ARGUMENT, "null index");
    idx->batching = enable ? 1 : 0;
    return NM_OK;
}

int nm_index_set_profiling(nm_index *idx, int enable) {
    if (!idx) return fail(NM_ERR_INVALID_ARGUMENT, "null index");
    idx->profiling = enable ? 1 : 0;
    return NM_OK;
}

int nm_index_stats(nm_index *idx, nm_stats *out) {
    if (!idx || !out) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
    out->searches = idx->searches;
    out->rows_scanned = idx->rows_scanned;
    out->bytes_streamed = idx->bytes_streamed;
    out->scan_launches = idx->scan_launches;
    out->merge_launches = idx->merge_launches;
    out->h2d_bytes = idx->h2d_bytes;
    out->d2h_bytes = idx->d2h_bytes;
    out->last_scan_ms = idx->last_scan_ms;
    out->prefilter_queries = idx->pf_queries;
    out->prefilter_fallbacks = idx->pf_fallbacks;
    out->prefilter_kept = idx->pf_kept;
    out->tc_queries = idx->tc_queries;
    out->tc_fallbacks = idx->tc_fallbacks;
    out->tc_survivors = idx->tc_survivors;
    out->coalesced_batches = idx->co_batches;
    out->coalesced_queries = idx->co_queries;
    out->filter_masks_built = idx->filter_masks_built;
    out->filter_mask_hits = idx->filter_mask_hits;
    {
        // fold finished profiling event pairs into the totals (waits for the streams)
        std::unique_lock<std::shared_mutex> g(idx->mu);
        for (auto &sh : idx->shards) {
            cudaSetDevice(sh->device);
            std::lock_guard<std::mutex> pg(sh->pool_mu);
            std::vector<Workspace *> all_ws;
            for (auto &e : sh->stream_ws) all_ws.push_back(e.second.get());
            for (auto &w : sh->pool) all_ws.push_back(w.get());
            for (Workspace *wsp : all_ws) {
                Workspace &ws = *wsp;
                for (size_t i = 0; i < ws.prof_used; ++i) {
                    float ms = 0.f;
                    if (cudaEventSynchronize(ws.prof_events[i].second) == cudaSuccess &&
                        cudaEventElapsedTime(&ms, ws.prof_events[i].first,
                                             ws.prof_events[i].second) == cudaSuccess) {
                        idx->profiled_scan_ms += ms;
                        idx->profiled_scans += 1;
                    }
                }
                ws.prof_used = 0;
            }
        }
        cudaGetLastError();
        out->profiled_scan_ms = idx->profiled_scan_ms;
        out->profiled_scans = idx->profiled_scans;
    }
    return NM_OK;
}

}  // extern "C"
