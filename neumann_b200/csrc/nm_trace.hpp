// nm_trace.hpp — NVTX ranges around the host-side phases of a search / a mutation: the
// tracing hook the reference gets from `#[instrument]` (vector_engine/src/lib.rs:1584, 1697,
// 1839, 1949, 2048).  NVTX v3 is header-only and resolves its injection library lazily: without
// a profiler attached a range costs one predictable branch.  Domain "neumann_b200"; ranges:
//   nm_search / nm_search_device / nm_search_masked / nm_search_filtered   whole call
//     stage_query     pinned copy + H2D of the queries
//     filter_mask     filter program -> device row mask (or cache hit)
//     scan            enqueue of the scan / batch / pre-filter kernels
//     exchange_merge  ncclAllGather + merge kernel (the fused path does it inside `scan`)
//     wait_download   D2H of the packed result + wait
//   nm_index_load / append / update / swap_remove / column_set        mutations
//     q8_refresh      upkeep of the int8 copy
#pragma once
#include <nvtx3/nvToolsExt.h>

namespace nmi {

inline nvtxDomainHandle_t trace_domain() {
    static nvtxDomainHandle_t d = nvtxDomainCreateA("neumann_b200");
    return d;
}

struct TraceRange {
    explicit TraceRange(const char *name) {
        nvtxEventAttributes_t a = {};
        a.version = NVTX_VERSION;
        a.size = NVTX_EVENT_ATTRIB_STRUCT_SIZE;
        a.messageType = NVTX_MESSAGE_TYPE_ASCII;
        a.message.ascii = name;
        nvtxDomainRangePushEx(trace_domain(), &a);
    }
    ~TraceRange() { nvtxDomainRangePop(trace_domain()); }
    TraceRange(const TraceRange &) = delete;
    TraceRange &operator=(const TraceRange &) = delete;
};

}  // namespace nmi

#define NM_TRACE_CAT2(a, b) a##b
#define NM_TRACE_CAT(a, b) NM_TRACE_CAT2(a, b)
#define NM_TRACE(name) ::nmi::TraceRange NM_TRACE_CAT(nm_trace_, __LINE__)(name)
