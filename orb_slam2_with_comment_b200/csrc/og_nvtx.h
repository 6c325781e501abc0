// og_nvtx.h — NVTX ranges around the C-ABI entries and the extraction stages (tracing).  NVTX 3 is header-only:
// without a profiler attached a range is one load and a not-taken branch; with nsys / ncu --nvtx the entries and stages show
// up by name on the host timeline.
#pragma once
#include <nvtx3/nvToolsExt.h>

namespace og {
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};
}  // namespace og
#define OG_NVTX_CAT2(a, b) a##b
#define OG_NVTX_CAT(a, b) OG_NVTX_CAT2(a, b)
#define OG_NVTX(name) og::NvtxRange OG_NVTX_CAT(og_nvtx_range_, __LINE__)(name)
