// rto_nvtx.h -- NVTX ranges around the library's entry points (tracing).  nvtx3 is header-only: without a profiler
// attached a range is one load and a branch; under Nsight Systems / Compute every C-ABI call shows up as a named range with the kernels,
// copies and host work it caused underneath.  RTO_NO_NVTX compiles them out.
#pragma once
#if !defined(RTO_NO_NVTX) && defined(__has_include)
#if !__has_include(<nvtx3/nvToolsExt.h>)
#define RTO_NO_NVTX 1      // host files compiled on their own with a plain g++ (tests/dc_asan): no CUDA include path, no ranges
#endif
#endif
#if !defined(RTO_NO_NVTX)
#include <nvtx3/nvToolsExt.h>
struct RtoRange {
	explicit RtoRange(const char* name) { nvtxRangePushA(name); }
	~RtoRange() { nvtxRangePop(); }
	RtoRange(const RtoRange&) = delete;
	RtoRange& operator=(const RtoRange&) = delete;
};
#define RTO_RANGE_CAT2(a, b) a##b
#define RTO_RANGE_CAT(a, b) RTO_RANGE_CAT2(a, b)
#define RTO_RANGE(name) RtoRange RTO_RANGE_CAT(rtoRange_, __LINE__)(name)
#else
#define RTO_RANGE(name) do {} while (0)
#endif
