// common.cuh -- error plumbing shared by the translation units of libvkit_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>  // header only: resolves the tool's injection library at run time, no link dependency
#include <stdio.h>
#include "../../include/vkit_b200.h"

namespace vkb {
void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return VKB_ERR_CUDA;
    }
    return VKB_OK;
}
// NVTX range over one C-ABI call (visible in Nsight Systems / compute timelines; a no-op costing
// one pointer test when no tool is attached).
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};
}  // namespace vkb

#define VKB_NVTX(name) vkb::NvtxRange vkb_nvtx_range_(name)

#define VKB_REQUIRE(cond, msg)                  \
    do {                                        \
        if (!(cond)) {                          \
            vkb::set_error("%s: %s", __func__, msg); \
            return VKB_ERR_INVALID;             \
        }                                       \
    } while (0)

#define VKB_CUDA(call)                                                    \
    do {                                                                  \
        cudaError_t e_ = (call);                                          \
        if (e_ != cudaSuccess) {                                          \
            vkb::set_error("%s: %s", #call, cudaGetErrorString(e_));      \
            return VKB_ERR_CUDA;                                          \
        }                                                                 \
    } while (0)
