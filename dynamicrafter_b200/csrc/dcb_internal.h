// Internal helpers shared by the translation units of libdcb200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

#include "../../include/dcb200.h"

namespace dcb {
int set_error(int code, const char* fmt, ...);   // records a thread-local message, returns code
constexpr int kMaxDevices = 16;                   // per-device caches (function attributes, SM counts) are indexed by ordinal & 15
int current_device();
int num_sms();                                    // SM count of the current device (cached per device)
int check_launch(const char* what);               // cudaGetLastError -> status
bool pdl_enabled(int cls);                        // programmatic dependent launch per kernel class (DCB_PDL bit mask)
constexpr int kPdlGemm = 1, kPdlLayerNorm = 2, kPdlGroupNorm = 4, kPdlAttention = 8, kPdlGnStream = 16;
// tcgen05 flash attention (flash_tc.cu); DCB_ERR_UNSUPPORTED if the arguments do not fit that kernel
int flash_tc_launch(const dcb_attention_args* a, cudaStream_t stream);
int flash_pp_launch(const dcb_attention_args* a, cudaStream_t stream);   // ping-pong variant (flash_pp.cu)
// tcgen05 text + image cross-attention and temporal attention (attn_tc.cu); DCB_ERR_UNSUPPORTED if the arguments do not fit
int cross_tc_launch(const dcb_attention_args* a, cudaStream_t stream);
int temporal_tc_launch(const void* q, const void* k, const void* v, void* o, long long ldq, long long ldk, long long ldv,
                       long long ldo, int B, int L, int HW, int heads, float softmax_scale, cudaStream_t stream);

// Launch with programmatic stream serialization (the kernel MUST call pdl_wait() before touching global memory).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(int cls, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    if (pdl_enabled(cls)) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
}  // namespace dcb
