// Internal helpers shared by the translation units of libdcb200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

#include "../../include/dcb200.h"

namespace dcb {
int set_error(int code, const char* fmt, ...);   // records a thread-local message, returns code
int num_sms();                                    // SM count of the current device (cached)
int check_launch(const char* what);               // cudaGetLastError -> status
// tcgen05 flash attention (flash_tc.cu); DCB_ERR_UNSUPPORTED if the arguments do not fit that kernel
int flash_tc_launch(const dcb_attention_args* a, cudaStream_t stream);
}  // namespace dcb
