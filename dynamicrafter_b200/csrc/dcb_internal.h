// Internal helpers shared by the translation units of libdcb200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

#include "../../include/dcb200.h"

namespace dcb {
int set_error(int code, const char* fmt, ...);   // records a thread-local message, returns code
int num_sms();                                    // SM count of the current device (cached)
int check_launch(const char* what);               // cudaGetLastError -> status
}  // namespace dcb
