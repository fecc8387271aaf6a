// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/sfb200.h"

namespace sfb {
// Records a thread-local error message and returns `code`.
int fail(int code, const char* fmt, ...);
// Counts one kernel launch and converts cudaGetLastError() into an sfb_status.
int check_launch(const char* what);
}  // namespace sfb
