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
// Whether kernels are launched with programmatic stream serialization (PDL); sfb_set_pdl().
extern int g_pdl;

// Launch with (optionally) the programmatic-dependent-launch attribute.  Every kernel launched
// through here calls pdl_wait() before touching global memory written by its predecessor.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
}  // namespace sfb
