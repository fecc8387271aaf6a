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
// Kernel attributes (cudaFuncSetAttribute) and SM counts are per DEVICE: a process that drives
// several GPUs must set them once on each.  `PerDeviceOnce::flag()` is the "already done" flag of
// the current device.
inline int current_device() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0) d = 0;
    return d > 63 ? 63 : d;
}
struct PerDeviceOnce {
    bool done[64] = {};
    bool& flag() { return done[current_device()]; }
};
// multiProcessorCount of the current device (cached; 148 on a full B200, fewer under MIG / green contexts)
int sm_count();
// Whether kernels are launched with programmatic stream serialization (PDL); sfb_set_pdl().
extern int g_pdl;

// Launch with (optionally) the programmatic-dependent-launch attribute.  Every kernel launched
// through here calls pdl_wait() before touching global memory written by its predecessor.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_cluster_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, dim3 cluster,
                                      size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int n = 0;
    if (g_pdl) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cluster.x * cluster.y * cluster.z > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = cluster.x;
        attr[n].val.clusterDim.y = cluster.y;
        attr[n].val.clusterDim.z = cluster.z;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, Args... args) {
    return launch_cluster_pdl(kernel, grid, block, dim3(1, 1, 1), smem, stream, args...);
}
}  // namespace sfb
