// Host-side pieces of the C ABI: error reporting, launch counter, TMA tensor-map encoding.
#include "host.h"

#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

namespace sfb {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};
int g_pdl = [] { const char* v = getenv("SFB_PDL"); return v ? atoi(v) : 1; }();

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(SFB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return SFB_OK;
}

int sm_count() {
    static int cached[64] = {};
    const int d = current_device();
    if (cached[d] <= 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || n <= 0) {
            cudaGetLastError();
            n = 148;
        }
        cached[d] = n;
    }
    return cached[d];
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) {
        cudaGetLastError();
        return nullptr;
    }
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    return fn;
}

}  // namespace sfb

using namespace sfb;

extern "C" int sfb_abi_version(void) { return SFB_ABI_VERSION; }
extern "C" const char* sfb_last_error(void) { return g_err; }
extern "C" uint64_t sfb_launch_count(void) { return g_launches.load(); }
extern "C" void sfb_set_pdl(int enable) { g_pdl = enable ? 1 : 0; }
extern "C" int sfb_sm_count(void) { return sm_count(); }

extern "C" int sfb_tmap_2d(void* out128, const void* base, uint64_t rows, uint64_t cols,
                           uint64_t pitch_elems, uint32_t box_rows) {
    auto enc = get_encode();
    if (!enc) return fail(SFB_ERR_NO_DRIVER, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (pitch_elems * 2) % 16 || box_rows == 0 ||
        box_rows > 256 || cols == 0 || rows == 0)
        return fail(SFB_ERR_INVALID, "sfb_tmap_2d: bad geometry rows=%llu cols=%llu pitch=%llu box=%u",
                    (unsigned long long)rows, (unsigned long long)cols,
                    (unsigned long long)pitch_elems, box_rows);
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {pitch_elems * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(reinterpret_cast<CUtensorMap*>(out128), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                     const_cast<void*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SFB_ERR_CUDA, "cuTensorMapEncodeTiled(2d) failed: %d", (int)r);
    return SFB_OK;
}

extern "C" int sfb_tmap_nhwc(void* out128, const void* base, uint32_t n, uint32_t h, uint32_t w,
                             uint32_t c, uint64_t pitch_elems, uint32_t box_n, uint32_t box_h,
                             uint32_t box_w, uint32_t stride) {
    auto enc = get_encode();
    if (!enc) return fail(SFB_ERR_NO_DRIVER, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (pitch_elems * 2) % 16 || pitch_elems < c ||
        (stride != 1 && stride != 2) || box_w * stride > 256 || box_h * stride > 256 ||
        box_n > 256 || !box_n || !box_h || !box_w)
        return fail(SFB_ERR_INVALID, "sfb_tmap_nhwc: bad geometry");
    cuuint64_t gdim[4] = {c, w, h, n};
    cuuint64_t gstr[3] = {pitch_elems * 2, (cuuint64_t)w * pitch_elems * 2,
                          (cuuint64_t)h * w * pitch_elems * 2};
    // with a traversal stride s the box spans box*s elements and yields `box` of them
    cuuint32_t box[4] = {64, box_w * stride, box_h * stride, box_n};
    cuuint32_t estr[4] = {1, stride, stride, 1};
    CUresult r = enc(reinterpret_cast<CUtensorMap*>(out128), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                     const_cast<void*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SFB_ERR_CUDA, "cuTensorMapEncodeTiled(nhwc) failed: %d", (int)r);
    return SFB_OK;
}

extern "C" int sfb_memset(void* p, int32_t value, size_t bytes, sfb_stream_t stream) {
    cudaError_t e = cudaMemsetAsync(p, value, bytes, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail(SFB_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    return SFB_OK;
}
