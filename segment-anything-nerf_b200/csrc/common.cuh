// Shared helpers for libsanerf_b200 (sm_100a only).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>

#include "sanerf_b200.h"

namespace sanerf {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

template <typename T>
__host__ __device__ constexpr T div_up(T a, T b) { return (a + b - 1) / b; }

// thread-local error text returned by sanerf_last_error()
char* error_buffer();
int fail(int status, const char* fmt, ...);

// Every launcher ends with this: surfaces launch-configuration errors (never syncs).
inline int check_launch(const char* what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(SANERF_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    }
    return SANERF_OK;
}

#define SANERF_REQUIRE_PTR(p)                                                        \
    do {                                                                             \
        if ((p) == nullptr) return ::sanerf::fail(SANERF_ERR_NULL_POINTER, #p " is NULL"); \
    } while (0)

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------
// Every kernel of the training / rendering steps is launched with programmatic stream serialization and starts with
// pdl_begin() = griddepcontrol.wait: the grid is set up while its predecessor drains and blocks there until the
// predecessor has completed and its writes are visible.  Along chains of 20-45 short dependent kernels (one CUDA graph per
// step) that removes ~1.7 % of the RGB step (0.876 -> 0.862 ms).  Letting dependents be SCHEDULED early as well
// (griddepcontrol.launch_dependents at kernel entry, -DSANERF_PDL_EARLY_TRIGGER) was measured slower (0.908 ms): parked
// CTAs take the thread slots the parallel branches of the step need.  A kernel launched through launch_pdl MUST call
// pdl_begin() before touching memory.  SANERF_PDL=0 in the environment disables the launch attribute.
#ifdef SANERF_HEAD_TRACE
// Diagnostic builds only: block 0 of every launch stamps %globaltimer when it passes griddepcontrol.wait, i.e. when the kernel
// really starts inside a replayed CUDA graph (where events cannot bracket it).  id = source line of the pdl_begin() call * 64 +
// a 6-bit hash of the file name; tools/graph_timeline.py maps ids back to kernels.  The buffer pointer is a per-translation-unit
// __device__ variable (no relocatable device code in this build), set through a registry of per-unit setters (api.cu).
struct KernelStamp { unsigned long long t; unsigned int id, pad; };
struct StampBuf { unsigned int count, pad; KernelStamp s[4096]; };
static __device__ StampBuf* tu_stamp_buf = nullptr;
void register_stamp_tu(void (*set)(StampBuf*));
namespace {
inline void set_tu_stamp_buf(StampBuf* b) { cudaMemcpyToSymbol(tu_stamp_buf, &b, sizeof(b)); }
struct StampReg { StampReg() { register_stamp_tu(&set_tu_stamp_buf); } };
static StampReg stamp_reg_instance;
}  // namespace
__host__ __device__ constexpr unsigned int file_tag(const char* f) {
    unsigned int h = 0;
    for (int i = 0; f[i] != 0; ++i) h = (f[i] == '/') ? 0u : h * 31u + (unsigned char)f[i];      // hash of the base name
    return h & 63u;
}
__device__ __forceinline__ void stamp_start(unsigned int id) {
    StampBuf* b = tu_stamp_buf;
    if (b != nullptr) {
        const unsigned int slot = atomicAdd(&b->count, 1u);
        if (slot < 4096u) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            b->s[slot].t = t;
            b->s[slot].id = id;
        }
    }
}
#define pdl_begin()                                                                                                      \
    do {                                                                                                                 \
        asm volatile("griddepcontrol.wait;" ::: "memory");                                                               \
        if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0 && threadIdx.y == 0)               \
            ::sanerf::stamp_start((unsigned int)__LINE__ * 64u + ::sanerf::file_tag(__FILE__));                          \
    } while (0)
#else
__device__ __forceinline__ void pdl_begin() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
#ifdef SANERF_PDL_EARLY_TRIGGER
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
#endif

bool pdl_enabled();

template <typename... P, typename... A>
inline void launch_pdl(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, std::forward<A>(args)...);      // errors surface in check_launch()
}
#define SANERF_LAUNCH(kernel, grid, block, smem, stream, ...) \
    ::sanerf::launch_pdl(kernel, dim3(grid), dim3(block), (size_t)(smem), stream, __VA_ARGS__)

// ---- vector reductions into global memory (sm_90+: red.global.add.v2/v4.f32) -------------
__device__ __forceinline__ void red_add_f32(float* addr, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void red_add_v2_f32(float* addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_v4_f32(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b),
                 "f"(c), "f"(d)
                 : "memory");
}
__device__ __forceinline__ void red_add_f16x2(__half* addr, __half2 v) {
    uint32_t u = *reinterpret_cast<uint32_t*>(&v);
    asm volatile("red.global.add.noftz.f16x2 [%0], %1;" ::"l"(addr), "r"(u) : "memory");
}
__device__ __forceinline__ void red_add_v2_f16x2(__half* addr, __half2 a, __half2 b) {
    uint32_t ua = *reinterpret_cast<uint32_t*>(&a), ub = *reinterpret_cast<uint32_t*>(&b);
    asm volatile("red.global.add.noftz.v2.f16x2 [%0], {%1, %2};" ::"l"(addr), "r"(ua), "r"(ub)
                 : "memory");
}
__device__ __forceinline__ void red_add_v4_f16x2(__half* addr, __half2 a, __half2 b, __half2 c,
                                                 __half2 d) {
    uint32_t ua = *reinterpret_cast<uint32_t*>(&a), ub = *reinterpret_cast<uint32_t*>(&b);
    uint32_t uc = *reinterpret_cast<uint32_t*>(&c), ud = *reinterpret_cast<uint32_t*>(&d);
    asm volatile("red.global.add.noftz.v4.f16x2 [%0], {%1, %2, %3, %4};" ::"l"(addr), "r"(ua),
                 "r"(ub), "r"(uc), "r"(ud)
                 : "memory");
}

// read-only 128-bit / 64-bit loads through the non-coherent path
__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float2 ldg_f2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }

// streaming (evict-first) stores for write-once outputs
__device__ __forceinline__ void st_cs_f4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace sanerf
