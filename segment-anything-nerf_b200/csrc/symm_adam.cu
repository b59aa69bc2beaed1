// Data-parallel parameter update as ONE kernel over NVLink / NVSwitch peer memory (sm_100a):
//
//   all ranks:  barrier  ->  this rank's 1/world slice:  g = sum over ranks of grad      (multimem.ld_reduce through the
//                                                         switch, or peer loads)
//                                                        Adam (+ optional per-step EMA) on the slice
//                                                        broadcast the new parameters     (multimem.st, or peer stores)
//               barrier  ->  clear the local gradient range
//
// It replaces the exchange the reference gets from DistributedDataParallel + torch.optim.Adam (one
// all-reduce of the gradient bucket, then identical full-size Adam passes on every rank) and this repository's own NCCL
// form (reduce-scatter -> Adam shard -> all-gather, sanerf_b200/parallel.py), which stays as the checker: no
// intermediate reduced-gradient buffer, no separate optimizer launch, no host launches between the three phases, and the
// optimizer traffic shrinks by the world size.  The flat parameter and gradient buffers live in symmetric memory
// (torch.distributed._symmetric_memory is only the allocator / rendezvous); the kernel receives the multicast addresses
// (NVLS) and / or the peers' unicast addresses.
//
// Cross-rank barrier: block b of rank r writes a monotonically increasing epoch into slot [b][r] of every peer's flag
// array (st.release.sys) and waits until its own slots [b][*] reached the epoch (ld.acquire.sys).  Blocks only ever wait
// for the SAME block index of the peers, and the grid is small enough to be co-resident, so progress needs nothing but
// every rank launching the kernel.  Calls that may run concurrently (the deferred table update beside the next step's
// front, the two tail ranges beside the hash-grid scatter) use different CHANNELS = disjoint slot ranges.  A bounded spin (about 4 s) raises an error flag instead of hanging the GPU.
#include "common.cuh"

namespace sanerf {

namespace symm {
constexpr uint32_t kThreads = 1024;
constexpr uint32_t kMaxWorld = 8;
constexpr uint32_t kMaxBlocks = 256;
// concurrent calls use disjoint flag slots: channel 0 (the deferred table update) owns 160 (one block per SM), channels 1
// and 2 own 32 each
constexpr uint32_t kChannels = 3;
__host__ __device__ constexpr uint32_t channel_slot0(uint32_t c) { return c == 0 ? 0u : 128u + 32u * c; }
__host__ __device__ constexpr uint32_t channel_blocks(uint32_t c) { return c == 0 ? 160u : 32u; }
}  // namespace symm

struct SymmAdamParams {
    float* param;            // local flat buffers (the same allocations the multicast / peer addresses map)
    float* grad;
    float* exp_avg;
    float* exp_avg_sq;
    float* ema;              // or NULL
    float* param_mc;         // multicast addresses of the flat buffers, or NULL: use the peer pointers
    float* grad_mc;
    float* param_peer[symm::kMaxWorld];
    float* grad_peer[symm::kMaxWorld];
    uint32_t* flags_peer[symm::kMaxWorld];   // [kMaxBlocks][kMaxWorld] epochs, one array per rank (symmetric)
    uint32_t* epoch;         // local [kMaxBlocks]: barriers completed so far by each block index
    uint32_t* error;         // local: set to 1 when a barrier timed out
    const float* dyn;        // {lr_t, 1-b1^t, 1-b2^t, 1-ema_t}
    const int32_t* gate;     // or NULL
    uint64_t start, stop;    // flat range, multiples of 4
    float beta1, beta2, eps, grad_scale;
    uint32_t world, rank, slot0;   // slot0: first flag / epoch slot of this call's channel
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// (ptxas emits LDGMC...STRONG.SYS for every semantics of multimem.ld_reduce, and a thread gets such a reduction back
// every ~2.6 us: on 32 x 512 threads a 25 MB slice took 255 us, and twice that on half the threads, whatever the
// unrolling - hence the wide launch below.)
__device__ __forceinline__ float4 multimem_ld_reduce_f4(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.weak.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc));       // no memory clobber: loads may be hoisted together
    return v;
}
__device__ __forceinline__ void multimem_st_f4(float* mc, float4 v) {
    asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// All threads of the block call it.  Everything this block (and, through the preceding __syncthreads / fences, this rank's
// earlier kernels) wrote is visible to the peers' blocks of the same index once they return, and vice versa.
__device__ __forceinline__ void rank_barrier(const SymmAdamParams& p, uint32_t& epoch_reg) {
    __threadfence_system();
    __syncthreads();
    epoch_reg += 1u;
    if (threadIdx.x < p.world) {
        const uint32_t peer = threadIdx.x;
        st_release_sys(p.flags_peer[peer] + (p.slot0 + blockIdx.x) * symm::kMaxWorld + p.rank, epoch_reg);
        const uint32_t* mine = p.flags_peer[p.rank] + (p.slot0 + blockIdx.x) * symm::kMaxWorld + peer;
        const long long t0 = clock64();
        while ((int32_t)(ld_acquire_sys(mine) - epoch_reg) < 0) {
            if (clock64() - t0 > 8000000000ll) {          // ~4 s at 1.9 GHz: a peer never arrived
                *p.error = 1u;
                break;
            }
        }
    }
    __syncthreads();
}

template <uint32_t WORLD>
__global__ void __launch_bounds__(symm::kThreads, 1) symm_adam_kernel(const SymmAdamParams p) {
    pdl_begin();
    if (p.gate != nullptr && *p.gate == 0) return;          // same value on every rank: all skip together
    uint32_t epoch_reg = p.epoch[p.slot0 + blockIdx.x];
    rank_barrier(p, epoch_reg);                              // every rank's gradient is complete

    const uint64_t n4 = (p.stop - p.start) / 4;
    const uint64_t per = (n4 + p.world - 1) / p.world;
    const uint64_t lo = p.start / 4 + (uint64_t)p.rank * per;
    const uint64_t hi4 = p.start / 4 + n4;
    const uint64_t hi = (lo + per < hi4) ? lo + per : hi4;
    const float lr = __ldg(p.dyn), bc1 = __ldg(p.dyn + 1), bc2 = __ldg(p.dyn + 2), ema_w = __ldg(p.dyn + 3);
    const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
        gg *= p.grad_scale;
        mm = p.beta1 * mm + (1.0f - p.beta1) * gg;
        vv = p.beta2 * vv + (1.0f - p.beta2) * gg * gg;
        pp -= step_size * mm / (sqrtf(vv) * inv_sqrt_bc2 + p.eps);
    };
    // Remote reads have a latency of microseconds: every load of a trip is issued before the first use (kU pieces x
    // WORLD peers of plain loads, or kU in-switch reductions), and the kernel is launched WIDE and SHORT (one block of up
    // to 1024 threads per SM) rather than narrow and long: ptxas emits LDGMC...STRONG.SYS for multimem.ld_reduce and a
    // thread gets one of them back every ~2.6 us, so the slice time is (elements per thread) x 2.6 us.
    constexpr uint32_t kU = 2;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i0 = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += kU * stride) {
        float4 g[kU], pp[kU], mm[kU], vv[kU], ee[kU];
        if (p.grad_mc != nullptr) {
#pragma unroll
            for (uint32_t u = 0; u < kU; ++u) {
                const uint64_t i = i0 + u * stride;
                g[u] = (i < hi) ? multimem_ld_reduce_f4(p.grad_mc + 4 * i) : zero;
            }
        } else {
            float4 t[kU][WORLD];
#pragma unroll
            for (uint32_t u = 0; u < kU; ++u) {
                const uint64_t i = i0 + u * stride;
#pragma unroll
                for (uint32_t w = 0; w < WORLD; ++w)
                    t[u][w] = (i < hi) ? __ldcs(reinterpret_cast<const float4*>(p.grad_peer[w]) + i) : zero;
            }
#pragma unroll
            for (uint32_t u = 0; u < kU; ++u) {
                g[u] = t[u][0];
#pragma unroll
                for (uint32_t w = 1; w < WORLD; ++w) {
                    g[u].x += t[u][w].x; g[u].y += t[u][w].y; g[u].z += t[u][w].z; g[u].w += t[u][w].w;
                }
            }
        }
#pragma unroll
        for (uint32_t u = 0; u < kU; ++u) {
            const uint64_t i = i0 + u * stride;
            if (i >= hi) break;
            pp[u] = __ldcs(reinterpret_cast<const float4*>(p.param) + i);
            mm[u] = __ldcs(reinterpret_cast<const float4*>(p.exp_avg) + i);
            vv[u] = __ldcs(reinterpret_cast<const float4*>(p.exp_avg_sq) + i);
            if (p.ema != nullptr) ee[u] = __ldcs(reinterpret_cast<const float4*>(p.ema) + i);
        }
#pragma unroll
        for (uint32_t u = 0; u < kU; ++u) {
            const uint64_t i = i0 + u * stride;
            if (i >= hi) break;
            upd(pp[u].x, g[u].x, mm[u].x, vv[u].x); upd(pp[u].y, g[u].y, mm[u].y, vv[u].y);
            upd(pp[u].z, g[u].z, mm[u].z, vv[u].z); upd(pp[u].w, g[u].w, mm[u].w, vv[u].w);
            __stcs(reinterpret_cast<float4*>(p.exp_avg) + i, mm[u]);
            __stcs(reinterpret_cast<float4*>(p.exp_avg_sq) + i, vv[u]);
            if (p.ema != nullptr) {
                float4 e = ee[u];
                e.x -= ema_w * (e.x - pp[u].x); e.y -= ema_w * (e.y - pp[u].y);
                e.z -= ema_w * (e.z - pp[u].z); e.w -= ema_w * (e.w - pp[u].w);
                __stcs(reinterpret_cast<float4*>(p.ema) + i, e);
            }
            if (p.param_mc != nullptr) {
                multimem_st_f4(p.param_mc + 4 * i, pp[u]);
            } else {
#pragma unroll
                for (uint32_t w = 0; w < WORLD; ++w) reinterpret_cast<float4*>(p.param_peer[w])[i] = pp[u];
            }
        }
    }
    rank_barrier(p, epoch_reg);                              // every rank holds the new parameters; nobody reads gradients any more
    if (threadIdx.x == 0) p.epoch[p.slot0 + blockIdx.x] = epoch_reg;
    // every slice of this rank's gradient has been consumed by its owner: clear the whole local range (local stores; sending
    // zeros to the peers instead would cost the link as many bytes as the parameters)
    for (uint64_t i = p.start / 4 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi4; i += stride)
        __stcs(reinterpret_cast<float4*>(p.grad) + i, zero);
}

}  // namespace sanerf

using namespace sanerf;

extern "C" int sanerf_symm_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, float* ema, void* param_mc,
                                     void* grad_mc, const uint64_t* param_peers, const uint64_t* grad_peers,
                                     const uint64_t* flag_peers, uint32_t* epoch, uint32_t* error, uint64_t start, uint64_t stop,
                                     uint32_t world, uint32_t rank, const float* dyn, float beta1, float beta2, float eps,
                                     float grad_scale, const int32_t* gate, uint32_t blocks, uint32_t threads, uint32_t channel, void* stream) {
    if (stop <= start) return SANERF_OK;
    SANERF_REQUIRE_PTR(param); SANERF_REQUIRE_PTR(grad); SANERF_REQUIRE_PTR(exp_avg); SANERF_REQUIRE_PTR(exp_avg_sq);
    SANERF_REQUIRE_PTR(flag_peers); SANERF_REQUIRE_PTR(epoch); SANERF_REQUIRE_PTR(error); SANERF_REQUIRE_PTR(dyn);
    if (world < 2 || world > symm::kMaxWorld || rank >= world) return fail(SANERF_ERR_INVALID_ARG, "symm_adam: world must be 2, 4 or 8");
    if ((start | stop) & 3u) return fail(SANERF_ERR_MISALIGNED, "symm_adam: range bounds must be multiples of 4 elements");
    if ((param_mc == nullptr && param_peers == nullptr) || (grad_mc == nullptr && grad_peers == nullptr))
        return fail(SANERF_ERR_INVALID_ARG, "symm_adam: each buffer needs its multicast address or its peer addresses");
    if (channel >= symm::kChannels || blocks == 0 || blocks > symm::channel_blocks(channel))
        return fail(SANERF_ERR_INVALID_ARG, "symm_adam: channel 0 (<= 160 blocks), 1 or 2 (<= 32 blocks)");
    if (threads == 0 || threads > symm::kThreads || (threads & 31u)) return fail(SANERF_ERR_INVALID_ARG, "symm_adam: threads 32..1024");
    SymmAdamParams p{};
    p.param = param; p.grad = grad; p.exp_avg = exp_avg; p.exp_avg_sq = exp_avg_sq; p.ema = ema;
    p.param_mc = static_cast<float*>(param_mc); p.grad_mc = static_cast<float*>(grad_mc);
    for (uint32_t w = 0; w < world; ++w) {
        p.param_peer[w] = param_peers ? reinterpret_cast<float*>(param_peers[w]) : nullptr;
        p.grad_peer[w] = grad_peers ? reinterpret_cast<float*>(grad_peers[w]) : nullptr;
        p.flags_peer[w] = reinterpret_cast<uint32_t*>(flag_peers[w]);
    }
    p.epoch = epoch; p.error = error; p.dyn = dyn; p.gate = gate; p.start = start; p.stop = stop;
    p.beta1 = beta1; p.beta2 = beta2; p.eps = eps; p.grad_scale = grad_scale; p.world = world; p.rank = rank;
    p.slot0 = symm::channel_slot0(channel);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (world) {
        case 2: SANERF_LAUNCH(symm_adam_kernel<2>, blocks, threads, 0, st, p); break;
        case 4: SANERF_LAUNCH(symm_adam_kernel<4>, blocks, threads, 0, st, p); break;
        case 8: SANERF_LAUNCH(symm_adam_kernel<8>, blocks, threads, 0, st, p); break;
        default: return fail(SANERF_ERR_INVALID_ARG, "symm_adam: world must be 2, 4 or 8");
    }
    return check_launch("symm_adam_kernel");
}
