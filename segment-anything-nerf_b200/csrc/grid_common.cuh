// Level geometry, cell location and corner indexing of the multiresolution grid.
//
// Semantics follow the reference kernels (gridencoder/src/gridencoder.cu:45-79 hash/index,
// :132-133 level size + resolution, :140-160 position) and must stay BIT-EXACT with them for
// every integer quantity: resolution, dense-vs-hash decision, uint32 wrap-around of strides
// and hashes, and the final `% level_rows`.
#pragma once

#include "common.cuh"

namespace sanerf {

// Spatial-hash multipliers (gridencoder.cu:49). Facts of the format, not tunables.
__device__ __constant__ const uint32_t kHashPrimes[7] = {1u,          2654435761u, 805459861u,
                                                         3674653429u, 2097192037u, 1434869437u,
                                                         2165219737u};

template <uint32_t D>
struct LevelGeom {
    uint32_t res;        // grid resolution of this level (device fp32 expression!)
    uint32_t rows;       // rows in this level's table slice ("hashmap_size")
    uint32_t mult[D];    // per-dimension multiplier: dense stride (0 if the stride loop stopped
                         // before this dimension) or hash prime
    uint32_t covered;    // number of dimensions the dense stride loop covered
    bool hashed;         // XOR-hash instead of strided sum
    bool wrap;           // index may exceed rows -> needs `% rows`
    bool pow2;           // rows is a power of two -> `& (rows-1)`
};

// gridencoder.cu:132-133 and the stride loop of get_grid_index (:63-76), hoisted per level.
template <uint32_t D>
__device__ __forceinline__ LevelGeom<D> level_geometry(const int32_t* __restrict__ offsets,
                                                       uint32_t level, float S, uint32_t H,
                                                       uint32_t gridtype) {
    LevelGeom<D> g;
    g.rows = (uint32_t)(__ldg(offsets + level + 1) - __ldg(offsets + level));
    // Same fp32 expression as the reference, evaluated on the device (MUFU.EX2 path); the host
    // computes table sizes in fp64 and can disagree by one.
    g.res = (uint32_t)ceilf(exp2f((float)level * S) * (float)H);

    uint32_t stride = 1;
    uint64_t wide = 1;  // same product without the uint32 wrap, to know when `%` is a no-op
    uint32_t covered = 0;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
        const bool take = (covered == d) && (stride <= g.rows);
        g.mult[d] = take ? stride : 0u;
        if (take) {
            stride *= g.res;  // uint32 wrap-around is part of the format
            wide = (wide > 0xffffffffull) ? wide : wide * (uint64_t)g.res;
            covered = d + 1;
        }
    }
    g.covered = covered;
    g.hashed = (gridtype == 0u) && (stride > g.rows);
    // index < rows is guaranteed only if every dimension was covered without overflow
    g.wrap = (stride > g.rows) || (wide != (uint64_t)stride);
    if (g.hashed) {
#pragma unroll
        for (uint32_t d = 0; d < D; ++d) g.mult[d] = kHashPrimes[d];
    }
    g.pow2 = (g.rows & (g.rows - 1u)) == 0u;
    return g;
}

template <uint32_t D>
__device__ __forceinline__ uint32_t finish_index(const LevelGeom<D>& g, uint32_t index) {
    if (g.wrap) index = g.pow2 ? (index & (g.rows - 1u)) : (index % g.rows);
    return index;
}

// Cell location of one sample on one level (gridencoder.cu:140-160).
template <uint32_t D>
struct Cell {
    uint32_t lo[D];   // contribution of the lower corner in each dim: pg[d] * mult[d]
    uint32_t hi[D];   // contribution of the upper corner: min(pg[d]+1, res-1) * mult[d]
    float f[D];       // interpolation weight of the upper corner (after optional smoothstep)
    float df[D];      // d f / d pos (1 for linear, 6f(1-f) for smoothstep)
};

template <uint32_t D>
__device__ __forceinline__ Cell<D> locate(const LevelGeom<D>& g, const float (&x)[D],
                                          bool align_corners, uint32_t interp) {
    Cell<D> c;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
        float pos;
        uint32_t pg;
        if (align_corners) {
            pos = x[d] * (float)(g.res - 1u);
            pg = min((uint32_t)floorf(pos), g.res - 2u);
        } else {
            // nvcc contracts `x*res - 0.5f` of the reference into one FMA; spell it out.
            pos = fminf(fmaxf(__fmaf_rn(x[d], (float)g.res, -0.5f), 0.0f), (float)(g.res - 1u));
            pg = (uint32_t)floorf(pos);
        }
        float f = pos - (float)pg;
        float df = 1.0f;
        if (interp == 1u) {
            df = (6.0f * f) * (1.0f - f);
            f = (f * f) * __fmaf_rn(-2.0f, f, 3.0f);
        }
        c.f[d] = f;
        c.df[d] = df;
        c.lo[d] = pg * g.mult[d];
        c.hi[d] = min(pg + 1u, g.res - 1u) * g.mult[d];
    }
    return c;
}

// Row (relative to the level's first row) and weight of corner `corner` (bit d = upper in dim d).
template <uint32_t D>
__device__ __forceinline__ uint32_t corner_row(const LevelGeom<D>& g, const Cell<D>& c,
                                               uint32_t corner) {
    uint32_t index = 0;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
        const uint32_t term = (corner & (1u << d)) ? c.hi[d] : c.lo[d];
        index = g.hashed ? (index ^ term) : (index + term);
    }
    return finish_index(g, index);
}

// Product in the reference's order: ((1 * a0) * a1) * a2 ...  (gridencoder.cu:172-184)
template <uint32_t D>
__device__ __forceinline__ float corner_weight(const Cell<D>& c, uint32_t corner) {
    float w = 1.0f;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) w *= (corner & (1u << d)) ? c.f[d] : (1.0f - c.f[d]);
    return w;
}

// Warp-aggregated scatter.  Samples are ray-ordered, so on coarse levels several CONSECUTIVE lanes of a warp fall
// into the same grid cell and would each fire the same 2^D reductions at the same addresses (the L2 atomic unit
// serialises them: level 0 of a proposal grid costs 4x more with ray-ordered than with shuffled samples).  This is a
// segmented suffix-reduction over runs of consecutive lanes with identical cell keys: afterwards the first lane of
// each run ("head", return value) holds the run's sums in v[] and is the only one that issues reductions.
// key[] = the per-dimension lower-corner contributions (Cell::lo): equal keys <=> equal rows for every corner.
// Must be called by all 32 lanes.
template <uint32_t NV, uint32_t D>
__device__ __forceinline__ bool warp_run_reduce(float (&v)[NV], const uint32_t (&key)[D], uint32_t lane) {
    constexpr uint32_t kFull = 0xffffffffu;
    bool head = (lane == 0u);
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) head |= (__shfl_up_sync(kFull, key[d], 1) != key[d]);
    const uint32_t heads = __ballot_sync(kFull, head);
    if (heads == kFull) return true;                                   // every run has length 1: nothing to merge
    const uint32_t later = (lane == 31u) ? 0u : (heads >> (lane + 1u));    // bit i: lane + 1 + i starts a new run
#pragma unroll
    for (uint32_t step = 1; step < 32u; step <<= 1) {
        const bool same = (lane + step < 32u) && ((later & ((1u << step) - 1u)) == 0u);
        if (__ballot_sync(kFull, same) == 0u) break;                   // longest run already folded
#pragma unroll
        for (uint32_t i = 0; i < NV; ++i) {
            const float t = __shfl_down_sync(kFull, v[i], step);
            if (same) v[i] += t;
        }
    }
    return head;
}

// Pair-merged gather for fp32 tables with two features per row: the two corners along the first dimension are adjacent
// rows whenever the lower one is even (dense levels: stride 1; hashed levels: prime 1, so x ^ h and (x + 1) ^ h differ in
// bit 0 only); both then sit in ONE aligned 16-byte slot and come back from one 128-bit load.  Random 8-byte gathers are
// bound by L1 wavefronts (one 128-byte line per lane and instruction), so every merged pair saves a quarter of a level's
// cost.  `slice` = first row of the level (levels start at multiples of 8 rows, grid.py:130, so slots stay aligned).
// Values and their use order are unchanged: results stay bit-identical.
__device__ __forceinline__ void gather_pair_f2(const float* __restrict__ slice, uint32_t r0, uint32_t r1, float2& v0,
                                               float2& v1) {
    if ((r0 ^ r1) == 1u) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(slice + (size_t)(r0 & ~1u) * 2));
        const bool lower = (r0 & 1u) == 0u;
        v0 = lower ? make_float2(t.x, t.y) : make_float2(t.z, t.w);
        v1 = lower ? make_float2(t.z, t.w) : make_float2(t.x, t.y);
    } else {
        v0 = __ldg(reinterpret_cast<const float2*>(slice + (size_t)r0 * 2));
        v1 = __ldg(reinterpret_cast<const float2*>(slice + (size_t)r1 * 2));
    }
}

// Inclusive range test of the reference (gridencoder.cu:109): NaN passes.
template <uint32_t D>
__device__ __forceinline__ bool out_of_range(const float (&x)[D]) {
    bool oob = false;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) oob |= (x[d] < 0.0f) || (x[d] > 1.0f);
    return oob;
}

// ---- table row loads / stores of C elements as the widest vectors -------------------------
template <typename T, uint32_t C>
struct RowIO;

template <uint32_t C>
struct RowIO<float, C> {
    static __device__ __forceinline__ void load(const float* __restrict__ p, float (&v)[C]) {
        if constexpr (C == 1) {
            v[0] = __ldg(p);
        } else if constexpr (C == 2) {
            float2 t = __ldg(reinterpret_cast<const float2*>(p));
            v[0] = t.x; v[1] = t.y;
        } else {
#pragma unroll
            for (uint32_t i = 0; i < C; i += 4) {
                float4 t = __ldg(reinterpret_cast<const float4*>(p + i));
                v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
            }
        }
    }
    static __device__ __forceinline__ void store(float* __restrict__ p, const float (&v)[C]) {
        if constexpr (C == 1) {
            p[0] = v[0];
        } else if constexpr (C == 2) {
            *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
        } else {
#pragma unroll
            for (uint32_t i = 0; i < C; i += 4)
                *reinterpret_cast<float4*>(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
    }
    // grad_table[row] += v  (vector reductions, no return value)
    static __device__ __forceinline__ void red(float* __restrict__ p, const float (&v)[C]) {
        if constexpr (C == 1) {
            red_add_f32(p, v[0]);
        } else if constexpr (C == 2) {
            red_add_v2_f32(p, v[0], v[1]);
        } else {
#pragma unroll
            for (uint32_t i = 0; i < C; i += 4) red_add_v4_f32(p + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
    }
};

template <uint32_t C>
struct RowIO<__half, C> {
    static __device__ __forceinline__ void load(const __half* __restrict__ p, float (&v)[C]) {
        if constexpr (C == 1) {
            v[0] = __half2float(__ldg(p));
        } else if constexpr (C == 2) {
            float2 t = __half22float2(__ldg(reinterpret_cast<const __half2*>(p)));
            v[0] = t.x; v[1] = t.y;
        } else if constexpr (C == 4) {
            uint2 raw = __ldg(reinterpret_cast<const uint2*>(p));
            float2 a = __half22float2(*reinterpret_cast<__half2*>(&raw.x));
            float2 b = __half22float2(*reinterpret_cast<__half2*>(&raw.y));
            v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
        } else {
#pragma unroll
            for (uint32_t i = 0; i < C; i += 8) {
                uint4 raw = __ldg(reinterpret_cast<const uint4*>(p + i));
                const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j) {
                    float2 t = __half22float2(h[j]);
                    v[i + 2 * j] = t.x; v[i + 2 * j + 1] = t.y;
                }
            }
        }
    }
    static __device__ __forceinline__ void store(__half* __restrict__ p, const float (&v)[C]) {
        if constexpr (C == 1) {
            p[0] = __float2half_rn(v[0]);
        } else {
#pragma unroll
            for (uint32_t i = 0; i < C; i += 2)
                *reinterpret_cast<__half2*>(p + i) = __floats2half2_rn(v[i], v[i + 1]);
        }
    }
    static __device__ __forceinline__ void red(__half* __restrict__ p, const float (&v)[C]) {
        if constexpr (C == 1) {
            atomicAdd(p, __float2half_rn(v[0]));
        } else if constexpr (C == 2) {
            red_add_f16x2(p, __floats2half2_rn(v[0], v[1]));
        } else if constexpr (C == 4) {
            red_add_v2_f16x2(p, __floats2half2_rn(v[0], v[1]), __floats2half2_rn(v[2], v[3]));
        } else {
#pragma unroll
            for (uint32_t i = 0; i < C; i += 8)
                red_add_v4_f16x2(p + i, __floats2half2_rn(v[i], v[i + 1]),
                                 __floats2half2_rn(v[i + 2], v[i + 3]),
                                 __floats2half2_rn(v[i + 4], v[i + 5]),
                                 __floats2half2_rn(v[i + 6], v[i + 7]));
        }
    }
};

template <typename T> __device__ __forceinline__ float to_float(T v);
template <> __device__ __forceinline__ float to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_float<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_float<__half>(float v) { return __float2half_rn(v); }

}  // namespace sanerf
