// group_plan.cuh — `.over()` / group_by key columns -> CSR groups, on the device.
//
// Replaces, for the batched route, what polars does in front of the reference's plugin (SURVEY.md §8 a3:
// GroupsProxy construction + the per-group gather feeding src/expressions.rs:22-103):
//   keys[n_keys][N]  ->  n_groups, group_offsets[G+1], row_index[N] (packed position -> original row,
//                        stable: rows of a group keep the frame's order, groups ascend by key tuple),
//                        group_first_row[G] (one representative row per group: its key values)
// The result is bit-identical to `np.unique(keys, return_inverse=True)` + `np.argsort(inv, kind="stable")`
// (the host plan it replaces, and what tests/test_ols.py:39-40,969-995 of the reference exercise: random,
// non-contiguous, unequal groups).
//
// Pipeline (all kernels HBM-bound integer work, grid = tiles of 4096 keys):
//   plan_prepass      order-preserving 64-bit radix image of the sort key, OR / AND of all images (the bits
//                     that vary decide which radix passes run), lexicographic "already sorted" flag
//   [per varying 8-bit digit, LSD]  radix_hist -> exclusive scan -> radix_scatter (stable: warp-level
//                     match_any ranking, per-warp digit counters in shared memory)
//   plan_flags        boundary flag per sorted position (any key differs from the previous row)
//   plan_count -> exclusive scan -> plan_write   group ids -> offsets / first rows / int64 row_index
// Sorted keys (GroupsSlice, the C2 bench shape) skip the sort: offsets only, row_index = NULL.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace b200 {

enum : int { KEY_I64 = 0, KEY_I32 = 1, KEY_U64 = 2, KEY_U32 = 3, KEY_F64 = 4, KEY_F32 = 5 };
constexpr int PLAN_MAX_KEYS = 8;
constexpr int PLAN_THREADS = 256;
constexpr int PLAN_ITEMS = 16;
constexpr int PLAN_TILE = PLAN_THREADS * PLAN_ITEMS;  // keys per block

struct PlanKeys {
    const void *col[PLAN_MAX_KEYS];
    int dtype[PLAN_MAX_KEYS];
    int n_keys;
};

// order-preserving image of a key in uint64: a < b  <=>  image(a) < image(b); equal keys <=> equal images
// (floats: -0.0 == +0.0 and every NaN is one key sorted last, as numpy.unique does)
__device__ __forceinline__ uint64_t key_image_f64(double v) {
    if (v != v) return ~0ull;
    if (v == 0.0) v = 0.0;
    const uint64_t b = static_cast<uint64_t>(__double_as_longlong(v));
    return (b >> 63) ? ~b : (b | (1ull << 63));
}

__device__ __forceinline__ uint64_t key_image(const void *col, int dtype, int64_t i) {
    switch (dtype) {
        case KEY_I64: return static_cast<uint64_t>(static_cast<const int64_t *>(col)[i]) ^ (1ull << 63);
        case KEY_I32: return static_cast<uint64_t>(static_cast<int64_t>(static_cast<const int32_t *>(col)[i])) ^ (1ull << 63);
        case KEY_U64: return static_cast<const uint64_t *>(col)[i];
        case KEY_U32: return static_cast<uint64_t>(static_cast<const uint32_t *>(col)[i]);
        case KEY_F64: return key_image_f64(static_cast<const double *>(col)[i]);
        default: return key_image_f64(static_cast<double>(static_cast<const float *>(col)[i]));
    }
}

struct PlanScalars {      // device-side scalars, copied to the host between phases
    unsigned long long key_or, key_and;
    unsigned int unsorted;
    unsigned int pad;
};

// images of key column `which` in the CURRENT order (idx == nullptr: identity) -> img[]; OR / AND over all
// images; with check_sorted: lexicographic comparison of every row with its predecessor over ALL keys
__global__ void __launch_bounds__(PLAN_THREADS) plan_prepass_kernel(const PlanKeys keys, int which, const uint32_t *__restrict__ idx,
                                                                    uint64_t *__restrict__ img, int64_t n, int check_sorted,
                                                                    PlanScalars *sc) {
    uint64_t vor = 0, vand = ~0ull;
    bool unsorted = false;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t src = idx ? static_cast<int64_t>(idx[i]) : i;
        const uint64_t k = key_image(keys.col[which], keys.dtype[which], src);
        img[i] = k;
        vor |= k;
        vand &= k;
        if (check_sorted && i > 0) {
            for (int j = 0; j < keys.n_keys; ++j) {
                const uint64_t a = key_image(keys.col[j], keys.dtype[j], i - 1), b = key_image(keys.col[j], keys.dtype[j], i);
                if (a != b) {
                    unsorted = unsorted || (b < a);
                    break;
                }
            }
        }
    }
    vor = __reduce_or_sync(0xffffffffu, static_cast<unsigned>(vor)) | (static_cast<uint64_t>(__reduce_or_sync(0xffffffffu, static_cast<unsigned>(vor >> 32))) << 32);
    vand = __reduce_and_sync(0xffffffffu, static_cast<unsigned>(vand)) | (static_cast<uint64_t>(__reduce_and_sync(0xffffffffu, static_cast<unsigned>(vand >> 32))) << 32);
    const unsigned any_unsorted = __ballot_sync(0xffffffffu, unsorted);
    if ((threadIdx.x & 31) == 0) {
        atomicOr(&sc->key_or, static_cast<unsigned long long>(vor));
        atomicAnd(&sc->key_and, static_cast<unsigned long long>(vand));
        if (any_unsorted) atomicOr(&sc->unsorted, 1u);
    }
}

// ---- exclusive scan of uint32 arrays (three small launches: segment sums, scan of the sums, apply) --------
constexpr int SCAN_SEG = 4096;  // elements per block

__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t *warp_sums, uint32_t *total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t s = lane < (PLAN_THREADS / 32) ? warp_sums[lane] : 0u;
        uint32_t si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= o) si += t;
        }
        if (lane < (PLAN_THREADS / 32)) warp_sums[lane] = si - s;
        if (lane == 31) *total = si;
    }
    __syncthreads();
    return warp_sums[w] + inc - v;
}

__global__ void __launch_bounds__(PLAN_THREADS) scan_seg_sum_kernel(const uint32_t *__restrict__ a, int64_t m, uint32_t *__restrict__ seg_sum) {
    __shared__ uint32_t ws[PLAN_THREADS / 32];
    const int64_t base = static_cast<int64_t>(blockIdx.x) * SCAN_SEG;
    uint32_t s = 0;
    for (int t = threadIdx.x; t < SCAN_SEG; t += PLAN_THREADS)
        if (base + t < m) s += a[base + t];
    s = __reduce_add_sync(0xffffffffu, s);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < PLAN_THREADS / 32; ++w) t += ws[w];
        seg_sum[blockIdx.x] = t;
    }
}

// one block: exclusive scan of seg_sum[0..ns) in place, grand total -> *total
__global__ void __launch_bounds__(PLAN_THREADS) scan_seg_scan_kernel(uint32_t *seg_sum, int64_t ns, uint32_t *total) {
    __shared__ uint32_t ws[PLAN_THREADS / 32];
    __shared__ uint32_t tile_total;
    uint32_t carry = 0;
    for (int64_t base = 0; base < ns; base += PLAN_THREADS) {
        const int64_t i = base + threadIdx.x;
        const uint32_t v = i < ns ? seg_sum[i] : 0u;
        const uint32_t ex = block_exclusive_scan_256(v, ws, &tile_total);
        if (i < ns) seg_sum[i] = carry + ex;
        carry += tile_total;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}

__global__ void __launch_bounds__(PLAN_THREADS) scan_seg_apply_kernel(uint32_t *__restrict__ a, int64_t m, const uint32_t *__restrict__ seg_off) {
    __shared__ uint32_t ws[PLAN_THREADS / 32];
    __shared__ uint32_t tile_total;
    const int64_t base = static_cast<int64_t>(blockIdx.x) * SCAN_SEG;
    uint32_t carry = seg_off[blockIdx.x];
    for (int t0 = 0; t0 < SCAN_SEG; t0 += PLAN_THREADS) {
        const int64_t i = base + t0 + threadIdx.x;
        const uint32_t v = i < m ? a[i] : 0u;
        const uint32_t ex = block_exclusive_scan_256(v, ws, &tile_total);
        if (i < m) a[i] = carry + ex;
        carry += tile_total;
        __syncthreads();
    }
}

// ---- LSD radix pass on (image, idx) pairs, 8-bit digit at `shift` ------------------------------------------
// hist[d * nb + b] = number of keys with digit d in tile b
__global__ void __launch_bounds__(PLAN_THREADS) radix_hist_kernel(const uint64_t *__restrict__ img, int64_t n, int shift, uint32_t *__restrict__ hist,
                                                                  int64_t nb) {
    __shared__ uint32_t sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t base = static_cast<int64_t>(blockIdx.x) * PLAN_TILE + static_cast<int64_t>(w) * (32 * PLAN_ITEMS);
#pragma unroll 4
    for (int r = 0; r < PLAN_ITEMS; ++r) {
        const int64_t i = base + r * 32 + lane;
        const unsigned d = i < n ? static_cast<unsigned>((img[i] >> shift) & 255u) : 256u;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (d < 256u && lane == __ffs(peers) - 1) atomicAdd(&sh[d], static_cast<uint32_t>(__popc(peers)));
    }
    __syncthreads();
    hist[static_cast<int64_t>(threadIdx.x) * nb + blockIdx.x] = sh[threadIdx.x];
}

// off = exclusive scan of hist (flattened, digit-major): global position of the first key of (digit, tile).
// Stable: inside a tile keys keep their order (warp sub-tiles in order, rounds in order, lanes in order).
__global__ void __launch_bounds__(PLAN_THREADS) radix_scatter_kernel(const uint64_t *__restrict__ img_in, const uint32_t *__restrict__ idx_in,
                                                                     uint64_t *__restrict__ img_out, uint32_t *__restrict__ idx_out, int64_t n,
                                                                     int shift, const uint32_t *__restrict__ off, int64_t nb) {
    __shared__ uint32_t wcnt[PLAN_THREADS / 32][256];
    for (int t = threadIdx.x; t < (PLAN_THREADS / 32) * 256; t += PLAN_THREADS) (&wcnt[0][0])[t] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t base = static_cast<int64_t>(blockIdx.x) * PLAN_TILE + static_cast<int64_t>(w) * (32 * PLAN_ITEMS);
    uint64_t k[PLAN_ITEMS];
    uint32_t rank[PLAN_ITEMS];
#pragma unroll
    for (int r = 0; r < PLAN_ITEMS; ++r) {
        const int64_t i = base + r * 32 + lane;
        k[r] = i < n ? img_in[i] : 0ull;
        const unsigned d = i < n ? static_cast<unsigned>((k[r] >> shift) & 255u) : 256u;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (d < 256u && lane == leader) {
            old = wcnt[w][d];
            wcnt[w][d] = old + static_cast<uint32_t>(__popc(peers));
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[r] = old + static_cast<uint32_t>(__popc(peers & ((1u << lane) - 1u)));
        __syncwarp();
    }
    __syncthreads();
    {   // thread d: positions of digit d for each warp of this tile
        uint32_t run = off[static_cast<int64_t>(threadIdx.x) * nb + blockIdx.x];
#pragma unroll
        for (int w2 = 0; w2 < PLAN_THREADS / 32; ++w2) {
            const uint32_t t = wcnt[w2][threadIdx.x];
            wcnt[w2][threadIdx.x] = run;
            run += t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < PLAN_ITEMS; ++r) {
        const int64_t i = base + r * 32 + lane;
        if (i < n) {
            const unsigned d = static_cast<unsigned>((k[r] >> shift) & 255u);
            const uint32_t pos = wcnt[w][d] + rank[r];
            img_out[pos] = k[r];
            idx_out[pos] = idx_in ? idx_in[i] : static_cast<uint32_t>(i);
        }
    }
}

// ---- boundaries -> groups ----------------------------------------------------------------------------------
// flag[i] = 1 when position i starts a group: i == 0 or any key differs from position i - 1.
// Single key: `img` is the sorted image array.  Several keys: compared through the original columns.
__global__ void __launch_bounds__(PLAN_THREADS) plan_flags_kernel(const PlanKeys keys, const uint64_t *__restrict__ img, const uint32_t *__restrict__ idx,
                                                                  int64_t n, uint8_t *__restrict__ flag) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        bool f = i == 0;
        if (!f) {
            if (img) {
                f = img[i] != img[i - 1];
            } else {
                const int64_t a = idx ? static_cast<int64_t>(idx[i - 1]) : i - 1, b = idx ? static_cast<int64_t>(idx[i]) : i;
                for (int j = 0; j < keys.n_keys && !f; ++j) f = key_image(keys.col[j], keys.dtype[j], a) != key_image(keys.col[j], keys.dtype[j], b);
            }
        }
        flag[i] = f ? 1 : 0;
    }
}

__global__ void __launch_bounds__(PLAN_THREADS) plan_count_kernel(const uint8_t *__restrict__ flag, int64_t n, uint32_t *__restrict__ cnt) {
    __shared__ uint32_t ws[PLAN_THREADS / 32];
    const int64_t base = static_cast<int64_t>(blockIdx.x) * PLAN_TILE;
    uint32_t s = 0;
    for (int t = threadIdx.x; t < PLAN_TILE; t += PLAN_THREADS)
        if (base + t < n) s += flag[base + t];
    s = __reduce_add_sync(0xffffffffu, s);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < PLAN_THREADS / 32; ++w) t += ws[w];
        cnt[blockIdx.x] = t;
    }
}

// cnt = exclusive scan of the per-tile boundary counts.  Writes offsets[g] (g < G; offsets[G] = n is written by
// block 0), first_row[g], and the int64 permutation row_index[i] = idx[i] (when idx != nullptr).
__global__ void __launch_bounds__(PLAN_THREADS) plan_write_kernel(const uint8_t *__restrict__ flag, const uint32_t *__restrict__ idx, int64_t n,
                                                                  const uint32_t *__restrict__ cnt, int64_t n_groups, int64_t *__restrict__ offsets,
                                                                  int64_t *__restrict__ first_row, int64_t *__restrict__ row_index) {
    __shared__ uint32_t ws[PLAN_THREADS / 32];
    __shared__ uint32_t tile_total;
    const int64_t base = static_cast<int64_t>(blockIdx.x) * PLAN_TILE;
    uint32_t carry = cnt[blockIdx.x];
    if (blockIdx.x == 0 && threadIdx.x == 0) offsets[n_groups] = n;
    for (int t0 = 0; t0 < PLAN_TILE; t0 += PLAN_THREADS) {
        const int64_t i = base + t0 + threadIdx.x;
        const uint32_t f = i < n ? flag[i] : 0u;
        const uint32_t ex = block_exclusive_scan_256(f, ws, &tile_total);
        if (i < n) {
            const int64_t orig = idx ? static_cast<int64_t>(idx[i]) : i;
            if (f) {
                const int64_t g = static_cast<int64_t>(carry) + ex;
                offsets[g] = i;
                first_row[g] = orig;
            }
            if (row_index) row_index[i] = orig;
        }
        carry += tile_total;
        __syncthreads();
    }
}

// group id of every ORIGINAL row (polars broadcasts per-group results back to rows under `.over()`)
__global__ void __launch_bounds__(PLAN_THREADS) plan_group_of_row_kernel(const int64_t *__restrict__ offsets, int64_t n_groups,
                                                                         const int64_t *__restrict__ row_index, int64_t n, int32_t *__restrict__ out) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < n; p += stride) {
        int64_t lo = 0, hi = n_groups;  // largest g with offsets[g] <= p
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (offsets[mid] <= p) lo = mid; else hi = mid;
        }
        out[row_index ? row_index[p] : p] = static_cast<int32_t>(lo);
    }
}

}  // namespace b200
