// ecb_sort.cuh — stable LSD radix sort of (u64 key, u32 value) pairs, 8 bits per pass.
//
// Used only by the per-cell (multisample) finalisation: ordering the cells by their nested
// first-occurrence key (alntools/bam_utils_multisample.py:513-551) and turning the (EC, cell) count
// pairs into the CSC N matrix (bam_utils_multisample.py:783-791: csr -> tocsc(), EC ids ascending
// inside a column).  Each pass: per-tile digit histogram -> device scan -> stable scatter.  Inside a
// tile every warp owns 256 consecutive keys and ranks them with __match_any_sync.
#pragma once
#include "ecb_common.cuh"

#define SORT_THREADS 256
#define SORT_ITEMS 8
#define SORT_TILE (SORT_THREADS * SORT_ITEMS)

__global__ void __launch_bounds__(SORT_THREADS) sort_hist_kernel(const u64* __restrict__ keys, u32 n, int shift,
                                                                  u32* __restrict__ hist, u32 n_tiles) {
  __shared__ u32 bins[256];
  bins[threadIdx.x] = 0;
  __syncthreads();
  const u32 base = blockIdx.x * SORT_TILE;
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; ++i) {
    const u32 idx = base + i * SORT_THREADS + threadIdx.x;
    if (idx < n) atomicAdd(&bins[(u32)(keys[idx] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * n_tiles + blockIdx.x] = bins[threadIdx.x];  // digit-major
}

__global__ void __launch_bounds__(SORT_THREADS) sort_scatter_kernel(const u64* __restrict__ keys_in,
                                                                     const u32* __restrict__ vals_in, u32 n, int shift,
                                                                     const u32* __restrict__ offsets, u32 n_tiles,
                                                                     u64* __restrict__ keys_out,
                                                                     u32* __restrict__ vals_out) {
  __shared__ u32 warp_cnt[SORT_THREADS / 32][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (SORT_THREADS / 32) * 256; i += SORT_THREADS) (&warp_cnt[0][0])[i] = 0;
  __syncthreads();
  const u32 seg = blockIdx.x * SORT_TILE + warp * (32 * SORT_ITEMS);  // this warp's 256 consecutive keys
  u64 k[SORT_ITEMS];
  u32 rank[SORT_ITEMS];
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; ++i) {
    const u32 idx = seg + i * 32 + lane;
    const bool in = idx < n;
    k[i] = in ? keys_in[idx] : ~0ull;
    const u32 d = in ? ((u32)(k[i] >> shift) & 255u) : 256u + lane;  // out-of-range lanes match nobody
    const u32 peers = __match_any_sync(ECB_FULL, d);
    u32 before = 0;
    if (in) before = warp_cnt[warp][d];
    __syncwarp();
    if (in && lane == __ffs(peers) - 1) warp_cnt[warp][d] = before + __popc(peers);
    __syncwarp();
    rank[i] = before + __popc(peers & ((1u << lane) - 1u));
  }
  __syncthreads();
  {  // thread d: exclusive prefix of digit d over the warps, plus the tile's global offset
    const u32 d = threadIdx.x;
    u32 run = offsets[(size_t)d * n_tiles + blockIdx.x];
#pragma unroll
    for (int w = 0; w < SORT_THREADS / 32; ++w) {
      const u32 cnt = warp_cnt[w][d];
      warp_cnt[w][d] = run;
      run += cnt;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; ++i) {
    const u32 idx = seg + i * 32 + lane;
    if (idx < n) {
      const u32 d = (u32)(k[i] >> shift) & 255u;
      const u32 dst = warp_cnt[warp][d] + rank[i];
      keys_out[dst] = k[i];
      vals_out[dst] = vals_in[idx];
    }
  }
}
