// ecb_scan.cuh — device-wide exclusive prefix sum over uint32 (three launches: block sums, scan of
// the block sums by one CTA, block-local scan + offset).  Used for EC ranking (popcounts of the
// first-occurrence bitmap), CSR row offsets and stream compaction.
#pragma once
#include "ecb_common.cuh"

#define SCAN_THREADS 256
#define SCAN_ITEMS 16
#define SCAN_BLOCK (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ u32 warp_incl_scan_u32(u32 v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    u32 o = __shfl_up_sync(ECB_FULL, v, d);
    if (lane >= d) v += o;
  }
  return v;
}

__device__ __forceinline__ u64 warp_incl_scan_u64(u64 v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    u64 o = __shfl_up_sync(ECB_FULL, v, d);
    if (lane >= d) v += o;
  }
  return v;
}

// Exclusive scan of one value per thread across a 256-thread CTA; returns the exclusive prefix and
// the CTA total.  `smem` needs 9 u32.
__device__ __forceinline__ u32 block_excl_scan_u32(u32 v, u32* smem, u32& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u32 inc = warp_incl_scan_u32(v, lane);
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    u32 w = lane < (SCAN_THREADS / 32) ? smem[lane] : 0u;
    u32 winc = warp_incl_scan_u32(w, lane);
    if (lane < (SCAN_THREADS / 32)) smem[lane] = winc - w;
    if (lane == (SCAN_THREADS / 32) - 1) smem[8] = winc;
  }
  __syncthreads();
  u32 res = smem[warp] + inc - v;
  total = smem[8];
  __syncthreads();
  return res;
}

// in may be a u32 array (POPC=false) or a bitmap whose popcounts are scanned (POPC=true).
template <bool POPC>
__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums_kernel(const u32* __restrict__ in, u64 n,
                                                                        u64* __restrict__ block_sums) {
  __shared__ u32 sm[9];
  const u64 base = (u64)blockIdx.x * SCAN_BLOCK;
  u32 s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    u64 idx = base + (u64)i * SCAN_THREADS + threadIdx.x;
    if (idx < n) {
      u32 v = in[idx];
      s += POPC ? (u32)__popc(v) : v;
    }
  }
  u32 total;
  block_excl_scan_u32(s, sm, total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// One CTA turns block_sums into exclusive offsets in place; total -> *total_out.
__global__ void __launch_bounds__(1024) scan_partials_kernel(u64* __restrict__ block_sums, u32 n_blocks,
                                                             u64* __restrict__ total_out) {
  __shared__ u64 warp_tot[32];
  __shared__ u64 carry_s;
  __shared__ u64 chunk_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (u32 start = 0; start < n_blocks; start += 1024) {
    u32 idx = start + threadIdx.x;
    u64 v = idx < n_blocks ? block_sums[idx] : 0ull;
    u64 inc = warp_incl_scan_u64(v, lane);
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      u64 w = warp_tot[lane];
      u64 winc = warp_incl_scan_u64(w, lane);
      warp_tot[lane] = winc - w;
      if (lane == 31) chunk_total = winc;
    }
    __syncthreads();
    u64 excl = carry_s + warp_tot[warp] + inc - v;
    if (idx < n_blocks) block_sums[idx] = excl;
    __syncthreads();
    if (threadIdx.x == 0) carry_s += chunk_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = carry_s;
}

template <bool POPC>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const u32* __restrict__ in, u64 n,
                                                                   const u64* __restrict__ block_offsets,
                                                                   u32* __restrict__ out, u32 base_offset) {
  __shared__ u32 sm[9];
  // blocked arrangement: thread t owns SCAN_ITEMS consecutive elements
  const u64 base = (u64)blockIdx.x * SCAN_BLOCK + (u64)threadIdx.x * SCAN_ITEMS;
  u32 v[SCAN_ITEMS];
  u32 s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    u64 idx = base + i;
    u32 x = 0;
    if (idx < n) {
      x = in[idx];
      if (POPC) x = (u32)__popc(x);
    }
    v[i] = x;
    s += x;
  }
  u32 total;
  u32 excl = block_excl_scan_u32(s, sm, total);
  u32 run = base_offset + (u32)block_offsets[blockIdx.x] + excl;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    u64 idx = base + i;
    if (idx < n) out[idx] = run;
    run += v[i];
  }
}
