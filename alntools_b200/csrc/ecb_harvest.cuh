// ecb_harvest.cuh — turn one read of every NEW equivalence class into its canonical row: distinct main
// targets ascending, data = OR of (1 << haplotype).
//
// Replaces alntools/bam_utils.py:788-825 (EC key -> per-haplotype (ec, target) COO lists) and the
// canonicalisation scipy performs in alntools/bin_utils.py:208-211 (sum of 2^h * data[h], tocsr()
// with sorted column indices).  Runs at the end of every push, while the caller's columns are still
// valid, over the ECs claimed in that push only (E rows, not R reads).  The grouping kernel recorded
// for every new EC the offset and the length of one read with that key.  A row's place in the arena is
// reserved when its length is known, with one atomic per warp (short reads) or per row (the rare long
// ones) on the arena cursor: no scan, no host round trip; the host only provides room for one row entry
// per alignment of the push, which is an upper bound.
//   k <= 8        one thread per EC: 8-element sorting network in registers
//   8 < k <= 32   one warp per EC: __match_any_sync / __reduce_or_sync, rank by counting
//   32 < k <= 1024  one warp per EC: warp-synchronous bitonic sort in shared memory (no CTA barrier)
//   k > 1024      one CTA per EC: bitonic sort in shared memory (up to ECB_MAX_READ_ALIGNMENTS)
#pragma once
#include "ecb_common.cuh"
#include "ecb_scan.cuh"

struct HarvestParams {
  const int32_t* rg;
  const int32_t* tg;
  const int32_t* hp;
  int n;
  const u32* ec_rep;
  const u32* ec_len;
  u32 e0, e1;        // provisional ids claimed in this push
  u32* row_len;      // [capacity] out: row length
  u32* row_off;      // [capacity] out: absolute offset inside the arena (reserved here)
  uint2* arena;      // (target, mask) pairs
  int n_targets, n_haps;  // bounds of the column values (checked here, off the streaming path)
  u32* long_list;    // provisional ids whose read has more than 32 alignments
  u32* mid_list;     // provisional ids whose read has 9..32 alignments
  u32* big_list;     // provisional ids whose read has 33..HARVEST_WSORT_MAX alignments
  EcbCounters* ctr;  // scratch[1] = #ECs with 8 < k <= 32, scratch[3] = #ECs with 32 < k <= HARVEST_WSORT_MAX,
                     // n_long = #ECs with more; arena_used = the arena cursor
};

// Reserve `cnt` arena entries for every thread of the (fully converged) 256-thread CTA: ONE atomic per CTA
// (all reservations hit the same cursor; one per warp - 127 k for 4 M ECs - serialise at the L2 and were
// measured to cost more than the scan they replaced).  `smem` needs 10 u32.
__device__ __forceinline__ u32 harvest_block_reserve(EcbCounters* ctr, u32 cnt, u32* smem) {
  u32 total;
  const u32 excl = block_excl_scan_u32(cnt, smem, total);
  if (threadIdx.x == 0) smem[9] = total ? (u32)atomicAdd((unsigned long long*)&ctr->arena_used, (unsigned long long)total) : 0u;
  __syncthreads();
  const u32 base = smem[9];
  __syncthreads();
  return base + excl;
}
// ... for one row, by one lane of a warp (the others get the value too).
__device__ __forceinline__ u32 harvest_row_reserve(EcbCounters* ctr, u32 cnt, int lane) {
  unsigned long long base = 0;
  if (lane == 0 && cnt) base = atomicAdd((unsigned long long*)&ctr->arena_used, (unsigned long long)cnt);
  return (u32)__shfl_sync(ECB_FULL, base, 0);
}

#define HARVEST_LONG_MAX 16384
#define HARVEST_WSORT_MAX 1024   // longest read whose row one warp sorts on its own

#define ECB_CSWAP(a, b)            \
  {                                \
    const u32 lo_ = min(a, b);     \
    const u32 hi_ = max(a, b);     \
    a = lo_;                       \
    b = hi_;                       \
  }

// Append `item` to list[] for the lanes where `take` holds: one atomic per warp.
__device__ __forceinline__ void warp_append(u32* list, u32* counter, bool take, u32 item) {
  const u32 act = __activemask();
  const u32 m = __ballot_sync(act, take);
  if (!m) return;
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(m) - 1;
  u32 base = 0;
  if (lane == leader) base = atomicAdd(counter, (u32)__popc(m));
  base = __shfl_sync(act, base, leader);
  if (take) list[base + (u32)__popc(m & ((1u << lane) - 1u))] = item;
}

__global__ void __launch_bounds__(256) ecb_harvest_short_kernel(const HarvestParams P) {
  // whole CTAs walk the id range together (the reservation below is a CTA-wide step)
  __shared__ u32 s_scan[10];
  const u32 n_new = P.e1 - P.e0;
  const u32 stride = gridDim.x * blockDim.x;
  for (u32 i0 = blockIdx.x * blockDim.x; i0 < n_new; i0 += stride) {
    const u32 i = i0 + threadIdx.x;
    const bool in = i < n_new;
    const u32 e = P.e0 + (in ? i : 0u);
    const u32 k = in ? P.ec_len[e] : 0u;
    warp_append(P.mid_list, &P.ctr->scratch[1], in && k > 8 && k <= 32, e);
    warp_append(P.big_list, &P.ctr->scratch[3], in && k > 32 && k <= HARVEST_WSORT_MAX, e);
    if (in && k > HARVEST_WSORT_MAX) {
      if (k > HARVEST_LONG_MAX) atomicOr(&P.ctr->error, ECB_DEVERR_READ_TOO_LONG);
      P.long_list[atomicAdd(&P.ctr->n_long, 1u)] = e;
    }
    const bool mine = in && k <= 8;
    const int s = (int)(mine ? P.ec_rep[e] : 0u);
    u32 c[8];
    bool bad = false;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      c[j] = 0xFFFFFFFFu;
      if (mine && (u32)j < k) {
        const int t = P.tg[s + j], h = P.hp[s + j];
        bad |= (u32)t >= (u32)P.n_targets || (u32)h >= (u32)P.n_haps;
        c[j] = ecb_code(t, h);
      }
    }
    if (bad) atomicOr(&P.ctr->error, ECB_DEVERR_VALUE_RANGE);
    // Batcher odd-even merge sort network for 8 keys (19 comparators)
    ECB_CSWAP(c[0], c[1]) ECB_CSWAP(c[2], c[3]) ECB_CSWAP(c[4], c[5]) ECB_CSWAP(c[6], c[7])
    ECB_CSWAP(c[0], c[2]) ECB_CSWAP(c[1], c[3]) ECB_CSWAP(c[4], c[6]) ECB_CSWAP(c[5], c[7])
    ECB_CSWAP(c[1], c[2]) ECB_CSWAP(c[5], c[6])
    ECB_CSWAP(c[0], c[4]) ECB_CSWAP(c[1], c[5]) ECB_CSWAP(c[2], c[6]) ECB_CSWAP(c[3], c[7])
    ECB_CSWAP(c[2], c[4]) ECB_CSWAP(c[3], c[5])
    ECB_CSWAP(c[1], c[2]) ECB_CSWAP(c[3], c[4]) ECB_CSWAP(c[5], c[6])
    // row length = distinct targets; then the row goes to the place reserved for it
    u32 cnt = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (c[j] != 0xFFFFFFFFu && (j == 0 || (c[j] >> 5) != (c[j - 1] >> 5))) ++cnt;
    const u32 off = harvest_block_reserve(P.ctr, cnt, s_scan);
    if (!mine) continue;
    uint2* out = P.arena + off;
    u32 w = 0, prev_t = 0xFFFFFFFFu, mask = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (c[j] != 0xFFFFFFFFu) {
        const u32 t = c[j] >> 5;
        if (t != prev_t) {
          if (w) out[w - 1] = make_uint2(prev_t, mask);
          ++w;
          prev_t = t;
          mask = 0;
        }
        mask |= 1u << (c[j] & 31u);
      }
    }
    if (w) out[w - 1] = make_uint2(prev_t, mask);
    P.row_len[e] = cnt;
    P.row_off[e] = off;
  }
}

// Rows of reads with 9..32 alignments (listed by the short kernel): one warp per EC, one alignment
// per lane.
__global__ void __launch_bounds__(256) ecb_harvest_warp_kernel(const HarvestParams P) {
  const u32 n_mid = P.ctr->scratch[1];   // listed by the short kernel (same stream, earlier launch)
  const int lane = threadIdx.x & 31;
  const u32 warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 n_warps = (gridDim.x * blockDim.x) >> 5;
  for (u32 i = warp_global; i < n_mid; i += n_warps) {
    const u32 e = P.mid_list[i];
    const int k = (int)P.ec_len[e];
    const int s = (int)P.ec_rep[e];
    const int t = lane < k ? P.tg[s + lane] : -1 - lane;
    const int h = lane < k ? P.hp[s + lane] : 0;
    if (lane < k && ((u32)t >= (u32)P.n_targets || (u32)h >= (u32)P.n_haps)) atomicOr(&P.ctr->error, ECB_DEVERR_VALUE_RANGE);
    const u32 hbit = lane < k ? (1u << (h & 31)) : 0u;
    const u32 grp = __match_any_sync(ECB_FULL, t);
    const u32 mask = __reduce_or_sync(grp, hbit);
    const bool leader = lane < k && lane == __ffs(grp) - 1;
    const u32 leaders = __ballot_sync(ECB_FULL, leader);
    u32 rank = 0;
    for (int j = 0; j < k; ++j) {
      const int tj = __shfl_sync(ECB_FULL, t, j);
      rank += ((leaders >> j) & 1u) && tj < t;
    }
    const u32 off = harvest_row_reserve(P.ctr, (u32)__popc(leaders), lane);
    if (leader) P.arena[(size_t)off + rank] = make_uint2((u32)t, mask);
    if (lane == 0) {
      P.row_len[e] = (u32)__popc(leaders);
      P.row_off[e] = off;
    }
  }
}

// Rows of reads with 33..HARVEST_WSORT_MAX alignments (heavy multimapping): one WARP per EC.  The
// element codes are sorted with a bitonic network in the warp's own slice of shared memory, so the only
// synchronisation is __syncwarp; then one entry per distinct target.
__global__ void __launch_bounds__(256) ecb_harvest_wsort_kernel(const HarvestParams P) {
  const u32 n_big = P.ctr->scratch[3];
  extern __shared__ u32 sm_all[];
  const int lane = threadIdx.x & 31;
  u32* codes = sm_all + (threadIdx.x >> 5) * HARVEST_WSORT_MAX;
  const u32 warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 n_warps = (gridDim.x * blockDim.x) >> 5;
  const u32 lt = (1u << lane) - 1u;
  for (u32 li = warp_global; li < n_big; li += n_warps) {
    const u32 e = P.big_list[li];
    const int s = (int)P.ec_rep[e];
    const u32 k = P.ec_len[e];
    u32 np2 = 64;
    while (np2 < k) np2 <<= 1;
    bool bad = false;
    for (u32 i = lane; i < np2; i += 32) {
      u32 code = 0xFFFFFFFFu;
      if (i < k) {
        const int t = P.tg[s + i], h = P.hp[s + i];
        bad |= (u32)t >= (u32)P.n_targets || (u32)h >= (u32)P.n_haps;
        code = ecb_code(t, h);
      }
      codes[i] = code;
    }
    if (bad) atomicOr(&P.ctr->error, ECB_DEVERR_VALUE_RANGE);
    __syncwarp();
    for (u32 size = 2; size <= np2; size <<= 1) {
      for (u32 stride = size >> 1; stride > 0; stride >>= 1) {
        for (u32 i = lane; i < (np2 >> 1); i += 32) {
          const u32 lo = 2 * i - (i & (stride - 1));
          const u32 hi = lo + stride;
          const bool up = (lo & size) == 0;
          const u32 a = codes[lo], b = codes[hi];
          if ((a > b) == up) {
            codes[lo] = b;
            codes[hi] = a;
          }
        }
        __syncwarp();
      }
    }
    // distinct targets first (the row's length), then the row at its reserved place
    u32 n_rows = 0;
    for (u32 base = 0; base < k; base += 32) {
      const u32 i = base + lane;
      const bool start = i < k && ((i == 0) || ((codes[i - 1] >> 5) != (codes[i] >> 5)));
      n_rows += (u32)__popc(__ballot_sync(ECB_FULL, start));
    }
    const u32 off = harvest_row_reserve(P.ctr, n_rows, lane);
    uint2* out = P.arena + (size_t)off;
    u32 run = 0;
    for (u32 base = 0; base < k; base += 32) {
      const u32 i = base + lane;
      u32 code = 0xFFFFFFFFu;
      bool start = false;
      if (i < k) {
        code = codes[i];
        start = (i == 0) || ((codes[i - 1] >> 5) != (code >> 5));
      }
      const u32 m = __ballot_sync(ECB_FULL, start);
      if (start) {
        u32 mask = 0;
        for (u32 j = i; j < k && (codes[j] >> 5) == (code >> 5); ++j) mask |= 1u << (codes[j] & 31u);
        out[run + (u32)__popc(m & lt)] = make_uint2(code >> 5, mask);
      }
      run += (u32)__popc(m);
    }
    if (lane == 0) {
      P.row_len[e] = run;
      P.row_off[e] = off;
    }
    __syncwarp();
  }
}

// Rows of long reads (HARVEST_WSORT_MAX+1 .. HARVEST_LONG_MAX alignments): one CTA per EC, bitonic sort of the element
// codes in shared memory, then one entry per distinct target.
__global__ void __launch_bounds__(256) ecb_harvest_long_kernel(const HarvestParams P) {
  const u32 n_long = P.ctr->n_long;
  extern __shared__ u32 sm_codes[];
  __shared__ u32 s_scan[9];
  __shared__ u32 s_run;
  __shared__ u32 s_off;
  for (u32 li = blockIdx.x; li < n_long; li += gridDim.x) {
    const u32 e = P.long_list[li];
    const int s = (int)P.ec_rep[e];
    const u32 k = min(P.ec_len[e], (u32)HARVEST_LONG_MAX);
    u32 np2 = 64;
    while (np2 < k) np2 <<= 1;
    for (u32 i = threadIdx.x; i < np2; i += blockDim.x) {
      u32 code = 0xFFFFFFFFu;
      if (i < k) {
        const int t = P.tg[s + i], h = P.hp[s + i];
        if ((u32)t >= (u32)P.n_targets || (u32)h >= (u32)P.n_haps) atomicOr(&P.ctr->error, ECB_DEVERR_VALUE_RANGE);
        code = ecb_code(t, h);
      }
      sm_codes[i] = code;
    }
    __syncthreads();
    for (u32 size = 2; size <= np2; size <<= 1) {
      for (u32 stride = size >> 1; stride > 0; stride >>= 1) {
        for (u32 i = threadIdx.x; i < (np2 >> 1); i += blockDim.x) {
          const u32 lo = 2 * i - (i & (stride - 1));
          const u32 hi = lo + stride;
          const bool up = (lo & size) == 0;
          const u32 a = sm_codes[lo], b = sm_codes[hi];
          if ((a > b) == up) {
            sm_codes[lo] = b;
            sm_codes[hi] = a;
          }
        }
        __syncthreads();
      }
    }
    // distinct targets first (the row's length), then the row at its reserved place
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    {
      u32 mine = 0;
      for (u32 i = threadIdx.x; i < k; i += blockDim.x)
        mine += ((i == 0) || ((sm_codes[i - 1] >> 5) != (sm_codes[i] >> 5))) ? 1u : 0u;
      if (mine) atomicAdd(&s_run, mine);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      s_off = (u32)atomicAdd((unsigned long long*)&P.ctr->arena_used, (unsigned long long)s_run);
      s_run = 0;
    }
    __syncthreads();
    const size_t out0 = (size_t)s_off;
    for (u32 base = 0; base < k; base += blockDim.x) {
      const u32 i = base + threadIdx.x;
      bool start = false;
      u32 code = 0xFFFFFFFFu;
      if (i < k) {
        code = sm_codes[i];
        start = (i == 0) || ((sm_codes[i - 1] >> 5) != (code >> 5));
      }
      u32 total;
      const u32 excl = block_excl_scan_u32(start ? 1u : 0u, s_scan, total);
      if (start) {
        u32 mask = 0;
        for (u32 j = i; j < k && (sm_codes[j] >> 5) == (code >> 5); ++j) mask |= 1u << (sm_codes[j] & 31u);
        P.arena[out0 + s_run + excl] = make_uint2(code >> 5, mask);
      }
      __syncthreads();
      if (threadIdx.x == 0) s_run += total;
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      P.row_len[e] = s_run;
      P.row_off[e] = s_off;
    }
    __syncthreads();
  }
}
