// ecb_finalize.cuh — EC id assignment and CSR/CSC emission.
//
// Replaces the id half of alntools/bam_utils.py:693-698 (ec_idx[k] = len(ec_idx): EC id = rank of the
// key's first occurrence) without a sort: every kept EC sets the bit of its first-occurrence position
// in a bitmap over the pushed alignment range, a device-wide scan of the word popcounts ranks the
// bits, and the rank IS the EC id.  Then alntools/bam_utils.py:827-847 + bin_utils.py:208-232:
// row lengths -> exclusive scan -> a_indptr, rows copied from the arena, counts scattered.
#pragma once
#include "ecb_common.cuh"

struct FinalizeParams {
  const EcbEntry* table;
  const u32* ec_slot;
  const u32* row_len;
  const u32* row_off;   // absolute arena offsets
  const uint2* arena;
  const u32* ec_keep;   // NULL = keep every EC (single-sample)
  u32 n_ec;             // provisional ids
  u64 min_base;
  u32* bitmap;
  const u32* word_rank;
  u64* first_rel;       // [n_ec]
  u32* ecid_of;         // [n_ec] final id or ECB_NONE
  int32_t* a_indptr;    // [E+1] holds row lengths before the scan
  int32_t* a_indices;
  int32_t* a_data;
  int32_t* n_indices;   // single-sample N matrix
  int32_t* n_data;
  u32* long_rows_flag;  // set when some row has more than 8 entries
};

__global__ void __launch_bounds__(256) ecb_fin_mark_kernel(const FinalizeParams P) {
  for (u32 e = blockIdx.x * blockDim.x + threadIdx.x; e < P.n_ec; e += gridDim.x * blockDim.x) {
    const u64 rel = P.table[P.ec_slot[e]].first - P.min_base;
    P.first_rel[e] = rel;
    if (P.ec_keep == nullptr || P.ec_keep[e]) atomicOr(&P.bitmap[rel >> 5], 1u << (rel & 31));
  }
}

template <bool SINGLE_SAMPLE>
__global__ void __launch_bounds__(256) ecb_fin_rank_kernel(const FinalizeParams P) {
  for (u32 e = blockIdx.x * blockDim.x + threadIdx.x; e < P.n_ec; e += gridDim.x * blockDim.x) {
    if (P.ec_keep != nullptr && !P.ec_keep[e]) {
      P.ecid_of[e] = ECB_NONE;
      continue;
    }
    const u64 rel = P.first_rel[e];
    const u32 w = (u32)(rel >> 5), b = (u32)(rel & 31);
    const u32 id = P.word_rank[w] + __popc(P.bitmap[w] & ((1u << b) - 1u));
    P.ecid_of[e] = id;
    const u32 len = P.row_len[e];
    P.a_indptr[id] = (int32_t)len;
    if (len > 8 && P.long_rows_flag && *P.long_rows_flag == 0u) *P.long_rows_flag = 1u;
    if (SINGLE_SAMPLE) {
      P.n_indices[id] = (int32_t)id;
      P.n_data[id] = (int32_t)(P.table[P.ec_slot[e]].countm1 + 1u);
    }
  }
}

// Copy rows from the arena to their CSR position.  Rows of up to 8 entries (the bulk): one thread per
// EC; longer rows: one warp per EC (second launch, only when such rows exist).
__global__ void __launch_bounds__(256) ecb_fin_rows_kernel(const FinalizeParams P) {
  for (u32 e = blockIdx.x * blockDim.x + threadIdx.x; e < P.n_ec; e += gridDim.x * blockDim.x) {
    const u32 id = P.ecid_of[e];
    if (id == ECB_NONE) continue;
    const u32 len = P.row_len[e];
    if (len > 8) continue;
    const uint2* src = P.arena + P.row_off[e];
    const size_t dst = (size_t)P.a_indptr[id];
    for (u32 j = 0; j < len; ++j) {
      const uint2 v = src[j];
      P.a_indices[dst + j] = (int32_t)v.x;
      P.a_data[dst + j] = (int32_t)v.y;
    }
  }
}

__global__ void __launch_bounds__(256) ecb_fin_rows_long_kernel(const FinalizeParams P) {
  const int lane = threadIdx.x & 31;
  const u32 warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 n_warps = (gridDim.x * blockDim.x) >> 5;
  for (u32 e = warp_global; e < P.n_ec; e += n_warps) {
    const u32 len = P.row_len[e];
    if (len <= 8) continue;
    const u32 id = P.ecid_of[e];
    if (id == ECB_NONE) continue;
    const size_t src = P.row_off[e];
    const size_t dst = (size_t)P.a_indptr[id];
    for (u32 j = lane; j < len; j += 32) {
      const uint2 v = P.arena[src + j];
      P.a_indices[dst + j] = (int32_t)v.x;
      P.a_data[dst + j] = (int32_t)v.y;
    }
  }
}

__global__ void ecb_set_pair_kernel(int32_t* p, int32_t a, int32_t b) {
  p[0] = a;
  p[1] = b;
}
