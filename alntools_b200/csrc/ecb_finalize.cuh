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
  u32* wide_count;      // number of rows with more than 8 entries ...
  u32* wide_list;       // ... and their provisional ids
  u32* count_of;        // [n_ec] reads per EC (copied out of the table by the mark kernel)
};

__global__ void __launch_bounds__(256) ecb_fin_mark_kernel(const FinalizeParams P) {
  for (u32 e = blockIdx.x * blockDim.x + threadIdx.x; e < P.n_ec; e += gridDim.x * blockDim.x) {
    const EcbEntry* en = P.table + P.ec_slot[e];
    const u64 rel = en->first - P.min_base;
    P.first_rel[e] = rel;
    if (P.count_of) P.count_of[e] = en->countm1 + 1u;
    if (P.ec_keep == nullptr || P.ec_keep[e]) atomicOr(&P.bitmap[rel >> 5], 1u << (rel & 31));
  }
}

__device__ __forceinline__ void fin_warp_append(u32* list, u32* counter, bool take, u32 item) {
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

template <bool SINGLE_SAMPLE>
__global__ void __launch_bounds__(256) ecb_fin_rank_kernel(const FinalizeParams P) {
  for (u32 e = blockIdx.x * blockDim.x + threadIdx.x; e < P.n_ec; e += gridDim.x * blockDim.x) {
    const bool kept = P.ec_keep == nullptr || P.ec_keep[e];
    if (P.wide_list) fin_warp_append(P.wide_list, P.wide_count, kept && P.row_len[e] > 8, e);
    if (!kept) {
      P.ecid_of[e] = ECB_NONE;
      continue;
    }
    const u64 rel = P.first_rel[e];
    const u32 w = (u32)(rel >> 5), b = (u32)(rel & 31);
    const u32 id = P.word_rank[w] + __popc(P.bitmap[w] & ((1u << b) - 1u));
    P.ecid_of[e] = id;
    const u32 len = P.row_len[e];
    if (P.a_indptr == nullptr) continue;   // ids only (slice assembly)
    P.a_indptr[id] = (int32_t)len;
    if (SINGLE_SAMPLE) {
      P.n_indices[id] = (int32_t)id;
      P.n_data[id] = (int32_t)P.count_of[e];
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

// Rows with more than 8 entries: one warp per row.  With a list (n_wide entries of P.wide_list; 0xFFFFFFFF =
// as many as *P.wide_count says) only those rows are visited; without one (list == NULL) every EC is looked at.
__global__ void __launch_bounds__(256) ecb_fin_rows_long_kernel(const FinalizeParams P, const u32* list, u32 n_wide) {
  const int lane = threadIdx.x & 31;
  const u32 warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 n_warps = (gridDim.x * blockDim.x) >> 5;
  if (n_wide == 0xFFFFFFFFu) n_wide = *P.wide_count;   // the list was counted on the device (same stream, earlier launch)
  const u32 n_items = list ? n_wide : P.n_ec;
  for (u32 i = warp_global; i < n_items; i += n_warps) {
    const u32 e = list ? list[i] : i;
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
