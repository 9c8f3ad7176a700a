// ecb_cells.cuh — per-cell (multisample) finalisation kernels.
//
// Replaces alntools/bam_utils_multisample.py:503-636 (file-ordered merge of ec[key][cell] counts, the
// insertion order of cr_totals that fixes the cell order, the minimum-count cell filter, dropping
// ECs left without cells) and :737-747,783-791 (N matrix rows -> CSC).
//
// Input: the (file, EC, cell) table filled by the grouping kernel: key_lo = EC slot << 32 | cell,
// key_hi = file (push) id, first = global position of the triple's first read, countm1 = reads - 1.
//
// Cell order in the reference = first time a cell is met while iterating files in order, inside a
// file the ECs in order of their first read in THAT file, inside a (file, EC) the cells in order of
// first read.  So a cell's sort key is the minimum over its triples of
//     ( first position of (file, EC) , first position of (file, EC, cell) )
// (positions are global and files occupy disjoint ranges, so the file index is implied).
#pragma once
#include "ecb_common.cuh"
#include "ecb_group.cuh"

// Cell totals are accumulated in this many replicas (picked by CTA) and summed afterwards: the
// cell-size distribution of real data is heavy-tailed and one counter per cell would serialise a
// large share of all updates on a few addresses.
#define CELLS_TOTAL_REPLICAS 32

struct CellScratch {
  DevBuf fe_table, pair_table, cell_key, cell_total, cell_new, sort_k[2], sort_v[2], hist, flags, offsets;
  u32 fe_slots = 0, pair_slots = 0;
};

struct CellResult {
  int64_t n_cells = 0;       // max cell id + 1
  int64_t n_kept_cells = 0;
  int64_t nnz_n = 0;
};

struct CellParams {
  const EcbEntry* ttable;  u32 t_slots;
  const EcbEntry* ec_table;
  EcbEntry* fe_table;      u32 fe_mask;
  EcbEntry* pair_table;    u32 pair_mask;
  u64* cell_key;           // [n_cells]
  u64* cell_total;         // [CELLS_TOTAL_REPLICAS * n_cells]: replica r of cell c at r * n_cells + c
  int32_t* cell_new;       // [n_cells] output column or -1
  u32* ec_keep;            // [n_prov]
  const u32* ecid_of;
  u64 min_base;
  u32 n_cells;
  EcbCounters* ctr;
};

// pass 1: largest cell id, and first position of every (file, EC).
__global__ void __launch_bounds__(256) cells_pass1_kernel(const CellParams P) {
  u32 local_max = 0;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < P.t_slots; i += gridDim.x * blockDim.x) {
    Key128 k;
    u64 first;
    u32 cm1, aux;
    load_entry_cg(P.ttable + i, k, first, cm1, aux);
    if (key_empty(k)) continue;
    local_max = max(local_max, (u32)(k.lo & 0xFFFFFFFFull) + 1u);
    bool claimed;
    u64 seen;
    const u32 s = table_find_or_claim(P.fe_table, P.fe_mask, Key128{k.lo >> 32, k.hi}, claimed, seen);
    if (s == ECB_NONE) {
      atomicOr(&P.ctr->error, ECB_DEVERR_EC_CAPACITY);
      continue;
    }
    if (first < seen) atomicMin(&P.fe_table[s].first, first);
  }
  local_max = __reduce_max_sync(ECB_FULL, local_max);
  if ((threadIdx.x & 31) == 0 && local_max) atomicMax(&P.ctr->scratch[0], local_max);
}

// pass 2: cell sort keys, cell totals, (EC, cell) counts summed over files.
__global__ void __launch_bounds__(256) cells_pass2_kernel(const CellParams P) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < P.t_slots; i += gridDim.x * blockDim.x) {
    Key128 k;
    u64 first;
    u32 cm1, aux;
    load_entry_cg(P.ttable + i, k, first, cm1, aux);
    if (key_empty(k)) continue;
    const u32 cell = (u32)(k.lo & 0xFFFFFFFFull);
    const u32 slot = (u32)(k.lo >> 32);
    const u32 count = cm1 + 1u;
    const u32 fs = table_find(P.fe_table, P.fe_mask, Key128{(u64)slot, k.hi});
    const u64 ec_first = P.fe_table[fs].first;
    const u64 cand = ((ec_first - P.min_base) << 32) | (first - P.min_base);
    if (cand < *reinterpret_cast<volatile u64*>(&P.cell_key[cell])) atomicMin(&P.cell_key[cell], cand);
    atomicAdd(&P.cell_total[(size_t)(blockIdx.x % CELLS_TOTAL_REPLICAS) * P.n_cells + cell], (u64)count);
    bool claimed;
    u64 seen;
    const u32 ps = table_find_or_claim(P.pair_table, P.pair_mask, Key128{k.lo, 0ull}, claimed, seen);
    if (ps == ECB_NONE) {
      atomicOr(&P.ctr->error, ECB_DEVERR_EC_CAPACITY);
      continue;
    }
    atomicAdd(&P.pair_table[ps].countm1, count);
  }
}

// cell_total[c] = sum of its replicas.
__global__ void __launch_bounds__(256) cells_total_reduce_kernel(u64* __restrict__ cell_total, u32 n_cells) {
  for (u32 c = blockIdx.x * blockDim.x + threadIdx.x; c < n_cells; c += gridDim.x * blockDim.x) {
    u64 t = 0;
    for (u32 r = 0; r < CELLS_TOTAL_REPLICAS; ++r) t += cell_total[(size_t)r * n_cells + c];
    cell_total[c] = t;
  }
}

// keys for the cell sort: absent cells get ~0 and end up last.
__global__ void __launch_bounds__(256) cells_sort_input_kernel(const u64* __restrict__ cell_key, u32 n_cells,
                                                               u64* __restrict__ keys, u32* __restrict__ vals) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n_cells; i += gridDim.x * blockDim.x) {
    keys[i] = cell_key[i];
    vals[i] = i;
  }
}

// flags[i] = 1 when the i-th cell in reference order passes the minimum-count filter.
__global__ void __launch_bounds__(256) cells_keep_flags_kernel(const u64* __restrict__ sorted_keys,
                                                               const u32* __restrict__ sorted_cells, u32 n_cells,
                                                               const u64* __restrict__ cell_total, u64 min_count,
                                                               u32* __restrict__ flags) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n_cells; i += gridDim.x * blockDim.x)
    flags[i] = (sorted_keys[i] != ~0ull && cell_total[sorted_cells[i]] >= min_count) ? 1u : 0u;
}

__global__ void __launch_bounds__(256) cells_assign_kernel(const u64* __restrict__ sorted_keys,
                                                           const u32* __restrict__ sorted_cells, u32 n_cells,
                                                           const u64* __restrict__ cell_total, u64 min_count,
                                                           const u32* __restrict__ new_idx, int32_t* __restrict__ cell_new,
                                                           int32_t* __restrict__ cell_order) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n_cells; i += gridDim.x * blockDim.x) {
    const u32 cell = sorted_cells[i];
    const bool keep = sorted_keys[i] != ~0ull && cell_total[cell] >= min_count;
    cell_new[cell] = keep ? (int32_t)new_idx[i] : -1;
    if (keep) cell_order[new_idx[i]] = (int32_t)cell;
  }
}

// ECs that keep at least one cell survive (bam_utils_multisample.py:616-632).
__global__ void __launch_bounds__(256) cells_ec_keep_kernel(const CellParams P) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i <= P.pair_mask; i += gridDim.x * blockDim.x) {
    Key128 k;
    u64 first;
    u32 cm1, aux;
    load_entry_cg(P.pair_table + i, k, first, cm1, aux);
    if (key_empty(k)) continue;
    if (P.cell_new[(u32)(k.lo & 0xFFFFFFFFull)] >= 0) P.ec_keep[P.ec_table[(u32)(k.lo >> 32)].aux] = 1u;
  }
}

__global__ void __launch_bounds__(256) cells_pair_flags_kernel(const CellParams P, u32* __restrict__ flags) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i <= P.pair_mask; i += gridDim.x * blockDim.x) {
    Key128 k;
    u64 first;
    u32 cm1, aux;
    load_entry_cg(P.pair_table + i, k, first, cm1, aux);
    flags[i] = (!key_empty(k) && P.cell_new[(u32)(k.lo & 0xFFFFFFFFull)] >= 0) ? 1u : 0u;
  }
}

// compact kept pairs into sort input: key = output column << 32 | EC id, value = count.
__global__ void __launch_bounds__(256) cells_pair_emit_kernel(const CellParams P, const u32* __restrict__ flags,
                                                              const u32* __restrict__ offsets, u64* __restrict__ keys,
                                                              u32* __restrict__ vals) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i <= P.pair_mask; i += gridDim.x * blockDim.x) {
    if (!flags[i]) continue;
    const EcbEntry e = P.pair_table[i];
    const u32 cell = (u32)(e.key_lo & 0xFFFFFFFFull);
    const u32 ecid = P.ecid_of[P.ec_table[(u32)(e.key_lo >> 32)].aux];
    keys[offsets[i]] = ((u64)(u32)P.cell_new[cell] << 32) | ecid;
    vals[offsets[i]] = e.countm1 + 1u;
  }
}

// sorted (column, EC) pairs -> CSC arrays.  Every kept column has at least one entry.
__global__ void __launch_bounds__(256) cells_csc_kernel(const u64* __restrict__ keys, const u32* __restrict__ vals,
                                                        u32 nnz, u32 n_cols, int32_t* __restrict__ indptr,
                                                        int32_t* __restrict__ indices, int32_t* __restrict__ data) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += gridDim.x * blockDim.x) {
    const u32 col = (u32)(keys[i] >> 32);
    if (i == 0 || (u32)(keys[i - 1] >> 32) != col) indptr[col] = (int32_t)i;
    indices[i] = (int32_t)(keys[i] & 0xFFFFFFFFull);
    data[i] = (int32_t)vals[i];
    if (i == nnz - 1) indptr[n_cols] = (int32_t)nnz;
  }
}
