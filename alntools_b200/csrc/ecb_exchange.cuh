// ecb_exchange.cuh — kernels of the multi-GPU exchange.
//
// Reads shard across GPUs by contiguous chunk (the reference's utils.partition + ordered imap,
// alntools/bam_utils.py:647-680); every GPU builds a local EC table.  Global dedup = the chunk merge
// of alntools/bam_utils.py:680-698 (sum the counts of equal keys, EC id = rank of the GLOBAL first
// occurrence): local ECs are hash-partitioned to an owner GPU (all-to-all), the owner merges them,
// and the global ids come from the same first-occurrence bitmap as on one GPU, OR-ed across ranks.
//
// Record layout of an exported EC ("meta", 5 x int64): key_lo, key_hi, first,
// count << 32 | row_len, row offset inside the partition (in (target, mask) pairs).
#pragma once
#include "ecb_common.cuh"
#include "ecb_group.cuh"

#define ECB_META_WORDS 5
#define ECB_MAX_WORLD 64

__device__ __forceinline__ u32 ecb_owner_of(const Key128& k, u32 world) {
  return fmix32((u32)(k.hi >> 32) ^ 0x5bd1e995u) % world;
}

struct ExportParams {
  const EcbEntry* table;
  const u32* ec_slot;
  const u32* row_len;
  const u32* row_off;
  const uint2* arena;
  u32 n_ec;
  u32 world;
  u32* counts;       // [2*world] ECs, rows per owner (count pass) / running cursors (fill pass)
  const u64* base;   // [2*world] partition bases (fill pass)
  long long* meta;
  int2* rows;
};

// Both export kernels aggregate per CTA in shared memory: a plain atomicAdd per EC on the `world`
// owner counters would serialise millions of operations on a handful of addresses.
__global__ void __launch_bounds__(256) ecb_export_count_kernel(const ExportParams P) {
  __shared__ u32 s_cnt[2 * ECB_MAX_WORLD];
  for (u32 i = threadIdx.x; i < 2 * P.world; i += blockDim.x) s_cnt[i] = 0u;
  __syncthreads();
  for (u32 e = blockIdx.x * blockDim.x + threadIdx.x; e < P.n_ec; e += gridDim.x * blockDim.x) {
    const EcbEntry* en = P.table + P.ec_slot[e];
    const u32 owner = ecb_owner_of(Key128{en->key_lo, en->key_hi}, P.world);
    atomicAdd(&s_cnt[owner], 1u);
    atomicAdd(&s_cnt[P.world + owner], P.row_len[e]);
  }
  __syncthreads();
  for (u32 i = threadIdx.x; i < 2 * P.world; i += blockDim.x)
    if (s_cnt[i]) atomicAdd(&P.counts[i], s_cnt[i]);
}

__global__ void __launch_bounds__(256) ecb_export_fill_kernel(const ExportParams P) {
  __shared__ u32 s_cnt[2 * ECB_MAX_WORLD];   // per tile: ECs / rows per owner, then their global bases
  const u32 tiles = (P.n_ec + blockDim.x - 1) / blockDim.x;
  for (u32 tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    for (u32 i = threadIdx.x; i < 2 * P.world; i += blockDim.x) s_cnt[i] = 0u;
    __syncthreads();
    const u32 e = tile * blockDim.x + threadIdx.x;
    const bool live = e < P.n_ec;
    EcbEntry en{};
    u32 owner = 0, len = 0, idx = 0, roff = 0;
    if (live) {
      en = P.table[P.ec_slot[e]];
      owner = ecb_owner_of(Key128{en.key_lo, en.key_hi}, P.world);
      len = P.row_len[e];
      idx = atomicAdd(&s_cnt[owner], 1u);
      roff = atomicAdd(&s_cnt[P.world + owner], len);
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < 2 * P.world; i += blockDim.x) {
      const u32 c = s_cnt[i];
      s_cnt[i] = c ? atomicAdd(&P.counts[i], c) : 0u;
    }
    __syncthreads();
    if (live) {
      idx += s_cnt[owner];
      roff += s_cnt[P.world + owner];
      long long* m = P.meta + (P.base[owner] + idx) * ECB_META_WORDS;
      m[0] = (long long)en.key_lo;
      m[1] = (long long)en.key_hi;
      m[2] = (long long)en.first;
      m[3] = (long long)(((u64)(en.countm1 + 1u) << 32) | len);
      m[4] = (long long)roff;
      const uint2* src = P.arena + P.row_off[e];
      int2* dst = P.rows + P.base[P.world + owner] + roff;
      for (u32 j = 0; j < len; ++j) dst[j] = make_int2((int)src[j].x, (int)src[j].y);
    }
    __syncthreads();
  }
}

// ---- fused partition + dispatch over peer memory ------------------------------------------------------
// Every rank owns an ARENA (plain cudaMalloc, mapped into its peers through CUDA IPC): a small header
// {records, rows, overflow} followed by a record area and a row area.  The kernel below does what
// ecb_export_* + an all-to-all did, in one pass: it finds the owner of every local EC and stores the
// record and its row straight into the owner's arena over NVLink (or into its own).  Space in the remote
// arena is reserved with ONE remote atomicAdd per CTA tile and owner; all other traffic is plain stores.
#define ECB_ARENA_HEADER_BYTES 256

struct ArenaTargets {
  unsigned long long* hdr[ECB_MAX_WORLD];   // [0] records, [1] rows, [2] overflow flag
  long long* meta[ECB_MAX_WORLD];
  int2* rows[ECB_MAX_WORLD];
  unsigned long long cap_ec, cap_rows;
};

#define ECB_XT_ROWS 2048   // rows of one tile staged in shared memory (more: stored one by one)

// One tile = 256 local ECs.  The tile's records (and rows) are first laid out in shared memory grouped
// by owner, then each owner's segment goes out with fully coalesced 8-byte stores - NVLink wants long
// contiguous writes, not one 8-byte word per lane every 40 bytes.
__global__ void __launch_bounds__(256) ecb_export_to_arenas_kernel(const ExportParams P, const ArenaTargets A) {
  __shared__ u32 s_cnt[2 * ECB_MAX_WORLD];                  // per tile: ECs / rows per owner
  __shared__ u32 s_off[2 * ECB_MAX_WORLD];                  // ... their exclusive prefix over the owners
  __shared__ unsigned long long s_base[2 * ECB_MAX_WORLD];  // where the tile's share starts in the owner's arena
  __shared__ u32 s_drop[ECB_MAX_WORLD];
  __shared__ u32 s_rows_total;
  __shared__ __align__(16) long long s_meta[256 * ECB_META_WORDS];
  __shared__ __align__(16) int2 s_rows[ECB_XT_ROWS];
  const u32 W = P.world;
  const u32 tiles = (P.n_ec + blockDim.x - 1) / blockDim.x;
  for (u32 tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    for (u32 i = threadIdx.x; i < 2 * W; i += blockDim.x) s_cnt[i] = 0u;
    __syncthreads();
    const u32 e = tile * blockDim.x + threadIdx.x;
    const bool live = e < P.n_ec;
    EcbEntry en{};
    u32 owner = 0, len = 0, idx = 0, roff = 0;
    if (live) {
      en = P.table[P.ec_slot[e]];
      owner = ecb_owner_of(Key128{en.key_lo, en.key_hi}, W);
      len = P.row_len[e];
      idx = atomicAdd(&s_cnt[owner], 1u);
      roff = atomicAdd(&s_cnt[W + owner], len);
    }
    __syncthreads();
    if (threadIdx.x < W) {   // reserve space in the owners' arenas: one remote atomic per owner and area
      const u32 o = threadIdx.x;
      const u32 ce = s_cnt[o], cr = s_cnt[W + o];
      s_drop[o] = 0u;
      if (ce) {
        const unsigned long long be = atomicAdd(A.hdr[o] + 0, (unsigned long long)ce);
        const unsigned long long br = atomicAdd(A.hdr[o] + 1, (unsigned long long)cr);
        s_base[o] = be;
        s_base[W + o] = br;
        if (be + ce > A.cap_ec || br + cr > A.cap_rows) {
          atomicExch(A.hdr[o] + 2, 1ull);
          s_drop[o] = 1u;
        }
      }
    }
    if (threadIdx.x == 32) {   // meanwhile: tile-local offsets of the owners' segments
      u32 ae = 0, ar = 0;
      for (u32 o = 0; o < W; ++o) {
        s_off[o] = ae;
        s_off[W + o] = ar;
        ae += s_cnt[o];
        ar += s_cnt[W + o];
      }
      s_rows_total = ar;
    }
    __syncthreads();
    const bool stage_rows = s_rows_total <= ECB_XT_ROWS;
    if (live && !s_drop[owner]) {
      const unsigned long long rat = s_base[W + owner] + roff;
      long long* m = s_meta + (size_t)(s_off[owner] + idx) * ECB_META_WORDS;
      m[0] = (long long)en.key_lo;
      m[1] = (long long)en.key_hi;
      m[2] = (long long)en.first;
      m[3] = (long long)(((u64)(en.countm1 + 1u) << 32) | len);
      m[4] = (long long)rat;
      const uint2* src = P.arena + P.row_off[e];
      if (stage_rows) {
        int2* dst = s_rows + s_off[W + owner] + roff;
        for (u32 j = 0; j < len; ++j) dst[j] = make_int2((int)src[j].x, (int)src[j].y);
      } else {
        int2* dst = A.rows[owner] + rat;
        for (u32 j = 0; j < len; ++j) dst[j] = make_int2((int)src[j].x, (int)src[j].y);
      }
    }
    __syncthreads();
    for (u32 o = 0; o < W; ++o) {
      if (s_cnt[o] == 0u || s_drop[o]) continue;
      const u32 n_words = s_cnt[o] * ECB_META_WORDS;
      const long long* src = s_meta + (size_t)s_off[o] * ECB_META_WORDS;
      long long* dst = A.meta[o] + s_base[o] * ECB_META_WORDS;
      for (u32 w = threadIdx.x; w < n_words; w += blockDim.x) dst[w] = src[w];
      if (stage_rows) {
        const u32 n_rows = s_cnt[W + o];
        const long long* rsrc = reinterpret_cast<const long long*>(s_rows + s_off[W + o]);
        long long* rdst = reinterpret_cast<long long*>(A.rows[o] + s_base[W + o]);
        for (u32 w = threadIdx.x; w < n_rows; w += blockDim.x) rdst[w] = rsrc[w];
      }
    }
    __syncthreads();
  }
}

#define ECB_SLICE_WORDS 2

// ---- second dispatch: every merged EC goes to the rank whose SHARD holds its first occurrence.  EC ids are ranks of first-occurrence positions and the shards partition the positions, so
// the ECs that arrive at a rank form one contiguous id range, ordered by their position inside the shard:
// the receiver ranks them with a bitmap over ITS OWN positions only.  Nothing global is built: no bitmap over
// all ranks' positions, no all-reduce of it, no padded arrays; the id range of a rank starts where the ranges
// of the shards in front of it end (one all-gather of a count).  Record: {position inside the shard |
// count << 32, row offset | len << 40} plus the row.  (Round 1 dispatched by EC-id range after ranking a bitmap
// over ALL ranks' positions that had been OR-ed by an all-reduce - 61 MB at 8 GPUs; that form is gone.)
struct OrderDispatchParams {
  const EcbEntry* table;
  const u32* ec_slot;
  const u32* row_len;
  const u32* row_off;
  const uint2* arena;
  u32 n_ec;
  u32 world;
  u64 lo[ECB_MAX_WORLD];      // first position of each shard, ascending; ranks without alignments come last
  u32 dest[ECB_MAX_WORLD];    // ... and the rank that holds it
  u32 n_shards;               // shards with alignments
};

__global__ void __launch_bounds__(256) ecb_order_dispatch_kernel(const OrderDispatchParams P, const ArenaTargets A) {
  __shared__ u32 s_cnt[2 * ECB_MAX_WORLD];
  __shared__ u32 s_off[2 * ECB_MAX_WORLD];
  __shared__ unsigned long long s_base[2 * ECB_MAX_WORLD];
  __shared__ u32 s_drop[ECB_MAX_WORLD];
  __shared__ u32 s_rows_total;
  __shared__ __align__(16) unsigned long long s_rec[256 * ECB_SLICE_WORDS];
  __shared__ __align__(16) int2 s_rows[ECB_XT_ROWS];
  const u32 W = P.world;
  const u32 tiles = (P.n_ec + blockDim.x - 1) / blockDim.x;
  for (u32 tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    for (u32 i = threadIdx.x; i < 2 * W; i += blockDim.x) s_cnt[i] = 0u;
    __syncthreads();
    const u32 e = tile * blockDim.x + threadIdx.x;
    const bool live = e < P.n_ec;
    u32 dest = 0, len = 0, idx = 0, roff = 0, count = 0;
    u64 rel = 0;
    if (live) {
      const EcbEntry en = P.table[P.ec_slot[e]];
      u32 k = 0;   // the last shard that starts at or before the position
      while (k + 1 < P.n_shards && P.lo[k + 1] <= en.first) ++k;
      dest = P.dest[k];
      rel = en.first - P.lo[k];
      count = en.countm1 + 1u;
      len = P.row_len[e];
      idx = atomicAdd(&s_cnt[dest], 1u);
      roff = atomicAdd(&s_cnt[W + dest], len);
    }
    __syncthreads();
    if (threadIdx.x < W) {
      const u32 o = threadIdx.x;
      const u32 ce = s_cnt[o], cr = s_cnt[W + o];
      s_drop[o] = 0u;
      if (ce) {
        const unsigned long long be = atomicAdd(A.hdr[o] + 0, (unsigned long long)ce);
        const unsigned long long br = atomicAdd(A.hdr[o] + 1, (unsigned long long)cr);
        s_base[o] = be;
        s_base[W + o] = br;
        if ((be + ce) * ECB_SLICE_WORDS > A.cap_ec * ECB_META_WORDS || br + cr > A.cap_rows) {
          atomicExch(A.hdr[o] + 2, 1ull);
          s_drop[o] = 1u;
        }
      }
    }
    if (threadIdx.x == 32) {
      u32 ae = 0, ar = 0;
      for (u32 o = 0; o < W; ++o) {
        s_off[o] = ae;
        s_off[W + o] = ar;
        ae += s_cnt[o];
        ar += s_cnt[W + o];
      }
      s_rows_total = ar;
    }
    __syncthreads();
    const bool stage_rows = s_rows_total <= ECB_XT_ROWS;
    if (live && !s_drop[dest]) {
      const unsigned long long rat = s_base[W + dest] + roff;
      unsigned long long* r = s_rec + (size_t)(s_off[dest] + idx) * ECB_SLICE_WORDS;
      r[0] = (rel > 0xFFFFFFFEull ? 0xFFFFFFFFull : rel) | ((unsigned long long)count << 32);
      r[1] = rat | ((unsigned long long)len << 40);
      const uint2* src = P.arena + P.row_off[e];
      int2* dst = stage_rows ? s_rows + s_off[W + dest] + roff : A.rows[dest] + rat;
      for (u32 j = 0; j < len; ++j) dst[j] = make_int2((int)src[j].x, (int)src[j].y);
    }
    __syncthreads();
    for (u32 o = 0; o < W; ++o) {
      if (s_cnt[o] == 0u || s_drop[o]) continue;
      const u32 n_words = s_cnt[o] * ECB_SLICE_WORDS;
      const unsigned long long* src = s_rec + (size_t)s_off[o] * ECB_SLICE_WORDS;
      unsigned long long* dst = reinterpret_cast<unsigned long long*>(A.meta[o]) + s_base[o] * ECB_SLICE_WORDS;
      for (u32 w = threadIdx.x; w < n_words; w += blockDim.x) dst[w] = src[w];
      if (stage_rows) {
        const u32 n_rows = s_cnt[W + o];
        const long long* rsrc = reinterpret_cast<const long long*>(s_rows + s_off[W + o]);
        long long* rdst = reinterpret_cast<long long*>(A.rows[o] + s_base[W + o]);
        for (u32 w = threadIdx.x; w < n_rows; w += blockDim.x) rdst[w] = rsrc[w];
      }
    }
    __syncthreads();
  }
}

// Slice assembly, pass 0: one bit per arrived EC at its position inside the shard.
__global__ void __launch_bounds__(256) ecb_order_mark_kernel(const unsigned long long* __restrict__ rec, u32 n_rec, u32 span,
                                                            u32* bitmap, u32* bad) {
  for (u32 r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rec; r += gridDim.x * blockDim.x) {
    const u32 rel = (u32)(rec[(size_t)r * ECB_SLICE_WORDS] & 0xFFFFFFFFull);
    if (rel >= span) {
      atomicOr(bad, 1u);
      continue;
    }
    const u32 bit = 1u << (rel & 31);
    if (atomicOr(&bitmap[rel >> 5], bit) & bit) atomicOr(bad, 2u);   // two ECs cannot start with the same read
  }
}

// ... pass 1: the rank of its bit is an EC's id inside the slice; row length, count and record index go there.
__global__ void __launch_bounds__(256) ecb_order_lens_kernel(const unsigned long long* __restrict__ rec, u32 n_rec, u32 span,
                                                            const u32* __restrict__ bitmap, const u32* __restrict__ word_rank,
                                                            int32_t* lens, int32_t* counts, u32* rec_of) {
  for (u32 r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rec; r += gridDim.x * blockDim.x) {
    const unsigned long long w0 = rec[(size_t)r * ECB_SLICE_WORDS], w1 = rec[(size_t)r * ECB_SLICE_WORDS + 1];
    const u32 rel = (u32)(w0 & 0xFFFFFFFFull);
    if (rel >= span) continue;
    const u32 w = rel >> 5, b = rel & 31;
    const u32 i = word_rank[w] + (u32)__popc(bitmap[w] & ((1u << b) - 1u));
    if (i >= n_rec) continue;   // (only after a duplicate position, which pass 0 has reported)
    lens[i] = (int32_t)(w1 >> 40);
    counts[i] = (int32_t)(w0 >> 32);
    rec_of[i] = r;
  }
}

// Slice assembly, pass 2: rows to their CSR position (thread per id).
__global__ void __launch_bounds__(256) ecb_slice_rows_kernel(const unsigned long long* __restrict__ rec,
                                                            const int2* __restrict__ rows, const u32* __restrict__ rec_of,
                                                            const int32_t* __restrict__ indptr, u32 slice_n,
                                                            int32_t* indices, int32_t* data) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < slice_n; i += gridDim.x * blockDim.x) {
    const unsigned long long w1 = rec[(size_t)rec_of[i] * ECB_SLICE_WORDS + 1];
    const int2* src = rows + (w1 & ((1ull << 40) - 1ull));
    const u32 len = (u32)(w1 >> 40);
    const size_t dst = (size_t)indptr[i];
    for (u32 j = 0; j < len; ++j) {
      indices[dst + j] = src[j].x;
      data[dst + j] = src[j].y;
    }
  }
}

struct PartTable {
  long long row_base[ECB_MAX_WORLD];  // first row of each source partition in the receive buffer
  long long ec_end[ECB_MAX_WORLD];    // cumulative record count after each source partition
  u32 n;
};

struct ImportParams {
  const long long* meta;
  const int2* rows;      // received rows
  PartTable parts;
  u32 n_rec;
  EcbEntry* table;
  u32 mask;
  u32* ec_slot;
  u32* ec_rep;     // provisional id -> record index of the import that created it
  u32* row_len;
  u32* row_off;
  uint2* arena;
  EcbCounters* ctr;
};

// Merge received records into the owner table: sum counts, min first (bam_utils.py:693-698); the records that
// bring a NEW key also get their provisional id and their row's place in the arena here - ids and arena space
// are reserved once per CTA tile (a reservation per warp would put a hundred thousand atomics on one address)
// - and their row is copied on the spot: one kernel, no scan, no host round trip in between.
__global__ void __launch_bounds__(256) ecb_import_insert_kernel(const ImportParams P) {
  __shared__ u32 s_scan[10];
  const u32 tiles = (P.n_rec + blockDim.x - 1) / blockDim.x;
  for (u32 tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const u32 i = tile * blockDim.x + threadIdx.x;
    const bool live = i < P.n_rec;
    bool claimed = false;
    u32 slot = ECB_NONE, len = 0;
    const long long* m = P.meta + (size_t)(live ? i : 0u) * ECB_META_WORDS;
    if (live) {
      const Key128 key{(u64)m[0], (u64)m[1]};
      const u64 first = (u64)m[2];
      const u32 count = (u32)((u64)m[3] >> 32);
      u64 seen;
      slot = table_find_or_claim<true>(P.table, P.mask, key, claimed, seen);
      if (slot == ECB_NONE) {
        atomicOr(&P.ctr->error, ECB_DEVERR_EC_CAPACITY);
        claimed = false;
      } else {
        EcbEntry* e = P.table + slot;
        atomicAdd(&e->countm1, count);
        if (first < seen) atomicMin(&e->first, first);
      }
      if (claimed) len = (u32)((u64)m[3] & 0xFFFFFFFFull);
    }
    // provisional ids and arena space of the tile's new ECs
    u32 n_new, n_rows;
    const u32 id_excl = block_excl_scan_u32(claimed ? 1u : 0u, s_scan, n_new);
    const u32 row_excl = block_excl_scan_u32(len, s_scan, n_rows);
    if (threadIdx.x == 0) {
      s_scan[8] = n_new ? atomicAdd(&P.ctr->n_ec, n_new) : 0u;
      s_scan[9] = n_rows ? (u32)atomicAdd((unsigned long long*)&P.ctr->arena_used, (unsigned long long)n_rows) : 0u;
    }
    __syncthreads();
    const u32 id = s_scan[8] + id_excl, off = s_scan[9] + row_excl;
    __syncthreads();
    if (claimed) {
      P.table[slot].aux = id;
      P.ec_slot[id] = slot;
      P.ec_rep[id] = i;
      P.row_len[id] = len;
      P.row_off[id] = off;
      u32 part = 0;
      while (part + 1 < P.parts.n && (long long)i >= P.parts.ec_end[part]) ++part;
      const int2* src = P.rows + P.parts.row_base[part] + m[4];
      uint2* dst = P.arena + off;
      for (u32 j = 0; j < len; ++j) dst[j] = make_uint2((u32)src[j].x, (u32)src[j].y);
    }
  }
}

// Shift the first-occurrence key of every EC of a context by `delta` (ecb_rebase): a rank that decodes one
// shard of a file learns the global position of its shard only when all ranks have counted theirs.
__global__ void __launch_bounds__(256) ecb_rebase_kernel(EcbEntry* table, const u32* __restrict__ ec_slot, u32 n_ec,
                                                         u64 delta) {
  for (u32 e = blockIdx.x * blockDim.x + threadIdx.x; e < n_ec; e += gridDim.x * blockDim.x)
    table[ec_slot[e]].first += delta;
}
