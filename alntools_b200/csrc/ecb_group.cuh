// ecb_group.cuh — the grouping + hash-insert kernel (the hot path), the overflow replay kernel and
// the table rehash kernels.
//
// What it replaces: alntools/bam_utils.py:258-344 (per-alignment loop: group consecutive alignments
// by read, collapse duplicate tids, ec[key] += 1) and the ordering half of :680-698 (EC id = rank of
// the key's first occurrence), on int32 columns.
//
// Shape of the kernel (HBM-bound integer work, no tensor cores):
//   * one CTA per contiguous chunk of the alignment stream (grid = resident CTAs x 148 SMs).  The CTA
//     walks its chunk in tiles of 1024 alignments; the three columns of the NEXT tile are fetched by
//     the TMA engine (cp.async.bulk, 3 x 4 KB, mbarrier completion) into a two-stage shared-memory
//     ring while the current tile is processed, so no thread spends registers or issue slots on the
//     streaming loads;
//   * a read is owned by the CTA in whose chunk it STARTS; the owner runs past its chunk end until
//     the read closes, the next CTA skips the leading partial read;
//   * inside a tile every warp owns 128 consecutive alignments as 4 rows of 32 (lane = consecutive
//     alignment), so read boundaries, read starts, duplicate (target, haplotype) pairs and per-read
//     sums are warp-wide bit tricks: one ballot gives every lane its read start, one 64-bit
//     __match_any_sync on (read start, element code) finds duplicates inside a read, one on the read
//     start groups the lanes of a read, and __reduce_add_sync adds the 128-bit element mixes of a
//     read (commutative set hash -> no per-read sort);
//   * closed reads are compacted into a shared-memory queue so the insert phase runs with full
//     warps; a per-CTA shared-memory cache absorbs the hot ECs (a few hundred keys carry more than
//     half of the reads), everything else goes to the HBM table: one 256-bit sector load per probe,
//     a 128-bit atomicCAS only when the slot looks empty, RED.ADD on the count and atomicMin on the
//     first-occurrence key only when it can lower it.  The cache is flushed at the end of the chunk.
#pragma once
#include "ecb_common.cuh"

#define ECB_ROWS 4                     // rows of 32 alignments per warp and tile
#define ECB_WARP_SPAN (32 * ECB_ROWS)  // 128
#define ECB_CACHE 512                  // per-CTA hot-EC cache entries

// A hot-cache entry that could not be flushed because the table was too full (replayed after growth).
struct EcbSpill {
  u64 lo, hi;
  u32 count, first, rep, len;
};

struct GroupParams {
  const int32_t* rg;
  const int32_t* tg;
  const int32_t* hp;
  const int32_t* cell;
  int n;           // alignments in this push
  int chunk_len;   // alignments per CTA, multiple of ECB_TILE
  u64 order_base;
  int drop_last;
  int use_cache;
  int n_targets;
  int n_haps;
  EcbEntry* table;
  u32 mask;
  u32* ec_slot;    // [capacity] provisional id -> table slot
  u32* ec_rep;     // [capacity] provisional id -> offset (in this push) of a read with that key
  u32* ec_len;     // [capacity] provisional id -> number of alignments of that read
  EcbCounters* ctr;
  u32* overflow_bits;  // [ceil(n/32)] reads that must be replayed after a table growth
  EcbSpill* spill;     // [grid * ECB_CACHE] cache entries that must be replayed after a table growth
  EcbEntry* ttable;    // (file, EC slot, cell) table, with cells only
  u32 tmask;
  u32 push_id;
};

// ---- TMA bulk copy + mbarrier (sm_90+/sm_100a) ----------------------------------------------------
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, u32 bytes, u64* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
  u32 done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

// Find `key` or claim an empty slot for it.  Returns the slot, or ECB_NONE when ECB_MAX_PROBE
// slots were tried.  first_seen = the entry's `first` as loaded (+inf when unknown/new).
__device__ __forceinline__ u32 table_find_or_claim(EcbEntry* table, u32 mask, const Key128& key,
                                                   bool& claimed, u64& first_seen) {
  u32 slot = key_slot_hash(key) & mask;
  claimed = false;
  first_seen = ~0ull;
  for (int p = 0; p < ECB_MAX_PROBE; ++p) {
    EcbEntry* e = table + slot;
    Key128 k;
    u64 first;
    u32 cm1, aux;
    load_entry_cg(e, k, first, cm1, aux);
    if (key_eq(k, key)) {
      first_seen = first;
      return slot;
    }
    if (key_empty(k)) {
      Key128 old = atomic_cas128(e, Key128{~0ull, ~0ull}, key);
      if (key_empty(old)) {
        claimed = true;
        return slot;
      }
      if (key_eq(old, key)) return slot;
    }
    slot = (slot + 1) & mask;
  }
  return ECB_NONE;
}

// Lookup only (no claim); ECB_NONE if absent.
__device__ __forceinline__ u32 table_find(const EcbEntry* table, u32 mask, const Key128& key) {
  u32 slot = key_slot_hash(key) & mask;
  for (u32 p = 0; p <= mask; ++p) {
    Key128 k;
    u64 first;
    u32 cm1, aux;
    load_entry_cg(table + slot, k, first, cm1, aux);
    if (key_eq(k, key)) return slot;
    if (key_empty(k)) return ECB_NONE;
    slot = (slot + 1) & mask;
  }
  return ECB_NONE;
}

__device__ __forceinline__ Key128 triple_key(u32 slot, u32 cell, u32 push_id) {
  return Key128{((u64)slot << 32) | cell, (u64)push_id};
}

// Insert `count` occurrences of (file, EC, cell).  The triple table is sized by the host so that it
// cannot fill up; running out of probes is reported as a device error.
__device__ __forceinline__ void triple_upsert(const GroupParams& P, u32 slot, u32 cell, u64 pos) {
  bool claimed;
  u64 first_seen;
  if ((int)cell < 0) {
    atomicOr(&P.ctr->error, ECB_DEVERR_VALUE_RANGE);
    return;
  }
  u32 ts = table_find_or_claim(P.ttable, P.tmask, triple_key(slot, cell, P.push_id), claimed, first_seen);
  if (ts == ECB_NONE) {
    atomicAdd(&P.ctr->n_triple_overflow, 1u);
    return;
  }
  EcbEntry* te = P.ttable + ts;
  atomicAdd(&te->countm1, 1u);
  if (pos < first_seen) atomicMin(&te->first, pos);
  if (claimed) atomicAdd(&P.ctr->n_triples, 1u);
}

// Provisional EC id for the lanes that just claimed a slot: one atomic per converged group of lanes.
__device__ __forceinline__ u32 alloc_ec_ids(EcbCounters* ctr, bool claimed) {
  const u32 act = __activemask();
  const u32 cm = __ballot_sync(act, claimed);
  if (!cm) return ECB_NONE;
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(cm) - 1;
  u32 base = 0;
  if (lane == leader) base = atomicAdd(&ctr->n_ec, (u32)__popc(cm));
  base = __shfl_sync(act, base, leader);
  return claimed ? base + (u32)__popc(cm & ((1u << lane) - 1u)) : ECB_NONE;
}

// Add `count` reads with key `key` to the HBM table.  s/len describe one read with that key (used
// when the key is new).  Returns the slot or ECB_NONE when the table is too full (caller records it).
__device__ __forceinline__ u32 global_upsert(const GroupParams& P, const Key128& key, u32 count, u32 first_local,
                                             u32 s, u32 len) {
  bool claimed;
  u64 first_seen;
  const u32 slot = table_find_or_claim(P.table, P.mask, key, claimed, first_seen);
  if (slot != ECB_NONE) {
    EcbEntry* e = P.table + slot;
    atomicAdd(&e->countm1, count);
    const u64 pos = P.order_base + first_local;
    if (pos < first_seen) atomicMin(&e->first, pos);
  }
  const u32 ecl = alloc_ec_ids(P.ctr, claimed);
  if (claimed) {
    P.table[slot].aux = ecl;
    P.ec_slot[ecl] = slot;
    P.ec_rep[ecl] = s;
    P.ec_len[ecl] = len;
  }
  return slot;
}

__device__ __forceinline__ Mix4 shfl_up_mix(const Mix4& m, int d) {
  Mix4 r;
  r.a = __shfl_up_sync(ECB_FULL, m.a, d);
  r.b = __shfl_up_sync(ECB_FULL, m.b, d);
  r.c = __shfl_up_sync(ECB_FULL, m.c, d);
  r.d = __shfl_up_sync(ECB_FULL, m.d, d);
  return r;
}

__device__ __forceinline__ Mix4 shfl_mix(const Mix4& m, int src) {
  Mix4 r;
  r.a = __shfl_sync(ECB_FULL, m.a, src);
  r.b = __shfl_sync(ECB_FULL, m.b, src);
  r.c = __shfl_sync(ECB_FULL, m.c, src);
  r.d = __shfl_sync(ECB_FULL, m.d, src);
  return r;
}

struct GroupSmem {
  // two-stage ring of the three columns of a tile (filled by TMA bulk copies)
  alignas(128) int32_t col[2][3][ECB_TILE];
  // queue of closed reads of the current tile
  alignas(16) uint4 qkey[ECB_TILE];
  u32 qpos[ECB_TILE];
  u32 qlen[ECB_TILE];
  // hot-EC cache of this CTA
  alignas(8) u64 c_lo[ECB_CACHE];
  u64 c_hi[ECB_CACHE];
  u64 c_replen[ECB_CACHE];  // len << 32 | offset of one read with this key
  u32 c_cnt[ECB_CACHE];
  u32 c_first[ECB_CACHE];  // len << 32 | offset of one read with this key
  // per-warp summaries
  int w_last_head[ECB_WARPS];
  u32 w_flag[ECB_WARPS];
  Mix4 w_sum[ECB_WARPS];
  alignas(8) u64 bar[2];
};

template <bool WITH_CELLS>
__global__ void __launch_bounds__(ECB_TILE_THREADS, 3) ecb_group_insert_kernel(const GroupParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  GroupSmem& S = *reinterpret_cast<GroupSmem*>(smem_raw);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 lt_mask = (1u << lane) - 1u;
  const int n = P.n;
  const long long cb64 = (long long)blockIdx.x * P.chunk_len;
  if (cb64 >= n) return;
  const int cb = (int)cb64;
  const int ce = min(cb + P.chunk_len, n);

  for (int i = tid; i < ECB_CACHE; i += ECB_TILE_THREADS) {
    S.c_lo[i] = ~0ull;
    S.c_hi[i] = ~0ull;
    S.c_cnt[i] = 0u;
    S.c_first[i] = 0xFFFFFFFFu;
  }
  if (tid == 0) {
    mbar_init(&S.bar[0], 1);
    mbar_init(&S.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0 && cb + ECB_TILE <= n) {
    mbar_expect_tx(&S.bar[0], 3 * ECB_TILE * 4);
    bulk_g2s(S.col[0][0], P.rg + cb, ECB_TILE * 4, &S.bar[0]);
    bulk_g2s(S.col[0][1], P.tg + cb, ECB_TILE * 4, &S.bar[0]);
    bulk_g2s(S.col[0][2], P.hp + cb, ECB_TILE * 4, &S.bar[0]);
  }

  int carry_head = -1;           // latest read start seen in [cb, current tile)
  Mix4 carry_sum = mix_zero();   // key contributions of the read that is open at the tile boundary
  u32 reads_counted = 0;

  for (int tile = 0;; ++tile) {
    const int tile_base = cb + tile * ECB_TILE;
    const int stg = tile & 1;
    const int wbase = tile_base + warp * ECB_WARP_SPAN;

    // ---- stage in: wait for the TMA copies of this tile (or fill a partial last tile by hand) ------
    if (tile_base + ECB_TILE <= n) {
      mbar_wait(&S.bar[stg], (u32)(tile >> 1) & 1u);
    } else {
      for (int i = tid; i < ECB_TILE; i += ECB_TILE_THREADS) {
        const int pos = tile_base + i;
        const bool in = pos < n;
        S.col[stg][0][i] = in ? P.rg[pos] : 0;
        S.col[stg][1][i] = in ? P.tg[pos] : 0;
        S.col[stg][2][i] = in ? P.hp[pos] : 0;
      }
      __syncthreads();
    }
    // prefetch the next tile into the other stage (its last readers finished before barrier (3) of
    // the previous iteration)
    if (tid == 0 && (long long)tile_base + 2 * ECB_TILE <= (long long)n) {
      const int nb = tile_base + ECB_TILE;
      mbar_expect_tx(&S.bar[stg ^ 1], 3 * ECB_TILE * 4);
      bulk_g2s(S.col[stg ^ 1][0], P.rg + nb, ECB_TILE * 4, &S.bar[stg ^ 1]);
      bulk_g2s(S.col[stg ^ 1][1], P.tg + nb, ECB_TILE * 4, &S.bar[stg ^ 1]);
      bulk_g2s(S.col[stg ^ 1][2], P.hp + nb, ECB_TILE * 4, &S.bar[stg ^ 1]);
    }

    const int32_t* s_rg = S.col[stg][0];
    const int32_t* s_tg = S.col[stg][1];
    const int32_t* s_hp = S.col[stg][2];

    // ---- head flags: lane = consecutive alignment, 4 rows per warp ---------------------------------
    int rgv[ECB_ROWS];
    u32 code[ECB_ROWS];
    u32 hb[ECB_ROWS];      // ballot of head flags per row
    bool range_bad = false;
    int before0 = 0;       // read_group of the alignment right before this warp's span
    if (lane == 0 && wbase > 0 && wbase - 1 < n) before0 = (warp > 0) ? s_rg[warp * ECB_WARP_SPAN - 1] : P.rg[wbase - 1];
#pragma unroll
    for (int r = 0; r < ECB_ROWS; ++r) {
      const int li = warp * ECB_WARP_SPAN + r * 32 + lane;
      const int pos = tile_base + li;
      rgv[r] = s_rg[li];
      const int t = s_tg[li], h = s_hp[li];
      const bool valid = pos < n;
      code[r] = valid ? ecb_code(t, h) : 0xFFFFFFFFu;
      range_bad |= valid && ((u32)t >= (u32)P.n_targets || (u32)h >= (u32)P.n_haps);
      int prev = __shfl_up_sync(ECB_FULL, rgv[r], 1);
      const int prev_row_last = __shfl_sync(ECB_FULL, r > 0 ? rgv[r - 1] : 0, 31);
      if (lane == 0) prev = (r == 0) ? before0 : prev_row_last;
      const bool hd = (pos <= n) && (pos == n || pos == 0 || rgv[r] != prev);  // n = virtual closing head
      hb[r] = __ballot_sync(ECB_FULL, hd);
    }
    if (range_bad) atomicOr(&P.ctr->error, ECB_DEVERR_VALUE_RANGE);
    int lh = -1;  // latest head inside this warp's span (warp-uniform)
#pragma unroll
    for (int r = 0; r < ECB_ROWS; ++r)
      if (hb[r]) lh = wbase + r * 32 + 31 - __clz(hb[r]);
    if (lane == 0) S.w_last_head[warp] = lh;
    // head flag of the position right after this warp's span (needed by lane 31 of the last row)
    bool next_span_head = false;
    if (lane == 31) {
      const int pos = wbase + ECB_WARP_SPAN;
      if (pos == n) next_span_head = true;
      else if (pos < n)
        next_span_head = ((warp < ECB_WARPS - 1) ? s_rg[(warp + 1) * ECB_WARP_SPAN] : P.rg[pos]) != rgv[ECB_ROWS - 1];
    }
    __syncthreads();  // (1) per-warp head summaries visible

    int open_st = carry_head;  // start of the read that is open at the beginning of this warp's span
    int tile_last_head = carry_head;
#pragma unroll
    for (int w = 0; w < ECB_WARPS; ++w) {
      const int v = S.w_last_head[w];
      if (w < warp) open_st = max(open_st, v);
      tile_last_head = max(tile_last_head, v);
    }

    // ---- per row: read start, duplicates, per-read sums --------------------------------------------
    Mix4 sum[ECB_ROWS];
    int st[ECB_ROWS];
    u32 pushb[ECB_ROWS];   // ballot: lane closes an owned read in this row
#pragma unroll
    for (int r = 0; r < ECB_ROWS; ++r) {
      const int rowbase = wbase + r * 32;
      const int pos = rowbase + lane;
      const bool valid = pos < n;
      const u32 m = hb[r] & (lt_mask | (1u << lane));
      st[r] = m ? rowbase + 31 - __clz(m) : open_st;
      const bool owned = st[r] >= 0 && st[r] < ce;
      // duplicates of (read, element) inside the row
      const u64 k64 = valid ? (((u64)(u32)st[r] << 32) | code[r]) : (0xFFFFFFFF00000000ull | (u32)lane);
      const u32 dm = __match_any_sync(ECB_FULL, k64);
      bool contrib = valid && owned && (lane == __ffs(dm) - 1);
      if (contrib && st[r] < rowbase) {  // the read began before this row: look at its earlier part
        for (int j = rowbase - 1; j >= st[r]; --j) {
          const int rel = j - tile_base;
          const u32 cj = rel >= 0 ? ecb_code(s_tg[rel], s_hp[rel]) : ecb_code(P.tg[j], P.hp[j]);
          if (cj == code[r]) {
            contrib = false;
            break;
          }
        }
      }
      Mix4 X = ecb_mix(code[r]);
      if (!contrib) X = mix_zero();
      // segmented inclusive scan over the lanes of each read (segments = reads; a read continuing
      // from the previous row forms the segment that starts at lane 0).  Only as many doubling
      // steps as the longest segment of the row needs (warp-uniform, from the head ballot).
      const int seg0 = max(st[r] - rowbase, 0);
      const u32 x1 = ~(hb[r] | 1u);
      const u32 x2 = x1 & (x1 >> 1);
      const u32 x4 = x2 & (x2 >> 2);
      const u32 x8 = x4 & (x4 >> 4);
      const u32 x16 = x8 & (x8 >> 8);
#define ECB_SEG_STEP(D)                                   \
  {                                                       \
    const Mix4 v = shfl_up_mix(X, D);                     \
    if (lane - D >= seg0) mix_add(X, v);                  \
  }
      if (x1) ECB_SEG_STEP(1)
      if (x2) ECB_SEG_STEP(2)
      if (x4) ECB_SEG_STEP(4)
      if (x8) ECB_SEG_STEP(8)
      if (x16) ECB_SEG_STEP(16)
#undef ECB_SEG_STEP
      if (r > 0 && !(hb[r] & 1u)) {  // the read of the previous row's last lane continues here
        const Mix4 cs = shfl_mix(sum[r - 1], 31);
        if (st[r] < rowbase) mix_add(X, cs);
      }
      sum[r] = X;
      // does this lane hold the last alignment of its read?
      bool nh = ((hb[r] >> 1) >> lane) & 1u;  // head flag of lane + 1
      if (lane == 31) nh = (r < ECB_ROWS - 1) ? ((hb[r + 1] & 1u) != 0) : next_span_head;
      const bool push = valid && nh && owned && !(P.drop_last && pos == n - 1);
      pushb[r] = __ballot_sync(ECB_FULL, push);
      if (hb[r]) open_st = rowbase + 31 - __clz(hb[r]);
    }
    // tail of the span: partial sum of the read that is still open at its end (without carry-in)
    {
      const Mix4 tail = shfl_mix(sum[ECB_ROWS - 1], 31);
      if (lane == 0) {
        S.w_sum[warp] = tail;
        S.w_flag[warp] = lh >= 0 ? 1u : 0u;
      }
    }
    __syncthreads();  // (2) per-warp tails visible

    Mix4 Xc = carry_sum;  // contributions of the open read before this warp's span
    Mix4 Xt = carry_sum;  // ... before the next tile
#pragma unroll
    for (int w = 0; w < ECB_WARPS; ++w) {
      const Mix4 ws = S.w_sum[w];
      const bool wf = S.w_flag[w] != 0;
      if (w < warp) {
        if (wf) Xc = ws; else mix_add(Xc, ws);
      }
      if (wf) Xt = ws; else mix_add(Xt, ws);
    }
    // closed reads go to this warp's private queue region (no CTA-wide compaction needed)
    u32 nq_w = 0;
#pragma unroll
    for (int r = 0; r < ECB_ROWS; ++r) {
      if ((pushb[r] >> lane) & 1u) {
        Mix4 y = sum[r];
        if (st[r] < wbase) mix_add(y, Xc);  // read started before this warp's span: add what earlier warps / tiles saw
        const Key128 k = mix_to_key(y);
        const u32 q = warp * ECB_WARP_SPAN + nq_w + __popc(pushb[r] & lt_mask);
        S.qkey[q] = make_uint4((u32)k.lo, (u32)(k.lo >> 32), (u32)k.hi, (u32)(k.hi >> 32));
        S.qpos[q] = (u32)st[r];
        S.qlen[q] = (u32)(wbase + r * 32 + lane - st[r] + 1);
      }
      nq_w += __popc(pushb[r]);
    }
    carry_sum = Xt;
    carry_head = tile_last_head;
    if (lane == 0) reads_counted += nq_w;
    __syncwarp();

    // ---- insert phase: every warp drains its own queue region with full warps -----------------------
    for (u32 q0 = 0; q0 < nq_w; q0 += 32) {
      const u32 qi = q0 + lane;
      if (qi < nq_w) {
        const u32 q = warp * ECB_WARP_SPAN + qi;
        const uint4 kq = S.qkey[q];
        const u32 s = S.qpos[q];
        const u32 len = S.qlen[q];
        Key128 key;
        key.lo = ((u64)kq.y << 32) | kq.x;
        key.hi = ((u64)kq.w << 32) | kq.z;
        bool cached = false;
        u32 cs = 0;
        if (!WITH_CELLS && P.use_cache && key.lo != ~0ull) {
          cs = (kq.x ^ (kq.z >> 7)) & (ECB_CACHE - 1);
          const u64 lo = *reinterpret_cast<volatile u64*>(&S.c_lo[cs]);
          bool lo_ok = lo == key.lo;
          if (!lo_ok && lo == ~0ull) {
            const u64 plo = atomicCAS(&S.c_lo[cs], ~0ull, key.lo);
            lo_ok = (plo == ~0ull || plo == key.lo);
          }
          if (lo_ok) {
            const u64 hi = *reinterpret_cast<volatile u64*>(&S.c_hi[cs]);
            cached = hi == key.hi;
            if (!cached && hi == ~0ull) {
              const u64 phi = atomicCAS(&S.c_hi[cs], ~0ull, key.hi);
              cached = (phi == ~0ull || phi == key.hi);
              // exactly one thread installs the high half: its read becomes the key's representative
              if (phi == ~0ull) S.c_replen[cs] = ((u64)len << 32) | s;
            }
          }
        }
        if (cached) {
          atomicAdd(&S.c_cnt[cs], 1u);
          if (s < *reinterpret_cast<volatile u32*>(&S.c_first[cs])) atomicMin(&S.c_first[cs], s);
        } else {
          const u32 slot = global_upsert(P, key, 1u, s, s, len);
          if (slot == ECB_NONE) {
            atomicOr(&P.overflow_bits[s >> 5], 1u << (s & 31));
            atomicAdd(&P.ctr->n_overflow, 1u);
          } else if (WITH_CELLS) {
            triple_upsert(P, slot, (u32)P.cell[s], P.order_base + s);
          }
        }
      }
    }

    // ---- continue? ----------------------------------------------------------------------------------
    const long long tile_end = (long long)tile_base + ECB_TILE;
    bool stop = tile_end > n;                                  // the virtual head at n was in this tile
    stop = stop || (tile_end >= ce && (carry_head < 0 || carry_head >= ce));  // no owned read is still open
    if (stop) {
      // a TMA prefetch of the next tile may still be in flight: it must land before the CTA retires
      if ((long long)tile_base + 2 * ECB_TILE <= (long long)n) mbar_wait(&S.bar[stg ^ 1], (u32)((tile + 1) >> 1) & 1u);
      break;
    }
  }

  // ---- flush the hot-EC cache into the HBM table -------------------------------------------------------
  __syncthreads();
  if (!WITH_CELLS && P.use_cache) {
    for (int i = tid; i < ECB_CACHE; i += ECB_TILE_THREADS) {
      const u32 cnt = S.c_cnt[i];
      if (cnt) {
        const Key128 key{S.c_lo[i], S.c_hi[i]};
        const u64 rl = S.c_replen[i];
        const u32 slot = global_upsert(P, key, cnt, S.c_first[i], (u32)rl, (u32)(rl >> 32));
        if (slot == ECB_NONE) {  // table too full: park the entry, the host grows the table and replays it
          const u32 si = atomicAdd(&P.ctr->n_spill, 1u);
          P.spill[si] = EcbSpill{key.lo, key.hi, cnt, S.c_first[i], (u32)rl, (u32)(rl >> 32)};
        }
      }
    }
  }
  if (lane == 0 && reads_counted) atomicAdd(&P.ctr->n_reads, (u64)reads_counted);
}

// Key of the read that starts at offset s, computed serially (replay / verification path).
__device__ inline Key128 ecb_serial_read_key(const int32_t* rg, const int32_t* tg, const int32_t* hp, int n,
                                             int s, int* len_out) {
  const int my = rg[s];
  Mix4 sum = mix_zero();
  int j = s;
  for (; j < n && rg[j] == my; ++j) {
    const u32 cj = ecb_code(tg[j], hp[j]);
    bool dup = false;
    for (int m = s; m < j; ++m)
      if (ecb_code(tg[m], hp[m]) == cj) {
        dup = true;
        break;
      }
    if (!dup) mix_add(sum, ecb_mix(cj));
  }
  if (len_out) *len_out = j - s;
  return mix_to_key(sum);
}

// Replay the reads flagged in overflow_bits after the table has been grown.  One thread per flagged
// read (rare path).  Bits of reads that went in are cleared; n_overflow counts the ones that did not.
template <bool WITH_CELLS>
__global__ void __launch_bounds__(256) ecb_replay_kernel(const GroupParams P) {
  const int n_words = (P.n + 31) >> 5;
  for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += gridDim.x * blockDim.x) {
    u32 bits = P.overflow_bits[w];
    while (bits) {
      const int b = __ffs(bits) - 1;
      bits &= bits - 1;
      const int s = (w << 5) + b;
      int len = 0;
      const Key128 key = ecb_serial_read_key(P.rg, P.tg, P.hp, P.n, s, &len);
      const u32 slot = global_upsert(P, key, 1u, (u32)s, (u32)s, (u32)len);
      if (slot == ECB_NONE) {
        atomicAdd(&P.ctr->n_overflow, 1u);
        continue;
      }
      if (WITH_CELLS) triple_upsert(P, slot, (u32)P.cell[s], P.order_base + (u32)s);
      atomicAnd(&P.overflow_bits[w], ~(1u << b));
    }
  }
}

// Replay parked hot-cache entries after the table has been grown.  Entries that still do not fit are
// compacted to the front of the list (n_spill counts them).
__global__ void __launch_bounds__(256) ecb_spill_replay_kernel(const GroupParams P, const EcbSpill* __restrict__ in,
                                                               u32 n_in) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n_in; i += gridDim.x * blockDim.x) {
    const EcbSpill e = in[i];
    const u32 slot = global_upsert(P, Key128{e.lo, e.hi}, e.count, e.first, e.rep, e.len);
    if (slot == ECB_NONE) {
      const u32 si = atomicAdd(&P.ctr->n_spill, 1u);
      P.spill[si] = e;
    }
  }
}

// Move every entry of `old_table` into `new_table` (capacity change).  remap (optional) records
// old slot -> new slot so that tables keyed by EC slot can be rewritten.
__global__ void __launch_bounds__(256) ecb_rehash_kernel(const EcbEntry* __restrict__ old_table, u32 old_slots,
                                                         EcbEntry* new_table, u32 new_mask, u32* ec_slot,
                                                         u32* remap, EcbCounters* ctr) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < old_slots; i += gridDim.x * blockDim.x) {
    Key128 k;
    u64 first;
    u32 cm1, aux;
    load_entry_cg(old_table + i, k, first, cm1, aux);
    if (remap) remap[i] = ECB_NONE;
    if (key_empty(k)) continue;
    u32 slot = key_slot_hash(k) & new_mask;
    for (;;) {
      Key128 old = atomic_cas128(new_table + slot, Key128{~0ull, ~0ull}, k);
      if (key_empty(old)) break;
      slot = (slot + 1) & new_mask;
    }
    EcbEntry* e = new_table + slot;
    e->first = first;
    e->countm1 = cm1;
    e->aux = aux;
    if (ec_slot && aux != ECB_NONE) ec_slot[aux] = slot;
    if (remap) remap[i] = slot;
  }
}

// Rewrite the EC-slot half of every triple key after an EC-table rehash and move the entry into a
// fresh triple table.
__global__ void __launch_bounds__(256) ecb_triple_remap_kernel(const EcbEntry* __restrict__ old_table, u32 old_slots,
                                                               EcbEntry* new_table, u32 new_mask,
                                                               const u32* __restrict__ remap) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < old_slots; i += gridDim.x * blockDim.x) {
    Key128 k;
    u64 first;
    u32 cm1, aux;
    load_entry_cg(old_table + i, k, first, cm1, aux);
    if (key_empty(k)) continue;
    if (remap) {
      const u32 old_slot = (u32)(k.lo >> 32);
      k.lo = ((u64)remap[old_slot] << 32) | (k.lo & 0xFFFFFFFFull);
    }
    u32 slot = key_slot_hash(k) & new_mask;
    for (;;) {
      Key128 old = atomic_cas128(new_table + slot, Key128{~0ull, ~0ull}, k);
      if (key_empty(old)) break;
      slot = (slot + 1) & new_mask;
    }
    EcbEntry* e = new_table + slot;
    e->first = first;
    e->countm1 = cm1;
    e->aux = aux;
  }
}
