// ecb_tile.cuh — the grouping + hash-insert kernel, second form: READS, not alignments, are the unit of
// the work that is done per lane.
//
// What it replaces: alntools/bam_utils.py:258-344 (per-alignment loop: group consecutive alignments by
// read, collapse duplicate tids, ec[key] += 1) and the ordering half of :680-698 (EC id = rank of the
// key's first occurrence), on int32 columns.  Same results as the window kernel of ecb_group.cuh (which
// stays for the per-cell path); about a third of its instructions per alignment.
//
// The window kernel spends most of its issue slots on cross-lane work with one ALIGNMENT per lane:
// read-start masks, shuffled duplicate tests, a segmented scan of four 32-bit sums, and then a cache
// look-up with only the lanes that close a read (14 of 32).  Here every warp runs a small pipeline over
// its chunk of the stream, and everything after the first stage has one READ per lane on full warps:
//
//   A  block of 128 alignments (one 128-bit load per column and lane): element codes -> the warp's ring in
//      shared memory, read starts (one shuffle + four ballots) -> compacted into the warp's list of heads;
//   B  32 reads at a time from the list: length = next head - head; reads are binned by length class
//      (1 | 2 | 3-4 | 5-8 | longer) into per-class queues (one ballot per class);
//   C  a class with 32 queued reads runs on a full warp, thread per read, fixed trip count: codes from the
//      ring, duplicate test by register compares, 128-bit set hash summed in registers - no shuffles;
//      reads longer than 8 alignments are taken one at a time by the whole warp from global memory
//      (match_any inside a block of 32) and parked in the lane that will commit them;
//   D  ONE commit site: hot-EC cache of the CTA (lock-free hit path), misses compacted into the warp's
//      queue (home slot prefetched into L2) and inserted into the HBM table 32 at a time.
//
// A read belongs to the chunk it starts in; the owner runs past the chunk end until the read closes.
// Nothing here carries the read length along: representative read = the first occurrence; its length is
// re-derived by the harvest kernels from read_group.
#pragma once
#include "ecb_common.cuh"
#include "ecb_group.cuh"

#ifndef ECB_T_WARPS
#define ECB_T_WARPS 24          // warps per CTA
#endif
#ifndef ECB_T_CACHE
#define ECB_T_CACHE 4096        // hot-EC cache entries per CTA (24 bytes each)
#endif
#ifndef ECB_T_RING
#define ECB_T_RING 512          // element codes a warp keeps (alignments back from its load frontier)
#endif
#define ECB_T_BLOCK 128         // alignments per stage-A block (4 per lane)
#define ECB_T_MQ 64             // per-warp miss queue; 32 are inserted at a time
#define ECB_T_QCAP 64           // per-class queue
#define ECB_T_HCAP 256          // list of heads (u16, relative to the chunk start, saturating)
#define ECB_T_CLASSES 5         // 1 | 2 | 3-4 | 5-8 | longer
#define ECB_T_THREADS (32 * ECB_T_WARPS)
#define ECB_T_MAX_CHUNK 8192    // heads are 16-bit offsets from the chunk start

// shared memory: [cache keys][cache counts][cache firsts][seen filter][per-warp state ...]
#define ECB_T_OFF_KEY 0
#define ECB_T_OFF_CNT (ECB_T_CACHE * 16)
#define ECB_T_OFF_FIRST (ECB_T_CACHE * 20)
#define ECB_T_OFF_SEEN (ECB_T_CACHE * 24)
#define ECB_T_OFF_WARP (ECB_T_CACHE * 24 + ECB_SEEN_WORDS * 4)
// per warp: [miss keys][ring][class queues][miss starts][heads][lengths of the 5-8 class]
#define ECB_T_W_MQK 0
#define ECB_T_W_RING (ECB_T_MQ * 16)
#define ECB_T_W_Q (ECB_T_W_RING + ECB_T_RING * 4)
#define ECB_T_W_MQS (ECB_T_W_Q + ECB_T_CLASSES * ECB_T_QCAP * 4)
#define ECB_T_W_HEADS (ECB_T_W_MQS + ECB_T_MQ * 4)
#define ECB_T_W_LEN (ECB_T_W_HEADS + ECB_T_HCAP * 2)
#define ECB_T_W_STRIDE (ECB_T_W_LEN + ECB_T_QCAP)
#define ECB_T_SMEM (ECB_T_OFF_WARP + ECB_T_WARPS * ECB_T_W_STRIDE)
static_assert(ECB_T_SMEM <= 232448, "shared memory of the tile kernel exceeds 227 KB");
static_assert((ECB_T_W_STRIDE & 15) == 0, "per-warp state must keep 16-byte alignment");

__device__ __forceinline__ u32 lds16(u32 a) {
  u32 v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts16(u32 a, u32 v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ u32 lds8(u32 a) {
  u32 v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts8(u32 a, u32 v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts32(u32 a, u32 v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

__device__ __forceinline__ int4 ld_col4(const int32_t* p, u64 pol) {
  int4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p), "l"(pol));
  return v;
}

// Four consecutive entries of a column from position p on; positions beyond n give `fill`.
__device__ __forceinline__ int4 tile_load4(const int32_t* __restrict__ col, int p, int n, int fill, u64 pol) {
  if (p + 3 < n) return ld_col4(col + p, pol);
  int4 v = make_int4(fill, fill, fill, fill);
  if (p < n) v.x = col[p];
  if (p + 1 < n) v.y = col[p + 1];
  if (p + 2 < n) v.z = col[p + 2];
  return v;
}

// One read per lane goes into the HBM table (entry idx of the warp's miss queue for the lanes that `has`
// one).  The whole warp calls it.  One 256-bit sector load per probe, a 128-bit compare-and-swap only when
// the slot looks empty, RED.ADD on the count, atomicMin on the first-occurrence key only when it can
// lower it; new ECs get provisional ids from one atomic per warp.
__device__ __forceinline__ void tile_insert(const GroupParams& P, u32 mqk, u32 mqs, u32 idx, bool has) {
  uint4 k4 = make_uint4(0u, 0u, 0u, 0u);
  u32 s = 0;
  if (has) {
    k4 = lds128(mqk + idx * 16u);
    s = lds32(mqs + idx * 4u);
  }
  const Key128 key = key_of(k4);
  const Key128 EMPTY{~0ull, ~0ull};
  const u32 h = ec_slot_hash(key) & P.mask;
  Key128 k = EMPTY;
  u64 f = ~0ull;
  u32 cm1, aux;
  if (has) load_entry_cg(P.table + h, k, f, cm1, aux);
  const bool eq = has && key_eq(k, key);
  const bool cas = has && !eq && key_empty(k);
  if (cas) {
    k = atomic_cas128(P.table + h, EMPTY, key);
    f = ~0ull;   // the loaded `first` belongs to the empty state
  }
  bool cl = cas && key_empty(k);
  u32 slot = (eq || cl || (cas && key_eq(k, key))) ? h : ECB_NONE;
  if (has && slot == ECB_NONE) {   // the home slot holds another key: walk on
    f = ~0ull;
    slot = table_probe_from(P.table, P.mask, key, (h + 1) & P.mask, ECB_MAX_PROBE - 1, cl, f);
  }
  if (slot != ECB_NONE) {
    EcbEntry* e = P.table + slot;
    atomicAdd(&e->countm1, 1u);
    const u64 pos = P.order_base + s;
    if (pos < f) atomicMin(&e->first, pos);
  } else if (has) {   // table too full: the host grows it and replays the flagged reads
    atomicOr(&P.overflow_bits[s >> 5], 1u << (s & 31));
    atomicAdd(&P.ctr->n_overflow, 1u);
  }
  __syncwarp();
  const u32 mc = __ballot_sync(ECB_FULL, cl);
  if (mc) {
    const int lane = threadIdx.x & 31;
    u32 base = 0;
    if (lane == 0) base = atomicAdd(&P.ctr->n_ec, (u32)__popc(mc));
    base = __shfl_sync(ECB_FULL, base, 0);
    if (cl) {
      const u32 id = base + (u32)__popc(mc & ((1u << lane) - 1u));
      P.table[slot].aux = id;
      P.ec_slot[id] = slot;
      P.ec_rep[id] = s;
    }
  }
}

// `count` reads with key `key`, the first of them at offset `first_local`, go into the HBM table (flush of
// the hot-EC cache; converged or not).
__device__ __forceinline__ u32 tile_upsert_counted(const GroupParams& P, const Key128& key, u32 count, u32 first_local) {
  bool claimed;
  u64 first_seen;
  const u32 slot = table_find_or_claim<true>(P.table, P.mask, key, claimed, first_seen);
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
    P.ec_rep[ecl] = first_local;
  }
  return slot;
}

// Key of the read that starts at s and is longer than 8 alignments: the whole warp walks it in blocks of
// 32 from global memory.  Duplicates inside a block come from one match, duplicates against the earlier
// blocks of the read from shuffled compares.  Returns the key (the same in every lane) and the length.
__device__ __noinline__ LongRead tile_long_read(const int32_t* __restrict__ rg, const int32_t* __restrict__ tg,
                                                const int32_t* __restrict__ hp, int n, int s, int n_targets,
                                                int n_haps, EcbCounters* ctr) {
  return ecb_long_read(rg, tg, hp, n, s, n_targets, n_haps, ctr);
}

__global__ void __launch_bounds__(ECB_T_THREADS, 1) ecb_group_tile_kernel(const GroupParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 lt_mask = (1u << lane) - 1u;
  const int n = P.n;
  const bool use_cache = P.use_cache != 0;
  const int32_t* __restrict__ const c_rg = P.rg;
  const int32_t* __restrict__ const c_tg = P.tg;
  const int32_t* __restrict__ const c_hp = P.hp;
  const u64 col_policy = make_evict_first_policy();

  const u32 sbase = smem_u32(smem_raw);
  const u32 a_key = sbase + ECB_T_OFF_KEY, a_cnt = sbase + ECB_T_OFF_CNT, a_first = sbase + ECB_T_OFF_FIRST;
  const u32 a_seen = sbase + ECB_T_OFF_SEEN;
  const u32 wbase = sbase + ECB_T_OFF_WARP + (u32)warp * ECB_T_W_STRIDE;
  const u32 mqk = wbase + ECB_T_W_MQK, mqs = wbase + ECB_T_W_MQS, ring = wbase + ECB_T_W_RING;
  const u32 qbase = wbase + ECB_T_W_Q, heads = wbase + ECB_T_W_HEADS, qlen = wbase + ECB_T_W_LEN;

  if (use_cache) {
    uint4* ck = reinterpret_cast<uint4*>(smem_raw + ECB_T_OFF_KEY);
    u32* cc = reinterpret_cast<u32*>(smem_raw + ECB_T_OFF_CNT);
    u32* cf = reinterpret_cast<u32*>(smem_raw + ECB_T_OFF_FIRST);
    u32* sn = reinterpret_cast<u32*>(smem_raw + ECB_T_OFF_SEEN);
    for (int i = tid; i < ECB_T_CACHE; i += ECB_T_THREADS) {
      ck[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
      cc[i] = 0u;
      cf[i] = 0xFFFFFFFFu;   // doubles as the entry's lock: the lane that swaps its offset in installs the key
    }
    for (int i = tid; i < ECB_SEEN_WORDS; i += ECB_T_THREADS) sn[i] = 0u;
  }
  __syncthreads();

  u32 qn[ECB_T_CLASSES], qh[ECB_T_CLASSES];   // fill and head index of the class queues (warp-uniform)
#pragma unroll
  for (int k = 0; k < ECB_T_CLASSES; ++k) qn[k] = qh[k] = 0u;
  u32 mqn = 0;             // reads parked in the miss queue (warp-uniform)
  u32 reads_counted = 0;   // per lane
  bool bad = false;        // a target / haplotype index outside its bounds was seen

  for (;;) {
    // ---- next chunk of the stream (dynamic: whichever warp is free takes it) ---------------------------
    u32 ci = 0;
    if (lane == 0) ci = atomicAdd(&P.ctr->chunk_next, 1u);
    ci = __shfl_sync(ECB_FULL, ci, 0);
    const long long cb64 = (long long)ci * P.chunk_len;
    if (cb64 >= n) break;
    const int cb = (int)cb64;
    const int ce = (int)min(cb64 + P.chunk_len, (long long)n);
    const u32 own_lim = (u32)(ce - cb);   // heads below this offset start reads of this chunk

    int b = cb;                 // load frontier: next block
    bool closed = false;        // a head at or beyond the chunk end has been listed (or there is nothing to own)
    u32 hn = 0, hh = 0;         // heads listed and not yet consumed, index of the oldest (warp-uniform)
    bool any_head = false;
    int carry = cb > 0 ? c_rg[cb - 1] : ECB_RG_SENTINEL;
    // the columns of the first block; from then on the next block is requested while the current one is used
    int4 r4 = tile_load4(c_rg, b + 4 * lane, n, ECB_RG_SENTINEL, col_policy);
    int4 t4 = tile_load4(c_tg, b + 4 * lane, n, 0, col_policy);
    int4 h4 = tile_load4(c_hp, b + 4 * lane, n, 0, col_policy);

    for (;;) {
      // ---- what next?  (all warp-uniform) --------------------------------------------------------------
      int sel = -1;   // 0..4: class pass, 5: bin reads, 6: load a block
#pragma unroll
      for (int k = 0; k < ECB_T_CLASSES; ++k)
        if (sel < 0 && qn[k] >= 32u) sel = k;
      if (sel < 0) {
        if (hn >= 33u) {
          sel = 5;
        } else if (!closed) {
          // the next block overwrites ring positions below lim: whatever still needs them goes first
          const int lim = b + ECB_T_BLOCK - ECB_T_RING;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (sel < 0 && qn[k] && (int)(lds32(qbase + (u32)(k * ECB_T_QCAP + qh[k]) * 4u) & 0x7FFFFFFFu) < lim) sel = k;
          if (sel < 0 && hn >= 2u && cb + (int)lds16(heads + hh * 2u) < lim) sel = 5;
          if (sel < 0) sel = 6;
        } else if (hn >= 2u) {
          sel = 5;
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (sel < 0 && qn[k]) sel = k;
          if (sel < 0) break;   // the chunk is done (long reads and misses may stay queued)
        }
      }

      if (sel == 6) {
        // ---- stage A: one block of 128 alignments ------------------------------------------------------
        const int p0 = b + 4 * lane;
        const int4 r = r4, t = t4, h = h4;
        bad |= (u32)t.x >= (u32)P.n_targets || (u32)t.y >= (u32)P.n_targets || (u32)t.z >= (u32)P.n_targets ||
               (u32)t.w >= (u32)P.n_targets || (u32)h.x >= (u32)P.n_haps || (u32)h.y >= (u32)P.n_haps ||
               (u32)h.z >= (u32)P.n_haps || (u32)h.w >= (u32)P.n_haps;
        sts128(ring + ((u32)p0 & (ECB_T_RING - 1)) * 4u,
               make_uint4(ecb_code(t.x, h.x), ecb_code(t.y, h.y), ecb_code(t.z, h.z), ecb_code(t.w, h.w)));
        int prev = __shfl_up_sync(ECB_FULL, r.w, 1);
        if (lane == 0) prev = carry;
        carry = __shfl_sync(ECB_FULL, r.w, 31);
        const bool h0 = r.x != prev, h1 = r.y != r.x, h2 = r.z != r.y, h3 = r.w != r.z;
        const u32 B0 = __ballot_sync(ECB_FULL, h0), B1 = __ballot_sync(ECB_FULL, h1);
        const u32 B2 = __ballot_sync(ECB_FULL, h2), B3 = __ballot_sync(ECB_FULL, h3);
        const u32 total = (u32)(__popc(B0) + __popc(B1) + __popc(B2) + __popc(B3));
        if (total) {
          u32 at = hh + hn + (u32)(__popc(B0 & lt_mask) + __popc(B1 & lt_mask) + __popc(B2 & lt_mask) + __popc(B3 & lt_mask));
          const u32 rel = (u32)(p0 - cb);
          if (h0) { sts16(heads + (at & (ECB_T_HCAP - 1)) * 2u, min(rel, 0xFFFFu)); ++at; }
          if (h1) { sts16(heads + (at & (ECB_T_HCAP - 1)) * 2u, min(rel + 1u, 0xFFFFu)); ++at; }
          if (h2) { sts16(heads + (at & (ECB_T_HCAP - 1)) * 2u, min(rel + 2u, 0xFFFFu)); ++at; }
          if (h3) { sts16(heads + (at & (ECB_T_HCAP - 1)) * 2u, min(rel + 3u, 0xFFFFu)); }
          hn += total;
          any_head = true;
        }
        if (b + ECB_T_BLOCK > ce) closed = (b >= ce) ? (total != 0u) : true;   // (b < ce < b + 128 only when ce == n)
        b += ECB_T_BLOCK;
        if (b >= ce && !any_head) closed = true;   // one long read of an earlier chunk covers this one
        if (!closed) {
          r4 = tile_load4(c_rg, b + 4 * lane, n, ECB_RG_SENTINEL, col_policy);
          t4 = tile_load4(c_tg, b + 4 * lane, n, 0, col_policy);
          h4 = tile_load4(c_hp, b + 4 * lane, n, 0, col_policy);
          // pull the lines two blocks further on into L2 (4 lines of 128 bytes per column and block)
          if (lane < 12) {
            const int pp = b + 2 * ECB_T_BLOCK + (lane & 3) * 32;
            if (pp < n) prefetch_l2((lane < 4 ? c_rg : (lane < 8 ? c_tg : c_hp)) + pp);
          }
        }
        __syncwarp();
        continue;
      }

      if (sel == 5) {
        // ---- stage B: up to 32 complete reads from the list of heads go to their class queues ----------
        const u32 cnt = min(32u, hn - 1u);
        const bool valid = (u32)lane < cnt;
        const u32 hs = lds16(heads + ((hh + (u32)lane) & (ECB_T_HCAP - 1)) * 2u);
        const u32 he = lds16(heads + ((hh + (u32)lane + 1u) & (ECB_T_HCAP - 1)) * 2u);
        const u32 len = he - hs;
        const u32 s = (u32)cb + hs;
        // the read that ends with the push is not counted when the caller says so (per-cell files); long
        // reads find their own end
        const bool dropped = P.drop_last && he != 0xFFFFu && (int)((u32)cb + he) == n;
        const bool own = valid && hs < own_lim;
        const bool ring_ok = own && !dropped;
        const u32 m0 = __ballot_sync(ECB_FULL, ring_ok && len == 1u);
        const u32 m1 = __ballot_sync(ECB_FULL, ring_ok && len == 2u);
        const u32 m2 = __ballot_sync(ECB_FULL, ring_ok && len - 3u < 2u);
        const u32 m3 = __ballot_sync(ECB_FULL, ring_ok && len - 5u < 4u);
        const u32 m4 = __ballot_sync(ECB_FULL, own && len > 8u);
        if (m0) {
          if ((m0 >> lane) & 1u) sts32(qbase + (u32)(0 * ECB_T_QCAP + ((qh[0] + qn[0] + __popc(m0 & lt_mask)) & (ECB_T_QCAP - 1))) * 4u, s);
          qn[0] += __popc(m0);
        }
        if (m1) {
          if ((m1 >> lane) & 1u) sts32(qbase + (u32)(1 * ECB_T_QCAP + ((qh[1] + qn[1] + __popc(m1 & lt_mask)) & (ECB_T_QCAP - 1))) * 4u, s);
          qn[1] += __popc(m1);
        }
        if (m2) {
          if ((m2 >> lane) & 1u)
            sts32(qbase + (u32)(2 * ECB_T_QCAP + ((qh[2] + qn[2] + __popc(m2 & lt_mask)) & (ECB_T_QCAP - 1))) * 4u, s | ((len - 3u) << 31));
          qn[2] += __popc(m2);
        }
        if (m3) {
          if ((m3 >> lane) & 1u) {
            const u32 at = (qh[3] + qn[3] + __popc(m3 & lt_mask)) & (ECB_T_QCAP - 1);
            sts32(qbase + (u32)(3 * ECB_T_QCAP + at) * 4u, s);
            sts8(qlen + at, len);
          }
          qn[3] += __popc(m3);
        }
        if (m4) {
          if ((m4 >> lane) & 1u) sts32(qbase + (u32)(4 * ECB_T_QCAP + ((qh[4] + qn[4] + __popc(m4 & lt_mask)) & (ECB_T_QCAP - 1))) * 4u, s);
          qn[4] += __popc(m4);
        }
        hh = (hh + cnt) & (ECB_T_HCAP - 1);
        hn -= cnt;
        __syncwarp();
        continue;
      }

      // ---- stage C: one class, up to 32 reads, one per lane ----------------------------------------------
      uint4 key = make_uint4(0u, 0u, 0u, 0u);
      u32 s = 0;
      bool has = false;
      if (sel == 0) {
        const u32 cnt = min(32u, qn[0]);
        has = (u32)lane < cnt;
        s = lds32(qbase + (u32)(0 * ECB_T_QCAP + ((qh[0] + (u32)lane) & (ECB_T_QCAP - 1))) * 4u);
        const u32 c0 = lds32(ring + (s & (ECB_T_RING - 1)) * 4u);
        key = key_words(ecb_mix(c0));
        qh[0] = (qh[0] + cnt) & (ECB_T_QCAP - 1);
        qn[0] -= cnt;
      } else if (sel == 1) {
        const u32 cnt = min(32u, qn[1]);
        has = (u32)lane < cnt;
        s = lds32(qbase + (u32)(1 * ECB_T_QCAP + ((qh[1] + (u32)lane) & (ECB_T_QCAP - 1))) * 4u);
        const u32 c0 = lds32(ring + (s & (ECB_T_RING - 1)) * 4u);
        const u32 c1 = lds32(ring + ((s + 1u) & (ECB_T_RING - 1)) * 4u);
        Mix4 X = ecb_mix(c0);
        const Mix4 Y = ecb_mix(c1);
        if (c1 != c0) mix_add(X, Y);
        key = key_words(X);
        qh[1] = (qh[1] + cnt) & (ECB_T_QCAP - 1);
        qn[1] -= cnt;
      } else if (sel == 2) {
        const u32 cnt = min(32u, qn[2]);
        has = (u32)lane < cnt;
        const u32 e = lds32(qbase + (u32)(2 * ECB_T_QCAP + ((qh[2] + (u32)lane) & (ECB_T_QCAP - 1))) * 4u);
        s = e & 0x7FFFFFFFu;
        const bool four = (e >> 31) != 0u;
        const u32 c0 = lds32(ring + (s & (ECB_T_RING - 1)) * 4u);
        const u32 c1 = lds32(ring + ((s + 1u) & (ECB_T_RING - 1)) * 4u);
        const u32 c2 = lds32(ring + ((s + 2u) & (ECB_T_RING - 1)) * 4u);
        const u32 c3 = lds32(ring + ((s + 3u) & (ECB_T_RING - 1)) * 4u);
        Mix4 X = ecb_mix(c0);
        const Mix4 Y1 = ecb_mix(c1), Y2 = ecb_mix(c2), Y3 = ecb_mix(c3);
        if (c1 != c0) mix_add(X, Y1);
        if (c2 != c0 && c2 != c1) mix_add(X, Y2);
        if (four && c3 != c0 && c3 != c1 && c3 != c2) mix_add(X, Y3);
        key = key_words(X);
        qh[2] = (qh[2] + cnt) & (ECB_T_QCAP - 1);
        qn[2] -= cnt;
      } else if (sel == 3) {
        const u32 cnt = min(32u, qn[3]);
        has = (u32)lane < cnt;
        const u32 at = (qh[3] + (u32)lane) & (ECB_T_QCAP - 1);
        s = lds32(qbase + (u32)(3 * ECB_T_QCAP + at) * 4u);
        const u32 len = has ? lds8(qlen + at) : 0u;
        u32 c[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) c[j] = lds32(ring + ((s + (u32)j) & (ECB_T_RING - 1)) * 4u);
        Mix4 X = mix_zero();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          bool take = (u32)j < len;
#pragma unroll
          for (int i = 0; i < j; ++i) take = take && c[j] != c[i];
          const Mix4 Y = ecb_mix(c[j]);
          if (take) mix_add(X, Y);
        }
        key = key_words(X);
        qh[3] = (qh[3] + cnt) & (ECB_T_QCAP - 1);
        qn[3] -= cnt;
      } else {
        // reads longer than 8 alignments: the warp takes them one at a time; lane i keeps the i-th key
        const u32 cnt = min(32u, qn[4]);
        for (u32 i = 0; i < cnt; ++i) {
          const u32 si = lds32(qbase + (u32)(4 * ECB_T_QCAP + ((qh[4] + i) & (ECB_T_QCAP - 1))) * 4u);
          const LongRead lr = tile_long_read(c_rg, c_tg, c_hp, n, (int)si, P.n_targets, P.n_haps, P.ctr);
          if ((u32)lane == i) {
            key = lr.key;
            s = si;
            has = !(P.drop_last && (int)si + lr.len == n);
          }
        }
        qh[4] = (qh[4] + cnt) & (ECB_T_QCAP - 1);
        qn[4] -= cnt;
      }

      // ---- stage D: hot-EC cache, then the miss queue ------------------------------------------------------
      bool miss = has;
      if (has) ++reads_counted;
      if (use_cache && has) {
        const u32 cidx = (key.y >> 7) & (ECB_T_CACHE - 1);
        const uint4 ck = lds128(a_key + cidx * 16u);
        u32 cf = lds32(a_first + cidx * 4u);
        bool hit = ck.x == key.x && ck.y == key.y && ck.z == key.z && ck.w == key.w;
        if (!hit && (ck.x & ck.y & ck.z & ck.w) == 0xFFFFFFFFu) {
          // empty entry.  Most keys occur once in a CTA's share of the stream and would only use up the
          // cache: a key is admitted when a read with the same hash bits has missed before.  Then one lane
          // swaps its offset into the entry's `first`, and the single 128-bit store of the key publishes
          // the entry (a reader that sees the key sees a valid `first`).
          const u32 sbit = 1u << (key.z & 31u);
          const bool again = (atoms_or(a_seen + ((key.z >> 5) & (ECB_SEEN_WORDS - 1)) * 4u, sbit) & sbit) != 0u;
          if (again && atoms_cas(a_first + cidx * 4u, 0xFFFFFFFFu, s) == 0xFFFFFFFFu) {
            sts128(a_key + cidx * 16u, key);
            hit = true;
            cf = s;
          }
        }
        if (hit) {
          reds_add(a_cnt + cidx * 4u, 1u);
          if (s < cf) reds_min(a_first + cidx * 4u, s);
          miss = false;
        }
      }
      const u32 mm = __ballot_sync(ECB_FULL, miss);
      if (mm) {
        if (miss) {
          const u32 q = mqn + __popc(mm & lt_mask);
          sts128(mqk + q * 16u, key);
          sts32(mqs + q * 4u, s);
          prefetch_l2(P.table + (ec_slot_hash(key_of(key)) & P.mask));
        }
        mqn += __popc(mm);
        __syncwarp();
        if (mqn >= 32u) {
          mqn -= 32u;
          tile_insert(P, mqk, mqs, mqn + (u32)lane, true);
          __syncwarp();
        }
      }
    }
  }

  // ---- leftovers: queued long reads, the miss queue, then the cache goes into the HBM table -------------
  while (qn[4]) {
    const u32 cnt = min(32u, qn[4]);
    uint4 key = make_uint4(0u, 0u, 0u, 0u);
    u32 s = 0;
    bool has = false;
    for (u32 i = 0; i < cnt; ++i) {
      const u32 si = lds32(qbase + (u32)(4 * ECB_T_QCAP + ((qh[4] + i) & (ECB_T_QCAP - 1))) * 4u);
      const LongRead lr = tile_long_read(c_rg, c_tg, c_hp, n, (int)si, P.n_targets, P.n_haps, P.ctr);
      if ((u32)lane == i) {
        key = lr.key;
        s = si;
        has = !(P.drop_last && (int)si + lr.len == n);
      }
    }
    qh[4] = (qh[4] + cnt) & (ECB_T_QCAP - 1);
    qn[4] -= cnt;
    // the cache is still live (other warps may be using it): these few reads go straight to the miss queue
    if (has) ++reads_counted;
    const u32 mm = __ballot_sync(ECB_FULL, has);
    if (has) {
      const u32 q = mqn + __popc(mm & lt_mask);
      sts128(mqk + q * 16u, key);
      sts32(mqs + q * 4u, s);
    }
    mqn += __popc(mm);
    __syncwarp();
    if (mqn >= 32u) {
      mqn -= 32u;
      tile_insert(P, mqk, mqs, mqn + (u32)lane, true);
      __syncwarp();
    }
  }
  if (mqn) tile_insert(P, mqk, mqs, (u32)lane, (u32)lane < mqn);
  if (bad) atomicOr(&P.ctr->error, ECB_DEVERR_VALUE_RANGE);
  reads_counted = __reduce_add_sync(ECB_FULL, reads_counted);
  if (lane == 0 && reads_counted) atomicAdd(&P.ctr->n_reads, (u64)reads_counted);
  __syncthreads();
  if (use_cache) {
    const uint4* ck = reinterpret_cast<const uint4*>(smem_raw + ECB_T_OFF_KEY);
    const u32* cc = reinterpret_cast<const u32*>(smem_raw + ECB_T_OFF_CNT);
    const u32* cf = reinterpret_cast<const u32*>(smem_raw + ECB_T_OFF_FIRST);
    for (int i = tid; i < ECB_T_CACHE; i += ECB_T_THREADS) {
      const u32 cnt = cc[i];
      if (cnt) {
        const Key128 key = key_of(ck[i]);
        const u32 first = cf[i];
        const u32 slot = tile_upsert_counted(P, key, cnt, first);
        if (slot == ECB_NONE) {  // table too full: park the entry, the host grows the table and replays it
          const u32 si = atomicAdd(&P.ctr->n_spill, 1u);
          P.spill[si] = EcbSpill{key.lo, key.hi, cnt, first, first, 0u};
        }
      }
    }
  }
}

// Length of the representative read of every EC claimed in this push (ids e0..e1): the grouping kernel
// records only where such a read starts.
__global__ void __launch_bounds__(256) ecb_rep_len_kernel(const int32_t* __restrict__ rg, int n,
                                                          const u32* __restrict__ ec_rep, u32* __restrict__ ec_len,
                                                          u32 e0, u32 e1) {
  for (u32 e = e0 + blockIdx.x * blockDim.x + threadIdx.x; e < e1; e += gridDim.x * blockDim.x) {
    const int s = (int)ec_rep[e];
    const int my = rg[s];
    int j = s + 1;
    while (j < n && rg[j] == my) ++j;
    ec_len[e] = (u32)(j - s);
  }
}
