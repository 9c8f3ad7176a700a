// ecb_group.cuh — the grouping + hash-insert kernel (the hot path), the overflow replay kernel and
// the table rehash kernel.
//
// What it replaces: alntools/bam_utils.py:258-344 (per-alignment loop: group consecutive alignments
// by read, collapse duplicate tids, ec[key] += 1) and the ordering half of :680-698 (EC id = rank of
// the key's first occurrence), on int32 columns.
//
// Shape of the kernel (HBM-bound integer work, no tensor cores):
//   * one CTA per contiguous chunk of the alignment stream (grid = resident CTAs x 148 SMs); the CTA
//     walks its chunk in tiles of 1024 alignments, 4 consecutive alignments per thread so each column
//     is read with one coalesced 128-bit load per thread; the next tile's loads are issued before the
//     current tile is processed;
//   * a read is owned by the CTA in whose chunk it STARTS; the owner runs past its chunk end until
//     the read closes, the next CTA skips the leading partial read;
//   * read boundaries: head flags from read_group, a block-wide max-scan gives every alignment the
//     position of its read's first alignment;
//   * duplicates inside a read (same (target, haplotype) twice) are found by a back-scan over the
//     codes staged in a two-tile shared-memory ring;
//   * the read key is the lane-wise sum of the 128-bit mixes of its distinct elements (commutative
//     set hash), reduced with a segmented block scan, so no per-read sort is needed;
//   * closed reads are compacted into a shared-memory queue so the latency-bound insert phase runs
//     with full warps; equal keys inside a warp are combined (__match_any_sync) before touching HBM;
//   * insert: one 256-bit sector load per probe, a 128-bit atomicCAS only when the slot looks empty,
//     then RED.ADD on the count and, only if it can lower it, atomicMin on the first-occurrence key.
#pragma once
#include "ecb_common.cuh"

struct GroupParams {
  const int32_t* rg;
  const int32_t* tg;
  const int32_t* hp;
  const int32_t* cell;
  int n;           // alignments in this push
  int chunk_len;   // alignments per CTA, multiple of ECB_TILE
  u64 order_base;
  int drop_last;
  int warp_aggregate;
  int n_targets;
  int n_haps;
  EcbEntry* table;
  u32 mask;
  u32* ec_slot;    // [capacity] provisional id -> table slot
  u32* ec_rep;     // [capacity] provisional id -> offset (in this push) of the claiming read
  EcbCounters* ctr;
  u32* overflow_bits;  // [ceil(n/32)] reads that must be replayed after a table growth
  EcbEntry* ttable;    // (file, EC slot, cell) table, with cells only
  u32 tmask;
  u32 push_id;
};

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

// Insert one (file, EC, cell) occurrence.  The triple table is sized so that it cannot fill up
// (see ensure_triple_capacity); running out of probes is reported as a device error.
__device__ __forceinline__ void triple_upsert(const GroupParams& P, u32 slot, u32 cell, u64 pos) {
  bool claimed;
  u64 first_seen;
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

__device__ __forceinline__ Mix4 shfl_up_mix(const Mix4& m, int d) {
  Mix4 r;
  r.a = __shfl_up_sync(ECB_FULL, m.a, d);
  r.b = __shfl_up_sync(ECB_FULL, m.b, d);
  r.c = __shfl_up_sync(ECB_FULL, m.c, d);
  r.d = __shfl_up_sync(ECB_FULL, m.d, d);
  return r;
}

template <bool WITH_CELLS>
__global__ void __launch_bounds__(ECB_TILE_THREADS, 2) ecb_group_insert_kernel(const GroupParams P) {
  __shared__ __align__(16) u32 s_codes[2][ECB_TILE];
  __shared__ __align__(16) uint4 s_qkey[ECB_TILE];
  __shared__ u32 s_qpos[ECB_TILE];
  __shared__ int s_warp_last_head[ECB_WARPS];
  __shared__ int s_warp_first_head[ECB_WARPS];
  __shared__ u32 s_warp_nq[ECB_WARPS];
  __shared__ u32 s_warp_flag[ECB_WARPS];
  __shared__ Mix4 s_warp_sum[ECB_WARPS];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = P.n;
  const long long cb64 = (long long)blockIdx.x * P.chunk_len;
  if (cb64 >= n) return;
  const int cb = (int)cb64;
  const int ce = min(cb + P.chunk_len, n);

  int carry_head = -1;           // latest read start seen in [cb, current tile)
  Mix4 carry_sum = mix_zero();   // key contributions of the read that is open at the tile boundary
  u64 reads_counted = 0;

  // software prefetch: registers for the tile being processed are filled one iteration ahead
  int4 nr = make_int4(0, 0, 0, 0), nt = nr, nh = nr;
  bool have_next = false;
  if (cb + ECB_TILE <= n) {
    nr = ld_stream_int4(P.rg + cb + tid * ECB_ITEMS);
    nt = ld_stream_int4(P.tg + cb + tid * ECB_ITEMS);
    nh = ld_stream_int4(P.hp + cb + tid * ECB_ITEMS);
    have_next = true;
  }

  for (int tile = 0;; ++tile) {
    const int tile_base = cb + tile * ECB_TILE;
    const int g0 = tile_base + tid * ECB_ITEMS;
    const int buf = tile & 1;

    int r[ECB_ITEMS], t[ECB_ITEMS], h[ECB_ITEMS];
    if (have_next) {
      r[0] = nr.x; r[1] = nr.y; r[2] = nr.z; r[3] = nr.w;
      t[0] = nt.x; t[1] = nt.y; t[2] = nt.z; t[3] = nt.w;
      h[0] = nh.x; h[1] = nh.y; h[2] = nh.z; h[3] = nh.w;
    } else {
#pragma unroll
      for (int i = 0; i < ECB_ITEMS; ++i) {
        const int pos = g0 + i;
        const bool in = pos < n;
        r[i] = in ? P.rg[pos] : 0;
        t[i] = in ? P.tg[pos] : 0;
        h[i] = in ? P.hp[pos] : 0;
      }
    }
    {
      const int nb = tile_base + ECB_TILE;
      have_next = (long long)nb + ECB_TILE <= (long long)n;
      if (have_next) {
        nr = ld_stream_int4(P.rg + nb + tid * ECB_ITEMS);
        nt = ld_stream_int4(P.tg + nb + tid * ECB_ITEMS);
        nh = ld_stream_int4(P.hp + nb + tid * ECB_ITEMS);
      }
    }

    // ---- head flags (position n is a virtual head that closes the last read) -------------------
    int rprev = __shfl_up_sync(ECB_FULL, r[ECB_ITEMS - 1], 1);
    if (lane == 0) rprev = (g0 > 0 && g0 - 1 < n) ? P.rg[g0 - 1] : 0;
    bool hd[ECB_ITEMS], valid[ECB_ITEMS];
    u32 code[ECB_ITEMS];
    bool range_bad = false;
#pragma unroll
    for (int i = 0; i < ECB_ITEMS; ++i) {
      const int pos = g0 + i;
      valid[i] = pos < n;
      const int before = (i == 0) ? rprev : r[i - 1];
      hd[i] = (pos <= n) && (pos == n || pos == 0 || r[i] != before);
      code[i] = valid[i] ? ecb_code(t[i], h[i]) : 0xFFFFFFFFu;
      range_bad |= valid[i] && ((u32)t[i] >= (u32)P.n_targets || (u32)h[i] >= (u32)P.n_haps);
    }
    if (range_bad) atomicOr(&P.ctr->error, ECB_DEVERR_VALUE_RANGE);
    *reinterpret_cast<uint4*>(&s_codes[buf][tid * ECB_ITEMS]) = make_uint4(code[0], code[1], code[2], code[3]);

    // ---- max-scan of head positions -------------------------------------------------------------
    int lh = -1;
#pragma unroll
    for (int i = 0; i < ECB_ITEMS; ++i)
      if (hd[i]) lh = g0 + i;
    int inc_head = lh;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int o = __shfl_up_sync(ECB_FULL, inc_head, d);
      if (lane >= d) inc_head = max(inc_head, o);
    }
    if (lane == 31) s_warp_last_head[warp] = inc_head;
    if (lane == 0) s_warp_first_head[warp] = hd[0] ? 1 : 0;
    __syncthreads();  // (1) codes + per-warp head summaries visible

    int excl_head = carry_head;
    int tile_last_head = carry_head;
#pragma unroll
    for (int w = 0; w < ECB_WARPS; ++w) {
      const int v = s_warp_last_head[w];
      if (w < warp) excl_head = max(excl_head, v);
      tile_last_head = max(tile_last_head, v);
    }
    {
      const int up = __shfl_up_sync(ECB_FULL, inc_head, 1);
      if (lane > 0) excl_head = max(excl_head, up);
    }
    int st[ECB_ITEMS];
    {
      int cur = excl_head;
#pragma unroll
      for (int i = 0; i < ECB_ITEMS; ++i) {
        if (hd[i]) cur = g0 + i;
        st[i] = cur;
      }
    }
    // head flag of the position right after this thread's items
    bool next_head;
    {
      const int v = __shfl_down_sync(ECB_FULL, hd[0] ? 1 : 0, 1);
      if (lane < 31) {
        next_head = v != 0;
      } else if (warp < ECB_WARPS - 1) {
        next_head = s_warp_first_head[warp + 1] != 0;
      } else {
        const int pos = tile_base + ECB_TILE;
        next_head = (pos > n) ? false : (pos == n) ? true : (P.rg[pos] != r[ECB_ITEMS - 1]);
      }
    }

    // ---- duplicate detection + contributions ------------------------------------------------------
    Mix4 c[ECB_ITEMS];
    bool owned[ECB_ITEMS];
#pragma unroll
    for (int i = 0; i < ECB_ITEMS; ++i) {
      const int pos = g0 + i;
      owned[i] = st[i] >= 0 && st[i] < ce;
      c[i] = mix_zero();
      if (valid[i] && owned[i]) {
        bool dup = false;
        for (int j = pos - 1; j >= st[i]; --j) {
          const int rel = j - tile_base;
          u32 cj;
          if (rel >= 0) cj = s_codes[buf][rel];
          else if (rel >= -ECB_TILE) cj = s_codes[buf ^ 1][rel + ECB_TILE];
          else cj = ecb_code(P.tg[j], P.hp[j]);
          if (cj == code[i]) {
            dup = true;
            break;
          }
        }
        if (!dup) c[i] = ecb_mix(code[i]);
      }
    }

    // ---- segmented sum of contributions ---------------------------------------------------------
    Mix4 x[ECB_ITEMS];
    Mix4 X = mix_zero();
    bool F = false;
#pragma unroll
    for (int i = 0; i < ECB_ITEMS; ++i) {
      if (hd[i]) {
        X = mix_zero();
        F = true;
      }
      mix_add(X, c[i]);
      x[i] = X;
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      Mix4 Xp = shfl_up_mix(X, d);
      const int Fp = __shfl_up_sync(ECB_FULL, F ? 1 : 0, d);
      if (lane >= d) {
        if (!F) mix_add(X, Xp);
        F = F || (Fp != 0);
      }
    }
    if (lane == 31) {
      s_warp_sum[warp] = X;
      s_warp_flag[warp] = F ? 1u : 0u;
    }
    Mix4 Xe = shfl_up_mix(X, 1);
    bool Fe = __shfl_up_sync(ECB_FULL, F ? 1 : 0, 1) != 0;
    if (lane == 0) {
      Xe = mix_zero();
      Fe = false;
    }

    // closed reads owned by this CTA go to the insert queue
    bool push[ECB_ITEMS];
    u32 nl = 0;
#pragma unroll
    for (int i = 0; i < ECB_ITEMS; ++i) {
      const int pos = g0 + i;
      const bool last = valid[i] && (i < ECB_ITEMS - 1 ? hd[i + 1] : next_head);
      push[i] = last && owned[i] && !(P.drop_last && pos == n - 1);
      nl += push[i] ? 1u : 0u;
    }
    u32 nl_inc = nl;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      u32 o = __shfl_up_sync(ECB_FULL, nl_inc, d);
      if (lane >= d) nl_inc += o;
    }
    if (lane == 31) s_warp_nq[warp] = nl_inc;
    __syncthreads();  // (2) per-warp sums / queue counts visible

    Mix4 Xc = carry_sum;        // contributions of the open read before this warp
    Mix4 Xt = carry_sum;        // ... before the next tile
    u32 qoff = nl_inc - nl, nq_total = 0;
#pragma unroll
    for (int w = 0; w < ECB_WARPS; ++w) {
      const Mix4 ws = s_warp_sum[w];
      const bool wf = s_warp_flag[w] != 0;
      if (w < warp) {
        if (wf) Xc = ws; else mix_add(Xc, ws);
        qoff += s_warp_nq[w];
      }
      if (wf) Xt = ws; else mix_add(Xt, ws);
      nq_total += s_warp_nq[w];
    }
    Mix4 Xin = Xe;
    if (!Fe) mix_add(Xin, Xc);
    {
      bool seen_head = false;
#pragma unroll
      for (int i = 0; i < ECB_ITEMS; ++i) {
        seen_head = seen_head || hd[i];
        if (push[i]) {
          Mix4 y = x[i];
          if (!seen_head) mix_add(y, Xin);
          const Key128 k = mix_to_key(y);
          s_qkey[qoff] = make_uint4((u32)k.lo, (u32)(k.lo >> 32), (u32)k.hi, (u32)(k.hi >> 32));
          s_qpos[qoff] = (u32)st[i];
          ++qoff;
        }
      }
    }
    carry_sum = Xt;
    carry_head = tile_last_head;
    if (tid == 0) reads_counted += nq_total;
    __syncthreads();  // (3) queue visible

    // ---- insert phase: full warps over the compacted queue ----------------------------------------
    for (u32 q0 = 0; q0 < nq_total; q0 += ECB_TILE_THREADS) {
      const u32 q = q0 + tid;
      const bool act = q < nq_total;
      const u32 amask = __ballot_sync(ECB_FULL, act);
      if (act) {
        const uint4 kq = s_qkey[q];
        const u32 s = s_qpos[q];
        Key128 key;
        key.lo = ((u64)kq.y << 32) | kq.x;
        key.hi = ((u64)kq.w << 32) | kq.z;
        u32 grp = 1u << lane;
        if (P.warp_aggregate) grp = __match_any_sync(amask, key.lo) & __match_any_sync(amask, key.hi);
        const int leader = __ffs(grp) - 1;
        u32 slot = ECB_NONE;
        bool claimed = false;
        if (lane == leader) {
          u64 first_seen;
          slot = table_find_or_claim(P.table, P.mask, key, claimed, first_seen);
          if (slot != ECB_NONE) {
            EcbEntry* e = P.table + slot;
            atomicAdd(&e->countm1, (u32)__popc(grp));
            const u64 pos = P.order_base + s;  // lowest lane of the group = earliest read
            if (pos < first_seen) atomicMin(&e->first, pos);
          }
        }
        __syncwarp(amask);
        const u32 cm = __ballot_sync(amask, claimed);
        if (cm) {  // hand out provisional EC ids, one atomic per warp
          const int cl = __ffs(cm) - 1;
          u32 base = 0;
          if (lane == cl) base = atomicAdd(&P.ctr->n_ec, (u32)__popc(cm));
          base = __shfl_sync(amask, base, cl);
          if (claimed) {
            const u32 ecl = base + (u32)__popc(cm & ((1u << lane) - 1u));
            P.table[slot].aux = ecl;
            P.ec_slot[ecl] = slot;
            P.ec_rep[ecl] = s;
          }
        }
        slot = __shfl_sync(amask, slot, leader);
        if (slot == ECB_NONE) {
          atomicOr(&P.overflow_bits[s >> 5], 1u << (s & 31));
          atomicAdd(&P.ctr->n_overflow, 1u);
        } else if (WITH_CELLS) {
          triple_upsert(P, slot, (u32)P.cell[s], P.order_base + s);
        }
      }
    }

    // ---- continue? ----------------------------------------------------------------------------------
    const long long tile_end = (long long)tile_base + ECB_TILE;
    if (tile_end > n) break;                                   // the virtual head at n was in this tile
    if (tile_end >= ce && (carry_head < 0 || carry_head >= ce)) break;  // no owned read is still open
  }
  if (tid == 0 && reads_counted) atomicAdd(&P.ctr->n_reads, reads_counted);
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
      const Key128 key = ecb_serial_read_key(P.rg, P.tg, P.hp, P.n, s, nullptr);
      bool claimed;
      u64 first_seen;
      const u32 slot = table_find_or_claim(P.table, P.mask, key, claimed, first_seen);
      if (slot == ECB_NONE) {
        atomicAdd(&P.ctr->n_overflow, 1u);
        continue;
      }
      EcbEntry* e = P.table + slot;
      atomicAdd(&e->countm1, 1u);
      const u64 pos = P.order_base + (u32)s;
      if (pos < first_seen) atomicMin(&e->first, pos);
      if (claimed) {
        const u32 ecl = atomicAdd(&P.ctr->n_ec, 1u);
        e->aux = ecl;
        P.ec_slot[ecl] = slot;
        P.ec_rep[ecl] = (u32)s;
      }
      if (WITH_CELLS) triple_upsert(P, slot, (u32)P.cell[s], pos);
      atomicAnd(&P.overflow_bits[w], ~(1u << b));
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
