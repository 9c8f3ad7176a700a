// ecb_strip.cuh — the strip form of the grouping kernel (ECB_OPT_STRIP_KERNEL).
//
// Same job and same table protocol as ecb_group_insert_kernel (ecb_group.cuh; replaces
// alntools/bam_utils.py:258-344 on int32 columns), different decomposition of the stream:
//
//   * window kernel: lane = ONE alignment; read boundaries, duplicate elimination and the key sums of a
//     32-alignment window are warp collectives (about 290 warp instructions per 30 alignments);
//   * strip kernel:  lane = a STRIP of 8 consecutive alignments kept in registers, plus the 8 alignments
//     behind it as look-ahead (they are the neighbour lane's strip; the loads hit L1).  A read belongs
//     to the lane whose strip holds its FIRST alignment, wherever it ends, so there is no carry between
//     lanes, tiles or work chunks: every lane walks its 16 positions alone - head bit, duplicate test
//     against the up to 7 earlier elements of the same read (register compares), running 128-bit sum -
//     and closes the reads that start in its strip.  Reads longer than 8 alignments cannot be seen
//     whole by one lane; their starts are handed to the warp-cooperative routine of the window kernel
//     (ecb_long_read) after the tile.  A warp instruction now works on 32 x 8 alignments.
//
// Closed reads go through the same per-CTA hot-EC cache, miss queue and batched HBM-table insert as in
// the window kernel (the code below repeats that block; see the header of ecb_group.cuh).
//
// The per-lane walk is plain C++ (`__host__ __device__`): tests/native/strip_host_test.cu runs it on the CPU
// against a serial statement of the grouping rule.
#pragma once
#include "ecb_group.cuh"

// Timing experiments (never shipped; results are wrong by construction): 1 = misses are dropped instead of
// being inserted into the HBM table, 2 = closed reads are only folded into a checksum (walk cost alone).
#ifndef ECB_STRIP_EXPERIMENT
#define ECB_STRIP_EXPERIMENT 0
#endif

#define ECB_STRIP 8                        // alignments per lane
#define ECB_TILE (32 * ECB_STRIP)          // alignments per warp and tile
#define ECB_STRIP_SPAN (2 * ECB_STRIP)     // positions a lane looks at (its strip + look-ahead)

// What a lane knows about its 16 positions: element codes and a bit per position that starts a read
// (bit i: position p0 + i; the first position beyond the push counts as a start, it closes the last read).
struct StripLane {
  u32 c[ECB_STRIP_SPAN];
  u32 hm;
};

// Walking state of a lane (plain 32-bit words: the compiler keeps them in registers and predicates).
struct StripWalk {
  Mix4 sum;        // key sum of the open read
  int st;          // index (0..7) of the open read's first alignment
  u32 open;        // 1: a read that started in this strip is open
  u32 has_long;    // 1: this strip starts a read with more than 8 alignments ...
  u32 long_start;  // ... at this index of the strip
};

__host__ __device__ __forceinline__ void strip_walk_init(StripWalk& W) {
  W.sum = mix_zero();
  W.st = 0;
  W.open = 0u;
  W.has_long = 0u;
  W.long_start = 0u;
}

// Head bits and codes from 16 + 1 read-group values (rgprev = the value in front of the strip; any value
// that differs from rg[0] when the strip starts the push) and the target / haplotype columns.  Positions
// at or beyond n must hold ECB_RG_SENTINEL in rg.
__host__ __device__ __forceinline__ void strip_lane_build(StripLane& L, int rgprev, const int (&rg)[ECB_STRIP_SPAN],
                                                          const int (&tg)[ECB_STRIP_SPAN],
                                                          const int (&hp)[ECB_STRIP_SPAN]) {
  u32 hm = rg[0] != rgprev ? 1u : 0u;
#pragma unroll
  for (int i = 1; i < ECB_STRIP_SPAN; ++i) hm |= rg[i] != rg[i - 1] ? (1u << i) : 0u;
#pragma unroll
  for (int i = 0; i < ECB_STRIP_SPAN; ++i) L.c[i] = ecb_code(tg[i], hp[i]);
  L.hm = hm;
}

// One position of the walk, in two halves (I is a compile-time constant after unrolling).
// strip_walk_closes: a read that started in this strip closes in front of position I; it is the read
// [W.st, I) of the strip and W.sum is its key sum - the caller uses them before strip_walk_advance.
__host__ __device__ __forceinline__ bool strip_walk_closes(const StripLane& L, int I, const StripWalk& W) {
  return (((L.hm >> I) & W.open) & 1u) != 0u;
}

// strip_walk_advance: take position I in.  n_own = number of positions of the strip that lie inside the
// push (<= 0: none).
__host__ __device__ __forceinline__ void strip_walk_advance(const StripLane& L, int I, int n_own, StripWalk& W) {
  if ((L.hm >> I) & 1u) {
    W.open = 0u;
    if (I < ECB_STRIP && I < n_own) {
      // the read is short iff another start (or the end of the push) follows within 8 positions
      if (((L.hm >> (I + 1)) & 0xFFu) == 0u) {
        W.has_long = 1u;
        W.long_start = (u32)I;
      } else {
        W.open = 1u;
        W.st = I;
        W.sum = mix_zero();
      }
    }
  }
  if (W.open) {
    // an element counts once per read: equal codes among the up to 7 earlier positions of this read
    u32 eqm = 0u;
#pragma unroll
    for (int k = 1; k < ECB_STRIP; ++k)
      if (I - k >= 0) eqm |= L.c[I] == L.c[I - k] ? (1u << (k - 1)) : 0u;
    if ((eqm & ((1u << (I - W.st)) - 1u)) == 0u) mix_add(W.sum, ecb_mix(L.c[I]));
  }
}

#ifdef __CUDACC__

__device__ __forceinline__ int4 ld_col4(const int32_t* p, u64 pol) {
  int4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p), "l"(pol));
  return v;
}

// Hot-EC cache probe, miss queue and batched table insert for the closed reads of the warp (one per lane
// at most; `ins` false = nothing).  Warp-uniform control flow: every lane calls it.  Same protocol as
// the block at the end of the window loop of ecb_group_insert_kernel.
struct StripSmemAddr {
  u32 qk, qr, a_key, a_lock, a_cnt, a_first, a_rep, a_seen;
};

// The batched table insert is reached from 17 places of the unrolled walk: one out-of-line copy keeps the
// kernel inside the instruction cache (it runs once per 64 misses).
// Behind a reference the compiler no longer knows that the pointers of the kernel parameter block point
// to global memory and emits generic atomics WITH return value (ATOM.E instead of RED.E, plus a shared-memory
// fallback loop for the 64-bit minimum): measured, the table insert then costs 1.3 ms instead of 0.25 ms on
// cfg2.  Address-space hints on a private copy of the block restore the fire-and-forget reductions of the
// window kernel.
__device__ __forceinline__ void strip_assume_global(const GroupParams& P) {
  __builtin_assume(__isGlobal(P.table));
  __builtin_assume(__isGlobal(P.ctr));
  __builtin_assume(__isGlobal(P.ec_slot));
  __builtin_assume(__isGlobal(P.ec_rep));
  __builtin_assume(__isGlobal(P.ec_len));
  __builtin_assume(__isGlobal(P.overflow_bits));
  __builtin_assume(__isGlobal(P.ttable));
  __builtin_assume(__isGlobal(P.cell));
}

template <bool WITH_CELLS>
__device__ __noinline__ void strip_insert64(const GroupParams& P, u32 qk, u32 qr, u32 base, int lane) {
  GroupParams Q{};   // a private copy of the fields the insert uses: values the hints can attach to
  Q.table = P.table; Q.mask = P.mask; Q.order_base = P.order_base; Q.ctr = P.ctr;
  Q.ec_slot = P.ec_slot; Q.ec_rep = P.ec_rep; Q.ec_len = P.ec_len; Q.overflow_bits = P.overflow_bits;
  Q.cell = P.cell; Q.ttable = P.ttable; Q.tmask = P.tmask; Q.push_id = P.push_id;
  strip_assume_global(Q);
  insert_misses<WITH_CELLS>(Q, qk, qr, base + lane, true, base + 32 + lane, true);
}

template <bool WITH_CELLS>
__device__ __forceinline__ void strip_commit(const GroupParams& P, const StripSmemAddr& A, bool use_cache, bool ins,
                                             const uint4& key, u32 s, u32 len, u32& qn, u32 lt_mask, int lane) {
  bool miss = ins;
  if (use_cache && ins) {
    const u32 cidx = (key.y >> 7) & (ECB_CACHE - 1);
    const uint4 ck = lds128(A.a_key + cidx * 16u);
    const u32 cf = lds32(A.a_first + cidx * 4u);
    bool hit = ck.x == key.x && ck.y == key.y && ck.z == key.z && ck.w == key.w;
    if (!hit && (ck.x & ck.y & ck.z & ck.w) == 0xFFFFFFFFu) {
#if ECB_ADMIT_SECOND
      const u32 sbit = 1u << (key.z & 31u);
      const bool again = (atoms_or(A.a_seen + ((key.z >> 5) & (ECB_SEEN_WORDS - 1)) * 4u, sbit) & sbit) != 0u;
#else
      const bool again = true;
#endif
      if (again && atoms_cas(A.a_lock + cidx * 4u, 0u, 1u) == 0u) {
        sts64(A.a_rep + cidx * 8u, s, len);
        sts128(A.a_key + cidx * 16u, key);
        hit = true;
      }
    }
    if (hit) {
      reds_add(A.a_cnt + cidx * 4u, 1u);
      if (s < cf) reds_min(A.a_first + cidx * 4u, s);
      miss = false;
    }
  }
  const u32 mm = __ballot_sync(ECB_FULL, miss);
  if (mm) {
    if (miss) {
      const u32 q = qn + __popc(mm & lt_mask);
      sts128(A.qk + q * 16u, key);
      sts64(A.qr + q * 8u, s, len);
      prefetch_l2(P.table + (ec_slot_hash(key_of(key)) & P.mask));
    }
    qn += __popc(mm);
    __syncwarp();
    if (qn >= 64u) {
      qn -= 64u;
#if ECB_STRIP_EXPERIMENT != 1
      strip_insert64<WITH_CELLS>(P, A.qk, A.qr, qn, lane);
#endif
      __syncwarp();
    }
  }
}

// Queue depth per warp: the miss queue of the window kernel (96 entries), or - DENSE, 24 warps - 128
// entries that hold two stacks: misses grow from entry 0 upwards, closed reads that have not been looked
// up in the cache yet grow from entry 127 downwards.
#define ECB_DQ 128

template <int WARPS, bool DENSE>
__device__ __forceinline__ StripSmemAddr strip_smem_addr(GroupSmem& S, int warp) {
  static_assert(!DENSE || WARPS * ECB_DQ <= ECB_GWARPS * ECB_MQ, "dense queues do not fit in the miss-queue area");
  StripSmemAddr A;
  A.qk = smem_u32(&S.q_key[0][0]) + (DENSE ? (u32)warp * ECB_DQ * 16u : (u32)warp * ECB_MQ * 16u);
  A.qr = smem_u32(&S.q_rep[0][0]) + (DENSE ? (u32)warp * ECB_DQ * 8u : (u32)warp * ECB_MQ * 8u);
  A.a_key = smem_u32(S.c_key);
  A.a_lock = smem_u32(S.c_lock);
  A.a_cnt = smem_u32(S.c_cnt);
  A.a_first = smem_u32(S.c_first);
  A.a_rep = smem_u32(S.c_rep);
  A.a_seen = smem_u32(S.seen);
  return A;
}

// DENSE: closed reads are first parked (one ballot, two stores) ...
__device__ __forceinline__ void strip_stage(const StripSmemAddr& A, bool ins, const uint4& key, u32 s, u32 len, u32& cn,
                                            u32 lt_mask) {
  const u32 mm = __ballot_sync(ECB_FULL, ins);
  if (mm) {
    if (ins) {
      const u32 j = (ECB_DQ - 1u) - (cn + __popc(mm & lt_mask));
      sts128(A.qk + j * 16u, key);
      sts64(A.qr + j * 8u, s, len);
    }
    cn += __popc(mm);
    __syncwarp();
  }
}

// ... and go through the cache 32 at a time, every lane busy (`count` < 32 only for the leftovers at the
// end).  Out of line: reached from every step of the unrolled walk.  cn = parked reads after this call.
// Returns the new fill of the miss stack.
template <bool WITH_CELLS, int WARPS>
__device__ __noinline__ u32 strip_drain(const GroupParams& P, u32 cn, u32 count, u32 qn, bool use_cache) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  GroupSmem& S = *reinterpret_cast<GroupSmem*>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const StripSmemAddr A = strip_smem_addr<WARPS, true>(S, warp);
  const bool has = (u32)lane < count;
  const u32 j = (ECB_DQ - 1u) - (cn + (u32)lane);
  uint4 key = make_uint4(0u, 0u, 0u, 0u);
  uint2 r = make_uint2(0u, 0u);
  if (has) {
    key = lds128(A.qk + j * 16u);
    r = lds64(A.qr + j * 8u);
  }
  __syncwarp();
  strip_commit<WITH_CELLS>(P, A, use_cache, has, key, r.x, r.y, qn, (1u << lane) - 1u, lane);
  return qn;
}

// WARPS: warps per CTA (one CTA per SM).  32 warps leave 64 registers per thread, 24 leave 80.
// DENSE (24 warps): cache look-ups and table inserts run on full warps (see strip_stage / strip_drain).
template <bool WITH_CELLS, int WARPS, bool DENSE = false>
__global__ void __launch_bounds__(32 * WARPS, 1) ecb_group_strip_kernel(const __grid_constant__ GroupParams P) {
  constexpr int THREADS = 32 * WARPS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  GroupSmem& S = *reinterpret_cast<GroupSmem*>(smem_raw);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 lt_mask = (1u << lane) - 1u;
  const int n = P.n;
  const bool use_cache = !WITH_CELLS && P.use_cache;
  const int32_t* __restrict__ const c_rg = P.rg;
  const int32_t* __restrict__ const c_tg = P.tg;
  const int32_t* __restrict__ const c_hp = P.hp;
  const u64 col_policy = make_evict_first_policy();
  // lanes 0..23 pull the next tile into L2: 3 columns x 8 lines of 128 bytes
  const int32_t* const pf_col = (lane < 8 ? c_rg : (lane < 16 ? c_tg : c_hp)) + (lane & 7) * 32;
  const bool pf_lane = lane < 24;

  // The out-of-line routines take the parameter block by reference.  A reference to the kernel parameter
  // itself turns every field access over there into a generic load from the parameter window; a copy in
  // shared memory is read with ordinary shared-memory latency (hypothesis for the slow insert of DESIGN.md
  // 4.1b, not measured yet).
  __shared__ GroupParams SP;
  for (int i = tid; i < (int)(sizeof(GroupParams) / 4); i += THREADS)
    reinterpret_cast<u32*>(&SP)[i] = reinterpret_cast<const u32*>(&P)[i];
  if (use_cache) {
    for (int i = tid; i < ECB_CACHE; i += THREADS) {
      S.c_key[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
      S.c_lock[i] = 0u;
      S.c_cnt[i] = 0u;
      S.c_first[i] = 0xFFFFFFFFu;
    }
    for (int i = tid; i < ECB_SEEN_WORDS; i += THREADS) S.seen[i] = 0u;
  }
  __syncthreads();

  const StripSmemAddr A = strip_smem_addr<WARPS, DENSE>(S, warp);
  u32 qn = 0;             // reads parked in this warp's miss queue (warp-uniform)
  u32 cn = 0;             // DENSE: closed reads parked in front of the cache (warp-uniform)
#if ECB_STRIP_EXPERIMENT == 2
  u32 sink = 0;
#endif
  u32 reads_counted = 0;  // per lane

  for (;;) {
    u32 ci = 0;
    if (lane == 0) ci = atomicAdd(&P.ctr->chunk_next, 1u);
    ci = __shfl_sync(ECB_FULL, ci, 0);
    const long long cb64 = (long long)ci * P.chunk_len;   // chunk_len is a multiple of ECB_TILE here
    if (cb64 >= n) break;
    const int cb = (int)cb64;
    const int ce = (int)min(cb64 + P.chunk_len, (long long)n);

    for (int tb = cb; tb < ce; tb += ECB_TILE) {
      const int p0 = tb + ECB_STRIP * lane;
      if (pf_lane && tb + ECB_TILE + (lane & 7) * 32 < n) prefetch_l2(pf_col + tb + ECB_TILE);

      // ---- the lane's 16 positions ------------------------------------------------------------------
      StripLane L;
      {
        int rg[ECB_STRIP_SPAN], tg[ECB_STRIP_SPAN], hp[ECB_STRIP_SPAN];
        int rgprev = ECB_RG_SENTINEL;
        if (tb + ECB_TILE + ECB_STRIP <= n) {   // warp-uniform: every position of every lane is inside the push
#pragma unroll
          for (int q = 0; q < ECB_STRIP_SPAN / 4; ++q) {
            const int4 a = ld_col4(c_rg + p0 + 4 * q, col_policy);
            const int4 b = ld_col4(c_tg + p0 + 4 * q, col_policy);
            const int4 d = ld_col4(c_hp + p0 + 4 * q, col_policy);
            rg[4 * q] = a.x; rg[4 * q + 1] = a.y; rg[4 * q + 2] = a.z; rg[4 * q + 3] = a.w;
            tg[4 * q] = b.x; tg[4 * q + 1] = b.y; tg[4 * q + 2] = b.z; tg[4 * q + 3] = b.w;
            hp[4 * q] = d.x; hp[4 * q + 1] = d.y; hp[4 * q + 2] = d.z; hp[4 * q + 3] = d.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < ECB_STRIP_SPAN; ++i) {
            const bool in = p0 + i < n;
            rg[i] = in ? c_rg[p0 + i] : ECB_RG_SENTINEL;
            tg[i] = in ? c_tg[p0 + i] : 0;
            hp[i] = in ? c_hp[p0 + i] : 0;
          }
        }
        if (p0 > 0 && p0 <= n) rgprev = c_rg[p0 - 1];
        strip_lane_build(L, rgprev, rg, tg, hp);
      }
      const int n_own = n - p0;

      // ---- walk: own strip, then the look-ahead for as long as some lane's last read is open ---------
      StripWalk W;
      strip_walk_init(W);
#pragma unroll
      for (int i = 0; i < ECB_STRIP_SPAN; ++i) {
        if (i >= ECB_STRIP && !__any_sync(ECB_FULL, W.open != 0u)) break;
        bool ins = strip_walk_closes(L, i, W);
        if (P.drop_last && p0 + i == n) ins = false;   // the read that ends the push is not counted
        if (ins) ++reads_counted;
        const uint4 key = key_words(W.sum);
        const u32 s = (u32)(p0 + W.st), len = (u32)(i - W.st);
#if ECB_STRIP_EXPERIMENT == 2
        if (ins) sink ^= key.x ^ key.y ^ key.z ^ key.w ^ s ^ len;
        ins = false;
#endif
        if constexpr (DENSE) {
          strip_stage(A, ins, key, s, len, cn, lt_mask);
          if (cn >= 32u) {
            cn -= 32u;
            qn = strip_drain<WITH_CELLS, WARPS>(SP, cn, 32u, qn, use_cache);
          }
        } else {
          strip_commit<WITH_CELLS>(P, A, use_cache, ins, key, s, len, qn, lt_mask, lane);
        }
        strip_walk_advance(L, i, n_own, W);
      }

      // ---- reads with more than 8 alignments: warp-cooperative, one after the other -----------------
      u32 lm = __ballot_sync(ECB_FULL, W.has_long != 0u);
      if (lm) {
        bool ins = false;
        uint4 key = make_uint4(0u, 0u, 0u, 0u);
        u32 s = 0u, len = 0u;
        int k = 0;
        while (lm) {
          const int src = __ffs(lm) - 1;
          lm &= lm - 1;
          const int start = tb + ECB_STRIP * src + (int)__shfl_sync(ECB_FULL, W.long_start, src);
          const LongRead lr = ecb_long_read(c_rg, c_tg, c_hp, n, start, P.n_targets, P.n_haps, P.ctr);
          if (lane == k) {   // the k-th long read of the tile is committed by lane k
            key = lr.key;
            s = (u32)start;
            len = (u32)lr.len;
            ins = !(P.drop_last && start + lr.len == n);
          }
          ++k;
        }
        if (ins) ++reads_counted;
        if constexpr (DENSE) {
          strip_stage(A, ins, key, s, len, cn, lt_mask);
          if (cn >= 32u) {
            cn -= 32u;
            qn = strip_drain<WITH_CELLS, WARPS>(SP, cn, 32u, qn, use_cache);
          }
        } else {
          strip_commit<WITH_CELLS>(P, A, use_cache, ins, key, s, len, qn, lt_mask, lane);
        }
      }
    }
  }

  // ---- leftovers of the queues, then the cache goes into the HBM table --------------------------------
  if constexpr (DENSE) if (cn) qn = strip_drain<WITH_CELLS, WARPS>(SP, 0u, cn, qn, use_cache);
#if ECB_STRIP_EXPERIMENT == 1
  qn = 0;
#elif ECB_STRIP_EXPERIMENT == 2
  if (sink == 0x12345678u) atomicAdd(&P.ctr->scratch[7], 1u);
#endif
  if (qn) insert_misses<WITH_CELLS>(P, A.qk, A.qr, lane, (u32)lane < qn, lane + 32, (u32)lane + 32u < qn);
  reads_counted = __reduce_add_sync(ECB_FULL, reads_counted);
  if (lane == 0 && reads_counted) atomicAdd(&P.ctr->n_reads, (u64)reads_counted);
  __syncthreads();
  if (use_cache) {
    for (int i = tid; i < ECB_CACHE; i += THREADS) {
      const u32 cnt = S.c_cnt[i];
      if (cnt) {
        const Key128 key = key_of(S.c_key[i]);
        const u32 first = S.c_first[i];
        const uint2 rep = S.c_rep[i];
        const u32 slot = global_upsert(P, key, cnt, first, rep.x, rep.y);
        if (slot == ECB_NONE) {  // table too full: park the entry, the host grows the table and replays it
          const u32 si = atomicAdd(&P.ctr->n_spill, 1u);
          P.spill[si] = EcbSpill{key.lo, key.hi, cnt, first, rep.x, rep.y};
        }
      }
    }
  }
}

#endif  // __CUDACC__
