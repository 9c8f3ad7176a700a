// ecb_api.cu — context management and the C ABI of libecb200.so (see include/ecb200.h).
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ecb200.h"
#include "ecb_common.cuh"
#include "ecb_scan.cuh"
#include "ecb_group.cuh"
#include "ecb_harvest.cuh"
#include "ecb_finalize.cuh"
#include "ecb_sort.cuh"
#include "ecb_cells.cuh"
#include "ecb_exchange.cuh"

namespace {

thread_local std::string g_create_error;

struct ecb_timer {
  cudaEvent_t a = nullptr, b = nullptr;
};

}  // namespace

struct ecb_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaMemPool_t pool = nullptr;   // private stream-ordered pool: nothing this context does changes the device's default pool
  int n_targets = 0, n_haps = 0, with_cells = 0;
  int64_t hint = 0;
  // options
  int result_on_device = 0;
  int64_t opt_table_slots = 0, opt_pair_slots = 0, opt_grid = 0, opt_chunk_len = 0;
  int use_cache = 1;
  int verify_keys = 0;
  int pageable_results = 0;
  // EC table
  DevBuf table;
  u32 table_slots = 0;
  DevBuf ec_slot, ec_rep, ec_len, row_len, row_off;  // [table_slots] u32 each
  DevBuf spill;
  bool group_attr_set = false;
  DevBuf arena;                              // uint2 pairs
  u64 arena_used = 0;
  DevBuf long_list, mid_list, big_list, count_of;
  // triple table (cells)
  DevBuf ttable;
  u32 ttable_slots = 0;
  u32 n_triples = 0;
  // counters
  EcbCounters* d_ctr = nullptr;
  EcbCounters* h_ctr = nullptr;  // pinned mirror
  u64* h_total = nullptr;        // pinned landing place of a scan total
  u32 n_ec = 0;                  // host copy after the last sync
  // staging
  DevBuf st_rg, st_tg, st_hp, st_cell;
  DevBuf st2_rg, st2_tg, st2_hp, st2_cell;   // second staging set of pipelined host pushes
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t copy_done[2] = {nullptr, nullptr};
  DevBuf overflow_bits;
  DevBuf scan_partials;  // u64 block sums + 1 total
  // push bookkeeping
  u64 min_base = ~0ull, max_end = 0;
  int64_t n_alignments = 0;
  u32 push_count = 0;
  // finalize scratch + results (device)
  DevBuf bitmap, word_rank, first_rel, ecid_of, ec_keep;
  DevBuf r_a_indptr, r_a_indices, r_a_data, r_n_indptr, r_n_indices, r_n_data, r_cell_order;
  // cells scratch
  CellScratch cells;
  // host (pinned) results
  void* h_res = nullptr;
  size_t h_res_bytes = 0;
  // stats
  ecb_stats stats{};
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool long_attr_set = false;
  // multi-GPU exchange
  DevBuf x_meta, x_rows, x_counts, x_base;
  // arena of the peer-memory exchange (plain cudaMalloc: must be exportable through CUDA IPC)
  void* xa_base = nullptr;
  int64_t xa_cap_ec = 0;
  std::vector<void*> xa_opened;
  int64_t x_part_ec[ECB_MAX_WORLD], x_part_rows[ECB_MAX_WORLD];
  u64 g_min_base = 0;
  u64 g_n_ec_total = 0;
  u32 g_n_wide = 0;
  std::string err;
};

namespace {

int fail(ecb_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf; else g_create_error = buf;
  return code;
}

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(c, ECB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                  __LINE__);                                                                       \
  } while (0)

#define CKR(expr)            \
  do {                       \
    int r_ = (expr);         \
    if (r_ != ECB_OK) return r_; \
  } while (0)

#define LAUNCH_CHECK(name)                                                                          \
  do {                                                                                              \
    c->stats.kernel_launches++;                                                                     \
    cudaError_t e_ = cudaGetLastError();                                                            \
    if (e_ != cudaSuccess) return fail(c, ECB_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e_)); \
  } while (0)

// Device memory comes from the stream-ordered allocator (a private pool per context, on the context's
// stream): growing a buffer costs no device-wide synchronisation, released blocks stay in the pool for the
// next buffer of this context, and ecb_destroy hands everything back to the driver.
int ensure(ecb_ctx* c, DevBuf& b, size_t bytes, bool preserve = false) {
  if (bytes <= b.bytes) return ECB_OK;
  size_t want = std::max(bytes, b.bytes + b.bytes / 2);
  want = (want + 255) & ~(size_t)255;
  void* np = nullptr;
  CK(cudaMallocFromPoolAsync(&np, want, c->pool, c->stream));
  if (preserve && b.p && b.bytes) CK(cudaMemcpyAsync(np, b.p, b.bytes, cudaMemcpyDeviceToDevice, c->stream));
  if (b.p) CK(cudaFreeAsync(b.p, c->stream));
  b.p = np;
  b.bytes = want;
  return ECB_OK;
}

void release(ecb_ctx* c, DevBuf& b) {
  if (b.p) cudaFreeAsync(b.p, c->stream);
  b.p = nullptr;
  b.bytes = 0;
}

u32 pow2_ceil(u64 v) {
  u64 p = 1;
  while (p < v) p <<= 1;
  return (u32)std::min<u64>(p, 1ull << 31);
}

int grid_for(u64 items, int per_block, int cap) {
  u64 g = (items + per_block - 1) / per_block;
  return (int)std::max<u64>(1, std::min<u64>(g, (u64)cap));
}

int sync_counters(ecb_ctx* c) {
  CK(cudaMemcpyAsync(c->h_ctr, c->d_ctr, sizeof(EcbCounters), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->stats.d2h_bytes += sizeof(EcbCounters);
  c->arena_used = c->h_ctr->arena_used;   // rows reserved by the harvest kernels so far
  return ECB_OK;
}

int check_device_error(ecb_ctx* c) {
  const u32 e = c->h_ctr->error;
  if (!e) return ECB_OK;
  if (e & ECB_DEVERR_VALUE_RANGE)
    return fail(c, ECB_ERR_INVALID, "target_idx/hap_idx outside [0,%d) x [0,%d)", c->n_targets, c->n_haps);
  if (e & ECB_DEVERR_READ_TOO_LONG)
    return fail(c, ECB_ERR_LIMIT, "a read has more than %d alignments", ECB_MAX_READ_ALIGNMENTS);
  if (e & ECB_DEVERR_VERIFY) return fail(c, ECB_ERR_LIMIT, "128-bit key collision detected by verification");
  return fail(c, ECB_ERR_LIMIT, "device error bits 0x%x", e);
}

// Device-wide exclusive scan (in may alias out).  total_out (host) may be NULL.
template <bool POPC>
int device_scan(ecb_ctx* c, const u32* in, u32* out, u64 n, u32 base_offset, u64* total_out) {
  const u32 n_blocks = (u32)((n + SCAN_BLOCK - 1) / SCAN_BLOCK);
  CKR(ensure(c, c->scan_partials, ((size_t)n_blocks + 2) * sizeof(u64)));
  u64* partials = (u64*)c->scan_partials.p;
  u64* d_total = partials + n_blocks;
  if (n_blocks == 0) {
    if (total_out) *total_out = 0;
    return ECB_OK;
  }
  scan_block_sums_kernel<POPC><<<n_blocks, SCAN_THREADS, 0, c->stream>>>(in, n, partials);
  LAUNCH_CHECK("scan_block_sums");
  scan_partials_kernel<<<1, 1024, 0, c->stream>>>(partials, n_blocks, d_total);
  LAUNCH_CHECK("scan_partials");
  scan_apply_kernel<POPC><<<n_blocks, SCAN_THREADS, 0, c->stream>>>(in, n, partials, out, base_offset);
  LAUNCH_CHECK("scan_apply");
  if (total_out) {
    CK(cudaMemcpyAsync(total_out, d_total, sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->stats.d2h_bytes += sizeof(u64);
  }
  return ECB_OK;
}

int alloc_ec_arrays(ecb_ctx* c, u32 slots, bool preserve) {
  CKR(ensure(c, c->ec_slot, (size_t)slots * 4, preserve));
  CKR(ensure(c, c->ec_rep, (size_t)slots * 4, preserve));
  CKR(ensure(c, c->ec_len, (size_t)slots * 4, preserve));
  CKR(ensure(c, c->row_len, (size_t)slots * 4, preserve));
  CKR(ensure(c, c->row_off, (size_t)slots * 4, preserve));
  CKR(ensure(c, c->long_list, (size_t)slots * 4, false));
  CKR(ensure(c, c->mid_list, (size_t)slots * 4, false));
  CKR(ensure(c, c->big_list, (size_t)slots * 4, false));
  return ECB_OK;
}

// first_n: alignments of the push that triggers the allocation (0 when unknown).  With a hint the
// table gets hint/8 slots (ECs are a small fraction of the alignments of a whole job); without one
// only the first push is known and a job's early part is EC-rich, so first_n/4.
int init_table(ecb_ctx* c, int64_t first_n = 0) {
  u64 want = c->opt_table_slots > 0 ? (u64)c->opt_table_slots
             : c->hint > 0          ? std::max<u64>(1u << 16, (u64)c->hint / 8)
                                    : std::max<u64>(1u << 16, (u64)first_n / 4);
  c->table_slots = std::max<u32>(1u << 10, pow2_ceil(want));
  CKR(ensure(c, c->table, (size_t)c->table_slots * sizeof(EcbEntry)));
  CK(cudaMemsetAsync(c->table.p, 0xFF, (size_t)c->table_slots * sizeof(EcbEntry), c->stream));
  CKR(alloc_ec_arrays(c, c->table_slots, false));
  if (c->with_cells) {
    u64 tw = c->opt_pair_slots > 0 ? (u64)c->opt_pair_slots : std::max<u64>(1u << 16, (u64)c->hint / 8);
    c->ttable_slots = std::max<u32>(1u << 10, pow2_ceil(tw));
    CKR(ensure(c, c->ttable, (size_t)c->ttable_slots * sizeof(EcbEntry)));
    CK(cudaMemsetAsync(c->ttable.p, 0xFF, (size_t)c->ttable_slots * sizeof(EcbEntry), c->stream));
  }
  CK(cudaMemsetAsync(c->d_ctr, 0, sizeof(EcbCounters), c->stream));
  memset(c->h_ctr, 0, sizeof(EcbCounters));
  c->n_ec = 0;
  c->n_triples = 0;
  c->arena_used = 0;
  return ECB_OK;
}

// Rebuild the (file, EC, cell) table with `new_slots` slots; remap (device, by old EC slot) may be NULL.
int rebuild_triple_table(ecb_ctx* c, u32 new_slots, const u32* remap) {
  DevBuf nt;
  CKR(ensure(c, nt, (size_t)new_slots * sizeof(EcbEntry)));
  CK(cudaMemsetAsync(nt.p, 0xFF, (size_t)new_slots * sizeof(EcbEntry), c->stream));
  ecb_triple_remap_kernel<<<grid_for(c->ttable_slots, 256, c->sm_count * 8), 256, 0, c->stream>>>(
      (const EcbEntry*)c->ttable.p, c->ttable_slots, (EcbEntry*)nt.p, new_slots - 1, remap);
  LAUNCH_CHECK("triple_remap");
  CK(cudaStreamSynchronize(c->stream));
  release(c, c->ttable);
  c->ttable = nt;
  c->ttable_slots = new_slots;
  return ECB_OK;
}

int grow_table(ecb_ctx* c, u32 new_slots) {
  DevBuf nt;
  CKR(ensure(c, nt, (size_t)new_slots * sizeof(EcbEntry)));
  CK(cudaMemsetAsync(nt.p, 0xFF, (size_t)new_slots * sizeof(EcbEntry), c->stream));
  CKR(alloc_ec_arrays(c, new_slots, true));
  DevBuf remap;
  if (c->with_cells) CKR(ensure(c, remap, (size_t)c->table_slots * 4));
  ecb_rehash_kernel<<<grid_for(c->table_slots, 256, c->sm_count * 8), 256, 0, c->stream>>>(
      (const EcbEntry*)c->table.p, c->table_slots, (EcbEntry*)nt.p, new_slots - 1, (u32*)c->ec_slot.p,
      (u32*)remap.p, c->d_ctr);
  LAUNCH_CHECK("rehash");
  CK(cudaStreamSynchronize(c->stream));
  release(c, c->table);
  c->table = nt;
  c->table_slots = new_slots;
  c->stats.table_grows++;
  if (c->with_cells) {
    CKR(rebuild_triple_table(c, c->ttable_slots, (const u32*)remap.p));
    release(c, remap);
  }
  return ECB_OK;
}

GroupParams make_group_params(ecb_ctx* c, const int32_t* rg, const int32_t* tg, const int32_t* hp,
                              const int32_t* cell, int64_t n, int64_t order_base, int drop_last, u32 push_id) {
  GroupParams P{};
  P.rg = rg; P.tg = tg; P.hp = hp; P.cell = cell;
  P.n = (int)n;
  P.order_base = (u64)order_base;
  P.drop_last = drop_last;
  P.use_cache = c->use_cache;
  P.n_targets = c->n_targets;
  P.n_haps = c->n_haps;
  P.table = (EcbEntry*)c->table.p;
  P.mask = c->table_slots - 1;
  P.ec_slot = (u32*)c->ec_slot.p;
  P.ec_rep = (u32*)c->ec_rep.p;
  P.ec_len = (u32*)c->ec_len.p;
  P.ctr = c->d_ctr;
  P.spill = (EcbSpill*)c->spill.p;
  P.overflow_bits = (u32*)c->overflow_bits.p;
  P.ttable = (EcbEntry*)c->ttable.p;
  P.tmask = c->ttable_slots ? c->ttable_slots - 1 : 0;
  P.push_id = push_id;
  return P;
}

// Opt-in shared memory of the grouping kernels.  The default kernels are prepared once per context; an
// experimental variant is prepared when it is first launched, so that nothing it needs can get in the way
// of the default path.
template <class K>
int allow_group_smem(ecb_ctx* c, K kernel) {
  CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GroupSmem)));
  return ECB_OK;
}

int group_prepare_launch(ecb_ctx* c) {
  if (!c->group_attr_set) {
    CKR(allow_group_smem(c, ecb_group_insert_kernel<true>));
    CKR(allow_group_smem(c, ecb_group_insert_kernel<false>));
    c->group_attr_set = true;
  }
  return ECB_OK;
}

// Launch geometry of the grouping kernel: one persistent CTA per SM, every warp takes chunks of
// `chunk_len` alignments from a shared counter.
void group_geometry(ecb_ctx* c, int64_t n, int* grid, int* chunk_len) {
  const int wpc = ECB_GWARPS;
  const int64_t warps = (int64_t)c->sm_count * wpc;
  int64_t cl = c->opt_chunk_len > 0 ? c->opt_chunk_len : std::min<int64_t>(4096, std::max<int64_t>(256, n / (warps * 4)));
  cl = (cl + 31) / 32 * 32;
  const int64_t chunks = (n + cl - 1) / cl;
  int64_t g = c->opt_grid > 0 ? c->opt_grid : std::min<int64_t>(c->sm_count, (chunks + wpc - 1) / wpc);
  *grid = (int)std::max<int64_t>(1, g);
  *chunk_len = (int)cl;
}

// Rows of the ECs claimed in this push (ids e0..e1).  No host round trip: the kernels reserve their rows on
// the device-side arena cursor, the lists of the longer reads are sized on the device, and all four kernels
// are launched whether their lists turn out empty or not (an empty list costs a few microseconds).
int harvest_new_rows(ecb_ctx* c, const int32_t* rg, const int32_t* tg, const int32_t* hp, int64_t n, u32 e0,
                     u32 e1) {
  if (e1 <= e0) return ECB_OK;
  // room for one row entry per alignment of this push: the rows of its new ECs come from distinct reads
  const u64 bound = c->arena_used + (u64)n;
  if (bound > 0xFFFFFFFFull) return fail(c, ECB_ERR_LIMIT, "row arena exceeds 2^32 entries");
  CKR(ensure(c, c->arena, (size_t)bound * sizeof(uint2), true));
  HarvestParams H{};
  H.rg = rg; H.tg = tg; H.hp = hp; H.n = (int)n;
  H.n_targets = c->n_targets; H.n_haps = c->n_haps;
  H.ec_rep = (const u32*)c->ec_rep.p;
  H.ec_len = (const u32*)c->ec_len.p;
  H.e0 = e0; H.e1 = e1;
  H.row_len = (u32*)c->row_len.p;
  H.row_off = (u32*)c->row_off.p;
  H.long_list = (u32*)c->long_list.p;
  H.mid_list = (u32*)c->mid_list.p;
  H.big_list = (u32*)c->big_list.p;
  H.ctr = c->d_ctr;
  H.arena = (uint2*)c->arena.p;
  const u32 n_new = e1 - e0;
  ecb_harvest_short_kernel<<<grid_for(n_new, 256, c->sm_count * 16), 256, 0, c->stream>>>(H);
  LAUNCH_CHECK("harvest_short");
  ecb_harvest_warp_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(H);
  LAUNCH_CHECK("harvest_warp");
  ecb_harvest_wsort_kernel<<<c->sm_count * 6, 256, (size_t)8 * HARVEST_WSORT_MAX * 4, c->stream>>>(H);
  LAUNCH_CHECK("harvest_wsort");
  if (!c->long_attr_set) {
    CK(cudaFuncSetAttribute(ecb_harvest_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HARVEST_LONG_MAX * 4));
    c->long_attr_set = true;
  }
  ecb_harvest_long_kernel<<<c->sm_count * 3, 256, HARVEST_LONG_MAX * 4, c->stream>>>(H);
  LAUNCH_CHECK("harvest_long");
  // the list counters go back to zero for the next push (the kernels above have read them by then)
  CK(cudaMemsetAsync(&c->d_ctr->scratch[1], 0, sizeof(u32), c->stream));
  CK(cudaMemsetAsync(&c->d_ctr->scratch[3], 0, sizeof(u32), c->stream));
  CK(cudaMemsetAsync(&c->d_ctr->n_long, 0, sizeof(u32), c->stream));
  return ECB_OK;
}

// Stable LSD radix sort of n (key, value) pairs on the low `bits` key bits.  Input in sort_k[0]/
// sort_v[0]; returns which ping-pong buffer holds the result.
int radix_sort(ecb_ctx* c, u32 n, int bits, int* result_buf) {
  CellScratch& S = c->cells;
  *result_buf = 0;
  if (n == 0) return ECB_OK;
  const u32 n_tiles = (n + SORT_TILE - 1) / SORT_TILE;
  CKR(ensure(c, S.sort_k[1], (size_t)n * 8));
  CKR(ensure(c, S.sort_v[1], (size_t)n * 4));
  CKR(ensure(c, S.hist, (size_t)n_tiles * 256 * 4));
  int cur = 0;
  for (int shift = 0; shift < bits; shift += 8) {
    sort_hist_kernel<<<n_tiles, SORT_THREADS, 0, c->stream>>>((const u64*)S.sort_k[cur].p, n, shift, (u32*)S.hist.p, n_tiles);
    LAUNCH_CHECK("sort_hist");
    CKR(device_scan<false>(c, (const u32*)S.hist.p, (u32*)S.hist.p, (u64)n_tiles * 256, 0, nullptr));
    sort_scatter_kernel<<<n_tiles, SORT_THREADS, 0, c->stream>>>((const u64*)S.sort_k[cur].p, (const u32*)S.sort_v[cur].p,
                                                                   n, shift, (const u32*)S.hist.p, n_tiles,
                                                                   (u64*)S.sort_k[cur ^ 1].p, (u32*)S.sort_v[cur ^ 1].p);
    LAUNCH_CHECK("sort_scatter");
    cur ^= 1;
  }
  *result_buf = cur;
  return ECB_OK;
}

int bits_for(u64 v) {
  int b = 1;
  while (b < 64 && (v >> b)) ++b;
  return b;
}

CellParams make_cell_params(ecb_ctx* c, const CellResult* cr) {
  CellScratch& S = c->cells;
  CellParams P{};
  P.ttable = (const EcbEntry*)c->ttable.p;
  P.t_slots = c->ttable_slots;
  P.ec_table = (const EcbEntry*)c->table.p;
  P.fe_table = (EcbEntry*)S.fe_table.p;
  P.fe_mask = S.fe_slots - 1;
  P.pair_table = (EcbEntry*)S.pair_table.p;
  P.pair_mask = S.pair_slots - 1;
  P.cell_key = (u64*)S.cell_key.p;
  P.cell_total = (u64*)S.cell_total.p;
  P.cell_new = (int32_t*)S.cell_new.p;
  P.ec_keep = (u32*)c->ec_keep.p;
  P.ecid_of = (const u32*)c->ecid_of.p;
  P.min_base = c->min_base;
  P.n_cells = (u32)cr->n_cells;
  P.ctr = c->d_ctr;
  return P;
}

// Cell order, minimum-count filter and the ec_keep flags (bam_utils_multisample.py:503-636).
int cells_prepare(ecb_ctx* c, int64_t min_cell_count, CellResult* cr) {
  CellScratch& S = c->cells;
  if (c->max_end - c->min_base > 0xFFFFFFFFull)
    return fail(c, ECB_ERR_LIMIT, "per-cell mode needs the pushed order range to span less than 2^32 positions");
  if (c->n_triples == 0) return fail(c, ECB_ERR_EMPTY, "no (EC, cell) counts were recorded");
  const u64 min_count = min_cell_count <= 0 ? 1ull : (u64)min_cell_count;  // :596-597
  S.fe_slots = std::max<u32>(1024u, pow2_ceil((u64)c->n_triples * 2));
  S.pair_slots = S.fe_slots;
  CKR(ensure(c, S.fe_table, (size_t)S.fe_slots * sizeof(EcbEntry)));
  CKR(ensure(c, S.pair_table, (size_t)S.pair_slots * sizeof(EcbEntry)));
  CK(cudaMemsetAsync(S.fe_table.p, 0xFF, (size_t)S.fe_slots * sizeof(EcbEntry), c->stream));
  CK(cudaMemsetAsync(S.pair_table.p, 0xFF, (size_t)S.pair_slots * sizeof(EcbEntry), c->stream));
  CK(cudaMemsetAsync(&c->d_ctr->scratch[0], 0, sizeof(u32), c->stream));
  CKR(ensure(c, c->ec_keep, (size_t)c->n_ec * 4));
  CK(cudaMemsetAsync(c->ec_keep.p, 0, (size_t)c->n_ec * 4, c->stream));

  CellParams P = make_cell_params(c, cr);
  const int g_t = grid_for(c->ttable_slots, 256, c->sm_count * 8);
  cells_pass1_kernel<<<g_t, 256, 0, c->stream>>>(P);
  LAUNCH_CHECK("cells_pass1");
  CKR(sync_counters(c));
  CKR(check_device_error(c));
  const u32 n_cells = c->h_ctr->scratch[0];
  if (n_cells == 0 || n_cells > 0x7FFFFFFFu) return fail(c, ECB_ERR_INVALID, "invalid cell ids (negative?)");
  cr->n_cells = n_cells;
  CKR(ensure(c, S.cell_key, (size_t)n_cells * 8));
  CKR(ensure(c, S.cell_total, (size_t)n_cells * 8 * CELLS_TOTAL_REPLICAS));
  CKR(ensure(c, S.cell_new, (size_t)n_cells * 4));
  CKR(ensure(c, c->r_cell_order, (size_t)n_cells * 4));
  CK(cudaMemsetAsync(S.cell_key.p, 0xFF, (size_t)n_cells * 8, c->stream));
  CK(cudaMemsetAsync(S.cell_total.p, 0, (size_t)n_cells * 8 * CELLS_TOTAL_REPLICAS, c->stream));
  P = make_cell_params(c, cr);
  cells_pass2_kernel<<<g_t, 256, 0, c->stream>>>(P);
  LAUNCH_CHECK("cells_pass2");
  cells_total_reduce_kernel<<<grid_for(n_cells, 256, c->sm_count * 8), 256, 0, c->stream>>>((u64*)S.cell_total.p, n_cells);
  LAUNCH_CHECK("cells_total_reduce");

  // order the cells by their nested first-occurrence key
  CKR(ensure(c, S.sort_k[0], (size_t)n_cells * 8));
  CKR(ensure(c, S.sort_v[0], (size_t)n_cells * 4));
  const int g_c = grid_for(n_cells, 256, c->sm_count * 8);
  cells_sort_input_kernel<<<g_c, 256, 0, c->stream>>>((const u64*)S.cell_key.p, n_cells, (u64*)S.sort_k[0].p, (u32*)S.sort_v[0].p);
  LAUNCH_CHECK("cells_sort_input");
  int rb = 0;
  CKR(radix_sort(c, n_cells, 64, &rb));
  CKR(ensure(c, S.flags, (size_t)n_cells * 4));
  CKR(ensure(c, S.offsets, (size_t)n_cells * 4));
  cells_keep_flags_kernel<<<g_c, 256, 0, c->stream>>>((const u64*)S.sort_k[rb].p, (const u32*)S.sort_v[rb].p, n_cells,
                                                      (const u64*)S.cell_total.p, min_count, (u32*)S.flags.p);
  LAUNCH_CHECK("cells_keep_flags");
  u64 n_kept = 0;
  CKR(device_scan<false>(c, (const u32*)S.flags.p, (u32*)S.offsets.p, n_cells, 0, &n_kept));
  CKR(check_device_error(c));
  if (n_kept == 0) return fail(c, ECB_ERR_EMPTY, "no cell reaches the minimum count %llu", (unsigned long long)min_count);
  cr->n_kept_cells = (int64_t)n_kept;
  cells_assign_kernel<<<g_c, 256, 0, c->stream>>>((const u64*)S.sort_k[rb].p, (const u32*)S.sort_v[rb].p, n_cells,
                                                  (const u64*)S.cell_total.p, min_count, (const u32*)S.offsets.p,
                                                  (int32_t*)S.cell_new.p, (int32_t*)c->r_cell_order.p);
  LAUNCH_CHECK("cells_assign");
  cells_ec_keep_kernel<<<grid_for(S.pair_slots, 256, c->sm_count * 8), 256, 0, c->stream>>>(P);
  LAUNCH_CHECK("cells_ec_keep");
  return ECB_OK;
}

// N matrix (CSC) from the kept (EC, cell) pairs (bam_utils_multisample.py:737-747,783-791).
int cells_emit(ecb_ctx* c, CellResult* cr, const u32* ecid_of, u64 n_ec_final) {
  CellScratch& S = c->cells;
  CellParams P = make_cell_params(c, cr);
  P.ecid_of = ecid_of;
  const int g_p = grid_for(S.pair_slots, 256, c->sm_count * 8);
  CKR(ensure(c, S.flags, (size_t)S.pair_slots * 4));
  CKR(ensure(c, S.offsets, (size_t)S.pair_slots * 4));
  cells_pair_flags_kernel<<<g_p, 256, 0, c->stream>>>(P, (u32*)S.flags.p);
  LAUNCH_CHECK("cells_pair_flags");
  u64 nnz = 0;
  CKR(device_scan<false>(c, (const u32*)S.flags.p, (u32*)S.offsets.p, S.pair_slots, 0, &nnz));
  if (nnz == 0 || nnz > 0x7FFFFFFFull) return fail(c, ECB_ERR_LIMIT, "N matrix has %llu non-zeros", (unsigned long long)nnz);
  cr->nnz_n = (int64_t)nnz;
  CKR(ensure(c, S.sort_k[0], (size_t)nnz * 8));
  CKR(ensure(c, S.sort_v[0], (size_t)nnz * 4));
  cells_pair_emit_kernel<<<g_p, 256, 0, c->stream>>>(P, (const u32*)S.flags.p, (const u32*)S.offsets.p,
                                                     (u64*)S.sort_k[0].p, (u32*)S.sort_v[0].p);
  LAUNCH_CHECK("cells_pair_emit");
  int rb = 0;
  CKR(radix_sort(c, (u32)nnz, 32 + bits_for((u64)cr->n_kept_cells), &rb));
  (void)n_ec_final;
  CKR(ensure(c, c->r_n_indptr, (size_t)(cr->n_kept_cells + 1) * 4));
  CKR(ensure(c, c->r_n_indices, (size_t)nnz * 4));
  CKR(ensure(c, c->r_n_data, (size_t)nnz * 4));
  cells_csc_kernel<<<grid_for(nnz, 256, c->sm_count * 8), 256, 0, c->stream>>>(
      (const u64*)S.sort_k[rb].p, (const u32*)S.sort_v[rb].p, (u32)nnz, (u32)cr->n_kept_cells,
      (int32_t*)c->r_n_indptr.p, (int32_t*)c->r_n_indices.p, (int32_t*)c->r_n_data.p);
  LAUNCH_CHECK("cells_csc");
  return ECB_OK;
}

void cells_release(ecb_ctx* c) {
  CellScratch& S = c->cells;
  DevBuf* bufs[] = {&S.fe_table, &S.pair_table, &S.cell_key, &S.cell_total, &S.cell_new, &S.sort_k[0], &S.sort_k[1],
                    &S.sort_v[0], &S.sort_v[1], &S.hist, &S.flags, &S.offsets};
  for (DevBuf* b : bufs) release(c, *b);
}

}  // namespace

extern "C" {

int ecb_version(void) { return 1000; }

const char* ecb_last_error(const ecb_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int ecb_create(ecb_ctx** out, int device, int n_targets, int n_haps, int with_cells, int64_t alignments_hint) {
  ecb_ctx* c = nullptr;
  if (!out) return fail(c, ECB_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (n_targets < 1 || n_targets > ECB_MAX_TARGETS) return fail(c, ECB_ERR_INVALID, "n_targets %d outside [1, 2^26]", n_targets);
  if (n_haps < 1 || n_haps > ECB_MAX_HAPS) return fail(c, ECB_ERR_INVALID, "n_haps %d outside [1, 31]", n_haps);
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    return fail(c, ECB_ERR_NO_DEVICE, "no CUDA device is visible; libecb200 has no CPU path");
  if (device < 0 || device >= count) return fail(c, ECB_ERR_INVALID, "device %d out of range (%d visible)", device, count);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(c, ECB_ERR_CUDA, "cudaGetDeviceProperties failed");
  if (prop.major != 10)
    return fail(c, ECB_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
  ecb_ctx* ctx = new ecb_ctx();
  c = ctx;
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->n_targets = n_targets;
  c->n_haps = n_haps;
  c->with_cells = with_cells ? 1 : 0;
  c->hint = std::max<int64_t>(alignments_hint, 0);
  auto bail = [&](int code) {
    g_create_error = c->err;
    ecb_destroy(c);
    return code;
  };
  if (cudaSetDevice(device) != cudaSuccess) { fail(c, ECB_ERR_CUDA, "cudaSetDevice failed"); return bail(ECB_ERR_CUDA); }
  {
    // a pool of this context's own, with released memory kept (the default pool would hand it back to the
    // driver at every synchronisation); the device's default pool - which torch and NCCL may share - is
    // left as it is
    cudaMemPoolProps props{};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    if (cudaMemPoolCreate(&c->pool, &props) != cudaSuccess) { fail(c, ECB_ERR_CUDA, "cudaMemPoolCreate failed"); return bail(ECB_ERR_CUDA); }
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) { fail(c, ECB_ERR_CUDA, "cudaStreamCreate failed"); return bail(ECB_ERR_CUDA); }
  c->stream = c->own_stream;
  for (auto& e : c->ev)
    if (cudaEventCreate(&e) != cudaSuccess) { fail(c, ECB_ERR_CUDA, "cudaEventCreate failed"); return bail(ECB_ERR_CUDA); }
  if (cudaMalloc(&c->d_ctr, sizeof(EcbCounters)) != cudaSuccess || cudaMallocHost(&c->h_ctr, sizeof(EcbCounters)) != cudaSuccess ||
      cudaMallocHost(&c->h_total, 64) != cudaSuccess) {
    fail(c, ECB_ERR_CUDA, "counter allocation failed");
    return bail(ECB_ERR_CUDA);
  }
  *out = c;
  return ECB_OK;
}

int ecb_set_option(ecb_ctx* c, int option, int64_t value) {
  if (!c) return ECB_ERR_INVALID;
  switch (option) {
    case ECB_OPT_RESULT_ON_DEVICE: c->result_on_device = value ? 1 : 0; break;
    case ECB_OPT_TABLE_SLOTS:
      if (c->table_slots) return fail(c, ECB_ERR_STATE, "table already allocated");
      c->opt_table_slots = value; break;
    case ECB_OPT_PAIR_SLOTS:
      if (c->ttable_slots) return fail(c, ECB_ERR_STATE, "table already allocated");
      c->opt_pair_slots = value; break;
    case ECB_OPT_GRID_CTAS: c->opt_grid = value; break;
    case ECB_OPT_HOT_CACHE: c->use_cache = value ? 1 : 0; break;
    case ECB_OPT_VERIFY_KEYS: c->verify_keys = value ? 1 : 0; break;
    case ECB_OPT_CHUNK_LEN: c->opt_chunk_len = value; break;
    case ECB_OPT_PAGEABLE_RESULTS:
      if (c->h_res) return fail(c, ECB_ERR_STATE, "result buffers already allocated");
      c->pageable_results = value ? 1 : 0; break;
    default: return fail(c, ECB_ERR_INVALID, "unknown option %d", option);
  }
  return ECB_OK;
}

int ecb_set_stream(ecb_ctx* c, void* cuda_stream) {
  if (!c) return ECB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
  return ECB_OK;
}

// One launch sequence over one contiguous piece of columns (the whole push, or one piece of a
// pipelined host push).
static int push_one(ecb_ctx* c, const int32_t* read_group, const int32_t* target_idx, const int32_t* hap_idx,
                    const int32_t* cell_idx, int64_t n, int64_t order_base, int drop_last_group, int on_device,
                    u32 push_id) {
  CK(cudaEventRecord(c->ev[0], c->stream));
  const int32_t *rg = read_group, *tg = target_idx, *hp = hap_idx, *cell = cell_idx;
  const size_t col_bytes = (size_t)n * 4;
  auto misaligned = [](const void* p) { return ((uintptr_t)p & 15) != 0; };
  const bool stage = !on_device || misaligned(rg) || misaligned(tg) || misaligned(hp);
  if (stage) {
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    CKR(ensure(c, c->st_rg, col_bytes));
    CKR(ensure(c, c->st_tg, col_bytes));
    CKR(ensure(c, c->st_hp, col_bytes));
    CK(cudaMemcpyAsync(c->st_rg.p, rg, col_bytes, kind, c->stream));
    CK(cudaMemcpyAsync(c->st_tg.p, tg, col_bytes, kind, c->stream));
    CK(cudaMemcpyAsync(c->st_hp.p, hp, col_bytes, kind, c->stream));
    rg = (const int32_t*)c->st_rg.p; tg = (const int32_t*)c->st_tg.p; hp = (const int32_t*)c->st_hp.p;
    if (cell) {
      CKR(ensure(c, c->st_cell, col_bytes));
      CK(cudaMemcpyAsync(c->st_cell.p, cell, col_bytes, kind, c->stream));
      cell = (const int32_t*)c->st_cell.p;
    }
    if (!on_device) c->stats.h2d_bytes += (int64_t)col_bytes * (cell ? 4 : 3);
  }

  // capacity: at most half full before the push starts, and - going by the ECs per alignment seen so
  // far - at most about 60 % full after it (a growth in the middle of a push costs a replay)
  while ((u64)c->n_ec * 2 > c->table_slots) CKR(grow_table(c, c->table_slots * 2));
  if (c->n_alignments > 0 && !c->opt_table_slots) {
    const double rate = (double)c->n_ec / (double)c->n_alignments;
    const u64 projected = (u64)c->n_ec + (u64)(rate * (double)n);
    if (projected * 5 > (u64)c->table_slots * 3 && c->table_slots < (1u << 31)) CKR(grow_table(c, pow2_ceil(projected * 2)));
  }
  if (c->with_cells) {
    // the triple table must be able to absorb one new entry per READ of this push without filling up
    // (it has no replay path): count the reads first - one pass over read_group, 4 bytes per alignment
    CK(cudaMemsetAsync(&c->d_ctr->scratch[4], 0, sizeof(u32), c->stream));
    ecb_count_reads_kernel<<<grid_for((u64)n, 1024, c->sm_count * 4), 256, 0, c->stream>>>(rg, (int)n, &c->d_ctr->scratch[4]);
    LAUNCH_CHECK("count_reads");
    CKR(sync_counters(c));
    const u64 reads = c->h_ctr->scratch[4];
    u64 need = ((u64)c->n_triples + reads) * 2;
    if (need > c->ttable_slots) CKR(rebuild_triple_table(c, pow2_ceil(need), nullptr));
  }
  const size_t ov_words = ((size_t)n + 31) / 32;
  CKR(ensure(c, c->overflow_bits, ov_words * 4));
  CK(cudaMemsetAsync(c->overflow_bits.p, 0, ov_words * 4, c->stream));

  const u32 e_before = c->n_ec;
  CKR(group_prepare_launch(c));
  int grid = 1, chunk_len = 32;
  group_geometry(c, n, &grid, &chunk_len);
  CKR(ensure(c, c->spill, (size_t)grid * ECB_CACHE * sizeof(EcbSpill)));
  GroupParams P = make_group_params(c, rg, tg, hp, cell, n, order_base, drop_last_group, push_id);
  P.chunk_len = chunk_len;
  CK(cudaMemsetAsync(&c->d_ctr->chunk_next, 0, sizeof(u32), c->stream));
  CK(cudaEventRecord(c->ev[1], c->stream));
  if (c->with_cells) ecb_group_insert_kernel<true><<<grid, ECB_GTHREADS, sizeof(GroupSmem), c->stream>>>(P);
  else ecb_group_insert_kernel<false><<<grid, ECB_GTHREADS, sizeof(GroupSmem), c->stream>>>(P);
  LAUNCH_CHECK("group_insert");
  CK(cudaEventRecord(c->ev[2], c->stream));
  CKR(sync_counters(c));
  CKR(check_device_error(c));
  DevBuf spill_in;
  while (c->h_ctr->n_overflow || c->h_ctr->n_spill) {  // table too full: grow, then replay just what did not fit
    c->stats.overflow_reads += c->h_ctr->n_overflow;
    const u32 n_spill = c->h_ctr->n_spill;
    if (c->table_slots >= (1u << 31)) return fail(c, ECB_ERR_LIMIT, "EC table cannot grow beyond 2^31 slots");
    if (n_spill) {
      CKR(ensure(c, spill_in, (size_t)n_spill * sizeof(EcbSpill)));
      CK(cudaMemcpyAsync(spill_in.p, c->spill.p, (size_t)n_spill * sizeof(EcbSpill), cudaMemcpyDeviceToDevice, c->stream));
    }
    CKR(grow_table(c, c->table_slots * 4u > c->table_slots ? c->table_slots * 4u : (1u << 31)));
    CK(cudaMemsetAsync(&c->d_ctr->n_overflow, 0, sizeof(u32), c->stream));
    CK(cudaMemsetAsync(&c->d_ctr->n_spill, 0, sizeof(u32), c->stream));
    P = make_group_params(c, rg, tg, hp, cell, n, order_base, drop_last_group, push_id);
    const int rg_grid = grid_for(ov_words, 256, c->sm_count * 8);
    if (c->with_cells) ecb_replay_kernel<true><<<rg_grid, 256, 0, c->stream>>>(P);
    else ecb_replay_kernel<false><<<rg_grid, 256, 0, c->stream>>>(P);
    LAUNCH_CHECK("replay");
    if (n_spill) {
      ecb_spill_replay_kernel<<<grid_for(n_spill, 256, c->sm_count * 8), 256, 0, c->stream>>>(
          P, (const EcbSpill*)spill_in.p, n_spill);
      LAUNCH_CHECK("spill_replay");
    }
    CKR(sync_counters(c));
    CKR(check_device_error(c));
  }
  release(c, spill_in);
  if (c->h_ctr->n_triple_overflow) return fail(c, ECB_ERR_LIMIT, "(file, EC, cell) table ran out of probes");
  c->n_ec = c->h_ctr->n_ec;
  c->n_triples = c->h_ctr->n_triples;
  CK(cudaEventRecord(c->ev[3], c->stream));
  CKR(harvest_new_rows(c, rg, tg, hp, n, e_before, c->n_ec));
  CK(cudaEventRecord(c->ev[4], c->stream));
  if (c->verify_keys) {
    VerifyParams V{};
    V.rg = rg; V.tg = tg; V.hp = hp; V.n = (int)n; V.drop_last = drop_last_group;
    V.table = (const EcbEntry*)c->table.p;
    V.mask = c->table_slots - 1;
    V.row_len = (const u32*)c->row_len.p;
    V.row_off = (const u32*)c->row_off.p;
    V.arena = (const uint2*)c->arena.p;
    V.ctr = c->d_ctr;
    ecb_verify_kernel<<<grid_for((u64)n, 256, c->sm_count * 16), 256, 0, c->stream>>>(V);
    LAUNCH_CHECK("verify");
  }
  // the push ends with its second (and last) host round trip: the caller's buffers are free again, the
  // counters (arena cursor, device-side errors of the harvest / verification) are current on the host
  CKR(sync_counters(c));
  CKR(check_device_error(c));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, c->ev[1], c->ev[2])); c->stats.group_ms = ms;
  CK(cudaEventElapsedTime(&ms, c->ev[3], c->ev[4])); c->stats.harvest_ms = ms;
  CK(cudaEventElapsedTime(&ms, c->ev[0], c->ev[4])); c->stats.push_ms = ms;

  c->min_base = std::min<u64>(c->min_base, (u64)order_base);
  c->max_end = std::max<u64>(c->max_end, (u64)order_base + (u64)n);
  c->n_alignments += n;
  c->stats.table_slots = c->table_slots;
  c->stats.table_used = c->n_ec;
  return ECB_OK;
}

// Host columns of a large push are sent in read-aligned pieces on a second stream, two staging sets
// deep, so that the copy of piece i+1 runs while piece i is grouped and harvested.  The pieces share
// the push id (the "file" of the per-cell path) and only the last one drops the last read.
#define ECB_PIPE_PIECE (8ll << 20)   // alignments per piece (96 MB of columns)

static int push_pipelined(ecb_ctx* c, const int32_t* rg, const int32_t* tg, const int32_t* hp, const int32_t* cell,
                          int64_t n, int64_t order_base, int drop_last_group, u32 push_id) {
  std::vector<int64_t> cut{0};
  while (cut.back() < n) {
    int64_t b = std::min<int64_t>(n, cut.back() + ECB_PIPE_PIECE);
    while (b < n && rg[b] == rg[b - 1]) ++b;   // never split a read
    cut.push_back(b);
  }
  const int pieces = (int)cut.size() - 1;
  int64_t longest = 0;
  for (int i = 0; i < pieces; ++i) longest = std::max(longest, cut[i + 1] - cut[i]);
  if (!c->copy_stream) {
    CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->copy_done[0], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->copy_done[1], cudaEventDisableTiming));
  }
  DevBuf* set[2][4] = {{&c->st_rg, &c->st_tg, &c->st_hp, &c->st_cell}, {&c->st2_rg, &c->st2_tg, &c->st2_hp, &c->st2_cell}};
  const int ncols = cell ? 4 : 3;
  for (int k = 0; k < 2; ++k)
    for (int j = 0; j < ncols; ++j) CKR(ensure(c, *set[k][j], (size_t)longest * 4));
  CK(cudaStreamSynchronize(c->stream));   // the staging sets exist before the copy stream touches them
  const int32_t* src[4] = {rg, tg, hp, cell};
  auto send = [&](int i) -> int {
    const size_t bytes = (size_t)(cut[i + 1] - cut[i]) * 4;
    for (int j = 0; j < ncols; ++j)
      CK(cudaMemcpyAsync(set[i & 1][j]->p, src[j] + cut[i], bytes, cudaMemcpyHostToDevice, c->copy_stream));
    CK(cudaEventRecord(c->copy_done[i & 1], c->copy_stream));
    c->stats.h2d_bytes += (int64_t)bytes * ncols;
    return ECB_OK;
  };
  CKR(send(0));
  for (int i = 0; i < pieces; ++i) {
    if (i + 1 < pieces) CKR(send(i + 1));   // its staging set was last read by piece i-1, which has completed
    CK(cudaStreamWaitEvent(c->stream, c->copy_done[i & 1], 0));
    const int rc = push_one(c, (const int32_t*)set[i & 1][0]->p, (const int32_t*)set[i & 1][1]->p,
                            (const int32_t*)set[i & 1][2]->p, cell ? (const int32_t*)set[i & 1][3]->p : nullptr,
                            cut[i + 1] - cut[i], order_base + cut[i], drop_last_group && i + 1 == pieces, 1, push_id);
    if (rc != ECB_OK) {
      cudaStreamSynchronize(c->copy_stream);
      return rc;
    }
  }
  return ECB_OK;
}

int ecb_push(ecb_ctx* c, const int32_t* read_group, const int32_t* target_idx, const int32_t* hap_idx,
             const int32_t* cell_idx, int64_t n, int64_t order_base, int drop_last_group, int on_device) {
  if (!c) return ECB_ERR_INVALID;
  if (n < 0 || order_base < 0) return fail(c, ECB_ERR_INVALID, "negative n or order_base");
  if (n > 0x7FFFFFFFll - 8192) return fail(c, ECB_ERR_LIMIT, "a push is limited to 2^31-8192 alignments; split it");
  if (n > 0 && (!read_group || !target_idx || !hap_idx)) return fail(c, ECB_ERR_INVALID, "NULL column");
  if (c->with_cells && n > 0 && !cell_idx) return fail(c, ECB_ERR_INVALID, "context has cells but cell_idx is NULL");
  if (!c->with_cells && cell_idx) return fail(c, ECB_ERR_INVALID, "cell_idx given but context was created without cells");
  CK(cudaSetDevice(c->device));
  if (!c->table_slots) CKR(init_table(c, n));
  const u32 push_id = c->push_count++;
  if (n == 0) return ECB_OK;
  if (!on_device && n >= 2 * ECB_PIPE_PIECE)
    return push_pipelined(c, read_group, target_idx, hap_idx, cell_idx, n, order_base, drop_last_group, push_id);
  return push_one(c, read_group, target_idx, hap_idx, cell_idx, n, order_base, drop_last_group, on_device, push_id);
}

int ecb_finalize(ecb_ctx* c, int64_t min_cell_count, ecb_result* out) {
  if (!c || !out) return ECB_ERR_INVALID;
  memset(out, 0, sizeof *out);
  CK(cudaSetDevice(c->device));
  if (!c->table_slots || c->n_ec == 0) return fail(c, ECB_ERR_EMPTY, "no equivalence classes (nothing pushed)");
  CK(cudaEventRecord(c->ev[0], c->stream));
  const u32 n_prov = c->n_ec;   // the host mirror of the counters is current: every push ends with a sync
  const u64 span = c->max_end - c->min_base;
  if (span > (1ull << 34)) return fail(c, ECB_ERR_LIMIT, "order_base range spans more than 2^34 positions");
  const size_t words = (size_t)((span + 31) / 32) + 1;

  CKR(ensure(c, c->bitmap, words * 4));
  CKR(ensure(c, c->word_rank, words * 4));
  CKR(ensure(c, c->first_rel, (size_t)n_prov * 8));
  CKR(ensure(c, c->ecid_of, (size_t)n_prov * 4));
  CK(cudaMemsetAsync(c->bitmap.p, 0, words * 4, c->stream));

  FinalizeParams F{};
  F.table = (const EcbEntry*)c->table.p;
  F.ec_slot = (const u32*)c->ec_slot.p;
  F.row_len = (const u32*)c->row_len.p;
  F.row_off = (const u32*)c->row_off.p;
  F.arena = (const uint2*)c->arena.p;
  F.n_ec = n_prov;
  F.min_base = c->min_base;
  F.bitmap = (u32*)c->bitmap.p;
  F.word_rank = (const u32*)c->word_rank.p;
  F.first_rel = (u64*)c->first_rel.p;
  F.ecid_of = (u32*)c->ecid_of.p;
  F.wide_count = &c->d_ctr->scratch[2];
  F.wide_list = (u32*)c->long_list.p;
  CKR(ensure(c, c->count_of, (size_t)n_prov * 4));
  F.count_of = (u32*)c->count_of.p;
  CK(cudaMemsetAsync(&c->d_ctr->scratch[2], 0, sizeof(u32), c->stream));

  CellResult cr{};
  if (c->with_cells) {
    CKR(cells_prepare(c, min_cell_count, &cr));   // cell order, kept cells, ec_keep flags
    F.ec_keep = (const u32*)c->ec_keep.p;
  }

  const int g_ec = grid_for(n_prov, 256, c->sm_count * 8);
  ecb_fin_mark_kernel<<<g_ec, 256, 0, c->stream>>>(F);
  LAUNCH_CHECK("fin_mark");
  u64 n_kept = n_prov;
  if (c->with_cells) {
    CKR(device_scan<true>(c, F.bitmap, (u32*)c->word_rank.p, words, 0, &n_kept));   // the cell filter decides: host round trip
    if (n_kept == 0) return fail(c, ECB_ERR_EMPTY, "no equivalence class survives the cell filter");
  } else {
    // single sample: every EC is kept, so E is known; the count of first-occurrence bits is checked against it
    // when the one round trip of this call brings it back (it differs when order_base ranges overlapped)
    CKR(device_scan<true>(c, F.bitmap, (u32*)c->word_rank.p, words, 0, nullptr));
    const u32 n_blocks = (u32)((words + SCAN_BLOCK - 1) / SCAN_BLOCK);
    CK(cudaMemcpyAsync(c->h_total + 1, (const u64*)c->scan_partials.p + n_blocks, sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
  }
  const u64 E = n_kept;
  if (E + 1 > 0x7FFFFFFFull) return fail(c, ECB_ERR_LIMIT, "more than 2^31 equivalence classes");

  CKR(ensure(c, c->r_a_indptr, (E + 1) * 4));
  F.a_indptr = (int32_t*)c->r_a_indptr.p;
  if (!c->with_cells) {
    CKR(ensure(c, c->r_n_indptr, 2 * 4));
    CKR(ensure(c, c->r_n_indices, E * 4));
    CKR(ensure(c, c->r_n_data, E * 4));
    F.n_indices = (int32_t*)c->r_n_indices.p;
    F.n_data = (int32_t*)c->r_n_data.p;
  }
  CK(cudaMemsetAsync((int32_t*)c->r_a_indptr.p + E, 0, 4, c->stream));
  if (c->with_cells) ecb_fin_rank_kernel<false><<<g_ec, 256, 0, c->stream>>>(F);
  else ecb_fin_rank_kernel<true><<<g_ec, 256, 0, c->stream>>>(F);
  LAUNCH_CHECK("fin_rank");
  // row lengths -> indptr; the rows follow at once: the arena's fill (known on the host since the last push)
  // bounds the number of non-zeros, so the result arrays need not wait for the exact figure
  CKR(device_scan<false>(c, (const u32*)c->r_a_indptr.p, (u32*)c->r_a_indptr.p, E + 1, 0, nullptr));
  {
    const u32 n_blocks = (u32)((E + 1 + SCAN_BLOCK - 1) / SCAN_BLOCK);
    CK(cudaMemcpyAsync(c->h_total, (const u64*)c->scan_partials.p + n_blocks, sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    c->stats.d2h_bytes += 2 * sizeof(u64);
  }
  const u64 z_bound = std::max<u64>(c->arena_used, 1);
  if (z_bound > 0x7FFFFFFFull) return fail(c, ECB_ERR_LIMIT, "A matrix may have more than 2^31-1 non-zeros");
  CKR(ensure(c, c->r_a_indices, z_bound * 4));
  CKR(ensure(c, c->r_a_data, z_bound * 4));
  F.a_indices = (int32_t*)c->r_a_indices.p;
  F.a_data = (int32_t*)c->r_a_data.p;
  ecb_fin_rows_kernel<<<grid_for(n_prov, 256, c->sm_count * 16), 256, 0, c->stream>>>(F);
  LAUNCH_CHECK("fin_rows");
  ecb_fin_rows_long_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(F, (const u32*)c->long_list.p, 0xFFFFFFFFu);   // (count on the device)
  LAUNCH_CHECK("fin_rows_long");
  CellResult cr2 = cr;
  if (c->with_cells) CKR(cells_emit(c, &cr2, (const u32*)c->ecid_of.p, E));   // N matrix as CSC
  else {
    ecb_set_pair_kernel<<<1, 1, 0, c->stream>>>((int32_t*)c->r_n_indptr.p, 0, (int32_t)E);
    LAUNCH_CHECK("set_pair");
  }
  CKR(sync_counters(c));   // the round trip of this call: non-zeros, first-occurrence bits, device-side errors
  CKR(check_device_error(c));
  const u64 Z = *c->h_total;
  if (!c->with_cells && c->h_total[1] != n_prov)
    return fail(c, ECB_ERR_INVALID, "order_base ranges of different pushes overlap (%llu first positions for %u ECs)",
                (unsigned long long)c->h_total[1], n_prov);
  if (Z > 0x7FFFFFFFull) return fail(c, ECB_ERR_LIMIT, "A matrix has more than 2^31-1 non-zeros");

  int64_t n_samples = 1, nnz_n = (int64_t)E;
  if (c->with_cells) {
    n_samples = cr2.n_kept_cells;
    nnz_n = cr2.nnz_n;
  }

  out->n_ec = (int64_t)E;
  out->nnz_a = (int64_t)Z;
  out->n_samples = n_samples;
  out->nnz_n = nnz_n;
  out->n_reads = (int64_t)c->h_ctr->n_reads;
  out->n_alignments = c->n_alignments;
  if (c->result_on_device) {
    out->a_indptr = (const int32_t*)c->r_a_indptr.p;
    out->a_indices = (const int32_t*)c->r_a_indices.p;
    out->a_data = (const int32_t*)c->r_a_data.p;
    out->n_indptr = (const int32_t*)c->r_n_indptr.p;
    out->n_indices = (const int32_t*)c->r_n_indices.p;
    out->n_data = (const int32_t*)c->r_n_data.p;
    out->cell_order = c->with_cells ? (const int32_t*)c->r_cell_order.p : nullptr;
  } else {
    const size_t sizes[7] = {(size_t)(E + 1) * 4, (size_t)Z * 4, (size_t)Z * 4, (size_t)(n_samples + 1) * 4,
                             (size_t)nnz_n * 4, (size_t)nnz_n * 4, c->with_cells ? (size_t)n_samples * 4 : 0};
    const void* src[7] = {c->r_a_indptr.p, c->r_a_indices.p, c->r_a_data.p, c->r_n_indptr.p,
                          c->r_n_indices.p, c->r_n_data.p, c->r_cell_order.p};
    size_t total = 0, offs[7];
    for (int i = 0; i < 7; ++i) { offs[i] = total; total += (sizes[i] + 63) & ~(size_t)63; }
    if (total > c->h_res_bytes) {
      if (c->h_res) {
        if (c->pageable_results) free(c->h_res); else CK(cudaFreeHost(c->h_res));
      }
      c->h_res = nullptr;
      c->h_res_bytes = 0;
      if (c->pageable_results) {
        c->h_res = malloc(total + total / 4);
        if (!c->h_res) return fail(c, ECB_ERR_CUDA, "out of host memory for %zu result bytes", total);
      } else {
        CK(cudaMallocHost(&c->h_res, total + total / 4));
      }
      c->h_res_bytes = total + total / 4;
    }
    char* base = (char*)c->h_res;
    for (int i = 0; i < 7; ++i)
      if (sizes[i]) {
        CK(cudaMemcpyAsync(base + offs[i], src[i], sizes[i], cudaMemcpyDeviceToHost, c->stream));
        c->stats.d2h_bytes += (int64_t)sizes[i];
      }
    out->a_indptr = (const int32_t*)(base + offs[0]);
    out->a_indices = (const int32_t*)(base + offs[1]);
    out->a_data = (const int32_t*)(base + offs[2]);
    out->n_indptr = (const int32_t*)(base + offs[3]);
    out->n_indices = (const int32_t*)(base + offs[4]);
    out->n_data = (const int32_t*)(base + offs[5]);
    out->cell_order = c->with_cells ? (const int32_t*)(base + offs[6]) : nullptr;
  }
  CK(cudaEventRecord(c->ev[1], c->stream));
  CK(cudaStreamSynchronize(c->stream));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]));
  c->stats.finalize_ms = ms;
  return ECB_OK;
}

int ecb_reset(ecb_ctx* c) {
  if (!c) return ECB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  if (c->table_slots) {
    CK(cudaMemsetAsync(c->table.p, 0xFF, (size_t)c->table_slots * sizeof(EcbEntry), c->stream));
    if (c->ttable_slots) CK(cudaMemsetAsync(c->ttable.p, 0xFF, (size_t)c->ttable_slots * sizeof(EcbEntry), c->stream));
    CK(cudaMemsetAsync(c->d_ctr, 0, sizeof(EcbCounters), c->stream));   // stream-ordered: no host sync needed
  }
  memset(c->h_ctr, 0, sizeof(EcbCounters));
  c->n_ec = 0;
  c->n_triples = 0;
  c->arena_used = 0;
  c->min_base = ~0ull;
  c->max_end = 0;
  c->n_alignments = 0;
  c->push_count = 0;
  const int64_t slots = c->stats.table_slots;
  c->stats = ecb_stats{};
  c->stats.table_slots = slots;
  return ECB_OK;
}

int ecb_get_stats(const ecb_ctx* c, ecb_stats* out) {
  if (!c || !out) return ECB_ERR_INVALID;
  *out = c->stats;
  out->table_slots = c->table_slots;
  out->table_used = c->n_ec;
  out->row_entries = (int64_t)c->arena_used;
  return ECB_OK;
}

int ecb_rebase(ecb_ctx* c, int64_t delta) {
  if (!c) return ECB_ERR_INVALID;
  if (delta < 0) return fail(c, ECB_ERR_INVALID, "negative delta");
  if (c->with_cells) return fail(c, ECB_ERR_INVALID, "ecb_rebase covers the single-sample path only");
  CK(cudaSetDevice(c->device));
  if (delta == 0 || !c->table_slots || c->n_ec == 0) return ECB_OK;
  ecb_rebase_kernel<<<grid_for(c->n_ec, 256, c->sm_count * 8), 256, 0, c->stream>>>(
      (EcbEntry*)c->table.p, (const u32*)c->ec_slot.p, c->n_ec, (u64)delta);
  LAUNCH_CHECK("rebase");
  c->min_base += (u64)delta;
  c->max_end += (u64)delta;
  return ECB_OK;
}

int ecb_export_partition(ecb_ctx* c, int world, ecb_export* out) {
  if (!c || !out) return ECB_ERR_INVALID;
  if (world < 1 || world > ECB_MAX_WORLD) return fail(c, ECB_ERR_INVALID, "world %d outside [1, %d]", world, ECB_MAX_WORLD);
  if (c->with_cells) return fail(c, ECB_ERR_INVALID, "the multi-GPU exchange covers the single-sample path only");
  memset(out, 0, sizeof *out);
  CK(cudaSetDevice(c->device));
  if (!c->table_slots) CKR(init_table(c));
  const u32 n_ec = c->n_ec;
  CKR(ensure(c, c->x_counts, (size_t)2 * world * 4));
  CKR(ensure(c, c->x_base, (size_t)2 * world * 8));
  CK(cudaMemsetAsync(c->x_counts.p, 0, (size_t)2 * world * 4, c->stream));
  ExportParams P{};
  P.table = (const EcbEntry*)c->table.p;
  P.ec_slot = (const u32*)c->ec_slot.p;
  P.row_len = (const u32*)c->row_len.p;
  P.row_off = (const u32*)c->row_off.p;
  P.arena = (const uint2*)c->arena.p;
  P.n_ec = n_ec;
  P.world = (u32)world;
  P.counts = (u32*)c->x_counts.p;
  P.base = (const u64*)c->x_base.p;
  std::vector<u32> h_counts(2 * world, 0);
  const int g = grid_for(std::max<u32>(n_ec, 1), 256, c->sm_count * 8);
  if (n_ec) {
    ecb_export_count_kernel<<<g, 256, 0, c->stream>>>(P);
    LAUNCH_CHECK("export_count");
    CK(cudaMemcpyAsync(h_counts.data(), c->x_counts.p, (size_t)2 * world * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  std::vector<u64> h_base(2 * world, 0);
  u64 ec_run = 0, row_run = 0;
  for (int w = 0; w < world; ++w) {
    h_base[w] = ec_run;
    h_base[world + w] = row_run;
    c->x_part_ec[w] = h_counts[w];
    c->x_part_rows[w] = h_counts[world + w];
    ec_run += h_counts[w];
    row_run += h_counts[world + w];
  }
  CKR(ensure(c, c->x_meta, std::max<u64>(ec_run, 1) * ECB_META_WORDS * 8));
  CKR(ensure(c, c->x_rows, std::max<u64>(row_run, 1) * 8));
  if (n_ec) {
    CK(cudaMemcpyAsync(c->x_base.p, h_base.data(), (size_t)2 * world * 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(c->x_counts.p, 0, (size_t)2 * world * 4, c->stream));
    P.meta = (long long*)c->x_meta.p;
    P.rows = (int2*)c->x_rows.p;
    ecb_export_fill_kernel<<<g, 256, 0, c->stream>>>(P);
    LAUNCH_CHECK("export_fill");
    CK(cudaStreamSynchronize(c->stream));
  }
  out->n_ec = (int64_t)ec_run;
  out->n_rows = (int64_t)row_run;
  out->meta = (const int64_t*)c->x_meta.p;
  out->rows = (const int32_t*)c->x_rows.p;
  out->part_ec_counts = c->x_part_ec;
  out->part_row_counts = c->x_part_rows;
  out->min_base = n_ec ? (int64_t)c->min_base : 0;
  out->max_end = n_ec ? (int64_t)c->max_end : 0;
  return ECB_OK;
}

int ecb_import_entries(ecb_ctx* c, const int64_t* meta_device, const int32_t* rows_device,
                       const int64_t* part_ec_counts, const int64_t* part_row_counts, int n_parts) {
  if (!c) return ECB_ERR_INVALID;
  if (n_parts < 1 || n_parts > ECB_MAX_WORLD || !part_ec_counts || !part_row_counts)
    return fail(c, ECB_ERR_INVALID, "bad partition description");
  CK(cudaSetDevice(c->device));
  if (!c->table_slots) CKR(init_table(c));
  PartTable parts{};
  parts.n = (u32)n_parts;
  u64 n_rec = 0, n_rows = 0;
  for (int p = 0; p < n_parts; ++p) {
    if (part_ec_counts[p] < 0 || part_row_counts[p] < 0) return fail(c, ECB_ERR_INVALID, "negative partition size");
    parts.row_base[p] = (long long)n_rows;
    n_rec += (u64)part_ec_counts[p];
    n_rows += (u64)part_row_counts[p];
    parts.ec_end[p] = (long long)n_rec;
  }
  if (n_rec == 0) return ECB_OK;
  if (!meta_device || (n_rows && !rows_device)) return fail(c, ECB_ERR_INVALID, "NULL exchange buffer");
  if (n_rec > 0x7FFFFFFFull) return fail(c, ECB_ERR_LIMIT, "too many records in one import");
  // the owner table must be able to take every incoming record as a new EC without filling up
  const u64 need = ((u64)c->n_ec + n_rec) * 2;
  if (need > c->table_slots) CKR(grow_table(c, pow2_ceil(need)));
  // room for every received row (each new EC brings one): the kernel reserves what it needs on the device
  if (c->arena_used + n_rows > 0xFFFFFFFFull) return fail(c, ECB_ERR_LIMIT, "row arena exceeds 2^32 entries");
  CKR(ensure(c, c->arena, std::max<u64>(c->arena_used + n_rows, 1) * sizeof(uint2), true));
  ImportParams P{};
  P.meta = (const long long*)meta_device;
  P.rows = (const int2*)rows_device;
  P.parts = parts;
  P.n_rec = (u32)n_rec;
  P.table = (EcbEntry*)c->table.p;
  P.mask = c->table_slots - 1;
  P.ec_slot = (u32*)c->ec_slot.p;
  P.ec_rep = (u32*)c->ec_rep.p;
  P.row_len = (u32*)c->row_len.p;
  P.row_off = (u32*)c->row_off.p;
  P.arena = (uint2*)c->arena.p;
  P.ctr = c->d_ctr;
  ecb_import_insert_kernel<<<grid_for(n_rec, 256, c->sm_count * 8), 256, 0, c->stream>>>(P);
  LAUNCH_CHECK("import_insert");
  CKR(sync_counters(c));   // the one round trip of the merge: ECs and arena fill are known to the host again
  CKR(check_device_error(c));
  c->n_ec = c->h_ctr->n_ec;
  return ECB_OK;
}

int ecb_arena_create(ecb_ctx* c, int64_t cap_records, void* ipc_handle_out, void** base_out) {
  if (!c || cap_records < 1 || !base_out) return ECB_ERR_INVALID;
  if (c->xa_base) return fail(c, ECB_ERR_STATE, "this context already has an arena");
  CK(cudaSetDevice(c->device));
  const size_t bytes = ECB_ARENA_HEADER_BYTES + (size_t)cap_records * ECB_KEYREC_WORDS * 8;
  CK(cudaMalloc(&c->xa_base, bytes));
  CK(cudaMemset(c->xa_base, 0, ECB_ARENA_HEADER_BYTES));
  c->xa_cap_ec = cap_records;
  if (ipc_handle_out) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles are documented as 64 bytes");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, c->xa_base));
    memcpy(ipc_handle_out, &h, sizeof h);
  }
  *base_out = c->xa_base;
  return ECB_OK;
}

int ecb_arena_open_peer(ecb_ctx* c, const void* ipc_handle, void** base_out) {
  if (!c || !ipc_handle || !base_out) return ECB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, sizeof h);
  void* p = nullptr;
  CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  c->xa_opened.push_back(p);
  *base_out = p;
  return ECB_OK;
}

int ecb_arena_reset(ecb_ctx* c) {
  if (!c || !c->xa_base) return ECB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  CK(cudaMemsetAsync(c->xa_base, 0, ECB_ARENA_HEADER_BYTES, c->stream));
  // on a caller-provided stream (ecb_set_stream) the call is stream-ordered - a collective on that stream is the
  // barrier in front of the peers' stores; a context on its own stream synchronises
  if (c->stream == c->own_stream) CK(cudaStreamSynchronize(c->stream));
  return ECB_OK;
}

static int arena_targets(ecb_ctx* c, ArenaTargets* A, int world, void* const* bases, int64_t cap_records) {
  for (int r = 0; r < world; ++r) {
    if (!bases[r]) return fail(c, ECB_ERR_INVALID, "arena base of rank %d is NULL", r);
    char* b = (char*)bases[r];
    A->hdr[r] = (unsigned long long*)b;
    A->rec[r] = (unsigned long long*)(b + ECB_ARENA_HEADER_BYTES);
  }
  A->cap_words = (unsigned long long)cap_records * ECB_KEYREC_WORDS;
  return ECB_OK;
}

int ecb_export_to_arenas(ecb_ctx* c, int world, void* const* arena_bases, int64_t cap_records, int64_t* min_base,
                         int64_t* max_end) {
  if (!c || !arena_bases || cap_records < 1) return ECB_ERR_INVALID;
  if (world < 1 || world > ECB_MAX_WORLD) return fail(c, ECB_ERR_INVALID, "world %d outside [1, %d]", world, ECB_MAX_WORLD);
  if (c->with_cells) return fail(c, ECB_ERR_INVALID, "the multi-GPU exchange covers the single-sample path only");
  CK(cudaSetDevice(c->device));
  if (!c->table_slots) CKR(init_table(c));
  if (min_base) *min_base = c->n_ec ? (int64_t)c->min_base : 0;
  if (max_end) *max_end = c->n_ec ? (int64_t)c->max_end : 0;
  if (c->n_ec == 0) return ECB_OK;
  KeyDispatchParams P{};
  P.table = (const EcbEntry*)c->table.p;
  P.ec_slot = (const u32*)c->ec_slot.p;
  P.n_ec = c->n_ec;
  P.world = (u32)world;
  ArenaTargets A{};
  CKR(arena_targets(c, &A, world, arena_bases, cap_records));
  ecb_key_dispatch_kernel<false><<<grid_for(c->n_ec, 256, c->sm_count * 8), 256, 0, c->stream>>>(P, A);
  LAUNCH_CHECK("export_to_arenas");
  // remote stores are complete (and visible to the owner) when the kernel is.  On a caller-provided stream the
  // caller's next collective on that stream is the barrier between the stores and the owner's merge - no host
  // round trip here; a context on its own stream synchronises
  if (c->stream == c->own_stream) CK(cudaStreamSynchronize(c->stream));
  return ECB_OK;
}

int ecb_import_arena(ecb_ctx* c) {
  if (!c || !c->xa_base) return ECB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  unsigned long long hdr[3] = {0, 0, 0};
  CK(cudaMemcpyAsync(hdr, c->xa_base, sizeof hdr, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (hdr[2] || hdr[0] > (unsigned long long)c->xa_cap_ec)
    return fail(c, ECB_ERR_LIMIT, "exchange arena too small: %llu records arrived, capacity %lld", hdr[0], (long long)c->xa_cap_ec);
  const u64 n_rec = hdr[0];
  if (n_rec == 0) return ECB_OK;
  if (n_rec > 0x7FFFFFFFull) return fail(c, ECB_ERR_LIMIT, "too many records in one import");
  if (!c->table_slots) CKR(init_table(c));
  // the owner table must be able to take every incoming record as a new EC without filling up
  const u64 need = ((u64)c->n_ec + n_rec) * 2;
  if (need > c->table_slots) CKR(grow_table(c, pow2_ceil(need)));
  KeyImportParams P{};
  P.rec = (const unsigned long long*)((const char*)c->xa_base + ECB_ARENA_HEADER_BYTES);
  P.n_rec = (u32)n_rec;
  P.table = (EcbEntry*)c->table.p;
  P.mask = c->table_slots - 1;
  P.ec_slot = (u32*)c->ec_slot.p;
  P.ctr = c->d_ctr;
  ecb_import_keys_kernel<<<grid_for(n_rec, 256, c->sm_count * 8), 256, 0, c->stream>>>(P);
  LAUNCH_CHECK("import_keys");
  CKR(sync_counters(c));   // the one round trip of the merge: the number of merged ECs is known to the host again
  CKR(check_device_error(c));
  c->n_ec = c->h_ctr->n_ec;
  return ECB_OK;
}

int ecb_order_dispatch(ecb_ctx* c, int world, void* const* arena_bases, int64_t cap_records, const int64_t* shard_lo,
                       const int64_t* shard_hi) {
  if (!c || !arena_bases || !shard_lo || !shard_hi || cap_records < 1) return ECB_ERR_INVALID;
  if (world < 1 || world > ECB_MAX_WORLD) return fail(c, ECB_ERR_INVALID, "world %d outside [1, %d]", world, ECB_MAX_WORLD);
  CK(cudaSetDevice(c->device));
  if (c->n_ec == 0) return ECB_OK;
  KeyDispatchParams P{};
  P.table = (const EcbEntry*)c->table.p;
  P.ec_slot = (const u32*)c->ec_slot.p;
  P.n_ec = c->n_ec;
  P.world = (u32)world;
  // shards with alignments, by position; they must not overlap (one read order over all ranks)
  std::vector<int> order;
  for (int r = 0; r < world; ++r)
    if (shard_hi[r] > shard_lo[r]) order.push_back(r);
  std::sort(order.begin(), order.end(), [&](int a, int b) { return shard_lo[a] < shard_lo[b]; });
  if (order.empty()) return fail(c, ECB_ERR_STATE, "no rank has pushed alignments");
  for (size_t k = 0; k < order.size(); ++k) {
    if (k && shard_lo[order[k]] < shard_hi[order[k - 1]])
      return fail(c, ECB_ERR_INVALID, "the ranks' order_base ranges overlap");
    if (shard_hi[order[k]] - shard_lo[order[k]] > 0xFFFFFFFEll)
      return fail(c, ECB_ERR_LIMIT, "a rank's alignments span more than 2^32 positions");
    P.lo[k] = (u64)shard_lo[order[k]];
    P.dest[k] = (u32)order[k];
  }
  P.n_shards = (u32)order.size();
  ArenaTargets A{};
  CKR(arena_targets(c, &A, world, arena_bases, cap_records));
  ecb_key_dispatch_kernel<true><<<grid_for(c->n_ec, 256, c->sm_count * 8), 256, 0, c->stream>>>(P, A);
  LAUNCH_CHECK("order_dispatch");
  if (c->stream == c->own_stream) CK(cudaStreamSynchronize(c->stream));   // (as ecb_export_to_arenas)
  return ECB_OK;
}

int ecb_order_build(ecb_ctx* c, const ecb_ctx* local, int64_t shard_lo, int64_t shard_hi, ecb_slice* out) {
  if (!c || !local || !out || !c->xa_base || shard_hi < shard_lo) return ECB_ERR_INVALID;
  if (local->device != c->device) return fail(c, ECB_ERR_INVALID, "the local context lives on another device");
  memset(out, 0, sizeof *out);
  CK(cudaSetDevice(c->device));
  unsigned long long hdr[3] = {0, 0, 0};
  CK(cudaMemcpyAsync(hdr, c->xa_base, sizeof hdr, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (hdr[2]) return fail(c, ECB_ERR_LIMIT, "exchange arena too small for the second dispatch");
  const u64 n_local = hdr[0];
  const u64 span = (u64)(shard_hi - shard_lo);
  if (span > 0xFFFFFFFEull) return fail(c, ECB_ERR_LIMIT, "a rank's alignments span more than 2^32 positions");
  if (n_local > span) return fail(c, ECB_ERR_INVALID, "%llu ECs arrived for a shard of %llu positions", hdr[0], (unsigned long long)span);
  if (n_local > local->n_ec) return fail(c, ECB_ERR_INVALID, "%llu ECs arrived but the local context holds %u", hdr[0], local->n_ec);
  if (n_local + 1 > 0x7FFFFFFFull) return fail(c, ECB_ERR_LIMIT, "slice exceeds the int32 fields of the EC file");
  const unsigned long long* rec = (const unsigned long long*)((const char*)c->xa_base + ECB_ARENA_HEADER_BYTES);
  const size_t words = (size_t)(span / 32) + 1;
  CKR(ensure(c, c->bitmap, words * 4));
  CKR(ensure(c, c->word_rank, words * 4));
  CKR(ensure(c, c->r_a_indptr, (n_local + 1) * 4));
  CKR(ensure(c, c->r_n_data, std::max<u64>(n_local, 1) * 4));
  CKR(ensure(c, c->first_rel, std::max<u64>(n_local, 1) * 8));   // scratch: local id per record, then per slice id
  // every row of the slice is a row of the local context: its arena fill bounds the non-zeros
  const u64 z_bound = std::max<u64>(local->arena_used, 1);
  CKR(ensure(c, c->r_a_indices, z_bound * 4));
  CKR(ensure(c, c->r_a_data, z_bound * 4));
  u32* lid_of_rec = (u32*)c->first_rel.p;
  u32* lid_of_id = lid_of_rec + n_local;
  CK(cudaMemsetAsync(c->bitmap.p, 0, words * 4, c->stream));
  CK(cudaMemsetAsync(c->r_a_indptr.p, 0, (n_local + 1) * 4, c->stream));
  CK(cudaMemsetAsync(&c->d_ctr->scratch[5], 0, sizeof(u32), c->stream));
  u64 nnz = 0;
  if (n_local) {
    const int g = grid_for(n_local, 256, c->sm_count * 8);
    ecb_order_mark_kernel<<<g, 256, 0, c->stream>>>(rec, (u32)n_local, (u32)span, (const EcbEntry*)local->table.p,
                                                    local->table_slots - 1, (u32*)c->bitmap.p, lid_of_rec, &c->d_ctr->scratch[5]);
    LAUNCH_CHECK("order_mark");
    CKR(device_scan<true>(c, (const u32*)c->bitmap.p, (u32*)c->word_rank.p, words, 0, nullptr));
    ecb_order_lens_kernel<<<g, 256, 0, c->stream>>>(rec, (u32)n_local, (u32)span, (const u32*)c->bitmap.p,
                                                    (const u32*)c->word_rank.p, lid_of_rec, (const u32*)local->row_len.p,
                                                    (int32_t*)c->r_a_indptr.p, (int32_t*)c->r_n_data.p, lid_of_id);
    LAUNCH_CHECK("order_lens");
    CKR(device_scan<false>(c, (const u32*)c->r_a_indptr.p, (u32*)c->r_a_indptr.p, n_local + 1, 0, nullptr));
    {
      const u32 n_blocks = (u32)((n_local + 1 + SCAN_BLOCK - 1) / SCAN_BLOCK);
      CK(cudaMemcpyAsync(c->h_total, (const u64*)c->scan_partials.p + n_blocks, sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    }
    ecb_order_rows_kernel<<<grid_for(n_local, 256, c->sm_count * 16), 256, 0, c->stream>>>(
        lid_of_id, (const u32*)local->row_len.p, (const u32*)local->row_off.p, (const uint2*)local->arena.p,
        (const int32_t*)c->r_a_indptr.p, (u32)n_local, (int32_t*)c->r_a_indices.p, (int32_t*)c->r_a_data.p);
    LAUNCH_CHECK("order_rows");
  }
  CKR(sync_counters(c));
  if (c->h_ctr->scratch[5])
    return fail(c, ECB_ERR_INVALID, "second dispatch: %s", (c->h_ctr->scratch[5] & 4u) ? "an EC arrived that the local context has never seen"
                : (c->h_ctr->scratch[5] & 2u) ? "two ECs share a first read" : "a record outside the shard's positions arrived");
  if (n_local) nnz = *c->h_total;
  if (nnz > 0x7FFFFFFFull) return fail(c, ECB_ERR_LIMIT, "slice exceeds the int32 fields of the EC file");
  out->id_base = 0;   // the caller knows where this rank's id range starts (the ECs of the shards in front)
  out->n_ec = (int64_t)n_local;
  out->nnz = (int64_t)nnz;
  out->a_indptr = (const int32_t*)c->r_a_indptr.p;
  out->a_indices = (const int32_t*)c->r_a_indices.p;
  out->a_data = (const int32_t*)c->r_a_data.p;
  out->n_data = (const int32_t*)c->r_n_data.p;
  return ECB_OK;
}

static FinalizeParams global_params(ecb_ctx* c) {
  FinalizeParams F{};
  F.table = (const EcbEntry*)c->table.p;
  F.ec_slot = (const u32*)c->ec_slot.p;
  F.row_len = (const u32*)c->row_len.p;
  F.row_off = (const u32*)c->row_off.p;
  F.arena = (const uint2*)c->arena.p;
  F.n_ec = c->n_ec;
  F.min_base = c->g_min_base;
  F.first_rel = (u64*)c->first_rel.p;
  F.ecid_of = (u32*)c->ecid_of.p;
  F.count_of = (u32*)c->count_of.p;
  return F;
}

int ecb_global_mark(ecb_ctx* c, int64_t min_base, uint32_t* bitmap_device, int64_t n_words) {
  if (!c || !bitmap_device || n_words <= 0 || min_base < 0) return ECB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  c->g_min_base = (u64)min_base;
  if (c->n_ec == 0) return ECB_OK;
  CKR(ensure(c, c->first_rel, (size_t)c->n_ec * 8));
  CKR(ensure(c, c->ecid_of, (size_t)c->n_ec * 4));
  CKR(ensure(c, c->count_of, (size_t)c->n_ec * 4));
  FinalizeParams F = global_params(c);
  F.bitmap = bitmap_device;
  ecb_fin_mark_kernel<<<grid_for(c->n_ec, 256, c->sm_count * 8), 256, 0, c->stream>>>(F);
  LAUNCH_CHECK("fin_mark");
  CK(cudaStreamSynchronize(c->stream));
  return ECB_OK;
}

int ecb_global_count(ecb_ctx* c, const uint32_t* bitmap_device, int64_t n_words, int64_t* n_ec_total) {
  if (!c || !bitmap_device || n_words <= 0 || !n_ec_total) return ECB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  CKR(ensure(c, c->word_rank, (size_t)n_words * 4));
  CKR(ensure(c, c->bitmap, (size_t)n_words * 4));
  CK(cudaMemcpyAsync(c->bitmap.p, bitmap_device, (size_t)n_words * 4, cudaMemcpyDeviceToDevice, c->stream));
  u64 total = 0;
  CKR(device_scan<true>(c, (const u32*)c->bitmap.p, (u32*)c->word_rank.p, (u64)n_words, 0, &total));
  c->g_n_ec_total = total;
  *n_ec_total = (int64_t)total;
  return ECB_OK;
}

int ecb_global_lens(ecb_ctx* c, int32_t* lens_device, int32_t* counts_device) {
  if (!c || !lens_device || !counts_device) return ECB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  if (c->n_ec == 0) return ECB_OK;
  FinalizeParams F = global_params(c);
  F.bitmap = (u32*)c->bitmap.p;
  F.word_rank = (const u32*)c->word_rank.p;
  F.a_indptr = lens_device;
  F.n_data = counts_device;
  CKR(ensure(c, c->r_n_indices, (size_t)c->g_n_ec_total * 4));
  F.n_indices = (int32_t*)c->r_n_indices.p;
  F.wide_count = &c->d_ctr->scratch[2];        // rows with more than 8 entries are listed for ecb_global_rows
  F.wide_list = (u32*)c->long_list.p;
  CK(cudaMemsetAsync(&c->d_ctr->scratch[2], 0, sizeof(u32), c->stream));
  ecb_fin_rank_kernel<true><<<grid_for(c->n_ec, 256, c->sm_count * 8), 256, 0, c->stream>>>(F);
  LAUNCH_CHECK("fin_rank");
  CKR(sync_counters(c));
  c->g_n_wide = c->h_ctr->scratch[2];
  return ECB_OK;
}

int ecb_global_indptr(ecb_ctx* c, int32_t* lens_device, int64_t n, int64_t* nnz) {
  if (!c || !lens_device || n < 1 || !nnz) return ECB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  u64 total = 0;
  CKR(device_scan<false>(c, (const u32*)lens_device, (u32*)lens_device, (u64)n, 0, &total));
  if (total > 0x7FFFFFFFull) return fail(c, ECB_ERR_LIMIT, "A matrix has more than 2^31-1 non-zeros");
  *nnz = (int64_t)total;
  return ECB_OK;
}

int ecb_global_rows(ecb_ctx* c, const int32_t* indptr_device, int32_t* indices_device, int32_t* data_device) {
  if (!c || !indptr_device || !indices_device || !data_device) return ECB_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  if (c->n_ec == 0) return ECB_OK;
  FinalizeParams F = global_params(c);
  F.a_indptr = const_cast<int32_t*>(indptr_device);
  F.a_indices = indices_device;
  F.a_data = data_device;
  ecb_fin_rows_kernel<<<grid_for(c->n_ec, 256, c->sm_count * 16), 256, 0, c->stream>>>(F);
  LAUNCH_CHECK("fin_rows");
  if (c->g_n_wide) {
    ecb_fin_rows_long_kernel<<<grid_for((u64)c->g_n_wide * 32, 256, c->sm_count * 8), 256, 0, c->stream>>>(
        F, (const u32*)c->long_list.p, c->g_n_wide);
    LAUNCH_CHECK("fin_rows_long");
  }
  CK(cudaStreamSynchronize(c->stream));
  return ECB_OK;
}

int ecb_destroy(ecb_ctx* c) {
  if (!c) return ECB_OK;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  DevBuf* bufs[] = {&c->table, &c->ec_slot, &c->ec_rep, &c->ec_len, &c->spill, &c->row_len, &c->row_off, &c->arena, &c->long_list, &c->mid_list, &c->big_list, &c->count_of,
                    &c->ttable, &c->st_rg, &c->st_tg, &c->st_hp, &c->st_cell, &c->st2_rg, &c->st2_tg, &c->st2_hp, &c->st2_cell,
                    &c->overflow_bits,
                    &c->scan_partials, &c->bitmap, &c->word_rank, &c->first_rel, &c->ecid_of, &c->ec_keep,
                    &c->r_a_indptr, &c->r_a_indices, &c->r_a_data, &c->r_n_indptr, &c->r_n_indices,
                    &c->r_n_data, &c->r_cell_order, &c->x_meta, &c->x_rows, &c->x_counts, &c->x_base};
  for (DevBuf* b : bufs) release(c, *b);
  cells_release(c);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->copy_stream) {
    cudaStreamSynchronize(c->copy_stream);
    cudaStreamDestroy(c->copy_stream);
    cudaEventDestroy(c->copy_done[0]);
    cudaEventDestroy(c->copy_done[1]);
  }
  for (void* p : c->xa_opened) cudaIpcCloseMemHandle(p);
  if (c->xa_base) cudaFree(c->xa_base);
  if (c->d_ctr) cudaFree(c->d_ctr);
  if (c->h_ctr) cudaFreeHost(c->h_ctr);
  if (c->h_total) cudaFreeHost(c->h_total);
  if (c->h_res) {
    if (c->pageable_results) free(c->h_res); else cudaFreeHost(c->h_res);
  }
  for (auto& e : c->ev)
    if (e) cudaEventDestroy(e);
  if (c->own_stream) {
    cudaStreamSynchronize(c->own_stream);
    cudaStreamDestroy(c->own_stream);
  }
  if (c->pool) cudaMemPoolDestroy(c->pool);
  delete c;
  return ECB_OK;
}

}  // extern "C"
