/*
 * ecb200.h — C ABI of libecb200.so: B200 (sm_100a) equivalence-class builder for alntools bam2ec.
 *
 * The reference (churchill-lab/alntools) is pure Python and has no FFI; the seam this library
 * replaces is the body of its per-alignment grouping loop, the chunk merge, the EC -> COO loop and
 * the scipy canonicalisation, fed with int32 columns instead of pysam objects:
 *
 *   ecb_push      replaces  alntools/bam_utils.py:258-344  (process_convert_bam hot loop: group by
 *                           read, set of tids per read, ec[key] += 1) and the chunk-ordered merge
 *                           alntools/bam_utils.py:680-698 (EC id = rank of first occurrence);
 *                           with cells: alntools/bam_utils_multisample.py:209-300 and :503-560.
 *   ecb_finalize  replaces  alntools/bam_utils.py:788-847 (EC -> per-haplotype COO -> APM) plus the
 *                           matrix part of alntools/bin_utils.py:208-232,246-275 (sum 2^h*data[h],
 *                           tocsr, N as CSC); with cells: bam_utils_multisample.py:595-636,702-791.
 *
 * Host code (alntools_b200/*.py) keeps the reference's convert()/methods/CLI signatures, decodes
 * the BAM, applies the filters (bam_utils.py:264-270) and the read-name comparison (:301-306), and
 * hands over columns.  No torch types cross this boundary: plain pointers and sizes only.
 *
 * All functions return 0 on success and a negative ecb_status on failure; the message is available
 * through ecb_last_error().  Nothing here ever falls back to a CPU implementation.
 *
 * Threading: a context is bound to one CUDA device and one stream and is not thread-safe; use one
 * context per GPU.  Calls block the calling thread until the work they describe is finished unless
 * stated otherwise (ctypes releases the GIL around them).
 */
#ifndef ECB200_H
#define ECB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ecb_ctx ecb_ctx;

enum ecb_status {
  ECB_OK = 0,
  ECB_ERR_INVALID = -1,     /* bad argument / contract violation */
  ECB_ERR_CUDA = -2,        /* CUDA runtime error (message has the CUDA string) */
  ECB_ERR_NO_DEVICE = -3,   /* no usable sm_100 device */
  ECB_ERR_LIMIT = -4,       /* a format limit was exceeded (int32 file fields, read too long, ...) */
  ECB_ERR_EMPTY = -5,       /* finalize with zero equivalence classes (the reference raises too:
                               alntools/matrix/Sparse3DMatrix.py:45-46) */
  ECB_ERR_STATE = -6        /* call sequence error */
};

/* Longest read (alignments sharing one name) whose canonical row can be built. */
#define ECB_MAX_READ_ALIGNMENTS 16384
/* target_idx < 2^26 and hap_idx < 31 (the EC file stores the haplotype mask in an int32,
 * alntools/bin_utils.py:232). */
#define ECB_MAX_TARGETS (1 << 26)
#define ECB_MAX_HAPS 31

enum ecb_option {
  ECB_OPT_RESULT_ON_DEVICE = 1, /* 1: ecb_result pointers are DEVICE pointers (no D2H copy) */
  ECB_OPT_TABLE_SLOTS = 2,      /* initial EC hash-table capacity (rounded up to a power of two) */
  ECB_OPT_PAIR_SLOTS = 3,       /* initial (file, EC, cell) table capacity */
  ECB_OPT_GRID_CTAS = 4,        /* CTAs of the grouping kernel (0 = one persistent CTA per SM) */
  ECB_OPT_HOT_CACHE = 5,        /* 1 (default): per-CTA shared-memory cache of hot ECs in front of the HBM table */
  ECB_OPT_VERIFY_KEYS = 6,      /* 1: every push re-derives each read's set of (target, haplotype) pairs and
                                   compares it with the row of the EC the read was counted in, turning
                                   a 128-bit key collision into ECB_ERR_LIMIT (slow; a debugging aid) */
  ECB_OPT_CHUNK_LEN = 7,        /* alignments per work chunk of the grouping kernel (0 = automatic) */
  ECB_OPT_PAGEABLE_RESULTS = 8, /* 1: host results go to ordinary (malloc) memory instead of pinned memory:
                                   cheaper for a context that finalizes once, slower when reused */
};

typedef struct ecb_result {
  int64_t n_ec;              /* E: number of equivalence classes (after the cell filter)          */
  int64_t nnz_a;             /* Z                                                                 */
  const int32_t* a_indptr;   /* [E+1]  CSR row offsets of the A matrix (E x T)                    */
  const int32_t* a_indices;  /* [Z]    main-target index, ascending inside a row                  */
  const int32_t* a_data;     /* [Z]    haplotype bitmask, bit h = sorted-haplotype index h        */
  int64_t n_samples;         /* S: 1 without cells; number of kept cells with cells               */
  int64_t nnz_n;
  const int32_t* n_indptr;   /* [S+1]  CSC column offsets of the N matrix (E x S)                 */
  const int32_t* n_indices;  /* [nnz_n] EC ids, ascending inside a column                         */
  const int32_t* n_data;     /* [nnz_n] read counts                                               */
  const int32_t* cell_order; /* [S] original cell_idx of every output column; NULL without cells  */
  int64_t n_reads;           /* reads counted over all pushes (after drop_last_group)             */
  int64_t n_alignments;      /* alignments pushed                                                 */
} ecb_result;

typedef struct ecb_stats {
  double group_ms;        /* device time of the grouping/insert kernel(s) of the last push        */
  double harvest_ms;      /* device time of the row-harvest kernels of the last push              */
  double push_ms;         /* device time of the whole last push (incl. H2D when host pointers)    */
  double finalize_ms;     /* device time of the last finalize (incl. D2H unless result on device) */
  int64_t kernel_launches;/* kernels launched by this library since create/reset                  */
  int64_t table_slots;    /* current EC table capacity                                            */
  int64_t table_used;     /* ECs in the table                                                     */
  int64_t table_grows;    /* rehash events                                                        */
  int64_t overflow_reads; /* reads that had to be replayed after a table growth                   */
  int64_t h2d_bytes;      /* bytes copied host->device since create/reset                         */
  int64_t d2h_bytes;      /* bytes copied device->host since create/reset                         */
  int64_t row_entries;    /* (target, mask) pairs reserved in the row arena (upper bound of nnz)   */
} ecb_stats;

/* Library/ABI version (major*1000 + minor). */
int ecb_version(void);

/* Create a context on CUDA device `device`.  n_targets/n_haps bound the column values;
 * with_cells selects the per-cell (multisample) path; alignments_hint sizes tables and staging. */
int ecb_create(ecb_ctx** out, int device, int n_targets, int n_haps, int with_cells,
               int64_t alignments_hint);

int ecb_set_option(ecb_ctx* ctx, int option, int64_t value);

/* Run all work of this context on an existing CUDA stream (a cudaStream_t passed as void*),
 * e.g. torch.cuda.current_stream().cuda_stream.  NULL restores the context's own stream; the legacy
 * default stream must be named explicitly (cudaStreamLegacy), because its plain handle is NULL too. */
int ecb_set_stream(ecb_ctx* ctx, void* cuda_stream);

/*
 * Push one contiguous run of VALID alignments (already filtered, name-grouped).
 *   read_group[n]  consecutive equal values form one read; values only need to change between reads
 *   target_idx[n]  main-target index of the alignment's reference   (0 <= v < n_targets)
 *   hap_idx[n]     index into the SORTED haplotype list             (0 <= v < n_haps)
 *   cell_idx[n]    cell id of the alignment's read (value on the read's first alignment is used);
 *                  NULL unless the context was created with_cells
 *   order_base     global index of this push's first alignment in reference read order: EC ids are
 *                  ranked by order_base + offset of the EC's first read, so pushes/shards may arrive
 *                  in any order as long as their [order_base, order_base+n) ranges do not overlap
 *   drop_last_group 1 reproduces alntools/bam_utils_multisample.py:306-308 (the last read of a file
 *                  is never flushed); one push = one file in that case
 *   on_device      0: host pointers (pinned preferred, copied with cudaMemcpyAsync);
 *                  1: device pointers on this context's device
 * Reads must not span pushes.  Buffers are caller-owned and may be reused after return.
 */
int ecb_push(ecb_ctx* ctx, const int32_t* read_group, const int32_t* target_idx,
             const int32_t* hap_idx, const int32_t* cell_idx, int64_t n, int64_t order_base,
             int drop_last_group, int on_device);

/* Build the A (CSR) and N (CSC) matrices.  min_cell_count follows
 * alntools/bam_utils_multisample.py:596-608 (<=0 means 1; ignored without cells).
 * Result buffers are library-owned (pinned host memory, or device memory with
 * ECB_OPT_RESULT_ON_DEVICE) and stay valid until the next finalize/reset/destroy. */
int ecb_finalize(ecb_ctx* ctx, int64_t min_cell_count, ecb_result* out);

/* Forget all pushed data but keep allocations (tables, staging) for the next job. */
int ecb_reset(ecb_ctx* ctx);

int ecb_get_stats(const ecb_ctx* ctx, ecb_stats* out);

int ecb_destroy(ecb_ctx* ctx);

/* Message of the last failing call on this context (or of ecb_create when ctx is NULL). */
const char* ecb_last_error(const ecb_ctx* ctx);

/* ---- multi-GPU exchange -------------------------------------------------------------------------
 * Reads shard across GPUs by contiguous chunk (alntools/utils.py:67-87 + bam_utils.py:647-680: each
 * worker gets a contiguous run of chunks and results are merged in chunk order).  Each rank pushes its
 * shard into a LOCAL context with order_base = global offset of the shard.  The merge of
 * alntools/bam_utils.py:680-698 then becomes:
 *   1. ecb_export_partition on the local context: local ECs packed into `world` partitions by owner
 *      rank (hash of the 128-bit key);
 *   2. the host exchanges the partitions (NCCL all-to-all through torch.distributed);
 *   3. ecb_import_entries on a fresh OWNER context: equal keys are merged (counts summed, smallest
 *      first-occurrence kept), rows adopted;
 *   4. ecb_global_mark / all-reduce(bitmap) / ecb_global_count: EC ids = rank of the first-occurrence
 *      bit in the bitmap OR-ed over all ranks;
 *   5. ecb_global_lens / all-reduce / ecb_global_indptr / ecb_global_rows / all-reduce: every rank
 *      scatters its owned rows and counts into zero-initialised global arrays; supports are disjoint,
 *      so a SUM all-reduce assembles the final CSR A matrix and the counts on every rank.
 * That is the form for any transport (NCCL, gloo in the CPU tests); it leaves the whole matrix on every rank
 * (or on rank 0).  Where the ranks can map each other's memory the exchange further down replaces all five
 * steps and leaves the final matrices partitioned, one contiguous EC-id range per rank.
 * All pointers marked "device" are device memory of the context's GPU owned by the caller.
 */
/* Add `delta` (>= 0) to the order key of everything pushed so far: a rank that decodes one shard of a file
 * pushes with order_base counted from its own first alignment and is moved to its global position once every
 * rank knows how many alignments the shards in front of it hold.  Single-sample contexts, before the exchange. */
int ecb_rebase(ecb_ctx* ctx, int64_t delta);

#define ECB_EXPORT_META_WORDS 5
typedef struct ecb_export {
  int64_t n_ec;                    /* local ECs exported                                           */
  int64_t n_rows;                  /* (target, mask) pairs exported                                */
  const int64_t* meta;             /* device [n_ec * 5]: key_lo, key_hi, first, count<<32|row_len,
                                      row offset inside its partition                              */
  const int32_t* rows;             /* device [n_rows * 2]: (target_idx, hap_mask), partition-major */
  const int64_t* part_ec_counts;   /* host [world]                                                 */
  const int64_t* part_row_counts;  /* host [world]                                                 */
  int64_t min_base, max_end;       /* order-key range this rank has pushed                         */
} ecb_export;

int ecb_export_partition(ecb_ctx* ctx, int world, ecb_export* out);

/* Merge n_parts received partitions (concatenated, in source-rank order) into this OWNER context. */
int ecb_import_entries(ecb_ctx* ctx, const int64_t* meta_device, const int32_t* rows_device,
                       const int64_t* part_ec_counts, const int64_t* part_row_counts, int n_parts);

/* ---- the exchange over peer memory (one process per GPU on an NVLink / NVSwitch box; the default of the
 * multi-GPU path) ---------------------------------------------------------------------------------------
 * Each rank creates an ARENA on its OWNER context (device memory from cudaMalloc, exportable through CUDA IPC),
 * hands the 64-byte IPC handle to its peers and maps theirs.  Two dispatches of fixed-size records go through
 * it, each one kernel that stores straight into the destination ranks' arenas over NVLink (one remote
 * atomicAdd per CTA tile and destination reserves the space):
 *   ecb_export_to_arenas (LOCAL context)   every local EC as {key, first, count} to its owner rank (hash of the
 *                                          key);
 *   ecb_import_arena     (OWNER context)   the owner merges what arrived: counts summed, smallest first
 *                                          occurrence kept (alntools/bam_utils.py:693-698);
 *   ecb_order_dispatch   (OWNER context)   every merged EC as {key, position inside the shard, count} to the rank
 *                                          whose shard [shard_lo[r], shard_hi[r]) of the global read order holds
 *                                          its first occurrence;
 *   ecb_order_build      (OWNER + LOCAL)   EC ids are ranks of first-occurrence positions and the shards
 *                                          partition the positions, so what arrived is one contiguous id range:
 *                                          it is ordered with a bitmap over this rank's OWN positions and the
 *                                          rows are taken from the LOCAL context - the rank that holds an EC's
 *                                          first occurrence has met the EC in its own reads.  Rows never travel;
 *                                          nothing global is built (no bitmap over all positions, no all-reduce,
 *                                          no ecb_global_* call).  id_base of the result is left 0: the slice
 *                                          starts after the ECs of the shards in front, which the caller learns
 *                                          from one all-gather of the slice sizes.
 * The caller brackets the stores with barriers:
 *     ecb_arena_reset (every rank) - barrier - ecb_export_to_arenas - barrier - ecb_import_arena -
 *     ecb_arena_reset - barrier - ecb_order_dispatch - barrier - ecb_order_build.
 * On a caller-provided stream (ecb_set_stream) reset, export and dispatch are stream-ordered and the barriers
 * are collectives on that stream (the host does not wait); a context on its own stream synchronises in each.
 * arena_bases[r] = base address of rank r's arena as seen from THIS process (own base for r == rank);
 * cap_records = records (of 32 bytes) every arena can take. */
int ecb_arena_create(ecb_ctx* owner_ctx, int64_t cap_records, void* ipc_handle_out /* 64 bytes */, void** base_out);
int ecb_arena_open_peer(ecb_ctx* owner_ctx, const void* ipc_handle /* 64 bytes */, void** base_out);
int ecb_arena_reset(ecb_ctx* owner_ctx);
int ecb_export_to_arenas(ecb_ctx* local_ctx, int world, void* const* arena_bases, int64_t cap_records,
                         int64_t* min_base, int64_t* max_end);
int ecb_import_arena(ecb_ctx* owner_ctx);

/* This rank's slice of the final matrices: a_indptr[n_ec + 1] (offsets local to the slice), a_indices, a_data
 * and n_data (read counts), all device memory of the owner context, valid until the next call on it. */
typedef struct ecb_slice {
  int64_t id_base, n_ec, nnz;
  const int32_t *a_indptr, *a_indices, *a_data, *n_data;
} ecb_slice;
int ecb_order_dispatch(ecb_ctx* owner_ctx, int world, void* const* arena_bases, int64_t cap_records,
                       const int64_t* shard_lo, const int64_t* shard_hi);
int ecb_order_build(ecb_ctx* owner_ctx, const ecb_ctx* local_ctx, int64_t shard_lo, int64_t shard_hi, ecb_slice* out);

/* Set the first-occurrence bits of the ECs this context owns in bitmap[n_words] (bit i = order key
 * min_base + i).  The caller zero-fills the bitmap and all-reduces it afterwards. */
int ecb_global_mark(ecb_ctx* ctx, int64_t min_base, uint32_t* bitmap_device, int64_t n_words);

/* Rank the bits of the (all-reduced) bitmap: *n_ec_total = number of set bits = global EC count. */
int ecb_global_count(ecb_ctx* ctx, const uint32_t* bitmap_device, int64_t n_words, int64_t* n_ec_total);

/* Scatter row lengths and read counts of the owned ECs into zero-filled lens[n_ec_total+1] and
 * counts[n_ec_total] at their global ids. */
int ecb_global_lens(ecb_ctx* ctx, int32_t* lens_device, int32_t* counts_device);

/* In-place exclusive scan lens -> indptr (n = n_ec_total + 1); *nnz = indptr[n_ec_total]. */
int ecb_global_indptr(ecb_ctx* ctx, int32_t* lens_device, int64_t n, int64_t* nnz);

/* Scatter the owned rows into zero-filled indices[nnz] / data[nnz] at indptr[global id]. */
int ecb_global_rows(ecb_ctx* ctx, const int32_t* indptr_device, int32_t* indices_device,
                    int32_t* data_device);

#ifdef __cplusplus
}
#endif
#endif /* ECB200_H */
