/*
 * ec_oracle.c — CPU ORACLE (C restatement) of the column-level EC build.  *** TEST INFRASTRUCTURE ***
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.  It restates,
 * on the int32 columns the GPU kernels consume, what the reference computes:
 *   - a read = maximal run of equal read_group values; its key = the SET of (target, haplotype)
 *     pairs of its alignments, duplicates collapsed   (alntools/bam_utils.py:306-325)
 *   - ec[key] += 1, EC id = rank of the key's first occurrence in read order (:310-312, :693-698)
 *   - A row of an EC: its targets ascending, data = OR of 1 << haplotype
 *     (alntools/bam_utils.py:788-825 + alntools/bin_utils.py:208-211 tocsr())
 *   - N for one sample: counts in EC order (alntools/bam_utils.py:845)
 *   - per-cell (multisample) form, ec_oracle_build_cells: the merge loops of
 *     alntools/bam_utils_multisample.py:503-636 and the N matrix of :737-747,783-791, see below
 * Pinned by tests/test_oracle_golden.py against oracle/ec_oracle.py, which is itself pinned against
 * EC files written by the unmodified reference (tests/golden/).
 *
 * Sequential scan + chained hash table keyed by the sorted unique code list; nothing clever.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int64_t key_off;  /* offset of the key (sorted unique codes) in the key arena */
  int32_t key_len;
  int32_t count;
  int64_t next;     /* chain */
} ec_rec;

typedef struct {
  int64_t n_ec, nnz;
  int32_t* indptr;   /* [n_ec+1] */
  int32_t* indices;  /* [nnz] */
  int32_t* data;     /* [nnz] */
  int32_t* counts;   /* [n_ec] */
  int64_t n_reads;
} ec_oracle_result;

static int cmp_i64(const void* a, const void* b) {
  int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
  return (x > y) - (x < y);
}

static uint64_t hash_codes(const int64_t* k, int n) {
  uint64_t h = 1469598103934665603ull;
  for (int i = 0; i < n; ++i) {
    uint64_t v = (uint64_t)k[i];
    for (int b = 0; b < 8; ++b) {
      h ^= (v >> (8 * b)) & 0xff;
      h *= 1099511628211ull;
    }
  }
  return h;
}

void ec_oracle_free(ec_oracle_result* r) {
  free(r->indptr); free(r->indices); free(r->data); free(r->counts);
  memset(r, 0, sizeof *r);
}

/* drop_last: do not count the last read (alntools/bam_utils_multisample.py:306-308). Returns 0 on success. */
int ec_oracle_build(const int32_t* rg, const int32_t* tg, const int32_t* hp, int64_t n, int drop_last,
                    ec_oracle_result* out) {
  memset(out, 0, sizeof *out);
  int64_t cap_rec = 1024, n_rec = 0;
  ec_rec* recs = (ec_rec*)malloc(cap_rec * sizeof(ec_rec));
  int64_t cap_arena = 4096, arena_used = 0;
  int64_t* arena = (int64_t*)malloc(cap_arena * sizeof(int64_t));
  int64_t n_buckets = 1 << 16;
  int64_t* buckets = (int64_t*)malloc(n_buckets * sizeof(int64_t));
  for (int64_t i = 0; i < n_buckets; ++i) buckets[i] = -1;
  int64_t cap_tmp = 1024;
  int64_t* tmp = (int64_t*)malloc(cap_tmp * sizeof(int64_t));
  int64_t n_reads = 0;

  int64_t s = 0;
  while (s < n) {
    int64_t e = s + 1;
    while (e < n && rg[e] == rg[s]) ++e;
    if (drop_last && e == n) break;
    int64_t k = e - s;
    if (k > cap_tmp) { cap_tmp = k * 2; tmp = (int64_t*)realloc(tmp, cap_tmp * sizeof(int64_t)); }
    for (int64_t i = 0; i < k; ++i) tmp[i] = (int64_t)tg[s + i] * 64 + hp[s + i];
    qsort(tmp, (size_t)k, sizeof(int64_t), cmp_i64);
    int m = 0;
    for (int64_t i = 0; i < k; ++i)
      if (i == 0 || tmp[i] != tmp[i - 1]) tmp[m++] = tmp[i];
    uint64_t h = hash_codes(tmp, m);
    int64_t b = (int64_t)(h & (uint64_t)(n_buckets - 1));
    int64_t r = buckets[b];
    while (r >= 0) {
      if (recs[r].key_len == m && memcmp(arena + recs[r].key_off, tmp, (size_t)m * 8) == 0) break;
      r = recs[r].next;
    }
    if (r < 0) {
      if (n_rec == cap_rec) { cap_rec *= 2; recs = (ec_rec*)realloc(recs, cap_rec * sizeof(ec_rec)); }
      if (arena_used + m > cap_arena) {
        while (arena_used + m > cap_arena) cap_arena *= 2;
        arena = (int64_t*)realloc(arena, cap_arena * sizeof(int64_t));
      }
      memcpy(arena + arena_used, tmp, (size_t)m * 8);
      r = n_rec++;
      recs[r].key_off = arena_used;
      recs[r].key_len = m;
      recs[r].count = 0;
      recs[r].next = buckets[b];
      buckets[b] = r;
      arena_used += m;
      if (n_rec * 2 > n_buckets) { /* rehash */
        n_buckets *= 4;
        buckets = (int64_t*)realloc(buckets, n_buckets * sizeof(int64_t));
        for (int64_t i = 0; i < n_buckets; ++i) buckets[i] = -1;
        for (int64_t i = 0; i < n_rec; ++i) {
          uint64_t hh = hash_codes(arena + recs[i].key_off, recs[i].key_len);
          int64_t bb = (int64_t)(hh & (uint64_t)(n_buckets - 1));
          recs[i].next = buckets[bb];
          buckets[bb] = i;
        }
      }
    }
    recs[r].count += 1;
    ++n_reads;
    s = e;
  }

  /* rows: targets ascending (codes are sorted by target*64+hap), mask = OR of 1<<hap */
  int64_t nnz = 0;
  for (int64_t i = 0; i < n_rec; ++i) {
    const int64_t* k = arena + recs[i].key_off;
    for (int j = 0; j < recs[i].key_len; ++j)
      if (j == 0 || (k[j] >> 6) != (k[j - 1] >> 6)) ++nnz;
  }
  out->n_ec = n_rec;
  out->nnz = nnz;
  out->n_reads = n_reads;
  out->indptr = (int32_t*)malloc((size_t)(n_rec + 1) * 4);
  out->indices = (int32_t*)malloc((size_t)(nnz ? nnz : 1) * 4);
  out->data = (int32_t*)malloc((size_t)(nnz ? nnz : 1) * 4);
  out->counts = (int32_t*)malloc((size_t)(n_rec ? n_rec : 1) * 4);
  int64_t z = 0;
  out->indptr[0] = 0;
  for (int64_t i = 0; i < n_rec; ++i) {
    const int64_t* k = arena + recs[i].key_off;
    for (int j = 0; j < recs[i].key_len; ++j) {
      if (j == 0 || (k[j] >> 6) != (k[j - 1] >> 6)) {
        out->indices[z] = (int32_t)(k[j] >> 6);
        out->data[z] = 0;
        ++z;
      }
      out->data[z - 1] |= (int32_t)(1u << (k[j] & 63));
    }
    out->indptr[i + 1] = (int32_t)z;
    out->counts[i] = recs[i].count;
  }
  free(recs); free(arena); free(buckets); free(tmp);
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Per-cell (multisample) form.  Follows the reference's loops, not a derived formula:
 *   worker  (bam_utils_multisample.py:288-300): per FILE an insertion-ordered map key -> insertion-ordered
 *           map cell -> count; the last read of every file is never flushed (:306-308, drop_last);
 *   merge   (:503-560): files in order, keys in the file's first-occurrence order, cells in first-occurrence
 *           order inside that (file, key): cr_totals[cell] += count (cells are ordered by their FIRST
 *           insertion here), final[key][cell] += count, EC ids by first insertion into final;
 *   filter  (:595-636): minimum_count <= 0 -> 1; cells with total >= minimum keep their order; an EC
 *           survives when one of its cells does; surviving ECs keep their order;
 *   N       (:737-747,783-791): CSC over the kept cells, EC ids ascending inside a column.
 * The cell of a read is the value of cell[] on its first alignment.
 */
typedef struct {
  int64_t n_ec, nnz_a;
  int32_t *a_indptr, *a_indices, *a_data;
  int64_t n_cells, nnz_n;
  int32_t *n_indptr, *n_indices, *n_data, *cell_order;
  int64_t n_reads;
} ec_cells_result;

void ec_oracle_cells_free(ec_cells_result* r) {
  free(r->a_indptr); free(r->a_indices); free(r->a_data);
  free(r->n_indptr); free(r->n_indices); free(r->n_data); free(r->cell_order);
  memset(r, 0, sizeof *r);
}

typedef struct { int64_t a; int32_t b; int32_t count; int64_t next_in_list; int64_t chain; } pair_rec;

typedef struct {   /* insertion-ordered multimap (a, b) -> count with chained buckets */
  pair_rec* v; int64_t n, cap; int64_t* buckets; int64_t n_buckets;
} pair_map;

static void pm_init(pair_map* m) {
  m->cap = 1024; m->n = 0; m->v = (pair_rec*)malloc(m->cap * sizeof(pair_rec));
  m->n_buckets = 1 << 12; m->buckets = (int64_t*)malloc(m->n_buckets * 8);
  for (int64_t i = 0; i < m->n_buckets; ++i) m->buckets[i] = -1;
}
static void pm_clear(pair_map* m) { m->n = 0; for (int64_t i = 0; i < m->n_buckets; ++i) m->buckets[i] = -1; }
static uint64_t pm_hash(int64_t a, int32_t b) {
  uint64_t h = (uint64_t)a * 0x9E3779B97F4A7C15ull ^ ((uint64_t)(uint32_t)b * 0xC2B2AE3D27D4EB4Full);
  return h ^ (h >> 29);
}
/* index of (a, b), inserted with count 0 when absent (*is_new says which) */
static int64_t pm_get(pair_map* m, int64_t a, int32_t b, int* is_new) {
  int64_t bk = (int64_t)(pm_hash(a, b) & (uint64_t)(m->n_buckets - 1));
  for (int64_t r = m->buckets[bk]; r >= 0; r = m->v[r].chain)
    if (m->v[r].a == a && m->v[r].b == b) { *is_new = 0; return r; }
  if (m->n == m->cap) { m->cap *= 2; m->v = (pair_rec*)realloc(m->v, m->cap * sizeof(pair_rec)); }
  int64_t r = m->n++;
  m->v[r].a = a; m->v[r].b = b; m->v[r].count = 0; m->v[r].next_in_list = -1;
  m->v[r].chain = m->buckets[bk]; m->buckets[bk] = r;
  if (m->n * 2 > m->n_buckets) {
    m->n_buckets *= 4; m->buckets = (int64_t*)realloc(m->buckets, m->n_buckets * 8);
    for (int64_t i = 0; i < m->n_buckets; ++i) m->buckets[i] = -1;
    for (int64_t i = 0; i < m->n; ++i) {
      int64_t bb = (int64_t)(pm_hash(m->v[i].a, m->v[i].b) & (uint64_t)(m->n_buckets - 1));
      m->v[i].chain = m->buckets[bb]; m->buckets[bb] = i;
    }
  }
  *is_new = 1;
  return r;
}

int ec_oracle_build_cells(int n_files, const int32_t* const* rg, const int32_t* const* tg, const int32_t* const* hp,
                          const int32_t* const* cell, const int64_t* n_rows, const int32_t* drop_last,
                          int64_t minimum_count, ec_cells_result* out) {
  memset(out, 0, sizeof *out);
  /* global EC table: key = sorted unique codes (as ec_oracle_build) */
  int64_t cap_rec = 1024, n_rec = 0;
  ec_rec* recs = (ec_rec*)malloc(cap_rec * sizeof(ec_rec));
  int64_t cap_arena = 4096, arena_used = 0;
  int64_t* arena = (int64_t*)malloc(cap_arena * 8);
  int64_t n_buckets = 1 << 16;
  int64_t* buckets = (int64_t*)malloc(n_buckets * 8);
  for (int64_t i = 0; i < n_buckets; ++i) buckets[i] = -1;
  int64_t cap_tmp = 1024;
  int64_t* tmp = (int64_t*)malloc(cap_tmp * 8);
  /* per EC: which file saw it last and its position in that file's key order */
  int64_t* ec_stamp = (int64_t*)malloc(cap_rec * 8);
  int64_t* ec_local = (int64_t*)malloc(cap_rec * 8);
  /* per file: keys in first-occurrence order, each with its list of (cell, count) in first-occurrence order */
  int64_t cap_loc = 1024, n_loc = 0;
  int64_t *loc_ec = (int64_t*)malloc(cap_loc * 8), *loc_head = (int64_t*)malloc(cap_loc * 8), *loc_tail = (int64_t*)malloc(cap_loc * 8);
  pair_map file_cells, final_pairs;
  pm_init(&file_cells); pm_init(&final_pairs);
  int32_t max_cell = -1;
  for (int f = 0; f < n_files; ++f)
    for (int64_t i = 0; i < n_rows[f]; ++i) {
      if (cell[f][i] < 0) return -1;
      if (cell[f][i] > max_cell) max_cell = cell[f][i];
    }
  int64_t n_cell_ids = (int64_t)max_cell + 1;
  int64_t* cell_total = (int64_t*)calloc((size_t)(n_cell_ids ? n_cell_ids : 1), 8);
  int32_t* cell_rank = (int32_t*)malloc((size_t)(n_cell_ids ? n_cell_ids : 1) * 4);   /* insertion order into cr_totals */
  for (int64_t i = 0; i < n_cell_ids; ++i) cell_rank[i] = -1;
  int32_t* cells_in_order = (int32_t*)malloc((size_t)(n_cell_ids ? n_cell_ids : 1) * 4);
  int64_t n_cells_seen = 0, n_reads = 0;

  for (int f = 0; f < n_files; ++f) {
    const int32_t *R = rg[f], *T = tg[f], *H = hp[f], *C = cell[f];
    const int64_t n = n_rows[f];
    n_loc = 0;
    pm_clear(&file_cells);
    int64_t s = 0;
    while (s < n) {   /* the worker's loop over this file */
      int64_t e = s + 1;
      while (e < n && R[e] == R[s]) ++e;
      if (drop_last[f] && e == n) break;
      int64_t k = e - s;
      if (k > cap_tmp) { cap_tmp = k * 2; tmp = (int64_t*)realloc(tmp, cap_tmp * 8); }
      for (int64_t i = 0; i < k; ++i) tmp[i] = (int64_t)T[s + i] * 64 + H[s + i];
      qsort(tmp, (size_t)k, 8, cmp_i64);
      int m = 0;
      for (int64_t i = 0; i < k; ++i)
        if (i == 0 || tmp[i] != tmp[i - 1]) tmp[m++] = tmp[i];
      uint64_t h = hash_codes(tmp, m);
      int64_t b = (int64_t)(h & (uint64_t)(n_buckets - 1));
      int64_t r = buckets[b];
      while (r >= 0) {
        if (recs[r].key_len == m && memcmp(arena + recs[r].key_off, tmp, (size_t)m * 8) == 0) break;
        r = recs[r].next;
      }
      if (r < 0) {
        if (n_rec == cap_rec) {
          cap_rec *= 2;
          recs = (ec_rec*)realloc(recs, cap_rec * sizeof(ec_rec));
          ec_stamp = (int64_t*)realloc(ec_stamp, cap_rec * 8);
          ec_local = (int64_t*)realloc(ec_local, cap_rec * 8);
        }
        if (arena_used + m > cap_arena) {
          while (arena_used + m > cap_arena) cap_arena *= 2;
          arena = (int64_t*)realloc(arena, cap_arena * 8);
        }
        memcpy(arena + arena_used, tmp, (size_t)m * 8);
        r = n_rec++;
        recs[r].key_off = arena_used; recs[r].key_len = m; recs[r].count = 0;
        recs[r].next = buckets[b]; buckets[b] = r;
        ec_stamp[r] = -1; ec_local[r] = -1;
        arena_used += m;
        if (n_rec * 2 > n_buckets) {
          n_buckets *= 4;
          buckets = (int64_t*)realloc(buckets, n_buckets * 8);
          for (int64_t i = 0; i < n_buckets; ++i) buckets[i] = -1;
          for (int64_t i = 0; i < n_rec; ++i) {
            uint64_t hh = hash_codes(arena + recs[i].key_off, recs[i].key_len);
            int64_t bb = (int64_t)(hh & (uint64_t)(n_buckets - 1));
            recs[i].next = buckets[bb]; buckets[bb] = i;
          }
        }
      }
      /* NOTE: EC ids of the merged result follow first insertion into `final`, which happens in the merge
         loop below (file order, then the file's key order) - the same order as first occurrence here. */
      if (ec_stamp[r] != f) {   /* first read with this key in this file */
        ec_stamp[r] = f;
        if (n_loc == cap_loc) {
          cap_loc *= 2;
          loc_ec = (int64_t*)realloc(loc_ec, cap_loc * 8);
          loc_head = (int64_t*)realloc(loc_head, cap_loc * 8);
          loc_tail = (int64_t*)realloc(loc_tail, cap_loc * 8);
        }
        ec_local[r] = n_loc;
        loc_ec[n_loc] = r; loc_head[n_loc] = -1; loc_tail[n_loc] = -1;
        ++n_loc;
      }
      const int64_t l = ec_local[r];
      int is_new;
      const int64_t pe = pm_get(&file_cells, l, C[s], &is_new);
      if (is_new) {   /* append to this key's cell list */
        if (loc_tail[l] >= 0) file_cells.v[loc_tail[l]].next_in_list = pe; else loc_head[l] = pe;
        loc_tail[l] = pe;
      }
      file_cells.v[pe].count += 1;
      ++n_reads;
      s = e;
    }
    /* the merge loop for this file (:513-551) */
    for (int64_t l = 0; l < n_loc; ++l)
      for (int64_t pe = loc_head[l]; pe >= 0; pe = file_cells.v[pe].next_in_list) {
        const int32_t c = file_cells.v[pe].b;
        if (cell_rank[c] < 0) { cell_rank[c] = (int32_t)n_cells_seen; cells_in_order[n_cells_seen++] = c; }
        cell_total[c] += file_cells.v[pe].count;
        int is_new;
        const int64_t gp = pm_get(&final_pairs, loc_ec[l], c, &is_new);
        final_pairs.v[gp].count += file_cells.v[pe].count;
      }
  }
  int rc = 0;
  if (n_rec == 0) { rc = -2; goto done; }   /* max() of an empty sequence (:593) */
  {
    if (minimum_count <= 0) minimum_count = 1;   /* :596-597 */
    int32_t* cell_new = (int32_t*)malloc((size_t)(n_cell_ids ? n_cell_ids : 1) * 4);
    for (int64_t i = 0; i < n_cell_ids; ++i) cell_new[i] = -1;
    int64_t n_kept = 0;
    out->cell_order = (int32_t*)malloc((size_t)(n_cells_seen ? n_cells_seen : 1) * 4);
    for (int64_t i = 0; i < n_cells_seen; ++i) {   /* :603-608 */
      const int32_t c = cells_in_order[i];
      if (cell_total[c] >= minimum_count) { cell_new[c] = (int32_t)n_kept; out->cell_order[n_kept++] = c; }
    }
    /* ECs that keep at least one cell, renumbered in order (:616-636) */
    int32_t* ec_new = (int32_t*)malloc((size_t)n_rec * 4);
    char* ec_keep = (char*)calloc((size_t)n_rec, 1);
    int64_t nnz_n = 0;
    for (int64_t p = 0; p < final_pairs.n; ++p)
      if (cell_new[final_pairs.v[p].b] >= 0) { ec_keep[final_pairs.v[p].a] = 1; ++nnz_n; }
    int64_t E = 0;
    for (int64_t i = 0; i < n_rec; ++i) ec_new[i] = ec_keep[i] ? (int32_t)E++ : -1;
    /* N as CSC: EC ids ascending inside a column -> bucket the pairs by EC first */
    int64_t* ec_ptr = (int64_t*)calloc((size_t)n_rec + 1, 8);
    for (int64_t p = 0; p < final_pairs.n; ++p) ec_ptr[final_pairs.v[p].a + 1] += 1;
    for (int64_t i = 0; i < n_rec; ++i) ec_ptr[i + 1] += ec_ptr[i];
    int64_t* by_ec = (int64_t*)malloc((size_t)(final_pairs.n ? final_pairs.n : 1) * 8);
    int64_t* fill = (int64_t*)malloc((size_t)n_rec * 8);
    memcpy(fill, ec_ptr, (size_t)n_rec * 8);
    for (int64_t p = 0; p < final_pairs.n; ++p) by_ec[fill[final_pairs.v[p].a]++] = p;
    out->n_indptr = (int32_t*)calloc((size_t)n_kept + 1, 4);
    out->n_indices = (int32_t*)malloc((size_t)(nnz_n ? nnz_n : 1) * 4);
    out->n_data = (int32_t*)malloc((size_t)(nnz_n ? nnz_n : 1) * 4);
    for (int64_t p = 0; p < final_pairs.n; ++p) {
      const int32_t cn = cell_new[final_pairs.v[p].b];
      if (cn >= 0) out->n_indptr[cn + 1] += 1;
    }
    for (int64_t i = 0; i < n_kept; ++i) out->n_indptr[i + 1] += out->n_indptr[i];
    int32_t* col_fill = (int32_t*)malloc((size_t)(n_kept ? n_kept : 1) * 4);
    memcpy(col_fill, out->n_indptr, (size_t)n_kept * 4);
    for (int64_t i = 0; i < n_rec; ++i)
      for (int64_t q = ec_ptr[i]; q < ec_ptr[i + 1]; ++q) {
        const pair_rec* pr = &final_pairs.v[by_ec[q]];
        const int32_t cn = cell_new[pr->b];
        if (cn < 0) continue;
        out->n_indices[col_fill[cn]] = ec_new[i];
        out->n_data[col_fill[cn]] = pr->count;
        col_fill[cn] += 1;
      }
    /* A rows of the kept ECs */
    int64_t nnz = 0;
    for (int64_t i = 0; i < n_rec; ++i) {
      if (!ec_keep[i]) continue;
      const int64_t* k = arena + recs[i].key_off;
      for (int j = 0; j < recs[i].key_len; ++j)
        if (j == 0 || (k[j] >> 6) != (k[j - 1] >> 6)) ++nnz;
    }
    out->a_indptr = (int32_t*)malloc((size_t)(E + 1) * 4);
    out->a_indices = (int32_t*)malloc((size_t)(nnz ? nnz : 1) * 4);
    out->a_data = (int32_t*)malloc((size_t)(nnz ? nnz : 1) * 4);
    int64_t z = 0, row = 0;
    out->a_indptr[0] = 0;
    for (int64_t i = 0; i < n_rec; ++i) {
      if (!ec_keep[i]) continue;
      const int64_t* k = arena + recs[i].key_off;
      for (int j = 0; j < recs[i].key_len; ++j) {
        if (j == 0 || (k[j] >> 6) != (k[j - 1] >> 6)) { out->a_indices[z] = (int32_t)(k[j] >> 6); out->a_data[z] = 0; ++z; }
        out->a_data[z - 1] |= (int32_t)(1u << (k[j] & 63));
      }
      out->a_indptr[++row] = (int32_t)z;
    }
    out->n_ec = E; out->nnz_a = nnz; out->n_cells = n_kept; out->nnz_n = nnz_n; out->n_reads = n_reads;
    free(cell_new); free(ec_new); free(ec_keep); free(ec_ptr); free(by_ec); free(fill); free(col_fill);
  }
done:
  free(recs); free(arena); free(buckets); free(tmp); free(ec_stamp); free(ec_local);
  free(loc_ec); free(loc_head); free(loc_tail); free(file_cells.v); free(file_cells.buckets);
  free(final_pairs.v); free(final_pairs.buckets); free(cell_total); free(cell_rank); free(cells_in_order);
  return rc;
}
