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
