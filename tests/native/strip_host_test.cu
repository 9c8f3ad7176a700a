// CPU check of the per-lane walk of the strip kernel (alntools_b200/csrc/ecb_strip.cuh): every tile and
// lane is emulated with the very functions the kernel calls, and the reads it closes (or hands to the
// long-read routine) are compared with a serial statement of the grouping rule of
// alntools/bam_utils.py:301-344 on columns: consecutive equal read_group values form a read, an element
// (target, haplotype) counts once per read.  No CUDA call is made; the file is built with nvcc only
// because the header is CUDA source.  usage: strip_host_test [cases]
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>
#include <vector>
#include "../../include/ecb200.h"
#include "../../alntools_b200/csrc/ecb_strip.cuh"

struct Closed {
  int len;
  Key128 key;
  int times;
};

static Key128 serial_key(const std::vector<int>& tg, const std::vector<int>& hp, int s, int e) {
  Mix4 sum = mix_zero();
  for (int i = s; i < e; ++i) {
    bool dup = false;
    for (int j = s; j < i; ++j) dup |= tg[j] == tg[i] && hp[j] == hp[i];
    if (!dup) mix_add(sum, ecb_mix(ecb_code(tg[i], hp[i])));
  }
  return mix_to_key(sum);
}

static int run_case(unsigned seed, int n, int mode, bool verbose) {
  std::mt19937 rng(seed);
  std::vector<int> rg(n), tg(n), hp(n);
  int g = (int)(rng() % 1000) - 500, i = 0;
  const int n_t = 1 + (int)(rng() % (mode == 2 ? 3 : 50));
  while (i < n) {
    int len;
    switch (mode) {
      case 0: len = 1 + (int)(rng() % 3); break;                          // short reads
      case 1: len = 1 + (int)(rng() % 12); break;                         // around the 8-alignment limit
      case 2: len = 1 + (int)(rng() % 9); break;                          // many duplicates (few targets)
      case 3: len = (rng() % 4) ? 1 + (int)(rng() % 4) : 5 + (int)(rng() % 80); break;  // long reads in between
      default: len = 8; break;                                            // exactly 8 everywhere
    }
    for (int k = 0; k < len && i < n; ++k, ++i) {
      rg[i] = g;
      tg[i] = (int)(rng() % n_t);
      hp[i] = (int)(rng() % 4);
    }
    g += 1 + (int)(rng() % 3);
  }
  std::map<int, Closed> got;  // start -> what the walk said
  int errors = 0;
  for (int tb = 0; tb < n; tb += ECB_TILE)
    for (int lane = 0; lane < 32; ++lane) {
      const int p0 = tb + ECB_STRIP * lane;
      int a[ECB_STRIP_SPAN], b[ECB_STRIP_SPAN], c[ECB_STRIP_SPAN];
      for (int k = 0; k < ECB_STRIP_SPAN; ++k) {
        const bool in = p0 + k < n;
        a[k] = in ? rg[p0 + k] : ECB_RG_SENTINEL;
        b[k] = in ? tg[p0 + k] : 0;
        c[k] = in ? hp[p0 + k] : 0;
      }
      const int rgprev = (p0 > 0 && p0 <= n) ? rg[p0 - 1] : ECB_RG_SENTINEL;
      StripLane L;
      strip_lane_build(L, rgprev, a, b, c);
      StripWalk W;
      strip_walk_init(W);
      for (int k = 0; k < ECB_STRIP_SPAN; ++k) {
        if (strip_walk_closes(L, k, W)) {
          Closed& e = got[p0 + W.st];
          e.len = k - W.st;
          e.key = mix_to_key(W.sum);
          e.times++;
        }
        strip_walk_advance(L, k, n - p0, W);
      }
      if (W.open) {
        if (verbose) printf("lane left a read open: tile %d lane %d\n", tb, lane);
        ++errors;
      }
      if (W.has_long) {
        const int s = p0 + (int)W.long_start;
        int e = s;
        while (e < n && rg[e] == rg[s]) ++e;
        Closed& r = got[s];
        r.len = e - s;
        r.key = serial_key(tg, hp, s, e);
        r.times++;
        if (e - s <= ECB_STRIP) {
          if (verbose) printf("short read %d (len %d) was handed to the long path\n", s, e - s);
          ++errors;
        }
      }
    }
  int reads = 0;
  for (int s = 0; s < n;) {
    int e = s;
    while (e < n && rg[e] == rg[s]) ++e;
    ++reads;
    auto it = got.find(s);
    if (it == got.end()) {
      if (verbose) printf("read at %d (len %d) was never closed\n", s, e - s);
      ++errors;
    } else {
      const Key128 want = serial_key(tg, hp, s, e);
      if (it->second.times != 1 || it->second.len != e - s || !key_eq(it->second.key, want)) {
        if (verbose) printf("read at %d: times %d len %d (want %d) key %s\n", s, it->second.times, it->second.len, e - s,
                            key_eq(it->second.key, want) ? "ok" : "DIFFERS");
        ++errors;
      }
    }
    s = e;
  }
  if ((int)got.size() != reads) {
    if (verbose) printf("%zu reads closed, %d expected\n", got.size(), reads);
    ++errors;
  }
  return errors;
}

int main(int argc, char** argv) {
  const int cases = argc > 1 ? atoi(argv[1]) : 400;
  int bad = 0;
  std::mt19937 rng(12345);
  for (int k = 0; k < cases; ++k) {
    const int mode = k % 5;
    int n;
    switch (k % 7) {
      case 0: n = 1 + (int)(rng() % 20); break;
      case 1: n = ECB_TILE - 9 + (int)(rng() % 20); break;
      case 2: n = 2 * ECB_TILE - 9 + (int)(rng() % 20); break;
      default: n = 1 + (int)(rng() % 5000); break;
    }
    const int e = run_case(1000u + (unsigned)k, n, mode, bad < 5);
    if (e) {
      printf("case %d (n %d, mode %d): %d errors\n", k, n, mode, e);
      ++bad;
    }
  }
  printf("%d cases, %d failed\n", cases, bad);
  return bad ? 1 : 0;
}
