// fast_inflate.h — whole-buffer DEFLATE (RFC 1951) decoder for BGZF blocks.
//
// A BGZF block is a complete raw-deflate stream of at most 64 KiB whose inflated size is known up front
// (ISIZE), source and destination are both in memory.  That removes everything zlib's streaming
// inflate pays for: no state machine, no window copy, matches are copied inside the destination, the
// bit buffer is 64 bits wide and refilled with one unaligned load, literal/length symbols come from one
// 11-bit table look-up (longer codes through a second-level table).
//
// Contract: fast_inflate() returns true only if the stream was well formed, ended exactly at a final
// block and produced exactly dst_len bytes.  On ANYTHING else - invalid code set, bad distance, output or
// input overrun - it returns false and the caller runs zlib on the same block, so error behaviour is
// zlib's.  It never reads outside [src, src + src_len) nor writes outside [dst, dst + dst_len).
#pragma once
#include <cstdint>
#include <cstring>

namespace fastinflate {

// Table entry, 32 bits:  [31:16] value (literal byte, length / distance base, or sub-table offset)
//                        [15:12] kind     [11:8] extra bits     [7:0] code bits to consume
enum : uint32_t { K_INVALID = 0, K_BASE = 1, K_EOB = 2, K_SUB = 4, K_LITERAL = 8 };   // one bit each
constexpr uint32_t F_LITERAL = K_LITERAL << 12, F_SUB = K_SUB << 12, F_BASE = K_BASE << 12;
static inline uint32_t make_entry(uint32_t value, uint32_t kind, uint32_t extra, uint32_t bits) {
  return (value << 16) | (kind << 12) | (extra << 8) | bits;
}

constexpr int LIT_BITS = 11;   // primary bits of the literal/length table
constexpr int DIST_BITS = 8;   // primary bits of the distance table
constexpr int LIT_TABLE = (1 << LIT_BITS) + 1024;   // primary + room for every sub-table (15-bit codes)
constexpr int DIST_TABLE = (1 << DIST_BITS) + 512;

struct Tables {
  uint32_t lit[LIT_TABLE];
  uint32_t dist[DIST_TABLE];
};

static const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59,
                                      67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769,
                                       1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8,
                                       9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

static inline uint32_t reverse_bits(uint32_t code, int len) {
  uint32_t r = 0;
  for (int i = 0; i < len; ++i) r |= ((code >> i) & 1u) << (len - 1 - i);
  return r;
}

// What a symbol of alphabet `kind_of` decodes to (without its code length).
static inline uint32_t litlen_symbol(int sym) {
  if (sym < 256) return make_entry((uint32_t)sym, K_LITERAL, 0, 0);
  if (sym == 256) return make_entry(0, K_EOB, 0, 0);
  if (sym > 285) return 0;   // 286, 287 never appear in a valid stream
  return make_entry(LEN_BASE[sym - 257], K_BASE, LEN_EXTRA[sym - 257], 0);
}
static inline uint32_t dist_symbol(int sym) {
  if (sym > 29) return 0;
  return make_entry(DIST_BASE[sym], K_BASE, DIST_EXTRA[sym], 0);
}

// Canonical Huffman decode table from code lengths (RFC 1951 3.2.2), LSB-first look-up.  Returns false
// for an over-subscribed or (except for the one-code case the format allows) incomplete code set.
template <class SymbolFn>
static bool build_table(const uint8_t* lens, int n_sym, int primary_bits, uint32_t* table, int table_cap,
                        SymbolFn symbol_of, bool allow_incomplete) {
  int count[16] = {0};
  for (int s = 0; s < n_sym; ++s) count[lens[s]]++;
  count[0] = 0;
  int used = 0;
  long left = 1;   // Kraft: codes still available
  for (int l = 1; l <= 15; ++l) {
    left <<= 1;
    left -= count[l];
    if (left < 0) return false;   // over-subscribed
    used += count[l];
  }
  if (left > 0) {   // incomplete
    if (!(allow_incomplete && used <= 1)) return false;
  }
  uint32_t next_code[16];
  uint32_t code = 0;
  for (int l = 1; l <= 15; ++l) {
    code = (code + (uint32_t)count[l - 1]) << 1;
    next_code[l] = code;
  }
  const int primary = 1 << primary_bits;
  for (int i = 0; i < primary; ++i) table[i] = 0;
  // sub-table sizes: longest code behind every primary prefix
  uint8_t sub_bits[1 << LIT_BITS];
  bool any_long = false;
  for (int l = primary_bits + 1; l <= 15; ++l) any_long |= count[l] != 0;
  if (any_long) memset(sub_bits, 0, (size_t)primary);
  uint32_t codes[288];
  for (int s = 0; s < n_sym; ++s) {
    const int l = lens[s];
    if (!l) continue;
    codes[s] = reverse_bits(next_code[l]++, l);
    if (l > primary_bits) {
      uint8_t& sb = sub_bits[codes[s] & (uint32_t)(primary - 1)];
      if (l - primary_bits > sb) sb = (uint8_t)(l - primary_bits);
    }
  }
  int next_sub = primary;
  if (any_long)
    for (int p = 0; p < primary; ++p)
      if (sub_bits[p]) {
        const int size = 1 << sub_bits[p];
        if (next_sub + size > table_cap) return false;
        table[p] = make_entry((uint32_t)next_sub, K_SUB, sub_bits[p], (uint32_t)primary_bits);
        for (int i = 0; i < size; ++i) table[next_sub + i] = 0;
        next_sub += size;
      }
  for (int s = 0; s < n_sym; ++s) {
    const int l = lens[s];
    if (!l) continue;
    const uint32_t e = symbol_of(s);
    if (!e) return false;   // a code for a symbol that must not occur
    if (l <= primary_bits) {
      const uint32_t ent = e | (uint32_t)l;
      for (uint32_t i = codes[s]; i < (uint32_t)primary; i += 1u << l) table[i] = ent;
    } else {
      const uint32_t link = table[codes[s] & (uint32_t)(primary - 1)];
      const uint32_t off = link >> 16, sb = (link >> 8) & 15u;
      const int sl = l - primary_bits;
      const uint32_t ent = e | (uint32_t)sl;
      for (uint32_t i = codes[s] >> primary_bits; i < (1u << sb); i += 1u << sl) table[off + i] = ent;
    }
  }
  return true;
}

static inline uint64_t load64(const uint8_t* p) {
  uint64_t v;
  memcpy(&v, p, 8);
  return v;   // little-endian hosts only (x86-64 / aarch64)
}

struct BitReader {
  const uint8_t* in;
  const uint8_t* in_end;
  uint64_t buf = 0;
  int cnt = 0;       // valid bits in buf; negative = bits were consumed that the input never had
  // at least 56 valid bits afterwards, or everything that is left (zeros are shifted in beyond the end)
  inline void refill() {
    if (cnt < 0) return;   // already past the end of the input: the caller will notice
    if (in + 8 <= in_end) {
      buf |= load64(in) << cnt;
      in += (63 - cnt) >> 3;
      cnt |= 56;
    } else {
      while (cnt <= 56 && in < in_end) {
        buf |= (uint64_t)*in++ << cnt;
        cnt += 8;
      }
    }
  }
  inline uint32_t peek(int n) const { return (uint32_t)(buf & ((1ull << n) - 1)); }
  inline void drop(int n) {
    buf >>= n;
    cnt -= n;   // may go negative: checked by the callers at their checkpoints (bad())
  }
  inline bool bad() const { return cnt < 0; }
  inline uint32_t take(int n) {
    const uint32_t v = peek(n);
    drop(n);
    return v;
  }
};

// One symbol from a two-level table; the code bits are dropped, the entry (with its extra-bit count) is returned.
static inline uint32_t decode_symbol(BitReader& br, const uint32_t* table, int primary_bits) {
  uint32_t e = table[br.peek(primary_bits)];
  if (((e >> 12) & 15u) == K_SUB) {
    br.drop(primary_bits);
    e = table[(e >> 16) + br.peek((int)((e >> 8) & 15u))];
  }
  br.drop((int)(e & 0xFFu));
  return e;
}

static bool read_dynamic_tables(BitReader& br, Tables& T) {
  static const uint8_t ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  br.refill();
  const int hlit = (int)br.take(5) + 257, hdist = (int)br.take(5) + 1, hclen = (int)br.take(4) + 4;
  if (hlit > 286 || hdist > 30) return false;
  uint8_t cl[19] = {0};
  for (int i = 0; i < hclen; ++i) {
    if ((i & 7) == 0) br.refill();
    cl[ORDER[i]] = (uint8_t)br.take(3);
  }
  // code-length code: at most 7 bits, one flat table
  uint32_t cl_table[128];
  {
    int count[8] = {0};
    for (int s = 0; s < 19; ++s) count[cl[s]]++;
    count[0] = 0;
    long left = 1;
    for (int l = 1; l <= 7; ++l) {
      left = (left << 1) - count[l];
      if (left < 0) return false;
    }
    if (left > 0) return false;   // zlib rejects an incomplete code-length code too
    uint32_t next_code[8], code = 0;
    for (int l = 1; l <= 7; ++l) {
      code = (code + (uint32_t)count[l - 1]) << 1;
      next_code[l] = code;
    }
    for (int i = 0; i < 128; ++i) cl_table[i] = 0;
    for (int s = 0; s < 19; ++s) {
      const int l = cl[s];
      if (!l) continue;
      const uint32_t rev = reverse_bits(next_code[l]++, l);
      for (uint32_t i = rev; i < 128; i += 1u << l) cl_table[i] = ((uint32_t)s << 8) | (uint32_t)l | 0x10000u;
    }
  }
  uint8_t lens[286 + 30 + 16];
  int n = 0;
  const int total = hlit + hdist;
  while (n < total) {
    br.refill();
    const uint32_t e = cl_table[br.peek(7)];
    if (!e) return false;
    br.drop((int)(e & 0xFFu));
    const int sym = (int)((e >> 8) & 0xFFu);
    if (sym < 16) {
      lens[n++] = (uint8_t)sym;
    } else {
      int rep;
      uint8_t val = 0;
      if (sym == 16) {
        if (n == 0) return false;
        val = lens[n - 1];
        rep = 3 + (int)br.take(2);
      } else if (sym == 17) {
        rep = 3 + (int)br.take(3);
      } else {
        rep = 11 + (int)br.take(7);
      }
      if (n + rep > total) return false;
      while (rep--) lens[n++] = val;
    }
    if (br.bad()) return false;
  }
  if (lens[256] == 0) return false;   // no end-of-block code
  if (!build_table(lens, hlit, LIT_BITS, T.lit, LIT_TABLE, litlen_symbol, false)) {
    // zlib accepts an incomplete literal/length code only when it has a single code; BGZF writers never
    // produce one - leave those to zlib
    return false;
  }
  if (!build_table(lens + hlit, hdist, DIST_BITS, T.dist, DIST_TABLE, dist_symbol, true)) return false;
  return true;
}

static void build_fixed_tables(Tables& T) {
  uint8_t lens[288];
  for (int s = 0; s < 144; ++s) lens[s] = 8;
  for (int s = 144; s < 256; ++s) lens[s] = 9;
  for (int s = 256; s < 280; ++s) lens[s] = 7;
  for (int s = 280; s < 288; ++s) lens[s] = 8;
  // symbols 286/287 have codes in the fixed set but must not occur: give them no entry
  auto sym = [](int s) -> uint32_t { return s > 285 ? make_entry(0, K_INVALID, 0, 0) | 0x80000000u : litlen_symbol(s); };
  build_table(lens, 288, LIT_BITS, T.lit, LIT_TABLE, sym, false);
  for (int i = 0; i < LIT_TABLE; ++i)
    if (T.lit[i] & 0x80000000u) T.lit[i] = 0;
  uint8_t dl[32];
  for (int s = 0; s < 32; ++s) dl[s] = 5;
  auto dsym = [](int s) -> uint32_t { return s > 29 ? 0x80000000u : dist_symbol(s); };
  build_table(dl, 32, DIST_BITS, T.dist, DIST_TABLE, dsym, false);
  for (int i = 0; i < DIST_TABLE; ++i)
    if (T.dist[i] & 0x80000000u) T.dist[i] = 0;
}

// Match copy inside the destination; the caller guarantees at least length + 16 writable bytes at out.
static inline void copy_match_fast(uint8_t* out, uint32_t distance, uint32_t length) {
  const uint8_t* from = out - distance;
  uint8_t* const stop = out + length;
  if (distance >= 16) {
    do {
      memcpy(out, from, 16);
      out += 16;
      from += 16;
    } while (out < stop);
  } else if (distance >= 8) {
    do {
      memcpy(out, from, 8);
      out += 8;
      from += 8;
    } while (out < stop);
  } else if (distance == 1) {
    memset(out, *from, length);
  } else {
    // 2..7: widen the period to at least 8 bytes, then go on in 8-byte steps
    uint8_t* o = out;
    for (uint32_t i = 0; i < 8 && o < stop; ++i) *o++ = *from++;
    if (o < stop) {
      const uint32_t period = distance * (8 / distance + (8 % distance ? 1 : 0));   // multiple of distance, >= 8
      // bytes [out, out + period) are final as soon as `period` bytes are written: finish them byte-wise
      while (o < out + period && o < stop) *o++ = *from++;
      const uint8_t* f = o - period;
      while (o < stop) {
        memcpy(o, f, 8);
        o += 8;
        f += 8;
      }
    }
  }
}

// Decode one Huffman-coded block body.  Returns false on any malformation / overrun.
// (A BMI2 build of this function - shrx/bzhi for the variable shifts - measured 4 % faster and a
// target_clones dispatch 20 % slower than this plain build, so there is one build.)
static bool inflate_codes(BitReader& br, const Tables& T, uint8_t* const dst, uint8_t*& out_ref, uint8_t* const out_end) {
  uint8_t* out = out_ref;
  // ---- fast loop: far enough from both ends that no single step can leave the buffers -------------------
  // one step writes at most 3 literals + a 258-byte match (+ 15 bytes of copy overshoot) and refills up to
  // three times (8 bytes read at br.in, which advances by at most 7 each time)
  constexpr long OUT_MARGIN = 3 + 258 + 16, IN_MARGIN = 32;
  if (out_end - out > OUT_MARGIN && br.in_end - br.in > IN_MARGIN && br.cnt >= 0) {
    uint8_t* const out_safe = out_end - OUT_MARGIN;
    const uint8_t* const in_safe = br.in_end - IN_MARGIN;
    uint64_t buf = br.buf;
    int cnt = br.cnt;
    const uint8_t* in = br.in;
#define FI_REFILL()               \
  buf |= load64(in) << cnt;       \
  in += (63 - cnt) >> 3;          \
  cnt |= 56;
    bool ok = true, done = false;
#define FI_PRIMARY(e) e = T.lit[buf & ((1u << LIT_BITS) - 1)];
#define FI_CONSUME(e)                                                  \
  if (e & F_SUB) {                                                     \
    buf >>= LIT_BITS;                                                  \
    cnt -= LIT_BITS;                                                   \
    e = T.lit[(e >> 16) + (uint32_t)(buf & ((1ull << ((e >> 8) & 15u)) - 1))]; \
  }                                                                    \
  buf >>= (e & 0xFFu);                                                 \
  cnt -= (int)(e & 0xFFu);
    // e = primary table entry of the upcoming symbol, looked up right after a refill, nothing consumed yet;
    // it is fetched BEFORE the match copy of the previous symbol so that the table load hides behind the copy
    uint32_t e;
    FI_REFILL()
    FI_PRIMARY(e)
    while (out < out_safe && in < in_safe) {
      FI_CONSUME(e)                       // >= 41 bits left
      if (e & F_LITERAL) {
        *out++ = (uint8_t)(e >> 16);
        FI_PRIMARY(e)
        FI_CONSUME(e)                     // >= 26
        if (e & F_LITERAL) {
          *out++ = (uint8_t)(e >> 16);
          FI_PRIMARY(e)
          FI_CONSUME(e)                   // >= 11
          if (e & F_LITERAL) {
            *out++ = (uint8_t)(e >> 16);
            FI_REFILL()
            FI_PRIMARY(e)
            continue;
          }
        }
        FI_REFILL()                       // a length / distance pair needs up to 33 bits
      }
      if (!(e & F_BASE)) {
        if (((e >> 12) & 15u) == K_EOB) done = true; else ok = false;
        break;
      }
      const uint32_t lx = (e >> 8) & 15u;
      const uint32_t length = (e >> 16) + (uint32_t)(buf & ((1ull << lx) - 1));
      buf >>= lx;
      cnt -= (int)lx;
      uint32_t d = T.dist[buf & ((1u << DIST_BITS) - 1)];
      if (d & F_SUB) {
        buf >>= DIST_BITS;
        cnt -= DIST_BITS;
        d = T.dist[(d >> 16) + (uint32_t)(buf & ((1ull << ((d >> 8) & 15u)) - 1))];
      }
      buf >>= (d & 0xFFu);
      cnt -= (int)(d & 0xFFu);
      if (!(d & F_BASE)) {
        ok = false;
        break;
      }
      const uint32_t dx = (d >> 8) & 15u;
      const uint32_t distance = (d >> 16) + (uint32_t)(buf & ((1ull << dx) - 1));
      buf >>= dx;
      cnt -= (int)dx;
      if (distance > (uint32_t)(out - dst)) {
        ok = false;
        break;
      }
      FI_REFILL()
      FI_PRIMARY(e)                       // next symbol's entry is on its way while the match is copied
      copy_match_fast(out, distance, length);
      out += length;
    }
#undef FI_PRIMARY
#undef FI_CONSUME
#undef FI_REFILL
    br.buf = buf;
    br.cnt = cnt;
    br.in = in;
    out_ref = out;
    if (!ok) return false;
    if (done) return true;
  }
  // ---- careful loop: every step checked against both ends ------------------------------------------------
  for (;;) {
    if (br.cnt < 0) return false;
    br.refill();
    uint32_t e = decode_symbol(br, T.lit, LIT_BITS);
    uint32_t kind = (e >> 12) & 15u;
    if (kind == K_LITERAL) {
      if (out >= out_end) return false;
      *out++ = (uint8_t)(e >> 16);
      out_ref = out;
      continue;
    }
    if (br.cnt < 0) return false;
    if (kind == K_EOB) {
      out_ref = out;
      return true;
    }
    if (kind != K_BASE) return false;
    br.refill();
    const uint32_t length = (e >> 16) + br.take((int)((e >> 8) & 15u));
    const uint32_t d = decode_symbol(br, T.dist, DIST_BITS);
    if (((d >> 12) & 15u) != K_BASE) return false;
    const uint32_t distance = (d >> 16) + br.take((int)((d >> 8) & 15u));
    if (br.cnt < 0) return false;
    if (distance > (uint32_t)(out - dst) || length > (uint32_t)(out_end - out)) return false;
    const uint8_t* from = out - distance;
    for (uint32_t i = 0; i < length; ++i) out[i] = from[i];
    out += length;
    out_ref = out;
  }
}

static bool fast_inflate(const uint8_t* src, size_t src_len, uint8_t* dst, size_t dst_len, Tables& scratch,
                         Tables& fixed, bool& fixed_ready) {
  BitReader br;
  br.in = src;
  br.in_end = src + src_len;
  uint8_t* out = dst;
  uint8_t* const out_end = dst + dst_len;
  for (;;) {
    br.refill();
    const uint32_t final_block = br.take(1), type = br.take(2);
    if (br.bad()) return false;
    if (type == 0) {
      // stored: skip to the byte boundary, LEN / NLEN, raw bytes
      br.drop(br.cnt & 7);
      br.refill();
      if (br.cnt < 32) return false;
      const uint32_t len = br.take(16), nlen = br.take(16);
      if ((len ^ 0xFFFFu) != nlen) return false;
      // hand the unread whole bytes of the bit buffer back to the input
      const uint8_t* p = br.in - (br.cnt >> 3);
      br.buf = 0;
      br.cnt = 0;
      if ((size_t)(br.in_end - p) < len || (size_t)(out_end - out) < len) return false;
      memcpy(out, p, len);
      out += len;
      br.in = p + len;
    } else if (type == 1) {
      if (!fixed_ready) {
        build_fixed_tables(fixed);
        fixed_ready = true;
      }
      if (!inflate_codes(br, fixed, dst, out, out_end)) return false;
    } else if (type == 2) {
      if (!read_dynamic_tables(br, scratch)) return false;
      if (!inflate_codes(br, scratch, dst, out, out_end)) return false;
    } else {
      return false;
    }
    if (final_block) break;
  }
  return out == out_end && !br.bad();
}

}  // namespace fastinflate
