/*
 * pz_device.cuh -- the device-side inflate engine (warp-per-stream).
 *
 * One warp decodes one zlib stream: all 32 lanes keep the same 64-bit bit buffer
 * (warp-uniform control flow, broadcast shared-memory lookups), lane 0 stores literals and
 * LZ77 matches are copied cooperatively (lane i moves byte i), straight into the stream's
 * slice of the HBM output blob -- the output itself is the history window.
 *
 * What it replaces in the reference (file:line relative to the pure-zlib checkout):
 *   bit reader            Monad.hs:199-263   -> PzCtx bit buffer over a cp.async-staged smem ring
 *   tree build + walk     HuffmanTree.hs:25-83, Deflate.hs:255-288 -> canonical counts + flat LUT
 *   block parser          Deflate.hs:65-156  -> pz_inflate_stream()
 *   symbol loop           Deflate.hs:106-120 -> fast LUT loop + exact bit-serial "careful" path
 *   output window         OutputWindow.hs:29-114 -> direct stores; the window is only *modelled*
 *                                              (fill/base counters) to reproduce its verdicts
 *   zlib framing          Zlib.hs:53-69, Deflate.hs:52-63
 *
 * The file also compiles with a host C++ compiler when PZ_HOSTSIM is defined: a "warp" is
 * then a single lane.  That build exists only for tests/hostsim (CPU-side differential
 * fuzzing of this logic against the oracle); the product library never contains it.
 */
#pragma once
#include <stdint.h>

#include "pzcuda.h"

#ifdef PZ_HOSTSIM
#include <string.h>
#define PZ_DEV static inline
#define PZ_WARP 1
PZ_DEV int pz_lane() { return 0; }
PZ_DEV void pz_syncwarp() {}
PZ_DEV unsigned pz_ballot(int p) { return p ? 1u : 0u; }
PZ_DEV unsigned pz_match_any(unsigned) { return 1u; }
PZ_DEV unsigned pz_lanemask_lt() { return 0u; }
PZ_DEV int pz_shfl(int v, int) { return v; }
PZ_DEV void pz_smem_inc(uint32_t *p) { ++*p; }
PZ_DEV int pz_popc(unsigned x) { return __builtin_popcount(x); }
PZ_DEV int pz_ffs(unsigned x) { return __builtin_ffs((int)x); }
PZ_DEV void pz_copy16_async(void *smem_dst, const void *gsrc) { memcpy(smem_dst, gsrc, 16); }
PZ_DEV void pz_async_wait_all() {}
#else
#define PZ_DEV __device__ __forceinline__
#define PZ_WARP 32
PZ_DEV int pz_lane() { return (int)(threadIdx.x & 31u); }
PZ_DEV void pz_syncwarp() { __syncwarp(); }
PZ_DEV unsigned pz_ballot(int p) { return __ballot_sync(0xffffffffu, p); }
PZ_DEV unsigned pz_match_any(unsigned v) { return __match_any_sync(0xffffffffu, v); }
PZ_DEV unsigned pz_lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}
PZ_DEV int pz_shfl(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
PZ_DEV void pz_smem_inc(uint32_t *p) { atomicAdd(p, 1u); }
PZ_DEV int pz_popc(unsigned x) { return __popc(x); }
PZ_DEV int pz_ffs(unsigned x) { return __ffs((int)x); }
PZ_DEV void pz_copy16_async(void *smem_dst, const void *gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
PZ_DEV void pz_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
#endif

/* ---- geometry ------------------------------------------------------------------------ */
#define PZ_LIT_BITS 10  /* first-level bits of the literal/length LUT */
#define PZ_DIST_BITS 8  /* first-level bits of the distance LUT        */
#define PZ_PRE_BITS 7   /* the code-length code never exceeds 7 bits    */
#define PZ_RING_WORDS 256u /* staged input: two 512-byte halves          */
#define PZ_HALF_WORDS 128u
#define PZ_MAX_LENS 464 /* 288 + 32 + 137 overshoot (Deflate.hs:124-156), padded */
#define PZ_WINDOW 131072u /* OutputWindow.hs:29-30 */
#define PZ_EXCESS 32768u  /* OutputWindow.hs:42-43 */

/* LUT entry: total bits [0,5) | code bits [8,12) | type [12,14) | value [16,32) */
#define PZ_T_LIT 0u
#define PZ_T_BASE 1u /* length / distance base + extra bits */
#define PZ_T_EOB 2u
#define PZ_T_SLOW 3u /* long code, dead prefix or a symbol the reference cannot index */
#define PZ_ENTRY(total, nbits, type, value) ((uint32_t)(total) | ((uint32_t)(nbits) << 8) | ((uint32_t)(type) << 12) | ((uint32_t)(value) << 16))
#define PZ_SLOW_ENTRY PZ_ENTRY(0, 0, PZ_T_SLOW, 0)

/* Canonical description of one prefix code: enough for the bit-serial walker to reproduce
 * the reference trie's accept / "Advanced to empty tree!" behaviour (HuffmanTree.hs:73-83). */
struct PzTree {
  uint16_t cnt[16];  /* cnt[l]  = codes of length l                                     */
  uint16_t used[16]; /* used[l] = l-bit prefixes that lead to longer codes              */
  uint16_t nsyms;
  uint16_t pad;
};

struct __attribute__((aligned(16))) PzWarpSmem {
  uint32_t lit_lut[1 << PZ_LIT_BITS];
  uint32_t dist_lut[1 << PZ_DIST_BITS]; /* the precode LUT aliases its first 128 entries */
  uint32_t ring[PZ_RING_WORDS];
  uint32_t scratch[32]; /* [0,16) per-length counters, [16,32) per-length offsets */
  uint16_t lit_perm[288];
  uint16_t dist_perm[176];
  uint16_t pre_perm[24];
  PzTree lit, dist, pre;
  uint8_t lens[PZ_MAX_LENS];
};

#ifdef PZ_HOSTSIM
static const uint16_t PZ_LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t PZ_LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t PZ_DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t PZ_DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static const uint8_t PZ_CL_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
#else
/* Deflate.hs:160-237 (length / distance tables) and Deflate.hs:290-292 (codeLengthOrder) */
static __constant__ uint16_t PZ_LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static __constant__ uint8_t PZ_LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static __constant__ uint16_t PZ_DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static __constant__ uint8_t PZ_DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static __constant__ uint8_t PZ_CL_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
#endif

/* ---- per-stream decoder state (registers; identical in every lane) -------------------- */
struct PzCtx {
  const uint8_t *in_al; /* input, rounded down to 16 bytes                                */
  uint64_t in_al_bytes; /* bytes readable from in_al (multiple of 16)                      */
  uint64_t end_bit;     /* first bit past the stream, counted from in_al                   */
  uint32_t *ring;
  uint64_t bitbuf;
  uint32_t cnt;      /* valid bits in bitbuf                                               */
  uint32_t word_idx; /* next 32-bit word (from in_al) to append to bitbuf                  */
  int32_t safe_word; /* word_idx <= safe_word: >= 48 input bits remain, no checks needed   */
  uint8_t *out;
  uint32_t pos;  /* bytes decoded                                                          */
  uint32_t base; /* bytes the reference would already have published (multiple of 32 KiB)  */
  uint32_t cap;
  int32_t status, detail;
  int64_t p0, p1;
};

PZ_DEV void pz_fail(PzCtx &c, int status, int detail, int64_t p0 = 0, int64_t p1 = 0) {
  c.status = status; c.detail = detail; c.p0 = p0; c.p1 = p1;
}

/* ---- staged input ---------------------------------------------------------------------
 * Block k = bytes [512k, 512k+512) of in_al, staged into ring half (k & 1) with cp.async
 * (16 bytes per lane).  Invariant: the block holding word_idx and its successor are loaded
 * or in flight. */
PZ_DEV void pz_ring_issue(PzCtx &c, uint32_t k) {
  uint32_t *dst = c.ring + (k & 1u) * PZ_HALF_WORDS;
  for (uint32_t p = (uint32_t)pz_lane(); p < 32u; p += PZ_WARP) {
    uint64_t bo = (uint64_t)k * 512u + p * 16u;
    if (bo < c.in_al_bytes) pz_copy16_async(dst + p * 4u, c.in_al + bo);
    /* past the stream: never consumed (the careful path counts bits), left as is */
  }
}
PZ_DEV void pz_ring_cross(PzCtx &c) {
  pz_async_wait_all();
  pz_syncwarp();
  pz_ring_issue(c, c.word_idx / PZ_HALF_WORDS + 1u);
}
PZ_DEV void pz_refill(PzCtx &c) { /* appends 32 bits; requires cnt <= 32 */
  uint32_t w = c.ring[c.word_idx & (PZ_RING_WORDS - 1u)];
  c.bitbuf |= (uint64_t)w << c.cnt;
  c.cnt += 32u;
  c.word_idx++;
  if ((c.word_idx & (PZ_HALF_WORDS - 1u)) == 0u) pz_ring_cross(c);
}
PZ_DEV void pz_consume(PzCtx &c, uint32_t n) { c.bitbuf >>= n; c.cnt -= n; }
PZ_DEV uint64_t pz_cur_bit(const PzCtx &c) { return (uint64_t)c.word_idx * 32u - c.cnt; }
PZ_DEV int64_t pz_avail(const PzCtx &c) { return (int64_t)(c.end_bit - pz_cur_bit(c)); }
PZ_DEV void pz_seek(PzCtx &c, uint64_t bit) {
  uint32_t word = (uint32_t)(bit >> 5);
  uint32_t k = word / PZ_HALF_WORDS;
  pz_syncwarp();
  pz_ring_issue(c, k);
  pz_ring_issue(c, k + 1u);
  pz_async_wait_all();
  pz_syncwarp();
  c.word_idx = word; c.bitbuf = 0; c.cnt = 0;
  pz_refill(c);
  pz_consume(c, (uint32_t)(bit & 31u));
}
/* nextBits n (Monad.hs:199-230), n <= 16: running past the input is the truncation verdict
 * (Zlib.hs:38-39). */
PZ_DEV bool pz_take(PzCtx &c, uint32_t n, uint32_t &v) {
  if (pz_avail(c) < (int64_t)n) { pz_fail(c, PZ_ERR_DECOMPRESSION, PZ_D_RAN_OUT); return false; }
  if (c.cnt < n) pz_refill(c);
  v = (uint32_t)c.bitbuf & ((1u << n) - 1u);
  pz_consume(c, n);
  return true;
}

/* ---- exact bit-serial walk (nextCode / advanceTree, Monad.hs:295-302, HuffmanTree.hs:73-83)
 * Consumes one bit per step and stops exactly where the reference's trie walk stops:
 * truncation if the input ends first, "Advanced to empty tree!" on an unused prefix. */
PZ_DEV int pz_walk(PzCtx &c, const PzTree *t, const uint16_t *perm) {
  uint32_t code = 0, first = 0, index = 0;
  for (int len = 1; len <= 15; len++) {
    uint32_t bit;
    if (!pz_take(c, 1, bit)) return -1;
    if (t->nsyms == 0) { pz_fail(c, PZ_ERR_HUFFMAN_TREE, PZ_D_ADVANCE_EMPTY_TREE); return -1; }
    code |= bit;
    uint32_t count = t->cnt[len];
    if (code - first < count) return perm[index + (code - first)];
    index += count; first += count;
    if (code - first >= t->used[len]) break;
    first <<= 1; code <<= 1;
  }
  pz_fail(c, PZ_ERR_HUFFMAN_TREE, PZ_D_ADVANCED_TO_EMPTY);
  return -1;
}

/* ---- table construction ---------------------------------------------------------------- */
template <int KIND> /* 0 = code-length code, 1 = literal/length, 2 = distance */
PZ_DEV uint32_t pz_make_entry(uint32_t sym, uint32_t nbits) {
  if (KIND == 0) return PZ_ENTRY(nbits, nbits, PZ_T_LIT, sym);
  if (KIND == 1) {
    if (sym < 256u) return PZ_ENTRY(nbits, nbits, PZ_T_LIT, sym);
    if (sym == 256u) return PZ_ENTRY(nbits, nbits, PZ_T_EOB, 0);
    if (sym > 285u) return PZ_SLOW_ENTRY; /* lengthArray ! 286/287 is a bounds error */
    return PZ_ENTRY(nbits + PZ_LEN_EXTRA[sym - 257u], nbits, PZ_T_BASE, PZ_LEN_BASE[sym - 257u]);
  }
  if (sym > 29u) return PZ_SLOW_ENTRY; /* distanceArray ! >=30 is a bounds error */
  return PZ_ENTRY(nbits + PZ_DIST_EXTRA[sym], nbits, PZ_T_BASE, PZ_DIST_BASE[sym]);
}

/* computeCodeValues (Deflate.hs:261-288) from the sorted symbol list: codes[s] for every
 * symbol with a non-zero length.  `mask` keeps only the low `len` bits, which is all the
 * trie insertion ever inspects (testBit, HuffmanTree.hs:52,64). */
PZ_DEV void pz_canon_codes(const uint8_t *lens, const PzTree *t, const uint16_t *perm, uint16_t *codes, bool mask) {
  uint32_t nc[16], start[16];
  uint32_t code = 0, acc = 0;
  nc[0] = 0; start[0] = 0;
#pragma unroll
  for (int l = 1; l <= 15; l++) {
    code = (code + (l > 1 ? t->cnt[l - 1] : 0u)) << 1;
    nc[l] = code;
    start[l] = acc;
    acc += t->cnt[l];
  }
  pz_syncwarp();
  for (int p = pz_lane(); p < (int)t->nsyms; p += PZ_WARP) {
    int s = perm[p];
    int l = lens[s];
    uint32_t cv = 0;
#pragma unroll
    for (int k = 1; k <= 15; k++) if (k == l) cv = nc[k] + ((uint32_t)p - start[k]);
    codes[s] = (uint16_t)(mask ? (cv & ((1u << l) - 1u)) : cv);
  }
  pz_syncwarp();
}

/* createHuffmanTree's verdict when the lengths over-subscribe the code space: replay the
 * reference's insertion order (descending symbol, HuffmanTree.hs:29-34) on the canonical
 * codes and report the first collision.  `codes` is n uint16 scratch. */
PZ_DEV int pz_tree_error(const uint8_t *lens, int n, const PzTree *t, const uint16_t *perm, uint16_t *codes, int64_t *val) {
  pz_canon_codes(lens, t, perm, codes, true);
  for (int i = n - 1; i >= 0; i--) {
    int li = lens[i];
    if (!li) continue;
    uint32_t ci = codes[i];
    for (int jb = i + 1; jb < n; jb += PZ_WARP) {
      int j = jb + pz_lane();
      int k = 0;
      if (j < n) {
        int lj = lens[j];
        if (lj) {
          uint32_t cj = codes[j];
          if (lj < li) { if (cj == (ci >> (li - lj))) k = PZ_D_VALUE_HIT; }
          else if (lj == li) { if (cj == ci) k = PZ_D_TWO_VALUES; }
          else { if ((cj >> (lj - li)) == ci) k = PZ_D_LEAF_IS_NODE; }
        }
      }
      unsigned b = pz_ballot(k != 0);
      if (b) { *val = i; return pz_shfl(k, pz_ffs(b) - 1); }
    }
  }
  return 0; /* not reached when the Kraft sum exceeds 1 */
}

/* computeHuffmanTree (Deflate.hs:255-259) for symbols 0..n-1 with lengths lens[]: canonical
 * counts, symbols sorted by (length, symbol), and the 2^BITS-entry LUT, all built
 * cooperatively.  Returns 0, or the HuffmanTreeError detail with *val. */
template <int BITS, int KIND>
PZ_DEV int pz_build(const uint8_t *lens, int n, PzTree *t, uint16_t *perm, uint32_t *lut, uint32_t *scratch, int64_t *val) {
  uint32_t *cnt32 = scratch, *offs = scratch + 16;
  const int lane = pz_lane();
  pz_syncwarp();
  for (int i = lane; i < 16; i += PZ_WARP) cnt32[i] = 0;
  pz_syncwarp();
  for (int i = lane; i < n; i += PZ_WARP) {
    int l = lens[i];
    if (l) pz_smem_inc(&cnt32[l]);
  }
  pz_syncwarp();
  /* every lane derives the same canonical description */
  int32_t left = 1;
  bool over = false;
  uint32_t acc = 0, cl[16];
#pragma unroll
  for (int l = 1; l <= 15; l++) {
    cl[l] = cnt32[l];
    left = left * 2 - (int32_t)cl[l];
    if (left < 0) over = true;
  }
  pz_syncwarp();
  uint32_t used = 0;
  if (lane == 0) { t->used[15] = 0; t->cnt[0] = 0; t->used[0] = 0; }
#pragma unroll
  for (int l = 14; l >= 1; l--) {
    used = (cl[l + 1] + used + 1u) >> 1;
    if (lane == 0) t->used[l] = (uint16_t)used;
  }
#pragma unroll
  for (int l = 1; l <= 15; l++) {
    if (lane == 0) { t->cnt[l] = (uint16_t)cl[l]; offs[l] = acc; }
    acc += cl[l];
  }
  if (lane == 0) t->nsyms = (uint16_t)acc;
  pz_syncwarp();
  /* stable counting sort by length: perm[] */
  for (int b = 0; b < n; b += PZ_WARP) {
    int i = b + lane;
    uint32_t l = i < n ? lens[i] : 0u;
    unsigned m = pz_match_any(l);
    uint32_t rank = (uint32_t)pz_popc(m & pz_lanemask_lt());
    if (l) perm[offs[l] + rank] = (uint16_t)i;
    pz_syncwarp();
    if (l && rank == 0) offs[l] += (uint32_t)pz_popc(m);
    pz_syncwarp();
  }
  if (over) return pz_tree_error(lens, n, t, perm, (uint16_t *)lut, val);
  /* LUT, entry-major: each lane walks the canonical code along the bits of its index */
  for (uint32_t e = (uint32_t)lane; e < (1u << BITS); e += PZ_WARP) {
    uint32_t code = 0, first = 0, index = 0, entry = PZ_SLOW_ENTRY;
    for (int len = 1; len <= BITS; len++) {
      code |= (e >> (len - 1)) & 1u;
      uint32_t count = t->cnt[len];
      if (code - first < count) { entry = pz_make_entry<KIND>(perm[index + (code - first)], (uint32_t)len); break; }
      index += count; first += count;
      if (code - first >= t->used[len]) break; /* dead prefix: the careful path reports it */
      first <<= 1; code <<= 1;
    }
    lut[e] = entry;
  }
  pz_syncwarp();
  return 0;
}

/* ---- output ---------------------------------------------------------------------------- */
/* moveWindow / emitExcess (Monad.hs:338-347, OutputWindow.hs:45-54): at most one 32 KiB
 * chunk leaves the window per call, only once 64 KiB have accumulated. */
PZ_DEV void pz_move_window(PzCtx &c) {
  if (c.pos - c.base >= 2u * PZ_EXCESS) c.base += PZ_EXCESS;
}

/* emitPastChunk (Monad.hs:324-333, OutputWindow.hs:82-101).  Returns false with the
 * verdict set when the reference would fault or the caller's buffer is full. */
template <bool COUNT_ONLY>
PZ_DEV bool pz_match(PzCtx &c, uint32_t len, uint32_t dist) {
  uint32_t fill = c.pos - c.base;
  if (dist > fill) { pz_fail(c, PZ_REF_BOTTOM, PZ_D_BOT_DIST_TOO_FAR, dist, fill); return false; }
  if (fill + len > PZ_WINDOW) { pz_fail(c, PZ_REF_BOTTOM, PZ_D_BOT_WINDOW_OVERFLOW); return false; }
  if (c.pos + len > c.cap) { pz_fail(c, PZ_OUTPUT_FULL, 0); return false; }
  if (!COUNT_ONLY) {
    uint8_t *dst = c.out + c.pos;
    const uint8_t *src = dst - dist;
    const uint32_t lane = (uint32_t)pz_lane();
    pz_syncwarp(); /* earlier stores by other lanes are ordered before the loads below */
    if (dist >= len) {
      for (uint32_t i = lane; i < len; i += PZ_WARP) dst[i] = src[i];
    } else if (dist >= PZ_WARP) {
      for (uint32_t i0 = 0; i0 < len; i0 += PZ_WARP) { /* each chunk may read the previous one */
        uint32_t i = i0 + lane;
        if (i < len) dst[i] = src[i];
        pz_syncwarp();
      }
    } else { /* dist < 32 and dist < len: replicate the dist-byte pattern (copyChunked) */
      uint32_t m = lane % dist;
      const uint32_t step = PZ_WARP % dist;
      for (uint32_t i = lane; i < len; i += PZ_WARP) {
        dst[i] = src[m];
        m += step;
        if (m >= dist) m -= dist;
      }
    }
  }
  c.pos += len;
  pz_move_window(c);
  return true;
}

/* emitByte (Monad.hs:309-315, OutputWindow.hs:64-68) with the window / capacity checks. */
template <bool COUNT_ONLY>
PZ_DEV bool pz_literal_checked(PzCtx &c, uint32_t b) {
  if (c.pos - c.base >= PZ_WINDOW) { pz_fail(c, PZ_REF_BOTTOM, PZ_D_BOT_WINDOW_OVERFLOW); return false; }
  if (c.pos >= c.cap) { pz_fail(c, PZ_OUTPUT_FULL, 0); return false; }
  if (!COUNT_ONLY && pz_lane() == 0) c.out[c.pos] = (uint8_t)b;
  c.pos++;
  return true;
}

/* ---- the careful symbol: exact verdict order, used near the end of the input and whenever
 * the LUT cannot answer (Deflate.hs:106-120).  Returns 1 = continue, 0 = end of block,
 * -1 = verdict set. */
template <bool COUNT_ONLY>
PZ_DEV int pz_dist_careful(PzCtx &c, PzWarpSmem *sm, uint32_t len) {
  int ds = pz_walk(c, &sm->dist, sm->dist_perm);
  if (ds < 0) return -1;
  if (ds > 29) { pz_fail(c, PZ_REF_BOTTOM, PZ_D_BOT_DIST_SYM, ds); return -1; }
  uint32_t ex;
  if (!pz_take(c, PZ_DIST_EXTRA[ds], ex)) return -1;
  return pz_match<COUNT_ONLY>(c, len, PZ_DIST_BASE[ds] + ex) ? 1 : -1;
}
template <bool COUNT_ONLY>
PZ_DEV int pz_symbol_careful(PzCtx &c, PzWarpSmem *sm) {
  int sym = pz_walk(c, &sm->lit, sm->lit_perm);
  if (sym < 0) return -1;
  if (sym < 256) return pz_literal_checked<COUNT_ONLY>(c, (uint32_t)sym) ? 1 : -1;
  if (sym == 256) return 0;
  if (sym > 285) { pz_fail(c, PZ_REF_BOTTOM, PZ_D_BOT_LENGTH_SYM, sym); return -1; }
  uint32_t ex;
  if (!pz_take(c, PZ_LEN_EXTRA[sym - 257], ex)) return -1;
  return pz_dist_careful<COUNT_ONLY>(c, sm, PZ_LEN_BASE[sym - 257] + ex);
}

/* runInflate (Deflate.hs:106-120) */
template <bool COUNT_ONLY>
PZ_DEV bool pz_run_inflate(PzCtx &c, PzWarpSmem *sm) {
  for (;;) {
    uint32_t lim = c.base + PZ_WINDOW;
    if (c.cap < lim) lim = c.cap;
    bool fast = (int32_t)c.word_idx <= c.safe_word && c.pos + 258u <= lim;
    if (fast) {
      if (c.cnt <= 32u) pz_refill(c);
      uint32_t e = sm->lit_lut[(uint32_t)c.bitbuf & ((1u << PZ_LIT_BITS) - 1u)];
      uint32_t type = (e >> 12) & 3u;
      if (type != PZ_T_SLOW) {
        uint32_t saved = (uint32_t)c.bitbuf;
        pz_consume(c, e & 31u);
        if (type == PZ_T_LIT) {
          if (!COUNT_ONLY && pz_lane() == 0) c.out[c.pos] = (uint8_t)(e >> 16);
          c.pos++;
          continue;
        }
        if (type == PZ_T_EOB) return true;
        uint32_t nb = (e >> 8) & 15u;
        uint32_t len = (e >> 16) + ((saved & ((1u << (e & 31u)) - 1u)) >> nb);
        if (c.cnt <= 32u) pz_refill(c);
        uint32_t d = sm->dist_lut[(uint32_t)c.bitbuf & ((1u << PZ_DIST_BITS) - 1u)];
        if (((d >> 12) & 3u) == PZ_T_SLOW) {
          if (pz_dist_careful<COUNT_ONLY>(c, sm, len) < 0) return false;
          continue;
        }
        saved = (uint32_t)c.bitbuf;
        pz_consume(c, d & 31u);
        uint32_t dist = (d >> 16) + ((saved & ((1u << (d & 31u)) - 1u)) >> ((d >> 8) & 15u));
        if (!pz_match<COUNT_ONLY>(c, len, dist)) return false;
        continue;
      }
    }
    int r = pz_symbol_careful<COUNT_ONLY>(c, sm);
    if (r < 0) return false;
    if (r == 0) return true;
  }
}

/* One symbol of the code-length code plus its repeat count (getCodeLengths, Deflate.hs:124-156). */
PZ_DEV int pz_pre_symbol(PzCtx &c, PzWarpSmem *sm, uint32_t *pre_lut) {
  if (pz_avail(c) >= PZ_PRE_BITS) { /* all peeked bits are real; the repeat field uses pz_take */
    if (c.cnt < (uint32_t)PZ_PRE_BITS) pz_refill(c);
    uint32_t e = pre_lut[(uint32_t)c.bitbuf & ((1u << PZ_PRE_BITS) - 1u)];
    if (((e >> 12) & 3u) != PZ_T_SLOW) { pz_consume(c, e & 31u); return (int)(e >> 16); }
  }
  return pz_walk(c, &sm->pre, sm->pre_perm);
}

/* inflateBlock's dynamic arm (Deflate.hs:83-101): returns false with the verdict set. */
PZ_DEV bool pz_dynamic_header(PzCtx &c, PzWarpSmem *sm) {
  uint32_t hlit, hdist, hclen, v;
  if (!pz_take(c, 5, hlit)) return false;
  if (!pz_take(c, 5, hdist)) return false;
  if (!pz_take(c, 4, hclen)) return false;
  hlit += 257u; hdist += 1u; hclen += 4u;
  const int lane = pz_lane();
  pz_syncwarp();
  for (int i = lane; i < 19; i += PZ_WARP) sm->lens[i] = 0;
  pz_syncwarp();
  for (uint32_t i = 0; i < hclen; i++) {
    if (!pz_take(c, 3, v)) return false;
    if (lane == 0) sm->lens[PZ_CL_ORDER[i]] = (uint8_t)v;
  }
  int64_t val = 0;
  uint32_t *pre_lut = sm->dist_lut;
  int e = pz_build<PZ_PRE_BITS, 0>(sm->lens, 19, &sm->pre, sm->pre_perm, pre_lut, sm->scratch, &val);
  if (e) { pz_fail(c, PZ_ERR_HUFFMAN_TREE, e, e == PZ_D_LEAF_IS_NODE ? val : 0); return false; }
  /* the code lengths; repeats are not clipped at hlit+hdist (Deflate.hs:153-156) */
  uint32_t n = 0, prev = 0;
  const uint32_t maxl = hlit + hdist;
  while (n < maxl) {
    int code = pz_pre_symbol(c, sm, pre_lut);
    if (code < 0) return false;
    if (code <= 15) {
      if (lane == 0) sm->lens[n] = (uint8_t)code;
      n++; prev = (uint32_t)code;
    } else {
      uint32_t num, fill;
      if (code == 16) { if (!pz_take(c, 2, num)) return false; num += 3u; fill = prev; }
      else if (code == 17) { if (!pz_take(c, 3, num)) return false; num += 3u; fill = 0; prev = 0; }
      else { if (!pz_take(c, 7, num)) return false; num += 11u; fill = 0; prev = 0; }
      for (uint32_t i = (uint32_t)lane; i < num; i += PZ_WARP) sm->lens[n + i] = (uint8_t)fill;
      n += num;
    }
  }
  pz_syncwarp();
  e = pz_build<PZ_LIT_BITS, 1>(sm->lens, (int)hlit, &sm->lit, sm->lit_perm, sm->lit_lut, sm->scratch, &val);
  if (e) { pz_fail(c, PZ_ERR_HUFFMAN_TREE, e, e == PZ_D_LEAF_IS_NODE ? val : 0); return false; }
  e = pz_build<PZ_DIST_BITS, 2>(sm->lens + hlit, (int)(n - hlit), &sm->dist, sm->dist_perm, sm->dist_lut, sm->scratch, &val);
  if (e) { pz_fail(c, PZ_ERR_HUFFMAN_TREE, e, e == PZ_D_LEAF_IS_NODE ? val : 0); return false; }
  return true;
}

/* buildFixedLitTree / buildFixedDistanceTree (Deflate.hs:241-251) */
PZ_DEV void pz_fixed_tables(PzWarpSmem *sm) {
  pz_syncwarp();
  for (int i = pz_lane(); i < 288 + 32; i += PZ_WARP)
    sm->lens[i] = (uint8_t)(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : i < 288 ? 8 : 5);
  pz_syncwarp();
  int64_t val;
  pz_build<PZ_LIT_BITS, 1>(sm->lens, 288, &sm->lit, sm->lit_perm, sm->lit_lut, sm->scratch, &val);
  pz_build<PZ_DIST_BITS, 2>(sm->lens + 288, 32, &sm->dist, sm->dist_perm, sm->dist_lut, sm->scratch, &val);
}

/* The stored arm (Deflate.hs:70-78, Monad.hs:265-293 for a single-chunk input). */
template <bool COUNT_ONLY>
PZ_DEV bool pz_stored_block(PzCtx &c) {
  uint32_t len, nlen;
  uint64_t cur = pz_cur_bit(c);
  uint32_t drop = (uint32_t)((8u - (cur & 7u)) & 7u); /* advanceToByte */
  if (drop) {
    /* the dropped bits belong to a byte that was already fetched: never a truncation */
    if (c.cnt < drop) pz_refill(c);
    pz_consume(c, drop);
  }
  if (pz_avail(c) < 32) { pz_fail(c, PZ_ERR_DECOMPRESSION, PZ_D_RAN_OUT); return false; }
  if (!pz_take(c, 16, len)) return false;
  if (!pz_take(c, 16, nlen)) return false;
  if (len != ((~nlen) & 0xffffu)) { pz_fail(c, PZ_ERR_FORMAT, PZ_D_LEN_NLEN); return false; }
  uint64_t boff = pz_cur_bit(c) >> 3;
  uint64_t remaining = (c.end_bit >> 3) - boff;
  /* getBlock takes the data only when strictly more than len bytes are left in the chunk */
  if ((uint64_t)len >= remaining) { pz_fail(c, PZ_ERR_DECOMPRESSION, PZ_D_RAN_OUT); return false; }
  uint32_t fill = c.pos - c.base;
  if (fill + len > PZ_WINDOW) { pz_fail(c, PZ_REF_BOTTOM, PZ_D_BOT_WINDOW_OVERFLOW); return false; }
  if (c.pos + len > c.cap) { pz_fail(c, PZ_OUTPUT_FULL, 0); return false; }
  if (!COUNT_ONLY) {
    const uint8_t *src = c.in_al + boff;
    uint8_t *dst = c.out + c.pos;
    for (uint32_t i = (uint32_t)pz_lane(); i < len; i += PZ_WARP) dst[i] = src[i];
  }
  c.pos += len;
  pz_seek(c, (boff + len) * 8u);
  return true;
}

/* `decompress` for one single-chunk stream: inflateWithHeaders (Zlib.hs:53-69), inflate
 * (Deflate.hs:39-63).  The Adler-32 comparison itself is done by the checksum kernels; this
 * function leaves the stored trailer in *adler_stored. */
template <bool COUNT_ONLY>
PZ_DEV void pz_inflate_stream(const uint8_t *in, uint64_t in_len, uint8_t *out, uint64_t out_cap, PzWarpSmem *sm, pz_result *res) {
  PzCtx c;
  uint64_t mis = (uint64_t)((uintptr_t)in & 15u);
  c.in_al = in - mis;
  c.in_al_bytes = (mis + in_len + 15u) & ~(uint64_t)15u;
  c.end_bit = (mis + in_len) * 8u;
  c.ring = sm->ring;
  c.out = out;
  c.pos = 0; c.base = 0;
  c.cap = out_cap > 0xfffdff00ull ? 0xfffdff00u : (uint32_t)out_cap; /* base + 128 KiB stays in 32 bits */
  c.status = PZ_OK; c.detail = 0; c.p0 = 0; c.p1 = 0;
  c.safe_word = c.end_bit >= 48u ? (int32_t)((c.end_bit - 48u) >> 5) : -1;
  uint32_t adler_stored = 0;
  pz_seek(c, mis * 8u);

  do {
    /* zlib header (Zlib.hs:53-69) */
    uint32_t cmf, flg;
    if (!pz_take(c, 8, cmf)) break;
    if (!pz_take(c, 8, flg)) break;
    if (((cmf << 8) | flg) % 31u != 0u) { pz_fail(c, PZ_ERR_HEADER, PZ_D_HDR_CHECKSUM); break; }
    if ((cmf & 15u) != 8u) { pz_fail(c, PZ_ERR_HEADER, PZ_D_HDR_METHOD, cmf & 15u); break; }
    if ((cmf >> 4) > 7u) { pz_fail(c, PZ_ERR_HEADER, PZ_D_HDR_WINDOW, cmf >> 4); break; }
    if (flg & 0x20u) { /* FDICT: the four DICTID bytes are skipped (Zlib.hs:68) */
      uint32_t skip;
      if (!pz_take(c, 16, skip)) break;
      if (!pz_take(c, 16, skip)) break;
    }
    bool fixed_ready = false;
    for (;;) { /* inflate's go loop (Deflate.hs:45-50) */
      uint32_t bfinal, btype;
      if (!pz_take(c, 1, bfinal)) break;
      if (!pz_take(c, 2, btype)) break;
      if (btype == 0u) {
        if (!pz_stored_block<COUNT_ONLY>(c)) break;
      } else if (btype == 1u) {
        if (!fixed_ready) { pz_fixed_tables(sm); fixed_ready = true; }
        if (!pz_run_inflate<COUNT_ONLY>(c, sm)) break;
      } else if (btype == 2u) {
        fixed_ready = false;
        if (!pz_dynamic_header(c, sm)) break;
        if (!pz_run_inflate<COUNT_ONLY>(c, sm)) break;
      } else {
        pz_fail(c, PZ_ERR_FORMAT, PZ_D_BAD_BTYPE, 3);
        break;
      }
      pz_move_window(c);
      if (bfinal) {
        /* checkChecksum (Deflate.hs:52-63): align, then four bytes, most significant first */
        uint64_t cur = pz_cur_bit(c);
        uint32_t drop = (uint32_t)((8u - (cur & 7u)) & 7u);
        if (drop) { if (c.cnt < drop) pz_refill(c); pz_consume(c, drop); }
        uint32_t hi, lo;
        if (pz_avail(c) < 32) { pz_fail(c, PZ_ERR_DECOMPRESSION, PZ_D_RAN_OUT); break; }
        if (!pz_take(c, 16, hi)) break;
        if (!pz_take(c, 16, lo)) break;
        adler_stored = ((hi & 0xffu) << 24) | ((hi >> 8) << 16) | ((lo & 0xffu) << 8) | (lo >> 8);
        break;
      }
    }
  } while (0);

  pz_syncwarp();
  if (pz_lane() == 0) {
    res->status = c.status;
    res->detail = c.detail;
    res->out_len = c.pos;
    res->adler_computed = 0;
    res->adler_stored = adler_stored;
    res->err_bitpos = pz_cur_bit(c) - mis * 8u;
    res->payload[0] = c.p0;
    /* payload[1]: bytes the reference has already published as 32 KiB chunks (the shim's
     * incremental driver needs it); DIST_TOO_FAR keeps the retained-byte count instead */
    res->payload[1] = (c.status == PZ_REF_BOTTOM && c.detail == PZ_D_BOT_DIST_TOO_FAR) ? c.p1 : (int64_t)c.base;
  }
}
