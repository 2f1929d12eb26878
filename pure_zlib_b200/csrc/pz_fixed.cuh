/*
 * pz_fixed.cuh -- K5: small streams made of fixed-Huffman blocks, one LANE per stream.
 *
 * K1 gives every stream a slot of shared memory (two look-up tables, an input ring, a token queue) and a lane of the one
 * hot warp of its SM: 28 chains per SM, each paced by the latency of its own symbol chain.  That is the right shape for a
 * few thousand long streams and the wrong one for a million 4 KiB records (BASELINE configs[2]): there the parallelism is
 * across streams, not inside them.  A record compressed with the fixed code (BTYPE 1, Deflate.hs:79-82, 241-251) needs no
 * table at all -- the code is arithmetic: 7 bits for 256..279, 8 for 0..143 and 280..287, 9 for 144..255, 5-bit distances --
 * so here a stream is ONE thread: it reads its bits with two aligned word loads per symbol, decodes the symbol from the
 * bit-reversed window, writes literals and copies matches itself (byte by byte: the replicate semantics of copyChunked,
 * OutputWindow.hs:94-101, come for free), and 2 048 of them per SM hide each other's latencies.  A loop iteration is either
 * "one symbol" or "up to eight bytes of the pending copy", so lanes of a warp that sit in a long match hold the others up
 * for at most a few iterations.
 *
 * Memory.  Thirty-two lanes of a warp write thirty-two different streams, so every access is a transaction of its own and
 * the kernel is bound by their number, not by instructions (the sizing pass, which touches no output, runs eight times
 * faster).  Output therefore leaves a lane as aligned 32-bit words: produced bytes are collected in a register (`acc`: the
 * bytes of the word that contains position pos) and stored when the word is full, and a match whose source lies at least
 * 11 bytes back -- every byte it reads is then in memory, not in the register -- is copied eight bytes at a time from
 * three aligned word loads.  Closer matches (the replicate case, a few per cent of text) take the byte-wise path between a
 * flush and a reload of the register.
 *
 * The kernel only ever reports SUCCESS.  Anything else -- a header that is not a plain zlib header, a block that is not a
 * fixed one, a symbol the reference cannot index, a distance beyond the output, the end of the input or of the caller's
 * buffer in sight, a gap of more than 32 KiB between two moveWindow calls (PzCtx::mark: the window's base is then no
 * longer the closed form) -- leaves the stream PENDING, and K1, launched behind, decodes it from its first byte with the
 * reference's exact order of checks.  So every verdict other than plain success still comes from the one place that
 * reproduces the reference, as with K2 (pz_stored.cuh).  Adler-32 is K3's, as for any other stream.
 *
 * DYNAMIC blocks (DYN = true, the second kernel of this file: K6).  zlib writes a dynamic block for 4 KiB of text unless told
 * otherwise, so the same one-thread-per-stream decoder also exists with tables: the thread parses the block header itself
 * (Deflate.hs:83-101, 124-156), builds a 10-bit literal/length and a 9-bit distance look-up table in its LOCAL memory
 * (3 KiB per thread, served by L1 / L2: a look-up costs a few hundred cycles, and a few hundred threads per SM wait
 * for one at any time -- K1 has 28 chains per SM, here there are up to 1 024) and decodes with them.  It takes exactly what
 * zlib's tree builder emits and nothing else: more than 286 / 30 codes, a repeat code that starts the list or runs past it,
 * an over-subscribed code, no end-of-block code, a code longer than its table's index, an unused table entry -- all K1's.
 * MEASURED (B200, config 3): K6 is slower than K1 on the same streams (28 ms against 25 ms for the dynamic quarter): a gigabyte
 * of tables is in flight, every look-up is a DRAM sector.  It is therefore opt-in (PZ_K6=1); see DESIGN.md, K5 / K6.
 *
 * Runs between K2 and K1 on batches of at least PZ_FIXED_MIN_STREAMS streams, for streams of at most PZ_FIXED_MAX_IN
 * compressed bytes (a lone thread is slower than K1's hot lane: it is the number of streams that makes this path fast).
 */
#pragma once
#include <stdint.h>

#include "pz_device.cuh"

#define PZ_FIXED_MAX_IN 16384u
#ifndef PZ_FIXED_MIN_STREAMS
#define PZ_FIXED_MIN_STREAMS 8192u
#endif
#define PZ_FIXED_THREADS 256

/* Deflate.hs:160-237, indexed by symbol - 257 / distance symbol: base | extra bits << 16 */
#ifdef PZ_HOSTSIM
#define PZ_FX_TABLE static const
#else
#define PZ_FX_TABLE static __constant__
#endif
PZ_FX_TABLE uint32_t PZ_FX_LEN[29] = {
    3 | 0 << 16, 4 | 0 << 16, 5 | 0 << 16, 6 | 0 << 16, 7 | 0 << 16, 8 | 0 << 16, 9 | 0 << 16, 10 | 0 << 16, 11 | 1 << 16, 13 | 1 << 16,
    15 | 1 << 16, 17 | 1 << 16, 19 | 2 << 16, 23 | 2 << 16, 27 | 2 << 16, 31 | 2 << 16, 35 | 3 << 16, 43 | 3 << 16, 51 | 3 << 16, 59 | 3 << 16,
    67 | 4 << 16, 83 | 4 << 16, 99 | 4 << 16, 115 | 4 << 16, 131 | 5 << 16, 163 | 5 << 16, 195 | 5 << 16, 227 | 5 << 16, 258 | 0 << 16};
PZ_FX_TABLE uint32_t PZ_FX_DIST[30] = {
    1 | 0 << 16, 2 | 0 << 16, 3 | 0 << 16, 4 | 0 << 16, 5 | 1 << 16, 7 | 1 << 16, 9 | 2 << 16, 13 | 2 << 16, 17 | 3 << 16, 25 | 3 << 16,
    33 | 4 << 16, 49 | 4 << 16, 65 | 5 << 16, 97 | 5 << 16, 129 | 6 << 16, 193 | 6 << 16, 257 | 7 << 16, 385 | 7 << 16, 513 | 8 << 16, 769 | 8 << 16,
    1025 | 9 << 16, 1537 | 9 << 16, 2049 | 10 << 16, 3073 | 10 << 16, 4097 | 11 << 16, 6145 | 11 << 16, 8193 | 12 << 16, 12289 | 12 << 16,
    16385 | 13 << 16, 24577 | 13 << 16};

/* One stream, start to finish, by the calling thread.  Returns true with *res filled (status PZ_OK last) if the stream
 * decoded completely; false -- nothing of *res touched -- if it is K1's (see the header of this file). */
#ifndef PZ_SM_LIT_BITS
#define PZ_SM_LIT_BITS 10 /* K6: index bits of a thread's literal/length table (entries: symbol | code length << 9) */
#endif
#ifndef PZ_SM_DIST_BITS
#define PZ_SM_DIST_BITS 9 /* ... and of its distance table (entries: symbol | code length << 5) */
#endif
template <bool COUNT_ONLY, bool DYN = false>
PZ_DEV bool pz_fixed_stream(const uint8_t *in, uint64_t n64, uint8_t *out, uint64_t cap64, pz_result *res) {
  /* K6's tables: thread-private, i.e. local memory */
  uint16_t lit_t[DYN ? (1 << PZ_SM_LIT_BITS) : 1];
  uint16_t dist_t[DYN ? (1 << PZ_SM_DIST_BITS) : 1];
  uint8_t lens[DYN ? 320 : 1];
  uint8_t pre_t[DYN ? 128 : 1];
  bool dyn_block = false;
  if (n64 < 8u || n64 > PZ_FIXED_MAX_IN) return false;
  const uint32_t n = (uint32_t)n64;
  /* inflateWithHeaders (Zlib.hs:53-69): only a header that passes every check; FDICT skips four bytes */
  const uint32_t cmf = in[0], flg = in[1];
  if (((cmf << 8) | flg) % 31u != 0u || (cmf & 15u) != 8u || (cmf >> 4) > 7u) return false;
  const uint32_t p = (flg & 0x20u) ? 6u : 2u;
  {
    const uint32_t bt = p < n ? (in[p] >> 1) & 3u : 3u; /* the first block decides whether the stream is tried at all */
    if (!(bt == 1u || (DYN && bt == 2u))) return false;
  }
  const uint32_t cap = (COUNT_ONLY || cap64 > 0xfffdff00ull) ? 0xfffdff00u : (uint32_t)cap64;
  /* bit positions count from the aligned word at or before the stream's first byte (the blob has 15 readable bytes of slack
   * at both ends, pzcuda.h) */
  const uint32_t mis = (uint32_t)((uintptr_t)in & 3u);
  const uint32_t *w = reinterpret_cast<const uint32_t *>(in - mis);
  uint32_t bp = (mis + p) * 8u;
  const uint32_t end_bit = (mis + n) * 8u;
  /* The bits at bp live in a 64-bit register that takes one aligned word whenever fewer than 33 are left (a symbol needs at
   * most 31): one load per 32 bits of input instead of two per symbol -- in this kernel every load of a lane is a
   * transaction of its own, and their number is what bounds it. */
  uint32_t wi = (bp >> 5) + 1u;                /* next word to take */
  uint64_t bb = (uint64_t)w[bp >> 5] >> (bp & 31u);
  uint32_t bc = 32u - (bp & 31u);              /* valid bits in bb */
  uint32_t pos = 0, mark = 0; /* bytes produced; bytes produced at the last moveWindow call (Monad.hs:338-347: after every match and block) */
  uint32_t rem = 0, dist = 0; /* the pending copy */
  uint32_t bfinal = 0;
  bool in_block = false;
  /* output as aligned words (see above): acc holds the bytes [pos & ~3, pos) of the output, which are NOT in memory yet */
  const bool wide = !COUNT_ONLY && ((uintptr_t)out & 3u) == 0u;
  uint32_t acc = 0;
  /* One iteration = at most one symbol decoded (lanes whose copy is still running skip that part) and at most eight bytes
   * produced -- the literal, or the next piece of the copy -- by ONE piece of code for both: the lanes of a warp are in
   * different streams, and every path that only some of them take is time the others wait (ncu: with separate literal, match
   * and copy paths 7.5 of 32 lanes were active on average). */
  bool done = false;
  for (;;) {
    uint64_t v = 0;   /* the bytes this iteration produces, first byte lowest */
    uint32_t c = 0;   /* how many: 0..8 */
    bool is_literal = false;
    if (rem == 0u) {
      /* the 32 stream bits at bp: a length / distance pair of the fixed code takes at most 8 + 5 + 5 + 13 = 31 of them */
      if (bc <= 32u) { bb |= (uint64_t)w[wi++] << bc; bc += 32u; }
      const uint32_t lo = (uint32_t)bb;
      if (!in_block) { /* inflateBlock (Deflate.hs:65-104): BFINAL, BTYPE */
        const uint32_t bt = (lo >> 1) & 3u;
        if (bp + 3u > end_bit || !(bt == 1u || (DYN && bt == 2u))) return false;
        bfinal = lo & 1u;
        bp += 3u; bb >>= 3; bc -= 3u;
        in_block = true;
        dyn_block = false;
      if (DYN && bt == 2u) { /* the dynamic arm (Deflate.hs:83-101): code lengths, then the two tables of this block */
        dyn_block = true;
#define PZ_SM_TAKE(nbits, v) do { if (bc <= 32u) { bb |= (uint64_t)w[wi++] << bc; bc += 32u; } (v) = (uint32_t)bb & ((1u << (nbits)) - 1u); bb >>= (nbits); bc -= (nbits); bp += (nbits); } while (0)
        uint32_t hlit, hdist, hclen, v;
        PZ_SM_TAKE(5, hlit); PZ_SM_TAKE(5, hdist); PZ_SM_TAKE(4, hclen);
        hlit += 257u; hdist += 1u; hclen += 4u;
        if (hlit > 286u || hdist > 30u) return false; /* the reference takes up to 288 / 32: K1's */
        /* code-length code: 3 bits per symbol in codeLengthOrder (Deflate.hs:290-292); complete or K1's */
        uint32_t pl[19];
#pragma unroll
        for (int k = 0; k < 19; k++) pl[k] = 0;
        uint32_t kraft = 0;
        for (uint32_t k = 0; k < hclen; k++) {
          PZ_SM_TAKE(3, v);
          pl[PZ_CL_ORDER[k]] = v;
          if (v) kraft += 128u >> v;
        }
        if (kraft != 128u || bp > end_bit) return false;
        {
          uint32_t code = 0;
          for (uint32_t l = 1; l <= 7u; l++) { /* canonical codes (computeCodeValues, Deflate.hs:261-288), most significant bit first */
            for (uint32_t sy = 0; sy < 19u; sy++) {
              if (pl[sy] != l) continue;
              for (uint32_t e = pz_brev(code) >> (32u - l); e < 128u; e += 1u << l) pre_t[e] = (uint8_t)(sy | (l << 5));
              code++;
            }
            code <<= 1;
          }
        }
        /* getCodeLengths (Deflate.hs:124-156) as zlib writes them: no repeat at the start, none past the end */
        const uint32_t total = hlit + hdist;
        uint32_t nl = 0, prev = 0;
        while (nl < total) {
          if (bc <= 32u) { bb |= (uint64_t)w[wi++] << bc; bc += 32u; }
          const uint32_t e = pre_t[(uint32_t)bb & 127u];
          const uint32_t sy = e & 31u, l = e >> 5;
          bb >>= l; bc -= l; bp += l;
          uint32_t rep = 1, val = sy;
          if (sy == 16u) { if (nl == 0u) return false; PZ_SM_TAKE(2, rep); rep += 3u; val = prev; }
          else if (sy == 17u) { PZ_SM_TAKE(3, rep); rep += 3u; val = 0; }
          else if (sy == 18u) { PZ_SM_TAKE(7, rep); rep += 11u; val = 0; }
          if (nl + rep > total || bp > end_bit) return false;
          for (uint32_t k = 0; k < rep; k++) lens[nl + k] = (uint8_t)val;
          nl += rep;
          prev = val;
        }
        if (lens[256] == 0u) return false; /* no end-of-block code: the reference would run on */
        /* the two tables: every entry starts as "not mine" (0), then each code of at most the table's index bits fills the
         * entries that end in its reversed bits; an over-subscribed code is K1's to reject (HuffmanTree.hs:25-71) */
        for (int pass = 0; pass < 2; pass++) {
          const uint32_t nsym = pass == 0 ? hlit : hdist, off = pass == 0 ? 0u : hlit;
          const uint32_t bits = pass == 0 ? PZ_SM_LIT_BITS : PZ_SM_DIST_BITS, shift = pass == 0 ? 9u : 5u;
          uint16_t *tab = pass == 0 ? lit_t : dist_t;
          uint32_t cnt[16];
#pragma unroll
          for (int k = 0; k < 16; k++) cnt[k] = 0;
          for (uint32_t sy = 0; sy < nsym; sy++) cnt[lens[off + sy]]++;
          cnt[0] = 0;
          uint32_t next[16], code = 0, kr = 0;
#pragma unroll
          for (int l = 1; l < 16; l++) { code = (code + cnt[l - 1]) << 1; next[l] = code; kr += cnt[l] << (15 - l); }
          if (kr > 32768u) return false;
          for (uint32_t e = 0; e < (1u << bits); e++) tab[e] = 0;
          for (uint32_t sy = 0; sy < nsym; sy++) {
            const uint32_t l = lens[off + sy];
            if (l == 0u) continue;
            uint32_t cv = 0;
#pragma unroll
            for (int k = 1; k < 16; k++) if ((uint32_t)k == l) cv = next[k]++;
            if (l > bits) continue; /* a longer code: its entries stay "not mine" */
            for (uint32_t e = pz_brev(cv) >> (32u - l); e < (1u << bits); e += 1u << l) tab[e] = (uint16_t)(sy | (l << shift));
          }
        }
#undef PZ_SM_TAKE
      }
        continue;
      }
      uint32_t sym, nb;
      if (DYN && dyn_block) { /* this block's own code: one look-up (local memory) */
        const uint32_t e = lit_t[lo & ((1u << PZ_SM_LIT_BITS) - 1u)];
        if (e == 0u) return false; /* a code longer than the table's index, or a prefix nothing is assigned to */
        sym = e & 511u; nb = e >> 9;
      } else {
        /* the fixed literal/length code, first bit of a code = its most significant one (HuffmanTree.hs:73-83) */
        const uint32_t r = pz_brev(lo) >> 23; /* the next nine stream bits as a number, first bit on top */
        if ((r >> 2) < 24u) { sym = 256u + (r >> 2); nb = 7u; }
        else if ((r >> 1) < 192u) { sym = (r >> 1) - 48u; nb = 8u; }
        else if ((r >> 1) < 200u) { sym = 280u + ((r >> 1) - 192u); nb = 8u; }
        else { sym = 144u + (r - 400u); nb = 9u; }
      }
      if (sym < 256u) { /* emitByte (Monad.hs:309-315) */
        if (bp + nb > end_bit || pos >= cap || pos + 1u - mark > PZ_EXCESS) return false;
        bp += nb; bb >>= nb; bc -= nb;
        v = sym; c = 1u; is_literal = true;
      } else if (sym == 256u) { /* end of block: moveWindow, then the next block or the trailer (Deflate.hs:45-50) */
        if (bp + nb > end_bit) return false;
        bp += nb; bb >>= nb; bc -= nb;
        mark = pos;
        in_block = false;
        if (bfinal) done = true;
      } else {
        if (sym > 285u) return false; /* lengthArray ! 286 / 287 (Deflate.hs:161,167): the exact kernel words it */
        const uint32_t le = PZ_FX_LEN[sym - 257u];
        const uint32_t len = (le & 0xffffu) + ((lo >> nb) & ~(0xffffffffu << (le >> 16)));
        nb += le >> 16; /* <= 13 (fixed code), <= 15 (a table's) */
        uint32_t dsym, lo2 = lo;
        if (DYN && dyn_block) { /* length and distance part may take 10 + 5 + 9 + 13 bits together: consume the first, look again */
          if (bp + nb > end_bit) return false;
          bp += nb; bb >>= nb; bc -= nb; nb = 0;
          if (bc <= 32u) { bb |= (uint64_t)w[wi++] << bc; bc += 32u; }
          lo2 = (uint32_t)bb;
          const uint32_t e = dist_t[lo2 & ((1u << PZ_SM_DIST_BITS) - 1u)];
          if (e == 0u) return false;
          dsym = e & 31u; nb = e >> 5;
        } else {
          dsym = pz_brev((lo >> nb) & 31u) >> 27;
          nb += 5u; /* <= 18 */
        }
        if (dsym > 29u) return false; /* distanceArray ! 30 / 31 (Deflate.hs:200,206) */
        const uint32_t de = PZ_FX_DIST[dsym];
        const uint32_t x = de >> 16; /* <= 13 */
        dist = (de & 0xffffu) + ((lo2 >> nb) & ~(0xffffffffu << x));
        nb += x; /* <= 31 */
        /* what pz_match() checks: the distance lies inside what exists (OutputWindow.hs:82-89: below 64 KiB of output that is
         * all of it, above it at least 32 KiB are retained), the bytes fit, and the gap rule */
        if (bp + nb > end_bit || dist > pos || len > cap - pos || pos + len - mark > PZ_EXCESS) return false;
        bp += nb; bb >>= nb; bc -= nb;
        if (COUNT_ONLY) { pos += len; mark = pos; }
        else rem = len;
      }
    }
    if (done) break;
    if (COUNT_ONLY) { pos += c; continue; } /* (c = 1 for a literal; matches were counted above) */
    /* ---- produce: the literal, or up to eight bytes of the copy (emitPastChunk, Monad.hs:324-333) ---- */
    if (rem != 0u) {
      c = rem < 8u ? rem : 8u;
      if (wide && dist >= 11u) { /* the source bytes [pos - dist, pos - dist + 8) lie below pos & ~3: all in memory */
        const uint8_t *q = out + pos - dist;
        const uint32_t *qa = reinterpret_cast<const uint32_t *>((uintptr_t)q & ~(uintptr_t)3);
        const uint32_t qs = (uint32_t)((uintptr_t)q & 3u) * 8u;
#ifdef PZ_EXP_K5_NOLOAD /* timing experiment (wrong output): what the history reads cost */
        const uint32_t s0 = (uint32_t)(uintptr_t)qa, s1 = s0 + 1u, s2 = s0 + 2u;
#else
        const uint32_t s0 = qa[0], s1 = qa[1], s2 = qa[2];
#endif
        v = (uint64_t)pz_funnel_r(s0, s1, qs) | ((uint64_t)pz_funnel_r(s1, s2, qs) << 32);
        if (c < 8u) v &= ~(~0ull << (8u * c));
      } else { /* a close match (or an output slice that is not word-aligned): byte by byte, in order -- a byte may be one
                  this loop has just written (dist < len) */
        uint8_t *d = out + pos;
        if (wide) { /* the pending bytes go to memory first: the copy may read them */
          const uint32_t a = pos & 3u;
#pragma unroll
          for (uint32_t j = 0; j < 3u; j++)
            if (j < a) d[(int32_t)j - (int32_t)a] = (uint8_t)(acc >> (8u * j));
        }
        const uint8_t *q = d - dist;
#pragma unroll
        for (uint32_t j = 0; j < 8u; j++)
#ifdef PZ_EXP_K5_NOLOAD
          if (j < c) d[j] = (uint8_t)(uintptr_t)q;
#elif defined(PZ_EXP_K5_NOSTORE)
          if (j < c && q[j] == 0x7fu && dist == 0xffffffu) d[j] = q[j];
#else
          if (j < c) d[j] = q[j];
#endif
        pos += c;
        rem -= c;
        if (rem == 0u) mark = pos;
        if (wide) { /* and the register takes the bytes of the word the output now ends in */
          const uint32_t a2 = pos & 3u;
          acc = a2 ? *reinterpret_cast<const uint32_t *>(out + (pos & ~3u)) & ~(0xffffffffu << (8u * a2)) : 0u;
        }
        continue;
      }
      rem -= c;
    }
    if (c == 0u) continue; /* a block header or an end of block: nothing to produce */
    if (wide) { /* append c bytes behind the a pending ones: up to 11 bytes = two full words and a rest */
      const uint32_t a = pos & 3u;
      const uint64_t comb = (uint64_t)acc | (v << (8u * a));
      const uint32_t over = a ? (uint32_t)(v >> (64u - 8u * a)) : 0u;
      const uint32_t total = a + c, full = total >> 2;
      uint32_t *dw = reinterpret_cast<uint32_t *>(out + (pos & ~3u));
#ifdef PZ_EXP_K5_NOSTORE /* timing experiment (wrong output): what the stores cost (the condition is never true, the values stay live) */
      if (full >= 1u && comb == 0x0123456789abcdefull) dw[0] = (uint32_t)comb;
      if (full >= 2u && comb == 0x0123456789abcdefull) dw[1] = (uint32_t)(comb >> 32);
#else
      if (full >= 1u) dw[0] = (uint32_t)comb;
      if (full >= 2u) dw[1] = (uint32_t)(comb >> 32);
#endif
      const uint32_t restw = full == 0u ? (uint32_t)comb : full == 1u ? (uint32_t)(comb >> 32) : over;
      const uint32_t rb = total & 3u;
      acc = rb ? restw & ~(0xffffffffu << (8u * rb)) : 0u;
    } else {
      out[pos] = (uint8_t)v; /* only literals come here: copies of an unaligned slice took the byte path above */
    }
    pos += c;
    if (!is_literal && rem == 0u) mark = pos;
  }
  /* checkChecksum (Deflate.hs:52-63): to the next byte boundary, four bytes, most significant first; K3 compares */
  const uint32_t tb = ((bp + 7u) >> 3) - mis; /* byte offset of the trailer in the stream */
  if (tb + 4u > n) return false;
  if (wide) { /* the last, incomplete word */
    const uint32_t a = pos & 3u;
    for (uint32_t j = 0; j < a; j++) out[(pos & ~3u) + j] = (uint8_t)(acc >> (8u * j));
  }
  res->detail = 0;
  res->out_len = pos;
  res->adler_computed = 0;
  res->adler_stored = ((uint32_t)in[tb] << 24) | ((uint32_t)in[tb + 1u] << 16) | ((uint32_t)in[tb + 2u] << 8) | in[tb + 3u];
  res->err_bitpos = (uint64_t)(tb + 4u) * 8u;
  res->payload[0] = 0;
  /* bytes the reference has published when it reaches the trailer: the window's base after the last moveWindow call, in
   * closed form because no gap exceeded 32 KiB (PzCtx::mark) */
  res->payload[1] = pos >= 2u * PZ_EXCESS ? (int64_t)((pos / PZ_EXCESS - 1u) * PZ_EXCESS) : 0;
#ifndef PZ_HOSTSIM
  __threadfence();
#endif
  res->status = PZ_OK;
  return true;
}

#ifndef PZ_HOSTSIM
/* the sizing pass has no K2 in front of it to mark the streams */
__global__ void __launch_bounds__(256)
pz_mark_pending_kernel(const PzJob job) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < job.count) job.res[job.first + k].status = PZ_ST_PENDING;
}

/* The streams a kernel of this file will try, packed: list[0] = how many, list[2 + i] = stream index.  Thread k of the decode
 * kernel takes list[2 + k], so its warps are full whatever the mix of the batch is (in BASELINE configs[2] every fourth record
 * is a dynamic one: with "thread k takes stream k", K6's warps would run with eight lanes).  want_dynamic = 0: streams whose
 * first block is a fixed one (K5); 1: whatever is still pending and small enough (K6). */
__global__ void __launch_bounds__(256)
pz_small_list_kernel(const PzJob job, uint32_t *list, uint32_t want_dynamic) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  bool take = false;
  uint32_t s = 0;
  if (k < job.count) {
    s = job.first + k;
    const uint64_t i0 = job.in_off[s], n = job.in_off[s + 1] - i0;
    if (job.res[s].status == PZ_ST_PENDING && n >= 8u && n <= PZ_FIXED_MAX_IN) {
      const uint8_t *in = job.in_blob + i0;
      const uint32_t p = (in[1] & 0x20u) ? 6u : 2u;
      const uint32_t bt = (in[p] >> 1) & 3u;
      take = bt == 1u || (want_dynamic != 0u && bt == 2u);
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, take);
  if (m == 0u) return;
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t base = 0;
  if (lane == (uint32_t)__ffs((int)m) - 1u) base = atomicAdd(list, (uint32_t)__popc(m));
  base = __shfl_sync(0xffffffffu, base, __ffs((int)m) - 1);
  if (take) list[2u + base + (uint32_t)__popc(m & ((1u << lane) - 1u))] = s;
}

/* DYN = false: K5 (fixed-Huffman streams, no tables, full occupancy); DYN = true: K6 (dynamic blocks too, 3.5 KiB of tables per
 * thread in local memory).  list: see pz_small_list_kernel. */
template <bool COUNT_ONLY, bool DYN = false>
__global__ void __launch_bounds__(PZ_FIXED_THREADS)
pz_fixed_kernel(const PzJob job, const uint32_t *__restrict__ list) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= list[0]) return;
  const uint32_t s = list[2u + k];
  pz_result *res = job.res + s;
  if (res->status != PZ_ST_PENDING) return;
  const uint64_t i0 = job.in_off[s], n = job.in_off[s + 1] - i0;
  uint8_t *out = nullptr;
  uint64_t cap = 0;
  if (!COUNT_ONLY) {
    const uint64_t o0 = job.out_off[s];
    cap = job.out_off[s + 1] - o0;
    out = job.out_blob + o0;
  }
  (void)pz_fixed_stream<COUNT_ONLY, DYN>(job.in_blob + i0, n, out, cap, res);
}
#endif
