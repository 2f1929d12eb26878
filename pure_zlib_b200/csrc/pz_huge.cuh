/*
 * pz_huge.cuh -- K4: the kernels around the block-parallel decode of ONE huge zlib stream.
 *
 * A single stream is one serial dependency chain for the symbol loop (Deflate.hs:106-120) and, through
 * the 32 KiB history, for the output too (OutputWindow.hs:82-101).  K4 cuts both chains at deflate
 * block boundaries:
 *
 *   K4a  pz_blk_search_kernel   every bit position of the stream is tested for "a dynamic block
 *        pz_blk_verify_kernel   header could start here" (BTYPE, HLIT/HDIST range, a complete
 *                               code-length code; then the full header: complete literal/length
 *                               and distance codes, end-of-block present).  Speculation: a
 *                               position that passes is only a CANDIDATE.
 *   K1   (pz_inflate_kernel in block-job mode, pz_device.cuh) decodes every candidate from its
 *        header to its end-of-block symbol -- first a sizing pass, from which the host builds
 *        the chain "block k ends exactly where candidate k+1 starts" beginning at the stream's
 *        first block (that is what makes the speculation self-synchronising: candidates that are
 *        not on the chain are simply never used), then a pass that writes 16-bit symbols: a
 *        byte, or a marker 256 + i for "byte i of the 32 KiB before this block".
 *   K4c  pz_blk_tails_kernel    LZ77 resolution, serial part, two levels: only the last 32 KiB of a
 *        pz_blk_windows_kernel  block can be referenced later, so only those "tails" are walked in
 *                               chain order -- group by group in parallel against an unknown
 *                               window, then one short walk over the groups.
 *   K4d  pz_blk_resolve_kernel  LZ77 resolution, parallel part: every symbol looked up on its own.
 *
 * Anything unusual -- a block the search cannot see, a verdict other than success anywhere, a
 * reference before the start of the stream, literal runs long enough to matter to the
 * reference's window model -- makes the host driver drop K4 for that stream and decode it with
 * the ordinary serial path, which reproduces the reference's verdicts exactly.
 */
#pragma once
#include <stdint.h>

#define PZ_HUGE_THREADS 256

/* the 96 stream bits starting at byte B (bits past the end read as zero) */
__device__ __forceinline__ void pz_load96(const uint8_t *in, uint64_t nbytes, uint64_t B, uint32_t &v0, uint32_t &v1, uint32_t &v2) {
  const uintptr_t addr = (uintptr_t)(in + B);
  const uint32_t *w = (const uint32_t *)(addr & ~(uintptr_t)3);
  const uint32_t sh = (uint32_t)(addr & 3u) * 8u;
  const uint8_t *end = in + nbytes;
  uint32_t x[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    /* a whole word is read only if it lies inside [in, end): the edges are assembled bytewise */
    const uint8_t *p = (const uint8_t *)(w + k);
    if (p >= in && p + 4 <= end) {
      x[k] = w[k];
    } else {
      x[k] = 0;
#pragma unroll
      for (int b = 0; b < 4; b++)
        if (p + b >= in && p + b < end) x[k] |= (uint32_t)p[b] << (8 * b);
    }
  }
  v0 = __funnelshift_r(x[0], x[1], sh);
  v1 = __funnelshift_r(x[1], x[2], sh);
  v2 = __funnelshift_r(x[2], x[3], sh);
}

/* K4a, first stage: one thread per byte of the stream, eight bit positions each.
 *
 * Two steps per warp, so that the expensive test runs on full warps: (1) every lane applies the cheap
 * tests (BTYPE, HLIT/HDIST range, room for the code-length code) to its eight positions -- 22 % of
 * all positions pass -- and the survivors are packed into a per-warp list in shared memory (the 96
 * stream bits of the byte and the bit offset); (2) the lanes walk that list together, one survivor
 * each per round, and test the code-length code for completeness (sum 2^(7-len) == 128).  Doing
 * (2) inside the loop over the eight offsets kept three lanes in four idle in every round. */
__global__ void __launch_bounds__(PZ_HUGE_THREADS)
pz_blk_search_kernel(const uint8_t *__restrict__ in, uint64_t nbytes, uint64_t first_bit, uint64_t last_bit,
                     uint32_t *__restrict__ cand, uint32_t *__restrict__ ncand, uint32_t cap) {
  __shared__ uint4 list[PZ_HUGE_THREADS / 32][256];
  __shared__ uint8_t pair[64]; /* Kraft weight (in 128ths) of two 3-bit code lengths at once */
  if (threadIdx.x < 64u) {
    const uint32_t a = threadIdx.x & 7u, b = threadIdx.x >> 3;
    pair[threadIdx.x] = (uint8_t)((a ? 128u >> a : 0u) + (b ? 128u >> b : 0u));
  }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint64_t B = (first_bit >> 3) + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t v0 = 0, v1 = 0, v2 = 0, mask = 0;
  if (B * 8u < last_bit) {
    pz_load96(in, nbytes, B, v0, v1, v2);
    /* the cheap tests on all eight offsets at once, one bit per offset: BTYPE (bits o+1, o+2) is dynamic (Deflate.hs:83),
     * HLIT (bits o+3..o+7) and HDIST (bits o+8..o+12) are at most 29 -- what zlib can emit -- i.e. their upper four bits
     * are not all ones.  Positions outside [first_bit, last_bit) are dropped in the second step. */
    const uint32_t dyn = ~(v0 >> 1) & (v0 >> 2);
    const uint32_t hlit_bad = (v0 >> 4) & (v0 >> 5) & (v0 >> 6) & (v0 >> 7);
    const uint32_t hdist_bad = (v0 >> 9) & (v0 >> 10) & (v0 >> 11) & (v0 >> 12);
    mask = dyn & ~hlit_bad & ~hdist_bad & 0xffu;
    /* an EMPTY stored block (00 00 ff ff behind the padding: what Z_SYNC_FLUSH / Z_FULL_FLUSH leave between the
     * pieces of a stream) is a candidate on the strength of those 32 bits alone */
    const uint32_t lw1 = (v0 >> 8) | (v1 << 24), lw2 = (v0 >> 16) | (v1 << 16); /* LEN | NLEN << 16 one / two bytes on */
    if (lw1 == 0xffff0000u || lw2 == 0xffff0000u) {
      for (uint32_t o = 0; o < 8u; o++) {
        const uint64_t pos = B * 8u + o;
        const uint32_t al = (o + 3u + 7u) >> 3; /* bytes from B to the byte boundary behind the three header bits */
        if (pos >= first_bit && pos < last_bit && ((v0 >> (o + 1u)) & 3u) == 0u && (al == 1u ? lw1 : lw2) == 0xffff0000u &&
            (B + al + 4u) * 8u <= last_bit) {
          const uint32_t k = atomicAdd(ncand, 1u);
          if (k < cap) cand[k] = (uint32_t)pos;
        }
      }
    }
  }
  /* pack the survivors of the warp */
  const uint32_t mine = (uint32_t)__popc(mask);
  uint32_t incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if ((int)lane >= d) incl += t;
  }
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  uint32_t at = incl - mine;
  while (mask) {
    const uint32_t o = (uint32_t)__ffs((int)mask) - 1u;
    mask &= mask - 1u;
    list[warp][at++] = make_uint4(v0, v1, v2, (uint32_t)(B * 8u) + o); /* bit positions fit 32 bits: streams are below 512 MiB */
  }
  __syncwarp();
  for (uint32_t i = lane; i < total; i += 32u) {
    const uint4 e = list[warp][i];
    const uint32_t o = e.w & 7u;
    const uint32_t h = __funnelshift_r(e.x, e.y, o);
    const uint32_t hclen = ((h >> 13) & 15u) + 4u;
    if (e.w < first_bit || (uint64_t)e.w + 17u + 3u * hclen > last_bit) continue; /* the stream's edges */
    /* the code-length code must be complete: sum 2^(7-len) == 128 */
    const uint32_t s = o + 17u; /* < 32 */
    uint32_t p0 = __funnelshift_r(e.x, e.y, s), p1 = __funnelshift_r(e.y, e.z, s);
    uint32_t sum = 0;
    for (uint32_t k = 0; k + 1u < hclen && sum <= 128u; k += 2u) { /* two lengths per step */
      sum += pair[p0 & 63u];
      p0 = __funnelshift_r(p0, p1, 6);
      p1 >>= 6;
    }
    if (hclen & 1u) sum += pair[p0 & 7u]; /* the odd one out (pair[l] = weight of l alone) */
    if (sum != 128u) continue;
    const uint32_t k = atomicAdd(ncand, 1u);
    if (k < cap) cand[k] = e.w;
  }
}

/* n <= 25 stream bits starting at bit position pos (bits past the end read as zero) */
__device__ __forceinline__ uint32_t pz_bits_at(const uint8_t *in, uint64_t nbytes, uint64_t pos, uint32_t n) {
  const uint64_t b = pos >> 3;
  uint32_t v = 0;
#pragma unroll
  for (int k = 0; k < 4; k++)
    if (b + k < nbytes) v |= (uint32_t)in[b + k] << (8 * k);
  return (v >> (pos & 7u)) & ((1u << n) - 1u);
}

/* The same with two aligned 32-bit loads instead of four guarded byte loads, for positions at least four bytes before
 * the end of the stream (the blob has 15 readable bytes of slack around it, pzcuda.h). */
__device__ __forceinline__ uint32_t pz_bits_fast(const uint8_t *in, uint64_t pos, uint32_t n) {
  const uintptr_t a = (uintptr_t)(in + (pos >> 3));
  const uint32_t *w = (const uint32_t *)(a & ~(uintptr_t)3);
  return __funnelshift_r(w[0], w[1], (uint32_t)(a & 3u) * 8u + (uint32_t)(pos & 7u)) & ((1u << n) - 1u);
}

/* K4a, second stage: one thread per first-stage candidate reads the whole dynamic header
 * (Deflate.hs:83-101,124-156) and keeps the candidate only if it is what zlib's tree builder
 * produces: complete code-length, literal/length and distance codes, no repeat that starts the
 * list or runs past it, an end-of-block code. */
__global__ void __launch_bounds__(PZ_HUGE_THREADS)
pz_blk_verify_kernel(const uint8_t *__restrict__ in, uint64_t nbytes, uint64_t last_bit, const uint32_t *__restrict__ cand,
                     uint32_t ncand, uint32_t *__restrict__ kept, uint32_t *__restrict__ nkept) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncand) return;
  uint64_t pos = cand[i];
  const uint32_t h = pz_bits_at(in, nbytes, pos, 17);
  if (((h >> 1) & 3u) == 0u) { /* the first stage only lets EMPTY stored blocks through: check LEN / NLEN again */
    const uint64_t b = (pos + 3u + 7u) >> 3;
    if (b + 4u <= nbytes && pz_bits_at(in, nbytes, b * 8u, 16) == 0u && pz_bits_at(in, nbytes, b * 8u + 16u, 16) == 0xffffu)
      kept[atomicAdd(nkept, 1u)] = cand[i];
    return;
  }
  const uint32_t hlit = ((h >> 3) & 31u) + 257u, hdist = ((h >> 8) & 31u) + 1u, hclen = ((h >> 13) & 15u) + 4u;
  pos += 17u;
  /* code-length code lengths, 3 bits per symbol, packed by symbol */
  const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  uint64_t pl = 0;
  for (uint32_t k = 0; k < hclen; k++) {
    pl |= (uint64_t)pz_bits_at(in, nbytes, pos, 3) << (3u * order[k]);
    pos += 3u;
  }
  /* The code-length code as a 7-bit lookup table (index = the next seven stream bits, first bit of a code = its most
   * significant one, HuffmanTree.hs:73-83), one table per thread in shared memory, entry e of thread t at e * 256 + t.
   * The first stage let this candidate through because the code is complete, so every entry gets written.  Almost every
   * candidate -- true or not -- is parsed to the end of its ~300 code lengths (random lengths of 8 or 9 bits take that
   * long to over-subscribe a code), so the table pays for itself many times: the bit-by-bit canonical decode it replaces
   * cost ~150 instructions per code length. */
  __shared__ uint8_t pz_pre_lut[128 * PZ_HUGE_THREADS];
  uint8_t *lut = pz_pre_lut + threadIdx.x;
  {
    uint32_t code = 0;
#pragma unroll 1
    for (uint32_t l = 1; l <= 7u; l++) {
      for (uint32_t s = 0; s < 19u; s++) {
        if (((pl >> (3u * s)) & 7u) != l) continue;
        for (uint32_t j = __brev(code) >> (32u - l); j < 128u; j += 1u << l) lut[j * PZ_HUGE_THREADS] = (uint8_t)(s | (l << 5));
        code++;
      }
      code <<= 1;
    }
  }
  const uint32_t total = hlit + hdist;
  uint32_t n = 0, prev = 0, sum_l = 0, sum_d = 0;
  bool eob = false;
  while (n < total) {
    if (pos + 7u + 7u > last_bit) return;
    const uint32_t e = lut[pz_bits_fast(in, pos, 7) * PZ_HUGE_THREADS]; /* pos + 14 <= last_bit: 32 bits before the end */
    const uint32_t sym = e & 31u;
    pos += e >> 5;
    uint32_t rep = 1, val = sym;
    if (sym == 16u) {
      if (n == 0u) return;
      rep = 3u + pz_bits_fast(in, pos, 2); pos += 2u; val = prev;
    } else if (sym == 17u) {
      rep = 3u + pz_bits_fast(in, pos, 3); pos += 3u; val = 0;
    } else if (sym == 18u) {
      rep = 11u + pz_bits_fast(in, pos, 7); pos += 7u; val = 0;
    }
    if (n + rep > total) return;
    prev = val;
    if (val) {
      for (uint32_t k = 0; k < rep; k++) {
        if (n + k < hlit) { sum_l += 32768u >> val; if (n + k == 256u) eob = true; }
        else sum_d += 32768u >> val;
      }
    }
    n += rep;
    if (sum_l > 32768u || sum_d > 32768u) return; /* over-subscribed already: most false candidates end here, a few symbols in */
  }
  /* the survivors (one position in 250 000) go into a list of their own, in no particular order: the host sorts it */
  if (sum_l == 32768u && sum_d == 32768u && eob) kept[atomicAdd(nkept, 1u)] = cand[i];
}

/* One-pass flow: the chain's blocks were decoded into scratch regions of their own (every candidate was, before the
 * chain was known); symbol e of the stream, which lies in chain block k, is scr[blk_src[k] + (e - blk_off[k])].
 * Moves them to their final positions so that the resolution kernels below see one contiguous stream of symbols.
 * Thread t of a CTA takes symbols c0 + t, c0 + t + THREADS, ...: consecutive lanes, consecutive symbols. */
__global__ void __launch_bounds__(PZ_HUGE_THREADS)
pz_blk_compact_kernel(const uint16_t *__restrict__ scr, uint16_t *__restrict__ sym, const uint64_t *__restrict__ blk_off,
                      const uint64_t *__restrict__ blk_src, uint32_t nblk, uint64_t total) {
  __shared__ uint32_t k0s;
  const uint64_t c0 = (uint64_t)blockIdx.x * (PZ_HUGE_THREADS * 8u);
  if (threadIdx.x == 0) { /* the block holding this CTA's first symbol: last k with blk_off[k] <= c0 */
    uint32_t lo = 0, hi = nblk;
    while (hi - lo > 1u) {
      const uint32_t mid = lo + (hi - lo) / 2u;
      if (blk_off[mid] <= c0) lo = mid; else hi = mid;
    }
    k0s = lo;
  }
  __syncthreads();
  uint32_t k = k0s;
  uint16_t v[8];
  uint64_t next = 0; /* first symbol behind block k; 0 = not looked up yet */
  const uint16_t *from = scr; /* scr + blk_src[k] - blk_off[k] */
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const uint64_t e = c0 + (uint64_t)j * PZ_HUGE_THREADS + threadIdx.x;
    v[j] = 0;
    if (e < total) {
      if (e >= next) {
        while (k + 1u < nblk && blk_off[k + 1u] <= e) k++;
        next = k + 1u < nblk ? blk_off[k + 1u] : total;
        from = scr + blk_src[k] - blk_off[k];
      }
      v[j] = from[e];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const uint64_t e = c0 + (uint64_t)j * PZ_HUGE_THREADS + threadIdx.x;
    if (e < total) sym[e] = v[j];
  }
}

/* K4c: the serial part of the LZ77 resolution, in two levels.
 *
 * Only the last 32 KiB of a block (its "tail") can be referenced by later blocks, so only tails
 * have to be resolved in chain order.  The chain is cut into G groups of consecutive blocks:
 *
 *   pz_blk_tails_kernel (one CTA per group, all groups at once) walks the tails of its group with a
 *     window of 16-bit symbols that starts out as "unknown": W[a mod 32768] = marker for byte a of
 *     the 32 KiB before the group.  Afterwards every tail symbol is a byte or a marker into the
 *     GROUP's initial window (written back in place).
 *   pz_blk_windows_kernel (one CTA) then walks the groups in order with a window of real bytes:
 *     group g's last 32 KiB, resolved against the window at its start, are the window at the
 *     start of group g+1.  The G windows are kept in gw[].
 *
 * After that every symbol of the stream can be resolved on its own (K4d). */
#define PZ_TAIL 32768u
#define PZ_TAILS_THREADS 1024
__global__ void __launch_bounds__(PZ_TAILS_THREADS, 1)
pz_blk_tails_kernel(uint16_t *sym, const uint64_t *__restrict__ blk_off, const uint32_t *__restrict__ blk_len,
                    const uint32_t *__restrict__ grp_first, uint32_t ngrp) {
  extern __shared__ uint16_t W[]; /* PZ_TAIL symbols */
  const uint32_t tid = threadIdx.x;
  const uint32_t g = blockIdx.x;
  const uint32_t k0 = grp_first[g], k1 = grp_first[g + 1];
  const uint64_t goff = blk_off[k0];
  /* byte a of [goff - 32768, goff) lives in slot a mod 32768; its marker index is a - (goff - 32768) */
  for (uint32_t i = tid; i < PZ_TAIL; i += PZ_TAILS_THREADS) {
    const uint64_t a = goff - PZ_TAIL + i; /* may wrap below zero: only the low 15 bits are used */
    W[(uint32_t)a & (PZ_TAIL - 1u)] = (uint16_t)(256u + i);
  }
  __syncthreads();
  for (uint32_t k = k0; k < k1; k++) {
    const uint64_t off = blk_off[k];
    const uint32_t len = blk_len[k], T = len < PZ_TAIL ? len : PZ_TAIL;
    const uint64_t t0 = off + (len - T);
    uint16_t v[PZ_TAIL / PZ_TAILS_THREADS];
#pragma unroll
    for (int j = 0; j < (int)(PZ_TAIL / PZ_TAILS_THREADS); j++) {
      const uint32_t i = tid + PZ_TAILS_THREADS * (uint32_t)j;
      v[j] = i < T ? sym[t0 + i] : (uint16_t)0;
    }
    bool ch[PZ_TAIL / PZ_TAILS_THREADS];
#pragma unroll
    for (int j = 0; j < (int)(PZ_TAIL / PZ_TAILS_THREADS); j++) {
      ch[j] = v[j] >= 256u;
      if (ch[j]) { /* byte (v - 256) of the 32 KiB before THIS block: absolute position off - 32768 + (v - 256) */
        const uint64_t a = off - PZ_TAIL + (uint32_t)(v[j] - 256u);
        v[j] = W[(uint32_t)a & (PZ_TAIL - 1u)];
      }
    }
    __syncthreads(); /* every read of the old window is done */
#pragma unroll
    for (int j = 0; j < (int)(PZ_TAIL / PZ_TAILS_THREADS); j++) {
      const uint32_t i = tid + PZ_TAILS_THREADS * (uint32_t)j;
      if (i < T) {
        W[(uint32_t)(t0 + i) & (PZ_TAIL - 1u)] = v[j];
        if (ch[j]) sym[t0 + i] = v[j];
      }
    }
    __syncthreads();
  }
}

/* gw[g * 32768 + i] = final byte a = goff_g - 32768 + i (slot order is by marker index, not by
 * a mod 32768); err is raised for a marker that points before the first byte of the stream. */
__global__ void __launch_bounds__(PZ_TAILS_THREADS, 1)
pz_blk_windows_kernel(const uint16_t *__restrict__ sym, const uint64_t *__restrict__ blk_off, const uint32_t *__restrict__ grp_first,
                      uint32_t ngrp, uint64_t total, uint8_t *__restrict__ gw, uint32_t *__restrict__ err) {
  __shared__ uint8_t W[PZ_TAIL]; /* final bytes before the current group, slot = a mod 32768 */
  const uint32_t tid = threadIdx.x;
  bool bad = false;
  for (uint32_t g = 0; g < ngrp; g++) {
    const uint64_t goff = blk_off[grp_first[g]];
    const uint64_t gend = g + 1u < ngrp ? blk_off[grp_first[g + 1u]] : total;
    /* publish the window at the start of group g */
    for (uint32_t i = tid; i < PZ_TAIL; i += PZ_TAILS_THREADS) {
      const uint64_t a = goff - PZ_TAIL + i;
      gw[(uint64_t)g * PZ_TAIL + i] = W[(uint32_t)a & (PZ_TAIL - 1u)];
    }
    /* then advance it over the group's last min(32768, length) bytes */
    const uint64_t glen = gend - goff;
    const uint32_t T = glen < PZ_TAIL ? (uint32_t)glen : PZ_TAIL;
    const uint64_t t0 = gend - T;
    uint8_t b[PZ_TAIL / PZ_TAILS_THREADS];
#pragma unroll
    for (int j = 0; j < (int)(PZ_TAIL / PZ_TAILS_THREADS); j++) {
      const uint32_t i = tid + PZ_TAILS_THREADS * (uint32_t)j;
      const uint32_t s = i < T ? sym[t0 + i] : 0u;
      if (s >= 256u) {
        const uint64_t a = goff - PZ_TAIL + (s - 256u);
        if (goff + (s - 256u) < PZ_TAIL) bad = true; /* a < 0 */
        b[j] = W[(uint32_t)a & (PZ_TAIL - 1u)];
      } else {
        b[j] = (uint8_t)s;
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < (int)(PZ_TAIL / PZ_TAILS_THREADS); j++) {
      const uint32_t i = tid + PZ_TAILS_THREADS * (uint32_t)j;
      if (i < T) W[(uint32_t)(t0 + i) & (PZ_TAIL - 1u)] = b[j];
    }
    __syncthreads();
  }
  if (bad) *err = 1u;
}

/* K4d: every symbol on its own.  Thread t handles symbols [8t, 8t+8).
 *   tail symbol      byte, or marker into its group's window gw[g]
 *   any other symbol byte, or marker for a byte before its block -- which is a tail symbol of an
 *                    earlier block: look that one up the same way */
__device__ __forceinline__ uint32_t pz_resolve_tail_sym(uint32_t s, uint32_t g, const uint8_t *__restrict__ gw, const uint64_t *__restrict__ blk_off,
                                                        const uint32_t *__restrict__ grp_first, bool &bad) {
  if (s < 256u) return s;
  if (blk_off[grp_first[g]] + (s - 256u) < PZ_TAIL) { bad = true; return 0u; } /* before the first byte of the stream */
  return gw[(uint64_t)g * PZ_TAIL + (s - 256u)];
}
__global__ void __launch_bounds__(PZ_HUGE_THREADS)
pz_blk_resolve_kernel(const uint16_t *__restrict__ sym, uint8_t *__restrict__ out, const uint64_t *__restrict__ blk_off,
                      const uint32_t *__restrict__ blk_len, const uint32_t *__restrict__ blk_grp, const uint32_t *__restrict__ grp_first,
                      uint32_t nblk, uint64_t total, const uint8_t *__restrict__ gw, uint32_t *__restrict__ err) {
  __shared__ uint32_t k0s;
  const uint64_t c0 = (uint64_t)blockIdx.x * (PZ_HUGE_THREADS * 8u);
  if (threadIdx.x == 0) { /* the block holding this CTA's first symbol: last k with blk_off[k] <= c0 */
    uint32_t lo = 0, hi = nblk;
    while (hi - lo > 1u) {
      const uint32_t mid = lo + (hi - lo) / 2u;
      if (blk_off[mid] <= c0) lo = mid; else hi = mid;
    }
    k0s = lo;
  }
  __syncthreads();
  uint32_t k = k0s;
  const uint64_t e0 = c0 + (uint64_t)threadIdx.x * 8u;
  if (e0 >= total) return;
  bool bad = false;
  uint32_t r[2] = {0u, 0u};
  uint32_t w[4] = {0u, 0u, 0u, 0u};
  const bool vec = e0 + 8u <= total && (((uintptr_t)(sym + e0)) & 15u) == 0u;
  if (vec) {
    const uint4 v = *reinterpret_cast<const uint4 *>(sym + e0);
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
  }
  /* the block holding e0 (the last k with blk_off[k] <= e0: empty blocks are skipped), kept in registers: eight
   * consecutive symbols rarely leave it */
  while (k + 1u < nblk && blk_off[k + 1u] <= e0) k++;
  uint64_t off = blk_off[k];
  uint64_t next = k + 1u < nblk ? blk_off[k + 1u] : total;
  uint32_t len = blk_len[k];
  uint64_t tail0 = off + (len - (len < PZ_TAIL ? len : PZ_TAIL)); /* first tail symbol of the block */
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const uint64_t e = e0 + (uint32_t)j;
    if (e >= total) break;
    if (e >= next) {
      do k++; while (k + 1u < nblk && blk_off[k + 1u] <= e);
      off = blk_off[k];
      next = k + 1u < nblk ? blk_off[k + 1u] : total;
      len = blk_len[k];
      tail0 = off + (len - (len < PZ_TAIL ? len : PZ_TAIL));
    }
    uint32_t s = vec ? (w[j >> 1] >> (16 * (j & 1))) & 0xffffu : (uint32_t)sym[e];
    if (e >= tail0) { /* a tail symbol */
      s = pz_resolve_tail_sym(s, blk_grp[k], gw, blk_off, grp_first, bad);
    } else if (s >= 256u) {
      const uint64_t a = off - PZ_TAIL + (s - 256u);
      if (off + (s - 256u) < PZ_TAIL) { bad = true; s = 0; }
      else {
        uint32_t ka = k; /* the block holding byte a: an earlier one */
        while (blk_off[ka] > a) ka--;
        s = pz_resolve_tail_sym(sym[a], blk_grp[ka], gw, blk_off, grp_first, bad);
      }
    }
    r[j >> 2] |= (s & 0xffu) << (8 * (j & 3));
  }
  if (e0 + 8u <= total && (((uintptr_t)(out + e0)) & 7u) == 0u) {
    *reinterpret_cast<uint2 *>(out + e0) = make_uint2(r[0], r[1]);
  } else {
    for (uint32_t j = 0; j < 8u && e0 + j < total; j++) out[e0 + j] = (uint8_t)(r[j >> 2] >> (8 * (j & 3)));
  }
  if (bad) *err = 1u;
}
