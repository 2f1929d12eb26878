/*
 * pz_stored.cuh -- K2: streams that consist of stored blocks only (BTYPE 0 throughout).
 *
 * Incompressible data leaves zlib as chains of ~16 KiB stored blocks (inflateBlock's first arm,
 * Deflate.hs:70-78).  Such a stream has no Huffman work at all: its cost is a serial walk over
 * the 5-byte block headers plus a raw copy, and the copy deserves the whole memory system
 * instead of one 8-lane group of K1.  Two kernels run before K1:
 *
 *   pz_stored_probe_kernel  one thread per stream: zlib header (Zlib.hs:53-69) and the first
 *                           block's type.  Marks the stream CANDIDATE or PENDING in res[].status.
 *   pz_stored_copy_kernel   one CTA per candidate: lane 0 of warp 0 walks the header chain
 *                           (LEN/NLEN, the reference's window model and truncation rule) and
 *                           publishes (src, dst, len) descriptors in shared memory; the other
 *                           seven warps copy them with 16-byte accesses.  Only a stream that
 *                           validates completely gets its verdict here; anything else -- a
 *                           Huffman block further on, a malformed header, a truncation, a window
 *                           or capacity problem -- is left PENDING and K1 decodes it from scratch,
 *                           so every verdict other than plain success still comes from the one
 *                           place that reproduces the reference's order of checks.
 *
 * K1 skips streams whose status is not PENDING.  Adler-32 is checked afterwards by K3 as for any
 * other stream.
 */
#pragma once
#include <stdint.h>

#include "pz_device.cuh"

/* PZ_ST_PENDING (-1, pz_device.cuh): K1 has to decode this stream */
#define PZ_ST_CANDIDATE (-2) /* starts with a stored block: K2 tries it */
#define PZ_ST_THREADS 256
#define PZ_ST_RING 63 /* descriptor ring: 7 copy warps x 9 */

__global__ void __launch_bounds__(256)
pz_stored_probe_kernel(const PzJob job) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= job.count) return;
  const uint32_t s = job.first + k;
  const uint64_t i0 = job.in_off[s], n = job.in_off[s + 1] - i0;
  const uint8_t *in = job.in_blob + i0;
  int32_t st = PZ_ST_PENDING;
  if (n >= 11u && n <= PZ_MAX_IN_BYTES) { /* header, one block header, trailer */
    const uint32_t cmf = in[0], flg = in[1];
    if (((cmf << 8) | flg) % 31u == 0u && (cmf & 15u) == 8u && (cmf >> 4) <= 7u) {
      const uint32_t p = (flg & 0x20u) ? 6u : 2u;
      if (p < n && ((in[p] >> 1) & 3u) == 0u) st = PZ_ST_CANDIDATE;
    }
  }
  job.res[s].status = st;
}

struct PzStoredShared {
  uint32_t src[PZ_ST_RING], dst[PZ_ST_RING], len[PZ_ST_RING];
  uint32_t head;    /* descriptors published */
  uint32_t ended;   /* 1 once the walker has stopped */
  uint32_t ok;      /* 1 if the whole stream validated */
  uint32_t done[7]; /* per copy warp: descriptors finished */
  uint32_t cand[PZ_ST_THREADS];
  uint32_t ncand;
  /* verdict fields of a validated stream */
  uint32_t out_len, adler_stored, end_byte, base;
};

/* 16 bytes from the unaligned address p: two aligned 16-byte loads and a funnel shift. */
__device__ __forceinline__ uint4 pz_load16_unaligned(const uint8_t *p) {
  const uint32_t o = (uint32_t)((uintptr_t)p & 15u);
  const uint4 *a = reinterpret_cast<const uint4 *>(p - o);
  const uint4 A = a[0];
  if (o == 0u) return A;
  const uint4 B = a[1];
  const uint32_t sh = (o & 3u) * 8u;
  uint32_t w0, w1, w2, w3, w4;
  switch (o >> 2) { /* uniform across the warp: every lane is a multiple of 16 bytes apart */
    case 0: w0 = A.x; w1 = A.y; w2 = A.z; w3 = A.w; w4 = B.x; break;
    case 1: w0 = A.y; w1 = A.z; w2 = A.w; w3 = B.x; w4 = B.y; break;
    case 2: w0 = A.z; w1 = A.w; w2 = B.x; w3 = B.y; w4 = B.z; break;
    default: w0 = A.w; w1 = B.x; w2 = B.y; w3 = B.z; w4 = B.w; break;
  }
  return make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
}

/* One warp copies len bytes; src and dst have arbitrary alignment.  With SUM the warp also adds up
 * what it copies: s1 += sum of the bytes, t += sum of byte * (pos0 + index), per lane (pos0 + len <=
 * 16384, so a lane's share -- 1/32 of at most 16 KiB -- keeps t below 2^32). */
__device__ __forceinline__ void pz_sum16(const uint4 w, uint32_t pos, uint32_t &s1, uint32_t &t) {
  uint32_t sum = __dp4a(w.x, 0x01010101u, 0u);
  sum = __dp4a(w.y, 0x01010101u, sum);
  sum = __dp4a(w.z, 0x01010101u, sum);
  sum = __dp4a(w.w, 0x01010101u, sum);
  uint32_t ks = __dp4a(w.x, 0x03020100u, 0u);
  ks = __dp4a(w.y, 0x07060504u, ks);
  ks = __dp4a(w.z, 0x0b0a0908u, ks);
  ks = __dp4a(w.w, 0x0f0e0d0cu, ks);
  s1 += sum;
  t += pos * sum + ks;
}
/* The 16 bytes at byte offset 4 * WO + sh / 8 of the 32 bytes (A, B): WO and sh are the same for every
 * vector of a copy (source and destination keep their relative alignment), so the loop below is
 * compiled once per WO and contains no branch between its loads. */
template <int WO>
__device__ __forceinline__ uint4 pz_shift16(const uint4 A, const uint4 B, const uint32_t sh) {
  const uint32_t w[8] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w};
  return make_uint4(__funnelshift_r(w[WO], w[WO + 1], sh), __funnelshift_r(w[WO + 1], w[WO + 2], sh),
                    __funnelshift_r(w[WO + 2], w[WO + 3], sh), __funnelshift_r(w[WO + 3], w[WO + 4], sh));
}
#ifndef PZ_COPY_DEPTH
#define PZ_COPY_DEPTH 4
#endif
/* PZ_COPY_DEPTH: 16-byte pieces per lane whose loads are all issued before the first is used */
template <int WO, bool ALIGNED, bool SUM>
__device__ __forceinline__ void pz_copy_vectors(uint4 *d16, const uint4 *a, const uint32_t n16, const uint32_t sh, const uint32_t pos0,
                                                uint32_t &s1, uint32_t &t) {
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t c = lane;
  for (; c + 32u * (PZ_COPY_DEPTH - 1) < n16; c += 32u * PZ_COPY_DEPTH) {
    uint4 A[PZ_COPY_DEPTH], B[PZ_COPY_DEPTH];
#pragma unroll
    for (int k = 0; k < PZ_COPY_DEPTH; k++) {
      A[k] = a[c + 32u * k];
      if (!ALIGNED) B[k] = a[c + 32u * k + 1u];
    }
#pragma unroll
    for (int k = 0; k < PZ_COPY_DEPTH; k++) {
      const uint4 v = ALIGNED ? A[k] : pz_shift16<WO>(A[k], B[k], sh);
      d16[c + 32u * k] = v;
      if (SUM) pz_sum16(v, pos0 + 16u * (c + 32u * k), s1, t);
    }
  }
  for (; c < n16; c += 32u) {
    const uint4 A = a[c];
    const uint4 v = ALIGNED ? A : pz_shift16<WO>(A, a[c + 1u], sh);
    d16[c] = v;
    if (SUM) pz_sum16(v, pos0 + 16u * c, s1, t);
  }
}
template <bool SUM>
__device__ __forceinline__ void pz_warp_copy(uint8_t *dst, const uint8_t *src, uint32_t len, uint32_t pos0, uint32_t &s1, uint32_t &t) {
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t head = (uint32_t)(-(intptr_t)dst) & 15u;
  if (head > len) head = len;
  if (lane < head) {
    const uint32_t b = src[lane];
    dst[lane] = (uint8_t)b;
    if (SUM) { s1 += b; t += b * (pos0 + lane); }
  }
  dst += head; src += head; len -= head; pos0 += head;
  const uint32_t n16 = len >> 4;
  uint4 *d16 = reinterpret_cast<uint4 *>(dst);
  const uint32_t o = (uint32_t)((uintptr_t)src & 15u), sh = (o & 3u) * 8u;
  const uint4 *a = reinterpret_cast<const uint4 *>(src - o);
  if (o == 0u) pz_copy_vectors<0, true, SUM>(d16, a, n16, 0u, pos0, s1, t);
  else switch (o >> 2) { /* uniform across the warp */
    case 0: pz_copy_vectors<0, false, SUM>(d16, a, n16, sh, pos0, s1, t); break;
    case 1: pz_copy_vectors<1, false, SUM>(d16, a, n16, sh, pos0, s1, t); break;
    case 2: pz_copy_vectors<2, false, SUM>(d16, a, n16, sh, pos0, s1, t); break;
    default: pz_copy_vectors<3, false, SUM>(d16, a, n16, sh, pos0, s1, t); break;
  }
  const uint32_t tail = len & 15u;
  if (lane < tail) {
    const uint32_t b = src[16u * n16 + lane];
    dst[16u * n16 + lane] = (uint8_t)b;
    if (SUM) { s1 += b; t += b * (pos0 + 16u * n16 + lane); }
  }
}
/* The same with the Adler-32 partial sums: the copy is cut at the 16 KiB segment boundaries of the
 * OUTPUT (K3's segments), and each piece adds (sum, sum of byte * index in the segment mod 65521)
 * to its segment's entry of parts[]. */
__device__ __forceinline__ void pz_warp_copy_summed(uint8_t *out, uint32_t dst, const uint8_t *src, uint32_t len, uint2 *parts) {
  const uint32_t lane = threadIdx.x & 31u;
  while (len) {
    const uint32_t seg = dst / PZ_ADLER_SEG, pos0 = dst % PZ_ADLER_SEG;
    const uint32_t piece = len < PZ_ADLER_SEG - pos0 ? len : PZ_ADLER_SEG - pos0;
    uint32_t s1 = 0, t = 0;
    pz_warp_copy<true>(out + dst, src, piece, pos0, s1, t);
    s1 = __reduce_add_sync(0xffffffffu, s1);
    t = __reduce_add_sync(0xffffffffu, t % 65521u);
    if (lane == 0) { atomicAdd(&parts[seg].x, s1); atomicAdd(&parts[seg].y, t % 65521u); }
    dst += piece; src += piece; len -= piece;
  }
}

__global__ void __launch_bounds__(PZ_ST_THREADS)
pz_stored_copy_kernel(const PzJob job, const uint32_t tile_streams) {
  __shared__ PzStoredShared sh;
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  volatile PzStoredShared *vs = &sh;
  for (uint32_t tile = blockIdx.x * tile_streams; tile < job.count; tile += gridDim.x * tile_streams) {
    /* candidates of this tile of tile_streams (<= 256) streams */
    if (tid == 0) sh.ncand = 0;
    __syncthreads();
    if (tid < tile_streams && tile + tid < job.count && job.res[job.first + tile + tid].status == PZ_ST_CANDIDATE)
      sh.cand[atomicAdd(&sh.ncand, 1u)] = job.first + tile + tid;
    __syncthreads();
    const uint32_t ncand = sh.ncand;
    for (uint32_t ci = 0; ci < ncand; ci++) {
      const uint32_t s = sh.cand[ci];
      const uint64_t i0 = job.in_off[s];
      const uint32_t n = (uint32_t)(job.in_off[s + 1] - i0);
      const uint8_t *in = job.in_blob + i0;
      const uint64_t o0 = job.out_off[s], ocap = job.out_off[s + 1] - o0;
      uint8_t *out = job.out_blob + o0;
      const uint32_t cap = ocap > 0xfffdff00ull ? 0xfffdff00u : (uint32_t)ocap;
      if (tid < 7) sh.done[tid] = 0;
      if (tid == 0) { sh.head = 0; sh.ended = 0; sh.ok = 0; }
      uint2 *const parts = job.parts ? job.parts + job.seg_off[s] : nullptr;
      if (parts) { /* this stream's partial sums start from zero (the copy warps add to them) */
        const uint32_t nseg = (uint32_t)((ocap + PZ_ADLER_SEG - 1u) / PZ_ADLER_SEG);
        for (uint32_t j = tid; j < nseg; j += PZ_ST_THREADS) parts[j] = make_uint2(0u, 0u);
        __threadfence();
      }
      __syncthreads();
      if (warp == 0) {
        if (lane == 0) { /* the walker: inflate's block loop restricted to stored blocks */
          uint32_t p = (in[1] & 0x20u) ? 6u : 2u; /* FDICT: four bytes skipped (Zlib.hs:68) */
          uint32_t pos = 0, base = 0, head = 0;
          bool ok = false;
          for (;;) {
            if (p >= n) break;                       /* truncated: K1 gives the verdict */
            const uint32_t b = in[p];
            if (((b >> 1) & 3u) != 0u) break;        /* a Huffman block (or BTYPE 3): K1 */
            p += 1u;                                 /* advanceToByte */
            if (n - p < 4u) break;
            const uint32_t len = in[p] | ((uint32_t)in[p + 1] << 8), nlen = in[p + 2] | ((uint32_t)in[p + 3] << 8);
            if (len != ((~nlen) & 0xffffu)) break;   /* Len/nlen mismatch */
            p += 4u;
            if (len >= n - p) break;                 /* getBlock needs strictly more than len bytes */
            if (pos - base + len > PZ_WINDOW) break; /* the reference's window would overflow */
            if (len > cap - pos) break;              /* PZ_OUTPUT_FULL */
            /* publish the descriptor once its ring slot has been consumed */
            /* descriptor `head` is the (head/7)-th of warp head%7 and reuses the slot of that warp's
             * descriptor nine earlier (63 = 7 x 9) */
            const uint32_t slot = head % PZ_ST_RING, owner = head % 7u, kth = head / 7u;
            while (head >= PZ_ST_RING && vs->done[owner] + 8u < kth) __nanosleep(50);
            vs->src[slot] = p; vs->dst[slot] = pos; vs->len[slot] = len;
            __threadfence_block();
            head++;
            vs->head = head;
            pos += len; p += len;
            if (pos - base >= 2u * PZ_EXCESS) base += PZ_EXCESS; /* moveWindow after every block */
            if (b & 1u) {                            /* BFINAL: the Adler-32 trailer, big-endian */
              if (n - p < 4u) break;
              sh.adler_stored = ((uint32_t)in[p] << 24) | ((uint32_t)in[p + 1] << 16) | ((uint32_t)in[p + 2] << 8) | in[p + 3];
              sh.out_len = pos; sh.end_byte = p + 4u; sh.base = base;
              ok = true;
              break;
            }
          }
          sh.ok = ok ? 1u : 0u;
          __threadfence_block();
          vs->ended = 1u;
        }
      } else { /* copy warps: warp w takes descriptors w-1, w-1+7, ... */
        const uint32_t me = warp - 1u;
        uint32_t mine = 0;
        for (uint32_t i = me;; i += 7u) {
          bool have;
          for (;;) { /* wait for descriptor i or the end of the walk */
            const uint32_t ended = vs->ended; /* read before head: head is final once ended is set */
            have = vs->head > i;
            if (have || ended) break;
            __nanosleep(100);
          }
          if (!have) break;
          const uint32_t slot = i % PZ_ST_RING;
          const uint32_t src = vs->src[slot], dst = vs->dst[slot], len = vs->len[slot];
          if (parts) pz_warp_copy_summed(out, dst, in + src, len, parts);
          else { uint32_t u0 = 0, u1 = 0; pz_warp_copy<false>(out + dst, in + src, len, 0u, u0, u1); }
          mine++;
          __syncwarp();
          if (lane == 0) vs->done[me] = mine;
        }
      }
      __syncthreads();
      if (tid == 0) {
        pz_result *res = job.res + s;
        if (sh.ok) {
          res->detail = 0;
          res->out_len = sh.out_len;
          res->adler_computed = parts ? PZ_ADLER_FUSED : 0u;
          res->adler_stored = sh.adler_stored;
          res->err_bitpos = (uint64_t)sh.end_byte * 8u;
          res->payload[0] = 0;
          res->payload[1] = (int64_t)sh.base;
          __threadfence();
          res->status = PZ_OK;
        } else {
          res->status = PZ_ST_PENDING;
        }
      }
      __syncthreads();
    }
  }
}
