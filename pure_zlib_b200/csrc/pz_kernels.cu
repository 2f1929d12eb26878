/*
 * pz_kernels.cu -- the sm_100a kernels of libpzcuda.so.
 *
 *   pz_inflate_kernel        K1: one persistent CTA per SM holding PZ_SLOTS zlib streams: a hot warp
 *                            (one lane per stream) runs the symbol loops, service warps parse
 *                            headers / build tables / feed the input rings, writer warps apply
 *                            the tokens to the output (pz_device.cuh)
 *   pz_stored_probe_kernel,  K2: streams made of stored blocks only are copied with 16-byte
 *   pz_stored_copy_kernel    accesses by whole CTAs before K1 runs (pz_stored.cuh)
 *   pz_fixed_kernel          K5 / K6: big batches of small streams: every fixed-Huffman stream is decoded by ONE thread
 *                            (no tables: the fixed code is arithmetic), then every small stream with dynamic blocks by
 *                            one thread with tables in its local memory, between K2 and K1 (pz_fixed.cuh)
 *   pz_adler_partial_kernel  K3a: one warp per 16 KiB segment of decoded output, dp4a sums
 *   pz_adler_finish_kernel   K3b: per stream, combines the segments (adler32-combine
 *                            identity) and compares with the stored trailer
 *   pz_code_values_kernel    canonical-code KAT hook (test/Test.hs:107-120)
 */
#include <cstdio>
#include <cstdlib>

#include "pz_device.cuh"
#include "pz_internal.h"
#include "pz_stored.cuh"
#include "pz_fixed.cuh"
#include "pz_huge.cuh"

/* Warp roles.  A CTA owns PZ_SLOTS stream slots: warp 0 is the hot warp (one lane per slot),
 * PZ_SERVICE_WARPS service warps serve four slots each (8 lanes per slot) and PZ_WRITER_WARPS
 * writer warps 32 / PZ_WGROUP slots each.  A warp's scheduler is its index modulo 4 when the CTA has the SM to itself, so the
 * multiples of four -- the hot warp's scheduler -- go to service warps, which sleep most of the
 * time, and the writers are dealt over the other three schedulers first. */
__device__ __forceinline__ void pz_role(uint32_t w, uint32_t &role, uint32_t &index) {
  if (w == 0u) { role = 0u; index = 0u; return; }
  if ((w & 3u) == 0u) { /* the hot warp's scheduler: service warps, then warps that exit at once */
    const uint32_t j = (w >> 2) - 1u;
    if (j + 1u <= PZ_HOT_SCHED_SERVICE) { role = 1u; index = j; }
    else { role = 3u; index = 0u; }
    return;
  }
  const uint32_t k = w - 1u - (w >> 2); /* rank among the other warps: writers first */
  if (k < PZ_WRITER_WARPS) { role = 2u; index = k; return; }
  role = 1u;
  index = PZ_HOT_SCHED_SERVICE + (k - PZ_WRITER_WARPS);
}

/* WIDE = block jobs (K4): 16-bit output symbols, and the gap rule of PzCtx::mark in the hot warp.
 * LEAN = plain batches: the hot warp decodes tokens without counting bytes, the writers check them (pz_hot_warp_lean); a
 * stream whose token fails a check is left PENDING for the exact kernel (LEAN = false), launched right behind. */
template <bool COUNT_ONLY, bool WIDE = false, bool LEAN = false>
__global__ void __launch_bounds__(PZ_THREADS_PER_CTA, 1)
pz_inflate_kernel(const PzJob job) {
  extern __shared__ __align__(16) unsigned char pz_smem_raw[];
  PzStreamSmem *slots = reinterpret_cast<PzStreamSmem *>(pz_smem_raw);
  /* empty token queues: every slot carries the phase the reader does NOT expect on lap 0 */
  for (uint32_t i = threadIdx.x; i < PZ_SLOTS * PZ_QLEN; i += PZ_THREADS_PER_CTA)
    slots[i / PZ_QLEN].q[i % PZ_QLEN] = make_uint2(0x80000000u, 0u);
  if (threadIdx.x < PZ_SLOTS) {
    slots[threadIdx.x].qtail = 0; slots[threadIdx.x].wpos = 0; slots[threadIdx.x].wmark = 0; slots[threadIdx.x].wbad = 0;
    slots[threadIdx.x].mail.state = PZ_MS_SERVICE;
    slots[threadIdx.x].mail.ring_hi = 0;
    slots[threadIdx.x].mail.hot_bp = 0;
  }
  __syncthreads();
  uint32_t role, index;
  pz_role(threadIdx.x >> 5, role, index);
  if (role == 0u) {
    if (LEAN) pz_hot_warp_lean(slots, PZ_SLOTS);
    else pz_hot_warp<COUNT_ONLY, WIDE>(slots, PZ_SLOTS);
    return;
  }
  if (role == 3u) return;
  /* slot s of CTA b takes streams b + grid * (s + PZ_SLOTS * k): a batch spreads over the SMs
   * before it fills the slots of any of them */
  const uint32_t slot = role == 1u ? index * PZ_SLOTS_PER_SERVICE + (threadIdx.x & 31u) / PZ_G : index * PZ_SLOTS_PER_WRITER + (threadIdx.x & 31u) / PZ_WG;
  const bool present = slot < PZ_SLOTS;
  PzStreamSmem *sm = slots + (present ? slot : 0u);
  if (role == 1u) {
    pz_decoder_warp<COUNT_ONLY>(job, job.first + blockIdx.x + gridDim.x * slot, gridDim.x * PZ_SLOTS, sm, present, LEAN);
  } else if (!COUNT_ONLY) {
    pz_writer_warp<WIDE, LEAN>(job, sm, present);
  }
}

/* ---- Adler-32 (Adler32.hs:17-57) as a segmented reduction ------------------------------
 * For a segment of L bytes d_0..d_{L-1}:  S1 = sum d_i,  S2 = sum (L - i) d_i.  Appending it
 * to a running (a, b):  b' = b + L*a + S2,  a' = a + S1  (mod 65521). */
#define PZ_ADLER_MOD 65521u

__global__ void __launch_bounds__(256)
pz_adler_partial_kernel(const uint8_t *__restrict__ out_blob, const uint64_t *__restrict__ out_off,
                        const uint64_t *__restrict__ seg_off, uint32_t n_total, uint64_t seg_first, uint64_t seg_count,
                        const pz_result *__restrict__ res, uint2 *__restrict__ parts) {
  const uint64_t gw = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gw >= seg_count) return;
  const uint64_t g = seg_first + gw;
  const uint32_t lane = threadIdx.x & 31u;
  /* the stream owning segment g: the last i with seg_off[i] <= g */
  uint32_t lo = 0, hi = n_total;
  while (hi - lo > 1) {
    uint32_t mid = lo + (hi - lo) / 2;
    if (seg_off[mid] <= g) lo = mid; else hi = mid;
  }
  const uint32_t i = lo;
  if (res[i].status != PZ_OK) return;
  if (res[i].adler_computed == PZ_ADLER_FUSED) return; /* K2 summed this stream while copying it */
  const uint64_t len = res[i].out_len;
  const uint64_t start = (g - seg_off[i]) * (uint64_t)PZ_ADLER_SEG;
  if (start >= len) return;
  const uint32_t L = (uint32_t)(len - start < PZ_ADLER_SEG ? len - start : PZ_ADLER_SEG);
  const uint8_t *p0 = out_blob + out_off[i] + start;
  const uint32_t head = (uint32_t)((uintptr_t)p0 & 15u);
  const uint4 *v = reinterpret_cast<const uint4 *>(p0 - head);
  const uint32_t nvec = (head + L + 15u) >> 4;
  uint32_t s1 = 0;
  uint64_t s2 = 0;
  for (uint32_t t = lane; t < nvec; t += 32u) {
    uint4 w = v[t];
    const int32_t q = (int32_t)(t * 16u) - (int32_t)head; /* index of this vector's first byte */
    if (q < 0 || q + 16 > (int32_t)L) {                   /* boundary vector: drop foreign bytes */
      uint32_t x[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int b = 0; b < 16; b++) {
        int32_t idx = q + b;
        if (idx < 0 || idx >= (int32_t)L) x[b >> 2] &= ~(0xffu << (8 * (b & 3)));
      }
      w = make_uint4(x[0], x[1], x[2], x[3]);
    }
    uint32_t sum = __dp4a(w.x, 0x01010101u, 0u);
    sum = __dp4a(w.y, 0x01010101u, sum);
    sum = __dp4a(w.z, 0x01010101u, sum);
    sum = __dp4a(w.w, 0x01010101u, sum);
    uint32_t ks = __dp4a(w.x, 0x03020100u, 0u);
    ks = __dp4a(w.y, 0x07060504u, ks);
    ks = __dp4a(w.z, 0x0b0a0908u, ks);
    ks = __dp4a(w.w, 0x0f0e0d0cu, ks);
    s1 += sum;
    s2 += (uint64_t)((int32_t)L - q) * sum - ks; /* sum_k (L - (q+k)) d_k */
  }
  s1 = __reduce_add_sync(0xffffffffu, s1);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s2 += __shfl_down_sync(0xffffffffu, s2, o);
  if (lane == 0) parts[g] = make_uint2(s1 % PZ_ADLER_MOD, (uint32_t)(s2 % PZ_ADLER_MOD));
}

/* One warp per stream.  A range of the stream is summarised as (s1, s2, n): s1 = sum of its bytes,
 * s2 = sum of (n - i) * byte_i, both mod 65521; appending range Y to range X gives
 * (s1x + s1y, s2x + n_y * s1x + s2y, n_x + n_y).  Each lane folds a contiguous run of segments, the
 * 32 partial summaries are folded with shuffles, and Adler = ((n + s2) mod p) << 16 | (1 + s1) mod p
 * (initialAdlerState a = 1, b = 0: Adler32.hs:19-20). */
__global__ void __launch_bounds__(128)
pz_adler_finish_kernel(const uint64_t *__restrict__ seg_off, uint32_t first, uint32_t count, pz_result *res,
                       const uint2 *__restrict__ parts, uint32_t compare) {
  const uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  if (k >= count) return;
  const uint32_t i = first + k;
  if (res[i].status != PZ_OK) return;
  const uint64_t len = res[i].out_len;
  const uint64_t nseg = (len + PZ_ADLER_SEG - 1) / PZ_ADLER_SEG;
  const uint2 *p = parts + seg_off[i];
  /* K2's form of a segment: (sum of bytes, sum of byte * index); sum (L - index) * byte follows */
  const bool fused = res[i].adler_computed == PZ_ADLER_FUSED;
  const uint64_t per = (nseg + 31u) / 32u;
  const uint64_t j0 = per * lane < nseg ? per * lane : nseg, j1 = j0 + per < nseg ? j0 + per : nseg;
  uint32_t s1 = 0, s2 = 0;
  uint64_t n = 0; /* bytes summarised by this lane */
  for (uint64_t j = j0; j < j1; j++) {
    const uint64_t rest = len - j * PZ_ADLER_SEG;
    const uint32_t L = (uint32_t)(rest < PZ_ADLER_SEG ? rest : PZ_ADLER_SEG);
    uint2 s = p[j];
    if (fused) {
      s.x %= PZ_ADLER_MOD;
      s.y = (uint32_t)(((uint64_t)L * s.x + PZ_ADLER_MOD - s.y % PZ_ADLER_MOD) % PZ_ADLER_MOD);
    }
    s2 = (uint32_t)(((uint64_t)s2 + (uint64_t)L * s1 + s.y) % PZ_ADLER_MOD);
    s1 = (s1 + s.x) % PZ_ADLER_MOD;
    n += L;
  }
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { /* lane l absorbs the summary of the 'o' lanes to its right, pairwise */
    const uint32_t t1 = __shfl_down_sync(0xffffffffu, s1, o), t2 = __shfl_down_sync(0xffffffffu, s2, o);
    const uint64_t tn = __shfl_down_sync(0xffffffffu, n, o);
    if ((lane & (2u * (uint32_t)o - 1u)) == 0u) {
      s2 = (uint32_t)(((uint64_t)s2 + (tn % PZ_ADLER_MOD) * s1 + t2) % PZ_ADLER_MOD);
      s1 = (s1 + t1) % PZ_ADLER_MOD;
      n += tn;
    }
  }
  if (lane != 0) return;
  const uint32_t a = (1u + s1) % PZ_ADLER_MOD, b = (uint32_t)((len % PZ_ADLER_MOD + s2) % PZ_ADLER_MOD);
  const uint32_t adler = (b << 16) | a; /* finalizeAdler (Adler32.hs:53-57) */
  res[i].adler_computed = adler;
  if (compare && adler != res[i].adler_stored) { /* checkChecksum (Deflate.hs:56-63); raw deflate has no trailer to compare with */
    res[i].status = PZ_ERR_CHECKSUM;
    res[i].detail = PZ_D_ADLER_MISMATCH;
  }
}


/* ---- CRC-32 (RFC 1952 section 8; gzip framing, an extension beyond the reference) as a segmented reduction ----------
 * Arithmetic: polynomials over GF(2) modulo P = 0xedb88320 (reflected: bit 31 is x^0).  For the register R(M, init) after
 * message M:  R(A || B, init) = R(A, init) * x^(8|B|) + R(B, 0)  (mod P), so a message is cut into pieces, each piece gives
 * its zero-init register, and pieces are put together with multiplications by x^(8 * bytes behind the piece):
 *   lane      512 consecutive bytes of a 16 KiB segment, four table look-ups per 32-bit word (tables in shared memory);
 *   segment   the 32 lanes' registers, each times x^(8 * bytes behind its piece in the segment), XORed (one shuffle tree);
 *   stream    the segments folded the same way by one warp (pz_crc_finish_kernel); CRC = ~(R + 0xffffffff * x^(8 len)). */
#define PZ_CRC_POLY 0xedb88320u
__device__ __forceinline__ uint32_t pz_gf_mul(uint32_t a, uint32_t b) { /* a * b mod P */
  uint32_t p = 0;
#pragma unroll 1
  for (uint32_t m = 0x80000000u; m != 0u && a != 0u; m >>= 1) {
    if (a & m) { p ^= b; a &= ~m; }
    b = (b >> 1) ^ (PZ_CRC_POLY & (0u - (b & 1u)));
  }
  return p;
}
/* x^(8 n) mod P from the table sq[j] = x^(2^j) mod P */
__device__ __forceinline__ uint32_t pz_gf_xpow8(uint64_t n, const uint32_t *sq) {
  uint32_t p = 0x80000000u; /* x^0 */
  uint32_t j = 3;
#pragma unroll 1
  for (; n != 0u; n >>= 1, j++) {
    if (n & 1u) p = pz_gf_mul(sq[j & 63u], p);
  }
  return p;
}
struct PzCrcSmem {
  uint32_t t[4][256]; /* t[0] = the byte table; t[k][i] = t[0][i] advanced by k more zero bytes */
  uint32_t sq[64];    /* x^(2^j) mod P (j < 64: byte counts below 2^61) */
  uint32_t behind[32];/* x^(8 * 512 * (31 - lane)): bytes behind lane's piece in a full segment */
};
__device__ void pz_crc_tables(PzCrcSmem &sm) {
  for (uint32_t i = threadIdx.x; i < 256u; i += blockDim.x) {
    uint32_t c = i;
    for (int k = 0; k < 8; k++) c = (c >> 1) ^ (PZ_CRC_POLY & (0u - (c & 1u)));
    sm.t[0][i] = c;
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < 256u; i += blockDim.x) {
    uint32_t c = sm.t[0][i];
    for (int k = 1; k < 4; k++) { c = sm.t[0][c & 0xffu] ^ (c >> 8); sm.t[k][i] = c; }
  }
  if (threadIdx.x == 0) {
    uint32_t v = 0x40000000u; /* x^1 */
    for (int j = 0; j < 64; j++) { sm.sq[j] = v; v = pz_gf_mul(v, v); }
  }
  __syncthreads();
  if (threadIdx.x < 32u) sm.behind[threadIdx.x] = pz_gf_xpow8(512u * (31u - threadIdx.x), sm.sq);
  __syncthreads();
}

/* parts[g] = (zero-init CRC register of segment g, its length): one warp per 16 KiB segment, as pz_adler_partial_kernel */
__global__ void __launch_bounds__(256)
pz_crc_partial_kernel(const uint8_t *__restrict__ out_blob, const uint64_t *__restrict__ out_off,
                      const uint64_t *__restrict__ seg_off, uint32_t n_total, uint64_t seg_first, uint64_t seg_count,
                      const pz_result *__restrict__ res, uint2 *__restrict__ parts) {
  __shared__ PzCrcSmem sm;
  pz_crc_tables(sm);
  const uint64_t gw = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gw >= seg_count) return;
  const uint64_t g = seg_first + gw;
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t lo = 0, hi = n_total;
  while (hi - lo > 1) {
    uint32_t mid = lo + (hi - lo) / 2;
    if (seg_off[mid] <= g) lo = mid; else hi = mid;
  }
  const uint32_t i = lo;
  if (res[i].status != PZ_OK) return;
  const uint64_t len = res[i].out_len;
  const uint64_t start = (g - seg_off[i]) * (uint64_t)PZ_ADLER_SEG;
  if (start >= len) return;
  const uint32_t L = (uint32_t)(len - start < PZ_ADLER_SEG ? len - start : PZ_ADLER_SEG);
  const uint8_t *p0 = out_blob + out_off[i] + start;
  /* lane's piece: bytes [512 lane, min(512 lane + 512, L)) */
  const uint32_t b0 = 512u * lane, b1 = b0 + 512u < L ? b0 + 512u : L;
  uint32_t c = 0;
  if (b0 < L) {
    const uint8_t *p = p0 + b0, *e = p0 + b1;
    while (p < e && ((uintptr_t)p & 3u)) { c = sm.t[0][(c ^ *p++) & 0xffu] ^ (c >> 8); }
    for (; p + 4 <= e; p += 4) {
      c ^= *reinterpret_cast<const uint32_t *>(p);
      c = sm.t[3][c & 0xffu] ^ sm.t[2][(c >> 8) & 0xffu] ^ sm.t[1][(c >> 16) & 0xffu] ^ sm.t[0][c >> 24];
    }
    while (p < e) { c = sm.t[0][(c ^ *p++) & 0xffu] ^ (c >> 8); }
    /* bytes of the segment behind this piece */
    c = pz_gf_mul(c, L == PZ_ADLER_SEG ? sm.behind[lane] : pz_gf_xpow8(L - b1, sm.sq));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c ^= __shfl_down_sync(0xffffffffu, c, o);
  if (lane == 0) parts[g] = make_uint2(c, L);
}

/* One warp per stream: folds the segments' registers, finishes the CRC and compares it and ISIZE with the trailer
 * (RFC 1952 2.3.1; the checks of checkChecksum, Deflate.hs:52-63, for this framing). */
__global__ void __launch_bounds__(128)
pz_crc_finish_kernel(const uint64_t *__restrict__ seg_off, uint32_t first, uint32_t count, pz_result *res,
                     const uint2 *__restrict__ parts, uint32_t compare) {
  __shared__ PzCrcSmem sm;
  pz_crc_tables(sm);
  const uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  if (k >= count) return;
  const uint32_t i = first + k;
  if (res[i].status != PZ_OK) return;
  const uint64_t len = res[i].out_len;
  const uint64_t nseg = (len + PZ_ADLER_SEG - 1) / PZ_ADLER_SEG;
  const uint2 *p = parts + seg_off[i];
  const uint64_t per = (nseg + 31u) / 32u;
  const uint64_t j0 = per * lane < nseg ? per * lane : nseg, j1 = j0 + per < nseg ? j0 + per : nseg;
  const uint32_t x16k = pz_gf_xpow8(PZ_ADLER_SEG, sm.sq);
  uint32_t c = 0;
  uint64_t n = 0; /* bytes this lane has folded */
  for (uint64_t j = j0; j < j1; j++) {
    const uint2 s = p[j];
    c = pz_gf_mul(c, s.y == PZ_ADLER_SEG ? x16k : pz_gf_xpow8(s.y, sm.sq)) ^ s.x;
    n += s.y;
  }
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { /* lane l absorbs the 'o' lanes to its right, pairwise */
    const uint32_t tc = __shfl_down_sync(0xffffffffu, c, o);
    const uint64_t tn = __shfl_down_sync(0xffffffffu, n, o);
    if ((lane & (2u * (uint32_t)o - 1u)) == 0u) {
      c = pz_gf_mul(c, pz_gf_xpow8(tn, sm.sq)) ^ tc;
      n += tn;
    }
  }
  if (lane != 0) return;
  const uint32_t crc = ~(c ^ pz_gf_mul(0xffffffffu, pz_gf_xpow8(len, sm.sq)));
  res[i].adler_computed = crc;
  if (!compare) return; /* the incremental driver folds piece checksums itself */
  if (crc != res[i].adler_stored) {
    res[i].status = PZ_ERR_CHECKSUM;
    res[i].detail = PZ_D_ADLER_MISMATCH;
  } else if ((uint32_t)res[i].payload[0] != (uint32_t)len) { /* ISIZE: the decoded length modulo 2^32 */
    res[i].status = PZ_ERR_CHECKSUM;
    res[i].detail = PZ_D_LENGTH_MISMATCH;
    res[i].payload[0] = (int64_t)(uint32_t)res[i].payload[0]; /* payload[1] stays the published-bytes count */
  }
}

/* host-buffer CRC-32 for pz_crc32(): one stream, the two kernels above */
__global__ void pz_code_values_kernel(const uint8_t *lens, int n, uint16_t *codes) {
  __shared__ PzStreamSmem sm; /* launched with one group (PZ_G threads) */
  for (int i = threadIdx.x; i < n; i += PZ_G) sm.lens[i] = lens[i];
  pz_syncwarp();
  int64_t val;
  (void)pz_build<PZ_LIT_BITS, 1>(sm.lens, n, &sm.lit, sm.lit_perm, sm.lit_lut, sm.scratch, &val);
  uint16_t *c16 = reinterpret_cast<uint16_t *>(sm.lit_lut); /* the LUT itself is not needed here */
  pz_canon_codes(sm.lens, &sm.lit, sm.lit_perm, c16, false);
  for (int i = threadIdx.x; i < n; i += PZ_G) codes[i] = sm.lens[i] ? c16[i] : 0;
}

/* Incremental contexts: piece j = g[3j+2] bytes from device address g[3j] to the pinned host address g[3j+1] (reachable
 * from the device: unified addressing), one CTA per piece.  Replaces one cudaMemcpyAsync per stream and pump. */
__global__ void __launch_bounds__(256)
pz_gather_kernel(const uint64_t *__restrict__ g) {
  const uint8_t *src = reinterpret_cast<const uint8_t *>((uintptr_t)g[3u * blockIdx.x]);
  uint8_t *dst = reinterpret_cast<uint8_t *>((uintptr_t)g[3u * blockIdx.x + 1u]);
  const uint64_t len = g[3u * blockIdx.x + 2u];
  if ((((uintptr_t)src | (uintptr_t)dst) & 15u) == 0u) {
    const uint64_t nv = len >> 4;
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
    uint4 *d4 = reinterpret_cast<uint4 *>(dst);
    for (uint64_t i = threadIdx.x; i < nv; i += 256u) d4[i] = s4[i];
    for (uint64_t i = (nv << 4) + threadIdx.x; i < len; i += 256u) dst[i] = src[i];
  } else {
    for (uint64_t i = threadIdx.x; i < len; i += 256u) dst[i] = src[i];
  }
}

/* ---- launch wrappers ------------------------------------------------------------------ */
static int g_inflate_ctas_per_sm[2] = {0, 0};
static int g_sm_count = 0;

/* Unit counters for launches that claim their streams (PzJob::next_unit): a ring of words per device, one word per launch,
 * zeroed on the launch's stream right before it (a word comes round again after 1024 launches). */
#include <atomic>
static uint32_t *g_claim_ring[PZ_MAX_DEVICES] = {};
static std::atomic<unsigned> g_claim_next{0};
static cudaError_t pz_claim_counter(cudaStream_t st, uint32_t **out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= PZ_MAX_DEVICES) { *out = nullptr; return cudaSuccess; }
  if (g_claim_ring[dev] == nullptr) { *out = nullptr; return cudaSuccess; } /* not configured: deal by index */
  uint32_t *p = g_claim_ring[dev] + (g_claim_next++ % 1024u);
  e = cudaMemsetAsync(p, 0, sizeof(uint32_t), st);
  *out = e == cudaSuccess ? p : nullptr;
  return e;
}

/* Function attributes belong to the CURRENT device's context: call once per device the library uses. */
cudaError_t pz_kernels_configure(void) {
  cudaError_t e;
  const size_t smem = sizeof(PzStreamSmem) * PZ_SLOTS;
  int dev = 0;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
  if ((e = cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
  if (dev >= 0 && dev < PZ_MAX_DEVICES && g_claim_ring[dev] == nullptr && (e = cudaMalloc(&g_claim_ring[dev], 1024 * sizeof(uint32_t))) != cudaSuccess) return e;
  { /* stream-ordered scratch (the lists of K5 / K6, the incremental contexts) is handed back to the pool, not to the driver */
    cudaMemPool_t pool;
    uint64_t keep = ~0ull;
    if ((e = cudaDeviceGetDefaultMemPool(&pool, dev)) != cudaSuccess) return e;
    if ((e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep)) != cudaSuccess) return e;
  }
  e = cudaFuncSetAttribute(pz_inflate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(pz_inflate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute((pz_inflate_kernel<false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute((pz_inflate_kernel<true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute((pz_inflate_kernel<true, true>), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute((pz_inflate_kernel<false, true>), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(pz_inflate_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(pz_inflate_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute((pz_inflate_kernel<false, false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute((pz_inflate_kernel<false, false, true>), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  /* K6's occupancy is set by a shared-memory request it never touches (its tables live in local memory: see pz_launch_inflate) */
  e = cudaFuncSetAttribute((pz_fixed_kernel<false, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute((pz_fixed_kernel<false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute((pz_fixed_kernel<true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(pz_blk_tails_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(PZ_TAIL * sizeof(uint16_t)));
  if (e != cudaSuccess) return e;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_inflate_ctas_per_sm[0], pz_inflate_kernel<false>, PZ_THREADS_PER_CTA, smem);
  if (e != cudaSuccess) return e;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_inflate_ctas_per_sm[1], pz_inflate_kernel<true>, PZ_THREADS_PER_CTA, smem);
  if (e != cudaSuccess) return e;
  if (g_inflate_ctas_per_sm[0] < 1 || g_inflate_ctas_per_sm[1] < 1) return cudaErrorLaunchOutOfResources;
  return cudaSuccess;
}

int pz_inflate_slots(void) { return g_sm_count * g_inflate_ctas_per_sm[0] * (int)PZ_SLOTS; }

/* One resident wave of persistent CTAs: one per SM (148 on B200), fewer for small batches. */
/* Launches the small-stream kernels add to a batch of `count` streams: 0 (not run), 2 (list + K5) or 4 (+ list + K6). */
int pz_small_launches(uint32_t count, uint32_t framing) {
  static const bool no_k5 = getenv("PZ_NO_K5") != nullptr, want_k6 = getenv("PZ_K6") != nullptr && getenv("PZ_NO_K6") == nullptr; /* A/B */
  if (no_k5 || count < PZ_FIXED_MIN_STREAMS || (framing & 0xffu) != PZ_FRAME_ZLIB) return 0;
  return want_k6 ? 4 : 2;
}

cudaError_t pz_launch_inflate(const uint8_t *d_in, const uint64_t *d_in_off, uint8_t *d_out, const uint64_t *d_out_off,
                              uint32_t first, uint32_t count, pz_result *d_res, cudaStream_t st, uint32_t *d_prog,
                              const uint32_t *d_in_ready, int phase, uint2 *d_parts, const uint64_t *d_seg_off, uint32_t framing) {
  if (count == 0) return cudaSuccess;
  const bool count_only = d_out == nullptr;
  const unsigned wave = (unsigned)(g_sm_count * g_inflate_ctas_per_sm[count_only ? 1 : 0]);
  const unsigned grid = count < wave ? count : wave; /* one stream per CTA before any CTA gets two */
  const size_t smem = sizeof(PzStreamSmem) * PZ_SLOTS;
  PzJob job;
  job.in_blob = d_in; job.in_off = d_in_off; job.out_blob = d_out; job.out_off = d_out_off; job.res = d_res;
  job.first = first; job.count = count; job.skip_done = count_only ? 0u : 1u;
  job.prog = count_only ? nullptr : d_prog;
  job.in_ready = d_in_ready;
  job.blk_start = nullptr; job.blk_out = nullptr; job.blk_len = nullptr; job.out16 = nullptr; job.blk_stream = 0; job.blk_cap = 0;
  job.parts = d_seg_off ? d_parts : nullptr; job.seg_off = d_seg_off;
  /* bit 8 of `framing`: the caller has marked the streams PENDING itself (K4 ran in between): K1 keeps skipping the others */
  const bool premarked = (framing & 0x100u) != 0u;
  framing &= 0xffu;
  job.framing = framing;
  if (framing != PZ_FRAME_ZLIB) { /* K2 reads zlib framing only (and it is K2 that marks streams PENDING) */
    if (phase == PZ_PHASE_K2) return cudaSuccess;
    if (!premarked) job.skip_done = 0;
    phase = PZ_PHASE_K1;
  }
  if (d_in_ready) job.skip_done = 0; /* K2 would read input that is not there yet: K1 decodes every stream */
  /* K6 (the same kernel with private dynamic-code tables in local memory) is OFF unless PZ_K6 is set: measured on B200
   * (profiles/r02n_*, r02k_*) it decodes the dynamic quarter of config 3 in 28 ms where K1 needs 25 ms, and sizes it in 17 ms
   * against 16 ms: with a thread per stream the 3.5 KiB of tables per thread are a gigabyte in flight, every look-up is a DRAM
   * sector (74 GB read for 1.07 GB decoded), and with few enough threads for the L2 to hold them nothing hides the latency
   * (profiles/r02l_*: 64 / 128 / 256 threads per SM = 238 / 137 / 99 ms).  Kept as a measured design alternative. */
  const int small = d_in_ready == nullptr ? pz_small_launches(count, framing) : 0;
  const bool k5 = small != 0, k6 = small == 4;
  /* K6 keeps 3.5 KiB of tables per thread in local memory: with every thread an SM can hold resident that is 7 MiB per SM,
   * a gigabyte in all, and every look-up would go to DRAM.  A shared-memory request the kernel never touches limits it to
   * PZ_K6_BLOCKS blocks of 256 threads per SM (default 2: 270 MB of tables in all, about twice the L2). */
  static const int k6_blocks = getenv("PZ_K6_BLOCKS") ? atoi(getenv("PZ_K6_BLOCKS")) : 4;
  static const unsigned k6_threads = getenv("PZ_K6_THREADS") ? (unsigned)atoi(getenv("PZ_K6_THREADS")) : 256u;
  static const int k5_blocks = getenv("PZ_K5_BLOCKS") ? atoi(getenv("PZ_K5_BLOCKS")) : 8; /* (A/B: fewer streams in flight = a smaller working set in L2) */
  const size_t k5_smem = k5_blocks >= 8 ? 0 : k5_blocks <= 1 ? 120 * 1024 : (size_t)(220 * 1024 / k5_blocks) - 2048;
  const size_t k6_smem = k6_blocks >= 8 ? 0 : k6_blocks <= 1 ? 120 * 1024 : (size_t)(220 * 1024 / k6_blocks) - 2048;
  static const unsigned k5_threads = getenv("PZ_K5_THREADS") ? (unsigned)atoi(getenv("PZ_K5_THREADS")) : PZ_FIXED_THREADS;
  const unsigned small_grid = (count + k5_threads - 1u) / k5_threads, k6_grid = (count + k6_threads - 1u) / k6_threads;
  /* K5 then K6 over packed lists of the streams each will try (pz_small_list_kernel); the list lives in stream-ordered memory */
  auto small_streams = [&](bool co) -> cudaError_t {
    uint32_t *list = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&list, ((size_t)count + 2u) * sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    for (int dyn = 0; dyn < (k6 ? 2 : 1); dyn++) {
      if ((e = cudaMemsetAsync(list, 0, 8, st)) != cudaSuccess) return e;
      pz_small_list_kernel<<<(count + 255u) / 256u, 256, 0, st>>>(job, list, (uint32_t)dyn);
      if (!dyn) {
        if (co) pz_fixed_kernel<true, false><<<small_grid, k5_threads, 0, st>>>(job, list);
        else pz_fixed_kernel<false, false><<<small_grid, k5_threads, k5_smem, st>>>(job, list);
      } else {
        if (co) pz_fixed_kernel<true, true><<<k6_grid, k6_threads, k6_smem, st>>>(job, list);
        else pz_fixed_kernel<false, true><<<k6_grid, k6_threads, k6_smem, st>>>(job, list);
      }
    }
    return cudaFreeAsync(list, st);
  };
  if (count_only) {
    if (k5) { /* K5 sizes the small fixed-Huffman streams, K1 what it left */
      pz_mark_pending_kernel<<<(count + 255u) / 256u, 256, 0, st>>>(job);
      cudaError_t e5 = small_streams(true);
      if (e5 != cudaSuccess) return e5;
      job.skip_done = 1;
      cudaError_t e = pz_claim_counter(st, &job.next_unit);
      if (e != cudaSuccess) return e;
    }
    pz_inflate_kernel<true><<<grid, PZ_THREADS_PER_CTA, smem, st>>>(job);
  } else if (d_in_ready) {
    pz_inflate_kernel<false><<<grid, PZ_THREADS_PER_CTA, smem, st>>>(job);
  } else {
    /* K2 first: all-stored streams are copied at memory speed and marked done; K1 takes the rest */
    if (phase != PZ_PHASE_K1) {
      pz_stored_probe_kernel<<<(count + 255u) / 256u, 256, 0, st>>>(job);
      const unsigned ctas = (unsigned)g_sm_count * 4u;
      unsigned tile = count / (ctas * 4u);
      tile = tile < 1u ? 1u : (tile > PZ_ST_THREADS ? PZ_ST_THREADS : tile);
      const unsigned tiles = (count + tile - 1u) / tile;
      pz_stored_copy_kernel<<<tiles < ctas ? tiles : ctas, PZ_ST_THREADS, 0, st>>>(job, tile);
      if (k5) { cudaError_t e5 = small_streams(false); if (e5 != cudaSuccess) return e5; }
    }
    if (phase != PZ_PHASE_K2 && job.skip_done && count > (unsigned)pz_inflate_slots()) { /* more streams than slots, some of them finished already: claim */
      cudaError_t e = pz_claim_counter(st, &job.next_unit);
      if (e != cudaSuccess) return e;
    }
    if (phase != PZ_PHASE_K2) {
      /* The lean kernel is OFF unless PZ_LEAN is set: measured on B200 (profiles/r02f_*, r02g_*) it is slower than the exact
       * kernel alone (8.4 against 7.8 ms on config 2): a symbol costs the lone hot warp the same ~145 issue slots whether
       * its body has 46 instructions or 67 (the chain of dependent ALU operations and two table loads sets the pace, not
       * the instruction count), and every hand-over now waits for the writer.  Kept as a measured design alternative. */
      static const bool lean = getenv("PZ_LEAN") != nullptr;
      if (lean && d_prog == nullptr) {
        pz_inflate_kernel<false, false, true><<<grid, PZ_THREADS_PER_CTA, smem, st>>>(job);
        job.skip_done = 1; /* only what the lean kernel left PENDING: streams a writer's check refused */
        if (job.next_unit != nullptr) { cudaError_t e = pz_claim_counter(st, &job.next_unit); if (e != cudaSuccess) return e; }
      }
      pz_inflate_kernel<false><<<grid, PZ_THREADS_PER_CTA, smem, st>>>(job);
    }
  }
  return cudaGetLastError();
}

/* Resumable contexts (the incremental driver): stream s reads [d_in_pairs[2s], d_in_pairs[2s+1]) and writes into
 * [d_out_pairs[2s], d_out_pairs[2s+1]) -- device ADDRESSES, every stream in buffers of its own -- starting from the
 * checkpoint d_resume[4s..] and leaving the next one in d_ckpt[4s..] (PzJob::resume, PzJob::ckpt).  K1 only. */
cudaError_t pz_launch_resume(const uint64_t *d_in_pairs, const uint64_t *d_out_pairs, uint32_t count, pz_result *d_res,
                             const uint32_t *d_resume, uint32_t *d_ckpt, cudaStream_t st, uint32_t framing) {
  if (count == 0) return cudaSuccess;
  const unsigned wave = (unsigned)(g_sm_count * g_inflate_ctas_per_sm[0]);
  const unsigned grid = count < wave ? count : wave;
  const size_t smem = sizeof(PzStreamSmem) * PZ_SLOTS;
  PzJob job;
  job.in_blob = nullptr; job.in_off = d_in_pairs; job.out_blob = nullptr; job.out_off = d_out_pairs; job.res = d_res;
  job.first = 0; job.count = count; job.skip_done = 0; job.prog = nullptr; job.in_ready = nullptr;
  job.blk_start = nullptr; job.blk_out = nullptr; job.blk_len = nullptr; job.out16 = nullptr; job.blk_stream = 0; job.blk_cap = 0;
  job.parts = nullptr; job.seg_off = nullptr;
  job.resume = d_resume; job.ckpt = d_ckpt; job.pair_off = 1u; job.framing = framing;
  pz_inflate_kernel<false><<<grid, PZ_THREADS_PER_CTA, smem, st>>>(job);
  return cudaGetLastError();
}

cudaError_t pz_launch_gather(const uint64_t *d_triples, uint32_t count, cudaStream_t st) {
  if (count == 0) return cudaSuccess;
  pz_gather_kernel<<<count, 256, 0, st>>>(d_triples);
  return cudaGetLastError();
}

/* ---- K4 launchers (pz_huge.cuh; block jobs run on K1) ---------------------------------------- */
cudaError_t pz_launch_blk_search(const uint8_t *d_stream, uint64_t nbytes, uint64_t first_bit, uint64_t last_bit, uint32_t *d_cand,
                                 uint32_t *d_ncand, uint32_t cap, cudaStream_t st) {
  const uint64_t bytes = (last_bit + 7u) / 8u - first_bit / 8u;
  if (bytes == 0) return cudaSuccess;
  const uint64_t grid = (bytes + PZ_HUGE_THREADS - 1) / PZ_HUGE_THREADS;
  pz_blk_search_kernel<<<(unsigned)grid, PZ_HUGE_THREADS, 0, st>>>(d_stream, nbytes, first_bit, last_bit, d_cand, d_ncand, cap);
  return cudaGetLastError();
}
cudaError_t pz_launch_blk_verify(const uint8_t *d_stream, uint64_t nbytes, uint64_t last_bit, const uint32_t *d_cand, uint32_t ncand,
                                 uint32_t *d_kept, uint32_t *d_nkept, cudaStream_t st) {
  if (ncand == 0) return cudaSuccess;
  pz_blk_verify_kernel<<<(ncand + PZ_HUGE_THREADS - 1) / PZ_HUGE_THREADS, PZ_HUGE_THREADS, 0, st>>>(d_stream, nbytes, last_bit, d_cand, ncand, d_kept, d_nkept);
  return cudaGetLastError();
}
/* d_in_off2 = {offset of the stream in d_in_blob, its end}.  d_blk_out == nullptr: sizing pass (every block bounded by
 * cap); else 16-bit symbols of block j go to d_sym16 + d_blk_out[j], bounded by d_blk_len[j]. */
cudaError_t pz_launch_blk_jobs(const uint8_t *d_in_blob, const uint64_t *d_in_off2, const uint32_t *d_blk_start, const uint64_t *d_blk_out,
                               const uint32_t *d_blk_len, uint32_t cap, uint16_t *d_sym16, uint32_t count, pz_result *d_res, cudaStream_t st,
                               uint32_t *d_counter) {
  if (count == 0) return cudaSuccess;
  const bool count_only = d_blk_out == nullptr;
  const unsigned wave = (unsigned)(g_sm_count * g_inflate_ctas_per_sm[count_only ? 1 : 0]);
  const unsigned grid = count < wave ? count : wave;
  const size_t smem = sizeof(PzStreamSmem) * PZ_SLOTS;
  PzJob job;
  job.in_blob = d_in_blob; job.in_off = d_in_off2; job.out_blob = nullptr; job.out_off = nullptr; job.res = d_res;
  job.first = 0; job.count = count; job.skip_done = 0; job.prog = nullptr; job.in_ready = nullptr;
  job.blk_start = d_blk_start; job.blk_out = d_blk_out; job.blk_len = d_blk_len; job.out16 = d_sym16; job.blk_stream = 0; job.blk_cap = cap;
  job.parts = nullptr; job.seg_off = nullptr;
  if (d_counter != nullptr && count > 1u) { /* blocks are claimed, not dealt (PzJob::next_unit) */
    cudaError_t e = cudaMemsetAsync(d_counter, 0, sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    job.next_unit = d_counter;
  }
  if (count_only) pz_inflate_kernel<true, true><<<grid, PZ_THREADS_PER_CTA, smem, st>>>(job);
  else pz_inflate_kernel<false, true><<<grid, PZ_THREADS_PER_CTA, smem, st>>>(job);
  return cudaGetLastError();
}
cudaError_t pz_launch_blk_compact(const uint16_t *d_scr, uint16_t *d_sym16, const uint64_t *d_blk_off, const uint64_t *d_blk_src, uint32_t nblk,
                                  uint64_t total, cudaStream_t st) {
  if (nblk == 0 || total == 0) return cudaSuccess;
  const uint64_t per_cta = PZ_HUGE_THREADS * 8u;
  pz_blk_compact_kernel<<<(unsigned)((total + per_cta - 1) / per_cta), PZ_HUGE_THREADS, 0, st>>>(d_scr, d_sym16, d_blk_off, d_blk_src, nblk, total);
  return cudaGetLastError();
}
cudaError_t pz_launch_blk_resolve(uint16_t *d_sym16, uint8_t *d_out, const uint64_t *d_blk_off, const uint32_t *d_blk_len, const uint32_t *d_blk_grp,
                                  const uint32_t *d_grp_first, uint32_t ngrp, uint32_t nblk, uint64_t total, uint8_t *d_gw, uint32_t *d_err,
                                  cudaStream_t st) {
  if (nblk == 0 || total == 0) return cudaSuccess;
  pz_blk_tails_kernel<<<ngrp, PZ_TAILS_THREADS, PZ_TAIL * sizeof(uint16_t), st>>>(d_sym16, d_blk_off, d_blk_len, d_grp_first, ngrp);
  pz_blk_windows_kernel<<<1, PZ_TAILS_THREADS, 0, st>>>(d_sym16, d_blk_off, d_grp_first, ngrp, total, d_gw, d_err);
  const uint64_t per_cta = PZ_HUGE_THREADS * 8u;
  pz_blk_resolve_kernel<<<(unsigned)((total + per_cta - 1) / per_cta), PZ_HUGE_THREADS, 0, st>>>(d_sym16, d_out, d_blk_off, d_blk_len, d_blk_grp, d_grp_first,
                                                                                               nblk, total, d_gw, d_err);
  return cudaGetLastError();
}

cudaError_t pz_launch_adler(const uint8_t *d_out, const uint64_t *d_out_off, const uint64_t *d_seg_off, uint32_t n_total,
                            uint32_t first, uint32_t count, uint64_t seg_first, uint64_t seg_count, pz_result *d_res,
                            uint2 *d_parts, cudaStream_t st, uint32_t framing) {
  if (count == 0) return cudaSuccess;
  const uint32_t compare = (framing & 0x200u) ? 0u : 1u; /* bit 9: checksums only, no trailer comparison */
  framing &= 0xffu;
  const uint64_t warps_per_cta = 8;
  const unsigned grid = (unsigned)((seg_count + warps_per_cta - 1) / warps_per_cta);
  if (framing == PZ_FRAME_GZIP) { /* the checksum of this framing is CRC-32 */
    if (seg_count) pz_crc_partial_kernel<<<grid, 256, 0, st>>>(d_out, d_out_off, d_seg_off, n_total, seg_first, seg_count, d_res, d_parts);
    pz_crc_finish_kernel<<<(count + 3) / 4, 128, 0, st>>>(d_seg_off, first, count, d_res, d_parts, compare);
    return cudaGetLastError();
  }
  if (seg_count) pz_adler_partial_kernel<<<grid, 256, 0, st>>>(d_out, d_out_off, d_seg_off, n_total, seg_first, seg_count, d_res, d_parts);
  pz_adler_finish_kernel<<<(count + 3) / 4, 128, 0, st>>>(d_seg_off, first, count, d_res, d_parts, framing == PZ_FRAME_ZLIB ? compare : 0u); /* one warp per stream */
  return cudaGetLastError();
}

cudaError_t pz_launch_code_values(const uint8_t *d_lens, int n, uint16_t *d_codes, cudaStream_t st) {
  pz_code_values_kernel<<<1, PZ_G, 0, st>>>(d_lens, n, d_codes);
  return cudaGetLastError();
}

#ifdef PZ_PHASES
extern "C" void pz_debug_phases(unsigned long long *out) { /* debug build only: read and clear pz_phase_ticks */
  unsigned long long z[16] = {0};
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, pz_phase_ticks, sizeof(z));
  cudaMemcpyToSymbol(pz_phase_ticks, z, sizeof(z));
}
#endif
