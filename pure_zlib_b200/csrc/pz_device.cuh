/*
 * pz_device.cuh -- the device-side inflate engine (K1 and the logic K4's block jobs share with it).
 *
 * Execution model.  ONE persistent CTA per SM owns PZ_SLOTS stream slots (shared memory: LUTs, a staged
 * input ring, a token queue and a hand-over block per slot, struct PzStreamSmem) and three kinds of warps
 * (roles are dealt by warp index in pz_kernels.cu:pz_role):
 *
 *   hot warp       pz_hot_warp(): one LANE per slot runs the symbol loop (runInflate, Deflate.hs:106-120) of
 *                  that slot's stream -- window -> literal/length LUT -> shift -> distance LUT -> shift, the
 *                  96 stream bits at the bit position held in registers -- PZ_TRIP = three symbols per trip
 *                  (pz_fast_trip), speculatively: a symbol the loop must not decide (long code, end of block,
 *                  end of input or output in sight, a verdict) commits nothing and returns the stream to its
 *                  service group through the slot's mailbox (PzMail).  It never touches the output: every
 *                  literal / match is one 32-bit token in the slot's queue.
 *   service warps  pz_decoder_warp(): groups of PZ_G lanes, one group per slot, run everything rare or
 *                  order-sensitive with the reference's exact verdict order -- zlib header (pz_begin), block
 *                  headers and code-length decode (pz_slow_step, pz_dynamic_header), LUT construction
 *                  (pz_build), the bit-serial walk (pz_walk / pz_symbol_careful), stored blocks, the trailer,
 *                  checkpoints of resumable streams -- and keep the input ring ahead of the hot lane with
 *                  cp.async (pz_service_poll).
 *   writer warps   pz_writer_warp(): groups of PZ_WG lanes pop tokens and produce the bytes straight into
 *                  the stream's slice of the HBM output blob, which is also the LZ77 history: batches of
 *                  literals and short disjoint matches are dealt to the lanes by output POSITION, all
 *                  history loads of a batch issued before its first store; overlapping or long copies,
 *                  stored runs and control tokens go through pz_writer_apply() (byte-serial replicate
 *                  semantics of copyChunked, OutputWindow.hs:94-101).
 *
 * All of the reference's verdicts are decided on the decoder side from counters alone (bit position, bytes
 * produced, the reference's window fill/base model); nothing ever reads the output to decide one.
 *
 * What it replaces in the reference (file:line relative to the pure-zlib checkout):
 *   bit reader            Monad.hs:199-263   -> bit position over a cp.async-staged smem ring,
 *                                               register windows assembled with funnel shifts
 *   tree build + walk     HuffmanTree.hs:25-83, Deflate.hs:255-288 -> canonical counts + flat LUT
 *   block parser          Deflate.hs:65-156  -> pz_slow_step()
 *   symbol loop           Deflate.hs:106-120 -> pz_fast_trip() + exact "careful" path
 *   output window         OutputWindow.hs:29-114 -> pz_writer_warp(): direct stores; the window
 *                                              is only *modelled* (fill/base counters in the
 *                                              decoder) to reproduce its verdicts
 *   zlib framing          Zlib.hs:53-69, Deflate.hs:52-63
 *
 * The file also compiles with a host C++ compiler when PZ_HOSTSIM is defined: a group is then
 * a single lane, the symbol loop runs in the same thread (pz_fast_loop) and every pushed token is
 * applied to the output at once.  That build exists only for tests/hostsim (CPU-side differential
 * fuzzing of this logic against the oracle); the product library never contains it.
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "pzcuda.h"

#ifndef PZ_GROUP
#define PZ_GROUP 8 /* lanes per stream on the device: 32, 16, 8 or 4 */
#endif

#ifndef PZ_DOZE_NS
#define PZ_DOZE_NS 1000 /* service warps between polls of their hot lanes */
#endif
/* cache operators of the writer's history loads and output stores in the hot loop ("" = default, ".cg", ".cs", ...) */
#ifndef PZ_LD_MOD
#define PZ_LD_MOD ".cg"
#endif
#ifndef PZ_ST_MOD
#define PZ_ST_MOD ""
#endif
#ifndef PZ_NAP_NS
#define PZ_NAP_NS 1000 /* writer warps waiting for a full batch */
#endif
#ifndef PZ_WGROUP
#define PZ_WGROUP 16 /* lanes per stream in the writer warps: 8 or 16 */
#endif

#ifdef PZ_HOSTSIM
#include <string.h>
struct uint2 { uint32_t x, y; }; /* PzJob::parts (never used by the host build) */
#define PZ_DEV static inline
#define PZ_COLD static
#define PZ_G 1
#define PZ_WG 1
PZ_DEV int pz_wlane() { return 0; }
PZ_DEV void pz_wsyncwarp() {}
PZ_DEV uint32_t pz_brev(uint32_t x) {
  x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
  x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
  x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
  x = ((x >> 8) & 0x00ff00ffu) | ((x & 0x00ff00ffu) << 8);
  return (x >> 16) | (x << 16);
}
PZ_DEV int pz_lane() { return 0; }
PZ_DEV void pz_syncwarp() {}
PZ_DEV unsigned pz_ballot(int p) { return p ? 1u : 0u; }
PZ_DEV unsigned pz_match_any(unsigned) { return 1u; }
PZ_DEV unsigned pz_lanemask_lt() { return 0u; }
PZ_DEV int pz_shfl(int v, int) { return v; }
PZ_DEV bool pz_warp_any(bool p) { return p; }
PZ_DEV void pz_smem_inc(uint32_t *p) { ++*p; }
PZ_DEV int pz_popc(unsigned x) { return __builtin_popcount(x); }
PZ_DEV int pz_ffs(unsigned x) { return __builtin_ffs((int)x); }
PZ_DEV uint32_t pz_funnel_r(uint32_t lo, uint32_t hi, uint32_t s) { return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (s & 31u)); }
PZ_DEV uint32_t pz_funnel_l(uint32_t lo, uint32_t hi, uint32_t s) { return (uint32_t)((((((uint64_t)hi) << 32) | lo) << (s & 31u)) >> 32); }
PZ_DEV void pz_copy16_async(void *smem_dst, const void *gsrc) { memcpy(smem_dst, gsrc, 16); }
PZ_DEV void pz_async_wait_all() {}
#else
#define PZ_DEV __device__ __forceinline__
#define PZ_COLD __device__ __noinline__ /* rare paths: kept out of line so the hot loops stay in the instruction cache */
#define PZ_G PZ_GROUP
#define PZ_WG PZ_WGROUP
/* the writer warps' groups (PZ_WG lanes per stream) */
PZ_DEV unsigned pz_wgshift() { return (threadIdx.x & 31u) & ~(unsigned)(PZ_WG - 1); }
PZ_DEV unsigned pz_wgmask() { return PZ_WG == 32 ? 0xffffffffu : (((1u << (PZ_WG & 31)) - 1u) << pz_wgshift()); }
PZ_DEV int pz_wlane() { return (int)(threadIdx.x & (unsigned)(PZ_WG - 1)); }
PZ_DEV void pz_wsyncwarp() { __syncwarp(pz_wgmask()); }
PZ_DEV uint32_t pz_brev(uint32_t x) { return __brev(x); }
PZ_DEV unsigned pz_gshift() { return (threadIdx.x & 31u) & ~(unsigned)(PZ_G - 1); }
PZ_DEV unsigned pz_gmask() { return PZ_G == 32 ? 0xffffffffu : (((1u << (PZ_G & 31)) - 1u) << pz_gshift()); }
PZ_DEV int pz_lane() { return (int)(threadIdx.x & (unsigned)(PZ_G - 1)); }
PZ_DEV void pz_syncwarp() { __syncwarp(pz_gmask()); }
PZ_DEV unsigned pz_ballot(int p) { return __ballot_sync(pz_gmask(), p) >> pz_gshift(); }
PZ_DEV unsigned pz_match_any(unsigned v) { return __match_any_sync(pz_gmask(), v) >> pz_gshift(); }
PZ_DEV unsigned pz_lanemask_lt() { return (1u << pz_lane()) - 1u; }
PZ_DEV int pz_shfl(int v, int src) { return __shfl_sync(pz_gmask(), v, src, PZ_G); }
PZ_DEV bool pz_warp_any(bool p) { return __any_sync(0xffffffffu, p) != 0; } /* all groups of the warp */
PZ_DEV void pz_smem_inc(uint32_t *p) { atomicAdd(p, 1u); }
PZ_DEV int pz_popc(unsigned x) { return __popc(x); }
PZ_DEV int pz_ffs(unsigned x) { return __ffs((int)x); }
PZ_DEV uint32_t pz_funnel_r(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_r(lo, hi, s); }
PZ_DEV uint32_t pz_funnel_l(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_l(lo, hi, s); }
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
#define PZ_DIST_BITS 9  /* first-level bits of the distance LUT        */
#define PZ_PRE_BITS 7   /* the code-length code never exceeds 7 bits    */
#define PZ_RING_WORDS 128u /* staged input: four 128-byte quarters       */
#define PZ_QUARTER_WORDS 32u
#define PZ_QUARTER_BYTES 128u
#define PZ_QUARTER_SHIFT 10 /* log2(bits per quarter) */
#define PZ_MAX_LENS 464 /* 288 + 32 + 137 overshoot (Deflate.hs:124-156), padded */
#define PZ_WINDOW 131072u /* OutputWindow.hs:29-30 */
#define PZ_EXCESS 32768u  /* OutputWindow.hs:42-43 */
#define PZ_STEP_BITS 48u  /* most bits one literal/length + distance pair can consume: 15+5+15+13 */
#define PZ_MAX_IN_BYTES 0x1ffffff0ull /* bit positions are 32-bit: streams below 512 MiB */

/* Token queue, decoder -> writer (one per stream, in shared memory).  A token is one 32-bit
 * word: phase [31] | type [29,31) | payload.  The phase bit flips every time the ring wraps, so a
 * slot is valid exactly when its phase matches the reader's lap: one store publishes a token. */
#ifndef PZ_QSHIFT
#define PZ_QSHIFT 6
#endif
#define PZ_QLEN (1u << PZ_QSHIFT)
#define PZ_Q_LIT 0u   /* payload [0,8): the byte                                             */
#define PZ_Q_MATCH 1u /* payload [16,25): length 3..258, [0,15): distance - 1                 */
#define PZ_Q_CTRL 2u  /* payload [26,29): operation, followed by raw 31-bit argument tokens   */
#define PZ_C_NEWSTREAM 1u /* + stream index: the writer switches to that stream's output slice */
#define PZ_C_STORED 2u    /* + byte offset from the stream's first byte, + length (0..65535)  */
#define PZ_C_EXIT 3u      /* the decoder group has no streams left                             */
#define PZ_C_FLUSH 4u     /* nothing to do: a token that is never part of a batch, so the writer does not wait for more (lean kernel: drains) */
#define PZ_Q_MATCHD 3u /* lean kernel, hot lane only: a match whose distance the WRITER computes: payload [16,25) length, [0,16) the
                          distance LUT entry; the queue entry's second word holds the 32 stream bits behind the length code */
#define PZ_TOKEN(type, payload) (((uint32_t)(type) << 29) | (uint32_t)(payload))

/* LUT entry: total bits [0,5) | code bits [8,12) | 0 | type [13,15) | value [16,31) | literal flag 31.
 * total == 0 marks an entry the hot loop must not act on (long code, dead prefix, end of
 * block, or a symbol the reference cannot index): the careful path decides.
 * Fields that are shift AMOUNTS sit where `entry >> k` has them in its low five bits with a zero above (the funnel shifts
 * take their amount modulo 32): no mask instruction between the table and the shift -- every instruction of the hot trip
 * is 0.4 % of K1. */
#define PZ_T_LIT 0u
#define PZ_T_BASE 1u /* length / distance base + extra bits */
#define PZ_T_SLOW 3u
#define PZ_ENTRY(total, nbits, type, value) ((uint32_t)(total) | ((uint32_t)(nbits) << 8) | ((uint32_t)(type) << 13) | ((uint32_t)(value) << 16))
#define PZ_SLOW_ENTRY PZ_ENTRY(0, 0, PZ_T_SLOW, 0)
#define PZ_LIT_FLAG 0x80000000u
/* Distance LUT entry (16 bits): total bits [0,5) | extra bits [5,9) | 0 | m [10,12);
 * distance = 1 + (m << extra) + extra-bit value (Deflate.hs:199-237: symbols 0,1 have m = 0,1 and
 * every later pair of symbols m = 2,3); the code's own length is total - extra.  0 = not for the hot loop. */
#define PZ_DENTRY(nbits, extra, m) ((uint32_t)((nbits) + (extra)) | ((uint32_t)(extra) << 5) | ((uint32_t)(m) << 10))
#define PZ_DENTRY_DEFINED 1

/* Canonical description of one prefix code: enough for the bit-serial walker to reproduce
 * the reference trie's accept / "Advanced to empty tree!" behaviour (HuffmanTree.hs:73-83). */
struct PzTree {
  uint16_t cnt[16];  /* cnt[l]  = codes of length l                                     */
  uint16_t used[16]; /* used[l] = l-bit prefixes that lead to longer codes              */
  uint16_t nsyms;
  uint16_t pad;
};

/* Hand-over block between a stream's service group and its lane of the hot warp (shared
 * memory; one owner at a time, ownership moves with `state`). */
#define PZ_MS_SERVICE 0u /* the service group owns the stream                         */
#define PZ_MS_HOT 1u     /* posted: the hot lane decodes symbols                      */
#define PZ_MS_DEAD 2u    /* the service group has no streams left                     */
struct PzMail {
  uint32_t state;
  uint32_t bp, pos, base, lim, safe_end, qhead; /* travel with the ownership           */
  uint32_t hot_bp;  /* hot lane -> service: bit position at the end of its last trip   */
  uint32_t ring_hi; /* service -> hot lane: ring quarters < ring_hi are resident       */
  uint32_t mark;    /* block jobs: PzCtx::mark, travels with the ownership               */
  uint32_t pad[2];
};

/* Shared memory of one stream (one slot of the CTA). */
struct __attribute__((aligned(16))) PzStreamSmem {
  uint32_t lit_lut[1 << PZ_LIT_BITS];
  uint16_t dist_lut[1 << PZ_DIST_BITS]; /* compact entries; the 128-entry 32-bit precode LUT aliases it */
  uint32_t ring[PZ_RING_WORDS + 4];     /* + a copy of words 0..3 so that ring[i+1] never wraps */
  uint32_t scratch[32]; /* [0,16) per-length counters, [16,32) per-length offsets */
  uint16_t lit_perm[288];
  uint16_t dist_perm[176];
  uint16_t pre_perm[24];
  PzTree lit, dist, pre;
  uint8_t lens[PZ_MAX_LENS];
  uint2 q[PZ_QLEN];    /* token queue: written by the decoder side, read by the writer (x: the token, y: see PZ_Q_MATCHD) */
  uint32_t qtail;      /* tokens consumed so far: written by the writer, read by the decoder side */
  /* lean kernel: what the writer knows and the decoder side does not, published with qtail by ONE 16-byte store */
  uint32_t wpos;       /* bytes written */
  uint32_t wmark;      /* bytes written when the last match ended (PzCtx::mark) */
  uint32_t wbad;       /* != 0: a token failed the writer's checks (distance beyond the output, capacity, gap rule): the stream
                          is decoded again by the exact kernel, nothing of the token was written */
  PzMail mail;
  uint32_t pad1[20]; /* slot stride = 16 (mod 128) bytes: the same field of consecutive slots falls
                       into different banks when the hot warp's lanes (one per slot) read it */
};
/* One CTA per SM holds PZ_SLOTS = 28 streams: 28 slots must fit the 227 KiB a CTA may own. */
static_assert(sizeof(PzStreamSmem) * 28 <= 232448, "PzStreamSmem no longer fits 28 streams per SM");
static_assert(sizeof(PzStreamSmem) % 128 == 16, "slot stride must be 16 (mod 128) bytes");
#ifndef PZ_HOSTSIM
static_assert(offsetof(PzStreamSmem, qtail) % 16 == 0, "qtail / wpos / wmark / wbad are published by one 16-byte store");
#endif
static_assert(sizeof(uint16_t) * (1 << PZ_DIST_BITS) >= sizeof(uint32_t) * (1 << PZ_PRE_BITS), "precode LUT must fit the distance LUT");

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

/* ---- the batch a kernel launch works on ------------------------------------------------ */
struct PzJob {
  const uint8_t *in_blob;
  const uint64_t *in_off;
  uint8_t *out_blob; /* nullptr = sizing pass */
  const uint64_t *out_off;
  pz_result *res;
  uint32_t first, count; /* streams [first, first+count) */
  uint32_t skip_done;    /* res[s].status != PZ_ST_PENDING: the stored-stream kernels dealt with s */
  uint32_t *prog;        /* optional (host-mapped): prog[s] = decoded bytes of stream s that are final, in
                            32 KiB steps, 0xffffffff once the stream is finished: the host driver drains
                            finished parts of the output over PCIe while the kernel is still running */
  const uint32_t *in_ready; /* optional: streams [0, *in_ready) have their input in device memory; the
                               host driver raises it while it is still copying the rest of the batch */
  /* Block jobs (K4, one huge stream decoded block-parallel): when blk_start != nullptr, unit j of
   * [first, first+count) is ONE deflate block of stream blk_stream, beginning at bit blk_start[j]
   * of that stream.  It is decoded from its header to its end-of-block symbol with an unknown
   * history: the writers produce 16-bit symbols (a byte, or 256 + index into the 32 KiB that
   * precede the block) at out16 + blk_out[j].  res[j] receives the block's verdict. */
  const uint32_t *blk_start;
  const uint64_t *blk_out; /* element offsets into out16 (nullptr in the sizing pass)            */
  const uint32_t *blk_len; /* exact decoded length of block j (nullptr: blk_cap bounds every block) */
  uint16_t *out16;
  uint32_t blk_stream, blk_cap;
  /* Optional (K2 only): the Adler-32 segment table of the batch.  A stream K2 copies is checksummed
   * while it is copied: parts[seg_off[s] + j] receives (sum of the bytes, sum of byte * position in
   * the segment, both of segment j of the output) and res[s].adler_computed = PZ_ADLER_FUSED tells
   * K3 that the partial sums exist in that form. */
  uint2 *parts;
  const uint64_t *seg_off;
  /* Resumable contexts (decompressIncremental, Monad.hs:163-197: the decoder stops at NeedMore and goes
   * on when the next chunk arrives).  ckpt[4s..4s+4) receives where stream s can be picked up again when
   * it stops: {bit of the header of the block it is in, bit of the symbol it could not finish (0 = pick
   * up at the header, PZ_CK_TRAILER = every block is done, word 0 is the bit after the last one), bytes
   * decoded, bytes the reference has published}, bits counted from the stream's first byte.  resume[4s..]
   * (word 0 != 0) restarts stream s from such a checkpoint: the first `bytes decoded` bytes of its
   * output slice are its history.  pair_off = 1: stream s owns in_off[2s], in_off[2s+1] and out_off[2s],
   * out_off[2s+1] (begin, end) instead of sharing its end with the next stream's begin, so that streams
   * can live in buffers of their own (in_blob / out_blob may then be null and the offsets addresses). */
  const uint32_t *resume = nullptr;
  uint32_t *ckpt = nullptr;
  uint32_t pair_off = 0;
  /* Optional: a zeroed device counter.  Units are then CLAIMED in order by whichever slot is free instead of being dealt
   * out by index.  Block jobs: blocks differ in length, and with two or three units per slot the longest deal sets the
   * time of the launch.  Batches behind K2 / K5: what those kernels left is no longer spread evenly over the slots (a
   * corpus whose every fourth record is a dynamic one leaves ALL of K1's work to the CTAs whose index is 3 mod 4). */
  uint32_t *next_unit = nullptr;
  /* Framing of every stream of the job (PZ_FRAME_*).  zlib (RFC 1950) is the reference's; gzip members (RFC 1952)
   * and raw deflate are the extension its README names as the first TODO (README.md:42-50). */
  uint32_t framing = 0;
};
#define PZ_FRAME_ZLIB 0u
#define PZ_FRAME_GZIP 1u /* header: magic, CM, FLG, MTIME/XFL/OS, FEXTRA / FNAME / FCOMMENT / FHCRC skipped; trailer: CRC-32, ISIZE (little-endian) */
#define PZ_FRAME_RAW 2u  /* no header, no trailer */
#define PZ_CK_TRAILER 0xffffffffu
/* four resume words denote a checkpoint unless they are all "nothing yet" (no header bit, no symbol, no byte decoded): the
 * first block of a raw deflate stream starts at bit 0, so word 0 alone cannot tell */
#define PZ_HAS_CKPT(rs) (((rs)[0] | (rs)[1] | (rs)[2]) != 0u)
#define PZ_ADLER_FUSED 0xffffffffu /* no Adler-32 value: both halves of one are below 65521 */
#define PZ_BLK_BIAS 65536u /* a block job counts its output from here: the window model then always
                              sees at least 32 KiB of history, as it would inside a long stream */
#define PZ_PROG_SHIFT 15
#define PZ_PROG_DONE 0xffffffffu
#define PZ_ST_PENDING (-1)

/* ---- per-stream decoder state (registers; identical in every lane of the group) -------- */
enum PzMode { PZ_M_IDLE = 0, PZ_M_HDR = 1, PZ_M_SYMS = 2, PZ_M_FAST = 3, PZ_M_DEAD = 4, PZ_M_WAIT = 5 /* posted to the hot lane */,
              PZ_M_DRAIN = 6 /* lean kernel: back from the hot lane, waiting for the writer's byte count */,
              PZ_M_FINDRAIN = 7 /* lean kernel: verdict reached, waiting for the writer to confirm every token */ };

struct PzCtx {
  const uint8_t *in_al; /* input, rounded down to 16 bytes                                */
  uint32_t in_al_bytes; /* bytes readable from in_al (multiple of 16)                      */
  uint32_t start_bit;   /* first bit of the stream, counted from in_al                     */
  uint32_t end_bit;     /* first bit past the stream, counted from in_al                   */
  uint32_t safe_end;    /* bp <= safe_end: a whole symbol pair (PZ_STEP_BITS) is available */
  uint32_t bp;          /* bit position of the reader, counted from in_al; bp <= end_bit   */
  uint32_t q;           /* ring quarter holding bp; quarters q and q+1 are resident        */
  uint32_t next_q;      /* first quarter not requested yet (q+3 in steady state)           */
  bool pending;         /* a quarter requested while the hot lane owns the stream has not been awaited */
  bool starved;         /* idle because the next stream's input has not reached the device yet */
  bool block_job;       /* the unit is one deflate block (PzJob::blk_start), not a zlib stream     */
  bool lean;            /* lean kernel: the hot lane does not count bytes, the writer does (pz_hot_warp_lean) */
  bool flush_sent;      /* lean kernel, draining: the FLUSH token is in the queue */
  uint32_t framing;     /* PzJob::framing */
  uint32_t pos;  /* bytes decoded                                                          */
  uint32_t base; /* bytes the reference would already have published (multiple of 32 KiB)  */
  uint32_t mark; /* block jobs: pos at the last moveWindow call (after a match, at a block's start).  The bytes
                    between two calls ("gap") decide whether the reference's window can overflow: while every gap is
                    <= 32 KiB the fill stays below 96 KiB + 258 whatever it was at the block's start, so a block job,
                    which does not know the real fill, is exact; a longer gap makes the job give up (K4 declines) */
  uint32_t cap;
  int32_t status, detail;
  int64_t p0, p1;
  /* state machine */
  uint32_t mode;
  uint32_t bfinal;
  uint32_t adler_stored;
  bool fixed_ready;
  bool need_careful; /* the hot loop met something only the careful path may decide */
  uint32_t next;     /* next stream index of this group */
  pz_result *res;
  /* checkpoint (PzJob::ckpt) */
  uint32_t hdr_bp;     /* bit position (from in_al) of the header of the current block */
  uint32_t sym_bp;     /* bit position of the symbol the careful path is deciding, 0 outside the symbol loop */
  uint32_t resume_sym; /* resuming inside a block: the symbol to go on with once the block's tables exist (from the stream's first byte) */
  uint32_t *ck;        /* this stream's four checkpoint words, or nullptr */
  /* token queue (decoder side) */
  uint32_t qhead;  /* tokens pushed so far */
  uint32_t qtailc; /* last value read from the writer's counter */
#ifdef PZ_HOSTSIM
  struct PzWriter *hw; /* host build: tokens are applied at once */
#endif
};

PZ_DEV void pz_fail(PzCtx &c, int status, int detail, int64_t p0 = 0, int64_t p1 = 0) {
  c.status = status; c.detail = detail; c.p0 = p0; c.p1 = p1;
}

/* ---- staged input ---------------------------------------------------------------------
 * Quarter k = bytes [128k, 128k+128) of in_al, staged into ring slot (k & 3) with cp.async
 * (16 bytes per request).  Invariant: quarters q and q+1 are resident, q+2 is in flight. */
PZ_DEV void pz_ring_issue(PzCtx &c, PzStreamSmem *sm, uint32_t k) {
  uint32_t *dst = sm->ring + (k & 3u) * PZ_QUARTER_WORDS;
  for (uint32_t p = (uint32_t)pz_lane(); p < PZ_QUARTER_BYTES / 16u; p += PZ_G) {
    uint32_t bo = k * PZ_QUARTER_BYTES + p * 16u;
    if (bo < c.in_al_bytes) {
      pz_copy16_async(dst + p * 4u, c.in_al + bo);
      if (((k & 3u) | p) == 0u) pz_copy16_async(sm->ring + PZ_RING_WORDS, c.in_al + bo);
    }
    /* past the stream: never consumed (every reader counts bits first), left as is */
  }
}
PZ_COLD void pz_cross(PzCtx &c, PzStreamSmem *sm) {
  while (c.q != (c.bp >> PZ_QUARTER_SHIFT)) {
    c.q++;
    pz_async_wait_all(); /* quarter q+1 was requested one crossing ago */
    pz_syncwarp();
    while (c.next_q < c.q + 3u) pz_ring_issue(c, sm, c.next_q++);
  }
}
PZ_DEV void pz_seek(PzCtx &c, PzStreamSmem *sm, uint32_t bit) {
  uint32_t k = bit >> PZ_QUARTER_SHIFT;
  pz_async_wait_all(); /* nothing may still be landing in the slots reused below */
  pz_syncwarp();
  pz_ring_issue(c, sm, k);
  pz_ring_issue(c, sm, k + 1u);
  pz_async_wait_all();
  pz_syncwarp();
  pz_ring_issue(c, sm, k + 2u);
  c.q = k; c.next_q = k + 3u; c.bp = bit;
}
/* The 32 stream bits starting at bit position bp (bits past end_bit are garbage). */
PZ_DEV uint32_t pz_peek(const uint32_t *ring, uint32_t bp) {
  uint32_t t = (bp >> 5) & (PZ_RING_WORDS - 1u);
  return pz_funnel_r(ring[t], ring[t + 1u], bp);
}
/* The 64 stream bits starting at bp: enough for a whole literal/length + distance pair. */
PZ_DEV void pz_peek64(const uint32_t *ring, uint32_t bp, uint32_t &lo, uint32_t &hi) {
  uint32_t t = (bp >> 5) & (PZ_RING_WORDS - 1u);
  uint32_t w0 = ring[t], w1 = ring[t + 1u], w2 = ring[t + 2u];
  lo = pz_funnel_r(w0, w1, bp);
  hi = pz_funnel_r(w1, w2, bp);
}
/* The 96 stream bits starting at bp (the hot loop's register window). */
PZ_DEV void pz_peek96(const uint32_t *ring, uint32_t bp, uint32_t &b0, uint32_t &b1, uint32_t &b2) {
  uint32_t t = (bp >> 5) & (PZ_RING_WORDS - 1u);
  uint32_t w0 = ring[t], w1 = ring[t + 1u], w2 = ring[t + 2u], w3 = ring[t + 3u];
  b0 = pz_funnel_r(w0, w1, bp);
  b1 = pz_funnel_r(w1, w2, bp);
  b2 = pz_funnel_r(w2, w3, bp);
}
/* Bits [bp+32, bp+96): the part of the register window that is NOT on the serial chain. */
PZ_DEV void pz_peek_tail(const uint32_t *ring, uint32_t bp, uint32_t &b1, uint32_t &b2) {
  uint32_t t = (bp >> 5) & (PZ_RING_WORDS - 1u);
  uint32_t w1 = ring[t + 1u], w2 = ring[t + 2u], w3 = ring[t + 3u];
  b1 = pz_funnel_r(w1, w2, bp);
  b2 = pz_funnel_r(w2, w3, bp);
}
PZ_DEV void pz_advance(PzCtx &c, PzStreamSmem *sm, uint32_t n) {
  c.bp += n;
  if ((c.bp >> PZ_QUARTER_SHIFT) != c.q) pz_cross(c, sm);
}
PZ_DEV uint32_t pz_avail(const PzCtx &c) { return c.end_bit - c.bp; }
/* nextBits n (Monad.hs:199-230), n <= 16: running past the input is the truncation verdict
 * (Zlib.hs:38-39). */
PZ_DEV bool pz_take(PzCtx &c, PzStreamSmem *sm, uint32_t n, uint32_t &v) {
  if (pz_avail(c) < n) { pz_fail(c, PZ_ERR_DECOMPRESSION, PZ_D_RAN_OUT); return false; }
  v = pz_peek(sm->ring, c.bp) & ((1u << n) - 1u);
  pz_advance(c, sm, n);
  return true;
}
/* advanceToByte (Monad.hs:303-307): the dropped bits belong to a byte that was already
 * fetched, so this is never a truncation. */
PZ_DEV void pz_align_byte(PzCtx &c, PzStreamSmem *sm) {
  uint32_t drop = (8u - (c.bp & 7u)) & 7u;
  if (drop) pz_advance(c, sm, drop);
}

/* ---- exact bit-serial walk (nextCode / advanceTree, Monad.hs:295-302, HuffmanTree.hs:73-83)
 * Consumes one bit per step and stops exactly where the reference's trie walk stops:
 * truncation if the input ends first, "Advanced to empty tree!" on an unused prefix. */
PZ_COLD int pz_walk(PzCtx &c, PzStreamSmem *sm, const PzTree *t, const uint16_t *perm) {
  const uint32_t av = pz_avail(c);
  const uint32_t w = pz_peek(sm->ring, c.bp); /* only the first min(av, 15) bits are looked at */
  uint32_t code = 0, first = 0, index = 0;
  for (uint32_t len = 1; len <= 15u; len++) {
    if (len > av) { pz_advance(c, sm, len - 1u); pz_fail(c, PZ_ERR_DECOMPRESSION, PZ_D_RAN_OUT); return -1; }
    if (t->nsyms == 0) { pz_advance(c, sm, len); pz_fail(c, PZ_ERR_HUFFMAN_TREE, PZ_D_ADVANCE_EMPTY_TREE); return -1; }
    code |= (w >> (len - 1u)) & 1u;
    uint32_t count = t->cnt[len];
    if (code - first < count) { pz_advance(c, sm, len); return perm[index + (code - first)]; }
    index += count; first += count;
    if (code - first >= t->used[len]) { pz_advance(c, sm, len); break; } /* used[15] == 0 */
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
    if (sym < 256u) return PZ_ENTRY(nbits, nbits, PZ_T_LIT, sym) | PZ_LIT_FLAG;
    if (sym == 256u) return PZ_SLOW_ENTRY; /* end of block: the careful path ends it */
    if (sym > 285u) return PZ_SLOW_ENTRY;  /* lengthArray ! 286/287 is a bounds error */
    return PZ_ENTRY(nbits + PZ_LEN_EXTRA[sym - 257u], nbits, PZ_T_BASE, PZ_LEN_BASE[sym - 257u]);
  }
  if (sym > 29u) return 0u; /* distanceArray ! >=30 is a bounds error */
  return PZ_DENTRY(nbits, PZ_DIST_EXTRA[sym], sym < 2u ? sym : 2u + (sym & 1u));
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
  for (int p = pz_lane(); p < (int)t->nsyms; p += PZ_G) {
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
PZ_COLD int pz_tree_error(const uint8_t *lens, int n, const PzTree *t, const uint16_t *perm, uint16_t *codes, int64_t *val) {
  pz_canon_codes(lens, t, perm, codes, true);
  for (int i = n - 1; i >= 0; i--) {
    int li = lens[i];
    if (!li) continue;
    uint32_t ci = codes[i];
    for (int jb = i + 1; jb < n; jb += PZ_G) {
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
 * cooperatively by the group.  Returns 0, or the HuffmanTreeError detail with *val. */
template <int BITS, int KIND, typename LutT>
PZ_COLD int pz_build(const uint8_t *lens, int n, PzTree *t, uint16_t *perm, LutT *lut, uint32_t *scratch, int64_t *val) {
  uint32_t *cnt32 = scratch, *offs = scratch + 16;
  const int lane = pz_lane();
  pz_syncwarp();
  for (int i = lane; i < 16; i += PZ_G) cnt32[i] = 0;
  pz_syncwarp();
  for (int i = lane; i < n; i += PZ_G) {
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
  /* canonical code of the symbol at sorted position p with length l: cnt32[l] + p
   * (computeCodeValues, Deflate.hs:261-288: next_code[l] minus the position of the first l-bit symbol) */
  uint32_t code = 0, nshort = 0;
#pragma unroll
  for (int l = 1; l <= 15; l++) {
    code = (code + (l > 1 ? cl[l - 1] : 0u)) << 1;
    if (lane == 0) { t->cnt[l] = (uint16_t)cl[l]; offs[l] = acc; cnt32[l] = code - acc; }
    acc += cl[l];
    if (l <= BITS) nshort = acc;
  }
  if (lane == 0) t->nsyms = (uint16_t)acc;
  pz_syncwarp();
  /* stable counting sort by length: perm[] */
  for (int b = 0; b < n; b += PZ_G) {
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
  /* LUT: every entry starts as "not for the hot loop" (long code or dead prefix: the careful path
   * decides), then each code of at most BITS bits fills the entries that end in its reversed bits */
  for (uint32_t e = (uint32_t)lane; e < (1u << BITS); e += PZ_G) lut[e] = (LutT)(KIND == 2 ? 0u : PZ_SLOW_ENTRY);
  pz_syncwarp();
  for (uint32_t p = (uint32_t)lane; p < nshort; p += PZ_G) {
    const uint32_t sym = perm[p];
    const uint32_t l = lens[sym];
    const uint32_t rev = pz_brev(cnt32[l] + p) >> (32u - l);
    const LutT entry = (LutT)pz_make_entry<KIND>(sym, l);
    for (uint32_t e = rev; e < (1u << BITS); e += 1u << l) lut[e] = entry;
  }
  pz_syncwarp();
  return 0;
}

/* The distance a PZ_DENTRY d and the 32 stream bits wd at its code stand for: 1 + (m << extra) + the extra bits' value.
 * sx = d >> 5 has the number of extra bits in its low five bits, d - sx the code's length (total - extra, modulo 32): both
 * go into funnel shifts as they are. */
PZ_DEV uint32_t pz_dist_value(uint32_t d, uint32_t wd) {
  const uint32_t sx = d >> 5;
  const uint32_t ex = pz_funnel_r(wd, 0u, d - sx) & ~pz_funnel_l(0u, 0xffffffffu, sx);
  return 1u + pz_funnel_l(0u, d >> 10, sx) + ex;
}

/* ---- output side: the writer ---------------------------------------------------------------- */
#ifdef PZ_HOSTSIM
PZ_DEV void pz_st8_if(bool p, uint8_t *a, uint32_t v) { if (p) *a = (uint8_t)v; }
PZ_DEV uint32_t pz_ld8_if(bool p, const uint8_t *a) { return p ? *a : 0u; }
PZ_DEV void pz_syncwarp_all() {}
PZ_DEV uint32_t pz_vload(const uint32_t *p) { return *p; }
PZ_DEV void pz_vstore(uint32_t *p, uint32_t v) { *p = v; }
PZ_DEV void pz_backoff() {}
PZ_DEV void pz_fence_cta() {}
#else
/* The writer's hot-loop global accesses are volatile asm WITHOUT a memory clobber: they keep
 * their order among themselves (which is all the copy semantics need), while the compiler
 * stays free to move shared-memory loads across them. */
PZ_DEV void pz_st8_if(bool p, uint8_t *a, uint32_t v) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %0, 0;\n\t@q st.global.u8 [%1], %2;\n\t}" ::"r"((int)p), "l"(a), "r"(v));
}
PZ_DEV uint32_t pz_ld8_if(bool p, const uint8_t *a) {
  /* a fresh register with no other definition: nothing may read (and so wait for) the loaded
   * byte before its store, and the store carries the same predicate as this load */
  uint32_t v;
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t@q ld.global" PZ_LD_MOD ".u8 %0, [%2];\n\t}" : "=r"(v) : "r"((int)p), "l"(a));
  return v;
}
PZ_DEV uint32_t pz_ld16_if(bool p, const uint16_t *a) {
  uint32_t v;
  asm volatile("{\n\t.reg .pred q;\n\t.reg .b16 h;\n\tmov.b16 h, 0;\n\tsetp.ne.b32 q, %1, 0;\n\t@q ld.global" PZ_LD_MOD ".u16 h, [%2];\n\tcvt.u32.u16 %0, h;\n\t}" : "=r"(v) : "r"((int)p), "l"(a));
  return v;
}
/* warp barrier between the stores and the loads of one iteration (all 32 lanes are converged
 * in the hot loops); same ordering rule as above */
PZ_DEV void pz_syncwarp_all() { asm volatile("bar.warp.sync 0xffffffff;"); }
PZ_DEV uint32_t pz_vload(const uint32_t *p) { return *(const volatile uint32_t *)p; }
PZ_DEV void pz_vstore(uint32_t *p, uint32_t v) { *(volatile uint32_t *)p = v; }
/* Orders this thread's earlier shared-memory accesses before its later ones as seen by the other
 * warps of the CTA (the hand-over of a stream between a service group and its hot lane).  An
 * acquire-release fence: __threadfence_block() is fence.sc.cta, which costs a MEMBAR.SC plus a
 * drain of the shared-memory pipe on every hand-over. */
PZ_DEV void pz_fence_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }
/* v = *p (shared memory, volatile) in the lanes where c holds, as ONE predicated load: no branch */
PZ_DEV void pz_vload_if(bool c, const uint32_t *p, uint32_t &v) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t@q ld.volatile.shared.u32 %0, [%2];\n\t}"
               : "+r"(v) : "r"((int)c), "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
}
PZ_DEV void pz_backoff() { __nanosleep(64); }
#endif

/* The LZ77 copy (OutputWindow.hs:82-101: copyChunked = byte-serial replicate semantics) of
 * len bytes from dist back, all lanes of the group, no deferral: the general case. */
/* Symbol i of a block job's output (i relative to the block; negative = one of the 32768 bytes
 * before it, which only the resolution pass knows: a marker 256 + index stands in for it). */
#define PZ_MARK(i) ((uint16_t)(256 + 32768 + (i)))
PZ_DEV uint16_t pz_sym16(const uint16_t *out16, int32_t i) { return i < 0 ? PZ_MARK(i) : out16[i]; }

template <bool WIDE>
PZ_DEV void pz_copy_match(uint8_t *out, uint16_t *out16, uint32_t pos, uint32_t len, uint32_t dist) {
  const uint32_t lane = (uint32_t)pz_wlane();
  uint8_t *dst = out + pos;
  const uint8_t *src = dst - dist;
  const int32_t s16 = (int32_t)pos - (int32_t)dist;
  pz_wsyncwarp(); /* earlier stores by other lanes are ordered before the loads below */
  if (dist >= len) {
    for (uint32_t i = lane; i < len; i += PZ_WG) {
      if (WIDE) out16[pos + i] = pz_sym16(out16, s16 + (int32_t)i);
      else dst[i] = src[i];
    }
  } else if (dist >= PZ_WG) {
    for (uint32_t i0 = 0; i0 < len; i0 += PZ_WG) { /* each chunk may read the previous one */
      uint32_t i = i0 + lane;
      if (i < len) {
        if (WIDE) out16[pos + i] = pz_sym16(out16, s16 + (int32_t)i);
        else dst[i] = src[i];
      }
      pz_wsyncwarp();
    }
  } else { /* dist < PZ_WG and dist < len: replicate the dist-byte pattern */
    uint32_t m = lane % dist;
    const uint32_t step = PZ_WG % dist;
    for (uint32_t i = lane; i < len; i += PZ_WG) {
      if (WIDE) out16[pos + i] = pz_sym16(out16, s16 + (int32_t)m);
      else dst[i] = src[m];
      m += step;
      if (m >= dist) m -= dist;
    }
  }
  pz_wsyncwarp();
}

/* Writer state of one stream slot (registers; identical in every lane of the group). */
struct PzWriter {
  const PzJob *job;
  uint8_t *out;      /* output slice of the current stream */
  uint16_t *out16;   /* block jobs: 16-bit symbols of the current block */
  const uint8_t *in; /* its first compressed byte (stored runs copy from the input) */
  uint32_t pos;      /* bytes written */
  uint32_t op, need, a0; /* control message being assembled: `need` argument tokens to go */
  uint32_t sidx, pub;    /* current stream; 32 KiB steps of it already published in job->prog */
  bool exited;
  /* lean kernel: the writer checks what the hot lane cannot (it does not count bytes) */
  uint32_t cap;  /* capacity of the current stream's output slice */
  uint32_t mark; /* pos when the last match ended (PzCtx::mark, seen from here: block ends are not tokens, so gaps look longer) */
  bool bad;      /* a token failed a check: the rest of the stream's tokens are dropped, the exact kernel decodes it again */
};
#ifdef PZ_HOSTSIM
PZ_DEV void pz_publish(PzWriter &, uint32_t) {}
#else
/* Tells the host driver that the first `value` bytes of the current stream are final. */
PZ_DEV void pz_publish(PzWriter &w, uint32_t value) {
  if (w.job->prog == nullptr || w.out == nullptr) return;
  pz_wsyncwarp();
  __threadfence_system(); /* the bytes first, then the word that announces them */
  if (pz_wlane() == 0) *(volatile uint32_t *)(w.job->prog + w.sidx) = value;
}
#endif
PZ_DEV void pz_writer_init(PzWriter &w, const PzJob *job) {
  w.job = job; w.out = nullptr; w.out16 = nullptr; w.in = nullptr; w.pos = 0; w.op = 0; w.need = 0; w.a0 = 0; w.sidx = 0; w.pub = 0; w.exited = false;
  w.cap = 0; w.mark = 0; w.bad = false;
}

/* Applies one token completely (no deferral): the writer's path for everything that is not a
 * literal or a short disjoint match.  `raw` is the token without its phase bit. */
template <bool WIDE, bool LEAN = false>
PZ_DEV void pz_writer_apply(PzWriter &w, uint32_t raw) {
  const uint32_t lane = (uint32_t)pz_wlane();
  if (w.need) { /* argument of a control message */
    if (w.op == PZ_C_NEWSTREAM) {
      if (LEAN) {
        const uint64_t o0 = w.job->out_off[raw << w.job->pair_off], o1 = w.job->out_off[(raw << w.job->pair_off) + 1u];
        w.cap = o1 - o0 > 0xfffdff00ull ? 0xfffdff00u : (uint32_t)(o1 - o0);
        w.mark = 0; w.bad = false;
      }
      pz_publish(w, PZ_PROG_DONE); /* the previous stream of this slot is complete */
      w.sidx = raw; w.pub = 0;
      if (WIDE) {
        w.out16 = w.job->out16 + w.job->blk_out[raw];
        w.in = w.job->in_blob + w.job->in_off[w.job->blk_stream];
      } else {
        w.out = w.job->out_blob + w.job->out_off[raw << w.job->pair_off];
        w.in = w.job->in_blob + w.job->in_off[raw << w.job->pair_off];
      }
      w.pos = 0;
      if (!WIDE && w.job->resume != nullptr && PZ_HAS_CKPT(w.job->resume + 4u * raw)) { /* the stream goes on behind its history */
        w.pos = w.job->resume[4u * raw + 2u];
        w.pub = w.pos >> PZ_PROG_SHIFT;
      }
      w.need = 0;
    } else if (w.need == 2u) {
      w.a0 = raw; w.need = 1;
    } else { /* emitBlock (Monad.hs:317-322): raw bytes of a stored block */
      if (LEAN) { /* (the service group checked this block with exact counters; only the gap rule can differ here) */
        if (!w.bad && (raw > w.cap - w.pos || w.pos + raw - w.mark > PZ_EXCESS)) w.bad = true;
        if (w.bad) { w.need = 0; return; }
      }
      const uint8_t *src = w.in + w.a0;
      uint8_t *dst = w.out + w.pos;
      pz_wsyncwarp();
      for (uint32_t i = lane; i < raw; i += PZ_WG) {
        if (WIDE) w.out16[w.pos + i] = src[i];
        else dst[i] = src[i];
      }
      pz_wsyncwarp();
      w.pos += raw; w.need = 0;
      if (LEAN) w.mark = w.pos; /* the block's end is a moveWindow call */
    }
    return;
  }
  const uint32_t type = (raw >> 29) & 3u;
  if (type == PZ_Q_LIT) {
    if (LEAN) {
      if (!w.bad && (w.pos >= w.cap || w.pos + 1u - w.mark > PZ_EXCESS)) w.bad = true;
      if (w.bad) return;
    }
    if (lane == 0) {
      if (WIDE) w.out16[w.pos] = (uint16_t)(raw & 0xffu);
      else w.out[w.pos] = (uint8_t)raw;
    }
    w.pos++;
  } else if (type == PZ_Q_MATCH) {
    const uint32_t len = (raw >> 16) & 0x1ffu, dist = (raw & 0x7fffu) + 1u;
    if (LEAN) { /* what pz_fast_trip checks with its own counters: distance inside what exists, room, and the gap rule of PzCtx::mark */
      if (!w.bad && (dist > w.pos || len > w.cap - w.pos || w.pos + len - w.mark > PZ_EXCESS)) w.bad = true;
      if (w.bad) return;
    }
    pz_copy_match<WIDE>(w.out, w.out16, w.pos, len, dist);
    w.pos += len;
    if (LEAN) w.mark = w.pos;
  } else {
    const uint32_t op = (raw >> 26) & 7u;
    if (op == PZ_C_EXIT) { pz_publish(w, PZ_PROG_DONE); w.exited = true; }
    else if (op == PZ_C_FLUSH) { /* nothing: its consumption is the message (pz_lean_drain) */ }
    else { w.op = op; w.need = op == PZ_C_NEWSTREAM ? 1u : 2u; }
  }
}

#ifndef PZ_HOSTSIM
/* The byte of a trip's output: a literal carried by the token word `inf` (bit 31 set, byte in
 * [0,8)) or the byte `x` loaded from the history.  Select and store are one asm statement so the
 * compiler cannot pull the use of `x` (and with it the wait for the load) ahead of later loads. */
PZ_DEV void pz_st8_sel(bool p, uint8_t *a, uint32_t inf, uint32_t x) {
  asm volatile(
      "{\n\t.reg .pred q, l;\n\t.reg .b32 t, u;\n\tsetp.ne.b32 q, %0, 0;\n\tsetp.lt.s32 l, %2, 0;\n\t"
      "and.b32 u, %2, 255;\n\tselp.b32 t, u, %3, l;\n\t@q st.global" PZ_ST_MOD ".u8 [%1], t;\n\t}" ::"r"((int)p),
      "l"(a), "r"(inf), "r"(x));
}
/* 16-bit flavour (block jobs): `alt` (a literal byte or a marker for a byte before the block) wins
 * over the loaded symbol `x` when bit 31 of alt is set. */
PZ_DEV void pz_st16_sel(bool p, uint16_t *a, uint32_t alt, uint32_t x) {
  asm volatile(
      "{\n\t.reg .pred q, l;\n\t.reg .b32 t, u;\n\t.reg .b16 h;\n\tsetp.ne.b32 q, %0, 0;\n\tsetp.lt.s32 l, %2, 0;\n\t"
      "and.b32 u, %2, 65535;\n\tselp.b32 t, u, %3, l;\n\tcvt.u16.u32 h, t;\n\t@q st.global" PZ_ST_MOD ".u16 [%1], h;\n\t}" ::"r"((int)p),
      "l"(a), "r"(alt), "r"(x));
}

/* The writer warp: runs until every group has seen its EXIT token.
 *
 * A trip looks at the next PZ_WG tokens of every stream of the warp, one token per lane.  The
 * longest prefix of literals and short, disjoint matches whose sources lie entirely before the
 * trip's first output byte (and whose bytes total at most PZ_TRIP_BYTES) forms the batch.  Its
 * output bytes are then dealt out to the lanes by POSITION: lane l produces bytes l, l+WG, l+2WG,
 * ... of the batch, finding the token that owns a byte with a popcount over the bitmap of token
 * start offsets and one shuffle.  All history loads of the trip are issued before the first
 * store, so one L2/HBM round trip is shared by the whole batch, and consecutive lanes touch
 * consecutive bytes.  Everything else (overlapping or long copies, stored runs, control tokens)
 * goes through pz_writer_apply(), one token per trip.
 *
 * The trip is one long dependent sequence (queue read, votes, prefix sum, bitmap, loads, stores:
 * about 2 000 cycles), and a stream cannot have two trips in flight, so the tokens one trip may
 * take bound the writer's rate per stream: with 8 lanes per stream the writers, not the hot warp,
 * set the pace of the kernel (ncu, profiles/); hence 16 lanes and up to 128 bytes per trip. */
#define PZ_TRIP_BYTES (8u * PZ_WG)
#define PZ_ROUNDS 8
#define PZ_BM_WORDS ((int)(PZ_TRIP_BYTES / 32u)) /* words of the token-start bitmap */
#define PZ_WGROUPS (32 / PZ_WG)                 /* streams per writer warp */
#define PZ_WGMASK ((1u << PZ_WG) - 1u)
/* qtail, wpos, wmark, wbad of a slot in one 16-byte store (lean kernel): the service group reads them together */
PZ_DEV void pz_wstat_store(PzStreamSmem *sm, bool p, uint32_t tail, uint32_t pos, uint32_t mark, uint32_t bad) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %0, 0;\n\t@q st.volatile.shared.v4.u32 [%1], {%2, %3, %4, %5};\n\t}" ::"r"((int)p),
               "r"((unsigned)__cvta_generic_to_shared(&sm->qtail)), "r"(tail), "r"(pos), "r"(mark), "r"(bad) : "memory");
}
template <bool WIDE, bool LEAN = false>
PZ_DEV void pz_writer_warp(const PzJob &job, PzStreamSmem *sm, bool present) {
  static_assert(PZ_WG == 8 || PZ_WG == 16, "the writer deals bytes to groups of 8 or 16 lanes");
  PzWriter w;
  pz_writer_init(w, &job);
  w.exited = !present;
  const uint32_t lane = (uint32_t)pz_wlane();
  const unsigned gsh = pz_wgshift();
  const uint32_t grp = gsh / PZ_WG;
  const unsigned all = 0xffffffffu; /* the groups run this loop converged: warp-wide collectives, group-sized segments */
  uint32_t tail = 0, naps = 0;
  if (!pz_warp_any(!w.exited)) return;
  for (;;) {
    pz_syncwarp_all(); /* the previous trip's stores are visible to the other lanes' loads */
    const uint32_t idx = tail + lane;
    uint32_t raw, len, dist;
    bool is_lit, is_match;
    if (LEAN) { /* queue entries are 8 bytes here: a match of the hot lane carries the distance entry and the bits behind the length code */
      uint32_t y;
      asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(raw), "=r"(y) : "r"((unsigned)__cvta_generic_to_shared(&sm->q[idx & (PZ_QLEN - 1u)])));
      const uint32_t t = (raw >> 29) & 3u;
      len = (raw >> 16) & 0x1ffu;
      const uint32_t dd = pz_dist_value(raw & 0xffffu, y);
      dist = t == PZ_Q_MATCHD ? dd : (raw & 0x7fffu) + 1u;
      if (t == PZ_Q_MATCHD) raw = (raw & 0x80000000u) | PZ_TOKEN(PZ_Q_MATCH, (len << 16) | (dist - 1u)); /* from here on an ordinary match token */
      is_lit = t == PZ_Q_LIT; is_match = t == PZ_Q_MATCH || t == PZ_Q_MATCHD;
    } else {
      raw = pz_vload(&sm->q[idx & (PZ_QLEN - 1u)].x);
      const uint32_t t = (raw >> 29) & 3u;
      len = (raw >> 16) & 0x1ffu; dist = (raw & 0x7fffu) + 1u;
      is_lit = t == PZ_Q_LIT; is_match = t == PZ_Q_MATCH;
    }
    const bool valid = !w.exited && (raw >> 31) == ((idx >> PZ_QSHIFT) & 1u);
    const bool fast = valid && w.need == 0u && !(LEAN && w.bad) && (is_lit || (is_match && dist >= len && len <= PZ_TRIP_BYTES));
    /* A trip costs the same whether it moves one token or PZ_WG per group, and the other warps
     * need the issue slots: unless some group has a full batch waiting (or a token that will not
     * join a batch anyway), sleep a little -- but never for long. */
    const unsigned vmask = (__ballot_sync(all, valid) >> gsh) & PZ_WGMASK;
    const unsigned smask = (__ballot_sync(all, valid && !fast) >> gsh) & PZ_WGMASK;
    if (naps < 8u && !pz_warp_any(vmask == PZ_WGMASK || smask != 0u)) { naps++; __nanosleep(PZ_NAP_NS); continue; }
    const uint32_t L = fast ? (is_lit ? 1u : len) : 0u;
    uint32_t E = L; /* inclusive prefix sum over the group: end offset of this lane's token */
#pragma unroll
    for (int o = 1; o < PZ_WG; o <<= 1) {
      const uint32_t t = __shfl_up_sync(all, E, o, PZ_WG);
      if ((int)lane >= o) E += t;
    }
    /* a match may not read what this trip produces: its source ends at E - dist <= 0 */
    bool ok = fast && (is_lit || dist >= E) && E <= PZ_TRIP_BYTES;
    /* lean kernel: the checks of pz_fast_trip that need the byte position -- the distance lies inside what exists, the token
     * fits the caller's buffer, and the gap rule (PzCtx::mark) with the trip's tokens taken as one gap.  A token that fails
     * ends the batch before it; it then comes first in a later trip and pz_writer_apply() decides. */
    if (LEAN) ok = ok && (is_lit || dist <= w.pos + (E - L)) && E <= w.cap - w.pos && w.pos + E - w.mark <= PZ_EXCESS;
    const unsigned bad = (__ballot_sync(all, !ok) >> gsh) & PZ_WGMASK;
    const uint32_t n = bad ? (uint32_t)__ffs((int)bad) - 1u : (uint32_t)PZ_WG;
    const bool slow = n == 0u && (vmask & 1u);
    naps = 0;
    const uint32_t r0 = (uint32_t)__shfl_sync(all, (int)raw, 0, PZ_WG);
    if (pz_warp_any(slow)) {
      if (slow) {
        pz_writer_apply<WIDE, LEAN>(w, r0 & 0x7fffffffu);
        tail++;
        if (LEAN) pz_wstat_store(sm, lane == 0u, tail, w.pos, w.mark, w.bad ? 1u : 0u);
        else pz_vstore(&sm->qtail, tail);
        if (!w.exited && (w.pos >> PZ_PROG_SHIFT) != w.pub) { w.pub = w.pos >> PZ_PROG_SHIFT; pz_publish(w, w.pub << PZ_PROG_SHIFT); }
      }
      if (!pz_warp_any(!w.exited)) break;
      continue;
    }
    const bool act = lane < n;
    const uint32_t P = E - L; /* first output byte of this lane's token, relative to the trip */
    const uint32_t Bn = (uint32_t)__shfl_sync(all, (int)E, (int)n - 1, PZ_WG);
    const uint32_t B = n ? Bn : 0u;
    /* bitmap of the tokens' first bytes, per group: one warp-wide OR per (group, word) */
    uint32_t bm[PZ_BM_WORDS];
#pragma unroll
    for (int k = 0; k < PZ_BM_WORDS; k++) bm[k] = 0u;
    const uint32_t my_bit = act ? (1u << (P & 31u)) : 0u;
#pragma unroll
    for (int g = 0; g < PZ_WGROUPS; g++) {
#pragma unroll
      for (int k = 0; k < PZ_BM_WORDS; k++) {
        const uint32_t v = __reduce_or_sync(all, (grp == (uint32_t)g && (P >> 5) == (uint32_t)k) ? my_bit : 0u);
        if (grp == (uint32_t)g) bm[k] = v;
      }
    }
    uint32_t cnt[PZ_BM_WORDS]; /* tokens starting in the words before word k, minus one */
    cnt[0] = 0xffffffffu;
#pragma unroll
    for (int k = 1; k < PZ_BM_WORDS; k++) cnt[k] = cnt[k - 1] + (uint32_t)__popc(bm[k - 1]);
    const uint32_t info = is_lit ? (0x80000000u | (raw & 0xffu)) : dist;
    const uint32_t max_b = __reduce_max_sync(0xffffffffu, B);
    uint8_t *const base = w.out + w.pos;
    uint32_t x[PZ_ROUNDS], inf[PZ_ROUNDS];
#ifdef PZ_EXP_NO_COPY /* timing experiment only: the batch is consumed, its bytes are not produced */
    if (false)
#endif
#pragma unroll
    for (int r = 0; r < PZ_ROUNDS; r++) {
      if ((uint32_t)(r * PZ_WG) < max_b) { /* warp-uniform */
        const uint32_t b = (uint32_t)(r * PZ_WG) + lane;
        const int k = (r * PZ_WG) / 32; /* the word of byte b: the same for every lane of the round */
        const int t = (int)(cnt[k] + (uint32_t)__popc(bm[k] & ((2u << (b & 31u)) - 1u)));
        inf[r] = (uint32_t)__shfl_sync(all, (int)info, t, PZ_WG);
        if (WIDE) {
          /* a source before the block is not loaded: its marker takes the literal's place */
          const int32_t si = (int32_t)(w.pos + b) - (int32_t)(inf[r] & 0xffffu);
          const bool lit = (int32_t)inf[r] < 0;
          x[r] = pz_ld16_if(b < B && !lit && si >= 0, w.out16 + si);
          if (!lit && si < 0) inf[r] = 0x80000000u | PZ_MARK(si);
        } else {
          x[r] = pz_ld8_if(b < B && (int32_t)inf[r] >= 0, base + (int32_t)(b - (inf[r] & 0xffffu)));
        }
      }
    }
#if defined(PZ_EXP_NO_COPY) || defined(PZ_EXP_NO_STORE) /* (NO_STORE: the loads stay, kept alive by a store that never happens) */
    if (max_b == 0x7fffffffu)
#endif
#pragma unroll
    for (int r = 0; r < PZ_ROUNDS; r++) {
      if ((uint32_t)(r * PZ_WG) < max_b) {
        const uint32_t b = (uint32_t)(r * PZ_WG) + lane;
        if (WIDE) pz_st16_sel(b < B, w.out16 + w.pos + b, inf[r], x[r]);
        else pz_st8_sel(b < B, base + b, inf[r], x[r]);
      }
    }
    if (LEAN) { /* the last match of the batch ends a gap */
      const unsigned mm = (__ballot_sync(all, act && !is_lit) >> gsh) & PZ_WGMASK;
      const uint32_t e_last = (uint32_t)__shfl_sync(all, (int)E, mm ? 31 - __clz((int)mm) : 0, PZ_WG);
      if (mm) w.mark = w.pos + e_last;
    }
    w.pos += B;
    tail += n;
    if (LEAN) pz_wstat_store(sm, lane == 0u, tail, w.pos, w.mark, w.bad ? 1u : 0u);
    else pz_vstore(&sm->qtail, tail);
    if (job.prog != nullptr && pz_warp_any((w.pos >> PZ_PROG_SHIFT) != w.pub)) {
      if ((w.pos >> PZ_PROG_SHIFT) != w.pub) { w.pub = w.pos >> PZ_PROG_SHIFT; pz_publish(w, w.pub << PZ_PROG_SHIFT); }
    }
  }
}
#endif /* !PZ_HOSTSIM */

/* ---- decoder side: emitting tokens --------------------------------------------------------- */
/* Pushes one token (31 bits, phase added here); waits while the queue is full. */
template <bool COUNT_ONLY>
PZ_DEV void pz_push(PzCtx &c, PzStreamSmem *sm, uint32_t v) {
#ifdef PZ_HOSTSIM
  (void)sm;
  if (!COUNT_ONLY) pz_writer_apply<false>(*c.hw, v);
#else
  if (!COUNT_ONLY) {
    while (c.qhead - c.qtailc >= PZ_QLEN) {
      c.qtailc = pz_vload(&sm->qtail);
      if (c.qhead - c.qtailc >= PZ_QLEN) pz_backoff();
    }
    pz_vstore(&sm->q[c.qhead & (PZ_QLEN - 1u)].x, v | (((c.qhead >> PZ_QSHIFT) & 1u) << 31));
    c.qhead++;
  }
#endif
}

/* moveWindow / emitExcess (Monad.hs:338-347, OutputWindow.hs:45-54): at most one 32 KiB
 * chunk leaves the window per call, only once 64 KiB have accumulated. */
PZ_DEV void pz_move_window(PzCtx &c) {
  if (c.pos - c.base >= 2u * PZ_EXCESS) c.base += PZ_EXCESS;
}

/* emitPastChunk (Monad.hs:324-333, OutputWindow.hs:82-101).  Returns false with the
 * verdict set when the reference would fault or the caller's buffer is full. */
template <bool COUNT_ONLY>
PZ_DEV bool pz_match(PzCtx &c, PzStreamSmem *sm, uint32_t len, uint32_t dist) {
  uint32_t fill = c.pos - c.base;
  if (dist > fill) { pz_fail(c, PZ_REF_BOTTOM, PZ_D_BOT_DIST_TOO_FAR, dist, fill); return false; }
  if (fill + len > PZ_WINDOW) { pz_fail(c, PZ_REF_BOTTOM, PZ_D_BOT_WINDOW_OVERFLOW); return false; }
  if (len > c.cap - c.pos) { pz_fail(c, PZ_OUTPUT_FULL, 0); return false; }
  if (c.block_job && c.pos + len - c.mark > PZ_EXCESS) { pz_fail(c, PZ_REF_BOTTOM, PZ_D_BOT_WINDOW_OVERFLOW); return false; } /* PzCtx::mark */
  pz_push<COUNT_ONLY>(c, sm, PZ_TOKEN(PZ_Q_MATCH, (len << 16) | (dist - 1u)));
  c.pos += len;
  c.mark = c.pos;
  pz_move_window(c);
  return true;
}

/* emitByte (Monad.hs:309-315, OutputWindow.hs:64-68) with the window / capacity checks. */
template <bool COUNT_ONLY>
PZ_DEV bool pz_literal_checked(PzCtx &c, PzStreamSmem *sm, uint32_t b) {
  if (c.pos - c.base >= PZ_WINDOW) { pz_fail(c, PZ_REF_BOTTOM, PZ_D_BOT_WINDOW_OVERFLOW); return false; }
  if (c.pos >= c.cap) { pz_fail(c, PZ_OUTPUT_FULL, 0); return false; }
  if (c.block_job && c.pos + 1u - c.mark > PZ_EXCESS) { pz_fail(c, PZ_REF_BOTTOM, PZ_D_BOT_WINDOW_OVERFLOW); return false; } /* PzCtx::mark */
  pz_push<COUNT_ONLY>(c, sm, PZ_TOKEN(PZ_Q_LIT, b & 0xffu));
  c.pos++;
  return true;
}

/* ---- the careful symbol: exact verdict order, used near the end of the input or of the
 * output and whenever the LUT cannot answer (Deflate.hs:106-120).  Returns 1 = continue,
 * 0 = end of block, -1 = verdict set. */
template <bool COUNT_ONLY>
PZ_DEV int pz_symbol_careful(PzCtx &c, PzStreamSmem *sm) {
  c.sym_bp = c.bp; /* checkpoint: nothing of this symbol is committed until it is complete */
  int sym = pz_walk(c, sm, &sm->lit, sm->lit_perm);
  if (sym < 0) return -1;
  if (sym < 256) return pz_literal_checked<COUNT_ONLY>(c, sm, (uint32_t)sym) ? 1 : -1;
  if (sym == 256) return 0;
  if (sym > 285) { pz_fail(c, PZ_REF_BOTTOM, PZ_D_BOT_LENGTH_SYM, sym); return -1; }
  uint32_t ex;
  if (!pz_take(c, sm, PZ_LEN_EXTRA[sym - 257], ex)) return -1;
  const uint32_t len = PZ_LEN_BASE[sym - 257] + ex;
  int ds = pz_walk(c, sm, &sm->dist, sm->dist_perm);
  if (ds < 0) return -1;
  if (ds > 29) { pz_fail(c, PZ_REF_BOTTOM, PZ_D_BOT_DIST_SYM, ds); return -1; }
  if (!pz_take(c, sm, PZ_DIST_EXTRA[ds], ex)) return -1;
  return pz_match<COUNT_ONLY>(c, sm, len, PZ_DIST_BASE[ds] + ex) ? 1 : -1;
}

/* ---- the decoder's hot loop (runInflate, Deflate.hs:106-120) --------------------------------
 * Every group in FAST mode decodes one symbol (a literal, or a length/distance pair) per
 * iteration and pushes its token.  The body is straight-line predicated code: the groups of a
 * warp take different "branches" (literal / match) in the same iteration, and a lone warp per
 * scheduler cannot afford branch latencies.
 *
 * The bit-position chain (window -> LUT -> bits -> next window) is the only serial part, so the
 * NEXT symbol's window and LUT entry are requested as soon as this symbol's bit count is
 * known -- speculatively: if this symbol turns out to be one the loop must not decide, the loop
 * ends and the look-ahead is dropped.
 *
 * The loop ends (at the end of a four-symbol trip) once ANY group of the warp has met something
 * it must not decide here (long code, end of block, end of input or output in sight, a verdict,
 * a full token queue): that group consumed nothing of it.  Groups that are not in FAST mode (no
 * streams left) idle along. */
struct PzFast { /* the registers of the hot loop */
  uint32_t bp, pos, base, lim, safe_end, qhead, qtailc, mark;
  uint32_t b0, b1, b2, e; /* the 96 stream bits at bp and the literal/length LUT entry of b0 (decoded ahead) */
  bool live;
#ifdef PZ_HOSTSIM
  PzWriter *hw;
#endif
};

/* The two table look-ups of the symbol chain.  On the device the address is ONE multiply-add behind the mask (index * size +
 * table base in the shared window) instead of the shift / mask / add ptxas derives from an array subscript: every
 * instruction between the two dependent loads of a symbol is four to five cycles of the chain. */
#if defined(PZ_HOSTSIM) || defined(PZ_NO_OPT_MAD) /* A/B on the GPU: 7.65 ms without, 7.50 ms with (config 2) */
PZ_DEV uint32_t pz_lit_at(const PzStreamSmem *sm, uint32_t bits) { return sm->lit_lut[bits & ((1u << PZ_LIT_BITS) - 1u)]; }
PZ_DEV uint32_t pz_dist_at(const PzStreamSmem *sm, uint32_t bits) { return sm->dist_lut[bits & ((1u << PZ_DIST_BITS) - 1u)]; }
#else
PZ_DEV uint32_t pz_lit_at(const PzStreamSmem *sm, uint32_t bits) {
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm->lit_lut);
  uint32_t a, v;
  asm("mad.lo.u32 %0, %1, 4, %2;" : "=r"(a) : "r"(bits & ((1u << PZ_LIT_BITS) - 1u)), "r"(base));
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
PZ_DEV uint32_t pz_dist_at(const PzStreamSmem *sm, uint32_t bits) {
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm->dist_lut);
  uint32_t a, v;
  asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(a) : "r"(bits & ((1u << PZ_DIST_BITS) - 1u)), "r"(base));
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
#endif

PZ_DEV void pz_fast_fetch(PzFast &f, const PzStreamSmem *sm, uint32_t bp) {
  pz_peek96(sm->ring, bp, f.b0, f.b1, f.b2);
  f.e = pz_lit_at(sm, f.b0);
}

/* One trip = PZ_TRIP symbols.  The serial chain of a stream is
 *     b0 -> literal/length LUT -> shift -> distance LUT -> shift -> b0' -> literal/length LUT ...
 * i.e. two dependent shared-memory loads and five ALU operations per symbol: the lane keeps the
 * 96 stream bits at bp in registers (b0, b1, b2), a symbol consumes at most 48 of them, so the
 * next 32-bit window b0' comes out of the registers by two funnel shifts.  The ring is re-read
 * only for the upper 64 bits of the next window, and nothing waits for that until the NEXT
 * symbol's first shift: the reload runs beside the chain instead of inside it.
 *
 * The chain runs SPECULATIVELY: it never waits for the verdict on the symbol it has just consumed.
 * The verdict (`alive`, sticky within the trip) only gates what is committed: the token, and the
 * position registers in `f`.  Once a symbol fails -- something the loop must not decide -- the chain keeps running on garbage for the rest of the trip (all table and ring
 * indices are masked, so that is harmless) and nothing more is committed.  Returns true if the
 * stream stopped inside this trip; f.b0/b1/b2/e are then stale. */
/* Symbols per trip.  Measured (profiles/r02aa_*, r02ab_*): 2 / 3 / 4 / 5 / 6 symbols give 146 / 154 / 140 / 134 / 132 GB/s on config 2.
 * The trip is straight-line code (80 instructions per symbol) issued by ONE warp, and ncu shows `no_inst` at every eighth
 * instruction of the four-symbol trip (325 instructions = 5.2 KB): it does not fit the instruction cache next to the
 * scheduler, three symbols (3.9 KB) do.  (Four was the optimum while the loop around the trip still cost two votes and two
 * convergence regions per trip, round 1; with one vote and one branch the shorter trip wins.) */
#ifndef PZ_TRIP
#define PZ_TRIP 3
#endif
template <bool COUNT_ONLY, bool BLK = false>
PZ_DEV bool pz_fast_trip(PzFast &f, PzStreamSmem *sm, const bool run) {
  uint32_t bp = f.bp, pos = f.pos, base = f.base, qhead = f.qhead, mark = f.mark;
  uint32_t b0 = f.b0, b1 = f.b1, b2 = f.b2, e = f.e;
  bool alive = run;
#ifndef PZ_HOSTSIM
  const uint32_t qbase = (uint32_t)__cvta_generic_to_shared(sm->q);
#endif
#pragma unroll
  for (int k = 0; k < PZ_TRIP; k++) {
    const uint32_t tb = e & 31u;
    const bool is_lit = (int32_t)e < 0;
    /* the funnel shifts take their amount modulo 32: e itself serves as tb (tb <= 20) */
    const uint32_t wd = pz_funnel_r(b0, b1, e);  /* the 32 bits after the literal/length symbol */
    const uint32_t wd1 = pz_funnel_r(b1, b2, e); /* and the 32 after those (off the chain)     */
    const uint32_t d = pz_dist_at(sm, wd);
    const uint32_t tb2 = d & (is_lit ? 0u : 31u); /* <= 28 */
    const uint32_t nb0 = pz_funnel_r(wd, wd1, tb2);
    const uint32_t ne = pz_lit_at(sm, nb0);
    const uint32_t nbp = bp + tb + tb2;
    uint32_t nb1, nb2;
    pz_peek_tail(sm->ring, nbp, nb1, nb2);
    /* off the chain: the symbol's values and its verdict */
    const uint32_t len = (e >> 16) + pz_funnel_r(b0 & ~(0xffffffffu << tb), 0u, e >> 8); /* (e >> 8) mod 32 = the code's bits */
    const uint32_t dm1 = pz_dist_value(d, wd) - 1u; /* the token carries distance - 1; the + 1 and - 1 fold away */
    const uint32_t room = f.lim - pos;
    /* (room in the token queue for the whole trip is the caller's to check: one test per trip instead of one per symbol) */
    const uint32_t adv = is_lit ? 1u : len;
    /* the bytes of the symbol fit (one for a literal).  Two spellings of the same test: ptxas makes the faster trip out of
     * the first in the decode kernel (6.67 against 6.69 ms) and out of the second in the sizing pass (5.38 against 5.75 ms) */
    const bool pre_ok = bp <= f.safe_end && (COUNT_ONLY ? room != 0u : adv <= room);
    /* dist <= pos - base (OutputWindow.hs:82-89) is dist <= pos here: base only ever moves when 64 KiB
     * have accumulated, so base > 0 implies pos - base >= 32 KiB >= any distance */
    const bool m_ok = tb != 0u && (d & 31u) != 0u && dm1 < pos && (COUNT_ONLY ? len <= room : true);
    alive = alive && pre_ok && (is_lit || m_ok);
    if (BLK) { /* PzCtx::mark: a gap of more than 32 KiB between two moveWindow calls is the careful path's to refuse */
      alive = alive && pos + adv - mark <= PZ_EXCESS;
      mark = is_lit ? mark : pos + adv;
    }
    if (!COUNT_ONLY) {
      const uint32_t tok = is_lit ? PZ_TOKEN(PZ_Q_LIT, (e >> 16) & 0xffu) : PZ_TOKEN(PZ_Q_MATCH, (len << 16) | dm1);
#ifdef PZ_HOSTSIM
      if (alive) pz_writer_apply<false>(*f.hw, tok);
#else
      { /* the queue slot's address is one multiply-add behind the mask, the store is predicated (no branch in the trip) */
        uint32_t a;
        asm("mad.lo.u32 %0, %1, 8, %2;" : "=r"(a) : "r"(qhead & (PZ_QLEN - 1u)), "r"(qbase));
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %0, 0;\n\t@q st.volatile.shared.u32 [%1], %2;\n\t}" ::"r"((int)alive), "r"(a),
                     "r"(tok | ((qhead << (31 - PZ_QSHIFT)) & 0x80000000u)) : "memory");
      }
      qhead += alive ? 1u : 0u;
#endif
    }
    pos += adv;
    if (!is_lit && pos - base >= 2u * PZ_EXCESS) base += PZ_EXCESS; /* moveWindow after every match */
    bp = nbp; b0 = nb0; b1 = nb1; b2 = nb2; e = ne;
    if (alive) { f.bp = bp; f.pos = pos; f.base = base; f.qhead = qhead; if (BLK) f.mark = mark; }
  }
  if (alive) { f.b0 = b0; f.b1 = b1; f.b2 = b2; f.e = e; } /* else: unchanged if the trip did not run, stale if it stopped */
  return run && !alive;
}

template <bool COUNT_ONLY>
PZ_DEV void pz_fast_loop(PzCtx &c, PzStreamSmem *sm) {
  PzFast f;
  f.live = c.mode == PZ_M_FAST;
  f.bp = c.bp; f.pos = c.pos; f.base = c.base; f.safe_end = c.safe_end;
  f.qhead = c.qhead; f.qtailc = c.qtailc; f.mark = c.mark;
#ifdef PZ_HOSTSIM
  f.hw = c.hw;
#endif
  /* first position this run may not write at: a stale `base` only makes it conservative */
  f.lim = f.base + PZ_WINDOW;
  if (c.cap < f.lim) f.lim = c.cap;
  pz_fast_fetch(f, sm, f.bp);
  bool stop;
  for (;;) {
    /* PZ_TRIP x 48 bits stay inside the resident quarters */
    stop = c.block_job ? pz_fast_trip<COUNT_ONLY, true>(f, sm, f.live) : pz_fast_trip<COUNT_ONLY, false>(f, sm, f.live);
    if (f.live && (f.bp >> PZ_QUARTER_SHIFT) != c.q) { c.bp = f.bp; pz_cross(c, sm); }
    if (pz_warp_any(stop)) break;
  }
  if (f.live) {
    c.bp = f.bp; c.pos = f.pos; c.base = f.base; c.qhead = f.qhead; c.mark = f.mark;
    c.mode = PZ_M_SYMS;
    c.need_careful = true;
  }
}

#ifndef PZ_HOSTSIM
/* ---- the hot warp ---------------------------------------------------------------------------
 * ONE warp per CTA runs the symbol loop of every stream of the CTA, one lane per stream slot:
 * the per-symbol chain (window -> LUT -> bits -> next window) is serial and the same for every
 * stream, so this is the cheapest way to issue it.  A lane works on a stream only while its
 * service group has posted it (PzMail.state == PZ_MS_HOT); as soon as the lane meets something
 * the loop must not decide (long code, end of block, end of input or output in sight, a verdict)
 * it writes the stream's position back and returns the ownership; the other lanes carry on.  The
 * lane only READS the staged input: the service group keeps the ring filled, following the bit
 * position the lane publishes after every trip. */
template <bool COUNT_ONLY, bool BLK>
PZ_DEV void pz_hot_warp(PzStreamSmem *slots, uint32_t n_slots) {
  const uint32_t lane = threadIdx.x & 31u;
  PzStreamSmem *sm = slots + (lane < n_slots ? lane : 0u);
  PzFast f;
  f.live = false;
  f.bp = 0; f.pos = 0; f.base = 0; f.lim = 0; f.safe_end = 0; f.qhead = 0; f.qtailc = 0; f.mark = 0; f.b0 = 0; f.b1 = 0; f.b2 = 0; f.e = 0;
  bool dead = lane >= n_slots;
  /* A lone warp pays every branch in full (nothing else issues on its scheduler while one resolves: ncu attributes a third
   * of a trip's cycles to the code AROUND the symbols -- the loop's back edge, two votes and two convergence regions for
   * "a posting has arrived" and "a stream has stopped", profiles/r02g_*).  So a trip has ONE vote and ONE warp-uniform branch:
   * the mailbox of a lane without a stream is read at the top (a predicated load whose latency the symbols hide), and
   * whatever needs attention -- a stream that stopped, a posting that arrived, a service group that has no streams left --
   * is looked at behind the trip, together.  While nothing is live the trips run empty; that is the idle loop. */
  for (;;) {
    uint32_t st = PZ_MS_SERVICE;
    pz_vload_if(!f.live && !dead, &sm->mail.state, st);
    /* PZ_TRIP symbols per trip: the input the trip can touch (PZ_TRIP x 48 bits + the 128-bit
     * look-ahead) lies in quarters q and q+1, which must be resident; a lane whose input is late
     * idles this trip */
    const uint32_t ring_hi = pz_vload(&sm->mail.ring_hi);
    if (!COUNT_ONLY) f.qtailc = pz_vload(&sm->qtail);
    /* ... and so does a lane whose token queue might not take a whole trip's tokens (the writer is behind) */
    const bool run = f.live && (f.bp >> PZ_QUARTER_SHIFT) + 1u < ring_hi && (COUNT_ONLY || f.qhead - f.qtailc <= PZ_QLEN - PZ_TRIP);
    const bool stop = pz_fast_trip<COUNT_ONLY, BLK>(f, sm, run);
    if (run) pz_vstore(&sm->mail.hot_bp, f.bp);
    const bool pick = st == PZ_MS_HOT, died = st == PZ_MS_DEAD;
    if (__any_sync(0xffffffffu, stop || pick || died)) {
      if (stop) { /* hand the stream back: the careful path decides the next symbol */
        pz_vstore(&sm->mail.bp, f.bp); pz_vstore(&sm->mail.pos, f.pos); pz_vstore(&sm->mail.base, f.base);
        pz_vstore(&sm->mail.qhead, f.qhead);
        if (BLK) pz_vstore(&sm->mail.mark, f.mark);
        pz_fence_cta();
        pz_vstore(&sm->mail.state, PZ_MS_SERVICE);
        f.live = false;
      }
      if (pick) {
        pz_fence_cta();
        f.bp = pz_vload(&sm->mail.bp); f.pos = pz_vload(&sm->mail.pos); f.base = pz_vload(&sm->mail.base);
        f.lim = pz_vload(&sm->mail.lim); f.safe_end = pz_vload(&sm->mail.safe_end); f.qhead = pz_vload(&sm->mail.qhead);
        if (BLK) f.mark = pz_vload(&sm->mail.mark);
        pz_fast_fetch(f, sm, f.bp);
        f.live = true;
      }
      dead = dead || died;
      if (__all_sync(0xffffffffu, dead)) break;
    }
  }
}


/* ---- the LEAN hot warp (plain batches: no resume, no block jobs, no sizing pass) ------------------------------------------
 * The symbol chain of a stream needs the bit position and the two tables, nothing else.  Everything the legacy trip computes
 * beside it to decide verdicts -- bytes produced, the window's base, room left in the window and in the caller's buffer, "is
 * the distance inside what exists" -- needs the BYTE position, and the writer knows that anyway.  So here the lane decodes
 * tokens and stops only for what the bit stream itself says (a table entry that is not for the loop, the end of the input in
 * sight), and the writer checks every token against its byte position (pz_writer_warp<.., true>): a token that fails leaves
 * the stream marked for the exact kernel, which decodes it again from scratch and words the verdict.  For the streams of a
 * well-formed batch nothing ever fails, and the lane's trip is 40 % shorter: 6 symbols fit where 4 did (the trip has to stay
 * inside the 6 KiB instruction cache of its scheduler: a lone warp pays every miss in full).
 * The distance of a match is left to the writer too (PZ_Q_MATCHD): the lane stores the distance entry and the bits behind the
 * length code, one 8-byte queue entry per symbol. */
#ifndef PZ_LEAN_TRIP
#define PZ_LEAN_TRIP 6
#endif
struct PzLean { /* the registers of the lean loop */
  uint32_t bp, safe_end, qhead;
  uint32_t b0, b1, b2, e;
  bool live;
};
PZ_DEV void pz_lean_fetch(PzLean &f, const PzStreamSmem *sm) {
  pz_peek96(sm->ring, f.bp, f.b0, f.b1, f.b2);
  f.e = pz_lit_at(sm, f.b0);
}
/* One trip: PZ_LEAN_TRIP symbols, speculative like pz_fast_trip (the chain never waits for the verdict on the symbol it has
 * just consumed; `alive`, sticky, only gates what is committed).  Returns true if the stream stopped inside the trip. */
PZ_DEV bool pz_lean_trip(PzLean &f, PzStreamSmem *sm, const bool run) {
  uint32_t bp = f.bp, qhead = f.qhead;
  uint32_t b0 = f.b0, b1 = f.b1, b2 = f.b2, e = f.e;
  bool alive = run;
  const uint32_t qbase = (uint32_t)__cvta_generic_to_shared(sm->q);
#pragma unroll
  for (int k = 0; k < PZ_LEAN_TRIP; k++) {
    const uint32_t tb = e & 31u;
    const bool is_lit = (int32_t)e < 0;
    const uint32_t wd = pz_funnel_r(b0, b1, e);
    const uint32_t wd1 = pz_funnel_r(b1, b2, e);
    const uint32_t d = pz_dist_at(sm, wd);
    const uint32_t tb2 = d & (is_lit ? 0u : 31u);
    const uint32_t nb0 = pz_funnel_r(wd, wd1, tb2);
    const uint32_t ne = pz_lit_at(sm, nb0);
    const uint32_t nbp = bp + tb + tb2;
    uint32_t nb1, nb2;
    pz_peek_tail(sm->ring, nbp, nb1, nb2);
    /* off the chain: the token */
    const uint32_t len = (e >> 16) + pz_funnel_r(b0 & ~(0xffffffffu << tb), 0u, e >> 8);
    alive = alive && tb != 0u && (is_lit || (d & 31u) != 0u) && bp <= f.safe_end;
    const uint32_t tok = (is_lit ? PZ_TOKEN(PZ_Q_LIT, (e >> 16) & 0xffu) : (PZ_TOKEN(PZ_Q_MATCHD, len << 16) | d)) | ((qhead << (31 - PZ_QSHIFT)) & 0x80000000u);
    {
      uint32_t a;
      asm("mad.lo.u32 %0, %1, 8, %2;" : "=r"(a) : "r"(qhead & (PZ_QLEN - 1u)), "r"(qbase));
      asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %0, 0;\n\t@q st.volatile.shared.v2.u32 [%1], {%2, %3};\n\t}" ::"r"((int)alive), "r"(a), "r"(tok), "r"(wd) : "memory");
    }
    qhead += alive ? 1u : 0u;
    bp = nbp; b0 = nb0; b1 = nb1; b2 = nb2; e = ne;
    if (alive) { f.bp = bp; f.qhead = qhead; }
  }
  if (alive) { f.b0 = b0; f.b1 = b1; f.b2 = b2; f.e = e; }
  return run && !alive;
}

PZ_DEV void pz_hot_warp_lean(PzStreamSmem *slots, uint32_t n_slots) {
  const uint32_t lane = threadIdx.x & 31u;
  PzStreamSmem *sm = slots + (lane < n_slots ? lane : 0u);
  PzLean f;
  f.live = false;
  f.bp = 0; f.safe_end = 0; f.qhead = 0; f.b0 = 0; f.b1 = 0; f.b2 = 0; f.e = 0;
  bool dead = lane >= n_slots;
  for (;;) { /* the same straight-line mailbox poll as pz_hot_warp */
    uint32_t st = PZ_MS_SERVICE;
    pz_vload_if(!f.live && !dead, &sm->mail.state, st);
    const bool pick = st == PZ_MS_HOT;
    dead = dead || st == PZ_MS_DEAD;
    const bool any_pick = __any_sync(0xffffffffu, pick), any_live = __any_sync(0xffffffffu, f.live);
    if (any_pick) {
      if (pick) {
        pz_fence_cta();
        f.bp = pz_vload(&sm->mail.bp); f.safe_end = pz_vload(&sm->mail.safe_end); f.qhead = pz_vload(&sm->mail.qhead);
        pz_lean_fetch(f, sm);
        f.live = true;
      }
    } else if (!any_live) {
      if (__all_sync(0xffffffffu, dead)) break;
      __nanosleep(100);
    }
    const uint32_t ring_hi = pz_vload(&sm->mail.ring_hi);
    const uint32_t qtail = pz_vload(&sm->qtail);
    /* a trip needs its input resident (quarters q and q + 1) and room for all its tokens: a lane without either idles this trip */
    const bool run = f.live && (f.bp >> PZ_QUARTER_SHIFT) + 1u < ring_hi && f.qhead - qtail <= PZ_QLEN - PZ_LEAN_TRIP;
    const bool stop = pz_lean_trip(f, sm, run);
    if (run) pz_vstore(&sm->mail.hot_bp, f.bp);
    if (__any_sync(0xffffffffu, stop)) {
      if (stop) { /* hand the stream back: the careful path decides the next symbol, once the writer has said where the bytes stand */
        pz_vstore(&sm->mail.bp, f.bp); pz_vstore(&sm->mail.qhead, f.qhead);
        pz_fence_cta();
        pz_vstore(&sm->mail.state, PZ_MS_SERVICE);
        f.live = false;
      }
    }
  }
}

/* Service group: posts its stream to the hot lane.  Everything requested from the ring has
 * landed before the ownership moves. */
PZ_DEV void pz_post_hot(PzCtx &c, PzStreamSmem *sm) {
  pz_async_wait_all();
  pz_syncwarp();
  if (pz_lane() == 0) {
    uint32_t lim = c.base + PZ_WINDOW; /* first position this run may not write at */
    if (c.cap < lim) lim = c.cap;
    pz_vstore(&sm->mail.bp, c.bp); pz_vstore(&sm->mail.pos, c.pos); pz_vstore(&sm->mail.base, c.base);
    pz_vstore(&sm->mail.lim, lim); pz_vstore(&sm->mail.safe_end, c.safe_end); pz_vstore(&sm->mail.qhead, c.qhead);
    pz_vstore(&sm->mail.hot_bp, c.bp); pz_vstore(&sm->mail.ring_hi, c.next_q); pz_vstore(&sm->mail.mark, c.mark);
    pz_fence_cta();
    pz_vstore(&sm->mail.state, PZ_MS_HOT);
  }
  pz_syncwarp();
  c.pending = false;
  c.mode = PZ_M_WAIT;
}

/* Service group while the hot lane owns the stream: one poll.  Takes the stream back if the lane
 * has returned it, otherwise keeps the ring ahead of the lane (quarters up to hot_q + 3; the slot
 * of quarter k is the slot of quarter k - 4, which the lane has left for good). */
PZ_DEV void pz_service_poll(PzCtx &c, PzStreamSmem *sm) {
  const uint32_t st = pz_vload(&sm->mail.state);
  if (c.pending) {
    pz_async_wait_all();
    pz_syncwarp();
    pz_fence_cta();
    if (pz_lane() == 0) pz_vstore(&sm->mail.ring_hi, c.next_q);
    c.pending = false;
  }
  if (st == PZ_MS_SERVICE) {
    pz_fence_cta();
    c.bp = pz_vload(&sm->mail.bp); c.pos = pz_vload(&sm->mail.pos); c.base = pz_vload(&sm->mail.base);
    c.qhead = pz_vload(&sm->mail.qhead);
    if (c.block_job) c.mark = pz_vload(&sm->mail.mark);
    c.q = c.bp >> PZ_QUARTER_SHIFT;
    /* the reader's invariant again: q and q+1 resident, q+2 requested */
    if (c.next_q < c.q + 2u) {
      while (c.next_q < c.q + 2u) pz_ring_issue(c, sm, c.next_q++);
      pz_async_wait_all();
      pz_syncwarp();
    }
    while (c.next_q < c.q + 3u) pz_ring_issue(c, sm, c.next_q++);
    c.mode = PZ_M_SYMS;
    c.need_careful = true;
    if (c.lean) { c.mode = PZ_M_DRAIN; c.flush_sent = false; } /* bytes, base and mark are the writer's to tell (the mailbox words are stale) */
    return;
  }
  const uint32_t hq = pz_vload(&sm->mail.hot_bp) >> PZ_QUARTER_SHIFT;
  if (c.next_q <= hq + 3u) {
    pz_ring_issue(c, sm, c.next_q++);
    c.pending = true;
  }
}
#endif /* !PZ_HOSTSIM */

#ifndef PZ_HOSTSIM
/* Lean kernel: the service group asks the writer where the bytes stand.  A FLUSH token goes behind whatever is queued (it
 * never joins a batch, so the writer does not wait for more tokens), and once the writer has consumed it the slot's status
 * block holds the byte count, the end of the last match and the verdict of the writer's checks.
 *   PZ_M_DRAIN      the stream came back from the hot lane: the careful path needs the counters
 *   PZ_M_FINDRAIN   a verdict is ready: it is only published once every token has passed the writer's checks
 * A failed check sends the stream to the exact kernel (status PENDING: the launch behind this one decodes it from scratch). */
PZ_DEV void pz_finish(PzCtx &c);
PZ_DEV bool pz_lean_drain(PzCtx &c, PzStreamSmem *sm) { /* true: still waiting for the writer */
  if (!c.flush_sent) {
    pz_push<false>(c, sm, PZ_TOKEN(PZ_Q_CTRL, PZ_C_FLUSH << 26));
    c.flush_sent = true;
  }
  uint32_t tail, wpos, wmark, wbad;
  asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(tail), "=r"(wpos), "=r"(wmark), "=r"(wbad)
               : "r"((unsigned)__cvta_generic_to_shared(&sm->qtail)) : "memory");
  if (tail != c.qhead) return true;
  c.flush_sent = false;
  if (wbad != 0u) {
    pz_syncwarp();
    if (pz_lane() == 0) c.res->status = PZ_ST_PENDING;
    c.mode = PZ_M_IDLE;
    return false;
  }
  if (c.mode == PZ_M_FINDRAIN) { pz_finish(c); return false; }
  c.pos = wpos;
  if (wmark > c.mark) c.mark = wmark;
  c.base = c.mark >= 2u * PZ_EXCESS ? (c.mark / PZ_EXCESS - 1u) * PZ_EXCESS : 0u; /* closed form of the window's base while every gap is <= 32 KiB (PzCtx::mark) */
  c.mode = PZ_M_SYMS;
  c.need_careful = true;
  return false;
}
#endif

/* One symbol of the code-length code (getCodeLengths, Deflate.hs:124-156). */
PZ_DEV int pz_pre_symbol(PzCtx &c, PzStreamSmem *sm, const uint32_t *pre_lut) {
  if (pz_avail(c) >= (uint32_t)PZ_PRE_BITS) { /* all peeked bits are real */
    uint32_t e = pre_lut[pz_peek(sm->ring, c.bp) & ((1u << PZ_PRE_BITS) - 1u)];
    if ((e & 31u) != 0u) { pz_advance(c, sm, e & 31u); return (int)((e >> 16) & 0xffu); }
  }
  return pz_walk(c, sm, &sm->pre, sm->pre_perm);
}

/* inflateBlock's dynamic arm (Deflate.hs:83-101): returns false with the verdict set. */
PZ_DEV bool pz_dynamic_header(PzCtx &c, PzStreamSmem *sm) {
  uint32_t hlit, hdist, hclen, v;
  if (!pz_take(c, sm, 5, hlit)) return false;
  if (!pz_take(c, sm, 5, hdist)) return false;
  if (!pz_take(c, sm, 4, hclen)) return false;
  hlit += 257u; hdist += 1u; hclen += 4u;
  const int lane = pz_lane();
  pz_syncwarp();
  for (int i = lane; i < 19; i += PZ_G) sm->lens[i] = 0;
  pz_syncwarp();
  for (uint32_t i = 0; i < hclen; i++) {
    if (!pz_take(c, sm, 3, v)) return false;
    if (lane == 0) sm->lens[PZ_CL_ORDER[i]] = (uint8_t)v;
  }
  int64_t val = 0;
  uint32_t *pre_lut = reinterpret_cast<uint32_t *>(sm->dist_lut);
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
      if (code == 16) { if (!pz_take(c, sm, 2, num)) return false; num += 3u; fill = prev; }
      else if (code == 17) { if (!pz_take(c, sm, 3, num)) return false; num += 3u; fill = 0; prev = 0; }
      else { if (!pz_take(c, sm, 7, num)) return false; num += 11u; fill = 0; prev = 0; }
      for (uint32_t i = (uint32_t)lane; i < num; i += PZ_G) sm->lens[n + i] = (uint8_t)fill;
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
PZ_DEV void pz_fixed_tables(PzStreamSmem *sm) {
  pz_syncwarp();
  for (int i = pz_lane(); i < 288 + 32; i += PZ_G)
    sm->lens[i] = (uint8_t)(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : i < 288 ? 8 : 5);
  pz_syncwarp();
  int64_t val;
  pz_build<PZ_LIT_BITS, 1>(sm->lens, 288, &sm->lit, sm->lit_perm, sm->lit_lut, sm->scratch, &val);
  pz_build<PZ_DIST_BITS, 2>(sm->lens + 288, 32, &sm->dist, sm->dist_perm, sm->dist_lut, sm->scratch, &val);
}

/* The stored arm (Deflate.hs:70-78, Monad.hs:265-293 for a single-chunk input). */
template <bool COUNT_ONLY>
PZ_DEV bool pz_stored_block(PzCtx &c, PzStreamSmem *sm) {
  uint32_t len, nlen;
  pz_align_byte(c, sm);
  if (pz_avail(c) < 32u) { pz_fail(c, PZ_ERR_DECOMPRESSION, PZ_D_RAN_OUT); return false; }
  if (!pz_take(c, sm, 16, len)) return false;
  if (!pz_take(c, sm, 16, nlen)) return false;
  if (len != ((~nlen) & 0xffffu)) { pz_fail(c, PZ_ERR_FORMAT, PZ_D_LEN_NLEN); return false; }
  uint32_t boff = c.bp >> 3;
  uint32_t remaining = (c.end_bit >> 3) - boff;
  /* getBlock takes the data only when strictly more than len bytes are left in the chunk (with zlib and gzip framing a
   * trailer always follows; a raw deflate stream may end with the last byte of a stored block) */
  if (len >= remaining && !(c.framing == PZ_FRAME_RAW && len == remaining)) { pz_fail(c, PZ_ERR_DECOMPRESSION, PZ_D_RAN_OUT); return false; }
  uint32_t fill = c.pos - c.base;
  if (fill + len > PZ_WINDOW) { pz_fail(c, PZ_REF_BOTTOM, PZ_D_BOT_WINDOW_OVERFLOW); return false; }
  if (len > c.cap - c.pos) { pz_fail(c, PZ_OUTPUT_FULL, 0); return false; }
  if (c.block_job && c.pos + len - c.mark > PZ_EXCESS) { pz_fail(c, PZ_REF_BOTTOM, PZ_D_BOT_WINDOW_OVERFLOW); return false; } /* PzCtx::mark */
  pz_push<COUNT_ONLY>(c, sm, PZ_TOKEN(PZ_Q_CTRL, PZ_C_STORED << 26));
  pz_push<COUNT_ONLY>(c, sm, boff - (c.start_bit >> 3)); /* offset from the stream's first byte */
  pz_push<COUNT_ONLY>(c, sm, len);
  c.pos += len;
  pz_seek(c, sm, (boff + len) * 8u);
  return true;
}

/* ---- the state machine ------------------------------------------------------------------ */
/* Publishes the verdict of the current stream and frees the group for its next one. */
PZ_DEV void pz_finish(PzCtx &c) {
#ifndef PZ_HOSTSIM
  /* lean kernel: the writer checks tokens the decoder side could not (pz_hot_warp_lean); the verdict waits until it has
   * confirmed the last one (pz_lean_drain, which calls this function again) */
  if (c.lean && c.mode != PZ_M_FINDRAIN) { c.mode = PZ_M_FINDRAIN; c.flush_sent = false; return; }
#endif
  pz_syncwarp();
  if (pz_lane() == 0) {
    pz_result *res = c.res;
    res->status = c.status;
    res->detail = c.block_job && c.status == PZ_OK ? (int32_t)c.bfinal : c.detail;
    res->out_len = c.block_job ? c.pos - PZ_BLK_BIAS : c.pos;
    res->adler_computed = 0;
    res->adler_stored = c.adler_stored;
    res->err_bitpos = c.bp - c.start_bit;
    res->payload[0] = c.p0;
    /* payload[1]: bytes the reference has already published as 32 KiB chunks (the shim's
     * incremental driver needs it); DIST_TOO_FAR keeps the retained-byte count instead */
    res->payload[1] = (c.status == PZ_REF_BOTTOM && c.detail == PZ_D_BOT_DIST_TOO_FAR) ? c.p1 : (int64_t)c.base;
    if (c.ck != nullptr) { /* PzJob::ckpt: c.pos / c.base are those of the last complete symbol */
      c.ck[0] = c.hdr_bp - c.start_bit;
      c.ck[1] = (c.sym_bp == 0u || c.sym_bp == PZ_CK_TRAILER) ? c.sym_bp : c.sym_bp - c.start_bit;
      c.ck[2] = c.pos;
      c.ck[3] = c.base;
    }
  }
  c.mode = PZ_M_IDLE;
}

/* checkChecksum (Deflate.hs:52-63): align, four bytes, most significant first.  The comparison is K3's. */
PZ_DEV void pz_trailer(PzCtx &c, PzStreamSmem *sm) {
  c.hdr_bp = c.bp; c.sym_bp = PZ_CK_TRAILER; /* a stream that stops here is picked up here */
  pz_align_byte(c, sm);
  uint32_t hi, lo;
  if (c.framing == PZ_FRAME_RAW) {
    /* nothing follows the last block */
  } else if (c.framing == PZ_FRAME_GZIP) { /* RFC 1952 2.3.1: CRC32, ISIZE, least significant byte first; K3 compares both */
    uint32_t s0, s1;
    if (pz_take(c, sm, 16, lo) && pz_take(c, sm, 16, hi)) {
      c.adler_stored = (hi << 16) | lo;
      if (pz_take(c, sm, 16, s0) && pz_take(c, sm, 16, s1)) c.p0 = (int64_t)((s1 << 16) | s0);
    }
  } else if (pz_avail(c) < 32u) pz_fail(c, PZ_ERR_DECOMPRESSION, PZ_D_RAN_OUT);
  else if (pz_take(c, sm, 16, hi) && pz_take(c, sm, 16, lo))
    c.adler_stored = ((hi & 0xffu) << 24) | ((hi >> 8) << 16) | ((lo & 0xffu) << 8) | (lo >> 8);
  pz_finish(c);
}

/* `decompress` for one single-chunk stream starts here: inflateWithHeaders (Zlib.hs:53-69). */
template <bool COUNT_ONLY>
PZ_DEV void pz_begin(PzCtx &c, PzStreamSmem *sm, uint32_t s, const uint8_t *in, uint64_t in_len, uint64_t out_cap, pz_result *res,
                     const uint32_t *rs = nullptr, uint32_t *ck = nullptr, uint32_t framing = PZ_FRAME_ZLIB) {
  uint32_t mis = (uint32_t)((uintptr_t)in & 15u);
  c.in_al = in - mis;
  c.res = res;
  c.framing = framing;
  c.pos = 0; c.base = 0; c.mark = 0;
  c.cap = out_cap > 0xfffdff00ull ? 0xfffdff00u : (uint32_t)out_cap; /* base + 128 KiB stays in 32 bits */
  c.status = PZ_OK; c.detail = 0; c.p0 = 0; c.p1 = 0;
  c.adler_stored = 0; c.bfinal = 0; c.need_careful = false;
  /* c.fixed_ready survives: the fixed-code LUTs stay valid until a dynamic header overwrites them */
  c.start_bit = mis * 8u;
  c.hdr_bp = c.start_bit; c.sym_bp = 0; c.resume_sym = 0; c.ck = ck;
  if (in_len > PZ_MAX_IN_BYTES) { /* the C ABI refuses such streams before launching */
    c.in_al_bytes = 0; c.end_bit = c.start_bit; c.safe_end = 0; c.bp = c.start_bit; c.q = 0;
    pz_fail(c, PZ_OUTPUT_FULL, 0);
    pz_finish(c);
    return;
  }
  c.in_al_bytes = (mis + (uint32_t)in_len + 15u) & ~15u;
  c.end_bit = (mis + (uint32_t)in_len) * 8u;
  c.safe_end = c.end_bit >= PZ_STEP_BITS ? c.end_bit - PZ_STEP_BITS : 0u; /* bp >= 16 once the header is read */
  /* the writer switches to this stream's output slice */
  pz_push<COUNT_ONLY>(c, sm, PZ_TOKEN(PZ_Q_CTRL, PZ_C_NEWSTREAM << 26));
  pz_push<COUNT_ONLY>(c, sm, s);
  if (rs != nullptr && PZ_HAS_CKPT(rs)) { /* PzJob::resume: the stream's header and rs[2] bytes of output are behind us */
    c.pos = rs[2]; c.base = rs[3];
    uint32_t bit = rs[0];
    if (bit > c.end_bit - c.start_bit) bit = c.end_bit - c.start_bit;
    pz_seek(c, sm, c.start_bit + bit);
    if (rs[1] == PZ_CK_TRAILER) { c.bfinal = 1; pz_trailer(c, sm); return; }
    c.resume_sym = rs[1];
    c.mode = PZ_M_HDR;
    return;
  }
  pz_seek(c, sm, c.start_bit);
  if (framing == PZ_FRAME_RAW) { c.mode = PZ_M_HDR; return; }
  if (framing == PZ_FRAME_GZIP) { /* the member header of RFC 1952 2.3, checks in stream order, optional fields skipped */
    uint32_t id1, id2, cm, fl, v;
    bool good = pz_take(c, sm, 8, id1) && pz_take(c, sm, 8, id2);
    if (good && (id1 != 0x1fu || id2 != 0x8bu)) { pz_fail(c, PZ_ERR_HEADER, PZ_D_HDR_GZIP_MAGIC, (id1 << 8) | id2); good = false; }
    good = good && pz_take(c, sm, 8, cm);
    if (good && cm != 8u) { pz_fail(c, PZ_ERR_HEADER, PZ_D_HDR_METHOD, cm); good = false; }
    good = good && pz_take(c, sm, 8, fl);
    if (good && (fl & 0xe0u)) { pz_fail(c, PZ_ERR_HEADER, PZ_D_HDR_GZIP_FLAGS, fl); good = false; }
    for (int k = 0; good && k < 6; k++) good = pz_take(c, sm, 8, v); /* MTIME, XFL, OS */
    if (good && (fl & 4u)) { /* FEXTRA */
      uint32_t xlen = 0;
      good = pz_take(c, sm, 16, xlen);
      for (uint32_t k = 0; good && k < xlen; k++) good = pz_take(c, sm, 8, v);
    }
    for (uint32_t bit = 8u; bit <= 16u; bit <<= 1) /* FNAME, FCOMMENT: zero-terminated */
      if (good && (fl & bit)) do { good = pz_take(c, sm, 8, v); } while (good && v != 0u);
    if (good && (fl & 2u)) good = pz_take(c, sm, 16, v); /* FHCRC: skipped, not verified */
    if (!good) { pz_finish(c); return; }
    c.mode = PZ_M_HDR;
    return;
  }
  uint32_t cmf, flg;
  bool ok = pz_take(c, sm, 8, cmf) && pz_take(c, sm, 8, flg);
  if (ok) {
    if (((cmf << 8) | flg) % 31u != 0u) { pz_fail(c, PZ_ERR_HEADER, PZ_D_HDR_CHECKSUM); ok = false; }
    else if ((cmf & 15u) != 8u) { pz_fail(c, PZ_ERR_HEADER, PZ_D_HDR_METHOD, cmf & 15u); ok = false; }
    else if ((cmf >> 4) > 7u) { pz_fail(c, PZ_ERR_HEADER, PZ_D_HDR_WINDOW, cmf >> 4); ok = false; }
    else if (flg & 0x20u) { /* FDICT: the four DICTID bytes are skipped (Zlib.hs:68) */
      uint32_t skip;
      ok = pz_take(c, sm, 16, skip) && pz_take(c, sm, 16, skip);
    }
  }
  if (!ok) { pz_finish(c); return; }
  c.mode = PZ_M_HDR;
}

/* A block job starts at a block header somewhere inside its stream (no zlib framing). */
template <bool COUNT_ONLY>
PZ_DEV void pz_begin_block(PzCtx &c, PzStreamSmem *sm, uint32_t j, const uint8_t *in, uint64_t in_len, uint32_t bit, uint32_t cap, pz_result *res) {
  uint32_t mis = (uint32_t)((uintptr_t)in & 15u);
  c.in_al = in - mis;
  c.res = res;
  c.pos = PZ_BLK_BIAS; c.base = 0; c.mark = PZ_BLK_BIAS;
  c.framing = PZ_FRAME_ZLIB; /* a block job never reads framing; its stored blocks always have bytes behind them or are declined */
  c.cap = PZ_BLK_BIAS + cap;
  c.status = PZ_OK; c.detail = 0; c.p0 = 0; c.p1 = 0;
  c.adler_stored = 0; c.bfinal = 0; c.need_careful = false;
  c.start_bit = mis * 8u;
  c.hdr_bp = c.start_bit; c.sym_bp = 0; c.resume_sym = 0; c.ck = nullptr;
  c.in_al_bytes = (mis + (uint32_t)in_len + 15u) & ~15u;
  c.end_bit = (mis + (uint32_t)in_len) * 8u;
  c.safe_end = c.end_bit >= PZ_STEP_BITS ? c.end_bit - PZ_STEP_BITS : 0u;
  pz_push<COUNT_ONLY>(c, sm, PZ_TOKEN(PZ_Q_CTRL, PZ_C_NEWSTREAM << 26));
  pz_push<COUNT_ONLY>(c, sm, j);
  if (bit > c.end_bit - c.start_bit) bit = c.end_bit - c.start_bit; /* the header read then reports the truncation */
  pz_seek(c, sm, c.start_bit + bit);
  c.mode = PZ_M_HDR;
}

/* End of a block (Deflate.hs:45-50): moveWindow, then either the next block or the trailer
 * (checkChecksum, Deflate.hs:52-63: align, four bytes, most significant first). */
PZ_DEV void pz_block_end(PzCtx &c, PzStreamSmem *sm) {
  pz_move_window(c);
  c.mark = c.pos; /* a moveWindow call ends a gap (PzCtx::mark) */
  if (c.block_job) { pz_finish(c); return; } /* err_bitpos = first bit after the block */
  if (!c.bfinal) { c.mode = PZ_M_HDR; return; }
  pz_trailer(c, sm);
}

/* One transition of a group that is not in the hot loop. */
template <bool COUNT_ONLY>
PZ_DEV void pz_slow_step(PzCtx &c, PzStreamSmem *sm, const PzJob &job, uint32_t stride) {
  if (c.mode == PZ_M_IDLE) {
    while (job.next_unit == nullptr && job.skip_done && c.next < job.first + job.count && job.res[c.next].status != PZ_ST_PENDING) {
#ifndef PZ_HOSTSIM
      if (job.prog != nullptr && pz_lane() == 0) *(volatile uint32_t *)(job.prog + c.next) = PZ_PROG_DONE; /* K2 wrote it before K1 started */
#endif
      c.next += stride;
    }
    while (job.next_unit != nullptr) { /* claim the next unit nobody has taken (and that nobody has finished: K2, K5) */
      uint32_t v = 0;
#ifdef PZ_HOSTSIM
      v = (*job.next_unit)++;
#else
      if (pz_lane() == 0) v = atomicAdd(job.next_unit, 1u);
      v = (uint32_t)pz_shfl((int)v, 0);
#endif
      c.next = v < job.count ? job.first + v : job.first + job.count;
      if (c.next >= job.first + job.count || !job.skip_done || job.res[c.next].status == PZ_ST_PENDING) break;
    }
    if (c.next >= job.first + job.count) {
      pz_push<COUNT_ONLY>(c, sm, PZ_TOKEN(PZ_Q_CTRL, PZ_C_EXIT << 26));
      c.mode = PZ_M_DEAD;
      return;
    }
    const uint32_t s = c.next;
#ifndef PZ_HOSTSIM
    /* the stream's input may still be on its way to the device: stay idle, the warp loop dozes */
    c.starved = job.in_ready != nullptr && *(const volatile uint32_t *)job.in_ready <= s;
    if (c.starved) return;
#endif
    c.next += stride;
    if (job.blk_start != nullptr) {
      const uint64_t b0 = job.in_off[job.blk_stream], b1 = job.in_off[job.blk_stream + 1];
      c.block_job = true;
      pz_begin_block<COUNT_ONLY>(c, sm, s, job.in_blob + b0, b1 - b0, job.blk_start[s], job.blk_len ? job.blk_len[s] : job.blk_cap, job.res + s);
      return;
    }
    const uint32_t e = s << job.pair_off;
    const uint64_t i0 = job.in_off[e], i1 = job.in_off[e + 1];
    if (COUNT_ONLY) {
      pz_begin<true>(c, sm, s, job.in_blob + i0, i1 - i0, ~0ull, job.res + s, nullptr, nullptr, job.framing);
    } else {
      const uint64_t o0 = job.out_off[e], o1 = job.out_off[e + 1];
      pz_begin<false>(c, sm, s, job.in_blob + i0, i1 - i0, o1 - o0, job.res + s, job.resume ? job.resume + 4u * s : nullptr,
                      job.ckpt ? job.ckpt + 4u * s : nullptr, job.framing);
    }
  } else if (c.mode == PZ_M_HDR) { /* inflateBlock (Deflate.hs:65-104) */
    uint32_t btype;
    c.hdr_bp = c.bp; c.sym_bp = 0;
    if (!pz_take(c, sm, 1, c.bfinal) || !pz_take(c, sm, 2, btype)) { pz_finish(c); return; }
    if (btype == 0u) {
      c.resume_sym = 0;
      if (!pz_stored_block<COUNT_ONLY>(c, sm)) { pz_finish(c); return; }
      pz_block_end(c, sm);
    } else if (btype == 1u || btype == 2u) {
      if (btype == 1u) {
        if (!c.fixed_ready) { pz_fixed_tables(sm); c.fixed_ready = true; }
      } else {
        c.fixed_ready = false;
        if (!pz_dynamic_header(c, sm)) { pz_finish(c); return; }
      }
      if (c.resume_sym != 0u) { /* PzJob::resume inside a block: its tables exist again, go on at the symbol */
        pz_seek(c, sm, c.start_bit + c.resume_sym);
        c.resume_sym = 0;
      }
      c.mode = PZ_M_SYMS;
    } else {
      pz_fail(c, PZ_ERR_FORMAT, PZ_D_BAD_BTYPE, 3);
      pz_finish(c);
    }
  } else { /* PZ_M_SYMS */
    if (!c.need_careful) {
      uint32_t lim = c.base + PZ_WINDOW;
      if (c.cap < lim) lim = c.cap;
      if (c.bp <= c.safe_end && c.pos < lim) {
        c.mode = PZ_M_FAST;
        return;
      }
    }
    c.need_careful = false;
    int r = pz_symbol_careful<COUNT_ONLY>(c, sm);
    if (r < 0) pz_finish(c);
    else if (r == 0) pz_block_end(c, sm);
  }
}

#if defined(PZ_PHASES) && !defined(PZ_HOSTSIM)
__device__ unsigned long long pz_phase_ticks[16]; /* [mode at the start of a step] = clocks, [8] = all, [9] = groups */
#endif
/* The service warp (device) / the whole decoder (host build): every group takes streams
 * first_stream, first_stream + stride, ... of the job (first_stream differs per group) through
 * the state machine.  On the device a stream that reaches the symbol loop is posted to its lane
 * of the hot warp and the group only feeds the input ring until the lane returns it. */
template <bool COUNT_ONLY>
PZ_DEV void pz_decoder_warp(const PzJob &job, uint32_t first_stream, uint32_t stride, PzStreamSmem *sm
#ifdef PZ_HOSTSIM
                            , PzWriter *hw
#else
                            , bool present, bool lean = false
#endif
) {
  PzCtx c;
  c.mode = PZ_M_IDLE;
  c.lean = false; c.flush_sent = false;
#ifndef PZ_HOSTSIM
  c.lean = lean;
#endif
  c.next = first_stream;
  c.in_al = nullptr; c.in_al_bytes = 0; c.bp = 0; c.q = 0; c.next_q = 0; c.pending = false; c.starved = false; c.block_job = false; c.res = nullptr;
  c.framing = PZ_FRAME_ZLIB;
  c.hdr_bp = 0; c.sym_bp = 0; c.resume_sym = 0; c.ck = nullptr; c.mark = 0;
  c.qhead = 0; c.qtailc = 0;
  c.fixed_ready = false;
#ifdef PZ_HOSTSIM
  c.hw = hw;
  for (;;) {
    while (c.mode != PZ_M_FAST && c.mode != PZ_M_DEAD) pz_slow_step<COUNT_ONLY>(c, sm, job, stride);
    if (!pz_warp_any(c.mode == PZ_M_FAST)) break; /* every group of the warp is out of streams */
    pz_fast_loop<COUNT_ONLY>(c, sm);
  }
#else
  if (!present) c.mode = PZ_M_DEAD; /* a group without a slot */
#ifdef PZ_PHASES /* debug build: where a service group's time goes, in SM clocks, summed over the groups (pz_phase_ticks) */
  long long ph_t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long ph_last = clock64();
  const long long ph_begin = ph_last;
#endif
  for (;;) {
    bool draining = false;
#ifdef PZ_PHASES
    const uint32_t ph_mode = c.mode;
#endif
    if (c.mode == PZ_M_WAIT) {
      pz_service_poll(c, sm);
    } else if (c.mode == PZ_M_DRAIN || c.mode == PZ_M_FINDRAIN) {
      draining = pz_lean_drain(c, sm);
    } else if (c.mode != PZ_M_DEAD) {
      pz_slow_step<COUNT_ONLY>(c, sm, job, stride);
      if (c.mode == PZ_M_FAST) pz_post_hot(c, sm);
      else if (c.mode == PZ_M_DEAD && pz_lane() == 0) pz_vstore(&sm->mail.state, PZ_MS_DEAD);
    }
#ifdef PZ_PHASES
    { const long long now = clock64(); ph_t[ph_mode & 7u] += now - ph_last; ph_last = now; }
#endif
    /* groups that only wait for their writer do not keep the warp spinning at full speed */
    if (pz_warp_any(draining) && !pz_warp_any(!draining && c.mode != PZ_M_WAIT && c.mode != PZ_M_DEAD && !(c.mode == PZ_M_IDLE && c.starved))) __nanosleep(100);
    if (!pz_warp_any(c.mode != PZ_M_WAIT && c.mode != PZ_M_DEAD && !(c.mode == PZ_M_IDLE && c.starved))) {
      if (!pz_warp_any(c.mode != PZ_M_DEAD)) break;
      /* every group waits for its hot lane: doze until one of them needs something (its stream
       * back, a ring quarter, or the end of a copy it has requested) */
      for (;;) {
        bool need = false;
        if (c.mode == PZ_M_WAIT)
          need = c.pending || pz_vload(&sm->mail.state) != PZ_MS_HOT ||
                 c.next_q <= (pz_vload(&sm->mail.hot_bp) >> PZ_QUARTER_SHIFT) + 3u;
        else if (c.mode == PZ_M_IDLE)
          need = *(const volatile uint32_t *)job.in_ready > c.next;
        if (pz_warp_any(need)) break;
        __nanosleep(PZ_DOZE_NS);
      }
    }
  }
#endif
#if defined(PZ_PHASES) && !defined(PZ_HOSTSIM)
  if (present && pz_lane() == 0) {
    for (int k = 0; k < 8; k++) atomicAdd(&pz_phase_ticks[k], (unsigned long long)ph_t[k]);
    atomicAdd(&pz_phase_ticks[8], (unsigned long long)(clock64() - ph_begin));
    atomicAdd(&pz_phase_ticks[9], 1ull);
  }
#endif
  pz_async_wait_all();
}
