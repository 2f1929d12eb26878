/* pz_internal.h -- launch wrappers shared by pz_kernels.cu and pz_abi.cu (not installed). */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pzcuda.h"

#ifndef PZ_GROUP
#define PZ_GROUP 8 /* lanes per stream (pz_device.cuh) */
#endif
#ifndef PZ_SLOTS
#define PZ_SLOTS 28u /* streams a CTA works on at a time (one lane of the hot warp each) */
#endif
#ifndef PZ_WGROUP
#define PZ_WGROUP 16 /* lanes per stream in the writer warps (pz_device.cuh) */
#endif
#define PZ_SLOTS_PER_SERVICE (32u / PZ_GROUP) /* four slots (8 lanes each) per service warp */
#define PZ_SERVICE_WARPS ((PZ_SLOTS + PZ_SLOTS_PER_SERVICE - 1u) / PZ_SLOTS_PER_SERVICE)
#define PZ_SLOTS_PER_WRITER (32u / PZ_WGROUP)
#define PZ_WRITER_WARPS ((PZ_SLOTS + PZ_SLOTS_PER_WRITER - 1u) / PZ_SLOTS_PER_WRITER)
#ifndef PZ_PAD_WARPS
#define PZ_PAD_WARPS 6u /* warps that exit at once: they only keep working warps off the hot warp's scheduler */
#endif
#define PZ_WARPS_PER_CTA (1u + PZ_SERVICE_WARPS + PZ_WRITER_WARPS + PZ_PAD_WARPS) /* hot + service + writer (+ idle) warps */
#define PZ_THREADS_PER_CTA (32u * PZ_WARPS_PER_CTA)
/* warps 4, 8, 12, ... share the hot warp's scheduler: service warps (which doze), then the idle ones */
#define PZ_HOT_SCHED_WARPS ((PZ_WARPS_PER_CTA - 1u) / 4u)
#define PZ_HOT_SCHED_SERVICE (PZ_HOT_SCHED_WARPS - PZ_PAD_WARPS)
static_assert(PZ_HOT_SCHED_SERVICE + (PZ_WARPS_PER_CTA - 1u - PZ_HOT_SCHED_WARPS - PZ_WRITER_WARPS) == PZ_SERVICE_WARPS,
              "warp roles do not add up: adjust PZ_PAD_WARPS");
#define PZ_MAX_STREAM_BYTES 0x1ffffff0ull /* == PZ_MAX_IN_BYTES in pz_device.cuh */
#define PZ_ADLER_SEG 16384u /* bytes per checksum segment (one warp each) */
#define PZ_FIXED_MIN_STREAMS 8192u /* batches from this size on run K5 (pz_fixed.cuh: a lane per small fixed-Huffman stream) between K2 and K1 */

/* Stream s = first + k decodes in_blob[in_off[s], in_off[s+1]) into out_blob[out_off[s], out_off[s+1]).
 * d_out == nullptr selects the sizing pass.  d_prog (optional, host-mapped, one word per stream, zeroed by
 * the caller) receives the progress words described at PzJob::prog.  d_in_ready (optional, device word): see
 * PzJob::in_ready; K2 is skipped then.  d_parts / d_seg_off (optional): the batch's Adler-32 segment table; K2 then
 * checksums the streams it copies while it copies them (PzJob::parts). */
cudaError_t pz_launch_inflate(const uint8_t *d_in, const uint64_t *d_in_off, uint8_t *d_out, const uint64_t *d_out_off,
                              uint32_t first, uint32_t count, pz_result *d_res, cudaStream_t st, uint32_t *d_prog = nullptr,
                              const uint32_t *d_in_ready = nullptr, int phase = 0, uint2 *d_parts = nullptr,
                              const uint64_t *d_seg_off = nullptr, uint32_t framing = 0 /* PZ_FRAME_*: 0 zlib, 1 gzip, 2 raw deflate */);
#define PZ_PHASE_ALL 0
#define PZ_PHASE_K2 1 /* only the stored-stream kernels (they also mark every other stream PENDING) */
#define PZ_PHASE_K1 2 /* only K1: decodes the streams that are still PENDING */
#define PZ_ST_PENDING_HOST (-1) /* == PZ_ST_PENDING in pz_device.cuh */
/* Resumable contexts: count streams in buffers of their own ((begin, end) device addresses), each continuing from
 * d_resume[4s..] and leaving its next checkpoint in d_ckpt[4s..] (PzJob::resume / ckpt in pz_device.cuh).  K1 only. */
cudaError_t pz_launch_resume(const uint64_t *d_in_pairs, const uint64_t *d_out_pairs, uint32_t count, pz_result *d_res,
                             const uint32_t *d_resume, uint32_t *d_ckpt, cudaStream_t st, uint32_t framing = 0);
/* count pieces of decoded output go to pinned host memory: (device source, host destination, bytes) triples of uint64 */
cudaError_t pz_launch_gather(const uint64_t *d_triples, uint32_t count, cudaStream_t st);
#define PZ_CK_TRAILER_HOST 0xffffffffu /* == PZ_CK_TRAILER in pz_device.cuh */
/* K4 (pz_huge.cuh): one huge stream decoded block-parallel; see pz_abi.cu for the driver */
cudaError_t pz_launch_blk_search(const uint8_t *d_stream, uint64_t nbytes, uint64_t first_bit, uint64_t last_bit, uint32_t *d_cand,
                                 uint32_t *d_ncand, uint32_t cap, cudaStream_t st);
/* d_kept (room for ncand entries) receives the candidates that pass, in no particular order; *d_nkept (zeroed by the caller) counts them */
cudaError_t pz_launch_blk_verify(const uint8_t *d_stream, uint64_t nbytes, uint64_t last_bit, const uint32_t *d_cand, uint32_t ncand,
                                 uint32_t *d_kept, uint32_t *d_nkept, cudaStream_t st);
cudaError_t pz_launch_blk_jobs(const uint8_t *d_in_blob, const uint64_t *d_in_off2, const uint32_t *d_blk_start, const uint64_t *d_blk_out,
                               const uint32_t *d_blk_len, uint32_t cap, uint16_t *d_sym16, uint32_t count, pz_result *d_res, cudaStream_t st,
                               uint32_t *d_counter = nullptr /* optional zeroable device word: blocks are claimed in order instead of dealt by index */);
/* one-pass flow: symbols of chain block k move from d_scr + d_blk_src[k] to d_sym16 + d_blk_off[k] (pz_blk_compact_kernel) */
cudaError_t pz_launch_blk_compact(const uint16_t *d_scr, uint16_t *d_sym16, const uint64_t *d_blk_off, const uint64_t *d_blk_src, uint32_t nblk,
                                  uint64_t total, cudaStream_t st);
/* groups of consecutive chain blocks: d_grp_first[ngrp + 1] (first block of each group), d_blk_grp[nblk]; d_gw = ngrp * 32768 bytes */
cudaError_t pz_launch_blk_resolve(uint16_t *d_sym16, uint8_t *d_out, const uint64_t *d_blk_off, const uint32_t *d_blk_len, const uint32_t *d_blk_grp,
                                  const uint32_t *d_grp_first, uint32_t ngrp, uint32_t nblk, uint64_t total, uint8_t *d_gw, uint32_t *d_err,
                                  cudaStream_t st);
/* Adler-32 of each decoded stream (segments [seg_off[first], seg_off[first+count])), then the
 * trailer comparison that completes the verdict (Deflate.hs:52-63). */
cudaError_t pz_launch_adler(const uint8_t *d_out, const uint64_t *d_out_off, const uint64_t *d_seg_off, uint32_t n_total,
                            uint32_t first, uint32_t count, uint64_t seg_first, uint64_t seg_count, pz_result *d_res,
                            uint2 *d_parts, cudaStream_t st, uint32_t framing = 0 /* gzip: CRC-32 + ISIZE instead of Adler-32; raw: no compare */);
/* canonical codes of lens[0..n) (n <= 288) into codes[0..n) */
cudaError_t pz_launch_code_values(const uint8_t *d_lens, int n, uint16_t *d_codes, cudaStream_t st);
cudaError_t pz_kernels_configure(void);
/* streams K1 decodes at the same time on this device (SMs x slots per CTA) */
int pz_inflate_slots(void);
int pz_small_launches(uint32_t count, uint32_t framing); /* launches K5 (/ K6) add to a batch: 0, 2 or 4 */
