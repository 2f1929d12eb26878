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
#define PZ_SERVICE_WARPS ((PZ_SLOTS + 3u) / 4u) /* four slots (8 lanes each) per service warp */
#define PZ_WARPS_PER_CTA (1u + 2u * PZ_SERVICE_WARPS) /* hot + service + writer warps */
#define PZ_THREADS_PER_CTA (32u * PZ_WARPS_PER_CTA)
#define PZ_MAX_STREAM_BYTES 0x1ffffff0ull /* == PZ_MAX_IN_BYTES in pz_device.cuh */
#define PZ_ADLER_SEG 16384u /* bytes per checksum segment (one warp each) */

/* Stream s = first + k decodes in_blob[in_off[s], in_off[s+1]) into out_blob[out_off[s], out_off[s+1]).
 * d_out == nullptr selects the sizing pass.  d_prog (optional, host-mapped, one word per stream, zeroed by
 * the caller) receives the progress words described at PzJob::prog.  d_in_ready (optional, device word): see
 * PzJob::in_ready; K2 is skipped then. */
cudaError_t pz_launch_inflate(const uint8_t *d_in, const uint64_t *d_in_off, uint8_t *d_out, const uint64_t *d_out_off,
                              uint32_t first, uint32_t count, pz_result *d_res, cudaStream_t st, uint32_t *d_prog = nullptr,
                              const uint32_t *d_in_ready = nullptr);
/* Adler-32 of each decoded stream (segments [seg_off[first], seg_off[first+count])), then the
 * trailer comparison that completes the verdict (Deflate.hs:52-63). */
cudaError_t pz_launch_adler(const uint8_t *d_out, const uint64_t *d_out_off, const uint64_t *d_seg_off, uint32_t n_total,
                            uint32_t first, uint32_t count, uint64_t seg_first, uint64_t seg_count, pz_result *d_res,
                            uint2 *d_parts, cudaStream_t st);
/* canonical codes of lens[0..n) (n <= 288) into codes[0..n) */
cudaError_t pz_launch_code_values(const uint8_t *d_lens, int n, uint16_t *d_codes, cudaStream_t st);
cudaError_t pz_kernels_configure(void);
/* streams K1 decodes at the same time on this device (SMs x slots per CTA) */
int pz_inflate_slots(void);
