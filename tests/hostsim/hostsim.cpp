/*
 * hostsim.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Compiles pure_zlib_b200/csrc/pz_device.cuh with a host compiler (PZ_HOSTSIM: a warp is
 * one lane) so the decoder LOGIC -- bit accounting, LUT construction, verdict order, the
 * window model -- can be fuzzed against the oracle on a machine without a GPU.  It proves
 * nothing about the CUDA build's warp-level code; the `-m gpu` tests do that.  Never linked
 * into libpzcuda.so.
 */
#define PZ_HOSTSIM 1
#include "../../pure_zlib_b200/csrc/pz_device.cuh"
#include "../../pure_zlib_b200/csrc/pz_fixed.cuh"

#include <stdlib.h>

/* resume / ckpt (four words each, may be null): PzJob::resume, PzJob::ckpt; `out` then holds the stream's history */
extern "C" int hs_inflate_framed(const uint8_t *in, uint64_t in_len, uint8_t *out, uint64_t out_cap, pz_result *res, int count_only,
                                 const uint32_t *resume, uint32_t *ckpt, uint32_t framing) {
  /* the device reads whole 16-byte pieces around the stream: give it a padded, aligned copy
   * with a deliberately odd misalignment so the (mis != 0) paths run */
  size_t mis = 5;
  uint8_t *buf = (uint8_t *)aligned_alloc(16, ((in_len + mis + 15) & ~(size_t)15) + 1024 + 16);
  memset(buf, 0xA5, ((in_len + mis + 15) & ~(size_t)15) + 1024 + 16);
  memcpy(buf + mis, in, in_len);
  PzStreamSmem *sm = (PzStreamSmem *)aligned_alloc(16, (sizeof(PzStreamSmem) + 15) & ~(size_t)15);
  memset(sm, 0xCD, sizeof(PzStreamSmem));
  /* a one-stream job through the same state machine the kernel runs */
  /* with a checkpoint array the job also addresses its buffers the way the incremental driver does:
   * (begin, end) pairs that are addresses, blobs null */
  const bool pairs = ckpt != nullptr && !count_only;
  uint64_t in_off[2] = {0, in_len}, out_off[2] = {0, out_cap};
  if (pairs) {
    in_off[0] = (uint64_t)(uintptr_t)(buf + mis); in_off[1] = in_off[0] + in_len;
    out_off[0] = (uint64_t)(uintptr_t)out; out_off[1] = out_off[0] + out_cap;
  }
  PzJob job;
  job.in_blob = pairs ? nullptr : buf + mis; job.in_off = in_off; job.out_blob = (count_only || pairs) ? nullptr : out; job.out_off = out_off;
  job.pair_off = pairs ? 1u : 0u; job.resume = resume; job.ckpt = ckpt; job.framing = framing;
  job.parts = nullptr; job.seg_off = nullptr;
  job.res = res; job.first = 0; job.count = 1; job.skip_done = 0; job.prog = nullptr; job.in_ready = nullptr; job.blk_start = nullptr; job.blk_out = nullptr; job.blk_len = nullptr; job.out16 = nullptr; job.blk_stream = 0; job.blk_cap = 0;
  PzWriter hw; /* tokens are applied as they are pushed */
  pz_writer_init(hw, &job);
  if (count_only) pz_decoder_warp<true>(job, 0, 1, sm, &hw);
  else pz_decoder_warp<false>(job, 0, 1, sm, &hw);
  free(sm);
  free(buf);
  return 0;
}

extern "C" int hs_inflate_resume(const uint8_t *in, uint64_t in_len, uint8_t *out, uint64_t out_cap, pz_result *res, int count_only,
                                 const uint32_t *resume, uint32_t *ckpt) {
  return hs_inflate_framed(in, in_len, out, out_cap, res, count_only, resume, ckpt, 0);
}

extern "C" int hs_inflate(const uint8_t *in, uint64_t in_len, uint8_t *out, uint64_t out_cap, pz_result *res, int count_only) {
  return hs_inflate_resume(in, in_len, out, out_cap, res, count_only, nullptr, nullptr);
}

extern "C" int hs_smem_bytes(void) { return (int)sizeof(PzStreamSmem); }

/* K5's per-stream logic (pz_fixed.cuh) on the host: 1 if the stream decoded completely (res filled), 0 if it is left to K1.
 * out_mis (0..3) is the misalignment of the output inside a buffer of its own: 0 takes the word-wide stores, the others the
 * byte-wide ones; the bytes around the output are checked to be untouched. */
extern "C" int hs_fixed(const uint8_t *in, uint64_t in_len, uint8_t *out, uint64_t out_cap, pz_result *res, int count_only, int out_mis, int dyn) {
  size_t mis = 3; /* an odd misalignment of the stream inside its (padded) buffer */
  uint8_t *buf = (uint8_t *)aligned_alloc(16, ((in_len + mis + 15) & ~(size_t)15) + 64);
  memset(buf, 0x5A, ((in_len + mis + 15) & ~(size_t)15) + 64);
  memcpy(buf + mis, in, in_len);
  uint8_t *obuf = (uint8_t *)aligned_alloc(16, ((out_cap + 15) & ~(size_t)15) + 64);
  memset(obuf, 0xEE, ((out_cap + 15) & ~(size_t)15) + 64);
  uint8_t *o = obuf + 16 + out_mis;
  const bool ok = dyn ? (count_only ? pz_fixed_stream<true, true>(buf + mis, in_len, nullptr, 0, res) : pz_fixed_stream<false, true>(buf + mis, in_len, o, out_cap, res))
                      : (count_only ? pz_fixed_stream<true>(buf + mis, in_len, nullptr, 0, res) : pz_fixed_stream<false>(buf + mis, in_len, o, out_cap, res));
  int rc = ok ? 1 : 0;
  if (!count_only) {
    for (int k = 0; k < 16 + out_mis; k++) if (obuf[k] != 0xEE) rc = -1;          /* nothing before the output ... */
    for (size_t k = out_cap; k < out_cap + 16; k++) if (o[k] != 0xEE) rc = -1;    /* ... nor behind its capacity */
    if (ok) memcpy(out, o, res->out_len < out_cap ? res->out_len : out_cap);
  }
  free(obuf);
  free(buf);
  return rc;
}
