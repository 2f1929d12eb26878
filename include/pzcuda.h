/*
 * pzcuda.h -- C ABI of libpzcuda.so, the B200-native zlib inflate engine that sits
 * beneath pure-zlib's `Codec.Compression.Zlib` API.
 *
 * The reference (GaloisInc/pure-zlib) has no FFI: its boundary is the Haskell module
 * API.  Every entry point below names the reference interface it replaces
 * (file:line relative to the reference checkout).  The Haskell shim
 * (haskell/Codec/Compression/Zlib.hs) binds exactly these symbols with
 * `foreign import ccall safe`; tests bind them through ctypes.
 *
 * There is NO CPU decode path in this library: every inflate call launches the
 * sm_100a kernels, and fails with PZ_E_CUDA if no device is usable.
 *
 * Thread safety: all entry points may be called concurrently from any number of
 * host threads (the reference's `decompress` is a pure function, Zlib.hs:32).
 */
#ifndef PZCUDA_H
#define PZCUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PZ_ABI_VERSION 2

/* ------------------------------------------------------------------------------------
 * Verdicts.  `status` is the constructor of the reference's DecompressionError
 * (Monad.hs:87-93) or one of the extra kinds below; `detail` selects the message
 * (the strings are listed next to each code; pz_strerror() reproduces them exactly).
 * ---------------------------------------------------------------------------------- */
enum pz_status {
  PZ_OK = 0,                /* Right bytes                      (Zlib.hs:46-47)            */
  PZ_ERR_HUFFMAN_TREE = 1,  /* Left (HuffmanTreeError s)        (Monad.hs:88)              */
  PZ_ERR_FORMAT = 2,        /* Left (FormatError s)             (Monad.hs:89)              */
  PZ_ERR_DECOMPRESSION = 3, /* Left (DecompressionError s)      (Monad.hs:90)              */
  PZ_ERR_HEADER = 4,        /* Left (HeaderError s)             (Monad.hs:91)              */
  PZ_ERR_CHECKSUM = 5,      /* Left (ChecksumError s)           (Monad.hs:92)              */
  PZ_REF_BOTTOM = 6,        /* the reference dies with an impure exception (bounds error
                               in Data.Array / Data.Vector); not a `Left`                  */
  PZ_OUTPUT_FULL = 7,       /* ABI-level: caller's capacity too small (no reference analogue;
                               the reference's output is unbounded)                        */
  PZ_NEED_MORE = 8          /* incremental only: input exhausted mid-stream (NeedMore)     */
};

enum pz_detail {
  PZ_D_NONE = 0,
  /* PZ_ERR_HUFFMAN_TREE (HuffmanTree.hs:36-83) */
  PZ_D_TWO_VALUES = 1,        /* "Two values point to the same place!"                     */
  PZ_D_VALUE_HIT = 2,         /* "HuffmanValue hit while inserting a value!"               */
  PZ_D_LEAF_IS_NODE = 3,      /* "Tried to add where the leaf is a node: <payload0>"       */
  PZ_D_ADVANCE_EMPTY_TREE = 4,/* "Tried to advance empty tree!"                            */
  PZ_D_ADVANCED_TO_EMPTY = 5, /* "Advanced to empty tree!"                                 */
  /* PZ_ERR_FORMAT (Deflate.hs:76-77,102-104) */
  PZ_D_LEN_NLEN = 1,          /* "Len/nlen mismatch in uncompressed block."                */
  PZ_D_BAD_BTYPE = 2,         /* "Unacceptable BTYPE: <payload0>"                          */
  /* PZ_ERR_DECOMPRESSION (Zlib.hs:38-39,48-49) */
  PZ_D_RAN_OUT = 1,           /* "Ran out of data mid-decompression 2."                    */
  PZ_D_DATA_REMAINING = 2,    /* "Finished with data remaining."                           */
  /* PZ_ERR_HEADER (Zlib.hs:62-67) */
  PZ_D_HDR_CHECKSUM = 1,      /* "Header checksum failed"                                  */
  PZ_D_HDR_METHOD = 2,        /* "Bad compression method: <payload0>"                      */
  PZ_D_HDR_WINDOW = 3,        /* "Window size too big: <payload0>"                         */
  PZ_D_HDR_GZIP_MAGIC = 4,    /* gzip framing (extension): "Not a gzip stream: <payload0 hex>" (ID1 ID2 != 1f 8b) */
  PZ_D_HDR_GZIP_FLAGS = 5,    /* gzip framing (extension): "Reserved gzip flags set: <payload0>" (FLG & 0xe0) */
  /* PZ_ERR_CHECKSUM (Deflate.hs:56-63) */
  PZ_D_ADLER_MISMATCH = 1,    /* "checksum mismatch: <adler_stored hex> != <adler_computed hex>" */
  PZ_D_LENGTH_MISMATCH = 2,   /* gzip framing (extension): "length mismatch: <payload0 = ISIZE> != <out_len mod 2^32>" */
  /* PZ_REF_BOTTOM */
  PZ_D_BOT_LENGTH_SYM = 1,    /* lengthArray ! payload0, payload0 in {286,287} (Deflate.hs:161,167) */
  PZ_D_BOT_DIST_SYM = 2,      /* distanceArray ! payload0, payload0 >= 30    (Deflate.hs:200,206)  */
  PZ_D_BOT_DIST_TOO_FAR = 3,  /* MV.slice (next - dist): payload0 = dist, payload1 = bytes retained (OutputWindow.hs:87) */
  PZ_D_BOT_WINDOW_OVERFLOW = 4/* write past the 128 KiB window (OutputWindow.hs:67,79,87)  */
};

/* One verdict per stream.  48 bytes, no padding. */
typedef struct pz_result {
  int32_t status;          /* enum pz_status                                               */
  int32_t detail;          /* enum pz_detail                                               */
  uint64_t out_len;        /* bytes decoded when the verdict was reached                   */
  uint32_t adler_computed; /* Adler-32 of the decoded bytes (valid for PZ_OK / PZ_ERR_CHECKSUM) */
  uint32_t adler_stored;   /* big-endian trailer word (Deflate.hs:55)                      */
  uint64_t err_bitpos;     /* bit offset in the stream at which decoding stopped (diagnostic) */
  int64_t payload[2];      /* message parameters, see enum pz_detail                       */
} pz_result;

/* Library-level return codes (negative = the call itself failed; verdicts are per stream). */
#define PZ_E_OK 0
#define PZ_E_CUDA (-1)     /* no usable device / CUDA error: the library has no CPU path    */
#define PZ_E_ARG (-2)
#define PZ_E_NOMEM (-3)
#define PZ_E_STATE (-4)

/* Flags for the batch calls. */
#define PZ_F_NO_ADLER 0x1u    /* skip the Adler-32 pass (verdicts then stop before the trailer compare) */
#define PZ_F_COUNT_ONLY 0x2u  /* sizing pass: decode symbols, write nothing */
#define PZ_F_INPUT_IN_PLACE 0x4u /* host blobs: let the kernel read a pinned, mapped in_blob over PCIe instead of copying it */
#define PZ_F_NO_HUGE 0x10u   /* never use the block-parallel path (K4) for huge streams */
#define PZ_F_NO_DRAIN 0x8u    /* host blobs: copy the output only after the kernel (no progressive 2-D copies) */
/* Framing (EXTENSION: the reference reads zlib streams only; gzip and raw deflate are the first TODO of its README,
 * lines 42-50, with docs/rfc1952.html shipped beside it).  Default: zlib (RFC 1950).
 *   PZ_F_GZIP  every stream is one gzip member (RFC 1952): header checked in stream order (magic, CM = 8, reserved FLG
 *              bits), FEXTRA / FNAME / FCOMMENT / FHCRC skipped (FHCRC not verified, in the spirit of the reference's
 *              FDICT), trailer CRC-32 then ISIZE; adler_computed / adler_stored of the verdict hold the CRC-32s.  Bytes
 *              behind the trailer (further members) are ignored like the reference ignores bytes behind a zlib trailer.
 *   PZ_F_RAW   raw deflate (RFC 1951): no header, no trailer, no comparison; adler_computed is the Adler-32 of the output. */
#define PZ_F_GZIP 0x20u
#define PZ_F_RAW 0x40u

#define PZ_MAX_DEVICES 16
typedef struct pz_config {
  int32_t device;          /* CUDA device ordinal of a single-device configuration, -1 = current device */
  int32_t n_devices;       /* > 0: host-buffer batches are sharded over devices[0 .. n_devices) (devices[0] is the
                              primary device: resident batches, incremental contexts); < 0: over every visible device;
                              0: one device (`device`), or what the environment variable PZ_DEVICES ("all", "0,1,3")
                              names when `device` is -1 */
  int32_t devices[PZ_MAX_DEVICES];
  int32_t reserved[6];
} pz_config;

/* ---- lifetime --------------------------------------------------------------------- */
/* Once-only, race-free initialisation.  Called implicitly (with cfg = NULL) by every other entry point.
 * Replaces nothing in the reference (it has no global state).
 * Multi-GPU: with more than one device configured, pz_inflate_batch, pz_inflate_sizes,
 * pz_decompress_batch and pz_inflate_batch_contig on HOST blobs cut the batch into contiguous ranges of streams
 * balanced by compressed bytes, one per device, each on a worker thread of its own with its own pinned staging, CUDA
 * streams and kernels -- ONE call from ONE host thread uses every device; no collective (streams are independent).  */
int pz_init(const pz_config *cfg);
/* Releases the CALLING thread's workspace (device and pinned buffers are kept per host thread). */
void pz_shutdown(void);
/* Devices host-buffer batches are sharded over (1 unless pz_init / PZ_DEVICES said otherwise). */
int pz_device_count(void);
int pz_abi_version(void);
/* Tuning knobs.  PZ_OPT_HUGE_BYTES: compressed size from which a stream is decoded block-parallel
 * (K4) instead of as one serial chain; default 4 MiB, environment PZ_HUGE_BYTES at start-up. */
#define PZ_OPT_HUGE_BYTES 1
/* PZ_OPT_STREAM_RESUME: accepted and ignored since ABI 2 (it made every pump of an incremental stream decode from the
 * stream's first byte again, for A/B timing; contexts no longer keep the bytes that would need). */
#define PZ_OPT_STREAM_RESUME 2
int pz_set_option(int key, uint64_t value);
/* Process-wide counters (diagnostics, tests): streams the block-parallel path decoded / declined. */
#define PZ_CTR_HUGE_DONE 1
#define PZ_CTR_HUGE_DECLINED 2
uint64_t pz_get_counter(int which);
/* Human-readable text of the last CUDA failure on this thread. */
const char *pz_last_error(void);

/* ---- batch inflate: the hot path --------------------------------------------------- *
 * Replaces `decompress` (Zlib.hs:32-51) applied to each of n independent single-chunk
 * streams.  Host pointers; the library stages through pinned memory, launches the
 * kernels, copies results back.  out[i] receives at most out_cap[i] bytes.               */
int pz_inflate_batch(const uint8_t *const *in, const size_t *in_len, uint8_t *const *out,
                     const size_t *out_cap, size_t n, pz_result *res, uint32_t flags);

/* Same, contiguous layout: stream i is in_blob[in_off[i] .. in_off[i+1]) and decodes into
 * out_blob[out_off[i] .. out_off[i+1]).  in_blob/out_blob may be HOST or DEVICE pointers
 * (detected with cudaPointerGetAttributes); the offset arrays and res are host memory.
 * `stream` is a cudaStream_t (NULL = default stream).                                    */
int pz_inflate_batch_contig(const uint8_t *in_blob, const uint64_t *in_off, uint8_t *out_blob,
                            const uint64_t *out_off, size_t n, pz_result *res, void *stream,
                            uint32_t flags);

/* Sizing pass: the decoded length of each stream (res[i].out_len), nothing written.
 * Lets the shim implement `decompress` without caller-supplied capacities.               */
int pz_inflate_sizes(const uint8_t *const *in, const size_t *in_len, size_t n, pz_result *res);
/* The same behind another framing (flags: PZ_F_GZIP or PZ_F_RAW, see there). */
int pz_inflate_sizes_framed(const uint8_t *const *in, const size_t *in_len, size_t n, pz_result *res, uint32_t flags);

/* `map decompress` in ONE call (Zlib.hs:32-51 for each of n single-chunk streams): sizing pass, output allocation and
 * decode inside the library, so the compressed bytes are packed once and nothing is copied on the host on the way out:
 * out[i] points at stream i's res[i].out_len decoded bytes inside one pinned block owned by *handle, valid until
 * pz_outputs_free(*handle) (the Haskell shim wraps the block in ONE ForeignPtr that every result ByteString shares).  */
typedef struct pz_outputs pz_outputs;
int pz_decompress_batch(const uint8_t *const *in, const size_t *in_len, size_t n, pz_result *res, uint8_t **out,
                        pz_outputs **handle, uint32_t flags);
void pz_outputs_free(pz_outputs *handle);

/* ---- resident batches: launch-only hot path ---------------------------------------- *
 * A plan owns the device-side descriptor tables (offsets, results, Adler partials), so
 * that pz_batch_run() is kernel launches only -- this is what bench.py times as `value`. */
typedef struct pz_batch pz_batch;
pz_batch *pz_batch_create(const uint64_t *in_off, const uint64_t *out_off, size_t n, uint32_t flags);
/* d_in / d_out are DEVICE pointers to the blobs described at creation. Asynchronous.     */
int pz_batch_run(pz_batch *b, const uint8_t *d_in, uint8_t *d_out, void *stream);
/* Synchronises `stream` and copies the n verdicts to host memory.                        */
int pz_batch_results(pz_batch *b, pz_result *res, void *stream);
/* Kernel launches issued by one pz_batch_run() (for bench.py's `gpu_launches`).          */
int pz_batch_launches(const pz_batch *b);
void pz_batch_destroy(pz_batch *b);

/* ---- pinned host memory ------------------------------------------------------------ *
 * Page-locked buffers for callers that want pz_inflate_batch_contig's host<->device copies
 * to overlap with decoding (the reference's strict ByteStrings are pinned ForeignPtrs;
 * this is the CUDA notion of the same thing).                                            */
void *pz_pinned_alloc(size_t bytes);
void pz_pinned_free(void *p);

/* ---- incremental decoder ----------------------------------------------------------- *
 * Replaces `decompressIncremental` / `ZlibDecoder` (Zlib.hs:29-30, Monad.hs:163-197,
 * 338-358).  The event sequence is the reference's: after each fed chunk the decoder
 * yields zero or more 32768-byte chunks, then PZ_S_NEED_MORE, or the final chunk
 * followed by PZ_S_DONE, or PZ_S_ERROR.                                                  */
typedef struct pz_stream pz_stream;
enum pz_stream_event { PZ_S_NEED_MORE = 0, PZ_S_CHUNK = 1, PZ_S_DONE = 2, PZ_S_ERROR = 3 };
pz_stream *pz_stream_new(void);
/* The same decoder behind another framing: flags = 0, PZ_F_GZIP or PZ_F_RAW (extension, see the flags). */
pz_stream *pz_stream_new_framed(uint32_t flags);
/* Supply the strict chunk that answers a NeedMore (Monad.hs:185-197).  The bytes are copied
 * (pinned staging, cudaMemcpyAsync): the call returns while they travel to the device.      */
int pz_stream_feed(pz_stream *s, const uint8_t *data, size_t len);
/* Next decoder state.  For PZ_S_CHUNK, `chunk` and `len` describe bytes owned by the stream and
 * valid until the next call on it (pz_stream_next, pz_stream_pump or pz_stream_feed).  For
 * PZ_S_ERROR, `res` (if non-NULL) receives the verdict.  If input has arrived since the last
 * decode, the call pumps the stream first.                                                 */
int pz_stream_next(pz_stream *s, const uint8_t **chunk, size_t *len, pz_result *res);
void pz_stream_free(pz_stream *s);
/* Decodes what has been fed to each of the n streams since its last decode, ALL streams in one
 * kernel launch (extension: many concurrent decompressIncremental consumers multiplexed onto
 * one device).  A stream's state lives on the device -- compressed bytes, decoded bytes (the
 * LZ77 history) and a checkpoint (block header, symbol, bytes decoded, bytes published) -- so
 * a pump decodes only what the new input adds (Monad.hs:163-197: the reference's coroutine
 * goes on where it stopped).  Streams with nothing new are skipped.  Afterwards
 * pz_stream_next() on any of them returns without touching the device.                      */
int pz_stream_pump(pz_stream *const *streams, size_t n);
/* The feed that goes with it: data[i] / len[i] is the next chunk of streams[i] (each stream at most once per call).
 * All chunks cross the bus in one copy.  Returns when the bytes are on the device.               */
int pz_stream_feed_many(pz_stream *const *streams, const uint8_t *const *data, const size_t *len, size_t n);
/* Introspection for tests and benchmarks. */
enum pz_stream_counter_id {
  PZ_SC_PUMPS = 0,      /* kernel launches that decoded this stream                          */
  PZ_SC_RESUMED = 1,    /* ... of which started from a checkpoint instead of the first byte  */
  PZ_SC_CKPT_BIT = 2,   /* compressed bits the checkpoint has behind it                      */
  PZ_SC_CKPT_BYTES = 3, /* decoded bytes the checkpoint has behind it                        */
  PZ_SC_DEVICE_BYTES = 4, /* device memory the context holds right now (input + history + room)    */
  PZ_SC_DEVICE_PEAK = 5,  /* ... and the most it has ever held                                      */
  PZ_SC_HOST_BYTES = 6    /* pinned host memory the context holds (decoded bytes not handed out yet, feed staging) */
};
uint64_t pz_stream_counter(const pz_stream *s, int which);

/* ---- auxiliaries ------------------------------------------------------------------- */
/* `show` of the DecompressionError this verdict denotes (Monad.hs:95-104), e.g.
 * "Checksum error: checksum mismatch: 680308b0 != 680308b1".  Returns the length that
 * was (or would have been) written, excluding the NUL.                                  */
size_t pz_strerror(const pz_result *r, char *buf, size_t cap);
/* KAT hook for `computeCodeValues` (Deflate.hs:261-288; test/Test.hs:107-120): input n
 * (symbol,length) pairs, output m (symbol,length,code) triples ascending by symbol;
 * returns m.  Runs the device table builder's canonical-code kernel.                    */
int pz_compute_code_values(const int32_t *sym, const int32_t *len, int n, int32_t *out_triples);
/* Adler-32 of a host buffer on the device (Adler32.hs:17-57); init = 1 for a fresh sum.  */
uint32_t pz_adler32(uint32_t init, const uint8_t *data, size_t len);
/* CRC-32 (RFC 1952 section 8) of a host buffer on the device, the checksum of the gzip framing; init = 0 for a fresh sum. */
uint32_t pz_crc32(uint32_t init, const uint8_t *data, size_t len);

#ifdef __cplusplus
}
#endif
#endif /* PZCUDA_H */
