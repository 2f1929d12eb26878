/*
 * pz_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of pure-zlib's `decompress` / `decompressIncremental`, used as the
 * parity checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  Nothing under pure_zlib_b200/ may include, link or call it.
 *
 * Parity pinning: checked against the reference's own 9 golden fixtures and 2 KATs
 * (test/Test.hs:13-120) by tests/test_oracle.py.  The reference itself cannot be run
 * here (no GHC), so behaviours those fixtures do not reach (fixed blocks, error
 * verdicts, chunked input) rest on this restatement following the cited lines, with
 * system zlib as a second opinion on valid streams: for those behaviours parity is
 * "unpinned by reference output".
 *
 * The numeric verdict codes deliberately equal the ones in include/pzcuda.h (a test
 * asserts this); the header is independent so the oracle builds without the product.
 */
#ifndef PZ_ORACLE_H
#define PZ_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct pzo_result {
  int32_t status;
  int32_t detail;
  uint64_t out_len;        /* bytes decoded (published + still in the window) at the verdict */
  uint32_t adler_computed;
  uint32_t adler_stored;
  uint64_t err_bitpos;     /* not compared */
  int64_t payload[2];
} pzo_result;

/* Event log of the incremental decoder: what a consumer of `ZlibDecoder` observes. */
enum { PZO_EV_NEED_MORE = 0, PZO_EV_CHUNK = 1, PZO_EV_DONE = 2, PZO_EV_ERROR = 3 };
typedef struct pzo_event {
  int32_t kind;
  int32_t pad;
  uint64_t len; /* chunk length for PZO_EV_CHUNK */
} pzo_event;

/* `decompress` (Zlib.hs:32-51) over a lazy ByteString given as `nchunks` strict chunks laid
 * end to end in `in` (chunk i has chunk_len[i] bytes).  Decoded bytes go to out[0..out_cap)
 * (decoding continues past out_cap, only the stores are dropped).  `ev`/`ev_cap`/`n_ev`
 * optionally receive the decoder-state sequence.  `published` receives the number of bytes
 * delivered through Chunk states.  Returns 0. */
int pzo_decompress(const uint8_t *in, const size_t *chunk_len, size_t nchunks, uint8_t *out,
                   size_t out_cap, pzo_result *res, pzo_event *ev, size_t ev_cap, size_t *n_ev,
                   uint64_t *published);

/* EXTENSION beyond the reference (gzip is the first TODO of its README, lines 42-50; docs/rfc1952.html ships with it):
 * the same decoder behind another framing.  framing 0 = zlib (pzo_decompress), 1 = gzip member (RFC 1952: header with
 * FEXTRA / FNAME / FCOMMENT / FHCRC skipped, CRC-32 + ISIZE trailer; res->adler_computed / adler_stored then hold the
 * CRC-32s), 2 = raw deflate (no header, no trailer).  Pinned against system zlib (wbits 31 / -15) on valid streams by
 * tests/test_oracle.py; error verdicts follow the reference's conventions and are this repository's definition. */
int pzo_decompress_framed(const uint8_t *in, const size_t *chunk_len, size_t nchunks, uint8_t *out,
                          size_t out_cap, pzo_result *res, pzo_event *ev, size_t ev_cap, size_t *n_ev,
                          uint64_t *published, int framing);

/* `computeCodeValues` (Deflate.hs:261-288). Returns the number of triples. */
int pzo_compute_code_values(const int32_t *sym, const int32_t *len, int n, int32_t *out_triples);

/* `createHuffmanTree . computeCodeValues` (Deflate.hs:255-259, HuffmanTree.hs:25-71) on
 * lengths lens[0..n) for symbols 0..n-1: returns 0 when the trie is built, else the detail
 * code (1,2,3) with *val = the symbol shown in message 3. */
int pzo_tree_check(const uint8_t *lens, int n, int64_t *val);

/* Adler-32 exactly as Adler32.hs:17-57 (init: a=1,b=0 packed as 1). */
uint32_t pzo_adler32(uint32_t init, const uint8_t *data, size_t len);

/* `show` of the verdict (Monad.hs:95-104); "" for PZ_OK; "_|_ <kind>" for bottoms. */
size_t pzo_strerror(const pzo_result *r, char *buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
