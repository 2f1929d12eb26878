/*
 * pz_oracle.c -- TEST INFRASTRUCTURE ONLY (see pz_oracle.h).
 *
 * A structural CPU restatement of pure-zlib 0.8.0's decoder: bit-serial reader, binary
 * trie walked one bit per step, 128 KiB sliding output window, per-byte Adler-32.  It is
 * written to be *obviously the same algorithm* as the reference, not to be fast; each
 * function cites the reference lines (relative to /root/reference) it restates.
 *
 * Haskell `Left e`  -> status 1..5 (+detail/payload)
 * impure exception  -> status 6 (PZ_REF_BOTTOM), SURVEY.md Appendix A.7
 */
#include "pz_oracle.h"

#include <setjmp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* verdict codes: numerically identical to include/pzcuda.h */
enum { ST_OK = 0, ST_HUFF = 1, ST_FORMAT = 2, ST_DECOMP = 3, ST_HEADER = 4, ST_CHECKSUM = 5, ST_BOTTOM = 6 };
enum { D_TWO_VALUES = 1, D_VALUE_HIT = 2, D_LEAF_IS_NODE = 3, D_ADV_EMPTY_TREE = 4, D_ADV_TO_EMPTY = 5 };
enum { D_LEN_NLEN = 1, D_BAD_BTYPE = 2 };
enum { D_RAN_OUT = 1, D_DATA_REMAINING = 2 };
enum { D_HDR_CHECKSUM = 1, D_HDR_METHOD = 2, D_HDR_WINDOW = 3, D_HDR_GZIP_MAGIC = 4, D_HDR_GZIP_FLAGS = 5 };
enum { D_ADLER_MISMATCH = 1, D_LENGTH_MISMATCH = 2 };
enum { FRAME_ZLIB = 0, FRAME_GZIP = 1, FRAME_RAW = 2 };
enum { D_BOT_LENGTH_SYM = 1, D_BOT_DIST_SYM = 2, D_BOT_DIST_TOO_FAR = 3, D_BOT_WINDOW_OVERFLOW = 4 };

#define WINDOW_SIZE (128 * 1024) /* OutputWindow.hs:29-30 */
#define EXCESS_CHUNK 32768       /* OutputWindow.hs:42-43 */
#define ADLER_MOD 65521u

/* ---------------------------------------------------------------- HuffmanTree.hs */

enum { T_EMPTY = 0, T_VALUE = 1, T_NODE = 2 };

/* Node storage is a bump allocator big enough for any tree the decoder can ask for:
 * at most 320+138 symbols of <= 15 bits.  Interior nodes <= leaves * 15. */
typedef struct {
  int *kind, *left, *right, *value;
  int n, cap, root;
} Tree;

static void tree_init(Tree *t) {
  t->cap = 1024;
  t->kind = malloc(sizeof(int) * t->cap);
  t->left = malloc(sizeof(int) * t->cap);
  t->right = malloc(sizeof(int) * t->cap);
  t->value = malloc(sizeof(int) * t->cap);
  t->n = 0;
  t->root = -1; /* -1 == HuffmanEmpty */
}
static void tree_free(Tree *t) {
  free(t->kind); free(t->left); free(t->right); free(t->value);
  t->kind = t->left = t->right = t->value = NULL;
}
static int tree_new(Tree *t, int kind, int l, int r, int v) {
  if (t->n == t->cap) {
    t->cap *= 2;
    t->kind = realloc(t->kind, sizeof(int) * t->cap);
    t->left = realloc(t->left, sizeof(int) * t->cap);
    t->right = realloc(t->right, sizeof(int) * t->cap);
    t->value = realloc(t->value, sizeof(int) * t->cap);
  }
  int i = t->n++;
  t->kind[i] = kind; t->left[i] = l; t->right[i] = r; t->value[i] = v;
  return i;
}

/* addHuffmanNode (HuffmanTree.hs:36-71).  `node` = -1 is HuffmanEmpty.  Returns the new
 * subtree in *out, or a detail code > 0 on Left. */
static int add_node(Tree *t, int val, int len, int code, int node, int *out) {
  if (node < 0) { /* HuffmanEmpty */
    if (len == 0) { *out = tree_new(t, T_VALUE, -1, -1, val); return 0; }
    int sub;
    int e = add_node(t, val, len - 1, code, -1, &sub);
    if (e) return e;
    if ((code >> (len - 1)) & 1) *out = tree_new(t, T_NODE, -1, sub, 0);
    else *out = tree_new(t, T_NODE, sub, -1, 0);
    return 0;
  }
  if (t->kind[node] == T_VALUE) return len == 0 ? D_TWO_VALUES : D_VALUE_HIT;
  /* HuffmanNode l r */
  if (len == 0) return D_LEAF_IS_NODE;
  int sub;
  if ((code >> (len - 1)) & 1) {
    int e = add_node(t, val, len - 1, code, t->right[node], &sub);
    if (e) return e;
    t->right[node] = sub; /* persistent update is unobservable: the old tree is dropped */
  } else {
    int e = add_node(t, val, len - 1, code, t->left[node], &sub);
    if (e) return e;
    t->left[node] = sub;
  }
  *out = node;
  return 0;
}

/* computeCodeValues (Deflate.hs:261-288): canonical codes, result ascending by symbol.
 * syms need not be sorted on input.  Returns the number of triples. */
typedef struct { int sym, len, code; } Triple;
static int cmp_triple(const void *a, const void *b) {
  return ((const Triple *)a)->sym - ((const Triple *)b)->sym;
}
static int compute_code_values(const int *sym, const int *len, int n, Triple *out) {
  int m = 0;
  for (int i = 0; i < n; i++) /* valsNo0s */
    if (len[i] != 0) { out[m].sym = sym[i]; out[m].len = len[i]; out[m].code = 0; m++; }
  if (m == 0) return 0;                 /* codeTree = step3 [] .. = empty: maxBits never forced */
  qsort(out, m, sizeof(Triple), cmp_triple); /* valsSort */
  int max_bits = 0;
  for (int i = 0; i < m; i++) if (out[i].len > max_bits) max_bits = out[i].len;
  long *bl_count = calloc(max_bits + 2, sizeof(long));
  long *next_code = calloc(max_bits + 2, sizeof(long));
  for (int i = 0; i < m; i++) bl_count[out[i].len]++;
  long code = 0;
  next_code[0] = 0;
  for (int bits = 1; bits <= max_bits; bits++) { /* step2 */
    long prev = bl_count[bits - 1]; /* bl_count[0] == 0: zero lengths were filtered */
    code = (code + prev) << 1;
    next_code[bits] = code;
  }
  for (int i = 0; i < m; i++) { /* step3, ascending symbol order */
    out[i].code = (int)next_code[out[i].len];
    next_code[out[i].len]++;
  }
  free(bl_count); free(next_code);
  return m;
}

/* computeHuffmanTree = createHuffmanTree . computeCodeValues (Deflate.hs:255-259).
 * createHuffmanTree is a foldr (HuffmanTree.hs:29-34): the LAST triple is inserted first. */
static int build_tree(Tree *t, const int *sym, const int *len, int n, int64_t *errval) {
  Triple *tr = malloc(sizeof(Triple) * (n > 0 ? n : 1));
  int m = compute_code_values(sym, len, n, tr);
  t->n = 0;
  t->root = -1;
  for (int i = m - 1; i >= 0; i--) {
    int nr;
    int e = add_node(t, tr[i].sym, tr[i].len, tr[i].code, t->root, &nr);
    if (e) { if (errval) *errval = tr[i].sym; free(tr); return e; }
    t->root = nr;
  }
  free(tr);
  return 0;
}

/* ---------------------------------------------------------------- decoder state */

typedef struct {
  /* the lazy input: a list of strict chunks (Zlib.hs:35) */
  const uint8_t *in;
  const size_t *chunk_len;
  size_t nchunks, next_chunk;
  size_t next_chunk_off;
  /* DecompressionState (Monad.hs:68-74) */
  int bitno;           /* dcsNextBitNo */
  uint8_t cur;         /* dcsCurByte   */
  uint32_t a, b;       /* dcsAdler32   */
  int framing;         /* FRAME_*: gzip / raw-deflate are extensions beyond the reference (its README's TODO, lines 42-50) */
  uint32_t crc;        /* running CRC-32 register (gzip), pre-conditioned with 0xffffffff */
  const uint8_t *inp;  /* dcsInput     */
  size_t inp_len;
  uint8_t *win;        /* dcsOutput: owWindow */
  size_t ow_next;      /*            owNext   */
  /* consumer side */
  uint8_t *out;
  size_t out_cap;
  uint64_t published;
  pzo_event *ev;
  size_t ev_cap, n_ev;
  uint64_t bytes_taken;
  uint8_t *sbuf; /* scratch for one stored block (+1 byte for the getBlock quirk) */
  pzo_result *res;
  jmp_buf jb;
} St;

static void log_event(St *s, int kind, uint64_t len) {
  if (s->ev && s->n_ev < s->ev_cap) {
    s->ev[s->n_ev].kind = kind; s->ev[s->n_ev].pad = 0; s->ev[s->n_ev].len = len;
  }
  s->n_ev++;
}

static uint64_t bitpos(const St *s) {
  return s->bytes_taken * 8 - (uint64_t)(s->bitno < 8 ? 8 - s->bitno : 0);
}

/* raise (Monad.hs:152-154) and the impure-exception exits */
static void raise_(St *s, int status, int detail, int64_t p0, int64_t p1) {
  s->res->status = status; s->res->detail = detail;
  s->res->payload[0] = p0; s->res->payload[1] = p1;
  s->res->err_bitpos = bitpos(s);
  longjmp(s->jb, 1);
}

/* getNextChunk (Monad.hs:185-197) as driven by `run` (Zlib.hs:38-42): yields NeedMore; the
 * driver answers with the next strict chunk; an empty chunk yields NeedMore again; no
 * chunk left is "Ran out of data mid-decompression 2.". */
static void get_next_chunk(St *s) {
  for (;;) {
    log_event(s, PZO_EV_NEED_MORE, 0);
    if (s->next_chunk == s->nchunks) raise_(s, ST_DECOMP, D_RAN_OUT, 0, 0);
    const uint8_t *c = s->in + s->next_chunk_off;
    size_t n = s->chunk_len[s->next_chunk];
    s->next_chunk_off += n;
    s->next_chunk++;
    if (n == 0) continue; /* S.uncons bstr == Nothing -> NeedMore again */
    s->bitno = 0; s->cur = c[0]; s->inp = c + 1; s->inp_len = n - 1;
    s->bytes_taken++;
    return;
  }
}

/* nextBits' (Monad.hs:210-230): LSB-first, at most the rest of the current byte per step */
static unsigned next_bits(St *s, int x) {
  unsigned acc = 0;
  int shift = 0;
  while (x != 0) {
    if (s->bitno == 8) {
      if (s->inp_len == 0) { get_next_chunk(s); continue; }
      s->bitno = 0; s->cur = s->inp[0]; s->inp++; s->inp_len--; s->bytes_taken++;
      continue;
    }
    int my = x < 8 - s->bitno ? x : 8 - s->bitno;
    unsigned base = (unsigned)s->cur >> s->bitno;
    unsigned mask = ~(0xFFu << my) & 0xFFu;
    acc |= (base & mask) << shift;
    s->bitno += my;
    x -= my; shift += my;
  }
  return acc;
}

/* nextByte (Monad.hs:232-249) */
static uint8_t next_byte(St *s) {
  for (;;) {
    if (s->bitno == 0) { s->bitno = 8; return s->cur; }
    if (s->bitno != 8) return (uint8_t)next_bits(s, 8);
    if (s->inp_len == 0) { get_next_chunk(s); continue; }
    s->bitno = 8; s->cur = s->inp[0]; s->inp++; s->inp_len--; s->bytes_taken++;
    return s->cur;
  }
}
/* nextWord16 little-endian (Monad.hs:251-255), nextWord32 big-endian (Monad.hs:257-263) */
static unsigned next_word16(St *s) { unsigned lo = next_byte(s); unsigned hi = next_byte(s); return (hi << 8) | lo; }
static uint32_t next_word32(St *s) {
  uint32_t a = next_byte(s), b = next_byte(s), c = next_byte(s), d = next_byte(s);
  return (a << 24) | (b << 16) | (c << 8) | d;
}

/* nextCode / advanceTree (Monad.hs:295-302, HuffmanTree.hs:73-83) */
static int next_code(St *s, const Tree *t) {
  int node = t->root;
  for (;;) {
    unsigned b = next_bits(s, 1);
    if (node < 0) raise_(s, ST_HUFF, D_ADV_EMPTY_TREE, 0, 0);
    /* a HuffmanValue at the root cannot be built (zero lengths are dropped) */
    int nx = b ? t->right[node] : t->left[node];
    if (nx < 0) raise_(s, ST_HUFF, D_ADV_TO_EMPTY, 0, 0);
    if (t->kind[nx] == T_VALUE) return t->value[nx];
    node = nx;
  }
}

/* ---------------------------------------------------------------- Adler32.hs */
/* CRC-32 (RFC 1952 section 8: reflected polynomial 0xedb88320), bit-serial: slow on purpose, obviously right */
static void crc_bytes(St *s, const uint8_t *p, size_t n) {
  uint32_t c = s->crc;
  for (size_t i = 0; i < n; i++) {
    c ^= p[i];
    for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0xedb88320u & (0u - (c & 1u)));
  }
  s->crc = c;
}
static void adler_byte(St *s, uint8_t v) { /* advanceAdler :22-27 */
  if (s->framing == FRAME_GZIP) crc_bytes(s, &v, 1);
  s->a = (s->a + v) % ADLER_MOD;
  s->b = (s->b + s->a) % ADLER_MOD;
}
static void adler_block(St *s, const uint8_t *p, size_t n) { /* advanceAdlerBlock :44-51 */
  if (s->framing == FRAME_GZIP) crc_bytes(s, p, n);
  while (n >= 5552) {
    uint64_t a = s->a, b = s->b;
    for (size_t i = 0; i < 5551; i++) { a += p[i]; b += a; }
    s->a = (uint32_t)(a % ADLER_MOD); s->b = (uint32_t)(b % ADLER_MOD);
    p += 5551; n -= 5551;
  }
  if (n == 0) return;
  if (n == 1) { /* advanceAdler on the one byte (the CRC has already taken it above: found by tools/fuzz_gpu.py on a 1-byte stored block) */
    s->a = (s->a + p[0]) % ADLER_MOD;
    s->b = (s->b + s->a) % ADLER_MOD;
    return;
  }
  uint64_t a = s->a, b = s->b; /* advanceAdlerLimited :37-42 */
  for (size_t i = 0; i < n; i++) { a += p[i]; b += a; }
  s->a = (uint32_t)(a % ADLER_MOD); s->b = (uint32_t)(b % ADLER_MOD);
}

/* ---------------------------------------------------------------- OutputWindow.hs */
static void publish(St *s, const uint8_t *p, size_t n) { /* Monad.hs:355-358 + Zlib.hs:43-45 */
  log_event(s, PZO_EV_CHUNK, n);
  for (size_t i = 0; i < n; i++)
    if (s->published + i < s->out_cap) s->out[s->published + i] = p[i];
  s->published += n;
}
static void add_byte(St *s, uint8_t b) { /* addByte :64-68, MV.write is bounds-checked */
  if (s->ow_next >= WINDOW_SIZE) raise_(s, ST_BOTTOM, D_BOT_WINDOW_OVERFLOW, 0, 0);
  s->win[s->ow_next++] = b;
}
static void add_chunk(St *s, const uint8_t *p, size_t n) { /* addChunk/copyChunk :70-80 */
  if (s->ow_next + n > WINDOW_SIZE) raise_(s, ST_BOTTOM, D_BOT_WINDOW_OVERFLOW, 0, 0);
  memcpy(s->win + s->ow_next, p, n);
  s->ow_next += n;
}
static const uint8_t *add_old_chunk(St *s, long dist, long len) { /* addOldChunk :82-89 */
  long next = (long)s->ow_next;
  /* copyChunked forces `MV.length src` first: the source slice is checked first */
  if (next - dist < 0) raise_(s, ST_BOTTOM, D_BOT_DIST_TOO_FAR, dist, next);
  if (next - dist + len > WINDOW_SIZE) raise_(s, ST_BOTTOM, D_BOT_WINDOW_OVERFLOW, 0, 0);
  if (next + len > WINDOW_SIZE) raise_(s, ST_BOTTOM, D_BOT_WINDOW_OVERFLOW, 0, 0);
  uint8_t *dest = s->win + next, *src = s->win + next - dist;
  long copied = 0, to_copy = len; /* copyChunked :94-101 */
  while (to_copy != 0) {
    long k = to_copy < dist ? to_copy : dist;
    memmove(dest + copied, src + copied, (size_t)k);
    copied += k; to_copy -= k;
  }
  s->ow_next = (size_t)(next + len);
  return dest;
}
static void move_window(St *s) { /* moveWindow Monad.hs:338-347 / emitExcess :45-54 */
  if (s->ow_next < EXCESS_CHUNK * 2) return;
  publish(s, s->win, EXCESS_CHUNK);
  size_t excess = s->ow_next - EXCESS_CHUNK;
  memmove(s->win, s->win + EXCESS_CHUNK, excess);
  s->ow_next = excess;
}

/* ---------------------------------------------------------------- Monad.hs emitters */
static void emit_byte(St *s, uint8_t b) { add_byte(s, b); adler_byte(s, b); }       /* :309-315 */
static void emit_past_chunk(St *s, long dist, long len) {                             /* :324-333 */
  const uint8_t *p = add_old_chunk(s, dist, len);
  adler_block(s, p, (size_t)len);
}

/* nextBlock (Monad.hs:265-293) followed by emitBlock (Monad.hs:317-322).  Reproduces the
 * chunk-boundary behaviour of getBlock: the fast path needs len < remaining (strict). */
static void stored_block(St *s, unsigned len16) {
  uint8_t *buf = s->sbuf;
  size_t nbuf = 0;
  long len = (long)len16;
  /* nextBitNo is always 8 here (nextByte leaves it at 8), so only getBlock is reachable */
  for (;;) {
    /* raw deflate (extension): nothing follows the last block, so a stored block may end exactly at the end of the
     * input; there the quirk above would turn every such stream into "Ran out of data" */
    if (len < (long)s->inp_len || (s->framing == FRAME_RAW && len == (long)s->inp_len)) {
      size_t take = len < 0 ? 0 : (size_t)len; /* S.splitAt of a negative count takes nothing */
      memcpy(buf + nbuf, s->inp, take); nbuf += take;
      s->inp += take; s->inp_len -= take; s->bytes_taken += take;
      s->bitno = 8;
      break;
    } else if (s->inp_len == 0) {
      get_next_chunk(s);
      buf[nbuf++] = s->cur; /* byte1 <- dcsCurByte; consumed as block data */
      len -= 1;
    } else {
      memcpy(buf + nbuf, s->inp, s->inp_len); nbuf += s->inp_len;
      len -= (long)s->inp_len;
      s->bytes_taken += s->inp_len;
      s->inp += s->inp_len; s->inp_len = 0; /* recursion continues on S.empty */
    }
  }
  add_chunk(s, buf, nbuf);
  adler_block(s, buf, nbuf);
}

/* ---------------------------------------------------------------- Deflate.hs */
static const int LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const int LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const int DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const int DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static const int CODE_LENGTH_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

static void compute_huffman_tree(St *s, Tree *t, const int *sym, const int *len, int n) {
  int64_t v = 0;
  int e = build_tree(t, sym, len, n, &v);
  if (e) raise_(s, ST_HUFF, e, e == D_LEAF_IS_NODE ? v : 0, 0);
}

/* runInflate (Deflate.hs:106-120) */
static void run_inflate(St *s, const Tree *lit, const Tree *dist) {
  for (;;) {
    int code = next_code(s, lit);
    if (code < 256) { emit_byte(s, (uint8_t)code); continue; }
    if (code == 256) return;
    /* getLength: lengthArray ! code, bounds (257,285) (Deflate.hs:160-197) */
    if (code > 285) raise_(s, ST_BOTTOM, D_BOT_LENGTH_SYM, code, 0);
    long len = LEN_BASE[code - 257] + (long)next_bits(s, LEN_EXTRA[code - 257]);
    int dcode = next_code(s, dist);
    /* getDistance: distanceArray ! dcode, bounds (0,29) (Deflate.hs:199-237) */
    if (dcode > 29) raise_(s, ST_BOTTOM, D_BOT_DIST_SYM, dcode, 0);
    long d = DIST_BASE[dcode] + (long)next_bits(s, DIST_EXTRA[dcode]);
    emit_past_chunk(s, d, len);
    move_window(s);
  }
}

/* inflateBlock (Deflate.hs:65-104) + getCodeLengths (Deflate.hs:124-156) */
static int inflate_block(St *s, const Tree *fixed_lit, const Tree *fixed_dist, Tree *t_code, Tree *t_lit, Tree *t_dist) {
  int bfinal = next_bits(s, 1) == 1;
  unsigned btype = next_bits(s, 2);
  if (btype == 0) {
    s->bitno = 8; /* advanceToByte (Monad.hs:304-307) */
    unsigned len = next_word16(s);
    unsigned nlen = next_word16(s);
    if (len != ((~nlen) & 0xFFFFu)) raise_(s, ST_FORMAT, D_LEN_NLEN, 0, 0);
    stored_block(s, len);
    return bfinal;
  }
  if (btype == 1) { run_inflate(s, fixed_lit, fixed_dist); return bfinal; }
  if (btype == 2) {
    int hlit = 257 + (int)next_bits(s, 5);
    int hdist = 1 + (int)next_bits(s, 5);
    int hclen = 4 + (int)next_bits(s, 4);
    int csym[19], clen[19];
    for (int i = 0; i < hclen; i++) { csym[i] = CODE_LENGTH_ORDER[i]; clen[i] = (int)next_bits(s, 3); }
    compute_huffman_tree(s, t_code, csym, clen, hclen);
    /* lens: keys 0.. may overshoot hlit+hdist by up to 137 (repeat codes are not clipped) */
    int lens[320 + 140];
    int have[320 + 140];
    memset(lens, 0, sizeof lens); memset(have, 0, sizeof have);
    int n = 0, maxl = hlit + hdist, prev = 0, top = 0;
    while (n < maxl) {
      int code = next_code(s, t_code);
      if (code <= 15) { lens[n] = code; have[n] = 1; n += 1; prev = code; }
      else {
        int num, val;
        if (code == 16) { num = 3 + (int)next_bits(s, 2); val = prev; }
        else if (code == 17) { num = 3 + (int)next_bits(s, 3); val = 0; prev = 0; }
        else { num = 11 + (int)next_bits(s, 7); val = 0; prev = 0; }
        for (int i = 0; i < num; i++) { lens[n + i] = val; have[n + i] = 1; }
        n += num;
      }
      if (n > top) top = n;
    }
    /* partition at hlit; distance keys are shifted down (Deflate.hs:95-97) */
    int lsym[288 + 8], llen[288 + 8], nl = 0;
    int dsym[320 + 140], dlen[320 + 140], nd = 0;
    for (int k = 0; k < top; k++) {
      if (!have[k]) continue;
      if (k < hlit) { lsym[nl] = k; llen[nl] = lens[k]; nl++; }
      else { dsym[nd] = k - hlit; dlen[nd] = lens[k]; nd++; }
    }
    compute_huffman_tree(s, t_lit, lsym, llen, nl);
    compute_huffman_tree(s, t_dist, dsym, dlen, nd);
    run_inflate(s, t_lit, t_dist);
    return bfinal;
  }
  raise_(s, ST_FORMAT, D_BAD_BTYPE, (int64_t)btype, 0);
  return 0;
}

/* inflateWithHeaders (Zlib.hs:53-69) then inflate (Deflate.hs:39-63) */
/* EXTENSION (not in the reference; its README lists gzip as the first TODO): the member header of RFC 1952 section 2.3,
 * parsed in the reference's style -- byte reads that run into the truncation verdict, checks in stream order, optional
 * fields skipped the way inflateWithHeaders skips FDICT.  FHCRC is skipped, not verified. */
static void gzip_header(St *s) {
  unsigned id1 = next_byte(s), id2 = next_byte(s);
  if (id1 != 0x1f || id2 != 0x8b) raise_(s, ST_HEADER, D_HDR_GZIP_MAGIC, (int64_t)((id1 << 8) | id2), 0);
  unsigned cm = next_byte(s);
  if (cm != 8) raise_(s, ST_HEADER, D_HDR_METHOD, cm, 0);
  unsigned flg = next_byte(s);
  if (flg & 0xe0) raise_(s, ST_HEADER, D_HDR_GZIP_FLAGS, flg, 0);
  for (int i = 0; i < 6; i++) (void)next_byte(s); /* MTIME, XFL, OS */
  if (flg & 4) { unsigned xlen = next_word16(s); for (unsigned i = 0; i < xlen; i++) (void)next_byte(s); }
  if (flg & 8) while (next_byte(s) != 0) {}
  if (flg & 16) while (next_byte(s) != 0) {}
  if (flg & 2) { (void)next_byte(s); (void)next_byte(s); }
}

static void inflate_with_headers(St *s, Tree *fixed_lit, Tree *fixed_dist, Tree *t_code, Tree *t_lit, Tree *t_dist) {
  if (s->framing == FRAME_GZIP) {
    gzip_header(s);
  } else if (s->framing == FRAME_ZLIB) {
    unsigned cmf = next_byte(s);
    unsigned flg = next_byte(s);
    unsigned both = (cmf << 8) | flg;
    unsigned cm = cmf & 0x0f, cinfo = cmf >> 4;
    int fdict = (flg >> 5) & 1;
    if (both % 31 != 0) raise_(s, ST_HEADER, D_HDR_CHECKSUM, 0, 0);
    if (cm != 8) raise_(s, ST_HEADER, D_HDR_METHOD, cm, 0);
    if (cinfo > 7) raise_(s, ST_HEADER, D_HDR_WINDOW, cinfo, 0);
    if (fdict) for (int i = 0; i < 4; i++) (void)next_byte(s);
  }
  /* buildFixedLitTree / buildFixedDistanceTree (Deflate.hs:241-251) */
  {
    int sym[288], len[288];
    for (int i = 0; i < 288; i++) { sym[i] = i; len[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8; }
    compute_huffman_tree(s, fixed_lit, sym, len, 288);
    for (int i = 0; i < 32; i++) { sym[i] = i; len[i] = 5; }
    compute_huffman_tree(s, fixed_dist, sym, len, 32);
  }
  for (;;) {
    int is_final = inflate_block(s, fixed_lit, fixed_dist, t_code, t_lit, t_dist);
    move_window(s);
    if (is_final) break;
  }
  /* checkChecksum (Deflate.hs:52-63) */
  s->bitno = 8;
  uint32_t ours = (s->b << 16) | s->a; /* finalizeAdler (Adler32.hs:53-57) */
  s->res->adler_computed = ours;
  if (s->framing == FRAME_GZIP) { /* RFC 1952 2.3.1: CRC32 then ISIZE, least significant byte first.  The WHOLE trailer is read
                                     before anything is compared (a member cut inside its trailer is "ran out of data", whatever
                                     its CRC: an incremental decoder must ask for more, not judge half a trailer); then the CRC,
                                     then the length */
    ours = ~s->crc;
    s->res->adler_computed = ours;
    uint32_t lo = next_word16(s), hi = next_word16(s);
    uint32_t theirs = (hi << 16) | lo;
    s->res->adler_stored = theirs;
    lo = next_word16(s); hi = next_word16(s);
    uint32_t isize = (hi << 16) | lo, total = (uint32_t)(s->published + s->ow_next);
    if (theirs != ours) raise_(s, ST_CHECKSUM, D_ADLER_MISMATCH, 0, 0);
    if (isize != total) raise_(s, ST_CHECKSUM, D_LENGTH_MISMATCH, isize, 0);
  } else if (s->framing == FRAME_ZLIB) {
    uint32_t theirs = next_word32(s);
    s->res->adler_stored = theirs;
    if (theirs != ours) raise_(s, ST_CHECKSUM, D_ADLER_MISMATCH, 0, 0);
  } /* raw deflate: no trailer, nothing to compare (adler_computed is still reported) */
  /* finalize (Monad.hs:349-353) */
  publish(s, s->win, s->ow_next);
  s->ow_next = 0; /* the window is not reused; zero so out_len below is not double counted */
}

int pzo_decompress(const uint8_t *in, const size_t *chunk_len, size_t nchunks, uint8_t *out,
                   size_t out_cap, pzo_result *res, pzo_event *ev, size_t ev_cap, size_t *n_ev,
                   uint64_t *published) {
  return pzo_decompress_framed(in, chunk_len, nchunks, out, out_cap, res, ev, ev_cap, n_ev, published, FRAME_ZLIB);
}

int pzo_decompress_framed(const uint8_t *in, const size_t *chunk_len, size_t nchunks, uint8_t *out,
                          size_t out_cap, pzo_result *res, pzo_event *ev, size_t ev_cap, size_t *n_ev,
                          uint64_t *published, int framing) {
  St *s = calloc(1, sizeof(St));
  s->framing = framing; s->crc = 0xffffffffu;
  Tree fl, fd, tc, tl, td;
  tree_init(&fl); tree_init(&fd); tree_init(&tc); tree_init(&tl); tree_init(&td);
  memset(res, 0, sizeof *res);
  s->in = in; s->chunk_len = chunk_len; s->nchunks = nchunks;
  s->bitno = 8; s->cur = 0; s->a = 1; s->b = 0; /* runDeflateM (Monad.hs:169-181) */
  s->inp = NULL; s->inp_len = 0;
  s->win = malloc(WINDOW_SIZE);
  s->sbuf = malloc(65536 + 8);
  s->out = out; s->out_cap = out_cap;
  s->ev = ev; s->ev_cap = ev_cap;
  s->res = res;
  if (setjmp(s->jb) == 0) {
    inflate_with_headers(s, &fl, &fd, &tc, &tl, &td);
    log_event(s, PZO_EV_DONE, 0);
    res->status = ST_OK;
    res->err_bitpos = bitpos(s);
    if (s->next_chunk < s->nchunks) { /* run Done (_:_) (Zlib.hs:48-49) */
      res->status = ST_DECOMP; res->detail = D_DATA_REMAINING;
    }
  } else {
    log_event(s, PZO_EV_ERROR, 0);
    /* bytes decoded but never published stay visible to the test-suite after `published` */
    for (size_t i = 0; i < s->ow_next; i++)
      if (s->published + i < out_cap) out[s->published + i] = s->win[i];
  }
  res->out_len = s->published + s->ow_next;
  if (res->status != ST_OK && res->status != ST_CHECKSUM && !(res->status == ST_DECOMP && res->detail == D_DATA_REMAINING))
    res->adler_computed = framing == FRAME_GZIP ? ~s->crc : (s->b << 16) | s->a;
  if (n_ev) *n_ev = s->n_ev;
  if (published) *published = s->published;
  tree_free(&fl); tree_free(&fd); tree_free(&tc); tree_free(&tl); tree_free(&td);
  free(s->win); free(s->sbuf); free(s);
  return 0;
}

int pzo_compute_code_values(const int32_t *sym, const int32_t *len, int n, int32_t *out_triples) {
  Triple *tr = malloc(sizeof(Triple) * (n > 0 ? n : 1));
  int m = compute_code_values((const int *)sym, (const int *)len, n, tr);
  for (int i = 0; i < m; i++) { out_triples[3 * i] = tr[i].sym; out_triples[3 * i + 1] = tr[i].len; out_triples[3 * i + 2] = tr[i].code; }
  free(tr);
  return m;
}

int pzo_tree_check(const uint8_t *lens, int n, int64_t *val) {
  int *sym = malloc(sizeof(int) * (n > 0 ? n : 1)), *len = malloc(sizeof(int) * (n > 0 ? n : 1));
  for (int i = 0; i < n; i++) { sym[i] = i; len[i] = lens[i]; }
  Tree t; tree_init(&t);
  int64_t v = 0;
  int e = build_tree(&t, sym, len, n, &v);
  if (val) *val = e == D_LEAF_IS_NODE ? v : 0;
  tree_free(&t); free(sym); free(len);
  return e;
}

uint32_t pzo_adler32(uint32_t init, const uint8_t *data, size_t len) {
  St s; memset(&s, 0, sizeof s);
  s.a = init & 0xffff; s.b = init >> 16;
  adler_block(&s, data, len);
  return (s.b << 16) | s.a;
}

size_t pzo_strerror(const pzo_result *r, char *buf, size_t cap) {
  char tmp[256];
  tmp[0] = 0;
  switch (r->status) {
  case ST_OK: break;
  case ST_HUFF: {
    const char *m = "?";
    char t2[96];
    if (r->detail == D_TWO_VALUES) m = "Two values point to the same place!";
    else if (r->detail == D_VALUE_HIT) m = "HuffmanValue hit while inserting a value!";
    else if (r->detail == D_LEAF_IS_NODE) { snprintf(t2, sizeof t2, "Tried to add where the leaf is a node: %lld", (long long)r->payload[0]); m = t2; }
    else if (r->detail == D_ADV_EMPTY_TREE) m = "Tried to advance empty tree!";
    else if (r->detail == D_ADV_TO_EMPTY) m = "Advanced to empty tree!";
    snprintf(tmp, sizeof tmp, "Huffman tree manipulation error: %s", m);
    break; }
  case ST_FORMAT:
    if (r->detail == D_LEN_NLEN) snprintf(tmp, sizeof tmp, "Block format error: Len/nlen mismatch in uncompressed block.");
    else snprintf(tmp, sizeof tmp, "Block format error: Unacceptable BTYPE: %lld", (long long)r->payload[0]);
    break;
  case ST_DECOMP:
    snprintf(tmp, sizeof tmp, "Decompression error: %s", r->detail == D_RAN_OUT ? "Ran out of data mid-decompression 2." : "Finished with data remaining.");
    break;
  case ST_HEADER:
    if (r->detail == D_HDR_CHECKSUM) snprintf(tmp, sizeof tmp, "Header error: Header checksum failed");
    else if (r->detail == D_HDR_METHOD) snprintf(tmp, sizeof tmp, "Header error: Bad compression method: %lld", (long long)r->payload[0]);
    else if (r->detail == D_HDR_GZIP_MAGIC) snprintf(tmp, sizeof tmp, "Header error: Not a gzip stream: %llx", (long long)r->payload[0]);
    else if (r->detail == D_HDR_GZIP_FLAGS) snprintf(tmp, sizeof tmp, "Header error: Reserved gzip flags set: %lld", (long long)r->payload[0]);
    else snprintf(tmp, sizeof tmp, "Header error: Window size too big: %lld", (long long)r->payload[0]);
    break;
  case ST_CHECKSUM:
    if (r->detail == D_LENGTH_MISMATCH) snprintf(tmp, sizeof tmp, "Checksum error: length mismatch: %lld != %lld", (long long)r->payload[0], (long long)(uint32_t)r->out_len);
    else snprintf(tmp, sizeof tmp, "Checksum error: checksum mismatch: %x != %x", r->adler_stored, r->adler_computed);
    break;
  case ST_BOTTOM:
    snprintf(tmp, sizeof tmp, "_|_ %s", r->detail == D_BOT_LENGTH_SYM ? "lengthArray index" : r->detail == D_BOT_DIST_SYM ? "distanceArray index" : r->detail == D_BOT_DIST_TOO_FAR ? "negative slice" : "window overflow");
    break;
  default: snprintf(tmp, sizeof tmp, "status %d", r->status);
  }
  size_t n = strlen(tmp);
  if (cap) { size_t k = n < cap - 1 ? n : cap - 1; memcpy(buf, tmp, k); buf[k] = 0; }
  return n;
}
