/*
 * pz_abi.cu -- the C ABI of libpzcuda.so (include/pzcuda.h) and the host driver.
 *
 * The driver replaces the reference's decoder monad as a *driver* (Monad.hs:163-197: the
 * NeedMore/Chunk coroutine) with: pinned staging, slices of the batch pipelined over several
 * CUDA streams (H2D, kernels, D2H overlap), and a resident-batch object whose run() is
 * kernel launches only.  There is no CPU decode path: without a device every call fails
 * with PZ_E_CUDA.
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "pz_internal.h"
#include "pzcuda.h"

namespace {

thread_local std::string g_last_error;

int fail_cuda(cudaError_t e, const char *what) {
  char buf[512];
  snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  g_last_error = buf;
  return PZ_E_CUDA;
}
#define PZ_CUDA(call)                                 \
  do {                                                \
    cudaError_t e__ = (call);                         \
    if (e__ != cudaSuccess) return fail_cuda(e__, #call); \
  } while (0)

std::atomic<uint64_t> g_huge_bytes{4ull << 20}; /* compressed size from which a stream goes to K4 (PZ_HUGE_BYTES overrides) */
std::atomic<uint64_t> g_huge_done{0}, g_huge_declined{0};
std::once_flag g_once;
int g_init_rc = PZ_E_STATE;
int g_device = -1;          /* the primary device: resident batches, incremental contexts, device-pointer calls */
std::vector<int> g_devices; /* every device host-buffer batches are sharded over (g_devices[0] == g_device) */

/* One worker thread per device of a multi-device configuration ("one host thread + CUDA context +
 * pinned staging + stream set per GPU"): the thread binds its device once, owns a thread_local Workspace on it, and
 * runs the shards it is handed.  The calling thread only waits, so its own current device is never touched. */
struct DeviceWorker {
  int dev = 0;
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::function<int()> job;
  bool has_job = false, done = false, stop = false;
  int rc = 0;
  std::string err;
  void loop();
  void submit(std::function<int()> f) {
    std::lock_guard<std::mutex> g(mu);
    job = std::move(f); has_job = true; done = false;
    cv.notify_all();
  }
  int wait(std::string *e) {
    std::unique_lock<std::mutex> g(mu);
    cv.wait(g, [&] { return done; });
    if (rc != PZ_E_OK && e) *e = err;
    return rc;
  }
};
std::vector<DeviceWorker *> g_workers; /* index k serves g_devices[k]; empty in a single-device configuration */
std::mutex g_multi_mu;                 /* one sharded call at a time (the workers hold one job each) */

void DeviceWorker::loop() {
  cudaSetDevice(dev);
  for (;;) {
    std::function<int()> f;
    {
      std::unique_lock<std::mutex> g(mu);
      cv.wait(g, [&] { return has_job || stop; });
      if (stop) return;
      f = std::move(job); has_job = false;
    }
    const int r = f();
    {
      std::lock_guard<std::mutex> g(mu);
      rc = r; err = r != PZ_E_OK ? g_last_error : std::string(); done = true;
    }
    cv.notify_all();
  }
}

void do_init(const pz_config *cfg) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_init_rc = fail_cuda(e != cudaSuccess ? e : cudaErrorNoDevice, "cudaGetDeviceCount");
    return;
  }
  /* the device list: pz_config.devices[], else the environment (PZ_DEVICES = "all" or "0,2,3"), else one device */
  std::vector<int> want;
  if (cfg && cfg->n_devices > 0) {
    for (int k = 0; k < cfg->n_devices && k < PZ_MAX_DEVICES; k++) want.push_back(cfg->devices[k]);
  } else if (cfg && cfg->n_devices < 0) {
    for (int d = 0; d < ndev; d++) want.push_back(d);
  } else if (const char *dv = (cfg && cfg->device >= 0) ? nullptr : getenv("PZ_DEVICES")) {
    if (!strcmp(dv, "all")) { for (int d = 0; d < ndev; d++) want.push_back(d); }
    else for (const char *p = dv; *p;) { char *q; const long d = strtol(p, &q, 10); if (q == p) break; want.push_back((int)d); p = *q ? q + 1 : q; }
  }
  int prev = 0;
  cudaGetDevice(&prev);
  if (want.empty()) want.push_back((cfg && cfg->device >= 0) ? cfg->device : prev);
  for (int d : want)
    if (d < 0 || d >= ndev) { g_last_error = "pz_init: device ordinal out of range"; g_init_rc = PZ_E_ARG; return; }
  for (size_t k = want.size(); k-- > 0;) { /* kernel attributes belong to a device's context; the primary device is bound last */
    e = cudaSetDevice(want[k]);
    if (e != cudaSuccess) { g_init_rc = fail_cuda(e, "cudaSetDevice"); return; }
    e = pz_kernels_configure();
    if (e != cudaSuccess) { g_init_rc = fail_cuda(e, "pz_kernels_configure"); return; }
  }
  g_devices = want;
  g_device = want[0];
  if (cfg == nullptr || cfg->device < 0) {
    /* implicit initialisation keeps the caller's current device when it is the primary one (the usual case: one device) */
    if (want.size() > 1 || want[0] != prev) cudaSetDevice(want.size() > 1 ? prev : want[0]);
  }
  if (want.size() > 1) {
    for (size_t k = 0; k < want.size(); k++) {
      DeviceWorker *w = new DeviceWorker();
      w->dev = want[k];
      w->th = std::thread([w] { w->loop(); });
      g_workers.push_back(w);
    }
  }
  if (const char *h = getenv("PZ_HUGE_BYTES")) { const long long v = atoll(h); if (v > 0) g_huge_bytes = (uint64_t)v; }
  g_init_rc = PZ_E_OK;
}

int ensure_init() {
  std::call_once(g_once, do_init, (const pz_config *)nullptr);
  if (g_init_rc != PZ_E_OK && g_last_error.empty()) g_last_error = "libpzcuda: initialisation failed (no usable CUDA device)";
  return g_init_rc;
}

inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }
/* PZ_F_GZIP / PZ_F_RAW -> PZ_FRAME_* of pz_device.cuh (0 zlib, 1 gzip, 2 raw deflate) */
inline uint32_t framing_of(uint32_t flags) { return (flags & PZ_F_GZIP) ? 1u : (flags & PZ_F_RAW) ? 2u : 0u; }

bool is_device_ptr(const void *p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

/* The device-side address of [p, p+len) if that host range is pinned and mapped (cudaHostAlloc /
 * cudaHostRegister), else nullptr. */
const uint8_t *mapped_device_ptr(const uint8_t *p, uint64_t len) {
  cudaPointerAttributes a0, a1;
  if (cudaPointerGetAttributes(&a0, p) != cudaSuccess || cudaPointerGetAttributes(&a1, p + (len ? len - 1 : 0)) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  if (a0.type != cudaMemoryTypeHost || a1.type != cudaMemoryTypeHost || !a0.devicePointer || !a1.devicePointer) return nullptr;
  if ((const uint8_t *)a1.devicePointer - (const uint8_t *)a0.devicePointer != (ptrdiff_t)(len ? len - 1 : 0)) return nullptr;
  return (const uint8_t *)a0.devicePointer;
}

/* Grow-only device / pinned buffers kept per host thread, so repeated calls do not pay
 * cudaMalloc.  Freed at thread exit. */
struct Buf {
  void *p = nullptr;
  size_t cap = 0;
  bool pinned = false;
  int reserve(size_t n) {
    if (n <= cap) return PZ_E_OK;
    release();
    size_t want = align_up(n + 256, 1 << 20);
    cudaError_t e = pinned ? cudaHostAlloc(&p, want, cudaHostAllocPortable) : cudaMalloc(&p, want); /* portable: every device of a multi-device configuration copies from / to it */
    if (e != cudaSuccess) { p = nullptr; cap = 0; fail_cuda(e, pinned ? "cudaHostAlloc" : "cudaMalloc"); return PZ_E_NOMEM; }
    cap = want;
    return PZ_E_OK;
  }
  void release() {
    if (p) { if (pinned) cudaFreeHost(p); else cudaFree(p); }
    p = nullptr; cap = 0;
  }
};

constexpr int kStreams = 4;
constexpr int kSlices = 8; /* slices of a host batch whose copies and kernels are pipelined */
constexpr int kGroups = 8; /* pieces a host batch travels in (progressive input) */
constexpr uint64_t kColumnBytes = 32768; /* granularity of the progress words (PZ_PROG_SHIFT) */
constexpr size_t PZ_EXCESS_CHUNK = 32768; /* excessChunkSize (OutputWindow.hs:42-43) */

struct Workspace {
  Buf d_in, d_out, d_in_off, d_out_off, d_seg_off, d_res, d_parts;
  Buf h_in, h_out; /* pinned staging for the pointer-array entry point */
  Buf h_res;       /* pinned landing zone for the verdicts: a D2H copy into pageable memory would block the host */
  Buf h_prog;      /* pinned, mapped: the kernel's progress words (PzJob::prog) */
  Buf d_ready;     /* device word: PzJob::in_ready */
  Buf k4_cand, k4_keep, k4_start, k4_off, k4_len, k4_res, k4_sym, k4_word, k4_grp, k4_gfirst, k4_gw, k4_scr, k4_src; /* K4 scratch (one huge stream at a time) */
  cudaStream_t streams[kStreams] = {};
  bool have_streams = false;
  Workspace() { h_in.pinned = true; h_out.pinned = true; h_res.pinned = true; h_prog.pinned = true; }
  int ensure_streams() {
    if (have_streams) return PZ_E_OK;
    for (int i = 0; i < kStreams; i++) PZ_CUDA(cudaStreamCreateWithFlags(&streams[i], cudaStreamNonBlocking));
    have_streams = true;
    return PZ_E_OK;
  }
  ~Workspace() {
    /* the CUDA context may already be gone at process exit; errors are ignored */
    d_in.release(); d_out.release(); d_in_off.release(); d_out_off.release(); d_seg_off.release();
    d_res.release(); d_parts.release(); h_in.release(); h_out.release(); h_res.release(); h_prog.release(); d_ready.release();
    k4_cand.release(); k4_keep.release(); k4_start.release(); k4_off.release(); k4_len.release(); k4_res.release(); k4_sym.release(); k4_word.release(); k4_grp.release(); k4_gfirst.release(); k4_gw.release(); k4_scr.release(); k4_src.release();
    if (have_streams) for (int i = 0; i < kStreams; i++) cudaStreamDestroy(streams[i]);
  }
};
thread_local Workspace g_ws;

/* Host-side copies of a batch (packing pointer arrays into pinned staging and back): one thread moves about 10 GB/s,
 * a 1 GiB batch would spend 100 ms there, several times the decode.  Jobs are dealt to up to 16 threads by bytes. */
struct CopyJob { void *dst; const void *src; size_t len; };
void parallel_copy(const std::vector<CopyJob> &jobs) {
  uint64_t total = 0;
  for (const CopyJob &j : jobs) total += j.len;
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  const unsigned nth = (unsigned)std::min<uint64_t>(std::min(16u, hw), total / (4u << 20) + 1u);
  if (nth <= 1) {
    for (const CopyJob &j : jobs) if (j.len) memcpy(j.dst, j.src, j.len);
    return;
  }
  /* thread t takes the bytes [total * t / nth, total * (t + 1) / nth) of the concatenation of all jobs */
  auto work = [&](unsigned t) {
    const uint64_t lo = total * t / nth, hi = total * (t + 1) / nth;
    uint64_t at = 0;
    for (const CopyJob &j : jobs) {
      const uint64_t b = at, e = at + j.len;
      at = e;
      if (e <= lo) continue;
      if (b >= hi) break;
      const uint64_t x = std::max(b, lo) - b, y = std::min(e, hi) - b;
      memcpy((uint8_t *)j.dst + x, (const uint8_t *)j.src + x, y - x);
    }
  };
  std::vector<std::thread> th;
  for (unsigned t = 1; t < nth; t++) th.emplace_back(work, t);
  work(0);
  for (std::thread &x : th) x.join();
}

/* seg_off[i] = first checksum segment of stream i; capacities bound the decoded length */
uint64_t build_seg_off(const uint64_t *out_off, size_t n, std::vector<uint64_t> &seg_off) {
  seg_off.resize(n + 1);
  uint64_t acc = 0;
  for (size_t i = 0; i < n; i++) {
    seg_off[i] = acc;
    acc += (out_off[i + 1] - out_off[i] + PZ_ADLER_SEG - 1) / PZ_ADLER_SEG;
  }
  seg_off[n] = acc;
  return acc;
}

bool offsets_ok(const uint64_t *off, size_t n) {
  for (size_t i = 0; i < n; i++) if (off[i + 1] < off[i]) return false;
  return true;
}

/* the kernels keep bit positions in 32 bits: one compressed stream must stay below 512 MiB */
bool in_sizes_ok(const uint64_t *off, size_t n) {
  for (size_t i = 0; i < n; i++) if (off[i + 1] - off[i] > PZ_MAX_STREAM_BYTES) return false;
  return true;
}


/* ---- K4 driver: one huge stream, block-parallel (kernels: pz_huge.cuh, block jobs on K1) --------
 * Returns 1 with *out filled when the stream was decoded, 0 when K4 declines (the caller then leaves
 * the stream to the serial path, which reproduces every verdict of the reference), < 0 on errors.
 * Declining is always safe, so every doubt -- an odd header, a block the search cannot see more
 * than a few times, any verdict other than success inside a block, a reference before the stream's
 * first byte -- ends here. */
constexpr uint32_t kBlockCap = 32u << 20; /* most bytes one speculative block may produce */
constexpr int kMaxGaps = 256; /* blocks the search cannot see (stored, fixed) that K4 will size one by one before giving up */

int huge_stream(const uint8_t *d_in_blob, const uint64_t *d_in_off_i, uint64_t in_byte_off, uint64_t in_len, uint8_t *d_out_i,
                uint64_t out_cap, pz_result *out, cudaStream_t st, uint32_t framing) {
  Workspace &ws = g_ws;
  if (in_len < 16 || in_len > PZ_MAX_STREAM_BYTES) return 0;
  static const bool trace = getenv("PZ_TRACE") != nullptr;
  const auto t_start = std::chrono::steady_clock::now();
  auto now_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count(); };
  const uint8_t *d_stream = d_in_blob + in_byte_off;
  uint8_t head[1024];
  const size_t head_len = (size_t)std::min<uint64_t>(sizeof head, in_len);
  PZ_CUDA(cudaMemcpyAsync(head, d_stream, head_len, cudaMemcpyDeviceToHost, st));
  PZ_CUDA(cudaStreamSynchronize(st));
  uint64_t first_bit = 0;
  uint32_t trailer_bytes = 0;
  if (framing == 0u) {
    const uint32_t cmf = head[0], flg = head[1];
    if (((cmf << 8) | flg) % 31u != 0u || (cmf & 15u) != 8u || (cmf >> 4) > 7u) return 0; /* Zlib.hs:53-69: the serial path words the verdict */
    first_bit = (flg & 0x20u) ? 48u : 16u;
    trailer_bytes = 4;
  } else if (framing == 1u) { /* gzip member header (RFC 1952 2.3): anything but a plain, complete header inside the first KiB is the serial path's */
    if (head[0] != 0x1f || head[1] != 0x8b || head[2] != 8 || (head[3] & 0xe0)) return 0;
    size_t at = 10;
    const uint32_t fl = head[3];
    if (fl & 4u) { if (at + 2 > head_len) return 0; at += 2u + ((size_t)head[at] | ((size_t)head[at + 1] << 8)); }
    for (uint32_t bit = 8u; bit <= 16u; bit <<= 1)
      if (fl & bit) { while (at < head_len && head[at] != 0) at++; at++; }
    if (fl & 2u) at += 2;
    if (at >= head_len) return 0;
    first_bit = at * 8u;
    trailer_bytes = 8;
  }
  const uint64_t total_bits = in_len * 8u, last_bit = total_bits - 8u * trailer_bytes; /* the trailer must follow the last block */
  int rc;
  /* K4a: candidates */
  const uint32_t cap = (uint32_t)(total_bits / 128u + 4096u);
  if ((rc = ws.k4_cand.reserve((size_t)cap * 4u)) != PZ_E_OK || (rc = ws.k4_keep.reserve((size_t)cap * 4u)) != PZ_E_OK || (rc = ws.k4_word.reserve(64)) != PZ_E_OK) return rc;
  uint32_t *d_word = (uint32_t *)ws.k4_word.p; /* [0] candidates of the first stage, [1] error flag of the resolution, [2] unit counter of the block jobs, [3] candidates kept */
  PZ_CUDA(cudaMemsetAsync(d_word, 0, 16, st));
  PZ_CUDA(pz_launch_blk_search(d_stream, in_len, first_bit, last_bit, (uint32_t *)ws.k4_cand.p, d_word, cap, st));
  uint32_t ncand = 0;
  PZ_CUDA(cudaMemcpyAsync(&ncand, d_word, 4, cudaMemcpyDeviceToHost, st));
  PZ_CUDA(cudaStreamSynchronize(st));
  if (ncand > cap) return 0;
  /* the second stage leaves its survivors (about one first-stage candidate in 200) in a list of their own: the host
   * never sees the 2 million first-stage candidates */
  PZ_CUDA(pz_launch_blk_verify(d_stream, in_len, last_bit, (const uint32_t *)ws.k4_cand.p, ncand, (uint32_t *)ws.k4_keep.p, d_word + 3, st));
  uint32_t nkept = 0;
  PZ_CUDA(cudaMemcpyAsync(&nkept, d_word + 3, 4, cudaMemcpyDeviceToHost, st));
  PZ_CUDA(cudaStreamSynchronize(st));
  std::vector<uint32_t> kept(nkept);
  if (nkept) {
    PZ_CUDA(cudaMemcpyAsync(kept.data(), ws.k4_keep.p, (size_t)nkept * 4u, cudaMemcpyDeviceToHost, st));
    PZ_CUDA(cudaStreamSynchronize(st));
  }
  std::vector<uint32_t> starts;
  starts.reserve(nkept + 16);
  if (trace) fprintf(stderr, "[pz-k4] %8.3f ms: search done, %u first-stage candidates, %u kept\n", now_ms(), ncand, nkept);
  starts.push_back((uint32_t)first_bit); /* the first block is known, whatever its type */
  for (uint32_t i = 0; i < nkept; i++) if (kept[i] != first_bit) starts.push_back(kept[i]);
  std::sort(starts.begin(), starts.end());
  std::vector<uint32_t> c_start;
  std::vector<uint32_t> c_len;
  std::vector<uint64_t> c_off;
  uint64_t total = 0, end_bit = 0;
  /* One pass instead of two (sizing, then decoding): every candidate is decoded at once, as 16-bit symbols, into a
   * scratch region of its own whose size is a guess -- 16 x the compressed bytes up to the next candidate -- and the
   * blocks that turn out to be the chain are then moved to their final positions (pz_blk_compact_kernel: 4 bytes of
   * traffic per decoded byte, against a second run of every symbol loop).  If the guess is too small for a chain block,
   * a chain block is not among the candidates, or the scratch does not fit, the two-pass flow below does the job. */
  bool symbols_ready = false;
  static const bool two_pass_only = getenv("PZ_K4_TWO_PASS") != nullptr;
  if (!two_pass_only) {
    const size_t nc = starts.size();
    std::vector<uint32_t> caps(nc);
    std::vector<uint64_t> soff(nc);
    uint64_t scr_total = 0;
    for (size_t j = 0; j < nc; j++) {
      const uint64_t next = j + 1 < nc ? starts[j + 1] : last_bit;
      const uint64_t span = (next - starts[j] + 7u) / 8u;
      const uint64_t cap = (std::min<uint64_t>(kBlockCap, 65536u + 16u * span) + 7u) & ~7ull;
      caps[j] = (uint32_t)cap; soff[j] = scr_total; scr_total += cap;
    }
    /* blocks the search cannot see (stored, fixed, unusual trees) are decoded one by one when the chain reaches
     * them, into a pool behind the candidates' regions */
    const uint64_t gap_pool = 256ull << 20;
    uint64_t gap_used = 0;
    int gaps = 0;
    bool ok = scr_total * 2u <= (48ull << 30);
    if (ok && ws.k4_scr.reserve((scr_total + gap_pool) * 2u + 64u) != PZ_E_OK) ok = false; /* no room: two passes need none */
    if (ok && ws.k4_src.reserve(64) != PZ_E_OK) ok = false;
    if (ok) {
      if ((rc = ws.k4_start.reserve(nc * 4u)) != PZ_E_OK || (rc = ws.k4_len.reserve(nc * 4u)) != PZ_E_OK || (rc = ws.k4_off.reserve(nc * 8u)) != PZ_E_OK ||
          (rc = ws.k4_res.reserve(nc * sizeof(pz_result))) != PZ_E_OK)
        return rc;
      PZ_CUDA(cudaMemcpyAsync(ws.k4_start.p, starts.data(), nc * 4u, cudaMemcpyHostToDevice, st));
      PZ_CUDA(cudaMemcpyAsync(ws.k4_len.p, caps.data(), nc * 4u, cudaMemcpyHostToDevice, st));
      PZ_CUDA(cudaMemcpyAsync(ws.k4_off.p, soff.data(), nc * 8u, cudaMemcpyHostToDevice, st));
      PZ_CUDA(pz_launch_blk_jobs(d_in_blob, d_in_off_i, (const uint32_t *)ws.k4_start.p, (const uint64_t *)ws.k4_off.p, (const uint32_t *)ws.k4_len.p, 0,
                                 (uint16_t *)ws.k4_scr.p, (uint32_t)nc, (pz_result *)ws.k4_res.p, st, d_word + 2));
      std::vector<pz_result> got(nc);
      PZ_CUDA(cudaMemcpyAsync(got.data(), ws.k4_res.p, nc * sizeof(pz_result), cudaMemcpyDeviceToHost, st));
      PZ_CUDA(cudaStreamSynchronize(st));
      if (trace) fprintf(stderr, "[pz-k4] %8.3f ms: %zu candidates decoded into scratch (%.2f GB)\n", now_ms(), nc, scr_total * 2e-9);
      std::vector<uint64_t> c_src;
      uint32_t at = (uint32_t)first_bit;
      for (;;) {
        auto it = std::lower_bound(starts.begin(), starts.end(), at);
        pz_result r;
        uint64_t src;
        if (it != starts.end() && *it == at) {
          r = got[it - starts.begin()];
          src = soff[it - starts.begin()];
        } else { /* a block the search did not see: decode it alone, now that its first bit is known */
          if (++gaps > kMaxGaps) return 0;
          const uint64_t next = it != starts.end() ? *it : last_bit;
          const uint64_t cap = (std::min<uint64_t>(kBlockCap, 65536u + 16u * ((next - at + 7u) / 8u)) + 7u) & ~7ull;
          if (gap_used + cap > gap_pool) { ok = false; break; }
          src = scr_total + gap_used;
          gap_used += cap;
          struct { uint64_t off; uint32_t start, len; } one = {src, at, (uint32_t)cap}; /* k4_src is free until the chain is complete */
          PZ_CUDA(cudaMemcpyAsync(ws.k4_src.p, &one, sizeof one, cudaMemcpyHostToDevice, st));
          const uint8_t *d_one = (const uint8_t *)ws.k4_src.p;
          PZ_CUDA(pz_launch_blk_jobs(d_in_blob, d_in_off_i, (const uint32_t *)(d_one + 8), (const uint64_t *)d_one, (const uint32_t *)(d_one + 12), 0,
                                     (uint16_t *)ws.k4_scr.p, 1, (pz_result *)ws.k4_res.p, st));
          PZ_CUDA(cudaMemcpyAsync(&r, ws.k4_res.p, sizeof r, cudaMemcpyDeviceToHost, st));
          PZ_CUDA(cudaStreamSynchronize(st));
        }
        if (r.status == PZ_OUTPUT_FULL) { ok = false; break; } /* the guess was too small */
        if (r.status != PZ_OK) return 0;
        if (total + r.out_len > out_cap) return 0; /* PZ_OUTPUT_FULL is the serial path's to report */
        c_start.push_back(at); c_off.push_back(total); c_len.push_back((uint32_t)r.out_len); c_src.push_back(src);
        total += r.out_len;
        end_bit = r.err_bitpos;
        if (r.detail != 0) break; /* BFINAL */
        if (end_bit >= last_bit || end_bit > 0xffffffffull) return 0;
        at = (uint32_t)end_bit;
      }
      if (ok) {
        const size_t nb = c_start.size();
        if ((rc = ws.k4_src.reserve(nb * 8u)) != PZ_E_OK || (rc = ws.k4_off.reserve(nb * 8u)) != PZ_E_OK) return rc;
        if (ws.k4_sym.reserve(total * 2u + 64u) != PZ_E_OK) return 0;
        PZ_CUDA(cudaMemcpyAsync(ws.k4_off.p, c_off.data(), nb * 8u, cudaMemcpyHostToDevice, st));
        PZ_CUDA(cudaMemcpyAsync(ws.k4_src.p, c_src.data(), nb * 8u, cudaMemcpyHostToDevice, st));
        PZ_CUDA(pz_launch_blk_compact((const uint16_t *)ws.k4_scr.p, (uint16_t *)ws.k4_sym.p, (const uint64_t *)ws.k4_off.p, (const uint64_t *)ws.k4_src.p,
                                      (uint32_t)nb, total, st));
        PZ_CUDA(cudaStreamSynchronize(st)); /* c_off / c_src leave scope */
        symbols_ready = true;
        if (trace) fprintf(stderr, "[pz-k4] %8.3f ms: chain of %zu blocks, %llu bytes, %d decoded one by one, symbols in place\n", now_ms(), nb, (unsigned long long)total, gaps);
      } else {
        c_start.clear(); c_len.clear(); c_off.clear(); total = 0; end_bit = 0;
        if (trace) fprintf(stderr, "[pz-k4] %8.3f ms: one pass not enough, sizing and decoding separately\n", now_ms());
      }
    }
  }
  if (!symbols_ready) {
  /* sizing pass over every candidate */
  std::vector<pz_result> sized(starts.size());
  auto size_jobs = [&](const uint32_t *h_start, size_t count, pz_result *h_res) -> int {
    int r;
    if ((r = ws.k4_start.reserve(count * 4u)) != PZ_E_OK || (r = ws.k4_res.reserve(count * sizeof(pz_result))) != PZ_E_OK) return r;
    PZ_CUDA(cudaMemcpyAsync(ws.k4_start.p, h_start, count * 4u, cudaMemcpyHostToDevice, st));
    PZ_CUDA(pz_launch_blk_jobs(d_in_blob, d_in_off_i, (const uint32_t *)ws.k4_start.p, nullptr, nullptr, kBlockCap, nullptr, (uint32_t)count,
                               (pz_result *)ws.k4_res.p, st, d_word + 2));
    PZ_CUDA(cudaMemcpyAsync(h_res, ws.k4_res.p, count * sizeof(pz_result), cudaMemcpyDeviceToHost, st));
    PZ_CUDA(cudaStreamSynchronize(st));
    return PZ_E_OK;
  };
  if ((rc = size_jobs(starts.data(), starts.size(), sized.data())) != PZ_E_OK) return rc;
  if (trace) fprintf(stderr, "[pz-k4] %8.3f ms: %zu candidates sized\n", now_ms(), starts.size());
  /* the chain: block k+1 starts at the bit where block k ended */
  uint32_t at = (uint32_t)first_bit;
  int gaps = 0;
  for (;;) {
    pz_result r;
    auto it = std::lower_bound(starts.begin(), starts.end(), at);
    if (it != starts.end() && *it == at) {
      r = sized[it - starts.begin()];
    } else { /* a block the search did not see (fixed or stored block, unusual trees): size it alone */
      if (++gaps > kMaxGaps) return 0;
      if ((rc = size_jobs(&at, 1, &r)) != PZ_E_OK) return rc;
    }
    if (r.status != PZ_OK) return 0;
    if (r.out_len > 0xffffffffull || total + r.out_len > out_cap) return 0; /* PZ_OUTPUT_FULL is the serial path's to report */
    c_start.push_back(at);
    c_off.push_back(total);
    c_len.push_back((uint32_t)r.out_len);
    total += r.out_len;
    end_bit = r.err_bitpos;
    if (r.detail != 0) break; /* BFINAL */
    if (end_bit >= last_bit || end_bit > 0xffffffffull) return 0;
    at = (uint32_t)end_bit;
  }
  if (trace) fprintf(stderr, "[pz-k4] %8.3f ms: chain of %zu blocks, %llu bytes, %d found one by one\n", now_ms(), c_start.size(), (unsigned long long)total, gaps);
  }
  const uint64_t tb = (end_bit + 7u) / 8u;
  if (tb + trailer_bytes > in_len) return 0;
  uint8_t trailer[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (trailer_bytes) PZ_CUDA(cudaMemcpyAsync(trailer, d_stream + tb, trailer_bytes, cudaMemcpyDeviceToHost, st));
  /* decode pass: 16-bit symbols of every chain block, at its final position */
  const size_t nb = c_start.size();
  if ((rc = ws.k4_start.reserve(nb * 4u)) != PZ_E_OK || (rc = ws.k4_len.reserve(nb * 4u)) != PZ_E_OK || (rc = ws.k4_off.reserve(nb * 8u)) != PZ_E_OK ||
      (rc = ws.k4_res.reserve(nb * sizeof(pz_result))) != PZ_E_OK || (rc = ws.k4_sym.reserve(total * 2u + 64u)) != PZ_E_OK)
    return rc == PZ_E_NOMEM ? 0 : rc; /* no room for the symbol buffer: the serial path needs none */
  PZ_CUDA(cudaMemcpyAsync(ws.k4_start.p, c_start.data(), nb * 4u, cudaMemcpyHostToDevice, st));
  PZ_CUDA(cudaMemcpyAsync(ws.k4_len.p, c_len.data(), nb * 4u, cudaMemcpyHostToDevice, st));
  PZ_CUDA(cudaMemcpyAsync(ws.k4_off.p, c_off.data(), nb * 8u, cudaMemcpyHostToDevice, st));
  if (!symbols_ready)
    PZ_CUDA(pz_launch_blk_jobs(d_in_blob, d_in_off_i, (const uint32_t *)ws.k4_start.p, (const uint64_t *)ws.k4_off.p, (const uint32_t *)ws.k4_len.p, 0,
                               (uint16_t *)ws.k4_sym.p, (uint32_t)nb, (pz_result *)ws.k4_res.p, st, d_word + 2));
  if (trace) { cudaStreamSynchronize(st); fprintf(stderr, "[pz-k4] %8.3f ms: symbols written\n", now_ms()); }
  /* groups for the two-level walk over the tails */
  /* groups = sqrt(gscale x blocks).  A step of the one-CTA walk over the groups (pz_blk_windows_kernel) costs about twice
   * a step of the parallel walks inside the groups (pz_blk_tails_kernel), hence fewer, longer groups than sqrt(2 x blocks),
   * which would balance equal steps: 2.0 / 1.0 / 0.5 / 0.25 give 35.1 / 34.8 / 34.6 / 34.7 ms on the 1 GiB stream */
  static const double gscale = getenv("PZ_K4_GSCALE") ? atof(getenv("PZ_K4_GSCALE")) : 0.5;
  uint32_t ngrp = 1;
  while ((double)ngrp * ngrp < gscale * (double)nb) ngrp++;
  ngrp = (uint32_t)std::min<size_t>(std::min<uint32_t>(ngrp, 1024u), nb);
  std::vector<uint32_t> g_first(ngrp + 1), b_grp(nb);
  for (uint32_t g = 0; g <= ngrp; g++) g_first[g] = (uint32_t)((uint64_t)nb * g / ngrp);
  for (uint32_t g = 0; g < ngrp; g++) for (uint32_t k = g_first[g]; k < g_first[g + 1]; k++) b_grp[k] = g;
  if ((rc = ws.k4_grp.reserve(nb * 4u)) != PZ_E_OK || (rc = ws.k4_gfirst.reserve((ngrp + 1) * 4u)) != PZ_E_OK || (rc = ws.k4_gw.reserve((size_t)ngrp * 32768u)) != PZ_E_OK) return rc;
  PZ_CUDA(cudaMemcpyAsync(ws.k4_grp.p, b_grp.data(), nb * 4u, cudaMemcpyHostToDevice, st));
  PZ_CUDA(cudaMemcpyAsync(ws.k4_gfirst.p, g_first.data(), (ngrp + 1) * 4u, cudaMemcpyHostToDevice, st));
  PZ_CUDA(cudaMemsetAsync(d_word + 1, 0, 4, st));
  PZ_CUDA(pz_launch_blk_resolve((uint16_t *)ws.k4_sym.p, d_out_i, (const uint64_t *)ws.k4_off.p, (const uint32_t *)ws.k4_len.p, (const uint32_t *)ws.k4_grp.p,
                                (const uint32_t *)ws.k4_gfirst.p, ngrp, (uint32_t)nb, total, (uint8_t *)ws.k4_gw.p, d_word + 1, st));
  std::vector<pz_result> done(symbols_ready ? 0 : nb);
  uint32_t err = 0;
  if (!symbols_ready) PZ_CUDA(cudaMemcpyAsync(done.data(), ws.k4_res.p, nb * sizeof(pz_result), cudaMemcpyDeviceToHost, st));
  PZ_CUDA(cudaMemcpyAsync(&err, d_word + 1, 4, cudaMemcpyDeviceToHost, st));
  PZ_CUDA(cudaStreamSynchronize(st));
  if (trace) fprintf(stderr, "[pz-k4] %8.3f ms: decoded and resolved (err %u)\n", now_ms(), err);
  if (err) return 0;
  for (size_t k = 0; k < done.size(); k++)
    if (done[k].status != PZ_OK || done[k].out_len != c_len[k]) return 0;
  memset(out, 0, sizeof *out);
  out->status = PZ_OK;
  out->out_len = total;
  out->adler_stored = ((uint32_t)trailer[0] << 24) | ((uint32_t)trailer[1] << 16) | ((uint32_t)trailer[2] << 8) | trailer[3];
  if (framing == 1u) { /* CRC-32 and ISIZE, least significant byte first; K3 compares both */
    out->adler_stored = ((uint32_t)trailer[3] << 24) | ((uint32_t)trailer[2] << 16) | ((uint32_t)trailer[1] << 8) | trailer[0];
    out->payload[0] = (int64_t)(((uint32_t)trailer[7] << 24) | ((uint32_t)trailer[6] << 16) | ((uint32_t)trailer[5] << 8) | trailer[4]);
  }
  out->err_bitpos = tb * 8u + 8u * trailer_bytes;
  /* bytes the reference has published as 32 KiB chunks when it reaches the trailer: one chunk per
   * moveWindow call that finds 64 KiB in the window (Monad.hs:338-347).  Every block job has checked that no
   * more than 32 KiB lie between two calls (PzCtx::mark), so the window never holds 96 KiB after a call, every
   * call that finds 64 KiB leaves less than 64 KiB, and the last call leaves [32 KiB, 64 KiB): the closed form
   * below is exact for every stream K4 accepts */
  out->payload[1] = total >= 65536u ? (int64_t)(32768u * ((total - 65536u) / 32768u + 1u)) : 0;
  return 1;
}

/* K2, then K4 for every huge stream K2 left pending, then K1 for whatever is still pending. */
int run_inflate(const uint8_t *d_in, const uint64_t *d_in_off, uint8_t *d_out, const uint64_t *d_out_off, uint32_t first, uint32_t count,
                pz_result *d_res, cudaStream_t st, const uint64_t *h_in_off, const uint64_t *h_out_off, uint32_t flags,
                uint2 *d_parts = nullptr, const uint64_t *d_seg_off = nullptr) {
  const bool count_only = d_out == nullptr;
  std::vector<uint32_t> huge;
  if (!count_only && !(flags & PZ_F_NO_HUGE) && h_in_off && h_out_off)
    for (uint32_t i = first; i < first + count; i++)
      if (h_in_off[i + 1] - h_in_off[i] >= g_huge_bytes) huge.push_back(i);
  if (huge.empty()) {
    PZ_CUDA(pz_launch_inflate(d_in, d_in_off, d_out, d_out_off, first, count, d_res, st, nullptr, nullptr, PZ_PHASE_ALL, d_parts, d_seg_off, framing_of(flags)));
    return PZ_E_OK;
  }
  const uint32_t framing = framing_of(flags);
  if (framing != 0u) { /* K2 reads zlib framing only, and it is K2 that marks the streams PENDING for the K1 launch below */
    std::vector<pz_result> pend(count);
    for (pz_result &r : pend) { memset(&r, 0, sizeof r); r.status = PZ_ST_PENDING_HOST; }
    PZ_CUDA(cudaMemcpyAsync(d_res + first, pend.data(), count * sizeof(pz_result), cudaMemcpyHostToDevice, st));
    PZ_CUDA(cudaStreamSynchronize(st));
  }
  PZ_CUDA(pz_launch_inflate(d_in, d_in_off, d_out, d_out_off, first, count, d_res, st, nullptr, nullptr, PZ_PHASE_K2, d_parts, d_seg_off, framing));
  /* K2's verdicts for the span of the huge streams, in one copy: a stored-block stream is done */
  std::vector<pz_result> after_k2(huge.back() - huge.front() + 1u);
  PZ_CUDA(cudaMemcpyAsync(after_k2.data(), d_res + huge.front(), after_k2.size() * sizeof(pz_result), cudaMemcpyDeviceToHost, st));
  PZ_CUDA(cudaStreamSynchronize(st));
  for (uint32_t i : huge) {
    if (after_k2[i - huge.front()].status != PZ_ST_PENDING_HOST) continue; /* K2 has copied it */
    pz_result r;
    const int k = huge_stream(d_in, d_in_off + i, h_in_off[i], h_in_off[i + 1] - h_in_off[i], d_out + h_out_off[i], h_out_off[i + 1] - h_out_off[i], &r, st, framing);
    if (k < 0) return k;
    (k == 1 ? g_huge_done : g_huge_declined)++;
    if (k == 1) {
      PZ_CUDA(cudaMemcpyAsync(&d_res[i], &r, sizeof r, cudaMemcpyHostToDevice, st));
      PZ_CUDA(cudaStreamSynchronize(st)); /* r lives on this stack frame */
    }
  }
  PZ_CUDA(pz_launch_inflate(d_in, d_in_off, d_out, d_out_off, first, count, d_res, st, nullptr, nullptr, PZ_PHASE_K1, nullptr, nullptr, framing | 0x100u));
  return PZ_E_OK;
}

}  // namespace

struct pz_batch {
  size_t n = 0;
  uint32_t flags = 0;
  uint64_t *d_in_off = nullptr, *d_out_off = nullptr, *d_seg_off = nullptr;
  pz_result *d_res = nullptr;
  uint2 *d_parts = nullptr;
  uint64_t total_segs = 0;
  int launches = 0;
  std::vector<uint64_t> h_in_off, h_out_off; /* host copies: the K4 driver works from them */
};

extern "C" {

int pz_abi_version(void) { return PZ_ABI_VERSION; }

uint64_t pz_get_counter(int which) {
  return which == PZ_CTR_HUGE_DONE ? g_huge_done.load() : which == PZ_CTR_HUGE_DECLINED ? g_huge_declined.load() : 0;
}

int pz_set_option(int key, uint64_t value) {
  if (key == PZ_OPT_HUGE_BYTES && value > 0) { g_huge_bytes = value; return PZ_E_OK; }
  if (key == PZ_OPT_STREAM_RESUME) return PZ_E_OK; /* accepted, ignored: contexts no longer keep what a restart from the first byte would need */
  return PZ_E_ARG;
}

const char *pz_last_error(void) { return g_last_error.c_str(); }

int pz_init(const pz_config *cfg) {
  std::call_once(g_once, do_init, cfg);
  return g_init_rc;
}

int pz_device_count(void) { return ensure_init() == PZ_E_OK ? (int)g_devices.size() : 0; }

void pz_shutdown(void) {
  g_ws.~Workspace();
  new (&g_ws) Workspace();
}

/* ---- resident batches ---------------------------------------------------------------- */
pz_batch *pz_batch_create(const uint64_t *in_off, const uint64_t *out_off, size_t n, uint32_t flags) {
  if (ensure_init() != PZ_E_OK) return nullptr;
  if (!in_off || n > 0xfffffff0ull || !offsets_ok(in_off, n) || !in_sizes_ok(in_off, n)) { g_last_error = "pz_batch_create: bad arguments (offsets must ascend; one stream is at most 512 MiB - 16)"; return nullptr; }
  const bool count_only = (flags & PZ_F_COUNT_ONLY) != 0;
  if (!count_only && (!out_off || !offsets_ok(out_off, n))) { g_last_error = "pz_batch_create: bad output offsets"; return nullptr; }
  pz_batch *b = new (std::nothrow) pz_batch();
  if (!b) return nullptr;
  b->n = n; b->flags = flags;
  b->h_in_off.assign(in_off, in_off + n + 1);
  if (!count_only) b->h_out_off.assign(out_off, out_off + n + 1);
  const size_t ob = (n + 1) * sizeof(uint64_t);
  cudaError_t e = cudaMalloc(&b->d_in_off, ob);
  if (e == cudaSuccess) e = cudaMalloc(&b->d_res, std::max<size_t>(n, 1) * sizeof(pz_result));
  if (e == cudaSuccess) e = cudaMemcpy(b->d_in_off, in_off, ob, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && !count_only) {
    e = cudaMalloc(&b->d_out_off, ob);
    if (e == cudaSuccess) e = cudaMemcpy(b->d_out_off, out_off, ob, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !(flags & PZ_F_NO_ADLER)) {
      std::vector<uint64_t> seg;
      b->total_segs = build_seg_off(out_off, n, seg);
      e = cudaMalloc(&b->d_seg_off, ob);
      if (e == cudaSuccess) e = cudaMemcpy(b->d_seg_off, seg.data(), ob, cudaMemcpyHostToDevice);
      if (e == cudaSuccess) e = cudaMalloc(&b->d_parts, std::max<uint64_t>(b->total_segs, 1) * sizeof(uint2));
    }
  }
  if (e != cudaSuccess) { fail_cuda(e, "pz_batch_create"); pz_batch_destroy(b); return nullptr; }
  const int k5 = pz_small_launches((uint32_t)n, framing_of(flags)); /* K5 (and K6 when asked for; pz_fixed.cuh), each behind its list kernel, run on big batches; the sizing pass marks the streams first */
  b->launches = (count_only ? 1 + (k5 ? k5 + 1 : 0) : 3 + k5) + ((count_only || (flags & PZ_F_NO_ADLER)) ? 0 : (b->total_segs ? 2 : 1)); /* K2 probe + K2 copy (+ K5) + K1, then K3a + K3b */
  return b;
}

int pz_batch_run(pz_batch *b, const uint8_t *d_in, uint8_t *d_out, void *stream) {
  if (!b) return PZ_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const bool count_only = (b->flags & PZ_F_COUNT_ONLY) != 0;
  if (!count_only && !d_out) return PZ_E_ARG;
  {
    const int rc = run_inflate(d_in, b->d_in_off, count_only ? nullptr : d_out, b->d_out_off, 0, (uint32_t)b->n, b->d_res, st, b->h_in_off.data(),
                               count_only ? nullptr : b->h_out_off.data(), b->flags,
                               (count_only || (b->flags & PZ_F_NO_ADLER)) ? nullptr : b->d_parts, b->d_seg_off);
    if (rc != PZ_E_OK) return rc;
  }
  if (!count_only && !(b->flags & PZ_F_NO_ADLER))
    PZ_CUDA(pz_launch_adler(d_out, b->d_out_off, b->d_seg_off, (uint32_t)b->n, 0, (uint32_t)b->n, 0, b->total_segs, b->d_res, b->d_parts, st, framing_of(b->flags)));
  return PZ_E_OK;
}

int pz_batch_results(pz_batch *b, pz_result *res, void *stream) {
  if (!b || !res) return PZ_E_ARG;
  PZ_CUDA(cudaMemcpyAsync(res, b->d_res, b->n * sizeof(pz_result), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  PZ_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return PZ_E_OK;
}

int pz_batch_launches(const pz_batch *b) { return b ? b->launches : 0; }

void pz_batch_destroy(pz_batch *b) {
  if (!b) return;
  cudaFree(b->d_in_off); cudaFree(b->d_out_off); cudaFree(b->d_seg_off); cudaFree(b->d_res); cudaFree(b->d_parts);
  delete b;
}

/* ---- pinned host memory for callers that want truly asynchronous staging -------------- */
void *pz_pinned_alloc(size_t bytes) {
  if (ensure_init() != PZ_E_OK) return nullptr;
  void *p = nullptr;
  cudaError_t e = cudaHostAlloc(&p, std::max<size_t>(bytes, 1), cudaHostAllocDefault);
  if (e != cudaSuccess) { fail_cuda(e, "cudaHostAlloc"); return nullptr; }
  return p;
}
void pz_pinned_free(void *p) { if (p) cudaFreeHost(p); }

/* ---- one-shot batch over contiguous blobs --------------------------------------------- */
}  // extern "C"

namespace {
/* The batch on ONE device: the calling thread's current one. */
int contig_one_device(const uint8_t *in_blob, const uint64_t *in_off, uint8_t *out_blob, const uint64_t *out_off,
                      size_t n, pz_result *res, void *stream, uint32_t flags) {
  int rc = PZ_E_OK;
  if (n == 0) return PZ_E_OK;
  const bool count_only = (flags & PZ_F_COUNT_ONLY) != 0;
  if (!in_blob || !in_off || !res || n > 0xfffffff0ull || !offsets_ok(in_off, n) || !in_sizes_ok(in_off, n)) return PZ_E_ARG;
  if (!count_only && (!out_blob || !out_off || !offsets_ok(out_off, n))) return PZ_E_ARG;
  const bool adler = !count_only && !(flags & PZ_F_NO_ADLER);

  const bool in_dev = is_device_ptr(in_blob);
  const bool out_dev = count_only ? true : is_device_ptr(out_blob);
  if (in_dev != out_dev && !count_only) { g_last_error = "pz_inflate_batch_contig: in_blob and out_blob must both be host or both be device"; return PZ_E_ARG; }

  Workspace &ws = g_ws;
  const size_t ob = (n + 1) * sizeof(uint64_t);
  std::vector<uint64_t> seg;
  uint64_t total_segs = 0;
  if ((rc = ws.d_in_off.reserve(ob)) != PZ_E_OK) return rc;
  if ((rc = ws.d_res.reserve(n * sizeof(pz_result))) != PZ_E_OK) return rc;
  if (!count_only && (rc = ws.d_out_off.reserve(ob)) != PZ_E_OK) return rc;
  if (adler) {
    total_segs = build_seg_off(out_off, n, seg);
    if ((rc = ws.d_seg_off.reserve(ob)) != PZ_E_OK) return rc;
    if ((rc = ws.d_parts.reserve(std::max<uint64_t>(total_segs, 1) * sizeof(uint2))) != PZ_E_OK) return rc;
  }
  uint64_t *d_in_off = (uint64_t *)ws.d_in_off.p, *d_out_off = (uint64_t *)ws.d_out_off.p, *d_seg_off = (uint64_t *)ws.d_seg_off.p;
  pz_result *d_res = (pz_result *)ws.d_res.p;
  uint2 *d_parts = (uint2 *)ws.d_parts.p;

  if (in_dev) {
    /* device-resident blobs: descriptor upload + launches on the caller's stream */
    cudaStream_t st = (cudaStream_t)stream;
    PZ_CUDA(cudaMemcpyAsync(d_in_off, in_off, ob, cudaMemcpyHostToDevice, st));
    if (!count_only) PZ_CUDA(cudaMemcpyAsync(d_out_off, out_off, ob, cudaMemcpyHostToDevice, st));
    if (adler) PZ_CUDA(cudaMemcpyAsync(d_seg_off, seg.data(), ob, cudaMemcpyHostToDevice, st));
    if ((rc = run_inflate(in_blob, d_in_off, count_only ? nullptr : out_blob, d_out_off, 0, (uint32_t)n, d_res, st, in_off, out_off, flags,
                          adler ? d_parts : nullptr, d_seg_off)) != PZ_E_OK) return rc;
    if (adler) PZ_CUDA(pz_launch_adler(out_blob, d_out_off, d_seg_off, (uint32_t)n, 0, (uint32_t)n, 0, total_segs, d_res, d_parts, st, framing_of(flags)));
    PZ_CUDA(cudaMemcpyAsync(res, d_res, n * sizeof(pz_result), cudaMemcpyDeviceToHost, st));
    PZ_CUDA(cudaStreamSynchronize(st));
    return PZ_E_OK;
  }

  /* Host blobs.  The decode of a stream is one serial chain, so the whole batch goes to the device
   * in ONE launch (slices would each take as long as the batch) and the PCIe traffic is overlapped
   * with it instead:
   *   input   a pinned, mapped in_blob is read by the kernel in place (its input rings are filled
   *           with 16-byte asynchronous copies, issued far ahead of the decoder, so PCIe latency is
   *           hidden); anything else is copied to the device first.
   *   output  the kernel announces the finished part of every stream (PzJob::prog, 32 KiB steps) in
   *           mapped host memory.  When all streams have the same capacity the output is a matrix
   *           of n rows; this thread polls the progress words and sends every finished block of
   *           columns home with one 2-D copy while the kernel is still decoding the next one.
   *           Otherwise the output is copied after the kernel, overlapped only with the checksum. */
  if ((rc = ws.ensure_streams()) != PZ_E_OK) return rc;
  const uint64_t in_base = in_off[0] & ~(uint64_t)15, in_end = in_off[n];
  const uint64_t out_base = count_only ? 0 : (out_off[0] & ~(uint64_t)15), out_end = count_only ? 0 : out_off[n];
  cudaStream_t s0 = ws.streams[0], s1 = ws.streams[1], s2 = ws.streams[2];

  /* column mode: equal capacities (rows of a matrix), big enough to be worth the polling */
  const uint64_t pitch = count_only ? 0 : out_off[1] - out_off[0];
  bool columns = !count_only && n >= 64 && pitch >= 4 * kColumnBytes && pitch < (1ull << 31) && !(flags & PZ_F_NO_DRAIN);
  for (size_t i = 1; columns && i < n; i++) columns = out_off[i + 1] - out_off[i] == pitch;
  /* progressive input: the batch travels in kGroups pieces and the kernel starts the streams of a
   * piece as soon as it has landed.  K2 cannot take part (it would read all inputs at once), so
   * batches with stored-block streams -- recognisable from their first block header -- wait for
   * the whole input instead. */
  const bool in_place = (flags & PZ_F_INPUT_IN_PLACE) != 0;
  /* (a sizing pass has no output to drain but is the same serial chains: it takes the one-launch path too -- cut into slices
   * it took 38 ms for config 2, eight kernels of one chain length each, where one launch fed in pieces takes 7) */
  bool progressive = (columns || count_only) && !in_place && n >= 8 * kGroups && n <= 2 * (size_t)pz_inflate_slots() && framing_of(flags) == 0u; /* (the peek below reads a zlib header) */
  for (size_t i = 0; progressive && i < n; i++) {
    const uint64_t len = in_off[i + 1] - in_off[i];
    const uint8_t *p = in_blob + in_off[i];
    progressive = len >= 3 && len < g_huge_bytes && ((p[(p[1] & 0x20u) && len >= 7 ? 6 : 2] >> 1) & 3u) != 0u;
  }
  /* Everything else -- many more streams than the device has slots, stored-block streams, ragged
   * capacities -- is cut into slices of consecutive streams: the kernels of the slices follow each
   * other on one CUDA stream while the copies of their neighbours run on two others. */
  columns = columns && progressive;
  const uint64_t total_bytes = (in_end - in_off[0]) + (count_only ? 0 : out_end - out_off[0]);
  const int n_slices = (!progressive && n >= 64 && total_bytes >= (64ull << 20)) ? kSlices : 1;

  const uint8_t *d_in = nullptr;
  if (in_place) d_in = mapped_device_ptr(in_blob + in_base, in_end - in_base);
  const bool staged = d_in == nullptr;
  if (!staged) {
    d_in -= in_base; /* the kernel indexes with the caller's offsets */
  } else {
    if ((rc = ws.d_in.reserve(in_end - in_base + 64)) != PZ_E_OK) return rc;
    d_in = (const uint8_t *)ws.d_in.p - in_base; /* keeps the host blob's offsets modulo 16 */
  }
  uint8_t *d_out = nullptr;
  if (!count_only) {
    if ((rc = ws.d_out.reserve(out_end - out_base + 64)) != PZ_E_OK) return rc;
    d_out = (uint8_t *)ws.d_out.p - out_base;
  }
  if ((rc = ws.h_res.reserve(n * sizeof(pz_result))) != PZ_E_OK) return rc;
  pz_result *h_res = (pz_result *)ws.h_res.p;
  volatile uint32_t *h_prog = nullptr;
  uint32_t *d_prog = nullptr, *d_ready = nullptr, *h_ready = nullptr;
  if (columns || progressive) {
    if ((rc = ws.h_prog.reserve((n + kGroups) * sizeof(uint32_t))) != PZ_E_OK) return rc;
    h_ready = (uint32_t *)ws.h_prog.p + n; /* the values the ready word takes, one per piece */
  }
  if (columns) {
    memset(ws.h_prog.p, 0, n * sizeof(uint32_t));
    h_prog = (volatile uint32_t *)ws.h_prog.p;
    PZ_CUDA(cudaHostGetDevicePointer((void **)&d_prog, ws.h_prog.p, 0));
  }
  if (progressive) {
    if ((rc = ws.d_ready.reserve(sizeof(uint32_t))) != PZ_E_OK) return rc;
    d_ready = (uint32_t *)ws.d_ready.p;
  }
  const int groups = progressive ? kGroups : 1;
  auto group_lo = [&](int g) { return (size_t)((uint64_t)n * g / groups); };

  /* PZ_TRACE=1: a timeline of this call on stderr (host clock for the copies, device clock for the kernel) */
  static const bool trace = getenv("PZ_TRACE") != nullptr;
  const auto t_start = std::chrono::steady_clock::now();
  auto now_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count(); };
  cudaEvent_t k1_start, k1_done, ready_zero;
  PZ_CUDA(cudaEventCreateWithFlags(&k1_start, trace ? cudaEventDefault : cudaEventDisableTiming));
  PZ_CUDA(cudaEventCreateWithFlags(&k1_done, trace ? cudaEventDefault : cudaEventDisableTiming));
  PZ_CUDA(cudaEventCreateWithFlags(&ready_zero, cudaEventDisableTiming));
  auto run = [&]() -> int {
    PZ_CUDA(cudaMemcpyAsync(d_in_off, in_off, ob, cudaMemcpyHostToDevice, s0));
    if (!count_only) PZ_CUDA(cudaMemcpyAsync(d_out_off, out_off, ob, cudaMemcpyHostToDevice, s0));
    if (adler) PZ_CUDA(cudaMemcpyAsync(d_seg_off, seg.data(), ob, cudaMemcpyHostToDevice, s0));
    if (progressive) {
      /* the copies go on their own stream: the kernel (on s0) waits for them, never the reverse */
      PZ_CUDA(cudaMemsetAsync(d_ready, 0, sizeof(uint32_t), s0));
      PZ_CUDA(cudaEventRecord(ready_zero, s0));
      PZ_CUDA(cudaStreamWaitEvent(s2, ready_zero, 0));
      for (int g = 0; g < groups; g++) {
        const uint64_t i0 = in_off[group_lo(g)], i1 = in_off[group_lo(g + 1)];
        if (i1 > i0) PZ_CUDA(cudaMemcpyAsync((void *)(d_in + i0), in_blob + i0, i1 - i0, cudaMemcpyHostToDevice, s2));
        h_ready[g] = (uint32_t)group_lo(g + 1);
        PZ_CUDA(cudaMemcpyAsync(d_ready, h_ready + g, sizeof(uint32_t), cudaMemcpyHostToDevice, s2));
      }
    } else {
      /* sliced: H2D on s2, kernels on s0, D2H on s1 */
      PZ_CUDA(cudaEventRecord(k1_start, s0));
      PZ_CUDA(cudaEventRecord(ready_zero, s0)); /* the offset tables are queued: s2 may not overtake a previous call's kernels */
      PZ_CUDA(cudaStreamWaitEvent(s2, ready_zero, 0));
      const uint64_t slice_bytes = total_bytes / n_slices + 1;
      std::vector<cudaEvent_t> evs;
      size_t first = 0;
      for (int k = 0; first < n; k++) {
        size_t last = first;
        uint64_t acc = 0;
        while (last < n && (acc < slice_bytes || last == first)) {
          acc += (in_off[last + 1] - in_off[last]) + (count_only ? 0 : out_off[last + 1] - out_off[last]);
          last++;
        }
        if (k == n_slices - 1) last = n;
        cudaEvent_t ev_in, ev_k;
        PZ_CUDA(cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming));
        evs.push_back(ev_in);
        PZ_CUDA(cudaEventCreateWithFlags(&ev_k, cudaEventDisableTiming));
        evs.push_back(ev_k);
        const uint64_t i0 = in_off[first], i1 = in_off[last];
        if (staged && i1 > i0) PZ_CUDA(cudaMemcpyAsync((void *)(d_in + i0), in_blob + i0, i1 - i0, cudaMemcpyHostToDevice, s2));
        PZ_CUDA(cudaEventRecord(ev_in, s2));
        PZ_CUDA(cudaStreamWaitEvent(s0, ev_in, 0));
        {
          const int r2 = run_inflate(d_in, d_in_off, d_out, d_out_off, (uint32_t)first, (uint32_t)(last - first), d_res, s0, in_off, out_off, flags,
                                     adler ? d_parts : nullptr, d_seg_off);
          if (r2 != PZ_E_OK) return r2;
        }
        if (adler)
          PZ_CUDA(pz_launch_adler(d_out, d_out_off, d_seg_off, (uint32_t)n, (uint32_t)first, (uint32_t)(last - first), seg[first],
                                  seg[last] - seg[first], d_res, d_parts, s0, framing_of(flags)));
        PZ_CUDA(cudaEventRecord(ev_k, s0));
        if (!count_only) {
          const uint64_t o0 = out_off[first], o1 = out_off[last];
          PZ_CUDA(cudaStreamWaitEvent(s1, ev_k, 0));
          if (o1 > o0) PZ_CUDA(cudaMemcpyAsync(out_blob + o0, d_out + o0, o1 - o0, cudaMemcpyDeviceToHost, s1));
        }
        first = last;
      }
      PZ_CUDA(cudaEventRecord(k1_done, s0));
      PZ_CUDA(cudaMemcpyAsync(h_res, d_res, n * sizeof(pz_result), cudaMemcpyDeviceToHost, s0));
      for (cudaStream_t st : {s0, s1, s2}) PZ_CUDA(cudaStreamSynchronize(st));
      for (cudaEvent_t e : evs) cudaEventDestroy(e);
      return PZ_E_OK;
    }
    PZ_CUDA(cudaEventRecord(k1_start, s0));
    PZ_CUDA(pz_launch_inflate(d_in, d_in_off, d_out, d_out_off, 0, (uint32_t)n, d_res, s0, d_prog, d_ready, PZ_PHASE_ALL, nullptr, nullptr, framing_of(flags)));
    PZ_CUDA(cudaEventRecord(k1_done, s0));
    if (adler) PZ_CUDA(pz_launch_adler(d_out, d_out_off, d_seg_off, (uint32_t)n, 0, (uint32_t)n, 0, total_segs, d_res, d_parts, s0, framing_of(flags)));
    PZ_CUDA(cudaMemcpyAsync(h_res, d_res, n * sizeof(pz_result), cudaMemcpyDeviceToHost, s0));
    if (count_only) return PZ_E_OK;
    /* drain: columns [0, sent[g]) of the rows of piece g are on their way home */
    uint64_t sent[kGroups] = {};
    bool finished = false;
    for (;;) {
      bool all_sent = true, progress = false;
      uint32_t lo[kGroups];
      for (int g = 0; g < groups; g++) {
        lo[g] = 0xffffffffu;
        if (sent[g] >= pitch || finished) continue;
        for (size_t i = group_lo(g), e = group_lo(g + 1); i < e; i++) { const uint32_t v = h_prog[i]; lo[g] = v < lo[g] ? v : lo[g]; }
      }
      for (int g = 0; g < groups; g++) {
        if (sent[g] >= pitch) continue;
        uint64_t ready = lo[g] == 0xffffffffu ? pitch : std::min<uint64_t>(lo[g], pitch) / kColumnBytes * kColumnBytes;
        if (ready > sent[g]) {
          const size_t r0 = group_lo(g), rows = group_lo(g + 1) - r0;
          PZ_CUDA(cudaMemcpy2DAsync(out_blob + out_off[r0] + sent[g], pitch, d_out + out_off[r0] + sent[g], pitch, ready - sent[g], rows,
                                    cudaMemcpyDeviceToHost, s1));
          if (trace) fprintf(stderr, "[pz] %8.3f ms: piece %d columns [%llu, %llu) issued\n", now_ms(), g, (unsigned long long)sent[g], (unsigned long long)ready);
          sent[g] = ready;
          progress = true;
        }
        all_sent = all_sent && sent[g] >= pitch;
      }
      if (all_sent) break;
      if (progress) continue;
      const cudaError_t q = cudaEventQuery(k1_done);
      if (q == cudaSuccess) finished = true; /* whatever is left is final now */
      else if (q != cudaErrorNotReady) return fail_cuda(q, "cudaEventQuery");
      else std::this_thread::yield();
    }
    return PZ_E_OK;
  };
  rc = run();
  for (cudaStream_t st : {s0, s1, s2}) {
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess && rc == PZ_E_OK) rc = fail_cuda(e, "cudaStreamSynchronize");
  }
  if (trace && rc == PZ_E_OK) {
    float k = 0;
    cudaEventElapsedTime(&k, k1_start, k1_done);
    fprintf(stderr, "[pz] %8.3f ms: call done; K1(+K2) %.3f ms on the device; input %s, output %s\n", now_ms(), k,
            !staged ? "read in place" : progressive ? "copied in pieces while the kernel runs" : "copied slice by slice",
            columns ? "drained in column blocks" : "copied slice by slice");
  }
  cudaEventDestroy(k1_start);
  cudaEventDestroy(k1_done);
  cudaEventDestroy(ready_zero);
  if (rc == PZ_E_OK) memcpy(res, h_res, n * sizeof(pz_result));
  return rc;
}

/* Contiguous ranges of streams, one per device, balanced by compressed bytes (the same rule as
 * pure_zlib_b200/shard.py:shard_ranges, which the multi-process bench uses). */
std::vector<size_t> shard_cuts(const uint64_t *in_off, size_t n, size_t parts) {
  std::vector<size_t> cuts(parts + 1, n);
  cuts[0] = 0;
  const uint64_t base = in_off[0], total = in_off[n] - base;
  for (size_t r = 1; r < parts; r++) {
    const uint64_t target = base + (uint64_t)((long double)total * r / parts);
    size_t k = (size_t)(std::lower_bound(in_off, in_off + n + 1, target) - in_off);
    cuts[r] = std::min(std::max(k, cuts[r - 1]), n);
  }
  return cuts;
}
}  // namespace

extern "C" {

int pz_inflate_batch_contig(const uint8_t *in_blob, const uint64_t *in_off, uint8_t *out_blob, const uint64_t *out_off,
                            size_t n, pz_result *res, void *stream, uint32_t flags) {
  int rc = ensure_init();
  if (rc != PZ_E_OK) return rc;
  if (n == 0) return PZ_E_OK;
  /* host blobs on a multi-device configuration: every device takes a contiguous range of the streams, on its own worker
   * thread with its own staging, CUDA streams and kernels; no data-path collective (the streams are independent) */
  if (g_workers.size() > 1 && n >= 2 * g_workers.size() && in_blob && in_off && res && offsets_ok(in_off, n) && !is_device_ptr(in_blob)) {
    const bool count_only = (flags & PZ_F_COUNT_ONLY) != 0;
    if (!count_only && (!out_blob || !out_off)) return PZ_E_ARG;
    std::lock_guard<std::mutex> g(g_multi_mu);
    const size_t parts = g_workers.size();
    const std::vector<size_t> cuts = shard_cuts(in_off, n, parts);
    for (size_t k = 0; k < parts; k++) {
      const size_t first = cuts[k], count = cuts[k + 1] - cuts[k];
      g_workers[k]->submit([=]() -> int {
        return count ? contig_one_device(in_blob, in_off + first, out_blob, count_only ? nullptr : out_off + first, count, res + first, nullptr, flags) : PZ_E_OK;
      });
    }
    for (size_t k = 0; k < parts; k++) {
      std::string err;
      const int r = g_workers[k]->wait(&err);
      if (r != PZ_E_OK && rc == PZ_E_OK) { rc = r; g_last_error = err; }
    }
    return rc;
  }
  return contig_one_device(in_blob, in_off, out_blob, out_off, n, res, stream, flags);
}

/* ---- pointer-array entry point: what the Haskell shim binds --------------------------- */
static int inflate_ptrs(const uint8_t *const *in, const size_t *in_len, uint8_t *const *out, const size_t *out_cap, size_t n,
                        pz_result *res, uint32_t flags) {
  int rc = ensure_init();
  if (rc != PZ_E_OK) return rc;
  if (n == 0) return PZ_E_OK;
  const bool count_only = (flags & PZ_F_COUNT_ONLY) != 0;
  if (!in || !in_len || !res || (!count_only && (!out || !out_cap))) return PZ_E_ARG;
  std::vector<uint64_t> in_off(n + 1), out_off(n + 1);
  uint64_t ia = 0, oa = 0;
  for (size_t i = 0; i < n; i++) {
    in_off[i] = ia; out_off[i] = oa;
    ia += in_len[i];
    if (!count_only) oa += out_cap[i];
  }
  in_off[n] = ia; out_off[n] = oa;
  /* streams are packed back to back; the kernel copes with any alignment */
  Workspace &ws = g_ws;
  if ((rc = ws.h_in.reserve(ia + 64)) != PZ_E_OK) return rc;
  if (!count_only && (rc = ws.h_out.reserve(oa + 64)) != PZ_E_OK) return rc;
  uint8_t *h_in = (uint8_t *)ws.h_in.p, *h_out = (uint8_t *)ws.h_out.p;
  std::vector<CopyJob> jobs(n);
  for (size_t i = 0; i < n; i++) jobs[i] = CopyJob{h_in + in_off[i], in[i], in_len[i]};
  parallel_copy(jobs);
  rc = pz_inflate_batch_contig(h_in, in_off.data(), h_out, out_off.data(), n, res, nullptr, flags);
  if (rc != PZ_E_OK) return rc;
  if (!count_only) {
    for (size_t i = 0; i < n; i++) jobs[i] = CopyJob{out[i], h_out + out_off[i], (size_t)std::min<uint64_t>(res[i].out_len, out_cap[i])};
    parallel_copy(jobs);
  }
  return PZ_E_OK;
}

int pz_inflate_batch(const uint8_t *const *in, const size_t *in_len, uint8_t *const *out, const size_t *out_cap, size_t n,
                     pz_result *res, uint32_t flags) {
  return inflate_ptrs(in, in_len, out, out_cap, n, res, flags & ~PZ_F_COUNT_ONLY);
}

int pz_inflate_sizes(const uint8_t *const *in, const size_t *in_len, size_t n, pz_result *res) {
  return inflate_ptrs(in, in_len, nullptr, nullptr, n, res, PZ_F_COUNT_ONLY);
}

int pz_inflate_sizes_framed(const uint8_t *const *in, const size_t *in_len, size_t n, pz_result *res, uint32_t flags) {
  return inflate_ptrs(in, in_len, nullptr, nullptr, n, res, (flags & (PZ_F_GZIP | PZ_F_RAW)) | PZ_F_COUNT_ONLY);
}

/* ---- incremental decoder (decompressIncremental, Zlib.hs:29-30) -----------------------
 * The reference's decoder is a coroutine (Monad.hs:163-197): it stops at NeedMore with its whole state
 * in a closure and goes on when the next chunk arrives.  Here the state of a stream lives on the DEVICE:
 * the compressed bytes fed so far, the decoded bytes (which are the LZ77 history) and a four-word
 * checkpoint (PzJob::ckpt: block header, symbol, bytes decoded, bytes published).  Every pump decodes
 * only what the new input adds: K1 rebuilds the tables of the block the checkpoint lies in (they live in
 * shared memory and do not survive a launch) and goes on at the symbol that could not be completed.
 * Any number of streams are pumped by ONE launch (pz_stream_pump).  Fed chunks travel through pinned
 * staging with cudaMemcpyAsync on the stream's own CUDA stream; published chunks come home into a pinned
 * buffer during the pump, so pz_stream_next never touches the device. */
}  // extern "C"

namespace {
/* Pinned blocks are expensive to create (cudaHostAlloc takes of the order of a millisecond), and incremental
 * streams come and go by the thousand: blocks are handed out in power-of-two size classes and go back to a
 * process-wide free list, never to the driver. */
struct PinnedCache {
  std::mutex mu;
  std::vector<uint8_t *> free_[48];
  static int cls(size_t n) { int c = 14; while (((size_t)1 << c) < n) c++; return c; } /* 16 KiB and up */
  uint8_t *get(size_t n, size_t *cap) {
    const int c = cls(n);
    *cap = (size_t)1 << c;
    {
      std::lock_guard<std::mutex> g(mu);
      if (!free_[c].empty()) { uint8_t *p = free_[c].back(); free_[c].pop_back(); return p; }
    }
    uint8_t *p = nullptr;
    cudaError_t e = cudaHostAlloc((void **)&p, *cap, cudaHostAllocPortable);
    if (e != cudaSuccess) { fail_cuda(e, "cudaHostAlloc"); return nullptr; }
    return p;
  }
  void put(uint8_t *p, size_t cap) {
    if (!p) return;
    std::lock_guard<std::mutex> g(mu);
    free_[cls(cap)].push_back(p);
  }
};
PinnedCache g_pinned;

struct PinnedBuf {
  uint8_t *p = nullptr;
  size_t cap = 0;
  int reserve_keep(size_t n, size_t keep_len) { /* grow, keeping the first keep_len bytes */
    if (n <= cap) return PZ_E_OK;
    size_t ncap = 0;
    uint8_t *q = g_pinned.get(std::max(n, cap * 2), &ncap);
    if (!q) return PZ_E_NOMEM;
    if (keep_len) memcpy(q, p, keep_len);
    g_pinned.put(p, cap);
    p = q; cap = ncap;
    return PZ_E_OK;
  }
  void release() { g_pinned.put(p, cap); p = nullptr; cap = 0; }
};
constexpr size_t kStageBytes = 1 << 20; /* most bytes of a feed that travel in one piece */
constexpr uint64_t kGatherMax = 1 << 20; /* pieces of decoded output up to this size go home by kernel (pz_gather_kernel) */

/* CUDA streams the incremental contexts share (round robin), created once per process: a pump then orders itself
 * behind the feeds of its streams with at most kPoolStreams event waits, however many contexts it decodes. */
constexpr int kPoolStreams = 8;
cudaStream_t g_pool[kPoolStreams] = {};
std::once_flag g_pool_once;
cudaError_t g_pool_rc = cudaSuccess;
std::atomic<unsigned> g_pool_next{0};
void pool_init() {
  for (int i = 0; i < kPoolStreams && g_pool_rc == cudaSuccess; i++) g_pool_rc = cudaStreamCreateWithFlags(&g_pool[i], cudaStreamNonBlocking);
  /* device buffers of the contexts come from the stream-ordered allocator; freed blocks stay with the process */
  cudaMemPool_t pool;
  int dev = 0;
  if (g_pool_rc == cudaSuccess) g_pool_rc = cudaGetDevice(&dev);
  if (g_pool_rc == cudaSuccess) g_pool_rc = cudaDeviceGetDefaultMemPool(&pool, dev);
  if (g_pool_rc == cudaSuccess) {
    uint64_t keep = ~0ull;
    g_pool_rc = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
}
}  // namespace

struct pz_stream {
  cudaStream_t st = nullptr;
  int pool_idx = 0;                /* st == g_pool[pool_idx] */
  uint32_t framing = 0;            /* PZ_FRAME_*: zlib, gzip, raw deflate */
  /* The device context is O(1) in the length of the stream (the reference's is its 128 KiB window, OutputWindow.hs:29-54):
   * the compressed bytes from the header of the block the decoder is in, the last 32 KiB of decoded bytes (the LZ77
   * history) plus room for what one pump adds, and a checkpoint.  Everything behind them is dropped as the stream goes. */
  /* input: d_in[0] is byte in_base of the stream; bytes [in_base, in_total) are on the device */
  uint8_t *d_in = nullptr;
  size_t d_in_cap = 0;
  uint64_t in_total = 0, in_base = 0;
  uint64_t lead = 0;               /* stream byte the checkpoint's bit positions are counted from (in_base <= lead) */
  uint32_t ck[4] = {0, 0, 0, 0};   /* where the next pump picks the stream up (PzJob::ckpt): {bit of the block header, bit of the
                                      symbol, -, -} counted from `lead`; the byte counters of the checkpoint are kept below */
  /* output: d_out[0] is decoded byte out_base; bytes [out_base, pos) are on the device, at least the last 32 KiB of them */
  uint8_t *d_out = nullptr;
  size_t d_out_cap = 0;
  uint64_t out_base = 0, pos = 0;  /* absolute counts: they may pass 4 GiB, the kernel sees them relative to base_abs */
  uint64_t base_abs = 0;           /* bytes the reference has published at the checkpoint (multiple of 32 KiB) */
  uint32_t sum = 1;                /* running checksum of the decoded bytes: Adler-32 (zlib, raw) or CRC-32 (gzip) */
  /* the feed on its way: pinned staging, reused once its copy has finished */
  PinnedBuf stage;
  cudaEvent_t staged = nullptr;
  /* decoded bytes on the host: [h_first, h_to) of the stream, h_first <= published; h_to == pos after every pump */
  PinnedBuf h_out;
  uint64_t h_first = 0, h_to = 0;
  uint64_t published = 0;    /* bytes already handed out as chunks */
  uint64_t publish_to = 0;   /* bytes the reference has published at the current state */
  bool dirty = false;        /* input arrived since the last decode */
  bool queued = false;       /* inside pz_stream_pump: already on the list of the pump */
  bool terminal = false;     /* verdict reached */
  bool final_pending = false;/* the final (possibly empty) chunk has not been delivered */
  bool done_delivered = false;
  pz_result verdict{};
  uint64_t pumps = 0, resumed = 0; /* launches that decoded this stream; those that started from a checkpoint */
  uint64_t feed_stamp = 0;         /* the pz_stream_feed_many call that last took a chunk for this stream */
  size_t peak_device = 0;          /* most device memory the context has held (PZ_SC_DEVICE_PEAK) */
};

namespace {
/* per-thread control tables of a pump: [in pairs | out pairs | resume] go up, [ckpt | res] come back */
struct PumpSpace {
  Buf h_ctl, d_ctl, d_parts, d_tab, h_tab, d_gather, h_feed, d_feed;
  cudaEvent_t fed[kPoolStreams] = {};
  PumpSpace() { h_ctl.pinned = true; h_feed.pinned = true; h_tab.pinned = true; }
  int ensure_events() {
    for (int i = 0; i < kPoolStreams; i++)
      if (!fed[i]) PZ_CUDA(cudaEventCreateWithFlags(&fed[i], cudaEventDisableTiming));
    return PZ_E_OK;
  }
  ~PumpSpace() {
    h_ctl.release(); d_ctl.release(); d_parts.release(); d_tab.release(); h_tab.release(); d_gather.release(); h_feed.release(); d_feed.release();
    for (int i = 0; i < kPoolStreams; i++) if (fed[i]) cudaEventDestroy(fed[i]);
  }
};
thread_local PumpSpace g_pump;

int grow_device(uint8_t *&d, size_t &cap, size_t want, size_t keep, cudaStream_t st) {
  if (want <= cap) return PZ_E_OK;
  size_t n = std::max<size_t>(align_up(want + 64, 1 << 16), cap * 2);
  uint8_t *q = nullptr;
  cudaError_t e = cudaMallocAsync((void **)&q, n, st);
  if (e != cudaSuccess) { fail_cuda(e, "cudaMallocAsync"); return PZ_E_NOMEM; }
  if (keep) PZ_CUDA(cudaMemcpyAsync(q, d, keep, cudaMemcpyDeviceToDevice, st));
  if (d) PZ_CUDA(cudaFreeAsync(d, st)); /* stream-ordered: after the copy, and after everything queued on st before */
  d = q; cap = n;
  return PZ_E_OK;
}

/* Replaces the buffer by one of `want` bytes holding bytes [from, from + keep) of the old one (a context gives memory back
 * when what it must keep has become much smaller than what it holds). */
int shrink_device(uint8_t *&d, size_t &cap, size_t want, size_t from, size_t keep, cudaStream_t st) {
  size_t n = align_up(want + 64, 1 << 16);
  uint8_t *q = nullptr;
  cudaError_t e = cudaMallocAsync((void **)&q, n, st);
  if (e != cudaSuccess) { fail_cuda(e, "cudaMallocAsync"); return PZ_E_NOMEM; }
  if (keep) PZ_CUDA(cudaMemcpyAsync(q, d + from, keep, cudaMemcpyDeviceToDevice, st));
  if (d) PZ_CUDA(cudaFreeAsync(d, st));
  d = q; cap = n;
  return PZ_E_OK;
}

constexpr uint64_t kHistory = 32768;          /* the farthest a match reaches back (Deflate.hs:199-237) */
constexpr uint64_t kRoomMin = 256 << 10;      /* decoded bytes a context has room for in one launch, at least ... */
constexpr uint64_t kDeadMin = 256 << 10;      /* consumed input a context may keep before it gives the memory back (an allocation, a copy and a free per
                                                 stream: with thousands of consumers in step that is tens of milliseconds of a round) */
constexpr uint64_t kRoomMax = 64 << 20;       /* ... and at most: a stream that expands further takes another launch of the same pump */

uint32_t adler_combine(uint32_t s1, uint32_t s2, uint64_t len2) { /* Adler-32 of A || B from those of A and B (a0 = 1, b0 = 0: Adler32.hs:19-20) */
  const uint32_t M = 65521u;
  const uint32_t a1 = s1 & 0xffffu, b1 = s1 >> 16, a2 = s2 & 0xffffu, b2 = s2 >> 16;
  const uint32_t a = (a1 + a2 + M - 1u) % M;
  const uint32_t b = (uint32_t)((b1 + b2 + (len2 % M) * ((a1 + M - 1u) % M)) % M);
  return (b << 16) | a;
}
uint32_t crc_gf_mul(uint32_t a, uint32_t b) {
  uint32_t p = 0;
  for (uint32_t m = 0x80000000u; m != 0u && a != 0u; m >>= 1) {
    if (a & m) { p ^= b; a &= ~m; }
    b = (b >> 1) ^ (0xedb88320u & (0u - (b & 1u)));
  }
  return p;
}
uint32_t crc_combine(uint32_t c1, uint32_t c2, uint64_t len2) { /* CRC-32 of A || B: crc(A) * x^(8|B|) + crc(B) over GF(2) mod P */
  uint32_t xp = 0x80000000u, sq = 0x40000000u;
  for (int k = 0; k < 3; k++) sq = crc_gf_mul(sq, sq);
  for (uint64_t n = len2; n != 0u; n >>= 1) {
    if (n & 1u) xp = crc_gf_mul(sq, xp);
    sq = crc_gf_mul(sq, sq);
  }
  return crc_gf_mul(xp, c1) ^ c2;
}

/* The state a verdict puts the stream in (what stream_next hands out afterwards).  r is in absolute byte counts. */
void stream_settle(pz_stream *s, const pz_result &r) {
  s->verdict = r;
  const bool need_more = r.status == PZ_ERR_DECOMPRESSION && r.detail == PZ_D_RAN_OUT;
  s->publish_to = s->base_abs;
  if (!need_more) {
    s->terminal = true;
    s->final_pending = (r.status == PZ_OK); /* finalize publishes what is left (Monad.hs:349-353) */
  }
}

/* One framing at a time: the streams of `act` all read the same framing (PzJob::framing belongs to the launch). */
int pump_framed(const std::vector<pz_stream *> &act, uint32_t framing) {
  if (act.empty()) return PZ_E_OK;
  cudaStream_t st = act[0]->st;
  PumpSpace &ps = g_pump;
  int rc;
  if ((rc = ps.ensure_events()) != PZ_E_OK) return rc;
  std::vector<pz_stream *> run = act;
  while (!run.empty()) {
    const size_t n = run.size();
    /* every stream's feeds (and buffer moves) must have landed before the launch on `st` reads them */
    {
      bool seen[kPoolStreams] = {};
      for (pz_stream *s : run) seen[s->pool_idx] = true;
      for (int k = 0; k < kPoolStreams; k++)
        if (seen[k] && g_pool[k] != st) {
          PZ_CUDA(cudaEventRecord(ps.fed[k], g_pool[k]));
          PZ_CUDA(cudaStreamWaitEvent(st, ps.fed[k], 0));
        }
    }
    const size_t up = n * 48, down = n * (16 + sizeof(pz_result));
    if ((rc = ps.h_ctl.reserve(up + down)) != PZ_E_OK) return rc;
    if ((rc = ps.d_ctl.reserve(up + down)) != PZ_E_OK) return rc;
    uint8_t *h = (uint8_t *)ps.h_ctl.p, *d = (uint8_t *)ps.d_ctl.p;
    uint64_t *in_pairs = (uint64_t *)h, *out_pairs = (uint64_t *)(h + 16 * n);
    uint32_t *resume = (uint32_t *)(h + 32 * n);
    std::vector<uint64_t> moves; /* histories moved to the front of their buffers: ONE kernel for all of them (a cudaMemcpyAsync per
                                    stream costs more host time than the copies take: 4096 consumers in step = 20 ms per round) */
    for (size_t i = 0; i < n; i++) {
      pz_stream *s = run[i];
      /* room for what this input may add before the kernel has to stop for a larger buffer */
      const uint64_t room = std::min<uint64_t>(kRoomMax, std::max<uint64_t>(kRoomMin, 4 * (s->in_total - s->lead)));
      uint64_t hist = s->pos - s->out_base; /* decoded bytes on the device */
      if (hist >= 2 * kHistory && hist + room + 64 > s->d_out_cap) { /* only the last 32 KiB can still be referenced: move them to the front */
        if (kHistory + room + 64 <= s->d_out_cap) { /* stays in this buffer: queued for the one move kernel below (disjoint: hist >= 64 KiB) */
          moves.push_back((uint64_t)(uintptr_t)(s->d_out + (hist - kHistory))); moves.push_back((uint64_t)(uintptr_t)s->d_out); moves.push_back(kHistory);
        } else { /* the buffer is about to be replaced: move now, stream-ordered before the copy into the new one */
          PZ_CUDA(cudaMemcpyAsync(s->d_out, s->d_out + (hist - kHistory), kHistory, cudaMemcpyDeviceToDevice, st));
        }
        s->out_base = s->pos - kHistory;
        hist = kHistory;
      }
      if ((rc = grow_device(s->d_out, s->d_out_cap, hist + room + 64, hist, st)) != PZ_E_OK) return rc; /* d_out is only ever touched by pumps, which end synchronised */
      if (!s->d_in && (rc = grow_device(s->d_in, s->d_in_cap, 64, 0, s->st)) != PZ_E_OK) return rc; /* a stream nothing was fed to yet */
      s->peak_device = std::max(s->peak_device, s->d_in_cap + s->d_out_cap);
      /* the kernel counts decoded bytes from base_abs (a multiple of 32 KiB): its `base` is 0, its position the window's fill */
      const uint64_t fill = s->pos - s->base_abs;
      in_pairs[2 * i] = (uint64_t)(uintptr_t)(s->d_in + (s->lead - s->in_base));
      in_pairs[2 * i + 1] = (uint64_t)(uintptr_t)(s->d_in + (s->in_total - s->in_base));
      const uint64_t out0 = (uint64_t)(uintptr_t)s->d_out + (s->base_abs - s->out_base); /* address of decoded byte base_abs (may lie before the buffer: never touched) */
      out_pairs[2 * i] = out0;
      out_pairs[2 * i + 1] = out0 + std::min<uint64_t>((s->out_base - s->base_abs) + (s->d_out_cap - 64), 0xfffdff00ull);
      resume[4 * i] = s->ck[0]; resume[4 * i + 1] = s->ck[1]; resume[4 * i + 2] = (uint32_t)fill; resume[4 * i + 3] = 0;
    }
    if (!moves.empty()) {
      if ((rc = ps.d_gather.reserve(moves.size() * 8u)) != PZ_E_OK) return rc;
      PZ_CUDA(cudaMemcpyAsync(ps.d_gather.p, moves.data(), moves.size() * 8u, cudaMemcpyHostToDevice, st));
      PZ_CUDA(pz_launch_gather((const uint64_t *)ps.d_gather.p, (uint32_t)(moves.size() / 3), st));
    }
    PZ_CUDA(cudaMemcpyAsync(d, h, up, cudaMemcpyHostToDevice, st));
    uint32_t *d_ck = (uint32_t *)(d + up);
    pz_result *d_res = (pz_result *)(d + up + 16 * n);
    PZ_CUDA(pz_launch_resume((const uint64_t *)d, (const uint64_t *)(d + 16 * n), (uint32_t)n, d_res, (const uint32_t *)(d + 32 * n), d_ck, st, framing));
    PZ_CUDA(cudaMemcpyAsync(h + up, d + up, down, cudaMemcpyDeviceToHost, st));
    PZ_CUDA(cudaStreamSynchronize(st));
    const uint32_t *ck = (const uint32_t *)(h + up);
    const pz_result *res = (const pz_result *)(h + up + 16 * n);
    /* What this launch decoded, per stream: its checksum (K3 over the new bytes only, folded into the running sum with the
     * combine identities) and its way home (ONE kernel stores the small pieces of all streams into their pinned buffers: a
     * cudaMemcpyAsync each costs more host time than the copy takes; long pieces go by copy engine). */
    std::vector<uint64_t> tab(2 * n + 1), gat; /* [0, n): where each new piece begins; [n, 2n]: its first checksum segment */
    std::vector<uint64_t> newpos(n);
    uint64_t segs = 0;
    bool any_new = false;
    if ((rc = ps.h_tab.reserve(n * sizeof(pz_result))) != PZ_E_OK) return rc;
    pz_result *piece = (pz_result *)ps.h_tab.p;
    for (size_t i = 0; i < n; i++) {
      pz_stream *s = run[i];
      newpos[i] = s->base_abs + ck[4 * i + 2];
      const uint64_t len = newpos[i] - s->pos;
      tab[i] = (uint64_t)(uintptr_t)(s->d_out + (s->pos - s->out_base));
      tab[n + i] = segs;
      memset(&piece[i], 0, sizeof(pz_result));
      piece[i].status = len ? PZ_OK : PZ_ST_PENDING_HOST;
      piece[i].out_len = len;
      if (len) { segs += (len + PZ_ADLER_SEG - 1) / PZ_ADLER_SEG; any_new = true; }
    }
    tab[2 * n] = segs;
    if (any_new) {
      if ((rc = ps.d_parts.reserve(std::max<uint64_t>(segs, 1) * sizeof(uint2))) != PZ_E_OK) return rc;
      if ((rc = ps.d_tab.reserve(tab.size() * 8u + n * sizeof(pz_result))) != PZ_E_OK) return rc;
      pz_result *d_piece = (pz_result *)((uint8_t *)ps.d_tab.p + tab.size() * 8u);
      PZ_CUDA(cudaMemcpyAsync(ps.d_tab.p, tab.data(), tab.size() * 8u, cudaMemcpyHostToDevice, st));
      PZ_CUDA(cudaMemcpyAsync(d_piece, piece, n * sizeof(pz_result), cudaMemcpyHostToDevice, st));
      PZ_CUDA(pz_launch_adler(nullptr, (const uint64_t *)ps.d_tab.p, (const uint64_t *)ps.d_tab.p + n, (uint32_t)n, 0, (uint32_t)n, 0, segs, d_piece,
                              (uint2 *)ps.d_parts.p, st, framing | 0x200u /* checksums only: nothing to compare with yet */));
      PZ_CUDA(cudaMemcpyAsync(piece, d_piece, n * sizeof(pz_result), cudaMemcpyDeviceToHost, st));
      for (size_t i = 0; i < n; i++) {
        pz_stream *s = run[i];
        const uint64_t len = newpos[i] - s->pos;
        if (!len) continue;
        /* host side: handed-out bytes in front of the buffer are dropped once they outweigh what is still owed */
        if (s->published == s->h_to) s->h_first = s->h_to;
        else if (s->published - s->h_first >= std::max<uint64_t>(128 << 10, s->h_to - s->published)) { /* (amortised: the move is at most what it frees) */
          memmove(s->h_out.p, s->h_out.p + (s->published - s->h_first), (size_t)(s->h_to - s->published));
          s->h_first = s->published;
        }
        const uint64_t keep = s->h_to - s->h_first;
        if ((rc = s->h_out.reserve_keep(std::max<size_t>((size_t)(keep + len), 64 << 10), (size_t)keep)) != PZ_E_OK) return rc; /* 64 KiB are buffered before the first chunk is due (Monad.hs:338-358) */
        if (len > kGatherMax) {
          PZ_CUDA(cudaMemcpyAsync(s->h_out.p + keep, s->d_out + (s->pos - s->out_base), len, cudaMemcpyDeviceToHost, st));
        } else {
          gat.push_back(tab[i]); gat.push_back((uint64_t)(uintptr_t)(s->h_out.p + keep)); gat.push_back(len);
        }
        s->h_to = newpos[i];
      }
      if (!gat.empty()) {
        if ((rc = ps.d_gather.reserve(gat.size() * 8u)) != PZ_E_OK) return rc;
        PZ_CUDA(cudaMemcpyAsync(ps.d_gather.p, gat.data(), gat.size() * 8u, cudaMemcpyHostToDevice, st));
        PZ_CUDA(pz_launch_gather((const uint64_t *)ps.d_gather.p, (uint32_t)(gat.size() / 3), st));
      }
      PZ_CUDA(cudaStreamSynchronize(st));
    }
    std::vector<pz_stream *> again;
    for (size_t i = 0; i < n; i++) {
      pz_stream *s = run[i];
      s->pumps++;
      if ((s->ck[0] | s->ck[1]) != 0 || s->lead != 0 || s->pos != 0) s->resumed++;
      /* the running checksum takes the new piece in */
      const uint64_t len = newpos[i] - s->pos;
      if (len) s->sum = framing == 1u ? crc_combine(s->sum, piece[i].adler_computed, len) : adler_combine(s->sum, piece[i].adler_computed, len);
      s->pos = newpos[i];
      s->base_abs += ck[4 * i + 3];
      /* the checkpoint moves on, and the input behind its block header is dead: bit positions are re-based on a byte
       * shortly before that header (never ON it: a checkpoint whose first word is 0 reads as "no checkpoint") */
      const uint64_t old_lead = s->lead;
      const uint64_t hdr_byte = s->lead + ck[4 * i] / 8u;
      const uint64_t new_lead = hdr_byte >= 17u ? (hdr_byte - 1u) & ~(uint64_t)15u : 0u;
      const uint32_t shift = new_lead > s->lead ? (uint32_t)((new_lead - s->lead) * 8u) : 0u;
      s->ck[0] = ck[4 * i] - shift;
      s->ck[1] = (ck[4 * i + 1] == 0u || ck[4 * i + 1] == PZ_CK_TRAILER_HOST) ? ck[4 * i + 1] : ck[4 * i + 1] - shift;
      if (shift) s->lead = new_lead;
      const uint64_t dead = s->lead - s->in_base, live = s->in_total - s->lead;
      if (dead >= std::max<uint64_t>(kDeadMin, live)) { /* give the dead input back (amortised: the copy is at most what was freed) */
        if ((rc = shrink_device(s->d_in, s->d_in_cap, std::max<uint64_t>(2 * live, 1 << 16), (size_t)dead, (size_t)live, s->st)) != PZ_E_OK) return rc;
        s->in_base = s->lead;
      }
      /* the verdict in absolute byte counts */
      pz_result r = res[i];
      r.out_len = s->pos;
      r.err_bitpos += old_lead * 8u;
      if (!(r.status == PZ_REF_BOTTOM && r.detail == PZ_D_BOT_DIST_TOO_FAR)) r.payload[1] = (int64_t)s->base_abs;
      r.adler_computed = s->sum;
      if (r.status == PZ_OUTPUT_FULL) { /* the buffer, not the stream: the next launch of this pump starts from the checkpoint with room again */
        if (s->pos >= (1ull << 62)) { stream_settle(s, r); continue; }
        again.push_back(s);
        continue;
      }
      if (r.status == PZ_OK && framing != 2u) { /* checkChecksum (Deflate.hs:52-63) against the running sum; gzip: then ISIZE */
        if (r.adler_stored != s->sum) { r.status = PZ_ERR_CHECKSUM; r.detail = PZ_D_ADLER_MISMATCH; }
        else if (framing == 1u && (uint32_t)r.payload[0] != (uint32_t)s->pos) { r.status = PZ_ERR_CHECKSUM; r.detail = PZ_D_LENGTH_MISMATCH; r.payload[0] = (int64_t)(uint32_t)r.payload[0]; }
      }
      stream_settle(s, r);
    }
    run.swap(again);
  }
  for (pz_stream *s : act) s->dirty = false;
  return PZ_E_OK;
}

int pump(pz_stream *const *all, size_t n_all) {
  std::vector<pz_stream *> act[3];
  for (size_t i = 0; i < n_all; i++) {
    pz_stream *s = all[i];
    if (!s) return PZ_E_ARG;
    if (s->dirty && !s->terminal && !s->queued) { s->queued = true; act[s->framing % 3u].push_back(s); } /* (a stream listed twice is pumped once) */
  }
  for (auto &v : act) for (pz_stream *s : v) s->queued = false;
  for (uint32_t f = 0; f < 3u; f++) {
    const int rc = pump_framed(act[f], f);
    if (rc != PZ_E_OK) return rc;
  }
  return PZ_E_OK;
}
}  // namespace

extern "C" {

pz_stream *pz_stream_new_framed(uint32_t flags) {
  if (ensure_init() != PZ_E_OK) return nullptr;
  std::call_once(g_pool_once, pool_init);
  if (g_pool_rc != cudaSuccess) { fail_cuda(g_pool_rc, "pz_stream_new (stream pool)"); return nullptr; }
  pz_stream *s = new (std::nothrow) pz_stream();
  if (!s) return nullptr;
  s->framing = framing_of(flags);
  s->sum = s->framing == 1u ? 0u : 1u; /* CRC-32 of nothing / initialAdlerState (Adler32.hs:19-20) */
  s->pool_idx = (int)(g_pool_next++ % kPoolStreams);
  s->st = g_pool[s->pool_idx];
  cudaError_t e = cudaEventCreateWithFlags(&s->staged, cudaEventDisableTiming);
  if (e != cudaSuccess) { fail_cuda(e, "pz_stream_new"); pz_stream_free(s); return nullptr; }
  return s;
}

pz_stream *pz_stream_new(void) { return pz_stream_new_framed(0); }

void pz_stream_free(pz_stream *s) {
  if (!s) return;
  if (s->staged) { cudaEventSynchronize(s->staged); cudaEventDestroy(s->staged); } /* the staging block may be reused at once */
  s->stage.release();
  s->h_out.release();
  if (s->d_in) cudaFreeAsync(s->d_in, s->st);
  if (s->d_out) cudaFreeAsync(s->d_out, s->st);
  delete s;
}

/* The chunk is copied into pinned staging and sent to the device asynchronously: the call returns while
 * the copy is in flight (the staging block is reused once its copy has finished). */
int pz_stream_feed(pz_stream *s, const uint8_t *data, size_t len) {
  if (!s || (len && !data)) return PZ_E_ARG;
  if (s->terminal) return PZ_E_STATE; /* the decoder never asked for this chunk */
  if ((s->in_total - s->lead) + len > PZ_MAX_STREAM_BYTES) return PZ_E_ARG; /* ONE deflate block (plus the chunk) beyond 512 MiB */
  s->dirty = true;
  if (len == 0) return PZ_E_OK; /* empty chunks are accepted and ignored (Monad.hs:193-195) */
  const size_t have = (size_t)(s->in_total - s->in_base);
  int rc = grow_device(s->d_in, s->d_in_cap, have + len + 64, have, s->st);
  if (rc != PZ_E_OK) return rc;
  for (size_t at = 0; at < len;) {
    const size_t piece = std::min(len - at, kStageBytes);
    PZ_CUDA(cudaEventSynchronize(s->staged));
    if ((rc = s->stage.reserve_keep(piece, 0)) != PZ_E_OK) return rc;
    memcpy(s->stage.p, data + at, piece);
    PZ_CUDA(cudaMemcpyAsync(s->d_in + (s->in_total - s->in_base), s->stage.p, piece, cudaMemcpyHostToDevice, s->st));
    PZ_CUDA(cudaEventRecord(s->staged, s->st));
    s->in_total += piece;
    at += piece;
  }
  return PZ_E_OK;
}

/* One chunk for each of n streams: all chunks are packed into ONE pinned staging buffer, cross the bus in one copy and
 * are dealt to the streams' input buffers by one kernel -- the cost of a round of feeds no longer grows with a
 * cudaMemcpyAsync and an event per stream.  Returns when the bytes are on the device. */
int pz_stream_feed_many(pz_stream *const *streams, const uint8_t *const *data, const size_t *len, size_t n) {
  if (n == 0) return PZ_E_OK;
  if (!streams || !data || !len) return PZ_E_ARG;
  int rc = ensure_init();
  if (rc != PZ_E_OK) return rc;
  PumpSpace &ps = g_pump;
  if ((rc = ps.ensure_events()) != PZ_E_OK) return rc;
  uint64_t total = 0;
  static std::atomic<uint64_t> g_feed_stamp{0};
  const uint64_t stamp = ++g_feed_stamp;
  for (size_t i = 0; i < n; i++) {
    pz_stream *s = streams[i];
    if (!s || (len[i] && !data[i])) return PZ_E_ARG;
    if (s->terminal) return PZ_E_STATE;
    if ((s->in_total - s->lead) + len[i] > PZ_MAX_STREAM_BYTES) return PZ_E_ARG;
    if (s->feed_stamp == stamp) return PZ_E_ARG; /* one chunk per stream and call */
    s->feed_stamp = stamp;
    total += (len[i] + 15u) & ~(uint64_t)15u;
  }
  cudaStream_t st = streams[0]->st;
  if ((rc = ps.h_feed.reserve(total + 16 + n * 24u)) != PZ_E_OK || (rc = ps.d_feed.reserve(total + 16 + n * 24u)) != PZ_E_OK) return rc;
  /* layout of both buffers: n (source, destination, length) triples, then the chunks, 16-byte aligned */
  uint64_t *tri = (uint64_t *)ps.h_feed.p;
  const uint64_t base = (n * 24u + 15u) & ~(uint64_t)15u;
  uint8_t *hp = (uint8_t *)ps.h_feed.p, *dp = (uint8_t *)ps.d_feed.p;
  uint64_t at = base;
  bool seen[kPoolStreams] = {};
  size_t m = 0;
  std::vector<CopyJob> jobs;
  for (size_t i = 0; i < n; i++) {
    pz_stream *s = streams[i];
    s->dirty = true;
    if (len[i] == 0) continue; /* empty chunks are accepted and ignored (Monad.hs:193-195) */
    const size_t have = (size_t)(s->in_total - s->in_base);
    if ((rc = grow_device(s->d_in, s->d_in_cap, have + len[i] + 64, have, s->st)) != PZ_E_OK) return rc;
    seen[s->pool_idx] = true;
    jobs.push_back(CopyJob{hp + at, data[i], len[i]});
    tri[3 * m] = (uint64_t)(uintptr_t)(dp + at); tri[3 * m + 1] = (uint64_t)(uintptr_t)(s->d_in + have); tri[3 * m + 2] = len[i];
    s->in_total += len[i];
    at += (len[i] + 15u) & ~(uint64_t)15u;
    m++;
  }
  if (m == 0) return PZ_E_OK;
  parallel_copy(jobs);
  /* the streams' buffers may just have moved (and earlier single feeds may still be in flight) on their own CUDA streams */
  for (int k = 0; k < kPoolStreams; k++)
    if (seen[k] && g_pool[k] != st) {
      PZ_CUDA(cudaEventRecord(ps.fed[k], g_pool[k]));
      PZ_CUDA(cudaStreamWaitEvent(st, ps.fed[k], 0));
    }
  PZ_CUDA(cudaMemcpyAsync(dp, hp, at, cudaMemcpyHostToDevice, st));
  PZ_CUDA(pz_launch_gather((const uint64_t *)dp, (uint32_t)m, st));
  PZ_CUDA(cudaStreamSynchronize(st)); /* the staging buffers are this thread's: free for the next call; later work on any stream is ordered behind */
  return PZ_E_OK;
}

int pz_stream_pump(pz_stream *const *streams, size_t n) {
  if (n && !streams) return PZ_E_ARG;
  int rc = ensure_init();
  if (rc != PZ_E_OK) return rc;
  return pump(streams, n);
}

uint64_t pz_stream_counter(const pz_stream *s, int which) {
  if (!s) return 0;
  switch (which) {
    case PZ_SC_PUMPS: return s->pumps;
    case PZ_SC_RESUMED: return s->resumed;
    case PZ_SC_CKPT_BIT: return s->lead * 8u + ((s->ck[1] != 0 && s->ck[1] != PZ_CK_TRAILER_HOST) ? s->ck[1] : s->ck[0]);
    case PZ_SC_CKPT_BYTES: return s->pos;
    case PZ_SC_DEVICE_BYTES: return s->d_in_cap + s->d_out_cap;
    case PZ_SC_DEVICE_PEAK: return s->peak_device;
    case PZ_SC_HOST_BYTES: return s->h_out.cap + s->stage.cap;
    default: return 0;
  }
}

int pz_stream_next(pz_stream *s, const uint8_t **chunk, size_t *len, pz_result *res) {
  if (!s) return PZ_E_ARG;
  if (chunk) *chunk = nullptr;
  if (len) *len = 0;
  if (s->done_delivered) {
    if (s->verdict.status == PZ_OK) return PZ_S_DONE;
    if (res) *res = s->verdict;
    return PZ_S_ERROR;
  }
  if (s->dirty && !s->terminal) {
    pz_stream *one = s;
    int rc = pump(&one, 1);
    if (rc != PZ_E_OK) return rc;
  }
  if (s->published < s->publish_to) { /* emitExcess hands out exactly 32 KiB at a time */
    if (chunk) *chunk = s->h_out.p + (s->published - s->h_first);
    if (len) *len = PZ_EXCESS_CHUNK;
    s->published += PZ_EXCESS_CHUNK;
    return PZ_S_CHUNK;
  }
  if (!s->terminal) return PZ_S_NEED_MORE;
  if (s->final_pending) {
    s->final_pending = false;
    if (chunk) *chunk = s->h_out.p ? s->h_out.p + (s->published - s->h_first) : (const uint8_t *)"";
    if (len) *len = (size_t)(s->verdict.out_len - s->published);
    s->published = s->verdict.out_len;
    return PZ_S_CHUNK;
  }
  s->done_delivered = true;
  if (s->verdict.status == PZ_OK) return PZ_S_DONE;
  if (res) *res = s->verdict;
  return PZ_S_ERROR;
}

/* ---- `map decompress` in one call (what the Haskell shim's decompressBatch binds) ---------------------- */
}  // extern "C"
struct pz_outputs {
  uint8_t *blob = nullptr;
  size_t cap = 0;
};
extern "C" {

int pz_decompress_batch(const uint8_t *const *in, const size_t *in_len, size_t n, pz_result *res, uint8_t **out,
                        pz_outputs **handle, uint32_t flags) {
  int rc = ensure_init();
  if (rc != PZ_E_OK) return rc;
  if (!handle) return PZ_E_ARG;
  *handle = nullptr;
  if (n == 0) return PZ_E_OK;
  if (!in || !in_len || !res || !out) return PZ_E_ARG;
  flags &= ~(uint32_t)PZ_F_COUNT_ONLY;
  std::vector<uint64_t> in_off(n + 1), out_off(n + 1);
  uint64_t ia = 0;
  for (size_t i = 0; i < n; i++) { in_off[i] = ia; ia += in_len[i]; }
  in_off[n] = ia;
  Workspace &ws = g_ws;
  if ((rc = ws.h_in.reserve(ia + 64)) != PZ_E_OK) return rc;
  uint8_t *h_in = (uint8_t *)ws.h_in.p;
  std::vector<CopyJob> jobs(n);
  for (size_t i = 0; i < n; i++) jobs[i] = CopyJob{h_in + in_off[i], in[i], in_len[i]};
  parallel_copy(jobs);
  /* the zlib format does not carry the decoded length: sizing pass first (Zlib.hs:32-51 returns a lazy ByteString of
   * whatever length comes out) */
  if ((rc = pz_inflate_batch_contig(h_in, in_off.data(), nullptr, nullptr, n, res, nullptr, flags | PZ_F_COUNT_ONLY)) != PZ_E_OK) return rc;
  uint64_t oa = 0;
  for (size_t i = 0; i < n; i++) { out_off[i] = oa; oa += (res[i].out_len + 15u) & ~(uint64_t)15u; }
  out_off[n] = oa;
  pz_outputs *h = new (std::nothrow) pz_outputs();
  if (!h) return PZ_E_NOMEM;
  h->blob = g_pinned.get((size_t)oa + 64, &h->cap);
  if (!h->blob) { delete h; return PZ_E_NOMEM; }
  rc = pz_inflate_batch_contig(h_in, in_off.data(), h->blob, out_off.data(), n, res, nullptr, flags);
  if (rc != PZ_E_OK) { pz_outputs_free(h); return rc; }
  for (size_t i = 0; i < n; i++) out[i] = h->blob + out_off[i];
  *handle = h;
  return PZ_E_OK;
}

void pz_outputs_free(pz_outputs *h) {
  if (!h) return;
  g_pinned.put(h->blob, h->cap);
  delete h;
}

/* ---- auxiliaries ----------------------------------------------------------------------- */
size_t pz_strerror(const pz_result *r, char *buf, size_t cap) {
  char tmp[256];
  tmp[0] = 0;
  if (!r) return 0;
  const long long p0 = (long long)r->payload[0];
  switch (r->status) {
    case PZ_OK: break;
    case PZ_ERR_HUFFMAN_TREE: {
      const char *m = "?";
      char t2[96];
      switch (r->detail) {
        case PZ_D_TWO_VALUES: m = "Two values point to the same place!"; break;
        case PZ_D_VALUE_HIT: m = "HuffmanValue hit while inserting a value!"; break;
        case PZ_D_LEAF_IS_NODE: snprintf(t2, sizeof t2, "Tried to add where the leaf is a node: %lld", p0); m = t2; break;
        case PZ_D_ADVANCE_EMPTY_TREE: m = "Tried to advance empty tree!"; break;
        case PZ_D_ADVANCED_TO_EMPTY: m = "Advanced to empty tree!"; break;
      }
      snprintf(tmp, sizeof tmp, "Huffman tree manipulation error: %s", m);
      break;
    }
    case PZ_ERR_FORMAT:
      if (r->detail == PZ_D_LEN_NLEN) snprintf(tmp, sizeof tmp, "Block format error: Len/nlen mismatch in uncompressed block.");
      else snprintf(tmp, sizeof tmp, "Block format error: Unacceptable BTYPE: %lld", p0);
      break;
    case PZ_ERR_DECOMPRESSION:
      snprintf(tmp, sizeof tmp, "Decompression error: %s",
               r->detail == PZ_D_RAN_OUT ? "Ran out of data mid-decompression 2." : "Finished with data remaining.");
      break;
    case PZ_ERR_HEADER:
      if (r->detail == PZ_D_HDR_CHECKSUM) snprintf(tmp, sizeof tmp, "Header error: Header checksum failed");
      else if (r->detail == PZ_D_HDR_METHOD) snprintf(tmp, sizeof tmp, "Header error: Bad compression method: %lld", p0);
      else if (r->detail == PZ_D_HDR_GZIP_MAGIC) snprintf(tmp, sizeof tmp, "Header error: Not a gzip stream: %llx", p0);
      else if (r->detail == PZ_D_HDR_GZIP_FLAGS) snprintf(tmp, sizeof tmp, "Header error: Reserved gzip flags set: %lld", p0);
      else snprintf(tmp, sizeof tmp, "Header error: Window size too big: %lld", p0);
      break;
    case PZ_ERR_CHECKSUM:
      if (r->detail == PZ_D_LENGTH_MISMATCH) snprintf(tmp, sizeof tmp, "Checksum error: length mismatch: %lld != %lld", p0, (long long)(uint32_t)r->out_len);
      else snprintf(tmp, sizeof tmp, "Checksum error: checksum mismatch: %x != %x", r->adler_stored, r->adler_computed);
      break;
    case PZ_REF_BOTTOM:
      snprintf(tmp, sizeof tmp, "_|_ %s",
               r->detail == PZ_D_BOT_LENGTH_SYM ? "lengthArray index" : r->detail == PZ_D_BOT_DIST_SYM ? "distanceArray index"
               : r->detail == PZ_D_BOT_DIST_TOO_FAR ? "negative slice" : "window overflow");
      break;
    case PZ_OUTPUT_FULL: snprintf(tmp, sizeof tmp, "output buffer full"); break;
    default: snprintf(tmp, sizeof tmp, "status %d", r->status);
  }
  size_t n = strlen(tmp);
  if (buf && cap) {
    size_t k = std::min(n, cap - 1);
    memcpy(buf, tmp, k);
    buf[k] = 0;
  }
  return n;
}

int pz_compute_code_values(const int32_t *sym, const int32_t *len, int n, int32_t *out_triples) {
  int rc = ensure_init();
  if (rc != PZ_E_OK) return rc;
  if (n < 0 || (n && (!sym || !len || !out_triples))) return PZ_E_ARG;
  uint8_t lens[288] = {0};
  for (int i = 0; i < n; i++) {
    if (sym[i] < 0 || sym[i] >= 288 || len[i] < 0 || len[i] > 15) return PZ_E_ARG;
    lens[sym[i]] = (uint8_t)len[i];
  }
  uint8_t *d_lens = nullptr;
  uint16_t *d_codes = nullptr, codes[288];
  PZ_CUDA(cudaMalloc(&d_lens, 288));
  PZ_CUDA(cudaMalloc(&d_codes, 288 * sizeof(uint16_t)));
  PZ_CUDA(cudaMemcpy(d_lens, lens, 288, cudaMemcpyHostToDevice));
  PZ_CUDA(pz_launch_code_values(d_lens, 288, d_codes, nullptr));
  PZ_CUDA(cudaMemcpy(codes, d_codes, sizeof codes, cudaMemcpyDeviceToHost));
  cudaFree(d_lens);
  cudaFree(d_codes);
  int m = 0;
  for (int s = 0; s < 288; s++)
    if (lens[s]) { out_triples[3 * m] = s; out_triples[3 * m + 1] = lens[s]; out_triples[3 * m + 2] = codes[s]; m++; }
  return m;
}

uint32_t pz_adler32(uint32_t init, const uint8_t *data, size_t len) {
  if (ensure_init() != PZ_E_OK) return 0;
  /* checksum of `data` from the initial state on the device, then zlib's combine identity
   * to continue from `init` (two modular adds and one multiply; no bytes touched here) */
  Workspace &ws = g_ws;
  uint64_t off[2] = {0, len};
  std::vector<uint64_t> seg;
  uint64_t total = build_seg_off(off, 1, seg);
  if (ws.d_out.reserve(len + 64) != PZ_E_OK || ws.d_out_off.reserve(16) != PZ_E_OK || ws.d_seg_off.reserve(16) != PZ_E_OK ||
      ws.d_res.reserve(sizeof(pz_result)) != PZ_E_OK || ws.d_parts.reserve(std::max<uint64_t>(total, 1) * sizeof(uint2)) != PZ_E_OK)
    return 0;
  pz_result r;
  memset(&r, 0, sizeof r);
  r.status = PZ_OK; r.out_len = len;
  if (len && cudaMemcpy(ws.d_out.p, data, len, cudaMemcpyHostToDevice) != cudaSuccess) return 0;
  if (cudaMemcpy(ws.d_out_off.p, off, 16, cudaMemcpyHostToDevice) != cudaSuccess) return 0;
  if (cudaMemcpy(ws.d_seg_off.p, seg.data(), 16, cudaMemcpyHostToDevice) != cudaSuccess) return 0;
  if (cudaMemcpy(ws.d_res.p, &r, sizeof r, cudaMemcpyHostToDevice) != cudaSuccess) return 0;
  if (pz_launch_adler((const uint8_t *)ws.d_out.p, (const uint64_t *)ws.d_out_off.p, (const uint64_t *)ws.d_seg_off.p, 1, 0, 1, 0, total,
                      (pz_result *)ws.d_res.p, (uint2 *)ws.d_parts.p, nullptr) != cudaSuccess)
    return 0;
  if (cudaMemcpy(&r, ws.d_res.p, sizeof r, cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
  const uint32_t M = 65521u;
  uint32_t a1 = init & 0xffffu, b1 = init >> 16, a2 = r.adler_computed & 0xffffu, b2 = r.adler_computed >> 16;
  uint32_t a = (a1 + a2 + M - 1u) % M;
  uint32_t b = (uint32_t)((b1 + b2 + (uint64_t)(len % M) * ((a1 + M - 1u) % M)) % M);
  return (b << 16) | a;
}

/* CRC-32 of a host buffer with the kernels of the gzip framing; `init` continues an earlier sum through the combine
 * identity crc(A || B) = crc(A) * x^(8|B|) + crc(B) over GF(2) modulo the CRC polynomial (no bytes touched here). */
uint32_t pz_crc32(uint32_t init, const uint8_t *data, size_t len) {
  if (ensure_init() != PZ_E_OK) return 0;
  Workspace &ws = g_ws;
  uint64_t off[2] = {0, len};
  std::vector<uint64_t> seg;
  uint64_t total = build_seg_off(off, 1, seg);
  if (ws.d_out.reserve(len + 64) != PZ_E_OK || ws.d_out_off.reserve(16) != PZ_E_OK || ws.d_seg_off.reserve(16) != PZ_E_OK ||
      ws.d_res.reserve(sizeof(pz_result)) != PZ_E_OK || ws.d_parts.reserve(std::max<uint64_t>(total, 1) * sizeof(uint2)) != PZ_E_OK)
    return 0;
  pz_result r;
  memset(&r, 0, sizeof r);
  r.status = PZ_OK; r.out_len = len; r.payload[0] = (int64_t)(uint32_t)len; /* ISIZE agrees: only the CRC is of interest */
  if (len && cudaMemcpy(ws.d_out.p, data, len, cudaMemcpyHostToDevice) != cudaSuccess) return 0;
  if (cudaMemcpy(ws.d_out_off.p, off, 16, cudaMemcpyHostToDevice) != cudaSuccess) return 0;
  if (cudaMemcpy(ws.d_seg_off.p, seg.data(), 16, cudaMemcpyHostToDevice) != cudaSuccess) return 0;
  if (cudaMemcpy(ws.d_res.p, &r, sizeof r, cudaMemcpyHostToDevice) != cudaSuccess) return 0;
  if (pz_launch_adler((const uint8_t *)ws.d_out.p, (const uint64_t *)ws.d_out_off.p, (const uint64_t *)ws.d_seg_off.p, 1, 0, 1, 0, total,
                      (pz_result *)ws.d_res.p, (uint2 *)ws.d_parts.p, nullptr, 1u) != cudaSuccess)
    return 0;
  if (cudaMemcpy(&r, ws.d_res.p, sizeof r, cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
  const uint32_t crc2 = r.adler_computed;
  if (init == 0u) return crc2;
  /* x^(8 len) mod P by square and multiply, then crc(init-part) * x^(8 len) + crc2 */
  auto mul = [](uint32_t a, uint32_t b) {
    uint32_t p = 0;
    for (uint32_t m = 0x80000000u; m != 0u && a != 0u; m >>= 1) {
      if (a & m) { p ^= b; a &= ~m; }
      b = (b >> 1) ^ (0xedb88320u & (0u - (b & 1u)));
    }
    return p;
  };
  uint32_t xp = 0x80000000u, sq = 0x40000000u; /* x^0, x^1 */
  for (int k = 0; k < 3; k++) sq = mul(sq, sq); /* x^8 */
  for (uint64_t n = len; n != 0u; n >>= 1) {
    if (n & 1u) xp = mul(sq, xp);
    sq = mul(sq, sq);
  }
  return mul(xp, init) ^ crc2;
}

}  // extern "C"
