// engine.cu — libzkemail_b200.so: the C ABI of include/zkemail_b200.h.
//
// Replaces the native execution of zkemail_core::verify_email (core/src/circuits.rs:9-29) and
// verify_email_with_regex (core/src/circuits.rs:31-68) for whole batches:
//   host  : header split, DKIM-Signature tag parsing, header selection, canonicalisation, base64
//           (dkim_host.hpp) on a thread pool, written straight into pinned staging blocks;
//   device: SHA-256 of every body / header preimage / from_domain / key (sha256.cuh), bh= compare,
//           RSA PKCS#1 v1.5 verification (rsa.cuh), regex DFA scans (dfa.cuh);
//   host  : per-email resolution of the reference's control flow (first passing signature wins,
//           every panic site becomes a status code).
// There is no CPU fallback for the arithmetic: without a CUDA device zkb_engine_create fails.
#include <cuda_runtime.h>

#include <stddef.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/zkemail_b200.h"
#include "common.cuh"
#include "dfa_host.hpp"
#include "dkim_host.hpp"
#include "kernels.h"
#include "keytab.hpp"
#include "regexc.hpp"

using namespace zkb;

namespace {

#define CK(expr)                                                                             \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      fprintf(stderr, "[zkemail_b200] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(_e),   \
              __FILE__, __LINE__, cudaGetErrorString(_e));                                   \
      return ZKB_E_CUDA;                                                                     \
    }                                                                                        \
  } while (0)

// inside a loop that owns in-flight work: record the error and leave the loop; the common exit path cleans up
#define CKB(expr)                                                                            \
  {                                                                                          \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      fprintf(stderr, "[zkemail_b200] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(_e),   \
              __FILE__, __LINE__, cudaGetErrorString(_e));                                   \
      rc = ZKB_E_CUDA;                                                                       \
      break;                                                                                 \
    }                                                                                        \
  }

// device temporaries of the kernel-level entry points: freed on every exit path
struct DevTmp {
  std::vector<void*> ptrs;
  ~DevTmp() { for (void* p : ptrs) cudaFree(p); }
  template <typename T> cudaError_t alloc(T** out, size_t bytes) {
    void* p = nullptr;
    cudaError_t r = cudaMalloc(&p, bytes ? bytes : 16);
    if (r == cudaSuccess) ptrs.push_back(p);
    *out = (T*)p;
    return r;
  }
};

// ------------------------------------------------------------------ thread pool
// Parallel phases follow one another within microseconds in the chunk pipeline, so workers spin on a
// generation counter for a short while before falling back to a condition variable.
class ThreadPool {
 public:
  explicit ThreadPool(int n) : n_(n < 1 ? 1 : n) {
    for (int i = 1; i < n_; i++) th_.emplace_back([this, i] { worker(i); });
  }
  ~ThreadPool() {
    { std::lock_guard<std::mutex> l(mu_); stop_.store(true); gen_.fetch_add(1); }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  int size() const { return n_; }
  // runs fn(tid) once on every thread of the pool (the caller is tid 0) and waits for all
  void run(const std::function<void(int)>& fn) {
    job_ = &fn;
    pending_.store(n_ - 1, std::memory_order_release);
    { std::lock_guard<std::mutex> l(mu_); gen_.fetch_add(1, std::memory_order_release); }
    if (sleepers_.load(std::memory_order_acquire) > 0) cv_.notify_all();
    fn(0);
    int spins = 0;
    while (pending_.load(std::memory_order_acquire) != 0) {
      if (++spins < 2000) cpu_relax(); else std::this_thread::yield();
    }
    job_ = nullptr;
  }
  // dynamic parallel-for over [0,n) in grains
  void parallel_for(size_t n, size_t grain, const std::function<void(size_t, size_t, int)>& body) {
    if (n == 0) return;
    std::atomic<size_t> next{0};
    run([&](int tid) {
      for (;;) {
        size_t lo = next.fetch_add(grain);
        if (lo >= n) break;
        body(lo, std::min(n, lo + grain), tid);
      }
    });
  }

 private:
  static void cpu_relax() {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
  void worker(int tid) {
    uint64_t seen = 0;
    for (;;) {
      int spins = 0;
      while (gen_.load(std::memory_order_acquire) == seen) {
        if (++spins < 20000) { cpu_relax(); continue; }
        std::unique_lock<std::mutex> l(mu_);
        sleepers_.fetch_add(1);
        cv_.wait(l, [&] { return gen_.load(std::memory_order_acquire) != seen; });
        sleepers_.fetch_sub(1);
      }
      seen = gen_.load(std::memory_order_acquire);
      if (stop_.load()) return;
      const std::function<void(int)>* job = job_;
      if (job) (*job)(tid);
      pending_.fetch_sub(1, std::memory_order_acq_rel);
    }
  }
  int n_;
  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable cv_;
  const std::function<void(int)>* volatile job_ = nullptr;
  std::atomic<uint64_t> gen_{0};
  std::atomic<int> pending_{0}, sleepers_{0};
  std::atomic<bool> stop_{false};
};

// ------------------------------------------------------------------ pinned staging blocks
struct PinBlock {
  uint8_t* p = nullptr;
  size_t cap = 0, used = 0;
  uint64_t dev_off = 0;  // where this block's bytes live in the chunk's device arena
};
class BlockPool {
 public:
  static constexpr size_t kBlock = 8u << 20;
  bool get(size_t min_cap, PinBlock& b) {
    size_t want = std::max(min_cap, kBlock);
    {
      std::lock_guard<std::mutex> l(mu_);
      for (size_t i = 0; want == kBlock && i < free_.size(); i++)
        if (free_[i].cap == kBlock) {
          b = free_[i]; free_[i] = free_.back(); free_.pop_back();
          b.used = 0;
          return true;
        }
    }
    void* p = nullptr;
    if (cudaHostAlloc(&p, want, cudaHostAllocDefault) != cudaSuccess) return false;
    b.p = (uint8_t*)p; b.cap = want; b.used = 0; b.dev_off = 0;
    return true;
  }
  void put(PinBlock& b) {
    if (!b.p) return;
    if (b.cap > kBlock) {   // one-off block for an oversized message: give the pinned memory back
      cudaFreeHost(b.p);
      b = PinBlock();
      return;
    }
    std::lock_guard<std::mutex> l(mu_);
    free_.push_back(b);
    b = PinBlock();
  }
  void release_all() {
    std::lock_guard<std::mutex> l(mu_);
    for (auto& b : free_) cudaFreeHost(b.p);
    free_.clear();
  }
 private:
  std::mutex mu_;
  std::vector<PinBlock> free_;
};

struct PinBuf {  // growable pinned buffer
  uint8_t* p = nullptr;
  size_t cap = 0;
  bool ensure(size_t n) {
    if (n <= cap) return true;
    if (p) cudaFreeHost(p);
    size_t want = std::max(n + n / 4, (size_t)1 << 16);
    void* q = nullptr;
    if (cudaHostAlloc(&q, want, cudaHostAllocDefault) != cudaSuccess) { p = nullptr; cap = 0; return false; }
    p = (uint8_t*)q; cap = want;
    return true;
  }
  void free() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};
struct DevBuf {  // growable device buffer
  uint8_t* p = nullptr;
  size_t cap = 0;
  bool ensure(size_t n) {
    if (n <= cap) return true;
    if (p) cudaFree(p);
    size_t want = std::max(n + n / 4, (size_t)1 << 16);
    void* q = nullptr;
    if (cudaMalloc(&q, want) != cudaSuccess) { p = nullptr; cap = 0; return false; }
    p = (uint8_t*)q; cap = want;
    return true;
  }
  void free() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// ------------------------------------------------------------------ records produced by the host pass
struct MsgRec {         // one SHA-256 message / DFA haystack
  uint64_t goff;        // byte offset in the device arena (after layout)
  uint32_t len, blk;    // blk == VIRT_BLK: a body canonicalised on the device; local = offset in the thread's
  uint32_t canon;       // virtual region (resident chunks can exceed 4 GiB), len = raw length (upper-bound
  uint64_t local;       // estimate), canon = its item; otherwise local = offset inside staging block blk
};
constexpr uint32_t VIRT_BLK = 0xFFFFFFFFu;
enum { SIG_OK = 0, SIG_SYNTAX = 1, SIG_BADLEN = 2 };
struct CandRec {        // one DKIM-Signature header that reaches the cryptographic checks
  uint32_t body_msg, hdr_msg;  // thread-local message indices
  uint32_t bh[8];              // decoded bh= as native SHA-256 state words
  uint32_t sig_off;            // word offset into the thread's signature words
  int32_t key_id;
  uint8_t bh_valid, sig_state, algo, haystack_only;
  uint8_t rsa_list;            // which of the 6 RSA launch lists (valid when the candidate goes to the device)
  uint32_t raw_blk, raw_local; // staged device front end: where the host copied the raw message (staging block, offset)
};
enum { STEP_ERR = 0, STEP_CAND = 1, STEP_SHA1 = 2 };
struct StepRec { uint32_t cand; uint8_t kind, detail; };
struct EmailRec {
  int32_t status = 0;
  uint32_t fe_cand = 0;     // device-front-end mode: thread-local candidate / slot index of this email
  uint32_t tid = 0, first_step = 0, n_steps = 0;
  uint32_t dom_msg = 0, key_msg = 0;
  int32_t canon_rc = 2;     // zo_canonicalize_signed_email return code equivalent
  uint32_t canon_cand = 0;  // thread-local candidate whose messages are the regex haystacks
};
struct KeyMeta { uint32_t limbs_class = 0, k = 0; bool generic = false; };

struct ThreadRecs {
  std::vector<PinBlock> blocks;
  std::vector<MsgRec> msgs;
  std::vector<CandRec> cands;
  std::vector<StepRec> steps;
  std::vector<uint32_t> sigw;
  std::vector<CanonItem> canon;     // bodies to canonicalise on the device (direct mode)
  std::vector<uint32_t> fe_emails;  // device-front-end mode: chunk-local email index of candidate j
  size_t fe_sig_words = 0;          // device-front-end mode: signature limbs are written by the device only
  uint64_t virt_used = 0, virt_base = 0;   // device-only arena region of this thread (canonical body slots)
  uint32_t canon_base = 0;
  std::vector<uint32_t> hist;       // messages per SHA block count (maintained by commit)
  uint64_t sha_blocks = 0, sha_bytes = 0;
  uint32_t rsa_cnt[6] = {0, 0, 0, 0, 0, 0};
  uint32_t msg_base = 0, cand_base = 0;
  void clear() {
    msgs.clear(); cands.clear(); steps.clear(); sigw.clear(); hist.clear(); canon.clear(); fe_emails.clear();
    virt_used = 0; fe_sig_words = 0;
    sha_blocks = sha_bytes = 0;
    for (auto& c : rsa_cnt) c = 0;
  }
};

struct DeviceChunk {   // everything one chunk needs in HBM
  DevBuf arena, meta, out, span;   // span: raw message bytes DMA'd from registered host memory (direct mode)
  const uint8_t* raw_base = nullptr;   // what raw offsets are relative to: span (direct mode) or the arena (staged raw messages)
  const CanonItem* canon_items = nullptr; uint32_t n_canon = 0;
  const uint32_t* msg_canon = nullptr;   // message index -> canon item (or ~0): bodies are canonicalised in the SHA order (similar lengths per warp)
  uint32_t* msg_len_rw = nullptr;
  const FeIn* fe_in = nullptr; FeOut* fe_out = nullptr; uint32_t n_fe = 0;   // device front end
  CanonItem* canon_rw = nullptr; uint32_t* cand_bh_rw = nullptr; uint32_t* sig_rw = nullptr;
  // pointers into meta / out
  const uint64_t* msg_off = nullptr; const uint32_t* msg_len = nullptr; const uint32_t* order = nullptr;
  const uint32_t* cand_body = nullptr; const uint32_t* cand_bh = nullptr;
  const uint32_t* sig_arena = nullptr;
  const RsaItem* rsa_items[6] = {nullptr}; uint32_t rsa_n[6] = {0};  // [class 32,64,128] x [e=65537, generic]
  const DfaItem* dfa_items = nullptr; uint32_t n_dfa = 0;
  uint32_t* digests = nullptr; uint32_t* cand_flags = nullptr; uint4* dfa_out = nullptr;
  uint32_t M = 0, C = 0, NE = 0, P = 0;
  size_t out_bytes = 0;
  uint32_t* recs = nullptr; size_t o_recs = 0; uint32_t rec_words = 0;   // device-built result records (assemble.cuh)
  const FeIn* asm_in = nullptr; const FeOut* asm_fo = nullptr; uint32_t n_asm = 0;   // what the records are built from
  void free() { arena.free(); meta.free(); out.free(); span.free(); }
};

struct Chunk {  // host view of one chunk
  size_t e0 = 0, ne = 0;
  std::vector<EmailRec> emails;
  std::vector<ThreadRecs> tr;
  zkb_batch_stats st;
  size_t arena_bytes = 0, meta_bytes = 0;
  uint32_t M = 0, C = 0;
  // offsets inside meta
  size_t o_msg_off = 0, o_msg_len = 0, o_order = 0, o_cand_body = 0, o_cand_bh = 0, o_sig = 0, o_rsa[6] = {0}, o_dfa = 0;
  uint32_t rsa_n[6] = {0}, n_dfa = 0;
  // direct mode: raw bodies stay in the caller's registered memory and are canonicalised on the device
  bool direct = false;
  const uint8_t* span_host = nullptr;
  size_t span_bytes = 0, o_canon = 0, o_msg_canon = 0;
  uint32_t n_canon = 0;
  // device front end (frontend.cuh): headers parsed / preimages built / base64 decoded on the device too
  bool fe = false;
  bool staged = false;   // fe without direct: the host copies each raw message into the staging blocks (pageable callers)
  size_t o_fein = 0;
  size_t o_asm_in = 0, o_asm_fo = 0;   // host front end: per-email inputs of the record kernel (n_asm = ne)
  const zkb_email_view* views = nullptr;   // the caller's views of this chunk (borrowed for the call)
  size_t upload_bytes = 0;                 // leading part of the meta buffer that is copied host -> device
};

struct Slot {
  DeviceChunk dev;
  PinBuf meta, result;
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  cudaEvent_t k0 = nullptr, k1 = nullptr, h0 = nullptr;   // ZKB_PROFILE: H2D start, kernels start / end
};

}  // namespace

struct zkb_regex_set {
  zkb_engine* eng = nullptr;
  size_t n_header = 0, n_body = 0;
  int header_present = 0, body_present = 0;
  struct Part { uint8_t* d_fwd = nullptr; uint8_t* d_rev = nullptr; uint32_t fwd_bytes = 0, rev_bytes = 0, elem = 2; bool body = false, direct = false; };
  std::vector<Part> parts;
  size_t n_active() const { return (header_present ? n_header : 0) + (body_present ? n_body : 0); }
};

struct zkb_engine {
  int device = 0, sm_count = 148;
  size_t smem_optin = 0;
  int64_t now_unix = 0;
  size_t chunk_emails = 65536;   // e2e pipeline chunk (also capped at 256 MB of raw bytes); resident batches use 4x
  uint32_t rsa_lanes = 4;   // lanes per 2048-bit signature (measured best on B200: 0.89 of the IMAD.WIDE peak)
  std::atomic<uint32_t> flags{0};   // ZKB_OPT_* (zkb_options.flags / zkb_engine_set_flags)
  bool has(uint32_t bit) const { return (flags.load(std::memory_order_relaxed) & bit) != 0; }
  ThreadPool* pool = nullptr;
  BlockPool blocks;
  Slot slots[3];
  std::mutex run_mu;
  std::vector<std::pair<const uint8_t*, size_t>> registered;  // cudaHostRegister'ed caller memory
  int borrowed_registrations = 0;   // trailing entries of `registered` page-locked by another engine of a zkb_multi (not unregistered here)
  // public keys
  std::mutex key_mu;
  std::unordered_map<std::string, int32_t> key_index;
  std::vector<uint32_t> keytab_host;
  std::vector<KeyMeta> key_meta;
  uint32_t* d_keytab = nullptr;
  size_t d_keytab_cap = 0, d_keytab_n = 0;
  std::atomic<int> live_batches{0};   // resident batches hold key ids: the table is not trimmed while any is alive
  cudaEvent_t ev[10] = {nullptr};
  cudaStream_t aux_stream = nullptr;          // zkb_batch_run_async: hashing of the next chunk beside the RSA of this one
  cudaStream_t aux_stream2 = nullptr;         // raw-resident batches: chunks alternate between the two side streams
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join2 = nullptr;
  std::vector<cudaEvent_t> ev_pre;            // one per resident chunk index: "pre phase of chunk k done"
  uint64_t last_h2d = 0, last_d2h = 0, last_fallback = 0;   // of the last zkb_verify_batch
  // ZKB_PROFILE accounting of the call in progress (calls on one engine are serialised by run_mu)
  double prof_parse = 0, prof_layout = 0, prof_prelude = 0;
  std::atomic<uint64_t> prof_busy_ns{0};
};

struct zkb_batch {
  zkb_engine* eng = nullptr;
  const zkb_regex_set* regex = nullptr;
  size_t n = 0;
  std::vector<Chunk*> chunks;
  std::vector<DeviceChunk*> dev;
  const zkb_email_view* emails = nullptr;       // borrowed until fetch
  const zkb_email_captures* captures = nullptr;  // borrowed until fetch
  float last_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // sha256, rsa, dfa, bh check, whole run, front end + canonicalisation, records
  bool ran = false;
  bool raw = false;   // raw messages resident: every run redoes the device front end and the canonicalisation
  uint64_t serial = 0;   // unique per prepared batch in this process (an address may be reused, a serial is not)
};

namespace {

inline int rsa_list_of(const KeyMeta& k) { return (k.limbs_class == 32 ? 0 : k.limbs_class == 64 ? 1 : 2) * 2 + (k.generic ? 1 : 0); }

// ------------------------------------------------------------------ per-thread parse context
struct ThreadCtx {
  zkb_engine* eng;
  ThreadRecs* tr;
  std::vector<HeaderField> hs;
  DkimSig sig;
  std::string scratch;
  std::vector<uint8_t> tmp;
  std::unordered_map<std::string, std::pair<int32_t, KeyMeta>> key_cache;
  struct PtrKey { const uint8_t* p = nullptr; size_t len = 0; int32_t id = 0; KeyMeta meta; };
  PtrKey ptr_cache[256];  // same buffer => same bytes: skips the content hash for pooled keys
  static size_t ptr_slot(const void* p) { return (size_t)(((uint64_t)(uintptr_t)p * 0x9E3779B97F4A7C15ull) >> 56); }
  struct PtrDom { const char* p = nullptr; size_t len = 0; uint32_t msg = 0; };
  PtrDom dom_cache[256];
  std::unordered_map<std::string, uint32_t> dom_msgs;
  std::unordered_map<int32_t, uint32_t> key_msgs;
  bool oom = false;
  bool direct = false;                 // bodies are canonicalised on the device from the raw span
  bool staged = false;                 // device front end over raw messages the host copies into the staging blocks
  const uint8_t* span_host = nullptr;

  // a device-only arena slot for a body the canon kernel will write (no host bytes)
  uint32_t add_virtual_body(const uint8_t* body, size_t body_len, bool relaxed, bool has_l, uint64_t l) {
    CanonItem it;
    memset(&it, 0, sizeof it);
    it.raw_off = (uint64_t)(body - span_host);
    it.raw_len = (uint32_t)body_len;
    it.flags = (relaxed ? 1u : 0u) | (has_l ? 2u : 0u);
    it.l = l > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)l;
    it.msg = (uint32_t)tr->msgs.size();   // thread-local; made global at layout
    MsgRec m;
    m.goff = 0; m.len = (uint32_t)body_len + 2; m.blk = VIRT_BLK; m.local = tr->virt_used;
    m.canon = (uint32_t)tr->canon.size();
    tr->virt_used += (((body_len + 2) >> 6) + 1) << 6;
    tr->canon.push_back(it);
    tr->msgs.push_back(m);
    const uint32_t nb = (uint32_t)(m.len >> 6) + 1 + ((m.len & 63) >= 56 ? 1u : 0u);  // estimate from the raw length
    if (tr->hist.size() <= nb) tr->hist.resize((size_t)nb + 1, 0u);
    tr->hist[nb]++;
    tr->sha_blocks += nb; tr->sha_bytes += body_len;
    return (uint32_t)tr->msgs.size() - 1;
  }

  // reserve `need` bytes (64-byte aligned start) in the thread's current staging block
  uint8_t* reserve(size_t need, uint32_t& blk, uint32_t& local) {
    if (tr->blocks.empty() || tr->blocks.back().used + need > tr->blocks.back().cap) {
      PinBlock b;
      if (!eng->blocks.get(need, b)) { oom = true; return nullptr; }
      tr->blocks.push_back(b);
    }
    PinBlock& b = tr->blocks.back();
    blk = (uint32_t)tr->blocks.size() - 1;
    local = (uint32_t)b.used;
    return b.p + b.used;
  }
  // commit a message of `len` bytes written at the reserved position
  uint32_t commit(uint32_t blk, uint32_t local, size_t len) {
    PinBlock& b = tr->blocks[blk];
    b.used = local + (((len >> 6) + 1) << 6);  // readable up to the block after the last full one
    MsgRec m;
    m.goff = 0; m.len = (uint32_t)len; m.blk = blk; m.local = local; m.canon = 0;
    tr->msgs.push_back(m);
    const uint32_t nb = (uint32_t)(len >> 6) + 1 + ((len & 63) >= 56 ? 1u : 0u);  // SHA-256 compressions
    if (tr->hist.size() <= nb) tr->hist.resize((size_t)nb + 1, 0u);
    tr->hist[nb]++;
    tr->sha_blocks += nb; tr->sha_bytes += len;
    return (uint32_t)tr->msgs.size() - 1;
  }
  uint32_t add_msg(const uint8_t* data, size_t len) {
    uint32_t blk, local;
    uint8_t* p = reserve(((len >> 6) + 1) << 6, blk, local);
    if (!p) return 0;
    if (len) memcpy(p, data, len);
    return commit(blk, local, len);
  }
};

// Copy of one raw message into pinned staging (dst 64-byte aligned, capacity rounded up to 64).  Streaming
// stores: the destination is read next by the DMA engine, not by a core, so it should not displace the cache
// or cost a read-for-ownership.  Callers fence once per pack pass (stage_fence).
#if defined(__x86_64__)
__attribute__((target("avx2"))) inline void stage_copy_avx2(uint8_t* dst, const uint8_t* src, size_t n) {
  size_t i = 0;
  for (; i + 64 <= n; i += 64) {
    const __m256i a = _mm256_loadu_si256((const __m256i*)(src + i));
    const __m256i b = _mm256_loadu_si256((const __m256i*)(src + i + 32));
    _mm256_stream_si256((__m256i*)(dst + i), a);
    _mm256_stream_si256((__m256i*)(dst + i + 32), b);
  }
  if (i < n) memcpy(dst + i, src + i, n - i);
}
inline void stage_copy(uint8_t* dst, const uint8_t* src, size_t n) {
  if (n >= 256 && cpu_has_avx2()) stage_copy_avx2(dst, src, n); else if (n) memcpy(dst, src, n);
}
inline void stage_fence() { _mm_sfence(); }
#else
inline void stage_copy(uint8_t* dst, const uint8_t* src, size_t n) { if (n) memcpy(dst, src, n); }
inline void stage_fence() {}
#endif

int32_t lookup_key(ThreadCtx& c, const uint8_t* der, size_t len, KeyMeta& meta) {
  ThreadCtx::PtrKey& pk = c.ptr_cache[ThreadCtx::ptr_slot(der)];
  if (pk.p == der && pk.len == len) { meta = pk.meta; return pk.id; }
  std::string k((const char*)der, len);
  auto it = c.key_cache.find(k);
  if (it != c.key_cache.end()) {
    meta = it->second.second;
    pk.p = der; pk.len = len; pk.id = it->second.first; pk.meta = meta;
    return it->second.first;
  }
  zkb_engine* e = c.eng;
  int32_t id = -1;
  meta = KeyMeta();
  bool known = false;
  {
    std::lock_guard<std::mutex> l(e->key_mu);
    auto g = e->key_index.find(k);
    if (g != e->key_index.end()) { id = g->second; meta = e->key_meta[id]; known = true; }
  }
  if (!known) {
    // DER decode and the Montgomery constants (~1 M limb operations for a 4096-bit modulus) run outside the lock;
    // rejected keys are remembered by the calling thread only (c.key_cache), never in the engine-wide table
    RsaKeyInfo info;
    if (parse_rsa_public_key(der, len, info)) {
      std::vector<uint32_t> ent(ZKB_KEY_STRIDE);
      build_key_entry(info, ent.data());
      std::lock_guard<std::mutex> l(e->key_mu);
      auto g = e->key_index.find(k);
      if (g != e->key_index.end()) id = g->second;   // another thread was faster
      else {
        id = (int32_t)e->key_meta.size();
        e->keytab_host.insert(e->keytab_host.end(), ent.begin(), ent.end());
        KeyMeta m;
        m.limbs_class = info.limbs_class; m.k = info.k; m.generic = info.e != 65537;
        e->key_meta.push_back(m);
        e->key_index.emplace(k, id);
      }
      meta = e->key_meta[id];
    }
  }
  c.key_cache.emplace(std::move(k), std::make_pair(id, meta));
  pk.p = der; pk.len = len; pk.id = id; pk.meta = meta;
  return id;
}

// Host pass over one email: everything of verify_dkim / canonicalize_signed_email that is byte
// shuffling.  Mirrors the control flow of cfdkim::verify_email_with_key.
void process_email(ThreadCtx& c, const zkb_email_view& em, bool want_regex, int tid, EmailRec& rec) {
  rec = EmailRec();
  rec.tid = (uint32_t)tid;
  rec.first_step = (uint32_t)c.tr->steps.size();
  const uint8_t* raw = em.raw_email;
  const size_t n = em.raw_email_len;
  size_t body_off = 0;
  if (n > 0xFFFFFF00u || !parse_headers(raw, n, c.hs, body_off)) { rec.status = ZKB_ST_MAIL_PARSE; return; }
  int32_t key_id = -1;
  KeyMeta km = KeyMeta();
  if (em.key_type_len == 3 && memcmp(em.key_type, "rsa", 3) == 0) {
    key_id = lookup_key(c, em.key, em.key_len, km);
    if (key_id < 0) { rec.status = ZKB_ST_KEY; return; }
  } else if (em.key_type_len == 7 && memcmp(em.key_type, "ed25519", 7) == 0) {
    rec.status = em.key_len == 32 ? ZKB_ST_UNSUPPORTED : ZKB_ST_KEY;
    return;
  } else { rec.status = ZKB_ST_KEY; return; }

  size_t body_len = 0;
  const uint8_t* body = nullptr;
  bool body_found = false;
  struct BodyKey { bool relaxed, has_l; uint64_t l; uint32_t msg; };
  BodyKey bodies[4];
  int n_bodies = 0;
  auto body_msg_for = [&](bool relaxed, bool has_l, uint64_t l) -> uint32_t {
    for (int i = 0; i < n_bodies; i++)
      if (bodies[i].relaxed == relaxed && bodies[i].has_l == has_l && bodies[i].l == l) return bodies[i].msg;
    if (!body_found) { body = find_body(raw, n, body_len, body_off); body_found = true; }
    uint32_t m;
    if (c.direct) {
      m = c.add_virtual_body(body, body_len, relaxed, has_l, l);
    } else {
      uint32_t blk, local;
      uint8_t* p = c.reserve((((body_len + 2) >> 6) + 1) << 6, blk, local);
      if (!p) return 0;
      size_t cl = relaxed ? canon_body_relaxed(body, body_len, p) : canon_body_simple(body, body_len, p);
      if (has_l && l < cl) cl = (size_t)l;
      m = c.commit(blk, local, cl);
    }
    if (n_bodies < 4) { bodies[n_bodies].relaxed = relaxed; bodies[n_bodies].has_l = has_l; bodies[n_bodies].l = l; bodies[n_bodies].msg = m; n_bodies++; }
    return m;
  };
  // builds the candidate for the signature currently held in c.sig
  auto make_cand = [&](bool hr, bool br, bool has_l, uint64_t l, uint8_t algo, bool haystack_only) -> uint32_t {
    CandRec cd;
    memset(&cd, 0, sizeof cd);
    cd.key_id = key_id; cd.algo = algo; cd.haystack_only = haystack_only ? 1 : 0;
    cd.body_msg = body_msg_for(br, has_l, l);
    uint32_t blk, local;
    size_t bound = preimage_bound(body_off, c.sig.n);
    uint8_t* p = c.reserve(((bound >> 6) + 1) << 6, blk, local);
    if (!p) return 0;
    size_t pl = build_header_preimage(raw, c.hs, c.sig, hr, p, c.scratch);
    cd.hdr_msg = c.commit(blk, local, pl);
    if (!haystack_only) {
      const Tag* tbh = c.sig.get("bh");
      uint8_t bh[48];
      if (tbh->val_len == 44 && base64_decode(c.sig.val(tbh), 44, bh) == 32) {
        cd.bh_valid = 1;
        for (int i = 0; i < 8; i++)
          cd.bh[i] = ((uint32_t)bh[4 * i] << 24) | ((uint32_t)bh[4 * i + 1] << 16) | ((uint32_t)bh[4 * i + 2] << 8) | bh[4 * i + 3];
      }
      const Tag* tb = c.sig.get("b");
      c.tmp.resize(tb->val_len + 4);
      long sl = base64_decode(c.sig.val(tb), tb->val_len, c.tmp.data());
      if (sl < 0) cd.sig_state = SIG_SYNTAX;
      else if ((size_t)sl != km.k) cd.sig_state = SIG_BADLEN;
      else if (algo == 1) {
        cd.sig_state = SIG_OK;
        cd.rsa_list = (uint8_t)rsa_list_of(km);
        c.tr->rsa_cnt[cd.rsa_list]++;
        cd.sig_off = (uint32_t)c.tr->sigw.size();
        c.tr->sigw.resize(c.tr->sigw.size() + km.limbs_class, 0u);
        uint32_t* w = c.tr->sigw.data() + cd.sig_off;
        for (long i = 0; i < sl; i++) {
          long bi = sl - 1 - i;
          w[bi >> 2] |= (uint32_t)c.tmp[i] << (8 * (bi & 3));
        }
      }
    }
    c.tr->cands.push_back(cd);
    return (uint32_t)c.tr->cands.size() - 1;
  };

  bool canon_seen = false;
  for (size_t hi = 0; hi < c.hs.size(); hi++) {
    const HeaderField& h = c.hs[hi];
    if (!ieq_ascii(raw + h.key_off, h.key_len, "DKIM-Signature", 14)) continue;
    int r = validate_dkim_header(raw + h.val_off, h.val_len, c.eng->now_unix, c.sig);
    if (r != ZKB_DKIM_PASS) {
      StepRec s; s.kind = STEP_ERR; s.detail = (uint8_t)r; s.cand = 0;
      c.tr->steps.push_back(s);
      continue;
    }
    // this header is the one canonicalize_signed_email would use if it is the first valid one
    const bool is_canon = want_regex && !canon_seen;
    canon_seen = true;
    bool hr = false, br = false;
    const bool canon_ok = parse_canon_tag(c.sig, hr, br);
    uint64_t lval = 0;
    bool has_l = false, l_ok = true;
    if (const Tag* tl = c.sig.get("l")) { has_l = true; l_ok = parse_usize_tag(c.sig, tl, lval); }
    long verify_cand = -1;
    {
      const Tag* td = c.sig.get("d");
      StepRec s; s.cand = 0; s.detail = 0;
      bool push = true;
      if (!ieq_ascii(c.sig.val(td), td->val_len, em.from_domain, em.from_domain_len)) push = false;  // skipped, not an error
      else if (!canon_ok) { s.kind = STEP_ERR; s.detail = ZKB_DKIM_CANON_TYPE; }
      else {
        const Tag* ta = c.sig.get("a");
        int algo = c.sig.val_is(ta, "rsa-sha1") ? 0 : c.sig.val_is(ta, "rsa-sha256") ? 1 : c.sig.val_is(ta, "ed25519-sha256") ? 2 : -1;
        if (algo < 0) { s.kind = STEP_ERR; s.detail = ZKB_DKIM_HASH_ALGO; }
        else if (algo == 0) { s.kind = STEP_SHA1; }
        else if (!l_ok) { s.kind = STEP_ERR; s.detail = ZKB_DKIM_LENGTH_TAG; }
        else {
          verify_cand = (long)make_cand(hr, br, has_l, lval, (uint8_t)algo, false);
          s.kind = STEP_CAND; s.cand = (uint32_t)verify_cand;
        }
      }
      if (push) c.tr->steps.push_back(s);
    }
    if (is_canon) {
      if (!canon_ok) rec.canon_rc = 3;
      else if (!l_ok) rec.canon_rc = 4;
      else {
        const Tag* tb = c.sig.get("b");
        c.tmp.resize(tb->val_len + 4);
        if (base64_decode(c.sig.val(tb), tb->val_len, c.tmp.data()) < 0) rec.canon_rc = 5;
        else {
          rec.canon_rc = 0;
          rec.canon_cand = verify_cand >= 0 ? (uint32_t)verify_cand : make_cand(hr, br, has_l, lval, 1, true);
        }
      }
    }
  }
  rec.n_steps = (uint32_t)c.tr->steps.size() - rec.first_step;
  // from_domain / key hashes (circuits.rs:16-17), one message per distinct value and thread
  ThreadCtx::PtrDom& pd = c.dom_cache[ThreadCtx::ptr_slot(em.from_domain)];
  if (pd.p == em.from_domain && pd.len == em.from_domain_len && pd.p) rec.dom_msg = pd.msg;
  else {
    std::string d(em.from_domain, em.from_domain_len);
    auto it = c.dom_msgs.find(d);
    if (it != c.dom_msgs.end()) rec.dom_msg = it->second;
    else { rec.dom_msg = c.add_msg((const uint8_t*)em.from_domain, em.from_domain_len); c.dom_msgs.emplace(std::move(d), rec.dom_msg); }
    pd.p = em.from_domain; pd.len = em.from_domain_len; pd.msg = rec.dom_msg;
  }
  auto kt = c.key_msgs.find(key_id);
  if (kt != c.key_msgs.end()) rec.key_msg = kt->second;
  else { rec.key_msg = c.add_msg(em.key, em.key_len); c.key_msgs.emplace(key_id, rec.key_msg); }
}

// Device-front-end mode: the host only resolves the key, stages distinct domains / keys for hashing and
// reserves slots; everything that reads the message bytes happens on the device (frontend.cuh).
void process_email_fe(ThreadCtx& c, const zkb_email_view& em, uint32_t local_idx, int tid, EmailRec& rec) {
  rec = EmailRec();
  rec.tid = (uint32_t)tid;
  rec.canon_rc = 0;
  if (em.raw_email_len > 0xFFFFFF00u) { rec.status = ZKB_ST_MAIL_PARSE; return; }
  int32_t key_id = -1;
  KeyMeta km = KeyMeta();
  if (em.key_type_len == 3 && memcmp(em.key_type, "rsa", 3) == 0) {
    key_id = lookup_key(c, em.key, em.key_len, km);
    if (key_id < 0) { rec.status = ZKB_ST_KEY; return; }
  } else if (em.key_type_len == 7 && memcmp(em.key_type, "ed25519", 7) == 0) {
    rec.status = em.key_len == 32 ? ZKB_ST_UNSUPPORTED : ZKB_ST_KEY;
    return;
  } else { rec.status = ZKB_ST_KEY; return; }
  ThreadCtx::PtrDom& pd = c.dom_cache[ThreadCtx::ptr_slot(em.from_domain)];
  if (pd.p == em.from_domain && pd.len == em.from_domain_len && pd.p) rec.dom_msg = pd.msg;
  else {
    std::string d(em.from_domain, em.from_domain_len);
    auto it = c.dom_msgs.find(d);
    if (it != c.dom_msgs.end()) rec.dom_msg = it->second;
    else { rec.dom_msg = c.add_msg((const uint8_t*)em.from_domain, em.from_domain_len); c.dom_msgs.emplace(std::move(d), rec.dom_msg); }
    pd.p = em.from_domain; pd.len = em.from_domain_len; pd.msg = rec.dom_msg;
  }
  auto kt = c.key_msgs.find(key_id);
  if (kt != c.key_msgs.end()) rec.key_msg = kt->second;
  else { rec.key_msg = c.add_msg(em.key, em.key_len); c.key_msgs.emplace(key_id, rec.key_msg); }
  // slots: canonical body (<= raw + 2) and header preimage (FE_PRE_CAP), both device-only
  CandRec cd;
  memset(&cd, 0, sizeof cd);
  cd.key_id = key_id; cd.algo = 1; cd.sig_state = SIG_OK; cd.rsa_list = (uint8_t)rsa_list_of(km);
  if (c.staged) {  // the only host touch of the message bytes: one copy into pinned memory (64-byte aligned, 16+ bytes of slack)
    const size_t need = ((em.raw_email_len + 16 + 63) >> 6) << 6;
    uint8_t* p = c.reserve(need, cd.raw_blk, cd.raw_local);
    if (!p) return;
    stage_copy(p, em.raw_email, em.raw_email_len);
    c.tr->blocks[cd.raw_blk].used = cd.raw_local + need;
  }
  auto virt_msg = [&](size_t cap_len, uint32_t est_len) {
    MsgRec m;
    m.goff = 0; m.len = est_len; m.blk = VIRT_BLK; m.local = c.tr->virt_used; m.canon = 0;
    c.tr->virt_used += ((cap_len >> 6) + 1) << 6;
    c.tr->msgs.push_back(m);
    const uint32_t nb = (est_len >> 6) + 1 + ((est_len & 63) >= 56 ? 1u : 0u);
    if (c.tr->hist.size() <= nb) c.tr->hist.resize((size_t)nb + 1, 0u);
    c.tr->hist[nb]++;
    c.tr->sha_blocks += nb; c.tr->sha_bytes += est_len;
    return (uint32_t)c.tr->msgs.size() - 1;
  };
  const uint32_t hdr_est = 640;   // ordering / statistics estimate; the device writes the real lengths
  cd.body_msg = virt_msg(em.raw_email_len + 2, em.raw_email_len > hdr_est + 400 ? (uint32_t)em.raw_email_len - hdr_est - 400 : 64);
  cd.hdr_msg = virt_msg(FE_PRE_CAP + 64, hdr_est);
  cd.sig_off = (uint32_t)c.tr->fe_sig_words;
  c.tr->fe_sig_words += km.limbs_class;
  c.tr->rsa_cnt[cd.rsa_list]++;
  rec.fe_cand = (uint32_t)c.tr->cands.size();
  rec.canon_cand = rec.fe_cand;
  c.tr->cands.push_back(cd);
  c.tr->fe_emails.push_back(local_idx);
  (void)km;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline double now_s2() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// Host pack of emails [e0, e0+ne): parallel parse + layout of the SoA meta buffers.
// ctxs: one ThreadCtx per pool thread, alive for the whole API call (key lookups are cached across chunks)
int pack_chunk(zkb_engine* e, const zkb_email_view* emails, size_t e0, size_t ne, const zkb_regex_set* rs, Chunk& ch, PinBuf& pin_meta,
               std::vector<ThreadCtx>& ctxs, bool allow_fe = false, size_t max_span = (size_t)3 << 30) {
  const int T = e->pool->size();
  ch.e0 = e0; ch.ne = ne; ch.views = emails + e0;
  ch.emails.assign(ne, EmailRec());
  ch.tr.resize(T);
  for (auto& t : ch.tr) t.clear();
  const bool want_regex = rs != nullptr;
  std::atomic<size_t> next{0};
  std::atomic<int> oom{0};
  const double tp0 = now_s2();
  // direct mode: every message of the chunk lies in one registered host range and the chunk's span is
  // not much larger than its payload -> one DMA of the raw span, bodies canonicalised on the device
  ch.direct = false; ch.span_host = nullptr; ch.span_bytes = 0; ch.n_canon = 0;
  if (ne && !e->registered.empty() && !e->has(ZKB_OPT_NO_DIRECT)) {
    const uint8_t* lo = emails[e0].raw_email;
    const uint8_t* hi = lo;
    size_t payload = 0;
    for (size_t i = 0; i < ne; i++) {
      const zkb_email_view& v = emails[e0 + i];
      lo = std::min(lo, v.raw_email); hi = std::max(hi, v.raw_email + v.raw_email_len);
      payload += v.raw_email_len;
    }
    for (auto& r : e->registered) {
      if (lo >= r.first && hi <= r.first + r.second && (size_t)(hi - lo) <= payload + payload / 2 + (1u << 20) &&
          (size_t)(hi - lo) < max_span) {
        const uint8_t* base = (const uint8_t*)((uintptr_t)lo & ~(uintptr_t)255);
        if (base < r.first) base = r.first;
        ch.direct = true; ch.span_host = base; ch.span_bytes = (size_t)(hi - base);
        break;
      }
    }
  }
  ch.fe = allow_fe && !e->has(ZKB_OPT_NO_DEVICE_FRONTEND);
  if (ch.fe && !ch.direct) {  // staged: arena offsets of the staged domains are 32-bit, so keep giant messages on the host path
    if (e->has(ZKB_OPT_NO_STAGED_FRONTEND)) ch.fe = false;
    for (size_t i = 0; i < ne && ch.fe; i++) if (emails[e0 + i].raw_email_len > ((size_t)64 << 20)) ch.fe = false;
  }
  ch.staged = ch.fe && !ch.direct;
  e->prof_prelude += now_s2() - tp0;
  const size_t grain = std::max<size_t>(16, std::min<size_t>(512, ne / (size_t)(T * 8) + 1));
  e->pool->run([&](int tid) {
    const double tb0 = now_s2();
    ThreadCtx& c = ctxs[tid];
    c.eng = e; c.tr = &ch.tr[tid];
    c.direct = ch.direct; c.span_host = ch.span_host; c.staged = ch.staged;
    c.oom = false;
    c.dom_msgs.clear(); c.key_msgs.clear();          // message indices are per chunk
    for (auto& dslot : c.dom_cache) dslot.p = nullptr;
    for (;;) {
      size_t lo = next.fetch_add(grain);
      if (lo >= ne) break;
      size_t hi = std::min(ne, lo + grain);
      for (size_t i = lo; i < hi; i++) {
        if (ch.staged && i + 1 < hi) {   // the next message is a cold heap allocation: start pulling its first lines in
          const uint8_t* nx = emails[e0 + i + 1].raw_email;
          const size_t nl = std::min<size_t>(emails[e0 + i + 1].raw_email_len, 512);
          for (size_t o = 0; o < nl; o += 64) __builtin_prefetch(nx + o, 0, 0);
        }
        if (ch.fe) process_email_fe(c, emails[e0 + i], (uint32_t)i, tid, ch.emails[i]);
        else process_email(c, emails[e0 + i], want_regex, tid, ch.emails[i]);
        if (c.oom) { oom = 1; return; }
      }
    }
    if (ch.staged) stage_fence();
    e->prof_busy_ns.fetch_add((uint64_t)((now_s2() - tb0) * 1e9));
  });
  if (oom) return ZKB_E_NOMEM;
  const double tp1 = now_s2();
  e->prof_parse += tp1 - tp0;
  // layout: device arena = concatenation of the used parts of all staging blocks
  uint64_t off = 0;
  uint32_t M = 0, C = 0, NC = 0;
  size_t sig_words = 0;
  for (auto& t : ch.tr) {
    for (auto& b : t.blocks) { b.dev_off = off; off += align_up(b.used, 128); }
    t.msg_base = M; t.cand_base = C; t.canon_base = NC;
    M += (uint32_t)t.msgs.size(); C += (uint32_t)t.cands.size(); NC += (uint32_t)t.canon.size();
    sig_words += t.sigw.size() + t.fe_sig_words;
  }
  const uint64_t staged_bytes = off;   // what travels host -> device out of the staging blocks
  for (auto& t : ch.tr) { t.virt_base = off; off += align_up(t.virt_used, 128); }  // device-only canonical body slots
  ch.n_canon = NC;
  ch.arena_bytes = (size_t)off + 128;
  ch.M = M; ch.C = C;
  // RSA lists and message-order buckets from the counters the threads kept while parsing
  uint32_t rn[6] = {0, 0, 0, 0, 0, 0};
  std::vector<uint32_t> rsa_start((size_t)T * 6, 0);
  for (int t = 0; t < T; t++)
    for (int k = 0; k < 6; k++) { rsa_start[(size_t)t * 6 + k] = rn[k]; rn[k] += ch.tr[t].rsa_cnt[k]; }
  uint32_t maxb = 0;
  uint64_t sha_blocks = 0, sha_bytes = 0;
  for (auto& t : ch.tr) { maxb = std::max<uint32_t>(maxb, t.hist.empty() ? 0u : (uint32_t)t.hist.size() - 1); sha_blocks += t.sha_blocks; sha_bytes += t.sha_bytes; }
  // order: descending block count; inside a bucket by thread, then by message index
  std::vector<std::vector<uint32_t>> ord_start(T);
  {
    for (int t = 0; t < T; t++) ord_start[t].assign(ch.tr[t].hist.size(), 0u);
    uint32_t acc = 0;
    for (uint32_t b = maxb + 1; b-- > 0;)
      for (int t = 0; t < T; t++)
        if (b < ch.tr[t].hist.size()) { ord_start[t][b] = acc; acc += ch.tr[t].hist[b]; }
  }
  const size_t P = rs ? rs->n_active() : 0;
  uint32_t n_dfa = 0;
  if (P) for (auto& er : ch.emails) if (er.status == ZKB_ST_OK && er.canon_rc == 0) n_dfa++;
  ch.n_dfa = n_dfa;
  // meta offsets.  In device-front-end mode the arrays the DEVICE writes (bh= words, signature limbs, canon
  // items) go last and are not uploaded.
  if (ch.fe) NC = C;   // one canon item per candidate, written by the device
  ch.n_canon = NC;
  size_t o = 0;
  ch.o_msg_off = o; o += align_up((size_t)M * 8, 16);
  ch.o_msg_len = o; o += align_up((size_t)M * 4, 16);
  ch.o_order = o; o += align_up((size_t)M * 4, 16);
  ch.o_msg_canon = o; o += align_up((size_t)M * 4, 16);
  ch.o_cand_body = o; o += align_up((size_t)C * 4, 16);
  for (int k = 0; k < 6; k++) { ch.o_rsa[k] = o; ch.rsa_n[k] = rn[k]; o += align_up((size_t)rn[k] * sizeof(RsaItem), 16); }
  ch.o_dfa = o; o += align_up((size_t)n_dfa * 2 * sizeof(DfaItem), 16);  // header haystack + body haystack per email
  ch.o_fein = o; o += align_up(ch.fe ? (size_t)C * sizeof(FeIn) : 0, 16);
  ch.o_asm_in = o; o += align_up(ch.fe ? 0 : ne * sizeof(FeIn), 16);
  ch.o_asm_fo = o; o += align_up(ch.fe ? 0 : ne * sizeof(FeOut), 16);
  if (ch.fe) ch.upload_bytes = o;
  ch.o_cand_bh = o; o += align_up((size_t)C * 32, 16);
  ch.o_sig = o; o += align_up(sig_words * 4, 16);
  ch.o_canon = o; o += align_up((size_t)NC * sizeof(CanonItem), 16);
  ch.meta_bytes = o + 16;
  if (!ch.fe) ch.upload_bytes = ch.meta_bytes;
  if (!pin_meta.ensure(ch.meta_bytes)) return ZKB_E_NOMEM;
  uint8_t* mh = pin_meta.p;  // every array below is written in full; padding bytes are never read
  uint64_t* msg_off = (uint64_t*)(mh + ch.o_msg_off);
  uint32_t* msg_len = (uint32_t*)(mh + ch.o_msg_len);
  uint32_t* order = (uint32_t*)(mh + ch.o_order);
  uint32_t* cand_body = (uint32_t*)(mh + ch.o_cand_body);
  uint32_t* cand_bh = (uint32_t*)(mh + ch.o_cand_bh);
  uint32_t* sigw = (uint32_t*)(mh + ch.o_sig);
  std::vector<size_t> sig_base(T, 0);
  { size_t sb = 0; for (int t = 0; t < T; t++) { sig_base[t] = sb; sb += ch.tr[t].sigw.size() + ch.tr[t].fe_sig_words; } }
  // one parallel pass: every thread lays out the records it produced
  e->pool->run([&](int tid) {
    ThreadRecs& t = ch.tr[tid];
    std::vector<uint32_t>& os = ord_start[tid];
    for (size_t i = 0; i < t.msgs.size(); i++) {
      MsgRec& m = t.msgs[i];
      m.goff = m.blk == VIRT_BLK ? t.virt_base + m.local : t.blocks[m.blk].dev_off + m.local;
      msg_off[t.msg_base + i] = m.goff;
      msg_len[t.msg_base + i] = m.len;
      const uint32_t nb = (m.len >> 6) + 1 + ((m.len & 63) >= 56 ? 1u : 0u);
      order[os[nb]++] = t.msg_base + (uint32_t)i;
    }
    uint32_t* msg_canon = (uint32_t*)(mh + ch.o_msg_canon);
    for (size_t i = 0; i < t.msgs.size(); i++) msg_canon[t.msg_base + i] = 0xFFFFFFFFu;
    uint32_t fill[6];
    for (int k = 0; k < 6; k++) fill[k] = rsa_start[(size_t)tid * 6 + k];
    for (size_t i = 0; i < t.cands.size(); i++) {
      const CandRec& cd = t.cands[i];
      cand_body[t.cand_base + i] = t.msg_base + cd.body_msg;
      if (ch.fe) msg_canon[t.msg_base + cd.body_msg] = t.cand_base + (uint32_t)i;   // the front end writes canon item [candidate]
      if (!ch.fe) memcpy(cand_bh + (size_t)(t.cand_base + i) * 8, cd.bh, 32);
      if (cd.haystack_only || cd.sig_state != SIG_OK || cd.algo != 1) continue;
      RsaItem it;
      it.sig_off = (uint32_t)(sig_base[tid] + cd.sig_off);
      it.key_id = (uint32_t)cd.key_id;
      it.digest_slot = t.msg_base + cd.hdr_msg;
      it.cand = t.cand_base + (uint32_t)i;
      ((RsaItem*)(mh + ch.o_rsa[cd.rsa_list]))[fill[cd.rsa_list]++] = it;
    }
    if (!t.sigw.empty()) memcpy(sigw + sig_base[tid], t.sigw.data(), t.sigw.size() * 4);
    CanonItem* ci = (CanonItem*)(mh + ch.o_canon) + t.canon_base;
    for (size_t i = 0; i < t.canon.size(); i++) { ci[i] = t.canon[i]; ci[i].msg += t.msg_base; msg_canon[ci[i].msg] = t.canon_base + (uint32_t)i; }
    if (ch.fe) {
      FeIn* fin = (FeIn*)(mh + ch.o_fein);
      for (size_t j = 0; j < t.fe_emails.size(); j++) {
        const zkb_email_view& em = emails[e0 + t.fe_emails[j]];
        const EmailRec& er = ch.emails[t.fe_emails[j]];
        const CandRec& cd = t.cands[j];
        const KeyMeta& km = e->key_meta[cd.key_id];
        FeIn fi;
        memset(&fi, 0, sizeof fi);
        fi.raw_off = ch.staged ? t.blocks[cd.raw_blk].dev_off + cd.raw_local : (uint64_t)(em.raw_email - ch.span_host);
        fi.raw_len = (uint32_t)em.raw_email_len;
        fi.dom_off = (uint32_t)t.msgs[er.dom_msg].goff;   // staged domains sit at the front of the arena (< 4 GiB)
        fi.dom_len = (uint32_t)em.from_domain_len;
        fi.k = km.k; fi.limbs = km.limbs_class;
        fi.sig_word_off = (uint32_t)(sig_base[tid] + cd.sig_off);
        fi.body_msg = t.msg_base + cd.body_msg; fi.pre_msg = t.msg_base + cd.hdr_msg;
        fi.cand = t.cand_base + (uint32_t)j;
        fi.email = t.fe_emails[j];
        fi.dom_msg = t.msg_base + er.dom_msg; fi.key_msg = t.msg_base + er.key_msg;
        fin[t.cand_base + j] = fi;
      }
    }
  });
  // DFA items: for each email with haystacks, slot 2*j = header preimage, 2*j+1 = canonical body (j = rank of the
  // email among those with haystacks: counted per grain, prefix-summed, filled in parallel)
  uint64_t dfa_bytes = 0;
  if (P) {
    DfaItem* items = (DfaItem*)(mh + ch.o_dfa);
    const size_t G = 4096, ng = (ne + G - 1) / G;
    std::vector<uint32_t> gstart(ng + 1, 0);
    e->pool->parallel_for(ng, 1, [&](size_t lo, size_t hi, int) {
      for (size_t g = lo; g < hi; g++) {
        uint32_t c = 0;
        for (size_t i = g * G; i < std::min(ne, (g + 1) * G); i++) c += (ch.emails[i].status == ZKB_ST_OK && ch.emails[i].canon_rc == 0) ? 1u : 0u;
        gstart[g + 1] = c;
      }
    });
    for (size_t g = 0; g < ng; g++) gstart[g + 1] += gstart[g];
    std::atomic<uint64_t> bytes_acc{0};
    const uint64_t nh_parts = rs->header_present ? rs->n_header : 0, nb_parts = rs->body_present ? rs->n_body : 0;
    e->pool->parallel_for(ng, 1, [&](size_t lo, size_t hi, int) {
      uint64_t local = 0;
      for (size_t g = lo; g < hi; g++) {
        uint32_t j = gstart[g];
        for (size_t i = g * G; i < std::min(ne, (g + 1) * G); i++) {
          EmailRec& er = ch.emails[i];
          if (er.status != ZKB_ST_OK || er.canon_rc != 0) continue;
          ThreadRecs& t = ch.tr[er.tid];
          const CandRec& cd = t.cands[er.canon_cand];
          const MsgRec& hm = t.msgs[cd.hdr_msg];
          const MsgRec& bm = t.msgs[cd.body_msg];
          items[2 * j].hay_off = hm.goff; items[2 * j].msg = t.msg_base + cd.hdr_msg; items[2 * j].out_slot = (uint32_t)i;
          items[2 * j + 1].hay_off = bm.goff; items[2 * j + 1].msg = t.msg_base + cd.body_msg; items[2 * j + 1].out_slot = (uint32_t)i;
          local += (uint64_t)hm.len * nh_parts + (uint64_t)bm.len * nb_parts;
          j++;
        }
      }
      bytes_acc.fetch_add(local);
    });
    dfa_bytes = bytes_acc.load();
  }
  if (!ch.fe && ne) {
    // host front end: per-email inputs of the record kernel.  An email decided by exactly one signature candidate
    // (the common case) gets its record on the device; everything else is marked for the host's resolve_email.
    FeIn* ain = (FeIn*)(mh + ch.o_asm_in);
    FeOut* afo = (FeOut*)(mh + ch.o_asm_fo);
    e->pool->parallel_for(ne, 4096, [&](size_t lo, size_t hi, int) {
      for (size_t i = lo; i < hi; i++) {
        const EmailRec& er = ch.emails[i];
        FeIn fi; FeOut fo;
        memset(&fi, 0, sizeof fi); memset(&fo, 0, sizeof fo);
        fi.email = (uint32_t)i;
        fo.flags = FE_FALLBACK;
        if (er.status == ZKB_ST_OK && er.n_steps == 1 && (!want_regex || er.canon_rc == 0)) {
          const ThreadRecs& t = ch.tr[er.tid];
          const StepRec& st = t.steps[er.first_step];
          if (st.kind == STEP_CAND && t.cands[st.cand].algo == 1) {
            const CandRec& cd = t.cands[st.cand];
            fi.cand = t.cand_base + st.cand;
            fi.body_msg = t.msg_base + cd.body_msg; fi.pre_msg = t.msg_base + cd.hdr_msg;
            fi.dom_msg = t.msg_base + er.dom_msg; fi.key_msg = t.msg_base + er.key_msg;
            fo.flags = (cd.bh_valid ? FE_BH_VALID : 0u) | (cd.sig_state == SIG_SYNTAX ? FE_SIG_SYNTAX : 0u) |
                       (cd.sig_state == SIG_BADLEN ? FE_SIG_BADLEN : 0u);
          }
        }
        ain[i] = fi; afo[i] = fo;
      }
    });
  }
  memset(&ch.st, 0, sizeof ch.st);
  ch.st.n_emails = ne; ch.st.n_candidates = C; ch.st.n_sha_messages = M;
  ch.st.sha_blocks = sha_blocks; ch.st.sha_bytes = sha_bytes;
  ch.st.rsa_items_1024 = rn[0] + rn[1]; ch.st.rsa_items_2048 = rn[2] + rn[3]; ch.st.rsa_items_other = rn[4] + rn[5];
  {
    uint64_t macs = 0;
    for (int k = 0; k < 6; k++) { uint64_t l = k < 2 ? 32 : k < 4 ? 64 : 128; macs += (uint64_t)rn[k] * 18ull * (2 * l * l + l); }
    ch.st.rsa_macs = macs;
  }
  ch.st.dfa_items = (uint64_t)n_dfa * P; ch.st.dfa_bytes = dfa_bytes;
  ch.st.arena_bytes = ch.arena_bytes;
  ch.st.h2d_bytes = staged_bytes + ch.upload_bytes + (ch.direct ? ch.span_bytes : 0);
  e->prof_layout += now_s2() - tp1;
  return ZKB_OK;
}

inline uint32_t rec_words_for(size_t P) { return (uint32_t)(36 + 4 * P); }   // 144-byte head + one uint4 per part
size_t out_layout(uint32_t M, uint32_t C, size_t ne, size_t P, size_t& o_flags, size_t& o_dfa, size_t n_fe = 0, size_t* o_fe = nullptr,
                  size_t* o_recs = nullptr) {
  size_t o = align_up((size_t)M * 32, 16);
  o_flags = o; o += align_up((size_t)C * 4, 16);
  o_dfa = o; o += ne * P * 16;
  if (o_fe) *o_fe = o;
  o += n_fe * sizeof(FeOut);
  if (o_recs) *o_recs = o;
  o += ne * (size_t)rec_words_for(P) * 4;   // one result record per email
  return o + 16;
}

// Device buffers + H2D of one packed chunk on `stream`.
int upload_chunk(zkb_engine* e, Chunk& ch, const zkb_regex_set* rs, DeviceChunk& d, PinBuf& pin_meta, cudaStream_t stream) {
  const size_t P = rs ? rs->n_active() : 0;
  if (!d.arena.ensure(ch.arena_bytes) || !d.meta.ensure(ch.meta_bytes)) return ZKB_E_NOMEM;
  size_t o_flags, o_dfa, o_fe = 0, o_recs = 0;
  d.out_bytes = out_layout(ch.M, ch.C, ch.ne, P, o_flags, o_dfa, ch.fe ? ch.C : 0, &o_fe, &o_recs);
  if (!d.out.ensure(d.out_bytes)) return ZKB_E_NOMEM;
  d.recs = nullptr; d.o_recs = o_recs; d.rec_words = rec_words_for(P);
  d.asm_in = nullptr; d.asm_fo = nullptr; d.n_asm = 0;
  for (auto& t : ch.tr)
    for (auto& b : t.blocks)
      if (b.used) CK(cudaMemcpyAsync(d.arena.p + b.dev_off, b.p, b.used, cudaMemcpyHostToDevice, stream));
  CK(cudaMemcpyAsync(d.meta.p, pin_meta.p, ch.upload_bytes, cudaMemcpyHostToDevice, stream));
  d.n_canon = 0; d.n_fe = 0; d.raw_base = nullptr;
  if (ch.direct && ch.n_canon) {
    if (!d.span.ensure(ch.span_bytes + 256)) return ZKB_E_NOMEM;
    CK(cudaMemcpyAsync(d.span.p, ch.span_host, ch.span_bytes, cudaMemcpyHostToDevice, stream));  // DMA from registered memory
    d.raw_base = d.span.p;
  } else if (ch.staged) {
    d.raw_base = d.arena.p;   // raw messages travelled inside the staging blocks
  }
  if (d.raw_base && ch.n_canon) {
    d.canon_items = (const CanonItem*)(d.meta.p + ch.o_canon);
    d.n_canon = ch.n_canon;
    d.msg_canon = (const uint32_t*)(d.meta.p + ch.o_msg_canon);
  }
  d.msg_len_rw = (uint32_t*)(d.meta.p + ch.o_msg_len);
  if (ch.fe && ch.C) {
    d.fe_in = (const FeIn*)(d.meta.p + ch.o_fein);
    d.fe_out = (FeOut*)(d.out.p + o_fe);
    d.n_fe = ch.C;
    d.canon_rw = (CanonItem*)(d.meta.p + ch.o_canon);
    d.cand_bh_rw = (uint32_t*)(d.meta.p + ch.o_cand_bh);
    d.sig_rw = (uint32_t*)(d.meta.p + ch.o_sig);
    d.asm_in = d.fe_in; d.asm_fo = d.fe_out; d.n_asm = d.n_fe;
  } else {
    d.asm_in = (const FeIn*)(d.meta.p + ch.o_asm_in); d.asm_fo = (const FeOut*)(d.meta.p + ch.o_asm_fo); d.n_asm = (uint32_t)ch.ne;
  }
  d.recs = (uint32_t*)(d.out.p + o_recs);
  CK(cudaMemsetAsync(d.out.p + o_flags, 0, align_up((size_t)ch.C * 4, 16), stream));
  if (P) CK(cudaMemsetAsync(d.out.p + o_dfa, 0, ch.ne * P * 16, stream));
  d.msg_off = (const uint64_t*)(d.meta.p + ch.o_msg_off);
  d.msg_len = (const uint32_t*)(d.meta.p + ch.o_msg_len);
  d.order = (const uint32_t*)(d.meta.p + ch.o_order);
  d.cand_body = (const uint32_t*)(d.meta.p + ch.o_cand_body);
  d.cand_bh = (const uint32_t*)(d.meta.p + ch.o_cand_bh);
  d.sig_arena = (const uint32_t*)(d.meta.p + ch.o_sig);
  for (int k = 0; k < 6; k++) { d.rsa_items[k] = (const RsaItem*)(d.meta.p + ch.o_rsa[k]); d.rsa_n[k] = ch.rsa_n[k]; }
  d.dfa_items = (const DfaItem*)(d.meta.p + ch.o_dfa); d.n_dfa = ch.n_dfa;
  d.digests = (uint32_t*)d.out.p;
  d.cand_flags = (uint32_t*)(d.out.p + o_flags);
  d.dfa_out = (uint4*)(d.out.p + o_dfa);
  d.M = ch.M; d.C = ch.C; d.NE = (uint32_t)ch.ne; d.P = (uint32_t)P;
  return ZKB_OK;
}

// The key table is append-only WITHIN a call (key ids live in the call's device records) and bounded ACROSS calls:
// once it holds more than kKeySoftCap keys and no resident batch refers to it, the next call starts from an empty
// table.  A service fed ever-new keys therefore holds at most max(kKeySoftCap, distinct keys of one call) entries
// (1056 bytes each on the host and in HBM) instead of growing for the engine's lifetime.
constexpr size_t kKeySoftCap = 16384;
void trim_keytab(zkb_engine* e) {
  std::lock_guard<std::mutex> l(e->key_mu);
  if (e->key_meta.size() <= kKeySoftCap || e->live_batches.load() != 0) return;
  e->key_index.clear(); e->keytab_host.clear(); e->key_meta.clear();
  e->d_keytab_n = 0;   // device rows are rewritten as the new ids are assigned (calls are serialised by run_mu)
}

int sync_keytab(zkb_engine* e, cudaStream_t stream) {
  std::lock_guard<std::mutex> l(e->key_mu);
  size_t nkeys = e->key_meta.size();
  if (nkeys == e->d_keytab_n) return ZKB_OK;
  if (nkeys > e->d_keytab_cap) {
    CK(cudaStreamSynchronize(stream));
    for (auto& s : e->slots) if (s.stream) CK(cudaStreamSynchronize(s.stream));
    if (e->d_keytab) cudaFree(e->d_keytab);
    size_t cap = std::max<size_t>(1024, nkeys * 2);
    CK(cudaMalloc((void**)&e->d_keytab, cap * ZKB_KEY_STRIDE * 4));
    e->d_keytab_cap = cap;
    e->d_keytab_n = 0;
  }
  // the table is small; key ids are append-only, so re-upload the new tail synchronously
  CK(cudaMemcpyAsync(e->d_keytab + e->d_keytab_n * ZKB_KEY_STRIDE, e->keytab_host.data() + e->d_keytab_n * ZKB_KEY_STRIDE,
                     (nkeys - e->d_keytab_n) * ZKB_KEY_STRIDE * 4, cudaMemcpyHostToDevice, stream));
  CK(cudaStreamSynchronize(stream));
  e->d_keytab_n = nkeys;
  return ZKB_OK;
}

// Enqueues every kernel of one chunk.  ev (optional): 5 events recorded around the kernel families.
// phase: 0 = everything in order; 1 = all but the RSA launches (front end, canonicalisation, SHA-256, bh= check,
// DFA scans); 2 = the RSA launches only.  Phases 1 / 2 let zkb_batch_run_async put the hashing of chunk k+1 next to
// the modular exponentiation of chunk k (different pipes: ALU vs FMA-heavy).
int launch_chunk(zkb_engine* e, const DeviceChunk& d, const zkb_regex_set* rs, cudaStream_t s, cudaEvent_t* ev, uint64_t* launches,
                 int phase = 0) {
  uint64_t nl = 0;
  const bool pre = phase != 2, rsa = phase != 1;
  if (pre && d.n_fe) {
    // without regex parts, signature headers of other domains in front of the candidate may be skipped on the device
    launch_frontend(d.raw_base, d.fe_in, d.n_fe, d.arena.p, d.msg_off, d.msg_len_rw, d.sig_rw, d.cand_bh_rw, d.canon_rw, d.fe_out, rs == nullptr,
                    (long long)(e->now_unix ? e->now_unix : (int64_t)time(nullptr)), s);
    nl++;
  }
  if (pre && d.n_canon) { launch_canon_body(d.raw_base, d.canon_items, d.n_canon, d.arena.p, d.msg_off, d.msg_len_rw, d.order, d.msg_canon, d.M, s); nl++; }
  if (ev) CK(cudaEventRecord(ev[0], s));
  if (pre && d.M) { launch_sha256(d.arena.p, d.msg_off, d.msg_len, d.order, d.M, d.digests, s); nl++; }
  if (ev) CK(cudaEventRecord(ev[1], s));
  if (pre && d.C) { launch_bh_check(d.digests, d.cand_body, d.cand_bh, d.C, d.cand_flags, s); nl++; }
  if (ev) CK(cudaEventRecord(ev[2], s));
  const int lanes = (int)e->rsa_lanes;
  for (int k = 0; k < 6; k++) {
    if (!rsa || !d.rsa_n[k]) continue;
    const bool generic = (k & 1) != 0;
    switch (k >> 1) {
      case 0: launch_rsa32(generic, std::max(2, lanes / 2), d.sig_arena, d.rsa_items[k], d.rsa_n[k], e->d_keytab, d.digests, d.cand_flags, s); break;
      case 1: launch_rsa64(generic, lanes, d.sig_arena, d.rsa_items[k], d.rsa_n[k], e->d_keytab, d.digests, d.cand_flags, s, !e->has(ZKB_OPT_NO_SQR)); break;
      default: launch_rsa128(generic, std::min(16, lanes * 2), d.sig_arena, d.rsa_items[k], d.rsa_n[k], e->d_keytab, d.digests, d.cand_flags, s); break;
    }
    nl++;
  }
  if (ev) CK(cudaEventRecord(ev[3], s));
  if (pre && rs && d.n_dfa) {
    size_t pi = 0;
    for (size_t p = 0; p < rs->parts.size(); p++) {
      const zkb_regex_set::Part& part = rs->parts[p];
      if (part.body ? !rs->body_present : !rs->header_present) continue;
      // items are interleaved (header, body) per email; out slot = email * P + pi
      launch_dfa_strided(part.elem, part.direct, d.arena.p, d.dfa_items, d.n_dfa, d.msg_len, part.body ? 1u : 0u, (uint32_t)d.P, (uint32_t)pi, part.d_fwd,
                         part.fwd_bytes, part.d_rev, part.rev_bytes, e->smem_optin, part.body ? 1 : 0, d.dfa_out, s);
      nl++; pi++;
    }
  }
  if (ev) CK(cudaEventRecord(ev[4], s));
  if (rsa && d.n_asm && d.recs) {
    // result records: needs the RSA flags and (phase 2 runs after the pre phase of the same chunk) the DFA results
    uint32_t body_mask = 0;
    if (rs) {
      uint32_t pi = 0;
      for (size_t p = 0; p < rs->parts.size(); p++) {
        if (rs->parts[p].body ? !rs->body_present : !rs->header_present) continue;
        if (rs->parts[p].body) body_mask |= 1u << pi;
        pi++;
      }
    }
    launch_assemble(d.asm_in, d.asm_fo, d.n_asm, d.cand_flags, d.digests, d.dfa_out, d.P, body_mask, rs != nullptr, d.recs, d.rec_words, s);
    nl++;
  }
  CK(cudaGetLastError());
  if (launches) *launches += nl;
  return ZKB_OK;
}

// ------------------------------------------------------------------ result resolution (host)
struct HayView { const uint8_t* p; uint32_t n; };

// cleaned [start,end) of a haystack (soft breaks removed, zero padded) for the capture check
void cleaned_span(const HayView& h, bool qp, uint32_t start, uint32_t end, std::string& out) {
  out.clear();
  if (!qp) { if (end <= h.n && start <= end) out.assign((const char*)h.p + start, end - start); return; }
  uint32_t c = 0, o = 0;
  while (c < end) {
    while (o + 2 < h.n && h.p[o] == '=' && h.p[o + 1] == '\r' && h.p[o + 2] == '\n') o += 3;
    char b = 0;
    if (o < h.n) b = (char)h.p[o++];
    if (c >= start) out.push_back(b);
    c++;
  }
}

// Host bytes of a haystack for the capture check.  Bodies canonicalised on the device (direct mode) are
// re-canonicalised here on demand from the caller's raw bytes: only emails with captures to check pay.
HayView hay_view(const Chunk& ch, const ThreadRecs& t, const MsgRec& m, std::vector<uint8_t>& scratch) {
  HayView v;
  if (m.blk != VIRT_BLK) { v.p = t.blocks[m.blk].p + m.local; v.n = m.len; return v; }
  const CanonItem& it = t.canon[m.canon];
  const uint8_t* raw = ch.span_host + it.raw_off;
  scratch.resize((size_t)it.raw_len + 64);
  size_t cl = (it.flags & 1u) ? canon_body_relaxed(raw, it.raw_len, scratch.data()) : canon_body_simple(raw, it.raw_len, scratch.data());
  if ((it.flags & 2u) && it.l < cl) cl = it.l;
  v.p = scratch.data(); v.n = (uint32_t)cl;
  return v;
}

void resolve_email(const zkb_engine* e, const Chunk& ch, size_t i, const uint8_t* outp, size_t o_flags, size_t o_dfa,
                   const zkb_regex_set* rs, const zkb_email_captures* caps, std::vector<uint8_t>& scratch,
                   zkb_result& res, std::string& s1, std::string& s2) {
  memset(&res, 0, rs ? sizeof res : offsetof(zkb_result, parts));  // parts[] stay untouched when n_parts = 0
  res.dkim_detail = ZKB_DKIM_NEUTRAL;
  const EmailRec& er = ch.emails[i];
  if (er.status != ZKB_ST_OK) { res.status = er.status; return; }
  const ThreadRecs& t = ch.tr[er.tid];
  const uint32_t* digests = (const uint32_t*)outp;
  const uint32_t* flags = (const uint32_t*)(outp + o_flags);
  auto put_digest = [&](uint8_t* dst, uint32_t gmsg) {
    const uint32_t* w = digests + (size_t)gmsg * 8;
    uint32_t be[8];
    for (int k = 0; k < 8; k++) be[k] = __builtin_bswap32(w[k]);   // state words -> digest bytes
    memcpy(dst, be, 32);
  };
  bool have_err = false, saw_sha1 = false, pass = false;
  int last_err = ZKB_DKIM_NEUTRAL;
  for (uint32_t k = 0; k < er.n_steps && !pass; k++) {
    const StepRec& st = t.steps[er.first_step + k];
    if (st.kind == STEP_ERR) { have_err = true; last_err = st.detail; continue; }
    if (st.kind == STEP_SHA1) { saw_sha1 = true; continue; }
    const CandRec& cd = t.cands[st.cand];
    const uint32_t f = flags[t.cand_base + st.cand];
    put_digest(res.body_hash, t.msg_base + cd.body_msg);
    put_digest(res.header_hash, t.msg_base + cd.hdr_msg);
    res.bh_ok = (cd.bh_valid && (f & ZKB_F_BH_OK)) ? 1 : 0;
    res.rsa_ok = 0;
    if (!res.bh_ok) { have_err = true; last_err = ZKB_DKIM_BODY_HASH; continue; }
    if (cd.sig_state == SIG_SYNTAX) { have_err = true; last_err = ZKB_DKIM_SIG_SYNTAX; continue; }
    if (cd.algo == 2) { have_err = true; last_err = ZKB_DKIM_ALGO_KEY_MISMATCH; continue; }
    if (cd.sig_state == SIG_OK && (f & ZKB_F_RSA_OK)) { res.rsa_ok = 1; pass = true; break; }
    have_err = true; last_err = ZKB_DKIM_SIG_MISMATCH;
  }
  if (!pass) {
    res.dkim_detail = have_err ? last_err : ZKB_DKIM_NEUTRAL;
    res.status = saw_sha1 ? ZKB_ST_UNSUPPORTED : ZKB_ST_DKIM_FAIL;
    return;
  }
  res.dkim_detail = ZKB_DKIM_PASS;
  put_digest(res.from_domain_hash, t.msg_base + er.dom_msg);
  put_digest(res.public_key_hash, t.msg_base + er.key_msg);
  if (!rs) return;
  if (er.canon_rc != 0) { res.status = ZKB_ST_CANONICALIZE; return; }
  const size_t P = rs->n_active();
  const uint4* dfa = (const uint4*)(outp + o_dfa) + i * P;
  const CandRec& cc = t.cands[er.canon_cand];
  size_t pi = 0;
  for (size_t p = 0; p < rs->parts.size(); p++) {
    const bool body = rs->parts[p].body;
    if (body ? !rs->body_present : !rs->header_present) continue;
    const uint4 r = dfa[pi++];
    uint32_t slot = res.n_parts;
    if (slot < ZKB_MAX_PARTS) {
      res.parts[slot].match_count = r.x; res.parts[slot].start = r.y; res.parts[slot].end = r.z; res.parts[slot].captures_ok = 0;
      res.n_parts++;
    }
    bool ok = r.x == 1;
    if (ok && caps) {
      const zkb_email_captures& ec = caps[ch.e0 + i];
      bool loaded = false;
      for (size_t q = 0; q < ec.n_caps && ok; q++) {
        if (ec.caps[q].part != p) continue;
        if (!loaded) {
          HayView hv = hay_view(ch, t, t.msgs[body ? cc.body_msg : cc.hdr_msg], scratch);
          cleaned_span(hv, body, r.y, r.z, s1);
          bool ascii = true;
          for (char ch2 : s1) if (ch2 & 0x80) { ascii = false; break; }
          if (!ascii) { utf8_lossy((const uint8_t*)s1.data(), s1.size(), s2); s1.swap(s2); }
          loaded = true;
        }
        if (ec.caps[q].len && s1.find(ec.caps[q].s, 0, ec.caps[q].len) == std::string::npos) ok = false;
      }
    }
    if (ok && slot < ZKB_MAX_PARTS) res.parts[slot].captures_ok = 1;
    if (!ok) { res.status = body ? ZKB_ST_REGEX_BODY : ZKB_ST_REGEX_HEADER; return; }
  }
}

// Device-front-end mode.  Returns false when the device declined the message (the caller re-runs it
// through the host front end); otherwise fills the record exactly as resolve_email would.
bool resolve_email_fe(const zkb_engine* e, const Chunk& ch, size_t i, const uint8_t* outp, size_t o_flags, size_t o_dfa, size_t o_fe,
                      const zkb_regex_set* rs, const zkb_email_captures* caps, std::vector<uint8_t>& scratch,
                      zkb_result& res, std::string& s1, std::string& s2) {
  memset(&res, 0, rs ? sizeof res : offsetof(zkb_result, parts));
  res.dkim_detail = ZKB_DKIM_NEUTRAL;
  const EmailRec& er = ch.emails[i];
  if (er.status != ZKB_ST_OK) { res.status = er.status; return true; }
  const ThreadRecs& t = ch.tr[er.tid];
  const uint32_t gc = t.cand_base + er.fe_cand;
  const FeOut& fo = ((const FeOut*)(outp + o_fe))[gc];
  if (fo.flags & FE_MAIL_PARSE) { res.status = ZKB_ST_MAIL_PARSE; return true; }
  if (fo.flags & FE_FALLBACK) return false;
  const CandRec& cd = t.cands[er.fe_cand];
  const uint32_t* digests = (const uint32_t*)outp;
  const uint32_t f = ((const uint32_t*)(outp + o_flags))[gc];
  auto put_digest = [&](uint8_t* dst, uint32_t gmsg) {
    const uint32_t* w = digests + (size_t)gmsg * 8;
    uint32_t be[8];
    for (int k = 0; k < 8; k++) be[k] = __builtin_bswap32(w[k]);   // state words -> digest bytes
    memcpy(dst, be, 32);
  };
  put_digest(res.body_hash, t.msg_base + cd.body_msg);
  put_digest(res.header_hash, t.msg_base + cd.hdr_msg);
  res.bh_ok = ((fo.flags & FE_BH_VALID) && (f & ZKB_F_BH_OK)) ? 1 : 0;
  int detail = ZKB_DKIM_PASS;
  if (!res.bh_ok) detail = ZKB_DKIM_BODY_HASH;
  else if (fo.flags & FE_SIG_SYNTAX) detail = ZKB_DKIM_SIG_SYNTAX;
  else if ((fo.flags & FE_SIG_BADLEN) || !(f & ZKB_F_RSA_OK)) detail = ZKB_DKIM_SIG_MISMATCH;
  // several signature headers: the reference would go on to the later ones; only a pass is final here
  if (detail != ZKB_DKIM_PASS && (fo.flags & FE_MULTI)) return false;
  res.dkim_detail = detail;
  if (detail != ZKB_DKIM_PASS) { res.status = ZKB_ST_DKIM_FAIL; return true; }
  res.rsa_ok = 1;
  put_digest(res.from_domain_hash, t.msg_base + er.dom_msg);
  put_digest(res.public_key_hash, t.msg_base + er.key_msg);
  if (!rs) return true;
  const size_t P = rs->n_active();
  const uint4* dfa = (const uint4*)(outp + o_dfa) + i * P;
  const zkb_email_view& view = ch.views[i];
  const uint8_t* raw = view.raw_email;
  size_t pi = 0;
  for (size_t p = 0; p < rs->parts.size(); p++) {
    const bool body = rs->parts[p].body;
    if (body ? !rs->body_present : !rs->header_present) continue;
    const uint4 r = dfa[pi++];
    uint32_t slot = res.n_parts;
    if (slot < ZKB_MAX_PARTS) {
      res.parts[slot].match_count = r.x; res.parts[slot].start = r.y; res.parts[slot].end = r.z; res.parts[slot].captures_ok = 0;
      res.n_parts++;
    }
    bool ok = r.x == 1;
    if (ok && caps) {
      const zkb_email_captures& ec = caps[ch.e0 + i];
      bool loaded = false;
      for (size_t q = 0; q < ec.n_caps && ok; q++) {
        if (ec.caps[q].part != p) continue;
        if (!loaded) {  // rebuild the haystack on the host from the raw message (only emails with captures pay)
          HayView hv;
          if (body && !(fo.flags & FE_HAS_L)) {
            scratch.resize((size_t)fo.body_len + 64);
            size_t cl = (fo.flags & FE_BODY_RELAXED) ? canon_body_relaxed(raw + fo.body_off, fo.body_len, scratch.data())
                                                     : canon_body_simple(raw + fo.body_off, fo.body_len, scratch.data());
            hv.p = scratch.data(); hv.n = (uint32_t)cl;
          } else {
            uint8_t *hp = nullptr, *bp = nullptr;
            size_t hl = 0, bl = 0;
            int detail2 = 0;
            if (zkb_host_canonicalize(raw, view.raw_email_len, e->now_unix, &hp, &hl, &bp, &bl, &detail2) != ZKB_OK) { ok = false; break; }
            if (body) scratch.assign(bp, bp + bl);   // l= bodies: the host canonicaliser applies the truncation
            else scratch.assign(hp, hp + hl);
            free(hp); free(bp);
            hv.p = scratch.data(); hv.n = (uint32_t)scratch.size();
          }
          cleaned_span(hv, body, r.y, r.z, s1);
          bool ascii = true;
          for (char ch2 : s1) if (ch2 & 0x80) { ascii = false; break; }
          if (!ascii) { utf8_lossy((const uint8_t*)s1.data(), s1.size(), s2); s1.swap(s2); }
          loaded = true;
        }
        if (ec.caps[q].len && s1.find(ec.caps[q].s, 0, ec.caps[q].len) == std::string::npos) ok = false;
      }
    }
    if (ok && slot < ZKB_MAX_PARTS) res.parts[slot].captures_ok = 1;
    if (!ok) { res.status = body ? ZKB_ST_REGEX_BODY : ZKB_ST_REGEX_HEADER; return true; }
  }
  return true;
}

// Expected-capture check of one part of one email against its haystack rebuilt on the host from the raw message
// (core/src/regex.rs:41-48; only emails that carry captures pay).  fo: the device front end's record of the message.
bool captures_hold(const zkb_engine* e, const zkb_email_view& view, const FeOut& fo, bool body, uint32_t start, uint32_t end,
                   const zkb_email_captures& ec, size_t part, std::vector<uint8_t>& scratch, std::string& s1, std::string& s2) {
  bool loaded = false;
  for (size_t q = 0; q < ec.n_caps; q++) {
    if (ec.caps[q].part != part) continue;
    if (!loaded) {
      const uint8_t* raw = view.raw_email;
      HayView hv;
      if (body && !(fo.flags & FE_HAS_L)) {
        scratch.resize((size_t)fo.body_len + 64);
        size_t cl = (fo.flags & FE_BODY_RELAXED) ? canon_body_relaxed(raw + fo.body_off, fo.body_len, scratch.data())
                                                 : canon_body_simple(raw + fo.body_off, fo.body_len, scratch.data());
        hv.p = scratch.data(); hv.n = (uint32_t)cl;
      } else {
        uint8_t *hp = nullptr, *bp = nullptr;
        size_t hl = 0, bl = 0;
        int detail2 = 0;
        if (zkb_host_canonicalize(raw, view.raw_email_len, e->now_unix, &hp, &hl, &bp, &bl, &detail2) != ZKB_OK) return false;
        if (body) scratch.assign(bp, bp + bl);   // l= bodies: the host canonicaliser applies the truncation
        else scratch.assign(hp, hp + hl);
        free(hp); free(bp);
        hv.p = scratch.data(); hv.n = (uint32_t)scratch.size();
      }
      cleaned_span(hv, body, start, end, s1);
      bool ascii = true;
      for (char ch2 : s1) if (ch2 & 0x80) { ascii = false; break; }
      if (!ascii) { utf8_lossy((const uint8_t*)s1.data(), s1.size(), s2); s1.swap(s2); }
      loaded = true;
    }
    if (ec.caps[q].len && s1.find(ec.caps[q].s, 0, ec.caps[q].len) == std::string::npos) return false;
  }
  return true;
}

// Device-front-end chunks: the records were built on the device (assemble.cuh); the host copies them into the caller's
// array, writes the few emails decided before the device (bad key, unsupported key type), collects the messages the
// device declined and, when the caller passed expected captures, applies the substring checks.
void resolve_chunk_recs(zkb_engine* e, const Chunk& ch, const uint8_t* outp, size_t o_recs, size_t o_fe, const zkb_regex_set* rs,
                        const zkb_email_captures* caps, zkb_result* out, std::vector<size_t>* fallback) {
  const size_t P = rs ? rs->n_active() : 0;
  const uint32_t rw = rec_words_for(P);
  const uint32_t* recs = (const uint32_t*)(outp + o_recs);
  std::mutex fb_mu;
  e->pool->parallel_for(ch.ne, 1024, [&](size_t lo, size_t hi, int) {
    std::string s1, s2;
    std::vector<uint8_t> scratch;
    std::vector<size_t> local_fb;
    for (size_t i = lo; i < hi; i++) {
      zkb_result& res = out[ch.e0 + i];
      const EmailRec& er = ch.emails[i];
      if (er.status != ZKB_ST_OK) {
        memset(&res, 0, rs ? sizeof res : offsetof(zkb_result, parts));
        res.status = er.status; res.dkim_detail = ZKB_DKIM_NEUTRAL;
        continue;
      }
      const uint32_t* r = recs + i * (size_t)rw;
      if ((int32_t)r[0] == ZKB_REC_REDO) { local_fb.push_back(ch.e0 + i); continue; }
      memcpy(&res, r, ZKB_REC_HEAD);
      if (!rs) continue;
      memcpy(res.parts, r + 36, P * 16);
      if (P < ZKB_MAX_PARTS) memset(&res.parts[P], 0, (ZKB_MAX_PARTS - P) * 16);
      if (!caps || res.n_parts == 0) continue;
      const zkb_email_captures& ec = caps[ch.e0 + i];
      if (!ec.n_caps) continue;
      const ThreadRecs& t = ch.tr[er.tid];
      const FeOut& fo = ((const FeOut*)(outp + o_fe))[t.cand_base + er.fe_cand];
      size_t slot = 0;
      for (size_t p = 0; p < rs->parts.size() && slot < res.n_parts; p++) {
        const bool body = rs->parts[p].body;
        if (body ? !rs->body_present : !rs->header_present) continue;
        if (res.parts[slot].captures_ok &&
            !captures_hold(e, ch.views[i], fo, body, res.parts[slot].start, res.parts[slot].end, ec, p, scratch, s1, s2)) {
          // the reference stops at this part (core/src/circuits.rs:45,54): later parts were never evaluated
          res.parts[slot].captures_ok = 0;
          res.status = body ? ZKB_ST_REGEX_BODY : ZKB_ST_REGEX_HEADER;
          res.n_parts = (uint32_t)slot + 1;
          memset(&res.parts[slot + 1], 0, (ZKB_MAX_PARTS - slot - 1) * 16);
          break;
        }
        slot++;
      }
    }
    if (!local_fb.empty() && fallback) { std::lock_guard<std::mutex> l(fb_mu); fallback->insert(fallback->end(), local_fb.begin(), local_fb.end()); }
  });
}

// fallback (optional): chunk-local indices of the messages the device front end declined
void resolve_chunk(zkb_engine* e, const Chunk& ch, const uint8_t* outp, const zkb_regex_set* rs, const zkb_email_captures* caps,
                   zkb_result* out, std::vector<size_t>* fallback = nullptr) {
  const size_t P = rs ? rs->n_active() : 0;
  size_t o_flags, o_dfa, o_fe = 0, o_recs = 0;
  out_layout(ch.M, ch.C, ch.ne, P, o_flags, o_dfa, ch.fe ? ch.C : 0, &o_fe, &o_recs);
  if (ch.fe && ch.C) { resolve_chunk_recs(e, ch, outp, o_recs, o_fe, rs, caps, out, fallback); return; }
  std::mutex fb_mu;
  e->pool->parallel_for(ch.ne, 256, [&](size_t lo, size_t hi, int) {
    std::string s1, s2;
    std::vector<uint8_t> scratch;
    std::vector<size_t> local_fb;
    const uint32_t rw = rec_words_for(P);
    const uint32_t* recs = (const uint32_t*)(outp + o_recs);
    for (size_t i = lo; i < hi; i++) {
      if (!ch.fe) {
        // host front end: emails decided by one signature candidate carry a device-built record; the rest (several
        // candidates, validation errors, expected captures to check) are resolved here
        zkb_result& res = out[ch.e0 + i];
        const uint32_t* r = recs + i * (size_t)rw;
        if (!caps && ch.emails[i].status == ZKB_ST_OK && (int32_t)r[0] != ZKB_REC_REDO) {
          memcpy(&res, r, ZKB_REC_HEAD);
          if (rs) {
            memcpy(res.parts, r + 36, P * 16);
            if (P < ZKB_MAX_PARTS) memset(&res.parts[P], 0, (ZKB_MAX_PARTS - P) * 16);
          }
          continue;
        }
        resolve_email(e, ch, i, outp, o_flags, o_dfa, rs, caps, scratch, res, s1, s2);
        continue;
      }
      if (!resolve_email_fe(e, ch, i, outp, o_flags, o_dfa, o_fe, rs, caps, scratch, out[ch.e0 + i], s1, s2)) local_fb.push_back(ch.e0 + i);
    }
    if (!local_fb.empty() && fallback) { std::lock_guard<std::mutex> l(fb_mu); fallback->insert(fallback->end(), local_fb.begin(), local_fb.end()); }
  });
}

void release_blocks(zkb_engine* e, Chunk& ch) {
  for (auto& t : ch.tr) { for (auto& b : t.blocks) e->blocks.put(b); t.blocks.clear(); }
}

int ensure_dfa_attr(zkb_engine* e) {
  CK(dfa_set_smem_limit(e->smem_optin));
  return ZKB_OK;
}

}  // namespace

// =================================================================== C ABI
extern "C" {

int zkb_abi_version(void) { return ZKB_ABI_VERSION; }

const char* zkb_strerror(int code) {
  switch (code) {
    case ZKB_OK: return "ok";
    case ZKB_E_INVALID: return "invalid argument";
    case ZKB_E_NO_DEVICE: return "no usable CUDA device (zkemail_b200 has no CPU fallback)";
    case ZKB_E_CUDA: return "CUDA runtime error";
    case ZKB_E_NOMEM: return "out of memory";
    case ZKB_E_REGEX: return "regex pattern or DFA table rejected";
    case ZKB_E_UNSUPPORTED: return "unsupported";
  }
  return "unknown error";
}

int zkb_engine_create(const zkb_options* opt, zkb_engine** out) {
  if (!out) return ZKB_E_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return ZKB_E_NO_DEVICE; }
  int dev = opt ? opt->device : 0;
  if (dev < 0 || dev >= ndev) return ZKB_E_INVALID;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major < 10) {
    fprintf(stderr, "[zkemail_b200] device %d is sm_%d%d; this library carries sm_100a code only\n", dev, prop.major, prop.minor);
    return ZKB_E_NO_DEVICE;
  }
  zkb_engine* e = new zkb_engine();
#define CKE(expr) do { if ((expr) != cudaSuccess) { fprintf(stderr, "[zkemail_b200] CUDA error at %s:%d: %s\n", __FILE__, __LINE__, cudaGetErrorString(cudaGetLastError())); zkb_engine_destroy(e); return ZKB_E_CUDA; } } while (0)
  e->device = dev;
  e->sm_count = prop.multiProcessorCount;
  e->smem_optin = prop.sharedMemPerBlockOptin;
  e->now_unix = opt ? opt->now_unix : 0;
  if (opt && opt->chunk_emails) e->chunk_emails = (size_t)opt->chunk_emails;
  if (opt && opt->rsa_lanes) e->rsa_lanes = opt->rsa_lanes;
  if (opt) e->flags.store(opt->flags);
  int nt = opt ? opt->host_threads : 0;
  if (nt <= 0) nt = (int)std::thread::hardware_concurrency();
  if (nt <= 0) nt = 1;
  if (nt > 256) nt = 256;
  e->pool = new ThreadPool(nt);
  for (auto& s : e->slots) {
    CKE(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    CKE(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    CKE(cudaEventCreate(&s.k0)); CKE(cudaEventCreate(&s.k1)); CKE(cudaEventCreate(&s.h0));
  }
  for (auto& ev : e->ev) CKE(cudaEventCreate(&ev));
  {
    int lo = 0, hi = 0;
    CKE(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CKE(cudaStreamCreateWithPriority(&e->aux_stream, cudaStreamNonBlocking, hi));
    CKE(cudaStreamCreateWithPriority(&e->aux_stream2, cudaStreamNonBlocking, hi));
    CKE(cudaEventCreateWithFlags(&e->ev_join2, cudaEventDisableTiming));
    CKE(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
    CKE(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
  }
  if (ensure_dfa_attr(e)) { zkb_engine_destroy(e); return ZKB_E_CUDA; }
#undef CKE
  *out = e;
  return ZKB_OK;
}

void zkb_engine_destroy(zkb_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  for (auto& s : e->slots) {
    s.dev.free(); s.meta.free(); s.result.free();
    if (s.stream) cudaStreamDestroy(s.stream);
    if (s.done) cudaEventDestroy(s.done);
    if (s.k0) cudaEventDestroy(s.k0);
    if (s.k1) cudaEventDestroy(s.k1);
    if (s.h0) cudaEventDestroy(s.h0);
  }
  for (auto& ev : e->ev) if (ev) cudaEventDestroy(ev);
  for (auto& ev : e->ev_pre) if (ev) cudaEventDestroy(ev);
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_join) cudaEventDestroy(e->ev_join);
  if (e->ev_join2) cudaEventDestroy(e->ev_join2);
  if (e->aux_stream) cudaStreamDestroy(e->aux_stream);
  if (e->aux_stream2) cudaStreamDestroy(e->aux_stream2);
  if (e->d_keytab) cudaFree(e->d_keytab);
  if (e->borrowed_registrations == 0)   // engines of a zkb_multi other than the first only borrow the first one's registrations
    for (auto& r : e->registered) cudaHostUnregister(const_cast<uint8_t*>(r.first));
  e->blocks.release_all();
  delete e->pool;
  delete e;
}

int zkb_host_register(zkb_engine* e, const void* p, size_t len) {
  if (!e || !p || !len) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lock(e->run_mu);
  CK(cudaSetDevice(e->device));
  CK(cudaHostRegister(const_cast<void*>(p), len, cudaHostRegisterPortable));
  e->registered.emplace_back((const uint8_t*)p, len);
  return ZKB_OK;
}

int zkb_host_unregister(zkb_engine* e, const void* p) {
  if (!e || !p) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lock(e->run_mu);
  CK(cudaSetDevice(e->device));
  for (size_t i = 0; i < e->registered.size(); i++)
    if (e->registered[i].first == (const uint8_t*)p) {
      CK(cudaDeviceSynchronize());
      CK(cudaHostUnregister(const_cast<void*>(p)));
      e->registered.erase(e->registered.begin() + (long)i);
      return ZKB_OK;
    }
  return ZKB_E_INVALID;
}

int zkb_engine_last_batch_bytes(const zkb_engine* e, uint64_t* h2d, uint64_t* d2h, uint64_t* host_front_end_emails) {
  if (!e) return ZKB_E_INVALID;
  if (h2d) *h2d = e->last_h2d;
  if (d2h) *d2h = e->last_d2h;
  if (host_front_end_emails) *host_front_end_emails = e->last_fallback;
  return ZKB_OK;
}

int zkb_engine_set_flags(zkb_engine* e, uint32_t flags) {
  if (!e || (flags & ~(uint32_t)ZKB_OPT_ALL)) return ZKB_E_INVALID;
  e->flags.store(flags);
  return ZKB_OK;
}

void* zkb_engine_stream(zkb_engine* e) { return e ? (void*)e->slots[0].stream : nullptr; }

int zkb_regex_set_create(zkb_engine* e, const zkb_dfa_view* parts, size_t n_header, size_t n_body, int header_present,
                         int body_present, zkb_regex_set** out) {
  if (!e || !out || (!parts && n_header + n_body)) return ZKB_E_INVALID;
  *out = nullptr;
  CK(cudaSetDevice(e->device));
  size_t active = (header_present ? n_header : 0) + (body_present ? n_body : 0);
  if (active > ZKB_MAX_PARTS) return ZKB_E_UNSUPPORTED;
  zkb_regex_set* rs = new zkb_regex_set();
  rs->eng = e; rs->n_header = n_header; rs->n_body = n_body;
  rs->header_present = header_present; rs->body_present = body_present;
  for (size_t p = 0; p < n_header + n_body; p++) {
    std::vector<uint8_t> fb, rb;
    uint32_t fe = 2;
    bool direct = false;
    if (!build_dfa_pair(parts[p].fwd, parts[p].fwd_len, parts[p].bwd, parts[p].bwd_len, fb, rb, fe, direct)) {
      zkb_regex_set_destroy(rs);
      return ZKB_E_REGEX;
    }
    zkb_regex_set::Part part;
    part.body = p >= n_header;
    part.elem = fe;
    part.direct = direct;
    part.fwd_bytes = (uint32_t)fb.size(); part.rev_bytes = (uint32_t)rb.size();
    if (cudaMalloc((void**)&part.d_fwd, fb.size()) != cudaSuccess || cudaMalloc((void**)&part.d_rev, rb.size()) != cudaSuccess) {
      if (part.d_fwd) cudaFree(part.d_fwd);
      zkb_regex_set_destroy(rs);
      return ZKB_E_NOMEM;
    }
    rs->parts.push_back(part);
    if (cudaMemcpy(part.d_fwd, fb.data(), fb.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(part.d_rev, rb.data(), rb.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
      zkb_regex_set_destroy(rs);
      return ZKB_E_CUDA;
    }
  }
  *out = rs;
  return ZKB_OK;
}

void zkb_regex_set_destroy(zkb_regex_set* s) {
  if (!s) return;
  for (auto& p : s->parts) { if (p.d_fwd) cudaFree(p.d_fwd); if (p.d_rev) cudaFree(p.d_rev); }
  delete s;
}

// Chunk boundaries: at most `max_emails` emails and about `max_bytes` raw bytes per chunk, so that
// large-body batches still pipeline (H2D of chunk k overlaps the host pack of chunk k+1).
static std::vector<size_t> chunk_bounds(const zkb_email_view* emails, size_t n, size_t max_emails, size_t max_bytes) {
  std::vector<size_t> b;
  b.push_back(0);
  size_t cnt = 0, bytes = 0;
  for (size_t i = 0; i < n; i++) {
    cnt++; bytes += emails[i].raw_email_len;
    if (cnt >= max_emails || bytes >= max_bytes) { b.push_back(i + 1); cnt = 0; bytes = 0; }
  }
  if (b.back() != n) b.push_back(n);
  return b;
}

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Pipelined end-to-end batch: pack chunk k+1 on the host while chunk k is on the device.
static int verify_batch_impl(zkb_engine* e, const zkb_email_view* emails, size_t n, const zkb_regex_set* regex,
                             const zkb_email_captures* captures, zkb_result* out, bool allow_fe, std::vector<size_t>* fallback) {
  // chunk size: at most chunk_emails messages and about 256 MB of raw bytes; batches of larger messages get up to
  // 1 GB per chunk so that a chunk still holds tens of thousands of messages (the per-message kernels map one lane
  // to a message: 16 K-message chunks of the mixed-size sweep left them latency-bound and the pipeline kernel-bound)
  size_t max_bytes = (size_t)256 << 20;
  if (n) {
    size_t sample = 0;
    const size_t step = n / 64 + 1;
    size_t cnt = 0;
    for (size_t i = 0; i < n; i += step) { sample += emails[i].raw_email_len; cnt++; }
    const size_t avg = sample / cnt;
    max_bytes = std::min<size_t>((size_t)1 << 30, std::max<size_t>(max_bytes, avg * e->chunk_emails));
  }
  const std::vector<size_t> bounds = chunk_bounds(emails, n, e->chunk_emails, max_bytes);
  const size_t nchunks = bounds.size() - 1;
  std::vector<ThreadCtx> ctxs(e->pool->size());
  Chunk chunks[3];
  bool busy[3] = {false, false, false};
  const size_t P = regex ? regex->n_active() : 0;
  const bool prof = e->has(ZKB_OPT_PROFILE);
  double t_pack = 0, t_upload = 0, t_wait = 0, t_resolve = 0, t_all = now_s();
  double t_gpu_h2d = 0, t_gpu_kern = 0;   // device-side stream time (ms) of the copies / kernels, summed over chunks
  e->prof_parse = e->prof_layout = e->prof_prelude = 0; e->prof_busy_ns = 0;
  auto finish = [&](int si) -> int {
    Slot& s = e->slots[si];
    Chunk& ch = chunks[si];
    double ta = now_s();
    CK(cudaEventSynchronize(s.done));
    double tb = now_s();
    t_wait += tb - ta;
    if (prof) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, s.h0, s.k0) == cudaSuccess) t_gpu_h2d += ms;
      if (cudaEventElapsedTime(&ms, s.k0, s.k1) == cudaSuccess) t_gpu_kern += ms;
    }
    resolve_chunk(e, ch, s.result.p, regex, captures, out, fallback);
    release_blocks(e, ch);
    busy[si] = false;
    t_resolve += now_s() - tb;
    return ZKB_OK;
  };
  int rc = ZKB_OK;
  for (size_t k = 0; k < nchunks && rc == ZKB_OK; k++) {
    int si = (int)(k % 3);
    if (busy[si]) rc = finish(si);
    if (rc) break;
    Slot& s = e->slots[si];
    Chunk& ch = chunks[si];
    size_t e0 = bounds[k], ne = bounds[k + 1] - bounds[k];
    double t0 = now_s();
    rc = pack_chunk(e, emails, e0, ne, regex, ch, s.meta, ctxs, allow_fe);
    if (rc) break;
    double t1 = now_s();
    t_pack += t1 - t0;
    rc = sync_keytab(e, s.stream);
    if (rc) break;
    if (prof) CKB(cudaEventRecord(s.h0, s.stream));
    rc = upload_chunk(e, ch, regex, s.dev, s.meta, s.stream);
    if (rc) break;
    if (prof) CKB(cudaEventRecord(s.k0, s.stream));
    rc = launch_chunk(e, s.dev, regex, s.stream, nullptr, nullptr);
    if (rc) break;
    if (prof) CKB(cudaEventRecord(s.k1, s.stream));
    if (!s.result.ensure(s.dev.out_bytes)) { rc = ZKB_E_NOMEM; break; }
    {
      // device front end: the per-email records are the result (the front-end records travel too when expected
      // captures have to be checked against haystacks rebuilt on the host); host front end: the raw kernel outputs
      size_t from = 0;
      if (ch.fe && ch.C) from = captures ? (size_t)((uint8_t*)s.dev.fe_out - s.dev.out.p) : s.dev.o_recs;
      const size_t nbytes = s.dev.out_bytes - from;
      e->last_h2d += ch.st.h2d_bytes; e->last_d2h += nbytes;
      CKB(cudaMemcpyAsync(s.result.p + from, s.dev.out.p + from, nbytes, cudaMemcpyDeviceToHost, s.stream));
    }
    CKB(cudaEventRecord(s.done, s.stream));
    busy[si] = true;
    t_upload += now_s() - t1;
    // resolve the chunk launched two iterations ago: two chunks stay in flight (one in H2D, one in kernels)
    // while the host packs the next, so the host only waits when the device is the slower side
    if (k >= 2) {
      int pj = (int)((k - 2) % 3);
      if (busy[pj]) rc = finish(pj);
    }
  }
  for (size_t k = nchunks >= 2 ? nchunks - 2 : 0; k < nchunks; k++) {  // drain in launch order
    int si = (int)(k % 3);
    if (busy[si]) { int r2 = finish(si); if (rc == ZKB_OK) rc = r2; }
  }
  for (int si = 0; si < 3; si++) {
    if (busy[si]) { int r2 = finish(si); if (rc == ZKB_OK) rc = r2; }
    // error exit: copies out of this slot's pinned blocks may still be queued; they must finish before the blocks
    // return to the pool and the stack-owned chunk records die
    if (rc != ZKB_OK && e->slots[si].stream) cudaStreamSynchronize(e->slots[si].stream);
    release_blocks(e, chunks[si]);
  }
  (void)P;
  if (prof)
    fprintf(stderr, "[zkb profile] n=%zu chunks=%zu threads=%d total=%.1fms pack=%.1f (parse %.1f [prelude %.1f, thread-busy %.1f], layout %.1f) upload+launch=%.1f wait_gpu=%.1f resolve=%.1f | stream time: h2d %.1f kernels %.1f\n",
            n, nchunks, e->pool->size(), 1e3 * (now_s() - t_all), 1e3 * t_pack, 1e3 * e->prof_parse, 1e3 * e->prof_prelude,
            1e-6 * (double)e->prof_busy_ns.load() / e->pool->size(), 1e3 * e->prof_layout, 1e3 * t_upload, 1e3 * t_wait, 1e3 * t_resolve,
            t_gpu_h2d, t_gpu_kern);
  return rc;
}

int zkb_verify_batch(zkb_engine* e, const zkb_email_view* emails, size_t n, const zkb_regex_set* regex,
                     const zkb_email_captures* captures, zkb_result* out) {
  if (!e || (!emails && n) || (!out && n)) return ZKB_E_INVALID;
  if (regex && regex->eng != e) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lock(e->run_mu);
  CK(cudaSetDevice(e->device));
  std::vector<size_t> fb;
  trim_keytab(e);
  e->last_h2d = e->last_d2h = e->last_fallback = 0;
  int rc = verify_batch_impl(e, emails, n, regex, captures, out, true, &fb);
  e->last_fallback = fb.size();
  if (rc || fb.empty()) return rc;
  // messages the device front end declined (irregular structure): the host front end implements every path
  std::sort(fb.begin(), fb.end());
  std::vector<zkb_email_view> v2(fb.size());
  std::vector<zkb_email_captures> c2(captures ? fb.size() : 0);
  std::vector<zkb_result> r2(fb.size());
  for (size_t i = 0; i < fb.size(); i++) { v2[i] = emails[fb[i]]; if (captures) c2[i] = captures[fb[i]]; }
  rc = verify_batch_impl(e, v2.data(), v2.size(), regex, captures ? c2.data() : nullptr, r2.data(), false, nullptr);
  if (rc) return rc;
  const size_t rec_bytes = regex ? sizeof(zkb_result) : offsetof(zkb_result, parts);
  for (size_t i = 0; i < fb.size(); i++) memcpy(&out[fb[i]], &r2[i], rec_bytes);
  if (e->has(ZKB_OPT_PROFILE)) fprintf(stderr, "[zkb profile] device front end declined %zu of %zu messages (host front end used)\n", fb.size(), n);
  return ZKB_OK;
}

int zkb_verify_one(zkb_engine* e, const zkb_email_view* email, const zkb_regex_set* regex, const zkb_email_captures* captures,
                   zkb_result* out) {
  return zkb_verify_batch(e, email, 1, regex, captures, out);
}

// ---- resident form: prepare (pack + H2D) / run (kernels) / fetch (D2H + resolve) ----
void zkb_batch_destroy(zkb_batch* b) {
  if (!b) return;
  if (b->eng) cudaSetDevice(b->eng->device);
  for (auto* c : b->chunks) { if (c) { release_blocks(b->eng, *c); delete c; } }
  for (auto* d : b->dev) { if (d) { d->free(); delete d; } }
  if (b->eng) b->eng->live_batches.fetch_sub(1);
  delete b;
}

static int batch_prepare_impl(zkb_engine* e, const zkb_email_view* emails, size_t n, const zkb_regex_set* regex,
                              const zkb_email_captures* captures, zkb_batch** out, bool raw) {
  if (!e || !out || (!emails && n)) return ZKB_E_INVALID;
  if (regex && regex->eng != e) return ZKB_E_INVALID;
  *out = nullptr;
  std::lock_guard<std::mutex> lock(e->run_mu);
  CK(cudaSetDevice(e->device));
  trim_keytab(e);
  zkb_batch* b = new zkb_batch();
  { static std::atomic<uint64_t> next_serial{1}; b->serial = next_serial.fetch_add(1); }
  e->live_batches.fetch_add(1);
  b->eng = e; b->regex = regex; b->n = n; b->emails = emails; b->captures = captures;
  b->raw = raw;
  // resident batches: fewer, larger launches (better SM balance).  When the whole batch lies in registered memory
  // (raw bytes are DMA'd, nothing is staged in pinned blocks) a chunk may hold up to 16 GB, so that batches of large
  // messages still give the lane-per-message kernels tens of thousands of lanes per launch (100 KB bodies: 29 K
  // lanes per 3 GB chunk left the SHA-256 kernel latency-bound at 1.5 warps per scheduler).
  size_t max_bytes = (size_t)3 << 30;
  if (n && !e->registered.empty() && !e->has(ZKB_OPT_NO_DIRECT)) {
    const uint8_t *lo = emails[0].raw_email, *hi = lo;
    for (size_t i = 0; i < n; i++) { lo = std::min(lo, emails[i].raw_email); hi = std::max(hi, emails[i].raw_email + emails[i].raw_email_len); }
    for (auto& r : e->registered)
      if (lo >= r.first && hi <= r.first + r.second) { max_bytes = (size_t)16 << 30; break; }
  }
  const std::vector<size_t> bounds = chunk_bounds(emails, n, e->chunk_emails * 4, max_bytes);
  std::vector<ThreadCtx> ctxs(e->pool->size());
  cudaStream_t s = e->slots[0].stream;
  int rc = ZKB_OK;
  for (size_t k = 0; k + 1 < bounds.size() && rc == ZKB_OK; k++) {
    Chunk* ch = new Chunk();
    DeviceChunk* d = new DeviceChunk();
    b->chunks.push_back(ch); b->dev.push_back(d);
    rc = pack_chunk(e, emails, bounds[k], bounds[k + 1] - bounds[k], regex, *ch, e->slots[0].meta, ctxs, raw, max_bytes + ((size_t)1 << 30));
    if (rc) break;
    rc = sync_keytab(e, s);
    if (rc) break;
    rc = upload_chunk(e, *ch, regex, *d, e->slots[0].meta, s);
    if (rc) break;
    if (raw && ch->fe) {
      // raw-resident form: the raw bytes (registered span or staged copies) stay in HBM and every run starts at the
      // device front end; nothing is precomputed here
      if (cudaStreamSynchronize(s) != cudaSuccess) { rc = ZKB_E_CUDA; break; }
      release_blocks(e, *ch);   // staged copies have reached the arena
      continue;
    }
    if (d->n_canon) {
      // resident form: device-side canonicalisation is part of getting the batch resident (it is what
      // the host threads do on the pageable path); zkb_batch_run then launches the verification kernels
      launch_canon_body(d->span.p, d->canon_items, d->n_canon, d->arena.p, d->msg_off, d->msg_len_rw, d->order, d->msg_canon, d->M, s);
      d->n_canon = 0;
    }
    if (cudaStreamSynchronize(s) != cudaSuccess) { rc = ZKB_E_CUDA; break; }
    d->span.free();
  }
  if (rc) { zkb_batch_destroy(b); return rc; }
  *out = b;
  return ZKB_OK;
}

int zkb_batch_prepare(zkb_engine* e, const zkb_email_view* emails, size_t n, const zkb_regex_set* regex,
                      const zkb_email_captures* captures, zkb_batch** out) {
  return batch_prepare_impl(e, emails, n, regex, captures, out, false);
}

int zkb_batch_prepare_raw(zkb_engine* e, const zkb_email_view* emails, size_t n, const zkb_regex_set* regex,
                          const zkb_email_captures* captures, zkb_batch** out) {
  return batch_prepare_impl(e, emails, n, regex, captures, out, true);
}

// The enqueue of zkb_batch_run_async with e->run_mu held.  after_chunk(k, s), if given, runs once the last kernel of
// resident chunk k (the one that writes its result records) has been enqueued on stream s (zkb_comm_run_allgather hooks
// the exchange of that chunk's records there).
using AfterChunk = std::function<int(size_t, cudaStream_t)>;
static int batch_run_async_locked(zkb_batch* b, const AfterChunk* after_chunk) {
  zkb_engine* e = b->eng;
  CK(cudaSetDevice(e->device));
  cudaStream_t s = e->slots[0].stream;
  const size_t nc = b->dev.size();
  if (nc < 2 || e->has(ZKB_OPT_NO_OVERLAP)) {
    for (auto& d : b->dev) {
      // flags and DFA outputs accumulate with atomicOr / plain stores: reset them for a re-run
      size_t o_flags, o_dfa;
      out_layout(d->M, d->C, d->NE, d->P, o_flags, o_dfa);
      CK(cudaMemsetAsync(d->out.p + o_flags, 0, align_up((size_t)d->C * 4, 16), s));
      int rc = launch_chunk(e, *d, b->regex, s, nullptr, nullptr);
      if (rc) return rc;
      if (after_chunk && (rc = (*after_chunk)((size_t)(&d - &b->dev[0]), s)) != 0) return rc;
    }
    b->ran = true;
    return ZKB_OK;
  }
  // Two streams: the engine stream carries the RSA launches of every chunk back to back; a high-priority side
  // stream carries everything else (SHA-256, bh= check, DFA scans) and runs ahead, so launches of different chunks
  // fill each other's tails and low-occupancy phases (large bodies give few lanes per chunk: +14 % on 100 KB bodies,
  // +15 % on the mixed-size sweep, +4 % with regex parts).  It does NOT buy pipe-level overlap of SHA-256 with RSA on
  // 4 KB mail (measured: +0.5 %, and capping the RSA CTAs per SM to make room for hashing CTAs, or keeping the SHA
  // additions off the FMA pipe, only lost time): IMAD.WIDE holds the issue port, the SM is saturated by either kernel.
  // The side stream forks from / joins the engine stream, so callers still order against that one stream.
  // Raw-resident batches (front end + canonicalisation in every run) alternate the chunks between TWO side streams: the
  // latency-bound canonicalisation and the issue-bound front end of chunk k + 1 then run beside the ALU-bound hashing of
  // chunk k instead of behind it (chunks share nothing but read-only tables).
  cudaStream_t auxs[2] = {e->aux_stream, (b->raw && e->aux_stream2 && !e->has(ZKB_OPT_NO_OVERLAP)) ? e->aux_stream2 : e->aux_stream};
  const bool two = auxs[1] != auxs[0];
  while (e->ev_pre.size() < nc) {
    cudaEvent_t ev = nullptr;
    CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    e->ev_pre.push_back(ev);
  }
  CK(cudaEventRecord(e->ev_fork, s));
  CK(cudaStreamWaitEvent(auxs[0], e->ev_fork, 0));
  if (two) CK(cudaStreamWaitEvent(auxs[1], e->ev_fork, 0));
  for (size_t k = 0; k < nc; k++) {
    DeviceChunk* d = b->dev[k];
    cudaStream_t aux = auxs[k & 1];
    size_t o_flags, o_dfa;
    out_layout(d->M, d->C, d->NE, d->P, o_flags, o_dfa);
    CK(cudaMemsetAsync(d->out.p + o_flags, 0, align_up((size_t)d->C * 4, 16), aux));
    int rc = launch_chunk(e, *d, b->regex, aux, nullptr, nullptr, 1);
    if (rc) return rc;
    CK(cudaEventRecord(e->ev_pre[k], aux));
    CK(cudaStreamWaitEvent(s, e->ev_pre[k], 0));
    rc = launch_chunk(e, *d, b->regex, s, nullptr, nullptr, 2);
    if (rc) return rc;
    if (after_chunk && (rc = (*after_chunk)(k, s)) != 0) return rc;
  }
  CK(cudaEventRecord(e->ev_join, auxs[0]));
  CK(cudaStreamWaitEvent(s, e->ev_join, 0));
  if (two) {
    CK(cudaEventRecord(e->ev_join2, auxs[1]));
    CK(cudaStreamWaitEvent(s, e->ev_join2, 0));
  }
  b->ran = true;
  return ZKB_OK;
}

int zkb_batch_run_async(zkb_batch* b) {
  if (!b) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lock(b->eng->run_mu);   // enqueue only; serialised with the other calls on this engine
  return batch_run_async_locked(b, nullptr);
}

int zkb_batch_run(zkb_batch* b) {
  if (!b) return ZKB_E_INVALID;
  zkb_engine* e = b->eng;
  std::lock_guard<std::mutex> lock(e->run_mu);
  CK(cudaSetDevice(e->device));
  cudaStream_t s = e->slots[0].stream;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  CK(cudaEventRecord(e->ev[5], s));
  for (auto* d : b->dev) {
    size_t o_flags, o_dfa;
    out_layout(d->M, d->C, d->NE, d->P, o_flags, o_dfa);
    CK(cudaMemsetAsync(d->out.p + o_flags, 0, align_up((size_t)d->C * 4, 16), s));
    CK(cudaEventRecord(e->ev[7], s));
    int rc = launch_chunk(e, *d, b->regex, s, e->ev, nullptr);
    if (rc) return rc;
    CK(cudaEventRecord(e->ev[8], s));
    CK(cudaEventSynchronize(e->ev[8]));
    float ms;
    CK(cudaEventElapsedTime(&ms, e->ev[0], e->ev[1])); acc[0] += ms;   // sha256
    CK(cudaEventElapsedTime(&ms, e->ev[2], e->ev[3])); acc[1] += ms;   // rsa
    CK(cudaEventElapsedTime(&ms, e->ev[3], e->ev[4])); acc[2] += ms;   // dfa
    CK(cudaEventElapsedTime(&ms, e->ev[1], e->ev[2])); acc[3] += ms;   // bh check
    CK(cudaEventElapsedTime(&ms, e->ev[7], e->ev[0])); acc[5] += ms;   // device front end + body canonicalisation (raw-resident batches)
    CK(cudaEventElapsedTime(&ms, e->ev[4], e->ev[8])); acc[6] += ms; // result records
  }
  CK(cudaEventRecord(e->ev[6], s));
  CK(cudaEventSynchronize(e->ev[6]));
  CK(cudaEventElapsedTime(&acc[4], e->ev[5], e->ev[6]));
  memcpy(b->last_ms, acc, sizeof acc);
  b->ran = true;
  return ZKB_OK;
}

int zkb_batch_last_timing(const zkb_batch* b, float ms[5]) {
  if (!b || !ms) return ZKB_E_INVALID;
  memcpy(ms, b->last_ms, 5 * sizeof(float));
  return ZKB_OK;
}

int zkb_batch_last_timing_ex(const zkb_batch* b, float* ms, size_t n) {
  if (!b || !ms) return ZKB_E_INVALID;
  for (size_t i = 0; i < n; i++) ms[i] = i < 8 ? b->last_ms[i] : 0.f;
  return ZKB_OK;
}

int zkb_batch_get_stats(const zkb_batch* b, zkb_batch_stats* out) {
  if (!b || !out) return ZKB_E_INVALID;
  memset(out, 0, sizeof *out);
  for (size_t i = 0; i < b->chunks.size(); i++) {
    const zkb_batch_stats& s = b->chunks[i]->st;
    out->n_emails += s.n_emails; out->n_candidates += s.n_candidates; out->n_sha_messages += s.n_sha_messages;
    out->sha_blocks += s.sha_blocks; out->sha_bytes += s.sha_bytes;
    out->rsa_items_1024 += s.rsa_items_1024; out->rsa_items_2048 += s.rsa_items_2048; out->rsa_items_other += s.rsa_items_other;
    out->rsa_macs += s.rsa_macs; out->dfa_items += s.dfa_items; out->dfa_bytes += s.dfa_bytes;
    out->arena_bytes += s.arena_bytes; out->h2d_bytes += s.h2d_bytes;
    out->d2h_bytes += b->dev[i]->out_bytes;
    const DeviceChunk& d = *b->dev[i];
    uint64_t nl = (d.M ? 1 : 0) + (d.C ? 1 : 0) + (d.n_fe ? 1 : 0) + (d.n_asm ? 1 : 0) + (d.n_canon ? 1 : 0);   // + front end, records, canonicalisation
    for (int k = 0; k < 6; k++) nl += d.rsa_n[k] ? 1 : 0;
    if (b->regex && d.n_dfa) nl += b->regex->n_active();
    out->kernel_launches += nl;
  }
  return ZKB_OK;
}

int zkb_batch_device_flags(const zkb_batch* b, size_t chunk, void** flags, size_t* n_flags, size_t* n_chunks) {
  if (!b) return ZKB_E_INVALID;
  if (n_chunks) *n_chunks = b->dev.size();
  if (chunk >= b->dev.size()) return flags ? ZKB_E_INVALID : ZKB_OK;
  if (flags) *flags = b->dev[chunk]->cand_flags;
  if (n_flags) *n_flags = b->dev[chunk]->C;
  return ZKB_OK;
}

int zkb_batch_fetch(zkb_batch* b, zkb_result* out) {
  if (!b || (!out && b->n)) return ZKB_E_INVALID;
  zkb_engine* e = b->eng;
  std::lock_guard<std::mutex> lock(e->run_mu);
  CK(cudaSetDevice(e->device));
  if (!b->ran) return ZKB_E_INVALID;
  cudaStream_t s = e->slots[0].stream;
  PinBuf& res = e->slots[0].result;
  std::vector<size_t> fb;
  for (size_t i = 0; i < b->chunks.size(); i++) {
    Chunk& ch = *b->chunks[i];
    DeviceChunk& d = *b->dev[i];
    if (!res.ensure(d.out_bytes)) return ZKB_E_NOMEM;
    CK(cudaMemcpyAsync(res.p, d.out.p, d.out_bytes, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    resolve_chunk(e, ch, res.p, b->regex, b->captures, out, b->raw ? &fb : nullptr);
  }
  if (fb.empty()) return ZKB_OK;
  // raw-resident batches: messages the device front end declined go through the host front end, as in zkb_verify_batch
  std::sort(fb.begin(), fb.end());
  std::vector<zkb_email_view> v2(fb.size());
  std::vector<zkb_email_captures> c2(b->captures ? fb.size() : 0);
  std::vector<zkb_result> r2(fb.size());
  for (size_t i = 0; i < fb.size(); i++) { v2[i] = b->emails[fb[i]]; if (b->captures) c2[i] = b->captures[fb[i]]; }
  int rc = verify_batch_impl(e, v2.data(), v2.size(), b->regex, b->captures ? c2.data() : nullptr, r2.data(), false, nullptr);
  if (rc) return rc;
  const size_t rec_bytes = b->regex ? sizeof(zkb_result) : offsetof(zkb_result, parts);
  for (size_t i = 0; i < fb.size(); i++) memcpy(&out[fb[i]], &r2[i], rec_bytes);
  return ZKB_OK;
}

// ---- regex compiler ----
int zkb_regex_compile(const char* pattern, size_t pattern_len, uint8_t** fwd, size_t* fwd_len, uint8_t** bwd, size_t* bwd_len,
                      char* err, size_t err_cap) {
  if (!pattern || !fwd || !fwd_len || !bwd || !bwd_len) return ZKB_E_INVALID;
  std::vector<uint8_t> f, r;
  std::string msg;
  if (!rx::compile(pattern, pattern_len, f, r, msg)) {
    if (err && err_cap) { strncpy(err, msg.c_str(), err_cap - 1); err[err_cap - 1] = 0; }
    return ZKB_E_REGEX;
  }
  *fwd = (uint8_t*)malloc(f.size()); *bwd = (uint8_t*)malloc(r.size());
  if (!*fwd || !*bwd) { free(*fwd); free(*bwd); return ZKB_E_NOMEM; }
  memcpy(*fwd, f.data(), f.size()); *fwd_len = f.size();
  memcpy(*bwd, r.data(), r.size()); *bwd_len = r.size();
  return ZKB_OK;
}
void zkb_free(void* p) { free(p); }

// ---- host-only canonicalize_signed_email ----
int zkb_host_canonicalize(const uint8_t* raw, size_t n, int64_t now_unix, uint8_t** hdr, size_t* hdr_len, uint8_t** body,
                          size_t* body_len, int* detail) {
  if ((!raw && n) || !hdr || !hdr_len || !body || !body_len) return ZKB_E_INVALID;
  int dummy;
  if (!detail) detail = &dummy;
  *hdr = *body = nullptr; *hdr_len = *body_len = 0;
  std::vector<HeaderField> hs;
  size_t body_off = 0;
  if (!parse_headers(raw, n, hs, body_off)) { *detail = 1; return ZKB_E_INVALID; }
  DkimSig sig;
  std::string scratch;
  *detail = 2;
  for (const HeaderField& h : hs) {
    if (!ieq_ascii(raw + h.key_off, h.key_len, "DKIM-Signature", 14)) continue;
    if (validate_dkim_header(raw + h.val_off, h.val_len, now_unix, sig) != ZKB_DKIM_PASS) continue;
    bool hr, br;
    if (!parse_canon_tag(sig, hr, br)) { *detail = 3; return ZKB_E_INVALID; }
    uint64_t l = 0;
    bool has_l = false;
    if (const Tag* tl = sig.get("l")) { has_l = true; if (!parse_usize_tag(sig, tl, l)) { *detail = 4; return ZKB_E_INVALID; } }
    const Tag* tb = sig.get("b");
    std::vector<uint8_t> tmp(tb->val_len + 4);
    if (base64_decode(sig.val(tb), tb->val_len, tmp.data()) < 0) { *detail = 5; return ZKB_E_INVALID; }
    size_t bl = 0;
    const uint8_t* b = find_body(raw, n, bl, body_off);
    uint8_t* ob = (uint8_t*)malloc(bl + 4);
    uint8_t* oh = (uint8_t*)malloc(preimage_bound(body_off, sig.n));
    if (!ob || !oh) { free(ob); free(oh); return ZKB_E_NOMEM; }
    size_t cl = br ? canon_body_relaxed(b, bl, ob) : canon_body_simple(b, bl, ob);
    if (has_l && l < cl) cl = (size_t)l;
    *body = ob; *body_len = cl;
    *hdr = oh; *hdr_len = build_header_preimage(raw, hs, sig, hr, oh, scratch);
    *detail = 0;
    return ZKB_OK;
  }
  return ZKB_E_INVALID;
}

// ---- kernel-level entry points (host buffers in / out) ----
int zkb_sha256_batch(zkb_engine* e, const uint8_t* data, size_t data_len, const uint64_t* off, const uint32_t* len, size_t n,
                     uint8_t* out) {
  if (!e || (!out && n) || (n && (!off || !len))) return ZKB_E_INVALID;
  if (n == 0) return ZKB_OK;
  std::lock_guard<std::mutex> lock(e->run_mu);
  CK(cudaSetDevice(e->device));
  // repack: every message 64-byte aligned and readable one block past its last full block
  std::vector<uint64_t> noff(n);
  uint64_t total = 0;
  for (size_t i = 0; i < n; i++) {
    if (off[i] + len[i] > data_len) return ZKB_E_INVALID;
    noff[i] = total; total += (((uint64_t)len[i] >> 6) + 1) << 6;
  }
  std::vector<uint8_t> host(total + 64, 0);
  for (size_t i = 0; i < n; i++) if (len[i]) memcpy(host.data() + noff[i], data + off[i], len[i]);
  uint8_t* d_arena; uint64_t* d_off; uint32_t* d_len; uint32_t* d_dig;
  DevTmp tmp;
  CK(tmp.alloc(&d_arena, host.size()));
  CK(tmp.alloc(&d_off, n * 8)); CK(tmp.alloc(&d_len, n * 4)); CK(tmp.alloc(&d_dig, n * 32));
  CK(cudaMemcpy(d_arena, host.data(), host.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_off, noff.data(), n * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_len, len, n * 4, cudaMemcpyHostToDevice));
  cudaStream_t s = e->slots[0].stream;
  launch_sha256(d_arena, d_off, d_len, nullptr, (uint32_t)n, d_dig, s);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(s));
  std::vector<uint32_t> dig(n * 8);
  CK(cudaMemcpy(dig.data(), d_dig, n * 32, cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < n * 8; i++) { out[4 * i] = (uint8_t)(dig[i] >> 24); out[4 * i + 1] = (uint8_t)(dig[i] >> 16); out[4 * i + 2] = (uint8_t)(dig[i] >> 8); out[4 * i + 3] = (uint8_t)dig[i]; }
  return ZKB_OK;
}

int zkb_rsa_verify_batch(zkb_engine* e, const uint8_t* const* key_der, const size_t* key_len, const uint8_t* digests,
                         const uint8_t* const* sig, const size_t* sig_len, size_t n, uint8_t* ok) {
  if (!e || (n && (!key_der || !key_len || !digests || !sig || !sig_len || !ok))) return ZKB_E_INVALID;
  if (n == 0) return ZKB_OK;
  std::lock_guard<std::mutex> lock(e->run_mu);
  CK(cudaSetDevice(e->device));
  trim_keytab(e);
  ThreadCtx c;
  ThreadRecs tr;
  c.eng = e; c.tr = &tr;
  std::vector<uint32_t> sigw, dig(n * 8);
  std::vector<RsaItem> lists[6];
  for (size_t i = 0; i < n; i++) {
    ok[i] = 0;
    for (int k = 0; k < 8; k++) dig[i * 8 + k] = ((uint32_t)digests[i * 32 + 4 * k] << 24) | ((uint32_t)digests[i * 32 + 4 * k + 1] << 16) | ((uint32_t)digests[i * 32 + 4 * k + 2] << 8) | digests[i * 32 + 4 * k + 3];
    KeyMeta km;
    int32_t id = lookup_key(c, key_der[i], key_len[i], km);
    if (id < 0) { ok[i] = 2; continue; }
    if (sig_len[i] != km.k) continue;
    RsaItem it;
    it.sig_off = (uint32_t)sigw.size(); it.key_id = (uint32_t)id; it.digest_slot = (uint32_t)i; it.cand = (uint32_t)i;
    sigw.resize(sigw.size() + km.limbs_class, 0u);
    for (size_t b = 0; b < sig_len[i]; b++) {
      size_t bi = sig_len[i] - 1 - b;
      sigw[it.sig_off + (bi >> 2)] |= (uint32_t)sig[i][b] << (8 * (bi & 3));
    }
    lists[rsa_list_of(km)].push_back(it);
  }
  cudaStream_t s = e->slots[0].stream;
  int rc = sync_keytab(e, s);
  if (rc) return rc;
  DeviceChunk d;
  uint32_t *d_sig = nullptr, *d_dig = nullptr, *d_flags = nullptr;
  RsaItem* d_items[6] = {nullptr};
  DevTmp tmp;
  CK(tmp.alloc(&d_sig, std::max<size_t>(16, sigw.size() * 4)));
  CK(tmp.alloc(&d_dig, n * 32)); CK(tmp.alloc(&d_flags, n * 4));
  CK(cudaMemcpy(d_sig, sigw.data(), sigw.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_dig, dig.data(), n * 32, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_flags, 0, n * 4));
  d.sig_arena = d_sig; d.digests = d_dig; d.cand_flags = d_flags;
  for (int k = 0; k < 6; k++) {
    d.rsa_n[k] = (uint32_t)lists[k].size();
    if (lists[k].empty()) continue;
    CK(tmp.alloc(&d_items[k], lists[k].size() * sizeof(RsaItem)));
    CK(cudaMemcpy(d_items[k], lists[k].data(), lists[k].size() * sizeof(RsaItem), cudaMemcpyHostToDevice));
    d.rsa_items[k] = d_items[k];
  }
  rc = launch_chunk(e, d, nullptr, s, nullptr, nullptr);
  if (rc) return rc;
  CK(cudaStreamSynchronize(s));
  std::vector<uint32_t> flags(n);
  CK(cudaMemcpy(flags.data(), d_flags, n * 4, cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < n; i++) if (ok[i] != 2) ok[i] = (flags[i] & ZKB_F_RSA_OK) ? 1 : 0;
  return ZKB_OK;
}

int zkb_dfa_scan_batch(zkb_engine* e, const zkb_dfa_view* part, const uint8_t* data, size_t data_len, const uint64_t* off,
                       const uint32_t* len, size_t n, int qp, uint32_t* out) {
  if (!e || !part || (n && (!off || !len || !out))) return ZKB_E_INVALID;
  if (n == 0) return ZKB_OK;
  zkb_regex_set* rs = nullptr;
  int rc = zkb_regex_set_create(e, part, qp ? 0 : 1, qp ? 1 : 0, 1, 1, &rs);
  if (rc) return rc;
  struct SetGuard { zkb_regex_set* p; ~SetGuard() { zkb_regex_set_destroy(p); } } guard{rs};
  std::lock_guard<std::mutex> lock(e->run_mu);
  CK(cudaSetDevice(e->device));
  std::vector<DfaItem> items(n);
  for (size_t i = 0; i < n; i++) {
    if (off[i] + len[i] > data_len) return ZKB_E_INVALID;
    items[i].hay_off = off[i]; items[i].msg = (uint32_t)i; items[i].out_slot = (uint32_t)i;
  }
  uint8_t* d_arena; DfaItem* d_items; uint4* d_out; uint32_t* d_hlen;
  DevTmp tmp;
  CK(tmp.alloc(&d_arena, data_len + 64)); CK(tmp.alloc(&d_items, n * sizeof(DfaItem))); CK(tmp.alloc(&d_out, n * 16));
  CK(tmp.alloc(&d_hlen, n * 4));
  CK(cudaMemcpy(d_hlen, len, n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_arena, data, data_len, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_items, items.data(), n * sizeof(DfaItem), cudaMemcpyHostToDevice));
  CK(cudaMemset(d_out, 0, n * 16));
  const zkb_regex_set::Part& p = rs->parts[0];
  cudaStream_t s = e->slots[0].stream;
  launch_dfa(p.elem, p.direct, d_arena, d_items, (uint32_t)n, d_hlen, p.d_fwd, p.fwd_bytes, p.d_rev, p.rev_bytes, e->smem_optin, qp, d_out, s);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(s));
  CK(cudaMemcpy(out, d_out, n * 16, cudaMemcpyDeviceToHost));
  return ZKB_OK;
}

int zkb_int_pipe_peaks(zkb_engine* e, double out[8]) {
  if (!e || !out) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lock(e->run_mu);
  CK(cudaSetDevice(e->device));
  for (int i = 0; i < 8; i++) out[i] = 0;
  uint32_t* d;
  DevTmp tmp;
  CK(tmp.alloc(&d, 64));
  cudaStream_t s = e->slots[0].stream;
  const int iters = 4096;
  const unsigned grid = (unsigned)e->sm_count * 8, block = 256;
  // thread-instructions per launch.  kind 0: each asm block is 16 PTX mad.lo/madc.hi instructions that
  // ptxas fuses pairwise into 8 IMAD.WIDE.U32(.X) (checked in SASS), 4 blocks per iteration; kinds 1-3:
  // 16 independent IADD3 / LOP3 / SHF per block (the two PTX adds fuse into one 3-input IADD3).
  const double per_thread[4] = {iters * 4.0 * 8.0, iters * 4.0 * 16.0, iters * 4.0 * 16.0, iters * 4.0 * 16.0};
  for (int kind = 0; kind < 4; kind++) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
      CK(cudaEventRecord(e->ev[5], s));
      launch_int_peak(kind, grid, block, d, 12345u + rep, iters, s);
      CK(cudaEventRecord(e->ev[6], s));
      CK(cudaEventSynchronize(e->ev[6]));
      float ms;
      CK(cudaEventElapsedTime(&ms, e->ev[5], e->ev[6]));
      if (rep > 0) best = std::min(best, ms);
    }
    out[kind] = per_thread[kind] * grid * block / (best * 1e-3) / 1e9;
  }
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, e->device);
  out[4] = khz / 1000.0;
  out[5] = e->sm_count;
  return ZKB_OK;
}

}  // extern "C"

#include "engine_multi.inc"
