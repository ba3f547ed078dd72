// hostbench.cpp — development aid: times the phases of the engine's host front end (dkim_host.hpp)
// on a pool of emails, without a GPU.  Build: g++ -O3 -std=c++17 -shared -fPIC -pthread -o /tmp/zk/libhostbench.so tools/hostbench.cpp
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>
#include "../zkemail.rs_b200/csrc/dkim_host.hpp"
using namespace zkb;
using clk = std::chrono::steady_clock;
extern "C" void hostbench(const zkb_email_view* em, size_t n, int threads, double* out /*8 phases, seconds summed over threads*/) {
  std::vector<std::thread> th;
  std::vector<std::vector<double>> acc(threads, std::vector<double>(8, 0.0));
  for (int t = 0; t < threads; t++) th.emplace_back([&, t] {
    std::vector<HeaderField> hs; DkimSig sig; std::string scratch; std::vector<uint8_t> buf(1 << 20), tmp(4096);
    auto& a = acc[t];
    for (size_t i = n * t / threads; i < n * (t + 1) / threads; i++) {
      const uint8_t* raw = em[i].raw_email; size_t len = em[i].raw_email_len; size_t body_off;
      auto t0 = clk::now();
      parse_headers(raw, len, hs, body_off);
      auto t1 = clk::now();
      for (auto& h : hs) if (ieq_ascii(raw + h.key_off, h.key_len, "DKIM-Signature", 14)) { validate_dkim_header(raw + h.val_off, h.val_len, 1, sig); break; }
      auto t2 = clk::now();
      size_t bl; const uint8_t* b = find_body(raw, len, bl);
      auto t3 = clk::now();
      if (buf.size() < bl + 64) buf.resize(bl + 64);
      size_t cl = canon_body_relaxed(b, bl, buf.data());
      auto t4 = clk::now();
      size_t pl = build_header_preimage(raw, hs, sig, true, buf.data() + ((cl + 63) & ~63), scratch);
      auto t5 = clk::now();
      const Tag* tb = sig.get("b"); tmp.resize(tb->val_len + 4);
      long sl = base64_decode(sig.val(tb), tb->val_len, tmp.data());
      const Tag* tbh = sig.get("bh"); uint8_t bh[48]; base64_decode(sig.val(tbh), tbh->val_len, bh);
      auto t6 = clk::now();
      std::string k((const char*)em[i].key, em[i].key_len); volatile size_t hv = std::hash<std::string>()(k); (void)hv;
      auto t7 = clk::now();
      (void)pl; (void)sl;
      auto d = [](clk::time_point x, clk::time_point y) { return std::chrono::duration<double>(y - x).count(); };
      a[0] += d(t0, t1); a[1] += d(t1, t2); a[2] += d(t2, t3); a[3] += d(t3, t4); a[4] += d(t4, t5); a[5] += d(t5, t6); a[6] += d(t6, t7);
    }
  });
  for (auto& x : th) x.join();
  for (int p = 0; p < 8; p++) { out[p] = 0; for (int t = 0; t < threads; t++) out[p] += acc[t][p]; }
}
