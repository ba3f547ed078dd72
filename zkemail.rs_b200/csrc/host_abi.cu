// host_abi.cu — C-ABI entry points of the Solidity-ABI packer (abi_pack.hpp).  Host code only.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <thread>
#include <vector>
#include "abi_pack.hpp"
#include "dkim_host.hpp"
#include "ra_wire.hpp"

extern "C" {

int zkb_abi_encode_batch(const zkb_output_view* outs, size_t n, int threads, uint8_t** blob, uint64_t* offsets) {
  if ((!outs && n) || !blob || !offsets) return ZKB_E_INVALID;
  *blob = nullptr;
  for (size_t i = 0; i < n; i++) {
    const zkb_output_view& o = outs[i];
    if (!o.from_domain_hash || !o.public_key_hash || (o.n_external_inputs && !o.external_inputs) ||
        (o.with_regex && o.n_matches && !o.matches))
      return ZKB_E_INVALID;
  }
  uint64_t total = 0;
  for (size_t i = 0; i < n; i++) { offsets[i] = total; total += zkb::abi::encoded_size(outs[i]); }
  offsets[n] = total;
  uint8_t* p = (uint8_t*)malloc(total ? (size_t)total : 1);
  if (!p) return ZKB_E_NOMEM;
  int T = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  T = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(T, 1), n / 4096 + 1));
  if (T == 1) {
    for (size_t i = 0; i < n; i++) zkb::abi::encode(outs[i], p + offsets[i]);
  } else {
    std::atomic<size_t> next{0};
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++)
      th.emplace_back([&] {
        for (;;) {
          const size_t lo = next.fetch_add(1024);
          if (lo >= n) break;
          const size_t hi = std::min(n, lo + 1024);
          for (size_t i = lo; i < hi; i++) zkb::abi::encode(outs[i], p + offsets[i]);
        }
      });
    for (auto& x : th) x.join();
  }
  *blob = p;
  return ZKB_OK;
}

int zkb_abi_decode(const uint8_t* data, size_t len, zkb_abi_decoded* out, zkb_span** spans) {
  if ((!data && len) || !out || !spans) return ZKB_E_INVALID;
  *spans = nullptr;
  std::vector<zkb_span> sp;
  if (!zkb::abi::decode(data, len, *out, sp)) return ZKB_E_INVALID;
  zkb_span* q = (zkb_span*)malloc(std::max<size_t>(1, sp.size()) * sizeof(zkb_span));
  if (!q) return ZKB_E_NOMEM;
  std::copy(sp.begin(), sp.end(), q);
  *spans = q;
  return ZKB_OK;
}

// helpers/src/generator.rs:16-31: get_all_headers("DKIM-Signature") then validate_header on each, in message order
int zkb_host_dkim_signatures(const uint8_t* raw, size_t n, int64_t now_unix, uint8_t** out, size_t* out_len, size_t* n_sigs) {
  using namespace zkb;
  if ((!raw && n) || !out || !out_len || !n_sigs) return ZKB_E_INVALID;
  *out = nullptr; *out_len = 0; *n_sigs = 0;
  std::vector<HeaderField> hs;
  size_t body_off = 0;
  if (!parse_headers(raw, n, hs, body_off)) return ZKB_E_INVALID;
  std::vector<uint8_t> buf;
  auto put32 = [&](uint32_t v) { for (int i = 0; i < 4; i++) buf.push_back((uint8_t)(v >> (8 * i))); };
  DkimSig sig;
  size_t cnt = 0;
  for (const HeaderField& h : hs) {
    if (!ieq_ascii(raw + h.key_off, h.key_len, "DKIM-Signature", 14)) continue;
    cnt++;
    if (validate_dkim_header(raw + h.val_off, h.val_len, now_unix, sig) != ZKB_DKIM_PASS) { put32(0); put32(0); put32(0); continue; }
    const Tag* td = sig.get("d");
    const Tag* ts = sig.get("s");
    put32(1); put32(td->val_len); put32(ts->val_len);
    buf.insert(buf.end(), sig.val(td), sig.val(td) + td->val_len);
    buf.insert(buf.end(), sig.val(ts), sig.val(ts) + ts->val_len);
  }
  uint8_t* p = (uint8_t*)malloc(std::max<size_t>(1, buf.size()));
  if (!p) return ZKB_E_NOMEM;
  if (!buf.empty()) memcpy(p, buf.data(), buf.size());
  *out = p; *out_len = buf.size(); *n_sigs = cnt;
  return ZKB_OK;
}

int zkb_regex_automata_to_zdf(const uint8_t* wire, size_t wire_len, int reverse, uint8_t** zdf, size_t* zdf_len) {
  if (!wire || !zdf || !zdf_len) return ZKB_E_INVALID;
  *zdf = nullptr; *zdf_len = 0;
  std::vector<uint8_t> z;
  if (!zkb::ra::to_zdf(wire, wire_len, reverse != 0, z)) return ZKB_E_REGEX;
  uint8_t* p = (uint8_t*)malloc(z.size());
  if (!p) return ZKB_E_NOMEM;
  memcpy(p, z.data(), z.size());
  *zdf = p; *zdf_len = z.size();
  return ZKB_OK;
}

}  // extern "C"
