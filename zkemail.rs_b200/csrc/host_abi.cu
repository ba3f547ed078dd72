// host_abi.cu — C-ABI entry points of the Solidity-ABI packer (abi_pack.hpp).  Host code only.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <thread>
#include <vector>
#include "abi_pack.hpp"

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

}  // extern "C"
