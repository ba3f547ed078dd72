/* C-ABI smoke test: include/zkemail_b200.h must compile as plain C11 and every call below must link against
 * libzkemail_b200.so.  Run by tests/test_c_abi.py.  Without a GPU the engine must refuse to start
 * (ZKB_E_NO_DEVICE: there is no CPU fallback); the host-only entry points must work. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "zkemail_b200.h"

static int fail(const char *what) { fprintf(stderr, "FAIL: %s\n", what); return 1; }

int main(void) {
  if (zkb_abi_version() <= 0) return fail("abi version");
  if (strcmp(zkb_strerror(ZKB_OK), "ok") != 0) return fail("strerror");
  if (sizeof(zkb_result) != 400) return fail("zkb_result layout");

  /* host-only: Solidity-ABI packer round trip (core/src/io.rs) */
  uint8_t h1[32], h2[32];
  memset(h1, 0x11, 32); memset(h2, 0x22, 32);
  zkb_str ext[2] = {{"name", 4}, {"value", 5}};
  zkb_str m[1] = {{"match", 5}};
  zkb_output_view v[2];
  memset(v, 0, sizeof v);
  v[0].from_domain_hash = h1; v[0].public_key_hash = h2; v[0].external_inputs = ext; v[0].n_external_inputs = 2;
  v[1] = v[0]; v[1].matches = m; v[1].n_matches = 1; v[1].with_regex = 1;
  uint8_t *blob = NULL;
  uint64_t offs[3];
  if (zkb_abi_encode_batch(v, 2, 1, &blob, offs) != ZKB_OK) return fail("encode");
  if (offs[0] != 0 || offs[1] != 352 || offs[2] <= offs[1]) return fail("offsets");
  zkb_abi_decoded dec;
  zkb_span *spans = NULL;
  if (zkb_abi_decode(blob + offs[1], (size_t)(offs[2] - offs[1]), &dec, &spans) != ZKB_OK) return fail("decode");
  if (!dec.with_regex || dec.n_external_inputs != 2 || dec.n_matches != 1 || memcmp(dec.from_domain_hash, h1, 32) != 0) return fail("decoded fields");
  if (spans[2].len != 5 || memcmp(blob + offs[1] + spans[2].off, "match", 5) != 0) return fail("decoded match");
  zkb_free(spans); zkb_free(blob);

  /* host-only: regex compiler, canonicaliser, signature listing */
  uint8_t *fwd = NULL, *bwd = NULL; size_t fl = 0, bl = 0; char err[128];
  if (zkb_regex_compile("ab+c", 4, &fwd, &fl, &bwd, &bl, err, sizeof err) != ZKB_OK || fl < ZKB_ZDF_HEADER) return fail("regex compile");
  zkb_free(fwd); zkb_free(bwd);
  const char *mail = "Subject: x\r\n\r\nbody\r\n";
  uint8_t *sig = NULL; size_t sl = 0, ns = 99;
  if (zkb_host_dkim_signatures((const uint8_t *)mail, strlen(mail), 0, &sig, &sl, &ns) != ZKB_OK || ns != 0) return fail("signature listing");
  zkb_free(sig);

  /* the engine itself: a device is mandatory */
  zkb_engine *e = NULL;
  int rc = zkb_engine_create(NULL, &e);
  if (rc == ZKB_OK) {
    zkb_email_view ev;
    memset(&ev, 0, sizeof ev);
    ev.from_domain = "example.com"; ev.from_domain_len = 11;
    ev.raw_email = (const uint8_t *)mail; ev.raw_email_len = strlen(mail);
    ev.key = (const uint8_t *)"\x30\x00"; ev.key_len = 2; ev.key_type = "rsa"; ev.key_type_len = 3;
    zkb_result r;
    if (zkb_verify_one(e, &ev, NULL, NULL, &r) != ZKB_OK || r.status != ZKB_ST_KEY) return fail("verify_one on a bad key");
    zkb_engine_destroy(e);
    printf("ok (device present)\n");
  } else if (rc == ZKB_E_NO_DEVICE) {
    printf("ok (no device: %s)\n", zkb_strerror(rc));
  } else {
    return fail("engine create");
  }
  return 0;
}
