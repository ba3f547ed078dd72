"""End-to-end use of the library on files, the way the reference's helpers are used from a host program:

    python examples/verify_eml.py mail.eml example.com keys.json [regex.json]

keys.json maps "selector" (or "domain/selector") to the DKIM TXT record of that selector, e.g.
    {"sel1": "v=DKIM1; k=rsa; p=MIIBIjANBg..."}
regex.json is a RegexConfig as the Rust helpers read it (helpers/src/structs.rs:9-13).

Steps (reference call sites in brackets): read the message [helpers/src/file.rs:4], generate the Email /
EmailWithRegex inputs offline [helpers/src/generator.rs:11-87], verify on the GPU [core/src/circuits.rs:9,31],
ABI-encode the output [core/src/io.rs:35].  Needs a B200; there is no CPU fallback."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkemail_rs_b200 as z  # noqa: E402
from zkemail_rs_b200.structs import RegexConfig  # noqa: E402


def main(argv):
    if len(argv) < 4:
        print(__doc__)
        return 2
    raw = z.read_email_file(argv[1])
    domain = argv[2]
    table = json.load(open(argv[3]))
    keys = z.StaticKeys({((k.split("/", 1)[0], k.split("/", 1)[1]) if "/" in k else (domain, k)): v for k, v in table.items()})
    if len(argv) > 4:
        cfg = z.read_json_file(argv[4], RegexConfig)
        inp = z.generate_email_with_regex_inputs(domain, raw, cfg, keys)
        out = z.verify_email_with_regex(inp)
    else:
        inp = z.generate_email_inputs(domain, raw, keys)
        out = z.verify_email(inp)
    print(json.dumps(z.to_serde(out)))
    print("abi:", z.VerificationOutput.from_output(out).abi_encode().hex())
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
