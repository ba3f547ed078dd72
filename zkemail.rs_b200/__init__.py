"""zkemail.rs_b200 — B200-native batched DKIM e-mail verification (the hot path of
zkemail_core::verify_email / verify_email_with_regex).  See DESIGN.md."""
from .structs import (  # noqa: F401
    DFA, CompiledRegex, Email, EmailVerifierOutput, EmailWithRegex, EmailWithRegexVerifierOutput,
    ExternalInput, PublicKey, RegexConfig, RegexInfo, RegexPattern,
)
from .engine import (  # noqa: F401,E402
    Engine, EngineUnavailable, RegexError, RegexSet, VerificationPanic, compile_regex,
    canonicalize_signed_email, compile_regex_parts, verify_email, verify_email_with_regex,
    OPT_NO_DIRECT, OPT_NO_DEVICE_FRONTEND, OPT_NO_STAGED_FRONTEND, OPT_NO_OVERLAP, OPT_PROFILE, OPT_NO_SQR,
)
from .abi_io import AbiDecodeError, VerificationOutput, abi_decode, abi_encode_batch  # noqa: F401,E402
from .generator import (  # noqa: F401,E402
    GeneratorError, StaticKeys, dkim_signatures, generate_email_inputs, generate_email_inputs_batch,
    generate_email_with_regex_inputs, parse_dkim_key_record, remove_quoted_printable_soft_breaks,
)
from .file import from_serde, read_email_file, read_json_file, to_serde, write_json_file  # noqa: F401,E402
from .multi import Comm, MultiBatch, MultiEngine, MultiRegexSet, plan_shards, records_to_results  # noqa: F401,E402
from .borsh_io import BorshError, from_borsh, to_borsh  # noqa: F401,E402
