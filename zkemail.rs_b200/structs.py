"""API structs of the hot path — field-for-field mirror of the reference's
core/src/structs.rs:8-75 (PublicKey, DFA, CompiledRegex, RegexInfo, ExternalInput, Email,
EmailWithRegex, EmailVerifierOutput, EmailWithRegexVerifierOutput) and
helpers/src/structs.rs:3-13 (RegexPattern, RegexConfig)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional


@dataclass
class PublicKey:  # core/src/structs.rs:8-11
    key: bytes  # PKCS#1 RSAPublicKey DER (helpers/src/dkim.rs:50,96-102)
    key_type: str  # "rsa" | "ed25519"


@dataclass
class DFA:  # core/src/structs.rs:16-19 ; here: ZDF1 tables (include/zkemail_b200.h)
    fwd: bytes
    bwd: bytes


@dataclass
class CompiledRegex:  # core/src/structs.rs:24-27
    verify_re: DFA
    captures: Optional[List[str]] = None


@dataclass
class RegexInfo:  # core/src/structs.rs:32-35
    header_parts: Optional[List[CompiledRegex]] = None
    body_parts: Optional[List[CompiledRegex]] = None


@dataclass
class ExternalInput:  # core/src/structs.rs:40-44
    name: str
    value: Optional[str] = None
    max_length: int = 0


@dataclass
class Email:  # core/src/structs.rs:49-54
    from_domain: str
    raw_email: bytes
    public_key: PublicKey
    external_inputs: List[ExternalInput] = field(default_factory=list)


@dataclass
class EmailWithRegex:  # core/src/structs.rs:59-62
    email: Email
    regex_info: RegexInfo


@dataclass
class EmailVerifierOutput:  # core/src/structs.rs:65-69
    from_domain_hash: bytes
    public_key_hash: bytes
    external_inputs: List[str]


@dataclass
class EmailWithRegexVerifierOutput:  # core/src/structs.rs:72-75
    email: EmailVerifierOutput
    regex_matches: List[str]


@dataclass
class RegexPattern:  # helpers/src/structs.rs:3-7
    pattern: str
    capture_indices: Optional[List[int]] = None


@dataclass
class RegexConfig:  # helpers/src/structs.rs:9-13
    header_parts: Optional[List[RegexPattern]] = None
    body_parts: Optional[List[RegexPattern]] = None
