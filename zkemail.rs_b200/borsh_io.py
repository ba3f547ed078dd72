"""borsh (de)serialisation of the input structs — what RISC Zero callers of the reference hold: under the `risc0`
feature every input struct derives BorshSerialize / BorshDeserialize (core/src/structs.rs:1-62, `#[cfg_attr(feature =
"risc0", derive(BorshSerialize, BorshDeserialize))]`), and a guest receives `Email` / `EmailWithRegex` as borsh bytes.

borsh 1.5 encodings used here (the crate's specification): structs = their fields in declaration order; `String` and
`Vec<T>` = u32 little-endian length + elements (`String` bytes must be UTF-8); `Option<T>` = one tag byte 0 / 1 + the
value; `usize` = u64 little-endian.  Decoding is strict like the crate's `try_from_slice`: truncated input, an Option tag
other than 0 / 1, invalid UTF-8 and trailing bytes are errors."""
from __future__ import annotations

import dataclasses
import struct
import typing
from typing import Any, Type, TypeVar

T = TypeVar("T")


class BorshError(ValueError):
    pass


def _enc(tp: Any, v: Any, out: bytearray) -> None:
    origin = typing.get_origin(tp)
    if origin is typing.Union:                       # Option<X>
        (inner,) = [a for a in typing.get_args(tp) if a is not type(None)]
        if v is None:
            out.append(0)
        else:
            out.append(1)
            _enc(inner, v, out)
    elif origin in (list, typing.List):
        (inner,) = typing.get_args(tp)
        out += struct.pack("<I", len(v))
        for x in v:
            _enc(inner, x, out)
    elif tp is bytes:
        out += struct.pack("<I", len(v)) + bytes(v)
    elif tp is str:
        b = v.encode("utf-8")
        out += struct.pack("<I", len(b)) + b
    elif tp is int:                                  # usize
        out += struct.pack("<Q", v)
    elif dataclasses.is_dataclass(tp):
        hints = typing.get_type_hints(tp)
        for f in dataclasses.fields(tp):
            _enc(hints[f.name], getattr(v, f.name), out)
    else:
        raise BorshError(f"no borsh encoding for {tp!r}")


class _Reader:
    def __init__(self, data: bytes):
        self.d, self.at = memoryview(data), 0

    def take(self, n: int) -> bytes:
        if n > len(self.d) - self.at:
            raise BorshError("unexpected end of input")
        b = bytes(self.d[self.at:self.at + n])
        self.at += n
        return b

    def u32(self) -> int:
        return struct.unpack("<I", self.take(4))[0]


def _dec(tp: Any, r: _Reader) -> Any:
    origin = typing.get_origin(tp)
    if origin is typing.Union:
        (inner,) = [a for a in typing.get_args(tp) if a is not type(None)]
        tag = r.take(1)[0]
        if tag == 0:
            return None
        if tag != 1:
            raise BorshError(f"invalid Option tag {tag}")
        return _dec(inner, r)
    if origin in (list, typing.List):
        (inner,) = typing.get_args(tp)
        n = r.u32()
        if n > len(r.d) - r.at:                      # every element takes at least one byte: cheap bound against bombs
            raise BorshError("unexpected end of input")
        return [_dec(inner, r) for _ in range(n)]
    if tp is bytes:
        return r.take(r.u32())
    if tp is str:
        try:
            return r.take(r.u32()).decode("utf-8")
        except UnicodeDecodeError as e:
            raise BorshError("invalid UTF-8 in a String") from e
    if tp is int:
        return struct.unpack("<Q", r.take(8))[0]
    if dataclasses.is_dataclass(tp):
        hints = typing.get_type_hints(tp)
        return tp(**{f.name: _dec(hints[f.name], r) for f in dataclasses.fields(tp)})
    raise BorshError(f"no borsh decoding for {tp!r}")


def to_borsh(obj: Any) -> bytes:
    """API struct -> the bytes `borsh::to_vec(&obj)` produces."""
    out = bytearray()
    _enc(type(obj), obj, out)
    return bytes(out)


def from_borsh(tp: Type[T], data: bytes) -> T:
    """`T::try_from_slice(data)`: strict, the whole slice must be consumed."""
    r = _Reader(data)
    v = _dec(tp, r)
    if r.at != len(data):
        raise BorshError("trailing bytes after the value")
    return v
