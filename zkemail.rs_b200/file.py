"""File and JSON helpers — mirror of the reference's helpers/src/file.rs:4-24 (read_email_file, read_json_file) plus
the serde JSON shape of the API structs (core/src/structs.rs:1-75, helpers/src/structs.rs:1-13), so that inputs written
by the Rust helpers (`serde_json`) load here and outputs written here load there.

serde_json's default encodings are used: `Vec<u8>` is an array of numbers, `Option<T>` is `null` or the value,
`usize` a number, structs are objects with the Rust field names."""
from __future__ import annotations

import dataclasses
import json
import typing
from typing import Any, Type, TypeVar

from . import structs as S

T = TypeVar("T")


def read_email_file(path) -> bytes:
    """helpers/src/file.rs:4-12."""
    try:
        with open(path, "rb") as f:
            return f.read()
    except OSError as e:
        raise OSError(f"Failed to open email file: {e}") from e


def to_serde(obj: Any) -> Any:
    """API struct -> the JSON value serde_json::to_value would produce."""
    if dataclasses.is_dataclass(obj):
        return {f.name: to_serde(getattr(obj, f.name)) for f in dataclasses.fields(obj)}
    if isinstance(obj, (bytes, bytearray, memoryview)):
        return list(bytes(obj))
    if isinstance(obj, (list, tuple)):
        return [to_serde(x) for x in obj]
    return obj


def _from(tp: Any, v: Any, where: str) -> Any:
    origin = typing.get_origin(tp)
    if origin is typing.Union:  # Optional[X]
        args = [a for a in typing.get_args(tp) if a is not type(None)]
        return None if v is None else _from(args[0], v, where)
    if v is None:
        raise ValueError(f"{where}: null where a value is required")
    if origin in (list, typing.List):
        if not isinstance(v, list):
            raise ValueError(f"{where}: expected an array")
        (inner,) = typing.get_args(tp)
        return [_from(inner, x, f"{where}[{i}]") for i, x in enumerate(v)]
    if tp is bytes:
        if not isinstance(v, list) or any(not isinstance(x, int) or isinstance(x, bool) or not 0 <= x <= 255 for x in v):
            raise ValueError(f"{where}: expected an array of bytes")
        return bytes(v)
    if tp is str:
        if not isinstance(v, str):
            raise ValueError(f"{where}: expected a string")
        return v
    if tp is int:
        if not isinstance(v, int) or isinstance(v, bool) or v < 0:
            raise ValueError(f"{where}: expected an unsigned integer")
        return v
    if dataclasses.is_dataclass(tp):
        if not isinstance(v, dict):
            raise ValueError(f"{where}: expected an object")
        hints = typing.get_type_hints(tp)
        kw = {}
        for f in dataclasses.fields(tp):
            is_opt = typing.get_origin(hints[f.name]) is typing.Union
            if f.name not in v and not is_opt:
                raise ValueError(f"{where}: missing field `{f.name}`")      # serde: missing field error
            kw[f.name] = _from(hints[f.name], v.get(f.name), f"{where}.{f.name}")
        return tp(**kw)
    raise TypeError(f"{where}: unsupported type {tp!r}")


def from_serde(cls: Type[T], value: Any) -> T:
    """JSON value -> API struct, with serde's strictness on shapes (missing non-Option fields are errors, unknown
    fields are ignored)."""
    return _from(cls, value, cls.__name__)


def read_json_file(path, cls: Type[T]) -> T:
    """helpers/src/file.rs:14-24 `read_json_file::<T>`."""
    try:
        with open(path, "r", encoding="utf-8") as f:
            data = json.load(f)
    except OSError as e:
        raise OSError(f"Failed to open file {path}: {e}") from e
    except json.JSONDecodeError as e:
        raise ValueError(f"Failed to parse JSON from {path}: {e}") from e
    try:
        return from_serde(cls, data)
    except (ValueError, TypeError) as e:
        raise ValueError(f"Failed to parse JSON from {path}: {e}") from e


def write_json_file(path, obj: Any) -> None:
    with open(path, "w", encoding="utf-8") as f:
        json.dump(to_serde(obj), f)


__all__ = ["read_email_file", "read_json_file", "write_json_file", "to_serde", "from_serde", "S"]
