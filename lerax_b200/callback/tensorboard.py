"""TensorBoardBackend (reference: callback/logging/tensorboard.py) without tensorboardX.

The reference delegates to `tensorboardX.SummaryWriter`, which is not installable here; this backend writes the
TensorBoard event-file format itself -- TFRecord framing (length, masked CRC-32C of the length, payload, masked
CRC-32C of the payload) around hand-encoded `Event` / `Summary` protobuf messages -- for what `LoggingCallback`
emits: scalars (`simple_value`) and the hyper-parameter table (a `text` plugin summary).  Same constructor and
directory layout (`log_dir / name / events.out.tfevents.*`); videos are not supported (no renderers in scope)."""

from __future__ import annotations

import os
import socket
import struct
import time
from pathlib import Path

from .logging import AbstractLoggingBackend

# ---- CRC-32C (Castagnoli), table driven
_POLY = 0x82F63B78
_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ _POLY if _c & 1 else _c >> 1
    _TABLE.append(_c)


def crc32c(data: bytes) -> int:
    c = 0xFFFFFFFF
    for b in data:
        c = _TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---- protobuf wire format (only what Event / Summary need)
def _varint(n: int) -> bytes:
    n &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _key(field: int, wire: int) -> bytes:
    return _varint((field << 3) | wire)


def _bytes_field(field: int, payload: bytes) -> bytes:
    return _key(field, 2) + _varint(len(payload)) + payload


def encode_scalar_event(tag: str, value: float, step: int, wall_time: float) -> bytes:
    """Event{wall_time=1, step=2, summary=5{value=1{tag=1, simple_value=2}}}."""
    val = _bytes_field(1, tag.encode()) + _key(2, 5) + struct.pack("<f", float(value))
    summary = _bytes_field(1, val)
    return _key(1, 1) + struct.pack("<d", wall_time) + _key(2, 0) + _varint(int(step)) + _bytes_field(5, summary)


def encode_text_event(tag: str, text: str, step: int, wall_time: float) -> bytes:
    """A `text` plugin summary: Value{tag=1, metadata=9{plugin_data=1{plugin_name=1}}, tensor=8{dtype=1:DT_STRING,
    string_val=8}}."""
    meta = _bytes_field(1, _bytes_field(1, b"text"))
    tensor = _key(1, 0) + _varint(7) + _bytes_field(8, text.encode())
    val = _bytes_field(1, tag.encode()) + _bytes_field(9, meta) + _bytes_field(8, tensor)
    summary = _bytes_field(1, val)
    return _key(1, 1) + struct.pack("<d", wall_time) + _key(2, 0) + _varint(int(step)) + _bytes_field(5, summary)


def encode_version_event(wall_time: float) -> bytes:
    return _key(1, 1) + struct.pack("<d", wall_time) + _bytes_field(3, b"brain.Event:2")


def tfrecord(payload: bytes) -> bytes:
    head = struct.pack("<Q", len(payload))
    return head + struct.pack("<I", masked_crc32c(head)) + payload + struct.pack("<I", masked_crc32c(payload))


class TensorBoardBackend(AbstractLoggingBackend):
    def __init__(self, log_dir: str | Path = "logs") -> None:
        self._log_dir = Path(log_dir)
        self._file = None
        self.path = None

    def open(self, name: str) -> None:
        d = self._log_dir / name
        d.mkdir(parents=True, exist_ok=True)
        self.path = d / f"events.out.tfevents.{int(time.time())}.{socket.gethostname()}.{os.getpid()}.lerax_b200"
        self._file = open(self.path, "ab")
        self._file.write(tfrecord(encode_version_event(time.time())))
        self._file.flush()

    def log_hparams(self, hparams: dict) -> None:
        rows = "\n".join(f"| {k} | {v} |" for k, v in sorted(hparams.items()))
        self._write(encode_text_event("hparams", f"| key | value |\n|---|---|\n{rows}", 0, time.time()))

    def log_scalars(self, scalars: dict, step: int) -> None:
        now = time.time()
        for tag, value in scalars.items():
            self._write(encode_scalar_event(tag, float(value), step, now), flush=False)
        self._file.flush()

    def _write(self, event: bytes, flush: bool = True) -> None:
        if self._file is None:
            raise RuntimeError("TensorBoardBackend.open(name) has not been called")
        self._file.write(tfrecord(event))
        if flush:
            self._file.flush()

    def close(self) -> None:
        if self._file is not None:
            self._file.close()
            self._file = None
