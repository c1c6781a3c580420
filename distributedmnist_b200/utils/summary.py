"""Scalar summaries: a TensorBoard event file (+ a JSON-lines mirror) per directory.

reference: ``tf.summary.FileWriter(eval_dir)`` with the two evaluator scalars
``Validation Accuracy`` / ``Validation Loss`` (src/nn_eval.py:107-110,133-134) and
the chief's merged training summary (src/distributed_train.py:225,382-390).

``events.out.tfevents.<time>.<host>`` is written in TensorBoard's own format -- TFRecord framing (length, masked
CRC-32C of the length, payload, masked CRC-32C of the payload) around hand-encoded ``Event`` protobufs
(``wall_time`` = field 1 double, ``step`` = 2 int64, ``file_version`` = 3 string, ``summary`` = 5 message with repeated
``Value{tag = 1 string, simple_value = 2 float}``) -- so ``tensorboard --logdir`` reads it unchanged; no TensorFlow or
protobuf package is involved.  The JSON-lines mirror keeps the files greppable (tools/benchmark.py-style plots).
"""
from __future__ import annotations

import json
import os
import socket
import struct
import time
from typing import Dict, Iterator, List, Tuple

# ---- CRC-32C (Castagnoli), table driven --------------------------------------------------------------------------------
_CRC_TABLE: List[int] = []


def _crc_table() -> List[int]:
    if not _CRC_TABLE:
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            _CRC_TABLE.append(c)
    return _CRC_TABLE


def crc32c(data: bytes) -> int:
    t = _crc_table()
    c = 0xFFFFFFFF
    for b in data:
        c = t[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _masked_crc(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---- minimal protobuf encoding ----------------------------------------------------------------------------------------------
def _varint(n: int) -> bytes:
    out = bytearray()
    n &= (1 << 64) - 1
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _field_bytes(num: int, payload: bytes) -> bytes:
    return _varint((num << 3) | 2) + _varint(len(payload)) + payload


def encode_event(wall_time: float, step: int = 0, scalars: Dict[str, float] = None, file_version: str = "") -> bytes:
    ev = _varint((1 << 3) | 1) + struct.pack("<d", float(wall_time))            # wall_time: double
    if step:
        ev += _varint((2 << 3) | 0) + _varint(int(step))                        # step: int64
    if file_version:
        ev += _field_bytes(3, file_version.encode())
    if scalars:
        summary = b""
        for tag, val in scalars.items():
            value = _field_bytes(1, tag.encode()) + _varint((2 << 3) | 5) + struct.pack("<f", float(val))
            summary += _field_bytes(1, value)
        ev += _field_bytes(5, summary)
    return ev


def _record(payload: bytes) -> bytes:
    hdr = struct.pack("<Q", len(payload))
    return hdr + struct.pack("<I", _masked_crc(hdr)) + payload + struct.pack("<I", _masked_crc(payload))


class SummaryWriter:
    def __init__(self, logdir: str, filename: str = "events.out.dmnist.jsonl"):
        os.makedirs(logdir, exist_ok=True)
        self.path = os.path.join(logdir, filename)
        self._f = open(self.path, "a", buffering=1)
        self.tfevents_path = os.path.join(logdir, "events.out.tfevents.%010d.%s" % (int(time.time()), socket.gethostname()))
        self._tf = open(self.tfevents_path, "ab")
        self._tf.write(_record(encode_event(time.time(), file_version="brain.Event:2")))
        self._tf.flush()

    def add_scalars(self, scalars: Dict[str, float], global_step: int) -> None:
        now = time.time()
        self._f.write(json.dumps({"wall_time": now, "step": int(global_step),
                                  "scalars": {k: float(v) for k, v in scalars.items()}}) + "\n")
        self._tf.write(_record(encode_event(now, int(global_step), scalars)))
        self._tf.flush()

    def close(self) -> None:
        self._f.close()
        self._tf.close()


def read_events(path: str):
    with open(path) as f:
        return [json.loads(line) for line in f if line.strip()]


# ---- reader for the tfevents file (tests, tools) ------------------------------------------------------------------------------
def _read_varint(buf: bytes, pos: int) -> Tuple[int, int]:
    n = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        n |= (b & 0x7F) << shift
        shift += 7
        if not b & 0x80:
            return n, pos


def _parse_fields(buf: bytes) -> Iterator[Tuple[int, int, object]]:
    pos = 0
    while pos < len(buf):
        key, pos = _read_varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 1:
            v, pos = buf[pos:pos + 8], pos + 8
        elif wt == 5:
            v, pos = buf[pos:pos + 4], pos + 4
        elif wt == 2:
            n, pos = _read_varint(buf, pos)
            v, pos = buf[pos:pos + n], pos + n
        else:
            raise ValueError("unsupported wire type %d" % wt)
        yield num, wt, v


def read_tfevents(path: str) -> List[Dict]:
    """Decode an event file written by :class:`SummaryWriter` (checks both CRCs of every record)."""
    out: List[Dict] = []
    with open(path, "rb") as f:
        data = f.read()
    pos = 0
    while pos < len(data):
        hdr = data[pos:pos + 8]
        (n,) = struct.unpack("<Q", hdr)
        (hcrc,) = struct.unpack("<I", data[pos + 8:pos + 12])
        payload = data[pos + 12:pos + 12 + n]
        (pcrc,) = struct.unpack("<I", data[pos + 12 + n:pos + 16 + n])
        if hcrc != _masked_crc(hdr) or pcrc != _masked_crc(payload):
            raise ValueError("corrupt record at byte %d" % pos)
        pos += 16 + n
        ev: Dict = {"step": 0, "scalars": {}}
        for num, _wt, v in _parse_fields(payload):
            if num == 1:
                ev["wall_time"] = struct.unpack("<d", v)[0]
            elif num == 2:
                ev["step"] = int(v)
            elif num == 3:
                ev["file_version"] = v.decode()
            elif num == 5:
                for n2, _w2, val in _parse_fields(v):
                    if n2 != 1:
                        continue
                    tag, sv = None, None
                    for n3, _w3, x in _parse_fields(val):
                        if n3 == 1:
                            tag = x.decode()
                        elif n3 == 2:
                            sv = struct.unpack("<f", x)[0]
                    ev["scalars"][tag] = sv
        out.append(ev)
    return out
