"""MAVLink 2 codec for the two messages at the edges of the MPC node (SURVEY.md section 8b "wire formats", 8f-4).

The reference node receives the vehicle state as ``MPC_FULL_STATE`` (id 367) straight from ``mavlink-router``
(scripts/router_sitl.conf:19; fields read at sde_control.py:246-249) and answers with ``MPC_MOTORS_CMD`` (id 368)
(scripts/router_sitl.conf:18; fields written at sde_control.py:607-613), through ``pymavlink`` and a custom dialect
built into the author's PX4 / mavlink forks.  Neither pymavlink nor the dialect XML is in the reference tree, so this
module implements the PUBLIC MAVLink 2 framing itself —

* frame: ``0xFD, len, incompat, compat, seq, sysid, compid, msgid[3], payload, crc[2]`` (little endian),
* checksum: CRC-16/MCRF4XX (X.25) over everything after the magic byte plus the message's CRC_EXTRA,
* CRC_EXTRA: derived from the message definition exactly as mavgen does (message name, then ``type name`` of every
  non-extension field in wire order, array lengths as a raw byte),
* wire order: fields sorted by element size, largest first, stable; trailing zero bytes of the payload are truncated,

(checked in tests/test_host.py against the published CRC_EXTRA values of HEARTBEAT / ATTITUDE / LOCAL_POSITION_NED and
the CRC-16/MCRF4XX check value) — and states the two message definitions from what the reference reads and writes:
field NAMES are the reference's; field TYPES and ORDER are the natural ones (uint64 timestamp, float32 values) and are
the one thing that cannot be pinned without the dialect XML.  A deployment pins them by passing its own
``MessageDef`` built from that XML; nothing else changes.
"""
from __future__ import annotations

import dataclasses
import struct

import numpy as np

_TYPES = {  # MAVLink type -> (struct code, size)
    "uint64_t": ("Q", 8), "int64_t": ("q", 8), "double": ("d", 8), "uint32_t": ("I", 4), "int32_t": ("i", 4), "float": ("f", 4),
    "uint16_t": ("H", 2), "int16_t": ("h", 2), "uint8_t": ("B", 1), "int8_t": ("b", 1), "char": ("c", 1),
}


def x25_crc(data: bytes, crc: int = 0xFFFF) -> int:
    """CRC-16/MCRF4XX as MAVLink accumulates it (init 0xFFFF, reflected polynomial 0x8408, no final xor)."""
    for b in data:
        tmp = (b ^ (crc & 0xFF)) & 0xFF
        tmp = (tmp ^ (tmp << 4)) & 0xFF
        crc = ((crc >> 8) ^ (tmp << 8) ^ (tmp << 3) ^ (tmp >> 4)) & 0xFFFF
    return crc


@dataclasses.dataclass(frozen=True)
class Field:
    name: str
    type: str
    length: int = 0          # 0: scalar; n > 0: array of n

    @property
    def size(self) -> int:
        return _TYPES[self.type][1]

    @property
    def count(self) -> int:
        return self.length or 1


@dataclasses.dataclass(frozen=True)
class MessageDef:
    name: str
    msgid: int
    fields: tuple            # declaration order (what the XML lists)

    @property
    def wire_fields(self) -> tuple:
        return tuple(sorted(self.fields, key=lambda f: -f.size))     # sorted() is stable: MAVLink's reordering rule

    @property
    def crc_extra(self) -> int:
        crc = x25_crc((self.name + " ").encode())
        for f in self.wire_fields:
            crc = x25_crc((f.type + " ").encode(), crc)
            crc = x25_crc((f.name + " ").encode(), crc)
            if f.length:
                crc = x25_crc(bytes([f.length]), crc)
        return (crc & 0xFF) ^ (crc >> 8)

    @property
    def payload_size(self) -> int:
        return sum(f.size * f.count for f in self.fields)

    def pack_payload(self, values: dict) -> bytes:
        out = b""
        for f in self.wire_fields:
            v = values[f.name]
            code = _TYPES[f.type][0]
            if f.length:
                v = list(np.asarray(v).reshape(-1))
                if len(v) != f.length:
                    raise ValueError(f"{self.name}.{f.name}: expected {f.length} values, got {len(v)}")
                out += struct.pack("<" + code * f.length, *[int(a) if code in "QqIiHhBb" else float(a) for a in v])
            else:
                out += struct.pack("<" + code, int(v) if code in "QqIiHhBb" else float(v))
        return out

    def unpack_payload(self, payload: bytes) -> dict:
        payload = payload + b"\x00" * (self.payload_size - len(payload))     # undo the zero truncation
        out, off = {}, 0
        for f in self.wire_fields:
            code = _TYPES[f.type][0]
            vals = struct.unpack_from("<" + code * f.count, payload, off)
            off += f.size * f.count
            out[f.name] = np.asarray(vals, np.float32 if code in "fd" else np.int64) if f.length else vals[0]
        return out


_F = lambda names: tuple(Field(n, "float") for n in names)

# sde_control.py:246-249 reads time_usec, x, y, z, vx, vy, vz, qw, qx, qy, qz, wx, wy, wz; the PlotJuggler layout
# (launch/pj_setpoint_layout.xml:55-78) additionally plots m1..m4 (the motor outputs PX4 reports back)
MPC_FULL_STATE = MessageDef("MPC_FULL_STATE", 367, (Field("time_usec", "uint64_t"),) + _F(
    ["x", "y", "z", "vx", "vy", "vz", "qw", "qx", "qy", "qz", "wx", "wy", "wz", "m1", "m2", "m3", "m4"]))
# sde_control.py:607-613 writes time_usec, motor_val_des[6], thrust_and_angrate_des[4], mpc_on, weight_motors
MPC_MOTORS_CMD = MessageDef("MPC_MOTORS_CMD", 368, (
    Field("time_usec", "uint64_t"), Field("motor_val_des", "float", 6), Field("thrust_and_angrate_des", "float", 4),
    Field("mpc_on", "uint8_t"), Field("weight_motors", "float")))
DIALECT = {m.msgid: m for m in (MPC_FULL_STATE, MPC_MOTORS_CMD)}


def encode(msg: MessageDef, values: dict, seq: int = 0, sysid: int = 1, compid: int = 1) -> bytes:
    """One MAVLink 2 frame (unsigned)."""
    payload = msg.pack_payload(values).rstrip(b"\x00") or b"\x00"          # v2: truncate trailing zeros, keep >= 1 byte
    head = struct.pack("<BBBBBB", len(payload), 0, 0, seq & 0xFF, sysid, compid) + struct.pack("<I", msg.msgid)[:3]
    crc = x25_crc(bytes([msg.crc_extra]), x25_crc(head + payload))
    return b"\xfd" + head + payload + struct.pack("<H", crc)


class Decoder:
    """Incremental frame parser: feed bytes as they arrive, get (MessageDef, values, header) tuples.  Frames of unknown
    ids are skipped, frames with a bad checksum are dropped and counted; resynchronises on the next 0xFD."""

    def __init__(self, dialect: dict | None = None):
        self.dialect, self.buf, self.bad_crc, self.skipped = dict(dialect or DIALECT), b"", 0, 0

    def feed(self, data: bytes) -> list:
        self.buf += data
        out = []
        while True:
            i = self.buf.find(b"\xfd")
            if i < 0:
                self.buf = b""
                break
            self.buf = self.buf[i:]
            if len(self.buf) < 12:
                break
            n, incompat = self.buf[1], self.buf[2]
            total = 10 + n + 2 + (13 if incompat & 1 else 0)
            if len(self.buf) < total:
                break
            frame, msgid = self.buf[:total], int.from_bytes(self.buf[7:10], "little")
            msg = self.dialect.get(msgid)
            if msg is None:
                self.skipped += 1
                self.buf = self.buf[total:]
                continue
            crc = x25_crc(bytes([msg.crc_extra]), x25_crc(frame[1:10 + n]))
            if crc != int.from_bytes(frame[10 + n:12 + n], "little"):
                self.bad_crc += 1
                self.buf = self.buf[1:]          # not a frame start after all (or corrupted): look for the next magic
                continue
            out.append((msg, msg.unpack_payload(frame[10:10 + n]), dict(seq=frame[4], sysid=frame[5], compid=frame[6])))
            self.buf = self.buf[total:]
        return out


def state_from_full_state(values: dict) -> tuple[np.ndarray, int]:
    """MPC_FULL_STATE -> (x[13] float32 in the order the node builds it, sde_control.py:246-247; time_usec)."""
    x = np.array([values[k] for k in ("x", "y", "z", "vx", "vy", "vz", "qw", "qx", "qy", "qz", "wx", "wy", "wz")], np.float32)
    return x, int(values["time_usec"])


def motors_cmd_values(time_usec: int, uopt, wopt, mpc_on: int, weight_motors: float) -> dict:
    """Arguments of mpc_motors_cmd_send (sde_control.py:607-613): motor values zero-padded to 6 (the node's _uopt)."""
    u = np.zeros(6, np.float32)
    uo = np.asarray(uopt, np.float32).reshape(-1)
    u[: len(uo)] = uo
    return dict(time_usec=int(time_usec), motor_val_des=u, thrust_and_angrate_des=np.asarray(wopt, np.float32).reshape(4),
                mpc_on=int(mpc_on), weight_motors=float(weight_motors))
