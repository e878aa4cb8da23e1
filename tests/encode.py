"""Test-side encoder for node trees, written from the format description alone (reference
src/FastNoise/Metadata.cpp:88-158,161-308): builds the encoded string of an arbitrary tree so the same tree can be
decoded by the product, by oracle/_ref and by the C port. Member tables come from tests/golden/metadata.json."""
import base64
import json
import os
import struct

_META = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metadata.json")))
BY_NAME = {m["name"]: m for m in _META}


class N:
    """N("FractalFBm", {"Octaves": 5}, [N("Perlin")], {"Gain": 0.6, "Weighted Strength": N("Value")})
    variables / hybrids are keyed by member name, or (name, dim) for per-dimension members."""

    def __init__(self, name, variables=None, lookups=None, hybrids=None):
        self.meta = BY_NAME[name]
        self.variables = variables or {}
        self.lookups = lookups or []
        self.hybrids = hybrids or {}

    @staticmethod
    def _index(members, key):
        name, dim = key if isinstance(key, tuple) else (key, -1)
        for i, m in enumerate(members):
            if m["name"] == name and m["dim"] == dim:
                return i
        raise KeyError(key)

    def _emit(self, out, done):
        if id(self) in done:                      # back-reference to an already completed node
            out.append(bytes([0xFF]) + struct.pack("<H", done[id(self)]))
            return
        out.append(bytes([self.meta["id"]]))
        for key, value in self.variables.items():
            i = self._index(self.meta["variables"], key)
            raw = struct.pack("<f", value) if self.meta["variables"][i]["type"] == 0 else struct.pack("<i", value)
            out.append(bytes([0 | (i << 3)]) + raw)
        if self.meta["lookups"]:
            assert len(self.lookups) == len(self.meta["lookups"]), self.meta["name"]
            out.append(bytes([1 | (len(self.lookups) << 3)]))
            for child in self.lookups:
                child._emit(out, done)
        for key, value in self.hybrids.items():
            i = self._index(self.meta["hybrids"], key)
            if isinstance(value, N):
                out.append(bytes([2 | (i << 3)]))
                value._emit(out, done)
            else:
                out.append(bytes([3 | (i << 3)]) + struct.pack("<f", value))
        out.append(b"END")
        done[id(self)] = len(done)

    def encode(self):
        parts = []
        self._emit(parts, {})
        raw = bytearray()
        pending = 0
        for part in parts:                          # fold runs of node ends into one byte (5-bit counter)
            if part == b"END":
                pending += 1
                if pending == 32:
                    raw.append(4 | (31 << 3))
                    pending = 0
                continue
            if pending:
                raw.append(4 | ((pending - 1) << 3))
                pending = 0
            raw += part
        if pending:
            raw.append(4 | ((pending - 1) << 3))
        return base64.b64encode(bytes(raw)).decode()


def default_tree(node_id, hybrid_nodes=False):
    """Node `node_id` with required inputs wired the way tests/trees.py does, as an encodable description."""
    meta = _META[node_id]
    lookups = []
    for i in range(len(meta["lookups"])):
        if node_id in (19, 20):
            lookups.append(N("DomainWarpGradient", {"Feature Scale": 53.0}, [N("Simplex", {"Feature Scale": 41.0})]))
        else:
            lookups.append(N("Simplex" if i == 0 else "Perlin", {"Feature Scale": 37.0 + 11 * i, **({"Seed Offset": i} if i else {})}))
    hybrids = {}
    if hybrid_nodes:
        for i, h in enumerate(meta["hybrids"]):
            key = (h["name"], h["dim"]) if h["dim"] >= 0 else h["name"]
            hybrids[key] = N("Perlin", {"Feature Scale": 19.0 + 7 * i, "Seed Offset": 5 + i, "Output Min": 0.2, "Output Max": 0.9})
    return N(meta["name"], {}, lookups, hybrids)
