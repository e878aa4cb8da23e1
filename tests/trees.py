"""Builds the same node tree inside any library that exports the fn* C ABI (the product or an oracle/_ref build),
using only the reflection entry points -- the way tests/FastNoiseBenchmark.cpp:14-46 of the reference auto-wires
nodes -- so a parity test needs no encoded string for every node type."""

WARP_IDS = (16, 17, 18)          # DomainWarpSimplex, DomainWarpSuperSimplex, DomainWarpGradient
WARP_FRACTAL_IDS = (19, 20)
UNSUPPORTED_IDS = ()             # every node type is lowered


class Tree:
    """Owns the handles it creates in `lib`."""

    def __init__(self, lib):
        self.lib = lib
        self.handles = []

    def new(self, node_id):
        h = self.lib.fnNewFromMetadata(node_id, 0xFFFFFFFF)
        assert h, node_id
        self.handles.append(h)
        return h

    def var_index(self, h, name, dim=-1):
        node_id = self.lib.fnGetMetadataID(h)
        for i in range(self.lib.fnGetMetadataVariableCount(node_id)):
            if self.lib.fnGetMetadataVariableName(node_id, i).decode() == name and self.lib.fnGetMetadataVariableDimensionIdx(node_id, i) == dim:
                return i
        raise KeyError(name)

    def set_float(self, h, name, value, dim=-1):
        assert self.lib.fnSetVariableFloat(h, self.var_index(h, name, dim), value)

    def set_int(self, h, name, value, dim=-1):
        assert self.lib.fnSetVariableIntEnum(h, self.var_index(h, name, dim), value)

    def leaf(self, node_id=6, scale=37.0, seed_offset=0):
        h = self.new(node_id)
        self.set_float(h, "Feature Scale", scale)
        if seed_offset:
            self.set_int(h, "Seed Offset", seed_offset)
        return h

    def small_range_leaf(self, node_id=8, scale=23.0, lo=0.2, hi=0.9, seed_offset=3):
        """A generator usable as a hybrid input: positive, bounded away from zero."""
        h = self.leaf(node_id, scale, seed_offset)
        self.set_float(h, "Output Min", lo)
        self.set_float(h, "Output Max", hi)
        return h

    def wire(self, node_id, hybrid_nodes=False):
        """Node `node_id` with every required input connected (and optionally every hybrid driven by a generator)."""
        lib = self.lib
        h = self.new(node_id)
        for i in range(lib.fnGetMetadataNodeLookupCount(node_id)):
            if node_id in WARP_FRACTAL_IDS:
                warp = self.new(18)
                assert lib.fnSetNodeLookup(warp, 0, self.leaf(6, 41.0))
                self.set_float(warp, "Feature Scale", 53.0)
                src = warp
            else:
                src = self.leaf(6 if i == 0 else 8, 37.0 + 11 * i, i)
            assert lib.fnSetNodeLookup(h, i, src), (node_id, i)
        if hybrid_nodes:
            for i in range(lib.fnGetMetadataHybridCount(node_id)):
                assert lib.fnSetHybridNodeLookup(h, i, self.small_range_leaf(8, 19.0 + 7 * i, 0.2, 0.9, 5 + i))
        return h

    def close(self):
        for h in self.handles:
            self.lib.fnDeleteNodeRef(h)
        self.handles = []
