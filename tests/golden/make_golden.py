"""Regenerates tests/golden/*.npz and metadata.json from the STRICT oracle builds of the reference
(oracle/_ref/libfn2ref_strict_avx2.so / _avx512.so = the reference's own node sources over oracle/shim, generic / table
gradient flavour -> outputs.npz / outputs_tables.npz). Run in the build
container, where /root/reference exists:   make -C oracle && python tests/golden/make_golden.py
The reference's own tests hold no expected values, so these vectors pin the oracle's
semantics, not Auburn's binaries -- "parity unpinned" upstream, see DESIGN.md.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests import helpers  # noqa: E402
from tests.cases import GOLDEN_CASES, run_case  # noqa: E402


def dump_metadata(lib):
    out = []
    for node_id in range(lib.fnGetMetadataCount()):
        node = {"id": node_id, "name": lib.fnGetMetadataName(node_id).decode(), "groups": [], "variables": [], "lookups": [], "hybrids": []}
        for g in range(lib.fnGetMetadataGroupCount(node_id)):
            node["groups"].append(lib.fnGetMetadataGroupName(node_id, g).decode())
        for i in range(lib.fnGetMetadataVariableCount(node_id)):
            t = lib.fnGetMetadataVariableType(node_id, i)
            v = {"name": lib.fnGetMetadataVariableName(node_id, i).decode(), "type": t, "dim": lib.fnGetMetadataVariableDimensionIdx(node_id, i),
                 "default_bits": int(np.float32(lib.fnGetMetadataVariableDefaultFloat(node_id, i)).view(np.int32)) if t == 0 else lib.fnGetMetadataVariableDefaultIntEnum(node_id, i),
                 "enums": [lib.fnGetMetadataEnumName(node_id, i, e).decode() for e in range(lib.fnGetMetadataEnumCount(node_id, i))]}
            if t == 1:
                v["min"] = int(np.float32(lib.fnGetMetadataVariableMinFloat(node_id, i)).view(np.int32))
                v["max"] = int(np.float32(lib.fnGetMetadataVariableMaxFloat(node_id, i)).view(np.int32))
            node["variables"].append(v)
        for i in range(lib.fnGetMetadataNodeLookupCount(node_id)):
            node["lookups"].append({"name": lib.fnGetMetadataNodeLookupName(node_id, i).decode(), "dim": lib.fnGetMetadataNodeLookupDimensionIdx(node_id, i)})
        for i in range(lib.fnGetMetadataHybridCount(node_id)):
            node["hybrids"].append({"name": lib.fnGetMetadataHybridName(node_id, i).decode(), "dim": lib.fnGetMetadataHybridDimensionIdx(node_id, i),
                                    "default": float(lib.fnGetMetadataHybridDefault(node_id, i))})
        out.append(node)
    return out


def main():
    for variant, suffix in (("strict_avx2", ""), ("strict_avx512", "_tables")):  # generic / table gradient flavour
        ref = helpers.load_ref(variant)
        assert ref is not None, "build oracle/_ref first (and run on a CPU with AVX-512 for the table flavour)"
        if not suffix:
            json.dump(dump_metadata(ref), open(os.path.join(HERE, "metadata.json"), "w"), indent=1)
        arrays = {}
        summary = {}
        for name, case in GOLDEN_CASES.items():
            out, mm = run_case(helpers.RefNode(ref, case["tree"]), case)
            arrays[name] = out
            summary[name] = {"min": float(mm[0]), "max": float(mm[1]), "fnv1a": "%08x" % helpers.fnv1a_words(out), "n": int(out.size)}
            print(variant, name, summary[name])
        np.savez_compressed(os.path.join(HERE, "outputs%s.npz" % suffix), **arrays)
        json.dump(summary, open(os.path.join(HERE, "summary%s.json" % suffix), "w"), indent=1)


if __name__ == "__main__":
    main()
