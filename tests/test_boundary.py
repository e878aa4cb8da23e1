"""Drop-in boundary, no GPU: exported symbols, reflection tables, encoded-tree codec, C++ API surface."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from tests import cases, helpers

ROOT = helpers.ROOT


def declared_symbols():
    names = []
    for header in ("FastNoise_C.h", "FastNoise_B200.h"):
        text = open(os.path.join(ROOT, "include", "FastNoise", header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names += re.findall(r"FASTNOISE_API\s+[\w\s\*]+?\b(fn\w+)\s*\(", text)
    return names


def test_library_exports_every_declared_symbol(fn):
    names = declared_symbols()
    assert len([n for n in names if not n.startswith("fnb200")]) == 45  # the reference's C ABI has 45 entry points
    lib = ctypes.CDLL(fn._lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(set(names)) == sorted(fn._lib.PROTOTYPES)  # the Python binding declares exactly the headers' symbols


def test_metadata_matches_reference_tables(fn):
    """Node ids, member order/names/types/dims/enum strings/defaults == the reference's (golden dump of oracle/_ref)."""
    _, _, golden = helpers.load_golden()
    from tests.golden.make_golden import dump_metadata
    ours = dump_metadata(fn.lib)
    assert len(ours) == len(golden) == 47
    assert sum(len(n["variables"]) for n in ours) == 93   # reference Metadata.cpp:24-33 reserves exactly these
    assert sum(len(n["lookups"]) for n in ours) == 32
    assert sum(len(n["hybrids"]) for n in ours) == 59
    for a, b in zip(ours, golden):
        assert a == b, (a["name"], a, b)


def test_metadata_matches_live_reference(fn, ref):
    from tests.golden.make_golden import dump_metadata
    assert dump_metadata(fn.lib) == dump_metadata(ref)


def test_reflection_error_conventions(fn):
    lib = fn.lib
    assert lib.fnGetMetadataName(47) == b"INVALID NODE ID"
    assert lib.fnGetMetadataName(-1) == b"INVALID NODE ID"
    assert lib.fnGetMetadataVariableCount(200) == -1
    assert lib.fnGetMetadataVariableName(6, 99) == b"INVALID VARIABLE INDEX"
    assert lib.fnGetMetadataEnumName(11, 7, 9) == b"INVALID ENUM INDEX"
    assert lib.fnGetMetadataNodeLookupName(13, 5) == b"INVALID NODE LOOKUP INDEX"
    assert lib.fnGetMetadataHybridName(13, 5) == b"INVALID HYBRID INDEX"
    assert lib.fnGetMetadataGroupName(6, 3) == b"INVALID GROUP INDEX"
    assert lib.fnNewFromMetadata(99, 0xFFFFFFFF) is None
    assert lib.fnNewFromMetadata(6, 0) is None          # FeatureSet::Invalid
    h = lib.fnNewFromMetadata(19, 0xFFFFFFFF)           # DomainWarpFractalProgressive needs a DomainWarp source
    simplex = lib.fnNewFromMetadata(6, 0xFFFFFFFF)
    warp = lib.fnNewFromMetadata(18, 0xFFFFFFFF)
    assert not lib.fnSetNodeLookup(h, 0, simplex)
    assert lib.fnSetNodeLookup(h, 0, warp)
    assert not lib.fnSetNodeLookup(h, 1, warp)
    assert not lib.fnSetVariableFloat(simplex, 17, 1.0)
    assert lib.fnGetActiveFeatureSet(simplex) == 0x100A
    for x in (h, simplex, warp):
        lib.fnDeleteNodeRef(x)


@pytest.mark.parametrize("bad", ["", "A", "BgQ", "BgQ=!", "@", "/w==", "DQkGDA", "DQkG", "////", "3/8A", "DQk=", "BgAAAAA="])
def test_decoder_rejects_malformed_strings(fn, bad):
    assert fn.lib.fnNewFromEncodedNodeTree(bad.encode(), 0xFFFFFFFF) is None


def test_decoder_agrees_with_reference_on_garbage(fn, ref):
    """Accept/reject decisions on random byte soup match the reference's decoder (both never crash)."""
    import base64
    rng = np.random.default_rng(5)
    checked = 0
    for _ in range(3000):
        n = int(rng.integers(1, 24))
        raw = bytes(int(b) for b in rng.integers(0, 47 if rng.random() < 0.7 else 256, n))
        text = base64.b64encode(raw)
        ours = fn.lib.fnNewFromEncodedNodeTree(text, 0xFFFFFFFF)
        theirs = ref.fnNewFromEncodedNodeTree(text, 0xFFFFFFFF)
        assert (ours is None) == (theirs is None), (raw.hex(), ours, theirs)
        if ours:
            assert fn.lib.fnGetMetadataID(ours) == ref.fnGetMetadataID(theirs)
            fn.lib.fnDeleteNodeRef(ours)
            checked += 1
    assert checked > 50


def test_canonical_encoding_round_trips(fn):
    """decode -> NodeData -> serialise reproduces strings written by the reference's real encoder, and is idempotent."""
    buf = ctypes.create_string_buffer(4096)
    for name, text in {**cases.DEMO_TREES, **fn.TREES}.items():
        n = fn.lib.fnb200CanonicalEncodedNodeTree(text.encode(), buf, 4096)
        assert n > 0, name
        canon = buf.value.decode()
        if name in cases.DEMO_TREES:
            assert canon == text, name
        n2 = fn.lib.fnb200CanonicalEncodedNodeTree(canon.encode(), buf, 4096)
        assert n2 == n and buf.value.decode() == canon
        assert fn.Generator(text).kernel_source(3) == fn.Generator(canon).kernel_source(3)


def test_every_config_tree_lowers_and_compiles_for_sm100a(fn):
    """NVRTC needs no GPU: the generated CUDA of every config compiles to an sm_100a cubin."""
    from fastnoise2_b200 import api
    kinds = {"c2_fbm5_perlin": [api.KIND_GRID2D, api.KIND_TILEABLE2D], "c5_supersimplex": [api.KIND_POS4D, api.KIND_GRID4D]}
    for name, text in {**fn.TREES, **cases.DEMO_TREES}.items():
        g = fn.Generator(text)
        for kind in kinds.get(name, [api.KIND_GRID3D]):
            src = g.kernel_source(kind)
            assert "fn2_entry" in src and "#include \"fn2_device.cuh\"" in src
            g.prepare(kind)


def test_packed_form_compiles_without_fused_packed_products(fn, tmp_path):
    """The opt-in two-points-per-value form (FASTNOISE2_B200_PACK) must compile for every dimensionality, must actually use
    the packed FP32 instructions, and -- in strict mode -- must not contain a single FMUL2: ptxas contracts a packed multiply
    with a following packed add into one FFMA2 whatever --fmad says, so packed products are issued as FFMA2( a, b, -0 ) and an
    FMUL2 in the SASS would mean a product that can be fused, i.e. a rounding the reference does not have."""
    import glob
    import shutil
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    code = r'''
import os, sys
os.environ["FASTNOISE2_B200_CACHE"] = sys.argv[1]
os.environ["FASTNOISE2_B200_PACK"] = "2"
sys.path.insert(0, %r)
import fastnoise2_b200 as fn
from fastnoise2_b200 import api
for name, kind in (("c4_terrain", api.KIND_GRID3D), ("c2_fbm5_perlin", api.KIND_GRID2D), ("c3a_cellular_distance", api.KIND_GRID3D), ("c5_supersimplex", api.KIND_GRID4D)):
    g = fn.Generator(fn.TREES[name])
    assert "typedef fn2::fpair F;" in g.kernel_source(kind)
    g.prepare(kind)
print("ok")
''' % ROOT
    result = subprocess.run([sys.executable, "-c", code, str(tmp_path)], capture_output=True, text=True)
    assert result.returncode == 0 and result.stdout.strip().endswith("ok"), result.stderr[-2000:]
    cubins = glob.glob(str(tmp_path / "*.cubin"))
    # per tree (all four are uniform-grid kinds): the 4-points-per-thread kernel (two points per value), the 1-point one and the
    # any-row-width one (one point per value each)
    assert len(cubins) == 12
    packed = 0
    for cubin in cubins:
        sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
        packed += "FFMA2" in sass and "FADD2" in sass
        assert "FMUL2" not in sass
    assert packed >= 4


def test_incomplete_tree_fails_loudly(fn):
    g = fn.Generator.from_metadata(13)  # FractalFBm without its required Source
    rc = fn.lib.fnb200Prepare(g.handle, 3)
    assert rc == -4 and "source node is not set" in fn.last_error()


def test_no_gpu_means_failure_not_fallback(fn):
    """Without a device a Gen* call must fail (empty min/max, error text), never compute on the CPU."""
    if fn.device_available():
        pytest.skip("a GPU is present")
    g = fn.Generator(fn.TREES["c1_simplex"])
    out = np.full(64, 7.0, np.float32)
    mm = np.zeros(2, np.float32)
    fn.lib.fnGenUniformGrid3D(g.handle, out.ctypes.data, 0, 0, 0, 4, 4, 4, 1, 1, 1, 1337, mm.ctypes.data)
    assert mm[0] > mm[1] and (out == 7.0).all() and fn.last_error()


CPP_PROGRAM = r"""
#include <FastNoise/FastNoise.h>
#include <FastNoise/Metadata.h>
#include <FastNoise/FastNoise_B200.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

static std::string Source( const FastNoise::Generator* g, int kind )
{
    long n = fnb200GetKernelSource( g, kind, nullptr, 0 );
    std::vector<char> buf( n + 1 );
    fnb200GetKernelSource( g, kind, buf.data(), n + 1 );
    return buf.data();
}

int main()
{
    // same calls as the reference's tests/FastNoiseCpp11Include.cpp
    auto node = FastNoise::New<FastNoise::FractalFBm>();
    if( (unsigned)node->GetActiveFeatureSet() != 0x100A ) return 1;
    node->SetSource( FastNoise::New<FastNoise::Simplex>() );
    node->SetGain( FastNoise::New<FastNoise::Value>() );
    node->SetOctaveCount( 5 );

    FastNoise::SmartNode<> base = FastNoise::New<FastNoise::Checkerboard>();
    auto checkerboard = FastNoise::SmartNode<FastNoise::Checkerboard>::DynamicCast( base );
    if( !checkerboard ) return 2;
    checkerboard->SetScale( 8.0f );
    if( FastNoise::SmartNode<FastNoise::Simplex>::DynamicCast( base ) ) return 3;
    if( base.use_count() != 2 ) return 4;

    // a tree built with typed setters lowers to the same program as its encoded form "DQAFAAAACQgM"
    auto fbm = FastNoise::New<FastNoise::FractalFBm>();
    fbm->SetSource( FastNoise::New<FastNoise::Perlin>() );
    fbm->SetOctaveCount( 5 );
    FastNoise::SmartNode<> decoded = FastNoise::NewFromEncodedNodeTree( "DQAFAAAACQgM" );
    if( !decoded || Source( fbm.get(), 2 ) != Source( decoded.get(), 2 ) ) return 5;

    // the config-4 terrain tree, built by hand
    auto warp = FastNoise::New<FastNoise::DomainWarpGradient>();
    warp->SetScale( 200.0f );
    warp->SetWarpAmplitude( 30.0f );
    auto simplex = FastNoise::New<FastNoise::Simplex>(); simplex->SetScale( 150.0f );
    auto fbm2 = FastNoise::New<FastNoise::FractalFBm>(); fbm2->SetSource( simplex ); fbm2->SetOctaveCount( 5 );
    auto perlin = FastNoise::New<FastNoise::Perlin>(); perlin->SetScale( 90.0f ); perlin->SetSeedOffset( 7 );
    auto ridged = FastNoise::New<FastNoise::FractalRidged>(); ridged->SetSource( perlin ); ridged->SetOctaveCount( 4 );
    auto maxSmooth = FastNoise::New<FastNoise::MaxSmooth>(); maxSmooth->SetLHS( fbm2 ); maxSmooth->SetRHS( ridged ); maxSmooth->SetSmoothness( 0.2f );
    auto remap = FastNoise::New<FastNoise::Remap>(); remap->SetSource( maxSmooth ); remap->SetFromMax( 2.0f ); remap->SetClampOutput( true );
    auto terrace = FastNoise::New<FastNoise::Terrace>(); terrace->SetSource( remap ); terrace->SetStepCount( 12.0f ); terrace->SetSmoothness( 0.25f );
    warp->SetSource( terrace );
    auto progressive = FastNoise::New<FastNoise::DomainWarpFractalProgressive>();
    progressive->SetSource( warp );
    progressive->SetOctaveCount( 3 );
    FastNoise::SmartNode<> c4 = FastNoise::NewFromEncodedNodeTree( "%s" );
    if( !c4 || Source( progressive.get(), 3 ) != Source( c4.get(), 3 ) ) return 6;

    // reflection
    const FastNoise::Metadata& md = FastNoise::Metadata::Get<FastNoise::CellularDistance>();
    if( md.id != 11 || std::strcmp( md.name, "CellularDistance" ) || md.memberVariables.size() != 8 ) return 7;
    if( FastNoise::Metadata::FormatMetadataNodeName( &md ) != "Cellular Distance" ) return 8;
    if( FastNoise::Metadata::FormatMetadataMemberName( FastNoise::Metadata::Get<FastNoise::Gradient>().memberHybrids[2] ) != "Z Offset" ) return 9;
    if( FastNoise::New<FastNoise::Simplex>( FastSIMD::FeatureSet::Invalid ) ) return 10;
    if( FastNoise::NewFromEncodedNodeTree( "not base64!" ) ) return 11;

    // NodeData serialiser
    FastNoise::NodeData fbmData( &FastNoise::Metadata::Get<FastNoise::FractalFBm>() );
    FastNoise::NodeData perlinData( &FastNoise::Metadata::Get<FastNoise::Perlin>() );
    fbmData.nodeLookups[0] = &perlinData;
    fbmData.variables[0] = 5;
    if( FastNoise::Metadata::SerialiseNodeData( &fbmData ) != "DQAF@BCQgM" ) return 12;
    fbmData.nodeLookups[0] = &fbmData;                         // a cycle: fixUp must cut it, plain serialise must fail safe
    if( !FastNoise::Metadata::SerialiseNodeData( &fbmData, true ).empty() ) return 13;

    std::puts( "ok" );
    return 0;
}
"""


def test_cpp_api_surface_compiles_and_behaves(fn, tmp_path):
    src = tmp_path / "api_check.cpp"
    src.write_text(CPP_PROGRAM % fn.TREES["c4_terrain"])
    exe = tmp_path / "api_check"
    libdir = os.path.dirname(fn._lib.LIB_PATH)
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), str(src), "-L" + libdir, "-lfastnoise2_b200",
                           "-Wl,-rpath," + libdir, "-o", str(exe)])
    result = subprocess.run([str(exe)], capture_output=True, text=True)
    assert result.returncode == 0 and result.stdout.strip() == "ok", (result.returncode, result.stdout, result.stderr)


def test_public_headers_parse_as_cpp11(tmp_path):
    """The reference keeps its public headers C++11-clean (tests/FastNoiseCpp11Include.cpp); so do these, a typed New<T>() included."""
    src = tmp_path / "inc.cpp"
    src.write_text("#include <FastNoise/FastNoise.h>\n#include <FastNoise/Metadata.h>\n#include <FastNoise/FastNoise_C.h>\n#include <FastNoise/FastNoise_B200.h>\n"
                   "int main() { auto n = FastNoise::New<FastNoise::FractalFBm>(); n->SetSource( FastNoise::New<FastNoise::Simplex>() ); return n ? 0 : 1; }\n")
    subprocess.check_call(["/usr/bin/g++", "-std=c++11", "-Wall", "-Werror", "-fsyntax-only", "-I" + os.path.join(ROOT, "include"), str(src)])


def test_canonical_reencoding_preserves_random_trees(fn, ref):
    """Codec fuzz: 300 random trees (all node types, generator hybrids, shared sub-trees -> back-references) are decoded and
    re-encoded by the product; the reference's own decoder must build a tree from the re-encoded string that evaluates to
    exactly the same values as the original string."""
    from tests.encode import N
    from tests.fuzz import TreeFuzzer, same

    def hits_metadata_default_quirk(node, seen):
        # DistanceToPoint is constructed with EuclideanSquared but its metadata default says Euclidean
        # (BasicGenerators.h:289 vs :305), and the serialiser omits values equal to the METADATA default: an explicit
        # Euclidean does not survive a round trip in the reference either. Reproduced, so such trees are skipped here.
        if id(node) in seen:
            return False
        seen.add(id(node))
        if node.meta["name"] == "DistanceToPoint" and node.variables.get("Distance Function") == 0:
            return True
        children = list(node.lookups) + [h for h in node.hybrids.values() if isinstance(h, N)]
        return any(hits_metadata_default_quirk(c, seen) for c in children)

    fuzzer = TreeFuzzer(31337)
    buf = ctypes.create_string_buffer(1 << 16)
    checked = 0
    for index in range(300):
        tree = fuzzer.tree(depth=3 if index % 4 else 4)
        if hits_metadata_default_quirk(tree, set()):
            continue
        checked += 1
        text = tree.encode()
        n = fn.lib.fnb200CanonicalEncodedNodeTree(text.encode(), buf, len(buf))
        assert n > 0, text
        canon = buf.value.decode()
        a, b = helpers.RefNode(ref, text), helpers.RefNode(ref, canon)
        assert same(a.grid3d(-3.3, 2.1, 0.7, 6, 6, 6, 1.9, 1.7, 2.3, index)[0], b.grid3d(-3.3, 2.1, 0.7, 6, 6, 6, 1.9, 1.7, 2.3, index)[0]), (text, canon)
        assert same(a.grid2d(-3.3, 2.1, 12, 12, 1.9, 1.7, index)[0], b.grid2d(-3.3, 2.1, 12, 12, 1.9, 1.7, index)[0]), (text, canon)
    assert checked > 250


def test_canonical_encoding_is_byte_identical_to_the_reference_serialiser(fn, ref):
    """The codec against the reference's own (Metadata::DeserialiseNodeData + SerialiseNodeData, exposed by the oracle build):
    400 random trees and 2000 corrupted ones must re-encode to the same bytes, or be rejected by both. (Offline:
    122 000 strings, no difference.)"""
    import base64
    from tests.fuzz import TreeFuzzer
    if not hasattr(ref, "fn2refCanonicalEncodedNodeTree"):
        pytest.skip("oracle/_ref predates fn2refCanonicalEncodedNodeTree: make -C oracle")
    fuzzer = TreeFuzzer(5150)
    rng = np.random.default_rng(5150)
    ours, theirs = ctypes.create_string_buffer(1 << 16), ctypes.create_string_buffer(1 << 16)
    accepted = rejected = 0
    for index in range(2400):
        text = fuzzer.tree(depth=3 if index % 3 else 4).encode().encode()
        if index >= 400:
            raw = bytearray(base64.b64decode(text))
            for _ in range(int(rng.integers(1, 4))):
                raw[int(rng.integers(len(raw)))] = int(rng.integers(256))
            text = base64.b64encode(bytes(raw))
        na = fn.lib.fnb200CanonicalEncodedNodeTree(text, ours, len(ours))
        nb = ref.fn2refCanonicalEncodedNodeTree(text, theirs, len(theirs))
        assert na == nb and ours.value == theirs.value, (text, ours.value, theirs.value)
        accepted, rejected = accepted + (na > 0), rejected + (na == 0)
    assert accepted > 1500 and rejected > 300, (accepted, rejected)


def test_decoder_agrees_with_reference_on_mutated_trees(fn, ref):
    """Decoder robustness: valid random trees with one to three bytes corrupted (or truncated) reach deep into both decoders;
    accept / reject decisions and the decoded root type must match the reference's, and neither may crash."""
    import base64
    from tests.fuzz import TreeFuzzer
    fuzzer = TreeFuzzer(99)
    rng = np.random.default_rng(99)
    accepted = rejected = 0
    for index in range(400):
        raw = bytearray(base64.b64decode(fuzzer.tree(depth=3).encode()))
        for _ in range(int(rng.integers(1, 4))):
            raw[int(rng.integers(len(raw)))] = int(rng.integers(256))
        if index % 5 == 0:
            raw = raw[:int(rng.integers(1, len(raw) + 1))]
        text = base64.b64encode(bytes(raw))
        ours = fn.lib.fnNewFromEncodedNodeTree(text, 0xFFFFFFFF)
        theirs = ref.fnNewFromEncodedNodeTree(text, 0xFFFFFFFF)
        assert (ours is None) == (theirs is None), (bytes(raw).hex(), ours, theirs)
        if ours:
            assert fn.lib.fnGetMetadataID(ours) == ref.fnGetMetadataID(theirs)
            fn.lib.fnDeleteNodeRef(ours)
            accepted += 1
        else:
            rejected += 1
    assert accepted > 20 and rejected > 20, (accepted, rejected)


def test_reflection_setters_build_the_same_program_as_the_decoder(fn):
    """150 random trees built twice -- through fnNewFromMetadata + fnSetVariable* / fnSetNodeLookup / fnSetHybrid* (what the
    Node Editor and every binding do) and by decoding the encoded string -- must lower to the same CUDA source, in 2D, 3D
    and 4D: member indices, setter side effects (scale -> frequency, output range, rotations ...) and shared sub-trees agree."""
    from tests.encode import N
    from tests.fuzz import TreeFuzzer
    lib = fn.lib
    created = []

    def build(node, done):
        if id(node) in done:
            return done[id(node)]
        meta = node.meta
        h = lib.fnNewFromMetadata(meta["id"], 0xFFFFFFFF)
        assert h
        created.append(h)
        for key, value in node.variables.items():
            i = N._index(meta["variables"], key)
            if meta["variables"][i]["type"] == 0:
                assert lib.fnSetVariableFloat(h, i, value)
            else:
                assert lib.fnSetVariableIntEnum(h, i, value)
        for i, child in enumerate(node.lookups):
            assert lib.fnSetNodeLookup(h, i, build(child, done)), (meta["name"], i)
        for key, value in node.hybrids.items():
            i = N._index(meta["hybrids"], key)
            if isinstance(value, N):
                assert lib.fnSetHybridNodeLookup(h, i, build(value, done))
            else:
                assert lib.fnSetHybridFloat(h, i, value)
        done[id(node)] = h
        return h

    fuzzer = TreeFuzzer(2718)
    for index in range(150):
        tree = fuzzer.tree(depth=3)
        decoded = fn.Generator(tree.encode())
        root = build(tree, {})
        for kind in (2, 3, 4):
            n = lib.fnb200GetKernelSource(root, kind, None, 0)
            assert n > 0, fn.last_error()
            buf = ctypes.create_string_buffer(n + 1)
            lib.fnb200GetKernelSource(root, kind, buf, n + 1)
            assert buf.value.decode() == decoded.kernel_source(kind), (index, kind, tree.encode())
    for h in created:
        lib.fnDeleteNodeRef(h)


def test_every_reference_setter_compiles_against_these_headers(tmp_path):
    """tests/golden/api_surface.cpp calls every public setter the reference declares for each of the 47 node classes (237
    calls: own and inherited, every overload, every Dim instantiation; generated from the reference's headers by
    tests/golden/make_api_surface.py) plus the Generator / SmartNode / Metadata / NodeData surface. It must compile, link
    and run against this repo's headers and library -- and, where the reference is present, pass a syntax check against the
    reference's own headers, which proves the file is what code written for the reference looks like."""
    source = os.path.join(ROOT, "tests", "golden", "api_surface.cpp")
    libdir = os.path.join(ROOT, "fastnoise2_b200")
    exe = str(tmp_path / "api_surface")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), source, "-L" + libdir, "-lfastnoise2_b200",
                           "-Wl,-rpath," + libdir, "-o", exe])
    assert subprocess.run([exe]).returncode == 0  # links and runs (without a GPU the Gen* calls fail cleanly, they never crash)
    if os.path.isdir("/root/reference/include"):
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-fsyntax-only", "-w", "-mavx2", "-mfma", "-I" + os.path.join(ROOT, "oracle", "shim"),
                               "-I/root/reference/include", source])


def test_c_abi_prototypes_are_compatible_with_the_reference_header(tmp_path):
    """Both FastNoise_C.h headers in ONE translation unit: a prototype here that differed from the reference's in any argument or
    return type would be a "conflicting types" error (C) / an ambiguous overload with C linkage (C++)."""
    reference = "/root/reference/include/FastNoise/FastNoise_C.h"
    if not os.path.exists(reference):
        pytest.skip("reference headers not present")
    ours = os.path.join(ROOT, "include", "FastNoise", "FastNoise_C.h")
    c = tmp_path / "both.c"
    c.write_text('#include <stdbool.h>\n#include "%s"\n#include "%s"\nint main(void) { return 0; }\n' % (reference, ours))
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-fsyntax-only", "-I" + os.path.join(ROOT, "include"), str(c)])  # macro redefinitions of the two Export.h are fine
    cpp = tmp_path / "both.cpp"
    cpp.write_text('#include "%s"\n#include "%s"\nint main() { return 0; }\n' % (reference, ours))
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-fsyntax-only", "-I" + os.path.join(ROOT, "include"), str(cpp)])
    # every function the reference declares is declared (and exported) here
    names = set(re.findall(r"FASTNOISE_API\s+[\w\s\*]+?\b(fn\w+)\s*\(", open(reference).read()))
    assert len(names) == 45
    exported = subprocess.run(["nm", "-D", "--defined-only", os.path.join(ROOT, "fastnoise2_b200", "libfastnoise2_b200.so")], capture_output=True, text=True).stdout
    missing = [n for n in sorted(names) if (" T " + n + "\n") not in exported]
    assert not missing, missing


def test_reflection_getters_match_live_reference_including_invalid_arguments(fn, ref):
    """Every reflection getter of the C ABI, for every node id from -2 to 48 and every member index from -2 to count+1, returns
    what the reference returns -- valid values, and the "INVALID ..." strings / -1 / 0 conventions for bad ids and indices.
    (Descriptions are this repo's own wording for valid members, so only their INVALID forms are compared.)"""
    a, b = fn.lib, ref
    ids = range(-2, 49)

    def both(name, *args):
        x, y = getattr(a, name)(*args), getattr(b, name)(*args)
        if isinstance(x, float) or isinstance(y, float):
            assert (x == y) or (x != x and y != y), (name, args, x, y)
        else:
            assert x == y, (name, args, x, y)
        return x

    assert both("fnGetMetadataCount") == 47
    for node in ids:
        both("fnGetMetadataName", node)
        groups = both("fnGetMetadataGroupCount", node)
        for g in range(-2, max(groups, 0) + 2):
            both("fnGetMetadataGroupName", node, g)
        variables = both("fnGetMetadataVariableCount", node)
        for v in range(-2, max(variables, 0) + 2):
            both("fnGetMetadataVariableName", node, v)
            both("fnGetMetadataVariableType", node, v)
            both("fnGetMetadataVariableDimensionIdx", node, v)
            both("fnGetMetadataVariableDefaultFloat", node, v)
            both("fnGetMetadataVariableDefaultIntEnum", node, v)
            both("fnGetMetadataVariableMinFloat", node, v)
            both("fnGetMetadataVariableMaxFloat", node, v)
            enums = both("fnGetMetadataEnumCount", node, v)
            for e in range(-2, max(enums, 0) + 2):
                both("fnGetMetadataEnumName", node, v, e)
        lookups = both("fnGetMetadataNodeLookupCount", node)
        for i in range(-2, max(lookups, 0) + 2):
            both("fnGetMetadataNodeLookupName", node, i)
            both("fnGetMetadataNodeLookupDimensionIdx", node, i)
        hybrids = both("fnGetMetadataHybridCount", node)
        for i in range(-2, max(hybrids, 0) + 2):
            both("fnGetMetadataHybridName", node, i)
            both("fnGetMetadataHybridDimensionIdx", node, i)
            both("fnGetMetadataHybridDefault", node, i)
    for node in (-2, -1, 47, 48):                   # descriptions: only the invalid-argument conventions are the reference's
        both("fnGetMetadataDescription", node)
        both("fnGetMetadataVariableDescription", node, 0)
        both("fnGetMetadataNodeLookupDescription", node, 0)
        both("fnGetMetadataHybridDescription", node, 0)
    for node in (6, 13):
        for bad in (-1, 99):
            both("fnGetMetadataVariableDescription", node, bad)
            both("fnGetMetadataNodeLookupDescription", node, bad)
            both("fnGetMetadataHybridDescription", node, bad)
    # setters and factories with invalid arguments
    for lib in (a, b):
        assert lib.fnNewFromMetadata(-1, 0xFFFFFFFF) is None and lib.fnNewFromMetadata(47, 0xFFFFFFFF) is None
        assert lib.fnNewFromEncodedNodeTree(b"", 0xFFFFFFFF) is None
    ha, hb = a.fnNewFromMetadata(13, 0xFFFFFFFF), b.fnNewFromMetadata(13, 0xFFFFFFFF)
    ca, cb = a.fnNewFromMetadata(0, 0xFFFFFFFF), b.fnNewFromMetadata(0, 0xFFFFFFFF)
    for index in (-1, 0, 1, 2, 7):
        assert bool(a.fnSetVariableFloat(ha, index, 2.0)) == bool(b.fnSetVariableFloat(hb, index, 2.0)), index
        assert bool(a.fnSetVariableIntEnum(ha, index, 3)) == bool(b.fnSetVariableIntEnum(hb, index, 3)), index
        assert bool(a.fnSetNodeLookup(ha, index, ca)) == bool(b.fnSetNodeLookup(hb, index, cb)), index
        assert bool(a.fnSetHybridFloat(ha, index, 0.5)) == bool(b.fnSetHybridFloat(hb, index, 0.5)), index
        assert bool(a.fnSetHybridNodeLookup(ha, index, ca)) == bool(b.fnSetHybridNodeLookup(hb, index, cb)), index
    assert a.fnGetMetadataID(ha) == b.fnGetMetadataID(hb) == 13
    for lib, h in ((a, ha), (a, ca), (b, hb), (b, cb)):
        lib.fnDeleteNodeRef(h)


def test_float_parameters_are_kernel_arguments_not_source_text(fn):
    """One compiled kernel serves every tree of the same SHAPE: a float edit through any setter (the reference applies it at once,
    FastNoise_C.cpp:221-239) leaves the generated source -- the cubin cache key -- unchanged and only moves a value of the kernel's
    parameter argument. Integers / enums, and floats the lowering looks at to pick a code shape, still change the text."""
    from tests.trees import Tree
    lib = fn.lib
    t = Tree(lib)
    fbm = t.new(13)                      # FractalFBm( Simplex )
    simplex = t.leaf(6, 37.0)
    assert lib.fnSetNodeLookup(fbm, 0, simplex)
    g = fn.Generator(handle=fbm)
    try:
        src0, par0 = g.kernel_source(3), g.kernel_params(3).copy()
        assert "tp[0]" in src0 and "0x1." not in src0.split("// FractalFBm")[1].split("extern")[0]  # no float literal of a parameter in the node functions
        def set_hybrid(node, name, value):
            node_id = lib.fnGetMetadataID(node)
            index = [lib.fnGetMetadataHybridName(node_id, i).decode() for i in range(lib.fnGetMetadataHybridCount(node_id))].index(name)
            assert lib.fnSetHybridFloat(node, index, value)

        t.set_float(simplex, "Feature Scale", 53.0)
        set_hybrid(fbm, "Gain", 0.625)
        t.set_float(fbm, "Lacunarity", 2.5)
        src1, par1 = g.kernel_source(3), g.kernel_params(3).copy()
        assert src1 == src0 and len(par1) == len(par0) and not np.array_equal(par0, par1)
        assert np.float32(0.625) in par1 and np.float32(2.5) in par1 and np.float32(1.0) / np.float32(53.0) in par1
        t.set_int(fbm, "Octaves", 5)         # structure: a different kernel
        assert g.kernel_source(3) != src0
        t.set_int(fbm, "Octaves", 3)
        assert g.kernel_source(3) == src0
        set_hybrid(fbm, "Weighted Strength", 0.5)  # 0 -> non-zero switches the weighted-amplitude code on: shape change
        assert g.kernel_source(3) != src0
    finally:
        g.handle = None
        t.close()
