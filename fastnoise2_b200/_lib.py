"""ctypes binding of libfastnoise2_b200.so (the C ABI of include/FastNoise/FastNoise_C.h + FastNoise_B200.h).

The library is the product; this module only declares prototypes. Importing it never falls back to any
other implementation: a missing or unloadable library raises ImportError.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfastnoise2_b200.so")

c_float_p = ctypes.POINTER(ctypes.c_float)
F, I, U, P, LL = ctypes.c_float, ctypes.c_int, ctypes.c_uint, ctypes.c_void_p, ctypes.c_longlong
S = ctypes.c_char_p

# name -> (restype, argtypes); every symbol declared in include/FastNoise/FastNoise_C.h and FastNoise_B200.h
PROTOTYPES = {
    # FastNoise_C.h
    "fnNewFromEncodedNodeTree": (P, [S, U]),
    "fnDeleteNodeRef": (None, [P]),
    "fnGetActiveFeatureSet": (U, [P]),
    "fnGetMetadataID": (I, [P]),
    "fnGenUniformGrid2D": (None, [P, P, F, F, I, I, F, F, I, P]),
    "fnGenUniformGrid3D": (None, [P, P, F, F, F, I, I, I, F, F, F, I, P]),
    "fnGenUniformGrid4D": (None, [P, P, F, F, F, F, I, I, I, I, F, F, F, F, I, P]),
    "fnGenPositionArray2D": (None, [P, P, I, P, P, F, F, I, P]),
    "fnGenPositionArray3D": (None, [P, P, I, P, P, P, F, F, F, I, P]),
    "fnGenPositionArray4D": (None, [P, P, I, P, P, P, P, F, F, F, F, I, P]),
    "fnGenTileable2D": (None, [P, P, I, I, F, F, I, P]),
    "fnGenSingle2D": (F, [P, F, F, I]),
    "fnGenSingle3D": (F, [P, F, F, F, I]),
    "fnGenSingle4D": (F, [P, F, F, F, F, I]),
    "fnGetMetadataCount": (I, []),
    "fnGetMetadataName": (S, [I]),
    "fnNewFromMetadata": (P, [I, U]),
    "fnGetMetadataVariableCount": (I, [I]),
    "fnGetMetadataVariableName": (S, [I, I]),
    "fnGetMetadataVariableType": (I, [I, I]),
    "fnGetMetadataVariableDimensionIdx": (I, [I, I]),
    "fnGetMetadataEnumCount": (I, [I, I]),
    "fnGetMetadataEnumName": (S, [I, I, I]),
    "fnSetVariableFloat": (ctypes.c_bool, [P, I, F]),
    "fnSetVariableIntEnum": (ctypes.c_bool, [P, I, I]),
    "fnGetMetadataNodeLookupCount": (I, [I]),
    "fnGetMetadataNodeLookupName": (S, [I, I]),
    "fnGetMetadataNodeLookupDimensionIdx": (I, [I, I]),
    "fnSetNodeLookup": (ctypes.c_bool, [P, I, P]),
    "fnGetMetadataHybridCount": (I, [I]),
    "fnGetMetadataHybridName": (S, [I, I]),
    "fnGetMetadataHybridDimensionIdx": (I, [I, I]),
    "fnSetHybridNodeLookup": (ctypes.c_bool, [P, I, P]),
    "fnSetHybridFloat": (ctypes.c_bool, [P, I, F]),
    "fnGetMetadataDescription": (S, [I]),
    "fnGetMetadataGroupCount": (I, [I]),
    "fnGetMetadataGroupName": (S, [I, I]),
    "fnGetMetadataVariableDescription": (S, [I, I]),
    "fnGetMetadataVariableDefaultFloat": (F, [I, I]),
    "fnGetMetadataVariableDefaultIntEnum": (I, [I, I]),
    "fnGetMetadataVariableMinFloat": (F, [I, I]),
    "fnGetMetadataVariableMaxFloat": (F, [I, I]),
    "fnGetMetadataNodeLookupDescription": (S, [I, I]),
    "fnGetMetadataHybridDescription": (S, [I, I]),
    "fnGetMetadataHybridDefault": (F, [I, I]),
    # FastNoise_B200.h
    "fnb200LastError": (S, []),
    "fnb200DeviceAvailable": (I, []),
    "fnb200SetMathMode": (I, [I]),
    "fnb200GetMathMode": (I, []),
    "fnb200Prepare": (I, [P, I]),
    "fnb200GetKernelSource": (ctypes.c_long, [P, I, P, ctypes.c_long]),
    "fnb200GetKernelParams": (ctypes.c_long, [P, I, P, ctypes.c_long]),
    "fnb200SelfTestInvSqrt": (LL, [F, F]),
    "fnb200SelfTestLeaf": (I, [I, I, I, I, F, P]),
    "fnb200DebugIntermediates": (I, [I, I, F, F, I, I, I, F, F, F, F, P]),
    "fnb200MinMaxReset": (I, [P, P]),
    "fnb200MinMaxRead": (I, [P, P, P]),
    "fnb200GenUniformGrid2D": (I, [P, P, F, F, I, I, F, F, I, I, I, P, P]),
    "fnb200GenUniformGrid3D": (I, [P, P, F, F, F, I, I, I, F, F, F, I, I, I, P, P]),
    "fnb200GenUniformGrid4D": (I, [P, P, F, F, F, F, I, I, I, I, F, F, F, F, I, I, I, P, P]),
    "fnb200GenTileable2D": (I, [P, P, I, I, F, F, I, I, I, P, P]),
    "fnb200GenPositionArray": (I, [P, I, P, LL, P, P, P, P, F, F, F, F, I, P, P]),
    "fnb200FillUniformPhilox": (I, [P, LL, LL, U, U, F, F, P]),
    "fnb200MeasureAluPeak": (ctypes.c_double, [I, I]),
    "fnb200CanonicalEncodedNodeTree": (ctypes.c_long, [S, P, ctypes.c_long]),
    "fnb200LaunchCount": (ctypes.c_ulonglong, []),
    "fnb200CommGetUniqueId": (I, [P]),
    "fnb200CommInitRank": (P, [I, I, P]),
    "fnb200CommDestroy": (I, [P]),
    "fnb200CommAllReduceMinMax": (I, [P, P, P]),
    "fnb200CommGatherSlabs": (I, [P, P, P, P, P]),
    "fnb200CommAllocVolume": (P, [P, LL, P]),
    "fnb200VolumeData": (P, [P]),
    "fnb200CommFreeVolume": (I, [P, P, P]),
    "fnb200VolumeMirrors": (I, [P, LL, P, P, P]),
    "fnb200GenUniformGrid3DShared": (I, [P, P, F, F, F, I, I, I, F, F, F, I, I, I, P, P]),
}


def bind(path):
    """Load a library exporting (a subset of) the fn* C ABI and attach prototypes to what it exports."""
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in PROTOTYPES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            continue
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


def load():
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (or python fastnoise2_b200/build.py)" % LIB_PATH)
    lib = bind(LIB_PATH)
    missing = [n for n in PROTOTYPES if not hasattr(lib, n)]
    if missing:
        raise ImportError("libfastnoise2_b200.so does not export: " + ", ".join(missing))
    return lib
