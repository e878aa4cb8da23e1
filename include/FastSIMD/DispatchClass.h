// fastnoise2_b200 -- backend selector types.
//
// The reference reaches its compute layer through FastSIMD::NewDispatchClass<T>( FeatureSet, allocator )
// (reference: include/FastNoise/FastNoise.h:42, src/FastNoise/Metadata.cpp:713-716) and every public
// factory takes a `FastSIMD::FeatureSet maxFeatureSet` argument. This header keeps that spelling so
// caller code compiles unchanged; the only level this library implements is the B200 (sm_100a) CUDA
// backend, there is no CPU level to fall back to.
#pragma once
#include <cstdint>

namespace FastSIMD
{
    enum class FeatureSet : std::uint32_t
    {
        Invalid = 0,
        B200_SM100A = 0x100A, ///< hand-written CUDA for sm_100a, tree JIT-specialised through NVRTC
        Max = ~0u             ///< "no limit": what callers pass by default (FastNoise_C.h: ~0u)
    };

    inline const char* GetFeatureSetString( FeatureSet featureSet )
    {
        switch( featureSet )
        {
        case FeatureSet::B200_SM100A: return "B200_SM100A";
        case FeatureSet::Max: return "Max";
        default: return "Invalid";
        }
    }
}
