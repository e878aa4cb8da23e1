// ORACLE / TEST INFRASTRUCTURE ONLY.
// Translation unit that pulls in the reference's UNMODIFIED hot-path sources from where they lie
// (/root/reference/src/FastNoise/FastSIMD_Build.inl -> include/FastNoise/Generators/*.inl) and
// instantiates every node class over the FastSIMD stand-in in oracle/shim/. Nothing from the
// reference is copied into this repository; this file only names it on the include path.
#include <FastSIMD/DispatchClass.h>
#include <FastNoise/Utility/Config.h>
#include <FastNoise/Metadata.h>

using namespace FastNoise;

template<typename T>
T* FastSIMD::NewDispatchClass( FeatureSet maxFeatureSet, MemoryAllocator allocator )
{
    return RegisterDispatchClass<T>::New( maxFeatureSet, allocator );
}

// FastSIMD_Build.inl:3-7 leaves this hook open for the dispatch library to define
#define FASTNOISE_REGISTER_NODE( CLASS ) \
    template struct FastSIMD::RegisterDispatchClass<FastNoise::CLASS>; \
    template FastNoise::CLASS* FastSIMD::NewDispatchClass<FastNoise::CLASS>( FastSIMD::FeatureSet, FastSIMD::MemoryAllocator )

#include "FastSIMD_Build.inl"
