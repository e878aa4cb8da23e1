// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Stand-in for the FastSIMD API subset that the FastNoise2 node sources are written against.
// The real dependency (Auburn/FastSIMD @ 16450dae9528727e500e7254f635a671f9c7ee2d, pinned in
// /root/reference/src/CMakeLists.txt:1-5) is fetched by CPM at configure time and is NOT present in
// this container, so its lane semantics are restated here. With this header on the include path
// the reference's own, unmodified hot-path translation unit (src/FastNoise/FastSIMD_Build.inl and
// the Generators/*.inl it includes) plus Metadata.cpp / SmartNode.cpp / FastNoise_C.cpp compile in
// place under /root/reference (recipe: oracle/Makefile, outputs only in oracle/_ref/).
//
// Two builds are made from this one header:
//   strict  (FN2SHIM_RELAXED=0, -O2 -ffp-contract=off):  every op is the IEEE-754 single op,
//           FMulAdd/FNMulAdd/FMulSub are single-rounding fused ops, InvSqrt = 1/sqrt, Reciprocal = 1/x.
//           This is the PARITY target for the CUDA path.
//   relaxed (FN2SHIM_RELAXED=1, -O3 -ffast-math): rsqrt/rcp approximations, mirrors the reference's
//           default build flags (src/CMakeLists.txt:65-69,96-102). Used for CPU TIMING only.
// Lanes: FN2SHIM_LANES = 8 (AVX2+FMA, one ymm) or 16 (AVX-512F, one zmm).
//
// Two FLAVOURS of the node sources' `if constexpr( SIMD & FeatureFlag::X )` chains:
//   generic (FN2SHIM_AVX512_PATHS=0): every flag reads false -- the select/xor code path FastSIMD's
//           Scalar/NEON/WASM levels run.
//   tables  (FN2SHIM_AVX512_PATHS=1, 16 lanes): FeatureFlag::AVX512_F reads true -- the permute-table
//           gradient forms an AVX-512 build of the reference runs (Utils.inl:38,84,124,167,196 and
//           DomainWarpSimplex.inl:13,70,126,198,254,370). The two flavours round differently (a table
//           weight multiplies inside the FMA chain), so each is its own parity target.
//
// Semantics that cannot be checked against upstream here ("parity unpinned"):
//   Convert<int32>(float) = round-to-nearest-even (cvtps2dq); Min/Max = x86 minps/maxps operand
//   order; Sin/Cos = the quadrant-folded polynomial below; Pow / Modulus = the restatements below.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <cmath>
#include <type_traits>
#include <new>
#include <immintrin.h>

#ifndef FN2SHIM_LANES
#define FN2SHIM_LANES 8
#endif
#ifndef FN2SHIM_RELAXED
#define FN2SHIM_RELAXED 0
#endif

#define FS_FORCEINLINE inline __attribute__(( always_inline ))
#define FS_VECTORCALL
#define FS_BIND_INTRINSIC( f ) []( auto... a ) { return f( a... ); }

namespace FastSIMD
{
    enum class FeatureFlag : std::uint32_t { Scalar = 1, x86 = 2, SSE41 = 4, AVX = 8, AVX2 = 16, AVX512_F = 32 };
    enum class FeatureSet : std::uint32_t { Invalid = 0, SHIM = 1, Max = ~0u };

#ifndef FN2SHIM_AVX512_PATHS
#define FN2SHIM_AVX512_PATHS 0
#endif
#if FN2SHIM_AVX512_PATHS
    // "tables" flavour: exactly the AVX512_F branches (the 11 gradient-table sites of Utils.inl and
    // DomainWarpSimplex.inl); every other flag sits in an else-branch of those chains or guards a vzeroupper.
    constexpr bool operator&( FeatureSet, FeatureFlag f ) { return f == FeatureFlag::AVX512_F; }
#else
    // generic code path everywhere
    constexpr bool operator&( FeatureSet, FeatureFlag ) { return false; }
#endif
    constexpr FeatureSet FeatureSetDefault() { return FeatureSet::SHIM; }
    constexpr bool IsRelaxed() { return FN2SHIM_RELAXED != 0; }

    using MemoryAllocator = void* (*)( std::size_t size, std::size_t align );

    template<typename T, FeatureSet SIMD>
    class DispatchClass;

    template<typename T>
    T* NewDispatchClass( FeatureSet maxFeatureSet = FeatureSet::Max, MemoryAllocator allocator = nullptr );

    template<typename T>
    struct RegisterDispatchClass
    {
        static T* New( FeatureSet maxFeatureSet, MemoryAllocator allocator )
        {
            if( maxFeatureSet == FeatureSet::Invalid )
            {
                return nullptr;
            }
            using Impl = DispatchClass<T, FeatureSetDefault()>;
            void* mem = allocator ? allocator( sizeof( Impl ), alignof( Impl ) ) : ::operator new( sizeof( Impl ), std::align_val_t( alignof( Impl ) ) );
            return new( mem ) Impl;
        }
    };
}

namespace FS
{
    using FastSIMD::FeatureSet;

    constexpr std::size_t kLanes = FN2SHIM_LANES;

    typedef float        vf_t __attribute__(( vector_size( 4 * FN2SHIM_LANES ) ));
    typedef std::int32_t vi_t __attribute__(( vector_size( 4 * FN2SHIM_LANES ) ));
    typedef std::uint32_t vu_t __attribute__(( vector_size( 4 * FN2SHIM_LANES ) ));

    template<typename T> struct VecOf;
    template<> struct VecOf<float> { using type = vf_t; };
    template<> struct VecOf<std::int32_t> { using type = vi_t; };

    // Only used as Cast<> target tags inside discarded x86 branches
    template<std::size_t BITS, bool FLOATMASK = true>
    struct Mask {};

    // lane mask: every lane is 0 or -1
    struct MaskV
    {
        vi_t m;
        FS_FORCEINLINE MaskV operator~() const { return { ~m }; }
        FS_FORCEINLINE MaskV operator&( MaskV o ) const { return { m & o.m }; }
        FS_FORCEINLINE MaskV operator|( MaskV o ) const { return { m | o.m }; }
        FS_FORCEINLINE MaskV operator^( MaskV o ) const { return { m ^ o.m }; }
        FS_FORCEINLINE MaskV& operator&=( MaskV o ) { m &= o.m; return *this; }
        FS_FORCEINLINE MaskV& operator|=( MaskV o ) { m |= o.m; return *this; }
    };

    template<typename T, std::size_t N, FeatureSet SIMD = FastSIMD::FeatureSetDefault()>
    struct Register
    {
        static_assert( N == kLanes, "shim supports exactly one native register per FS::Register" );
        static constexpr std::size_t ElementCount = N;
        using ElementType = T;
        using MaskType = MaskV;
        using Vec = typename VecOf<T>::type;
        static constexpr bool kInt = std::is_integral<T>::value;

        Vec v;

        Register() = default;
        FS_FORCEINLINE Register( Vec raw, int ) : v( raw ) {}

        // broadcast from any arithmetic scalar (sources pass int, unsigned, float and double constants); lane by lane like
        // set1_ps, so that a -0.0f constant keeps its sign ( 0 + -0.0f would be +0.0f )
        static FS_FORCEINLINE Vec Splat( T s )
        {
            Vec r;
            for( int lane = 0; lane < (int)( sizeof( Vec ) / sizeof( T ) ); lane++ ) r[lane] = s;
            return r;
        }
        template<typename U, typename = std::enable_if_t<std::is_arithmetic<U>::value>>
        FS_FORCEINLINE Register( U s ) : v( Splat( (T)s ) ) {}

        FS_FORCEINLINE vi_t bits() const { return (vi_t)v; }
        static FS_FORCEINLINE Register fromBits( vi_t b ) { return Register( (Vec)b, 0 ); }

        // arithmetic (int32 wraps: done on unsigned lanes)
        FS_FORCEINLINE Register operator+( Register o ) const { if constexpr( kInt ) return Register( (Vec)( (vu_t)v + (vu_t)o.v ), 0 ); else return Register( v + o.v, 0 ); }
        FS_FORCEINLINE Register operator-( Register o ) const { if constexpr( kInt ) return Register( (Vec)( (vu_t)v - (vu_t)o.v ), 0 ); else return Register( v - o.v, 0 ); }
        FS_FORCEINLINE Register operator*( Register o ) const { if constexpr( kInt ) return Register( (Vec)( (vu_t)v * (vu_t)o.v ), 0 ); else return Register( v * o.v, 0 ); }
        FS_FORCEINLINE Register operator/( Register o ) const { return Register( v / o.v, 0 ); }
        FS_FORCEINLINE Register operator-() const
        {
            if constexpr( kInt ) return Register( (Vec)( (vu_t){} - (vu_t)v ), 0 );
            else return fromBits( bits() ^ (vi_t)( (vu_t){} + 0x80000000u ) ); // sign flip, also of zero
        }
        FS_FORCEINLINE Register& operator+=( Register o ) { return *this = *this + o; }
        FS_FORCEINLINE Register& operator-=( Register o ) { return *this = *this - o; }
        FS_FORCEINLINE Register& operator*=( Register o ) { return *this = *this * o; }
        FS_FORCEINLINE Register& operator/=( Register o ) { return *this = *this / o; }

        // bitwise, on the raw lanes (valid for float registers too)
        FS_FORCEINLINE Register operator&( Register o ) const { return fromBits( bits() & o.bits() ); }
        FS_FORCEINLINE Register operator|( Register o ) const { return fromBits( bits() | o.bits() ); }
        FS_FORCEINLINE Register operator^( Register o ) const { return fromBits( bits() ^ o.bits() ); }
        FS_FORCEINLINE Register operator~() const { return fromBits( ~bits() ); }
        FS_FORCEINLINE Register& operator&=( Register o ) { return *this = *this & o; }
        FS_FORCEINLINE Register& operator|=( Register o ) { return *this = *this | o; }
        FS_FORCEINLINE Register& operator^=( Register o ) { return *this = *this ^ o; }

        // shifts: << logical, >> arithmetic (FS::BitShiftRightZeroExtend is the logical one)
        FS_FORCEINLINE Register operator<<( int s ) const { return fromBits( (vi_t)( (vu_t)bits() << s ) ); }
        FS_FORCEINLINE Register operator>>( int s ) const { return fromBits( bits() >> s ); }
        FS_FORCEINLINE Register& operator<<=( int s ) { return *this = *this << s; }
        FS_FORCEINLINE Register& operator>>=( int s ) { return *this = *this >> s; }

        // comparisons -> lane mask
        FS_FORCEINLINE MaskV operator<( Register o ) const { return { (vi_t)( v < o.v ) }; }
        FS_FORCEINLINE MaskV operator>( Register o ) const { return { (vi_t)( v > o.v ) }; }
        FS_FORCEINLINE MaskV operator<=( Register o ) const { return { (vi_t)( v <= o.v ) }; }
        FS_FORCEINLINE MaskV operator>=( Register o ) const { return { (vi_t)( v >= o.v ) }; }
        FS_FORCEINLINE MaskV operator==( Register o ) const { return { (vi_t)( v == o.v ) }; }
        FS_FORCEINLINE MaskV operator!=( Register o ) const { return { (vi_t)( v != o.v ) }; }
    };

#define FN2_T template<typename T, std::size_t N, FeatureSet S>
#define FN2_R Register<T, N, S>
#define FN2_SCALAR_LHS template<typename T, std::size_t N, FeatureSet S, typename U, typename = std::enable_if_t<std::is_arithmetic<U>::value>>

    FN2_SCALAR_LHS FS_FORCEINLINE FN2_R operator+( U a, FN2_R b ) { return FN2_R( a ) + b; }
    FN2_SCALAR_LHS FS_FORCEINLINE FN2_R operator-( U a, FN2_R b ) { return FN2_R( a ) - b; }
    FN2_SCALAR_LHS FS_FORCEINLINE FN2_R operator*( U a, FN2_R b ) { return FN2_R( a ) * b; }

    template<typename T>
    constexpr std::size_t NativeRegisterCount() { return kLanes / 2; } // Generator.inl:10 doubles it

    using f32v = Register<float, kLanes>;
    using i32v = Register<std::int32_t, kLanes>;
}

#include "Functions.h"
