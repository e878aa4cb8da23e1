// TEST INFRASTRUCTURE ONLY -- a minimal stand-in for the CUDA device environment, so that the device node library
// (fastnoise2_b200/csrc/device/*.cuh) and the per-tree code that lower.cpp generates can be compiled by a HOST compiler and
// executed on the CPU by tests/test_host_emulation.py. This checks the lowering (emitter + scalar instantiation of the node
// library) against the oracle for hundreds of random trees without a GPU. It is never part of the product: nothing under
// fastnoise2_b200/ includes it, and the product has no CPU path.
//
// "Threads" run one after another: threadIdx/blockIdx are globals set by the driver in the test, __syncthreads() is a
// no-op, warp reductions return the caller's own value, shared memory is static storage. The min/max reduction is not
// exercised (the driver passes a null accumulator and the test computes min/max from the output).
#pragma once
#define FN2_HOST_EMULATION 1
#include <algorithm>
#include <cmath>
#include <math.h>
#include <cstdint>
#include <cstring>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline __attribute__(( always_inline ))
#define __noinline__ __attribute__(( noinline ))
#define __constant__ static const
#define __shared__ static
#define __restrict__ __restrict
#define __launch_bounds__( ... )
#define FN2_GRID_CONSTANT
#define __align__( n ) __attribute__(( aligned( n ) ))

struct Fn2HostDim { unsigned x, y, z; };
static Fn2HostDim threadIdx, blockIdx, blockDim, gridDim;

struct float2 { float x, y; };
struct alignas( 16 ) float4 { float x, y, z, w; };
static inline float2 make_float2( float x, float y ) { return float2 { x, y }; }
static inline float4 make_float4( float x, float y, float z, float w ) { return float4 { x, y, z, w }; }

static inline float __fmaf_rn( float a, float b, float c ) { return std::fmaf( a, b, c ); }
static inline float __fdiv_rn( float a, float b ) { return a / b; }
static inline float __fsqrt_rn( float v ) { return std::sqrt( v ); }
static inline float __frcp_rn( float v ) { return 1.0f / v; }
static inline float rsqrtf( float v ) { return 1.0f / std::sqrt( v ); }
static inline int __float_as_int( float v ) { int b; std::memcpy( &b, &v, 4 ); return b; }
static inline float __int_as_float( int b ) { float v; std::memcpy( &v, &b, 4 ); return v; }
static inline float __int2float_rn( int v ) { return (float)v; }
// cvt.rni.s32.f32: nearest even, saturating, NaN -> 0
static inline int __float2int_rn( float v )
{
    if( v != v ) return 0;
    if( v >= 2147483648.0f ) return INT32_MAX;
    if( v <= -2147483648.0f ) return INT32_MIN;
    return (int)std::lrintf( v );
}
static inline unsigned __umulhi( unsigned a, unsigned b ) { return (unsigned)( ( (unsigned long long)a * b ) >> 32 ); }
static inline void __syncthreads() {}
static inline unsigned __reduce_max_sync( unsigned, unsigned v ) { return v; }
static inline unsigned atomicMax( unsigned* p, unsigned v ) { unsigned old = *p; if( v > old ) *p = v; return old; }
static inline float4 __ldcs( const float4* p ) { return *p; }
// packed FP32 (only reached when the test forces the two-points-per-value form)
static inline float2 __ffma2_rn( float2 a, float2 b, float2 c ) { return float2 { std::fmaf( a.x, b.x, c.x ), std::fmaf( a.y, b.y, c.y ) }; }
static inline float2 __fadd2_rn( float2 a, float2 b ) { return float2 { a.x + b.x, a.y + b.y }; }
static inline float2 __fmul2_rn( float2 a, float2 b ) { return float2 { a.x * b.x, a.y * b.y }; }
// floorf / rintf / truncf / fabsf / fminf / fmaxf come from <cmath> in the global namespace, with CUDA's semantics
