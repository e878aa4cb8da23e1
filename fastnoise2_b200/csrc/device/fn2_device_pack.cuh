// fastnoise2_b200 -- two points per value: the packed-FP32 path of sm_100a.
//
// These kernels are bound by instruction issue, and almost half of what they issue is FP32 arithmetic. Blackwell has
// packed FP32 instructions -- FFMA2 / FMUL2 / FADD2 work on two floats held in an aligned register pair, with the same
// negate / |abs| / immediate operand forms as the scalar ones -- that do the math of two lanes in ONE issue slot (the FMA
// pipe's lane throughput is unchanged; measured by fnb200MeasureAluPeak modes 4-6). So a thread that owns several points
// evaluates them two at a time: `fpair` holds the same quantity for two neighbouring points, every FP32 add / multiply / fma
// becomes one packed instruction, and everything that has no packed form (integer hashing, compares, selects,
// conversions, table loads, MUFU) is done lane by lane exactly as in the scalar code. Each lane's operation sequence is
// the scalar one, so results stay bit-identical to the strict oracle.
//
// The node library is written once, as templates over the value type F (float or fpair); `pack<F>` names the matching
// integer and predicate types. This header supplies fpair / ipair / bpair and the pack overload of every primitive of
// fn2_device.cuh; the scalar `select` / comparison helpers it also defines let one source text serve both widths.
#pragma once

namespace fn2
{
    struct fpair
    {
        float x, y;
        __device__ __forceinline__ fpair() {}
        __device__ __forceinline__ fpair( float v ) : x( v ), y( v ) {}           // broadcast (implicit: literals mix freely)
        __device__ __forceinline__ fpair( float a, float b ) : x( a ), y( b ) {}
    };
    struct ipair
    {
        i32 x, y;
        __device__ __forceinline__ ipair() {}
        __device__ __forceinline__ ipair( i32 v ) : x( v ), y( v ) {}
        __device__ __forceinline__ ipair( i32 a, i32 b ) : x( a ), y( b ) {}
    };
    struct bpair
    {
        bool x, y;
        __device__ __forceinline__ bpair() {}
        __device__ __forceinline__ bpair( bool v ) : x( v ), y( v ) {}
        __device__ __forceinline__ bpair( bool a, bool b ) : x( a ), y( b ) {}
    };

    template<class F> struct pack;
    template<> struct pack<float> { typedef i32 I; typedef bool B; static constexpr int width = 1; };
    template<> struct pack<fpair> { typedef ipair I; typedef bpair B; static constexpr int width = 2; };

    FN2_DI float2 raw( fpair a ) { return make_float2( a.x, a.y ); }
    FN2_DI fpair wrap( float2 a ) { return fpair( a.x, a.y ); }
    FN2_DI float lane( float v, int ) { return v; }
    FN2_DI float lane( fpair v, int k ) { return k ? v.y : v.x; }

    // ------------------------------------------------------------------ packed FP32 arithmetic (one instruction for both lanes)
    FN2_DI fpair operator-( fpair a ) { return fpair( -a.x, -a.y ); } // folds into the consumer's negate modifier
    FN2_DI fpair operator+( fpair a, fpair b ) { return wrap( __fadd2_rn( raw( a ), raw( b ) ) ); }
    FN2_DI fpair operator-( fpair a, fpair b ) { return wrap( __fadd2_rn( raw( a ), raw( -b ) ) ); }
#if FN2_RELAXED
    FN2_DI fpair operator*( fpair a, fpair b ) { return wrap( __fmul2_rn( raw( a ), raw( b ) ) ); }
#else
    // ptxas (12.9) contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 whatever --fmad says and in spite of the explicit
    // .rn -- a single rounding where the reference has two. A packed product is therefore issued as FFMA2( a, b, -0.0 )
    // with the -0.0 read from constant memory so that ptxas cannot fold it away: RN( a * b + -0 ) == RN( a * b ) for every
    // input including signed zeros, the instruction count and pipe are those of FMUL2, and there is no FMUL2 left to fuse.
    __constant__ float kOpaqueNegZero = -0.0f;
    FN2_DI fpair operator*( fpair a, fpair b )
    {
        const float z = kOpaqueNegZero;
        return wrap( __ffma2_rn( raw( a ), raw( b ), make_float2( z, z ) ) );
    }
#endif
    FN2_DI fpair& operator+=( fpair& a, fpair b ) { return a = a + b; }
    FN2_DI fpair& operator-=( fpair& a, fpair b ) { return a = a - b; }
    FN2_DI fpair& operator*=( fpair& a, fpair b ) { return a = a * b; }
    FN2_DI fpair fma_( fpair a, fpair b, fpair c ) { return wrap( __ffma2_rn( raw( a ), raw( b ), raw( c ) ) ); }
    FN2_DI fpair fnma_( fpair a, fpair b, fpair c ) { return wrap( __ffma2_rn( raw( -a ), raw( b ), raw( c ) ) ); }
    FN2_DI fpair fms_( fpair a, fpair b, fpair c ) { return wrap( __ffma2_rn( raw( a ), raw( b ), raw( -c ) ) ); }
    FN2_DI fpair abs_( fpair v ) { return fpair( fabsf( v.x ), fabsf( v.y ) ); }  // folds into the consumer's |abs| modifier

    // ------------------------------------------------------------------ lane-by-lane forms of everything else
#define FN2_LANES1( RET, NAME, T ) FN2_DI RET NAME( T a ) { return RET( NAME( a.x ), NAME( a.y ) ); }
#define FN2_LANES2( RET, NAME, T, U ) FN2_DI RET NAME( T a, U b ) { return RET( NAME( a.x, b.x ), NAME( a.y, b.y ) ); }
    FN2_LANES1( ipair, fbits, fpair )
    FN2_LANES1( fpair, bitsf, ipair )
    FN2_LANES2( fpair, fxor, fpair, ipair )
    FN2_DI fpair fxor( fpair v, i32 bits ) { return fpair( fxor( v.x, bits ), fxor( v.y, bits ) ); }
    FN2_LANES2( ipair, iadd, ipair, ipair )
    FN2_LANES2( ipair, isub, ipair, ipair )
    FN2_LANES2( ipair, imul, ipair, ipair )
    FN2_DI ipair shl( ipair v, int s ) { return ipair( shl( v.x, s ), shl( v.y, s ) ); }
    FN2_LANES1( ipair, cvt_i, fpair )
    FN2_LANES1( fpair, cvt_f, ipair )
    FN2_LANES1( fpair, floor_, fpair )
    FN2_LANES1( fpair, round_, fpair )
    FN2_LANES2( fpair, min_, fpair, fpair )
    FN2_LANES2( fpair, max_, fpair, fpair )
    FN2_LANES2( fpair, min_x86, fpair, fpair )
    FN2_LANES2( fpair, max_x86, fpair, fpair )
    FN2_LANES1( fpair, sqrt_, fpair )
    FN2_LANES2( fpair, div_, fpair, fpair )
    FN2_LANES1( fpair, inv_sqrt, fpair )
    FN2_LANES1( fpair, reciprocal, fpair )

    template<int SHIFT, int BITS>
    FN2_DI fpair field_f( ipair v )
    {
        constexpr i32 mask = ( ( 1 << BITS ) - 1 ) << SHIFT;
        constexpr i32 magic = ( 127 + 23 - SHIFT ) << 23;
        const i32 mx = FN2_LOP3_AND_OR( v.x, mask, magic ), my = FN2_LOP3_AND_OR( v.y, mask, magic );
        return fpair( __int_as_float( mx ), __int_as_float( my ) ) - fpair( __int_as_float( magic ) ); // the subtraction is packed
    }
    template<int SHIFT, int BITS>
    FN2_DI fpair field_minus( ipair v, float centre )
    {
        constexpr i32 mask = ( ( 1 << BITS ) - 1 ) << SHIFT;
        constexpr i32 magic = ( 127 + 23 - SHIFT ) << 23;
        const i32 mx = FN2_LOP3_AND_OR( v.x, mask, magic ), my = FN2_LOP3_AND_OR( v.y, mask, magic );
        return fpair( __int_as_float( mx ), __int_as_float( my ) ) - fpair( __int_as_float( magic ) + centre );
    }

    FN2_DI fpair inv_sqrt_bounded( fpair v ) // see the scalar form: MUFU per lane, the Newton steps packed
    {
#if FN2_RELAXED || defined( FN2_HOST_EMULATION )
        return fpair( inv_sqrt_bounded( v.x ), inv_sqrt_bounded( v.y ) );
#else
        const fpair y( rsqrt_approx( v.x ), rsqrt_approx( v.y ) );
        fpair g = v * y;
        const fpair h = fpair( 0.5f ) * y;
        g = fma_( fnma_( g, g, v ), h, g );
        const fpair r( rcp_approx( g.x ), rcp_approx( g.y ) );
        return fma_( r, fnma_( g, r, fpair( 1.0f ) ), r );
#endif
    }

    // integer operators, lane by lane (there are no packed int32 instructions)
#define FN2_IOP( OP ) \
    FN2_DI ipair operator OP( ipair a, ipair b ) { return ipair( a.x OP b.x, a.y OP b.y ); }
    FN2_IOP( & )
    FN2_IOP( | )
    FN2_IOP( ^ )
    FN2_IOP( + )
    FN2_IOP( - )
#undef FN2_IOP
    FN2_DI ipair operator~( ipair a ) { return ipair( ~a.x, ~a.y ); }
    FN2_DI ipair operator>>( ipair a, int s ) { return ipair( a.x >> s, a.y >> s ); } // arithmetic, like the scalar i32 >>
    FN2_DI ipair& operator&=( ipair& a, ipair b ) { return a = a & b; }
    FN2_DI ipair& operator|=( ipair& a, ipair b ) { return a = a | b; }
    FN2_DI ipair& operator^=( ipair& a, ipair b ) { return a = a ^ b; }

    // comparisons -> lane predicates
#define FN2_CMP( OP ) \
    FN2_DI bpair operator OP( fpair a, fpair b ) { return bpair( a.x OP b.x, a.y OP b.y ); } \
    FN2_DI bpair operator OP( ipair a, ipair b ) { return bpair( a.x OP b.x, a.y OP b.y ); }
    FN2_CMP( < )
    FN2_CMP( > )
    FN2_CMP( <= )
    FN2_CMP( >= )
    FN2_CMP( == )
    FN2_CMP( != )
#undef FN2_CMP
    FN2_DI bpair operator&&( bpair a, bpair b ) { return bpair( a.x && b.x, a.y && b.y ); }
    FN2_DI bpair operator||( bpair a, bpair b ) { return bpair( a.x || b.x, a.y || b.y ); }
    FN2_DI bpair operator!( bpair a ) { return bpair( !a.x, !a.y ); }

    // select( c, a, b ) = c ? a : b, for both widths
    FN2_DI float select( bool c, float a, float b ) { return c ? a : b; }
    FN2_DI i32 select( bool c, i32 a, i32 b ) { return c ? a : b; }
    FN2_DI fpair select( bpair c, fpair a, fpair b ) { return fpair( c.x ? a.x : b.x, c.y ? a.y : b.y ); }
    FN2_DI ipair select( bpair c, ipair a, ipair b ) { return ipair( c.x ? a.x : b.x, c.y ? a.y : b.y ); }
    FN2_DI bpair select( bpair c, bpair a, bpair b ) { return bpair( c.x ? a.x : b.x, c.y ? a.y : b.y ); }
    FN2_DI bool select( bool c, bool a, bool b ) { return c ? a : b; }
    // FS::SelectHighBit
    FN2_DI float sel_hi( i32 selector, float a, float b ) { return selector < 0 ? a : b; }
    FN2_DI fpair sel_hi( ipair selector, fpair a, fpair b ) { return fpair( selector.x < 0 ? a.x : b.x, selector.y < 0 ? a.y : b.y ); }
    // mask & value (FS::Masked) and the sign bit of a predicate
    FN2_DI i32 signbit_if( bool c ) { return c ? (i32)0x80000000u : 0; }
    FN2_DI ipair signbit_if( bpair c ) { return ipair( signbit_if( c.x ), signbit_if( c.y ) ); }
#undef FN2_LANES1
#undef FN2_LANES2
}
