// ORACLE / TEST INFRASTRUCTURE ONLY -- see DispatchClass.h in this directory.
// Free functions of the FastSIMD API subset used by the FastNoise2 node sources
// (census: grep -rhoE 'FS::[A-Za-z0-9_]+' over /root/reference/include + src).
#pragma once

namespace FS
{
    // ---------------------------------------------------------------- native helpers
#if FN2SHIM_LANES == 8
    using nf_t = __m256;
    using ni_t = __m256i;
    FS_FORCEINLINE vf_t n_fma( vf_t a, vf_t b, vf_t c )  { return (vf_t)_mm256_fmadd_ps( (nf_t)a, (nf_t)b, (nf_t)c ); }
    FS_FORCEINLINE vf_t n_fnma( vf_t a, vf_t b, vf_t c ) { return (vf_t)_mm256_fnmadd_ps( (nf_t)a, (nf_t)b, (nf_t)c ); }
    FS_FORCEINLINE vf_t n_fms( vf_t a, vf_t b, vf_t c )  { return (vf_t)_mm256_fmsub_ps( (nf_t)a, (nf_t)b, (nf_t)c ); }
    FS_FORCEINLINE vf_t n_floor( vf_t a ) { return (vf_t)_mm256_round_ps( (nf_t)a, _MM_FROUND_TO_NEG_INF | _MM_FROUND_NO_EXC ); }
    FS_FORCEINLINE vf_t n_round( vf_t a ) { return (vf_t)_mm256_round_ps( (nf_t)a, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC ); }
    FS_FORCEINLINE vf_t n_trunc( vf_t a ) { return (vf_t)_mm256_round_ps( (nf_t)a, _MM_FROUND_TO_ZERO | _MM_FROUND_NO_EXC ); }
    FS_FORCEINLINE vi_t n_cvt_f2i( vf_t a ) { return (vi_t)_mm256_cvtps_epi32( (nf_t)a ); } // round-to-nearest-even (MXCSR default)
    FS_FORCEINLINE vf_t n_cvt_i2f( vi_t a ) { return (vf_t)_mm256_cvtepi32_ps( (ni_t)a ); }
    FS_FORCEINLINE vf_t n_min( vf_t a, vf_t b ) { return (vf_t)_mm256_min_ps( (nf_t)a, (nf_t)b ); }
    FS_FORCEINLINE vf_t n_max( vf_t a, vf_t b ) { return (vf_t)_mm256_max_ps( (nf_t)a, (nf_t)b ); }
    FS_FORCEINLINE vf_t n_sqrt( vf_t a ) { return (vf_t)_mm256_sqrt_ps( (nf_t)a ); }
    FS_FORCEINLINE vf_t n_rsqrt_approx( vf_t a ) { return (vf_t)_mm256_rsqrt_ps( (nf_t)a ); }
    FS_FORCEINLINE vf_t n_rcp_approx( vf_t a ) { return (vf_t)_mm256_rcp_ps( (nf_t)a ); }
    FS_FORCEINLINE vf_t n_blend( vi_t m, vf_t a, vf_t b ) { return (vf_t)_mm256_blendv_ps( (nf_t)b, (nf_t)a, (nf_t)m ); }
    FS_FORCEINLINE bool n_any( vi_t m ) { return _mm256_movemask_ps( (nf_t)m ) != 0; }
#elif FN2SHIM_LANES == 16
    using nf_t = __m512;
    using ni_t = __m512i;
    FS_FORCEINLINE vf_t n_fma( vf_t a, vf_t b, vf_t c )  { return (vf_t)_mm512_fmadd_ps( (nf_t)a, (nf_t)b, (nf_t)c ); }
    FS_FORCEINLINE vf_t n_fnma( vf_t a, vf_t b, vf_t c ) { return (vf_t)_mm512_fnmadd_ps( (nf_t)a, (nf_t)b, (nf_t)c ); }
    FS_FORCEINLINE vf_t n_fms( vf_t a, vf_t b, vf_t c )  { return (vf_t)_mm512_fmsub_ps( (nf_t)a, (nf_t)b, (nf_t)c ); }
    FS_FORCEINLINE vf_t n_floor( vf_t a ) { return (vf_t)_mm512_roundscale_ps( (nf_t)a, _MM_FROUND_TO_NEG_INF | _MM_FROUND_NO_EXC ); }
    FS_FORCEINLINE vf_t n_round( vf_t a ) { return (vf_t)_mm512_roundscale_ps( (nf_t)a, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC ); }
    FS_FORCEINLINE vf_t n_trunc( vf_t a ) { return (vf_t)_mm512_roundscale_ps( (nf_t)a, _MM_FROUND_TO_ZERO | _MM_FROUND_NO_EXC ); }
    FS_FORCEINLINE vi_t n_cvt_f2i( vf_t a ) { return (vi_t)_mm512_cvtps_epi32( (nf_t)a ); }
    FS_FORCEINLINE vf_t n_cvt_i2f( vi_t a ) { return (vf_t)_mm512_cvtepi32_ps( (ni_t)a ); }
    FS_FORCEINLINE vf_t n_min( vf_t a, vf_t b ) { return (vf_t)_mm512_min_ps( (nf_t)a, (nf_t)b ); }
    FS_FORCEINLINE vf_t n_max( vf_t a, vf_t b ) { return (vf_t)_mm512_max_ps( (nf_t)a, (nf_t)b ); }
    FS_FORCEINLINE vf_t n_sqrt( vf_t a ) { return (vf_t)_mm512_sqrt_ps( (nf_t)a ); }
    FS_FORCEINLINE vf_t n_rsqrt_approx( vf_t a ) { return (vf_t)_mm512_rsqrt14_ps( (nf_t)a ); }
    FS_FORCEINLINE vf_t n_rcp_approx( vf_t a ) { return (vf_t)_mm512_rcp14_ps( (nf_t)a ); }
    FS_FORCEINLINE vf_t n_blend( vi_t m, vf_t a, vf_t b ) { return m ? a : b; }
    FS_FORCEINLINE bool n_any( vi_t m ) { return _mm512_test_epi32_mask( (ni_t)m, (ni_t)m ) != 0; }
#else
#error "FN2SHIM_LANES must be 8 or 16"
#endif

    FS_FORCEINLINE vi_t n_blend_i( vi_t m, vi_t a, vi_t b ) { return (vi_t)n_blend( m, (vf_t)a, (vf_t)b ); }

    template<typename T, std::size_t N, FeatureSet S>
    FS_FORCEINLINE FN2_R n_select( vi_t m, FN2_R a, FN2_R b )
    {
        return FN2_R::fromBits( n_blend_i( m, a.bits(), b.bits() ) );
    }

    // ---------------------------------------------------------------- fused multiply-add family (single rounding)
    FS_FORCEINLINE f32v FMulAdd( f32v a, f32v b, f32v c )  { return f32v( n_fma( a.v, b.v, c.v ), 0 ); }   //  a*b + c
    FS_FORCEINLINE f32v FNMulAdd( f32v a, f32v b, f32v c ) { return f32v( n_fnma( a.v, b.v, c.v ), 0 ); }  // -a*b + c
    FS_FORCEINLINE f32v FMulSub( f32v a, f32v b, f32v c )  { return f32v( n_fms( a.v, b.v, c.v ), 0 ); }   //  a*b - c

    // ---------------------------------------------------------------- min/max/abs/rounding
    FS_FORCEINLINE f32v Min( f32v a, f32v b ) { return f32v( n_min( a.v, b.v ), 0 ); }
    FS_FORCEINLINE f32v Max( f32v a, f32v b ) { return f32v( n_max( a.v, b.v ), 0 ); }
    FS_FORCEINLINE i32v Min( i32v a, i32v b ) { return i32v( a.v < b.v ? a.v : b.v, 0 ); }
    FS_FORCEINLINE i32v Max( i32v a, i32v b ) { return i32v( a.v > b.v ? a.v : b.v, 0 ); }
    FS_FORCEINLINE f32v Abs( f32v a ) { return f32v::fromBits( a.bits() & 0x7fffffff ); }
    FS_FORCEINLINE i32v Abs( i32v a ) { return i32v( a.v < 0 ? -a.v : a.v, 0 ); }
    FS_FORCEINLINE f32v Floor( f32v a ) { return f32v( n_floor( a.v ), 0 ); }
    FS_FORCEINLINE f32v Round( f32v a ) { return f32v( n_round( a.v ), 0 ); } // ties to even
    FS_FORCEINLINE f32v Sqrt( f32v a ) { return f32v( n_sqrt( a.v ), 0 ); }

    FS_FORCEINLINE f32v InvSqrt( f32v a )
    {
#if FN2SHIM_RELAXED
        return f32v( n_rsqrt_approx( a.v ), 0 );
#else
        return f32v( ( vf_t {} + 1.0f ) / n_sqrt( a.v ), 0 ); // IEEE sqrt then IEEE divide
#endif
    }

    FS_FORCEINLINE f32v Reciprocal( f32v a )
    {
#if FN2SHIM_RELAXED
        return f32v( n_rcp_approx( a.v ), 0 );
#else
        return f32v( ( vf_t {} + 1.0f ) / a.v, 0 );
#endif
    }

    // ---------------------------------------------------------------- select / masked family
    FN2_T FS_FORCEINLINE FN2_R Select( MaskV m, FN2_R a, FN2_R b ) { return n_select( m.m, a, b ); }
    FN2_T FS_FORCEINLINE FN2_R Masked( MaskV m, FN2_R a ) { return FN2_R::fromBits( a.bits() & m.m ); }
    FN2_T FS_FORCEINLINE FN2_R InvMasked( MaskV m, FN2_R a ) { return FN2_R::fromBits( a.bits() & ~m.m ); }
    FN2_T FS_FORCEINLINE FN2_R MaskedAdd( MaskV m, FN2_R a, FN2_R b ) { return n_select( m.m, a + b, a ); }
    FN2_T FS_FORCEINLINE FN2_R InvMaskedAdd( MaskV m, FN2_R a, FN2_R b ) { return n_select( m.m, a, a + b ); }
    FN2_T FS_FORCEINLINE FN2_R MaskedSub( MaskV m, FN2_R a, FN2_R b ) { return n_select( m.m, a - b, a ); }
    FN2_T FS_FORCEINLINE FN2_R InvMaskedSub( MaskV m, FN2_R a, FN2_R b ) { return n_select( m.m, a, a - b ); }
    FN2_T FS_FORCEINLINE FN2_R MaskedMul( MaskV m, FN2_R a, FN2_R b ) { return n_select( m.m, a * b, a ); }
    FN2_T FS_FORCEINLINE FN2_R MaskedIncrement( MaskV m, FN2_R a ) { return n_select( m.m, a + FN2_R( 1 ), a ); }
    FN2_T FS_FORCEINLINE FN2_R MaskedDecrement( MaskV m, FN2_R a ) { return n_select( m.m, a - FN2_R( 1 ), a ); }
    FN2_T FS_FORCEINLINE FN2_R Increment( FN2_R a ) { return a + FN2_R( 1 ); }
    FN2_T FS_FORCEINLINE FN2_R Decrement( FN2_R a ) { return a - FN2_R( 1 ); }

    FS_FORCEINLINE bool AnyMask( MaskV m ) { return n_any( m.m ); }

    // the selector's sign bit picks `a`; selector may be an int or a float register
    template<typename TS, typename T, std::size_t N, FeatureSet S>
    FS_FORCEINLINE FN2_R SelectHighBit( Register<TS, N, S> sel, FN2_R a, FN2_R b )
    {
        return n_select( sel.bits() >> 31, a, b );
    }

    FN2_T FS_FORCEINLINE FN2_R BitwiseAndNot( FN2_R a, FN2_R b ) { return a & ~b; }
    FS_FORCEINLINE MaskV BitwiseAndNot( MaskV a, MaskV b ) { return a & ~b; }
    FS_FORCEINLINE i32v BitShiftRightZeroExtend( i32v a, int s ) { return i32v::fromBits( (vi_t)( (vu_t)a.bits() >> s ) ); }
    FS_FORCEINLINE f32v SignBit( f32v a ) { return f32v::fromBits( a.bits() & (vi_t)( vu_t {} + 0x80000000u ) ); }

    // ---------------------------------------------------------------- convert (value) / cast (bits)
    template<typename TO, typename T, std::size_t N, FeatureSet S>
    FS_FORCEINLINE Register<TO, N, S> Convert( FN2_R a )
    {
        if constexpr( std::is_same<TO, T>::value ) return a;
        else if constexpr( std::is_integral<TO>::value ) return Register<TO, N, S>( n_cvt_f2i( a.v ), 0 ); // nearest-even, NOT truncation
        else return Register<TO, N, S>( n_cvt_i2f( a.v ), 0 );
    }

    template<typename TO, typename T, std::size_t N, FeatureSet S, typename = std::enable_if_t<std::is_arithmetic<TO>::value>>
    FS_FORCEINLINE Register<TO, N, S> Cast( FN2_R a ) { return Register<TO, N, S>::fromBits( a.bits() ); }

    template<typename TO, typename = std::enable_if_t<std::is_arithmetic<TO>::value>>
    FS_FORCEINLINE Register<TO, kLanes> Cast( MaskV m ) { return Register<TO, kLanes>::fromBits( m.m ); }

    // Cast<FS::Mask<32>>( int32v ): lane sign bit -> mask (only reachable from x86 branches)
    template<typename TO, typename T, std::size_t N, FeatureSet S, typename = std::enable_if_t<!std::is_arithmetic<TO>::value>, typename = void>
    FS_FORCEINLINE MaskV Cast( FN2_R a ) { return { a.bits() >> 31 }; }

    // ---------------------------------------------------------------- memory
    template<typename R>
    FS_FORCEINLINE R Load( const typename R::ElementType* p ) { R r; std::memcpy( &r.v, p, sizeof( r.v ) ); return r; }
    FN2_T FS_FORCEINLINE void Store( T* p, FN2_R a ) { std::memcpy( p, &a.v, sizeof( a.v ) ); }
    template<typename R>
    FS_FORCEINLINE R LoadIncremented() { R r; for( std::size_t i = 0; i < R::ElementCount; i++ ) r.v[i] = (typename R::ElementType)i; return r; }
    FN2_T FS_FORCEINLINE T Extract0( FN2_R a ) { return a.v[0]; }

    // ---------------------------------------------------------------- transcendental restatements (parity unpinned vs upstream)
    // cos: |x| reduced to [0,2pi) with a float floor, folded to [0,pi/2] by quadrant, even degree-4 polynomial in a^2.
    FS_FORCEINLINE f32v Cos( f32v x )
    {
        f32v v = Abs( x );
        v -= Floor( v * f32v( 0.1591549f ) ) * f32v( 6.283185f );
        MaskV geHalfPi = v >= f32v( 1.570796f );
        MaskV gePi15 = v >= f32v( 4.7123889f );
        f32v a = Select( gePi15, f32v( 6.283185f ) - v, Select( geHalfPi, f32v( 3.141593f ) - v, v ) );
        f32v c = a * a;
        f32v r = FMulAdd( c, FMulAdd( c, f32v( 0.03679168f ), f32v( -0.49558072f ) ), f32v( 0.99940307f ) );
        return Select( BitwiseAndNot( geHalfPi, gePi15 ), -r, r );
    }
    FS_FORCEINLINE f32v Sin( f32v x ) { return Cos( f32v( 1.570796f ) - x ); }

    // pow(x,y) for x > 0: exp2( y * log2(x) ), fixed operation sequence so the CUDA path can mirror it bit for bit.
    inline float Fn2PowScalar( float x, float y )
    {
        std::int32_t xb; std::memcpy( &xb, &x, 4 );
        std::int32_t e = ( ( xb >> 23 ) & 0xff ) - 127;
        std::int32_t mb = ( xb & 0x007fffff ) | 0x3f800000;
        float m; std::memcpy( &m, &mb, 4 );
        if( m > 1.41421354f ) { m = m * 0.5f; e += 1; }
        float s = ( m - 1.0f ) / ( m + 1.0f );
        float s2 = s * s;
        float p = std::fma( s2, 0.111111112f, 0.142857149f );
        p = std::fma( s2, p, 0.2f );
        p = std::fma( s2, p, 0.333333343f );
        p = std::fma( s2, p, 1.0f );
        float lg = std::fma( ( s + s ) * p, 1.44269502f, (float)e );
        float z = y * lg;
        z = z < -126.0f ? -126.0f : z;
        z = z > 127.0f ? 127.0f : z;
        float n = std::floor( z + 0.5f );
        float u = ( z - n ) * 0.693147182f;
        float q = std::fma( u, 1.98412701e-4f, 1.38888892e-3f );
        q = std::fma( u, q, 8.33333377e-3f );
        q = std::fma( u, q, 4.16666679e-2f );
        q = std::fma( u, q, 0.166666672f );
        q = std::fma( u, q, 0.5f );
        q = std::fma( u, q, 1.0f );
        q = std::fma( u, q, 1.0f );
        std::int32_t sb = ( (std::int32_t)n + 127 ) << 23;
        float sc; std::memcpy( &sc, &sb, 4 );
        return q * sc;
    }
    FS_FORCEINLINE f32v Pow( f32v a, f32v b )
    {
        f32v r;
        for( std::size_t i = 0; i < kLanes; i++ ) r.v[i] = Fn2PowScalar( a.v[i], b.v[i] );
        return r;
    }

    // modulus: a - trunc(a/b)*b
    FS_FORCEINLINE f32v Modulus( f32v a, f32v b ) { return FNMulAdd( f32v( n_trunc( ( a / b ).v ), 0 ), b, a ); }
    FS_FORCEINLINE i32v Modulus( i32v a, i32v b ) { return i32v( a.v % b.v, 0 ); }

    // ---------------------------------------------------------------- ISA-specific branches
#if FN2SHIM_AVX512_PATHS
    // "tables" flavour: the node sources' AVX512_F branches are live. They build 16-entry lane tables with
    // FS::Constant and index them with the real _mm512_permutexvar_ps / _mm512_permutex2var_ps / _mm512_rol_epi32
    // intrinsics through FS::NativeExec; a 16-lane GCC vector IS a zmm register, so the arguments are only re-typed.
    static_assert( kLanes == 16, "the AVX512_F code paths index 16-entry tables: build with FN2SHIM_LANES=16" );
    template<typename T, typename... A>
    FS_FORCEINLINE Register<T, kLanes> Constant( A... a )
    {
        static_assert( sizeof...( A ) == kLanes, "one table entry per lane" );
        typename VecOf<T>::type v = { (T)a... };
        return Register<T, kLanes>( v, 0 );
    }
    FS_FORCEINLINE __m512 Fn2Native( f32v a ) { return (__m512)a.v; }
    FS_FORCEINLINE __m512i Fn2Native( i32v a ) { return (__m512i)a.v; }
    template<typename R, typename F, typename... A>
    FS_FORCEINLINE R NativeExec( F&& f, A... a ) { return R( (typename R::Vec)f( Fn2Native( a )... ), 0 ); }
#else
    // generic flavour: these names only appear inside discarded `if constexpr` branches
    template<typename T, typename... A>
    Register<T, kLanes> Constant( A... ) { return {}; }
    template<typename R, typename F, typename... A>
    R NativeExec( F&&, A&&... ) { return {}; }
#endif
}
