// fastnoise2_b200 -- device node library: DomainWarpGradient displacement field.
// Restates include/FastNoise/Generators/DomainWarp.inl:33-206: a value-noise-like vector field whose
// components are bit-fields of the corner hashes, blended with Hermite weights. `seed` already includes the
// node's seed offset, (x,y,..) is the position times the warp frequency, (xOut,yOut,..) is the position being
// displaced and (ax,ay,..) the per-axis amplitude scaling. Returns the squared length of the unit-range warp
// vector (the "strength" the fractal warps feed back into the amplitude).
#pragma once

namespace fn2
{
    template<class F>
    FN2_DI F warp_gradient2( i32 seed, F warpAmp, F x, F y, F& xOut, F& yOut, float ax, float ay )
    {
        typedef typename pack<F>::I I;
        typedef typename pack<F>::B B;
        F xs = floor_( x ), ys = floor_( y );
        I x0 = imul( cvt_i( xs ), kPX ), y0 = imul( cvt_i( ys ), kPY );
        I x1 = iadd( x0, kPX ), y1 = iadd( y0, kPY );
        F xs1 = hermite( x - xs ), ys1 = hermite( y - ys );
        F xs0 = 1.0f - xs1, ys0 = 1.0f - ys1;
        const F normalise = 1.0f / ( 0xffff / 2.0f );

        F xWarp, yWarp;
        {
            I h = hash_hb( seed, x0, y0 );
            F c = xs0 * ys0;
            xWarp = c * field_f<0, 16>( h );
            yWarp = c * cvt_f( lsr( h, 16 ) );
        }
#define FN2_CORNER( XS, YS, XP, YP ) { I h = hash_hb( seed, XP, YP ); F c = XS * YS; \
            xWarp = fma_( c, field_f<0, 16>( h ), xWarp ); yWarp = fma_( c, cvt_f( lsr( h, 16 ) ), yWarp ); }
        FN2_CORNER( xs1, ys0, x1, y0 )
        FN2_CORNER( xs0, ys1, x0, y1 )
        FN2_CORNER( xs1, ys1, x1, y1 )
#undef FN2_CORNER

        xWarp = fms_( xWarp, normalise, 1.0f );
        yWarp = fms_( yWarp, normalise, 1.0f );
        xOut = fma_( xWarp, warpAmp * ax, xOut );
        yOut = fma_( yWarp, warpAmp * ay, yOut );
        return fma_( xWarp, xWarp, yWarp * yWarp );
    }

    template<class F>
    FN2_DI F warp_gradient3( i32 seed, F warpAmp, F x, F y, F z, F& xOut, F& yOut, F& zOut, float ax, float ay, float az )
    {
        typedef typename pack<F>::I I;
        typedef typename pack<F>::B B;
        F xs = floor_( x ), ys = floor_( y ), zs = floor_( z );
        I x0 = imul( cvt_i( xs ), kPX ), y0 = imul( cvt_i( ys ), kPY ), z0 = imul( cvt_i( zs ), kPZ );
        I x1 = iadd( x0, kPX ), y1 = iadd( y0, kPY ), z1 = iadd( z0, kPZ );
        F xs1 = hermite( x - xs ), ys1 = hermite( y - ys ), zs1 = hermite( z - zs );
        F xs0 = 1.0f - xs1, ys0 = 1.0f - ys1, zs0 = 1.0f - zs1;
        const F normalise = 1.0f / ( 0x3ff / 2.0f );

        F xWarp, yWarp, zWarp;
        {
            I h = hash_hb( seed, x0, y0, z0 );
            F c = xs0 * ys0 * zs0;
            xWarp = c * field_f<0, 10>( h );
            yWarp = c * field_f<11, 10>( h );
            zWarp = c * cvt_f( lsr( h, 22 ) );
        }
#define FN2_CORNER( XS, YS, ZS, XP, YP, ZP ) { I h = hash_hb( seed, XP, YP, ZP ); F c = XS * YS * ZS; \
            xWarp = fma_( c, field_f<0, 10>( h ), xWarp ); yWarp = fma_( c, field_f<11, 10>( h ), yWarp ); zWarp = fma_( c, cvt_f( lsr( h, 22 ) ), zWarp ); }
        FN2_CORNER( xs1, ys0, zs0, x1, y0, z0 )
        FN2_CORNER( xs0, ys1, zs0, x0, y1, z0 )
        FN2_CORNER( xs1, ys1, zs0, x1, y1, z0 )
        FN2_CORNER( xs0, ys0, zs1, x0, y0, z1 )
        FN2_CORNER( xs1, ys0, zs1, x1, y0, z1 )
        FN2_CORNER( xs0, ys1, zs1, x0, y1, z1 )
        FN2_CORNER( xs1, ys1, zs1, x1, y1, z1 )
#undef FN2_CORNER

        xWarp = fms_( xWarp, normalise, 1.0f );
        yWarp = fms_( yWarp, normalise, 1.0f );
        zWarp = fms_( zWarp, normalise, 1.0f );
        xOut = fma_( xWarp, warpAmp * ax, xOut );
        yOut = fma_( yWarp, warpAmp * ay, yOut );
        zOut = fma_( zWarp, warpAmp * az, zOut );
        return fma_( xWarp, xWarp, fma_( yWarp, yWarp, zWarp * zWarp ) );
    }

#if FN2_AXIS_PACK
    // One point per value with the FP32 chains packed inside the point (see "axis-packed forms" in fn2_device.cuh): (y, z) through
    // the cell-relative offsets and Hermite curves, the corner weights of an x0 / x1 corner pair in one register pair, and the x and y
    // components of the displacement through field conversion, accumulation and the final scaling. [DomainWarp.inl:77-132]
    template<int SHIFT_X, int BITS_X, int SHIFT_Y, int BITS_Y>
    FN2_DI fpair field_f2( i32 v )
    {
        constexpr i32 maskX = ( ( 1 << BITS_X ) - 1 ) << SHIFT_X, magicX = ( 127 + 23 - SHIFT_X ) << 23;
        constexpr i32 maskY = ( ( 1 << BITS_Y ) - 1 ) << SHIFT_Y, magicY = ( 127 + 23 - SHIFT_Y ) << 23;
        const i32 mx = FN2_LOP3_AND_OR( v, maskX, magicX ), my = FN2_LOP3_AND_OR( v, maskY, magicY );
        return fpair( __int_as_float( mx ), __int_as_float( my ) ) - fpair( __int_as_float( magicX ), __int_as_float( magicY ) );
    }

    FN2_DI float warp_gradient3( i32 seed, float warpAmp, float x, float y, float z, float& xOut, float& yOut, float& zOut, float ax, float ay, float az )
    {
        const float xs = floor_( x );
        const fpair yz( y, z ), yzs( floor_( y ), floor_( z ) );
        const i32 x0 = imul( cvt_i( xs ), kPX ), y0 = imul( cvt_i( yzs.x ), kPY ), z0 = imul( cvt_i( yzs.y ), kPZ );
        const i32 x1 = iadd( x0, kPX ), y1 = iadd( y0, kPY ), z1 = iadd( z0, kPZ );
        const float xs1 = hermite( x - xs );
        const fpair wx( 1.0f - xs1, xs1 );                       // (xs0, xs1): lanes = the x0 / x1 corner of a pair
        const fpair yzs1 = hermite( yz - yzs );                  // (ys1, zs1)
        const fpair yzs0 = fpair( 1.0f ) - yzs1;                 // (ys0, zs0)
        const fpair wy0 = wx * fpair( yzs0.x ), wy1 = wx * fpair( yzs1.x );
        const fpair c00 = wy0 * fpair( yzs0.y ), c10 = wy1 * fpair( yzs0.y ), c01 = wy0 * fpair( yzs1.y ), c11 = wy1 * fpair( yzs1.y );
        const float normalise = 1.0f / ( 0x3ff / 2.0f );

        fpair xyWarp;
        float zWarp;
        {
            const i32 h = hash_hb( seed, x0, y0, z0 );
            xyWarp = fpair( c00.x ) * field_f2<0, 10, 11, 10>( h );
            zWarp = c00.x * cvt_f( lsr( h, 22 ) );
        }
#define FN2_CORNER( C, XP, YP, ZP ) { const i32 h = hash_hb( seed, XP, YP, ZP ); \
            xyWarp = fma_( fpair( C ), field_f2<0, 10, 11, 10>( h ), xyWarp ); zWarp = fma_( C, cvt_f( lsr( h, 22 ) ), zWarp ); }
        FN2_CORNER( c00.y, x1, y0, z0 )
        FN2_CORNER( c10.x, x0, y1, z0 )
        FN2_CORNER( c10.y, x1, y1, z0 )
        FN2_CORNER( c01.x, x0, y0, z1 )
        FN2_CORNER( c01.y, x1, y0, z1 )
        FN2_CORNER( c11.x, x0, y1, z1 )
        FN2_CORNER( c11.y, x1, y1, z1 )
#undef FN2_CORNER

        xyWarp = fms_( xyWarp, fpair( normalise ), fpair( 1.0f ) );
        zWarp = fms_( zWarp, normalise, 1.0f );
        const fpair xyOut = fma_( xyWarp, fpair( warpAmp * ax, warpAmp * ay ), fpair( xOut, yOut ) );
        xOut = xyOut.x;
        yOut = xyOut.y;
        zOut = fma_( zWarp, warpAmp * az, zOut );
        return fma_( xyWarp.x, xyWarp.x, fma_( xyWarp.y, xyWarp.y, zWarp * zWarp ) );
    }
#endif

    template<class F>
    FN2_DI F warp_gradient4( i32 seed, F warpAmp, F x, F y, F z, F w,
                                 F& xOut, F& yOut, F& zOut, F& wOut, float ax, float ay, float az, float aw )
    {
        typedef typename pack<F>::I I;
        typedef typename pack<F>::B B;
        F xs = floor_( x ), ys = floor_( y ), zs = floor_( z ), ws = floor_( w );
        I x0 = imul( cvt_i( xs ), kPX ), y0 = imul( cvt_i( ys ), kPY ), z0 = imul( cvt_i( zs ), kPZ ), w0 = imul( cvt_i( ws ), kPW );
        I x1 = iadd( x0, kPX ), y1 = iadd( y0, kPY ), z1 = iadd( z0, kPZ ), w1 = iadd( w0, kPW );
        F xs1 = hermite( x - xs ), ys1 = hermite( y - ys ), zs1 = hermite( z - zs ), ws1 = hermite( w - ws );
        F xs0 = 1.0f - xs1, ys0 = 1.0f - ys1, zs0 = 1.0f - zs1, ws0 = 1.0f - ws1;
        const F normalise = 1.0f / ( 0xff / 2.0f );

        F xWarp, yWarp, zWarp, wWarp;
        {
            I h = hash_hb( seed, x0, y0, z0, w0 );
            F c = xs0 * ys0 * zs0 * ws0;
            xWarp = c * field_f<0, 8>( h );
            yWarp = c * field_f<8, 8>( h );
            zWarp = c * cvt_f( ( h >> 16 ) & 0xff );
            wWarp = c * cvt_f( lsr( h, 24 ) );
        }
#define FN2_CORNER( XS, YS, ZS, WS, XP, YP, ZP, WP ) { I h = hash_hb( seed, XP, YP, ZP, WP ); F c = XS * YS * ZS * WS; \
            xWarp = fma_( c, field_f<0, 8>( h ), xWarp ); yWarp = fma_( c, field_f<8, 8>( h ), yWarp ); \
            zWarp = fma_( c, cvt_f( ( h >> 16 ) & 0xff ), zWarp ); wWarp = fma_( c, cvt_f( lsr( h, 24 ) ), wWarp ); }
        FN2_CORNER( xs1, ys0, zs0, ws0, x1, y0, z0, w0 )
        FN2_CORNER( xs0, ys1, zs0, ws0, x0, y1, z0, w0 )
        FN2_CORNER( xs1, ys1, zs0, ws0, x1, y1, z0, w0 )
        FN2_CORNER( xs0, ys0, zs1, ws0, x0, y0, z1, w0 )
        FN2_CORNER( xs1, ys0, zs1, ws0, x1, y0, z1, w0 )
        FN2_CORNER( xs0, ys1, zs1, ws0, x0, y1, z1, w0 )
        FN2_CORNER( xs1, ys1, zs1, ws0, x1, y1, z1, w0 )
        FN2_CORNER( xs0, ys0, zs0, ws1, x0, y0, z0, w1 )
        FN2_CORNER( xs1, ys0, zs0, ws1, x1, y0, z0, w1 )
        FN2_CORNER( xs0, ys1, zs0, ws1, x0, y1, z0, w1 )
        FN2_CORNER( xs1, ys1, zs0, ws1, x1, y1, z0, w1 )
        FN2_CORNER( xs0, ys0, zs1, ws1, x0, y0, z1, w1 )
        FN2_CORNER( xs1, ys0, zs1, ws1, x1, y0, z1, w1 )
        FN2_CORNER( xs0, ys1, zs1, ws1, x0, y1, z1, w1 )
        FN2_CORNER( xs1, ys1, zs1, ws1, x1, y1, z1, w1 )
#undef FN2_CORNER

        xWarp = fms_( xWarp, normalise, 1.0f );
        yWarp = fms_( yWarp, normalise, 1.0f );
        zWarp = fms_( zWarp, normalise, 1.0f );
        wWarp = fms_( wWarp, normalise, 1.0f );
        xOut = fma_( xWarp, warpAmp * ax, xOut );
        yOut = fma_( yWarp, warpAmp * ay, yOut );
        zOut = fma_( zWarp, warpAmp * az, zOut );
        wOut = fma_( wWarp, warpAmp * aw, wOut );
        return fma_( xWarp, xWarp, fma_( yWarp, yWarp, fma_( zWarp, zWarp, wWarp * wWarp ) ) );
    }
}
