// fastnoise2_b200 -- device node library: CellularValue / CellularDistance / CellularLookup.
// Restates include/FastNoise/Generators/Cellular.inl:18-807. Each thread walks the 3^N neighbourhood of its
// point; cell points are the cell corner plus a jitter vector taken from bit-fields of the cell hash and
// normalised with FS::InvSqrt. Template parameters are the node's uniform settings, which the reference
// branches on per call and the generated program fixes at compile time:
//   DF     distance function (FastNoise::DistanceFunction value)
//   SIZEJ  "size jitter" active (constant != 0 or a generator attached, Cellular.inl:27)
//   TOP    highest index of the sorted-distance / value array that the node reads
//   PF     callable returning the Minkowski exponent for a delta vector (evaluated AT the delta, Utils.inl:330)
// Inputs: `seed` already has the node's seed offset added, positions are already multiplied by the
// frequency, `jitter` is kJitterND * GridJitter and `sizeJitter` is SizeJitter * (-1 / fieldMax).
#pragma once

namespace fn2
{
    constexpr float kJitter2D = 0.437016f;
    constexpr float kJitter3D = 0.396144f;
    constexpr float kJitter4D = 0.366025f;
    constexpr float kSizeJitterScale2D = -1.f / 0x3ff;
    constexpr float kSizeJitterScale3D = -1.f / 0xffff;
    constexpr float kSizeJitterScale4D = -1.f / 0xfffff;
    constexpr float kFltMax = 3.402823466e+38f;

    // Jitter direction of a cell from its hash and the factor that scales it to length `jitter`.
    // [Cellular.inl:57-60] 2D: two 11-bit fields
    template<class F>
    FN2_DI F cell_vector( typename pack<F>::I hash, F jitter, F& xd, F& yd )
    {
        typedef typename pack<F>::I I;
        typedef typename pack<F>::B B;
        xd = cvt_f( hash & 0x7ff ) - ( 0x7ff / 2.0f );
        yd = cvt_f( lsr( hash, 21 ) ) - ( 0x7ff / 2.0f );
        return jitter * inv_sqrt_bounded( fma_( xd, xd, yd * yd ) );
    }
    // [Cellular.inl:148-152] 3D: three 10-bit fields at bits 0, 11, 22
    template<class F>
    FN2_DI F cell_vector( typename pack<F>::I hash, F jitter, F& xd, F& yd, F& zd )
    {
        typedef typename pack<F>::I I;
        typedef typename pack<F>::B B;
        xd = cvt_f( hash & 0x3ff ) - ( 0x3ff / 2.0f );
        yd = field_minus<11, 10>( hash, 0x3ff / 2.0f );
        zd = cvt_f( lsr( hash, 22 ) ) - ( 0x3ff / 2.0f );
        return jitter * inv_sqrt_bounded( fma_( xd, xd, fma_( yd, yd, zd * zd ) ) );
    }
    // [Cellular.inl:247-252] 4D: four bytes
    template<class F>
    FN2_DI F cell_vector( typename pack<F>::I hash, F jitter, F& xd, F& yd, F& zd, F& wd )
    {
        typedef typename pack<F>::I I;
        typedef typename pack<F>::B B;
        xd = cvt_f( hash & 0xff ) - ( 0xff / 2.0f );
        yd = field_minus<8, 8>( hash, 0xff / 2.0f );
        zd = cvt_f( ( hash >> 16 ) & 0xff ) - ( 0xff / 2.0f );
        wd = cvt_f( lsr( hash, 24 ) ) - ( 0xff / 2.0f );
        return jitter * inv_sqrt_bounded( fma_( xd, xd, fma_( yd, yd, fma_( zd, zd, wd * wd ) ) ) );
    }

    // per-cell distance multiplier of the size jitter [Cellular.inl:66-70, 160-164, 260-264]
    template<class F>
    FN2_DI F size_jitter2( typename pack<F>::I hash, F sizeJitter ) { typedef typename pack<F>::I I; typedef typename pack<F>::B B; (void)sizeof( I ); (void)sizeof( B ); return fnma_( sizeJitter, field_f<11, 10>( hash ), 1.0f ); }
    template<class F>
    FN2_DI F size_jitter3( typename pack<F>::I hash, F sizeJitter ) { typedef typename pack<F>::I I; typedef typename pack<F>::B B; (void)sizeof( I ); (void)sizeof( B ); return fnma_( sizeJitter, cvt_f( hash & 0xffff ), 1.0f ); }
    template<class F>
    FN2_DI F size_jitter4( typename pack<F>::I hash, F sizeJitter ) { typedef typename pack<F>::I I; typedef typename pack<F>::B B; (void)sizeof( I ); (void)sizeof( B ); return fnma_( sizeJitter, cvt_f( hash & 0xfffff ), 1.0f ); }

    // insert into the ascending (distance, hash) list, entries 0..TOP [Cellular.inl:72-89]
    template<int TOP, class F>
    FN2_DI void value_insert( F ( &distance )[4], typename pack<F>::I ( &valueHash )[4], F newDistance, typename pack<F>::I newHash )
    {
        typedef typename pack<F>::I I;
        typedef typename pack<F>::B B;
#pragma unroll
        for( int i = 0; i <= TOP; i++ )
        {
            B closer = newDistance < distance[i];
            F localDistance = distance[i];
            I localHash = valueHash[i];
            distance[i] = select( closer, newDistance, distance[i] );
            valueHash[i] = select( closer, newHash, valueHash[i] );
            newDistance = select( closer, localDistance, newDistance );
            newHash = select( closer, localHash, newHash );
        }
    }

    // sorting network step of CellularDistance, entries 0..TOP [Cellular.inl:344-349]. A candidate distance is never NaN while
    // the position and the jitter / size jitter / Minkowski P inputs are finite, and FMNMX then equals minps / maxps; ORDERED
    // (chosen by the lowering when one of those inputs is a node that may return NaN or inf) keeps the reference's operand
    // order, in which a NaN candidate shifts the list: min( d, NaN ) = NaN, max( NaN, d[i-1] ) = d[i-1].
    template<int TOP, bool ORDERED, class F>
    FN2_DI void distance_insert( F ( &distance )[4], F newDistance )
    {
#pragma unroll
        for( int i = TOP; i > 0; i-- )
        {
            distance[i] = ORDERED ? max_x86( min_x86( distance[i], newDistance ), distance[i - 1] ) : max_( min_( distance[i], newDistance ), distance[i - 1] );
        }
        distance[0] = ORDERED ? min_x86( distance[0], newDistance ) : min_( distance[0], newDistance );
    }

    // ================================================================== CellularValue [Cellular.inl:18-292]
    // returns the hash of the TOP-th closest cell (caller converts and range-maps it)
    template<int DF, bool SIZEJ, int TOP, class PF, class F>
    FN2_DI typename pack<F>::I cellular_value2( i32 seed, F x, F y, F jitter, F sizeJitter, PF minkowskiP )
    {
        typedef typename pack<F>::I I;
        typedef typename pack<F>::B B;
        F distance[4] = { FN2_INF, FN2_INF, FN2_INF, FN2_INF };
        I valueHash[4] = { 0, 0, 0, 0 };

        I xc = iadd( cvt_i( x ), -1 ), ycBase = iadd( cvt_i( y ), -1 );
        F xcf = cvt_f( xc ) - x, ycfBase = cvt_f( ycBase ) - y;
        xc = imul( xc, kPX );
        ycBase = imul( ycBase, kPY );

        for( int xi = 0; xi < 3; xi++ )
        {
            F ycf = ycfBase;
            I yc = ycBase;
            for( int yi = 0; yi < 3; yi++ )
            {
                I hash = hash_hb( seed, xc, yc );
                F xd, yd;
                F invMag = cell_vector( hash, jitter, xd, yd );
                xd = fma_( xd, invMag, xcf );
                yd = fma_( yd, invMag, ycf );
                F newDistance = calc_distance<DF, true>( DF == 5 ? minkowskiP( xd, yd ) : 0.0f, xd, yd );
                if( SIZEJ ) newDistance *= size_jitter2( hash, sizeJitter );
                value_insert<TOP>( distance, valueHash, newDistance, hash );
                ycf += 1.0f;
                yc = iadd( yc, kPY );
            }
            xcf += 1.0f;
            xc = iadd( xc, kPX );
        }
        return valueHash[TOP];
    }

    template<int DF, bool SIZEJ, int TOP, class PF, class F>
    FN2_DI typename pack<F>::I cellular_value3( i32 seed, F x, F y, F z, F jitter, F sizeJitter, PF minkowskiP )
    {
        typedef typename pack<F>::I I;
        typedef typename pack<F>::B B;
        F distance[4] = { FN2_INF, FN2_INF, FN2_INF, FN2_INF };
        I valueHash[4] = { 0, 0, 0, 0 };

        I xc = iadd( cvt_i( x ), -1 ), ycBase = iadd( cvt_i( y ), -1 ), zcBase = iadd( cvt_i( z ), -1 );
        F xcf = cvt_f( xc ) - x, ycfBase = cvt_f( ycBase ) - y, zcfBase = cvt_f( zcBase ) - z;
        xc = imul( xc, kPX );
        ycBase = imul( ycBase, kPY );
        zcBase = imul( zcBase, kPZ );

        for( int xi = 0; xi < 3; xi++ )
        {
            F ycf = ycfBase;
            I yc = ycBase;
            for( int yi = 0; yi < 3; yi++ )
            {
                F zcf = zcfBase;
                I zc = zcBase;
                for( int zi = 0; zi < 3; zi++ )
                {
                    I hash = hash_hb( seed, xc, yc, zc );
                    F xd, yd, zd;
                    F invMag = cell_vector( hash, jitter, xd, yd, zd );
                    xd = fma_( xd, invMag, xcf );
                    yd = fma_( yd, invMag, ycf );
                    zd = fma_( zd, invMag, zcf );
                    F newDistance = calc_distance<DF, false>( DF == 5 ? minkowskiP( xd, yd, zd ) : 0.0f, xd, yd, zd );
                    if( SIZEJ ) newDistance *= size_jitter3( hash, sizeJitter );
                    value_insert<TOP>( distance, valueHash, newDistance, hash );
                    zcf += 1.0f;
                    zc = iadd( zc, kPZ );
                }
                ycf += 1.0f;
                yc = iadd( yc, kPY );
            }
            xcf += 1.0f;
            xc = iadd( xc, kPX );
        }
        return valueHash[TOP];
    }

    template<int DF, bool SIZEJ, int TOP, class PF, class F>
    FN2_DI typename pack<F>::I cellular_value4( i32 seed, F x, F y, F z, F w, F jitter, F sizeJitter, PF minkowskiP )
    {
        typedef typename pack<F>::I I;
        typedef typename pack<F>::B B;
        F distance[4] = { FN2_INF, FN2_INF, FN2_INF, FN2_INF };
        I valueHash[4] = { 0, 0, 0, 0 };

        I xc = iadd( cvt_i( x ), -1 ), ycBase = iadd( cvt_i( y ), -1 ), zcBase = iadd( cvt_i( z ), -1 ), wcBase = iadd( cvt_i( w ), -1 );
        F xcf = cvt_f( xc ) - x, ycfBase = cvt_f( ycBase ) - y, zcfBase = cvt_f( zcBase ) - z, wcfBase = cvt_f( wcBase ) - w;
        xc = imul( xc, kPX );
        ycBase = imul( ycBase, kPY );
        zcBase = imul( zcBase, kPZ );
        wcBase = imul( wcBase, kPW );

        for( int xi = 0; xi < 3; xi++ )
        {
            F ycf = ycfBase;
            I yc = ycBase;
            for( int yi = 0; yi < 3; yi++ )
            {
                F zcf = zcfBase;
                I zc = zcBase;
                for( int zi = 0; zi < 3; zi++ )
                {
                    F wcf = wcfBase;
                    I wc = wcBase;
                    for( int wi = 0; wi < 3; wi++ )
                    {
                        I hash = hash_hb( seed, xc, yc, zc, wc );
                        F xd, yd, zd, wd;
                        F invMag = cell_vector( hash, jitter, xd, yd, zd, wd );
                        xd = fma_( xd, invMag, xcf );
                        yd = fma_( yd, invMag, ycf );
                        zd = fma_( zd, invMag, zcf );
                        wd = fma_( wd, invMag, wcf );
                        F newDistance = calc_distance<DF, false>( DF == 5 ? minkowskiP( xd, yd, zd, wd ) : 0.0f, xd, yd, zd, wd );
                        if( SIZEJ ) newDistance *= size_jitter4( hash, sizeJitter );
                        value_insert<TOP>( distance, valueHash, newDistance, hash );
                        wcf += 1.0f;
                        wc = iadd( wc, kPW );
                    }
                    zcf += 1.0f;
                    zc = iadd( zc, kPZ );
                }
                ycf += 1.0f;
                yc = iadd( yc, kPY );
            }
            xcf += 1.0f;
            xc = iadd( xc, kPX );
        }
        return valueHash[TOP];
    }

    // ================================================================== CellularDistance [Cellular.inl:295-529]
    // fills distance[0..TOP] with the ascending cell distances
    template<int DF, bool SIZEJ, int TOP, bool ORDERED = false, class PF, class F>
    FN2_DI void cellular_distance2( i32 seed, F x, F y, F jitter, F sizeJitter, PF minkowskiP, F ( &distance )[4] )
    {
        typedef typename pack<F>::I I;
        typedef typename pack<F>::B B;
        distance[0] = distance[1] = distance[2] = distance[3] = FN2_INF;

        I xc = iadd( cvt_i( x ), -1 ), ycBase = iadd( cvt_i( y ), -1 );
        F xcf = cvt_f( xc ), ycfBase = cvt_f( ycBase );
        xc = imul( xc, kPX );
        ycBase = imul( ycBase, kPY );

        for( int xi = 0; xi < 3; xi++ )
        {
            F xcfOffset = xcf - x;
            F ycf = ycfBase;
            I yc = ycBase;
            for( int yi = 0; yi < 3; yi++ )
            {
                I hash = hash_hb( seed, xc, yc );
                F xd, yd;
                F invMag = cell_vector( hash, jitter, xd, yd );
                xd = fma_( xd, invMag, xcfOffset );
                yd = fma_( yd, invMag, ycf - y );
                F newDistance = calc_distance<DF, false>( DF == 5 ? minkowskiP( xd, yd ) : 0.0f, xd, yd );
                if( SIZEJ ) newDistance *= size_jitter2( hash, sizeJitter );
                distance_insert<TOP, ORDERED>( distance, newDistance );
                ycf += 1.0f;
                yc = iadd( yc, kPY );
            }
            xcf += 1.0f;
            xc = iadd( xc, kPX );
        }
    }

    template<int DF, bool SIZEJ, int TOP, bool ORDERED = false, class PF, class F>
    FN2_DI void cellular_distance3( i32 seed, F x, F y, F z, F jitter, F sizeJitter, PF minkowskiP, F ( &distance )[4] )
    {
        typedef typename pack<F>::I I;
        typedef typename pack<F>::B B;
        distance[0] = distance[1] = distance[2] = distance[3] = FN2_INF;

        I xc = iadd( cvt_i( x ), -1 ), ycBase = iadd( cvt_i( y ), -1 ), zcBase = iadd( cvt_i( z ), -1 );
        F xcf = cvt_f( xc ), ycfBase = cvt_f( ycBase ), zcfBase = cvt_f( zcBase );
        xc = imul( xc, kPX );
        ycBase = imul( ycBase, kPY );
        zcBase = imul( zcBase, kPZ );

        for( int xi = 0; xi < 3; xi++ )
        {
            F xcfOffset = xcf - x;
            F ycf = ycfBase;
            I yc = ycBase;
            for( int yi = 0; yi < 3; yi++ )
            {
                F ycfOffset = ycf - y;
                F zcf = zcfBase;
                I zc = zcBase;
                for( int zi = 0; zi < 3; zi++ )
                {
                    I hash = hash_hb( seed, xc, yc, zc );
                    F xd, yd, zd;
                    F invMag = cell_vector( hash, jitter, xd, yd, zd );
                    xd = fma_( xd, invMag, xcfOffset );
                    yd = fma_( yd, invMag, ycfOffset );
                    zd = fma_( zd, invMag, zcf - z );
                    F newDistance = calc_distance<DF, false>( DF == 5 ? minkowskiP( xd, yd, zd ) : 0.0f, xd, yd, zd );
                    if( SIZEJ ) newDistance *= size_jitter3( hash, sizeJitter );
                    distance_insert<TOP, ORDERED>( distance, newDistance );
                    zcf += 1.0f;
                    zc = iadd( zc, kPZ );
                }
                ycf += 1.0f;
                yc = iadd( yc, kPY );
            }
            xcf += 1.0f;
            xc = iadd( xc, kPX );
        }
    }

    template<int DF, bool SIZEJ, int TOP, bool ORDERED = false, class PF, class F>
    FN2_DI void cellular_distance4( i32 seed, F x, F y, F z, F w, F jitter, F sizeJitter, PF minkowskiP, F ( &distance )[4] )
    {
        typedef typename pack<F>::I I;
        typedef typename pack<F>::B B;
        distance[0] = distance[1] = distance[2] = distance[3] = FN2_INF;

        I xc = iadd( cvt_i( x ), -1 ), ycBase = iadd( cvt_i( y ), -1 ), zcBase = iadd( cvt_i( z ), -1 ), wcBase = iadd( cvt_i( w ), -1 );
        F xcf = cvt_f( xc ), ycfBase = cvt_f( ycBase ), zcfBase = cvt_f( zcBase ), wcfBase = cvt_f( wcBase );
        xc = imul( xc, kPX );
        ycBase = imul( ycBase, kPY );
        zcBase = imul( zcBase, kPZ );
        wcBase = imul( wcBase, kPW );

        for( int xi = 0; xi < 3; xi++ )
        {
            F xcfOffset = xcf - x;
            F ycf = ycfBase;
            I yc = ycBase;
            for( int yi = 0; yi < 3; yi++ )
            {
                F ycfOffset = ycf - y;
                F zcf = zcfBase;
                I zc = zcBase;
                for( int zi = 0; zi < 3; zi++ )
                {
                    F zcfOffset = zcf - z;
                    F wcf = wcfBase;
                    I wc = wcBase;
                    for( int wi = 0; wi < 3; wi++ )
                    {
                        I hash = hash_hb( seed, xc, yc, zc, wc );
                        F xd, yd, zd, wd;
                        F invMag = cell_vector( hash, jitter, xd, yd, zd, wd );
                        xd = fma_( xd, invMag, xcfOffset );
                        yd = fma_( yd, invMag, ycfOffset );
                        zd = fma_( zd, invMag, zcfOffset );
                        wd = fma_( wd, invMag, wcf - w );
                        F newDistance = calc_distance<DF, false>( DF == 5 ? minkowskiP( xd, yd, zd, wd ) : 0.0f, xd, yd, zd, wd );
                        if( SIZEJ ) newDistance *= size_jitter4( hash, sizeJitter );
                        distance_insert<TOP, ORDERED>( distance, newDistance );
                        wcf += 1.0f;
                        wc = iadd( wc, kPW );
                    }
                    zcf += 1.0f;
                    zc = iadd( zc, kPZ );
                }
                ycf += 1.0f;
                yc = iadd( yc, kPY );
            }
            xcf += 1.0f;
            xc = iadd( xc, kPX );
        }
    }

    // ================================================================== CellularLookup [Cellular.inl:569-806]
    // writes the jittered centre of the closest cell
    template<int DF, bool SIZEJ, class PF, class F>
    FN2_DI void cellular_lookup2( i32 seed, F x, F y, F jitter, F sizeJitter, PF minkowskiP, F& cellX, F& cellY )
    {
        typedef typename pack<F>::I I;
        typedef typename pack<F>::B B;
        F distance = kFltMax;
        cellX = cellY = 0.0f;

        I xc = iadd( cvt_i( x ), -1 ), ycBase = iadd( cvt_i( y ), -1 );
        F xcf = cvt_f( xc ), ycfBase = cvt_f( ycBase );
        xc = imul( xc, kPX );
        ycBase = imul( ycBase, kPY );

        for( int xi = 0; xi < 3; xi++ )
        {
            F ycf = ycfBase;
            I yc = ycBase;
            for( int yi = 0; yi < 3; yi++ )
            {
                I hash = hash_hb( seed, xc, yc );
                F xd, yd;
                F invMag = cell_vector( hash, jitter, xd, yd );
                F localX = fma_( xd, invMag, xcf );
                F localY = fma_( yd, invMag, ycf );
                xd = localX - x;
                yd = localY - y;
                F newDistance = calc_distance<DF, false>( DF == 5 ? minkowskiP( xd, yd ) : 0.0f, xd, yd );
                if( SIZEJ ) newDistance *= size_jitter2( hash, sizeJitter );
                B closer = newDistance < distance;
                distance = min_( newDistance, distance );
                cellX = select( closer, localX, cellX );
                cellY = select( closer, localY, cellY );
                ycf += 1.0f;
                yc = iadd( yc, kPY );
            }
            xcf += 1.0f;
            xc = iadd( xc, kPX );
        }
    }

    template<int DF, bool SIZEJ, class PF, class F>
    FN2_DI void cellular_lookup3( i32 seed, F x, F y, F z, F jitter, F sizeJitter, PF minkowskiP, F& cellX, F& cellY, F& cellZ, typename pack<F>::I* dbg = nullptr )
    {
        typedef typename pack<F>::I I;
        typedef typename pack<F>::B B;
        F distance = kFltMax;
        cellX = cellY = cellZ = 0.0f;

        I xc = iadd( cvt_i( x ), -1 ), ycBase = iadd( cvt_i( y ), -1 ), zcBase = iadd( cvt_i( z ), -1 );
        F xcf = cvt_f( xc ), ycfBase = cvt_f( ycBase ), zcfBase = cvt_f( zcBase );
        xc = imul( xc, kPX );
        ycBase = imul( ycBase, kPY );
        zcBase = imul( zcBase, kPZ );
        if( dbg ) // integer intermediates for the parity tests (fnb200DebugIntermediates): primed base cell, then the closest cell's hash
        {
            dbg[0] = xc; dbg[1] = ycBase; dbg[2] = zcBase;
            dbg[3] = 0;
        }

        for( int xi = 0; xi < 3; xi++ )
        {
            F ycf = ycfBase;
            I yc = ycBase;
            for( int yi = 0; yi < 3; yi++ )
            {
                F zcf = zcfBase;
                I zc = zcBase;
                for( int zi = 0; zi < 3; zi++ )
                {
                    I hash = hash_hb( seed, xc, yc, zc );
                    F xd, yd, zd;
                    F invMag = cell_vector( hash, jitter, xd, yd, zd );
                    F localX = fma_( xd, invMag, xcf );
                    F localY = fma_( yd, invMag, ycf );
                    F localZ = fma_( zd, invMag, zcf );
                    xd = localX - x;
                    yd = localY - y;
                    zd = localZ - z;
                    F newDistance = calc_distance<DF, false>( DF == 5 ? minkowskiP( xd, yd, zd ) : 0.0f, xd, yd, zd );
                    if( SIZEJ ) newDistance *= size_jitter3( hash, sizeJitter );
                    B closer = newDistance < distance;
                    distance = min_( newDistance, distance );
                    cellX = select( closer, localX, cellX );
                    cellY = select( closer, localY, cellY );
                    cellZ = select( closer, localZ, cellZ );
                    if( dbg ) dbg[3] = select( closer, hash, dbg[3] );
                    zcf += 1.0f;
                    zc = iadd( zc, kPZ );
                }
                ycf += 1.0f;
                yc = iadd( yc, kPY );
            }
            xcf += 1.0f;
            xc = iadd( xc, kPX );
        }
    }

    template<int DF, bool SIZEJ, class PF, class F>
    FN2_DI void cellular_lookup4( i32 seed, F x, F y, F z, F w, F jitter, F sizeJitter, PF minkowskiP,
                                  F& cellX, F& cellY, F& cellZ, F& cellW )
    {
        typedef typename pack<F>::I I;
        typedef typename pack<F>::B B;
        F distance = kFltMax;
        cellX = cellY = cellZ = cellW = 0.0f;

        I xc = iadd( cvt_i( x ), -1 ), ycBase = iadd( cvt_i( y ), -1 ), zcBase = iadd( cvt_i( z ), -1 ), wcBase = iadd( cvt_i( w ), -1 );
        F xcf = cvt_f( xc ), ycfBase = cvt_f( ycBase ), zcfBase = cvt_f( zcBase ), wcfBase = cvt_f( wcBase );
        xc = imul( xc, kPX );
        ycBase = imul( ycBase, kPY );
        zcBase = imul( zcBase, kPZ );
        wcBase = imul( wcBase, kPW );

        for( int xi = 0; xi < 3; xi++ )
        {
            F ycf = ycfBase;
            I yc = ycBase;
            for( int yi = 0; yi < 3; yi++ )
            {
                F zcf = zcfBase;
                I zc = zcBase;
                for( int zi = 0; zi < 3; zi++ )
                {
                    F wcf = wcfBase;
                    I wc = wcBase;
                    for( int wi = 0; wi < 3; wi++ )
                    {
                        I hash = hash_hb( seed, xc, yc, zc, wc );
                        F xd, yd, zd, wd;
                        F invMag = cell_vector( hash, jitter, xd, yd, zd, wd );
                        F localX = fma_( xd, invMag, xcf );
                        F localY = fma_( yd, invMag, ycf );
                        F localZ = fma_( zd, invMag, zcf );
                        F localW = fma_( wd, invMag, wcf );
                        xd = localX - x;
                        yd = localY - y;
                        zd = localZ - z;
                        wd = localW - w;
                        F newDistance = calc_distance<DF, false>( DF == 5 ? minkowskiP( xd, yd, zd, wd ) : 0.0f, xd, yd, zd, wd );
                        if( SIZEJ ) newDistance *= size_jitter4( hash, sizeJitter );
                        B closer = newDistance < distance;
                        distance = min_( newDistance, distance );
                        cellX = select( closer, localX, cellX );
                        cellY = select( closer, localY, cellY );
                        cellZ = select( closer, localZ, cellZ );
                        cellW = select( closer, localW, cellW );
                        wcf += 1.0f;
                        wc = iadd( wc, kPW );
                    }
                    zcf += 1.0f;
                    zc = iadd( zc, kPZ );
                }
                ycf += 1.0f;
                yc = iadd( yc, kPY );
            }
            xcf += 1.0f;
            xc = iadd( xc, kPX );
        }
    }
}
