// fastnoise2_b200 -- device node library: SuperSimplex 2D/3D/4D (un-normalised values).
// Restates include/FastNoise/Generators/Simplex.inl:293-936 as scalar per-thread code; the vertex
// contributions are accumulated in the reference's order so the fma chain rounds identically.
#pragma once

namespace fn2
{
    constexpr float kSuperSimplexBound2 = 0.14084912836551666259765625f;
    constexpr float kSuperSimplexBound3 = 0.0069091119803488254547119140625f;
    constexpr float kSuperSimplexBound4 = (float)0.011310013942420482635498046875;

    FN2_DI float sq2( float v ) { return ( v * v ) * ( v * v ); }
    FN2_DI fpair sq2( fpair v ) { return ( v * v ) * ( v * v ); }
    // base +- prime, by sign mask [Simplex.inl:475 and siblings]
    FN2_DI i32 step_prime( i32 base, i32 prime, bool negative ) { return iadd( base, negative ? isub( 0, prime ) : prime ); }
    FN2_DI ipair step_prime( ipair base, i32 prime, bpair negative ) { return ipair( step_prime( base.x, prime, negative.x ), step_prime( base.y, prime, negative.y ) ); }

    // Vertex producers shared with DomainWarpSuperSimplex [DomainWarpSimplex.inl:1133-1833 repeats the geometry below],
    // vertices in the reference's accumulation order, falloff already raised to the 4th power.
    // [Simplex.inl:293-377]
    template<int SCALE, class F>
    FN2_DI void supersimplex2_vertices( i32 seed, F x, F y, Verts2<F> ( &v )[4] )
    {
        typedef typename pack<F>::I I;
        typedef typename pack<F>::B B;
        constexpr double kSkew2 = 1.0 / ( kRoot3 + 1.0 );
        constexpr double kUnskew2 = -1.0 / ( kRoot3 + 3.0 );
        constexpr float kR2 = (float)( 2.0 / 3.0 );

        F skewDelta = (float)kSkew2 * ( x + y );
        F xSkewed = x + skewDelta, ySkewed = y + skewDelta;
        F xBase = floor_( xSkewed ), yBase = floor_( ySkewed );
        F dxS = xSkewed - xBase, dyS = ySkewed - yBase;
        I xP = imul( cvt_i( xBase ), kPX ), yP = imul( cvt_i( yBase ), kPY );

        B forwardXY = dxS + dyS > 1.0f;
        F boundaryXY = select( forwardXY, F( -1.0f ), F( 0.0f ) );
        B forwardX = fma_( dxS, -2.0f, dyS ) < boundaryXY;
        B forwardY = fma_( dyS, -2.0f, dxS ) < boundaryXY;

        F unskewDelta = (float)kUnskew2 * ( dxS + dyS );
        F dxBase = dxS + unskewDelta, dyBase = dyS + unskewDelta;

        F falloffBase0 = fnma_( dxBase, dxBase, fnma_( dyBase, dyBase, kR2 ) );
        v[0] = { hash_hbs<SCALE>( seed, xP, yP ), dxBase, dyBase, pow4( falloffBase0 ) }; // vertex <0,0> (no clamp, as there)
        { // vertex <1,1>
            F falloff = fma_( unskewDelta, (float)( -4.0 * ( kRoot3 + 2.0 ) / ( kRoot3 + 3.0 ) ), falloffBase0 - kR2 );
            v[1] = { hash_hbs<SCALE>( seed, iadd( xP, kPX ), iadd( yP, kPY ) ), dxBase - (float)( 2 * kUnskew2 + 1 ), dyBase - (float)( 2 * kUnskew2 + 1 ), pow4( falloff ) };
        }

        F xyDelta = select( forwardXY, F( (float)( kUnskew2 + 1 ) ), F( (float)( -kUnskew2 ) ) );
        dxBase -= xyDelta;
        dyBase -= xyDelta;

        { // vertex <1,0> or <-1,0> or <1,2>
            I xv = select( forwardX, iadd( xP, imul( kPX, 2 ) ), xP );
            xv = select( forwardXY, xv, isub( xv, kPX ) );
            I yv = select( forwardXY, iadd( yP, kPY ), yP );
            F dx = dxBase - ( select( forwardX, F( (float)( 1 + 2 * kUnskew2 ) ), F( -1.0f ) ));
            F dy = select( forwardX, dyBase - (float)( 2 * kUnskew2 ), dyBase );
            F falloff = max_( fnma_( dx, dx, fnma_( dy, dy, kR2 ) ), 0.0f );
            v[2] = { hash_hbs<SCALE>( seed, xv, yv ), dx, dy, pow4( falloff ) };
        }
        { // vertex <0,1> or <0,-1> or <2,1>
            I xv = select( forwardXY, iadd( xP, kPX ), xP );
            I yv = select( forwardY, iadd( yP, imul( kPY, 2 ) ), yP );
            yv = select( forwardXY, yv, isub( yv, kPY ) );
            F dx = select( forwardY, dxBase - (float)( 2 * kUnskew2 ), dxBase );
            F dy = dyBase - ( select( forwardY, F( (float)( 1 + 2 * kUnskew2 ) ), F( -1.0f ) ));
            F falloff = max_( fnma_( dx, dx, fnma_( dy, dy, kR2 ) ), 0.0f );
            v[3] = { hash_hbs<SCALE>( seed, xv, yv ), dx, dy, pow4( falloff ) };
        }
    }

    template<class F>
    FN2_DI F supersimplex2( i32 seed, F x, F y )
    {
        typedef typename pack<F>::I I;
        typedef typename pack<F>::B B;
        Verts2<F> v[4];
        supersimplex2_vertices<kGradHashScale>( seed, x, y, v );
        F value = v[0].falloff * grad_dot2( v[0].hb, v[0].dx, v[0].dy );
#pragma unroll
        for( int k = 1; k < 4; k++ ) value = fma_( v[k].falloff, grad_dot2( v[k].hb, v[k].dx, v[k].dy ), value );
        return value;
    }

    // [Simplex.inl:379-576]
    template<int SCALE, class F>
    FN2_DI void supersimplex3_vertices( i32 seed, F x, F y, F z, Verts3<F> ( &v )[8] )
    {
        typedef typename pack<F>::I I;
        typedef typename pack<F>::B B;
        constexpr double kSkew3 = 1.0 / 3.0;
        constexpr double kReflectUnskew3 = -1.0 / 2.0;
        constexpr double kTwiceUnskew3 = -1.0 / 4.0;
        constexpr double kDistanceSquaredA = 3.0 / 4.0;
        constexpr double kDistanceSquaredB = 1.0;
        constexpr float kR2 = (float)kDistanceSquaredA;

        F skewDelta = (float)kSkew3 * ( x + y + z );
        F xSkewed = x + skewDelta, ySkewed = y + skewDelta, zSkewed = z + skewDelta;
        F xBase = floor_( xSkewed ), yBase = floor_( ySkewed ), zBase = floor_( zSkewed );
        F dxS = xSkewed - xBase, dyS = ySkewed - yBase, dzS = zSkewed - zBase;

        { // closest lattice vertex of the unit cell [:405-441]
            F tud = (float)kTwiceUnskew3 * ( dxS + dyS + dzS );
            F xN = dxS + tud, yN = dyS + tud, zN = dzS + tud;
            F xyzN = -tud;

            F maxScore = 0.375f;
            I bits = select( xyzN > maxScore, I( -1 ), I( 0 ) );
            maxScore = max_( maxScore, xyzN );
#define FN2_CONSIDER( SCORE, BITS ) { F sc = SCORE; bits = select( sc > maxScore, BITS, bits ); maxScore = max_( maxScore, sc ); }
            FN2_CONSIDER( xN, 1 )
            FN2_CONSIDER( yN, 2 )
            FN2_CONSIDER( zN, 4 )
            maxScore += 0.125f - xyzN;
            FN2_CONSIDER( -zN, 3 )
            FN2_CONSIDER( -yN, 5 )
            FN2_CONSIDER( -xN, 6 )

            B moveX = ( bits & 1 ) != 0, moveY = ( bits & 2 ) != 0, moveZ = ( bits & 4 ) != 0;
            xBase = select( moveX, xBase + 1.0f, xBase );
            yBase = select( moveY, yBase + 1.0f, yBase );
            zBase = select( moveZ, zBase + 1.0f, zBase );
            dxS = select( moveX, dxS - 1.0f, dxS );
            dyS = select( moveY, dyS - 1.0f, dyS );
            dzS = select( moveZ, dzS - 1.0f, dzS );
        }

        I xP = imul( cvt_i( xBase ), kPX ), yP = imul( cvt_i( yBase ), kPY ), zP = imul( cvt_i( zBase ), kPZ );

        F sum = dxS + dyS + dzS;
        F tud = (float)kTwiceUnskew3 * sum;
        F xN = dxS + tud, yN = dyS + tud, zN = dzS + tud;
        F xyzN = -tud;

        F ud = (float)kReflectUnskew3 * sum;
        F dxB = dxS + ud, dyB = dyS + ud, dzB = dzS + ud;
        F coordSum = (float)( 1 + 3 * kReflectUnskew3 ) * sum;

        F stemA, stemB;
        { // <0,0,0>
            F fb = fnma_( dzB, dzB, fnma_( dyB, dyB, fnma_( dxB, dxB, kR2 ) ) ) * 0.5f;
            stemA = fb - (float)( kDistanceSquaredA * 0.5 );
            stemB = fb - (float)( kDistanceSquaredB * 0.5 );
            v[0] = { hash_hbs<SCALE>( seed, xP, yP, zP ), dxB, dyB, dzB, sq2( fb ) };
        }
        // The reference flips signs with an XOR mask (`c ^ sign`, `d ^ sign`). A flipped constant or coordinate is the value times
        // s = -1 or +1, and that product is exact, so every `a - ( c ^ sign )` is ONE fused operation fnma_( c, s, a ) with the
        // same single rounding, and `fma( c ^ sign, coordSum, stem )` is fma( c, coordSum * s, stem ): the sign work moves from
        // the half-rate ALU pipe (SEL + LOP3 per use) to the FMA pipe and the XORs disappear. Stepped lattice coordinates
        // (base +- prime by the same mask) are selected from the two neighbours computed once per axis.
        const I xPp = iadd( xP, kPX ), yPp = iadd( yP, kPY ), zPp = iadd( zP, kPZ );
        const I xPm = isub( xP, kPX ), yPm = isub( yP, kPY ), zPm = isub( zP, kPZ );
        { // <1,1,1> or <-1,-1,-1>
            B sm = xyzN < 0.0f;
            F s = select( sm, F( -1.0f ), F( 1.0f ) );
            constexpr float o = (float)( 3 * kReflectUnskew3 + 1 );
            F fb = max_( fma_( F( o ), coordSum * s, stemA ), 0.0f );
            v[1] = { hash_hbs<SCALE>( seed, select( sm, xPm, xPp ), select( sm, yPm, yPp ), select( sm, zPm, zPp ) ), fnma_( F( o ), s, dxB ), fnma_( F( o ), s, dyB ), fnma_( F( o ), s, dzB ), sq2( fb ) };
        }
        constexpr float o2 = (float)( 2 * kReflectUnskew3 );
        { // <1,1,0> or <-1,-1,0>
            B sm = xyzN < zN;
            F s = select( sm, F( -1.0f ), F( 1.0f ) );
            F fb = min_( fms_( dzB, s, stemB ), 0.0f );
            v[2] = { hash_hbs<SCALE>( seed, select( sm, xPm, xPp ), select( sm, yPm, yPp ), zP ), dxB, dyB, fnma_( F( o2 ), s, dzB ), sq2( fb ) };
        }
        { // <1,0,1> or <-1,0,-1>
            B sm = xyzN < yN;
            F s = select( sm, F( -1.0f ), F( 1.0f ) );
            F fb = min_( fms_( dyB, s, stemB ), 0.0f );
            v[3] = { hash_hbs<SCALE>( seed, select( sm, xPm, xPp ), yP, select( sm, zPm, zPp ) ), dxB, fnma_( F( o2 ), s, dyB ), dzB, sq2( fb ) };
        }
        { // <0,1,1> or <0,-1,-1>
            B sm = xyzN < xN;
            F s = select( sm, F( -1.0f ), F( 1.0f ) );
            F fb = min_( fms_( dxB, s, stemB ), 0.0f );
            v[4] = { hash_hbs<SCALE>( seed, xP, select( sm, yPm, yPp ), select( sm, zPm, zPp ) ), fnma_( F( o2 ), s, dxB ), dyB, dzB, sq2( fb ) };
        }
        constexpr float o1 = (float)kReflectUnskew3;
        { // <1,0,0> or <-1,0,0>
            B sm = xN < 0.0f;
            F s = select( sm, F( -1.0f ), F( 1.0f ) );
            F fb = max_( fma_( dxB, s, fma_( F( o1 ), coordSum * s, stemA ) ), 0.0f );
            v[5] = { hash_hbs<SCALE>( seed, select( sm, xPm, xPp ), yP, zP ), fma_( F( o1 ), s, dxB ), fnma_( F( o1 ), s, dyB ), fnma_( F( o1 ), s, dzB ), sq2( fb ) };
        }
        { // <0,1,0> or <0,-1,0>
            B sm = yN < 0.0f;
            F s = select( sm, F( -1.0f ), F( 1.0f ) );
            F fb = max_( fma_( dyB, s, fma_( F( o1 ), coordSum * s, stemA ) ), 0.0f );
            v[6] = { hash_hbs<SCALE>( seed, xP, select( sm, yPm, yPp ), zP ), fnma_( F( o1 ), s, dxB ), fma_( F( o1 ), s, dyB ), fnma_( F( o1 ), s, dzB ), sq2( fb ) };
        }
        { // <0,0,1> or <0,0,-1>
            B sm = zN < 0.0f;
            F s = select( sm, F( -1.0f ), F( 1.0f ) );
            F fb = max_( fma_( dzB, s, fma_( F( o1 ), coordSum * s, stemA ) ), 0.0f );
            v[7] = { hash_hbs<SCALE>( seed, xP, yP, select( sm, zPm, zPp ) ), fnma_( F( o1 ), s, dxB ), fnma_( F( o1 ), s, dyB ), fma_( F( o1 ), s, dzB ), sq2( fb ) };
        }
    }

    template<class F>
    FN2_DI F supersimplex3( i32 seed, F x, F y, F z )
    {
        typedef typename pack<F>::I I;
        typedef typename pack<F>::B B;
        Verts3<F> v[8];
        supersimplex3_vertices<kGradHashScale>( seed, x, y, z, v );
        F value = v[0].falloff * grad_dot3( v[0].hb, v[0].dx, v[0].dy, v[0].dz );
#pragma unroll
        for( int k = 1; k < 8; k++ ) value = fma_( v[k].falloff, grad_dot3( v[k].hb, v[k].dx, v[k].dy, v[k].dz ), value );
        return value;
    }

    // [Simplex.inl:578-936]
    template<int SCALE, class F>
    FN2_DI void supersimplex4_vertices( i32 seed, F x, F y, F z, F w, Verts4<F> ( &v )[16] )
    {
        typedef typename pack<F>::I I;
        typedef typename pack<F>::B B;
        constexpr double kSkew4 = 1.0 / ( kRoot5 + 1.0 );
        constexpr double kUnskew4 = -1.0 / ( kRoot5 + 5.0 );
        constexpr double kTwiceUnskew4 = -1.0 / 5.0;
        constexpr double kDistanceSquaredA = 4.0 / 5.0;
        constexpr double kDistanceSquaredB = 6.0 / 5.0;
        constexpr float kR2 = (float)kDistanceSquaredA;

        F skewDelta = (float)kSkew4 * ( x + y + z + w );
        F xSkewed = x + skewDelta, ySkewed = y + skewDelta, zSkewed = z + skewDelta, wSkewed = w + skewDelta;
        F xBase = floor_( xSkewed ), yBase = floor_( ySkewed ), zBase = floor_( zSkewed ), wBase = floor_( wSkewed );
        F dxS = xSkewed - xBase, dyS = ySkewed - yBase, dzS = zSkewed - zBase, dwS = wSkewed - wBase;

        { // closest lattice vertex of the unit cell [:608-659]
            F tud = (float)kTwiceUnskew4 * ( dxS + dyS + dzS + dwS );
            F xN = dxS + tud, yN = dyS + tud, zN = dzS + tud, wN = dwS + tud;
            F xyzwN = -tud;

            F maxScore = 0.6f - xyzwN;
            I bits = select( 0.2f > maxScore, I( -1 ), I( 0 ) );
            maxScore = max_( maxScore, 0.2f );
            FN2_CONSIDER( -wN, 7 )
            FN2_CONSIDER( -zN, 11 )
            FN2_CONSIDER( -yN, 13 )
            FN2_CONSIDER( -xN, 14 )
            maxScore += xyzwN - 0.2f;
            FN2_CONSIDER( xN, 1 )
            FN2_CONSIDER( yN, 2 )
            FN2_CONSIDER( zN, 4 )
            FN2_CONSIDER( wN, 8 )
            maxScore += 0.2f - xN;
            FN2_CONSIDER( yN, 3 )
            FN2_CONSIDER( zN, 5 )
            FN2_CONSIDER( wN, 9 )
            maxScore += xN;
            FN2_CONSIDER( yN + zN, 6 )
            maxScore -= wN;
            FN2_CONSIDER( yN, 10 )
            FN2_CONSIDER( zN, 12 )
#undef FN2_CONSIDER

            B moveX = ( bits & 1 ) != 0, moveY = ( bits & 2 ) != 0, moveZ = ( bits & 4 ) != 0, moveW = ( bits & 8 ) != 0;
            xBase = select( moveX, xBase + 1.0f, xBase );
            yBase = select( moveY, yBase + 1.0f, yBase );
            zBase = select( moveZ, zBase + 1.0f, zBase );
            wBase = select( moveW, wBase + 1.0f, wBase );
            dxS = select( moveX, dxS - 1.0f, dxS );
            dyS = select( moveY, dyS - 1.0f, dyS );
            dzS = select( moveZ, dzS - 1.0f, dzS );
            dwS = select( moveW, dwS - 1.0f, dwS );
        }

        I xP = imul( cvt_i( xBase ), kPX ), yP = imul( cvt_i( yBase ), kPY ), zP = imul( cvt_i( zBase ), kPZ ), wP = imul( cvt_i( wBase ), kPW );

        F sum = dxS + dyS + dzS + dwS;
        F tud = (float)kTwiceUnskew4 * sum;
        F xN = dxS + tud, yN = dyS + tud, zN = dzS + tud, wN = dwS + tud;
        F xyzwN = -tud;

        F ud = (float)kUnskew4 * sum;
        F dxB = dxS + ud, dyB = dyS + ud, dzB = dzS + ud, dwB = dwS + ud;
        F coordSum = (float)( 1 + 4 * kUnskew4 ) * sum;

        F stemA, stemB;
        { // <0,0,0,0>
            F fb = fnma_( dwB, dwB, fnma_( dzB, dzB, fnma_( dyB, dyB, fnma_( dxB, dxB, kR2 ) ) ) ) * 0.5f;
            stemA = fb - (float)( kDistanceSquaredA * 0.5 );
            stemB = fb - (float)( kDistanceSquaredB * 0.5 );
            v[0] = { hash_hbs<SCALE>( seed, xP, yP, zP, wP ), dxB, dyB, dzB, dwB, sq2( fb ) };
        }

        // One lattice vertex: MX..MW say which axes step (by +-1 according to the sign mask), K is how many. Coordinates on stepping
        // axes use offset1 = (K*u + 1)^sign, the others offset0 = (K*u)^sign [Simplex.inl:692-931]. As in 3D above the sign mask is
        // applied as the exact factor s = -1 / +1 inside fused operations: OFF picks the offset that multiplies coordSum, STEM the
        // falloff stem, FALLOFF finishes the vertex-class specific falloff base from `base` and `s`.
        const I xPp = iadd( xP, kPX ), yPp = iadd( yP, kPY ), zPp = iadd( zP, kPZ ), wPp = iadd( wP, kPW );
        const I xPm = isub( xP, kPX ), yPm = isub( yP, kPY ), zPm = isub( zP, kPZ ), wPm = isub( wP, kPW );
#define FN2_SS4_VERTEX( SLOT, COND, MX, MY, MZ, MW, K, OFF, STEM, FALLOFF ) \
        { \
            B sm = COND; \
            F s = select( sm, F( -1.0f ), F( 1.0f ) ); \
            F base = fma_( F( (float)( K * kUnskew4 + OFF ) ), coordSum * s, STEM ); \
            F fb = max_( FALLOFF, 0.0f ); \
            v[SLOT] = { hash_hbs<SCALE>( seed, MX ? select( sm, xPm, xPp ) : xP, MY ? select( sm, yPm, yPp ) : yP, MZ ? select( sm, zPm, zPp ) : zP, MW ? select( sm, wPm, wPp ) : wP ), \
                        fnma_( F( (float)( K * kUnskew4 + MX ) ), s, dxB ), fnma_( F( (float)( K * kUnskew4 + MY ) ), s, dyB ), \
                        fnma_( F( (float)( K * kUnskew4 + MZ ) ), s, dzB ), fnma_( F( (float)( K * kUnskew4 + MW ) ), s, dwB ), sq2( fb ) }; \
        }

        FN2_SS4_VERTEX( 1, xyzwN < 0.0f, 1, 1, 1, 1, 4, 1, stemA, base )
        FN2_SS4_VERTEX( 2, xyzwN < wN, 1, 1, 1, 0, 3, 1, stemB, fnma_( dwB, s, base ) )
        FN2_SS4_VERTEX( 3, xyzwN < zN, 1, 1, 0, 1, 3, 1, stemB, fnma_( dzB, s, base ) )
        FN2_SS4_VERTEX( 4, xyzwN < yN, 1, 0, 1, 1, 3, 1, stemB, fnma_( dyB, s, base ) )
        FN2_SS4_VERTEX( 5, xyzwN < xN, 0, 1, 1, 1, 3, 1, stemB, fnma_( dxB, s, base ) )
        FN2_SS4_VERTEX( 6, xN < 0.0f, 1, 0, 0, 0, 1, 0, stemA, fma_( dxB, s, base ) )
        FN2_SS4_VERTEX( 7, xN < -yN, 1, 1, 0, 0, 2, 0, stemB, fma_( dxB + dyB, s, base ) )
        FN2_SS4_VERTEX( 8, xN < -zN, 1, 0, 1, 0, 2, 0, stemB, fma_( dxB + dzB, s, base ) )
        FN2_SS4_VERTEX( 9, xN < -wN, 1, 0, 0, 1, 2, 0, stemB, fma_( dxB + dwB, s, base ) )
        FN2_SS4_VERTEX( 10, yN < 0.0f, 0, 1, 0, 0, 1, 0, stemA, fma_( dyB, s, base ) )
        FN2_SS4_VERTEX( 11, yN < -zN, 0, 1, 1, 0, 2, 0, stemB, fma_( dyB + dzB, s, base ) )
        FN2_SS4_VERTEX( 12, yN < -wN, 0, 1, 0, 1, 2, 0, stemB, fma_( dyB + dwB, s, base ) )
        FN2_SS4_VERTEX( 13, zN < 0.0f, 0, 0, 1, 0, 1, 0, stemA, fma_( dzB, s, base ) )
        FN2_SS4_VERTEX( 14, zN < -wN, 0, 0, 1, 1, 2, 0, stemB, fma_( dzB + dwB, s, base ) )
        FN2_SS4_VERTEX( 15, wN < 0.0f, 0, 0, 0, 1, 1, 0, stemA, fma_( dwB, s, base ) )
#undef FN2_SS4_VERTEX
    }

    template<class F>
    FN2_DI F supersimplex4( i32 seed, F x, F y, F z, F w )
    {
        typedef typename pack<F>::I I;
        typedef typename pack<F>::B B;
        Verts4<F> v[16];
        supersimplex4_vertices<kGradHashScale>( seed, x, y, z, w, v );
        F value = v[0].falloff * grad_dot4( v[0].hb, v[0].dx, v[0].dy, v[0].dz, v[0].dw );
#pragma unroll
        for( int k = 1; k < 16; k++ ) value = fma_( v[k].falloff, grad_dot4( v[k].hb, v[k].dx, v[k].dy, v[k].dz, v[k].dw ), value );
        return value;
    }
}
