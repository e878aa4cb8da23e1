// fastnoise2_b200 -- device node library: DomainWarpSimplex / DomainWarpSuperSimplex displacement fields.
// Restates include/FastNoise/Generators/DomainWarpSimplex.inl (both flavours, see FN2_TABLES in fn2_device.cuh): the lattice
// geometry is the Simplex / SuperSimplex vertex list (fn2_device.cuh, fn2_device_supersimplex.cuh); each vertex
// contributes a VECTOR built from its hash by one of two schemes
//   SCHEME 0  OrthogonalGradientMatrix  [:199-686]
//   SCHEME 1  GradientOuterProduct      [:7-196]
// `seed` is the seed the reference's Warp_ND receives (DomainWarpSimplex::Warp adds the node's seed offset first,
// DomainWarpSuperSimplex::Warp does not -- the caller in lower.cpp reproduces that). (x,..) is position x frequency,
// (xOut,..) the position being displaced, boundAx = kBounding * axisScale folded on the host in float.
#pragma once

namespace fn2
{
    constexpr float kSkew2f = (float)( 1.0 / ( kRoot3 + 1.0 ) );
    constexpr float kSkew4f = (float)( 1.0 / ( kRoot5 + 1.0 ) );
    constexpr float kRoot2f = (float)kRoot2;
    constexpr float kRoot3f = (float)kRoot3;

    // ------------------------------------------------------------------ GradientOuterProduct
    // 2D [:7-60]
    FN2_DI void vec_gop2( i32 hash31, float fX, float fY, float multiplier, float& valueX, float& valueY )
    {
        i32 hashShifted = lsr( hash31, 1 );
        i32 indexGradient = imul( hashShifted, 12 >> 2 );                                  // [0,12) in the upper four bits
        i32 indexOuter = imul( hashShifted, (i32)0xAAAAAAABu ) & (i32)0xC0000003u;          // [0,12) in bits 0,1,30,31
#if FN2_TABLES
        // [:13-25]
        const float2 g = table_pair( kTabGop2, ( ( indexGradient >> 28 ) & 15 ) * 8 );
        const float2 o = table_pair( kTabGop2, ( ( indexOuter | ( indexOuter >> 28 ) ) & 15 ) * 8 );
        multiplier *= fma_( fY, g.y, fX * g.x );
        valueX = fma_( multiplier, o.x, valueX );
        valueY = fma_( multiplier, o.y, valueY );
        return;
#endif
        {
            float u = sel_hi( shl( indexGradient, 2 ), fY, fX );
            float v = sel_hi( shl( indexGradient, 2 ), fX, fY );
            float a = u * sel_hi( indexGradient, 2.0f, kRoot3f );
            float b = fxor( v, shl( indexGradient >> 30, 31 ) );
            float dot = indexGradient >= 0 ? a + b : a;
            multiplier *= fxor( dot, shl( ( indexGradient >> 28 ) ^ indexOuter, 31 ) );
        }
        {
            float a = multiplier * sel_hi( indexOuter, 2.0f, kRoot3f );
            float b = fxor( indexOuter >= 0 ? multiplier : 0.0f, shl( indexOuter >> 30, 31 ) );
            valueX += sel_hi( shl( indexOuter, 30 ), b, a );
            valueY += sel_hi( shl( indexOuter, 30 ), a, b );
        }
    }

    // 3D [:62-119]
    FN2_DI void vec_gop3( i32 hash31, float fX, float fY, float fZ, float multiplier, float& valueX, float& valueY, float& valueZ )
    {
        i32 hashShifted = lsr( hash31, 1 );
        i32 indexGradient = lsr( imul( hashShifted, 12 >> 2 ), 28 );                        // [0,12)
        i32 indexOuter = imul( hashShifted, (i32)0xAAAAAAABu ) & (i32)0xC0000003u;
        indexOuter |= indexOuter >> 28;
#if FN2_TABLES
        // [:70-81] the 12 directions are the first 12 entries of the common 3D gradient table
        const int slotG = indexGradient * 8, slotO = ( indexOuter & 15 ) * 8;
        const float2 g = table_pair( kTabCommon3XY, slotG );
        const float2 o = table_pair( kTabCommon3XY, slotO );
        multiplier *= fma_( table_first( kTabCommon3Z, slotG ), fZ, fma_( fY, g.y, fX * g.x ) );
        valueX = fma_( multiplier, o.x, valueX );
        valueY = fma_( multiplier, o.y, valueY );
        valueZ = fma_( multiplier, table_first( kTabCommon3Z, slotO ), valueZ );
        return;
#endif
        {
            float u = indexGradient >= 8 ? fY : fX;
            float v = indexGradient >= 4 ? fZ : fY;
            multiplier *= fxor( u, shl( indexGradient, 31 ) ) + fxor( v, shl( indexGradient >> 1, 31 ) );
        }
        {
            indexOuter &= 0xF;
            float signed0 = fxor( multiplier, shl( indexOuter, 31 ) );
            float signed1 = fxor( multiplier, shl( indexOuter >> 1, 31 ) );
            bool notYZ = indexOuter < 8;
            bool notXY = indexOuter >= 4;
            valueX = notYZ ? valueX + signed0 : valueX;
            valueZ = notXY ? valueZ + signed1 : valueZ;
            valueY = ( notYZ && notXY ) ? valueY : valueY + ( notXY ? signed0 : signed1 );
        }
    }

    // one of the 20 4D directions from an index held in the upper five bits [:149-166]
    FN2_DI void gop4_direction( i32 index, float& gX, float& gY, float& gZ, float& gW )
    {
        i32 indexA = index & ( 0x03 << 27 );
        i32 indexB = ( index >> 2 ) & ( 0x07 << 27 );
        indexB ^= indexA;
        bool extra = indexB >= ( 0x04 << 27 );
        bool equal = indexA == indexB;
        if( equal ) indexA = -1; // forces the decrement conditions to fail
        float neutral = ( equal || extra ) ? ( extra ? kSkew4f * -1.0f : kSkew4f ) : 0.0f;
        // MaskedIncrement( indexB == k, MaskedDecrement( indexA == k, neutral ) ) per axis
        float dX = indexA == ( 0 << 27 ) ? neutral - 1.0f : neutral;
        float dY = indexA == ( 1 << 27 ) ? neutral - 1.0f : neutral;
        float dZ = indexA == ( 2 << 27 ) ? neutral - 1.0f : neutral;
        float dW = indexA == ( 3 << 27 ) ? neutral - 1.0f : neutral;
        gX = indexB == ( 0 << 27 ) ? dX + 1.0f : dX;
        gY = indexB == ( 1 << 27 ) ? dY + 1.0f : dY;
        gZ = indexB == ( 2 << 27 ) ? dZ + 1.0f : dZ;
        gW = indexB == ( 3 << 27 ) ? dW + 1.0f : dW;
    }

    // 4D [:121-196]
    FN2_DI void vec_gop4( i32 hash, float fX, float fY, float fZ, float fW, float multiplier, float& valueX, float& valueY, float& valueZ, float& valueW )
    {
        i32 hashShifted = lsr( hash, 2 );
        i32 indexGradient = imul( hashShifted, 20 >> 2 );                                   // [0,20) in the upper five bits
        i32 indexOuter = imul( hashShifted, (i32)0xCCCCCCCDu );
        indexOuter = imul( indexOuter & (i32)0xE0000003u, 3 | ( 1 << 27 ) );                // [0,20) in the upper five bits, independent of the above
        float gX, gY, gZ, gW;
        gop4_direction( indexGradient, gX, gY, gZ, gW );
        multiplier *= fma_( gW, fW, fma_( gZ, fZ, fma_( gY, fY, gX * fX ) ) );
        gop4_direction( indexOuter, gX, gY, gZ, gW );
        valueX = fma_( multiplier, gX, valueX );
        valueY = fma_( multiplier, gY, valueY );
        valueZ = fma_( multiplier, gZ, valueZ );
        valueW = fma_( multiplier, gW, valueW );
    }

    // ------------------------------------------------------------------ OrthogonalGradientMatrix
    // 2D [:199-243]
    FN2_DI void vec_ogm2( i32 hash31, float fX, float fY, float multiplier, float& valueX, float& valueY )
    {
        i32 index = imul( lsr( hash31, 1 ), 12 >> 2 );                                      // [0,12) in the upper four bits
        i32 ofThree = lsr( index, 30 );
        float masked = index >= 0 ? kSkew2f : 0.0f;
        float gX = fxor( ofThree != 0 ? masked + 1.0f : masked, shl( index >> 28, 31 ) );
        float gY = fxor( ofThree != 1 ? masked + 1.0f : masked, shl( index >> 29, 31 ) );
        valueX = fma_( multiplier, fma_( fY, gY, fX * gX ), valueX );
        multiplier = fxor( multiplier, shl( hash31, 31 ) );
        valueY = fma_( multiplier, fms_( fY, gX, fX * gY ), valueY );
    }

    // 3D [:245-322]
    FN2_DI void vec_ogm3( i32 hash31, float fX, float fY, float fZ, float multiplier, float& valueX, float& valueY, float& valueZ )
    {
        constexpr float kA = 2.224744871391589f, kB = -0.224744871391589f, kC = -1.0f, kDE = 1.0f, kF = 2.0f;

        i32 hashShifted = lsr( hash31, 1 );
        i32 facet = imul( hashShifted, (i32)0xAAAAAAABu );   // [0,3) in the highest two bits, [0,8) in the lowest three
        i32 permute2 = shl( facet, 29 );                     // one bit, in the sign position
        i32 permute3 = lsr( imul( hashShifted, 3 ), 30 );    // [0,3)
        i32 finalSign = shl( hash31, 31 );

        i32 sign0 = shl( facet, 31 );
        i32 sign1 = shl( facet, 30 ) & (i32)0x80000000u;
        i32 notYZ = facet, notXY = shl( facet, 1 );

        float aX = sel_hi( notYZ, kC, fxor( kA, sign0 ) );
        float aY = sel_hi( notYZ | notXY, sel_hi( notXY, fxor( kB, sign1 ), fxor( kA, sign0 ) ), kC );
        float aZ = sel_hi( notXY, kC, fxor( kB, sign1 ) );
        float valueA = fma_( aZ, fZ, fma_( fY, aY, fX * aX ) );

        float bX = sel_hi( notYZ, kC, fxor( kB, sign0 ) );
        float bY = sel_hi( notYZ | notXY, sel_hi( notXY, fxor( kA, sign1 ), fxor( kB, sign0 ) ), kC );
        float bZ = sel_hi( notXY, kC, fxor( kA, sign1 ) );
        float valueB = fma_( bZ, fZ, fma_( fY, bY, fX * bX ) );

        float cX = sel_hi( notYZ, kF, fxor( kDE, sign0 ) );
        float cY = sel_hi( notYZ | notXY, sel_hi( notXY, fxor( kDE, sign1 ), fxor( kDE, sign0 ) ), kF );
        float cZ = sel_hi( notXY, kF, fxor( kDE, sign1 ) );
        float valueC = fma_( cZ, fZ, fma_( fY, cY, fX * cX ) );

        float valueAB = sel_hi( permute2, valueB, valueA );
        float valueBA = sel_hi( permute2, valueA, valueB );

        multiplier = fxor( multiplier, finalSign );
        valueX = fma_( multiplier, permute3 == 0 ? valueC : valueAB, valueX );
        valueY = fma_( multiplier, permute3 == 1 ? valueC : ( permute3 == 2 ? valueBA : valueAB ), valueY );
        valueZ = fma_( multiplier, permute3 == 2 ? valueC : valueBA, valueZ );
    }

    // 4D [:324-686]
    FN2_DI void vec_ogm4( i32 hash31, float fX, float fY, float fZ, float fW, float multiplier, float& valueX, float& valueY, float& valueZ, float& valueW )
    {
        constexpr float kIdxNegAB = -0.375999676691291f;  // kComponentPairwiseIndexedNegativeAB
        constexpr float kFillAB = 0.222726847849776f;     // kComponentPairwiseUnindexedFillerAB
        constexpr float kIdxPosD = -kSkew4f;              // kComponentPairwiseIndexedPositiveD
        constexpr float kUnidxD = kSkew4f;                // kComponentPairwiseUnindexedD
        constexpr float kToSingleAB = -0.124000323308709f;
        constexpr float kToSingleD = 0.190983005625053f;
        constexpr float kSingleToExtra = kSkew4f;
        constexpr float kABToC = 0.437016024448821f;
        constexpr float kToDiagonal = -kRoot2f;

        i32 hashShifted = lsr( hash31, 2 );
        i32 indexBasis = imul( hashShifted, 20 >> 2 );                    // [0,20) << 27
        i32 permute3 = imul( hashShifted, (i32)0xD5555556u ) >> 29;       // arithmetic shift, as written there
        i32 permute8 = indexBasis >> 24;
        i32 finalSign = shl( hash31, 31 );

        i32 indexPositive = indexBasis & ( 0x03 << 27 );
        i32 indexNegative = ( indexBasis >> 2 ) & ( 0x03 << 27 );
        indexNegative ^= indexPositive;

        bool singleCase = indexPositive == indexNegative;
#if FN2_TABLES
        // [:370-460] the 32-entry tables carry the "extra" constants in entries 16-19 only, i.e. the comparison as it was meant;
        // every other entry is what the select logic below produces, so the tables are not materialised here.
        const bool extraCase = lsr( indexBasis, 27 ) >= 16;
        bool singleNonExtraCase = singleCase && !extraCase;
#else
        // `indexBasis >= int32v( 0x10 << 27 )`: the literal is 2^31, which the int32 broadcast turns into INT_MIN, so the signed
        // comparison holds for every lane. Reproduced as is (the reference's AVX2 branch has the same expression).
        const bool extraCase = true;
        bool singleNonExtraCase = indexBasis < ( 0x04 << 27 );
#endif
        if( singleNonExtraCase ) indexPositive = -1; // forces the indexPositive checks to fail

        float singleOffsetAB = extraCase ? kToSingleAB + kSingleToExtra : kToSingleAB;
        float idxNegAB = singleCase ? kIdxNegAB + singleOffsetAB : kIdxNegAB;
        float fillAB = singleCase ? kFillAB + singleOffsetAB : kFillAB;
        float idxNegC = singleCase ? idxNegAB : idxNegAB + kABToC;
        float fillC = singleCase ? fillAB : fillAB + kABToC;
        float singleOffsetD = extraCase ? kToSingleD + kSingleToExtra : kToSingleD;
        float idxPosD = singleCase ? kIdxPosD + singleOffsetD : kIdxPosD;
        float unidxD = singleCase ? kUnidxD + singleOffsetD : kUnidxD;

        int counter = -1;
        float valueA = 0, valueB = 0, valueC = 0, valueD = 0;
        const float f[4] = { fX, fY, fZ, fW };
#pragma unroll
        for( int axis = 0; axis < 4; axis++ )
        {
            bool indexedPositive = indexPositive == ( axis << 27 );
            bool indexed = indexedPositive || indexNegative == ( axis << 27 );

            float idxAB = indexedPositive ? idxNegAB + 1.0f : idxNegAB;
            float idxC = indexedPositive ? idxNegC + 1.0f : idxNegC;
            float idxD = !indexedPositive ? idxPosD - 1.0f : idxPosD;

            // which unindexed component takes the diagonal correction depends on how many axes were indexed so far
            float unA, unB, unC;
            if( axis == 0 ) { unA = fillAB + kToDiagonal; unB = fillAB; unC = fillC; }
            else if( axis == 1 ) { unA = counter == -2 ? fillAB + kToDiagonal : fillAB; unB = counter == -1 ? fillAB + kToDiagonal : fillAB; unC = fillC; }
            else if( axis == 2 ) { unA = counter == -3 ? fillAB + kToDiagonal : fillAB; unB = counter == -2 ? fillAB + kToDiagonal : fillAB; unC = counter == -1 ? fillC + kToDiagonal : fillC; }
            else { unA = fillAB; unB = counter == -3 ? fillAB + kToDiagonal : fillAB; unC = counter == -2 ? fillC + kToDiagonal : fillC; }

            float gA = indexed ? idxAB : unA;
            float gB = indexed ? idxAB : unB;
            float gC = indexed ? idxC : unC;
            float gD = indexed ? idxD : unidxD;

            if( axis == 0 )
            {
                valueA = gA * f[0]; valueB = gB * f[0]; valueC = gC * f[0]; valueD = gD * f[0];
            }
            else
            {
                valueA = fma_( gA, f[axis], valueA ); valueB = fma_( gB, f[axis], valueB );
                valueC = fma_( gC, f[axis], valueC ); valueD = fma_( gD, f[axis], valueD );
            }
            if( indexed ) counter -= 1;
        }

        i32 valueIndexX = permute8 >> 1;
        i32 valueIndexY = iadd( iadd( valueIndexX, 1 ), permute3 );
        i32 valueIndexZ = permute8 & 1;
        valueIndexZ = iadd( iadd( valueIndexX, 1 ), valueIndexZ >= permute3 ? iadd( valueIndexZ, 1 ) : valueIndexZ );
        i32 valueIndexSum = iadd( iadd( valueIndexX, valueIndexY ), valueIndexZ );

        multiplier = fxor( multiplier, finalSign );
#define FN2_PICK( IDX ) sel_hi( shl( IDX, 31 ), sel_hi( shl( IDX, 30 ), valueD, valueB ), sel_hi( shl( IDX, 30 ), valueC, valueA ) )
        valueX = fma_( multiplier, FN2_PICK( valueIndexX ), valueX );
        valueY = fma_( multiplier, FN2_PICK( valueIndexY ), valueY );
        valueZ = fma_( multiplier, FN2_PICK( valueIndexZ ), valueZ );
#undef FN2_PICK
        valueW = fma_( multiplier, sel_hi( shl( valueIndexSum, 31 ), sel_hi( shl( valueIndexSum, 30 ), valueD, valueB ), sel_hi( shl( valueIndexSum, 30 ), valueA, valueC ) ), valueW );
    }

    template<int SCHEME> FN2_DI void vec_contribution( i32 h, float fX, float fY, float m, float& vX, float& vY )
    {
        if( SCHEME == 1 ) vec_gop2( h, fX, fY, m, vX, vY ); else vec_ogm2( h, fX, fY, m, vX, vY );
    }
    template<int SCHEME> FN2_DI void vec_contribution( i32 h, float fX, float fY, float fZ, float m, float& vX, float& vY, float& vZ )
    {
        if( SCHEME == 1 ) vec_gop3( h, fX, fY, fZ, m, vX, vY, vZ ); else vec_ogm3( h, fX, fY, fZ, m, vX, vY, vZ );
    }
    template<int SCHEME> FN2_DI void vec_contribution( i32 h, float fX, float fY, float fZ, float fW, float m, float& vX, float& vY, float& vZ, float& vW )
    {
        if( SCHEME == 1 ) vec_gop4( h, fX, fY, fZ, fW, m, vX, vY, vZ, vW ); else vec_ogm4( h, fX, fY, fZ, fW, m, vX, vY, vZ, vW );
    }

    // ------------------------------------------------------------------ the warps
    // LATTICE 0 = Simplex vertices [DomainWarpSimplex.inl:771-1086], 1 = SuperSimplex vertices [:1133-1833].
    // REFLECT = the constant of the "match gradient orientation" step applied for the outer-product scheme in 3D
    // (-2/2 for Simplex :911-917, -2/3 for SuperSimplex :1426-1433).
    template<int LATTICE, int SCHEME>
    FN2_DI float warp_simplex2( i32 seed, float warpAmp, float x, float y, float& xOut, float& yOut, float boundAx, float boundAy )
    {
        float valueX = 0.0f, valueY = 0.0f;
        if( LATTICE == 0 )
        {
            Verts2<float> v[3];
            simplex2_vertices<1>( seed, x, y, v );
#pragma unroll
            for( int k = 0; k < 3; k++ ) vec_contribution<SCHEME>( hash_fold( v[k].hb ), v[k].dx, v[k].dy, v[k].falloff, valueX, valueY );
        }
        else
        {
            Verts2<float> v[4];
            supersimplex2_vertices<1>( seed, x, y, v );
#pragma unroll
            for( int k = 0; k < 4; k++ ) vec_contribution<SCHEME>( hash_fold( v[k].hb ), v[k].dx, v[k].dy, v[k].falloff, valueX, valueY );
        }
        valueX *= boundAx;
        valueY *= boundAy;
        xOut = fma_( valueX, warpAmp, xOut );
        yOut = fma_( valueY, warpAmp, yOut );
        return fma_( valueY, valueY, valueX * valueX );
    }

    template<int LATTICE, int SCHEME>
    FN2_DI float warp_simplex3( i32 seed, float warpAmp, float x, float y, float z, float& xOut, float& yOut, float& zOut, float boundAx, float boundAy, float boundAz )
    {
        float valueX = 0.0f, valueY = 0.0f, valueZ = 0.0f;
        if( LATTICE == 0 )
        {
            Verts3<float> v[4];
            simplex3_vertices<1>( seed, x, y, z, v );
#pragma unroll
            for( int k = 0; k < 4; k++ ) vec_contribution<SCHEME>( hash_fold( v[k].hb ), v[k].dx, v[k].dy, v[k].dz, v[k].falloff, valueX, valueY, valueZ );
        }
        else
        {
            Verts3<float> v[8];
            supersimplex3_vertices<1>( seed, x, y, z, v );
#pragma unroll
            for( int k = 0; k < 8; k++ ) vec_contribution<SCHEME>( hash_fold( v[k].hb ), v[k].dx, v[k].dy, v[k].dz, v[k].falloff, valueX, valueY, valueZ );
        }
        if( SCHEME != 0 ) // match gradient orientation
        {
            const float kReflect3D = LATTICE == 0 ? (float)( -2.0 / 2.0 ) : (float)( -2.0 / 3.0 );
            float delta = kReflect3D * ( valueX + valueY + valueZ );
            valueX += delta;
            valueY += delta;
            valueZ += delta;
        }
        valueX *= boundAx;
        valueY *= boundAy;
        valueZ *= boundAz;
        xOut = fma_( valueX, warpAmp, xOut );
        yOut = fma_( valueY, warpAmp, yOut );
        zOut = fma_( valueZ, warpAmp, zOut );
        return fma_( valueZ, valueZ, fma_( valueY, valueY, valueX * valueX ) );
    }

    template<int LATTICE, int SCHEME>
    FN2_DI float warp_simplex4( i32 seed, float warpAmp, float x, float y, float z, float w, float& xOut, float& yOut, float& zOut, float& wOut,
                                float boundAx, float boundAy, float boundAz, float boundAw )
    {
        float valueX = 0.0f, valueY = 0.0f, valueZ = 0.0f, valueW = 0.0f;
        if( LATTICE == 0 )
        {
            Verts4<float> v[5];
            simplex4_vertices<1>( seed, x, y, z, w, v );
#pragma unroll
            for( int k = 0; k < 5; k++ ) vec_contribution<SCHEME>( hash_fold( v[k].hb ), v[k].dx, v[k].dy, v[k].dz, v[k].dw, v[k].falloff, valueX, valueY, valueZ, valueW );
        }
        else
        {
            Verts4<float> v[16];
            supersimplex4_vertices<1>( seed, x, y, z, w, v );
#pragma unroll
            for( int k = 0; k < 16; k++ ) vec_contribution<SCHEME>( hash_fold( v[k].hb ), v[k].dx, v[k].dy, v[k].dz, v[k].dw, v[k].falloff, valueX, valueY, valueZ, valueW );
        }
        valueX *= boundAx;
        valueY *= boundAy;
        valueZ *= boundAz;
        valueW *= boundAw;
        xOut = fma_( valueX, warpAmp, xOut );
        yOut = fma_( valueY, warpAmp, yOut );
        zOut = fma_( valueZ, warpAmp, zOut );
        wOut = fma_( valueW, warpAmp, wOut );
        return fma_( valueW, valueW, fma_( valueZ, valueZ, fma_( valueY, valueY, valueX * valueX ) ) );
    }

    // pack forms, lane by lane: the vector contributions are index arithmetic and selects almost throughout
    template<int LATTICE, int SCHEME>
    FN2_DI fpair warp_simplex2( i32 seed, fpair warpAmp, fpair x, fpair y, fpair& xOut, fpair& yOut, float boundAx, float boundAy )
    {
        fpair r;
        r.x = warp_simplex2<LATTICE, SCHEME>( seed, warpAmp.x, x.x, y.x, xOut.x, yOut.x, boundAx, boundAy );
        r.y = warp_simplex2<LATTICE, SCHEME>( seed, warpAmp.y, x.y, y.y, xOut.y, yOut.y, boundAx, boundAy );
        return r;
    }
    template<int LATTICE, int SCHEME>
    FN2_DI fpair warp_simplex3( i32 seed, fpair warpAmp, fpair x, fpair y, fpair z, fpair& xOut, fpair& yOut, fpair& zOut, float boundAx, float boundAy, float boundAz )
    {
        fpair r;
        r.x = warp_simplex3<LATTICE, SCHEME>( seed, warpAmp.x, x.x, y.x, z.x, xOut.x, yOut.x, zOut.x, boundAx, boundAy, boundAz );
        r.y = warp_simplex3<LATTICE, SCHEME>( seed, warpAmp.y, x.y, y.y, z.y, xOut.y, yOut.y, zOut.y, boundAx, boundAy, boundAz );
        return r;
    }
    template<int LATTICE, int SCHEME>
    FN2_DI fpair warp_simplex4( i32 seed, fpair warpAmp, fpair x, fpair y, fpair z, fpair w, fpair& xOut, fpair& yOut, fpair& zOut, fpair& wOut,
                                float boundAx, float boundAy, float boundAz, float boundAw )
    {
        fpair r;
        r.x = warp_simplex4<LATTICE, SCHEME>( seed, warpAmp.x, x.x, y.x, z.x, w.x, xOut.x, yOut.x, zOut.x, wOut.x, boundAx, boundAy, boundAz, boundAw );
        r.y = warp_simplex4<LATTICE, SCHEME>( seed, warpAmp.y, x.y, y.y, z.y, w.y, xOut.y, yOut.y, zOut.y, wOut.y, boundAx, boundAy, boundAz, boundAw );
        return r;
    }
}
