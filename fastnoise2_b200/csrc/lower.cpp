// fastnoise2_b200 -- lowering: node tree -> CUDA source of one fused kernel.
//
// The reference evaluates a tree by virtual calls, one per child per SIMD vector
// (Generators/Generator.inl:48-67 GetSourceValue). Here the tree rooted at the node a Gen* call was made on
// is turned into straight-line device code: one `n<K>_<D>` function per (node, dimensionality) reached,
// with every uniform parameter folded in as a literal and every uniform branch of the reference (distance
// function, return type, interpolation, "is this hybrid a node?") resolved while generating. The functions
// are written against the value types F / I / B -- float / int / bool for one point per value, or the
// fpair / ipair / bpair of device/fn2_device_pack.cuh for two points per value (packed FP32), chosen per
// kernel by LowerTree -- and call the hand-written bodies in device/fn2_device*.cuh. Shared sub-trees are emitted once. The entry
// kernel appended at the end is the batch driver for one call shape (grid 2/3/4D, tileable, positions).
#include <FastNoise/FastNoise.h>
#include <FastNoise/Metadata.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <sstream>
#include <algorithm>
#include <vector>
#include <unordered_map>

#include "internal.h"
#include "node_slots.h"

using namespace FastNoise;
using namespace FastNoise::Slots;

namespace
{
    // exact float literal
    std::string Lit( float v )
    {
        char buf[64];
        if( std::isnan( v ) )
        {
            std::snprintf( buf, sizeof( buf ), "__int_as_float(0x%08x)", (unsigned)Impl::Bits( v ) );
        }
        else if( std::isinf( v ) )
        {
            std::snprintf( buf, sizeof( buf ), v < 0 ? "(-FN2_INF)" : "FN2_INF" );
        }
        else
        {
            std::snprintf( buf, sizeof( buf ), "%af", v );
            if( v < 0 || ( v == 0 && std::signbit( v ) ) ) return std::string( "(" ) + buf + ")";
        }
        return buf;
    }

    std::string Int( int v )
    {
        char buf[32];
        std::snprintf( buf, sizeof( buf ), v < 0 ? "(%d)" : "%d", v );
        if( v == INT32_MIN ) return "(-2147483647 - 1)";
        return buf;
    }

    const char* const kAxis[4] = { "x", "y", "z", "w" };

    int EnvFlag( const char* name, int fallback )
    {
        const char* v = std::getenv( name );
        return v && *v ? std::atoi( v ) : fallback;
    }

    // Code-shape options of the generated text (each measured on B200, profiles/r2_perf_sweep_c_code_shape.txt). Both are bit-neutral.
    //   one-body loops: a fractal's first octave runs inside its octave loop, so the source sub-tree appears once in the code
    //   convergent walk: every thread of a block runs the same number of groups (threads past the end re-evaluate the last group
    //                    and skip the store), so the loop control and everything uniform inside it stays on the uniform datapath
    bool OneBodyLoops() { return EnvFlag( "FASTNOISE2_B200_ONE_BODY", 0 ) != 0; }
    bool ConvergentWalk() { return EnvFlag( "FASTNOISE2_B200_CONVERGENT", 1 ) != 0; }

    // "x, y, z"
    std::string Pos( int dim, const char* prefix = "" )
    {
        std::string s;
        for( int d = 0; d < dim; d++ )
        {
            if( d ) s += ", ";
            s += prefix;
            s += kAxis[d];
        }
        return s;
    }

    // "F x, F y, F z" -- F is float or fn2::fpair, see the typedefs LowerTree emits
    std::string PosDecl( int dim, const char* type = "F ", const char* prefix = "" )
    {
        std::string s;
        for( int d = 0; d < dim; d++ )
        {
            if( d ) s += ", ";
            s += type;
            s += prefix;
            s += kAxis[d];
        }
        return s;
    }

    struct Emitter
    {
        std::ostringstream out;
        std::unordered_map<const Generator*, int> index;
        std::set<std::pair<int, int>> genDone, warpDone;
        std::string error;
        int depth = 0;
        bool paramMode = false;     // float node parameters travel as kernel arguments (tp[k]) instead of literals
        std::vector<float> params;  // their values, in slot order

        // A float node parameter (or a value folded from node parameters on the host): a literal, or -- in parameter mode -- a
        // slot of the kernel's TreeParams argument, so that one compiled kernel serves every tree of the same shape. Values the
        // emitter LOOKS at to choose between code shapes (a zero weighted strength, a clamp bound that cannot be NaN ...) still
        // change the text, and with it the kernel, when the choice changes.
        std::string Par( float v )
        {
            if( !paramMode ) return Lit( v );
            params.push_back( v );
            return "tp[" + std::to_string( params.size() - 1 ) + "]";
        }

        int IndexOf( const Generator* g )
        {
            auto it = index.find( g );
            if( it != index.end() ) return it->second;
            int k = (int)index.size();
            index.emplace( g, k );
            return k;
        }

        void Fail( const std::string& why )
        {
            if( error.empty() ) error = why;
        }

        // ---------------------------------------------------------------- naming + recursion
        std::string Gen( const Generator* g, int dim )
        {
            if( !g )
            {
                Fail( "a required source node is not set" );
                return "fn2_missing";
            }
            int k = IndexOf( g );
            if( genDone.insert( { k, dim } ).second )
            {
                if( ++depth > 256 )
                {
                    Fail( "node tree is too deep or cyclic" );
                }
                else
                {
                    std::string body = Body( g, k, dim ); // emits the children first
                    out << "// " << g->GetMetadata().name << " " << dim << "D\n"
                        << "FN2_DI F n" << k << "_" << dim << "( int seed, " << PosDecl( dim ) << ", FN2_TP )\n{\n" << body << "}\n\n";
                }
                --depth;
            }
            return "n" + std::to_string( k ) + "_" + std::to_string( dim );
        }

        // call expression of a required source
        std::string Call( const Generator* g, int dim, const std::string& seed, const std::string& pos )
        {
            return Gen( g, dim ) + "( " + seed + ", " + pos + ", tp )";
        }

        std::string Src( const Generator* owner, int lookup, int dim, const std::string& seed, const std::string& pos )
        {
            return Call( owner->State().lookups[lookup].get(), dim, seed, pos );
        }

        // hybrid: literal or call
        std::string Hyb( const Generator* owner, int hybrid, int dim, const std::string& seed, const std::string& pos )
        {
            const Impl::NodeState::Hybrid& h = owner->State().hybrids[hybrid];
            if( h.node ) return Call( h.node.get(), dim, seed, pos );
            return Par( h.constant );
        }

        static bool HybIsConst( const Generator* owner, int hybrid, float value )
        {
            const Impl::NodeState::Hybrid& h = owner->State().hybrids[hybrid];
            return !h.node && h.constant == value;
        }

        // FS::Min / FS::Max with the reference's operand order. The one-instruction form (FMNMX) equals minps / maxps unless the
        // second operand is NaN or the operands are zeros of opposite sign (fn2_device.cuh, min_x86): it is emitted when the
        // second operand is a constant that rules both out, the compare + select form otherwise.
        static bool ClampSafe( float second, bool isMax )
        {
            return second == second && !( second == 0.0f && std::signbit( second ) == isMax );
        }
        static bool HybClampSafe( const Generator* owner, int hybrid, bool isMax )
        {
            const Impl::NodeState::Hybrid& h = owner->State().hybrids[hybrid];
            return !h.node && ClampSafe( h.constant, isMax );
        }
        static std::string MinMax( bool isMax, const std::string& first, const std::string& second, bool safe )
        {
            return std::string( isMax ? "fn2::max_" : "fn2::min_" ) + ( safe ? "( " : "x86( " ) + first + ", " + second + " )";
        }

        // Can this node return only finite values or NaN, never +-inf? True for the lattice and hash leaves with a sane output
        // range: their native value is bounded for every input (a non-finite position turns into NaN inside them).
        // Used by the fractals: Lerp( 1, t, 0 ) is 1 for every finite t and NaN for an infinite one.
        static bool FiniteOrNaN( const Generator* node )
        {
            if( !node ) return false;
            const Impl::NodeState& s = node->State();
            switch( s.nodeId )
            {
            case N_Perlin: case N_Simplex: case N_SuperSimplex: case N_Value: case N_White: case N_Checkerboard: case N_SineWave: case N_CellularValue:
                return std::fabs( s.f[F_RangeScale] ) < 0x1p100f && std::fabs( s.f[F_RangeMin] ) < 0x1p100f;
            default:
                return false;
            }
        }

        // A hybrid input that is finite at every finite position: a finite constant, or one of the bounded leaves with a finite
        // frequency. The Cellular nodes keep FMNMX in their distance sorting network when all their inputs are like that.
        static bool HybAlwaysFinite( const Generator* owner, int hybrid )
        {
            const Impl::NodeState::Hybrid& h = owner->State().hybrids[hybrid];
            if( !h.node ) return std::isfinite( h.constant );
            return FiniteOrNaN( h.node.get() ) && std::isfinite( h.node->State().f[F_Frequency] );
        }

        // ScaleOutput [BasicGenerators.inl:27-30]: fma( (1/(max-min)) * rangeScale, value - min, rangeMin ), factor folded on the host in float
        std::string Range( const Impl::NodeState& s, const std::string& value, float nativeMin, float nativeMax )
        {
            float k = ( 1.0f / ( nativeMax - nativeMin ) ) * s.f[F_RangeScale];
            return "fn2::fma_( " + Par( k ) + ", " + value + " - " + Lit( nativeMin ) + ", " + Par( s.f[F_RangeMin] ) + " )";
        }

        std::string ScalePositions( const Impl::NodeState& s, int dim )
        {
            std::string code;
            const std::string f = Par( s.f[F_Frequency] );
            for( int d = 0; d < dim; d++ ) code += std::string( "    " ) + kAxis[d] + " *= " + f + ";\n";
            return code;
        }

        static std::string AddSeedOffset( const Impl::NodeState& s )
        {
            return "    seed = fn2::iadd( seed, " + Int( s.i[I_SeedOffset] ) + " );\n";
        }

        // ---------------------------------------------------------------- domain warps
        static bool IsWarp( std::uint8_t id ) { return id == N_DomainWarpGradient || id == N_DomainWarpSimplex || id == N_DomainWarpSuperSimplex; }

        // w<K>_<D>( seed, amp, scaled pos..., pos refs... ) -> strength, the reference's virtual Warp() [DomainWarp.inl:24-26]
        std::string Warp( const Generator* g, int dim )
        {
            int k = IndexOf( g );
            std::string name = "w" + std::to_string( k ) + "_" + std::to_string( dim );
            if( !warpDone.insert( { k, dim } ).second ) return name;

            const Impl::NodeState& s = g->State();
            out << "// " << g->GetMetadata().name << "::Warp " << dim << "D\n"
                << "FN2_DI F " << name << "( int seed, F warpAmp, " << PosDecl( dim, "F ", "s" ) << ", " << PosDecl( dim, "F& " ) << ", FN2_TP )\n{\n";
            if( s.nodeId == N_DomainWarpGradient )
            {
                out << AddSeedOffset( s ) // Warp() adds the offset itself [DomainWarp.inl:35,79,136]
                    << "    return fn2::warp_gradient" << dim << "( seed, warpAmp, " << Pos( dim, "s" ) << ", " << Pos( dim );
                for( int d = 0; d < dim; d++ ) out << ", " << Par( s.f[F_K0 + d] );
                out << " );\n";
            }
            else
            {
                // DomainWarpSimplex::Warp adds the seed offset itself, DomainWarpSuperSimplex::Warp does not
                // [DomainWarpSimplex.inl:733,746,759 vs :1094-1128]
                const bool super = s.nodeId == N_DomainWarpSuperSimplex;
                const int scheme = s.i[I_K0] == 1 ? 1 : 0; // `default:` falls to OrthogonalGradientMatrix
                // kBounding per lattice / dimension / scheme [:819-821,919-921,1073-1075,1215-1217,1435-1437,1822-1824]
                static const float kBound[2][3][2] = {
                    { { 70.1480577066486f, (float)( 49.918426513671875 / 2.0 ) }, { 16.281631889139874f, (float)( 32.69428253173828125 / 1.4142135623730951 ) },
                      { 30.88161777516092f, (float)( 33.653125584827855 / 1.4142135623730951 ) } },
                    { { 12.814453124999995f, (float)( 9.28993664146183 / 2.0 ) }, { 37.63698669623629f, (float)( 144.736422163332608 / 1.4142135623730951 ) },
                      { 48.80058117543753f, (float)( 115.21625311930542 / 1.4142135623730951 ) } } };
                if( !super ) out << AddSeedOffset( s );
                out << "    return fn2::warp_simplex" << dim << "<" << ( super ? 1 : 0 ) << ", " << scheme << ">( seed, warpAmp, " << Pos( dim, "s" ) << ", " << Pos( dim );
                for( int d = 0; d < dim; d++ ) out << ", " << Par( kBound[super ? 1 : 0][dim - 2][scheme] * s.f[F_K0 + d] );
                out << " );\n";
            }
            out << "}\n\n";
            return name;
        }

        // "sx*f, sy*f" helper for warp calls
        static std::string Scaled( int dim, const std::string& factor, const char* prefix = "" )
        {
            std::string s;
            for( int d = 0; d < dim; d++ )
            {
                if( d ) s += ", ";
                s += std::string( prefix ) + kAxis[d] + " * " + factor;
            }
            return s;
        }

        // ---------------------------------------------------------------- per node bodies
        std::string Body( const Generator* g, int k, int dim )
        {
            const Impl::NodeState& s = g->State();
            const std::string P = Pos( dim );
            const std::string D = std::to_string( dim );
            std::ostringstream b;

            auto cellularPrologue = [&]( bool ranged ) // common head of the three Cellular nodes [Cellular.inl:22-29]
            {
                (void)ranged;
                const char* jitterK = dim == 2 ? "fn2::kJitter2D" : ( dim == 3 ? "fn2::kJitter3D" : "fn2::kJitter4D" );
                const char* sizeK = dim == 2 ? "fn2::kSizeJitterScale2D" : ( dim == 3 ? "fn2::kSizeJitterScale3D" : "fn2::kSizeJitterScale4D" );
                bool sizeJ = s.hybrids[2].node || s.hybrids[2].constant != 0.0f;
                b << "    const int sourceSeed = seed;\n" << AddSeedOffset( s );
                b << "    const F jitter = " << jitterK << " * " << Hyb( g, 1, dim, "sourceSeed", P ) << ";\n";
                if( sizeJ ) b << "    const F sizeJitter = " << Hyb( g, 2, dim, "sourceSeed", P ) << " * " << sizeK << ";\n";
                else b << "    const F sizeJitter = 0.0f;\n";
                b << ScalePositions( s, dim );
                // Minkowski exponent is sampled AT the cell delta vector [Utils.inl:330]
                b << "    auto minkowskiP = [=]( " << PosDecl( dim, "F ", "d" ) << " ) -> F { return " << Hyb( g, 0, dim, "sourceSeed", Pos( dim, "d" ) ) << "; };\n";
                return sizeJ;
            };

            switch( s.nodeId )
            {
            case N_Constant:
                b << "    return " << Par( s.f[F_K0] ) << ";\n";
                break;

            case N_White:
                b << AddSeedOffset( s ) << "    return " << Range( s, "fn2::white( seed, " + P + " )", -2147483648.f, 2147483648.f ) << ";\n";
                break;

            case N_Checkerboard:
                b << ScalePositions( s, dim ) << "    return " << Range( s, "fn2::checkerboard( " + P + " )", 0, 2 ) << ";\n";
                break;

            case N_SineWave:
                b << ScalePositions( s, dim ) << "    return " << Range( s, "fn2::sinewave( " + P + " )", -1, 1 ) << ";\n";
                break;

            case N_Gradient: // [BasicGenerators.inl:98-108]
                b << "    F r = 0.0f;\n";
                for( int d = 0; d < dim; d++ )
                {
                    b << "    r = fn2::fma_( " << kAxis[d] << " + " << Hyb( g, d, dim, "seed", P ) << ", " << Par( s.f[F_K0 + d] ) << ", r );\n";
                }
                b << "    return r;\n";
                break;

            case N_DistanceToPoint: // [BasicGenerators.inl:115-125]
            {
                std::string O = Pos( dim, "o" );
                for( int d = 0; d < dim; d++ ) b << "    const F o" << kAxis[d] << " = " << kAxis[d] << ";\n";
                for( int d = 0; d < dim; d++ ) b << "    " << kAxis[d] << " -= " << Hyb( g, d, dim, "seed", O ) << ";\n";
                int df = s.i[I_K0];
                if( df < 0 || df > 5 ) df = 0; // `default:` label of the switch [Utils.inl:286]
                b << "    return fn2::calc_distance<" << df << ", true>( F( " << ( df == 5 ? Hyb( g, 4, dim, "seed", P ) : "0.0f" ) << " ), " << P << " );\n";
                break;
            }

            case N_Simplex:
            case N_SuperSimplex:
            case N_Perlin:
            {
                float bound;
                const char* fn;
                if( s.nodeId == N_Simplex ) { fn = "simplex"; bound = dim == 2 ? 0.0261208079755306243896484375f : ( dim == 3 ? 0.030586399137973785400390625f : 0.036730043590068817138671875f ); }
                else if( s.nodeId == N_SuperSimplex ) { fn = "supersimplex"; bound = dim == 2 ? 0.14084912836551666259765625f : ( dim == 3 ? 0.0069091119803488254547119140625f : (float)0.011310013942420482635498046875 ); }
                else { fn = "perlin"; bound = dim == 2 ? 1.726796627044677734375f : ( dim == 3 ? (float)1.0363423824310302734375 : 1.33858621120452880859375f ); }
                b << AddSeedOffset( s ) << ScalePositions( s, dim )
                  << "    return " << Range( s, std::string( "fn2::" ) + fn + D + "( seed, " + P + " )", -bound, bound ) << ";\n";
                break;
            }

            case N_Value:
                b << AddSeedOffset( s ) << ScalePositions( s, dim )
                  << "    return " << Range( s, "fn2::value" + D + "( seed, " + P + " )", -2147483648.f, 2147483648.f ) << ";\n";
                break;

            case N_CellularValue: // [Cellular.inl:18-292]
            {
                bool sizeJ = cellularPrologue( true );
                int df = s.i[I_K0];
                if( df < 0 || df > 5 ) df = 0;
                int top = s.i[I_K0 + 1];
                b << "    const I cellHash = fn2::cellular_value" << D << "<" << df << ", " << ( sizeJ ? "true" : "false" ) << ", " << top
                  << ">( seed, " << P << ", jitter, sizeJitter, minkowskiP );\n";
                b << "    return " << Range( s, "fn2::cvt_f( cellHash )", -2147483648.f, 2147483648.f ) << ";\n";
                break;
            }

            case N_CellularDistance: // [Cellular.inl:295-566]
            {
                bool sizeJ = cellularPrologue( true );
                int df = s.i[I_K0];
                // an out-of-range distance function enters CalcDistance<false> at `default:` and falls through to the
                // EuclideanSquared case [Utils.inl:286-302], and GetReturn's Euclidean test fails [Cellular.inl:533]
                if( df < 0 || df > 5 ) df = 1;
                int i0 = s.i[I_K0 + 1], i1 = s.i[I_K0 + 2], ret = s.i[I_K0 + 3];
                if( ret < 0 || ret > 4 ) ret = 0;
                int top = ( df == 0 || ret != 0 ) ? std::max( i0, i1 ) : i0;
                b << "    F dist[4];\n";
                // a NaN candidate distance (possible only through a non-finite jitter / size jitter / Minkowski P input) moves
                // through the reference's min / max network in operand order: fn2::distance_insert<TOP, ORDERED>
                const bool ordered = !HybAlwaysFinite( g, 1 ) || ( sizeJ && !HybAlwaysFinite( g, 2 ) ) || ( df == 5 && !HybAlwaysFinite( g, 0 ) );
                b << "    fn2::cellular_distance" << D << "<" << df << ", " << ( sizeJ ? "true" : "false" ) << ", " << top << ( ordered ? ", true" : "" )
                  << ">( seed, " << P << ", jitter, sizeJitter, minkowskiP, dist );\n";
                if( df == 0 ) // GetReturn [Cellular.inl:533-537]; with index0 == index1 the entry is normalised twice, as there
                {
                    b << "    dist[" << i0 << "] *= fn2::inv_sqrt( dist[" << i0 << "] );\n";
                    b << "    dist[" << i1 << "] *= fn2::inv_sqrt( dist[" << i1 << "] );\n";
                }
                float jitterK = dim == 2 ? 0.437016f : ( dim == 3 ? 0.396144f : 0.366025f );
                float maxDist = 1 + jitterK;
                maxDist *= maxDist;
                std::string d0 = "dist[" + std::to_string( i0 ) + "]", d1 = "dist[" + std::to_string( i1 ) + "]";
                switch( ret )
                {
                case 0: b << "    return " << Range( s, d0, 0, maxDist ) << ";\n"; break;
                case 1: b << "    return " << Range( s, "( " + d0 + " + " + d1 + " )", 0, maxDist * 2 ) << ";\n"; break;
                case 2: b << "    return " << Range( s, "fn2::abs_( " + d0 + " - " + d1 + " )", 0, maxDist ) << ";\n"; break;
                case 3: b << "    return " << Range( s, "( " + d0 + " * " + d1 + " )", 0, maxDist * maxDist ) << ";\n"; break;
                case 4: b << "    return " << Range( s, "( " + d0 + " * fn2::reciprocal( " + d1 + " ) )", 0, maxDist ) << ";\n"; break;
                }
                break;
            }

            case N_CellularLookup: // [Cellular.inl:569-806]
            {
                bool sizeJ = cellularPrologue( false );
                int df = s.i[I_K0];
                if( df < 0 || df > 5 ) df = 0;
                std::string C = Pos( dim, "c" );
                b << "    F " << C << ";\n";
                b << "    fn2::cellular_lookup" << D << "<" << df << ", " << ( sizeJ ? "true" : "false" ) << ">( seed, " << P << ", jitter, sizeJitter, minkowskiP, " << C << " );\n";
                b << "    return " << Src( g, 0, dim, "fn2::iadd( sourceSeed, 1 )", Scaled( dim, Par( s.f[F_Scale] ), "c" ) ) << ";\n";
                break;
            }

            case N_FractalFBm:   // [Fractal.inl:14-36]
            case N_FractalRidged: // [Fractal.inl:44-66]
            {
                const bool ridged = s.nodeId == N_FractalRidged;
                // with a constant weighted strength of 0 the factor is fma( 0, t - 1, 1 ): 1 for every finite t, NaN when the
                // source returned +-inf (then the reference's sum ends as NaN, not inf) -- skipped only over sources that cannot
                const bool weighted = !( HybIsConst( g, 1, 0.0f ) && FiniteOrNaN( g->State().lookups[0].get() ) );
                int octaves = s.i[I_K0];
                std::string sample = Src( g, 0, dim, "seed", P );
                if( ridged ) sample = "fn2::abs_( " + sample + " )";
                b << "    const F gain = " << Hyb( g, 0, dim, "seed", P ) << ";\n";
                if( weighted ) b << "    const F weightedStrength = " << Hyb( g, 1, dim, "seed", P ) << ";\n";
                const std::string term = ridged ? "( noise * -2.0f + 1.0f ) * amp" : "noise * amp";
                const std::string lacunarity = Par( s.f[F_K0] );
                auto nextOctave = [&]( const char* indent )
                {
                    b << indent << "seed = fn2::iadd( seed, 1 );\n";
                    if( weighted ) b << indent << "amp *= fn2::lerp( 1.0f, " << ( ridged ? "1.0f - noise" : "( noise + 1.0f ) * 0.5f" ) << ", weightedStrength );\n";
                    b << indent << "amp *= gain;\n";
                    for( int d = 0; d < dim; d++ ) b << indent << kAxis[d] << " *= " << lacunarity << ";\n";
                };
                b << "    F amp = 1.0f;\n";
                if( OneBodyLoops() && octaves > 1 )
                {
                    // One copy of the source in the code instead of two (first octave + loop body): the first octave runs inside
                    // the loop. sum starts at -0.0f, the identity of IEEE addition for EVERY addend (-0 + v == v bit for bit, signed
                    // zeros included), so the first `sum +=` is the reference's `sum = noise * amp`.
                    b << "    F noise;\n    F sum = -0.0f;\n";
                    b << "    int octave = 0;\n    FN2_ONE_BODY_LOOP\n    do\n    {\n";
                    b << "        noise = " << sample << ";\n";
                    b << "        sum += " << term << ";\n";
                    nextOctave( "        " ); // also after the last octave (unused there): a conditional step would be peeled into a second copy
                    b << "    } while( ++octave < " << octaves << " );\n";
                }
                else
                {
                    b << "    F noise = " << sample << ";\n";
                    b << "    F sum = " << term << ";\n";
                    if( octaves > 1 )
                    {
                        b << "    FN2_OCTAVE_LOOP\n    for( int octave = 1; octave < " << octaves << "; octave++ )\n    {\n";
                        nextOctave( "        " );
                        b << "        noise = " << sample << ";\n";
                        b << "        sum += " << term << ";\n";
                        b << "    }\n";
                    }
                }
                b << "    return sum;\n";
                break;
            }

            case N_PingPong: // [Modifiers.inl:185-201]
                b << "    const F strength = " << Hyb( g, 0, dim, "seed", P ) << ";\n";
                b << "    F t = " << Src( g, 0, dim, "seed", P ) << " * strength;\n";
                b << "    t -= fn2::floor_( t * 0.5f ) * 2.0f;\n";
                b << "    return fn2::select( t < 1.0f, t, 2.0f - t );\n";
                break;

            case N_DomainWarpSimplex:
            case N_DomainWarpSuperSimplex:
            case N_DomainWarpGradient: // DomainWarp::GenT [DomainWarp.inl:9-17]
            {
                std::string warp = Warp( g, dim );
                b << "    const int sourceSeed = seed;\n" << AddSeedOffset( s );
                b << "    const F warpAmp = " << Hyb( g, 0, dim, "sourceSeed", P ) << ";\n";
                b << "    " << warp << "( seed, warpAmp, " << Scaled( dim, Par( s.f[F_Frequency] ) ) << ", " << P << ", tp );\n";
                b << "    return " << Src( g, 0, dim, "sourceSeed", P ) << ";\n";
                break;
            }

            case N_DomainWarpFractalProgressive: // [DomainWarpFractal.inl:9-35]
            case N_DomainWarpFractalIndependent: // [DomainWarpFractal.inl:43-73]
            {
                const Generator* warpNode = s.lookups[0].get();
                if( !warpNode || !IsWarp( warpNode->State().nodeId ) )
                {
                    Fail( "DomainWarpFractal needs a DomainWarp node as its source" );
                    b << "    return 0.0f;\n";
                    break;
                }
                const Impl::NodeState& ws = warpNode->State();
                const bool independent = s.nodeId == N_DomainWarpFractalIndependent;
                const bool weighted = !HybIsConst( g, 1, 0.0f );
                int octaves = s.i[I_K0];
                std::string warp = Warp( warpNode, dim );
                std::string N = independent ? Pos( dim, "n" ) : P; // position the frequency is applied to
                b << "    F amp = " << Hyb( warpNode, 0, dim, "seed", P ) << ";\n";
                if( weighted ) b << "    const F weightedStrength = " << Hyb( g, 1, dim, "seed", P ) << ";\n";
                b << "    F freq = " << Par( ws.f[F_Frequency] ) << ";\n";
                b << "    int seedInc = seed;\n";
                b << "    const F gain = " << Hyb( g, 0, dim, "seed", P ) << ";\n";
                if( independent ) for( int d = 0; d < dim; d++ ) b << "    const F n" << kAxis[d] << " = " << kAxis[d] << ";\n";
                const std::string warpCall = warp + "( seedInc, amp, " + Scaled( dim, "freq", independent ? "n" : "" ) + ", " + P + ", tp )";
                auto nextWarpOctave = [&]( const char* indent )
                {
                    if( weighted ) b << indent << "strength = fn2::fast_length_sqrt( strength );\n";
                    b << indent << "seedInc = fn2::iadd( seedInc, 1 );\n";
                    b << indent << "freq *= " << Par( s.f[F_K0] ) << ";\n";
                    if( weighted ) b << indent << "amp *= fn2::lerp( 1.0f, 1.0f - strength, weightedStrength );\n";
                    b << indent << "amp *= gain;\n";
                };
                if( OneBodyLoops() && octaves > 1 )
                {
                    b << "    F strength;\n";
                    b << "    int octave = 0;\n    FN2_ONE_BODY_LOOP\n    do\n    {\n";
                    b << "        strength = " << warpCall << ";\n";
                    nextWarpOctave( "        " );
                    b << "    } while( ++octave < " << octaves << " );\n";
                }
                else
                {
                    b << "    F strength = " << warpCall << ";\n";
                    if( octaves > 1 )
                    {
                        b << "    FN2_OCTAVE_LOOP\n    for( int octave = 1; octave < " << octaves << "; octave++ )\n    {\n";
                        nextWarpOctave( "        " );
                        b << "        strength = " << warpCall << ";\n";
                        b << "    }\n";
                    }
                }
                b << "    (void)strength;\n";
                b << "    return " << Src( warpNode, 0, dim, "seed", P ) << ";\n";
                break;
            }

            case N_Add:
                b << "    return " << Src( g, 0, dim, "seed", P ) << " + " << Hyb( g, 0, dim, "seed", P ) << ";\n";
                break;
            case N_Multiply:
                b << "    return " << Src( g, 0, dim, "seed", P ) << " * " << Hyb( g, 0, dim, "seed", P ) << ";\n";
                break;
            case N_Min:
                b << "    return " << MinMax( false, Src( g, 0, dim, "seed", P ), Hyb( g, 0, dim, "seed", P ), HybClampSafe( g, 0, false ) ) << ";\n";
                break;
            case N_Max:
                b << "    return " << MinMax( true, Src( g, 0, dim, "seed", P ), Hyb( g, 0, dim, "seed", P ), HybClampSafe( g, 0, true ) ) << ";\n";
                break;
            case N_Subtract:
                b << "    return " << Hyb( g, 0, dim, "seed", P ) << " - " << Hyb( g, 1, dim, "seed", P ) << ";\n";
                break;
            case N_Divide:
                b << "    return fn2::div_( " << Hyb( g, 0, dim, "seed", P ) << ", " << Hyb( g, 1, dim, "seed", P ) << " );\n";
                break;
            case N_Modulus:
                b << "    return fn2::modulus( " << Hyb( g, 0, dim, "seed", P ) << ", " << Hyb( g, 1, dim, "seed", P ) << " );\n";
                break;

            case N_Abs:
                b << "    return fn2::abs_( " << Src( g, 0, dim, "seed", P ) << " );\n";
                break;

            case N_SignedSquareRoot: // [Modifiers.inl:324-335]
                b << "    const F v = " << Src( g, 0, dim, "seed", P ) << ";\n";
                b << "    return fn2::bitsf( fn2::fbits( fn2::fast_length_sqrt( fn2::abs_( v ) ) ) | ( fn2::fbits( v ) & (int)0x80000000u ) );\n";
                break;

            case N_MinSmooth: // [Blends.inl:126-143]
            case N_MaxSmooth: // [Blends.inl:145-162]
            {
                const char* neg = s.nodeId == N_MaxSmooth ? "-" : "";
                b << "    const F a = " << neg << Src( g, 0, dim, "seed", P ) << ";\n";
                b << "    const F b = " << neg << Hyb( g, 0, dim, "seed", P ) << ";\n";
                // |smoothness| is +0 or larger: only a NaN can make the two forms differ
                const Impl::NodeState::Hybrid& sm = s.hybrids[1];
                if( !sm.node && sm.constant == sm.constant )
                {
                    // constant smoothness: the clamp and the reciprocal are folded here, in the same float arithmetic
                    const float smoothness = std::fmax( 0x1p-126f, std::fabs( sm.constant ) );
                    b << "    const F smoothness = " << Par( smoothness ) << ";\n";
                    b << "    F h = fn2::max_( smoothness - fn2::abs_( a - b ), 0.0f );\n";
                    b << "    h *= " << Par( 1.0f / smoothness ) << ";\n";
                }
                else
                {
                    b << "    const F smoothness = " << MinMax( true, "0x1p-126f", "fn2::abs_( " + Hyb( g, 1, dim, "seed", P ) + " )", HybClampSafe( g, 1, false ) && HybClampSafe( g, 1, true ) ) << ";\n";
                    b << "    F h = fn2::max_( smoothness - fn2::abs_( a - b ), 0.0f );\n";
                    b << "    h *= fn2::reciprocal( smoothness );\n";
                }
                b << "    return " << neg << "fn2::fnma_( " << Lit( 1.0f / 6.0f ) << ", h * h * h * smoothness, " << MinMax( false, "a", "b", HybClampSafe( g, 0, false ) && HybClampSafe( g, 0, true ) ) << " );\n";
                break;
            }

            case N_PowFloat: // [Blends.inl:72-82]
                b << "    const F v = fn2::max_( fn2::abs_( " << Hyb( g, 0, dim, "seed", P ) << " ), 0x1p-126f );\n";
                b << "    return fn2::pow_pos( v, " << Hyb( g, 1, dim, "seed", P ) << " );\n";
                break;

            case N_PowInt: // [Blends.inl:84-101]
                b << "    const F v = " << Src( g, 0, dim, "seed", P ) << ";\n";
                b << "    F p = v * v;\n";
                b << "    for( int i = 2; i < " << Int( s.i[I_K0] ) << "; i++ ) p *= v;\n";
                b << "    return p;\n";
                break;

            case N_DomainScale:
                b << "    return " << Src( g, 0, dim, "seed", Scaled( dim, Par( s.f[F_K0] ) ) ) << ";\n";
                break;

            case N_DomainOffset: // [Modifiers.inl:15-28] every offset is sampled at the un-offset position
            {
                std::string O = Pos( dim, "o" );
                for( int d = 0; d < dim; d++ ) b << "    const F o" << kAxis[d] << " = " << kAxis[d] << " + " << Hyb( g, d, dim, "seed", P ) << ";\n";
                b << "    return " << Src( g, 0, dim, "seed", O ) << ";\n";
                break;
            }

            case N_DomainRotate: // [Modifiers.inl:30-61], matrix as Modifiers.h CalculateRotation
            {
                const float yc = s.f[F_K0], ys = s.f[F_K0 + 1], pc = s.f[F_K0 + 2], ps = s.f[F_K0 + 3], rc = s.f[F_K0 + 4], rs = s.f[F_K0 + 5];
                if( dim == 4 )
                {
                    b << "    return " << Src( g, 0, 4, "seed", P ) << ";\n";
                }
                else if( dim == 2 && ps == 0.0f && rs == 0.0f )
                {
                    const std::string sinYaw = Par( ys ), cosYaw = Par( yc );
                    b << "    return " << Src( g, 0, 2, "seed",
                        "fn2::fnma_( y, " + sinYaw + ", x * " + cosYaw + " ), fn2::fma_( x, " + sinYaw + ", y * " + cosYaw + " )" ) << ";\n";
                }
                else if( dim == 2 )
                {
                    b << "    return " << Gen( g, 3 ) << "( seed, x, y, 0.0f, tp );\n";
                }
                else
                {
                    const float xa = yc * pc, xb = yc * ps * rs - ys * rc, xc = yc * ps * rc + ys * rs;
                    const float ya = ys * pc, yb = ys * ps * rs + yc * rc, ycc = ys * ps * rc - yc * rs;
                    const float za = -ps, zb = pc * rs, zc = pc * rc;
                    auto row = [&]( float a, float bb, float c )
                    {
                        return "fn2::fma_( x, " + Par( a ) + ", fn2::fma_( y, " + Par( bb ) + ", z * " + Par( c ) + " ) )";
                    };
                    b << "    return " << Src( g, 0, 3, "seed", row( xa, xb, xc ) + ", " + row( ya, yb, ycc ) + ", " + row( za, zb, zc ) ) << ";\n";
                }
                break;
            }

            case N_DomainAxisScale:
                for( int d = 0; d < dim; d++ ) b << "    " << kAxis[d] << " *= " << Par( s.f[F_K0 + d] ) << ";\n";
                b << "    return " << Src( g, 0, dim, "seed", P ) << ";\n";
                break;

            case N_SeedOffset:
                b << "    return " << Src( g, 0, dim, "fn2::iadd( seed, " + Int( s.i[I_K0] ) + " )", P ) << ";\n";
                break;

            case N_ConvertRGBA8: // [Modifiers.inl:103-127]
            {
                const float lo = s.f[F_K0], hi = s.f[F_K0 + 1];
                b << "    F v = " << Src( g, 0, dim, "seed", P ) << ";\n";
                const std::string loText = Par( lo );
                b << "    v = " << MinMax( false, "v", Par( hi ), ClampSafe( hi, false ) ) << ";\n";
                b << "    v = " << MinMax( true, "v", loText, ClampSafe( lo, true ) ) << ";\n";
                b << "    v -= " << loText << ";\n";
                b << "    v *= " << Par( 255.0f / ( hi - lo ) ) << ";\n";
                b << "    const I byteVal = fn2::cvt_i( v );\n";
                b << "    return fn2::bitsf( (int)0xff000000u | byteVal | fn2::shl( byteVal, 8 ) | fn2::shl( byteVal, 16 ) );\n";
                break;
            }

            case N_GeneratorCache:
                b << "    return " << Src( g, 0, dim, "seed", P ) << ";\n";
                break;

            case N_Fade: // [Blends.inl:164-199]
            {
                b << "    F fade = " << Hyb( g, 0, dim, "seed", P ) << ";\n";
                b << "    const F fadeMin = " << Hyb( g, 1, dim, "seed", P ) << ";\n";
                b << "    const F fadeMax = " << Hyb( g, 2, dim, "seed", P ) << ";\n";
                b << "    const F fadeRange = fadeMax - fadeMin;\n";
                b << "    fade = fn2::div_( fade - fadeMin, fadeRange );\n";
                b << "    fade = fn2::max_x86( 0.0f, fn2::min_x86( 1.0f, fade ) );\n";
                if( s.i[I_K0] == 1 ) b << "    fade = fn2::hermite( fade );\n";
                if( s.i[I_K0] == 2 ) b << "    fade = fn2::quintic( fade );\n";
                b << "    fade = fn2::select( fadeRange == 0.0f, F( 0.5f ), fade );\n";
                b << "    return fn2::lerp( " << Src( g, 0, dim, "seed", P ) << ", " << Src( g, 1, dim, "seed", P ) << ", fade );\n";
                break;
            }

            case N_Remap: // [Modifiers.inl:76-100]
                b << "    const F source = " << Src( g, 0, dim, "seed", P ) << ";\n";
                b << "    const F fromMin = " << Hyb( g, 0, dim, "seed", P ) << ";\n";
                b << "    const F fromMax = " << Hyb( g, 1, dim, "seed", P ) << ";\n";
                b << "    const F toMin = " << Hyb( g, 2, dim, "seed", P ) << ";\n";
                b << "    const F toMax = " << Hyb( g, 3, dim, "seed", P ) << ";\n";
                b << "    F result = toMin + ( fn2::div_( source - fromMin, fromMax - fromMin ) * ( toMax - toMin ) );\n";
                if( s.i[I_K0] == 1 )
                {
                    b << "    result = " << MinMax( true, "result", "toMin", HybClampSafe( g, 2, true ) ) << ";\n";
                    b << "    result = " << MinMax( false, "result", "toMax", HybClampSafe( g, 3, false ) ) << ";\n";
                }
                b << "    return result;\n";
                break;

            case N_Terrace: // [Modifiers.inl:130-182]
            {
                b << "    F source = " << Src( g, 0, dim, "seed", P ) << ";\n";
                b << "    source *= " << Par( s.f[F_K0] ) << ";\n";
                b << "    F rounded = fn2::round_( source );\n";
                const bool dynamic = (bool)s.hybrids[0].node;
                if( dynamic || s.hybrids[0].constant != 0.0f )
                {
                    if( dynamic ) b << "    const F smoothness = " << Hyb( g, 0, dim, "seed", P ) << ";\n";
                    b << "    F diff = rounded - source;\n";
                    b << "    const B diffSign = diff < 0.0f;\n";
                    b << "    diff = fn2::abs_( diff );\n";
                    b << "    diff = 0.5f - diff;\n";
                    if( dynamic ) b << "    diff *= 1.0f + fn2::reciprocal( smoothness );\n";
                    else b << "    diff *= " << Par( s.f[F_K0 + 2] ) << ";\n";
                    b << "    diff = fn2::min_( diff, 0.5f );\n";
                    b << "    diff = fn2::select( diffSign, 0.5f - diff, diff - 0.5f );\n";
                    if( dynamic ) b << "    rounded = fn2::select( smoothness != 0.0f, rounded + diff, rounded );\n";
                    else b << "    rounded += diff;\n";
                }
                b << "    return rounded * " << Par( s.f[F_K0 + 1] ) << ";\n";
                break;
            }

            case N_AddDimension: // [Modifiers.inl:219-237]
                if( dim == 4 ) b << "    return " << Src( g, 0, 4, "seed", P ) << ";\n";
                else b << "    return " << Src( g, 0, dim + 1, "seed", P + ", " + Hyb( g, 0, dim, "seed", P ) ) << ";\n";
                break;

            case N_RemoveDimension: // [Modifiers.inl:239-277]
            {
                int removed = s.i[I_K0];
                if( dim == 2 || removed < 0 || removed >= dim )
                {
                    b << "    return " << Src( g, 0, dim, "seed", P ) << ";\n";
                }
                else
                {
                    std::string kept;
                    for( int d = 0; d < dim; d++ )
                    {
                        if( d == removed ) continue;
                        if( !kept.empty() ) kept += ", ";
                        kept += kAxis[d];
                    }
                    b << "    return " << Src( g, 0, dim - 1, "seed", kept ) << ";\n";
                }
                break;
            }

            case N_DomainRotatePlane: // [Modifiers.inl:352-409]
            {
                bool xz = dim != 2 && s.i[I_K0] == 1; // 2D input always uses the XY preset
                if( dim != 2 && s.i[I_K0] != 0 && s.i[I_K0] != 1 )
                {
                    // an out-of-range rotation type falls through the reference's switch: the coordinates pass unchanged
                    b << "    return " << Src( g, 0, dim, "seed", P ) << ";\n";
                    break;
                }
                if( dim == 2 ) b << "    const F z = 0.0f;\n";
                const char* A = xz ? "z" : "y"; // the axis paired with x
                const char* B = xz ? "y" : "z"; // the axis that is "improved away"
                b << "    const F pair = x + " << A << ";\n";
                b << "    const F s2 = pair * -0.211324865405187f;\n";
                b << "    F nb = " << B << " * 0.577350269189626f;\n";
                b << "    const F nx = x + s2 - nb;\n";
                b << "    const F na = " << A << " + s2 - nb;\n";
                b << "    nb = nb + pair * 0.577350269189626f;\n";
                std::string rotated = xz ? "nx, nb, na" : "nx, na, nb";
                if( dim == 4 ) rotated += ", w";
                b << "    return " << Src( g, 0, dim == 2 ? 3 : dim, "seed", rotated ) << ";\n";
                break;
            }

            default:
                Fail( "unknown node id " + std::to_string( (int)s.nodeId ) );
                b << "    return 0.0f;\n";
            }
            (void)k;
            return b.str();
        }
    };

    // ------------------------------------------------------------------ entry kernels
    // Entry kernels are persistent: the launch fills the GPU once (backend.cpp sizes the grid to the resident blocks) and every
    // thread walks the call's point groups with a grid stride, so the gradient tables are staged and the block min/max is reduced
    // once per thread, not once per group. A group is `ppt` consecutive points of one x-row (4: rows are a multiple of 4 wide and
    // the thread stores one 16-byte vector; 1: any width, a warp stores 32 consecutive floats), evaluated FN2_W at a time
    // (FN2_W = 2: one fpair per two points, see fn2_device_pack.cuh).
    // (the argument struct is __grid_constant__ too: its peer-pointer array is indexed at run time, and without the qualifier the
    // tileable entry copied the whole struct to local memory at the top of every thread)
    const char* const kEntryHead = "( float* __restrict__ out, const FN2_GRID_CONSTANT fn2::";
    std::string EntryOpen( const char* argsType, int ppt, bool wide )
    {
        std::ostringstream k;
        k << "extern \"C\" __global__ void FN2_LAUNCH_BOUNDS fn2_entry" << kEntryHead << argsType
          << " a, unsigned* __restrict__ mm, int seed, const FN2_GRID_CONSTANT fn2::TreeParams params )\n{\n"
          << "    const float* __restrict__ tp = params.v;\n"
          << "    float vmin = FN2_INF, vmax = -FN2_INF;\n"
          << "    bool any = false;\n"
          << "    fn2::stage_tables();\n";
        const char* T = wide ? "unsigned long long" : "unsigned";
        k << "    const " << T << " stride = (" << T << ")gridDim.x * FN2_BLOCK;\n";
        if( ConvergentWalk() )
        {
            k << "    for( " << T << " base = (" << T << ")blockIdx.x * FN2_BLOCK; base < a.groups; base += stride )\n    {\n"
              << "        FN2_GROUP_BARRIER\n"
              << "        const bool live = base + threadIdx.x < a.groups;\n"
              << "        const " << T << " group = live ? base + threadIdx.x : a.groups - 1;\n";
        }
        else
        {
            k << "    for( " << T << " group = (" << T << ")blockIdx.x * FN2_BLOCK + threadIdx.x; group < a.groups; group += stride )\n    {\n"
              << "        const bool live = true;\n";
        }
        k << "        const " << T << " first = group * " << ppt << "u;\n"
          << "        any = true;\n";
        return k.str();
    }

    std::string EntryClose( int ppt )
    {
        std::ostringstream k;
        // (a thread past the end holds the values of the call's last group: real points, harmless in the min/max)
        k << "#pragma unroll\n"
          << "        for( int p = 0; p < " << ppt << "; p++ )\n        {\n"
          << "            vmin = fn2::min_( vmin, v[p] );\n"
          << "            vmax = fn2::max_( vmax, v[p] );\n"
          << "        }\n"
          << "        if( live )\n        {\n"
          << "            fn2::store_points<" << ppt << ">( out, first, v, a.vecStore != 0 );\n"
          << "            for( unsigned k = 0; k < a.mirrors; k++ ) fn2::store_points<" << ppt << ">( a.mirror[k], first, v, a.vecStore != 0 ); // fused all-gather: peer copies\n"
          << "        }\n"
          << "    }\n"
          << "    fn2::reduce_minmax( vmin, vmax, any, mm );\n}\n";
        return k.str();
    }

    // Rows of any width with four points per thread (chunk sizes such as the reference mesh preview's (128+2)^3,
    // tools/NodeEditor/MeshNoisePreview.cpp:446, as uniform grids): a group is four
    // consecutive points of the row-major output, so it may run over the end of a row -- every point carries its own row, decomposed
    // with carries instead of divisions. The last group of a call may be partial (its points are clamped and not stored).
    std::string RaggedGridEntry( int dim, const std::string& root )
    {
        std::ostringstream k;
        k << EntryOpen( "GridArgs", 4, false )
          << "        const unsigned rowLocal = fn2::fast_div( first, a.divX );\n"
          << "        const unsigned ix = first - rowLocal * (unsigned)a.count[0];\n"
          << "        const unsigned row = rowLocal + a.rowBegin;\n";
        if( dim == 2 )
        {
            k << "        const unsigned iy = row;\n";
        }
        else if( dim == 3 )
        {
            k << "        const unsigned iz = fn2::fast_div( row, a.divY );\n"
              << "        const unsigned iy = row - iz * (unsigned)a.count[1];\n";
        }
        else
        {
            k << "        const unsigned plane = fn2::fast_div( row, a.divY );\n"
              << "        const unsigned iy = row - plane * (unsigned)a.count[1];\n"
              << "        const unsigned iw = fn2::fast_div( plane, a.divZ );\n"
              << "        const unsigned iz = plane - iw * (unsigned)a.count[2];\n";
        }
        k << "        const unsigned left = a.total - first; // points of this group inside the call (>= 1)\n"
          << "        float v[4];\n"
          << "#pragma unroll\n"
          << "        for( int p = 0; p < 4; p++ )\n        {\n"
          << "            const unsigned q = (unsigned)p < left ? (unsigned)p : left - 1u; // past the end of the call: the last point again\n"
          << "            unsigned jx = ix + q;\n"
          << "            unsigned carry = jx >= (unsigned)a.count[0] ? 1u : 0u; // rows are at least four points wide: one carry at most\n"
          << "            jx -= carry ? (unsigned)a.count[0] : 0u;\n"
          << "            unsigned jy = iy + carry;\n";
        if( dim >= 3 )
        {
            k << "            carry = jy >= (unsigned)a.count[1] ? 1u : 0u;\n"
              << "            jy -= carry ? (unsigned)a.count[1] : 0u;\n"
              << "            unsigned jz = iz + carry;\n";
        }
        if( dim == 4 )
        {
            k << "            carry = jz >= (unsigned)a.count[2] ? 1u : 0u;\n"
              << "            jz -= carry ? (unsigned)a.count[2] : 0u;\n"
              << "            const unsigned jw = iw + carry;\n";
        }
        // positions exactly as Generator.inl:160-162: fma( float(index), step, offset )
        for( int d = 0; d < dim; d++ )
        {
            k << "            const F p" << kAxis[d] << " = fn2::fma_( fn2::cvt_f( (int)j" << kAxis[d] << " ), a.step[" << d << "], a.off[" << d << "] );\n";
        }
        k << "            v[p] = " << root << "( seed, " << Pos( dim, "p" ) << ", tp );\n"
          << "        }\n"
          << "#pragma unroll\n"
          << "        for( int p = 0; p < 4; p++ )\n        {\n"
          << "            vmin = fn2::min_( vmin, v[p] );\n"
          << "            vmax = fn2::max_( vmax, v[p] );\n"
          << "        }\n"
          << "        if( live )\n        {\n"
          << "            if( left >= 4u )\n            {\n"
          << "                fn2::store_points<4>( out, first, v, a.vecStore != 0 );\n"
          << "                for( unsigned k = 0; k < a.mirrors; k++ ) fn2::store_points<4>( a.mirror[k], first, v, a.vecStore != 0 );\n"
          << "            }\n            else\n            {\n"
          << "                for( unsigned p = 0; p < left; p++ )\n                {\n"
          << "                    const float value = p == 0 ? v[0] : p == 1 ? v[1] : v[2];\n"
          << "                    out[first + p] = value;\n"
          << "                    for( unsigned k = 0; k < a.mirrors; k++ ) a.mirror[k][first + p] = value;\n"
          << "                }\n            }\n"
          << "        }\n"
          << "    }\n"
          << "    fn2::reduce_minmax( vmin, vmax, any, mm );\n}\n";
        return k.str();
    }

    std::string GridEntry( int dim, const std::string& root, int ppt )
    {
        std::ostringstream k;
        k << EntryOpen( "GridArgs", ppt, false )
          << "        const unsigned rowLocal = fn2::fast_div( first, a.divX );\n"
          << "        const unsigned ix = first - rowLocal * (unsigned)a.count[0];\n"
          << "        const unsigned row = rowLocal + a.rowBegin;\n";
        if( dim == 2 )
        {
            k << "        const unsigned iy = row;\n";
        }
        else if( dim == 3 )
        {
            k << "        const unsigned iz = fn2::fast_div( row, a.divY );\n"
              << "        const unsigned iy = row - iz * (unsigned)a.count[1];\n";
        }
        else
        {
            k << "        const unsigned plane = fn2::fast_div( row, a.divY );\n"
              << "        const unsigned iy = row - plane * (unsigned)a.count[1];\n"
              << "        const unsigned iw = fn2::fast_div( plane, a.divZ );\n"
              << "        const unsigned iz = plane - iw * (unsigned)a.count[2];\n";
        }
        // positions exactly as Generator.inl:160-162: fma( float(index), step, offset )
        for( int d = 1; d < dim; d++ )
        {
            k << "        const F p" << kAxis[d] << " = fn2::fma_( fn2::cvt_f( (int)i" << kAxis[d] << " ), a.step[" << d << "], a.off[" << d << "] );\n";
        }
        k << "        float v[" << ppt << "];\n"
          << "#pragma unroll\n"
          << "        for( int p = 0; p < " << ppt << "; p += FN2_W )\n        {\n"
          << "            const F px = fn2::index_positions<F>( ix + p, a.step[0], a.off[0] );\n"
          << "            fn2::spill_lanes( " << root << "( seed, " << Pos( dim, "p" ) << ", tp ), v + p );\n"
          << "        }\n"
          << EntryClose( ppt );
        return k.str();
    }

    std::string TileEntry( const std::string& root, int ppt )
    {
        std::ostringstream k;
        k << EntryOpen( "TileArgs", ppt, false )
          << "        const unsigned rowLocal = fn2::fast_div( first, a.divX );\n"
          << "        const unsigned ix = first - rowLocal * (unsigned)a.xSize;\n"
          << "        const unsigned iy = rowLocal + a.rowBegin;\n"
          // torus mapping [Generator.inl:400-406]
          << "        const float yF = fn2::cvt_f( (int)iy ) * a.yMul;\n"
          << "        const F py = fn2::cos_poly( yF ) * a.yFreq;\n"
          << "        const F pw = fn2::sin_poly( yF ) * a.yFreq;\n"
          << "        float v[" << ppt << "];\n"
          << "#pragma unroll\n"
          << "        for( int p = 0; p < " << ppt << "; p += FN2_W )\n        {\n"
          << "            const F xF = fn2::index_positions<F>( ix + p, a.xMul );\n"
          << "            const F px = fn2::cos_poly( xF ) * a.xFreq;\n"
          << "            const F pz = fn2::sin_poly( xF ) * a.xFreq;\n"
          << "            fn2::spill_lanes( " << root << "( seed, px, py, pz, pw, tp ), v + p );\n"
          << "        }\n"
          << EntryClose( ppt );
        return k.str();
    }

    std::string PositionsEntry( int dim, const std::string& root, int ppt )
    {
        std::ostringstream k;
        k << EntryOpen( "PosArgs", ppt, true )
          << "        const bool full = first + " << ppt << "u <= a.count;\n"
          << "        float in[" << dim << "][" << ppt << "];\n"
          << "        float v[" << ppt << "];\n";
        if( ppt == 4 )
        {
            k << "        if( full && a.vecIO )\n        {\n";
            for( int d = 0; d < dim; d++ )
            {
                k << "            { const float4 t = __ldcs( reinterpret_cast<const float4*>( a.pos[" << d << "] + first ) ); in[" << d << "][0] = t.x; in[" << d << "][1] = t.y; in[" << d << "][2] = t.z; in[" << d << "][3] = t.w; }\n";
            }
            k << "        }\n        else\n";
        }
        k << "        {\n"
          << "#pragma unroll\n"
          << "            for( int p = 0; p < " << ppt << "; p++ )\n            {\n"
          << "                const unsigned long long i = first + p < a.count ? first + p : a.count - 1;\n";
        for( int d = 0; d < dim; d++ ) k << "                in[" << d << "][p] = a.pos[" << d << "][i];\n";
        k << "            }\n        }\n"
          << "#pragma unroll\n"
          << "        for( int p = 0; p < " << ppt << "; p += FN2_W )\n        {\n";
        // pos = offset + array value [Generator.inl:327-330]
        std::string args;
        for( int d = 0; d < dim; d++ )
        {
            if( d ) args += ", ";
            args += "a.off[" + std::to_string( d ) + "] + fn2::lane_values<F>( in[" + std::to_string( d ) + "] + p )";
        }
        k << "            fn2::spill_lanes( " << root << "( seed, " << args << ", tp ), v + p );\n"
          << "        }\n"
          << "#pragma unroll\n"
          << "        for( int p = 0; p < " << ppt << "; p++ )\n        {\n"
          << "            if( first + p < a.count )\n            {\n"
          << "                vmin = fn2::min_( vmin, v[p] );\n"
          << "                vmax = fn2::max_( vmax, v[p] );\n"
          << "            }\n"
          << "        }\n"
          << "        if( !live )\n        {\n        }\n        else if( full )\n        {\n"
          << "            fn2::store_points<" << ppt << ">( out + first, 0u, v, a.vecIO != 0 );\n"
          << "        }\n        else\n        {\n"
          << "            for( int p = 0; p < " << ppt << "; p++ ) if( first + p < a.count ) out[first + p] = v[p];\n"
          << "        }\n"
          << "    }\n"
          << "    fn2::reduce_minmax( vmin, vmax, any, mm );\n}\n";
        return k.str();
    }
}

namespace FastNoise
{
    namespace Impl
    {
        LoweredProgram LowerTree( const Generator* root, int entryKind, int pointsPerThread, int mathMode, bool paramMode )
        {
            LoweredProgram program;
            program.pointsPerThread = pointsPerThread;
            if( !root || !EntryKindValid( entryKind ) || ( pointsPerThread != 1 && pointsPerThread != 4 && pointsPerThread != kRaggedPointsPerThread ) ||
                ( pointsPerThread == kRaggedPointsPerThread && ( entryKind < kGrid2D || entryKind > kGrid4D ) ) )
            {
                program.error = "invalid lowering request";
                return program;
            }

            const int dim = EntryDims( entryKind );
            Emitter emitter;
            emitter.paramMode = paramMode;
            std::string rootFn = emitter.Gen( root, dim );
            if( !emitter.error.empty() )
            {
                program.error = emitter.error;
                return program;
            }

            // Two points per value (packed FP32) or one? Packing removes ~20 % of the issued instructions, but a value then takes
            // a register pair, the kernels end up at 100-250 registers and, with the octave loops unrolled, ~90 KB of code.
            // Measured on B200: for trees evaluated in 3D +3 % on some boxes and -5 % on others with the
            // SAME cubin (the scalar kernel is stable), a wash in 2D, a loss in 4D; the JIT takes up to 6x longer. It is
            // therefore opt-in: FASTNOISE2_B200_PACK=1 enables it for 3D calls, =2 for every dimension.
            const int entryDim = entryKind == kTileable2D ? 4 : ( entryKind >= 10 ? entryKind - 10 : entryKind );
            const char* packEnv = std::getenv( "FASTNOISE2_B200_PACK" );
            const int packMode = packEnv && *packEnv ? std::atoi( packEnv ) : 0;
            const bool pack = packMode >= 2 || ( packMode == 1 && entryDim == 3 );
            const bool ragged = pointsPerThread == kRaggedPointsPerThread;
            const int width = pointsPerThread == 4 && pack ? 2 : 1;

            std::ostringstream src;
            src << "// generated by fastnoise2_b200 -- entry kind " << entryKind << ", " << ( ragged ? 4 : pointsPerThread ) << " point(s) per thread" << ( ragged ? " (rows of any width), " : ", " )
                << ( ( mathMode & 1 ) ? "relaxed" : "strict" ) << " math, " << ( ( mathMode & 2 ) ? "generic" : "table" ) << " gradients\n"
                << "#define FN2_RELAXED " << ( mathMode & 1 ) << "\n"
                << "#define FN2_TABLES " << ( ( mathMode & 2 ) ? 0 : 1 ) << "\n"
                << "#ifndef FN2_BLOCK\n#define FN2_BLOCK 256\n#endif\n"
                // backend.cpp recompiles with a minimum of two blocks per SM when ptxas takes more than 128 registers this way
                << "#ifndef FN2_LAUNCH_BOUNDS\n#define FN2_LAUNCH_BOUNDS __launch_bounds__( FN2_BLOCK )\n#endif\n"
                << "#ifndef FN2_OCTAVE_LOOP\n#define FN2_OCTAVE_LOOP " << ( width == 2 ? "_Pragma( \"unroll\" )" : "" ) << "\n#endif\n"
                // one-body loops stay rolled: unrolling or peeling them would bring the second copy of the source back
                << "#ifndef FN2_ONE_BODY_LOOP\n#define FN2_ONE_BODY_LOOP " << ( EnvFlag( "FASTNOISE2_B200_ONE_BODY", 0 ) == 2 ? "" : "_Pragma( \"unroll 1\" )" ) << "\n#endif\n"
                << "#ifndef FN2_GROUP_BARRIER\n#define FN2_GROUP_BARRIER " << ( ConvergentWalk() && EnvFlag( "FASTNOISE2_B200_GROUP_BARRIER", 0 ) ? "__syncthreads();" : "" ) << "\n#endif\n"
                << "#define FN2_W " << width << "\n"

                << "#include \"fn2_device.cuh\"\n\n"
                << "// float node parameters: kernel arguments (one compiled kernel per tree SHAPE) -- " << emitter.params.size() << " in this tree\n"
                << "namespace fn2 { struct TreeParams { float v[" << std::max<std::size_t>( 1, emitter.params.size() ) << "]; }; }\n"
                << "#define FN2_TP const float* __restrict__ tp\n\n"
                << "// value types of the node functions below: one point (float) or two points (packed FP32) per value\n"
                << ( width == 2 ? "typedef fn2::fpair F;\ntypedef fn2::ipair I;\ntypedef fn2::bpair B;\n\n" : "typedef float F;\ntypedef fn2::i32 I;\ntypedef bool B;\n\n" )
                << emitter.out.str();

            if( entryKind == kTileable2D ) src << TileEntry( rootFn, pointsPerThread );
            else if( entryKind >= kPositions2D ) src << PositionsEntry( dim, rootFn, pointsPerThread );
            else if( pointsPerThread == kRaggedPointsPerThread ) src << RaggedGridEntry( dim, rootFn );
            else src << GridEntry( dim, rootFn, pointsPerThread );

            program.source = src.str();
            program.params = std::move( emitter.params );
            return program;
        }
    }
}
