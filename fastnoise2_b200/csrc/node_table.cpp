// fastnoise2_b200 -- the node table: one declarative description of all 47 node types.
//
// From this table come (a) the FastNoise::Metadata records and therefore the member indices of the
// encoded-tree format, (b) default parameter blocks, (c) the single implementation of every parameter
// write. Node order, member order, names, defaults and enum strings restate what the reference
// registers through its MetadataT<> constructor chains (include/FastNoise/Generators/*.h, order of
// src/FastNoise/FastSIMD_Build.inl:86-143; cross-checked in SURVEY.md section 8b). Descriptions are
// this project's own wording.
#include <FastNoise/FastNoise.h>
#include <FastNoise/Metadata.h>

#include <cmath>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <vector>

#include "node_slots.h"
#include "internal.h"

using namespace FastNoise;
using namespace FastNoise::Slots;

namespace
{
    using MV = Metadata::MemberVariable;

    enum class VarOp : std::uint8_t
    {
        Float,     // f[slot] = v
        Int,       // i[slot] = v
        Index4,    // i[slot] = clamp( v, 0, 3 )       (Cellular.h:68,118-119)
        Scale,     // f[F_Scale] = v, f[F_Frequency] = 1/v (BasicGenerators.h SetScale)
        OutMin,    // rangeScale += rangeMin - v; rangeMin = v (BasicGenerators.h SetOutputMin)
        OutMax,    // rangeScale = v - rangeMin
        AngleCS,   // f[slot] = cos v, f[slot+1] = sin v  (Modifiers.h:84-86)
        StepCount, // f[slot] = v, f[slot+1] = 1/v        (Modifiers.h Terrace::SetStepCount)
    };

    enum class HybOp : std::uint8_t { Plain, TerraceSmoothness };

    struct VarDef { MV pub; VarOp op; int slot; std::vector<const char*> enumNames; };
    struct SrcDef { Metadata::MemberNodeLookup pub; bool needsDomainWarp; };
    struct HybDef { Metadata::MemberHybrid pub; HybOp op; };

    struct NodeDef
    {
        Metadata meta;
        std::vector<VarDef> vars;
        std::vector<SrcDef> srcs;
        std::vector<HybDef> hybs;
        std::vector<const char*> groups;
        Impl::NodeState defaults; // only f/i/hybrid constants are meaningful

        // contiguous public views, filled by Seal()
        std::vector<MV> pubVars;
        std::vector<Metadata::MemberNodeLookup> pubSrcs;
        std::vector<Metadata::MemberHybrid> pubHybs;
    };

    // ------------------------------------------------------------------ builder
    struct Builder
    {
        NodeDef& n;

        Builder& Group( const char* g ) { n.groups.push_back( g ); return *this; }
        Builder& Desc( const char* d ) { n.meta.description = d; return *this; }

        MV& AddVar( const char* name, const char* desc, MV::eType type, VarOp op, int slot, int dim = -1 )
        {
            VarDef v;
            v.pub.name = name;
            v.pub.description = desc;
            v.pub.type = type;
            v.pub.dimensionIdx = dim;
            v.pub.valueUiDragSpeed = type == MV::EFloat ? 0.02f : ( type == MV::EInt ? 0.2f : 0.0f );
            v.op = op;
            v.slot = slot;
            n.vars.push_back( v );
            return n.vars.back().pub;
        }

        Builder& VarF( const char* name, const char* desc, float def, int slot, float drag = 0.02f, int dim = -1 )
        {
            MV& v = AddVar( name, desc, MV::EFloat, VarOp::Float, slot, dim );
            v.valueDefault = def; v.valueMin = 0.0f; v.valueMax = 0.0f; v.valueUiDragSpeed = drag;
            n.defaults.f[slot] = def;
            return *this;
        }

        Builder& VarF4( const char* name, const char* desc, float def, int slot, float drag = 0.02f )
        {
            for( int d = 0; d < 4; d++ ) VarF( name, desc, def, slot + d, drag, d );
            return *this;
        }

        Builder& VarI( const char* name, const char* desc, int def, int slot, int minV = 0, int maxV = 0, VarOp op = VarOp::Int )
        {
            MV& v = AddVar( name, desc, MV::EInt, op, slot );
            v.valueDefault = def; v.valueMin = minV; v.valueMax = maxV;
            n.defaults.i[slot] = def;
            return *this;
        }

        Builder& VarEnum( const char* name, const char* desc, int def, int slot, std::initializer_list<const char*> names )
        {
            MV& v = AddVar( name, desc, MV::EEnum, VarOp::Int, slot );
            v.valueDefault = def; v.valueMin = 0; v.valueMax = 0;
            n.vars.back().enumNames.assign( names.begin(), names.end() );
            n.defaults.i[slot] = def;
            return *this;
        }

        Builder& Src( const char* name, const char* desc = "", bool needsDomainWarp = false )
        {
            SrcDef s;
            s.pub.name = name;
            s.pub.description = desc;
            s.needsDomainWarp = needsDomainWarp;
            n.srcs.push_back( s );
            return *this;
        }

        Builder& Hyb( const char* name, const char* desc, float def, float drag = 0.02f, int dim = -1, HybOp op = HybOp::Plain )
        {
            HybDef h;
            h.pub.name = name;
            h.pub.description = desc;
            h.pub.valueDefault = def;
            h.pub.valueUiDragSpeed = drag;
            h.pub.dimensionIdx = dim;
            h.op = op;
            n.defaults.hybrids[n.hybs.size()].constant = def;
            n.hybs.push_back( h );
            return *this;
        }

        Builder& Hyb4( const char* name, const char* desc, float def, float drag = 0.02f )
        {
            for( int d = 0; d < 4; d++ ) Hyb( name, desc, def, drag, d );
            return *this;
        }

        // ---- mixins, in the order the reference's base-class metadata constructors run
        Builder& Scalable()
        {
            MV& v = AddVar( "Feature Scale", "Size of one noise feature in world units (1 / frequency)", MV::EFloat, VarOp::Scale, F_Scale );
            v.valueDefault = 100.0f; v.valueMin = 0.0f; v.valueMax = 0.0f; v.valueUiDragSpeed = 0.25f;
            n.defaults.f[F_Scale] = 100.0f;
            n.defaults.f[F_Frequency] = 1.0f / 100;
            return *this;
        }

        Builder& Seeded()
        {
            return VarI( "Seed Offset", "Added to the seed for this node only; child nodes still receive the caller's seed", 0, I_SeedOffset );
        }

        Builder& Range()
        {
            MV& lo = AddVar( "Output Min", "Lower bound of the output range", MV::EFloat, VarOp::OutMin, F_RangeMin );
            lo.valueDefault = -1.0f; lo.valueMin = 0.0f; lo.valueMax = 0.0f;
            MV& hi = AddVar( "Output Max", "Upper bound of the output range", MV::EFloat, VarOp::OutMax, F_RangeMin );
            hi.valueDefault = 1.0f; hi.valueMin = 0.0f; hi.valueMax = 0.0f;
            n.defaults.f[F_RangeMin] = -1.0f;
            n.defaults.f[F_RangeScale] = 2.0f;
            return *this;
        }

        Builder& DistanceFunctionVar( const char* desc, int metadataDefault )
        {
            VarEnum( "Distance Function", desc, metadataDefault, I_K0,
                { "Euclidean", "Euclidean Squared", "Manhattan", "Hybrid", "Max Axis", "Minkowski" } );
            // every owner of this member starts at EuclideanSquared, whatever the table advertises
            // (DistanceToPoint: BasicGenerators.h:289 vs :305)
            n.defaults.i[I_K0] = (int)DistanceFunction::EuclideanSquared;
            return *this;
        }

        Builder& CellularCommon()
        {
            Group( "Coherent Noise" );
            DistanceFunctionVar( "Metric used to find the closest cells; Hybrid = Euclidean Squared + Manhattan", (int)DistanceFunction::EuclideanSquared );
            Hyb( "Minkowski P", "Exponent of the Minkowski metric only: 1 = Manhattan, 2 = Euclidean", 1.5f );
            Hyb( "Grid Jitter", "How far cell points move off the regular grid; 0 = grid, above 1 shows artifacts", 1.0f );
            Hyb( "Size Jitter", "Random per-cell multiplier on the measured distance, varies cell size", 0.0f );
            return *this;
        }

        Builder& FractalCommon( const char* sourceName, const char* sourceDesc, bool warpSource )
        {
            Src( sourceName, sourceDesc, warpSource );
            Hyb( "Gain", "Amplitude multiplier applied from one octave to the next", 0.5f );
            Hyb( "Weighted Strength", "How much an octave's output scales the amplitude of the following octaves", 0.0f );
            VarI( "Octaves", "Number of layers summed; each one is sampled at a higher frequency", 3, I_K0, 2, 16 );
            VarF( "Lacunarity", "Frequency multiplier from one octave to the next", 2.0f, F_K0 );
            return *this;
        }

        Builder& DomainWarpCommon()
        {
            Scalable().Seeded();
            Group( "Domain Warp" );
            Src( "Source" );
            Hyb( "Warp Amplitude", "Largest distance a position can be displaced", 50.0f, 0.1f );
            VarF4( "Amplitude Scaling", "", 1.0f, F_K0 );
            return *this;
        }

        Builder& OperatorSourceLHS( const char* group )
        {
            Group( group );
            Src( "LHS" );
            Hyb( "RHS", "", 0.0f );
            return *this;
        }

        Builder& OperatorHybridLHS( const char* group )
        {
            Group( group );
            Hyb( "LHS", "", 0.0f );
            Hyb( "RHS", "", 0.0f );
            return *this;
        }
    };

    // ------------------------------------------------------------------ the table
    struct Table
    {
        std::deque<NodeDef> nodes; // deque: stable addresses
        std::vector<const Metadata*> all;
        Metadata::Vector<const Metadata*> allView;

        Builder Node( const char* name )
        {
            nodes.emplace_back();
            NodeDef& n = nodes.back();
            n.meta.name = name;
            n.meta.id = (Metadata::node_id)( nodes.size() - 1 );
            n.defaults.nodeId = n.meta.id;
            return Builder { n };
        }

        Table()
        {
            // 0..5 basic generators
            Node( "Constant" ).Group( "Basic Generators" ).VarF( "Value", "The output", 1.0f, F_K0 )
                .Desc( "Same value at every position" );
            Node( "White" ).Seeded().Range().Group( "Basic Generators" )
                .Desc( "Uncorrelated random value per distinct position" );
            Node( "Checkerboard" ).Scalable().Range().Group( "Basic Generators" )
                .Desc( "Alternating cells of Feature Scale size in every dimension" );
            Node( "SineWave" ).Scalable().Range().Group( "Basic Generators" )
                .Desc( "Product of one sine per dimension" );
            Node( "Gradient" ).Group( "Basic Generators" )
                .VarF4( "Multiplier", "See node description", 0.0f, F_K0, 0.001f )
                .Hyb4( "Offset", "See node description", 0.0f, 0.25f )
                .Desc( "Sum over dimensions of (position + offset) * multiplier" );
            Node( "DistanceToPoint" ).Group( "Basic Generators" )
                .DistanceFunctionVar( "", (int)DistanceFunction::Euclidean )
                .Hyb4( "Point", "Target point, in the current (possibly warped) domain", 0.0f )
                .Hyb( "Minkowski P", "Exponent of the Minkowski metric only: 1 = Manhattan, 2 = Euclidean", 1.5f )
                .Desc( "Distance from the sample position to a point" );

            // 6..9 coherent noise
            Node( "Simplex" ).Scalable().Seeded().Range().Group( "Coherent Noise" )
                .Desc( "Gradient noise on a simplex lattice" );
            Node( "SuperSimplex" ).Scalable().Seeded().Range().Group( "Coherent Noise" )
                .Desc( "Smoother, costlier gradient noise on a simplex lattice (OpenSimplex2S family)" );
            Node( "Perlin" ).Scalable().Seeded().Range().Group( "Coherent Noise" )
                .Desc( "Gradient noise on a cubic lattice" );
            Node( "Value" ).Scalable().Seeded().Range().Group( "Coherent Noise" )
                .Desc( "Interpolated random values on a cubic lattice" );

            // 10..12 cellular
            Node( "CellularValue" ).Scalable().Seeded().Range().CellularCommon()
                .VarI( "Value Index", "Which of the closest cells supplies the value (0 = closest)", 0, I_K0 + 1, 0, 3, VarOp::Index4 )
                .Desc( "Random value of the Nth closest cell" );
            Node( "CellularDistance" ).Scalable().Seeded().Range().CellularCommon()
                .VarI( "Distance Index 0", "Nth closest cell", 0, I_K0 + 1, 0, 3, VarOp::Index4 )
                .VarI( "Distance Index 1", "Nth closest cell", 1, I_K0 + 2, 0, 3, VarOp::Index4 )
                .VarEnum( "Return Type", "How the two distances are combined", 0, I_K0 + 3, { "Index0", "Index0Add1", "Index0Sub1", "Index0Mul1", "Index0Div1" } )
                .Desc( "Distance to the Nth closest cell, optionally combined with a second one" );
            Node( "CellularLookup" ).Scalable().Seeded().CellularCommon()
                .Src( "Lookup", "Sampled at the centre of the closest cell" )
                .Desc( "Value of the lookup source at the closest cell's centre" );

            // 13..15
            Node( "FractalFBm" ).Group( "Fractal" ).FractalCommon( "Source", "", false )
                .Desc( "Fractional Brownian motion: sum of octaves with decreasing amplitude" );
            Node( "PingPong" ).Group( "Modifiers" ).Src( "Source" )
                .Hyb( "Ping Pong Strength", "Multiplier applied before the value is folded back into range", 2.0f )
                .Desc( "Scales the source and reflects it back and forth between the range ends" );
            Node( "FractalRidged" ).Group( "Fractal" ).FractalCommon( "Source", "", false )
                .Desc( "Sum of octaves of 1 - 2|noise|: sharp ridges" );

            // 16..18 domain warps
            Node( "DomainWarpSimplex" ).DomainWarpCommon()
                .VarEnum( "Vectorization Scheme", "How the lattice noise is turned into a displacement vector", 0, I_K0, { "Orthogonal Gradient Matrix", "Gradient Outer Product" } )
                .Desc( "Displaces positions with simplex-lattice vector noise before sampling Source" );
            Node( "DomainWarpSuperSimplex" ).DomainWarpCommon()
                .VarEnum( "Vectorization Scheme", "How the lattice noise is turned into a displacement vector", 0, I_K0, { "Orthogonal Gradient Matrix", "Gradient Outer Product" } )
                .Desc( "Smoother, costlier variant of DomainWarpSimplex" );
            Node( "DomainWarpGradient" ).DomainWarpCommon()
                .Desc( "Displaces positions with cubic-lattice vector noise before sampling Source" );

            // 19..20 warp fractals: groups are { "Domain Warp", "Fractal" } (DomainWarpFractal.h:33-34)
            Node( "DomainWarpFractalProgressive" ).Group( "Domain Warp" ).Group( "Fractal" )
                .FractalCommon( "Domain Warp Source", "Warp algorithm applied once per octave", true )
                .Desc( "Each octave warps the position produced by the previous octave" );
            Node( "DomainWarpFractalIndependent" ).Group( "Domain Warp" ).Group( "Fractal" )
                .FractalCommon( "Domain Warp Source", "Warp algorithm applied once per octave", true )
                .Desc( "Every octave warps the original position; displacements accumulate" );

            // 21..24 operators
            Node( "Add" ).OperatorSourceLHS( "Operators" );
            Node( "Subtract" ).OperatorHybridLHS( "Operators" );
            Node( "Multiply" ).OperatorSourceLHS( "Operators" );
            Node( "Divide" ).OperatorHybridLHS( "Operators" );

            // 25..32
            Node( "Abs" ).Group( "Modifiers" ).Src( "Source" ).Desc( "Absolute value of the source" );
            Node( "Min" ).OperatorSourceLHS( "Blends" );
            Node( "Max" ).OperatorSourceLHS( "Blends" );
            Node( "MinSmooth" ).OperatorSourceLHS( "Blends" ).Hyb( "Smoothness", "", 0.1f );
            Node( "MaxSmooth" ).OperatorSourceLHS( "Blends" ).Hyb( "Smoothness", "", 0.1f );
            Node( "SignedSquareRoot" ).Group( "Modifiers" ).Src( "Source" ).Desc( "sqrt(|source|) with the source's sign" );
            Node( "PowFloat" ).Group( "Blends" ).Hyb( "Value", "", 2.0f ).Hyb( "Pow", "", 2.0f );
            Node( "PowInt" ).Group( "Blends" ).Src( "Value" ).VarI( "Pow", "", 2, I_K0, 2, 0 );

            // 33..36 domain modifiers
            Node( "DomainScale" ).Group( "Domain Modifiers" ).Src( "Source" ).VarF( "Scaling", "", 1.0f, F_K0 )
                .Desc( "Multiplies every coordinate by one factor before sampling the source" );
            Node( "DomainOffset" ).Group( "Domain Modifiers" ).Src( "Source" ).Hyb4( "Offset", "", 0.0f, 0.25f )
                .Desc( "Adds a per-axis offset to the position before sampling the source" );
            {
                Builder b = Node( "DomainRotate" );
                b.Group( "Domain Modifiers" ).Src( "Source" );
                const char* names[3] = { "Yaw", "Pitch", "Roll" };
                for( int a = 0; a < 3; a++ )
                {
                    MV& v = b.AddVar( names[a], "", MV::EFloat, VarOp::AngleCS, F_K0 + a * 2 );
                    v.valueDefault = 0.0f; v.valueMin = 0.0f; v.valueMax = 0.0f;
                    b.n.defaults.f[F_K0 + a * 2] = 1.0f; // cos 0
                    b.n.defaults.f[F_K0 + a * 2 + 1] = 0.0f;
                }
                b.Desc( "Rotates the position about the origin (2D: in-plane unless pitch/roll are set, 3D: yaw/pitch/roll, 4D: unchanged)" );
            }
            Node( "DomainAxisScale" ).Group( "Domain Modifiers" ).Src( "Source" ).VarF4( "Scaling", "", 1.0f, F_K0 )
                .Desc( "Multiplies each coordinate by its own factor before sampling the source" );

            // 37..39
            Node( "SeedOffset" ).Group( "Modifiers" ).Src( "Source" ).VarI( "Seed Offset", "", 1, I_K0 )
                .Desc( "Adds a constant to the seed passed to the source" );
            Node( "ConvertRGBA8" ).Group( "Modifiers" ).Src( "Source" ).VarF( "Min", "", -1.0f, F_K0 ).VarF( "Max", "", 1.0f, F_K0 + 1 )
                .Desc( "Clamps to [Min, Max], scales to 0..255 and packs a grey opaque RGBA8 pixel into the float's bits" );
            Node( "GeneratorCache" ).Group( "Modifiers" ).Src( "Source" )
                .Desc( "Marks a shared sub-tree; evaluation result is identical to the source" );

            // 40..44
            Node( "Fade" ).Group( "Blends" ).Src( "A", "From" ).Src( "B", "To" )
                .Hyb( "Fade", "", 0.0f ).Hyb( "Fade Min", "", -1.0f ).Hyb( "Fade Max", "", 1.0f )
                .VarEnum( "Interpolation", "Easing curve", 0, I_K0, { "Linear", "Hermite", "Quintic" } );
            Node( "Remap" ).Group( "Modifiers" ).Src( "Source" )
                .Hyb( "From Min", "", -1.0f ).Hyb( "From Max", "", 1.0f ).Hyb( "To Min", "", 0.0f ).Hyb( "To Max", "", 1.0f )
                .VarEnum( "Clamp Output", "Clamp the result between To Min and To Max", 0, I_K0, { "False", "True" } )
                .Desc( "Linear map of the source from one range onto another" );
            {
                Builder b = Node( "Terrace" );
                b.Group( "Modifiers" ).Src( "Source" );
                MV& v = b.AddVar( "Step Count", "More steps = lower steps", MV::EFloat, VarOp::StepCount, F_K0 );
                v.valueDefault = 1.0f; v.valueMin = 0.0f; v.valueMax = 0.0f;
                b.n.defaults.f[F_K0] = 1.0f;     // step count
                b.n.defaults.f[F_K0 + 1] = 1.0f; // its reciprocal
                b.n.defaults.f[F_K0 + 2] = 0.0f; // 1 + 1/smoothness, only meaningful for a non-zero constant smoothness
                b.Hyb( "Smoothness", "Width of the blend between neighbouring steps", 0.0f, 0.02f, -1, HybOp::TerraceSmoothness );
                b.Desc( "Quantises the source into steps" );
            }
            Node( "AddDimension" ).Group( "Domain Modifiers" ).Src( "Source" )
                .Hyb( "New Dimension Position", "Coordinate given to the appended axis", 0.0f )
                .Desc( "Samples the source with one more coordinate appended" );
            Node( "RemoveDimension" ).Group( "Domain Modifiers" ).Src( "Source" )
                .VarEnum( "Remove Dimension", "", (int)Dim::Y, I_K0, { "X", "Y", "Z", "W" } )
                .Desc( "Samples the source with one coordinate dropped" );

            // 45..46
            Node( "Modulus" ).OperatorHybridLHS( "Operators" );
            Node( "DomainRotatePlane" ).Group( "Domain Modifiers" ).Src( "Source" )
                .VarEnum( "Rotation Type", "", 0, I_K0, { "Improve XY Planes", "Improve XZ Planes" } )
                .Desc( "Fixed rotation that hides lattice alignment in the chosen plane of 3D noise" );

            Seal();
        }

        void Seal()
        {
            for( NodeDef& n : nodes )
            {
                const Metadata::node_id id = n.meta.id;
                for( std::size_t k = 0; k < n.vars.size(); k++ )
                {
                    VarDef& v = n.vars[k];
                    v.pub.enumNames = Metadata::Vector<const char*>( v.enumNames.data(), v.enumNames.size() );
                    v.pub.setFunc = [id, k]( Generator* g, MV::ValueUnion value )
                    {
                        if( !g || g->State().nodeId != id ) return false;
                        Impl::SetVariable( g, (int)k, value.i );
                        return true;
                    };
                    n.pubVars.push_back( v.pub );
                }
                for( std::size_t k = 0; k < n.srcs.size(); k++ )
                {
                    n.srcs[k].pub.setFunc = [id, k]( Generator* g, SmartNodeArg<> s )
                    {
                        return g && g->State().nodeId == id && Impl::SetLookup( g, (int)k, s.get() );
                    };
                    n.pubSrcs.push_back( n.srcs[k].pub );
                }
                for( std::size_t k = 0; k < n.hybs.size(); k++ )
                {
                    n.hybs[k].pub.setValueFunc = [id, k]( Generator* g, float value )
                    {
                        if( !g || g->State().nodeId != id ) return false;
                        Impl::SetHybridValue( g, (int)k, value );
                        return true;
                    };
                    n.hybs[k].pub.setNodeFunc = [id, k]( Generator* g, SmartNodeArg<> s )
                    {
                        return g && g->State().nodeId == id && Impl::SetHybridNode( g, (int)k, s.get() );
                    };
                    n.pubHybs.push_back( n.hybs[k].pub );
                }
                n.meta.memberVariables = Metadata::Vector<MV>( n.pubVars.data(), n.pubVars.size() );
                n.meta.memberNodeLookups = Metadata::Vector<Metadata::MemberNodeLookup>( n.pubSrcs.data(), n.pubSrcs.size() );
                n.meta.memberHybrids = Metadata::Vector<Metadata::MemberHybrid>( n.pubHybs.data(), n.pubHybs.size() );
                n.meta.groups = Metadata::Vector<const char*>( n.groups.data(), n.groups.size() );
                all.push_back( &n.meta );
            }
            allView = Metadata::Vector<const Metadata*>( all.data(), all.size() );
        }
    };

    const Table& GetTable()
    {
        static const Table table;
        return table;
    }

    bool IsDomainWarp( std::uint8_t id )
    {
        return id == N_DomainWarpSimplex || id == N_DomainWarpSuperSimplex || id == N_DomainWarpGradient;
    }
}

// ---------------------------------------------------------------------- Metadata statics
const Metadata::Vector<const Metadata*>& Metadata::GetAll()
{
    return GetTable().allView;
}

const Metadata* Metadata::GetFromId( node_id nodeId )
{
    const Table& t = GetTable();
    return nodeId < t.all.size() ? t.all[nodeId] : nullptr;
}

SmartNode<> Metadata::CreateNode( FastSIMD::FeatureSet maxFeatureSet ) const
{
    if( maxFeatureSet == FastSIMD::FeatureSet::Invalid )
    {
        return nullptr;
    }
    return SmartNode<>::FromRaw( Impl::CreateNode( id ) );
}

NodeData::NodeData( const Metadata* data ) : metadata( data )
{
    if( !metadata ) return;
    for( const Metadata::MemberVariable& v : metadata->memberVariables ) variables.push_back( v.valueDefault );
    nodeLookups.assign( metadata->memberNodeLookups.size(), nullptr );
    for( const Metadata::MemberHybrid& h : metadata->memberHybrids ) hybrids.emplace_back( nullptr, h.valueDefault );
}

// ---------------------------------------------------------------------- node instances
namespace FastNoise
{
    namespace Impl
    {
        static std::atomic<std::uint64_t> gTreeStamp { 1 };

        std::uint64_t TreeStamp() { return gTreeStamp.load( std::memory_order_acquire ); }
        static void Touch() { gTreeStamp.fetch_add( 1, std::memory_order_acq_rel ); }

        bool LookupNeedsDomainWarp( std::uint8_t nodeId, int lookupIndex )
        {
            const NodeDef& def = GetTable().nodes[nodeId];
            return lookupIndex < (int)def.srcs.size() && def.srcs[lookupIndex].needsDomainWarp;
        }

        Generator* CreateNode( std::uint8_t nodeId )
        {
#define FN2_CASE( CLASS ) case CLASS::kNodeId: return new CLASS();
            switch( nodeId )
            {
                FN2_CASE( Constant ) FN2_CASE( White ) FN2_CASE( Checkerboard ) FN2_CASE( SineWave ) FN2_CASE( Gradient ) FN2_CASE( DistanceToPoint )
                FN2_CASE( Simplex ) FN2_CASE( SuperSimplex ) FN2_CASE( Perlin ) FN2_CASE( Value )
                FN2_CASE( CellularValue ) FN2_CASE( CellularDistance ) FN2_CASE( CellularLookup )
                FN2_CASE( FractalFBm ) FN2_CASE( PingPong ) FN2_CASE( FractalRidged )
                FN2_CASE( DomainWarpSimplex ) FN2_CASE( DomainWarpSuperSimplex ) FN2_CASE( DomainWarpGradient )
                FN2_CASE( DomainWarpFractalProgressive ) FN2_CASE( DomainWarpFractalIndependent )
                FN2_CASE( Add ) FN2_CASE( Subtract ) FN2_CASE( Multiply ) FN2_CASE( Divide )
                FN2_CASE( Abs ) FN2_CASE( Min ) FN2_CASE( Max ) FN2_CASE( MinSmooth ) FN2_CASE( MaxSmooth )
                FN2_CASE( SignedSquareRoot ) FN2_CASE( PowFloat ) FN2_CASE( PowInt )
                FN2_CASE( DomainScale ) FN2_CASE( DomainOffset ) FN2_CASE( DomainRotate ) FN2_CASE( DomainAxisScale )
                FN2_CASE( SeedOffset ) FN2_CASE( ConvertRGBA8 ) FN2_CASE( GeneratorCache )
                FN2_CASE( Fade ) FN2_CASE( Remap ) FN2_CASE( Terrace ) FN2_CASE( AddDimension ) FN2_CASE( RemoveDimension )
                FN2_CASE( Modulus ) FN2_CASE( DomainRotatePlane )
            default: return nullptr;
            }
#undef FN2_CASE
        }

        void SetVariable( Generator* node, int variableIndex, std::int32_t rawBits )
        {
            if( !node ) return;
            const NodeDef& def = GetTable().nodes[node->mState.nodeId];
            if( variableIndex < 0 || variableIndex >= (int)def.vars.size() ) return;
            const VarDef& v = def.vars[variableIndex];
            NodeState& s = node->mState;
            float asFloat;
            std::memcpy( &asFloat, &rawBits, 4 );

            switch( v.op )
            {
            case VarOp::Float: s.f[v.slot] = asFloat; break;
            case VarOp::Int: s.i[v.slot] = rawBits; break;
            case VarOp::Index4: s.i[v.slot] = std::min( std::max( rawBits, 0 ), 3 ); break;
            case VarOp::Scale:
                s.f[F_Scale] = asFloat;
                s.f[F_Frequency] = 1.0f / asFloat;
                break;
            case VarOp::OutMin:
                s.f[F_RangeScale] += s.f[F_RangeMin] - asFloat;
                s.f[F_RangeMin] = asFloat;
                break;
            case VarOp::OutMax:
                s.f[F_RangeScale] = asFloat - s.f[F_RangeMin];
                break;
            case VarOp::AngleCS:
                s.f[v.slot] = std::cos( asFloat );
                s.f[v.slot + 1] = std::sin( asFloat );
                break;
            case VarOp::StepCount:
                s.f[v.slot] = asFloat;
                s.f[v.slot + 1] = 1 / asFloat;
                break;
            }
            Touch();
        }

        bool SetLookup( Generator* node, int lookupIndex, const Generator* source )
        {
            if( !node ) return false;
            const NodeDef& def = GetTable().nodes[node->mState.nodeId];
            if( lookupIndex < 0 || lookupIndex >= (int)def.srcs.size() ) return false;
            if( source && def.srcs[lookupIndex].needsDomainWarp && !IsDomainWarp( source->State().nodeId ) ) return false;
            node->mState.lookups[lookupIndex] = SmartNode<const Generator>::FromRaw( source );
            Touch();
            return true;
        }

        bool SetHybridNode( Generator* node, int hybridIndex, const Generator* source )
        {
            if( !node ) return false;
            const NodeDef& def = GetTable().nodes[node->mState.nodeId];
            if( hybridIndex < 0 || hybridIndex >= (int)def.hybs.size() ) return false;
            node->mState.hybrids[hybridIndex].node = SmartNode<const Generator>::FromRaw( source ); // constant is kept
            Touch();
            return true;
        }

        void SetHybridValue( Generator* node, int hybridIndex, float value )
        {
            if( !node ) return;
            const NodeDef& def = GetTable().nodes[node->mState.nodeId];
            if( hybridIndex < 0 || hybridIndex >= (int)def.hybs.size() ) return;
            NodeState::Hybrid& h = node->mState.hybrids[hybridIndex];
            h.node = nullptr; // assigning a float replaces the whole hybrid, as `mGain = value` does in the reference
            h.constant = value;
            if( def.hybs[hybridIndex].op == HybOp::TerraceSmoothness && value != 0.0f )
            {
                node->mState.f[F_K0 + 2] = 1 + 1 / value;
            }
            Touch();
        }

        const NodeState& DefaultState( std::uint8_t nodeId )
        {
            return GetTable().nodes[nodeId].defaults;
        }
    }

    namespace Internal
    {
        std::int32_t NodeRefAdd( const Generator* node, std::int32_t delta )
        {
            if( delta == 0 ) return (std::int32_t)node->mRefs.load( std::memory_order_relaxed );
            if( delta > 0 ) return (std::int32_t)node->mRefs.fetch_add( (std::uint32_t)delta, std::memory_order_relaxed );
            return (std::int32_t)node->mRefs.fetch_sub( (std::uint32_t)-delta, std::memory_order_acq_rel );
        }

        void NodeDestroy( const Generator* node )
        {
            delete node;
        }
    }

    void SmartNodeManager::SetMemoryPoolSize( std::uint32_t ) {}

    Generator::Generator( std::uint8_t nodeId )
    {
        const Impl::NodeState& d = Impl::DefaultState( nodeId );
        mState.nodeId = nodeId;
        std::memcpy( mState.f, d.f, sizeof( mState.f ) );
        std::memcpy( mState.i, d.i, sizeof( mState.i ) );
        for( int k = 0; k < Impl::NodeState::kHybrids; k++ ) mState.hybrids[k].constant = d.hybrids[k].constant;
    }

    Generator::~Generator()
    {
        Impl::ReleaseCompiled( mCompiled.exchange( nullptr ) );
    }

    FastSIMD::FeatureSet Generator::GetActiveFeatureSet() const
    {
        return FastSIMD::FeatureSet::B200_SM100A;
    }

    const Metadata& Generator::GetMetadata() const
    {
        return *Metadata::GetFromId( mState.nodeId );
    }
}
