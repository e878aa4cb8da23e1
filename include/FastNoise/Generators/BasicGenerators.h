// fastnoise2_b200 -- parameter mixins and the basic generator nodes.
// Setter names/meaning follow the reference's BasicGenerators.h; node ids are the registration
// order of src/FastNoise/FastSIMD_Build.inl:86-143. Each layer only records where its members sit
// in the node's reflection tables (kVar/kSrc/kHyb ranges); the values live in Impl::NodeState.
#pragma once
#include "Generator.h"

// boilerplate for a concrete node type
#define FN2_NODE( CLASS, ID ) \
    public: static constexpr std::uint8_t kNodeId = ID; \
    protected: CLASS() : CLASS( kNodeId ) {} \
    friend FASTNOISE_API Generator* Impl::CreateNode( std::uint8_t ); \
    public:

namespace FastNoise
{
    /// Root of the member-index bookkeeping.
    class NodeBase : public Generator
    {
    protected:
        explicit NodeBase( std::uint8_t id ) : Generator( id ) {}
        static constexpr int kVarEnd = 0, kSrcEnd = 0, kHybEnd = 0;
    };

    /// "Feature Scale": positions are multiplied by 1/scale before the noise is sampled.
    class ScalableGenerator : public NodeBase
    {
    public:
        void SetScale( float value ) { SetVar( kVarScale, value ); }

    protected:
        explicit ScalableGenerator( std::uint8_t id ) : NodeBase( id ) {}
        static constexpr int kVarScale = NodeBase::kVarEnd, kVarEnd = kVarScale + 1;
    };

    /// "Seed Offset": added to the seed for this node only, children still see the caller's seed.
    template<typename PARENT>
    class Seeded : public PARENT
    {
    public:
        void SetSeedOffset( int value ) { this->SetVar( kVarSeedOffset, value ); }

    protected:
        explicit Seeded( std::uint8_t id ) : PARENT( id ) {}
        static constexpr int kVarSeedOffset = PARENT::kVarEnd, kVarEnd = kVarSeedOffset + 1;
    };

    /// "Output Min"/"Output Max": affine map of the node's native range.
    template<typename PARENT>
    class VariableRange : public PARENT
    {
    public:
        void SetOutputMin( float value ) { this->SetVar( kVarOutMin, value ); }
        void SetOutputMax( float value ) { this->SetVar( kVarOutMin + 1, value ); }

    protected:
        explicit VariableRange( std::uint8_t id ) : PARENT( id ) {}
        static constexpr int kVarOutMin = PARENT::kVarEnd, kVarEnd = kVarOutMin + 2;
    };

    class Constant : public NodeBase
    {
        FN2_NODE( Constant, 0 )
        void SetValue( float value ) { SetVar( 0, value ); }
    protected:
        explicit Constant( std::uint8_t id ) : NodeBase( id ) {}
    };

    class White : public VariableRange<Seeded<NodeBase>>
    {
        FN2_NODE( White, 1 )
    protected:
        explicit White( std::uint8_t id ) : VariableRange( id ) {}
    };

    class Checkerboard : public VariableRange<ScalableGenerator>
    {
        FN2_NODE( Checkerboard, 2 )
    protected:
        explicit Checkerboard( std::uint8_t id ) : VariableRange( id ) {}
    };

    class SineWave : public VariableRange<ScalableGenerator>
    {
        FN2_NODE( SineWave, 3 )
    protected:
        explicit SineWave( std::uint8_t id ) : VariableRange( id ) {}
    };

    /// sum over dimensions of (position + offset) * multiplier
    class Gradient : public NodeBase
    {
        FN2_NODE( Gradient, 4 )
        template<Dim D> void SetMultiplier( float multiplier ) { SetVar( (int)D, multiplier ); }
        template<Dim D> void SetOffset( float offset ) { SetHyb( (int)D, offset ); }
        template<Dim D> void SetOffset( SmartNodeArg<> gen ) { SetHyb( (int)D, gen.get() ); }
    protected:
        explicit Gradient( std::uint8_t id ) : NodeBase( id ) {}
    };

    class DistanceToPoint : public NodeBase
    {
        FN2_NODE( DistanceToPoint, 5 )
        /// Accepted for source compatibility and ignored: the reference declares it (BasicGenerators.h:266) but never
        /// evaluates, reflects or serialises that source.
        void SetSource( SmartNodeArg<> ) {}
        void SetDistanceFunction( DistanceFunction value ) { SetVar( 0, (int)value ); }
        void SetMinkowskiP( SmartNodeArg<> gen ) { SetHyb( 4, gen.get() ); }
        void SetMinkowskiP( float value ) { SetHyb( 4, value ); }
        void SetPoint( float x, float y, float z = 0, float w = 0 ) { SetHyb( 0, x ); SetHyb( 1, y ); SetHyb( 2, z ); SetHyb( 3, w ); }
        template<Dim D> void SetPoint( float value ) { SetHyb( (int)D, value ); }
        template<Dim D> void SetPoint( SmartNodeArg<> gen ) { SetHyb( (int)D, gen.get() ); }
    protected:
        explicit DistanceToPoint( std::uint8_t id ) : NodeBase( id ) {}
    };
}
