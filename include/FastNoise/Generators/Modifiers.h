// fastnoise2_b200 -- domain and value modifier node types (reference: Generators/Modifiers.h).
// Member indices used below are the reflection-table positions listed in SURVEY.md section 8(b).
#pragma once
#include "BasicGenerators.h"

namespace FastNoise
{
    /// Common shape: one required "Source" input at lookup 0.
    class SingleSourceNode : public NodeBase
    {
    public:
        void SetSource( SmartNodeArg<> gen ) { SetSrc( 0, gen.get() ); }
    protected:
        explicit SingleSourceNode( std::uint8_t id ) : NodeBase( id ) {}
    };

    class DomainScale : public SingleSourceNode
    {
        FN2_NODE( DomainScale, 33 )
        void SetScaling( float value ) { SetVar( 0, value ); }
    protected:
        explicit DomainScale( std::uint8_t id ) : SingleSourceNode( id ) {}
    };

    class DomainOffset : public SingleSourceNode
    {
        FN2_NODE( DomainOffset, 34 )
        template<Dim D> void SetOffset( float value ) { SetHyb( (int)D, value ); }
        template<Dim D> void SetOffset( SmartNodeArg<> gen ) { SetHyb( (int)D, gen.get() ); }
    protected:
        explicit DomainOffset( std::uint8_t id ) : SingleSourceNode( id ) {}
    };

    /// Angles in radians. 2D input with pitch = roll = 0 rotates in the plane, otherwise the source is sampled in 3D.
    class DomainRotate : public SingleSourceNode
    {
        FN2_NODE( DomainRotate, 35 )
        void SetYaw( float value ) { SetVar( 0, value ); }
        void SetPitch( float value ) { SetVar( 1, value ); }
        void SetRoll( float value ) { SetVar( 2, value ); }
    protected:
        explicit DomainRotate( std::uint8_t id ) : SingleSourceNode( id ) {}
    };

    class DomainAxisScale : public SingleSourceNode
    {
        FN2_NODE( DomainAxisScale, 36 )
        template<Dim D> void SetScaling( float value ) { SetVar( (int)D, value ); }
    protected:
        explicit DomainAxisScale( std::uint8_t id ) : SingleSourceNode( id ) {}
    };

    class SeedOffset : public SingleSourceNode
    {
        FN2_NODE( SeedOffset, 37 )
        void SetOffset( int value ) { SetVar( 0, value ); }
    protected:
        explicit SeedOffset( std::uint8_t id ) : SingleSourceNode( id ) {}
    };

    /// Packs the clamped source into a grey RGBA8 pixel carried in the float's bits.
    class ConvertRGBA8 : public SingleSourceNode
    {
        FN2_NODE( ConvertRGBA8, 38 )
        void SetMinMax( float min, float max ) { SetVar( 0, min ); SetVar( 1, max ); }
    protected:
        explicit ConvertRGBA8( std::uint8_t id ) : SingleSourceNode( id ) {}
    };

    /// On this backend a pure pass-through: a GPU thread owns its point, there is nothing to cache across lanes.
    class GeneratorCache : public SingleSourceNode
    {
        FN2_NODE( GeneratorCache, 39 )
    protected:
        explicit GeneratorCache( std::uint8_t id ) : SingleSourceNode( id ) {}
    };

    class Remap : public SingleSourceNode
    {
        FN2_NODE( Remap, 41 )
        void SetFromMin( float value ) { SetHyb( 0, value ); }
        void SetFromMin( SmartNodeArg<> gen ) { SetHyb( 0, gen.get() ); }
        void SetFromMax( float value ) { SetHyb( 1, value ); }
        void SetFromMax( SmartNodeArg<> gen ) { SetHyb( 1, gen.get() ); }
        void SetToMin( float value ) { SetHyb( 2, value ); }
        void SetToMin( SmartNodeArg<> gen ) { SetHyb( 2, gen.get() ); }
        void SetToMax( float value ) { SetHyb( 3, value ); }
        void SetToMax( SmartNodeArg<> gen ) { SetHyb( 3, gen.get() ); }
        void SetClampOutput( Boolean value ) { SetVar( 0, (int)value ); }
        void SetClampOutput( bool value ) { SetVar( 0, value ? 1 : 0 ); }
    protected:
        explicit Remap( std::uint8_t id ) : SingleSourceNode( id ) {}
    };

    class Terrace : public SingleSourceNode
    {
        FN2_NODE( Terrace, 42 )
        void SetStepCount( float stepCount ) { SetVar( 0, stepCount ); }
        void SetSmoothness( float smoothness ) { SetHyb( 0, smoothness ); }
        void SetSmoothness( SmartNodeArg<> gen ) { SetHyb( 0, gen.get() ); }
    protected:
        explicit Terrace( std::uint8_t id ) : SingleSourceNode( id ) {}
    };

    class PingPong : public SingleSourceNode
    {
        FN2_NODE( PingPong, 14 )
        void SetPingPongStrength( float value ) { SetHyb( 0, value ); }
        void SetPingPongStrength( SmartNodeArg<> gen ) { SetHyb( 0, gen.get() ); }
    protected:
        explicit PingPong( std::uint8_t id ) : SingleSourceNode( id ) {}
    };

    /// Appends one coordinate (constant or generated) and samples the source in the higher dimension.
    class AddDimension : public SingleSourceNode
    {
        FN2_NODE( AddDimension, 43 )
        void SetNewDimensionPosition( float value ) { SetHyb( 0, value ); }
        void SetNewDimensionPosition( SmartNodeArg<> gen ) { SetHyb( 0, gen.get() ); }
    protected:
        explicit AddDimension( std::uint8_t id ) : SingleSourceNode( id ) {}
    };

    class RemoveDimension : public SingleSourceNode
    {
        FN2_NODE( RemoveDimension, 44 )
        void SetRemoveDimension( Dim dimension ) { SetVar( 0, (int)dimension ); }
    protected:
        explicit RemoveDimension( std::uint8_t id ) : SingleSourceNode( id ) {}
    };

    class SignedSquareRoot : public SingleSourceNode
    {
        FN2_NODE( SignedSquareRoot, 30 )
    protected:
        explicit SignedSquareRoot( std::uint8_t id ) : SingleSourceNode( id ) {}
    };

    class Abs : public SingleSourceNode
    {
        FN2_NODE( Abs, 25 )
    protected:
        explicit Abs( std::uint8_t id ) : SingleSourceNode( id ) {}
    };

    enum class PlaneRotationType { ImproveXYPlanes, ImproveXZPlanes };

    class DomainRotatePlane : public SingleSourceNode
    {
        FN2_NODE( DomainRotatePlane, 46 )
        void SetRotationType( PlaneRotationType type ) { SetVar( 0, (int)type ); }
    protected:
        explicit DomainRotatePlane( std::uint8_t id ) : SingleSourceNode( id ) {}
    };
}
