// fastnoise2_b200 -- operator and blend node types (reference: Generators/Blends.h)
#pragma once
#include <climits>
#include "BasicGenerators.h"

namespace FastNoise
{
    /// LHS must be a node, RHS is a constant or a node.
    class OperatorSourceLHS : public NodeBase
    {
    public:
        void SetLHS( SmartNodeArg<> gen ) { SetSrc( 0, gen.get() ); }
        void SetRHS( SmartNodeArg<> gen ) { SetHyb( 0, gen.get() ); }
        void SetRHS( float value ) { SetHyb( 0, value ); }
    protected:
        explicit OperatorSourceLHS( std::uint8_t id ) : NodeBase( id ) {}
    };

    /// Both sides are a constant or a node.
    class OperatorHybridLHS : public NodeBase
    {
    public:
        void SetLHS( SmartNodeArg<> gen ) { SetHyb( 0, gen.get() ); }
        void SetLHS( float value ) { SetHyb( 0, value ); }
        void SetRHS( SmartNodeArg<> gen ) { SetHyb( 1, gen.get() ); }
        void SetRHS( float value ) { SetHyb( 1, value ); }
    protected:
        explicit OperatorHybridLHS( std::uint8_t id ) : NodeBase( id ) {}
    };

#define FN2_OPERATOR_NODE( CLASS, BASE, ID ) \
    class CLASS : public BASE \
    { \
        FN2_NODE( CLASS, ID ) \
    protected: \
        explicit CLASS( std::uint8_t id ) : BASE( id ) {} \
    }

    FN2_OPERATOR_NODE( Add, OperatorSourceLHS, 21 );
    FN2_OPERATOR_NODE( Subtract, OperatorHybridLHS, 22 );
    FN2_OPERATOR_NODE( Multiply, OperatorSourceLHS, 23 );
    FN2_OPERATOR_NODE( Divide, OperatorHybridLHS, 24 );
    FN2_OPERATOR_NODE( Min, OperatorSourceLHS, 26 );
    FN2_OPERATOR_NODE( Max, OperatorSourceLHS, 27 );
    FN2_OPERATOR_NODE( Modulus, OperatorHybridLHS, 45 );

#undef FN2_OPERATOR_NODE

    /// Polynomial smooth minimum of LHS and RHS with blend width Smoothness.
    class MinSmooth : public OperatorSourceLHS
    {
        FN2_NODE( MinSmooth, 28 )
        void SetSmoothness( SmartNodeArg<> gen ) { SetHyb( 1, gen.get() ); }
        void SetSmoothness( float value ) { SetHyb( 1, value ); }
    protected:
        explicit MinSmooth( std::uint8_t id ) : OperatorSourceLHS( id ) {}
    };

    class MaxSmooth : public OperatorSourceLHS
    {
        FN2_NODE( MaxSmooth, 29 )
        void SetSmoothness( SmartNodeArg<> gen ) { SetHyb( 1, gen.get() ); }
        void SetSmoothness( float value ) { SetHyb( 1, value ); }
    protected:
        explicit MaxSmooth( std::uint8_t id ) : OperatorSourceLHS( id ) {}
    };

    /// |Value| ^ Pow
    class PowFloat : public NodeBase
    {
        FN2_NODE( PowFloat, 31 )
        void SetValue( SmartNodeArg<> gen ) { SetHyb( 0, gen.get() ); }
        void SetValue( float value ) { SetHyb( 0, value ); }
        void SetPow( SmartNodeArg<> gen ) { SetHyb( 1, gen.get() ); }
        void SetPow( float value ) { SetHyb( 1, value ); }
    protected:
        explicit PowFloat( std::uint8_t id ) : NodeBase( id ) {}
    };

    /// Value multiplied by itself Pow-1 times (Pow >= 2)
    class PowInt : public NodeBase
    {
        FN2_NODE( PowInt, 32 )
        void SetValue( SmartNodeArg<> gen ) { SetSrc( 0, gen.get() ); }
        void SetPow( int value ) { SetVar( 0, value ); }
    protected:
        explicit PowInt( std::uint8_t id ) : NodeBase( id ) {}
    };

    /// lerp( A, B, ease( clamp01( (Fade - FadeMin) / (FadeMax - FadeMin) ) ) )
    class Fade : public NodeBase
    {
        FN2_NODE( Fade, 40 )
        enum class Interpolation { Linear, Hermite, Quintic };
        void SetA( SmartNodeArg<> gen ) { SetSrc( 0, gen.get() ); }
        void SetB( SmartNodeArg<> gen ) { SetSrc( 1, gen.get() ); }
        void SetFade( SmartNodeArg<> gen ) { SetHyb( 0, gen.get() ); }
        void SetFade( float value ) { SetHyb( 0, value ); }
        void SetFadeMin( SmartNodeArg<> gen ) { SetHyb( 1, gen.get() ); }
        void SetFadeMin( float value ) { SetHyb( 1, value ); }
        void SetFadeMax( SmartNodeArg<> gen ) { SetHyb( 2, gen.get() ); }
        void SetFadeMax( float value ) { SetHyb( 2, value ); }
        void SetInterpolation( Interpolation interpolation ) { SetVar( 0, (int)interpolation ); }
    protected:
        explicit Fade( std::uint8_t id ) : NodeBase( id ) {}
    };
}
