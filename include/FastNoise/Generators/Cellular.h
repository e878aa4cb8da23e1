// fastnoise2_b200 -- Cellular (Voronoi) node types, ids 10-12 (reference: Generators/Cellular.h)
#pragma once
#include "BasicGenerators.h"

namespace FastNoise
{
    template<typename PARENT = VariableRange<Seeded<ScalableGenerator>>>
    class Cellular : public PARENT
    {
    public:
        void SetDistanceFunction( DistanceFunction value ) { this->SetVar( kVarDistFunc, (int)value ); }
        void SetMinkowskiP( SmartNodeArg<> gen ) { this->SetHyb( 0, gen.get() ); }
        void SetMinkowskiP( float value ) { this->SetHyb( 0, value ); }
        void SetGridJitter( SmartNodeArg<> gen ) { this->SetHyb( 1, gen.get() ); }
        void SetGridJitter( float value ) { this->SetHyb( 1, value ); }
        void SetSizeJitter( SmartNodeArg<> gen ) { this->SetHyb( 2, gen.get() ); }
        void SetSizeJitter( float value ) { this->SetHyb( 2, value ); }

    protected:
        explicit Cellular( std::uint8_t id ) : PARENT( id ) {}
        static constexpr int kVarDistFunc = PARENT::kVarEnd, kVarEnd = kVarDistFunc + 1;
    };

    class CellularValue : public Cellular<>
    {
        FN2_NODE( CellularValue, 10 )
        static const int kMaxDistanceCount = 4;
        void SetValueIndex( int value ) { SetVar( kVarEnd, value ); } ///< clamped to [0, kMaxDistanceCount)
    protected:
        explicit CellularValue( std::uint8_t id ) : Cellular( id ) {}
    };

    class CellularDistance : public Cellular<>
    {
        FN2_NODE( CellularDistance, 11 )
        enum class ReturnType { Index0, Index0Add1, Index0Sub1, Index0Mul1, Index0Div1 };
        static const int kMaxDistanceCount = 4;
        void SetDistanceIndex0( int value ) { SetVar( kVarEnd, value ); }
        void SetDistanceIndex1( int value ) { SetVar( kVarEnd + 1, value ); }
        void SetReturnType( ReturnType value ) { SetVar( kVarEnd + 2, (int)value ); }
    protected:
        explicit CellularDistance( std::uint8_t id ) : Cellular( id ) {}
    };

    class CellularLookup : public Cellular<Seeded<ScalableGenerator>>
    {
        FN2_NODE( CellularLookup, 12 )
        void SetLookup( SmartNodeArg<> gen ) { SetSrc( 0, gen.get() ); }
    protected:
        explicit CellularLookup( std::uint8_t id ) : Cellular( id ) {}
    };
}
