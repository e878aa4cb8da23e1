// fastnoise2_b200 -- DomainWarp base and DomainWarpGradient (18) (reference: Generators/DomainWarp.h)
#pragma once
#include "BasicGenerators.h"

namespace FastNoise
{
    /// Moves the sample position, then evaluates Source there. Output value is Source's.
    class DomainWarp : public Seeded<ScalableGenerator>
    {
    public:
        void SetSource( SmartNodeArg<> gen ) { SetSrc( 0, gen.get() ); }
        void SetWarpAmplitude( SmartNodeArg<> gen ) { SetHyb( 0, gen.get() ); }
        void SetWarpAmplitude( float value ) { SetHyb( 0, value ); }
        template<Dim D> void SetAmplitudeScaling( float value ) { SetVar( kVarAxisScale + (int)D, value ); }

    protected:
        explicit DomainWarp( std::uint8_t id ) : Seeded( id ) {}
        static constexpr int kVarAxisScale = Seeded::kVarEnd, kVarEnd = kVarAxisScale + 4;
    };

    class DomainWarpGradient : public DomainWarp
    {
        FN2_NODE( DomainWarpGradient, 18 )
    protected:
        explicit DomainWarpGradient( std::uint8_t id ) : DomainWarp( id ) {}
    };
}
