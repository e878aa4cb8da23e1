// fastnoise2_b200 -- fractal node types: FractalFBm (13), FractalRidged (15) (reference: Generators/Fractal.h)
#pragma once
#include "Generator.h"
#include "BasicGenerators.h"

namespace FastNoise
{
    /// Octave loop over one source of type T (any Generator, or a DomainWarp for the warp fractals).
    template<typename T = Generator>
    class Fractal : public NodeBase
    {
    public:
        void SetSource( SmartNodeArg<T> gen ) { SetSrc( 0, gen.get() ); }
        void SetGain( float value ) { SetHyb( 0, value ); }
        void SetGain( SmartNodeArg<> gen ) { SetHyb( 0, gen.get() ); }
        void SetWeightedStrength( float value ) { SetHyb( 1, value ); }
        void SetWeightedStrength( SmartNodeArg<> gen ) { SetHyb( 1, gen.get() ); }
        void SetOctaveCount( int value ) { SetVar( 0, value ); }
        void SetLacunarity( float value ) { SetVar( 1, value ); }

    protected:
        explicit Fractal( std::uint8_t id ) : NodeBase( id ) {}
    };

    class FractalFBm : public Fractal<>
    {
        FN2_NODE( FractalFBm, 13 )
    protected:
        explicit FractalFBm( std::uint8_t id ) : Fractal( id ) {}
    };

    class FractalRidged : public Fractal<>
    {
        FN2_NODE( FractalRidged, 15 )
    protected:
        explicit FractalRidged( std::uint8_t id ) : Fractal( id ) {}
    };
}
