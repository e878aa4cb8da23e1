// fastnoise2_b200 -- DomainWarpFractalProgressive (19) / Independent (20) (reference: Generators/DomainWarpFractal.h)
#pragma once
#include "Fractal.h"
#include "DomainWarp.h"

namespace FastNoise
{
    class DomainWarpFractalProgressive : public Fractal<DomainWarp>
    {
        FN2_NODE( DomainWarpFractalProgressive, 19 )
    protected:
        explicit DomainWarpFractalProgressive( std::uint8_t id ) : Fractal( id ) {}
    };

    class DomainWarpFractalIndependent : public Fractal<DomainWarp>
    {
        FN2_NODE( DomainWarpFractalIndependent, 20 )
    protected:
        explicit DomainWarpFractalIndependent( std::uint8_t id ) : Fractal( id ) {}
    };
}
