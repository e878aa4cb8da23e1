// fastnoise2_b200 -- Simplex (id 6) and SuperSimplex (id 7) node types (reference: Generators/Simplex.h)
#pragma once
#include "BasicGenerators.h"

namespace FastNoise
{
    class Simplex : public VariableRange<Seeded<ScalableGenerator>>
    {
        FN2_NODE( Simplex, 6 )
    protected:
        explicit Simplex( std::uint8_t id ) : VariableRange( id ) {}
    };

    class SuperSimplex : public VariableRange<Seeded<ScalableGenerator>>
    {
        FN2_NODE( SuperSimplex, 7 )
    protected:
        explicit SuperSimplex( std::uint8_t id ) : VariableRange( id ) {}
    };
}
