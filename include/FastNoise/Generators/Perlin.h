// fastnoise2_b200 -- Perlin node type, id 8 (reference: Generators/Perlin.h)
#pragma once
#include "BasicGenerators.h"

namespace FastNoise
{
    class Perlin : public VariableRange<Seeded<ScalableGenerator>>
    {
        FN2_NODE( Perlin, 8 )
    protected:
        explicit Perlin( std::uint8_t id ) : VariableRange( id ) {}
    };
}
