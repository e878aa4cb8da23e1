// fastnoise2_b200 -- Value node type, id 9 (reference: Generators/Value.h)
#pragma once
#include "BasicGenerators.h"

namespace FastNoise
{
    class Value : public VariableRange<Seeded<ScalableGenerator>>
    {
        FN2_NODE( Value, 9 )
    protected:
        explicit Value( std::uint8_t id ) : VariableRange( id ) {}
    };
}
