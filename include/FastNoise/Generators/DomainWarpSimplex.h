// fastnoise2_b200 -- DomainWarpSimplex (16), DomainWarpSuperSimplex (17) (reference: Generators/DomainWarpSimplex.h)
#pragma once
#include "DomainWarp.h"

namespace FastNoise
{
    enum class VectorizationScheme { OrthogonalGradientMatrix, GradientOuterProduct };

    constexpr static const char* kVectorizationScheme_Strings[] = { "Orthogonal Gradient Matrix", "Gradient Outer Product" };

    class DomainWarpSimplex : public DomainWarp
    {
        FN2_NODE( DomainWarpSimplex, 16 )
        void SetVectorizationScheme( VectorizationScheme value ) { SetVar( kVarEnd, (int)value ); }
    protected:
        explicit DomainWarpSimplex( std::uint8_t id ) : DomainWarp( id ) {}
    };

    class DomainWarpSuperSimplex : public DomainWarpSimplex
    {
        FN2_NODE( DomainWarpSuperSimplex, 17 )
    protected:
        explicit DomainWarpSuperSimplex( std::uint8_t id ) : DomainWarpSimplex( id ) {}
    };
}
