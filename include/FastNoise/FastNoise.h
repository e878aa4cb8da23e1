// fastnoise2_b200 -- umbrella header: every node type plus the two factories
// (reference: include/FastNoise/FastNoise.h:36-65).
#pragma once
#include "Utility/Config.h"

// node classes, by family
#include "Generators/BasicGenerators.h"   // Constant White Checkerboard SineWave Gradient DistanceToPoint
#include "Generators/Blends.h"            // Add ... Fade, PowFloat / PowInt, Min/MaxSmooth
#include "Generators/Cellular.h"          // CellularValue / Distance / Lookup
#include "Generators/DomainWarp.h"        // DomainWarpGradient
#include "Generators/DomainWarpFractal.h" // DomainWarpFractalProgressive / Independent
#include "Generators/DomainWarpSimplex.h" // DomainWarpSimplex / SuperSimplex
#include "Generators/Fractal.h"           // FractalFBm / Ridged
#include "Generators/Modifiers.h"         // Domain*, Remap, Terrace, SeedOffset, ...
#include "Generators/Perlin.h"
#include "Generators/Simplex.h"           // Simplex / SuperSimplex
#include "Generators/Value.h"

namespace FastNoise
{
    /// Create a node with default parameters, e.g. `auto fbm = FastNoise::New<FastNoise::FractalFBm>();`
    /// Returns an empty handle for FeatureSet::Invalid.
    template<typename T>
    SmartNode<T> New( FastSIMD::FeatureSet maxFeatureSet )
    {
        static_assert( std::is_base_of<Generator, T>::value, "New<T>: T must be a FastNoise node class, e.g. FastNoise::Simplex" );
        if( maxFeatureSet == FastSIMD::FeatureSet::Invalid )
        {
            return nullptr;
        }
        return SmartNode<T>::FromRaw( static_cast<T*>( Impl::CreateNode( T::kNodeId ) ) );
    }

    /// Build a node tree from the string the Node Editor exports, e.g. "DQkGDA==" (FBm over Simplex).
    /// nullptr on any malformed input; never throws.
    FASTNOISE_API SmartNode<> NewFromEncodedNodeTree( const char* encodedNodeTreeString, FastSIMD::FeatureSet maxFeatureSet = FastSIMD::FeatureSet::Max );
}
