// fastnoise2_b200 -- forward declarations shared by every public header
// (same role as the reference's Utility/Config.h).
#pragma once
#include "Export.h"
#include <FastSIMD/DispatchClass.h>

// The reference computes OutputMinMax in every batch call (Utility/Config.h:5); so does this backend.
#define FASTNOISE_CALC_MIN_MAX true

namespace FastNoise
{
    class Generator;
    struct Metadata;

    template<typename T = Generator>
    class SmartNode;

    /// `const SmartNode<const T>&` -- the argument type of every source setter.
    template<typename T = Generator>
    using SmartNodeArg = const SmartNode<const T>&;

    template<typename T>
    SmartNode<T> New( FastSIMD::FeatureSet maxFeatureSet = FastSIMD::FeatureSet::Max );
}

#include "SmartNode.h"
