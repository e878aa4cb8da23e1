// fastnoise2_b200 -- forward declarations shared by every public header
// (same role as the reference's Utility/Config.h).
#pragma once
#include <FastSIMD/DispatchClass.h>
#include "Export.h"

// Every batch call of this backend produces OutputMinMax on the device (block reduction + two atomics per block),
// as the reference does on the CPU (Utility/Config.h:5); consumers test this macro.
#ifndef FASTNOISE_CALC_MIN_MAX
#define FASTNOISE_CALC_MIN_MAX true
#endif

namespace FastNoise
{
    struct Metadata;
    class Generator;
    template<typename T = Generator> class SmartNode;

    /// `const SmartNode<const T>&` -- the argument type of every source setter, e.g. `SetSource( SmartNodeArg<> gen )`.
    template<typename T = Generator> using SmartNodeArg = const SmartNode<const T>&;

    /// Defined in FastNoise.h. FeatureSet::Max (or anything but Invalid) selects the one level there is: B200_SM100A.
    template<typename T> SmartNode<T> New( FastSIMD::FeatureSet maxFeatureSet = FastSIMD::FeatureSet::Max );
}

#include "SmartNode.h"
