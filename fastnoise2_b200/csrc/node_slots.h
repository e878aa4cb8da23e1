// fastnoise2_b200 internal -- where each node type keeps its parameters inside Impl::NodeState.
// Shared by the reflection table (node_table.cpp: which setter writes which slot) and the lowering
// pass (lower.cpp: which slot feeds which generated constant).
#pragma once
#include <cstdint>

namespace FastNoise
{
    namespace Slots
    {
        // node ids = registration order of the reference (src/FastNoise/FastSIMD_Build.inl:86-143)
        enum NodeId : std::uint8_t
        {
            N_Constant = 0, N_White, N_Checkerboard, N_SineWave, N_Gradient, N_DistanceToPoint,
            N_Simplex = 6, N_SuperSimplex, N_Perlin, N_Value,
            N_CellularValue = 10, N_CellularDistance, N_CellularLookup,
            N_FractalFBm = 13, N_PingPong, N_FractalRidged,
            N_DomainWarpSimplex = 16, N_DomainWarpSuperSimplex, N_DomainWarpGradient,
            N_DomainWarpFractalProgressive = 19, N_DomainWarpFractalIndependent,
            N_Add = 21, N_Subtract, N_Multiply, N_Divide,
            N_Abs = 25, N_Min, N_Max, N_MinSmooth, N_MaxSmooth, N_SignedSquareRoot, N_PowFloat, N_PowInt,
            N_DomainScale = 33, N_DomainOffset, N_DomainRotate, N_DomainAxisScale,
            N_SeedOffset = 37, N_ConvertRGBA8, N_GeneratorCache,
            N_Fade = 40, N_Remap, N_Terrace, N_AddDimension, N_RemoveDimension,
            N_Modulus = 45, N_DomainRotatePlane,
            NodeCount = 47
        };

        // float slots f[]
        enum : int
        {
            F_Scale = 0,      // ScalableGenerator: feature scale
            F_Frequency = 1,  //                    1.0f / scale (computed when the scale is set)
            F_RangeMin = 2,   // VariableRange: output min
            F_RangeScale = 3, //                output max - min, updated the way the reference's setters do
            F_K0 = 4          // first node-specific float
        };
        // node-specific floats, counted from F_K0:
        //   Constant: value | Gradient: multiplier X..W | Fractal*: lacunarity | DomainWarp*: axis scale X..W
        //   DomainScale: scale | DomainAxisScale: scale X..W | ConvertRGBA8: min, max
        //   DomainRotate: yawCos, yawSin, pitchCos, pitchSin, rollCos, rollSin
        //   Terrace: stepCount, 1/stepCount, 1 + 1/smoothness

        // int slots i[]
        enum : int
        {
            I_SeedOffset = 0, // Seeded
            I_K0 = 1          // first node-specific int
        };
        // node-specific ints, counted from I_K0:
        //   DistanceToPoint / Cellular*: distance function | CellularValue: +1 value index
        //   CellularDistance: +1 index0, +2 index1, +3 return type | Fractal*: octaves
        //   DomainWarpSimplex*: vectorization scheme | PowInt: pow | SeedOffset: offset | Fade: interpolation
        //   Remap: clamp | RemoveDimension: dimension | DomainRotatePlane: rotation type
    }
}
