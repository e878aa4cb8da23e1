// fastnoise2_b200 internal -- backend entry points used by the C ABI and the Generator virtuals.
#pragma once
#include <string>
#include <vector>

#include <FastNoise/Generators/Generator.h>
#include "internal.h"

typedef struct CUstream_st* cudaStream_t;

namespace FastNoise
{
    namespace Impl
    {
        // host mirrors of the kernel argument structs in device/fn2_device_entry.cuh (layouts must match)
        struct FastDivHost { unsigned mul, shift, d, pad; };

        struct GridArgsHost
        {
            float off[4];
            float step[4];
            int count[4];
            unsigned rowBegin, total, vecStore, groups;
            FastDivHost divX, divY, divZ;
            float* mirror[7];
            unsigned mirrors, pad2;
        };

        struct TileArgsHost
        {
            float xFreq, yFreq, xMul, yMul;
            int xSize, ySize;
            unsigned rowBegin, total, vecStore, groups, pad[2];
            FastDivHost divX;
            float* mirror[7];
            unsigned mirrors, pad2;
        };

        /// Copies of the output in other GPUs' memory that a grid launch also writes (fused all-gather over peer memory).
        struct Mirrors
        {
            int count = 0;
            float* out[7] = {};
        };

        struct PosArgsHost
        {
            const float* pos[4];
            float off[4];
            unsigned long long count, groups;
            unsigned vecIO, pad;
        };

        /// One uniform-grid style call: kind = kGrid2D/3D/4D or kTileable2D; unused axes have count 1.
        struct GridCall
        {
            int kind;
            float offset[4];
            float step[4];
            int count[4];
            int seed;
        };

        const std::string& LastError();
        long long SelfTestInvSqrt( float lo, float hi ); // device_check.cu
        int SelfTestLeaf( int packed, int which, int n, int seed, float step, float* hostOut ); // device_check.cu
        int DebugIntermediates( int mode, int seed, float frequency, float jitter, int nx, int ny, int nz, float ox, float oy, float oz, float step, int* hostOut ); // device_check.cu
        int MathMode();
        void SetMathMode( int mode );
        unsigned long long LaunchCount();
        void CountLaunch();
        bool DeviceAvailable();
        cudaStream_t ThreadStream();
        OutputMinMax DecodeMinMax( const unsigned keys[2] );

        std::string KernelSource( const Generator* root, int kind, int pointsPerThread );
        std::vector<float> KernelParams( const Generator* root, int kind, int pointsPerThread ); ///< current values of the kernel's TreeParams argument
        int PrepareOffline( const Generator* root, int kind, int pointsPerThread ); ///< lower + NVRTC into the disk cache; needs no GPU
        int PrepareAll( const Generator* root, int kind ); ///< PrepareOffline for every points-per-thread variant a launch can pick

        /// Rows [rowBegin, rowBegin + rowCount) of the full grid into deviceOut (first element = first point of rowBegin). 0 or FNB200_E*.
        int LaunchGridRows( const Generator* root, const GridCall& call, float* deviceOut, unsigned long long rowBegin, unsigned long long rowCount,
                            unsigned* deviceMinMax, cudaStream_t stream, const Mirrors* mirrors = nullptr );
        int LaunchPositions( const Generator* root, int dims, float* deviceOut, unsigned long long count, const float* const pos[4],
                             const float offset[4], int seed, unsigned* deviceMinMax, cudaStream_t stream );

        /// Reference-shaped synchronous calls; `out` / `pos` may be host or device memory.
        OutputMinMax GenGrid( const Generator* root, const GridCall& call, float* out );
        OutputMinMax GenPositions( const Generator* root, int dims, float* out, long long count, const float* const pos[4], const float offset[4], int seed );

        // kernels.cu
        int FillUniformPhilox( float* deviceOut, long long count, long long firstIndex, unsigned key, unsigned axis, float lo, float hi, cudaStream_t stream );
        double MeasureAluPeak( int mode, int iters );
    }
}
