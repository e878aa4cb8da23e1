// fastnoise2_b200 -- FastNoise::Generator, the abstract node interface of the hot path.
//
// Method names, argument order and meaning, output layout and return conventions follow the
// reference's Generator (include/FastNoise/Generators/Generator.h:173-450). What is behind them is
// new: the first Gen* call on a node lowers the tree rooted at that node into CUDA source, compiles
// it for sm_100a (NVRTC, cached by source hash) and launches it; one thread evaluates the whole tree
// for its output points in registers. `out` (and position arrays) may be host pointers -- staged
// through pinned buffers in pipelined chunks -- or device pointers, written in place.
//
// Unlike the reference there is one concrete parameter block per node (Impl::NodeState) instead of
// per-class members plus a per-ISA DispatchClass; typed setters and the reflection table both write
// it through Impl::Set* so they cannot disagree.
#pragma once
#include <atomic>
#include <cmath>
#include <cstdint>
#include <algorithm>

#include "FastNoise/Utility/Config.h"

namespace FastNoise
{
    enum class Dim { X, Y, Z, W, Count };

    constexpr static const char* kDim_Strings[] = { "X", "Y", "Z", "W" };

    /// Distance metrics of the Cellular nodes and DistanceToPoint (values are the encoded ones).
    enum class DistanceFunction { Euclidean, EuclideanSquared, Manhattan, Hybrid, MaxAxis, Minkowski };

    constexpr static const char* kDistanceFunction_Strings[] = { "Euclidean", "Euclidean Squared", "Manhattan", "Hybrid", "Max Axis", "Minkowski" };

    enum class Boolean { False, True };

    constexpr static const char* kBoolean_Strings[] = { "False", "True" };

    static constexpr float kInfinity = __builtin_inff();

    /// Smallest/largest value written by a batch call. Default-constructed = empty range (min > max),
    /// which is also what a failed call returns.
    struct OutputMinMax
    {
        float min = kInfinity;
        float max = -kInfinity;

        OutputMinMax& operator<<( float v )
        {
            min = std::min( min, v );
            max = std::max( max, v );
            return *this;
        }

        OutputMinMax& operator<<( const OutputMinMax& v )
        {
            min = std::min( min, v.min );
            max = std::max( max, v.max );
            return *this;
        }
    };

    namespace Impl
    {
        struct CompiledTree;

        /// Uniform parameter block of a node instance. Slot meaning per node type: csrc/node_slots.h.
        struct NodeState
        {
            static constexpr int kFloats = 16, kInts = 6, kLookups = 2, kHybrids = 5;

            struct Hybrid
            {
                SmartNode<const Generator> node; ///< evaluated per point when set
                float constant = 0;              ///< used otherwise
            };

            std::uint8_t nodeId = 0;
            float f[kFloats] = {};
            std::int32_t i[kInts] = {};
            SmartNode<const Generator> lookups[kLookups];
            Hybrid hybrids[kHybrids];
        };

        // One implementation of every parameter write, shared by typed setters, Metadata setFuncs and the C ABI.
        FASTNOISE_API void SetVariable( Generator* node, int variableIndex, std::int32_t rawBits );
        FASTNOISE_API bool SetLookup( Generator* node, int lookupIndex, const Generator* source );
        FASTNOISE_API bool SetHybridNode( Generator* node, int hybridIndex, const Generator* source );
        FASTNOISE_API void SetHybridValue( Generator* node, int hybridIndex, float value );
        FASTNOISE_API Generator* CreateNode( std::uint8_t nodeId );

        inline std::int32_t Bits( float v ) { std::int32_t b; __builtin_memcpy( &b, &v, 4 ); return b; }
        inline std::int32_t Bits( int v ) { return v; }
    }

    class FASTNOISE_API Generator
    {
    public:
        Generator( const Generator& ) = delete;
        Generator( Generator&& ) = delete;
        virtual ~Generator();

        /// Always FeatureSet::B200_SM100A.
        virtual FastSIMD::FeatureSet GetActiveFeatureSet() const;
        virtual const Metadata& GetMetadata() const;

        /// out[y * xCount + x] = tree( fma( float(x), xStepSize, xOffset ), fma( float(y), yStepSize, yOffset ) )
        virtual OutputMinMax GenUniformGrid2D( float* out, float xOffset, float yOffset, int xCount, int yCount, float xStepSize, float yStepSize, int seed ) const;

        /// out[(z * yCount + y) * xCount + x]
        virtual OutputMinMax GenUniformGrid3D( float* out, float xOffset, float yOffset, float zOffset, int xCount, int yCount, int zCount, float xStepSize, float yStepSize, float zStepSize, int seed ) const;

        /// out[((w * zCount + z) * yCount + y) * xCount + x]
        virtual OutputMinMax GenUniformGrid4D( float* out, float xOffset, float yOffset, float zOffset, float wOffset, int xCount, int yCount, int zCount, int wCount, float xStepSize, float yStepSize, float zStepSize, float wStepSize, int seed ) const;

        /// Seamlessly tiling 2D image: the grid is wrapped on a 4D torus and the tree is evaluated in 4D.
        virtual OutputMinMax GenTileable2D( float* out, int xSize, int ySize, float xStepSize, float yStepSize, int seed ) const;

        /// out[i] = tree( xPosArray[i] + xOffset, ... ), struct-of-arrays positions
        virtual OutputMinMax GenPositionArray2D( float* out, int count, const float* xPosArray, const float* yPosArray, float xOffset, float yOffset, int seed ) const;
        virtual OutputMinMax GenPositionArray3D( float* out, int count, const float* xPosArray, const float* yPosArray, const float* zPosArray, float xOffset, float yOffset, float zOffset, int seed ) const;
        virtual OutputMinMax GenPositionArray4D( float* out, int count, const float* xPosArray, const float* yPosArray, const float* zPosArray, const float* wPosArray, float xOffset, float yOffset, float zOffset, float wOffset, int seed ) const;

        /// One point = one tiny launch plus a 4-byte read-back; as in the reference, only for occasional samples.
        virtual float GenSingle2D( float x, float y, int seed ) const;
        virtual float GenSingle3D( float x, float y, float z, int seed ) const;
        virtual float GenSingle4D( float x, float y, float z, float w, int seed ) const;

        /// Parameter block (read by the lowering pass; written only through Impl::Set*).
        const Impl::NodeState& State() const { return mState; }

    protected:
        explicit Generator( std::uint8_t nodeId );

        template<typename V>
        void SetVar( int variableIndex, V value ) { Impl::SetVariable( this, variableIndex, Impl::Bits( value ) ); }
        void SetSrc( int lookupIndex, const Generator* source ) { Impl::SetLookup( this, lookupIndex, source ); }
        void SetHyb( int hybridIndex, const Generator* source ) { Impl::SetHybridNode( this, hybridIndex, source ); }
        void SetHyb( int hybridIndex, float value ) { Impl::SetHybridValue( this, hybridIndex, value ); }

    private:
        Impl::NodeState mState;
        mutable std::atomic<std::uint32_t> mRefs { 0 };
        mutable std::atomic<Impl::CompiledTree*> mCompiled { nullptr };
        mutable std::atomic<std::uint64_t> mCompiledStamp { 0 };

        friend std::int32_t Internal::NodeRefAdd( const Generator*, std::int32_t );
        friend void Internal::NodeDestroy( const Generator* );
        friend struct Impl::CompiledTree;
        friend FASTNOISE_API void Impl::SetVariable( Generator*, int, std::int32_t );
        friend FASTNOISE_API bool Impl::SetLookup( Generator*, int, const Generator* );
        friend FASTNOISE_API bool Impl::SetHybridNode( Generator*, int, const Generator* );
        friend FASTNOISE_API void Impl::SetHybridValue( Generator*, int, float );
        friend FASTNOISE_API Generator* Impl::CreateNode( std::uint8_t );
        friend class Backend;
    };

    /// Kept for source compatibility with code that names the reference's source-holder types.
    template<typename T>
    struct PerDimensionVariable
    {
        using Type = T;
        T varArray[(int)Dim::Count];

        template<typename U = T>
        PerDimensionVariable( U value = 0 )
        {
            for( T& element : varArray ) element = value;
        }

        T& operator[]( std::size_t i ) { return varArray[i]; }
        const T& operator[]( std::size_t i ) const { return varArray[i]; }
    };
}
