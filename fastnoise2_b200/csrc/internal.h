// fastnoise2_b200 internal -- declarations shared between the library's translation units.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include <FastNoise/Generators/Generator.h>

namespace FastNoise
{
    namespace Impl
    {
        // node_table.cpp
        std::uint64_t TreeStamp(); ///< bumped by every parameter write anywhere; compiled programs remember the stamp they saw
        bool LookupNeedsDomainWarp( std::uint8_t nodeId, int lookupIndex );
        const NodeState& DefaultState( std::uint8_t nodeId );

        // codec.cpp
        std::vector<std::uint8_t> Base64Decode( const char* text );
        std::string Base64Encode( const std::vector<std::uint8_t>& bytes );

        // lower.cpp
        enum EntryKind : int
        {
            kGrid2D = 2, kGrid3D = 3, kGrid4D = 4, kTileable2D = 5,
            kPositions2D = 12, kPositions3D = 13, kPositions4D = 14
        };
        inline int EntryDims( int kind ) { return kind == kTileable2D ? 4 : ( kind >= 12 ? kind - 10 : kind ); }
        inline bool EntryKindValid( int kind ) { return ( kind >= 2 && kind <= 5 ) || ( kind >= 12 && kind <= 14 ); }
        /// "points per thread" code of the uniform-grid entry for rows that are not a multiple of four wide: four consecutive points
        /// of the row-major output per thread, each with its own row (a group may run over the end of a row)
        constexpr int kRaggedPointsPerThread = 5;

        struct LoweredProgram
        {
            std::string source;  ///< complete CUDA translation unit (device library + generated tree + entry kernel)
            std::string error;   ///< non-empty if the tree cannot be lowered
            int pointsPerThread; ///< 1, 4 or kRaggedPointsPerThread (only the grid/positions index math depends on it)
            std::vector<float> params; ///< parameter mode: the values of the kernel's TreeParams argument, in slot order
        };
        /// Generates the CUDA source that evaluates the tree rooted at `root` for one entry kind.
        LoweredProgram LowerTree( const Generator* root, int entryKind, int pointsPerThread, int mathMode, bool paramMode );
        const char* DeviceLibrarySource();

        // backend.cpp
        void ReleaseCompiled( CompiledTree* compiled );
        void SetLastError( const std::string& message );
    }
}
