/* fastnoise2_b200 -- C ABI of the node-graph evaluation path.
 *
 * These 45 entry points are the drop-in boundary: names, argument order and return conventions are
 * those of the reference's include/FastNoise/FastNoise_C.h:27-651 (implemented there by
 * src/FastNoise/FastNoise_C.cpp), so an existing cgo / P/Invoke / ctypes binding of FastNoise2 can be
 * pointed at libfastnoise2_b200.so unchanged. Each declaration names the reference line it replaces.
 *
 * Conventions (reference FastNoise_C.cpp:32-55,122-150,221-239):
 *   - node handles are opaque `void*` holding one reference; release with fnDeleteNodeRef
 *   - decode / create failures return NULL, never throw
 *   - `outputMinMax` is NULL or float[2] = { min, max }
 *   - name getters return "INVALID ..." literals, count/type getters return -1 on a bad id/index
 *   - setters return false when the index or the source node type does not fit
 *   - `maxFeatureSet`: pass ~0u (no limit). 0 (Invalid) yields NULL.
 * Extension: `noiseOut` / `*PosArray` may be CUDA device pointers (detected with
 * cudaPointerGetAttributes); host pointers are staged through pinned buffers. Device-side variants
 * with explicit slab origins and streams are in FastNoise_B200.h.
 */
#ifndef FASTNOISE_B200_C_H
#define FASTNOISE_B200_C_H

#include "Utility/Export.h"

#ifdef __cplusplus
extern "C" {
#else
#include <stdbool.h>
#endif

/* --- node lifetime (reference FastNoise_C.h:27,32,38,44) --- */
FASTNOISE_API void* fnNewFromEncodedNodeTree( const char* encodedString, unsigned maxFeatureSet );
FASTNOISE_API void fnDeleteNodeRef( void* node );
FASTNOISE_API unsigned fnGetActiveFeatureSet( const void* node );
FASTNOISE_API int fnGetMetadataID( const void* node );

/* --- batch generation: the hot path (reference FastNoise_C.h:66,99,134,164,195,228,250) --- */
FASTNOISE_API void fnGenUniformGrid2D( const void* node, float* noiseOut, float xOffset, float yOffset, int xCount, int yCount,
                                       float xStepSize, float yStepSize, int seed, float* outputMinMax );
FASTNOISE_API void fnGenUniformGrid3D( const void* node, float* noiseOut, float xOffset, float yOffset, float zOffset,
                                       int xCount, int yCount, int zCount, float xStepSize, float yStepSize, float zStepSize,
                                       int seed, float* outputMinMax );
FASTNOISE_API void fnGenUniformGrid4D( const void* node, float* noiseOut, float xOffset, float yOffset, float zOffset, float wOffset,
                                       int xCount, int yCount, int zCount, int wCount,
                                       float xStepSize, float yStepSize, float zStepSize, float wStepSize, int seed, float* outputMinMax );
FASTNOISE_API void fnGenPositionArray2D( const void* node, float* noiseOut, int count, const float* xPosArray, const float* yPosArray,
                                         float xOffset, float yOffset, int seed, float* outputMinMax );
FASTNOISE_API void fnGenPositionArray3D( const void* node, float* noiseOut, int count,
                                         const float* xPosArray, const float* yPosArray, const float* zPosArray,
                                         float xOffset, float yOffset, float zOffset, int seed, float* outputMinMax );
FASTNOISE_API void fnGenPositionArray4D( const void* node, float* noiseOut, int count,
                                         const float* xPosArray, const float* yPosArray, const float* zPosArray, const float* wPosArray,
                                         float xOffset, float yOffset, float zOffset, float wOffset, int seed, float* outputMinMax );
FASTNOISE_API void fnGenTileable2D( const void* node, float* noiseOut, int xSize, int ySize, float xStepSize, float yStepSize,
                                    int seed, float* outputMinMax );

/* --- single samples (reference FastNoise_C.h:267,282,298) --- */
FASTNOISE_API float fnGenSingle2D( const void* node, float x, float y, int seed );
FASTNOISE_API float fnGenSingle3D( const void* node, float x, float y, float z, int seed );
FASTNOISE_API float fnGenSingle4D( const void* node, float x, float y, float z, float w, int seed );

/* --- reflection: node types (reference FastNoise_C.h:307,318,332) --- */
FASTNOISE_API int fnGetMetadataCount();
FASTNOISE_API const char* fnGetMetadataName( int id );
FASTNOISE_API void* fnNewFromMetadata( int id, unsigned maxFeatureSet );

/* --- reflection: variables (reference FastNoise_C.h:343-427) --- */
FASTNOISE_API int fnGetMetadataVariableCount( int id );
FASTNOISE_API const char* fnGetMetadataVariableName( int id, int variableIndex );
FASTNOISE_API int fnGetMetadataVariableType( int id, int variableIndex ); /* 0 float, 1 int, 2 enum */
FASTNOISE_API int fnGetMetadataVariableDimensionIdx( int id, int variableIndex );
FASTNOISE_API int fnGetMetadataEnumCount( int id, int variableIndex );
FASTNOISE_API const char* fnGetMetadataEnumName( int id, int variableIndex, int enumIndex );
FASTNOISE_API bool fnSetVariableFloat( void* node, int variableIndex, float value );
FASTNOISE_API bool fnSetVariableIntEnum( void* node, int variableIndex, int value );

/* --- reflection: required node inputs (reference FastNoise_C.h:438-472) --- */
FASTNOISE_API int fnGetMetadataNodeLookupCount( int id );
FASTNOISE_API const char* fnGetMetadataNodeLookupName( int id, int nodeLookupIndex );
FASTNOISE_API int fnGetMetadataNodeLookupDimensionIdx( int id, int nodeLookupIndex );
FASTNOISE_API bool fnSetNodeLookup( void* node, int nodeLookupIndex, const void* nodeLookup );

/* --- reflection: hybrid (constant-or-node) inputs (reference FastNoise_C.h:486-534) --- */
FASTNOISE_API int fnGetMetadataHybridCount( int id );
FASTNOISE_API const char* fnGetMetadataHybridName( int id, int hybridIndex );
FASTNOISE_API int fnGetMetadataHybridDimensionIdx( int id, int hybridIndex );
FASTNOISE_API bool fnSetHybridNodeLookup( void* node, int hybridIndex, const void* nodeLookup );
FASTNOISE_API bool fnSetHybridFloat( void* node, int hybridIndex, float value );

/* --- reflection: descriptions, groups, defaults and limits (reference FastNoise_C.h:545-651) --- */
FASTNOISE_API const char* fnGetMetadataDescription( int id );
FASTNOISE_API int fnGetMetadataGroupCount( int id );
FASTNOISE_API const char* fnGetMetadataGroupName( int id, int groupIndex );
FASTNOISE_API const char* fnGetMetadataVariableDescription( int id, int variableIndex );
FASTNOISE_API float fnGetMetadataVariableDefaultFloat( int id, int variableIndex );
FASTNOISE_API int fnGetMetadataVariableDefaultIntEnum( int id, int variableIndex );
FASTNOISE_API float fnGetMetadataVariableMinFloat( int id, int variableIndex );
FASTNOISE_API float fnGetMetadataVariableMaxFloat( int id, int variableIndex );
FASTNOISE_API const char* fnGetMetadataNodeLookupDescription( int id, int nodeLookupIndex );
FASTNOISE_API const char* fnGetMetadataHybridDescription( int id, int hybridIndex );
FASTNOISE_API float fnGetMetadataHybridDefault( int id, int hybridIndex );

#ifdef __cplusplus
}
#endif

#endif
