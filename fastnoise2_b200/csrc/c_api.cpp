// fastnoise2_b200 -- C ABI: the 45 reference entry points (include/FastNoise/FastNoise_C.h) and the
// device-side extension (include/FastNoise/FastNoise_B200.h). Handle convention as in the reference's
// FastNoise_C.cpp:9-55: a handle is a Generator* that owns one reference; unlike the reference,
// fnDeleteNodeRef destroys the node when that was the last reference (the reference only decrements).
#include <FastNoise/FastNoise_C.h>
#include <FastNoise/FastNoise_B200.h>
#include <FastNoise/FastNoise.h>
#include <FastNoise/Metadata.h>

#include <cuda_runtime_api.h>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "backend.h"

using namespace FastNoise;

namespace
{
    Generator* Gen( void* p ) { return static_cast<Generator*>( p ); }
    const Generator* Gen( const void* p ) { return static_cast<const Generator*>( p ); }

    const Metadata* Meta( int id )
    {
        return id >= 0 && id <= 255 ? Metadata::GetFromId( (Metadata::node_id)id ) : nullptr;
    }

    void* Detach( SmartNode<>&& node ) // handle = pointer + one manual reference
    {
        if( !node ) return nullptr;
        Internal::NodeRefAdd( node.get(), 1 );
        return node.get();
    }

    void StoreMinMax( float* out2, const OutputMinMax& mm )
    {
        if( out2 )
        {
            out2[0] = mm.min;
            out2[1] = mm.max;
        }
    }
}

extern "C"
{
    void* fnNewFromEncodedNodeTree( const char* encodedString, unsigned maxFeatureSet )
    {
        return Detach( NewFromEncodedNodeTree( encodedString, (FastSIMD::FeatureSet)maxFeatureSet ) );
    }

    void fnDeleteNodeRef( void* node )
    {
        if( node && Internal::NodeRefAdd( Gen( node ), -1 ) == 1 ) Internal::NodeDestroy( Gen( node ) );
    }

    unsigned fnGetActiveFeatureSet( const void* node ) { return (unsigned)Gen( node )->GetActiveFeatureSet(); }
    int fnGetMetadataID( const void* node ) { return Gen( node )->GetMetadata().id; }

    void fnGenUniformGrid2D( const void* node, float* noiseOut, float xOffset, float yOffset, int xCount, int yCount, float xStepSize, float yStepSize, int seed, float* outputMinMax )
    {
        StoreMinMax( outputMinMax, Gen( node )->GenUniformGrid2D( noiseOut, xOffset, yOffset, xCount, yCount, xStepSize, yStepSize, seed ) );
    }

    void fnGenUniformGrid3D( const void* node, float* noiseOut, float xOffset, float yOffset, float zOffset, int xCount, int yCount, int zCount, float xStepSize, float yStepSize, float zStepSize, int seed, float* outputMinMax )
    {
        StoreMinMax( outputMinMax, Gen( node )->GenUniformGrid3D( noiseOut, xOffset, yOffset, zOffset, xCount, yCount, zCount, xStepSize, yStepSize, zStepSize, seed ) );
    }

    void fnGenUniformGrid4D( const void* node, float* noiseOut, float xOffset, float yOffset, float zOffset, float wOffset, int xCount, int yCount, int zCount, int wCount, float xStepSize, float yStepSize, float zStepSize, float wStepSize, int seed, float* outputMinMax )
    {
        StoreMinMax( outputMinMax, Gen( node )->GenUniformGrid4D( noiseOut, xOffset, yOffset, zOffset, wOffset, xCount, yCount, zCount, wCount, xStepSize, yStepSize, zStepSize, wStepSize, seed ) );
    }

    void fnGenPositionArray2D( const void* node, float* noiseOut, int count, const float* xPosArray, const float* yPosArray, float xOffset, float yOffset, int seed, float* outputMinMax )
    {
        StoreMinMax( outputMinMax, Gen( node )->GenPositionArray2D( noiseOut, count, xPosArray, yPosArray, xOffset, yOffset, seed ) );
    }

    void fnGenPositionArray3D( const void* node, float* noiseOut, int count, const float* xPosArray, const float* yPosArray, const float* zPosArray, float xOffset, float yOffset, float zOffset, int seed, float* outputMinMax )
    {
        StoreMinMax( outputMinMax, Gen( node )->GenPositionArray3D( noiseOut, count, xPosArray, yPosArray, zPosArray, xOffset, yOffset, zOffset, seed ) );
    }

    void fnGenPositionArray4D( const void* node, float* noiseOut, int count, const float* xPosArray, const float* yPosArray, const float* zPosArray, const float* wPosArray, float xOffset, float yOffset, float zOffset, float wOffset, int seed, float* outputMinMax )
    {
        StoreMinMax( outputMinMax, Gen( node )->GenPositionArray4D( noiseOut, count, xPosArray, yPosArray, zPosArray, wPosArray, xOffset, yOffset, zOffset, wOffset, seed ) );
    }

    void fnGenTileable2D( const void* node, float* noiseOut, int xSize, int ySize, float xStepSize, float yStepSize, int seed, float* outputMinMax )
    {
        StoreMinMax( outputMinMax, Gen( node )->GenTileable2D( noiseOut, xSize, ySize, xStepSize, yStepSize, seed ) );
    }

    float fnGenSingle2D( const void* node, float x, float y, int seed ) { return Gen( node )->GenSingle2D( x, y, seed ); }
    float fnGenSingle3D( const void* node, float x, float y, float z, int seed ) { return Gen( node )->GenSingle3D( x, y, z, seed ); }
    float fnGenSingle4D( const void* node, float x, float y, float z, float w, int seed ) { return Gen( node )->GenSingle4D( x, y, z, w, seed ); }

    // ------------------------------------------------------------------ reflection
    int fnGetMetadataCount() { return (int)Metadata::GetAll().size(); }

    const char* fnGetMetadataName( int id )
    {
        const Metadata* m = Meta( id );
        return m ? m->name : "INVALID NODE ID";
    }

    void* fnNewFromMetadata( int id, unsigned maxFeatureSet )
    {
        const Metadata* m = Meta( id );
        return m ? Detach( m->CreateNode( (FastSIMD::FeatureSet)maxFeatureSet ) ) : nullptr;
    }

#define FN2_MEMBER( LIST, INDEX, BAD_ID, BAD_INDEX, EXPR ) \
    const Metadata* m = Meta( id ); \
    if( !m ) return BAD_ID; \
    if( (size_t)INDEX >= m->LIST.size() ) return BAD_INDEX; \
    const auto& member = m->LIST[INDEX]; \
    return EXPR;

    int fnGetMetadataVariableCount( int id ) { const Metadata* m = Meta( id ); return m ? (int)m->memberVariables.size() : -1; }
    const char* fnGetMetadataVariableName( int id, int variableIndex ) { FN2_MEMBER( memberVariables, variableIndex, "INVALID NODE ID", "INVALID VARIABLE INDEX", member.name ) }
    int fnGetMetadataVariableType( int id, int variableIndex ) { FN2_MEMBER( memberVariables, variableIndex, -1, -1, (int)member.type ) }
    int fnGetMetadataVariableDimensionIdx( int id, int variableIndex ) { FN2_MEMBER( memberVariables, variableIndex, -1, -1, member.dimensionIdx ) }
    int fnGetMetadataEnumCount( int id, int variableIndex ) { FN2_MEMBER( memberVariables, variableIndex, -1, -1, (int)member.enumNames.size() ) }

    const char* fnGetMetadataEnumName( int id, int variableIndex, int enumIndex )
    {
        const Metadata* m = Meta( id );
        if( !m ) return "INVALID NODE ID";
        if( (size_t)variableIndex >= m->memberVariables.size() ) return "INVALID VARIABLE INDEX";
        const auto& names = m->memberVariables[variableIndex].enumNames;
        return (size_t)enumIndex < names.size() ? names[enumIndex] : "INVALID ENUM INDEX";
    }

    bool fnSetVariableFloat( void* node, int variableIndex, float value )
    {
        const Metadata& m = Gen( node )->GetMetadata();
        return (size_t)variableIndex < m.memberVariables.size() && m.memberVariables[variableIndex].setFunc( Gen( node ), value );
    }

    bool fnSetVariableIntEnum( void* node, int variableIndex, int value )
    {
        const Metadata& m = Gen( node )->GetMetadata();
        return (size_t)variableIndex < m.memberVariables.size() && m.memberVariables[variableIndex].setFunc( Gen( node ), value );
    }

    int fnGetMetadataNodeLookupCount( int id ) { const Metadata* m = Meta( id ); return m ? (int)m->memberNodeLookups.size() : -1; }
    const char* fnGetMetadataNodeLookupName( int id, int nodeLookupIndex ) { FN2_MEMBER( memberNodeLookups, nodeLookupIndex, "INVALID NODE ID", "INVALID NODE LOOKUP INDEX", member.name ) }
    int fnGetMetadataNodeLookupDimensionIdx( int id, int nodeLookupIndex ) { FN2_MEMBER( memberNodeLookups, nodeLookupIndex, -1, -1, member.dimensionIdx ) }

    bool fnSetNodeLookup( void* node, int nodeLookupIndex, const void* nodeLookup )
    {
        const Metadata& m = Gen( node )->GetMetadata();
        if( (size_t)nodeLookupIndex >= m.memberNodeLookups.size() ) return false;
        SmartNode<const Generator> source = SmartNode<const Generator>::FromRaw( Gen( nodeLookup ) );
        return m.memberNodeLookups[nodeLookupIndex].setFunc( Gen( node ), source );
    }

    int fnGetMetadataHybridCount( int id ) { const Metadata* m = Meta( id ); return m ? (int)m->memberHybrids.size() : -1; }
    const char* fnGetMetadataHybridName( int id, int hybridIndex ) { FN2_MEMBER( memberHybrids, hybridIndex, "INVALID NODE ID", "INVALID HYBRID INDEX", member.name ) }
    int fnGetMetadataHybridDimensionIdx( int id, int hybridIndex ) { FN2_MEMBER( memberHybrids, hybridIndex, -1, -1, member.dimensionIdx ) }

    bool fnSetHybridNodeLookup( void* node, int hybridIndex, const void* nodeLookup )
    {
        const Metadata& m = Gen( node )->GetMetadata();
        if( (size_t)hybridIndex >= m.memberHybrids.size() ) return false;
        SmartNode<const Generator> source = SmartNode<const Generator>::FromRaw( Gen( nodeLookup ) );
        return m.memberHybrids[hybridIndex].setNodeFunc( Gen( node ), source );
    }

    bool fnSetHybridFloat( void* node, int hybridIndex, float value )
    {
        const Metadata& m = Gen( node )->GetMetadata();
        return (size_t)hybridIndex < m.memberHybrids.size() && m.memberHybrids[hybridIndex].setValueFunc( Gen( node ), value );
    }

    const char* fnGetMetadataDescription( int id ) { const Metadata* m = Meta( id ); return m ? m->description : "INVALID NODE ID"; }
    int fnGetMetadataGroupCount( int id ) { const Metadata* m = Meta( id ); return m ? (int)m->groups.size() : -1; }

    const char* fnGetMetadataGroupName( int id, int groupIndex )
    {
        const Metadata* m = Meta( id );
        if( !m ) return "INVALID NODE ID";
        return (size_t)groupIndex < m->groups.size() ? m->groups[groupIndex] : "INVALID GROUP INDEX";
    }

    const char* fnGetMetadataVariableDescription( int id, int variableIndex ) { FN2_MEMBER( memberVariables, variableIndex, "INVALID NODE ID", "INVALID VARIABLE INDEX", member.description ) }
    float fnGetMetadataVariableDefaultFloat( int id, int variableIndex ) { FN2_MEMBER( memberVariables, variableIndex, 0.0f, 0.0f, member.valueDefault.f ) }
    int fnGetMetadataVariableDefaultIntEnum( int id, int variableIndex ) { FN2_MEMBER( memberVariables, variableIndex, 0, 0, member.valueDefault.i ) }
    float fnGetMetadataVariableMinFloat( int id, int variableIndex ) { FN2_MEMBER( memberVariables, variableIndex, 0.0f, 0.0f, member.valueMin.f ) }
    float fnGetMetadataVariableMaxFloat( int id, int variableIndex ) { FN2_MEMBER( memberVariables, variableIndex, 0.0f, 0.0f, member.valueMax.f ) }
    const char* fnGetMetadataNodeLookupDescription( int id, int nodeLookupIndex ) { FN2_MEMBER( memberNodeLookups, nodeLookupIndex, "INVALID NODE ID", "INVALID NODE LOOKUP INDEX", member.description ) }
    const char* fnGetMetadataHybridDescription( int id, int hybridIndex ) { FN2_MEMBER( memberHybrids, hybridIndex, "INVALID NODE ID", "INVALID HYBRID INDEX", member.description ) }
    float fnGetMetadataHybridDefault( int id, int hybridIndex ) { FN2_MEMBER( memberHybrids, hybridIndex, 0.0f, 0.0f, member.valueDefault ) }

#undef FN2_MEMBER

    // ------------------------------------------------------------------ FastNoise_B200.h
    const char* fnb200LastError( void ) { return Impl::LastError().c_str(); }
    int fnb200DeviceAvailable( void ) { return Impl::DeviceAvailable() ? 1 : 0; }
    int fnb200SetMathMode( int mode ) { Impl::SetMathMode( mode ); return FNB200_OK; }
    int fnb200GetMathMode( void ) { return Impl::MathMode(); }
    unsigned long long fnb200LaunchCount( void ) { return Impl::LaunchCount(); }

    // which points-per-thread variant fnb200GetKernelSource / fnb200GetKernelParams describe: 4 unless FASTNOISE2_B200_PPT says 1
    // or 5 (the any-width form of the uniform-grid entries, tests only)
    static int DefaultPpt( int kind )
    {
        const char* v = std::getenv( "FASTNOISE2_B200_PPT" );
        int p = v && *v ? std::atoi( v ) : 4;
        if( p == Impl::kRaggedPointsPerThread && kind >= Impl::kGrid2D && kind <= Impl::kGrid4D ) return p;
        return p >= 4 ? 4 : 1;
    }

    int fnb200Prepare( const void* node, int kind )
    {
        if( !node || !Impl::EntryKindValid( kind ) )
        {
            Impl::SetLastError( "fnb200Prepare: bad node or kind" );
            return FNB200_EINVAL;
        }
        return Impl::PrepareAll( Gen( node ), kind );
    }

    long fnb200GetKernelParams( const void* node, int kind, float* buf, long bufCount )
    {
        if( !node || !Impl::EntryKindValid( kind ) ) return FNB200_EINVAL;
        const std::vector<float> params = Impl::KernelParams( Gen( node ), kind, DefaultPpt( kind ) );
        for( long i = 0; buf && i < bufCount && i < (long)params.size(); i++ ) buf[i] = params[i];
        return (long)params.size();
    }

    long fnb200GetKernelSource( const void* node, int kind, char* buf, long bufSize )
    {
        if( !node || !Impl::EntryKindValid( kind ) ) return FNB200_EINVAL;
        std::string source = Impl::KernelSource( Gen( node ), kind, DefaultPpt( kind ) );
        if( buf && bufSize > 0 )
        {
            long n = std::min<long>( bufSize - 1, (long)source.size() );
            std::memcpy( buf, source.data(), (size_t)n );
            buf[n] = 0;
        }
        return (long)source.size();
    }

    // NULL is CUDA's default (legacy) stream, as everywhere else in CUDA; torch's default stream is that stream too
    static cudaStream_t StreamOf( void* stream ) { return (cudaStream_t)stream; }

    static int CudaRc( cudaError_t e, const char* what )
    {
        if( e == cudaSuccess ) return FNB200_OK;
        Impl::SetLastError( std::string( what ) + ": " + cudaGetErrorString( e ) );
        cudaGetLastError();
        return FNB200_ECUDA;
    }

    int fnb200MinMaxReset( void* deviceMinMax, void* stream )
    {
        if( !deviceMinMax ) return FNB200_EINVAL;
        return CudaRc( cudaMemsetAsync( deviceMinMax, 0, 2 * sizeof( unsigned ), StreamOf( stream ) ), "cudaMemsetAsync" );
    }

    int fnb200MinMaxRead( const void* deviceMinMax, void* stream, float* outputMinMax )
    {
        if( !deviceMinMax || !outputMinMax ) return FNB200_EINVAL;
        unsigned keys[2] = { 0, 0 };
        cudaStream_t s = StreamOf( stream );
        if( int rc = CudaRc( cudaMemcpyAsync( keys, deviceMinMax, sizeof( keys ), cudaMemcpyDeviceToHost, s ), "cudaMemcpyAsync" ) ) return rc;
        if( int rc = CudaRc( cudaStreamSynchronize( s ), "cudaStreamSynchronize" ) ) return rc;
        StoreMinMax( outputMinMax, Impl::DecodeMinMax( keys ) );
        return FNB200_OK;
    }

    static int SlabLaunch( const void* node, Impl::GridCall& call, int outerAxis, float* deviceOut, int slabBegin, int slabEnd, void* deviceMinMax, void* stream, void* sharedVolume = nullptr )
    {
        Impl::Mirrors mirrors;
        if( sharedVolume )
        {
            // the slab is generated at its place inside the shared volume and stored into every rank's copy
            long long first = slabBegin;
            for( int d = 0; d < outerAxis; d++ ) first *= call.count[d];
            if( slabBegin < 0 || fnb200VolumeMirrors( sharedVolume, first, &deviceOut, mirrors.out, &mirrors.count ) != FNB200_OK )
            {
                Impl::SetLastError( "shared-volume call: bad volume or slab" );
                return FNB200_EINVAL;
            }
        }
        if( !node || !deviceOut || slabBegin < 0 || slabEnd > call.count[outerAxis] || slabBegin >= slabEnd )
        {
            Impl::SetLastError( "slab call: bad node, output or slab range" );
            return FNB200_EINVAL;
        }
        unsigned long long rowsPerOuter = 1;
        for( int d = 1; d < outerAxis; d++ ) rowsPerOuter *= (unsigned long long)call.count[d];
        for( int d = 0; d <= outerAxis; d++ )
        {
            if( call.count[d] <= 0 )
            {
                Impl::SetLastError( "slab call: grid size must be positive" );
                return FNB200_EINVAL;
            }
        }
        return Impl::LaunchGridRows( Gen( node ), call, deviceOut, rowsPerOuter * (unsigned)slabBegin, rowsPerOuter * (unsigned)( slabEnd - slabBegin ),
                                     (unsigned*)deviceMinMax, StreamOf( stream ), sharedVolume ? &mirrors : nullptr );
    }

    static Impl::GridCall EmptyCall( int kind, int seed )
    {
        Impl::GridCall c {};
        c.kind = kind;
        c.seed = seed;
        for( int d = 0; d < 4; d++ ) c.count[d] = 1;
        return c;
    }

    int fnb200GenUniformGrid2D( const void* node, float* deviceOut, float xOffset, float yOffset, int xCount, int yCount, float xStepSize, float yStepSize,
                                int seed, int slabBegin, int slabEnd, void* deviceMinMax, void* stream )
    {
        Impl::GridCall c = EmptyCall( Impl::kGrid2D, seed );
        c.offset[0] = xOffset; c.offset[1] = yOffset;
        c.count[0] = xCount; c.count[1] = yCount;
        c.step[0] = xStepSize; c.step[1] = yStepSize;
        return SlabLaunch( node, c, 1, deviceOut, slabBegin, slabEnd, deviceMinMax, stream );
    }

    int fnb200GenUniformGrid3D( const void* node, float* deviceOut, float xOffset, float yOffset, float zOffset, int xCount, int yCount, int zCount,
                                float xStepSize, float yStepSize, float zStepSize, int seed, int slabBegin, int slabEnd, void* deviceMinMax, void* stream )
    {
        Impl::GridCall c = EmptyCall( Impl::kGrid3D, seed );
        c.offset[0] = xOffset; c.offset[1] = yOffset; c.offset[2] = zOffset;
        c.count[0] = xCount; c.count[1] = yCount; c.count[2] = zCount;
        c.step[0] = xStepSize; c.step[1] = yStepSize; c.step[2] = zStepSize;
        return SlabLaunch( node, c, 2, deviceOut, slabBegin, slabEnd, deviceMinMax, stream );
    }

    int fnb200GenUniformGrid3DShared( const void* node, void* sharedVolume, float xOffset, float yOffset, float zOffset, int xCount, int yCount, int zCount,
                                      float xStepSize, float yStepSize, float zStepSize, int seed, int slabBegin, int slabEnd, void* deviceMinMax, void* stream )
    {
        Impl::GridCall c = EmptyCall( Impl::kGrid3D, seed );
        c.offset[0] = xOffset; c.offset[1] = yOffset; c.offset[2] = zOffset;
        c.count[0] = xCount; c.count[1] = yCount; c.count[2] = zCount;
        c.step[0] = xStepSize; c.step[1] = yStepSize; c.step[2] = zStepSize;
        if( !sharedVolume )
        {
            Impl::SetLastError( "fnb200GenUniformGrid3DShared: null volume" );
            return FNB200_EINVAL;
        }
        float* placeholder = reinterpret_cast<float*>( sharedVolume ); // replaced by the slab's address inside the volume
        return SlabLaunch( node, c, 2, placeholder, slabBegin, slabEnd, deviceMinMax, stream, sharedVolume );
    }

    int fnb200GenUniformGrid4D( const void* node, float* deviceOut, float xOffset, float yOffset, float zOffset, float wOffset,
                                int xCount, int yCount, int zCount, int wCount, float xStepSize, float yStepSize, float zStepSize, float wStepSize,
                                int seed, int slabBegin, int slabEnd, void* deviceMinMax, void* stream )
    {
        Impl::GridCall c = EmptyCall( Impl::kGrid4D, seed );
        c.offset[0] = xOffset; c.offset[1] = yOffset; c.offset[2] = zOffset; c.offset[3] = wOffset;
        c.count[0] = xCount; c.count[1] = yCount; c.count[2] = zCount; c.count[3] = wCount;
        c.step[0] = xStepSize; c.step[1] = yStepSize; c.step[2] = zStepSize; c.step[3] = wStepSize;
        return SlabLaunch( node, c, 3, deviceOut, slabBegin, slabEnd, deviceMinMax, stream );
    }

    int fnb200GenTileable2D( const void* node, float* deviceOut, int xSize, int ySize, float xStepSize, float yStepSize,
                             int seed, int slabBegin, int slabEnd, void* deviceMinMax, void* stream )
    {
        Impl::GridCall c = EmptyCall( Impl::kTileable2D, seed );
        c.count[0] = xSize; c.count[1] = ySize;
        c.step[0] = xStepSize; c.step[1] = yStepSize;
        return SlabLaunch( node, c, 1, deviceOut, slabBegin, slabEnd, deviceMinMax, stream );
    }

    int fnb200GenPositionArray( const void* node, int dims, float* deviceOut, long long count,
                                const float* xPos, const float* yPos, const float* zPos, const float* wPos,
                                float xOffset, float yOffset, float zOffset, float wOffset, int seed, void* deviceMinMax, void* stream )
    {
        const float* pos[4] = { xPos, yPos, zPos, wPos };
        const float off[4] = { xOffset, yOffset, zOffset, wOffset };
        if( !node || !deviceOut || dims < 2 || dims > 4 || count < 0 )
        {
            Impl::SetLastError( "fnb200GenPositionArray: bad argument" );
            return FNB200_EINVAL;
        }
        for( int d = 0; d < dims; d++ )
        {
            if( !pos[d] )
            {
                Impl::SetLastError( "fnb200GenPositionArray: null position array" );
                return FNB200_EINVAL;
            }
        }
        return Impl::LaunchPositions( Gen( node ), dims, deviceOut, (unsigned long long)count, pos, off, seed, (unsigned*)deviceMinMax, StreamOf( stream ) );
    }

    int fnb200FillUniformPhilox( float* deviceOut, long long count, long long firstIndex, unsigned key, unsigned axis, float lo, float hi, void* stream )
    {
        if( !deviceOut || count < 0 ) return FNB200_EINVAL;
        return Impl::FillUniformPhilox( deviceOut, count, firstIndex, key, axis, lo, hi, StreamOf( stream ) );
    }

    long fnb200CanonicalEncodedNodeTree( const char* encodedString, char* buf, long bufSize )
    {
        std::vector<std::unique_ptr<NodeData>> records;
        NodeData* root = Metadata::DeserialiseNodeData( encodedString, records );
        std::string text = root ? Metadata::SerialiseNodeData( root, false ) : std::string();
        if( buf && bufSize > 0 )
        {
            long n = std::min<long>( bufSize - 1, (long)text.size() );
            std::memcpy( buf, text.data(), (size_t)n );
            buf[n] = 0;
        }
        return (long)text.size();
    }

    double fnb200MeasureAluPeak( int mode, int iters ) { return Impl::MeasureAluPeak( mode, iters ); }
    long long fnb200SelfTestInvSqrt( float lo, float hi ) { return Impl::SelfTestInvSqrt( lo, hi ); }
    int fnb200DebugIntermediates( int mode, int seed, float frequency, float jitter, int xCount, int yCount, int zCount, float xOffset, float yOffset, float zOffset, float step, int* hostOut )
    {
        return Impl::DebugIntermediates( mode, seed, frequency, jitter, xCount, yCount, zCount, xOffset, yOffset, zOffset, step, hostOut );
    }
    int fnb200SelfTestLeaf( int packed, int which, int n, int seed, float step, float* hostOut ) { return Impl::SelfTestLeaf( packed, which, n, seed, step, hostOut ); }
}
