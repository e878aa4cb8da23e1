// fastnoise2_b200 -- the two cross-GPU steps of the path, one process per GPU:
//   * OutputMinMax of a volume whose slabs were generated on several GPUs = one all-reduce (max) of the two order-preserving
//     unsigned keys of the device accumulator, enqueued on the generating stream -- no host round trip;
//   * the OPTIONAL assembly of the whole volume on every GPU (only when a caller wants one contiguous buffer) = an
//     all-gather of ragged slabs over NVLink.
// Both are NCCL collectives. NCCL is bound at run time (dlopen of libnccl.so.2 -- inside a PyTorch process that is the copy
// torch already loaded), so the library has no link-time dependency on it and single-GPU consumers never touch it.
// The reference has no counterpart: its model is "the caller splits one volume into calls" (Generator.h:166-167).
#include <FastNoise/FastNoise_B200.h>

#include <cuda_runtime_api.h>
#include <dlfcn.h>

#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "backend.h"

using namespace FastNoise;

namespace
{
    // the slice of nccl.h this file needs (NCCL 2.x ABI)
    typedef struct ncclComm* ncclComm_t;
    typedef struct { char internal[128]; } ncclUniqueId;
    enum { ncclSuccessValue = 0 };
    enum { ncclFloat32Value = 7, ncclUint32Value = 3 }; // ncclDataType_t
    enum { ncclMaxValue = 2 };                          // ncclRedOp_t

    struct Nccl
    {
        void* handle = nullptr;
        int ( *GetUniqueId )( ncclUniqueId* ) = nullptr;
        int ( *CommInitRank )( ncclComm_t*, int, ncclUniqueId, int ) = nullptr;
        int ( *CommDestroy )( ncclComm_t ) = nullptr;
        int ( *CommCount )( const ncclComm_t, int* ) = nullptr;
        int ( *CommUserRank )( const ncclComm_t, int* ) = nullptr;
        int ( *AllReduce )( const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t ) = nullptr;
        int ( *AllGather )( const void*, void*, size_t, int, ncclComm_t, cudaStream_t ) = nullptr;
        int ( *Broadcast )( const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t ) = nullptr;
        int ( *GroupStart )() = nullptr;
        int ( *GroupEnd )() = nullptr;
        const char* ( *GetErrorString )( int ) = nullptr;
        std::string error;
    };

    Nccl& Library()
    {
        static Nccl nccl;
        static std::once_flag once;
        std::call_once( once, []
        {
            for( const char* name : { "libnccl.so.2", "libnccl.so" } )
            {
                nccl.handle = dlopen( name, RTLD_NOW | RTLD_GLOBAL );
                if( nccl.handle ) break;
            }
            if( !nccl.handle )
            {
                nccl.error = std::string( "NCCL is not available: " ) + dlerror();
                return;
            }
            bool ok = true;
            auto bind = [&]( auto& fn, const char* symbol )
            {
                fn = reinterpret_cast<std::decay_t<decltype( fn )>>( dlsym( nccl.handle, symbol ) );
                if( !fn )
                {
                    ok = false;
                    nccl.error = std::string( "NCCL lacks " ) + symbol;
                }
            };
            bind( nccl.GetUniqueId, "ncclGetUniqueId" );
            bind( nccl.CommInitRank, "ncclCommInitRank" );
            bind( nccl.CommDestroy, "ncclCommDestroy" );
            bind( nccl.CommCount, "ncclCommCount" );
            bind( nccl.CommUserRank, "ncclCommUserRank" );
            bind( nccl.AllReduce, "ncclAllReduce" );
            bind( nccl.AllGather, "ncclAllGather" );
            bind( nccl.Broadcast, "ncclBroadcast" );
            bind( nccl.GroupStart, "ncclGroupStart" );
            bind( nccl.GroupEnd, "ncclGroupEnd" );
            bind( nccl.GetErrorString, "ncclGetErrorString" );
            if( !ok ) nccl.handle = nullptr;
        } );
        return nccl;
    }

    int Fail( const std::string& message, int code = FNB200_ECUDA )
    {
        Impl::SetLastError( message );
        return code;
    }

    int Rc( Nccl& nccl, int result, const char* what )
    {
        if( result == ncclSuccessValue ) return FNB200_OK;
        return Fail( std::string( what ) + ": " + ( nccl.GetErrorString ? nccl.GetErrorString( result ) : "NCCL error" ) );
    }
}

extern "C"
{
    int fnb200CommGetUniqueId( void* uniqueId128 )
    {
        Nccl& nccl = Library();
        if( !nccl.handle ) return Fail( nccl.error );
        if( !uniqueId128 ) return Fail( "fnb200CommGetUniqueId: null buffer", FNB200_EINVAL );
        return Rc( nccl, nccl.GetUniqueId( static_cast<ncclUniqueId*>( uniqueId128 ) ), "ncclGetUniqueId" );
    }

    void* fnb200CommInitRank( int worldSize, int rank, const void* uniqueId128 )
    {
        Nccl& nccl = Library();
        if( !nccl.handle )
        {
            Fail( nccl.error );
            return nullptr;
        }
        if( worldSize < 1 || rank < 0 || rank >= worldSize || !uniqueId128 )
        {
            Fail( "fnb200CommInitRank: bad world size, rank or id", FNB200_EINVAL );
            return nullptr;
        }
        ncclComm_t comm = nullptr;
        if( Rc( nccl, nccl.CommInitRank( &comm, worldSize, *static_cast<const ncclUniqueId*>( uniqueId128 ), rank ), "ncclCommInitRank" ) != FNB200_OK ) return nullptr;
        return comm;
    }

    int fnb200CommDestroy( void* comm )
    {
        Nccl& nccl = Library();
        if( !nccl.handle || !comm ) return FNB200_EINVAL;
        return Rc( nccl, nccl.CommDestroy( static_cast<ncclComm_t>( comm ) ), "ncclCommDestroy" );
    }

    int fnb200CommAllReduceMinMax( void* comm, void* deviceMinMax, void* stream )
    {
        Nccl& nccl = Library();
        if( !nccl.handle ) return Fail( nccl.error );
        if( !comm || !deviceMinMax ) return Fail( "fnb200CommAllReduceMinMax: null communicator or accumulator", FNB200_EINVAL );
        // the accumulator holds order-preserving keys of -min and +max (zero = empty): MAX over ranks reduces both
        return Rc( nccl, nccl.AllReduce( deviceMinMax, deviceMinMax, 2, ncclUint32Value, ncclMaxValue, static_cast<ncclComm_t>( comm ), (cudaStream_t)stream ), "ncclAllReduce" );
    }

    // ------------------------------------------------------------------ fused generate + all-gather over peer memory
    // A "shared volume" is one buffer of `floats` floats on every rank, each allocated by this library (cudaMalloc) and mapped
    // into every other rank through CUDA IPC, so a rank's generating kernel can store its slab into ALL copies as it produces it
    // (NVLink P2P stores): the all-gather costs no pass of its own. Collective: every rank of `comm` calls it with the same size.
    struct SharedVolume
    {
        int world = 0, rank = 0;
        long long floats = 0;
        float* copy[8] = {}; // copy[rank] is local, the others are peers' memory
    };

    void* fnb200CommAllocVolume( void* comm, long long floats, void* stream )
    {
        Nccl& nccl = Library();
        if( !nccl.handle ) { Fail( nccl.error ); return nullptr; }
        if( !comm || floats <= 0 ) { Fail( "fnb200CommAllocVolume: bad argument", FNB200_EINVAL ); return nullptr; }
        ncclComm_t c = static_cast<ncclComm_t>( comm );
        std::unique_ptr<SharedVolume> v( new SharedVolume );
        if( Rc( nccl, nccl.CommCount( c, &v->world ), "ncclCommCount" ) || Rc( nccl, nccl.CommUserRank( c, &v->rank ), "ncclCommUserRank" ) ) return nullptr;
        if( v->world > 8 ) { Fail( "fnb200CommAllocVolume: at most 8 ranks", FNB200_EINVAL ); return nullptr; }
        v->floats = floats;
        cudaStream_t s = (cudaStream_t)stream;
        cudaIpcMemHandle_t mine;
        cudaIpcMemHandle_t* gathered = nullptr; // device staging for the handle exchange
        std::vector<cudaIpcMemHandle_t> all( v->world );
        bool ok = cudaMalloc( (void**)&v->copy[v->rank], (size_t)floats * sizeof( float ) ) == cudaSuccess &&
                  cudaIpcGetMemHandle( &mine, v->copy[v->rank] ) == cudaSuccess &&
                  cudaMalloc( (void**)&gathered, sizeof( mine ) * v->world ) == cudaSuccess &&
                  cudaMemcpyAsync( gathered + v->rank, &mine, sizeof( mine ), cudaMemcpyHostToDevice, s ) == cudaSuccess;
        if( ok ) ok = Rc( nccl, nccl.AllGather( gathered + v->rank, gathered, sizeof( mine ), 0 /* ncclInt8 */, c, s ), "ncclAllGather(ipc handles)" ) == FNB200_OK;
        ok = ok && cudaMemcpyAsync( all.data(), gathered, sizeof( mine ) * v->world, cudaMemcpyDeviceToHost, s ) == cudaSuccess && cudaStreamSynchronize( s ) == cudaSuccess;
        for( int r = 0; ok && r < v->world; r++ )
        {
            if( r != v->rank ) ok = cudaIpcOpenMemHandle( (void**)&v->copy[r], all[r], cudaIpcMemLazyEnablePeerAccess ) == cudaSuccess;
        }
        if( gathered ) cudaFree( gathered );
        if( !ok )
        {
            Fail( std::string( "fnb200CommAllocVolume: " ) + cudaGetErrorString( cudaGetLastError() ) );
            for( int r = 0; r < v->world; r++ ) if( v->copy[r] && r != v->rank ) cudaIpcCloseMemHandle( v->copy[r] );
            if( v->copy[v->rank] ) cudaFree( v->copy[v->rank] );
            return nullptr;
        }
        return v.release();
    }

    float* fnb200VolumeData( void* volume ) { return volume ? static_cast<SharedVolume*>( volume )->copy[static_cast<SharedVolume*>( volume )->rank] : nullptr; }

    int fnb200CommFreeVolume( void* comm, void* volume, void* stream )
    {
        Nccl& nccl = Library();
        if( !nccl.handle || !comm || !volume ) return FNB200_EINVAL;
        std::unique_ptr<SharedVolume> v( static_cast<SharedVolume*>( volume ) );
        for( int r = 0; r < v->world; r++ ) if( r != v->rank && v->copy[r] ) cudaIpcCloseMemHandle( v->copy[r] );
        // nobody may still be writing into my copy: a barrier (1-element all-reduce) before the owner frees it
        unsigned* token = nullptr;
        if( cudaMalloc( (void**)&token, sizeof( unsigned ) ) == cudaSuccess )
        {
            cudaMemsetAsync( token, 0, sizeof( unsigned ), (cudaStream_t)stream );
            nccl.AllReduce( token, token, 1, ncclUint32Value, ncclMaxValue, static_cast<ncclComm_t>( comm ), (cudaStream_t)stream );
            cudaStreamSynchronize( (cudaStream_t)stream );
            cudaFree( token );
        }
        cudaFree( v->copy[v->rank] );
        cudaGetLastError();
        return FNB200_OK;
    }

    // Elements [firstFloat, firstFloat + n) of every copy: where a launch that starts at `firstFloat` of the volume must store
    int fnb200VolumeMirrors( void* volume, long long firstFloat, float** local, float** mirrors /* [7] */, int* mirrorCount )
    {
        if( !volume || !local || !mirrors || !mirrorCount ) return FNB200_EINVAL;
        SharedVolume* v = static_cast<SharedVolume*>( volume );
        if( firstFloat < 0 || firstFloat > v->floats ) return FNB200_EINVAL;
        *local = v->copy[v->rank] + firstFloat;
        *mirrorCount = 0;
        for( int r = 0; r < v->world; r++ ) if( r != v->rank ) mirrors[( *mirrorCount )++] = v->copy[r] + firstFloat;
        return FNB200_OK;
    }

    int fnb200CommGatherSlabs( void* comm, const float* deviceSlab, float* deviceVolume, const long long* floatsPerRank, void* stream )
    {
        Nccl& nccl = Library();
        if( !nccl.handle ) return Fail( nccl.error );
        if( !comm || !deviceVolume || !floatsPerRank ) return Fail( "fnb200CommGatherSlabs: null argument", FNB200_EINVAL );
        ncclComm_t c = static_cast<ncclComm_t>( comm );
        int world = 0, rank = 0;
        if( int rc = Rc( nccl, nccl.CommCount( c, &world ), "ncclCommCount" ) ) return rc;
        if( int rc = Rc( nccl, nccl.CommUserRank( c, &rank ), "ncclCommUserRank" ) ) return rc;
        bool equal = true;
        long long offsetOfMine = 0;
        for( int r = 0; r < world; r++ )
        {
            if( floatsPerRank[r] < 0 ) return Fail( "fnb200CommGatherSlabs: negative slab size", FNB200_EINVAL );
            equal = equal && floatsPerRank[r] == floatsPerRank[0];
            if( r < rank ) offsetOfMine += floatsPerRank[r];
        }
        // a slab generated in place (straight into its position of the volume) is sent from there
        const float* mine = deviceSlab ? deviceSlab : deviceVolume + offsetOfMine;
        if( equal ) return Rc( nccl, nccl.AllGather( mine, deviceVolume, (size_t)floatsPerRank[0], ncclFloat32Value, c, (cudaStream_t)stream ), "ncclAllGather" );

        // ragged slabs (the outermost axis does not divide by the number of ranks): one broadcast per slab, fused into one group
        if( int rc = Rc( nccl, nccl.GroupStart(), "ncclGroupStart" ) ) return rc;
        long long offset = 0;
        int first = FNB200_OK;
        for( int r = 0; r < world; r++ )
        {
            const int rc = Rc( nccl, nccl.Broadcast( r == rank ? mine : deviceVolume + offset, deviceVolume + offset, (size_t)floatsPerRank[r], ncclFloat32Value, r, c, (cudaStream_t)stream ), "ncclBroadcast" );
            if( first == FNB200_OK ) first = rc;
            offset += floatsPerRank[r];
        }
        const int end = Rc( nccl, nccl.GroupEnd(), "ncclGroupEnd" );
        return first != FNB200_OK ? first : end;
    }
}
