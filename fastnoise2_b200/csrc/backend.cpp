// fastnoise2_b200 -- the B200 backend behind FastNoise::Generator::Gen*.
//
// What the reference does with FastSIMD::NewDispatchClass + virtual Gen() calls is done here in three
// steps: (1) lower the tree to CUDA source (lower.cpp), (2) compile it for sm_100a with NVRTC -- cached in
// memory and on disk by source hash, so a given tree shape/parameter set is compiled once per machine --
// and load the cubin as a CUDA library, (3) launch the entry kernel over the requested points.
// Host output / position pointers go through a double-buffered device staging pipeline so copies overlap
// the kernels; device pointers are used in place. There is no CPU evaluation path: if CUDA or NVRTC is
// unavailable every Gen* call fails (empty OutputMinMax, fnb200LastError() says why).
#include "backend.h"

#include <cuda_runtime.h>
#include <nvrtc.h>
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <unordered_map>
#include <vector>

#include "embedded_device_sources.h"

using namespace FastNoise;
using namespace FastNoise::Impl;

namespace
{
    constexpr int kBlock = 256;

    thread_local std::string tLastError;
    std::atomic<int> gMathMode { 0 };
    std::atomic<unsigned long long> gLaunches { 0 };

    bool Check( cudaError_t e, const char* what )
    {
        if( e == cudaSuccess ) return true;
        SetLastError( std::string( what ) + ": " + cudaGetErrorString( e ) );
        cudaGetLastError(); // clear the sticky-less error state
        return false;
    }

    std::uint64_t Fnv64( const std::string& text, std::uint64_t h = 1469598103934665603ull )
    {
        for( unsigned char c : text )
        {
            h ^= c;
            h *= 1099511628211ull;
        }
        return h;
    }

    int EnvInt( const char* name, int fallback )
    {
        const char* v = std::getenv( name );
        return v && *v ? std::atoi( v ) : fallback;
    }

    // ------------------------------------------------------------------ compiled-kernel cache
    struct Module
    {
        cudaLibrary_t library = nullptr;
        cudaKernel_t kernel = nullptr;
    };

    std::string CacheDir()
    {
        static const std::string dir = []
        {
            std::string d;
            if( const char* env = std::getenv( "FASTNOISE2_B200_CACHE" ) )
            {
                d = env;
            }
            else
            {
                Dl_info info;
                if( dladdr( (const void*)&CacheDir, &info ) && info.dli_fname )
                {
                    std::string lib = info.dli_fname;
                    std::size_t slash = lib.rfind( '/' );
                    d = ( slash == std::string::npos ? std::string( "." ) : lib.substr( 0, slash ) ) + "/_kcache";
                }
            }
            if( !d.empty() )
            {
                mkdir( d.c_str(), 0755 );
                if( access( d.c_str(), W_OK ) != 0 ) d.clear();
            }
            if( d.empty() )
            {
                const char* home = std::getenv( "HOME" );
                d = std::string( home ? home : "/tmp" ) + "/.cache";
                mkdir( d.c_str(), 0755 );
                d += "/fastnoise2_b200";
                mkdir( d.c_str(), 0755 );
            }
            return d;
        }();
        return dir;
    }

    std::vector<std::string> CompileOptions( int mathMode )
    {
        std::vector<std::string> o = {
            "--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device",
            "-DFN2_BLOCK=" + std::to_string( kBlock ),
            ( mathMode & 1 ) ? "--fmad=true" : "--fmad=false",
        };
        if( const char* extra = std::getenv( "FASTNOISE2_B200_NVRTC_FLAGS" ) )
        {
            std::istringstream in( extra );
            std::string flag;
            while( in >> flag ) o.push_back( flag );
        }
        return o;
    }

    // NVRTC: source -> cubin (no GPU needed). Empty vector + error text on failure.
    std::vector<char> CompileToCubin( const std::string& source, int mathMode, std::string& error )
    {
        const int n = kEmbeddedDeviceSourceCount;
        nvrtcProgram prog = nullptr;
        nvrtcResult r = nvrtcCreateProgram( &prog, source.c_str(), "fn2_tree.cu", n, kEmbeddedDeviceSourceText, kEmbeddedDeviceSourceName );
        if( r != NVRTC_SUCCESS )
        {
            error = std::string( "nvrtcCreateProgram: " ) + nvrtcGetErrorString( r );
            return {};
        }
        std::vector<std::string> opts = CompileOptions( mathMode );
        std::vector<const char*> optPtrs;
        for( const std::string& s : opts ) optPtrs.push_back( s.c_str() );
        r = nvrtcCompileProgram( prog, (int)optPtrs.size(), optPtrs.data() );
        if( r != NVRTC_SUCCESS )
        {
            std::size_t logSize = 0;
            nvrtcGetProgramLogSize( prog, &logSize );
            std::string log( logSize, '\0' );
            if( logSize ) nvrtcGetProgramLog( prog, &log[0] );
            error = std::string( "nvrtcCompileProgram: " ) + nvrtcGetErrorString( r ) + "\n" + log;
            nvrtcDestroyProgram( &prog );
            return {};
        }
        std::size_t size = 0;
        nvrtcGetCUBINSize( prog, &size );
        std::vector<char> cubin( size );
        if( size ) nvrtcGetCUBIN( prog, cubin.data() );
        nvrtcDestroyProgram( &prog );
        if( cubin.empty() ) error = "NVRTC produced an empty cubin";
        return cubin;
    }

    std::vector<char> CubinFor( const std::string& source, int mathMode, std::string& error )
    {
        std::string key = source;
        for( const std::string& o : CompileOptions( mathMode ) ) key += "\n//" + o;
        for( int i = 0; i < kEmbeddedDeviceSourceCount; i++ ) key += kEmbeddedDeviceSourceText[i];
        char name[64];
        std::snprintf( name, sizeof( name ), "/%016llx.cubin", (unsigned long long)Fnv64( key ) );
        const std::string path = CacheDir() + name;

        {
            std::ifstream in( path, std::ios::binary );
            if( in )
            {
                std::vector<char> cubin( ( std::istreambuf_iterator<char>( in ) ), std::istreambuf_iterator<char>() );
                if( !cubin.empty() ) return cubin;
            }
        }

        std::vector<char> cubin = CompileToCubin( source, mathMode, error );
        if( !cubin.empty() )
        {
            const std::string tmp = path + ".tmp" + std::to_string( (long)getpid() );
            std::ofstream outFile( tmp, std::ios::binary );
            if( outFile )
            {
                outFile.write( cubin.data(), (std::streamsize)cubin.size() );
                outFile.close();
                std::rename( tmp.c_str(), path.c_str() );
            }
        }
        return cubin;
    }

    std::mutex gModuleMutex;
    std::map<std::pair<std::uint64_t, int>, Module> gModules; // (source hash, device) -> loaded kernel

    const Module* LoadModule( const std::string& source, int mathMode, int device, std::string& error )
    {
        const std::uint64_t hash = Fnv64( source, Fnv64( "math mode " + std::to_string( mathMode ) ) );
        std::lock_guard<std::mutex> lock( gModuleMutex );
        auto it = gModules.find( { hash, device } );
        if( it != gModules.end() ) return &it->second;

        std::vector<char> cubin = CubinFor( source, mathMode, error );
        if( cubin.empty() ) return nullptr;

        Module m;
        cudaError_t e = cudaLibraryLoadData( &m.library, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0 );
        if( e != cudaSuccess )
        {
            error = std::string( "cudaLibraryLoadData: " ) + cudaGetErrorString( e );
            cudaGetLastError();
            return nullptr;
        }
        e = cudaLibraryGetKernel( &m.kernel, m.library, "fn2_entry" );
        if( e != cudaSuccess )
        {
            error = std::string( "cudaLibraryGetKernel: " ) + cudaGetErrorString( e );
            cudaGetLastError();
            cudaLibraryUnload( m.library );
            return nullptr;
        }
        return &gModules.emplace( std::make_pair( hash, device ), m ).first->second;
    }

    // ------------------------------------------------------------------ per-thread CUDA context
    struct ThreadContext
    {
        int device = -1;
        cudaStream_t compute = nullptr, copyOut = nullptr, copyIn = nullptr;
        cudaEvent_t computed[2] = {}, drained[2] = {}, uploaded[2] = {};
        float* outBuf[2] = {};
        std::size_t outCap = 0; // floats per buffer
        float* inBuf[2][4] = {};
        std::size_t inCap = 0;
        unsigned* minMax = nullptr;     // device accumulator
        unsigned* minMaxHost = nullptr; // pinned

        bool Ensure()
        {
            int current = 0;
            if( !Check( cudaGetDevice( &current ), "cudaGetDevice" ) ) return false;
            if( device == current && compute ) return true;
            Release();
            device = current;
            bool ok = Check( cudaStreamCreateWithFlags( &compute, cudaStreamNonBlocking ), "cudaStreamCreate" ) &&
                      Check( cudaStreamCreateWithFlags( &copyOut, cudaStreamNonBlocking ), "cudaStreamCreate" ) &&
                      Check( cudaStreamCreateWithFlags( &copyIn, cudaStreamNonBlocking ), "cudaStreamCreate" );
            for( int i = 0; ok && i < 2; i++ )
            {
                ok = Check( cudaEventCreateWithFlags( &computed[i], cudaEventDisableTiming ), "cudaEventCreate" ) &&
                     Check( cudaEventCreateWithFlags( &drained[i], cudaEventDisableTiming ), "cudaEventCreate" ) &&
                     Check( cudaEventCreateWithFlags( &uploaded[i], cudaEventDisableTiming ), "cudaEventCreate" );
            }
            ok = ok && Check( cudaMalloc( &minMax, 2 * sizeof( unsigned ) ), "cudaMalloc" ) &&
                 Check( cudaMallocHost( &minMaxHost, 2 * sizeof( unsigned ) ), "cudaMallocHost" );
            if( !ok ) Release();
            return ok;
        }

        bool EnsureOut( std::size_t floats )
        {
            if( floats <= outCap ) return true;
            for( float*& b : outBuf ) { if( b ) cudaFree( b ); b = nullptr; }
            outCap = 0;
            for( float*& b : outBuf ) if( !Check( cudaMalloc( &b, floats * sizeof( float ) ), "cudaMalloc(staging out)" ) ) return false;
            outCap = floats;
            return true;
        }

        bool EnsureIn( std::size_t floats, int dims )
        {
            if( floats <= inCap ) return true;
            for( auto& pair : inBuf ) for( float*& b : pair ) { if( b ) cudaFree( b ); b = nullptr; }
            inCap = 0;
            for( auto& pair : inBuf ) for( int d = 0; d < 4; d++ )
            {
                (void)dims;
                if( !Check( cudaMalloc( &pair[d], floats * sizeof( float ) ), "cudaMalloc(staging in)" ) ) return false;
            }
            inCap = floats;
            return true;
        }

        void Release()
        {
            if( device < 0 ) return;
            for( float*& b : outBuf ) { if( b ) cudaFree( b ); b = nullptr; }
            for( auto& pair : inBuf ) for( float*& b : pair ) { if( b ) cudaFree( b ); b = nullptr; }
            outCap = inCap = 0;
            if( minMax ) cudaFree( minMax );
            if( minMaxHost ) cudaFreeHost( minMaxHost );
            minMax = minMaxHost = nullptr;
            for( int i = 0; i < 2; i++ )
            {
                if( computed[i] ) cudaEventDestroy( computed[i] );
                if( drained[i] ) cudaEventDestroy( drained[i] );
                if( uploaded[i] ) cudaEventDestroy( uploaded[i] );
                computed[i] = drained[i] = uploaded[i] = nullptr;
            }
            if( compute ) cudaStreamDestroy( compute );
            if( copyOut ) cudaStreamDestroy( copyOut );
            if( copyIn ) cudaStreamDestroy( copyIn );
            compute = copyOut = copyIn = nullptr;
            device = -1;
            cudaGetLastError();
        }

        ~ThreadContext() { /* leaked on purpose at thread exit: the CUDA runtime may already be torn down */ }
    };

    thread_local ThreadContext tContext;

    bool IsDevicePointer( const void* p )
    {
        cudaPointerAttributes attr {};
        if( cudaPointerGetAttributes( &attr, p ) != cudaSuccess )
        {
            cudaGetLastError();
            return false;
        }
        return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
    }

    FastDivHost MakeFastDiv( unsigned d )
    {
        // q = ( umulhi( n, mul ) + n ) >> shift, exact for n < 2^31
        FastDivHost fd { 0, 0, d, 0 };
        if( d <= 1 ) return fd;
        unsigned shift = 0;
        while( ( 1ull << shift ) < d ) shift++;
        fd.shift = shift;
        fd.mul = (unsigned)( ( ( ( 1ull << shift ) - d ) << 32 ) / d + 1 );
        return fd;
    }

    float DecodeKey( unsigned key )
    {
        unsigned bits = ( key & 0x80000000u ) ? ( key ^ 0x80000000u ) : ~key;
        float v;
        std::memcpy( &v, &bits, 4 );
        return v;
    }
}

namespace FastNoise
{
    namespace Impl
    {
        void SetLastError( const std::string& message ) { tLastError = message; }
        const std::string& LastError() { return tLastError; }
        int MathMode() { return gMathMode.load(); }
        void SetMathMode( int mode ) { gMathMode.store( mode & 3 ); }
        unsigned long long LaunchCount() { return gLaunches.load(); }
        void CountLaunch() { gLaunches.fetch_add( 1, std::memory_order_relaxed ); }

        OutputMinMax DecodeMinMax( const unsigned keys[2] )
        {
            OutputMinMax mm;
            if( keys[0] ) mm.min = -DecodeKey( keys[0] );
            if( keys[1] ) mm.max = DecodeKey( keys[1] );
            return mm;
        }

        // --------------------------------------------------------------- per-root compiled programs
        struct CompiledTree
        {
            struct Variant
            {
                std::uint64_t stamp = 0;
                std::string source;
                std::string error;
                int mathMode = 0;
            };
            std::mutex mutex;
            std::map<int, Variant> variants; // key = kind * 16 + ppt

            static CompiledTree* Of( const Generator* root )
            {
                CompiledTree* tree = root->mCompiled.load( std::memory_order_acquire );
                if( tree ) return tree;
                CompiledTree* fresh = new CompiledTree;
                if( root->mCompiled.compare_exchange_strong( tree, fresh, std::memory_order_acq_rel ) ) return fresh;
                delete fresh;
                return tree;
            }
        };

        void ReleaseCompiled( CompiledTree* compiled ) { delete compiled; }

        // source text for (root, kind, ppt), re-lowered only when some node parameter changed since last time
        static bool SourceFor( const Generator* root, int kind, int ppt, std::string& source, int& mathMode )
        {
            CompiledTree* tree = CompiledTree::Of( root );
            std::lock_guard<std::mutex> lock( tree->mutex );
            CompiledTree::Variant& v = tree->variants[kind * 16 + ppt];
            const std::uint64_t stamp = TreeStamp();
            const int mode = MathMode();
            if( v.stamp != stamp || v.mathMode != mode || ( v.source.empty() && v.error.empty() ) )
            {
                LoweredProgram lowered = LowerTree( root, kind, ppt, mode );
                v.source = std::move( lowered.source );
                v.error = std::move( lowered.error );
                v.stamp = stamp;
                v.mathMode = mode;
            }
            if( !v.error.empty() )
            {
                SetLastError( "cannot lower node tree: " + v.error );
                return false;
            }
            source = v.source;
            mathMode = v.mathMode;
            return true;
        }

        std::string KernelSource( const Generator* root, int kind, int ppt )
        {
            std::string source;
            int mode;
            return SourceFor( root, kind, ppt, source, mode ) ? source : std::string();
        }

        int PrepareOffline( const Generator* root, int kind, int ppt )
        {
            std::string source, error;
            int mode;
            if( !SourceFor( root, kind, ppt, source, mode ) ) return -4;
            if( CubinFor( source, mode, error ).empty() )
            {
                SetLastError( error );
                return -3;
            }
            return 0;
        }

        // 0 or FNB200_E* (-2 CUDA, -3 compile, -4 unsupported tree)
        static int ModuleFor( const Generator* root, int kind, int ppt, const Module*& module )
        {
            std::string source, error;
            int mode;
            module = nullptr;
            if( !SourceFor( root, kind, ppt, source, mode ) ) return -4;
            int device = 0;
            if( !Check( cudaGetDevice( &device ), "cudaGetDevice" ) ) return -2;
            module = LoadModule( source, mode, device, error );
            if( module ) return 0;
            SetLastError( error );
            return error.compare( 0, 5, "nvrtc" ) == 0 || error.compare( 0, 5, "NVRTC" ) == 0 ? -3 : -2;
        }

        static int MaxPointsPerThread()
        {
            int p = EnvInt( "FASTNOISE2_B200_PPT", 4 );
            return p >= 8 ? 8 : ( p >= 4 ? 4 : ( p >= 2 ? 2 : 1 ) );
        }

        // Threads the current device keeps resident (SM count x 2048): below that a call cannot fill the GPU, so it is worth
        // trading the instruction-level parallelism of several points per thread for more threads (chunk-sized calls such
        // as 32^3 are latency-bound: one point per thread finishes ~3x sooner than four).
        static unsigned long long ResidentThreads()
        {
            static const unsigned long long value = []
            {
                int device = 0, sms = 148;
                if( cudaGetDevice( &device ) == cudaSuccess ) cudaDeviceGetAttribute( &sms, cudaDevAttrMultiProcessorCount, device );
                return (unsigned long long)( sms > 0 ? sms : 148 ) * 2048ull;
            }();
            return value;
        }

        static int PointsPerThreadFor( unsigned long long points, int ppt )
        {
            if( EnvInt( "FASTNOISE2_B200_FILL", 1 ) == 0 ) return ppt; // tests: keep the large-call code path for small calls
            while( ppt > 1 && points / (unsigned)ppt < ResidentThreads() ) ppt >>= 1;
            return ppt;
        }

        // --------------------------------------------------------------- launches (device pointers)
        static bool Launch( const Module* m, unsigned long long threads, void** args, cudaStream_t stream )
        {
            const unsigned long long blocks = ( threads + kBlock - 1 ) / kBlock;
            if( blocks == 0 ) return true;
            if( blocks > 0x7fffffffull )
            {
                SetLastError( "launch too large" );
                return false;
            }
            CountLaunch();
            return Check( cudaLaunchKernel( (const void*)m->kernel, dim3( (unsigned)blocks ), dim3( kBlock ), args, 0, stream ), "cudaLaunchKernel" );
        }

        int LaunchGridRows( const Generator* root, const GridCall& call, float* deviceOut, unsigned long long rowBegin, unsigned long long rowCount,
                            unsigned* deviceMinMax, cudaStream_t stream )
        {
            const unsigned nx = (unsigned)call.count[0];
            const unsigned long long points = rowCount * nx;
            if( points == 0 ) return 0;
            if( points >= 0x80000000ull || rowBegin + rowCount > 0xffffffffull )
            {
                SetLastError( "a single launch is limited to 2^31 - 1 points" );
                return -1;
            }
            int ppt = MaxPointsPerThread();
            while( ppt > 1 && nx % (unsigned)ppt != 0 ) ppt >>= 1;
            ppt = PointsPerThreadFor( points, ppt );

            const Module* m = nullptr;
            if( int rc = ModuleFor( root, call.kind, ppt, m ) ) return rc;

            int seed = call.seed;
            if( call.kind == kTileable2D )
            {
                TileArgsHost a {};
                // [Generator.inl:386-392], same float arithmetic
                const float pi2Recip = 0.15915493667f;
                const float xSizePi = (float)call.count[0] * pi2Recip;
                const float ySizePi = (float)call.count[1] * pi2Recip;
                a.xFreq = xSizePi * call.step[0];
                a.yFreq = ySizePi * call.step[1];
                a.xMul = 1 / xSizePi;
                a.yMul = 1 / ySizePi;
                a.xSize = call.count[0];
                a.ySize = call.count[1];
                a.rowBegin = (unsigned)rowBegin;
                a.total = (unsigned)points;
                a.vecStore = ( (std::uintptr_t)deviceOut % 16 ) == 0;
                a.divX = MakeFastDiv( nx );
                void* args[] = { &deviceOut, &a, &deviceMinMax, &seed };
                return Launch( m, ( points + ppt - 1 ) / ppt, args, stream ) ? 0 : -2;
            }

            GridArgsHost a {};
            for( int d = 0; d < 4; d++ )
            {
                a.off[d] = call.offset[d];
                a.step[d] = call.step[d];
                a.count[d] = call.count[d];
            }
            a.rowBegin = (unsigned)rowBegin;
            a.total = (unsigned)points;
            a.vecStore = ( (std::uintptr_t)deviceOut % 16 ) == 0;
            a.divX = MakeFastDiv( nx );
            a.divY = MakeFastDiv( (unsigned)call.count[1] );
            a.divZ = MakeFastDiv( (unsigned)call.count[2] );
            void* args[] = { &deviceOut, &a, &deviceMinMax, &seed };
            return Launch( m, ( points + ppt - 1 ) / ppt, args, stream ) ? 0 : -2;
        }

        int LaunchPositions( const Generator* root, int dims, float* deviceOut, unsigned long long count, const float* const pos[4],
                             const float offset[4], int seed, unsigned* deviceMinMax, cudaStream_t stream )
        {
            if( count == 0 ) return 0;
            const int ppt = PointsPerThreadFor( count, MaxPointsPerThread() > 4 ? 4 : MaxPointsPerThread() );
            const Module* m = nullptr;
            if( int rc = ModuleFor( root, 10 + dims, ppt, m ) ) return rc;

            PosArgsHost a {};
            bool aligned = ( (std::uintptr_t)deviceOut % 16 ) == 0;
            for( int d = 0; d < 4; d++ )
            {
                a.pos[d] = d < dims ? pos[d] : nullptr;
                a.off[d] = offset[d];
                if( d < dims ) aligned = aligned && ( (std::uintptr_t)pos[d] % 16 ) == 0;
            }
            a.count = count;
            a.vecIO = aligned;
            void* args[] = { &deviceOut, &a, &deviceMinMax, &seed };
            return Launch( m, ( count + ppt - 1 ) / ppt, args, stream ) ? 0 : -2;
        }

        // --------------------------------------------------------------- reference-shaped calls (host or device pointers)
        static std::size_t ChunkFloats()
        {
            // staging buffer size per pipeline slot; two slots are allocated per calling thread
            return (std::size_t)EnvInt( "FASTNOISE2_B200_STAGING_MB", 64 ) * 1024 * 1024 / sizeof( float );
        }

        static bool ReadMinMax( ThreadContext& ctx, OutputMinMax& mm )
        {
            if( !Check( cudaMemcpyAsync( ctx.minMaxHost, ctx.minMax, 2 * sizeof( unsigned ), cudaMemcpyDeviceToHost, ctx.compute ), "cudaMemcpyAsync(minmax)" ) ) return false;
            if( !Check( cudaStreamSynchronize( ctx.compute ), "cudaStreamSynchronize" ) ) return false;
            mm = DecodeMinMax( ctx.minMaxHost );
            return true;
        }

        OutputMinMax GenGrid( const Generator* root, const GridCall& call, float* out )
        {
            OutputMinMax failed;
            tLastError.clear(); // an all-NaN output also leaves min > max: callers tell the two apart by the error text
            const int dims = call.kind == kTileable2D ? 2 : call.kind;
            long long rows = 1;
            for( int d = 0; d < dims; d++ )
            {
                if( call.count[d] <= 0 )
                {
                    SetLastError( "grid size must be positive on every axis" );
                    return failed;
                }
                if( d ) rows *= call.count[d];
            }
            const long long nx = call.count[0];
            if( rows * nx >= 0x80000000ll )
            {
                // same limit as the reference, whose index arithmetic is 32-bit [Generator.inl:150]
                SetLastError( "grid has 2^31 or more points; use the slab entry points of FastNoise_B200.h" );
                return failed;
            }
            if( !out )
            {
                SetLastError( "output pointer is null" );
                return failed;
            }

            ThreadContext& ctx = tContext;
            if( !ctx.Ensure() ) return failed;
            if( !Check( cudaMemsetAsync( ctx.minMax, 0, 2 * sizeof( unsigned ), ctx.compute ), "cudaMemsetAsync" ) ) return failed;

            OutputMinMax mm;
            if( IsDevicePointer( out ) )
            {
                if( LaunchGridRows( root, call, out, 0, (unsigned long long)rows, ctx.minMax, ctx.compute ) != 0 ) return failed;
                return ReadMinMax( ctx, mm ) ? mm : failed;
            }

            // host destination: rows are generated chunk by chunk into two device buffers; the D2H copy of
            // chunk i runs on its own stream while chunk i+1 is being computed.
            const long long chunkRows = std::max<long long>( 1, std::min<long long>( rows, (long long)( ChunkFloats() / (std::size_t)nx ) ) );
            if( !ctx.EnsureOut( (std::size_t)( chunkRows * nx ) ) ) return failed;

            int slot = 0;
            bool used[2] = { false, false };
            for( long long row = 0; row < rows; row += chunkRows, slot ^= 1 )
            {
                const long long n = std::min( chunkRows, rows - row );
                if( used[slot] && !Check( cudaStreamWaitEvent( ctx.compute, ctx.drained[slot], 0 ), "cudaStreamWaitEvent" ) ) return failed;
                if( LaunchGridRows( root, call, ctx.outBuf[slot], (unsigned long long)row, (unsigned long long)n, ctx.minMax, ctx.compute ) != 0 ) return failed;
                if( !Check( cudaEventRecord( ctx.computed[slot], ctx.compute ), "cudaEventRecord" ) ||
                    !Check( cudaStreamWaitEvent( ctx.copyOut, ctx.computed[slot], 0 ), "cudaStreamWaitEvent" ) ||
                    !Check( cudaMemcpyAsync( out + row * nx, ctx.outBuf[slot], (std::size_t)( n * nx ) * sizeof( float ), cudaMemcpyDeviceToHost, ctx.copyOut ), "cudaMemcpyAsync(D2H)" ) ||
                    !Check( cudaEventRecord( ctx.drained[slot], ctx.copyOut ), "cudaEventRecord" ) )
                {
                    return failed;
                }
                used[slot] = true;
            }
            if( !ReadMinMax( ctx, mm ) ) return failed;
            if( !Check( cudaStreamSynchronize( ctx.copyOut ), "cudaStreamSynchronize" ) ) return failed;
            return mm;
        }

        OutputMinMax GenPositions( const Generator* root, int dims, float* out, long long count, const float* const pos[4], const float offset[4], int seed )
        {
            OutputMinMax failed;
            tLastError.clear();
            if( count <= 0 || !out )
            {
                SetLastError( count <= 0 ? "count must be positive" : "output pointer is null" );
                return failed;
            }
            for( int d = 0; d < dims; d++ )
            {
                if( !pos[d] )
                {
                    SetLastError( "position array is null" );
                    return failed;
                }
            }

            ThreadContext& ctx = tContext;
            if( !ctx.Ensure() ) return failed;
            if( !Check( cudaMemsetAsync( ctx.minMax, 0, 2 * sizeof( unsigned ), ctx.compute ), "cudaMemsetAsync" ) ) return failed;

            const bool outOnDevice = IsDevicePointer( out );
            bool inOnDevice = true;
            for( int d = 0; d < dims; d++ ) inOnDevice = inOnDevice && IsDevicePointer( pos[d] );

            OutputMinMax mm;
            if( outOnDevice && inOnDevice )
            {
                if( LaunchPositions( root, dims, out, (unsigned long long)count, pos, offset, seed, ctx.minMax, ctx.compute ) != 0 ) return failed;
                return ReadMinMax( ctx, mm ) ? mm : failed;
            }

            // staged pipeline: H2D of chunk i+1 (copyIn), kernel of chunk i (compute), D2H of chunk i-1 (copyOut)
            const long long chunk = std::min<long long>( count, (long long)( ChunkFloats() / 2 ) & ~3ll );
            if( !outOnDevice && !ctx.EnsureOut( (std::size_t)chunk ) ) return failed;
            if( !inOnDevice && !ctx.EnsureIn( (std::size_t)chunk, dims ) ) return failed;

            int slot = 0;
            bool used[2] = { false, false };
            for( long long first = 0; first < count; first += chunk, slot ^= 1 )
            {
                const long long n = std::min( chunk, count - first );
                const float* devPos[4] = { nullptr, nullptr, nullptr, nullptr };
                if( inOnDevice )
                {
                    for( int d = 0; d < dims; d++ ) devPos[d] = pos[d] + first;
                }
                else
                {
                    // the slot's input buffers are free once the kernel that last read them has finished
                    if( used[slot] && !Check( cudaStreamWaitEvent( ctx.copyIn, ctx.computed[slot], 0 ), "cudaStreamWaitEvent" ) ) return failed;
                    for( int d = 0; d < dims; d++ )
                    {
                        if( !Check( cudaMemcpyAsync( ctx.inBuf[slot][d], pos[d] + first, (std::size_t)n * sizeof( float ), cudaMemcpyHostToDevice, ctx.copyIn ), "cudaMemcpyAsync(H2D)" ) ) return failed;
                        devPos[d] = ctx.inBuf[slot][d];
                    }
                    if( !Check( cudaEventRecord( ctx.uploaded[slot], ctx.copyIn ), "cudaEventRecord" ) ||
                        !Check( cudaStreamWaitEvent( ctx.compute, ctx.uploaded[slot], 0 ), "cudaStreamWaitEvent" ) ) return failed;
                }

                float* devOut = outOnDevice ? out + first : ctx.outBuf[slot];
                if( !outOnDevice && used[slot] && !Check( cudaStreamWaitEvent( ctx.compute, ctx.drained[slot], 0 ), "cudaStreamWaitEvent" ) ) return failed;
                if( LaunchPositions( root, dims, devOut, (unsigned long long)n, devPos, offset, seed, ctx.minMax, ctx.compute ) != 0 ) return failed;
                if( !Check( cudaEventRecord( ctx.computed[slot], ctx.compute ), "cudaEventRecord" ) ) return failed;
                if( !outOnDevice )
                {
                    if( !Check( cudaStreamWaitEvent( ctx.copyOut, ctx.computed[slot], 0 ), "cudaStreamWaitEvent" ) ||
                        !Check( cudaMemcpyAsync( out + first, devOut, (std::size_t)n * sizeof( float ), cudaMemcpyDeviceToHost, ctx.copyOut ), "cudaMemcpyAsync(D2H)" ) ||
                        !Check( cudaEventRecord( ctx.drained[slot], ctx.copyOut ), "cudaEventRecord" ) ) return failed;
                }
                used[slot] = true;
            }
            if( !ReadMinMax( ctx, mm ) ) return failed;
            if( !Check( cudaStreamSynchronize( ctx.copyOut ), "cudaStreamSynchronize" ) ) return failed;
            if( !Check( cudaStreamSynchronize( ctx.copyIn ), "cudaStreamSynchronize" ) ) return failed;
            return mm;
        }

        cudaStream_t ThreadStream()
        {
            return tContext.Ensure() ? tContext.compute : nullptr;
        }

        bool DeviceAvailable()
        {
            int count = 0;
            if( cudaGetDeviceCount( &count ) != cudaSuccess || count == 0 )
            {
                cudaGetLastError();
                return false;
            }
            int device = 0, major = 0;
            if( cudaGetDevice( &device ) != cudaSuccess || cudaDeviceGetAttribute( &major, cudaDevAttrComputeCapabilityMajor, device ) != cudaSuccess )
            {
                cudaGetLastError();
                return false;
            }
            int nvrtcMajor = 0, nvrtcMinor = 0;
            return major == 10 && nvrtcVersion( &nvrtcMajor, &nvrtcMinor ) == NVRTC_SUCCESS;
        }
    }

    // ------------------------------------------------------------------ Generator virtuals
    static Impl::GridCall MakeGrid( int kind, int seed )
    {
        Impl::GridCall c {};
        c.kind = kind;
        c.seed = seed;
        for( int d = 0; d < 4; d++ ) { c.count[d] = 1; c.step[d] = 0; c.offset[d] = 0; }
        return c;
    }

    OutputMinMax Generator::GenUniformGrid2D( float* out, float xOffset, float yOffset, int xCount, int yCount, float xStepSize, float yStepSize, int seed ) const
    {
        Impl::GridCall c = MakeGrid( Impl::kGrid2D, seed );
        c.offset[0] = xOffset; c.offset[1] = yOffset;
        c.count[0] = xCount; c.count[1] = yCount;
        c.step[0] = xStepSize; c.step[1] = yStepSize;
        return Impl::GenGrid( this, c, out );
    }

    OutputMinMax Generator::GenUniformGrid3D( float* out, float xOffset, float yOffset, float zOffset, int xCount, int yCount, int zCount, float xStepSize, float yStepSize, float zStepSize, int seed ) const
    {
        Impl::GridCall c = MakeGrid( Impl::kGrid3D, seed );
        c.offset[0] = xOffset; c.offset[1] = yOffset; c.offset[2] = zOffset;
        c.count[0] = xCount; c.count[1] = yCount; c.count[2] = zCount;
        c.step[0] = xStepSize; c.step[1] = yStepSize; c.step[2] = zStepSize;
        return Impl::GenGrid( this, c, out );
    }

    OutputMinMax Generator::GenUniformGrid4D( float* out, float xOffset, float yOffset, float zOffset, float wOffset, int xCount, int yCount, int zCount, int wCount, float xStepSize, float yStepSize, float zStepSize, float wStepSize, int seed ) const
    {
        Impl::GridCall c = MakeGrid( Impl::kGrid4D, seed );
        c.offset[0] = xOffset; c.offset[1] = yOffset; c.offset[2] = zOffset; c.offset[3] = wOffset;
        c.count[0] = xCount; c.count[1] = yCount; c.count[2] = zCount; c.count[3] = wCount;
        c.step[0] = xStepSize; c.step[1] = yStepSize; c.step[2] = zStepSize; c.step[3] = wStepSize;
        return Impl::GenGrid( this, c, out );
    }

    OutputMinMax Generator::GenTileable2D( float* out, int xSize, int ySize, float xStepSize, float yStepSize, int seed ) const
    {
        Impl::GridCall c = MakeGrid( Impl::kTileable2D, seed );
        c.count[0] = xSize; c.count[1] = ySize;
        c.step[0] = xStepSize; c.step[1] = yStepSize;
        return Impl::GenGrid( this, c, out );
    }

    OutputMinMax Generator::GenPositionArray2D( float* out, int count, const float* xPosArray, const float* yPosArray, float xOffset, float yOffset, int seed ) const
    {
        const float* pos[4] = { xPosArray, yPosArray, nullptr, nullptr };
        const float off[4] = { xOffset, yOffset, 0, 0 };
        return Impl::GenPositions( this, 2, out, count, pos, off, seed );
    }

    OutputMinMax Generator::GenPositionArray3D( float* out, int count, const float* xPosArray, const float* yPosArray, const float* zPosArray, float xOffset, float yOffset, float zOffset, int seed ) const
    {
        const float* pos[4] = { xPosArray, yPosArray, zPosArray, nullptr };
        const float off[4] = { xOffset, yOffset, zOffset, 0 };
        return Impl::GenPositions( this, 3, out, count, pos, off, seed );
    }

    OutputMinMax Generator::GenPositionArray4D( float* out, int count, const float* xPosArray, const float* yPosArray, const float* zPosArray, const float* wPosArray, float xOffset, float yOffset, float zOffset, float wOffset, int seed ) const
    {
        const float* pos[4] = { xPosArray, yPosArray, zPosArray, wPosArray };
        const float off[4] = { xOffset, yOffset, zOffset, wOffset };
        return Impl::GenPositions( this, 4, out, count, pos, off, seed );
    }

    float Generator::GenSingle2D( float x, float y, int seed ) const
    {
        float out = 0;
        const float* pos[4] = { &x, &y, nullptr, nullptr };
        const float off[4] = { 0, 0, 0, 0 };
        Impl::GenPositions( this, 2, &out, 1, pos, off, seed );
        return out;
    }

    float Generator::GenSingle3D( float x, float y, float z, int seed ) const
    {
        float out = 0;
        const float* pos[4] = { &x, &y, &z, nullptr };
        const float off[4] = { 0, 0, 0, 0 };
        Impl::GenPositions( this, 3, &out, 1, pos, off, seed );
        return out;
    }

    float Generator::GenSingle4D( float x, float y, float z, float w, int seed ) const
    {
        float out = 0;
        const float* pos[4] = { &x, &y, &z, &w };
        const float off[4] = { 0, 0, 0, 0 };
        Impl::GenPositions( this, 4, &out, 1, pos, off, seed );
        return out;
    }
}
