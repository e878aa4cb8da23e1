// fastnoise2_b200 -- the B200 backend behind FastNoise::Generator::Gen*.
//
// What the reference does with FastSIMD::NewDispatchClass + virtual Gen() calls is done here in three
// steps: (1) lower the tree to CUDA source (lower.cpp) -- float node parameters become kernel arguments, so the text
// depends on the SHAPE of the tree only, (2) compile it for sm_100a with NVRTC -- cached in memory and on disk by
// source hash, so a given tree shape is compiled once per machine -- and load the cubin as a CUDA library,
// (3) launch the persistent entry kernel over the requested points with the tree's current parameter values.
// Host output / position pointers go through a double-buffered device staging pipeline so copies overlap
// the kernels -- split over every visible GPU for large calls, each GPU draining its rows over its own PCIe link --
// and device pointers are used in place. There is no CPU evaluation path: if CUDA or NVRTC is
// unavailable every Gen* call fails (empty OutputMinMax, fnb200LastError() says why).
#include "backend.h"

#include <cuda_runtime.h>
#include <nvrtc.h>
#include <dlfcn.h>
#if defined( __x86_64__ )
#include <immintrin.h>
#endif
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <thread>
#include <unordered_map>
#include <vector>

#include "embedded_device_sources.h"

using namespace FastNoise;
using namespace FastNoise::Impl;

namespace
{
    constexpr int kDefaultBlock = 256;
    constexpr int kItersPerThread = 8; // groups one thread walks in a large call (see Launch)
    constexpr std::size_t kConvergedCodeBytes = 48 * 1024; // kernels larger than this walk one group per thread (see Launch)

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

    // Threads per block of the entry kernels (FN2_BLOCK in the generated text). 256 from measurements
    // (profiles/r2_perf_sweep_c_code_shape.txt); FASTNOISE2_B200_BLOCK, read once, overrides it for A/B runs.
    int BlockThreads()
    {
        static const int value = []
        {
            const int v = EnvInt( "FASTNOISE2_B200_BLOCK", kDefaultBlock );
            return v == 64 || v == 128 || v == 512 || v == 1024 ? v : kDefaultBlock;
        }();
        return value;
    }

    // ------------------------------------------------------------------ compiled-kernel cache
    struct Module
    {
        cudaLibrary_t library = nullptr;
        cudaKernel_t kernel = nullptr;
        int residentBlocks = 0; // blocks of BlockThreads() threads the whole device keeps resident: the persistent launch size
        std::size_t codeBytes = 0; // size of fn2_entry's machine code: what the instruction caches have to hold
    };

    // Size of the section that holds `kernel`'s code in a cubin (ELF64): .text.<kernel>. 0 if it cannot be found.
    std::size_t KernelCodeBytes( const std::vector<char>& cubin, const char* kernel )
    {
        struct Header { unsigned char ident[16]; std::uint16_t type, machine; std::uint32_t version; std::uint64_t entry, phoff, shoff; std::uint32_t flags; std::uint16_t ehsize, phentsize, phnum, shentsize, shnum, shstrndx; };
        struct Section { std::uint32_t name, type; std::uint64_t flags, addr, offset, size; std::uint32_t link, info; std::uint64_t addralign, entsize; };
        if( cubin.size() < sizeof( Header ) || std::memcmp( cubin.data(), "\177ELF", 4 ) != 0 || cubin[4] != 2 ) return 0;
        Header h;
        std::memcpy( &h, cubin.data(), sizeof( h ) );
        if( h.shentsize != sizeof( Section ) || h.shoff + (std::uint64_t)h.shnum * sizeof( Section ) > cubin.size() || h.shstrndx >= h.shnum ) return 0;
        auto section = [&]( unsigned i ) { Section s; std::memcpy( &s, cubin.data() + h.shoff + (std::uint64_t)i * sizeof( Section ), sizeof( s ) ); return s; };
        const Section names = section( h.shstrndx );
        if( names.offset + names.size > cubin.size() ) return 0;
        const std::string want = std::string( ".text." ) + kernel;
        for( unsigned i = 0; i < h.shnum; i++ )
        {
            const Section s = section( i );
            if( s.name >= names.size ) continue;
            const char* text = cubin.data() + names.offset + s.name;
            if( std::strncmp( text, want.c_str(), names.size - s.name ) == 0 ) return (std::size_t)s.size;
        }
        return 0;
    }

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
            "--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device", "--ptxas-options=-v",
            "-DFN2_BLOCK=" + std::to_string( BlockThreads() ),
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

    std::vector<char> CompileOnce( const std::string& source, int mathMode, const char* extraOption, std::string& error, int* registers );

    // NVRTC: source -> cubin (no GPU needed). Empty vector + error text on failure.
    // Kernels are compiled with __launch_bounds__( 256 ) and no minimum block count: a minimum is an occupancy TARGET that costs the
    // light trees 3-7 % (profiles/r2_perf_sweep_b_launch_geometry.txt). Left alone, though, ptxas takes 145 registers for SuperSimplex
    // 4D x 4 points -- one block per SM, -18 % -- so a kernel that comes out above 128 registers is compiled once more with a minimum
    // of two blocks per SM (a 128-register cap).
    std::vector<char> CompileToCubin( const std::string& source, int mathMode, std::string& error )
    {
        int registers = 0;
        std::vector<char> cubin = CompileOnce( source, mathMode, nullptr, error, &registers );
        if( !cubin.empty() && registers > 128 )
        {
            std::string capError;
            std::vector<char> capped = CompileOnce( source, mathMode, "-DFN2_LAUNCH_BOUNDS=__launch_bounds__(FN2_BLOCK,2)", capError, nullptr );
            if( !capped.empty() ) return capped;
        }
        return cubin;
    }

    std::vector<char> CompileOnce( const std::string& source, int mathMode, const char* extraOption, std::string& error, int* registers )
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
        if( extraOption ) opts.push_back( extraOption );
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
        if( registers )
        {
            // "ptxas info    : Used 64 registers, ..." for fn2_entry, the only kernel of the program
            std::size_t logSize = 0;
            nvrtcGetProgramLogSize( prog, &logSize );
            std::string log( logSize, '\0' );
            if( logSize ) nvrtcGetProgramLog( prog, &log[0] );
            const std::size_t at = log.find( "Used " );
            *registers = at == std::string::npos ? 0 : std::atoi( log.c_str() + at + 5 );
        }
        std::size_t size = 0;
        nvrtcGetCUBINSize( prog, &size );
        std::vector<char> cubin( size );
        if( size ) nvrtcGetCUBIN( prog, cubin.data() );
        nvrtcDestroyProgram( &prog );
        if( cubin.empty() ) error = "NVRTC produced an empty cubin";
        return cubin;
    }

    std::uint64_t LibraryTextHash()
    {
        static const std::uint64_t h = []
        {
            std::uint64_t v = 1469598103934665603ull;
            for( int i = 0; i < kEmbeddedDeviceSourceCount; i++ ) v = Fnv64( kEmbeddedDeviceSourceText[i], v );
            return v;
        }();
        return h;
    }

    std::vector<char> CubinFor( const std::string& source, int mathMode, std::string& error )
    {
        std::uint64_t key = Fnv64( source, LibraryTextHash() );
        for( const std::string& o : CompileOptions( mathMode ) ) key = Fnv64( "\n//" + o, key );
        char name[64];
        std::snprintf( name, sizeof( name ), "/%016llx.cubin", (unsigned long long)key );
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

    // the current device must be `device`
    const Module* LoadModule( const std::string& source, std::uint64_t sourceHash, int mathMode, int device, std::string& error )
    {
        std::lock_guard<std::mutex> lock( gModuleMutex );
        auto it = gModules.find( { sourceHash, device } );
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
        int perSm = 0, sms = 0;
        if( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &perSm, (const void*)m.kernel, BlockThreads(), 0 ) != cudaSuccess || perSm < 1 )
        {
            cudaGetLastError();
            perSm = 2;
        }
        if( cudaDeviceGetAttribute( &sms, cudaDevAttrMultiProcessorCount, device ) != cudaSuccess || sms < 1 )
        {
            cudaGetLastError();
            sms = 148;
        }
        m.residentBlocks = perSm * sms;
        m.codeBytes = KernelCodeBytes( cubin, "fn2_entry" );
        return &gModules.emplace( std::make_pair( sourceHash, device ), m ).first->second;
    }

    // ------------------------------------------------------------------ host-side copy workers (pageable destinations)
    // cudaMemcpyAsync into pageable memory is staged by the driver through one small pinned buffer and runs at a fraction of the
    // link rate. Pageable destinations therefore go device -> pinned bounce buffer (DMA at link rate) -> destination, the second
    // hop done by a few worker threads while the next chunk is computed and transferred.
    // bounce buffer -> caller's pageable memory. The destination is written once and not read here, so it is written with
    // non-temporal stores: no read-for-ownership of the destination lines and the bounce buffer stays the only thing in the caches
    // (glibc's memcpy only switches to streaming stores above a size threshold that the per-worker slices do not always reach).
#if defined( __x86_64__ )
    __attribute__(( target( "avx2" ) )) void StreamCopyAvx2( char* dst, const char* src, std::size_t bytes )
    {
        std::size_t head = ( 32 - ( (std::uintptr_t)dst & 31 ) ) & 31;
        if( head > bytes ) head = bytes;
        std::memcpy( dst, src, head );
        dst += head; src += head; bytes -= head;
        std::size_t i = 0;
        for( ; i + 128 <= bytes; i += 128 )
        {
            const __m256i a = _mm256_loadu_si256( (const __m256i*)( src + i ) ), b = _mm256_loadu_si256( (const __m256i*)( src + i + 32 ) );
            const __m256i c = _mm256_loadu_si256( (const __m256i*)( src + i + 64 ) ), d = _mm256_loadu_si256( (const __m256i*)( src + i + 96 ) );
            _mm256_stream_si256( (__m256i*)( dst + i ), a );
            _mm256_stream_si256( (__m256i*)( dst + i + 32 ), b );
            _mm256_stream_si256( (__m256i*)( dst + i + 64 ), c );
            _mm256_stream_si256( (__m256i*)( dst + i + 96 ), d );
        }
        _mm_sfence();
        std::memcpy( dst + i, src + i, bytes - i );
    }
#endif

    // `streaming`: the chunk this slice belongs to is too large to stay in the last-level cache anyway. Small destinations (a
    // 128^3 chunk is 8 MB) are copied with plain stores: they are usually cache-resident -- the caller reads them next -- and
    // non-temporal stores to cached lines are slower (128^3 pageable: 855 us streamed, 435 us plain; profiles/r2_call_latency.txt)
    void CopyToPageable( char* dst, const char* src, std::size_t bytes, bool streaming )
    {
#if defined( __x86_64__ )
        static const bool stream = __builtin_cpu_supports( "avx2" ) && EnvInt( "FASTNOISE2_B200_STREAM_COPY", 1 ) != 0;
        if( stream && streaming && bytes >= 4096 )
        {
            StreamCopyAvx2( dst, src, bytes );
            return;
        }
#endif
        std::memcpy( dst, src, bytes );
    }

    class CopyPool
    {
    public:
        struct Ticket
        {
            std::mutex mutex;
            std::condition_variable done;
            int pending = 0;
            void Wait()
            {
                std::unique_lock<std::mutex> lock( mutex );
                done.wait( lock, [this] { return pending == 0; } );
            }
        };

        static CopyPool& Get()
        {
            static CopyPool* pool = new CopyPool; // never destroyed: worker threads may outlive static destruction
            return *pool;
        }

        int Threads() const { return (int)mWorkers.size(); }

        // after `ready` (an event recorded behind the DMA into `src`) has completed, copy bytes src -> dst on the workers
        void Submit( cudaEvent_t ready, const char* src, char* dst, std::size_t bytes, Ticket& ticket )
        {
            const int parts = (int)std::max<std::size_t>( 1, std::min<std::size_t>( mWorkers.size(), bytes / ( 1u << 20 ) ) );
            const bool streaming = bytes >= ( 32u << 20 );
            {
                std::lock_guard<std::mutex> lock( ticket.mutex );
                ticket.pending += parts;
            }
            const std::size_t slice = ( ( bytes / parts ) + 63 ) & ~std::size_t( 63 );
            std::lock_guard<std::mutex> lock( mMutex );
            for( int p = 0; p < parts; p++ )
            {
                const std::size_t begin = std::min( bytes, slice * p ), end = p == parts - 1 ? bytes : std::min( bytes, slice * ( p + 1 ) );
                mJobs.push_back( [=, &ticket]
                {
                    cudaEventSynchronize( ready );
                    if( end > begin ) CopyToPageable( dst + begin, src + begin, end - begin, streaming );
                    std::lock_guard<std::mutex> l( ticket.mutex );
                    if( --ticket.pending == 0 ) ticket.done.notify_all();
                } );
            }
            mWake.notify_all();
        }

    private:
        CopyPool()
        {
            int n = EnvInt( "FASTNOISE2_B200_COPY_THREADS", 0 );
            if( n <= 0 ) n = (int)std::min<unsigned>( 16, std::max<unsigned>( 2, std::thread::hardware_concurrency() * 3 / 4 ) ); // 12 of 16 cores measured best (profiles/r2_pageable_probe.txt)
            for( int i = 0; i < n; i++ )
            {
                mWorkers.emplace_back( [this]
                {
                    for( ;; )
                    {
                        std::function<void()> job;
                        {
                            std::unique_lock<std::mutex> lock( mMutex );
                            mWake.wait( lock, [this] { return !mJobs.empty(); } );
                            job = std::move( mJobs.front() );
                            mJobs.pop_front();
                        }
                        job();
                    }
                } );
                mWorkers.back().detach();
            }
        }

        std::mutex mMutex;
        std::condition_variable mWake;
        std::deque<std::function<void()>> mJobs;
        std::vector<std::thread> mWorkers;
    };

    // ------------------------------------------------------------------ per-thread, per-device CUDA resources
    struct DeviceContext
    {
        int device = -1;
        cudaStream_t compute = nullptr, copyOut = nullptr, copyIn = nullptr;
        cudaEvent_t computed[2] = {}, drained[2] = {}, uploaded[2] = {}, callerDone = nullptr;
        float* outBuf[2] = {};
        std::size_t outCap = 0; // floats per buffer
        float* inBuf[2][4] = {};
        std::size_t inCap = 0;
        int inDims = 0;
        float* bounce[2] = {}; // pinned host, pageable destinations only
        std::size_t bounceCap = 0;
        CopyPool::Ticket bounceBusy[2];
        unsigned* minMax = nullptr;     // device accumulator
        unsigned* minMaxHost = nullptr; // pinned

        // the current device must be `device`
        bool Ensure( int dev )
        {
            if( device == dev && compute ) return true;
            Release();
            device = dev;
            bool ok = Check( cudaStreamCreateWithFlags( &compute, cudaStreamNonBlocking ), "cudaStreamCreate" ) &&
                      Check( cudaStreamCreateWithFlags( &copyOut, cudaStreamNonBlocking ), "cudaStreamCreate" ) &&
                      Check( cudaStreamCreateWithFlags( &copyIn, cudaStreamNonBlocking ), "cudaStreamCreate" ) &&
                      Check( cudaEventCreateWithFlags( &callerDone, cudaEventDisableTiming ), "cudaEventCreate" );
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
            if( floats <= inCap && dims <= inDims ) return true;
            floats = std::max( floats, inCap );
            dims = std::max( dims, inDims );
            for( auto& pair : inBuf ) for( float*& b : pair ) { if( b ) cudaFree( b ); b = nullptr; }
            inCap = 0;
            inDims = 0;
            for( auto& pair : inBuf ) for( int d = 0; d < dims; d++ )
            {
                if( !Check( cudaMalloc( &pair[d], floats * sizeof( float ) ), "cudaMalloc(staging in)" ) ) return false;
            }
            inCap = floats;
            inDims = dims;
            return true;
        }

        bool EnsureBounce( std::size_t floats )
        {
            if( floats <= bounceCap ) return true;
            for( float*& b : bounce ) { if( b ) cudaFreeHost( b ); b = nullptr; }
            bounceCap = 0;
            for( float*& b : bounce ) if( !Check( cudaMallocHost( &b, floats * sizeof( float ) ), "cudaMallocHost(bounce)" ) ) return false;
            bounceCap = floats;
            return true;
        }

        void Release()
        {
            if( device < 0 ) return;
            int previous = -1;
            if( cudaGetDevice( &previous ) == cudaSuccess && previous != device ) cudaSetDevice( device );
            for( CopyPool::Ticket& t : bounceBusy ) t.Wait();
            for( float*& b : outBuf ) { if( b ) cudaFree( b ); b = nullptr; }
            for( auto& pair : inBuf ) for( float*& b : pair ) { if( b ) cudaFree( b ); b = nullptr; }
            for( float*& b : bounce ) { if( b ) cudaFreeHost( b ); b = nullptr; }
            outCap = inCap = bounceCap = 0;
            inDims = 0;
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
            if( callerDone ) cudaEventDestroy( callerDone );
            callerDone = nullptr;
            if( compute ) cudaStreamDestroy( compute );
            if( copyOut ) cudaStreamDestroy( copyOut );
            if( copyIn ) cudaStreamDestroy( copyIn );
            compute = copyOut = copyIn = nullptr;
            device = -1;
            if( previous >= 0 ) cudaSetDevice( previous );
            cudaGetLastError();
        }
    };

    // One DeviceContext per device this thread has used. Released when the thread exits: a consumer that spawns short-lived worker
    // threads must not leak staging memory per thread. At process exit the CUDA runtime may already be unloading -- every call in
    // Release() then fails with cudaErrorCudartUnloading, which is ignored.
    struct ThreadContexts
    {
        std::vector<std::unique_ptr<DeviceContext>> perDevice;

        // makes `device` current
        DeviceContext* Get( int device )
        {
            if( !Check( cudaSetDevice( device ), "cudaSetDevice" ) ) return nullptr;
            if( (std::size_t)device >= perDevice.size() ) perDevice.resize( device + 1 );
            if( !perDevice[device] ) perDevice[device].reset( new DeviceContext );
            return perDevice[device]->Ensure( device ) ? perDevice[device].get() : nullptr;
        }

        ~ThreadContexts()
        {
            for( auto& c : perDevice ) if( c ) c->Release();
        }
    };

    thread_local ThreadContexts tContexts;

    enum class Memory { Device, Pinned, Pageable };

    Memory Classify( const void* p )
    {
        cudaPointerAttributes attr {};
        if( cudaPointerGetAttributes( &attr, p ) != cudaSuccess )
        {
            cudaGetLastError();
            return Memory::Pageable;
        }
        if( attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged ) return Memory::Device;
        return attr.type == cudaMemoryTypeHost ? Memory::Pinned : Memory::Pageable;
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

        static bool ParamMode()
        {
            return EnvInt( "FASTNOISE2_B200_LITERALS", 0 ) == 0;
        }

        // --------------------------------------------------------------- per-root compiled programs
        struct CompiledTree
        {
            struct Variant
            {
                std::uint64_t stamp = 0;
                std::string source;
                std::uint64_t sourceHash = 0;
                std::vector<float> params;
                std::string error;
                int mathMode = 0;
                bool paramMode = false;
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

        // What one launch needs of (root, kind, ppt): the program text (re-lowered only when some node parameter changed since
        // last time; a float edit gives the same text back, so the loaded kernel is found again by its hash) and the values
        // of the kernel's TreeParams argument.
        struct Program
        {
            std::string source;
            std::uint64_t sourceHash = 0;
            std::vector<float> params;
            int mathMode = 0;
        };

        static bool ProgramFor( const Generator* root, int kind, int ppt, Program& program, bool wantSource )
        {
            CompiledTree* tree = CompiledTree::Of( root );
            std::lock_guard<std::mutex> lock( tree->mutex );
            CompiledTree::Variant& v = tree->variants[kind * 16 + ppt];
            const std::uint64_t stamp = TreeStamp();
            const int mode = MathMode();
            const bool paramMode = ParamMode();
            if( v.stamp != stamp || v.mathMode != mode || v.paramMode != paramMode || ( v.source.empty() && v.error.empty() ) )
            {
                LoweredProgram lowered = LowerTree( root, kind, ppt, mode, paramMode );
                if( lowered.source != v.source || v.mathMode != mode )
                {
                    v.source = std::move( lowered.source );
                    // (extra NVRTC flags are part of the key: the loaded-kernel table must not hand back a kernel built with others)
                    const char* extraFlags = std::getenv( "FASTNOISE2_B200_NVRTC_FLAGS" );
                    v.sourceHash = Fnv64( v.source, Fnv64( "math mode " + std::to_string( mode ) + " flags " + ( extraFlags ? extraFlags : "" ) + " block " + std::to_string( BlockThreads() ) ) );
                }
                v.params = std::move( lowered.params );
                v.error = std::move( lowered.error );
                v.stamp = stamp;
                v.mathMode = mode;
                v.paramMode = paramMode;
            }
            if( !v.error.empty() )
            {
                SetLastError( "cannot lower node tree: " + v.error );
                return false;
            }
            if( wantSource ) program.source = v.source;
            program.sourceHash = v.sourceHash;
            program.params = v.params;
            program.mathMode = v.mathMode;
            return true;
        }

        std::string KernelSource( const Generator* root, int kind, int ppt )
        {
            Program program;
            return ProgramFor( root, kind, ppt, program, true ) ? program.source : std::string();
        }

        std::vector<float> KernelParams( const Generator* root, int kind, int ppt )
        {
            Program program;
            return ProgramFor( root, kind, ppt, program, false ) ? program.params : std::vector<float>();
        }

        int PrepareOffline( const Generator* root, int kind, int ppt )
        {
            Program program;
            std::string error;
            if( !ProgramFor( root, kind, ppt, program, true ) ) return -4;
            if( CubinFor( program.source, program.mathMode, error ).empty() )
            {
                SetLastError( error );
                return -3;
            }
            return 0;
        }

        // 0 or FNB200_E* (-2 CUDA, -3 compile, -4 unsupported tree); the current device is the one launched on
        static int ModuleFor( const Generator* root, int kind, int ppt, const Module*& module, Program& program )
        {
            module = nullptr;
            if( !ProgramFor( root, kind, ppt, program, false ) ) return -4;
            int device = 0;
            if( !Check( cudaGetDevice( &device ), "cudaGetDevice" ) ) return -2;
            {
                std::lock_guard<std::mutex> lock( gModuleMutex );
                auto it = gModules.find( { program.sourceHash, device } );
                if( it != gModules.end() ) module = &it->second;
            }
            if( module ) return 0;
            std::string error;
            if( !ProgramFor( root, kind, ppt, program, true ) ) return -4;
            module = LoadModule( program.source, program.sourceHash, program.mathMode, device, error );
            if( module ) return 0;
            SetLastError( error );
            return error.compare( 0, 5, "nvrtc" ) == 0 || error.compare( 0, 5, "NVRTC" ) == 0 ? -3 : -2;
        }

        // Points per thread: 4 (one 16-byte store per thread, neighbouring points share what does not depend on x) needs rows that
        // are a multiple of 4 wide and a call large enough to fill the GPU that way; everything else runs one point per thread.
        static int MaxPointsPerThread()
        {
            return EnvInt( "FASTNOISE2_B200_PPT", 4 ) >= 4 ? 4 : 1;
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
            return ppt > 1 && points / (unsigned)ppt < ResidentThreads() ? 1 : ppt;
        }

        // --------------------------------------------------------------- launches (device pointers)
        static bool Launch( const Module* m, unsigned long long groups, void** args, cudaStream_t stream )
        {
            const unsigned block = (unsigned)BlockThreads();
            unsigned long long blocks = ( groups + block - 1 ) / block;
            if( blocks == 0 ) return true;
            // Grid-stride launch: every thread walks about kItersPerThread groups, so the per-block work (table staging, the min/max
            // reduction) is paid once per that many groups, while the grid stays many waves deep -- the hardware block scheduler
            // keeps balancing the SMs to the end of the call. (A grid of exactly the resident blocks, every thread walking the whole
            // call, measured 11 % SLOWER on the heavy trees: warps that the issue arbiter favours finish early and the tail runs
            // under-occupied. profiles/r2_perf_sweep_b_launch_geometry.txt) Never fewer blocks than the device keeps resident.
            // Above kConvergedCodeBytes of loop body one group per thread: the warps of a block then run through the code together,
            // one instruction stream per block. Walking several groups they drift apart, every warp fetches its own stream, and the
            // instruction cache thrashes (terrain tree, 55 KB: ICC hit rate 70 %, "no instruction" the top stall, -11 %).
            const unsigned long long resident = (unsigned long long)std::max( 1, m->residentBlocks );
            const int iters = EnvInt( "FASTNOISE2_B200_ITERS", m->codeBytes > kConvergedCodeBytes ? 1 : kItersPerThread );
            blocks = iters <= 0 ? std::min( blocks, resident ) : std::min( blocks, std::max( resident, ( blocks + iters - 1 ) / iters ) );
            if( blocks > 0x7fffffffull )
            {
                SetLastError( "launch too large" );
                return false;
            }
            CountLaunch();
            return Check( cudaLaunchKernel( (const void*)m->kernel, dim3( (unsigned)blocks ), dim3( block ), args, 0, stream ), "cudaLaunchKernel" );
        }

        int LaunchGridRows( const Generator* root, const GridCall& call, float* deviceOut, unsigned long long rowBegin, unsigned long long rowCount,
                            unsigned* deviceMinMax, cudaStream_t stream, const Mirrors* mirrors )
        {
            const unsigned nx = (unsigned)call.count[0];
            const unsigned long long points = rowCount * nx;
            if( points == 0 ) return 0;
            // the kernel's index arithmetic (fn2::fast_div) is exact below 2^31, for points and for rows of the full grid alike
            if( points >= 0x80000000ull || rowBegin + rowCount >= 0x80000000ull )
            {
                SetLastError( "a single launch is limited to 2^31 - 1 points, and a grid to 2^31 - 1 rows" );
                return -1;
            }
            // four points per thread for large calls: groups inside one row when rows are a multiple of four wide, else (uniform
            // grids at least four wide) four consecutive points of the output with per-point rows
            int ppt = PointsPerThreadFor( points, MaxPointsPerThread() );
            const int ragged = EnvInt( "FASTNOISE2_B200_RAGGED", 1 ); // 0: never, 1: by code size (below), 2: always
            if( ppt == 4 && nx % 4u != 0 ) ppt = call.kind != kTileable2D && nx >= 4 && ragged ? kRaggedPointsPerThread : 1;

            const Module* m = nullptr;
            Program program;
            if( int rc = ModuleFor( root, call.kind, ppt, m, program ) ) return rc;
            if( ppt == kRaggedPointsPerThread && ragged == 1 && m->codeBytes > kConvergedCodeBytes )
            {
                // Heavy trees: four inlined copies of the tree with per-point rows lose to the one-point kernel, whose code fits
                // the instruction cache (terrain tree 1030 x 1030 x 250: 13.8 vs 13.4 ms; light trees gain 5-19 % from the
                // four-point form, profiles/r2_odd_width.txt)
                ppt = 1;
                if( int rc = ModuleFor( root, call.kind, ppt, m, program ) ) return rc;
            }
            if( program.params.empty() ) program.params.push_back( 0.0f );

            int seed = call.seed;
            const unsigned long long groups = ppt == kRaggedPointsPerThread ? ( points + 3 ) / 4 : points / (unsigned)ppt;
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
                a.groups = (unsigned)groups;
                a.divX = MakeFastDiv( nx );
                for( int k = 0; mirrors && k < mirrors->count && k < 7; k++ )
                {
                    a.mirror[a.mirrors++] = mirrors->out[k];
                    a.vecStore = a.vecStore && ( (std::uintptr_t)mirrors->out[k] % 16 ) == 0;
                }
                void* args[] = { &deviceOut, &a, &deviceMinMax, &seed, program.params.data() };
                return Launch( m, groups, args, stream ) ? 0 : -2;
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
            a.groups = (unsigned)groups;
            a.divX = MakeFastDiv( nx );
            a.divY = MakeFastDiv( (unsigned)call.count[1] );
            a.divZ = MakeFastDiv( (unsigned)call.count[2] );
            for( int k = 0; mirrors && k < mirrors->count && k < 7; k++ )
            {
                a.mirror[a.mirrors++] = mirrors->out[k];
                a.vecStore = a.vecStore && ( (std::uintptr_t)mirrors->out[k] % 16 ) == 0;
            }
            void* args[] = { &deviceOut, &a, &deviceMinMax, &seed, program.params.data() };
            return Launch( m, groups, args, stream ) ? 0 : -2;
        }

        int LaunchPositions( const Generator* root, int dims, float* deviceOut, unsigned long long count, const float* const pos[4],
                             const float offset[4], int seed, unsigned* deviceMinMax, cudaStream_t stream )
        {
            if( count == 0 ) return 0;
            const int ppt = PointsPerThreadFor( count, MaxPointsPerThread() );
            const Module* m = nullptr;
            Program program;
            if( int rc = ModuleFor( root, 10 + dims, ppt, m, program ) ) return rc;
            if( program.params.empty() ) program.params.push_back( 0.0f );

            PosArgsHost a {};
            bool aligned = ( (std::uintptr_t)deviceOut % 16 ) == 0;
            for( int d = 0; d < 4; d++ )
            {
                a.pos[d] = d < dims ? pos[d] : nullptr;
                a.off[d] = offset[d];
                if( d < dims ) aligned = aligned && ( (std::uintptr_t)pos[d] % 16 ) == 0;
            }
            a.count = count;
            a.groups = ( count + ppt - 1 ) / ppt;
            a.vecIO = aligned;
            void* args[] = { &deviceOut, &a, &deviceMinMax, &seed, program.params.data() };
            return Launch( m, a.groups, args, stream ) ? 0 : -2;
        }

        int PrepareAll( const Generator* root, int kind )
        {
            // every variant the launchers above can choose
            const int only = EnvInt( "FASTNOISE2_B200_PREPARE_ONLY", 0 ); // tools: compile one points-per-thread form (4, 1 or 5)
            for( int ppt : { 4, 1, kRaggedPointsPerThread } )
            {
                if( only && ppt != only ) continue;
                if( ppt > 1 && MaxPointsPerThread() < 4 ) continue;
                if( ppt == kRaggedPointsPerThread && ( kind < kGrid2D || kind > kGrid4D ) ) continue;
                if( int rc = PrepareOffline( root, kind, ppt ) ) return rc;
            }
            return 0;
        }

        // --------------------------------------------------------------- reference-shaped calls (host or device pointers)
        static std::size_t ChunkFloats()
        {
            // staging buffer size per pipeline slot; two slots are allocated per calling thread and device
            return (std::size_t)EnvInt( "FASTNOISE2_B200_STAGING_MB", 64 ) * 1024 * 1024 / sizeof( float );
        }

        // A reference-shaped call has no stream argument: work the caller queued earlier on the legacy default stream (and, through
        // it, on every blocking stream) is ordered before this call's kernels; producers on NON-blocking streams must be
        // synchronised by the caller.
        static bool OrderAfterCaller( DeviceContext& ctx )
        {
            return Check( cudaEventRecord( ctx.callerDone, cudaStreamLegacy ), "cudaEventRecord(legacy stream)" ) &&
                   Check( cudaStreamWaitEvent( ctx.compute, ctx.callerDone, 0 ), "cudaStreamWaitEvent" );
        }

        static bool ReadMinMax( DeviceContext& ctx, unsigned keys[2] )
        {
            if( !Check( cudaMemcpyAsync( ctx.minMaxHost, ctx.minMax, 2 * sizeof( unsigned ), cudaMemcpyDeviceToHost, ctx.compute ), "cudaMemcpyAsync(minmax)" ) ) return false;
            if( !Check( cudaStreamSynchronize( ctx.compute ), "cudaStreamSynchronize" ) ) return false;
            keys[0] = std::max( keys[0], ctx.minMaxHost[0] );
            keys[1] = std::max( keys[1], ctx.minMaxHost[1] );
            return true;
        }

        // Devices a host-destination call of `points` points is spread over: the current one, plus every other visible sm_100 GPU
        // when the call is large enough for the extra PCIe links to matter (FASTNOISE2_B200_FANOUT=0 disables, =N caps).
        static std::vector<int> FanoutDevices( int current, unsigned long long points )
        {
            std::vector<int> devices { current };
            const int cap = EnvInt( "FASTNOISE2_B200_FANOUT", 1024 );
            const unsigned long long minPoints = (unsigned long long)EnvInt( "FASTNOISE2_B200_FANOUT_MIN_MPOINTS", 32 ) << 20;
            int count = 0;
            if( cap <= 1 || points < minPoints || cudaGetDeviceCount( &count ) != cudaSuccess ) return devices;
            for( int d = 0; d < count && (int)devices.size() < cap; d++ )
            {
                int major = 0;
                if( d != current && cudaDeviceGetAttribute( &major, cudaDevAttrComputeCapabilityMajor, d ) == cudaSuccess && major == 10 ) devices.push_back( d );
            }
            cudaGetLastError();
            return devices;
        }

        // smaller pageable transfers go straight through cudaMemcpyAsync: waking the copy workers costs more than the driver's staging
        constexpr long long kBounceMinFloats = 1 << 20;

        struct RestoreDevice
        {
            int device;
            ~RestoreDevice() { cudaSetDevice( device ); }
        };

        // One device's share of a staged call: units [begin, end) (rows or positions), issued chunk by chunk.
        struct Shard
        {
            DeviceContext* ctx = nullptr;
            long long begin = 0, end = 0, next = 0;
            int slot = 0;
            bool used[2] = { false, false };
        };

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
                // same limit as the reference, whose index arithmetic is 32-bit [Generator.inl:150]
                if( rows * call.count[0] >= 0x80000000ll )
                {
                    SetLastError( "grid has 2^31 or more points; use the slab entry points of FastNoise_B200.h" );
                    return failed;
                }
            }
            const long long nx = call.count[0];
            if( !out )
            {
                SetLastError( "output pointer is null" );
                return failed;
            }

            int current = 0;
            if( !Check( cudaGetDevice( &current ), "cudaGetDevice" ) ) return failed;
            RestoreDevice restore { current };
            unsigned keys[2] = { 0, 0 };

            const Memory where = Classify( out );
            if( where == Memory::Device )
            {
                DeviceContext* ctx = tContexts.Get( current );
                if( !ctx ) return failed;
                if( !OrderAfterCaller( *ctx ) ) return failed;
                if( !Check( cudaMemsetAsync( ctx->minMax, 0, 2 * sizeof( unsigned ), ctx->compute ), "cudaMemsetAsync" ) ) return failed;
                if( LaunchGridRows( root, call, out, 0, (unsigned long long)rows, ctx->minMax, ctx->compute ) != 0 ) return failed;
                return ReadMinMax( *ctx, keys ) ? DecodeMinMax( keys ) : failed;
            }

            // host destination: every device generates a contiguous share of the rows chunk by chunk into two device buffers; the D2H
            // copy of chunk i runs on its own stream while chunk i+1 is being computed. Pageable destinations are reached through
            // pinned bounce buffers and the copy workers.
            const bool pageable = where == Memory::Pageable && rows * nx >= kBounceMinFloats && EnvInt( "FASTNOISE2_B200_BOUNCE", 1 ) != 0;
            const std::vector<int> devices = FanoutDevices( current, (unsigned long long)( rows * nx ) );
            const long long chunkRows = std::max<long long>( 1, std::min<long long>( ( rows + (long long)devices.size() - 1 ) / (long long)devices.size(), (long long)( ChunkFloats() / (std::size_t)nx ) ) );
            std::vector<Shard> shards( devices.size() );
            for( std::size_t i = 0; i < devices.size(); i++ )
            {
                Shard& s = shards[i];
                s.begin = s.next = rows * (long long)i / (long long)devices.size();
                s.end = rows * (long long)( i + 1 ) / (long long)devices.size();
                s.ctx = tContexts.Get( devices[i] );
                if( !s.ctx || !s.ctx->EnsureOut( (std::size_t)( chunkRows * nx ) ) ) return failed;
                if( pageable && !s.ctx->EnsureBounce( (std::size_t)( chunkRows * nx ) ) ) return failed;
                if( !Check( cudaMemsetAsync( s.ctx->minMax, 0, 2 * sizeof( unsigned ), s.ctx->compute ), "cudaMemsetAsync" ) ) return failed;
            }

            bool ok = true;
            for( bool work = true; work && ok; )
            {
                work = false;
                for( Shard& s : shards )
                {
                    if( s.next >= s.end ) continue;
                    work = true;
                    DeviceContext& ctx = *s.ctx;
                    const long long row = s.next, n = std::min( chunkRows, s.end - row );
                    const int slot = s.slot;
                    const std::size_t bytes = (std::size_t)( n * nx ) * sizeof( float );
                    ok = Check( cudaSetDevice( ctx.device ), "cudaSetDevice" );
                    if( ok && s.used[slot] ) ok = Check( cudaStreamWaitEvent( ctx.compute, ctx.drained[slot], 0 ), "cudaStreamWaitEvent" );
                    ok = ok && LaunchGridRows( root, call, ctx.outBuf[slot], (unsigned long long)row, (unsigned long long)n, ctx.minMax, ctx.compute ) == 0;
                    if( ok && pageable ) ctx.bounceBusy[slot].Wait(); // the workers have emptied this slot's bounce buffer
                    float* target = pageable ? ctx.bounce[slot] : out + row * nx;
                    ok = ok && Check( cudaEventRecord( ctx.computed[slot], ctx.compute ), "cudaEventRecord" ) &&
                         Check( cudaStreamWaitEvent( ctx.copyOut, ctx.computed[slot], 0 ), "cudaStreamWaitEvent" ) &&
                         Check( cudaMemcpyAsync( target, ctx.outBuf[slot], bytes, cudaMemcpyDeviceToHost, ctx.copyOut ), "cudaMemcpyAsync(D2H)" ) &&
                         Check( cudaEventRecord( ctx.drained[slot], ctx.copyOut ), "cudaEventRecord" );
                    if( ok && pageable ) CopyPool::Get().Submit( ctx.drained[slot], (const char*)ctx.bounce[slot], (char*)( out + row * nx ), bytes, ctx.bounceBusy[slot] );
                    s.used[slot] = true;
                    s.slot ^= 1;
                    s.next += n;
                    if( !ok ) break;
                }
            }
            // drain every device even after a failure: the workers and DMA engines still reference the caller's buffer
            for( Shard& s : shards )
            {
                if( !s.ctx ) continue;
                cudaSetDevice( s.ctx->device );
                if( ok ) ok = ReadMinMax( *s.ctx, keys );
                if( !Check( cudaStreamSynchronize( s.ctx->copyOut ), "cudaStreamSynchronize" ) ) ok = false;
                for( CopyPool::Ticket& t : s.ctx->bounceBusy ) t.Wait();
            }
            return ok ? DecodeMinMax( keys ) : failed;
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

            int current = 0;
            if( !Check( cudaGetDevice( &current ), "cudaGetDevice" ) ) return failed;
            RestoreDevice restore { current };
            unsigned keys[2] = { 0, 0 };

            const Memory outWhere = Classify( out );
            const bool outOnDevice = outWhere == Memory::Device;
            bool inOnDevice = true;
            for( int d = 0; d < dims; d++ ) inOnDevice = inOnDevice && Classify( pos[d] ) == Memory::Device;

            if( outOnDevice && inOnDevice )
            {
                DeviceContext* ctx = tContexts.Get( current );
                if( !ctx ) return failed;
                if( !OrderAfterCaller( *ctx ) ) return failed;
                if( !Check( cudaMemsetAsync( ctx->minMax, 0, 2 * sizeof( unsigned ), ctx->compute ), "cudaMemsetAsync" ) ) return failed;
                if( LaunchPositions( root, dims, out, (unsigned long long)count, pos, offset, seed, ctx->minMax, ctx->compute ) != 0 ) return failed;
                return ReadMinMax( *ctx, keys ) ? DecodeMinMax( keys ) : failed;
            }

            // staged pipeline per device: H2D of chunk i+1 (copyIn), kernel of chunk i (compute), D2H of chunk i-1 (copyOut).
            // Device-resident arrays on either side pin the call to the current device.
            const bool pageable = outWhere == Memory::Pageable && count >= kBounceMinFloats && EnvInt( "FASTNOISE2_B200_BOUNCE", 1 ) != 0;
            const std::vector<int> devices = ( outOnDevice || inOnDevice ) ? std::vector<int> { current } : FanoutDevices( current, (unsigned long long)count );
            const long long perDevice = ( count + (long long)devices.size() - 1 ) / (long long)devices.size();
            const long long chunk = std::max<long long>( 4, std::min<long long>( ( perDevice + 3 ) & ~3ll, (long long)( ChunkFloats() / 2 ) & ~3ll ) );
            std::vector<Shard> shards( devices.size() );
            for( std::size_t i = 0; i < devices.size(); i++ )
            {
                Shard& s = shards[i];
                s.begin = s.next = std::min( count, ( ( count * (long long)i / (long long)devices.size() ) + 3 ) & ~3ll ); // 16-byte aligned shares
                s.end = i + 1 == devices.size() ? count : std::min( count, ( ( count * (long long)( i + 1 ) / (long long)devices.size() ) + 3 ) & ~3ll );
                s.ctx = tContexts.Get( devices[i] );
                if( !s.ctx ) return failed;
                if( !outOnDevice && !s.ctx->EnsureOut( (std::size_t)chunk ) ) return failed;
                if( !inOnDevice && !s.ctx->EnsureIn( (std::size_t)chunk, dims ) ) return failed;
                if( pageable && !s.ctx->EnsureBounce( (std::size_t)chunk ) ) return failed;
                if( ( outOnDevice || inOnDevice ) && !OrderAfterCaller( *s.ctx ) ) return failed;
                if( !Check( cudaMemsetAsync( s.ctx->minMax, 0, 2 * sizeof( unsigned ), s.ctx->compute ), "cudaMemsetAsync" ) ) return failed;
            }

            bool ok = true;
            for( bool work = true; work && ok; )
            {
                work = false;
                for( Shard& s : shards )
                {
                    if( s.next >= s.end ) continue;
                    work = true;
                    DeviceContext& ctx = *s.ctx;
                    const long long first = s.next, n = std::min( chunk, s.end - first );
                    const int slot = s.slot;
                    ok = Check( cudaSetDevice( ctx.device ), "cudaSetDevice" );
                    const float* devPos[4] = { nullptr, nullptr, nullptr, nullptr };
                    if( inOnDevice )
                    {
                        for( int d = 0; d < dims; d++ ) devPos[d] = pos[d] + first;
                        if( ok && !outOnDevice && s.used[slot] ) ok = Check( cudaStreamWaitEvent( ctx.compute, ctx.drained[slot], 0 ), "cudaStreamWaitEvent" );
                    }
                    else
                    {
                        // the slot's input buffers are free once the kernel that last read them has finished
                        if( ok && s.used[slot] ) ok = Check( cudaStreamWaitEvent( ctx.copyIn, ctx.computed[slot], 0 ), "cudaStreamWaitEvent" );
                        for( int d = 0; ok && d < dims; d++ )
                        {
                            ok = Check( cudaMemcpyAsync( ctx.inBuf[slot][d], pos[d] + first, (std::size_t)n * sizeof( float ), cudaMemcpyHostToDevice, ctx.copyIn ), "cudaMemcpyAsync(H2D)" );
                            devPos[d] = ctx.inBuf[slot][d];
                        }
                        ok = ok && Check( cudaEventRecord( ctx.uploaded[slot], ctx.copyIn ), "cudaEventRecord" ) &&
                             Check( cudaStreamWaitEvent( ctx.compute, ctx.uploaded[slot], 0 ), "cudaStreamWaitEvent" );
                        if( ok && !outOnDevice && s.used[slot] ) ok = Check( cudaStreamWaitEvent( ctx.compute, ctx.drained[slot], 0 ), "cudaStreamWaitEvent" );
                    }

                    float* devOut = outOnDevice ? out + first : ctx.outBuf[slot];
                    ok = ok && LaunchPositions( root, dims, devOut, (unsigned long long)n, devPos, offset, seed, ctx.minMax, ctx.compute ) == 0;
                    ok = ok && Check( cudaEventRecord( ctx.computed[slot], ctx.compute ), "cudaEventRecord" );
                    if( ok && !outOnDevice )
                    {
                        if( pageable ) ctx.bounceBusy[slot].Wait();
                        float* target = pageable ? ctx.bounce[slot] : out + first;
                        ok = Check( cudaStreamWaitEvent( ctx.copyOut, ctx.computed[slot], 0 ), "cudaStreamWaitEvent" ) &&
                             Check( cudaMemcpyAsync( target, devOut, (std::size_t)n * sizeof( float ), cudaMemcpyDeviceToHost, ctx.copyOut ), "cudaMemcpyAsync(D2H)" ) &&
                             Check( cudaEventRecord( ctx.drained[slot], ctx.copyOut ), "cudaEventRecord" );
                        if( ok && pageable ) CopyPool::Get().Submit( ctx.drained[slot], (const char*)ctx.bounce[slot], (char*)( out + first ), (std::size_t)n * sizeof( float ), ctx.bounceBusy[slot] );
                    }
                    s.used[slot] = true;
                    s.slot ^= 1;
                    s.next += n;
                    if( !ok ) break;
                }
            }
            for( Shard& s : shards )
            {
                if( !s.ctx ) continue;
                cudaSetDevice( s.ctx->device );
                if( ok ) ok = ReadMinMax( *s.ctx, keys );
                if( !Check( cudaStreamSynchronize( s.ctx->copyOut ), "cudaStreamSynchronize" ) ) ok = false;
                if( !Check( cudaStreamSynchronize( s.ctx->copyIn ), "cudaStreamSynchronize" ) ) ok = false;
                for( CopyPool::Ticket& t : s.ctx->bounceBusy ) t.Wait();
            }
            return ok ? DecodeMinMax( keys ) : failed;
        }

        cudaStream_t ThreadStream()
        {
            int current = 0;
            if( cudaGetDevice( &current ) != cudaSuccess ) return nullptr;
            DeviceContext* ctx = tContexts.Get( current );
            return ctx ? ctx->compute : nullptr;
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
