// ORACLE / TEST INFRASTRUCTURE ONLY.
// Small C entry points added next to the reference's own C ABI (FastNoise_C.cpp) inside each
// oracle/_ref/libfn2ref_*.so: build identification and a threaded driver that follows the
// reference's documented threading model (one shared const tree, one slab per thread,
// Generator.h:166-167) for the CPU baseline timing; and the reference's own serialiser behind a C entry
// point, so that the product's codec can be compared with it byte for byte.
#include <FastNoise/FastNoise_C.h>
#include <FastNoise/Metadata.h>
#include <FastSIMD/DispatchClass.h>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

extern "C"
{
    int fn2refLanes() { return FN2SHIM_LANES; }
    int fn2refRelaxed() { return FN2SHIM_RELAXED; }
    // Metadata::DeserialiseNodeData + Metadata::SerialiseNodeData of the reference (Metadata.cpp:161-308, 508-632):
    // the canonical spelling of a tree. Returns the length (0 = rejected); same contract as fnb200CanonicalEncodedNodeTree.
    long fn2refCanonicalEncodedNodeTree( const char* encoded, char* buf, long bufSize )
    {
        std::vector<std::unique_ptr<FastNoise::NodeData>> records;
        FastNoise::NodeData* root = FastNoise::Metadata::DeserialiseNodeData( encoded, records );
        std::string text = root ? FastNoise::Metadata::SerialiseNodeData( root, false ) : std::string();
        if( buf && bufSize > 0 )
        {
            long n = std::min<long>( bufSize - 1, (long)text.size() );
            std::memcpy( buf, text.data(), (size_t)n );
            buf[n] = 0;
        }
        return (long)text.size();
    }

    int fn2refTables() { return FN2SHIM_AVX512_PATHS; } // 1: the reference's AVX512_F gradient-table branches are live

    // Generates nx*ny*nz points (row-major, x fastest) into out, splitting z (or y when nz==1)
    // into `threads` contiguous slabs, each through the ordinary fnGenUniformGrid3D/2D call with a
    // float slab offset (the reference API has no integer slab origin). Returns seconds of wall time.
    // dim = 2 uses fnGenUniformGrid2D (nz ignored).
    double fn2refTimedGridMT( const void* node, float* out, int dim, int nx, int ny, int nz,
                              float x0, float y0, float z0, float step, int seed, int threads, float* minmax )
    {
        if( dim == 2 ) nz = 1;
        int slabs = dim == 2 ? ny : nz;
        threads = std::max( 1, std::min( threads, slabs ) );
        std::vector<float> mm( (size_t)threads * 2 );
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> pool;
        for( int t = 0; t < threads; t++ )
        {
            int begin = (int)( (int64_t)slabs * t / threads );
            int end = (int)( (int64_t)slabs * ( t + 1 ) / threads );
            pool.emplace_back( [=, &mm]()
            {
                if( dim == 2 )
                {
                    fnGenUniformGrid2D( node, out + (size_t)begin * nx, x0, y0 + begin * step, nx, end - begin, step, step, seed, &mm[t * 2] );
                }
                else
                {
                    fnGenUniformGrid3D( node, out + (size_t)begin * nx * ny, x0, y0, z0 + begin * step, nx, ny, end - begin, step, step, step, seed, &mm[t * 2] );
                }
            } );
        }
        for( auto& th : pool ) th.join();
        auto t1 = std::chrono::steady_clock::now();
        if( minmax )
        {
            minmax[0] = mm[0]; minmax[1] = mm[1];
            for( int t = 1; t < threads; t++ ) { minmax[0] = std::min( minmax[0], mm[t * 2] ); minmax[1] = std::max( minmax[1], mm[t * 2 + 1] ); }
        }
        return std::chrono::duration<double>( t1 - t0 ).count();
    }

    // Same idea for GenPositionArray4D / 3D / 2D: contiguous index ranges per thread.
    double fn2refTimedPositionsMT( const void* node, float* out, int dim, int count,
                                   const float* xs, const float* ys, const float* zs, const float* ws,
                                   int seed, int threads, float* minmax )
    {
        threads = std::max( 1, std::min( threads, count ) );
        std::vector<float> mm( (size_t)threads * 2 );
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> pool;
        for( int t = 0; t < threads; t++ )
        {
            int begin = (int)( (int64_t)count * t / threads );
            int end = (int)( (int64_t)count * ( t + 1 ) / threads );
            pool.emplace_back( [=, &mm]()
            {
                int n = end - begin;
                if( dim == 2 ) fnGenPositionArray2D( node, out + begin, n, xs + begin, ys + begin, 0, 0, seed, &mm[t * 2] );
                else if( dim == 3 ) fnGenPositionArray3D( node, out + begin, n, xs + begin, ys + begin, zs + begin, 0, 0, 0, seed, &mm[t * 2] );
                else fnGenPositionArray4D( node, out + begin, n, xs + begin, ys + begin, zs + begin, ws + begin, 0, 0, 0, 0, seed, &mm[t * 2] );
            } );
        }
        for( auto& th : pool ) th.join();
        auto t1 = std::chrono::steady_clock::now();
        if( minmax )
        {
            minmax[0] = mm[0]; minmax[1] = mm[1];
            for( int t = 1; t < threads; t++ ) { minmax[0] = std::min( minmax[0], mm[t * 2] ); minmax[1] = std::max( minmax[1], mm[t * 2 + 1] ); }
        }
        return std::chrono::duration<double>( t1 - t0 ).count();
    }

    // GenTileable2D has no slab form (Generator.inl:370-433 walks the whole image), so a threaded baseline of the tileable config
    // splits the ROWS across threads and evaluates each slab through the reference's own GenPositionArray4D over the torus
    // positions of Generator.inl:386-406: x -> ( cos, sin )( xIdx / ( xSize / 2pi ) ) * xSize / 2pi * xStep, likewise y. The
    // positions are tabulated once per column / row with libm (outside the reference's FS::Cos / FS::Sin, which only the
    // single-threaded fnGenTileable2D exercises); this is a TIMING driver -- parity tests use fnGenTileable2D itself.
    double fn2refTimedTileableMT( const void* node, float* out, int xSize, int ySize, float xStep, float yStep, int seed, int threads, float* minmax )
    {
        threads = std::max( 1, std::min( threads, ySize ) );
        const float pi2Recip = 0.15915493667f;
        const float xSizePi = (float)xSize * pi2Recip, ySizePi = (float)ySize * pi2Recip;
        const float xFreq = xSizePi * xStep, yFreq = ySizePi * yStep, xMul = 1 / xSizePi, yMul = 1 / ySizePi;
        std::vector<float> colCos( xSize ), colSin( xSize ), rowCos( ySize ), rowSin( ySize );
        for( int x = 0; x < xSize; x++ ) { colCos[x] = std::cos( (float)x * xMul ) * xFreq; colSin[x] = std::sin( (float)x * xMul ) * xFreq; }
        for( int y = 0; y < ySize; y++ ) { rowCos[y] = std::cos( (float)y * yMul ) * yFreq; rowSin[y] = std::sin( (float)y * yMul ) * yFreq; }
        std::vector<float> mm( (size_t)threads * 2 );
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> pool;
        for( int t = 0; t < threads; t++ )
        {
            int begin = (int)( (int64_t)ySize * t / threads );
            int end = (int)( (int64_t)ySize * ( t + 1 ) / threads );
            pool.emplace_back( [=, &mm, &colCos, &colSin, &rowCos, &rowSin]()
            {
                // a few rows at a time keeps the position arrays in cache, like the reference's own per-vector positions
                const int rowsPerCall = std::max( 1, 16384 / xSize );
                std::vector<float> xs( (size_t)rowsPerCall * xSize ), ys( xs.size() ), zs( xs.size() ), ws( xs.size() );
                float lo = INFINITY, hi = -INFINITY;
                for( int row = begin; row < end; row += rowsPerCall )
                {
                    const int rows = std::min( rowsPerCall, end - row );
                    for( int r = 0; r < rows; r++ )
                    {
                        float* px = &xs[(size_t)r * xSize]; float* py = &ys[(size_t)r * xSize]; float* pz = &zs[(size_t)r * xSize]; float* pw = &ws[(size_t)r * xSize];
                        for( int x = 0; x < xSize; x++ ) { px[x] = colCos[x]; py[x] = rowCos[row + r]; pz[x] = colSin[x]; pw[x] = rowSin[row + r]; }
                    }
                    float part[2];
                    fnGenPositionArray4D( node, out + (size_t)row * xSize, rows * xSize, xs.data(), ys.data(), zs.data(), ws.data(), 0, 0, 0, 0, seed, part );
                    lo = std::min( lo, part[0] );
                    hi = std::max( hi, part[1] );
                }
                mm[t * 2] = lo;
                mm[t * 2 + 1] = hi;
            } );
        }
        for( auto& th : pool ) th.join();
        auto t1 = std::chrono::steady_clock::now();
        if( minmax )
        {
            minmax[0] = mm[0]; minmax[1] = mm[1];
            for( int t = 1; t < threads; t++ ) { minmax[0] = std::min( minmax[0], mm[t * 2] ); minmax[1] = std::max( minmax[1], mm[t * 2 + 1] ); }
        }
        return std::chrono::duration<double>( t1 - t0 ).count();
    }
}
