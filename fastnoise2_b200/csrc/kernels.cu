// fastnoise2_b200 -- ahead-of-time compiled helper kernels (sm_100a): synthetic position generator for the
// position-array benchmark and the ALU issue-rate probe that provides the roofline denominator
// (MEASURED_PEAKS.json has HBM and tensor peaks only; this workload is bound by FP32/INT32 issue).
#include <cuda_runtime.h>
#include <cstdint>

#include "backend.h"

namespace
{
    // Philox4x32-10 (Salmon et al., SC'11), counter = (index lo, index hi, axis, 0), key = (key, 0x9E3779B9)
    __device__ __forceinline__ void PhiloxRound( uint32_t ( &c )[4], uint32_t ( &k )[2] )
    {
        const uint32_t hi0 = __umulhi( 0xD2511F53u, c[0] ), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi( 0xCD9E8D57u, c[2] ), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }

    __global__ void __launch_bounds__( 256 ) PhiloxFillKernel( float* __restrict__ out, long long count, long long firstIndex, uint32_t key, uint32_t axis, float lo, float span )
    {
        const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
        if( i >= count ) return;
        const unsigned long long index = (unsigned long long)( firstIndex + i );
        uint32_t c[4] = { (uint32_t)index, (uint32_t)( index >> 32 ), axis, 0u };
        uint32_t k[2] = { key, 0x9E3779B9u };
#pragma unroll
        for( int r = 0; r < 10; r++ ) PhiloxRound( c, k );
        // 24 random mantissa bits -> [0,1), then affine map; fmaf keeps it a single rounding
        const float u = (float)( c[0] >> 8 ) * ( 1.0f / 16777216.0f );
        out[i] = fmaf( u, span, lo );
    }

    // ---- ALU issue-rate probe: 8 independent dependency chains per thread so the pipes stay full
    template<int MODE>
    __global__ void __launch_bounds__( 256 ) AluProbeKernel( float* __restrict__ sink, int iters, float a, float b, uint32_t m )
    {
        float f[8];
        uint32_t u[8];
#pragma unroll
        for( int k = 0; k < 8; k++ )
        {
            f[k] = a + (float)( threadIdx.x + k );
            u[k] = m + threadIdx.x * 8u + k;
        }
        for( int it = 0; it < iters; it++ )
        {
#pragma unroll
            for( int rep = 0; rep < 16; rep++ )
            {
#pragma unroll
                for( int k = 0; k < 8; k++ )
                {
                    if( MODE == 0 ) f[k] = fmaf( f[k], a, b );                       // FFMA
                    if( MODE == 1 ) u[k] = u[k] * m + 0x9E3779B9u;                   // IMAD
                    if( MODE == 2 ) u[k] = ( u[k] ^ m ) & ( u[( k + 1 ) & 7] | 0x55555555u ); // LOP3
                    if( MODE == 3 )
                    {
                        f[k] = fmaf( f[k], a, b );
                        u[k] = ( u[k] ^ m ) & ( u[( k + 1 ) & 7] | 0x55555555u );
                    }
                    if( MODE >= 4 && MODE <= 8 && ( k & 1 ) == 0 )                    // packed FP32: one FFMA2 per pair of chains
                    {
                        const float2 r = MODE == 8 ? __ffma2_rn( make_float2( f[k], f[k + 1] ), make_float2( 1.0001f, 1.0001f ), make_float2( 0.5f, 0.5f ) )
                                                   : __ffma2_rn( make_float2( f[k], f[k + 1] ), make_float2( a, a ), make_float2( b, b ) );
                        f[k] = r.x;
                        f[k + 1] = r.y;
                    }
                    if( MODE == 5 && ( k & 1 ) == 0 ) u[k] = ( u[k] ^ m ) & ( u[( k + 2 ) & 7] | 0x55555555u );   // + 1 LOP3 per FFMA2
                    if( MODE == 6 ) u[k] = ( u[k] ^ m ) & ( u[( k + 1 ) & 7] | 0x55555555u );                     // + 2 LOP3 per FFMA2
                    // The question behind the axis-packed leaves: does an FFMA2 take ONE issue slot? 1 FFMA2 + 1 LOP3 per pair of chains
                    // needs 2 cycles of the FMA pipe, 2 of the ALU pipe and 2 issue slots if it does (3 if the FFMA2 takes two);
                    // mode 9 is the scalar counterpart (2 FFMA + 1 LOP3: 3 issue slots), mode 8 the FFMA2 with a broadcast operand.
                    if( MODE == 7 && ( k & 1 ) == 0 ) u[k] = ( u[k] ^ m ) & u[( k + 2 ) & 7];                     // 1 FFMA2 + 1 LOP3 (one instruction)
                    if( MODE == 8 && ( k & 1 ) == 0 ) u[k] = ( u[k] ^ m ) & u[( k + 2 ) & 7];
                    if( MODE == 10 && ( k & 1 ) == 0 )                                                            // FFMA2 with three register-PAIR operands + 1 LOP3
                    {
                        const float2 r = __ffma2_rn( make_float2( f[k], f[k + 1] ), make_float2( f[( k + 2 ) & 7], f[( k + 3 ) & 7] ), make_float2( f[( k + 4 ) & 7], f[( k + 5 ) & 7] ) );
                        f[k] = r.x;
                        f[k + 1] = r.y;
                        u[k] = ( u[k] ^ m ) & u[( k + 2 ) & 7];
                    }
                    if( MODE == 9 )
                    {
                        f[k] = fmaf( f[k], a, b );
                        if( ( k & 1 ) == 0 ) u[k] = ( u[k] ^ m ) & u[( k + 2 ) & 7];
                    }
                }
            }
        }
        float acc = 0;
#pragma unroll
        for( int k = 0; k < 8; k++ ) acc += f[k] + (float)u[k];
        if( acc == 12345.678f ) sink[0] = acc; // keeps the chains alive without real traffic
    }
}

namespace FastNoise
{
    namespace Impl
    {
        int FillUniformPhilox( float* deviceOut, long long count, long long firstIndex, unsigned key, unsigned axis, float lo, float hi, cudaStream_t stream )
        {
            if( count <= 0 ) return 0;
            const long long blocks = ( count + 255 ) / 256;
            CountLaunch();
            PhiloxFillKernel<<<(unsigned)blocks, 256, 0, stream>>>( deviceOut, count, firstIndex, key, axis, lo, hi - lo );
            return cudaGetLastError() == cudaSuccess ? 0 : -2;
        }

        double MeasureAluPeak( int mode, int iters )
        {
            int device = 0, sms = 0;
            if( cudaGetDevice( &device ) != cudaSuccess || cudaDeviceGetAttribute( &sms, cudaDevAttrMultiProcessorCount, device ) != cudaSuccess ) return 0;
            float* sink = nullptr;
            if( cudaMalloc( &sink, 64 ) != cudaSuccess ) return 0;
            cudaEvent_t t0, t1;
            cudaEventCreate( &t0 );
            cudaEventCreate( &t1 );
            const int blocks = sms * 8; // 8 x 256 threads = full occupancy at <= 32 registers
            auto launch = [&]( int n )
            {
                CountLaunch();
                switch( mode )
                {
                case 0: AluProbeKernel<0><<<blocks, 256>>>( sink, n, 1.0001f, 0.5f, 0x2545F491u ); break;
                case 1: AluProbeKernel<1><<<blocks, 256>>>( sink, n, 1.0001f, 0.5f, 0x2545F491u ); break;
                case 2: AluProbeKernel<2><<<blocks, 256>>>( sink, n, 1.0001f, 0.5f, 0x2545F491u ); break;
                case 3: AluProbeKernel<3><<<blocks, 256>>>( sink, n, 1.0001f, 0.5f, 0x2545F491u ); break;
                case 4: AluProbeKernel<4><<<blocks, 256>>>( sink, n, 1.0001f, 0.5f, 0x2545F491u ); break;
                case 5: AluProbeKernel<5><<<blocks, 256>>>( sink, n, 1.0001f, 0.5f, 0x2545F491u ); break;
                case 6: AluProbeKernel<6><<<blocks, 256>>>( sink, n, 1.0001f, 0.5f, 0x2545F491u ); break;
                case 7: AluProbeKernel<7><<<blocks, 256>>>( sink, n, 1.0001f, 0.5f, 0x2545F491u ); break;
                case 8: AluProbeKernel<8><<<blocks, 256>>>( sink, n, 1.0001f, 0.5f, 0x2545F491u ); break;
                case 9: AluProbeKernel<9><<<blocks, 256>>>( sink, n, 1.0001f, 0.5f, 0x2545F491u ); break;
                default: AluProbeKernel<10><<<blocks, 256>>>( sink, n, 1.0001f, 0.5f, 0x2545F491u ); break;
                }
            };
            launch( 16 ); // warm-up
            cudaDeviceSynchronize();
            double best = 0;
            for( int trial = 0; trial < 3; trial++ )
            {
                cudaEventRecord( t0 );
                launch( iters );
                cudaEventRecord( t1 );
                cudaEventSynchronize( t1 );
                float ms = 0;
                cudaEventElapsedTime( &ms, t0, t1 );
                // lane-operations per repetition (an FFMA2 counts as 2); modes 7-9: 8 FP lane-ops + 4 LOP3
                const double perRep = mode == 3 ? 16 : ( mode == 5 ? 12 : ( mode == 6 ? 16 : ( mode >= 7 ? 12 : 8 ) ) );
                const double opsPerThread = (double)iters * 16 * perRep;
                const double rate = opsPerThread * blocks * 256 / ( ms * 1e-3 );
                if( ms > 0 && rate > best ) best = rate;
            }
            cudaEventDestroy( t0 );
            cudaEventDestroy( t1 );
            cudaFree( sink );
            return cudaGetLastError() == cudaSuccess ? best : 0;
        }
    }
}
