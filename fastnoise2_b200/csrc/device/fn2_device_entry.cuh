// fastnoise2_b200 -- device node library: batch-driver plumbing shared by every generated entry kernel.
// Replaces the reference's per-vector driver loops (Generators/Generator.inl:79-510: index vectors with
// AxisReset wrap, tail Load/StoreRemaining, lane-wise min/max) by exact integer index decomposition per
// thread, bounds-checked vector stores and a shuffle + shared-memory + atomic min/max reduction.
#pragma once

namespace fn2
{
    // n / d for n < 2^31 through a host-computed reciprocal: q = (umulhi(n, mul) + n) >> shift
    struct FastDiv
    {
        u32 mul, shift, d, pad;
    };
    FN2_DI u32 fast_div( u32 n, const FastDiv& fd ) { return ( __umulhi( n, fd.mul ) + n ) >> fd.shift; }

    struct GridArgs
    {
        float off[4];   // world position of index 0 on each axis
        float step[4];  // distance between samples on each axis
        int count[4];   // FULL grid size on each axis
        u32 rowBegin;   // first x-row of this launch, counted over the FULL grid (row = linear index / count[0])
        u32 total;      // points in this launch ( < 2^31 )
        u32 vecStore;   // 1: output pointer is 16-byte aligned, vector stores allowed
        u32 groups;     // total / points per thread: the grid-stride loop bound
        FastDiv divX, divY, divZ;
        // Fused all-gather: the same elements are also stored, as they are produced, into the copies of the volume that other GPUs
        // hold (peer memory over NVLink; fnb200GenUniformGrid*Shared). mirror[k] points at THIS launch's first element in copy k.
        float* mirror[7];
        u32 mirrors;
        u32 pad2;
    };

    struct TileArgs
    {
        float xFreq, yFreq, xMul, yMul; // [Generator.inl:386-392] computed on the host in the same float arithmetic
        int xSize, ySize;
        u32 rowBegin;
        u32 total;
        u32 vecStore;
        u32 groups;
        u32 pad[2];
        FastDiv divX;
        float* mirror[7];
        u32 mirrors;
        u32 pad2;
    };

    struct PosArgs
    {
        const float* pos[4];
        float off[4];
        unsigned long long count;
        unsigned long long groups; // ceil( count / points per thread )
        u32 vecIO; // 1: out and every position array are 16-byte aligned
        u32 pad;
    };

    // order-preserving float -> unsigned map; 0 is below every float, so zeroed memory is the empty accumulator
    FN2_DI u32 order_key( float v )
    {
        u32 b = (u32)fbits( v );
        return b ^ ( ( (i32)b >> 31 ) | 0x80000000u );
    }

    // Block-wide min/max into mm[0] = max key(-v), mm[1] = max key(v). Every thread of the block must call it.
    FN2_DI void reduce_minmax( float vmin, float vmax, bool valid, u32* __restrict__ mm )
    {
        if( !mm ) return;
        u32 kmin = valid ? order_key( -vmin ) : 0u;
        u32 kmax = valid ? order_key( vmax ) : 0u;
        kmin = __reduce_max_sync( 0xffffffffu, kmin );
        kmax = __reduce_max_sync( 0xffffffffu, kmax );

        __shared__ u32 sMin[32], sMax[32];
        const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
        if( lane == 0 )
        {
            sMin[warp] = kmin;
            sMax[warp] = kmax;
        }
        __syncthreads();
        if( warp == 0 )
        {
            const u32 warps = ( blockDim.x + 31u ) >> 5;
            kmin = lane < warps ? sMin[lane] : 0u;
            kmax = lane < warps ? sMax[lane] : 0u;
            kmin = __reduce_max_sync( 0xffffffffu, kmin );
            kmax = __reduce_max_sync( 0xffffffffu, kmax );
            if( lane == 0 )
            {
                if( kmin ) atomicMax( mm, kmin );
                if( kmax ) atomicMax( mm + 1, kmax );
            }
        }
    }

    // width-generic glue between the per-point arrays of the entry kernels and the value type of the node functions
    template<class F> FN2_DI F index_positions( u32 index, float step, float offset );        // fma( float(index + lane), step, offset )
    template<> __device__ __forceinline__ float index_positions<float>( u32 index, float step, float offset ) { return fma_( cvt_f( (i32)index ), step, offset ); }
    template<> __device__ __forceinline__ fpair index_positions<fpair>( u32 index, float step, float offset )
    {
        return fma_( fpair( cvt_f( (i32)index ), cvt_f( (i32)( index + 1u ) ) ), fpair( step ), fpair( offset ) );
    }
    template<class F> FN2_DI F index_positions( u32 index, float scale );                      // float(index + lane) * scale
    template<> __device__ __forceinline__ float index_positions<float>( u32 index, float scale ) { return cvt_f( (i32)index ) * scale; }
    template<> __device__ __forceinline__ fpair index_positions<fpair>( u32 index, float scale ) { return fpair( cvt_f( (i32)index ), cvt_f( (i32)( index + 1u ) ) ) * fpair( scale ); }
    template<class F> FN2_DI F lane_values( const float* v );
    template<> __device__ __forceinline__ float lane_values<float>( const float* v ) { return v[0]; }
    template<> __device__ __forceinline__ fpair lane_values<fpair>( const float* v ) { return fpair( v[0], v[1] ); }
    FN2_DI void spill_lanes( float r, float* v ) { v[0] = r; }
    FN2_DI void spill_lanes( fpair r, float* v ) { v[0] = r.x; v[1] = r.y; }

    template<int P>
    FN2_DI void store_points( float* __restrict__ out, u32 first, const float ( &v )[P], bool vec )
    {
        if( P == 8 && vec )
        {
            *reinterpret_cast<float4*>( out + first ) = make_float4( v[0], v[1], v[2], v[3] );
            *reinterpret_cast<float4*>( out + first + 4 ) = make_float4( v[P > 4 ? 4 : 0], v[P > 5 ? 5 : 0], v[P > 6 ? 6 : 0], v[P > 7 ? 7 : 0] );
        }
        else if( P == 4 && vec )
        {
            *reinterpret_cast<float4*>( out + first ) = make_float4( v[0], v[1], v[2], v[3] );
        }
        else if( P == 2 && vec )
        {
            *reinterpret_cast<float2*>( out + first ) = make_float2( v[0], v[1] );
        }
        else
        {
#pragma unroll
            for( int p = 0; p < P; p++ ) out[first + p] = v[p];
        }
    }
}
