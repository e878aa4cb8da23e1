// TEST INFRASTRUCTURE ONLY -- appended after a generated tree source: runs the entry kernel "thread" by "thread" on the host.
// FN2_HOST_ENTRY selects the driver for the entry kind of that source (0 uniform grid, 1 tileable, 2 position array); the
// arguments are filled the way csrc/backend.cpp LaunchGridRows / LaunchPositions fill them.
static fn2::FastDiv fn2_host_fast_div( unsigned d )
{
    fn2::FastDiv fd { 0, 0, d, 0 };
    if( d <= 1 ) return fd;
    unsigned shift = 0;
    while( ( 1ull << shift ) < d ) shift++;
    fd.shift = shift;
    fd.mul = (unsigned)( ( ( ( 1ull << shift ) - d ) << 32 ) / d + 1 );
    return fd;
}

// The launch is smaller than the call on purpose (at most 3 blocks), so that the kernels' grid-stride loops iterate.
template<class ARGS>
static void fn2_host_launch( float* out, ARGS& a, unsigned long long points, int pointsPerThread, int seed, const float* params, int paramCount )
{
    const unsigned long long groups = ( points + pointsPerThread - 1 ) / pointsPerThread;
    a.groups = groups;
    fn2::TreeParams tp {};
    for( int i = 0; i < paramCount && i < (int)( sizeof( tp.v ) / sizeof( float ) ); i++ ) tp.v[i] = params[i];
    blockDim = { FN2_BLOCK, 1, 1 };
    gridDim = { (unsigned)std::min<unsigned long long>( 3, ( groups + FN2_BLOCK - 1 ) / FN2_BLOCK ), 1, 1 };
    for( unsigned b = 0; b < gridDim.x; b++ )
    {
        for( unsigned t = 0; t < FN2_BLOCK; t++ )
        {
            blockIdx = { b, 0, 0 };
            threadIdx = { t, 0, 0 };
            fn2_entry( out, a, nullptr, seed, tp );
        }
    }
}

#if FN2_HOST_ENTRY == 0
extern "C" void fn2_host_grid( float* out, int dims, const float* off, const int* count, const float* step, int seed, int pointsPerThread, const float* params, int paramCount )
{
    fn2::GridArgs a {};
    unsigned long long points = 1;
    for( int d = 0; d < 4; d++ )
    {
        a.off[d] = d < dims ? off[d] : 0.0f;
        a.step[d] = d < dims ? step[d] : 0.0f;
        a.count[d] = d < dims ? count[d] : 1;
        points *= (unsigned)a.count[d];
    }
    a.total = (unsigned)points;
    a.divX = fn2_host_fast_div( (unsigned)a.count[0] );
    a.divY = fn2_host_fast_div( (unsigned)a.count[1] );
    a.divZ = fn2_host_fast_div( (unsigned)a.count[2] );
    fn2_host_launch( out, a, points, pointsPerThread, seed, params, paramCount );
}
#elif FN2_HOST_ENTRY == 1
extern "C" void fn2_host_tileable( float* out, int xSize, int ySize, float xStep, float yStep, int seed, int pointsPerThread, const float* params, int paramCount )
{
    fn2::TileArgs a {};
    const float pi2Recip = 0.15915493667f; // [Generator.inl:386-392], same float arithmetic as backend.cpp
    const float xSizePi = (float)xSize * pi2Recip, ySizePi = (float)ySize * pi2Recip;
    a.xFreq = xSizePi * xStep;
    a.yFreq = ySizePi * yStep;
    a.xMul = 1 / xSizePi;
    a.yMul = 1 / ySizePi;
    a.xSize = xSize;
    a.ySize = ySize;
    a.total = (unsigned)( xSize * ySize );
    a.divX = fn2_host_fast_div( (unsigned)xSize );
    fn2_host_launch( out, a, (unsigned long long)xSize * ySize, pointsPerThread, seed, params, paramCount );
}
#else
extern "C" void fn2_host_positions( float* out, int dims, long long count, const float* x, const float* y, const float* z, const float* w,
                                    const float* off, int seed, int pointsPerThread, const float* params, int paramCount )
{
    fn2::PosArgs a {};
    const float* pos[4] = { x, y, z, w };
    for( int d = 0; d < 4; d++ )
    {
        a.pos[d] = d < dims ? pos[d] : nullptr;
        a.off[d] = d < dims ? off[d] : 0.0f;
    }
    a.count = (unsigned long long)count;
    a.vecIO = 0;
    fn2_host_launch( out, a, (unsigned long long)count, pointsPerThread, seed, params, paramCount );
}
#endif
