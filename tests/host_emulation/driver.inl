// TEST INFRASTRUCTURE ONLY -- appended after a generated tree source: runs the entry kernel "thread" by "thread" on the host.
// Grid entries only (kinds 2, 3, 4); the arguments are filled the way csrc/backend.cpp LaunchGridRows does.
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

extern "C" void fn2_host_grid( float* out, int dims, const float* off, const int* count, const float* step, int seed, int pointsPerThread )
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
    a.rowBegin = 0;
    a.total = (unsigned)points;
    a.vecStore = 0;
    a.divX = fn2_host_fast_div( (unsigned)a.count[0] );
    a.divY = fn2_host_fast_div( (unsigned)a.count[1] );
    a.divZ = fn2_host_fast_div( (unsigned)a.count[2] );
    const unsigned long long threads = ( points + pointsPerThread - 1 ) / pointsPerThread;
    blockDim = { FN2_BLOCK, 1, 1 };
    gridDim = { (unsigned)( ( threads + FN2_BLOCK - 1 ) / FN2_BLOCK ), 1, 1 };
    for( unsigned b = 0; b < gridDim.x; b++ )
    {
        for( unsigned t = 0; t < FN2_BLOCK; t++ )
        {
            blockIdx = { b, 0, 0 };
            threadIdx = { t, 0, 0 };
            fn2_entry( out, a, nullptr, seed );
        }
    }
}
