/* ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into, loaded by or called from the product library.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use anything under oracle/.
 *
 * fn2_oracle.c: plain-C, scalar, one-point-at-a-time restatement of FastNoise2's node-graph evaluation
 * (reference: /root/reference/include/FastNoise/Generators/ .inl files, Generator.inl batch drivers, and the
 * encoded-tree decoder of src/FastNoise/Metadata.cpp + Base64.h). Every function cites the reference lines it
 * follows. It evaluates a decoded tree recursively per point -- no SIMD lanes, no code generation -- so it shares
 * neither structure nor code with the CUDA path it checks.
 *
 * Arithmetic contract = the STRICT semantics of oracle/shim (the stand-in for the absent FastSIMD dependency,
 * Auburn/FastSIMD @ 16450dae9528727e500e7254f635a671f9c7ee2d): IEEE single ops, fmaf exactly where the reference has
 * FS::FMulAdd/FNMulAdd/FMulSub, Convert<int> = cvtps2dq (round-to-nearest-even, 0x80000000 when out of range;
 * fn2oSetConvertMode( 1 ) switches to the saturating conversion of ARM / GPU builds for tests of that regime),
 * InvSqrt = 1/sqrt, and BOTH gradient flavours of the reference (fn2oSetFlavor: 0 = generic select/xor code paths,
 * 1 = the AVX512_F permute-table code paths). Pinned by tests/test_oracle.py against tests/golden/ (vectors produced
 * by the reference's own sources compiled over that shim, one set per flavour) and against those sources directly on
 * every node type and on fuzzed trees. The reference's tests hold no expected values: PARITY UNPINNED upstream.
 *
 * Build: gcc -O2 -std=c99 -ffp-contract=off -mfma -fPIC -shared fn2_oracle.c -o libfn2oracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int32_t i32;
typedef uint32_t u32;

/* ------------------------------------------------------------------ node table (reference headers) */
enum
{
    N_Constant, N_White, N_Checkerboard, N_SineWave, N_Gradient, N_DistanceToPoint, N_Simplex, N_SuperSimplex, N_Perlin, N_Value,
    N_CellularValue, N_CellularDistance, N_CellularLookup, N_FractalFBm, N_PingPong, N_FractalRidged,
    N_DomainWarpSimplex, N_DomainWarpSuperSimplex, N_DomainWarpGradient, N_DomainWarpFractalProgressive, N_DomainWarpFractalIndependent,
    N_Add, N_Subtract, N_Multiply, N_Divide, N_Abs, N_Min, N_Max, N_MinSmooth, N_MaxSmooth, N_SignedSquareRoot, N_PowFloat, N_PowInt,
    N_DomainScale, N_DomainOffset, N_DomainRotate, N_DomainAxisScale, N_SeedOffset, N_ConvertRGBA8, N_GeneratorCache,
    N_Fade, N_Remap, N_Terrace, N_AddDimension, N_RemoveDimension, N_Modulus, N_DomainRotatePlane, N_COUNT
};

/* how a variable write changes the node (the reference's setters) */
enum { V_F, V_I, V_IDX4, V_SCALE, V_OUTMIN, V_OUTMAX, V_YAW, V_PITCH, V_ROLL, V_STEPS };

#define MAXV 8
#define MAXH 5
#define MAXL 2

typedef struct
{
    int nvars, nlookups, nhybrids;
    unsigned char vop[MAXV];  /* V_* per variable index */
    unsigned char vslot[MAXV];/* which f[]/i[] entry a V_F / V_I / V_IDX4 write lands in */
    float hdef[MAXH];
} Spec;

typedef struct Node
{
    int id, supported;
    /* ScalableGenerator / Seeded / VariableRange state [BasicGenerators.h] */
    float scale, freq, rangeMin, rangeScale;
    i32 seedOffset;
    float f[8]; /* node specific floats */
    i32 i[4];   /* node specific ints */
    struct Node* lookups[MAXL];
    struct Node* hnode[MAXH];
    float hval[MAXH];
    int refs;
} Node;

static Spec gSpec[N_COUNT];
static int gSpecReady = 0;

static void spec_vars( int id, int n, const unsigned char* ops, const unsigned char* slots )
{
    gSpec[id].nvars = n;
    for( int k = 0; k < n; k++ ) { gSpec[id].vop[k] = ops[k]; gSpec[id].vslot[k] = slots[k]; }
}
static void spec_hybrids( int id, int n, const float* defs )
{
    gSpec[id].nhybrids = n;
    for( int k = 0; k < n; k++ ) gSpec[id].hdef[k] = defs[k];
}

static void build_specs( void )
{
    if( gSpecReady ) return;
    memset( gSpec, 0, sizeof( gSpec ) );
#define VARS( ID, ... ) { static const unsigned char o[] = { __VA_ARGS__ }; static const unsigned char s[MAXV] = { 0, 1, 2, 3, 4, 5, 6, 7 }; (void)s; spec_vars( ID, (int)sizeof( o ), o, slots ); }
    /* variable tables: op per index; slot = running count of node-specific floats / ints */
    {   const unsigned char slots[MAXV] = { 0 }; VARS( N_Constant, V_F ) }
    {   const unsigned char slots[MAXV] = { 0 }; VARS( N_White, V_I /*seed off handled below*/, V_OUTMIN, V_OUTMAX ) }
    {   const unsigned char slots[MAXV] = { 0 }; VARS( N_Checkerboard, V_SCALE, V_OUTMIN, V_OUTMAX ) VARS( N_SineWave, V_SCALE, V_OUTMIN, V_OUTMAX ) }
    {   const unsigned char slots[MAXV] = { 0, 1, 2, 3 }; VARS( N_Gradient, V_F, V_F, V_F, V_F ) }
    {   const unsigned char slots[MAXV] = { 0 }; VARS( N_DistanceToPoint, V_I ) }
    {   const unsigned char slots[MAXV] = { 0 };
        VARS( N_Simplex, V_SCALE, V_I, V_OUTMIN, V_OUTMAX ) VARS( N_SuperSimplex, V_SCALE, V_I, V_OUTMIN, V_OUTMAX )
        VARS( N_Perlin, V_SCALE, V_I, V_OUTMIN, V_OUTMAX ) VARS( N_Value, V_SCALE, V_I, V_OUTMIN, V_OUTMAX ) }
    {   const unsigned char slots[MAXV] = { 0, 0, 0, 0, 0, 1, 2, 3 };
        VARS( N_CellularValue, V_SCALE, V_I, V_OUTMIN, V_OUTMAX, V_I, V_IDX4 )
        VARS( N_CellularDistance, V_SCALE, V_I, V_OUTMIN, V_OUTMAX, V_I, V_IDX4, V_IDX4, V_I ) }
    {   const unsigned char slots[MAXV] = { 0, 0, 0 }; VARS( N_CellularLookup, V_SCALE, V_I, V_I ) }
    {   const unsigned char slots[MAXV] = { 0, 0 };
        VARS( N_FractalFBm, V_I, V_F ) VARS( N_FractalRidged, V_I, V_F )
        VARS( N_DomainWarpFractalProgressive, V_I, V_F ) VARS( N_DomainWarpFractalIndependent, V_I, V_F ) }
    {   const unsigned char slots[MAXV] = { 0, 0, 0, 1, 2, 3, 0 };
        VARS( N_DomainWarpGradient, V_SCALE, V_I, V_F, V_F, V_F, V_F )
        VARS( N_DomainWarpSimplex, V_SCALE, V_I, V_F, V_F, V_F, V_F, V_I ) VARS( N_DomainWarpSuperSimplex, V_SCALE, V_I, V_F, V_F, V_F, V_F, V_I ) }
    {   const unsigned char slots[MAXV] = { 0 };
        VARS( N_PowInt, V_I ) VARS( N_DomainScale, V_F ) VARS( N_SeedOffset, V_I ) VARS( N_Fade, V_I ) VARS( N_Remap, V_I )
        VARS( N_RemoveDimension, V_I ) VARS( N_DomainRotatePlane, V_I ) VARS( N_Terrace, V_STEPS ) }
    {   const unsigned char slots[MAXV] = { 0, 0, 0 }; VARS( N_DomainRotate, V_YAW, V_PITCH, V_ROLL ) }
    {   const unsigned char slots[MAXV] = { 0, 1, 2, 3 }; VARS( N_DomainAxisScale, V_F, V_F, V_F, V_F ) }
    {   const unsigned char slots[MAXV] = { 0, 1 }; VARS( N_ConvertRGBA8, V_F, V_F ) }
#undef VARS
    /* White's first variable is the seed offset of Seeded<Generator>; Seeded nodes store it in seedOffset: mark with slot 255 */
    gSpec[N_White].vslot[0] = 255;
    { static const int seeded[] = { N_Simplex, N_SuperSimplex, N_Perlin, N_Value, N_CellularValue, N_CellularDistance, N_CellularLookup,
                                    N_DomainWarpGradient, N_DomainWarpSimplex, N_DomainWarpSuperSimplex };
      for( unsigned k = 0; k < sizeof( seeded ) / sizeof( int ); k++ ) gSpec[seeded[k]].vslot[1] = 255; }

    { static const int one[] = { N_CellularLookup, N_FractalFBm, N_PingPong, N_FractalRidged, N_DomainWarpSimplex, N_DomainWarpSuperSimplex, N_DomainWarpGradient,
                                 N_DomainWarpFractalProgressive, N_DomainWarpFractalIndependent, N_Add, N_Multiply, N_Abs, N_Min, N_Max, N_MinSmooth, N_MaxSmooth,
                                 N_SignedSquareRoot, N_PowInt, N_DomainScale, N_DomainOffset, N_DomainRotate, N_DomainAxisScale, N_SeedOffset, N_ConvertRGBA8,
                                 N_GeneratorCache, N_Remap, N_Terrace, N_AddDimension, N_RemoveDimension, N_DomainRotatePlane };
      for( unsigned k = 0; k < sizeof( one ) / sizeof( int ); k++ ) gSpec[one[k]].nlookups = 1; }
    gSpec[N_Fade].nlookups = 2;

    { static const float z4[] = { 0, 0, 0, 0 }; spec_hybrids( N_Gradient, 4, z4 ); spec_hybrids( N_DomainOffset, 4, z4 ); }
    { static const float d[] = { 0, 0, 0, 0, 1.5f }; spec_hybrids( N_DistanceToPoint, 5, d ); }
    { static const float d[] = { 1.5f, 1.0f, 0.0f }; spec_hybrids( N_CellularValue, 3, d ); spec_hybrids( N_CellularDistance, 3, d ); spec_hybrids( N_CellularLookup, 3, d ); }
    { static const float d[] = { 0.5f, 0.0f }; spec_hybrids( N_FractalFBm, 2, d ); spec_hybrids( N_FractalRidged, 2, d );
      spec_hybrids( N_DomainWarpFractalProgressive, 2, d ); spec_hybrids( N_DomainWarpFractalIndependent, 2, d ); }
    { static const float d[] = { 2.0f }; spec_hybrids( N_PingPong, 1, d ); }
    { static const float d[] = { 50.0f }; spec_hybrids( N_DomainWarpSimplex, 1, d ); spec_hybrids( N_DomainWarpSuperSimplex, 1, d ); spec_hybrids( N_DomainWarpGradient, 1, d ); }
    { static const float d[] = { 0.0f }; spec_hybrids( N_Add, 1, d ); spec_hybrids( N_Multiply, 1, d ); spec_hybrids( N_Min, 1, d ); spec_hybrids( N_Max, 1, d );
      spec_hybrids( N_Terrace, 1, d ); spec_hybrids( N_AddDimension, 1, d ); }
    { static const float d[] = { 0.0f, 0.0f }; spec_hybrids( N_Subtract, 2, d ); spec_hybrids( N_Divide, 2, d ); spec_hybrids( N_Modulus, 2, d ); }
    { static const float d[] = { 0.0f, 0.1f }; spec_hybrids( N_MinSmooth, 2, d ); spec_hybrids( N_MaxSmooth, 2, d ); }
    { static const float d[] = { 2.0f, 2.0f }; spec_hybrids( N_PowFloat, 2, d ); }
    { static const float d[] = { 0.0f, -1.0f, 1.0f }; spec_hybrids( N_Fade, 3, d ); }
    { static const float d[] = { -1.0f, 1.0f, 0.0f, 1.0f }; spec_hybrids( N_Remap, 4, d ); }
    gSpecReady = 1;
}

static Node* node_new( int id )
{
    Node* n = (Node*)calloc( 1, sizeof( Node ) );
    n->id = id;
    n->supported = 1;
    n->scale = 100.0f; n->freq = 1.0f / 100; n->rangeMin = -1.0f; n->rangeScale = 2.0f; n->refs = 1;
    for( int k = 0; k < gSpec[id].nhybrids; k++ ) n->hval[k] = gSpec[id].hdef[k];
    switch( id ) /* member defaults of the node classes */
    {
    case N_Constant: n->f[0] = 1.0f; break;
    case N_DistanceToPoint: case N_CellularValue: case N_CellularDistance: case N_CellularLookup: n->i[0] = 1; break; /* EuclideanSquared */
    case N_FractalFBm: case N_FractalRidged: case N_DomainWarpFractalProgressive: case N_DomainWarpFractalIndependent: n->i[0] = 3; n->f[0] = 2.0f; break;
    case N_DomainWarpGradient: case N_DomainWarpSimplex: case N_DomainWarpSuperSimplex: case N_DomainAxisScale: n->f[0] = n->f[1] = n->f[2] = n->f[3] = 1.0f; break;
    case N_PowInt: n->i[0] = 2; break;
    case N_DomainScale: n->f[0] = 1.0f; break;
    case N_DomainRotate: n->f[0] = 1.0f; n->f[2] = 1.0f; n->f[4] = 1.0f; break; /* cos yaw/pitch/roll; sines at f[1],f[3],f[5] */
    case N_SeedOffset: n->i[0] = 1; break;
    case N_ConvertRGBA8: n->f[0] = -1.0f; n->f[1] = 1.0f; break;
    case N_Terrace: n->f[0] = 1.0f; n->f[1] = 1.0f; n->f[2] = 0.0f; break; /* steps, 1/steps, 1 + 1/smoothness */
    case N_RemoveDimension: n->i[0] = 1; break; /* Dim::Y */
    }
    if( id == N_CellularDistance ) { n->i[1] = 0; n->i[2] = 1; n->i[3] = 0; }
    return n;
}

static void node_release( Node* n )
{
    if( !n || --n->refs > 0 ) return;
    for( int k = 0; k < MAXL; k++ ) node_release( n->lookups[k] );
    for( int k = 0; k < MAXH; k++ ) node_release( n->hnode[k] );
    free( n );
}

static void set_variable( Node* n, int idx, i32 bits )
{
    const Spec* s = &gSpec[n->id];
    float v; memcpy( &v, &bits, 4 );
    int slot = s->vslot[idx];
    switch( s->vop[idx] )
    {
    case V_F: n->f[slot] = v; break;
    case V_I: if( slot == 255 ) n->seedOffset = bits; else n->i[slot] = bits; break;
    case V_IDX4: n->i[slot] = bits < 0 ? 0 : ( bits > 3 ? 3 : bits ); break;        /* Cellular.h:68,118 */
    case V_SCALE: n->scale = v; n->freq = 1.0f / v; break;                          /* BasicGenerators.h SetScale */
    case V_OUTMIN: n->rangeScale += n->rangeMin - v; n->rangeMin = v; break;        /* VariableRange::SetOutputMin */
    case V_OUTMAX: n->rangeScale = v - n->rangeMin; break;
    case V_YAW: n->f[0] = cosf( v ); n->f[1] = sinf( v ); break;                    /* Modifiers.h:84-86 (std::cos on float) */
    case V_PITCH: n->f[2] = cosf( v ); n->f[3] = sinf( v ); break;
    case V_ROLL: n->f[4] = cosf( v ); n->f[5] = sinf( v ); break;
    case V_STEPS: n->f[0] = v; n->f[1] = 1 / v; break;
    }
}

static void set_hybrid_value( Node* n, int idx, float v )
{
    node_release( n->hnode[idx] );
    n->hnode[idx] = NULL;
    n->hval[idx] = v;
    if( n->id == N_Terrace && idx == 0 && v != 0.0f ) n->f[2] = 1 + 1 / v;
}

/* ------------------------------------------------------------------ decoder [Base64.h:94-192, Metadata.cpp:310-497] */
static int sextet( unsigned char c )
{
    if( c >= 'A' && c <= 'Z' ) return c - 'A';
    if( c >= 'a' && c <= 'z' ) return c - 'a' + 26;
    if( c >= '0' && c <= '9' ) return c - '0' + 52;
    return c == '+' ? 62 : ( c == '/' ? 63 : -1 );
}

typedef struct { const unsigned char* p; size_t n, pos; int pendingEnds; Node** done; int ndone, cap; } Stream;

static int rd( Stream* s, void* out, size_t n )
{
    if( s->pendingEnds > 0 || s->pos + n > s->n ) return 0;
    memcpy( out, s->p + s->pos, n ); s->pos += n; return 1;
}
static int rd_tag( Stream* s, int* type, int* index )
{
    unsigned char b;
    if( s->pendingEnds > 0 ) { s->pendingEnds--; *type = 4; *index = 0; return 1; }
    if( !rd( s, &b, 1 ) ) return 0;
    *type = b & 7; *index = b >> 3;
    if( *type == 4 ) s->pendingEnds = *index;
    return 1;
}

static int is_warp( int id ) { return id == N_DomainWarpSimplex || id == N_DomainWarpSuperSimplex || id == N_DomainWarpGradient; }

static Node* read_node( Stream* s )
{
    unsigned char id;
    if( !rd( s, &id, 1 ) ) return NULL;
    if( id == 0xFF )
    {
        uint16_t ref;
        if( !rd( s, &ref, 2 ) || ref >= s->ndone ) return NULL;
        s->done[ref]->refs++;
        return s->done[ref];
    }
    if( id >= N_COUNT ) return NULL;
    Node* n = node_new( id );
    const Spec* sp = &gSpec[id];
    int type, index;
    if( !rd_tag( s, &type, &index ) ) goto fail;
    while( type == 0 )
    {
        i32 bits;
        if( !rd( s, &bits, 4 ) ) goto fail;
        if( index < sp->nvars ) set_variable( n, index, bits );
        if( !rd_tag( s, &type, &index ) ) goto fail;
    }
    if( type == 1 )
    {
        for( int k = 0; k < index; k++ )
        {
            Node* child = read_node( s );
            if( !child ) goto fail;
            if( k < sp->nlookups )
            {
                int needWarp = ( id == N_DomainWarpFractalProgressive || id == N_DomainWarpFractalIndependent );
                if( needWarp && !is_warp( child->id ) ) { node_release( child ); goto fail; }
                n->lookups[k] = child;
            }
            else node_release( child );
        }
        for( int k = index; k < sp->nlookups; k++ )
        {
            int needWarp = ( id == N_DomainWarpFractalProgressive || id == N_DomainWarpFractalIndependent );
            if( needWarp ) goto fail;
            n->lookups[k] = node_new( N_Constant ); /* Metadata.cpp:446-452 */
        }
        if( !rd_tag( s, &type, &index ) ) goto fail;
    }
    while( type != 4 )
    {
        if( type == 2 )
        {
            Node* child = read_node( s );
            if( index < sp->nhybrids )
            {
                if( !child ) goto fail;
                node_release( n->hnode[index] );
                n->hnode[index] = child;
            }
            else node_release( child );
        }
        else if( type == 3 )
        {
            float v;
            if( !rd( s, &v, 4 ) ) goto fail;
            if( index < sp->nhybrids ) set_hybrid_value( n, index, v );
        }
        if( !rd_tag( s, &type, &index ) ) goto fail;
    }
    if( s->ndone == s->cap ) { s->cap = s->cap ? s->cap * 2 : 16; s->done = (Node**)realloc( s->done, sizeof( Node* ) * (size_t)s->cap ); }
    s->done[s->ndone++] = n;
    return n;
fail:
    node_release( n );
    return NULL;
}

static int tree_supported( const Node* n )
{
    if( !n ) return 1;
    if( !n->supported ) return 0;
    for( int k = 0; k < MAXL; k++ ) if( !tree_supported( n->lookups[k] ) ) return 0;
    for( int k = 0; k < MAXH; k++ ) if( !tree_supported( n->hnode[k] ) ) return 0;
    return 1;
}

void* fn2oDecode( const char* text )
{
    build_specs();
    if( !text ) return NULL;
    size_t len = strlen( text ), ns = 0, pad = 0;
    unsigned char* six = (unsigned char*)malloc( len * 66 + 4 );
    for( size_t k = 0; k < len; k++ )
    {
        unsigned char c = (unsigned char)text[k];
        if( c == '@' )
        {
            int extra = k + 1 < len ? sextet( (unsigned char)text[++k] ) : -1;
            if( extra < 0 ) { free( six ); return NULL; }
            for( int r = 0; r < extra + 3; r++ ) six[ns++] = 0;
        }
        else if( c == '=' ) { six[ns++] = 0; pad++; }
        else
        {
            int v = sextet( c );
            if( v < 0 || pad ) { free( six ); return NULL; }
            six[ns++] = (unsigned char)v;
        }
    }
    if( ns == 0 || ns % 4 || pad > 2 ) { free( six ); return NULL; }
    size_t nb = ns / 4 * 3;
    unsigned char* bytes = (unsigned char*)malloc( nb );
    for( size_t k = 0, o = 0; k < ns; k += 4 )
    {
        u32 t = ( (u32)six[k] << 18 ) | ( (u32)six[k + 1] << 12 ) | ( (u32)six[k + 2] << 6 ) | six[k + 3];
        bytes[o++] = (unsigned char)( t >> 16 ); bytes[o++] = (unsigned char)( t >> 8 ); bytes[o++] = (unsigned char)t;
    }
    Stream s = { bytes, nb - pad, 0, 0, NULL, 0, 0 };
    Node* root = read_node( &s );
    free( s.done ); free( bytes ); free( six );
    return root;
}

void fn2oFree( void* node ) { node_release( (Node*)node ); }
int fn2oSupported( const void* node ) { return tree_supported( (const Node*)node ); }
int fn2oNodeId( const void* node ) { return ( (const Node*)node )->id; }

/* ------------------------------------------------------------------ scalar primitives (strict FastSIMD semantics, oracle/shim) */
static const i32 kPrime[4] = { (i32)0xF797C5C7u, (i32)0x6C060C89u, (i32)0x465FD04Fu, (i32)0xF7A62279u }; /* Utils.inl:8-13 */
#define HASH_MUL 0xB7E0A5F5u                                                                              /* Utils.inl:18 */

static i32 imul( i32 a, i32 b ) { return (i32)( (u32)a * (u32)b ); }
static i32 iadd( i32 a, i32 b ) { return (i32)( (u32)a + (u32)b ); }
static i32 isub( i32 a, i32 b ) { return (i32)( (u32)a - (u32)b ); }
static i32 shl( i32 a, int s ) { return (i32)( (u32)a << s ); }
static i32 lsr( i32 a, int s ) { return (i32)( (u32)a >> s ); }
static i32 asr( i32 a, int s ) { return a >> s; } /* arithmetic on every compiler this builds with */
static i32 f2b( float v ) { i32 b; memcpy( &b, &v, 4 ); return b; }
static float b2f( i32 b ) { float v; memcpy( &v, &b, 4 ); return v; }
static float fxor( float v, i32 bits ) { return b2f( f2b( v ) ^ bits ); }
/* FS::Convert<int32_t> = cvtps2dq: round to nearest even; NaN and anything outside [-2^31, 2^31) give the "integer indefinite" 0x80000000 */
/* fn2oSetConvertMode( 1 ): the SATURATING conversion of ISAs without an "indefinite" value (ARM fcvtns, the GPU's cvt.rni.s32.f32):
 * NaN -> 0, too large -> INT_MAX, too small -> INT_MIN. The reference's result for such positions (beyond +-2^31 cells, or NaN)
 * depends on the ISA it was built for; the tests use mode 1 to pin the product there too. Mode 0 (default) = x86. */
static int gSaturatingConvert = 0;
void fn2oSetConvertMode( int saturating ) { gSaturatingConvert = saturating != 0; }
static i32 cvt_i( float v )
{
    if( v >= -2147483648.0f && v < 2147483648.0f ) return (i32)lrintf( v );
    if( !gSaturatingConvert ) return (i32)0x80000000u;
    return v != v ? 0 : ( v > 0.0f ? 0x7fffffff : (i32)0x80000000u );
}
static float fmin_x86( float a, float b ) { return a < b ? a : b; } /* minps operand order */
static float fmax_x86( float a, float b ) { return a > b ? a : b; }
static float fnma( float a, float b, float c ) { return fmaf( -a, b, c ); }
static float inv_sqrt( float v ) { return 1.0f / sqrtf( v ); }
static float reciprocal( float v ) { return 1.0f / v; }
static float lerp( float a, float b, float t ) { return fmaf( t, b - a, a ); }                      /* Utils.inl:252 */
static float hermite( float t ) { return t * t * fnma( t, 2.0f, 3.0f ); }                            /* Utils.inl:257 */
static float quintic( float t ) { return t * t * t * fmaf( t, fmaf( t, 6.0f, -15.0f ), 10.0f ); }   /* Utils.inl:262 */

static i32 xor_primed( i32 seed, const i32* p, int dim ) { i32 h = seed; for( int d = 0; d < dim; d++ ) h ^= p[d]; return h; }
static i32 hash_primes( i32 seed, const i32* p, int dim ) { u32 h = (u32)xor_primed( seed, p, dim ) * HASH_MUL; return asr( (i32)h, 15 ) ^ (i32)h; } /* Utils.inl:221 */
static i32 hash_hb( i32 seed, const i32* p, int dim ) { return (i32)( (u32)xor_primed( seed, p, dim ) * HASH_MUL ); }                                /* Utils.inl:232 */
static float value_coord( i32 seed, const i32* p, int dim ) { u32 h = (u32)xor_primed( seed, p, dim ); h *= h * HASH_MUL; return (float)(i32)h; }     /* Utils.inl:242 */

/* gradient dot products.
 * The reference has two FLAVOURS of every gradient helper, picked at compile time by `if constexpr( SIMD & FeatureFlag::... )`:
 *   generic (gTables = 0): select + sign-xor forms -- what its Scalar/NEON/WASM (and, for most helpers, SSE/AVX2) levels run;
 *   tables  (gTables = 1): the AVX512_F branches -- a 16/32-entry weight table indexed by the hash, then one FMA chain.
 * They round differently (a table weight multiplies inside the chain), so each flavour is its own parity target:
 * oracle/_ref/libfn2ref_strict_avx2.so pins the first, libfn2ref_strict_avx512.so the second. */
static int gTables = 0;
void fn2oSetFlavor( int tables ) { gTables = tables != 0; }
int fn2oGetFlavor( void ) { return gTables; }

#define R2P1 ( 1.0f + (float)1.4142135623730950488016887242097 )
static const float kPerlin2X[8] = { R2P1, -R2P1, R2P1, -R2P1, 1, -1, 1, -1 };            /* Utils.inl:169, low 3 index bits (the table repeats) */
static const float kPerlin2Y[8] = { 1, 1, -1, -1, R2P1, R2P1, -R2P1, -R2P1 };            /* Utils.inl:170 */
static const float kCommon3[3][16] = {                                                    /* Utils.inl:126-128 */
    { 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0, 1, 0, -1, 0 },
    { 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1 },
    { 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1, 0, 1, 0, -1 } };
static const float kPerlin4[4][32] = {                                                    /* Utils.inl:198-201, entries 0-15 = first operand, 16-31 = second */
    { 0, 0, 0, 0, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1,   1, -1, 1, -1, 1, -1, 1, -1, 1, -1, 1, -1, 1, -1, 1, -1 },
    { 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0, 0, 0, 0, 0,   1, 1, -1, -1, 1, 1, -1, -1, 1, 1, -1, -1, 1, 1, -1, -1 },
    { 1, 1, -1, -1, 1, 1, -1, -1, 1, 1, -1, -1, 1, 1, -1, -1,   0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, -1, -1, -1, -1 },
    { 1, 1, 1, 1, -1, -1, -1, -1, 1, 1, 1, 1, -1, -1, -1, -1,   1, 1, 1, 1, -1, -1, -1, -1, 0, 0, 0, 0, 0, 0, 0, 0 } };

static float grad2( i32 h, float x, float y ) /* Utils.inl:162-190 */
{
    if( gTables ) return fmaf( kPerlin2X[h & 7], x, y * kPerlin2Y[h & 7] );
    x = fxor( x, shl( h, 31 ) ); y = fxor( y, shl( asr( h, 1 ), 31 ) );
    int swap = shl( h, 29 ) < 0;
    return fmaf( R2P1, swap ? y : x, swap ? x : y );
}
static float grad3( i32 h, float x, float y, float z ) /* Utils.inl:121-158 */
{
    if( gTables ) return fmaf( kCommon3[0][h & 15], x, fmaf( y, kCommon3[1][h & 15], z * kCommon3[2][h & 15] ) );
    i32 a = h & 13;
    float u = a > 7 ? y : x;
    float v = a == 12 ? x : z;
    if( a < 2 ) v = y;
    return fxor( u, shl( h, 31 ) ) + fxor( v, shl( asr( h, 1 ), 31 ) );
}
static float grad4( i32 h, float x, float y, float z, float w ) /* Utils.inl:193-218 */
{
    if( gTables ) return fmaf( kPerlin4[0][h & 31], x, fmaf( y, kPerlin4[1][h & 31], fmaf( z, kPerlin4[2][h & 31], w * kPerlin4[3][h & 31] ) ) );
    i32 p = h & ( 3 << 3 );
    float a = p > 0 ? x : y, b = shl( h, 27 ) < 0 ? y : z, c = p > ( 2 << 3 ) ? z : w;
    return fxor( a, shl( h, 31 ) ) + fxor( b, shl( asr( h, 1 ), 31 ) ) + fxor( c, shl( asr( h, 2 ), 31 ) );
}
static float grad_n( int dim, i32 h, const float* d ) { return dim == 2 ? grad2( h, d[0], d[1] ) : ( dim == 3 ? grad3( h, d[0], d[1], d[2] ) : grad4( h, d[0], d[1], d[2], d[3] ) ); }

/* shim restatements of FS::Cos / Sin / Pow / Modulus (oracle/shim/FastSIMD/Functions.h) */
static float cos_poly( float x )
{
    float v = fabsf( x );
    v -= floorf( v * 0.1591549f ) * 6.283185f;
    int ge1 = v >= 1.570796f, ge3 = v >= 4.7123889f;
    float a = ge3 ? 6.283185f - v : ( ge1 ? 3.141593f - v : v );
    float c = a * a;
    float r = fmaf( c, fmaf( c, 0.03679168f, -0.49558072f ), 0.99940307f );
    return ( ge1 && !ge3 ) ? -r : r;
}
static float sin_poly( float x ) { return cos_poly( 1.570796f - x ); }
static float pow_pos( float x, float y )
{
    i32 xb = f2b( x ), e = ( ( xb >> 23 ) & 0xff ) - 127;
    float m = b2f( ( xb & 0x007fffff ) | 0x3f800000 );
    if( m > 1.41421354f ) { m = m * 0.5f; e += 1; }
    float s = ( m - 1.0f ) / ( m + 1.0f ), s2 = s * s;
    float p = fmaf( s2, 0.111111112f, 0.142857149f );
    p = fmaf( s2, p, 0.2f ); p = fmaf( s2, p, 0.333333343f ); p = fmaf( s2, p, 1.0f );
    float z = y * fmaf( ( s + s ) * p, 1.44269502f, (float)e );
    z = z < -126.0f ? -126.0f : z; z = z > 127.0f ? 127.0f : z;
    float n = floorf( z + 0.5f ), u = ( z - n ) * 0.693147182f;
    float q = fmaf( u, 1.98412701e-4f, 1.38888892e-3f );
    q = fmaf( u, q, 8.33333377e-3f ); q = fmaf( u, q, 4.16666679e-2f ); q = fmaf( u, q, 0.166666672f );
    q = fmaf( u, q, 0.5f ); q = fmaf( u, q, 1.0f ); q = fmaf( u, q, 1.0f );
    return q * b2f( shl( (i32)n + 127, 23 ) );
}

/* ------------------------------------------------------------------ evaluation */
static float eval( const Node* n, i32 seed, int dim, const float* p );

static float hybrid( const Node* n, int idx, i32 seed, int dim, const float* p ) /* Generator.inl:48-58 */
{
    return n->hnode[idx] ? eval( n->hnode[idx], seed, dim, p ) : n->hval[idx];
}

static float range_out( const Node* n, float v, float lo, float hi ) /* ScaleOutput, BasicGenerators.inl:27-30 */
{
    return fmaf( ( 1.0f / ( hi - lo ) ) * n->rangeScale, v - lo, n->rangeMin );
}

/* CalcDistance [Utils.inl:281-335]; minkowski exponent is sampled AT the delta vector */
static float calc_distance( const Node* n, int pIdx, int df, int doSqrt, i32 seed, int dim, const float* d )
{
    float r;
    switch( df )
    {
    default:
    case 0:
        if( doSqrt ) { r = d[0] * d[0]; for( int k = 1; k < dim; k++ ) r = fmaf( d[k], d[k], r ); return sqrtf( r ); }
        /* fall through */
    case 1: r = d[0] * d[0]; for( int k = 1; k < dim; k++ ) r = fmaf( d[k], d[k], r ); return r;
    case 2: { float t = fabsf( d[dim - 1] ); for( int k = dim - 2; k >= 1; k-- ) t = fabsf( d[k] ) + t; return fabsf( d[0] ) + t; } /* right fold */
    case 3: r = fmaf( d[0], d[0], fabsf( d[0] ) ); for( int k = 1; k < dim; k++ ) r += fmaf( d[k], d[k], fabsf( d[k] ) ); return r;
    case 4: r = fabsf( d[0] ); for( int k = 1; k < dim; k++ ) r = fmax_x86( fabsf( d[k] ), r ); return r;
    case 5:
    {
        float mp = hybrid( n, pIdx, seed, dim, d );
        float t = pow_pos( fabsf( d[dim - 1] ), mp );
        for( int k = dim - 2; k >= 1; k-- ) t = pow_pos( fabsf( d[k] ), mp ) + t;
        return pow_pos( pow_pos( fabsf( d[0] ), mp ) + t, reciprocal( mp ) );
    }
    }
}

/* ---- lattice noises; positions are already scaled */
static float noise_value( i32 seed, int dim, const float* p ) /* Value.inl:7-95 */
{
    i32 c0[4], c1[4]; float t[4];
    for( int d = 0; d < dim; d++ )
    {
        float fl = floorf( p[d] );
        c0[d] = imul( cvt_i( fl ), kPrime[d] ); c1[d] = iadd( c0[d], kPrime[d] );
        t[d] = hermite( p[d] - fl );
    }
    float v[16];
    for( int corner = 0; corner < ( 1 << dim ); corner++ )
    {
        i32 q[4];
        for( int d = 0; d < dim; d++ ) q[d] = ( corner >> d ) & 1 ? c1[d] : c0[d];
        v[corner] = value_coord( seed, q, dim );
    }
    for( int d = 0; d < dim; d++ ) for( int k = 0; k < ( 1 << ( dim - d - 1 ) ); k++ ) v[k] = lerp( v[2 * k], v[2 * k + 1], t[d] );
    return v[0];
}

static float noise_perlin( i32 seed, int dim, const float* p ) /* Perlin.inl:7-125 */
{
    i32 c0[4], c1[4]; float f0[4], f1[4], t[4];
    for( int d = 0; d < dim; d++ )
    {
        float fl = floorf( p[d] );
        c0[d] = imul( cvt_i( fl ), kPrime[d] ); c1[d] = iadd( c0[d], kPrime[d] );
        f0[d] = p[d] - fl; f1[d] = f0[d] - 1.0f; t[d] = quintic( f0[d] );
    }
    float v[16];
    for( int corner = 0; corner < ( 1 << dim ); corner++ )
    {
        i32 q[4]; float f[4];
        for( int d = 0; d < dim; d++ ) { int hi = ( corner >> d ) & 1; q[d] = hi ? c1[d] : c0[d]; f[d] = hi ? f1[d] : f0[d]; }
        v[corner] = grad_n( dim, hash_primes( seed, q, dim ), f );
    }
    for( int d = 0; d < dim; d++ ) for( int k = 0; k < ( 1 << ( dim - d - 1 ) ); k++ ) v[k] = lerp( v[2 * k], v[2 * k + 1], t[d] );
    return v[0];
}

static float pow4( float v ) { v *= v; v *= v; return v; }

/* ---- simplex-lattice geometry. The Simplex / SuperSimplex noises and the DomainWarpSimplex / DomainWarpSuperSimplex
 * displacement fields visit the same lattice vertices [Simplex.inl:7-936, DomainWarpSimplex.inl:771-1833]; a producer lists
 * them in the reference's order (hash of the lattice point, offset to it, falloff^4) and each consumer folds them its way. */
typedef struct { i32 hash; float d[4]; float falloff; } Vert;

/* Debug sink (fn2oDebugGrid3D): when set, the 3D Simplex lattice and the CellularLookup search write their INTEGER intermediates
 * here, 8 ints per point -- primed base-cell coordinates, HashPrimesHB [Utils.inl:232-240] of the simplex vertices / of the closest
 * cell -- so the tests can compare them with the GPU's (fnb200DebugIntermediates) directly instead of inferring them from floats. */
static i32* gDebug = 0;

static int simplex_vertices( i32 seed, int dim, const float* p, Vert* out )
{
    if( dim == 2 ) /* Simplex.inl:7-56 */
    {
        const double R3 = 1.7320508075688772935274463415059, SK = 1.0 / ( R3 + 1.0 ), US = -1.0 / ( R3 + 3.0 );
        float x = p[0], y = p[1];
        float sd = (float)SK * ( x + y ), xs = x + sd, ys = y + sd, xb = floorf( xs ), yb = floorf( ys ), dxs = xs - xb, dys = ys - yb;
        i32 P[2] = { imul( cvt_i( xb ), kPrime[0] ), imul( cvt_i( yb ), kPrime[1] ) };
        int xge = dxs >= dys;
        float ud = (float)US * ( dxs + dys ), dx0 = dxs + ud, dy0 = dys + ud;
        float dx1 = ( xge ? dx0 : dx0 + 1.0f ) - (float)( US + 1 ), dy1 = ( xge ? dy0 + 1.0f : dy0 ) - (float)( US + 1 );
        float dx2 = dx0 - (float)( US * 2 + 1 ), dy2 = dy0 - (float)( US * 2 + 1 );
        float f0 = fnma( dx0, dx0, fnma( dy0, dy0, 0.5f ) ), f1 = fnma( dx1, dx1, fnma( dy1, dy1, 0.5f ) );
        float f2 = f0 + fmaf( ud, (float)( -4.0 * ( R3 + 2.0 ) / ( R3 + 3.0 ) ), (float)( -2.0 / 3.0 ) );
        i32 Q1[2] = { xge ? iadd( P[0], kPrime[0] ) : P[0], xge ? P[1] : iadd( P[1], kPrime[1] ) }, Q2[2] = { iadd( P[0], kPrime[0] ), iadd( P[1], kPrime[1] ) };
        out[0].hash = hash_primes( seed, P, 2 ); out[0].d[0] = dx0; out[0].d[1] = dy0; out[0].falloff = pow4( fmax_x86( f0, 0.0f ) );
        out[1].hash = hash_primes( seed, Q1, 2 ); out[1].d[0] = dx1; out[1].d[1] = dy1; out[1].falloff = pow4( fmax_x86( f1, 0.0f ) );
        out[2].hash = hash_primes( seed, Q2, 2 ); out[2].d[0] = dx2; out[2].d[1] = dy2; out[2].falloff = pow4( fmax_x86( f2, 0.0f ) );
        return 3;
    }
    if( dim == 3 ) /* Simplex.inl:65-135 */
    {
        const double SK = 1.0 / 3.0, RU = -1.0 / 2.0;
        const float R2 = (float)0.6;
        float x = p[0], y = p[1], z = p[2];
        float sd = (float)SK * ( x + y + z ), xs = x + sd, ys = y + sd, zs = z + sd;
        float xb = floorf( xs ), yb = floorf( ys ), zb = floorf( zs ), a = xs - xb, b = ys - yb, c = zs - zb;
        i32 P[3] = { imul( cvt_i( xb ), kPrime[0] ), imul( cvt_i( yb ), kPrime[1] ), imul( cvt_i( zb ), kPrime[2] ) };
        int xy = a >= b, yz = b >= c, xz = a >= c;
        float ud = (float)RU * ( a + b + c ), dx0 = a + ud, dy0 = b + ud, dz0 = c + ud;
        int mX1 = xy && xz, mY1 = yz && !xy, mZ1 = xz || yz, nX2 = xy || xz, nY2 = xy && !yz, nZ2 = xz && yz;
        float k3 = (float)( RU * 3 + 1 ), dx3 = dx0 - k3, dy3 = dy0 - k3, dz3 = dz0 - k3;
        float dx1 = mX1 ? dx3 - 1.0f : dx3, dy1 = mY1 ? dy3 - 1.0f : dy3, dz1 = mZ1 ? dz3 : dz3 - 1.0f;
        float dx2 = !nX2 ? dx0 + 1.0f : dx0, dy2 = nY2 ? dy0 + 1.0f : dy0, dz2 = nZ2 ? dz0 + 1.0f : dz0;
        float f0 = fnma( dz0, dz0, fnma( dy0, dy0, fnma( dx0, dx0, R2 ) ) ), f1 = fnma( dz1, dz1, fnma( dy1, dy1, fnma( dx1, dx1, R2 ) ) );
        float f2 = fnma( dz2, dz2, fnma( dy2, dy2, fnma( dx2, dx2, R2 ) ) ), f3 = f0 - ( ud + (float)( 3.0 / 4.0 ) );
        i32 Q1[3] = { mX1 ? iadd( P[0], kPrime[0] ) : P[0], mY1 ? iadd( P[1], kPrime[1] ) : P[1], mZ1 ? P[2] : iadd( P[2], kPrime[2] ) };
        i32 Q2[3] = { nX2 ? iadd( P[0], kPrime[0] ) : P[0], nY2 ? P[1] : iadd( P[1], kPrime[1] ), nZ2 ? P[2] : iadd( P[2], kPrime[2] ) };
        i32 Q3[3] = { iadd( P[0], kPrime[0] ), iadd( P[1], kPrime[1] ), iadd( P[2], kPrime[2] ) };
        const i32* Q[4] = { P, Q1, Q2, Q3 };
        float dd[4][3] = { { dx0, dy0, dz0 }, { dx1, dy1, dz1 }, { dx2, dy2, dz2 }, { dx3, dy3, dz3 } }, ff[4] = { f0, f1, f2, f3 };
        if( gDebug )
        {
            for( int k = 0; k < 3; k++ ) gDebug[k] = P[k];
            for( int v = 0; v < 4; v++ ) gDebug[3 + v] = hash_hb( seed, Q[v], 3 );
        }
        for( int v = 0; v < 4; v++ )
        {
            out[v].hash = hash_primes( seed, Q[v], 3 );
            for( int k = 0; k < 3; k++ ) out[v].d[k] = dd[v][k];
            out[v].falloff = pow4( fmax_x86( ff[v], 0.0f ) );
        }
        return 4;
    }
    { /* Simplex.inl:144-279 */
        const double R5 = 2.2360679774997896964091736687313, SK = 1.0 / ( R5 + 1.0 ), US = -1.0 / ( R5 + 5.0 );
        const float R2 = (float)0.6;
        float sd = (float)SK * ( p[0] + p[1] + p[2] + p[3] ), ds[4], d0[4]; i32 P[4];
        for( int k = 0; k < 4; k++ ) { float s = p[k] + sd, b = floorf( s ); ds[k] = s - b; P[k] = imul( cvt_i( b ), kPrime[k] ); }
        float ud = (float)US * ( ds[0] + ds[1] + ds[2] + ds[3] );
        for( int k = 0; k < 4; k++ ) d0[k] = ds[k] + ud;
        int rank[4] = { 0, 0, 0, 0 }; /* how many of the other coordinates each one is >= to [:180-207] */
        for( int a = 0; a < 4; a++ ) for( int b = a + 1; b < 4; b++ ) { if( d0[a] >= d0[b] ) rank[a]++; else rank[b]++; }
        float rawF0 = 0;
        for( int v = 0; v < 5; v++ ) /* vertex v = origin plus the v highest-ranked axes */
        {
            float* d = out[v].d; i32 Q[4];
            for( int k = 0; k < 4; k++ )
            {
                int stepped = v == 4 || ( v > 0 && rank[k] > 3 - v );
                Q[k] = stepped ? iadd( P[k], kPrime[k] ) : P[k];
                if( v == 0 ) d[k] = d0[k];
                else if( v == 4 ) d[k] = d0[k] - (float)( US * 4 + 1 );
                else d[k] = ( stepped ? d0[k] - 1.0f : d0[k] ) - (float)( US * v );
            }
            float fall;
            if( v == 4 ) fall = rawF0 + fmaf( ud, (float)( -4.0 * ( R5 + 3.0 ) / ( R5 + 5.0 ) ), (float)( -4.0 / 5.0 ) );
            else fall = fnma( d[3], d[3], fnma( d[2], d[2], fnma( d[1], d[1], fnma( d[0], d[0], R2 ) ) ) );
            if( v == 0 ) rawF0 = fall;
            out[v].falloff = pow4( fmax_x86( fall, 0.0f ) );
            out[v].hash = hash_primes( seed, Q, 4 );
        }
        return 5;
    }
}

static float noise_simplex( i32 seed, int dim, const float* p )
{
    Vert v[5];
    int n = simplex_vertices( seed, dim, p, v );
    float g[5];
    for( int k = 0; k < n; k++ ) g[k] = grad_n( dim, v[k].hash, v[k].d );
    if( dim == 3 ) /* fma( g3, f3, fma( g2, f2, fma( g1, f1, g0 * f0 ) ) ) [Simplex.inl:137] */
    {
        float value = g[0] * v[0].falloff;
        for( int k = 1; k < n; k++ ) value = fmaf( g[k], v[k].falloff, value );
        return value;
    }
    /* 2D and 4D fold from the last vertex inwards [Simplex.inl:58, 281] */
    float value = g[n - 1] * v[n - 1].falloff;
    for( int k = n - 2; k >= 0; k-- ) value = fmaf( g[k], v[k].falloff, value );
    return value;
}

/* ---- SuperSimplex [Simplex.inl:293-936] */
static float sq2( float v ) { return ( v * v ) * ( v * v ); }

/* closest-vertex search shared by 3D and 4D: a running best score and the axis mask of its vertex */
typedef struct { float best; i32 bits; } Closest;
static void consider( Closest* c, float score, i32 bits ) { if( score > c->best ) c->bits = bits; c->best = fmax_x86( c->best, score ); }

static int supersimplex_vertices( i32 seed, int dim, const float* p, Vert* out )
{
    if( dim == 2 ) /* :293-372 */
    {
        const double R3 = 1.7320508075688772935274463415059, SK = 1.0 / ( R3 + 1.0 ), US = -1.0 / ( R3 + 3.0 );
        const float R2 = (float)( 2.0 / 3.0 );
        float x = p[0], y = p[1];
        float sd = (float)SK * ( x + y ), xs = x + sd, ys = y + sd, xb = floorf( xs ), yb = floorf( ys ), dxs = xs - xb, dys = ys - yb;
        i32 X = imul( cvt_i( xb ), kPrime[0] ), Y = imul( cvt_i( yb ), kPrime[1] );
        int fwdXY = dxs + dys > 1.0f;
        float bound = fwdXY ? -1.0f : 0.0f;
        int fwdX = fmaf( dxs, -2.0f, dys ) < bound, fwdY = fmaf( dys, -2.0f, dxs ) < bound;
        float ud = (float)US * ( dxs + dys ), dx = dxs + ud, dy = dys + ud;
        i32 Q[2] = { X, Y };
        float base0 = fnma( dx, dx, fnma( dy, dy, R2 ) );
        out[0].hash = hash_primes( seed, Q, 2 ); out[0].d[0] = dx; out[0].d[1] = dy; out[0].falloff = pow4( base0 );
        Q[0] = iadd( X, kPrime[0] ); Q[1] = iadd( Y, kPrime[1] );
        out[1].hash = hash_primes( seed, Q, 2 ); out[1].d[0] = dx - (float)( 2 * US + 1 ); out[1].d[1] = dy - (float)( 2 * US + 1 );
        out[1].falloff = pow4( fmaf( ud, (float)( -4.0 * ( R3 + 2.0 ) / ( R3 + 3.0 ) ), base0 - R2 ) );
        float delta = fwdXY ? (float)( US + 1 ) : (float)( -US );
        dx -= delta; dy -= delta;
        {
            i32 qx = fwdX ? iadd( X, imul( kPrime[0], 2 ) ) : X;
            Q[0] = fwdXY ? qx : isub( qx, kPrime[0] ); Q[1] = fwdXY ? iadd( Y, kPrime[1] ) : Y;
            float ex = dx - ( fwdX ? (float)( 1 + 2 * US ) : -1.0f ), ey = fwdX ? dy - (float)( 2 * US ) : dy;
            out[2].hash = hash_primes( seed, Q, 2 ); out[2].d[0] = ex; out[2].d[1] = ey;
            out[2].falloff = pow4( fmax_x86( fnma( ex, ex, fnma( ey, ey, R2 ) ), 0.0f ) );
        }
        {
            i32 qy = fwdY ? iadd( Y, imul( kPrime[1], 2 ) ) : Y;
            Q[0] = fwdXY ? iadd( X, kPrime[0] ) : X; Q[1] = fwdXY ? qy : isub( qy, kPrime[1] );
            float ex = fwdY ? dx - (float)( 2 * US ) : dx, ey = dy - ( fwdY ? (float)( 1 + 2 * US ) : -1.0f );
            out[3].hash = hash_primes( seed, Q, 2 ); out[3].d[0] = ex; out[3].d[1] = ey;
            out[3].falloff = pow4( fmax_x86( fnma( ex, ex, fnma( ey, ey, R2 ) ), 0.0f ) );
        }
        return 4;
    }
    if( dim == 3 ) /* :379-572 */
    {
        const double SK = 1.0 / 3.0, RU = -1.0 / 2.0, TU = -1.0 / 4.0, A = 3.0 / 4.0, B = 1.0;
        float x = p[0], y = p[1], z = p[2];
        float sd = (float)SK * ( x + y + z ), s[3] = { x + sd, y + sd, z + sd }, base[3], d[3];
        for( int k = 0; k < 3; k++ ) { base[k] = floorf( s[k] ); d[k] = s[k] - base[k]; }
        {
            float tud = (float)TU * ( d[0] + d[1] + d[2] ), nx = d[0] + tud, ny = d[1] + tud, nz = d[2] + tud, nn = -tud;
            Closest c = { 0.375f, 0 };
            c.bits = nn > c.best ? -1 : 0; c.best = fmax_x86( c.best, nn );
            consider( &c, nx, 1 ); consider( &c, ny, 2 ); consider( &c, nz, 4 );
            c.best += 0.125f - nn;
            consider( &c, -nz, 3 ); consider( &c, -ny, 5 ); consider( &c, -nx, 6 );
            for( int k = 0; k < 3; k++ ) if( c.bits & ( 1 << k ) ) { base[k] = base[k] + 1.0f; d[k] = d[k] - 1.0f; }
        }
        i32 P[3];
        for( int k = 0; k < 3; k++ ) P[k] = imul( cvt_i( base[k] ), kPrime[k] );
        float sum = d[0] + d[1] + d[2], tud = (float)TU * sum, N[3] = { d[0] + tud, d[1] + tud, d[2] + tud }, nn = -tud;
        float ud = (float)RU * sum, D[3] = { d[0] + ud, d[1] + ud, d[2] + ud };
        float coordSum = (float)( 1 + 3 * RU ) * sum;

        float fb = fnma( D[2], D[2], fnma( D[1], D[1], fnma( D[0], D[0], (float)A ) ) ) * 0.5f;
        float stemA = fb - (float)( A * 0.5 ), stemB = fb - (float)( B * 0.5 );
        out[0].hash = hash_primes( seed, P, 3 ); out[0].falloff = sq2( fb );
        for( int k = 0; k < 3; k++ ) out[0].d[k] = D[k];

        /* the 7 neighbours, in the reference's order: mask of stepping axes */
        static const int kMask[7] = { 7, 3, 5, 6, 1, 2, 4 };
        for( int v = 0; v < 7; v++ )
        {
            int mask = kMask[v], axes = ( mask & 1 ) + ( ( mask >> 1 ) & 1 ) + ( ( mask >> 2 ) & 1 );
            int neg;
            if( axes == 3 ) neg = nn < 0.0f;
            else if( axes == 2 ) neg = nn < N[mask == 3 ? 2 : ( mask == 5 ? 1 : 0 )]; /* compared with the axis that does NOT step */
            else neg = N[mask == 1 ? 0 : ( mask == 2 ? 1 : 2 )] < 0.0f;
            i32 sign = neg ? (i32)0x80000000u : 0, Q[3];
            for( int k = 0; k < 3; k++ ) Q[k] = ( mask >> k ) & 1 ? iadd( P[k], neg ? isub( 0, kPrime[k] ) : kPrime[k] ) : P[k];
            float* e = out[v + 1].d; float fall;
            if( axes == 3 )
            {
                float o = fxor( (float)( 3 * RU + 1 ), sign );
                for( int k = 0; k < 3; k++ ) e[k] = D[k] - o;
                fall = fmax_x86( fmaf( o, coordSum, stemA ), 0.0f );
            }
            else if( axes == 2 )
            {
                int still = mask == 3 ? 2 : ( mask == 5 ? 1 : 0 );
                float o0 = fxor( (float)( 2 * RU ), sign );
                for( int k = 0; k < 3; k++ ) e[k] = k == still ? D[k] - o0 : D[k];
                fall = fmin_x86( fxor( D[still], sign ) - stemB, 0.0f );
            }
            else
            {
                int moving = mask == 1 ? 0 : ( mask == 2 ? 1 : 2 );
                float o0 = fxor( (float)RU, sign );
                for( int k = 0; k < 3; k++ ) e[k] = k == moving ? D[k] + o0 : D[k] - o0;
                fall = fmax_x86( fmaf( o0, coordSum, stemA ) + fxor( D[moving], sign ), 0.0f );
            }
            out[v + 1].hash = hash_primes( seed, Q, 3 );
            out[v + 1].falloff = sq2( fall );
        }
        return 8;
    }
    { /* :578-931 */
        const double R5 = 2.2360679774997896964091736687313, SK = 1.0 / ( R5 + 1.0 ), US = -1.0 / ( R5 + 5.0 ), TU = -1.0 / 5.0, A = 4.0 / 5.0, B = 6.0 / 5.0;
        float sd = (float)SK * ( p[0] + p[1] + p[2] + p[3] ), base[4], d[4];
        for( int k = 0; k < 4; k++ ) { float s = p[k] + sd; base[k] = floorf( s ); d[k] = s - base[k]; }
        {
            float tud = (float)TU * ( d[0] + d[1] + d[2] + d[3] ), nx = d[0] + tud, ny = d[1] + tud, nz = d[2] + tud, nw = d[3] + tud, nn = -tud;
            Closest c; c.best = 0.6f - nn; c.bits = 0.2f > c.best ? -1 : 0; c.best = fmax_x86( c.best, 0.2f );
            consider( &c, -nw, 7 ); consider( &c, -nz, 11 ); consider( &c, -ny, 13 ); consider( &c, -nx, 14 );
            c.best += nn - 0.2f;
            consider( &c, nx, 1 ); consider( &c, ny, 2 ); consider( &c, nz, 4 ); consider( &c, nw, 8 );
            c.best += 0.2f - nx;
            consider( &c, ny, 3 ); consider( &c, nz, 5 ); consider( &c, nw, 9 );
            c.best += nx;
            consider( &c, ny + nz, 6 );
            c.best -= nw;
            consider( &c, ny, 10 ); consider( &c, nz, 12 );
            for( int k = 0; k < 4; k++ ) if( c.bits & ( 1 << k ) ) { base[k] = base[k] + 1.0f; d[k] = d[k] - 1.0f; }
        }
        i32 P[4];
        for( int k = 0; k < 4; k++ ) P[k] = imul( cvt_i( base[k] ), kPrime[k] );
        float sum = d[0] + d[1] + d[2] + d[3], tud = (float)TU * sum, N[4], D[4], nn = -tud, ud = (float)US * sum;
        for( int k = 0; k < 4; k++ ) { N[k] = d[k] + tud; D[k] = d[k] + ud; }
        float coordSum = (float)( 1 + 4 * US ) * sum;

        float fb = fnma( D[3], D[3], fnma( D[2], D[2], fnma( D[1], D[1], fnma( D[0], D[0], (float)A ) ) ) ) * 0.5f;
        float stemA = fb - (float)( A * 0.5 ), stemB = fb - (float)( B * 0.5 );
        out[0].hash = hash_primes( seed, P, 4 ); out[0].falloff = sq2( fb );
        for( int k = 0; k < 4; k++ ) out[0].d[k] = D[k];

        /* the 15 neighbours in the reference's accumulation order, as masks of stepping axes (bit k = axis k) */
        static const int kMask[15] = { 15, 7, 11, 13, 14, 1, 3, 5, 9, 2, 6, 10, 4, 12, 8 };
        for( int v = 0; v < 15; v++ )
        {
            int mask = kMask[v], axes = 0, lo = -1, hi = -1, still = -1;
            for( int k = 0; k < 4; k++ ) { if( ( mask >> k ) & 1 ) { axes++; if( lo < 0 ) lo = k; hi = k; } else still = k; }
            int neg;
            if( axes == 4 ) neg = nn < 0.0f;
            else if( axes == 3 ) neg = nn < N[still];
            else if( axes == 2 ) neg = N[lo] < -N[hi];
            else neg = N[lo] < 0.0f;
            i32 sign = neg ? (i32)0x80000000u : 0, Q[4];
            for( int k = 0; k < 4; k++ ) Q[k] = ( mask >> k ) & 1 ? iadd( P[k], neg ? isub( 0, kPrime[k] ) : kPrime[k] ) : P[k];
            float o1 = fxor( (float)( axes * US + 1 ), sign ), o0 = fxor( (float)( axes * US ), sign ), fall;
            float* e = out[v + 1].d;
            for( int k = 0; k < 4; k++ ) e[k] = D[k] - ( ( mask >> k ) & 1 ? o1 : o0 );
            if( axes == 4 ) fall = fmaf( o1, coordSum, stemA );
            else if( axes == 3 ) fall = fmaf( o1, coordSum, stemB ) - fxor( D[still], sign );
            else if( axes == 2 ) fall = fmaf( o0, coordSum, stemB ) + fxor( D[lo] + D[hi], sign );
            else fall = fmaf( o0, coordSum, stemA ) + fxor( D[lo], sign );
            out[v + 1].hash = hash_primes( seed, Q, 4 );
            out[v + 1].falloff = sq2( fmax_x86( fall, 0.0f ) );
        }
        return 16;
    }
}

static float noise_supersimplex( i32 seed, int dim, const float* p )
{
    Vert v[16];
    int n = supersimplex_vertices( seed, dim, p, v );
    float value = v[0].falloff * grad_n( dim, v[0].hash, v[0].d );
    for( int k = 1; k < n; k++ ) value = fmaf( v[k].falloff, grad_n( dim, v[k].hash, v[k].d ), value );
    return value;
}

/* ---- vector contributions of the simplex warps, generic code paths [DomainWarpSimplex.inl:7-686] */
static float sel_hi( i32 selector, float a, float b ) { return selector < 0 ? a : b; }
#define SKEW2F ( (float)( 1.0 / ( 1.7320508075688772935274463415059 + 1.0 ) ) )
#define SKEW4F ( (float)( 1.0 / ( 2.2360679774997896964091736687313 + 1.0 ) ) )
#define ROOT2F ( (float)1.4142135623730950488016887242097 )
#define ROOT3F ( (float)1.7320508075688772935274463415059 )

static void gop4_direction( i32 index, float* g ) /* :149-166 */
{
    i32 a = index & ( 0x03 << 27 ), b = ( asr( index, 2 ) & ( 0x07 << 27 ) ) ^ a;
    int extra = b >= ( 0x04 << 27 ), equal = a == b;
    if( equal ) a = -1;
    float neutral = ( equal || extra ) ? ( extra ? SKEW4F * -1.0f : SKEW4F ) : 0.0f;
    for( int k = 0; k < 4; k++ )
    {
        float t = a == shl( k, 27 ) ? neutral - 1.0f : neutral;
        g[k] = b == shl( k, 27 ) ? t + 1.0f : t;
    }
}

static void vec_contribution( int scheme, int dim, i32 hash, const float* f, float m, float* value )
{
    /* tables flavour [:13-25, :70-81]: 12 directions in the low entries of a 16-entry table, one FMA chain */
    static const float kGop2X[12] = { ROOT3F, -ROOT3F, 1, -1, ROOT3F, -ROOT3F, -1, 1, 2, -2, 0, 0 };
    static const float kGop2Y[12] = { 1, -1, ROOT3F, -ROOT3F, -1, 1, ROOT3F, -ROOT3F, 0, 0, 2, -2 };
    static const float kGop3[3][12] = { { 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0 }, { 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1 }, { 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1 } };
    if( gTables && scheme == 1 && dim == 2 )
    {
        i32 hs = lsr( hash, 1 ), ig = asr( imul( hs, 3 ), 28 ) & 15, io = imul( hs, (i32)0xAAAAAAABu ) & (i32)0xC0000003u;
        io = ( io | asr( io, 28 ) ) & 15;
        m *= fmaf( f[1], kGop2Y[ig], f[0] * kGop2X[ig] );
        value[0] = fmaf( m, kGop2X[io], value[0] );
        value[1] = fmaf( m, kGop2Y[io], value[1] );
    }
    else if( gTables && scheme == 1 && dim == 3 )
    {
        i32 hs = lsr( hash, 1 ), ig = lsr( imul( hs, 3 ), 28 ), io = imul( hs, (i32)0xAAAAAAABu ) & (i32)0xC0000003u;
        io = ( io | asr( io, 28 ) ) & 15;
        m *= fmaf( kGop3[2][ig], f[2], fmaf( f[1], kGop3[1][ig], f[0] * kGop3[0][ig] ) );
        for( int k = 0; k < 3; k++ ) value[k] = fmaf( m, kGop3[k][io], value[k] );
    }
    else if( scheme == 1 && dim == 2 ) /* GradientOuterProduct 2D :7-60 */
    {
        i32 hs = lsr( hash, 1 ), ig = imul( hs, 3 ), io = imul( hs, (i32)0xAAAAAAABu ) & (i32)0xC0000003u;
        float u = sel_hi( shl( ig, 2 ), f[1], f[0] ), v = sel_hi( shl( ig, 2 ), f[0], f[1] );
        float a = u * sel_hi( ig, 2.0f, ROOT3F ), b = fxor( v, shl( asr( ig, 30 ), 31 ) );
        m *= fxor( ig >= 0 ? a + b : a, shl( asr( ig, 28 ) ^ io, 31 ) );
        a = m * sel_hi( io, 2.0f, ROOT3F );
        b = fxor( io >= 0 ? m : 0.0f, shl( asr( io, 30 ), 31 ) );
        value[0] += sel_hi( shl( io, 30 ), b, a );
        value[1] += sel_hi( shl( io, 30 ), a, b );
    }
    else if( scheme == 1 && dim == 3 ) /* :62-119 */
    {
        i32 hs = lsr( hash, 1 ), ig = lsr( imul( hs, 3 ), 28 ), io = imul( hs, (i32)0xAAAAAAABu ) & (i32)0xC0000003u;
        io |= asr( io, 28 );
        float u = ig >= 8 ? f[1] : f[0], v = ig >= 4 ? f[2] : f[1];
        m *= fxor( u, shl( ig, 31 ) ) + fxor( v, shl( asr( ig, 1 ), 31 ) );
        io &= 0xF;
        float s0 = fxor( m, shl( io, 31 ) ), s1 = fxor( m, shl( asr( io, 1 ), 31 ) );
        int notYZ = io < 8, notXY = io >= 4;
        if( notYZ ) value[0] = value[0] + s0;
        if( notXY ) value[2] = value[2] + s1;
        if( !( notYZ && notXY ) ) value[1] = value[1] + ( notXY ? s0 : s1 );
    }
    else if( scheme == 1 ) /* 4D :121-196 */
    {
        i32 hs = lsr( hash, 2 ), ig = imul( hs, 5 ), io = imul( hs, (i32)0xCCCCCCCDu );
        io = imul( io & (i32)0xE0000003u, 3 | ( 1 << 27 ) );
        float g[4];
        gop4_direction( ig, g );
        m *= fmaf( g[3], f[3], fmaf( g[2], f[2], fmaf( g[1], f[1], g[0] * f[0] ) ) );
        gop4_direction( io, g );
        for( int k = 0; k < 4; k++ ) value[k] = fmaf( m, g[k], value[k] );
    }
    else if( dim == 2 ) /* OrthogonalGradientMatrix 2D :199-243 */
    {
        i32 index = imul( lsr( hash, 1 ), 3 ), ofThree = lsr( index, 30 );
        float masked = index >= 0 ? SKEW2F : 0.0f;
        float gX = fxor( ofThree != 0 ? masked + 1.0f : masked, shl( asr( index, 28 ), 31 ) );
        float gY = fxor( ofThree != 1 ? masked + 1.0f : masked, shl( asr( index, 29 ), 31 ) );
        value[0] = fmaf( m, fmaf( f[1], gY, f[0] * gX ), value[0] );
        m = fxor( m, shl( hash, 31 ) );
        value[1] = fmaf( m, fmaf( f[1], gX, -( f[0] * gY ) ), value[1] ); /* FMulSub */
    }
    else if( dim == 3 ) /* :245-322 */
    {
        const float kA = 2.224744871391589f, kB = -0.224744871391589f, kC = -1.0f, kDE = 1.0f, kF = 2.0f;
        i32 hs = lsr( hash, 1 ), facet = imul( hs, (i32)0xAAAAAAABu ), permute2 = shl( facet, 29 ), permute3 = lsr( imul( hs, 3 ), 30 );
        i32 sign0 = shl( facet, 31 ), sign1 = shl( facet, 30 ) & (i32)0x80000000u, notYZ = facet, notXY = shl( facet, 1 );
        float comp[3][2] = { { kA, kB }, { kB, kA }, { kDE, kDE } }, other[3] = { kC, kC, kF }, val[3];
        for( int r = 0; r < 3; r++ ) /* rows A, B, C of the facet basis */
        {
            float gX = sel_hi( notYZ, other[r], fxor( comp[r][0], sign0 ) );
            float gY = sel_hi( notYZ | notXY, sel_hi( notXY, fxor( comp[r][1], sign1 ), fxor( comp[r][0], sign0 ) ), other[r] );
            float gZ = sel_hi( notXY, other[r], fxor( comp[r][1], sign1 ) );
            val[r] = fmaf( gZ, f[2], fmaf( f[1], gY, f[0] * gX ) );
        }
        float vAB = sel_hi( permute2, val[1], val[0] ), vBA = sel_hi( permute2, val[0], val[1] ), vC = val[2];
        m = fxor( m, shl( hash, 31 ) );
        value[0] = fmaf( m, permute3 == 0 ? vC : vAB, value[0] );
        value[1] = fmaf( m, permute3 == 1 ? vC : ( permute3 == 2 ? vBA : vAB ), value[1] );
        value[2] = fmaf( m, permute3 == 2 ? vC : vBA, value[2] );
    }
    else /* 4D :324-686 */
    {
        const float kIdxNegAB = -0.375999676691291f, kFillAB = 0.222726847849776f, kIdxPosD = -SKEW4F, kUnidxD = SKEW4F;
        const float kToSingleAB = -0.124000323308709f, kToSingleD = 0.190983005625053f, kSingleToExtra = SKEW4F, kABToC = 0.437016024448821f, kToDiag = -ROOT2F;
        i32 hs = lsr( hash, 2 ), basis = imul( hs, 5 ), permute3 = asr( imul( hs, (i32)0xD5555556u ), 29 ), permute8 = asr( basis, 24 );
        i32 pos = basis & ( 0x03 << 27 ), neg = ( asr( basis, 2 ) & ( 0x03 << 27 ) ) ^ pos;
        /* generic flavour: `indexBasis >= int32v( 0x10 << 27 )` compares with INT_MIN after the int32 broadcast, true for every lane.
         * tables flavour [:370-460]: the 32-entry tables hold the "extra" constants only in entries 16-19, i.e. the comparison as meant;
         * everything else in the tables is what the select logic below yields (checked against libfn2ref_strict_avx512.so). */
        int single = pos == neg, extra = gTables ? lsr( basis, 27 ) >= 16 : 1;
        if( gTables ? !extra && single : basis < ( 0x04 << 27 ) ) pos = -1;
        float offAB = extra ? kToSingleAB + kSingleToExtra : kToSingleAB, offD = extra ? kToSingleD + kSingleToExtra : kToSingleD;
        float idxNegAB = single ? kIdxNegAB + offAB : kIdxNegAB, fillAB = single ? kFillAB + offAB : kFillAB;
        float idxNegC = single ? idxNegAB : idxNegAB + kABToC, fillC = single ? fillAB : fillAB + kABToC;
        float idxPosD = single ? kIdxPosD + offD : kIdxPosD, unidxD = single ? kUnidxD + offD : kUnidxD;
        int counter = -1;
        float val[4] = { 0, 0, 0, 0 }; /* A, B, C, D */
        for( int axis = 0; axis < 4; axis++ )
        {
            int indexedPos = pos == shl( axis, 27 ), indexed = indexedPos || neg == shl( axis, 27 );
            float idxAB = indexedPos ? idxNegAB + 1.0f : idxNegAB, idxC = indexedPos ? idxNegC + 1.0f : idxNegC, idxD = !indexedPos ? idxPosD - 1.0f : idxPosD;
            /* the diagonal correction moves A -> B -> C as indexed axes are consumed: on axis k component j takes it when counter == j - k - 1
             * (component A on axis 0 unconditionally; components that cannot be reached on an axis never take it) */
            float un[3] = { fillAB, fillAB, fillC };
            static const int kWhen[4][3] = { { 99, 0, 0 }, { -2, -1, 0 }, { -3, -2, -1 }, { 0, -3, -2 } }; /* 99 = always, 0 = never */
            for( int j = 0; j < 3; j++ ) if( kWhen[axis][j] == 99 || ( kWhen[axis][j] != 0 && counter == kWhen[axis][j] ) ) un[j] = un[j] + kToDiag;
            float g[4] = { indexed ? idxAB : un[0], indexed ? idxAB : un[1], indexed ? idxC : un[2], indexed ? idxD : unidxD };
            for( int j = 0; j < 4; j++ ) val[j] = axis == 0 ? g[j] * f[0] : fmaf( g[j], f[axis], val[j] );
            if( indexed ) counter -= 1;
        }
        i32 ix = asr( permute8, 1 ), iy = iadd( iadd( ix, 1 ), permute3 ), iz = permute8 & 1;
        iz = iadd( iadd( ix, 1 ), iz >= permute3 ? iadd( iz, 1 ) : iz );
        i32 isum = iadd( iadd( ix, iy ), iz );
        m = fxor( m, shl( hash, 31 ) );
        i32 idx[3] = { ix, iy, iz };
        for( int k = 0; k < 3; k++ ) value[k] = fmaf( m, sel_hi( shl( idx[k], 31 ), sel_hi( shl( idx[k], 30 ), val[3], val[1] ), sel_hi( shl( idx[k], 30 ), val[2], val[0] ) ), value[k] );
        value[3] = fmaf( m, sel_hi( shl( isum, 31 ), sel_hi( shl( isum, 30 ), val[3], val[1] ), sel_hi( shl( isum, 30 ), val[0], val[2] ) ), value[3] );
    }
}

/* ------------------------------------------------------------------ Cellular [Cellular.inl:18-807] */
static const float kJitter[5] = { 0, 0, 0.437016f, 0.396144f, 0.366025f };

/* jitter vector of a cell from its hash bit-fields, and the factor that scales it to length `jitter` */
static float cell_vector( i32 hash, int dim, float jitter, float* v )
{
    if( dim == 2 ) /* :57-60 */
    {
        v[0] = (float)( hash & 0x7ff ) - ( 0x7ff / 2.0f );
        v[1] = (float)lsr( hash, 21 ) - ( 0x7ff / 2.0f );
        return jitter * inv_sqrt( fmaf( v[0], v[0], v[1] * v[1] ) );
    }
    if( dim == 3 ) /* :148-152 */
    {
        v[0] = (float)( hash & 0x3ff ) - ( 0x3ff / 2.0f );
        v[1] = (float)( asr( hash, 11 ) & 0x3ff ) - ( 0x3ff / 2.0f );
        v[2] = (float)lsr( hash, 22 ) - ( 0x3ff / 2.0f );
        return jitter * inv_sqrt( fmaf( v[0], v[0], fmaf( v[1], v[1], v[2] * v[2] ) ) );
    }
    v[0] = (float)( hash & 0xff ) - ( 0xff / 2.0f ); /* :247-252 */
    v[1] = (float)( asr( hash, 8 ) & 0xff ) - ( 0xff / 2.0f );
    v[2] = (float)( asr( hash, 16 ) & 0xff ) - ( 0xff / 2.0f );
    v[3] = (float)lsr( hash, 24 ) - ( 0xff / 2.0f );
    return jitter * inv_sqrt( fmaf( v[0], v[0], fmaf( v[1], v[1], fmaf( v[2], v[2], v[3] * v[3] ) ) ) );
}

static float size_jitter_factor( i32 hash, int dim, float sizeJitter )
{
    i32 field = dim == 2 ? ( asr( hash, 11 ) & 0x3ff ) : ( dim == 3 ? ( hash & 0xffff ) : ( hash & 0xfffff ) );
    return fnma( sizeJitter, (float)field, 1.0f );
}

/* mode 0 = CellularValue, 1 = CellularDistance, 2 = CellularLookup */
static float cellular( const Node* n, int mode, i32 seed, int dim, const float* in )
{
    i32 sourceSeed = seed;
    seed = iadd( seed, n->seedOffset );
    float jitter = kJitter[dim] * hybrid( n, 1, sourceSeed, dim, in );
    int sizeJ = n->hnode[2] || n->hval[2] != 0.0f;
    float sizeJitter = 0;
    if( sizeJ ) sizeJitter = hybrid( n, 2, sourceSeed, dim, in ) * ( dim == 2 ? -1.f / 0x3ff : ( dim == 3 ? -1.f / 0xffff : -1.f / 0xfffff ) );

    float p[4];
    for( int d = 0; d < dim; d++ ) p[d] = in[d] * n->freq;

    float distance[4] = { INFINITY, INFINITY, INFINITY, INFINITY };
    i32 valueHash[4] = { 0, 0, 0, 0 };
    float best = 3.402823466e+38f, cell[4] = { 0, 0, 0, 0 };

    i32 c0[4]; float cf0[4];
    for( int d = 0; d < dim; d++ ) { c0[d] = iadd( cvt_i( p[d] ), -1 ); cf0[d] = (float)c0[d]; }

    if( gDebug && dim == 3 && mode == 2 )
    {
        for( int d = 0; d < 3; d++ ) gDebug[d] = imul( c0[d], kPrime[d] );
        gDebug[3] = 0;
    }
    int total = dim == 2 ? 9 : ( dim == 3 ? 27 : 81 );
    for( int it = 0; it < total; it++ )
    {
        /* axis 0 is the OUTER loop of the reference, the last axis the innermost */
        int step[4], rem = it;
        for( int d = dim - 1; d >= 0; d-- ) { step[d] = rem % 3; rem /= 3; }

        i32 cp[4]; float delta[4], local[4], vec[4];
        for( int d = 0; d < dim; d++ )
        {
            /* primed cell coordinate: (c - 1) * prime, then + prime per step, all mod 2^32 */
            i32 c = imul( c0[d], kPrime[d] );
            for( int s = 0; s < step[d]; s++ ) c = iadd( c, kPrime[d] );
            cp[d] = c;
        }
        i32 hash = hash_hb( seed, cp, dim );
        float invMag = cell_vector( hash, dim, jitter, vec );
        for( int d = 0; d < dim; d++ )
        {
            if( mode == 0 )
            {   /* CellularValue: xcf = float(xc) - x once, then += 1 per step [:44-45,91-96] */
                float cf = cf0[d] - p[d];
                for( int s = 0; s < step[d]; s++ ) cf += 1.0f;
                delta[d] = fmaf( vec[d], invMag, cf );
            }
            else
            {   /* Distance: (float(xc) + steps) - x [:325,335]; Lookup: cell point first, then - x [:611-614] */
                float cf = cf0[d];
                for( int s = 0; s < step[d]; s++ ) cf += 1.0f;
                if( mode == 1 ) delta[d] = fmaf( vec[d], invMag, cf - p[d] );
                else { local[d] = fmaf( vec[d], invMag, cf ); delta[d] = local[d] - p[d]; }
            }
        }
        float nd = calc_distance( n, 0, n->i[0], mode == 0 && dim == 2, sourceSeed, dim, delta ); /* only CellularValue 2D asks for the sqrt [:63] */
        if( sizeJ ) nd *= size_jitter_factor( hash, dim, sizeJitter );

        if( mode == 0 ) /* ordered insert, entries 0..valueIndex [:72-89] */
        {
            i32 nh = hash;
            for( int i = 0; ; i++ )
            {
                int closer = nd < distance[i];
                float ld = distance[i]; i32 lh = valueHash[i];
                if( closer ) { distance[i] = nd; valueHash[i] = nh; }
                if( i >= n->i[1] ) break;
                if( closer ) { nd = ld; nh = lh; }
            }
        }
        else if( mode == 1 ) /* sorting network [:344-349] */
        {
            for( int i = 3; i > 0; i-- ) distance[i] = fmax_x86( fmin_x86( distance[i], nd ), distance[i - 1] );
            distance[0] = fmin_x86( distance[0], nd );
        }
        else
        {
            int closer = nd < best;
            best = fmin_x86( nd, best );
            if( closer ) for( int d = 0; d < dim; d++ ) cell[d] = local[d];
            if( closer && gDebug && dim == 3 ) { gDebug[3] = hash; for( int d = 0; d < 3; d++ ) gDebug[4 + d] = f2b( local[d] ); }
        }
    }

    if( mode == 0 ) return range_out( n, (float)valueHash[n->i[1]], -2147483648.f, 2147483648.f );
    if( mode == 2 )
    {
        float q[4];
        for( int d = 0; d < dim; d++ ) q[d] = cell[d] * n->scale;
        return eval( n->lookups[0], isub( sourceSeed, -1 ), dim, q ); /* [:633] */
    }
    /* GetReturn [:531-565] */
    int i0 = n->i[1], i1 = n->i[2];
    if( n->i[0] == 0 ) { distance[i0] *= inv_sqrt( distance[i0] ); distance[i1] *= inv_sqrt( distance[i1] ); }
    float maxDist = 1 + kJitter[dim];
    maxDist *= maxDist;
    switch( n->i[3] )
    {
    default:
    case 0: return range_out( n, distance[i0], 0, maxDist );
    case 1: return range_out( n, distance[i0] + distance[i1], 0, maxDist * 2 );
    case 2: return range_out( n, fabsf( distance[i0] - distance[i1] ), 0, maxDist );
    case 3: return range_out( n, distance[i0] * distance[i1], 0, maxDist * maxDist );
    case 4: return range_out( n, distance[i0] * reciprocal( distance[i1] ), 0, maxDist );
    }
}

/* ------------------------------------------------------------------ DomainWarpGradient::Warp [DomainWarp.inl:33-206] */
static float warp_gradient( const Node* n, i32 seed, float amp, int dim, const float* scaled, float* out )
{
    seed = iadd( seed, n->seedOffset );
    i32 c0[4], c1[4]; float w0[4], w1[4];
    for( int d = 0; d < dim; d++ )
    {
        float fl = floorf( scaled[d] );
        c0[d] = imul( cvt_i( fl ), kPrime[d] ); c1[d] = iadd( c0[d], kPrime[d] );
        w1[d] = hermite( scaled[d] - fl ); w0[d] = 1.0f - w1[d];
    }
    float warp[4] = { 0, 0, 0, 0 };
    float fieldMax = dim == 2 ? 0xffff : ( dim == 3 ? 0x3ff : 0xff );
    for( int corner = 0; corner < ( 1 << dim ); corner++ )
    {
        i32 q[4]; float contrib = 0;
        for( int d = 0; d < dim; d++ )
        {
            int hi = ( corner >> d ) & 1;
            q[d] = hi ? c1[d] : c0[d];
            contrib = d == 0 ? ( hi ? w1[0] : w0[0] ) : contrib * ( hi ? w1[d] : w0[d] );
        }
        i32 h = hash_hb( seed, q, dim );
        float field[4];
        if( dim == 2 ) { field[0] = (float)( h & 0xffff ); field[1] = (float)lsr( h, 16 ); }
        else if( dim == 3 ) { field[0] = (float)( h & 0x3ff ); field[1] = (float)( asr( h, 11 ) & 0x3ff ); field[2] = (float)lsr( h, 22 ); }
        else { field[0] = (float)( h & 0xff ); field[1] = (float)( asr( h, 8 ) & 0xff ); field[2] = (float)( asr( h, 16 ) & 0xff ); field[3] = (float)lsr( h, 24 ); }
        for( int d = 0; d < dim; d++ ) warp[d] = corner == 0 ? contrib * field[d] : fmaf( contrib, field[d], warp[d] );
    }
    float normalise = 1.0f / ( fieldMax / 2.0f );
    for( int d = 0; d < dim; d++ )
    {
        warp[d] = fmaf( warp[d], normalise, -1.0f );            /* FMulSub */
        out[d] = fmaf( warp[d], amp * n->f[d], out[d] );
    }
    float strength = warp[dim - 1] * warp[dim - 1];
    for( int d = dim - 2; d >= 0; d-- ) strength = fmaf( warp[d], warp[d], strength );
    return strength;
}

/* DomainWarpSimplex / DomainWarpSuperSimplex::Warp [DomainWarpSimplex.inl:728-1088, 1091-1836] */
static float warp_simplex( const Node* n, i32 seed, float amp, int dim, const float* scaled, float* out )
{
    int super = n->id == N_DomainWarpSuperSimplex, scheme = n->i[0] == 1 ? 1 : 0;
    static const float kBound[2][3][2] = {
        { { 70.1480577066486f, (float)( 49.918426513671875 / 2.0 ) }, { 16.281631889139874f, (float)( 32.69428253173828125 / 1.4142135623730951 ) },
          { 30.88161777516092f, (float)( 33.653125584827855 / 1.4142135623730951 ) } },
        { { 12.814453124999995f, (float)( 9.28993664146183 / 2.0 ) }, { 37.63698669623629f, (float)( 144.736422163332608 / 1.4142135623730951 ) },
          { 48.80058117543753f, (float)( 115.21625311930542 / 1.4142135623730951 ) } } };
    if( !super ) seed = iadd( seed, n->seedOffset ); /* only DomainWarpSimplex::Warp adds the offset itself [:733 vs :1094] */
    Vert v[16];
    int count = super ? supersimplex_vertices( seed, dim, scaled, v ) : simplex_vertices( seed, dim, scaled, v );
    float value[4] = { 0, 0, 0, 0 };
    for( int k = 0; k < count; k++ ) vec_contribution( scheme, dim, v[k].hash, v[k].d, v[k].falloff, value );
    if( dim == 3 && scheme != 0 ) /* match gradient orientation [:911-917, 1426-1433] */
    {
        float delta = ( super ? (float)( -2.0 / 3.0 ) : (float)( -2.0 / 2.0 ) ) * ( value[0] + value[1] + value[2] );
        for( int k = 0; k < 3; k++ ) value[k] += delta;
    }
    for( int k = 0; k < dim; k++ )
    {
        value[k] *= kBound[super][dim - 2][scheme] * n->f[k];
        out[k] = fmaf( value[k], amp, out[k] );
    }
    float strength = value[0] * value[0];
    for( int k = 1; k < dim; k++ ) strength = fmaf( value[k], value[k], strength );
    return strength;
}

static float warp_call( const Node* warpNode, i32 seed, float amp, int dim, const float* scaled, float* out )
{
    if( warpNode->id == N_DomainWarpGradient ) return warp_gradient( warpNode, seed, amp, dim, scaled, out );
    return warp_simplex( warpNode, seed, amp, dim, scaled, out );
}

/* ------------------------------------------------------------------ the recursive evaluator */
static float eval( const Node* n, i32 seed, int dim, const float* pin )
{
    float p[4] = { 0, 0, 0, 0 };
    for( int d = 0; d < dim; d++ ) p[d] = pin[d];
    const Node* src = n->lookups[0];

    switch( n->id )
    {
    case N_Constant: return n->f[0];

    case N_White: /* BasicGenerators.inl:48-58 */
    {
        i32 q[4];
        seed = iadd( seed, n->seedOffset );
        for( int d = 0; d < dim; d++ ) { i32 b = f2b( p[d] ); q[d] = imul( b ^ asr( b, 16 ), kPrime[d] ); }
        return range_out( n, value_coord( seed, q, dim ), -2147483648.f, 2147483648.f );
    }
    case N_Checkerboard: /* :66-75 */
    {
        i32 v = 0;
        for( int d = 0; d < dim; d++ ) v ^= cvt_i( p[d] * n->freq );
        return range_out( n, b2f( shl( v & 1, 30 ) ), 0, 2 );
    }
    case N_SineWave: /* :83-90, right fold of the product */
    {
        float v = sin_poly( p[dim - 1] * n->freq );
        for( int d = dim - 2; d >= 0; d-- ) v = sin_poly( p[d] * n->freq ) * v;
        return range_out( n, v, -1, 1 );
    }
    case N_Gradient: /* :98-108 */
    {
        float r = 0.0f;
        for( int d = 0; d < dim; d++ ) r = fmaf( p[d] + hybrid( n, d, seed, dim, p ), n->f[d], r );
        return r;
    }
    case N_DistanceToPoint: /* :115-125 */
    {
        float q[4];
        for( int d = 0; d < dim; d++ ) q[d] = p[d] - hybrid( n, d, seed, dim, p );
        return calc_distance( n, 4, n->i[0], 1, seed, dim, q );
    }
    case N_Simplex:
    {
        static const float kBound[5] = { 0, 0, 0.0261208079755306243896484375f, 0.030586399137973785400390625f, 0.036730043590068817138671875f };
        seed = iadd( seed, n->seedOffset );
        for( int d = 0; d < dim; d++ ) p[d] *= n->freq;
        float v = noise_simplex( seed, dim, p );
        return range_out( n, v, -kBound[dim], kBound[dim] );
    }
    case N_SuperSimplex:
    {
        static const float kBound[5] = { 0, 0, 0.14084912836551666259765625f, 0.0069091119803488254547119140625f, (float)0.011310013942420482635498046875 };
        seed = iadd( seed, n->seedOffset );
        for( int d = 0; d < dim; d++ ) p[d] *= n->freq;
        float v = noise_supersimplex( seed, dim, p );
        return range_out( n, v, -kBound[dim], kBound[dim] );
    }
    case N_Perlin:
    {
        static const float kBound[5] = { 0, 0, 1.726796627044677734375f, (float)1.0363423824310302734375, 1.33858621120452880859375f };
        seed = iadd( seed, n->seedOffset );
        for( int d = 0; d < dim; d++ ) p[d] *= n->freq;
        return range_out( n, noise_perlin( seed, dim, p ), -kBound[dim], kBound[dim] );
    }
    case N_Value:
        seed = iadd( seed, n->seedOffset );
        for( int d = 0; d < dim; d++ ) p[d] *= n->freq;
        return range_out( n, noise_value( seed, dim, p ), -2147483648.f, 2147483648.f );

    case N_CellularValue: return cellular( n, 0, seed, dim, p );
    case N_CellularDistance: return cellular( n, 1, seed, dim, p );
    case N_CellularLookup: return cellular( n, 2, seed, dim, p );

    case N_FractalFBm:    /* Fractal.inl:14-36 */
    case N_FractalRidged: /* Fractal.inl:44-66 */
    {
        int ridged = n->id == N_FractalRidged;
        float gain = hybrid( n, 0, seed, dim, p ), ws = hybrid( n, 1, seed, dim, p ), amp = 1.0f;
        float noise = eval( src, seed, dim, p );
        if( ridged ) noise = fabsf( noise );
        float sum = ridged ? ( noise * -2.0f + 1.0f ) * amp : noise * amp;
        for( int o = 1; o < n->i[0]; o++ )
        {
            seed = isub( seed, -1 );
            amp *= lerp( 1.0f, ridged ? 1.0f - noise : ( noise + 1.0f ) * 0.5f, ws );
            amp *= gain;
            for( int d = 0; d < dim; d++ ) p[d] *= n->f[0];
            noise = eval( src, seed, dim, p );
            if( ridged ) noise = fabsf( noise );
            sum += ridged ? ( noise * -2.0f + 1.0f ) * amp : noise * amp;
        }
        return sum;
    }
    case N_PingPong: /* Modifiers.inl:185-201 */
    {
        float strength = hybrid( n, 0, seed, dim, p );
        float t = eval( src, seed, dim, p ) * strength;
        t -= floorf( t * 0.5f ) * 2.0f;
        return t < 1.0f ? t : 2.0f - t;
    }
    case N_DomainWarpSimplex:
    case N_DomainWarpSuperSimplex:
    case N_DomainWarpGradient: /* DomainWarp::GenT, DomainWarp.inl:9-17 (Gradient and Simplex add the offset again inside Warp) */
    {
        i32 sourceSeed = seed;
        float scaled[4];
        seed = iadd( seed, n->seedOffset );
        float amp = hybrid( n, 0, sourceSeed, dim, p );
        for( int d = 0; d < dim; d++ ) scaled[d] = p[d] * n->freq;
        warp_call( n, seed, amp, dim, scaled, p );
        return eval( src, sourceSeed, dim, p );
    }
    case N_DomainWarpFractalProgressive: /* DomainWarpFractal.inl:9-35 */
    case N_DomainWarpFractalIndependent: /* DomainWarpFractal.inl:43-73 */
    {
        const Node* w = src;
        int independent = n->id == N_DomainWarpFractalIndependent;
        float noisePos[4], scaled[4];
        for( int d = 0; d < dim; d++ ) noisePos[d] = p[d];
        float amp = hybrid( w, 0, seed, dim, p ), ws = hybrid( n, 1, seed, dim, p ), freq = w->freq;
        i32 seedInc = seed;
        float gain = hybrid( n, 0, seed, dim, p );
        for( int d = 0; d < dim; d++ ) scaled[d] = ( independent ? noisePos[d] : p[d] ) * freq;
        float strength = warp_call( w, seedInc, amp, dim, scaled, p );
        for( int o = 1; o < n->i[0]; o++ )
        {
            strength = sqrtf( strength ); /* FastLengthSqrt, strict */
            seedInc = isub( seedInc, -1 );
            freq *= n->f[0];
            amp *= lerp( 1.0f, 1.0f - strength, ws );
            amp *= gain;
            for( int d = 0; d < dim; d++ ) scaled[d] = ( independent ? noisePos[d] : p[d] ) * freq;
            strength = warp_call( w, seedInc, amp, dim, scaled, p );
        }
        return eval( w->lookups[0], seed, dim, p );
    }

    /* Blends.inl */
    case N_Add: return eval( src, seed, dim, p ) + hybrid( n, 0, seed, dim, p );
    case N_Multiply: return eval( src, seed, dim, p ) * hybrid( n, 0, seed, dim, p );
    case N_Min: return fmin_x86( eval( src, seed, dim, p ), hybrid( n, 0, seed, dim, p ) );
    case N_Max: return fmax_x86( eval( src, seed, dim, p ), hybrid( n, 0, seed, dim, p ) );
    case N_Subtract: return hybrid( n, 0, seed, dim, p ) - hybrid( n, 1, seed, dim, p );
    case N_Divide: return hybrid( n, 0, seed, dim, p ) / hybrid( n, 1, seed, dim, p );
    case N_Modulus: { float a = hybrid( n, 0, seed, dim, p ), b = hybrid( n, 1, seed, dim, p ); return fnma( truncf( a / b ), b, a ); }
    case N_Abs: return fabsf( eval( src, seed, dim, p ) );
    case N_SignedSquareRoot: { float v = eval( src, seed, dim, p ); return b2f( f2b( sqrtf( fabsf( v ) ) ) | ( f2b( v ) & (i32)0x80000000u ) ); }
    case N_MinSmooth:
    case N_MaxSmooth: /* :126-162 */
    {
        int neg = n->id == N_MaxSmooth;
        float a = eval( src, seed, dim, p ), b = hybrid( n, 0, seed, dim, p );
        if( neg ) { a = -a; b = -b; }
        float smooth = fmax_x86( 1.17549435e-38f, fabsf( hybrid( n, 1, seed, dim, p ) ) );
        float h = fmax_x86( smooth - fabsf( a - b ), 0.0f );
        h *= reciprocal( smooth );
        float r = fnma( 1.0f / 6.0f, h * h * h * smooth, fmin_x86( a, b ) );
        return neg ? -r : r;
    }
    case N_PowFloat: { float v = fmax_x86( fabsf( hybrid( n, 0, seed, dim, p ) ), 1.17549435e-38f ); return pow_pos( v, hybrid( n, 1, seed, dim, p ) ); }
    case N_PowInt: { float v = eval( src, seed, dim, p ), r = v * v; for( int k = 2; k < n->i[0]; k++ ) r *= v; return r; }
    case N_Fade: /* :164-199 */
    {
        float fade = hybrid( n, 0, seed, dim, p ), lo = hybrid( n, 1, seed, dim, p ), hi = hybrid( n, 2, seed, dim, p );
        float range = hi - lo;
        fade = ( fade - lo ) / range;
        fade = fmax_x86( 0.0f, fmin_x86( 1.0f, fade ) );
        if( n->i[0] == 1 ) fade = hermite( fade );
        if( n->i[0] == 2 ) fade = quintic( fade );
        if( range == 0.0f ) fade = 0.5f;
        return lerp( eval( n->lookups[0], seed, dim, p ), eval( n->lookups[1], seed, dim, p ), fade );
    }

    /* Modifiers.inl */
    case N_DomainScale: for( int d = 0; d < dim; d++ ) p[d] *= n->f[0]; return eval( src, seed, dim, p );
    case N_DomainOffset: { float q[4]; for( int d = 0; d < dim; d++ ) q[d] = p[d] + hybrid( n, d, seed, dim, p ); return eval( src, seed, dim, q ); }
    case N_DomainRotate: /* :30-61 */
    {
        float yc = n->f[0], ys = n->f[1], pc = n->f[2], ps = n->f[3], rc = n->f[4], rs = n->f[5];
        if( dim == 4 ) return eval( src, seed, 4, p );
        if( dim == 2 && ps == 0.0f && rs == 0.0f )
        {
            float q[2] = { fnma( p[1], ys, p[0] * yc ), fmaf( p[0], ys, p[1] * yc ) };
            return eval( src, seed, 2, q );
        }
        float x = p[0], y = p[1], z = dim == 2 ? 0.0f : p[2];
        float xa = yc * pc, xb = yc * ps * rs - ys * rc, xc = yc * ps * rc + ys * rs; /* Modifiers.h CalculateRotation */
        float ya = ys * pc, yb = ys * ps * rs + yc * rc, ycc = ys * ps * rc - yc * rs;
        float za = -ps, zb = pc * rs, zc = pc * rc;
        float q[3] = { fmaf( x, xa, fmaf( y, xb, z * xc ) ), fmaf( x, ya, fmaf( y, yb, z * ycc ) ), fmaf( x, za, fmaf( y, zb, z * zc ) ) };
        return eval( src, seed, 3, q );
    }
    case N_DomainAxisScale: for( int d = 0; d < dim; d++ ) p[d] *= n->f[d]; return eval( src, seed, dim, p );
    case N_SeedOffset: return eval( src, iadd( seed, n->i[0] ), dim, p );
    case N_ConvertRGBA8: /* :103-127 */
    {
        float v = eval( src, seed, dim, p );
        v = fmin_x86( v, n->f[1] ); v = fmax_x86( v, n->f[0] ); v -= n->f[0];
        v *= 255.0f / ( n->f[1] - n->f[0] );
        i32 b = cvt_i( v );
        return b2f( (i32)0xff000000u | b | shl( b, 8 ) | shl( b, 16 ) );
    }
    case N_GeneratorCache: return eval( src, seed, dim, p );
    case N_Remap: /* :76-100 */
    {
        float v = eval( src, seed, dim, p );
        float fromMin = hybrid( n, 0, seed, dim, p ), fromMax = hybrid( n, 1, seed, dim, p ), toMin = hybrid( n, 2, seed, dim, p ), toMax = hybrid( n, 3, seed, dim, p );
        float r = toMin + ( ( v - fromMin ) / ( fromMax - fromMin ) * ( toMax - toMin ) );
        if( n->i[0] == 1 ) { r = fmax_x86( r, toMin ); r = fmin_x86( r, toMax ); }
        return r;
    }
    case N_Terrace: /* :130-182 */
    {
        float v = eval( src, seed, dim, p );
        v *= n->f[0];
        float rounded = rintf( v );
        int dynamic = n->hnode[0] != NULL;
        if( dynamic || n->hval[0] != 0.0f )
        {
            float smooth = dynamic ? hybrid( n, 0, seed, dim, p ) : 0.0f;
            float diff = rounded - v;
            int neg = diff < 0.0f;
            diff = fabsf( diff );
            diff = 0.5f - diff;
            diff *= dynamic ? 1.0f + reciprocal( smooth ) : n->f[2];
            diff = fmin_x86( diff, 0.5f );
            diff = neg ? 0.5f - diff : diff - 0.5f;
            if( !dynamic || smooth != 0.0f ) rounded = rounded + diff;
        }
        return rounded * n->f[1];
    }
    case N_AddDimension:
        if( dim == 4 ) return eval( src, seed, 4, p );
        p[dim] = hybrid( n, 0, seed, dim, p );
        return eval( src, seed, dim + 1, p );
    case N_RemoveDimension: /* :239-277 */
    {
        int removed = n->i[0];
        if( dim == 2 || removed < 0 || removed >= dim ) return eval( src, seed, dim, p );
        float q[4]; int k = 0;
        for( int d = 0; d < dim; d++ ) if( d != removed ) q[k++] = p[d];
        return eval( src, seed, dim - 1, q );
    }
    case N_DomainRotatePlane: /* :352-409 */
    {
        int xz = dim != 2 && n->i[0] == 1;
        if( dim != 2 && n->i[0] != 0 && n->i[0] != 1 ) return eval( src, seed, dim, p ); /* no case of the switch [:361-388]: unchanged */
        float x = p[0], y = p[1], z = dim == 2 ? 0.0f : p[2], q[4];
        float a = xz ? z : y, b = xz ? y : z;
        float pair = x + a, s2 = pair * -0.211324865405187f, nb = b * 0.577350269189626f;
        float nx = x + s2 - nb, na = a + s2 - nb;
        nb = nb + pair * 0.577350269189626f;
        q[0] = nx; q[1] = xz ? nb : na; q[2] = xz ? na : nb; q[3] = p[3];
        return eval( src, seed, dim == 2 ? 3 : dim, q );
    }
    }
    return 0.0f;
}

/* ------------------------------------------------------------------ batch drivers [Generator.inl:79-433] */
typedef struct { float mn, mx; } MinMax;
static void mm_add( MinMax* m, float v ) { m->mn = m->mn < v ? m->mn : v; m->mx = m->mx > v ? m->mx : v; }

/* Rows [rowBegin, rowBegin + rowCount) of a dims-D uniform grid of size n[], x fastest: out[0] is the first point of
 * rowBegin. Positions come from GLOBAL integer indices, fma( float(i), step, offset ) [Generator.inl:160-162]. */
void fn2oGrid( const void* node, float* out, int dims, const float* off, const int* n, const float* step, int seed,
               long long rowBegin, long long rowCount, float* minmax )
{
    MinMax mm = { INFINITY, -INFINITY };
    for( long long r = 0; r < rowCount; r++ )
    {
        long long row = rowBegin + r;
        int idx[4] = { 0, 0, 0, 0 };
        for( int d = 1; d < dims; d++ ) { idx[d] = (int)( row % n[d] ); row /= n[d]; }
        float p[4];
        for( int d = 1; d < dims; d++ ) p[d] = fmaf( (float)idx[d], step[d], off[d] );
        for( int x = 0; x < n[0]; x++ )
        {
            p[0] = fmaf( (float)x, step[0], off[0] );
            float v = eval( (const Node*)node, seed, dims, p );
            out[r * n[0] + x] = v;
            mm_add( &mm, v );
        }
    }
    if( minmax ) { minmax[0] = mm.mn; minmax[1] = mm.mx; }
}

/* Integer intermediates of a 3D uniform grid, 8 ints per point (see gDebug): the tree is evaluated point by point with the sink set.
 * Meant for trees whose root is Simplex or CellularLookup; not thread-safe. */
void fn2oDebugGrid3D( const void* node, int* out8, const float* off, const int* n, const float* step, int seed )
{
    size_t i = 0;
    for( int z = 0; z < n[2]; z++ )
        for( int y = 0; y < n[1]; y++ )
            for( int x = 0; x < n[0]; x++, i++ )
            {
                float p[3] = { fmaf( (float)x, step[0], off[0] ), fmaf( (float)y, step[1], off[1] ), fmaf( (float)z, step[2], off[2] ) };
                for( int k = 0; k < 8; k++ ) out8[i * 8 + k] = 0;
                gDebug = out8 + i * 8;
                (void)eval( (const Node*)node, seed, 3, p );
                gDebug = 0;
            }
}

/* GenTileable2D [Generator.inl:370-433], rows [rowBegin, rowBegin + rowCount) */
void fn2oTileable( const void* node, float* out, int xSize, int ySize, float xStep, float yStep, int seed, int rowBegin, int rowCount, float* minmax )
{
    MinMax mm = { INFINITY, -INFINITY };
    float pi2Recip = 0.15915493667f;
    float xSizePi = (float)xSize * pi2Recip, ySizePi = (float)ySize * pi2Recip;
    float xFreq = xSizePi * xStep, yFreq = ySizePi * yStep, xMul = 1 / xSizePi, yMul = 1 / ySizePi;
    for( int r = 0; r < rowCount; r++ )
    {
        float yF = (float)( rowBegin + r ) * yMul;
        for( int x = 0; x < xSize; x++ )
        {
            float xF = (float)x * xMul;
            float p[4] = { cos_poly( xF ) * xFreq, cos_poly( yF ) * yFreq, sin_poly( xF ) * xFreq, sin_poly( yF ) * yFreq };
            float v = eval( (const Node*)node, seed, 4, p );
            out[(size_t)r * xSize + x] = v;
            mm_add( &mm, v );
        }
    }
    if( minmax ) { minmax[0] = mm.mn; minmax[1] = mm.mx; }
}

/* GenPositionArray2D/3D/4D [Generator.inl:256-350]: pos = offset + array value */
void fn2oPositions( const void* node, float* out, int dims, long long count, const float* xs, const float* ys, const float* zs, const float* ws,
                    const float* off, int seed, float* minmax )
{
    MinMax mm = { INFINITY, -INFINITY };
    const float* arr[4] = { xs, ys, zs, ws };
    for( long long i = 0; i < count; i++ )
    {
        float p[4] = { 0, 0, 0, 0 };
        for( int d = 0; d < dims; d++ ) p[d] = off[d] + arr[d][i];
        float v = eval( (const Node*)node, seed, dims, p );
        out[i] = v;
        mm_add( &mm, v );
    }
    if( minmax ) { minmax[0] = mm.mn; minmax[1] = mm.mx; }
}
