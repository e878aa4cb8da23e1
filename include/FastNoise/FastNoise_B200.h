/* fastnoise2_b200 -- device-side extension of the C ABI.
 *
 * The reference API (FastNoise_C.h) takes host pointers and has no notion of a device, a stream or a
 * slab of a larger grid. These entry points expose what the B200 backend adds, with plain C types only:
 *   - outputs / position arrays are CUDA device pointers and are written / read in place
 *   - `stream` is a cudaStream_t passed as void* (NULL = CUDA's default stream); calls only enqueue work and return
 *   - uniform grids take the FULL grid description plus an integer slab [slabBegin, slabEnd) along the
 *     outermost axis (y for 2D and tileable, z for 3D, w for 4D). Positions are derived from GLOBAL
 *     integer indices, fma( float(index), step, offset ), exactly as the reference's single-call loop
 *     does (Generator.inl:160-162), so N ranks each generating one slab produce bit-identical values
 *     to one call over the whole grid. `noiseOut` points at the first element OF THE SLAB.
 *   - min/max: `deviceMinMax` (may be NULL) is a device unsigned[2] accumulator, see fnb200MinMax*.
 * All functions return 0 on success or a negative FNB200_E* code; fnb200LastError() has the text.
 */
#ifndef FASTNOISE_B200_EXT_H
#define FASTNOISE_B200_EXT_H

#include "Utility/Export.h"

#ifdef __cplusplus
extern "C" {
#endif

enum
{
    FNB200_OK = 0,
    FNB200_EINVAL = -1,   /* bad argument (null node, negative count, empty slab ...) */
    FNB200_ECUDA = -2,    /* CUDA runtime error (no device, launch failure, out of memory) */
    FNB200_ECOMPILE = -3, /* NVRTC rejected the generated program */
    FNB200_EUNSUPPORTED = -4 /* tree contains a node type this build cannot lower */
};

/* Thread-local description of the last failure in this library. Every Gen* call clears it on entry, so "" after a call
 * that returned min > max means the output is genuine (every value NaN), not a failure. */
FASTNOISE_API const char* fnb200LastError( void );

/* 1 if a usable CUDA device (compute capability 10.x) and NVRTC are present. Never throws. */
FASTNOISE_API int fnb200DeviceAvailable( void );

/* Numerics mode for trees compiled AFTER the call (process-wide). Two independent bits:
 *   bit 0  0 = strict  : no FMA contraction beyond the reference's explicit FMulAdd sites, IEEE div/sqrt (parity mode)
 *          1 = relaxed : nvcc FMA contraction on, approximate rsqrt/rcp where the reference's RELAXED build uses them
 *   bit 1  0 = table gradients   : the reference's AVX512_F code path (permute tables; Utils.inl:124,167,196 and
 *                                  DomainWarpSimplex.inl:13,70,126,198,254,370), tables staged in shared memory
 *          1 = generic gradients : the select/xor code path its non-AVX-512 levels run
 * The two gradient flavours round differently (the reference's own output differs between its ISA levels in the same
 * way); each is bit-exact against the corresponding build of the reference sources. Default 0 = strict + tables. */
enum fnb200_math_mode { FNB200_MATH_STRICT = 0, FNB200_MATH_RELAXED = 1, FNB200_MATH_STRICT_GENERIC = 2, FNB200_MATH_RELAXED_GENERIC = 3 };
FASTNOISE_API int fnb200SetMathMode( int mode );
FASTNOISE_API int fnb200GetMathMode( void );

/* Lower + compile the tree for one entry kind ahead of time (otherwise done on first use).
 * kind: 2,3,4 = uniform grid 2D/3D/4D; 5 = tileable 2D; 12,13,14 = position array 2D/3D/4D. */
FASTNOISE_API int fnb200Prepare( const void* node, int kind );

/* Copies the generated CUDA source of (node, kind) into buf (NUL-terminated, truncated to bufSize); returns full length. */
FASTNOISE_API long fnb200GetKernelSource( const void* node, int kind, char* buf, long bufSize );

/* The float node parameters of the tree travel to the kernel as ONE argument (struct of floats) so that a compiled kernel serves
 * every tree of the same shape: changing a float through any setter costs no recompile. Copies the current values for
 * (node, kind), in the order the generated source indexes them (tp[k]), into buf (up to bufCount); returns their number. */
FASTNOISE_API long fnb200GetKernelParams( const void* node, int kind, float* buf, long bufCount );

/* Min/max accumulator helpers. The accumulator is two unsigned ints in device memory holding
 * order-preserving encodings of -min and +max; zero bytes = empty. Decode after the stream is idle. */
FASTNOISE_API int fnb200MinMaxReset( void* deviceMinMax, void* stream );
FASTNOISE_API int fnb200MinMaxRead( const void* deviceMinMax, void* stream, float* outputMinMax /* float[2] */ );

FASTNOISE_API int fnb200GenUniformGrid2D( const void* node, float* deviceOut, float xOffset, float yOffset, int xCount, int yCount,
                                          float xStepSize, float yStepSize, int seed, int slabBegin, int slabEnd,
                                          void* deviceMinMax, void* stream );
FASTNOISE_API int fnb200GenUniformGrid3D( const void* node, float* deviceOut, float xOffset, float yOffset, float zOffset,
                                          int xCount, int yCount, int zCount, float xStepSize, float yStepSize, float zStepSize,
                                          int seed, int slabBegin, int slabEnd, void* deviceMinMax, void* stream );
FASTNOISE_API int fnb200GenUniformGrid4D( const void* node, float* deviceOut, float xOffset, float yOffset, float zOffset, float wOffset,
                                          int xCount, int yCount, int zCount, int wCount,
                                          float xStepSize, float yStepSize, float zStepSize, float wStepSize,
                                          int seed, int slabBegin, int slabEnd, void* deviceMinMax, void* stream );
FASTNOISE_API int fnb200GenTileable2D( const void* node, float* deviceOut, int xSize, int ySize, float xStepSize, float yStepSize,
                                       int seed, int slabBegin, int slabEnd, void* deviceMinMax, void* stream );
/* dims = 2, 3 or 4; unused position arrays may be NULL. `count` positions starting at the given pointers. */
FASTNOISE_API int fnb200GenPositionArray( const void* node, int dims, float* deviceOut, long long count,
                                          const float* xPos, const float* yPos, const float* zPos, const float* wPos,
                                          float xOffset, float yOffset, float zOffset, float wOffset,
                                          int seed, void* deviceMinMax, void* stream );

/* Synthetic input for the position-array benchmark: fills out[i] (device) with uniform floats in
 * [lo, hi) from Philox4x32-10( key, counter = firstIndex + i, stream id = axis ). Reproducible on any rank. */
FASTNOISE_API int fnb200FillUniformPhilox( float* deviceOut, long long count, long long firstIndex, unsigned key, unsigned axis,
                                           float lo, float hi, void* stream );

/* ALU roofline probe: runs `iters` rounds of independent FFMA / IMAD / LOP3 (mode 0/1/2) or a 50:50
 * FFMA+LOP3 mix (mode 3) on every SM and returns achieved lane-ops per second (0 on failure). */
FASTNOISE_API double fnb200MeasureAluPeak( int mode, int iters );

/* Self-test: compares the check-free inverse square root the Cellular nodes use (valid on [0.5, 2^23)) with the IEEE
 * 1/sqrt sequence for EVERY float in [lo, hi). Returns the number of differing values (0 = identical), < 0 on a CUDA error. */
FASTNOISE_API long long fnb200SelfTestInvSqrt( float lo, float hi );

/* Self-test: evaluates leaf function `which` (0-23: simplex / supersimplex / perlin / value / white / cellular / gradient warp /
 * misc, see csrc/device_check.cu) of the ahead-of-time compiled device library at n points (n even) into hostOut, either one
 * point per value (packed = 0) or two points per value with packed FP32 instructions (packed = 1). The two must agree bit for bit. */
FASTNOISE_API int fnb200SelfTestLeaf( int packed, int which, int n, int seed, float step, float* hostOut );

/* Debug: the INTEGER intermediates of the lattice / cell search for a 3D uniform grid (positions fma( float(i), step, offset ) * frequency),
 * 8 ints per point into hostOut, written by the same device functions the generated kernels call:
 *   mode 0 (Simplex)  : primed base-cell coordinates x, y, z; HashPrimesHB of the 4 simplex vertices; [7] = vertex 0's product as the gradient lookup sees it
 *   mode 1 (Cellular) : primed base-cell coordinates x, y, z; HashPrimesHB of the closest cell; bit patterns of that cell's point x, y, z; 0
 * `jitter` is the Cellular jitter modifier (1 = default). The parity tests compare these with the oracle's (fn2oDebugGrid3D). */
FASTNOISE_API int fnb200DebugIntermediates( int mode, int seed, float frequency, float jitter, int xCount, int yCount, int zCount,
                                            float xOffset, float yOffset, float zOffset, float step, int* hostOut );

/* Decodes `encodedString` into NodeData records (Metadata::DeserialiseNodeData) and serialises them again
 * (Metadata::SerialiseNodeData): the canonical spelling of a tree. Writes a NUL-terminated string into buf
 * (truncated to bufSize) and returns its full length, 0 if the input does not decode. */
FASTNOISE_API long fnb200CanonicalEncodedNodeTree( const char* encodedString, char* buf, long bufSize );

/* ---- several GPUs, one process per GPU (the reference's model is "the caller splits one volume into calls", Generator.h:166-167)
 * Each rank generates its slab with fnb200GenUniformGrid* (GLOBAL integer indices: bit-identical to one call over the whole grid).
 * Two cross-rank steps exist, both NCCL collectives enqueued on a stream; NCCL is bound at run time (libnccl.so.2), no link dependency.
 *   fnb200CommGetUniqueId  : rank 0 fills a 128-byte id; the host application broadcasts it (MPI, torch.distributed, a file ...)
 *   fnb200CommInitRank     : every rank, with its GPU current -> communicator handle (NULL on failure)
 *   fnb200CommAllReduceMinMax : OutputMinMax of the whole volume -- MAX over ranks of the accumulator's two unsigned keys, in place,
 *                               on `stream`; read it afterwards with fnb200MinMaxRead. No host round trip.
 *   fnb200CommGatherSlabs  : OPTIONAL assembly of the whole volume on every rank (only for callers that need one contiguous
 *                            buffer): all-gather of ragged slabs, floatsPerRank[r] floats from rank r, in rank order, into
 *                            deviceVolume. deviceSlab = NULL means "my slab already lies at its place inside deviceVolume". */
FASTNOISE_API int fnb200CommGetUniqueId( void* uniqueId128 );
FASTNOISE_API void* fnb200CommInitRank( int worldSize, int rank, const void* uniqueId128 );
FASTNOISE_API int fnb200CommDestroy( void* comm );
FASTNOISE_API int fnb200CommAllReduceMinMax( void* comm, void* deviceMinMax, void* stream );
FASTNOISE_API int fnb200CommGatherSlabs( void* comm, const float* deviceSlab, float* deviceVolume, const long long* floatsPerRank, void* stream );

/* Fused generate + all-gather. A shared volume is one buffer of `floats` floats per rank, allocated by the library and mapped into every
 * other rank (CUDA IPC over NVLink; at most 8 ranks of one node). fnb200GenUniformGrid3DShared generates this rank's z-slab at its place
 * in the volume and stores every value into ALL ranks' copies as it is produced (peer stores from the generating kernel), so no separate
 * gather pass exists. A rank may read the whole volume once every rank's kernel has finished -- e.g. after fnb200CommAllReduceMinMax on
 * the same stream, which completes only when all ranks have reached it. Alloc / Free are collective (all ranks, same size). */
FASTNOISE_API void* fnb200CommAllocVolume( void* comm, long long floats, void* stream );
FASTNOISE_API float* fnb200VolumeData( void* sharedVolume );          /* this rank's copy (device pointer) */
FASTNOISE_API int fnb200CommFreeVolume( void* comm, void* sharedVolume, void* stream );
FASTNOISE_API int fnb200VolumeMirrors( void* sharedVolume, long long firstFloat, float** local, float** mirrors /* [7] */, int* mirrorCount );
FASTNOISE_API int fnb200GenUniformGrid3DShared( const void* node, void* sharedVolume, float xOffset, float yOffset, float zOffset,
                                                int xCount, int yCount, int zCount, float xStepSize, float yStepSize, float zStepSize,
                                                int seed, int slabBegin, int slabEnd, void* deviceMinMax, void* stream );

/* Number of kernels this library has launched in the calling process (all threads). */
FASTNOISE_API unsigned long long fnb200LaunchCount( void );

#ifdef __cplusplus
}
#endif

#endif
