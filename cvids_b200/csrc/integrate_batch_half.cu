// The fused multi-frame kernels with HALF-brick tasks (8x8x4 voxels per warp, 8 voxels per lane): depth-only batches.
#define CHS_BATCH_VARIANT half
#ifndef CHS_BRICK_SLICES
#define CHS_BRICK_SLICES 4
#endif
#ifndef CHS_FAST_THREADS
#define CHS_FAST_THREADS 128
#endif
#ifndef CHS_FAST_MIN_CTAS
#define CHS_FAST_MIN_CTAS 4
#endif
#include "integrate_batch_impl.cuh"
