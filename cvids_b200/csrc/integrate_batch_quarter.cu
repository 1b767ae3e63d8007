// The fused multi-frame kernels with QUARTER-brick tasks (8x8x2 voxels per warp, 4 voxels per lane): colour batches.
#define CHS_BATCH_VARIANT quarter
#define CHS_BRICK_SLICES 2
#define CHS_FAST_THREADS 256
#define CHS_FAST_MIN_CTAS 2
#include "integrate_batch_impl.cuh"
