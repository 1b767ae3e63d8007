/* chisel_oracle.h -- C ABI of the CPU restatement of OpenChisel's hot path.
 *
 * TEST INFRASTRUCTURE, NOT A PRODUCT COMPONENT. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load the library built from chisel_oracle.c, and only as the
 * checker or the timed CPU baseline -- never as a fallback for the CUDA path.
 *
 * Parity pin: the reference ships no tests or golden vectors. The restatement is
 * pinned (a) here in the build container against oracle/_ref (the unmodified reference sources compiled
 * against oracle/eigen_shim) by tests/test_oracle_vs_ref.py, bit for bit, and (b) everywhere against the
 * golden fixtures under tests/golden/ that oracle/_ref produced (tests/golden/make_golden.py).
 *
 * The entry points mirror oracle/ref_driver.cpp one for one (prefix orc_ instead of ref_).
 */
#ifndef CVIDS_CHISEL_ORACLE_H
#define CVIDS_CHISEL_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* pose: row-major 3x4 [R|t], camera -> world.   cam: fx fy cx cy W H near far (floats). */

void *orc_create(int chunkSize, float resolution, int useColor);
void orc_destroy(void *h);
void orc_reset(void *h);
/* truncKind: 0 constant (param = metres), 1 quadratic (param = scale), 2 inverse (param = scale) */
void orc_setup_integrator(void *h, int truncKind, float truncParam, float weight, int carve, float carveDist);
float orc_truncation(int truncKind, float truncParam, float depth);

void orc_integrate_depth(void *h, const float *depth, int W, int H, const float *pose, const float *cam);
void orc_integrate_color(void *h, const float *depth, int W, int H, const float *pose, const float *cam,
                         const uint8_t *color, int cW, int cH, int channels, const float *cpose,
                         const float *ccam, int unused);
int orc_candidate_ids(void *h, const float *pose, const float *cam, int *out, int cap);
void orc_frustum(const float *pose, const float *cam, float *corners, float *lines, float *planes);
void orc_update_meshes(void *h, int unused);

int orc_num_chunks(void *h);
void orc_chunk_ids(void *h, int *out);
int orc_chunk_voxels(void *h, const int *id, float *sdf, float *weight, uint8_t *rgbw);
void orc_set_chunk_voxels(void *h, const int *id, const float *sdf, const float *weight, const uint8_t *rgbw); /* test-only state injection */
int orc_num_dirty(void *h);
void orc_dirty_ids(void *h, int *out);
int orc_num_meshes(void *h);
void orc_mesh_ids(void *h, int *out);
int orc_mesh_sizes(void *h, const int *id, long *sizes);
int orc_mesh_data(void *h, const int *id, float *verts, float *normals, float *colors, float *grids, long *indices);
void orc_last_counts(void *h, long *out3);

/* Per-frame counters of the last integrate call (SURVEY.md section 8(d) definitions):
 * [0] candidates  [1] visited voxels  [2] N_upd  [3] N_carve  [4] N_col  [5] N_new (surviving)
 * [6] updated chunks  [7] garbage chunks */
void orc_frame_counters(void *h, long *out8);

#ifdef __cplusplus
}
#endif
#endif
