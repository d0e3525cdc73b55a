/*
 * spnb.h -- C ABI of libspnb.so, the B200 (sm_100a) particle-interaction library.
 *
 * This is the drop-in seam for SmoothParticleNets' native layer: it replaces the reference's
 * `extern "C"` CUDA entry points in src/gpu_kernels.h:7-138 (and through them the pybind glue in
 * src/cuda_layer_funcs.cpp and the vendored CUB 1.3.2).  Every function takes raw DEVICE pointers,
 * plain sizes and a cudaStream_t passed as void*; there are no torch types here.
 *
 * Conventions shared by all entry points
 *   - all tensors are contiguous float32, including index-valued ones (idxs, neighbors, cellStarts,
 *     cellEnds, SDF idxs/offsets/shapes), exactly as in the reference;
 *     `cellIDs` holds uint32 keys in a float-typed buffer like the reference GPU path
 *     (gpu_kernels.cu:291);
 *   - return value: 1 on success, 0 on failure (the reference's convention, gpu_kernels.cu:26-34);
 *     on failure spnb_last_error() describes the problem;
 *   - everything is stream-ordered: no host synchronisation, no device allocation, no global device
 *     state is touched, so calls can be captured in CUDA graphs;
 *   - outputs are OVERWRITTEN (the reference accumulates into caller-zeroed buffers); callers need
 *     not pre-fill them;
 *   - there is no CPU fallback.
 */
#ifndef SPNB_H_
#define SPNB_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPNB_MAX_NDIM 19       /* ndim < MAX_CARTESIAN_DIM (src/constants.h:7) */
#define SPNB_NUM_KERNEL_FNS 12 /* kernels.py:123: alphabetical ids 0..11 */

/* Library version and the last error message of the calling thread ("" if none). */
int spnb_version(void);
const char* spnb_last_error(void);

/* Replaces spn_max_cartesian_dim (cpu_layer_funcs.cpp:29-32). */
int spnb_max_cartesian_dim(void);

/* Number of CUDA kernels this process has enqueued through the library so far (host-side counter;
 * bench accounting, no reference counterpart). */
unsigned long long spnb_launch_count(void);

/* ---- hash-grid neighbour search ------------------------------------------------------------- */

/* Bytes of device scratch needed by spnb_grid_bounds / spnb_hashgrid_order for these sizes.
 * Replaces get_radixsort_buffer_size (gpu_kernels.h:113). */
size_t spnb_hashgrid_workspace_bytes(int batch_size, int N, int ndims, int max_grid_dim);

/* Per-scene grid bounds, bit-identical to the float32 torch ops of ParticleCollision.forward
 * (ParticleCollision.py:174-181):
 *   grid_dims = ceil(clamp((max-min)/radius, 0, max_grid_dim)),
 *   low       = (min+max)/2 - grid_dims*radius/2.
 * low, grid_dims: [batch_size, ndims]. */
int spnb_grid_bounds(const float* locs, int batch_size, int N, int ndims, float radius,
                     int max_grid_dim, float* low, float* grid_dims, void* workspace,
                     size_t workspace_bytes, void* stream);

/* Cell keys + STABLE sort of (key, original index) per scene.  Replaces cuda_hashgrid_order
 * (gpu_kernels.h:63-74; kernel_compute_cellIDs + CUB SortPairs, gpu_kernels.cu:238-337).
 *   cellIDs: [batch_size, N] sorted uint32 keys;  idxs: [batch_size, N] float, idxs[b,i] = original
 *   index of the particle now at sorted position i (ties by ascending original index). */
int spnb_hashgrid_order(const float* locs, const float* low, const float* grid_dims, float* cellIDs,
                        float* idxs, void* workspace, size_t workspace_bytes, int batch_size, int N,
                        int ndims, float cellEdge, int max_grid_dim, void* stream);

/* Cell table + fixed-width neighbour lists.  Replaces cuda_compute_collisions
 * (gpu_kernels.h:76-95; kernel_fill_cells + kernel_compute_collisions, gpu_kernels.cu:339-485).
 *   cellStarts/cellEnds: [batch_size, ncells] float scratch tables (only the first
 *   prod(grid_dims[b]) entries of each row are defined afterwards);
 *   collisions: [batch_size, M, max_collisions] float: accepted particle indices in the reference's
 *   visiting order, then -1 up to the end of the row.
 *   trunc_flag (optional, may be NULL): device int that is atomically OR-ed with 1 if any row was cut
 *   at max_collisions or a query lies beyond a clamped grid, i.e. the neighbour relation may not be symmetric
 *   (consumers use it to choose the atomics-free symmetric backward). */
int spnb_compute_collisions(const float* qlocs, const float* locs, const float* low,
                            const float* grid_dims, const float* cellIDs, float* cellStarts,
                            float* cellEnds, float* collisions, int batch_size, int M, int N,
                            int ndims, int max_collisions, int ncells, float cellEdge, float radius,
                            int include_self, int* trunc_flag, void* stream);

/* The same lists in two forms from one kernel, for the particles as their own queries (qlocs == locs,
 * ndims <= 3, max_collisions <= 512): the float rows of spnb_compute_collisions() -- bit-identical -- and the
 * compact "tile lists" the ConvSP group kernels consume (layout in
 * smoothparticlenets_b200/csrc/tile_lists.cuh): per 64 consecutive queries the ranges of the sorted order that
 * hold their neighbours (a "tile", staged into shared memory with TMA bulk copies) and 16-bit lists of tile
 * slots, stored by list-length rank.  pos4: [batch_size, N, 4] float4 plane of the SORTED positions (x, y, z, 0),
 * as written by spnb_reorder_data_pos4(); the kernel stages its candidates from it.  tile_lists: device buffer of
 * spnb_tile_lists_bytes() bytes (0: sizes not supported).  The first int of the buffer is a device-side flag:
 * non-zero = unusable for this call (bit 0: the neighbour relation may not be symmetric -- a list is full / was
 * cut at max_collisions, or a query lies beyond a clamped grid; bit 1: a tile exceeds the staging capacity);
 * consumers test it on the device and fall back to the float lists.  sym_flag (optional) receives bit 0 too. */
size_t spnb_tile_lists_bytes(int batch_size, int N, int ndims, int max_collisions);
int spnb_compute_collisions_tiled(const float* pos4, const float* locs, const float* low, const float* grid_dims,
                                  const float* cellIDs, float* cellStarts, float* cellEnds, float* collisions,
                                  int batch_size, int N, int ndims, int max_collisions, int ncells,
                                  float cellEdge, float radius, int include_self, int* sym_flag,
                                  void* tile_lists, size_t tile_lists_bytes, void* stream);

/* Row permutation.  Replaces cuda_reorder_data (gpu_kernels.h:97-111).
 *   reverse == 0: nlocs[b,i] = locs[b,idxs[b,i]];  reverse != 0: nlocs[b,idxs[b,i]] = locs[b,i];
 *   same for data -> ndata when data != NULL (nchannels columns). */
int spnb_reorder_data(const float* locs, const float* data, const float* idxs, float* nlocs,
                      float* ndata, int batch_size, int N, int ndims, int nchannels, int reverse,
                      void* stream);
/* reverse == 0 variant that also writes the reordered positions as a float4 plane pos4[b,i] = (x, y, z, 0)
 * (ndims <= 4): the 16-byte-aligned layout TMA bulk copies and 128-bit shared-memory gathers want. */
int spnb_reorder_data_pos4(const float* locs, const float* data, const float* idxs, float* nlocs,
                           float* ndata, float* pos4, int batch_size, int N, int ndims, int nchannels,
                           void* stream);

/* ---- ConvSP ----------------------------------------------------------------------------------- */

/* Forward.  Replaces cuda_convsp with NULL gradients (gpu_kernels.h:7-33) plus the Python bias add
 * (convsp.py:172):  out[b,m,o] = bias[o] + sum over neighbours/kernel cells (bias may be NULL). */
int spnb_convsp_forward(const float* qlocs, const float* locs, const float* data,
                        const float* neighbors, const float* weight, const float* bias,
                        int batch_size, int M, int N, int nchannels, int ndims, int max_neighbors,
                        int nkernels, int ncells, float radius, const float* kernel_size,
                        const float* dilation, int dis_norm, int kernel_fn, float* out, void* stream);

/* Forward for wide channel counts (nchannels >= 32, nkernels <= 256, ndims <= 3; BASELINE.json config
 * 3): same result as spnb_convsp_forward, computed in the factored form
 *   G[q,cell,c] = sum_j W*norm*data[j,c],  out[q,o] = bias[o] + sum_{cell,c} weight[o,c,cell]*G[q,cell,c]
 * with the weights transposed into `workspace` (spnb_convsp_forward_wide_workspace_bytes() bytes, 0 =
 * shape not supported: use spnb_convsp_forward). */
size_t spnb_convsp_forward_wide_workspace_bytes(int nkernels, int nchannels, int ndims, int ncells);
int spnb_convsp_forward_wide(const float* qlocs, const float* locs, const float* data,
                             const float* neighbors, const float* weight, const float* bias,
                             int batch_size, int M, int N, int nchannels, int ndims, int max_neighbors,
                             int nkernels, int ncells, float radius, const float* kernel_size,
                             const float* dilation, int dis_norm, int kernel_fn, float* out,
                             void* workspace, size_t workspace_bytes, void* stream);

/* Backward.  Replaces cuda_convsp with non-NULL gradients.  Any of dqlocs/dlocs/ddata/dweight may be
 * NULL (not computed).  sym_flag: optional device int; when non-NULL, qlocs == locs, and *sym_flag
 * == 0 at execution time, the neighbour relation is taken to be symmetric and dlocs/ddata are
 * computed by gathers (no atomics); otherwise the scatter path with float atomics runs.
 *   workspace: device scratch of spnb_convsp_backward_workspace_bytes() bytes (may be NULL when that
 *   is 0). */
size_t spnb_convsp_backward_workspace_bytes(int nkernels, int nchannels, int ncells);
int spnb_convsp_backward(const float* qlocs, const float* locs, const float* data,
                         const float* neighbors, const float* weight, int batch_size, int M, int N,
                         int nchannels, int ndims, int max_neighbors, int nkernels, int ncells,
                         float radius, const float* kernel_size, const float* dilation, int dis_norm,
                         int kernel_fn, const float* grad_out, float* dqlocs, float* dlocs,
                         float* ddata, float* dweight, const int* sym_flag, void* workspace,
                         void* stream);

/* Backward for a BLOCK of the particles as the queries (slab decomposition): the queries are the
 * particles query_offset .. query_offset + M - 1 of `locs` [B, N, D] (own block inside [left halo | own | right
 * halo]), `neighbors` [B, M, K] their lists (indices into the N particles), grad_out_all [B, N, nkernels] the
 * output gradients of ALL N particles (the halo rows' come from their owners).  With a symmetric neighbour relation
 * (*sym_flag == 0 on the device, required) every gradient of a block particle is a gather over its own list:
 * dlocs_block [B, M, D] = d/dqlocs + d/dlocs of the block's particles, ddata_block [B, M, nchannels]; nothing is
 * scattered, no partial sums travel back to other ranks.  One scene per call (batch_size 1), kernel_size 1 shapes of
 * the fast path only (returns 0 otherwise).  No reference counterpart: the reference has no multi-GPU path. */
int spnb_convsp_backward_block(const float* locs, const float* data, const float* neighbors, const float* weight,
                               int batch_size, int M, int N, int nchannels, int ndims, int max_neighbors,
                               int nkernels, int ncells, float radius, int dis_norm, int kernel_fn,
                               const float* grad_out_all, int query_offset, const int* sym_flag,
                               float* dlocs_block, float* ddata_block, void* stream);

/* Backward for wide channel counts (nchannels in {32, 64}, nkernels <= 64, ndims <= 3; BASELINE.json config 3):
 * same gradients as spnb_convsp_backward, computed in the factored form
 *   dG[q,cell,c] = sum_o go[q,o]*weight[o,c,cell]        dweight[o,c,cell] = sum_q go[q,o]*G[q,cell,c]
 * (two tensor-core contractions) plus one walk over the neighbour lists for ddata / dlocs / dqlocs.  Any of the
 * four outputs may be NULL; dqlocs == dlocs (one buffer for both roles) requires M == N.  `workspace`:
 * spnb_convsp_backward_wide_workspace_bytes() bytes (0 = shape not supported: use spnb_convsp_backward). */
size_t spnb_convsp_backward_wide_workspace_bytes(int nkernels, int nchannels, int ndims, int ncells);
int spnb_convsp_backward_wide(const float* qlocs, const float* locs, const float* data,
                              const float* neighbors, const float* weight, int batch_size, int M, int N,
                              int nchannels, int ndims, int max_neighbors, int nkernels, int ncells,
                              float radius, const float* kernel_size, const float* dilation, int dis_norm,
                              int kernel_fn, const float* grad_out, float* dqlocs, float* dlocs, float* ddata,
                              float* dweight, void* workspace, size_t workspace_bytes, void* stream);

/* ---- fused ConvSP group (no reference counterpart) ------------------------ */

/* Several ConvSP layers with kernel_size 1 that share (locs, neighbors, radius) and use the particles
 * as their own queries (qlocs == locs), evaluated in one walk over the neighbour lists.  Per layer the
 * result is identical to spnb_convsp_forward / spnb_convsp_backward (within fp32 rounding).  Only a
 * fixed set of channel layouts is compiled in; for any other layout the functions return 0 /
 * workspace size 0 ("unsupported group signature") and the caller uses the per-layer entry points. */
typedef struct SpnbGroupLayer {
    const float* data;     /* [B,N,nchannels]; may be the locs pointer itself */
    const float* weight;   /* [nkernels,nchannels,1] */
    const float* bias;     /* [nkernels] or NULL (forward) */
    float* out;            /* forward: [B,N,nkernels] */
    const float* grad_out; /* backward: [B,N,nkernels] */
    float* ddata;          /* backward: [B,N,nchannels] or NULL */
    int nchannels, nkernels, kernel_fn, dis_norm;
} SpnbGroupLayer;

/* tile_lists: NULL, or the buffer spnb_build_tile_lists() filled for these neighbors (same
 * batch_size, N, max_neighbors): the group then runs from the compact lists when their flag allows. */
size_t spnb_convsp_group_workspace_bytes(const float* locs, int batch_size, int N, int ndims, float radius,
                                         int nlayers, const SpnbGroupLayer* layers, int backward);
int spnb_convsp_group_forward(const float* locs, const float* neighbors, int batch_size, int N, int ndims,
                              int max_neighbors, float radius, int nlayers, const SpnbGroupLayer* layers,
                              void* workspace, size_t workspace_bytes, const void* tile_lists, void* stream);
/* dlocs [B,N,ndims] receives d/dlocs through the geometry of ALL layers (query + neighbour roles);
 * gradients that reach locs through a data tensor aliasing it are in that layer's ddata. */
int spnb_convsp_group_backward(const float* locs, const float* neighbors, int batch_size, int N, int ndims,
                               int max_neighbors, float radius, int nlayers, const SpnbGroupLayer* layers,
                               float* dlocs, const int* sym_flag, void* workspace, size_t workspace_bytes,
                               const void* tile_lists, void* stream);

/* ---- fused elementwise stages of the PBF solver iteration (no reference counterpart; 8(f) rank 1) -- */

/* The per-particle arithmetic examples/fluid_sim.py:367-397 performs with torch ops between the ConvSP
 * layers of one solver iteration, one launch per stage (csrc/fluid_glue.cu has the formulas).  Vectors
 * are [BN, ndims], scalars [BN]; gradient INPUTS may be NULL (= zero). */
int spnb_pbf_stage1_forward(const float* x, const float* density, const float* nj, const float* ni_s, float* p,
                            float* xp, float* nij, long long BN, int ndims, float stiffness, float rho0,
                            void* stream);
int spnb_pbf_stage1_backward(const float* x, const float* density, const float* ni_s, const float* g_p,
                             const float* g_xp, const float* g_nij, float* g_x, float* g_density, float* g_nj,
                             float* g_ni_s, long long BN, int ndims, float stiffness, float rho0, void* stream);
int spnb_pbf_stage2_forward(const float* x, const float* p, const float* nij, const float* njp,
                            const float* nip_s, const float* nj_c, const float* ni_cs, float* d0, float* nrm,
                            long long BN, int ndims, float cohesion, float radius, float surface_tension,
                            float rho0, float constraint_scale, void* stream);
int spnb_pbf_stage2_backward(const float* x, const float* p, const float* nij, const float* nip_s,
                             const float* ni_cs, const float* g_d0, const float* g_nrm, float* g_x, float* g_p,
                             float* g_nij, float* g_njp, float* g_nip_s, float* g_nj_c, float* g_ni_cs,
                             long long BN, int ndims, float cohesion, float radius, float surface_tension,
                             float rho0, float constraint_scale, void* stream);
int spnb_pbf_stage3_forward(const float* x, const float* d0, const float* cd, const float* nrm,
                            const float* ncount, float* xnew, long long BN, int ndims, float relaxation,
                            float damp, void* stream);
int spnb_pbf_stage3_backward(const float* d0, const float* cd, const float* nrm, const float* ncount,
                             const float* g, float* g_d0, float* g_nrm, float* g_ncount, long long BN,
                             int ndims, float relaxation, float damp, void* stream);

/* out = inputs[0] + ... + inputs[n-1] (1 <= n <= 8; inputs_host: HOST array of n device pointers): the
 * backward of a fan-out, one pass instead of autograd's n-1 pairwise adds. */
int spnb_sum_n(const float* const* inputs_host, int n, float* out, long long nfloats, void* stream);

/* The ends of the step (fluid_sim.py:355-365, 412-424): gravity + velocity cap + position update
 * (gravity: HOST array of ndims floats), velocity from the position change ((a - b) / dt; backward != 0:
 * o = a / dt and, when o2 != NULL, o2 = -(a / dt)), and the XSPH viscosity update w0 + c*(vj - w0*vi_s). */
int spnb_pbf_integrate_forward(const float* x, const float* v, float* v2, float* x1, long long BN, int ndims,
                               const float* gravity_host, float dt, float cap, void* stream);
int spnb_pbf_integrate_backward(const float* v, const float* g_v2, const float* g_x1, float* g_v, long long BN,
                                int ndims, const float* gravity_host, float dt, float cap, void* stream);
int spnb_pbf_velocity(const float* a, const float* b, float* o, float* o2, long long n_floats, float dt,
                      int backward, void* stream);
int spnb_pbf_viscosity_forward(const float* w0, const float* vj, const float* vi_s, float* w1, long long BN,
                               int ndims, float c, void* stream);
int spnb_pbf_viscosity_backward(const float* w0, const float* vi_s, const float* g, float* g_w0, float* g_vj,
                                float* g_vi_s, long long BN, int ndims, float c, void* stream);

/* ---- ConvSDF ---------------------------------------------------------------------------------- */

/* Forward (bias added in the kernel, as common_funcs.h:834-835).  Replaces cuda_convsdf with NULL
 * gradients (gpu_kernels.h:35-59).  sdfs_len = number of floats in the SDF atlas (bounds guard). */
int spnb_convsdf_forward(const float* locs, int batch_size, int N, int ndims, const float* idxs,
                         const float* poses, const float* scales, int M, int pose_len,
                         const float* sdfs, size_t sdfs_len, const float* sdf_offsets,
                         const float* sdf_shapes, int nsdfs, const float* weight, const float* bias,
                         int nkernels, int ncells, const float* kernel_size, const float* dilation,
                         float max_distance, float* out, void* stream);

/* Backward.  dlocs [B,N,D], dweight [O,ncells], dposes [B,M,pose_len] (translation columns only,
 * rotation columns are left 0: the reference overwrites them with finite differences in Python,
 * convsdf.py:211-224).  Any of the three may be NULL. */
int spnb_convsdf_backward(const float* locs, int batch_size, int N, int ndims, const float* idxs,
                          const float* poses, const float* scales, int M, int pose_len,
                          const float* sdfs, size_t sdfs_len, const float* sdf_offsets,
                          const float* sdf_shapes, int nsdfs, const float* weight, int nkernels,
                          int ncells, const float* kernel_size, const float* dilation,
                          float max_distance, const float* grad_out, float* dlocs, float* dweight,
                          float* dposes, void* stream);

/* Same as spnb_convsdf_backward, and additionally fills the rotation columns of dposes with the
 * analytic derivative of the forward formula (rotate_point, common_funcs.h:203-235) with respect to
 * the 2-D angle / the four quaternion components taken as independent variables.  The reference
 * has no counterpart: it estimates these columns with forward differences (eps = 1e-3,
 * convsdf.py:211-224), which costs M * R extra forward passes.  Opt-in through
 * ConvSDF(compute_pose_grads="analytic"). */
int spnb_convsdf_backward_analytic(const float* locs, int batch_size, int N, int ndims,
                                   const float* idxs, const float* poses, const float* scales, int M,
                                   int pose_len, const float* sdfs, size_t sdfs_len,
                                   const float* sdf_offsets, const float* sdf_shapes, int nsdfs,
                                   const float* weight, int nkernels, int ncells,
                                   const float* kernel_size, const float* dilation,
                                   float max_distance, const float* grad_out, float* dlocs,
                                   float* dweight, float* dposes, void* stream);

/* ---- ParticleProjection / ImageProjection (3-D particles in camera space) -------------------------------- */

/* Gaussian splat of every particle into out [batch_size, height, width] (zero-filled here).  Replaces
 * cuda_particleprojection with dlocs == NULL (gpu_kernels.h:111-123; compute_particle_projection,
 * common_funcs.h:979-1052).  depth_mask [batch_size, height, width]: a particle behind a positive mask value
 * does not contribute to that pixel. */
int spnb_particleprojection_forward(const float* locs, int batch_size, int N, float camera_fl, int width,
                                    int height, float filter_std, float filter_scale, const float* depth_mask,
                                    float* out, void* stream);
/* dlocs [batch_size, N, 3] = d(loss)/d(locs) given grad_out = d(loss)/d(out); written, not accumulated.
 * Replaces cuda_particleprojection with dlocs != NULL (where `out` carries grad_out). */
int spnb_particleprojection_backward(const float* locs, int batch_size, int N, float camera_fl, int width,
                                     int height, float filter_std, float filter_scale, const float* depth_mask,
                                     const float* grad_out, float* dlocs, void* stream);

/* Bilinear sample of image [batch_size, channels, height, width] at each particle's pixel position into out
 * [batch_size, N, channels] (0 for particles behind the camera, outside the image or behind the depth mask).
 * Replaces cuda_imageprojection with NULL gradients (gpu_kernels.h:125-138; compute_image_projection,
 * common_funcs.h:1087-1183). */
int spnb_imageprojection_forward(const float* locs, const float* image, int batch_size, int N, float camera_fl,
                                 int width, int height, int channels, const float* depth_mask, float* out,
                                 void* stream);
/* dlocs [batch_size, N, 3] written, dimage [batch_size, channels, height, width] zero-filled here and
 * accumulated; either may be NULL. */
int spnb_imageprojection_backward(const float* locs, const float* image, int batch_size, int N, float camera_fl,
                                  int width, int height, int channels, const float* depth_mask,
                                  const float* grad_out, float* dlocs, float* dimage, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPNB_H_ */
