/*
 * mq3d.h -- C ABI of the B200-native volumetric reconstruction hot path.
 *
 * This is the drop-in boundary for the Open3D calls that the reference pipeline
 * (lszmer/metaquest-3d-reconstruction, pure Python) makes on its hot path.  Each entry point
 * names the reference call site it replaces (paths relative to the reference's scripts/).
 * Conventions:
 *   - every function returns an int status (MQ3D_OK == 0); mq3d_last_error() gives the message of
 *     the last failure on the calling thread;
 *   - pointers named *_dev are DEVICE pointers valid on the grid's CUDA device; everything else
 *     (camera matrices, counts) is host memory.  No torch/Open3D types appear here;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Kernels are enqueued on
 *     it; functions documented as "synchronises" wait for that stream before returning;
 *   - one handle <-> one CUDA device; handles are not thread-safe (the reference caller is a single
 *     frame-sequential Python thread, processing/reconstruction/utils/o3d_utils.py:231-236);
 *   - there is no CPU fallback: without a CUDA device every call fails with MQ3D_ERR_CUDA.
 *
 * Camera matrices follow the reference call sites: intrinsic K is row-major double[9], extrinsic E
 * is row-major double[16] world->camera (o3d_utils.py:203-210), depth_scale is a float (always 1.0
 * in the reference), block_resolution is 16.
 */
#ifndef MQ3D_H
#define MQ3D_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MQ3D_OK 0
#define MQ3D_ERR_CUDA 1             /* CUDA runtime failure (message has the cudaError string) */
#define MQ3D_ERR_INVALID 2          /* bad argument (shape, null pointer, unsupported resolution) */
#define MQ3D_ERR_NO_BLOCK_TOUCHED 3 /* Open3D: "No block is touched in TSDF volume ..." */
#define MQ3D_ERR_STATE 4            /* call order violated (e.g. fill before count) */

#define MQ3D_ATTR_TSDF_WEIGHT 1 /* 'tsdf','weight' float32 x1 (o3d_utils.py:171-179) */
#define MQ3D_ATTR_COLOR 2       /* + 'color' float32 x3 (Open3D colour Integrate overload) */

typedef struct mq3d_grid mq3d_grid;   /* VoxelBlockGrid state: spatial hash + block pool */
typedef struct mq3d_scene mq3d_scene; /* RaycastingScene state: triangles + LBVH */

const char *mq3d_last_error(void);
int mq3d_version(void);
/* Self-test used by the parity suite: number of floats with bit pattern in [lo_bits, hi_bits] (both
 * signs) for which the kernels' branch-free reciprocal differs from the IEEE round-to-nearest
 * reciprocal.  Must be 0 over the normal range [0x00800000, 0x7E800000]. */
int mq3d_selftest_rcp(unsigned lo_bits, unsigned hi_bits, unsigned long long *n_bad_out);

/* ---- voxel block grid life cycle ------------------------------------------------------------
 * Replaces o3d.t.geometry.VoxelBlockGrid(attr_names, attr_dtypes, attr_channels, voxel_size,
 * block_resolution, block_count, device)            (o3d_utils.py:171-179).
 * block_count is an initial capacity; the pool and hash grow x2 when exceeded (Open3D rehash). */
int mq3d_grid_create(float voxel_size, int block_resolution, int64_t block_count, int attr_mask,
                     int device, mq3d_grid **out);
int mq3d_grid_destroy(mq3d_grid *g);
int mq3d_grid_reset(mq3d_grid *g, void *stream);           /* drop all blocks, keep capacity */
int mq3d_grid_reserve(mq3d_grid *g, int64_t block_count, void *stream);
int mq3d_grid_num_blocks(mq3d_grid *g, int64_t *n, void *stream); /* synchronises */
int mq3d_grid_info(mq3d_grid *g, float *voxel_size, int *resolution, int64_t *capacity,
                   int *attr_mask, int *device);
/* Raw views of the block pool (valid until the next call that may grow the pool):
 * keys int32 [capacity][3]; tsdf/weight float32 [capacity][16][16][16] (z,y,x); color float32
 * [capacity][16][16][16][3] or NULL.  First num_blocks entries are active.  This is the
 * VoxelBlockGrid.save payload (dataio/reconstruction_data_io.py:51-55). */
int mq3d_grid_pool(mq3d_grid *g, int32_t **keys_dev, float **tsdf_dev, float **weight_dev,
                   float **color_dev);
/* Copy the active blocks (device to device) into caller buffers sized by mq3d_grid_num_blocks:
 * keys int32 [n][3], tsdf/weight float32 [n][4096], color float32 [n][4096][3] (NULL to skip). */
int mq3d_grid_export(mq3d_grid *g, int32_t *keys_dev, float *tsdf_dev, float *weight_dev,
                     float *color_dev, void *stream);
/* VoxelBlockGrid.load (reconstruction_data_io.py:42-48): insert n blocks with their values. */
int mq3d_grid_import(mq3d_grid *g, const int32_t *keys_dev, const float *tsdf_dev,
                     const float *weight_dev, const float *color_dev, int64_t n, void *stream);
/* Multi-GPU partition: this grid keeps only blocks whose super-tile
 * (tile_blocks^3 blocks) hashes to `rank` of `world`, plus the one-block ghost shell around them.
 * world == 1 disables filtering.  Must be called on an empty grid. */
int mq3d_grid_set_partition(mq3d_grid *g, int rank, int world, int tile_blocks);
/* integrate_ghosts = 1 (default): frames are integrated into owned + ghost blocks, no exchange is ever
 * needed (the north-star scheme).  integrate_ghosts = 0: only owned blocks are integrated (no redundant
 * work) and the ghost shell is fetched once from the owners before extraction: the owner calls
 * mq3d_grid_ghost_select(dest) (first with keys_dev == NULL for the count, then with buffers:
 * keys int32 [n][3], tsdf/weight float32 [n][4096], color float32 [n][4096][3] or NULL), the host moves
 * the payload (NCCL send/recv) and the receiver calls mq3d_grid_import.  Values are the owner's, i.e.
 * bit-identical to what redundant integration would have produced. */
int mq3d_grid_set_ghost_mode(mq3d_grid *g, int integrate_ghosts);
int mq3d_grid_ghost_select(mq3d_grid *g, int dest_rank, int64_t *n_out, int32_t *keys_dev, float *tsdf_dev,
                           float *weight_dev, float *color_dev, void *stream);
/* Ghost block counts for every destination rank in one pass (counts_out: host int64[world], entry of the
 * own rank is 0).  Passing a count obtained here as *n_out to mq3d_grid_ghost_select (with buffers) makes
 * that call fully asynchronous: no host synchronisation, payload valid in stream order. */
int mq3d_grid_ghost_counts(mq3d_grid *g, int64_t *counts_out, void *stream);
/* Ghost shell by peer memory (CUDA IPC over NVLink / NVSwitch), the alternative to select + send/recv +
 * import: every rank publishes a MQ3D_PEER_DESC_BYTES descriptor of its pool (IPC handles, block count),
 * the host all-gathers the descriptors (stream-ordered after integration, so the collective is also the
 * "all ranks finished integrating" barrier) and calls mq3d_grid_ghost_pull with the [world] array: one
 * kernel lists the peers' owned blocks inside this rank's shell straight from the peers' key arrays,
 * another copies tsdf | weight | colour from the owners' pools into the local pool -- all enqueued without a host
 * round trip (list length and activation stay on the device; the pool is reserved for the upper bound first).
 * n_pulled may be NULL; if given the call synchronises to read the number of blocks fetched.  The caller must
 * fence (any stream-ordered collective) before any rank changes its grid again.  Pools exported this
 * way are retired instead of freed on growth, until the grid is destroyed.  Grids living in the same
 * process (rank-by-rank emulation) are read through their plain device pointers. */
#define MQ3D_PEER_DESC_BYTES 512
int mq3d_grid_peer_descriptor(mq3d_grid *g, void *desc_out, void *stream);
int mq3d_grid_ghost_pull(mq3d_grid *g, const void *descs, int64_t *n_pulled, void *stream);

/* ---- K1: raw NDC depth -> linear metres + confidence mask -----------------------------------
 * Replaces DepthDataIO.load_depth_map's convert_depth_to_linear + is_depth_map_valid
 * (dataio/depth_data_io.py:33-53,80-85; utils/depth_utils.py:21-46) and the masking in
 * o3d_utils.load_depth_map (o3d_utils.py:131-142), for n_frames frames in one launch.
 * near/far: host double[n_frames].  conf_dev (float64) / count_dev (int32) may be NULL (no mask);
 * has_conf_dev (uint8[n_frames], may be NULL = all present) marks frames whose confidence map
 * exists (missing map => unfiltered depth, o3d_utils.py:137-139).
 * frame_valid_dev int32[n_frames] receives is_depth_map_valid per frame. */
int mq3d_depth_prepare(const float *raw_dev, int n_frames, int width, int height, const double *near_z,
                       const double *far_z, const double *conf_dev, const int32_t *count_dev,
                       const uint8_t *has_conf_dev, double conf_thr, int32_t count_thr,
                       float *out_dev, int32_t *frame_valid_dev, void *stream);

/* ---- K2: frustum block activation -----------------------------------------------------------
 * Replaces vbg.compute_unique_block_coordinates(depth, intrinsic, extrinsic, depth_scale,
 * depth_max, trunc_voxel_multiplier) (o3d_utils.py:212-219).  Does NOT allocate blocks (Open3D
 * uses a scratch frustum hashmap).  out_keys_dev: int32[(W/4)*(H/4)*4][3]; *out_n = number of
 * unique keys (order unspecified).  Synchronises.  Zero touched blocks =>
 * MQ3D_ERR_NO_BLOCK_TOUCHED (the reference's RuntimeError). */
int mq3d_touch(mq3d_grid *g, const float *depth_dev, int width, int height, const double K[9],
               const double E[16], float depth_scale, float depth_max, float trunc_voxel_multiplier,
               int32_t *out_keys_dev, int64_t *out_n, void *stream);

/* ---- K3: projective TSDF / weight (/ colour) update ----------------------------------------
 * Replaces vbg.integrate(block_coords, depth, intrinsic, extrinsic, depth_scale, depth_max,
 * trunc_voxel_multiplier) (o3d_utils.py:221-229) and Open3D's colour overload when
 * color_dev != NULL (uint8 [CH][CW][3], colour intrinsic Kc, identity colour extrinsic).
 * Activates the listed blocks (zero-initialised) then updates every voxel of every listed block. */
int mq3d_integrate(mq3d_grid *g, const int32_t *keys_dev, int64_t n_keys, const float *depth_dev,
                   int width, int height, const uint8_t *color_dev, int color_width, int color_height,
                   const double Kd[9], const double Kc[9], const double E[16], float depth_scale,
                   float depth_max, float trunc_voxel_multiplier, void *stream);

typedef struct {
    int64_t frames_integrated; /* frames with frame_valid != 0 */
    int64_t block_visits;      /* sum over frames of |unique touched blocks| (Open3D's |block_coords|) */
    int64_t blocks_loaded;     /* block residencies (one 16^3 tile load+store each) */
    int64_t num_blocks;        /* active blocks after the call */
    int64_t batches;
    int64_t voxel_updates;     /* voxel visits that passed every reject (updated-voxel count) */
    double touch_ms;           /* device time of the K2 launches (CUDA events on `stream`) */
    double integrate_ms;       /* device time of the K3 launches (CUDA events on `stream`) */
    int64_t slow_div_batches;  /* batches holding a depth in (0, 2^-75): integrated with the guarded division */
} mq3d_seq_stats;

/* Fused replacement of the whole per-frame loop of integrate() (o3d_utils.py:231-236):
 * for f in frames (in order): touch -> activate -> integrate.  Frames are processed in batches of
 * `batch_frames` (<= 256): each touched block is loaded once per batch and the frames that touched
 * it are applied in frame order, which is bit-identical to the frame-sequential loop.
 * depth_dev: float32 [n_frames][H][W] linear (K1 output); frame_valid_dev: int32[n_frames] or
 * NULL; color_dev: uint8 [n_frames][CH][CW][3] or NULL; Kd/Kc: host double[n_frames][9];
 * E: host double[n_frames][16].  All batches are enqueued without intermediate host synchronisation
 * (overflow of the pool / hash table is detected on the device; the call then grows the grid and resumes
 * from the failing batch); synchronises once before returning.  A valid frame whose (unpartitioned) touch
 * yields no block => MQ3D_ERR_NO_BLOCK_TOUCHED after the remaining frames were integrated. */
int mq3d_integrate_sequence(mq3d_grid *g, const float *depth_dev, const int32_t *frame_valid_dev,
                            int n_frames, int width, int height, const uint8_t *color_dev,
                            int color_width, int color_height, const double *Kd, const double *Kc,
                            const double *E, float depth_scale, float depth_max,
                            float trunc_voxel_multiplier, int batch_frames, mq3d_seq_stats *stats,
                            void *stream);
/* Batch gates for the NEXT mq3d_integrate_sequence / _rgbx call on this grid: events[i] (a cudaEvent_t, or NULL) is
 * waited for on the call's stream before batch i is enqueued, so one call can consume frames that another stream
 * is still producing (host->device upload + mq3d_depth_prepare of chunk i+1 while chunk i integrates) without a
 * host synchronisation per chunk -- the streaming form of the frame loop in integrate() (o3d_utils.py:231-236),
 * where load_depth_map of the next frame could overlap the integration of the current one.  The list is copied;
 * it is consumed (cleared) by that call.  n_events = 0 clears it. */
int mq3d_grid_set_batch_gates(mq3d_grid *g, const void *const *events, int n_events);
/* Colour on the depth pixel grid.  The colour branch of Open3D's Integrate reads, for a voxel that projects
 * to depth pixel (ui, vi), the colour pixel round(Project_colourK(Unproject_depthK(ui, vi, 1))) under an
 * identity extrinsic -- a function of the depth pixel alone.  mq3d_color_resample evaluates it once per
 * frame: rgbx_dev uint32 [n_frames][H][W] = R | G << 8 | B << 16, byte 3 = 0xFF where the projection leaves
 * the colour image.  color_src: uint8 [n_frames][CH][CW][3] in device memory OR in pinned (mapped) host
 * memory -- then only the W x H sampled pixels per frame cross PCIe instead of CW x CH.  Asynchronous on
 * `stream`.  mq3d_integrate_sequence does this internally per batch; mq3d_integrate_sequence_rgbx takes
 * the resampled frames instead (so that a copy stream can prepare batch k+1 while batch k integrates). */
int mq3d_color_resample(const uint8_t *color_src, int n_frames, int color_width, int color_height,
                        int width, int height, const double *Kd, const double *Kc, uint32_t *rgbx_dev,
                        int device, void *stream);
int mq3d_integrate_sequence_rgbx(mq3d_grid *g, const float *depth_dev, const int32_t *frame_valid_dev,
                                 int n_frames, int width, int height, const uint32_t *rgbx_dev,
                                 const double *Kd, const double *E, float depth_scale, float depth_max,
                                 float trunc_voxel_multiplier, int batch_frames, mq3d_seq_stats *stats,
                                 void *stream);

/* ---- K5: marching cubes / point cloud -------------------------------------------------------
 * Replaces vbg.extract_triangle_mesh(weight_threshold, estimated_vertex_number=-1)
 * (reconstruct_scene.py:105-108,186-189) and vbg.extract_point_cloud(weight_threshold=3.0)
 * (reconstruct_scene.py:90).  Two calls: *_count classifies and returns sizes (synchronises);
 * *_fill writes into caller-allocated device buffers sized by those counts.
 * vertices/normals float32 [V][3]; triangles int32 [T][3]; vertex_keys (optional, may be NULL)
 * int32 [V][4] = (voxel x,y,z, axis) of the lattice edge carrying the vertex.
 * Output order is deterministic: block order, then edge/cube order inside the block. */
int mq3d_extract_mesh_count(mq3d_grid *g, float weight_threshold, int64_t *n_vertices,
                            int64_t *n_triangles, void *stream);
int mq3d_extract_mesh_fill(mq3d_grid *g, float *vertices_dev, float *normals_dev,
                           int32_t *triangles_dev, int32_t *vertex_keys_dev, void *stream);
/* Single-call form of the pair above for callers that can bound the mesh size (e.g. from the previous extraction):
 * classification, scan and emission are enqueued back to back without a host round trip; the kernels read the
 * totals on the device and write nothing when the mesh exceeds cap_vertices / cap_triangles.  Synchronises once,
 * returns the true sizes; if they exceed the capacities the caller allocates exact buffers and calls
 * mq3d_extract_mesh_fill (and _colors) -- the classification is kept.  vertex_keys_dev / colors_dev may be NULL. */
int mq3d_extract_mesh(mq3d_grid *g, float weight_threshold, float *vertices_dev, float *normals_dev,
                      int32_t *triangles_dev, int32_t *vertex_keys_dev, float *colors_dev, int64_t cap_vertices,
                      int64_t cap_triangles, int64_t *n_vertices, int64_t *n_triangles, void *stream);
int mq3d_extract_points_count(mq3d_grid *g, float weight_threshold, int64_t *n_points, void *stream);
int mq3d_extract_points_fill(mq3d_grid *g, float *points_dev, float *normals_dev,
                             int32_t *point_keys_dev, void *stream);
/* Vertex / point colours of a grid with the colour attribute (north-star row A3c; Open3D's colour branch
 * of ExtractTriangleMesh / ExtractPointCloud, what mesh.vertex.colors / pcd.point.colors hold after
 * reconstruct_scene.py:90,105-108 on a coloured grid): float32 [V][3] in [0,1] =
 * ((1 - ratio) * c_owner + ratio * c_neighbour) / 255, same order as the matching *_fill.  Valid
 * between the matching *_count and the next change of the grid. */
int mq3d_extract_mesh_colors(mq3d_grid *g, float *colors_dev, void *stream);
int mq3d_extract_points_colors(mq3d_grid *g, float *colors_dev, void *stream);

/* ---- N1: mesh component filter on the device -------------------------------------------------
 * Replaces filter_mesh_components (processing/reconstruction/utils/o3d_utils.py:241-321), which the reference runs
 * on a legacy CPU mesh between marching cubes and the colour-view raycast (reconstruct_scene.py:115-118,192-195):
 * cluster_connected_triangles (triangles sharing an edge), keep clusters with >= min_triangle_count triangles (or
 * the largest one if none qualifies), remove_unreferenced_vertices (only if triangles were removed),
 * remove_degenerate_triangles, remove_duplicated_triangles (equal up to rotation, first kept),
 * remove_duplicated_vertices (identical coordinates, first kept -- this also welds the vertices that several ranks
 * of a multi-GPU run emit on ghost edges).  Order of the surviving vertices / triangles is preserved.
 * Inputs: vertices float32 [V][3], triangles int32 [T][3], optional per-vertex normals / colors float32 [V][3]
 * (NULL to skip; outputs must then be NULL too).  Outputs are caller-allocated with the INPUT sizes; the valid
 * prefix is out_n_vertices / out_n_triangles.  remove_non_manifold_edges: edges with more than two triangles are
 * counted in info.non_manifold_edges (a marching-cubes mesh has none); if non-zero the caller completes that step.
 * V, T > 0.  Synchronises. */
typedef struct {
    int64_t components;          /* connected components of the input */
    int64_t components_kept;
    int64_t largest_component;   /* triangles of the largest component */
    int64_t fallback_largest;    /* 1: no component reached min_triangle_count, the largest one was kept */
    int64_t input_triangles;
    int64_t removed_triangles;   /* triangles of the dropped components */
    int64_t non_manifold_edges;  /* edges of the result carried by more than two triangles */
} mq3d_mesh_filter_info;
int mq3d_mesh_filter(const float *vertices_dev, const float *normals_dev, const float *colors_dev, int64_t n_vertices,
                     const int32_t *triangles_dev, int64_t n_triangles, int64_t min_triangle_count,
                     float *out_vertices_dev, float *out_normals_dev, float *out_colors_dev,
                     int32_t *out_triangles_dev, int64_t *out_n_vertices, int64_t *out_n_triangles,
                     mq3d_mesh_filter_info *info, int device, void *stream);

/* ---- N4: odometry information matrix -----------------------------------------------------------
 * Replaces o3d.t.pipelines.odometry.compute_odometry_information_matrix(source_depth, target_depth, intrinsic,
 * source_to_target, dist_threshold, depth_scale, depth_max) as build_pose_graph_for_fragment calls it for every
 * consecutive frame pair and every overlapping key-frame pair
 * (processing/reconstruction/depth_optimization/make_fragments.py:142-150,228-233).  Depth images float32 [H][W]
 * on the device; K row-major double[9], source_to_target row-major double[16] (host).  info_out: host double[36],
 * row-major symmetric 6 x 6 (rotation block first).  Synchronises. */
int mq3d_odometry_information(const float *source_depth_dev, const float *target_depth_dev, int width, int height,
                              const double K[9], const double source_to_target[16], float dist_threshold,
                              float depth_scale, float depth_max, double info_out[36], int device, void *stream);

/* ---- K4: multi-view depth confidence --------------------------------------------------------
 * Replaces build_confidence_map over all reference frames of one side
 * (processing/reconstruction/confidence_estimation/estimate_depth_confidences.py:15-79 and
 * compute_pixel_error_map.py:120-220), float64 arithmetic on float32 inputs.
 * depths_dev: float32 [N][H][W] linear; frame_valid_dev: int32[N] or NULL; K: host float[N][9]
 * (cx already mirrored, o3d_utils.py:14-19); Ecw / Ecw_inv: host float[N][16] camera->world and its
 * float32 inverse.  Outputs conf_dev float64 [N][H][W], count_dev int32 [N][H][W]. */
int mq3d_confidence(const float *depths_dev, const int32_t *frame_valid_dev, int n_frames, int width,
                    int height, const float *K, const float *Ecw, const float *Ecw_inv,
                    int target_frame_range, double depth_max, double error_threshold,
                    double *conf_dev, int32_t *count_dev, void *stream);

/* ---- K6: colour-aligned depth raycast -------------------------------------------------------
 * Replaces o3d.t.geometry.RaycastingScene(): add_triangles(mesh), create_rays_pinhole(K, E,
 * width_px, height_px), cast_rays(rays)['t_hit'] (reconstruct_scene.py:197-198;
 * o3d_utils.py:324-342).  t_hit float32 [H][W], +inf on miss, in units of the (unnormalised) ray
 * direction, i.e. z-depth for pinhole rays. */
int mq3d_scene_create(int device, mq3d_scene **out);
int mq3d_scene_destroy(mq3d_scene *s);
int mq3d_scene_add_triangles(mq3d_scene *s, const float *vertices_dev, int64_t n_vertices,
                             const int32_t *triangles_dev, int64_t n_triangles, void *stream);
int mq3d_scene_create_rays_pinhole(const double K[9], const double E[16], int width, int height,
                                   float *rays_dev, void *stream);
int mq3d_scene_cast_rays(mq3d_scene *s, const float *rays_dev, int64_t n_rays, float *t_hit_dev,
                         void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MQ3D_H */
