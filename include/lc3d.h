/*
 * lc3d.h — C ABI of the B200-native fine-registration hot path.
 *
 * The reference (Eberty/LowCost3DReconstruction) has no FFI: its four hot-path
 * tools call PCL classes directly.  This header is the boundary inserted at those
 * PCL call sites.  Every entry point cites the reference line(s)
 * whose PCL call it replaces.  Plain pointers and sizes only; no C++/torch types.
 *
 * Threading: one lc3d_ctx = one device + one stream; a ctx is not thread-safe,
 * distinct ctxs are independent.  All calls are synchronous on return.
 * Ownership: the caller owns every host buffer; the library owns device memory.
 * Errors: 0 = ok, <0 = error (message via lc3d_last_error).  There is no CPU
 * fallback: without a usable CUDA device lc3d_create fails.
 */
#ifndef LC3D_H_
#define LC3D_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LC3D_OK 0
#define LC3D_ERR_INVALID (-1)
#define LC3D_ERR_CUDA (-2)
#define LC3D_ERR_NOMEM (-3)
#define LC3D_ERR_INTERNAL (-4)

/* A host point cloud described by base pointers + byte strides, so that both the
 * PCL in-memory layout (pcl::PointXYZRGBNormal, 48-byte AoS: xyz at +0, normal at
 * +16, rgba at +32, curvature at +36) and packed SoA arrays are accepted.
 * normal / rgba / curvature may be NULL when the operation does not need them.
 * Alignment: every stride must be a multiple of 4 bytes (the records are unpacked on the
 * device with 4-byte loads); a layout that violates this is rejected with LC3D_ERR_INVALID.
 * The base pointers themselves need no particular alignment. */
typedef struct lc3d_cloud {
  int64_t n;
  const float* xyz;
  int64_t xyz_stride; /* bytes between consecutive points */
  const float* normal;
  int64_t normal_stride;
  const uint32_t* rgba;
  int64_t rgba_stride;
  const float* curvature;
  int64_t curvature_stride;
} lc3d_cloud;

typedef struct lc3d_ctx lc3d_ctx;
/* A cloud staged on the device (SoA float4) — lets callers keep views resident in
 * HBM across calls (chain registration re-uses each view as target then source). */
typedef struct lc3d_dcloud lc3d_dcloud;

/* device: CUDA ordinal.  stream: a cudaStream_t to run on (e.g. torch's current
 * stream, so that the caller's CUDA events bracket our kernels) or NULL for a
 * private stream. */
int lc3d_create(int device, void* stream, lc3d_ctx** out);
void lc3d_destroy(lc3d_ctx* ctx);
/* ctx may be NULL: returns the last error of a failed lc3d_create on this thread. */
const char* lc3d_last_error(const lc3d_ctx* ctx);
const char* lc3d_version(void);
/* Diagnostics of the most recently built spatial index: out[0]=cell edge, out[1..3]=grid
 * dims, out[4]=cells, out[5]=indexed points, out[6]=bricks/occupied super-cells (or 0). */
void lc3d_debug_grid_info(const lc3d_ctx* ctx, double out[8]);
/* Number of kernels this ctx has launched since creation (bench.py "gpu_launches"). */
int64_t lc3d_launch_count(const lc3d_ctx* ctx);
/* Device allocations (cudaMalloc of a scratch / cloud buffer) made so far by this process: a
 * diagnostic for "the steady state allocates nothing" (cudaFree synchronises the whole device). */
int64_t lc3d_debug_alloc_count(void);

/* Page-locks / releases a caller-owned host range (cudaHostRegister): copies from / to it are
 * then true asynchronous DMA at full PCIe rate instead of staged pageable copies.  Worth it for
 * buffers that cross PCIe more than once (pinning costs about as much as one pageable copy).
 * Optional: every entry point also accepts pageable memory. */
int lc3d_host_register(void* ptr, uint64_t bytes);
int lc3d_host_unregister(void* ptr);

int lc3d_cloud_upload(lc3d_ctx* ctx, const lc3d_cloud* host, lc3d_dcloud** out);
void lc3d_cloud_free(lc3d_ctx* ctx, lc3d_dcloud* dc);
int64_t lc3d_dcloud_size(const lc3d_dcloud* dc);

/* ------------------------------------------------------------------ ICP ---- */

#define LC3D_ICP_POINT_TO_POINT 0 /* pcl::IterativeClosestPoint + TransformationEstimationSVD */
#define LC3D_ICP_POINT_TO_PLANE 1 /* pcl::IterativeClosestPointWithNormals (PointToPlaneLLS)  */

/* pcl::registration::DefaultConvergenceCriteria::ConvergenceState */
#define LC3D_STATE_NOT_CONVERGED 0
#define LC3D_STATE_ITERATIONS 1
#define LC3D_STATE_TRANSFORM 2
#define LC3D_STATE_ABS_MSE 3
#define LC3D_STATE_REL_MSE 4
#define LC3D_STATE_NO_CORRESPONDENCES 5

/* Mirrors the setters at pcl_tools/fine_registration.cpp:112-118. */
typedef struct lc3d_icp_params {
  double max_correspondence_distance; /* :112 setMaxCorrespondenceDistance  (CLI default 0.1)  */
  double transformation_epsilon;      /* :116 setTransformationEpsilon      (CLI default 1e-9) */
  double euclidean_fitness_epsilon;   /* :118 setEuclideanFitnessEpsilon    (CLI default 1e-3) */
  int32_t max_iterations;             /* :114 setMaximumIterations          (CLI default 50)   */
  int32_t mode;                       /* LC3D_ICP_POINT_TO_POINT | LC3D_ICP_POINT_TO_PLANE     */
  int32_t compute_fitness;            /* :126 getFitnessScore() (one extra unbounded NN pass)  */
  int32_t dump_iteration;             /* parity hook: iteration (0-based) whose correspondences
                                         are copied to corr_index/corr_dist2; -1 = none         */
} lc3d_icp_params;

typedef struct lc3d_icp_result {
  float transformation[16]; /* :124 getFinalTransformation(), row-major 4x4 */
  double fitness;           /* :126 getFitnessScore(): mean squared NN distance, all source pts */
  double last_mse;          /* mean squared correspondence distance of the last iteration      */
  int64_t last_correspondences;
  int32_t converged;        /* :125 hasConverged() */
  int32_t iterations;       /* nr_iterations_ */
  int32_t state;            /* LC3D_STATE_* */
  int32_t reserved;
  /* device-side timings, CUDA events on the ctx stream, milliseconds */
  float ms_upload;
  float ms_index;   /* spatial index build over the target */
  float ms_loop;    /* the ICP loop (all iterations) */
  float ms_fitness; /* getFitnessScore pass */
  float ms_download;
  float ms_total;
} lc3d_icp_result;

/* Optional outputs of an alignment (any pointer may be NULL).
 * registered_xyz / registered_normal: n_src x 3 packed floats — the `registered`
 * cloud of fine_registration.cpp:121 (source transformed by the final matrix,
 * normals rotated by its 3x3 block).  corr_index / corr_dist2: n_src entries, the
 * correspondences of iteration params.dump_iteration in source order: index of
 * the matched target point or -1 if rejected, and its squared distance. */
typedef struct lc3d_icp_outputs {
  float* registered_xyz;
  float* registered_normal;
  int32_t* corr_index;
  float* corr_dist2;
} lc3d_icp_outputs;

/* Replaces icp.setInputSource/Target + align + getFinalTransformation +
 * hasConverged + getFitnessScore (pcl_tools/fine_registration.cpp:105-126).
 * Host buffers in, host buffers out; the whole loop runs on the device. */
int lc3d_icp_align(lc3d_ctx* ctx, const lc3d_cloud* source, const lc3d_cloud* target,
                   const lc3d_icp_params* params, lc3d_icp_result* result,
                   const lc3d_icp_outputs* outputs);

/* Same, on clouds already resident in HBM.  Only the result record crosses PCIe. */
int lc3d_icp_align_resident(lc3d_ctx* ctx, const lc3d_dcloud* source, const lc3d_dcloud* target,
                            const lc3d_icp_params* params, lc3d_icp_result* result,
                            const lc3d_icp_outputs* outputs);

/* ------------------------------------- one pair over several GPUs -- */

/* Second multi-GPU mode (the first is independent pairs, one per GPU): ONE large pair with the
 * SOURCE sharded over `world` processes, one GPU each (at most 8), every rank holding the whole
 * target.  Each rank's correspondence kernel writes its partial estimator sums into an exchange
 * buffer the other ranks have mapped with CUDA IPC; the solve kernel of every rank reads all
 * ranks' sums directly from peer memory over NVLink / NVSwitch, adds them in rank order and
 * solves, so every rank obtains the bit-identical pose and convergence decision without a
 * separate all-reduce or broadcast.  Correspondences are exactly those of the unsharded run;
 * the pose differs from it only by the grouping of the fp64 sums.
 *   1. every rank: lc3d_shard_export (allocates the exchange buffer, returns its 64-byte IPC handle)
 *   2. the caller exchanges the handles (any transport: 64 bytes per rank)
 *   3. every rank: lc3d_shard_connect with all handles
 *   4. every rank, collectively, once per alignment: lc3d_icp_align_sharded.  The callers must
 *      synchronise (a barrier) between two consecutive sharded alignments.
 * fitness_sum_count (may be NULL) receives this shard's {sum of squared NN distances, points}: the
 * fitness of the whole pair is sum(sums) / sum(counts) over the ranks. */
int lc3d_shard_export(lc3d_ctx* ctx, int64_t max_shard_points, unsigned char handle_out[64]);
int lc3d_shard_connect(lc3d_ctx* ctx, int32_t rank, int32_t world, const unsigned char* handles);
int lc3d_icp_align_sharded(lc3d_ctx* ctx, const lc3d_dcloud* source_shard, const lc3d_dcloud* target,
                           const lc3d_icp_params* params, lc3d_icp_result* result,
                           const lc3d_icp_outputs* outputs, double fitness_sum_count[2]);
void lc3d_shard_close(lc3d_ctx* ctx);

/* ------------------------------------------------------ neighbour search ---- */

/* Exact k nearest neighbours of every query among `cloud` (pcl::search::KdTree
 * nearestKSearch, normal_estimation.cpp:89-96; implicit in outlier_removal.cpp:80-84).
 * out_index/out_dist2: n_query x k, ascending distance; ties broken by lower index.
 * queries == NULL means the cloud queries itself (self is then neighbour 0). */
int lc3d_knn(lc3d_ctx* ctx, const lc3d_cloud* cloud, const lc3d_cloud* queries, int32_t k,
             int32_t* out_index, float* out_dist2);

/* Exact 1-NN with a max squared-distance gate (max_dist <= 0 or +inf: unbounded).
 * out_index -1 where no neighbour passes the gate. */
int lc3d_nn(lc3d_ctx* ctx, const lc3d_cloud* cloud, const lc3d_cloud* queries, double max_dist,
            int32_t* out_index, float* out_dist2);

/* ------------------------------------------------------- normal estimation -- */

/* Replaces pcl::NormalEstimation setKSearch + setViewPoint + compute
 * (pcl_tools/normal_estimation.cpp:84-108).  out_normal: n x 3, out_curvature: n.
 * The tool-level global flip (normal_estimation.cpp:112-118) stays in the caller. */
int lc3d_normals(lc3d_ctx* ctx, const lc3d_cloud* cloud, int32_t k, const float viewpoint[3],
                 float* out_normal, float* out_curvature);

/* pcl::compute3DCentroid (normal_estimation.cpp:101): float32 mean of xyz. */
int lc3d_centroid(lc3d_ctx* ctx, const lc3d_cloud* cloud, float out_centroid[4]);

/* ------------------------------------------------------------- VoxelGrid ---- */

/* Replaces pcl::VoxelGrid setLeafSize + filter (pcl_tools/cloud_downsampling.cpp:73-76).
 * Outputs are caller-allocated for the worst case (n points): out_xyz n x 3,
 * out_normal n x 3 (or NULL), out_rgba n (or NULL), out_curvature n (or NULL),
 * out_voxel_of_point n (or NULL; voxel rank in output order of each input point).
 * *out_count receives the number of output points.  When PCL's index-overflow
 * guard trips (dx*dy*dz > INT32_MAX) the output is the unfiltered input, as in PCL. */
int lc3d_voxel_grid(lc3d_ctx* ctx, const lc3d_cloud* cloud, const float leaf[3], float* out_xyz,
                    float* out_normal, uint32_t* out_rgba, float* out_curvature,
                    int32_t* out_voxel_of_point, int64_t* out_count);

/* --------------------------------------------- StatisticalOutlierRemoval ---- */

/* Replaces pcl::StatisticalOutlierRemoval setMeanK + setStddevMulThresh
 * [+ setNegative] + filter (pcl_tools/outlier_removal.cpp:80-84, :91-93).
 * out_kept_index: caller-allocated n entries, receives the kept indices in input
 * order; *out_count their number.  out_mean_dist (n, or NULL): per-point mean
 * neighbour distance.  out_stats (or NULL): {mean, stddev, threshold}. */
int lc3d_sor(lc3d_ctx* ctx, const lc3d_cloud* cloud, int32_t mean_k, double stddev_mul,
             int32_t negative, int32_t* out_kept_index, int64_t* out_count, float* out_mean_dist,
             double out_stats[3]);

/* ------------------------------------------------------- accumulate_clouds --- */

/* Replaces the per-target-point pcl::CropBox (negative) loop of
 * pcl_tools/accumulate_clouds.cpp:100-111: a source point is dropped iff
 * some target point t has it inside the axis-aligned box [t - radius, t + radius] (corners
 * rounded to float, bounds inclusive, as CropBox evaluates them); non-finite source points are
 * dropped.  O(N) on the grid index instead of O(N*M).  out_kept_index: caller-allocated
 * source->n entries, receives the surviving source indices in source order. */
int lc3d_box_dedup(lc3d_ctx* ctx, const lc3d_cloud* source, const lc3d_cloud* target, double radius,
                   int32_t* out_kept_index, int64_t* out_count);

/* ------------------------------------------------------- cluster_extraction --- */

/* Replaces pcl::EuclideanClusterExtraction::extract as pcl_tools/cluster_extraction.cpp:88-101
 * drives it (setClusterTolerance / setMinClusterSize / setMaxClusterSize / extract; SURVEY 8f
 * rank 4).  Clusters = connected components of the graph with an edge wherever the float32
 * squared distance is < (float)(tolerance^2) (KdTreeFLANN::radiusSearch is strict), kept iff
 * min_size <= size <= max_size, ranked by size descending (ties: the cluster holding the lower
 * point index first — PCL leaves ties unspecified).  Non-finite points belong to no cluster.
 * out_labels: caller-allocated cloud->n entries, rank of the point's cluster or -1;
 * out_sizes (may be NULL): the first min(count, sizes_cap) cluster sizes by rank. */
int lc3d_euclidean_clusters(lc3d_ctx* ctx, const lc3d_cloud* cloud, double tolerance, int64_t min_size,
                            int64_t max_size, int32_t* out_labels, int64_t* out_sizes, int64_t sizes_cap,
                            int64_t* out_count);

/* ------------------------------------------------- in-process view pipeline -- */

/* The per-view stages that scripts/alignment.sh:99-100 (outlier_removal, normals) and the
 * turntable chain of BASELINE configs[2] (VoxelGrid 2 mm + SOR k=50 + normals) run before the
 * pairwise ICP, chained on the device: no PLY file and no host copy between the stages (SURVEY §8f
 * rank 3).  Each stage is exactly the computation of its stage-by-stage entry point
 * (lc3d_voxel_grid / lc3d_sor / lc3d_normals), so the result is bit-identical to calling them
 * in sequence through host buffers.  A stage is skipped when its parameter is <= 0. */
typedef struct lc3d_prepare_params {
  float leaf_size;        /* pcl::VoxelGrid cubic leaf (cloud_downsampling.cpp:74), <= 0: skip          */
  int32_t sor_mean_k;     /* StatisticalOutlierRemoval meanK (outlier_removal.cpp:81), <= 0: skip      */
  double sor_stddev_mul;  /* setStddevMulThresh (outlier_removal.cpp:82)                                */
  int32_t normals_k;      /* NormalEstimation setKSearch (normal_estimation.cpp:96), <= 0: no normals  */
  float viewpoint[3];     /* setViewPoint (normal_estimation.cpp:98-105)                                */
} lc3d_prepare_params;

/* xyz of `cloud` in, resident cloud out (xyz, and normals + curvature when normals_k > 0), ready
 * for lc3d_icp_align_resident.  counts (may be NULL): points after VoxelGrid, after SOR, final. */
int lc3d_prepare_view(lc3d_ctx* ctx, const lc3d_cloud* cloud, const lc3d_prepare_params* params,
                      lc3d_dcloud** out, int64_t counts[3]);

/* Copies a resident cloud back: out_xyz n x 3; out_normal n x 3 and out_curvature n may be NULL
 * (they are left untouched when the cloud carries no normals). */
int lc3d_cloud_download(lc3d_ctx* ctx, const lc3d_dcloud* dc, float* out_xyz, float* out_normal,
                        float* out_curvature);

/* ------------------------------------------------------- view chain on one device ----- */

/* The pair block of a turntable chain as a task graph on ONE device (scripts/alignment.sh:99-113 and
 * the :123-126 TODO, per GPU): every view prepared once (lc3d_prepare_view), every pair aligned as
 * soon as both of its views are ready (lc3d_icp_align_resident), on prepare_threads + align_threads
 * host threads that each own an lc3d_ctx (own stream, own scratch).  The executor keeps its contexts
 * between runs, so a steady stream of chains allocates nothing. */
typedef struct lc3d_chain lc3d_chain;
int lc3d_chain_create(int device, int32_t prepare_threads, int32_t align_threads, lc3d_chain** out);
void lc3d_chain_destroy(lc3d_chain* chain);
const char* lc3d_chain_last_error(const lc3d_chain* chain);
/* views[0..n_views): host clouds (xyz) of consecutive views; pair i (0 <= i < n_views - 1) registers
 * view i+1 (source) onto view i (target).  results[n_views - 1]; points_per_view[n_views] (may be
 * NULL) = points of every prepared view.  warm != 0: instead of the task graph, EVERY context
 * prepares every view / aligns every pair once (scratch buffers grow to their final sizes; the
 * results are still written).  Returns LC3D_OK or the first error (lc3d_chain_last_error). */
int lc3d_chain_run(lc3d_chain* chain, const lc3d_cloud* views, int32_t n_views, const lc3d_prepare_params* prepare,
                   const lc3d_icp_params* icp, lc3d_icp_result* results, int64_t* points_per_view, int32_t warm);

/* ------------------------------------------------------------ transform ----- */

/* pcl::transformPointCloudWithNormals (pcl_tools/transform.cpp:84-90; SURVEY §8f
 * rank 1): xyz' = T * (xyz,1), n' = R * n, float32. matrix: row-major 4x4. */
int lc3d_transform(lc3d_ctx* ctx, const lc3d_cloud* cloud, const float matrix[16], float* out_xyz,
                   float* out_normal);

#ifdef __cplusplus
}
#endif
#endif /* LC3D_H_ */
