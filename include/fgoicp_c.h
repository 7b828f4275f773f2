/* fgoicp_c.h -- C ABI of the B200-native Go-ICP hot path.
 *
 * This is the drop-in boundary beneath the reference's C++ class icp::FastGoICP
 * (reference fgoicp/fgoicp.hpp:10-108).  The reference has no FFI of its own; each entry
 * point below replaces one reference operator so that parity can be checked call by call.
 * Citations are file:line inside the reference repository (solemnwind/fast-go-icp).
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are HOST pointers unless the parameter
 *     name starts with d_ (device pointer on the context's device);
 *   - 3x3 matrices are 9 floats in COLUMN-major order (glm::mat3 memory order), y = R*x + t;
 *   - a "cube" is 4 floats (cx, cy, cz, half_span); rotation cubes live in the
 *     quaternion-vector unit ball (reference fgoicp/common.hpp:30-57), translation cubes in
 *     the normalised frame [-1,1]^3 (reference fgoicp/fgoicp.cpp:113);
 *   - every function returns an int status: 0 = ok, negative = error; the message of the
 *     last error on the calling thread is available from fgoicp_last_error();
 *   - nothing throws, nothing calls exit(); a context is used from one host thread at a time;
 *   - there is NO CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef FGOICP_C_H
#define FGOICP_C_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FGOICP_OK            0
#define FGOICP_ERR_ARG      (-1)
#define FGOICP_ERR_CUDA     (-2)
#define FGOICP_ERR_STATE    (-3)
#define FGOICP_ERR_NOMEM    (-4)

/* How the nearest-distance grid is sampled by the bound kernels (all three return the
 * reference's texture semantics: tri-linear filter of SQUARED distances at q - res/2 with
 * 8-bit weights and clamp-to-edge, reference fgoicp/registration.cu:214-234, 320-328). */
#define FGOICP_SAMPLER_GRID    0   /* 8 scalar gathers from the dense x-fastest grid in HBM   */
#define FGOICP_SAMPLER_PACKED  1   /* one 32-byte gather from the corner-packed cell grid     */
#define FGOICP_SAMPLER_TEX     2   /* hardware tex3D on a cudaArray (the reference's own path) */

/* ctx_create flags */
#define FGOICP_BUILD_PACKED    (1u << 0)   /* build the corner-packed grid (8x the dense grid) */
#define FGOICP_BUILD_TEX       (1u << 1)   /* build the cudaArray + texture object             */
#define FGOICP_BUILD_BRUTE_LUT (1u << 2)   /* build the grid by tiled brute force (test hook)  */
#define FGOICP_BUILD_KEEP_ORDER (1u << 3)  /* keep the data points in the caller's order on the device (default: Morton order) */
#define FGOICP_BUILD_DEFAULT   (FGOICP_BUILD_PACKED | FGOICP_BUILD_TEX)

typedef struct fgoicp_ctx fgoicp_ctx;

typedef struct fgoicp_info
{
    uint64_t nt, ns;            /* model / data point counts                                   */
    int32_t  dims[3];           /* grid nodes per axis = ceil(range / res)  (registration.cu:186) */
    float    resolution;
    float    scale;             /* 1 / resolution (registration.cu:198)                        */
    float    offset[3];         /* -bbox_min      (registration.cu:199-201)                    */
    uint64_t grid_bytes;        /* dense grid bytes in HBM                                     */
    uint64_t packed_bytes;      /* corner-packed grid bytes in HBM (0 if not built)            */
    int32_t  device;
    int32_t  sm_count;
    int32_t  sampler;           /* currently selected FGOICP_SAMPLER_*                         */
    int32_t  has_packed, has_tex;
    float    build_ms;          /* device time of the grid build                               */
} fgoicp_info;

const char* fgoicp_last_error(void);
const char* fgoicp_version(void);

/* FastGoICP constructor preprocessing on the device (fgoicp/fgoicp.hpp:13-19):
 *   center_point_cloud(source), center_point_cloud(target)   fgoicp/fgoicp.cpp:176-195
 *   scale_point_clouds(target, source)                       fgoicp/fgoicp.cpp:197-220, 271-287
 *   get_point_cloud_ranges(target)                           fgoicp/fgoicp.cpp:222-268
 * Both clouds (n*3 floats each) are centred and scaled IN PLACE.  With flags = FGOICP_PRE_REFERENCE every output is
 * bit-identical to the reference's host code: the centroid is the serial fp32 sum in index order, the scale is
 * 1 / max|coordinate| of the centred SOURCE.
 *   FGOICP_PRE_TREE_CENTROID  centroid from a deterministic parallel fp64 reduction (fixed partition and combination
 *                             order) rounded once to fp32 -- for clouds of millions of points; last-bit differences
 *                             from the reference's centroid;
 *   FGOICP_PRE_SCALE_BOTH     scale = 1 / max|coordinate| over BOTH centred clouds, so the target also lies inside
 *                             [-1,1]^3, the domain the translation search covers (fgoicp.cpp:113; the reference
 *                             scales by the source alone and lets the target stick out).
 * fgoicp_preprocess takes host buffers (copies in, computes on `device`, copies back); fgoicp_preprocess_dev takes
 * device buffers and a CUDA stream (cudaStream_t as void*, NULL = default stream) and returns after the stream has
 * finished.  The outputs feed fgoicp_ctx_create (bbox_min / bbox_max) and restore_translation (fgoicp.hpp:87-90). */
#define FGOICP_PRE_REFERENCE      0u
#define FGOICP_PRE_TREE_CENTROID  (1u << 0)
#define FGOICP_PRE_SCALE_BOTH     (1u << 1)

typedef struct fgoicp_normalisation
{
    float offset_pcs[3];        /* -centroid of the source, as center_point_cloud returns it (fgoicp.cpp:194) */
    float offset_pct[3];        /* -centroid of the target                                                     */
    float scale;                /* scaling factor applied to both clouds                                       */
    float bbox_min[3];          /* per-axis range of the centred, scaled target                                */
    float bbox_max[3];
    float device_ms;            /* device time of the preprocessing kernels                                    */
} fgoicp_normalisation;

int fgoicp_preprocess(float* model_xyz, size_t nt, float* data_xyz, size_t ns,
                      int device, unsigned flags, fgoicp_normalisation* out);
int fgoicp_preprocess_dev(float* d_model_xyz, size_t nt, float* d_data_xyz, size_t ns,
                          int device, unsigned flags, void* cuda_stream, fgoicp_normalisation* out);

/* Registration + NearestNeighborLUT constructor (registration.hpp:68-80, registration.cu:180-207,
 * 258-318): uploads both clouds (already centred and scaled by the caller, fgoicp.hpp:16-18) and
 * builds the nearest-SQUARED-distance grid over [bbox_min, bbox_max] of the model cloud.
 * model_xyz: nt*3 floats, data_xyz: ns*3 floats. */
int fgoicp_ctx_create(const float* model_xyz, size_t nt,
                      const float* data_xyz, size_t ns,
                      const float bbox_min[3], const float bbox_max[3],
                      float lut_resolution, int device, unsigned flags,
                      fgoicp_ctx** out);
int fgoicp_ctx_destroy(fgoicp_ctx* ctx);
int fgoicp_ctx_info(const fgoicp_ctx* ctx, fgoicp_info* out);
int fgoicp_set_sampler(fgoicp_ctx* ctx, int sampler);
/* Trimmed registration (EXTENSION -- the reference parses `trim` and ignores it, src/utilities.hpp:94,
 * fgoicp/fgoicp.hpp:73): with trim_fraction rho > 0 every sum over the data points -- per-cube upper and lower
 * bounds, the exact SSE, the centroids and cross-covariance of the ICP -- runs over the
 * K = ns - floor(ns * rho) points with the smallest residual only.  rho = 0 (default) is the reference's behaviour.
 * *inliers (optional) receives K.  Affects every later call on the context; the bound kernel then keeps one key per
 * data point (its signed residual: both bounds are monotone in it, so ONE exact select serves both sums) in shared memory
 * (ns <= 51,200) or, for larger clouds, in an L2-resident scratch slice per thread block, and the inner searches run
 * round-synchronously.  Switching trimming on allocates the per-level scratch of that schedule (~45 KB per rotation cube,
 * sized for 4,096 cubes) so that the searches allocate nothing. */
int fgoicp_set_trim(fgoicp_ctx* ctx, float trim_fraction, uint64_t* inliers);
/* CUDA stream (cudaStream_t passed as void*) every later call on this context enqueues on;
 * NULL selects the context's own stream.  Lets a host framework time calls with its own events. */
int fgoicp_set_stream(fgoicp_ctx* ctx, void* cuda_stream);

/* Test hooks ------------------------------------------------------------------------------- */
/* Nearest-neighbour engine behind fgoicp_sse / fgoicp_nn / fgoicp_icp: 0 = uniform cell grid in HBM (default; far
 * queries cull whole blocks of cells through their bounding boxes), 1 = tiled brute force, 2 = cell grid with the
 * bounding-box culling applied to EVERY query, 3 = cell grid without it.  All are exact and return identical indices. */
int fgoicp_set_nn_mode(fgoicp_ctx* ctx, int mode);
/* Dense grid download, x fastest: out[(z*dims[1]+y)*dims[0]+x]  (registration.cu:276-277). */
int fgoicp_lut_download(fgoicp_ctx* ctx, float* out, size_t out_floats);
/* NearestNeighborLUT::search for n query points with a chosen sampler (registration.cu:320-328). */
int fgoicp_lut_sample(fgoicp_ctx* ctx, const float* q_xyz, size_t n, int sampler, float* out_d2);
/* Evaluation order of fgoicp_bounds_multi / _multi_dev: 1 = z-phase-ordered kernel (default with the packed
 * sampler; same results, L2-friendly), 0 = plain kernel. */
int fgoicp_set_phased(fgoicp_ctx* ctx, int on);
/* Schedule of the inner searches behind fgoicp_bnb_r3_batch / fgoicp_so3_level_*: 0 = default (one persistent
 * thread-block cluster per search), 1 = same, explicitly, 2 = round-synchronous (all searches advance one
 * iteration per round, bounds of each round through the phase-ordered kernel; selecting it allocates its per-level
 * scratch).  Same results either way. */
int fgoicp_set_bnb_mode(fgoicp_ctx* ctx, int mode);
/* Driver of the ICP refinements behind fgoicp_icp / fgoicp_icp_batch / fgoicp_so3_level_ub:
 *   2 = one persistent cooperative kernel per batch: the whole loop of icp3d.cu:85-108 on the device, stages separated
 *       by grid-wide barriers, ONE host synchronisation per batch;
 *   1 = one launch per stage and iteration with a host poll every 8 iterations (also used by trimmed runs);
 *   0 = automatic (default): the kernel for batches that fit the pool of instance slots, the launches for batches that
 *       refill their slots many times (pure throughput; measured ~10 % faster there).
 * Same results in every mode, bit for bit. */
int fgoicp_set_icp_mode(fgoicp_ctx* ctx, int mode);
/* Measurement hook: useful GB/s of independent random gathers of width_bytes (16/32/64/128) over a
 * buffer of `bytes` bytes -- the gather roofline the bound kernels are compared with. */
int fgoicp_gather_probe(fgoicp_ctx* ctx, size_t bytes, int width_bytes, int blocks_per_sm, float* out_gbps);
/* sinf(span * sqrt3 * pi / 2) exactly as the device evaluates it in the bound kernels
 * (registration.cu:41-42). */
int fgoicp_rot_sin(fgoicp_ctx* ctx, const float* spans, int n, float* out);

/* Registration::compute_sse_error(rnode, tnodes, fix_rot, pool)  (registration.cu:88-152):
 * one rotation (matrix R of the cube centre, half-span rot_span) against T translation cubes.
 * Outputs lb[T], ub[T] (the reference returns {lower, upper}, registration.cu:151). */
int fgoicp_bounds_batch(fgoicp_ctx* ctx, const float R[9], float rot_span, int fix_rot,
                        const float* t_xyz_span, int T, float* lb, float* ub);

/* Same operator over Rn rotation cubes, each with its own list of T translation cubes:
 * rot_xyz_span[Rn][4], t_xyz_span[Rn][T][4] -> lb[Rn][T], ub[Rn][T].  One fused launch.
 * Rotation cubes whose centre lies outside the unit ball get R = I (common.hpp:41). */
int fgoicp_bounds_multi(fgoicp_ctx* ctx, const float* rot_xyz_span, int Rn, int fix_rot,
                        const float* t_xyz_span, int T, float* lb, float* ub);
/* Device-resident form of the same call (inputs and outputs already in HBM); asynchronous on
 * the context stream.  d_best_ub (optional, 1 float) receives min over all ub. */
int fgoicp_bounds_multi_dev(fgoicp_ctx* ctx, const float* d_rot_xyz_span, int Rn, int fix_rot,
                            const float* d_t_xyz_span, int T, float* d_lb, float* d_ub,
                            float* d_best_ub);

/* Registration::compute_sse_error(R, t)  (registration.cu:62-86, 154-174): exact
 * sum_i min_j |R p_i + t - m_j|^2. */
int fgoicp_sse(fgoicp_ctx* ctx, const float R[9], const float t[3], float* sse);

/* Nearest model point of every transformed data point (test hook for K5 / K7):
 * rooted = 0: squared-distance compare, lowest index wins ties (registration.cu:160-172);
 * rooted = 1: compare sqrt(d2) as kernFindNearestNeighbor does (icp3d.cu:17-26).
 * idx[ns] (may be NULL), d2[ns] (may be NULL). */
int fgoicp_nn(fgoicp_ctx* ctx, const float R[9], const float t[3], int rooted,
              int32_t* idx, float* d2);

/* IterativeClosestPoint3D(reg, pct, pcs, max_iter, thr, R0, t0).run()  (icp3d.cu:55-108). */
int fgoicp_icp(fgoicp_ctx* ctx, const float R0[9], const float t0[3], int max_iter, float thr,
               float* sse, float R[9], float t[3], int* iters);

/* n independent refinements (one IterativeClosestPoint3D(...).run() each, icp3d.cu:55-108) from the seed poses
 * R0s[n][9], t0s[n][3], run concurrently on a pool of instance slots; a slot that finishes takes the next pending
 * seed on the device.  Every output equals what fgoicp_icp returns for that seed alone.
 * Outputs (each may be NULL): sse[n], R[n][9], t[n][3], iters[n]. */
int fgoicp_icp_batch(fgoicp_ctx* ctx, const float* R0s, const float* t0s, int n, int max_iter, float thr,
                     float* sse, float* R, float* t, int* iters);

/* FastGoICP::branch_and_bound_R3(rnode, fix_rot)  (fgoicp.cpp:102-174) as a GPU-resident
 * best-first search: pool, batch selection, bound evaluation, pruning and child spawning all
 * stay on the device.  best_sse seeds best_error (fgoicp.cpp:104).  Outputs the reference's
 * {best_ub, best_t} plus the number of (cube x point) bound evaluations spent. */
int fgoicp_bnb_r3(fgoicp_ctx* ctx, const float rot_xyz_span[4], int fix_rot,
                  float best_sse, float sse_threshold,
                  float* best_ub, float best_t[3], uint64_t* evals);
/* Rn independent inner searches in one launch (one thread-block per rotation cube). */
int fgoicp_bnb_r3_batch(fgoicp_ctx* ctx, const float* rot_xyz_span, int Rn, int fix_rot,
                        float best_sse, float sse_threshold,
                        float* best_ub, float* best_t, uint64_t* evals);

/* One level of the outer SO(3) search (fgoicp.cpp:49-97) over this process's shard of the
 * level's child cubes: inner search with fixed rotation, ICP on promising cubes, inner search
 * with rotation uncertainty.
 *   in : cubes[n][4] (children whose centre is inside the unit ball), best_sse (global, level start)
 *   out: ub[n], bt[n][3] (fix_rot = true result), lb[n] (fix_rot = false result, computed with
 *        *io_best_sse after this shard's ICPs), io_best_sse/io_best_R/io_best_t updated if an ICP
 *        improved on them, n_icp, evals.
 * The caller min-reduces best_sse over shards between levels (NCCL across processes). */
typedef struct fgoicp_level_stats
{
    uint64_t evals;        /* cube x point bound evaluations                                  */
    uint32_t n_icp;        /* ICP refinements run                                             */
    uint32_t icp_iters;    /* total ICP iterations                                            */
    float    ms_bnb_ub;    /* device ms: inner searches, fixed rotation                       */
    float    ms_icp;       /* device ms: ICP                                                  */
    float    ms_bnb_lb;    /* device ms: inner searches, rotation uncertainty                 */
    int32_t  best_icp_index; /* index (into cubes) of the ICP that improved io_best_sse, -1 if none */
} fgoicp_level_stats;

int fgoicp_so3_level_ub(fgoicp_ctx* ctx, const float* cubes, int n,
                        float best_sse, float sse_threshold,
                        float* ub, float* bt,
                        float* io_best_sse, float io_best_R[9], float io_best_t[3],
                        fgoicp_level_stats* stats);
int fgoicp_so3_level_lb(fgoicp_ctx* ctx, const float* cubes, int n,
                        float best_sse, float sse_threshold,
                        float* lb, fgoicp_level_stats* stats);

#ifdef __cplusplus
}
#endif

#endif /* FGOICP_C_H */
