/*
 * pano_b200.h — C ABI of the B200-native SIFT + match + blend engine.
 *
 * This is the drop-in boundary for the hot path of ppwwyyxx/OpenPano
 * (SURVEY.md §8b).  The reference has no FFI layer: its seams are four C++
 * classes.  Each entry point below names the reference interface it replaces
 * (paths relative to the reference's src/).  Plain pointers and sizes only; no
 * torch / C++ types.  All functions return 0 on success or a negative
 * pano_status; they never call exit().
 *
 * Threading: a pano_ctx owns one CUDA stream, a private stream-ordered memory
 * pool and pinned scratch; calls on one ctx must be serialized by the caller
 * (create one ctx per host thread, or use the *_batch entry points, which is
 * how the reference's `#pragma omp parallel for` over images maps to this
 * engine).  Different contexts may be driven from different host threads at
 * the same time; every entry point makes its context's device current for the
 * calling thread, so a ctx can be used from any thread on a multi-GPU host.
 * Work of different contexts overlaps on the device; order it with pano_event_*.
 */
#ifndef PANO_B200_H
#define PANO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef enum pano_status {
  PANO_OK = 0,
  PANO_ERR_CUDA = -1,        /* a CUDA runtime call failed: pano_last_error() */
  PANO_ERR_INVALID = -2,     /* bad argument (null pointer, non-positive size…) */
  PANO_ERR_CAPACITY = -3,    /* a fixed-capacity device list overflowed */
  PANO_ERR_NO_DEVICE = -4,   /* no CUDA device / extension built without one */
  PANO_ERR_NO_FEATURE = -5   /* reference: error_exit("Cannot find feature…"), stitcherbase.cc:20 */
} pano_status;

/* Snapshot of the reference's mutable config globals (lib/config.hh:24-68,
 * defaults from config.cfg:2-69) that the hot path reads. */
typedef struct pano_params {
  int   sift_working_size;         /* SIFT_WORKING_SIZE 800 */
  int   num_octave;                /* NUM_OCTAVE 4 */
  int   num_scale;                 /* NUM_SCALE 7 */
  float scale_factor;              /* SCALE_FACTOR 1.4142135623 */
  float gauss_sigma;               /* GAUSS_SIGMA 1.4142135623 */
  int   gauss_window_factor;       /* GAUSS_WINDOW_FACTOR 6 */
  float judge_extrema_diff_thres;  /* JUDGE_EXTREMA_DIFF_THRES 2e-3 */
  float contrast_thres;            /* CONTRAST_THRES 4e-2 */
  float pre_color_thres;           /* PRE_COLOR_THRES 5e-2 */
  float edge_ratio;                /* EDGE_RATIO 6 */
  int   calc_offset_depth;         /* CALC_OFFSET_DEPTH 4 */
  float offset_thres;              /* OFFSET_THRES 0.5 */
  float ori_radius;                /* ORI_RADIUS 4.5 */
  int   ori_hist_smooth_count;     /* ORI_HIST_SMOOTH_COUNT 2 */
  int   desc_hist_scale_factor;    /* DESC_HIST_SCALE_FACTOR 3 */
  int   desc_int_factor;           /* DESC_INT_FACTOR 512 */
  float match_reject_next_ratio;   /* MATCH_REJECT_NEXT_RATIO 0.8 */
  float focal_length;              /* FOCAL_LENGTH 37 */
  int   ordered_input;             /* ORDERED_INPUT 0 */
  int   lazy_read;                 /* LAZY_READ 1 */
  int   multiband;                 /* MULTIBAND 0 */
  int   max_output_size;           /* MAX_OUTPUT_SIZE 8000 */
} pano_params;

/* Fills *p with the defaults of the reference's config.cfg. */
void pano_params_default(pano_params* p);

/* One scale-space point: POD image of the reference's SSPoint
 * (feature/feature.hh:33-39). */
typedef struct pano_sspoint {
  int    x, y;            /* Coor coor: integer coordinate in the octave */
  double real_x, real_y;  /* Vec2D real_coor in [0,1) */
  int    pyr_id, scale_id;
  float  dir;
  float  scale_factor;
} pano_sspoint;

/* ---------------------------------------------------------------- context */

typedef struct pano_ctx pano_ctx;

/* Creates an engine context on CUDA device `device`.  `cuda_stream` may be
 * NULL (the ctx creates its own non-blocking stream) or a cudaStream_t cast to
 * void* (e.g. torch.cuda.current_stream().cuda_stream) on which every kernel
 * of this ctx is then launched. */
int  pano_create(pano_ctx** out, int device, void* cuda_stream);
void pano_destroy(pano_ctx* ctx);
/* Device memory: every buffer a ctx needs comes stream-ordered from its own pool, and freed blocks
 * are kept by the ctx for the next request of about the same size (a stitch job asks for the same
 * sizes every time; re-arranging the pool for a 0.9 GB arena was seen to block the host for up to
 * 1.5 s).  The environment variable PANO_CACHE_MB bounds what a ctx keeps (default 32768, 0 = keep
 * nothing); pano_trim hands everything it keeps back to the pool, e.g. before a long idle period. */
int  pano_trim(pano_ctx* ctx);
/* Message of the last failure on this ctx ("" if none).  ctx may be NULL for
 * the last pano_create failure. */
const char* pano_last_error(const pano_ctx* ctx);
/* Blocks until all work queued on the ctx stream has finished. */
int  pano_sync(pano_ctx* ctx);
/* The stream (cudaStream_t as void*) the ctx launches on. */
void* pano_stream(pano_ctx* ctx);

/* Per-kernel timing (CUDA events on the ctx stream).  When enabled, every
 * kernel launch is bracketed by events; pano_kernel_times reports, per kernel
 * name, the launch count and summed duration since the last reset.  Used by
 * bench.py for the roofline line; adds host overhead, so it is off by default
 * and never on inside a timed end-to-end region. */
int  pano_profile_enable(pano_ctx* ctx, int on);
int  pano_profile_reset(pano_ctx* ctx);
/* names: buffer of cap entries × 64 bytes; returns number of distinct kernels. */
int  pano_profile_read(pano_ctx* ctx, int cap, char* names, int* launches, double* total_ms);
/* Total number of kernels this ctx has launched since creation. */
long long pano_launch_count(const pano_ctx* ctx);
/* Diagnostics: how many descriptor rows the last pano_match_pairs_dev call had to
 * re-scan exactly because the tensor-core nomination was not certain. */
int pano_match_last_exact_rows(const pano_ctx* ctx);
/* Diagnostics: rows of the LARGER sets that the last pano_match_pairs_dev call nominated on request
 * ("columns on demand": on large runs only the smaller set of each pair goes through the first tensor
 * pass; matcher.cc:57-61 walks column j only for rows that passed their own ratio test).  0 when
 * both sets went through the first pass. */
int pano_match_last_nominated_rows(const pano_ctx* ctx);

/* ---------------------------------------------------------------- features
 * Replaces FeatureDetector::detect_feature / SIFTDetector::do_detect_feature
 * (feature/feature.hh:42-57, feature/feature.cc:20-47): working-size resize,
 * ScaleSpace (feature/dog.cc:96-114), DOGSpace (dog.cc:131-143),
 * ExtremaDetector::get_extrema (extrema.cc:36-61), OrientationAssign::work
 * (orientation.cc:22-32), SIFT::get_descriptor (sift.cc:77-85).
 *
 * A pano_featureset holds, per image, the descriptors (n×128 f32, row-major)
 * and coordinates (n×2 f64, image-centred input pixels: (c-0.5)*w) ON THE
 * DEVICE so that matching runs without a host round trip; download copies
 * them out in the reference's Descriptor layout (feature.hh:18-30). */
typedef struct pano_featureset pano_featureset;

/* Host images: n pointers to H×W×3 f32 RGB in [0,1] (Mat32f layout,
 * lib/mat.h:7-60).  Includes H2D of the images.  Pageable buffers are staged and may
 * be reused as soon as the call returns; PINNED buffers (pano_host_alloc /
 * cudaHostAlloc) are read by an asynchronous copy and must stay untouched until the
 * first pano_featureset_count / _download / pano_match_* on the result, or pano_sync.
 *
 * Capacity: per-image keypoint lists start at 8192 entries and are NOT a limit — when
 * an image overflows them (the reference's vectors are unbounded, extrema.cc:56-57) the
 * batch is run again with doubled lists at the first count query, transparently.
 * One parameter limit the reference does not have: descriptor windows of at most 127 pixels
 * radius (sqrt(1/2) * GAUSS_SIGMA * max(1, SCALE_FACTOR) * DESC_HIST_SCALE_FACTOR * 5 <= 127;
 * the defaults give 21) — wider settings return PANO_ERR_INVALID. */
int pano_sift_detect_batch(pano_ctx* ctx, int n, const float* const* rgb_hwc,
                           const int* w, const int* h, const pano_params* p,
                           pano_featureset** out);
/* Same, images already resident in device memory (device pointers).  The images must
 * stay valid until the first count query / download / match on the featureset (they are
 * read again if a keypoint list has to grow). */
int pano_sift_detect_batch_dev(pano_ctx* ctx, int n, const float* const* d_rgb_hwc,
                               const int* w, const int* h, const pano_params* p,
                               pano_featureset** out);
/* Single-image convenience = detect_feature(const Mat32f&). */
int pano_sift_detect(pano_ctx* ctx, const float* rgb_hwc, int w, int h,
                     const pano_params* p, pano_featureset** out);

/* Builds a featureset from host descriptors = PairWiseMatcher ctor
 * (feature/matcher.hh:40-46; matcher.cc:73-88 without the kd-forest).
 * desc[i]: n_kp[i]×128 f32; coor[i] may be NULL. */
int pano_featureset_upload(pano_ctx* ctx, int n_images, const int* n_kp,
                           const float* const* desc, const double* const* coor_xy,
                           pano_featureset** out);
/* The same from DEVICE pointers, queued on the context's stream without a host
 * sync (the sources must stay valid until the stream has passed this call): the
 * receiving side of the multi-GPU descriptor exchange. */
int pano_featureset_import_dev(pano_ctx* ctx, int n_images, const int* n_kp,
                               const float* const* d_desc, const double* const* d_coor_xy,
                               pano_featureset** out);
/* Copies image i's rows device-to-device (the sending side); either may be NULL. */
int pano_featureset_export_dev(pano_featureset* fs, int image, double* d_coor_xy, float* d_desc);
/* All images at once: image 0's rows, then image 1's, ... packed back to back (one launch instead
 * of two copies per image); destinations must be 16-byte aligned. */
int pano_featureset_export_all_dev(pano_featureset* fs, double* d_coor_xy, float* d_desc);
int pano_featureset_num_images(const pano_featureset* fs);
/* Number of descriptors of image i (synchronizes on first use). */
int pano_featureset_count(pano_featureset* fs, int image);
/* Copies image i's results to host: coor_xy (2·n f64) and desc (128·n f32);
 * either may be NULL. */
int pano_featureset_download(pano_featureset* fs, int image, double* coor_xy, float* desc);
/* SIFT sets only: image i's keypoints as SIFTDetector::do_detect_feature returns them
 * (feature/sift.cc:150, Descriptor::coor = SSPoint::real_coor in [0,1)), i.e. BEFORE
 * FeatureDetector::detect_feature scales them to image-centred pixels (feature.cc:20-28).
 * A FeatureDetector subclass returns these and lets the base class do its scaling. */
int pano_featureset_download_real(pano_featureset* fs, int image, double* real_xy);
void pano_featureset_free(pano_featureset* fs);

/* --------------------------------------------------------------- matching
 * Replaces PairWiseMatcher::match(i, j) (feature/matcher.cc:90-135) with the
 * exact rule of FeatureMatcher::match (matcher.cc:15-71), which is the parity
 * contract: loop over the smaller set, exact fp32 top-2 with
 * lowest-index ties, two-way ratio test with REJECT_RATIO_SQR = ratio².
 * Output pairs are (idx in image i, idx in image j), ascending index of the
 * smaller set. */
typedef struct pano_matches {
  int  n_pairs;      /* number of image pairs */
  int* count;        /* [n_pairs] matches of pair k */
  int* offset;       /* [n_pairs+1] prefix into idx */
  int* idx;          /* [2*offset[n_pairs]] (first, second) */
} pano_matches;

int  pano_match_pairs(pano_ctx* ctx, pano_featureset* fs, int n_pairs,
                      const int* image_ij /* 2*n_pairs */, const pano_params* p,
                      pano_matches* out);
void pano_matches_free(pano_matches* m);
/* Device-resident variant for the timed-in-HBM bench leg: results stay on the
 * device, only the total match count is returned. */
int  pano_match_pairs_dev(pano_ctx* ctx, pano_featureset* fs, int n_pairs,
                          const int* image_ij, const pano_params* p, int* total_matches);
/* Row-sharded forms (multi-GPU brute-force match, BASELINE config 4): the loop of
 * FeatureMatcher::match over the smaller set (matcher.cc:32: `#pragma omp parallel for` over k)
 * is split into n_shards contiguous shares; this call decides share `shard` of EVERY pair —
 * rows [n_small*shard/n_shards, n_small*(shard+1)/n_shards).  Each share still scans all of the
 * larger set and, for its surviving rows, all of the smaller set (matcher.cc:57-61), so every
 * rank holds both descriptor sets and no collective is needed; the shards' lists, concatenated
 * in shard order, are pano_match_pairs' lists.  (0, 1) is the unsharded call. */
int  pano_match_pairs_shard(pano_ctx* ctx, pano_featureset* fs, int n_pairs, const int* image_ij,
                            const pano_params* p, int shard, int n_shards, pano_matches* out);
int  pano_match_pairs_dev_shard(pano_ctx* ctx, pano_featureset* fs, int n_pairs, const int* image_ij,
                                const pano_params* p, int shard, int n_shards, int* total_matches);
/* FeatureMatcher(f1,f2).match() on two host descriptor arrays
 * (matcher.hh:27-38); pairs_out holds 2*min(n,m) ints. */
int  pano_match_bruteforce(pano_ctx* ctx, const float* desc_a, int n,
                           const float* desc_b, int m, const pano_params* p,
                           int* pairs_out, int* n_pairs_out);

/* ------------------------------------------------- RANSAC inlier scoring
 * Replaces the scoring half of TransformEstimation::get_transform
 * (stitch/transform_estimate.cc:68-85): get_inliers (:132-148) for every hypothesis
 * of every pair, the FIRST hypothesis with the largest inlier count (update_max,
 * lib/utils.hh:58-63) and its inlier flags.  Hypothesis generation (sampling with
 * the caller's seeded generator + the normalised DLT, :89-130) stays on the host. */
typedef struct pano_ransac_pair {
  int n_match;
  const double* kp1_xy;   /* 2*n_match: kp1[match.data[i].first]  (image-centred pixels) */
  const double* kp2_xy;   /* 2*n_match: kp2[match.data[i].second] */
  int n_hyp;
  const double* homos;    /* 9*n_hyp: Homography::data, row-major, image 2 -> image 1 */
  float inlier_thres;     /* ransac_inlier_thres (transform_estimate.cc:47) */
} pano_ransac_pair;
/* best_hyp[k] (-1 when pair k has no hypothesis), best_count[k]; hyp_counts[k] (n_hyp ints)
 * and inlier_flags[k] (n_match bytes) are optional per pair (array or entries may be NULL). */
int pano_ransac_score_pairs(pano_ctx* ctx, int n_pairs, const pano_ransac_pair* pairs, int* best_hyp,
                            int* best_count, int* const* hyp_counts, unsigned char* const* inlier_flags);

/* ------------------------------------- bundle-adjustment Jacobian assembly
 * Replaces the per-point part of IncrementalBundleAdjuster::calcJacobianSymbolic
 * (stitch/incremental_bundle_adjuster.cc:306-383): the two rows of J of every point match
 * (:355-361) and the sums of J^T J (:363-382), bit-identical to the reference's loops.
 * The per-pair 3x3 algebra in front of them (:288-304 and the loop-invariant products inside
 * the loop) is Eigen-backed in the reference (Homography::operator*, inverse,
 * Camera::rotation_to_angle) and stays with the caller, who hands in per pair:
 *   m[0]      Hto_to_from = (fromK * c_from.R) * (toRinv * toKinv)            (:304)
 *   m[1]      c_from.R * toRinv * toKinv                                       (:323)
 *   m[2]      toRinv * toKinv                                                  (:332)
 *   m[3..5]   fromK * dRfromdvi[k]                                             (:333-335)
 *   m[6]      toKinv                                                           (:339,349)
 *   m[7..9]   m * dKdfocal, m * dKdppx, m * dKdppy,  m = fromK * c_from.R * toRinv * toKinv  (:338-345)
 *   m[10..12] (fromK * c_from.R) * dRtodviT[k]                                 (:348-352)
 * row-major Homography::data each. */
typedef struct pano_ba_pair {
  int from, to;         /* camera slots: index_map[pair.from], index_map[pair.to] */
  int match_begin;      /* match_cnt_prefix_sum[pair_idx]: pairs are consecutive, the first starts at 0 */
  int n_match;          /* pair.m.match.size() */
  double m[13][9];
} pano_ba_pair;
/* pts_to: p.first of every match (2 doubles each), all pairs concatenated.
 * j_rows (optional, 24 doubles per match): J(idx, param_idx_from + 0..5), J(idx, param_idx_to + 0..5),
 *   then the same 12 entries of row idx + 1 — every other entry of those rows is zero.
 * jtj: (6 n_cam)^2 doubles, row-major, fully written (JtJ.setZero() + the sums). */
int pano_ba_jacobian(pano_ctx* ctx, int n_cam, int n_pair, const pano_ba_pair* pairs, const double* pts_to,
                     double* j_rows, double* jtj);

/* ---------------------------------------------------------- cylinder warp
 * Replaces CylinderWarper(h_factor).warp(Mat32f&, vector<Vec2D>&)
 * (stitch/warp.hh:41-66, warp.cc:25-75). */
/* Output shape and offset for an input of w×h (CylinderProject::project(shape),
 * warp.cc:46-67); host-only arithmetic. */
int pano_cyl_warp_shape(int w, int h, double h_factor, const pano_params* p,
                        int* out_w, int* out_h, double* offset_x, double* offset_y);
/* out_hwc: out_h×out_w×3 f32 (Color::NO = -1 where unmapped); kpts_xy (n_kpts
 * pairs, image-centred) are rewritten in place; may be NULL/0. */
int pano_cyl_warp(pano_ctx* ctx, const float* rgb_hwc, int w, int h, double h_factor,
                  const pano_params* p, float* out_hwc, int out_w, int out_h,
                  double* kpts_xy, int n_kpts);

/* The warp loop of CylinderStitcher::build_warp (cylstitcher.cc:65-67: `REP(k, n) warper.warp(*imgs[k].img,
 * keypoints[k])`) as ONE launch over device-resident images: source and destination stay in HBM (the
 * warped images feed pano_blend_dev), the call is asynchronous on the ctx stream; only the keypoints —
 * a few thousand f64 pairs per image, host arithmetic — are rewritten in place before it returns. */
typedef struct pano_cyl_job {
  const float* d_rgb_hwc;   /* device, h×w×3 f32 */
  int w, h;
  float* d_out_hwc;         /* device, out_h×out_w×3 f32 (sizes from pano_cyl_warp_shape) */
  int out_w, out_h;
  double* kpts_xy;          /* host, n_kpts pairs, image-centred; may be NULL/0 */
  int n_kpts;
} pano_cyl_job;
int pano_cyl_warp_batch_dev(pano_ctx* ctx, int n, const pano_cyl_job* jobs, double h_factor, const pano_params* p);

/* ----------------------------------------------------------------- blend
 * Replaces BlenderBase::add_image + run (stitch/blender.hh:14-59) for
 * LinearBlender (blender.cc:24-96) and MultiBandBlender (multiband.cc:19-151).
 * The reference passes an opaque std::function per image; it is always the
 * closed form built at stitcher_image.cc:142-151 (and cylstitcher.cc:176-178),
 * so the ABI takes that form's parameters. */
typedef enum pano_projection { PANO_PROJ_FLAT = 0, PANO_PROJ_CYLINDRICAL = 1, PANO_PROJ_SPHERICAL = 2 } pano_projection;

typedef struct pano_blend_image {
  const float* rgb_hwc;   /* H×W×3 f32, host (pano_blend) or device (pano_blend_dev) */
  int w, h;
  int x0, y0, x1, y1;     /* Range{upper_left, bottom_right}, both inclusive (blender.hh:19-26) */
  double homo_inv[9];     /* ImageComponent::homo_inv (stitcher_image.hh:40-43) */
} pano_blend_image;

typedef struct pano_blend_geom {
  int    projection;             /* pano_projection */
  double res_x, res_y;           /* `resolution` (stitcher_image.cc:119) */
  double proj_min_x, proj_min_y; /* proj_range.min */
} pano_blend_geom;

/* target_size = componentwise max of bottom_right (blender.cc:21, multiband.cc:16). */
int pano_blend_target_size(int n, const pano_blend_image* imgs, int* out_w, int* out_h);
/* bands == 0: LinearBlender::run; bands > 0: MultiBandBlender{bands}::run.
 * out_hwc: out_h×out_w×3 f32, -1 where no image contributes. */
int pano_blend(pano_ctx* ctx, int n, const pano_blend_image* imgs, const pano_blend_geom* g,
               int bands, const pano_params* p, float* out_hwc, int out_w, int out_h);
/* Device-resident variant: imgs[k].rgb_hwc and d_out_hwc are device pointers. */
int pano_blend_dev(pano_ctx* ctx, int n, const pano_blend_image* imgs, const pano_blend_geom* g,
                   int bands, const pano_params* p, float* d_out_hwc, int out_w, int out_h);

/* Rows [row0, row1) of the same mosaic into d_out_rows ((row1-row0)×out_w×3 f32): the
 * strip partition of the canvas across GPUs (every output pixel of
 * LinearBlender::run is independent, blender.cc:37-96, so concatenated strips are
 * bit-identical to pano_blend_dev).  bands > 0 (MultiBandBlender): the strip is computed
 * from every image's ROI clipped to [row0 - H, row1 + H), H = the summed half-widths of
 * the level blurs (6+6+6+9 = 27 rows for 5 bands): a band at level l depends on level 0
 * only within that radius, so concatenated strips are bit-identical as well. */
int pano_blend_rows_dev(pano_ctx* ctx, int n, const pano_blend_image* imgs, const pano_blend_geom* g,
                        int bands, const pano_params* p, float* d_out_rows, int out_w, int out_h,
                        int row0, int row1);

/* --------------------------------------------------------- multi-GPU
 * One process (or host thread) per GPU, one pano_ctx each.  The path shards on
 * independent units — images k mod G for SIFT (stitcherbase.cc:14), the pair list of
 * stitcher.cc:98-112 for matching, canvas rows for the blend (blender.cc:79, multiband.cc:75,127)
 * — and has two exchange steps, both NCCL over NVLink on the context's stream:
 *   C1  pano_comm_allgather_features   the descriptor sets every pair task needs
 *   C2  pano_comm_allgather_dev        the row strips of pano_blend_rows_dev -> the mosaic
 * libnccl.so.2 is loaded at run time by the first pano_comm_* call; single-GPU hosts never need it. */
typedef struct pano_comm pano_comm;
/* ncclGetUniqueId: rank 0 calls this and hands the 128 bytes to every rank (file, socket, MPI...). */
int  pano_comm_unique_id(unsigned char id[128]);
/* ncclCommInitRank on ctx's device; collective over all `world` ranks. */
int  pano_comm_create(pano_ctx* ctx, int world, int rank, const unsigned char id[128], pano_comm** out);
/* Wraps a communicator the host already has (ncclComm_t cast to void*); not destroyed by pano_comm_destroy. */
int  pano_comm_adopt(pano_ctx* ctx, void* nccl_comm, int world, int rank, pano_comm** out);
void pano_comm_destroy(pano_comm* c);
int  pano_comm_world(const pano_comm* c);
int  pano_comm_rank(const pano_comm* c);
/* C1: `local` = this rank's images (k = rank, rank + world, ... in ascending k; NULL if it owns none)
 * of n_images_total; *all receives a featureset of all n_images_total images on every rank,
 * bit-identical to detecting them on one GPU.  Device to device, no host staging of descriptors. */
int  pano_comm_allgather_features(pano_comm* c, pano_featureset* local, int n_images_total, pano_featureset** all);
/* C2 (and any other equal-sized exchange): bytes_per_rank from d_send of every rank, concatenated by
 * rank into d_recv (world * bytes_per_rank). */
int  pano_comm_allgather_dev(pano_comm* c, const void* d_send, void* d_recv, size_t bytes_per_rank);

/* ------------------------------------------------ 8-bit image boundary
 * The byte formats either side of the path: decoded 8-bit
 * pixels in, 8-bit mosaic out, so 3 B/px cross PCIe instead of 12.  All
 * pointers are device pointers; work is queued on the context's stream. */
/* read_img's conversion loop (lib/imgio.cc:75-88): channels == 3 -> every
 * sample is (float)((double)v / 255.0); channels == 1 -> the grey value is
 * replicated to R,G,B WITHOUT the division (imgio.cc:84-87).  d_pix is
 * h×w×channels interleaved u8; d_out_hwc is h×w×3 f32. */
int pano_rgb8_to_mat32f_dev(pano_ctx* ctx, const unsigned char* d_pix, int w, int h, int channels,
                            float* d_out_hwc);
/* The same for n images in one launch (the calc_feature loop reads every image,
 * stitcherbase.cc:14-17).  d_pix[i] must be 4-byte, d_out_hwc[i] 16-byte aligned;
 * the pointer arrays themselves are host arrays of device pointers. */
int pano_rgb8_to_mat32f_batch_dev(pano_ctx* ctx, int n, const unsigned char* const* d_pix, const int* w,
                                  const int* h, const int* channels, float* const* d_out_hwc);
/* crop()'s rectangle (lib/imgproc.cc:200-235): the largest axis-aligned
 * rectangle of pixels whose max(r,g,b) >= 0, first maximum in (line, column)
 * order.  d_rect receives {x0, y0, width, height} (device int[4]). */
int pano_crop_rect_dev(pano_ctx* ctx, const float* d_mat_hwc, int w, int h, int* d_rect);
/* write_rgb's conversion loop (lib/imgio.cc:98-113) applied to the sub-rectangle
 * d_rect = {x0,y0,cw,ch} (device int[4]; NULL = whole image): every sample is
 * (unsigned char)((v < 0 ? 1 : v) * 255), i.e. Color::NO turns white.  Output
 * is packed ch×cw×3 u8 at d_out (capacity h*w*3). */
int pano_mat32f_to_rgb8_dev(pano_ctx* ctx, const float* d_mat_hwc, int w, int h, const int* d_rect,
                            unsigned char* d_out);

/* ------------------------------------------------------- device utilities */
int pano_dev_alloc(pano_ctx* ctx, size_t bytes, void** d_ptr);
int pano_dev_free(pano_ctx* ctx, void* d_ptr);
int pano_dev_upload(pano_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);
int pano_dev_download(pano_ctx* ctx, void* h_dst, const void* d_src, size_t bytes);
/* Stream-ordered variants: the host buffer must be pinned and stay valid until
 * pano_sync(); used by the end-to-end path to overlap copies with kernels. */
int pano_dev_upload_async(pano_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);
int pano_dev_download_async(pano_ctx* ctx, void* h_dst, const void* d_src, size_t bytes);

/* Page-locked host memory (cudaHostAlloc) for buffers handed to the async copies
 * and to pano_sift_detect_batch. */
int pano_host_alloc(size_t bytes, void** h_ptr);
int pano_host_free(void* h_ptr);

/* Cross-context ordering.  Several contexts on one device (e.g. an upload ctx, a
 * compute ctx and a download ctx, each with its own stream) can overlap copies
 * with kernels: an event recorded on one ctx's stream can be waited for by
 * another ctx's stream (device-side) or by the host. */
typedef struct pano_event pano_event;
int  pano_event_create(pano_ctx* ctx, pano_event** out);
int  pano_event_record(pano_ctx* ctx, pano_event* ev);        /* on ctx's stream */
int  pano_event_wait(pano_ctx* ctx, pano_event* ev);          /* ctx's stream waits (no host block) */
int  pano_event_sync(pano_event* ev);                         /* host blocks until the event completed */
void pano_event_destroy(pano_event* ev);

/* ------------------------------------------------------ stage inspection
 * Parity-test hooks: run the SIFT chain on ONE host image and keep every
 * intermediate on the device so tests can compare each stage with the oracle
 * (SURVEY.md §4 "per-stage oracle tests"). */
typedef struct pano_sift_trace pano_sift_trace;
int  pano_sift_trace_run(pano_ctx* ctx, const float* rgb_hwc, int w, int h,
                         const pano_params* p, pano_sift_trace** out);
int  pano_sift_trace_working_size(const pano_sift_trace* t, int* w0, int* h0);
int  pano_sift_trace_octave_size(const pano_sift_trace* t, int octave, int* w, int* h);
/* kind: 0 working RGB (3ch, octave ignored), 1 gaussian level i∈[0,nscale),
 * 2 |DoG| level i∈[0,nscale-1), 3 mag level i∈[1,nscale), 4 ort level. */
int  pano_sift_trace_plane(pano_sift_trace* t, int kind, int octave, int level, float* out);
/* stage: 0 raw extrema (x,y,pyr_id,scale_id valid), 1 refined+edge-tested
 * keypoints, 2 oriented keypoints.  Returns count; copies min(count,cap). */
int  pano_sift_trace_points(pano_sift_trace* t, int stage, int cap, pano_sspoint* out);
int  pano_sift_trace_descriptors(pano_sift_trace* t, int cap, double* coor_xy, float* desc);
void pano_sift_trace_free(pano_sift_trace* t);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* PANO_B200_H */
