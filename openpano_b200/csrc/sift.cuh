// sift.cuh — host-visible structures of the batched SIFT pipeline.
#pragma once
#include "common.cuh"
#include "match_tc.cuh"

#define SIFT_MAX_OCT 8
#define SIFT_MAX_LEVELS 8        // nscale-1 blurred levels
#define SIFT_MAX_TAPS 32         // kw <= 31
#define SIFT_CAP_DEFAULT 8192    // raw extrema / descriptors per image a batch starts with; the
                                 // capacity is a RUNTIME value that doubles on overflow (the reference's
                                 // vectors are unbounded, extrema.cc:56-57): see featureset_sync_counts
#define SIFT_CAP_MAX (1 << 20)
#define SIFT_MAX_PEAKS 18        // a peak needs two lower neighbours: <= 36/2

struct ImgMeta {
  const float* src;   // input RGB (device)
  int in_w, in_h;
  int w0, h0;         // working size
  float ifx, ify;     // 1/fx (rows), 1/fy (cols) of the working resize
  long long work_off; // working RGB offset in the arena (floats)
};

struct OctMeta {
  int img, oct;
  int w, h;
  int pitch;            // floats per plane row: w rounded up to 32, so rows start on 128-byte lines
                        // (TMA needs 16-byte global strides; warps read / write whole lines)
  float ifx, ify;       // octave resize from the working image (oct > 0)
  long long gauss_off;  // nscale planes: grey + blurred levels
  long long dog_off;    // nscale-1 planes
  long long plane;      // floats per plane = pitch * h
};

struct GaussTable {
  int nlev;
  int rmax;
  int center[SIFT_MAX_LEVELS];
  float taps[SIFT_MAX_LEVELS][SIFT_MAX_TAPS];
};

// All device state of one batched SIFT run.  Kept alive by pano_sift_trace for
// stage inspection; freed right after the run otherwise.
struct SiftWork {
  int n_img = 0, n_oct = 0, n_scale = 0;
  std::vector<ImgMeta> h_img;
  std::vector<OctMeta> h_oct;     // n_img * n_oct
  float* arena = nullptr;
  size_t arena_floats = 0;
  ImgMeta* d_img = nullptr;
  OctMeta* d_oct = nullptr;
  int2* d_tilespan = nullptr;
  TmaDesc* d_maps = nullptr;      // [n_img * n_oct]
  int n_tiles = 0;
  int cap = SIFT_CAP_DEFAULT;     // per-image capacity of every candidate / descriptor list below
  // keypoint state, all [n_img * cap] unless noted
  int* cand_count = nullptr;      // [n_img] + work counters (see sift.cu)
  uint32_t* cand_keys = nullptr;
  uint32_t* sorted_keys = nullptr;
  pano_sspoint* refined = nullptr;  // valid flag in .dir < 0 ? no: see kp_valid
  unsigned char* kp_valid = nullptr;
  int* npeaks = nullptr;
  float* dirs = nullptr;            // [n_img * cap * SIFT_MAX_PEAKS]
  int* n_desc = nullptr;            // [n_img]
  int* n_refined = nullptr;         // [n_img] (for traces)
  int* desc_cand = nullptr;         // [n_img * cap] candidate index of descriptor
  float* desc_dir = nullptr;        // [n_img * cap]
};

struct pano_featureset {
  pano_ctx* ctx = nullptr;
  int n_images = 0;
  float* d_desc = nullptr;    // rows of 128 f32
  double* d_coor = nullptr;   // rows of 2 f64 (may be null for uploaded sets)
  double* d_real = nullptr;   // SIFT sets: unscaled real_coor in [0,1), rows of 2 f64 (second half of d_coor's block)
  int* d_count = nullptr;     // [n_images]
  std::vector<long long> base;  // first row of image i
  std::vector<int> h_count;
  bool counts_on_host = false;
  cudaEvent_t counts_ready = nullptr;   // unused by the SIFT path (kept for uploaded sets)
  unsigned counts_token = 0;            // completion marker of the count read-back (ctx_signal)
  bool counts_pending = false;
  int* h_count_pinned = nullptr;
  size_t h_count_cap = 0;
  TcOperands tc;              // fp16 tensor-core operands of the descriptors (lazy)
  bool tc_ready = false;
  int error = 0;              // sticky failure of the count read-back (returned by every later use)
  // What a capacity overflow needs to run the batch again with larger lists (SIFT sets only):
  // the sources must stay valid until the counts have been read once (pano_b200.h).
  int cap = 0;                          // per-image row capacity of d_desc / d_coor (0: not a SIFT set)
  std::vector<const float*> src;        // device images
  std::vector<int> src_w, src_h;
  pano_params src_params;
  float* owned_block = nullptr;         // staged upload of pano_sift_detect_batch, freed after the count sync
};

int sift_run_batch(pano_ctx* ctx, int n, const float* const* d_src, const int* w, const int* h,
                   const pano_params* p, pano_featureset* fs, SiftWork** keep, int cap);
void sift_work_free(pano_ctx* ctx, SiftWork* wk);
int featureset_sync_counts(pano_featureset* fs);
