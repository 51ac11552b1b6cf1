// match_tc.cuh — interface between the tensor-core nomination pass (match_tc.cu)
// and the exact decision logic (match.cu).
#pragma once
#include "common.cuh"
#include <vector>

struct TcImage {
  long long row0;   // first descriptor row of this image in the featureset buffer
  int n;            // real rows
  int n_pad;        // rows padded to a multiple of 256
  int blk0;         // first 128-row block of this image in the operand buffers
};

struct TcTask {       // one CTA: 128 query rows against every target row
  int q_blk;          // operand block holding the query rows
  int q_row0, q_n;    // first query row (within its image) of this block, rows in the image
  int t_blk0, t_blocks;  // target image: first block, number of blocks (even)
  long long res_off;  // where the query image's results start
};

struct TcTop2 { float m1, m2; int idx; int pad; };   // approximate best / second-best d^2 and argmin

struct TcOperands {
  TcImage* d_imgs = nullptr;
  float* d_norms = nullptr;        // |x|^2 per descriptor row (f32)
  unsigned* d_maxnorm = nullptr;   // bits of max |x|^2
  unsigned char* qbuf = nullptr;   // query-form fp16 blocks
  unsigned char* tbuf = nullptr;   // target-form fp16 blocks
};

int  tc_prepare(pano_ctx* ctx, const float* d_desc, const std::vector<TcImage>& imgs, TcOperands* ops);
void tc_release(pano_ctx* ctx, TcOperands* ops);
int  tc_run_top2(pano_ctx* ctx, const TcOperands* ops, const TcTask* d_tasks, int n_tasks, TcTop2* d_res);
