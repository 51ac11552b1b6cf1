// comm_test.cc — the multi-GPU exchange steps of the C ABI driven from C++ with no Python:
// two host threads, one pano_ctx + pano_comm each (GPU 0 and GPU 1), NCCL between them.
//   SIFT on the owned images (k mod 2) -> pano_comm_allgather_features (C1) -> the dealt half of
//   the pair list -> pano_blend_rows_dev on half of the canvas -> pano_comm_allgather_dev (C2)
// and everything must equal the same job on GPU 0 alone, bit for bit.
//   comm_test <stack.bin>   (format: tests/adaptor/adaptor_test.cc)
// Built by oracle/Makefile (test infrastructure; needs only include/pano_b200.h + the library).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "pano_b200.h"

struct Item { int x0, y0, x1, y1; double hi[9]; };
struct Job {
  int n, w, h;
  std::vector<std::vector<float>> imgs;
  std::vector<Item> items;
  double res, min_x, min_y;
};

#define CK(ctx, call) do { int _rc = (call); if (_rc != 0) { fprintf(stderr, "%s -> %d: %s\n", #call, _rc, pano_last_error(ctx)); exit(3); } } while (0)

struct Result {
  std::vector<int> counts;
  std::vector<std::vector<float>> desc;
  std::vector<std::vector<int>> matches;      // per pair task: flat (i, j) list
  std::vector<float> mosaic;
  int ow = 0, oh = 0;
};

static std::vector<pano_blend_image> blend_images(const Job& J, const std::vector<float*>& d_imgs) {
  std::vector<pano_blend_image> b(J.n);
  for (int k = 0; k < J.n; ++k) {
    b[k].rgb_hwc = d_imgs[k]; b[k].w = J.w; b[k].h = J.h;
    b[k].x0 = J.items[k].x0; b[k].y0 = J.items[k].y0; b[k].x1 = J.items[k].x1; b[k].y1 = J.items[k].y1;
    memcpy(b[k].homo_inv, J.items[k].hi, sizeof(double) * 9);
  }
  return b;
}

static void run_rank(const Job& J, const std::vector<int>& pairs, int world, int rank, const unsigned char* id, Result* out) {
  pano_ctx* ctx = nullptr;
  if (pano_create(&ctx, rank, nullptr)) { fprintf(stderr, "pano_create(%d): %s\n", rank, pano_last_error(nullptr)); exit(3); }
  pano_comm* comm = nullptr;
  if (world > 1) CK(ctx, pano_comm_create(ctx, world, rank, id, &comm));
  pano_params p;
  pano_params_default(&p);
  // every image on this device (the blend reads a strip of each), SIFT only on the owned ones
  std::vector<float*> d_imgs(J.n);
  const size_t bytes = (size_t)J.w * J.h * 3 * sizeof(float);
  for (int k = 0; k < J.n; ++k) {
    CK(ctx, pano_dev_alloc(ctx, bytes, (void**)&d_imgs[k]));
    CK(ctx, pano_dev_upload(ctx, d_imgs[k], J.imgs[k].data(), bytes));
  }
  std::vector<const float*> own;
  std::vector<int> ws, hs;
  for (int k = rank; k < J.n; k += world) { own.push_back(d_imgs[k]); ws.push_back(J.w); hs.push_back(J.h); }
  pano_featureset *fs_local = nullptr, *fs_all = nullptr;
  if (!own.empty()) CK(ctx, pano_sift_detect_batch_dev(ctx, (int)own.size(), own.data(), ws.data(), hs.data(), &p, &fs_local));
  if (world > 1) { CK(ctx, pano_comm_allgather_features(comm, fs_local, J.n, &fs_all)); pano_featureset_free(fs_local); }
  else fs_all = fs_local;
  out->counts.resize(J.n); out->desc.resize(J.n);
  for (int k = 0; k < J.n; ++k) {
    out->counts[k] = pano_featureset_count(fs_all, k);
    out->desc[k].resize((size_t)out->counts[k] * 128 + 1);
    std::vector<double> xy((size_t)out->counts[k] * 2 + 2);
    CK(ctx, pano_featureset_download(fs_all, k, xy.data(), out->desc[k].data()));
  }
  // pair tasks dealt round-robin
  const int n_pairs = (int)pairs.size() / 2;
  std::vector<int> mine_ij, mine_idx;
  for (int t = rank; t < n_pairs; t += world) { mine_ij.push_back(pairs[2 * t]); mine_ij.push_back(pairs[2 * t + 1]); mine_idx.push_back(t); }
  out->matches.assign(n_pairs, {});
  if (!mine_idx.empty()) {
    pano_matches m;
    CK(ctx, pano_match_pairs(ctx, fs_all, (int)mine_idx.size(), mine_ij.data(), &p, &m));
    for (size_t q = 0; q < mine_idx.size(); ++q)
      out->matches[mine_idx[q]].assign(m.idx + 2 * m.offset[q], m.idx + 2 * (m.offset[q] + m.count[q]));
    pano_matches_free(&m);
  }
  pano_featureset_free(fs_all);
  // blend: rows [rank*rows_per, ...) of the canvas, then C2
  std::vector<pano_blend_image> b = blend_images(J, d_imgs);
  pano_blend_geom g{PANO_PROJ_FLAT, J.res, J.res, J.min_x, J.min_y};
  int ow, oh;
  CK(ctx, pano_blend_target_size(J.n, b.data(), &ow, &oh));
  out->ow = ow; out->oh = oh;
  for (int bands : {0, 3}) {
    const int rows_per = (oh + world - 1) / world;
    const int row0 = std::min(oh, rank * rows_per), row1 = std::min(oh, (rank + 1) * rows_per);
    float *d_strip = nullptr, *d_mosaic = nullptr;
    const size_t strip_bytes = (size_t)rows_per * ow * 3 * sizeof(float);
    CK(ctx, pano_dev_alloc(ctx, strip_bytes, (void**)&d_strip));
    CK(ctx, pano_dev_alloc(ctx, strip_bytes * world, (void**)&d_mosaic));
    CK(ctx, pano_blend_rows_dev(ctx, J.n, b.data(), &g, bands, &p, d_strip, ow, oh, row0, row1));
    if (world > 1) CK(ctx, pano_comm_allgather_dev(comm, d_strip, d_mosaic, strip_bytes));
    std::vector<float> host((size_t)oh * ow * 3);
    CK(ctx, pano_dev_download(ctx, host.data(), world > 1 ? d_mosaic : d_strip, host.size() * sizeof(float)));
    out->mosaic.insert(out->mosaic.end(), host.begin(), host.end());
    pano_dev_free(ctx, d_strip); pano_dev_free(ctx, d_mosaic);
  }
  for (auto d : d_imgs) pano_dev_free(ctx, d);
  pano_comm_destroy(comm);
  pano_destroy(ctx);
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: comm_test stack.bin\n"); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  Job J;
  int hdr[3];
  if (fread(hdr, 4, 3, f) != 3) return 2;
  J.n = hdr[0]; J.w = hdr[1]; J.h = hdr[2];
  J.imgs.resize(J.n); J.items.resize(J.n);
  for (int k = 0; k < J.n; ++k) {
    J.imgs[k].resize((size_t)J.w * J.h * 3);
    if (fread(J.imgs[k].data(), 4, J.imgs[k].size(), f) != J.imgs[k].size()) return 2;
  }
  for (int k = 0; k < J.n; ++k) { if (fread(&J.items[k].x0, 4, 4, f) != 4 || fread(J.items[k].hi, 8, 9, f) != 9) return 2; }
  double geo[3];
  if (fread(geo, 8, 3, f) != 3) return 2;
  J.res = geo[0]; J.min_x = geo[1]; J.min_y = geo[2];
  fclose(f);
  std::vector<int> pairs;
  for (int i = 0; i < J.n; ++i) for (int j = i + 1; j < J.n; ++j) { pairs.push_back(i); pairs.push_back(j); }   // stitcher.cc:98-100

  Result one;
  run_rank(J, pairs, 1, 0, nullptr, &one);
  unsigned char id[128];
  if (pano_comm_unique_id(id)) { fprintf(stderr, "pano_comm_unique_id: %s\n", pano_last_error(nullptr)); return 3; }
  Result r[2];
  std::thread t0(run_rank, std::cref(J), std::cref(pairs), 2, 0, id, &r[0]);
  std::thread t1(run_rank, std::cref(J), std::cref(pairs), 2, 1, id, &r[1]);
  t0.join(); t1.join();

  int fail = 0;
  for (int q = 0; q < 2; ++q) {
    if (r[q].counts != one.counts) { printf("FAIL rank %d: feature counts differ\n", q); ++fail; }
    for (int k = 0; k < J.n; ++k)
      if (r[q].desc[k] != one.desc[k]) { printf("FAIL rank %d: descriptors of image %d differ after C1\n", q, k); ++fail; }
    if (r[q].mosaic != one.mosaic) { printf("FAIL rank %d: gathered mosaics differ from one GPU\n", q); ++fail; }
  }
  size_t nm = 0;
  for (size_t t = 0; t < one.matches.size(); ++t) {
    const std::vector<int>& got = r[t % 2].matches[t];
    if (got != one.matches[t]) { printf("FAIL pair task %zu: %zu vs %zu matches\n", t, got.size() / 2, one.matches[t].size() / 2); ++fail; }
    nm += got.size() / 2;
  }
  long long nfeat = 0;
  for (int c : one.counts) nfeat += c;
  printf("2 ranks: %lld features, %zu matches over %zu pair tasks, mosaic %dx%d (linear + 3 bands)\n", nfeat, nm,
         one.matches.size(), one.ow, one.oh);
  printf(fail ? "COMM TEST FAILED (%d)\n" : "COMM TEST OK\n", fail);
  return fail ? 1 : 0;
}
