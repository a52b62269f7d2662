// sched_test.cpp — C++ test of the multi-GPU scheduler entry points (include/sfmgpu.h: sfmgpu_sched_*), no Python, no torch.
// One host thread per rank (= per GPU; NCCL refuses two ranks on one device, so world = min(requested, visible GPUs)).
// Every rank generates ITS frames of one synthetic sequence (pair block + halo frame), runs the two-view unit with the
// RANSAC stage on its block and takes part in the gather; the root compares the gathered arrays with a single-GPU run
// of the whole sequence, byte for byte.  usage: sched_test [world=2] [frames=13] [w=320] [h=240] [corners=300]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/sfmgpu.h"

#define CK(ctx, expr)                                                                         \
  do {                                                                                        \
    int rc__ = (expr);                                                                        \
    if (rc__ != 0) {                                                                          \
      fprintf(stderr, "FAILED %s -> %d: %s\n", #expr, rc__, sfmgpu_last_error(ctx));          \
      return 1;                                                                               \
    }                                                                                         \
  } while (0)

struct Results {
  std::vector<double> li, lj, R, t;
  std::vector<int32_t> nk, nc, st, bn, inl;
  void size(int P, int cap) {
    li.assign((size_t)P * cap * 2, -1.0);
    lj.assign((size_t)P * cap * 2, -1.0);
    R.assign((size_t)P * 9, -1.0);
    t.assign((size_t)P * 3, -1.0);
    nk.assign(P, -1);
    nc.assign(P, -1);
    st.assign(P, -1);
    bn.assign(P, -1);
    inl.assign((size_t)P * cap, -1);
  }
};

static const double K[9] = {1520.4, 0, 302.32, 0, 1525.9, 246.87, 0, 0, 1.0};
static const uint32_t SEED = 20261018u;

static int run_rank(int world, int rank, int device, const void* id, int frames, int w, int h, int cap, Results* out) {
  sfmgpu_ctx* ctx = nullptr;
  if (sfmgpu_create(device, &ctx) != 0) {
    fprintf(stderr, "rank %d: no CUDA device %d\n", rank, device);
    return 1;
  }
  // KLT kernel selection depends on the batch size (lane-per-feature from 6000 features on, warp-per-feature below); the
  // kernels agree to ~1e-13 px, not bit for bit, so a byte comparison between shardings pins ONE kernel family
  CK(ctx, sfmgpu_klt_set_mode(ctx, 2));
  sfmgpu_sched* s = nullptr;
  CK(ctx, sfmgpu_sched_create(ctx, world, rank, id, nullptr, &s));
  int p0, p1, f0, f1;
  CK(ctx, sfmgpu_sched_pair_shard(s, frames, &p0, &p1, &f0, &f1));
  const int np = p1 - p0, nf = f1 - f0;
  sfmgpu_lkcfg cfg;
  sfmgpu_lkcfg_default(&cfg);
  cfg.max_tracks = cap;
  sfmgpu_ransac_cfg rc = {200, 2e-3, 80, 120};
  sfmgpu_frames* fr = nullptr;
  sfmgpu_pairs* pr = nullptr;
  CK(ctx, sfmgpu_frames_create(ctx, w, h, nf > 2 ? nf : 2, cfg.pyr_levels, &fr));
  CK(ctx, sfmgpu_pairs_create(ctx, np > 1 ? np : 1, cap, &pr));
  CK(ctx, sfmgpu_pairs_set_ransac(ctx, pr, K, &rc));
  if (nf > 0) {
    CK(ctx, sfmgpu_frames_synth(ctx, fr, 0, nf, SEED, f0));  // this rank's frames of THE sequence
    CK(ctx, sfmgpu_pyramid_build(ctx, fr, 0, nf));
  }
  CK(ctx, sfmgpu_pair_frontend(ctx, fr, 0, np, &cfg, pr));
  const int P = frames - 1;
  if (rank == 0) out->size(P, cap);
  CK(ctx, sfmgpu_sched_gather_pairs(ctx, s, pr, P, 0, 1, rank == 0 ? out->li.data() : nullptr, rank == 0 ? out->lj.data() : nullptr,
                                    rank == 0 ? out->nk.data() : nullptr, rank == 0 ? out->nc.data() : nullptr,
                                    rank == 0 ? out->st.data() : nullptr, rank == 0 ? out->bn.data() : nullptr,
                                    rank == 0 ? out->inl.data() : nullptr, rank == 0 ? out->R.data() : nullptr,
                                    rank == 0 ? out->t.data() : nullptr));
  sfmgpu_sched_destroy(ctx, s);
  sfmgpu_pairs_destroy(ctx, pr);
  sfmgpu_frames_destroy(ctx, fr);
  sfmgpu_destroy(ctx);
  return 0;
}

int main(int argc, char** argv) {
  int want = argc > 1 ? atoi(argv[1]) : 2;
  const int frames = argc > 2 ? atoi(argv[2]) : 13, w = argc > 3 ? atoi(argv[3]) : 320, h = argc > 4 ? atoi(argv[4]) : 240,
            cap = argc > 5 ? atoi(argv[5]) : 300;
  // how many devices are there?  sfmgpu_create fails beyond the last one
  int ndev = 0;
  for (; ndev < 16; ndev++) {
    sfmgpu_ctx* c = nullptr;
    if (sfmgpu_create(ndev, &c) != 0) break;
    sfmgpu_destroy(c);
  }
  if (ndev == 0) {
    fprintf(stderr, "no CUDA device: the library has no CPU fallback\n");
    return 2;
  }
  const int world = want < ndev ? want : ndev;
  // shard arithmetic (pure): blocks are contiguous, cover [0, n), sizes differ by at most one
  for (int n : {0, 1, 7, 1999})
    for (int W : {1, 2, 3, 8}) {
      int prev = 0;
      for (int r = 0; r < W; r++) {
        int a, b;
        if (sfmgpu_sched_shard(n, W, r, &a, &b) != 0 || a != prev || b < a || b - a > n / W + 1) {
          fprintf(stderr, "shard(%d, %d, %d) wrong\n", n, W, r);
          return 1;
        }
        prev = b;
      }
      if (prev != n) return 1;
    }
  unsigned char id[SFMGPU_SCHED_ID_BYTES] = {0};
  if (world > 1 && sfmgpu_sched_unique_id(id) != 0) {
    fprintf(stderr, "sfmgpu_sched_unique_id failed (NCCL library not found?)\n");
    return 1;
  }
  Results got, ref;
  std::vector<int> rcs(world, 0);
  std::vector<std::thread> pool;
  for (int r = 0; r < world; r++) pool.emplace_back([&, r]() { rcs[r] = run_rank(world, r, r, id, frames, w, h, cap, &got); });
  for (auto& th : pool) th.join();
  for (int r = 0; r < world; r++)
    if (rcs[r] != 0) return 1;
  // the whole sequence on ONE GPU through the same entry points (world = 1: no communicator)
  if (run_rank(1, 0, 0, nullptr, frames, w, h, cap, &ref) != 0) return 1;
  const int P = frames - 1;
  bool ok = got.nk == ref.nk && got.nc == ref.nc && got.st == ref.st && got.bn == ref.bn;
  long tracks = 0, poses = 0;
  for (int p = 0; p < P && ok; p++) {
    const size_t o = (size_t)p * cap;
    tracks += ref.nc[p];
    ok = ok && memcmp(&got.li[2 * o], &ref.li[2 * o], (size_t)ref.nk[p] * 16) == 0 && memcmp(&got.lj[2 * o], &ref.lj[2 * o], (size_t)ref.nk[p] * 16) == 0;
    if (ref.st[p] == 2) {
      poses++;
      ok = ok && memcmp(&got.inl[o], &ref.inl[o], (size_t)ref.bn[p] * 4) == 0 && memcmp(&got.R[9 * p], &ref.R[9 * p], 72) == 0 &&
           memcmp(&got.t[3 * p], &ref.t[3 * p], 24) == 0;
    }
  }
  printf("sched_test: world %d (of %d GPUs), %d pairs, %ld feature-tracks, %ld poses: %s\n", world, ndev, P, tracks, poses,
         ok ? "gathered == single-GPU, OK" : "MISMATCH");
  return ok ? 0 : 1;
}
