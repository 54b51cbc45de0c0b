// Stand-alone timing probe of the fused sequence mixer (seq_mixer.cuh): random data, knock-outs of single stages.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DW2S_SEQ_PROBE -o seq_probe seq_probe.cu && ./seq_probe
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../wav2sleep_b200/csrc/seq_mixer.cuh"
using namespace w2s;

int main(int argc, char** argv) {
  const int S = argc > 1 ? atoi(argv[1]) : 1200;
  const int n_blocks = 2, n_dil = 6, L = n_blocks * n_dil;
  const int nc_force = argc > 2 ? atoi(argv[2]) : 0;
  const int Bmax = 32;
  __half *w, *x, *buf;
  float *ln, *head, *logits;
  cudaMalloc(&w, (size_t)L * 7 * 128 * 128 * 2);
  cudaMalloc(&x, (size_t)Bmax * S * 128 * 2);
  cudaMalloc(&buf, seq_fused_workspace_bytes(Bmax, S, n_dil));
  cudaMalloc(&ln, 128 * 4);
  cudaMalloc(&head, 8 * 128 * 4);
  cudaMalloc(&logits, (size_t)Bmax * S * 8 * 4);
  std::vector<__half> hw((size_t)L * 7 * 128 * 128), hx((size_t)Bmax * S * 128);
  srand(1);
  for (auto& v : hw) v = __float2half((rand() / (float)RAND_MAX - 0.5f) * 0.1f);
  for (auto& v : hx) v = __float2half((rand() / (float)RAND_MAX - 0.5f) * 2.0f);
  cudaMemcpy(w, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(x, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
  std::vector<float> ones(128, 1.0f);
  cudaMemcpy(ln, ones.data(), 512, cudaMemcpyHostToDevice);
  cudaMemset(head, 0, 8 * 128 * 4);
  SeqArgs a;
  memset(&a, 0, sizeof(a));
  for (int l = 0; l < L; ++l) {
    a.w[l] = w + (size_t)l * 7 * 128 * 128;
    a.ln_w[l] = ln;
    a.ln_b[l] = ln;
  }
  a.x = x; a.buf = buf; a.head_w = head; a.head_b = head; a.logits = logits; a.n_classes = 4;
  a.n_blocks = n_blocks; a.n_dil = n_dil; a.S = S; a.ln_eps = 1e-5f;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int dbgs[] = {0, 4, 7, 256, 4 + 256, 4 + 256 + 224};
  for (int B : {16}) {
    SeqGeom g;
    if (!seq_fused_geometry(B, S, n_dil, g, nc_force)) return 1;
    printf("B=%d S=%d nc=%d rpc=%d ntiles=%d AR=%d stages=%d smem=%zu active clusters=%d\n", B, S, g.nc, g.rpc, g.ntiles, g.AR,
           g.n_stages, g.smem, seq_max_active_clusters(g.nc, g.smem));
    for (int dbg : dbgs) {
      a.dbg = dbg;
      for (int i = 0; i < 3; ++i) launch_seq_mixer(a, g, B, 0);
      cudaEventRecord(e0);
      const int reps = 10;
      for (int i = 0; i < reps; ++i) launch_seq_mixer(a, g, B, 0);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      printf("B=%2d dbg=%2d  %8.1f us per launch  (%s)\n", B, dbg, ms * 1000 / reps, cudaGetErrorString(e));
      if (e != cudaSuccess) return 2;
    }
  }
  return 0;
}
