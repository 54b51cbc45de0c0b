// Stand-alone probe for the tcgen05 conventions conv_igemm.cuh relies on (no torch, no library).
//   usage: umma_probe <variant> [rowshift]
//     variant 0: LBO = chunk stride (K direction), SBO = 128 (8-row group)   <- what the kernels use
//     variant 1: LBO / SBO swapped
//   rowshift: start the A operand `rowshift` rows into the staged tile (conv tap addressing)
// Computes D[128 x N] = A[128 x K] * B[N x K]^T with K = 32, N = 32 from chunk-major smem and checks against the CPU.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../wav2sleep_b200/csrc/common.cuh"

using namespace w2s;

constexpr int M = 128, N = 32, K = 32, ROWS = 160;  // ROWS staged rows (>= M + shift)

__global__ void probe_kernel(const __half* A /*[ROWS][K]*/, const __half* B /*[N][K]*/, float* D /*[M][N]*/, int variant,
                             int rowshift) {
  __shared__ __align__(128) uint8_t sA[(K / 8) * ROWS * 16];
  __shared__ __align__(128) uint8_t sB[(K / 8) * N * 16];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tmem_alloc(&tmem_slot, 32);
  if (tid == 32) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  for (int id = tid; id < ROWS * (K / 8); id += blockDim.x) {
    const int row = id / (K / 8), c = id % (K / 8);
    *reinterpret_cast<uint4*>(sA + ((size_t)c * ROWS + row) * 16) = *reinterpret_cast<const uint4*>(A + row * K + c * 8);
  }
  for (int id = tid; id < N * (K / 8); id += blockDim.x) {
    const int row = id / (K / 8), c = id % (K / 8);
    *reinterpret_cast<uint4*>(sB + ((size_t)c * N + row) * 16) = *reinterpret_cast<const uint4*>(B + row * K + c * 8);
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_f16(M, N, false);
    for (int kk = 0; kk < K / 16; ++kk) {
      const uint32_t a_addr = smem_u32(sA) + ((2 * kk) * ROWS + rowshift) * 16;
      const uint32_t b_addr = smem_u32(sB) + (2 * kk) * N * 16;
      uint64_t da, db;
      if (variant == 0) {
        da = umma_smem_desc(a_addr, ROWS * 16, 128);
        db = umma_smem_desc(b_addr, N * 16, 128);
      } else {
        da = umma_smem_desc(a_addr, 128, ROWS * 16);
        db = umma_smem_desc(b_addr, 128, N * 16);
      }
      umma_f16(tmem, da, db, idesc, kk > 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  for (int cg = 0; cg < N / 16; ++cg) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + cg * 16, v);
    for (int k = 0; k < 16; ++k) D[(warp * 32 + lane) * N + cg * 16 + k] = v[k];
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int rowshift = argc > 2 ? atoi(argv[2]) : 0;
  std::vector<__half> hA(ROWS * K), hB(N * K);
  std::vector<float> fA(ROWS * K), fB(N * K);
  srand(1);
  for (int i = 0; i < ROWS * K; ++i) {
    fA[i] = (float)((rand() % 17) - 8) / 8.0f;
    hA[i] = __float2half(fA[i]);
  }
  for (int i = 0; i < N * K; ++i) {
    fB[i] = (float)((rand() % 13) - 6) / 4.0f;
    hB[i] = __float2half(fB[i]);
  }
  __half *dA, *dB;
  float* dD;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, M * N * 4);
  probe_kernel<<<1, 128>>>(dA, dB, dD, variant, rowshift);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("variant %d shift %d: CUDA error %s\n", variant, rowshift, cudaGetErrorString(e));
    return 2;
  }
  std::vector<float> hD(M * N);
  cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)fA[(m + rowshift) * K + k] * fB[n * K + k];
      const double err = fabs(ref - hD[m * N + n]);
      if (err > maxerr) maxerr = err;
    }
  printf("variant %d shift %d: max abs err %.6f -> %s\n", variant, rowshift, maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH");
  return maxerr < 1e-3 ? 0 : 1;
}
