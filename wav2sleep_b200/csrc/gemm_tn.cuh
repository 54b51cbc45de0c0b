// Weight-gradient GEMM of the training path:  C[M, N] += sum over rows r of X[r, :]^T (x) Y[map(r), :]
//
// X = output-side gradient rows [B, LX, M] (fp16, channels-last), Y = input-side activation rows [B, LY, N];
// map(r) = l * y_stride + y_offset inside the same sample (rows outside [0, LY) contribute zero: conv padding).
// This is dW of every Linear (y_stride 1, y_offset 0) and of every conv tap (reference autograd of
// nn.Conv1d / nn.Linear: blocks.py:154-163, wav2sleep.py:230,286-296,41): the reduction runs over positions
// (up to 2e7 rows) while M, N <= 128, so it is a streaming, HBM-bound kernel.  Warp-level mma.sync.m16n8k16 with
// both operands transposed on load (ldmatrix.trans from row-major [row][channel] tiles); each CTA reduces a
// contiguous row range into registers and finishes with fp32 atomics into C (stride-addressed, so C can be a
// tap slice of a [Cout, Cin, taps] conv weight gradient).
//
// Measured and rejected (round 2, same-box A/B on the 4-signal and the ECG-only training step): (i) forcing 2 or 3
// resident CTAs per SM through __launch_bounds__ (spills: 37.0 -> 37.9 / 44.4 ms per step); (ii) replacing the register
// double buffering by a 3-5 stage cp.async ring with one CTA per SM (equal step time; 16-channel layers 20 % slower,
// 128-channel layers 25 % faster in isolation).  Inside the step these kernels share the GPU with the other encoders'
// streams, so their isolated rate is not what bounds the step.
#pragma once
#include "common.cuh"

namespace w2s {

struct GemmTNArgs {
  const act_t* X;   // [B, LX, M]
  const act_t* Y;   // [B, LY, N]
  float* C;         // element (m, n) of tap t at C[m * ldc_m + n * ldc_n + t * ldc_t]
  const uint8_t* row_mask;  // [B] or null
  int B, LX, LY;
  int y_stride, y_offset;   // Y row of X row l, tap t:  l * y_stride + y_offset + t
  int grid_taps, tap_stride;  // TAPS == 1 kernels: blockIdx.y = tap, Y offset += tap * tap_stride, C += tap * ldc_t
  long long ldc_m, ldc_n, ldc_t;
  float scale;      // multiplies the contribution (1.0)
};

constexpr int kGemmTNThreads = 256;
// X rows per smem tile (tiles never straddle two samples): long tiles for narrow operands, so that every staging
// round moves >= 16 KB per CTA and the per-tile synchronisation is amortised.
constexpr int gemm_tn_rows(int m, int n) { return (m + n) <= 64 ? 512 : 128; }

template <int M, int N, int TAPS = 1>
struct GemmTNCfg {
  // warp tile; fused-tap kernels keep TAPS accumulator sets, so their warp tile is halved to stay within registers
  static constexpr int MW = (TAPS > 1 && M > 32) ? 32 : (M < 64 ? M : 64);
  static constexpr int NW = N < 32 ? N : 32;
  static constexpr int WARPS_MN = (M / MW) * (N / NW);
  static constexpr int KG = 8 / WARPS_MN;          // warps that split the rows of a tile
  static constexpr int LDX = M + 8, LDY = N + 8;   // padded smem row strides (halfs)
  static constexpr int MT = MW / 16, NT = NW / 8;
  static constexpr int ROWS = gemm_tn_rows(M, N);
  static_assert(WARPS_MN <= 8 && 8 % WARPS_MN == 0, "warp layout");
  static_assert(ROWS % (16 * KG) == 0, "row tile vs k-groups");
};

W2S_DEVINL void ldmatrix_x4_trans(uint32_t (&r)[4], const __half* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
W2S_DEVINL void ldmatrix_x2_trans(uint32_t (&r)[2], const __half* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];"
               : "=r"(r[0]), "=r"(r[1])
               : "r"(smem_u32(p)));
}
W2S_DEVINL void mma_16816_f16(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// TAPS > 1: the taps of a convolution (consecutive Y rows, y_stride must be 1) are reduced in one pass over X and a
// haloed Y tile: X and Y are read once instead of TAPS times.
template <int M, int N, int TAPS>
__global__ void __launch_bounds__(kGemmTNThreads) gemm_tn_kernel(const GemmTNArgs p) {
  using Cfg = GemmTNCfg<M, N, TAPS>;
  constexpr int LDX = Cfg::LDX, LDY = Cfg::LDY, MT = Cfg::MT, NT = Cfg::NT, KG = Cfg::KG;
  constexpr int kGemmTNRows = Cfg::ROWS;
  // Staged Y rows per tile: with taps the haloed run of consecutive rows (y_stride == 1); without, exactly the rows
  // the X rows map to (row k of the tile <-> Y row ybase + k * y_stride), stored compactly.
  constexpr int YR = TAPS == 1 ? kGemmTNRows : kGemmTNRows + TAPS - 1;
  constexpr int XN = kGemmTNRows * (M / 8) / kGemmTNThreads;
  constexpr int YN = (YR * (N / 8) + kGemmTNThreads - 1) / kGemmTNThreads;
  static_assert(kGemmTNRows * (M / 8) % kGemmTNThreads == 0, "X staging");
  extern __shared__ __align__(16) uint8_t gemm_tn_smem[];
  __half* sX = reinterpret_cast<__half*>(gemm_tn_smem);
  __half* sY = sX + kGemmTNRows * LDX;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wmn = warp % Cfg::WARPS_MN, kg = warp / Cfg::WARPS_MN;
  const int m0 = (wmn / (N / Cfg::NW)) * Cfg::MW;
  const int n0 = (wmn % (N / Cfg::NW)) * Cfg::NW;
  const int ystep = TAPS == 1 ? p.y_stride : 1;

  float acc[TAPS][MT][NT][4];
#pragma unroll
  for (int t = 0; t < TAPS; ++t)
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[t][i][j][q] = 0.0f;

  const int tiles_per_sample = (p.LX + kGemmTNRows - 1) / kGemmTNRows;
  // masked samples are compacted away before the tiles are split over the CTAs (balanced work under the modality
  // masker); batches larger than the list fall back to skipping tile by tile
  constexpr int kMaxLive = 512;
  __shared__ uint16_t sLive[kMaxLive];
  __shared__ int sNLive;
  const bool compact = p.row_mask != nullptr && p.B <= kMaxLive;
  if (compact) {
    if (warp == 0) {
      const int n = build_live_list(p.row_mask, p.B, sLive, lane);
      if (lane == 0) sNLive = n;
    }
    __syncthreads();
  }
  const long long tiles = (long long)tiles_per_sample * (compact ? sNLive : p.B);
  const long long t_begin = tiles * blockIdx.x / gridDim.x, t_end = tiles * (blockIdx.x + 1) / gridDim.x;
  auto sample_of = [&](long long t) { return compact ? (int)sLive[t / tiles_per_sample] : (int)(t / tiles_per_sample); };
  if (t_begin >= t_end) return;  // no work (e.g. every sample masked): also no atomics into C

  // next tile of this CTA whose sample is not masked (uniform over the CTA)
  auto next_tile = [&](long long t) {
    while (!compact && t < t_end && p.row_mask != nullptr && p.row_mask[(int)(t / tiles_per_sample)]) ++t;
    return t;
  };
  // Register double buffering: the global loads of tile i+1 are issued before the MMAs of tile i and land while they
  // run; they are written to shared memory after the MMAs are done.
  uint4 xv[XN], yv[YN];
  auto load_tile = [&](long long tile) {
    const int b = sample_of(tile);
    const int l0 = (int)(tile % tiles_per_sample) * kGemmTNRows;
    const act_t* Xb = p.X + ((size_t)b * p.LX) * M;
    const act_t* Yb = p.Y + ((size_t)b * p.LY) * N;
    const int ybase = l0 * p.y_stride + p.y_offset + (int)blockIdx.y * p.tap_stride;
#pragma unroll
    for (int i = 0; i < XN; ++i) {
      const int id = tid + i * kGemmTNThreads;
      const int k = id / (M / 8), c = id % (M / 8);
      xv[i] = make_uint4(0u, 0u, 0u, 0u);
      if (l0 + k < p.LX) xv[i] = __ldg(reinterpret_cast<const uint4*>(Xb + (size_t)(l0 + k) * M) + c);
    }
#pragma unroll
    for (int i = 0; i < YN; ++i) {
      const int id = tid + i * kGemmTNThreads;
      const int k = id / (N / 8), c = id % (N / 8);
      const long long ly = (long long)ybase + (long long)k * ystep;
      yv[i] = make_uint4(0u, 0u, 0u, 0u);
      if (id < YR * (N / 8) && ly >= 0 && ly < p.LY && (TAPS > 1 || l0 + k < p.LX))
        yv[i] = __ldg(reinterpret_cast<const uint4*>(Yb + (size_t)ly * N) + c);
    }
  };
  auto store_tile = [&]() {
#pragma unroll
    for (int i = 0; i < XN; ++i) {
      const int id = tid + i * kGemmTNThreads;
      *reinterpret_cast<uint4*>(sX + (id / (M / 8)) * LDX + (id % (M / 8)) * 8) = xv[i];
    }
#pragma unroll
    for (int i = 0; i < YN; ++i) {
      const int id = tid + i * kGemmTNThreads;
      if (id < YR * (N / 8)) *reinterpret_cast<uint4*>(sY + (id / (N / 8)) * LDY + (id % (N / 8)) * 8) = yv[i];
    }
  };

  long long tile = next_tile(t_begin);
  if (tile < t_end) load_tile(tile);
  while (tile < t_end) {
    __syncthreads();  // the MMAs of the previous tile are done with the staging buffers
    store_tile();
    __syncthreads();
    tile = next_tile(tile + 1);
    if (tile < t_end) load_tile(tile);
    constexpr int KSTEPS = kGemmTNRows / 16 / KG;
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
      const int k0 = (kg * KSTEPS + ks) * 16;
      uint32_t a[MT][4];
#pragma unroll
      for (int i = 0; i < MT; ++i)
        ldmatrix_x4_trans(a[i], sX + (k0 + (lane >> 4) * 8 + (lane & 7)) * LDX + m0 + i * 16 + ((lane >> 3) & 1) * 8);
#pragma unroll
      for (int t = 0; t < TAPS; ++t) {
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          uint32_t bf[2];
          const int yrow = k0 + ((lane >> 3) & 1) * 8 + (lane & 7) + t;
          ldmatrix_x2_trans(bf, sY + yrow * LDY + n0 + j * 8);
#pragma unroll
          for (int i = 0; i < MT; ++i) mma_16816_f16(acc[t][i][j], a[i], bf);
        }
      }
    }
  }
  // ---- accumulate into C ----
  // k-group partials are first combined in shared memory (the staging buffers are free now), so that C sees one
  // atomic per element and CTA instead of one per warp: C is tiny (<= 48 KB) and contended by every CTA.
  if (KG > 1) {
    float* sC = reinterpret_cast<float*>(gemm_tn_smem);
    __syncthreads();
    for (int i = tid; i < M * N * TAPS; i += kGemmTNThreads) sC[i] = 0.0f;
    __syncthreads();
#pragma unroll
    for (int t = 0; t < TAPS; ++t)
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const int m = m0 + i * 16 + (lane >> 2);
          const int n = n0 + j * 8 + (lane & 3) * 2;
          float* c0 = sC + (t * M + m) * N + n;
          atomicAdd(c0, acc[t][i][j][0]);
          atomicAdd(c0 + 1, acc[t][i][j][1]);
          atomicAdd(c0 + 8 * N, acc[t][i][j][2]);
          atomicAdd(c0 + 8 * N + 1, acc[t][i][j][3]);
        }
    __syncthreads();
    for (int i = tid; i < M * N * TAPS; i += kGemmTNThreads) {
      const int t = i / (M * N), m = (i / N) % M, n = i % N;
      atomicAdd(p.C + m * p.ldc_m + n * p.ldc_n + (t + (int)blockIdx.y) * p.ldc_t, sC[i] * p.scale);
    }
  } else {
#pragma unroll
    for (int t = 0; t < TAPS; ++t)
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const int m = m0 + i * 16 + (lane >> 2);
          const int n = n0 + j * 8 + (lane & 3) * 2;
          float* c0 = p.C + m * p.ldc_m + n * p.ldc_n + (t + (int)blockIdx.y) * p.ldc_t;
          atomicAdd(c0, acc[t][i][j][0] * p.scale);
          atomicAdd(c0 + p.ldc_n, acc[t][i][j][1] * p.scale);
          atomicAdd(c0 + 8 * p.ldc_m, acc[t][i][j][2] * p.scale);
          atomicAdd(c0 + 8 * p.ldc_m + p.ldc_n, acc[t][i][j][3] * p.scale);
        }
  }
}

template <int M, int N, int TAPS>
inline cudaError_t launch_gemm_tn(const GemmTNArgs& a, int sm_count, cudaStream_t stream) {
  using Cfg = GemmTNCfg<M, N, TAPS>;
  constexpr int kGemmTNRows = Cfg::ROWS;
  const long long tiles = (long long)((a.LX + kGemmTNRows - 1) / kGemmTNRows) * a.B;
  // every CTA ends with M*N*TAPS atomics into the same small C: give each CTA enough tiles to amortise them
  constexpr int kMinTiles = (M * N >= 64 * 64) ? 8 : 2;
  long long grid = 4LL * sm_count;
  if (grid > tiles / kMinTiles) grid = tiles / kMinTiles;
  if (grid < 1) grid = 1;
  constexpr int y_rows = TAPS == 1 ? kGemmTNRows : kGemmTNRows + TAPS - 1;
  int smem = (kGemmTNRows * Cfg::LDX + y_rows * Cfg::LDY) * 2;
  if (Cfg::KG > 1 && smem < M * N * TAPS * 4) smem = M * N * TAPS * 4;
  static int configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tn_kernel<M, N, TAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  gemm_tn_kernel<M, N, TAPS><<<dim3((unsigned)grid, TAPS == 1 && a.grid_taps > 1 ? a.grid_taps : 1), kGemmTNThreads, smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace w2s
