// First encoder layer: raw fp32 signal (Cin = 1) -> 16 channels, k=3, pad=1, stride 1, no bias.
// reference: SignalEncoders.forward (models/wav2sleep.py:146-161) for the -inf handling,
//            ConvBlock1D.conv1 / downsample of block 0 (models/blocks.py:39-53).
//
// A 3-tap FIR with 16 outputs per sample is not a tensor-core shape (K = 3); it runs on CUDA cores in fp32
// and is purely HBM-bound: 4 B read, 32 B (+16 B residual branch) written per sample.
// Outputs:  y1   [B, T, 16]   fp16 pre-norm conv1 output
//           r0   [B, T/2, 16] fp16 1x1 stride-2 residual branch  w_ds[c] * x[2j]
//           stats[B, 16, 2]   sum / sum-of-squares of y1 over T (InstanceNorm, utils.py:89-92)
//           row_mask[B]       1 where the sample's signal is missing (x[b,0] is +-inf)
#pragma once
#include "common.cuh"

namespace w2s {

struct FirstConvArgs {
  const float* x;       // [B, T]
  const float* w;       // [16, 3]  (conv1.conv.weight[:, 0, :])
  const float* w_ds;    // [16]     (downsample.weight[:, 0, 0])
  act_t* y1;            // [B, T, 16]
  act_t* r0;            // [B, T/2, 16]
  double* stats;        // [B, 16, 2]  zeroed by caller (fp64 accumulators)
  uint8_t* row_mask;    // [B]
  int T;
};

constexpr int kFirstConvThreads = 256;
constexpr int kFirstConvPerThread = 4;
constexpr int kFirstConvPos = kFirstConvThreads * kFirstConvPerThread;

__global__ void __launch_bounds__(kFirstConvThreads) first_conv_kernel(const FirstConvArgs p) {
  const int b = blockIdx.y;
  const float* xb = p.x + (size_t)b * p.T;
  const bool masked = isinf(__ldg(xb));
  if (blockIdx.x == 0 && threadIdx.x == 0) p.row_mask[b] = masked ? 1 : 0;
  if (masked) return;

  __shared__ float sw[16 * 3 + 16];
  __shared__ float sSum[8][16], sSq[8][16];  // per-warp partials, summed in fixed order
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid < 48) sw[tid] = __ldg(p.w + tid);
  if (tid >= 64 && tid < 80) sw[48 + tid - 64] = __ldg(p.w_ds + tid - 64);
  __syncthreads();

  float acc[16], acc2[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = acc2[c] = 0.0f;

  const int p0 = blockIdx.x * kFirstConvPos;
#pragma unroll
  for (int k = 0; k < kFirstConvPerThread; ++k) {
    const int pos = p0 + k * kFirstConvThreads + tid;
    if (pos < p.T) {
      float xm = pos > 0 ? __ldg(xb + pos - 1) : 0.0f;
      float x0 = __ldg(xb + pos);
      float xp = pos + 1 < p.T ? __ldg(xb + pos + 1) : 0.0f;
      xm = isinf(xm) ? 0.0f : xm;  // wav2sleep.py:151
      x0 = isinf(x0) ? 0.0f : x0;
      xp = isinf(xp) ? 0.0f : xp;
      float v[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        v[c] = fmaf(sw[c * 3 + 2], xp, fmaf(sw[c * 3 + 1], x0, sw[c * 3] * xm));
        acc[c] += v[c];
        acc2[c] = fmaf(v[c], v[c], acc2[c]);
      }
      store_h16(p.y1 + ((size_t)b * p.T + pos) * 16, v);
      if ((pos & 1) == 0) {
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] = sw[48 + c] * x0;
        store_h16(p.r0 + ((size_t)b * (p.T >> 1) + (pos >> 1)) * 16, v);
      }
    }
  }
  butterfly16(acc, lane);
  butterfly16(acc2, lane);
  if ((lane & 1) == 0) {
    const int c = butterfly16_channel(lane);
    sSum[tid >> 5][c] = acc[0];
    sSq[tid >> 5][c] = acc2[0];
  }
  __syncthreads();
  if (tid < 16) {
    float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
    for (int w = 0; w < kFirstConvThreads / 32; ++w) {
      s0 += sSum[w][tid];
      s1 += sSq[w][tid];
    }
    atomicAdd(&p.stats[((size_t)b * 16 + tid) * 2 + 0], (double)s0);
    atomicAdd(&p.stats[((size_t)b * 16 + tid) * 2 + 1], (double)s1);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Block-0 fusion: conv1 (Cin = 1) is never materialised at inference.  Its whole-night InstanceNorm statistics follow
// in closed form from four sums over the raw signal (S0 = sum x, R_k = sum x[p] x[p+k], k = 0..2) and the two edge samples:
//   sum_p y_c   = w0 (S0 - x[T-1]) + w1 S0 + w2 (S0 - x[0])
//   sum_p y_c^2 = w0^2 (R0 - x[T-1]^2) + w1^2 R0 + w2^2 (R0 - x[0]^2) + 2 (w0 w1 + w1 w2) R1 + 2 w0 w2 R2
// with y_c[p] = w0 x[p-1] + w1 x[p] + w2 x[p+1] and zero padding.  One pass over 4 B/sample instead of writing and
// re-reading 32 B/sample.
// ------------------------------------------------------------------------------------------------------------------
struct XStatsArgs {
  const float* x;     // [B, T]
  double* xs;         // [B, 4] S0, R0, R1, R2 (zeroed by caller)
  uint8_t* row_mask;  // [B]
  int T;
};
// (blockIdx.z selects the argument set: two signals of equal length in one launch, w2s_encoder_fwd_pair)
__global__ void __launch_bounds__(256) x_stats_kernel(const XStatsArgs p0, const XStatsArgs p1) {
  const XStatsArgs p = blockIdx.z ? p1 : p0;
  const int b = blockIdx.y;
  const float* xb = p.x + (size_t)b * p.T;
  const bool masked = isinf(__ldg(xb));
  if (blockIdx.x == 0 && threadIdx.x == 0) p.row_mask[b] = masked ? 1 : 0;
  if (masked) return;
  double s0 = 0.0, r0 = 0.0, r1 = 0.0, r2 = 0.0;
  // 4 samples per thread and iteration: one coalesced 16-byte load + the two samples after it (an 8-byte load that
  // hits the line the neighbouring lane just fetched); T % 4 == 0.  Few, long-lived blocks: the fp64 atomics at the
  // end go to only 4 addresses per night.
  const int n4 = p.T >> 2;
  auto fin = [](float v) { return isinf(v) ? 0.0f : v; };
#pragma unroll 4
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += gridDim.x * blockDim.x) {
    float4 f = __ldg(reinterpret_cast<const float4*>(xb) + q);
    float2 g = make_float2(0.0f, 0.0f);
    if (q + 1 < n4) g = __ldg(reinterpret_cast<const float2*>(xb + 4 * (size_t)(q + 1)));
    f.x = fin(f.x); f.y = fin(f.y); f.z = fin(f.z); f.w = fin(f.w);
    g.x = fin(g.x); g.y = fin(g.y);
    s0 += (double)((f.x + f.y) + (f.z + f.w));
    r0 += (double)fmaf(f.x, f.x, fmaf(f.y, f.y, fmaf(f.z, f.z, f.w * f.w)));
    r1 += (double)fmaf(f.x, f.y, fmaf(f.y, f.z, fmaf(f.z, f.w, f.w * g.x)));
    r2 += (double)fmaf(f.x, f.z, fmaf(f.y, f.w, fmaf(f.z, g.x, f.w * g.y)));
  }
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    r0 += __shfl_xor_sync(0xffffffffu, r0, o);
    r1 += __shfl_xor_sync(0xffffffffu, r1, o);
    r2 += __shfl_xor_sync(0xffffffffu, r2, o);
  }
  __shared__ double red[8][4];
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    red[w][0] = s0; red[w][1] = r0; red[w][2] = r1; red[w][3] = r2;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    atomicAdd(&p.xs[b * 4 + threadIdx.x], t);
  }
}
// stats1[b, c] = (sum y_c, sum y_c^2) from the four sums; one thread per (b, c)
struct XFinalizeArgs {
  const float* x;
  const double* xs;
  const float* w;
  const uint8_t* row_mask;
  double* stats1;
};
__global__ void x_stats_finalize_kernel(const XFinalizeArgs f0, const XFinalizeArgs f1, int B, int T) {
  const XFinalizeArgs f = blockIdx.y ? f1 : f0;
  const float* x = f.x;
  const double* xs = f.xs;
  const float* w = f.w;
  const uint8_t* row_mask = f.row_mask;
  double* stats1 = f.stats1;
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= B * 16) return;
  const int b = id / 16, c = id % 16;
  if (row_mask[b]) return;
  auto fix = [](float v) { return isinf(v) ? 0.0 : (double)v; };
  const double xf = fix(x[(size_t)b * T]), xl = fix(x[(size_t)b * T + T - 1]);
  const double S0 = xs[b * 4], R0 = xs[b * 4 + 1], R1 = xs[b * 4 + 2], R2 = xs[b * 4 + 3];
  const double w0 = w[c * 3], w1 = w[c * 3 + 1], w2 = w[c * 3 + 2];
  stats1[((size_t)b * 16 + c) * 2 + 0] = w0 * (S0 - xl) + w1 * S0 + w2 * (S0 - xf);
  stats1[((size_t)b * 16 + c) * 2 + 1] = w0 * w0 * (R0 - xl * xl) + w1 * w1 * R0 + w2 * w2 * (R0 - xf * xf) +
                                        2.0 * (w0 * w1 + w1 * w2) * R1 + 2.0 * w0 * w2 * R2;
}

inline cudaError_t launch_first_conv(const FirstConvArgs& a, int B, cudaStream_t stream) {
  dim3 grid((a.T + kFirstConvPos - 1) / kFirstConvPos, B);
  first_conv_kernel<<<grid, kFirstConvThreads, 0, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace w2s
