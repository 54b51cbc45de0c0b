// Input staging on device: whole-night z-score of a raw signal + -inf fill of missing rows.
//
// Replaces the per-item CPU work of ParquetDataset.__getitem__ (data/dataset.py:76-87 `_zscore_normalize`,
// :170-173 -inf padding) for batches that arrive on the GPU as raw fp32 / fp16 / int16 samples: shipping int16 halves
// the host->device bytes of the forward's inputs, which is what limits end-to-end throughput on 8 GPUs.
//   pass 1: per-night sum / sum of squares (fp64) and a non-finite flag
//   pass 2: out = (x - mean) / max(std, 1e-6), std = unbiased (N - 1) as torch.std; nights with a non-finite sample
//           are passed through unchanged (dataset.py:81-83); nights flagged absent become rows of -inf
#pragma once
#include "common.cuh"

namespace w2s {

enum { STAGE_F32 = 0, STAGE_F16 = 1, STAGE_I16 = 2 };

template <int DT>
W2S_DEVINL float4 stage_load4(const void* raw, size_t i4) {  // 4 consecutive samples as fp32
  if (DT == STAGE_F32) return __ldg(reinterpret_cast<const float4*>(raw) + i4);
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(raw) + i4);
  if (DT == STAGE_F16) {
    const float2 a = unpack_h2(u.x), b = unpack_h2(u.y);
    return make_float4(a.x, a.y, b.x, b.y);
  }
  return make_float4((float)(short)(u.x & 0xffff), (float)(short)(u.x >> 16), (float)(short)(u.y & 0xffff),
                     (float)(short)(u.y >> 16));
}

struct StageArgs {
  const void* raw;         // [B, T]
  float* out;              // [B, T] fp32
  const uint8_t* present;  // [B] (0 = signal missing for this night) or null
  double* ws;              // [B, 3]: sum, sumsq, non-finite count
  long long T;             // multiple of 4
};

template <int DT>
__global__ void __launch_bounds__(256) stage_stats_kernel(const StageArgs p) {
  const int b = blockIdx.y;
  if (p.present != nullptr && !p.present[b]) return;
  const long long n4 = p.T / 4;
  double s = 0.0, q = 0.0;
  int bad = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = stage_load4<DT>(p.raw, (size_t)b * n4 + i);
    s += (double)((v.x + v.y) + (v.z + v.w));
    q += (double)((v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w));
    bad |= !(isfinite(v.x) && isfinite(v.y) && isfinite(v.z) && isfinite(v.w));
  }
  __shared__ double rs[8], rq[8];
  __shared__ int rb[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
    bad |= __shfl_xor_sync(0xffffffffu, bad, o);
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { rs[w] = s; rq[w] = q; rb[w] = bad; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) { s += rs[k]; q += rq[k]; bad |= rb[k]; }
    s = rs[0] + (s - rs[0]);  // (keeps the fixed summation order explicit)
    atomicAdd(&p.ws[b * 3 + 0], s);
    atomicAdd(&p.ws[b * 3 + 1], q);
    if (bad) atomicAdd(&p.ws[b * 3 + 2], 1.0);
  }
}

template <int DT>
__global__ void __launch_bounds__(256) stage_apply_kernel(const StageArgs p) {
  const int b = blockIdx.y;
  const long long n4 = p.T / 4;
  const bool absent = p.present != nullptr && !p.present[b];
  float mean = 0.0f, inv = 1.0f;
  if (!absent && p.ws[b * 3 + 2] == 0.0) {
    const double n = (double)p.T;
    const double m = p.ws[b * 3 + 0] / n;
    const double var = n > 1.0 ? fmax((p.ws[b * 3 + 1] - n * m * m) / (n - 1.0), 0.0) : 0.0;  // torch.std: unbiased
    const double sd = sqrt(var);
    mean = (float)m;
    inv = (float)(1.0 / (sd > 1e-6 ? sd : 1e-6));
  }
  float4* out = reinterpret_cast<float4*>(p.out) + (size_t)b * n4;
  const float ninf = __int_as_float(0xff800000);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    if (absent) {
      out[i] = make_float4(ninf, ninf, ninf, ninf);
      continue;
    }
    const float4 v = stage_load4<DT>(p.raw, (size_t)b * n4 + i);
    out[i] = make_float4((v.x - mean) * inv, (v.y - mean) * inv, (v.z - mean) * inv, (v.w - mean) * inv);
  }
}

}  // namespace w2s
