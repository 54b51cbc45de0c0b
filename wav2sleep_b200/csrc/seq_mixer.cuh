// Sequence mixer + classifier as ONE persistent kernel: a thread-block cluster per night walks all dilated layers.
// reference: SequenceCNN.forward (models/wav2sleep.py:379-390) = n_blocks x DilatedConvBlock.forward
//            (models/blocks.py:115-126): n_dil x [Conv1d(128->128, k7, dilation 2^k, no bias) -> ConvLayerNorm
//            (models/utils.py:9-23) -> GELU] -> (+ block input) -> GELU; then Wav2Sleep.classifier (wav2sleep.py:41,66).
//
// Why one kernel: a layer is 4.4 GFLOP and 10 MB of traffic for 16 nights - about 3 us of tensor time - but every layer
// needs the whole previous layer (taps reach 96 epochs away), so the layer-per-launch version (conv_igemm.cuh, 12
// launches of 160 one-tile CTAs, each re-staging its input and 229 KB of weights synchronously) is pure latency:
// 30 us per layer.  Here the layers of one night stay inside one cluster of <= 8 CTAs:
//   * every CTA owns a contiguous range of <= 512 epochs (up to four 128-row UMMA tiles, accumulators in TMEM);
//   * activations travel between layers through an L2-resident scratch in the UMMA "chunk-major" layout
//     [16-byte channel chunk][epoch + zero padding][8 halfs], so the next layer's operand (own rows + the halo rows of the
//     neighbour CTAs) is 16 contiguous byte ranges fetched by cp.async.bulk, and an epilogue warp's stores of one chunk
//     are 512 contiguous bytes; the zero padding of the convolution is real zero rows around every night;
//   * weights stream through a ring of 16 KB stages (half the K range of one tap) that runs ahead across layer
//     boundaries;
//   * a layer boundary is one cluster barrier (release / acquire) instead of a kernel boundary.
// Roles: warp 0 weight producer, warp 1 activation producer, warp 2 MMA issuer, warp 3 LayerNorm-parameter stager,
// warps 4..19 epilogue: warp e owns TMEM lane quadrant e % 4 and the 32-column quarter e / 4 of every tile; a row's
// LayerNorm statistics are the Chan-combination of the four quarters' (mean, M2), exchanged through shared memory
// (16 warps rather than 4 or 8: the epilogue is a dependent chain per thread - MUFU, LDS, TMEM round trips - and sits on
// the critical path of every layer, so it is spread over as many warps as the register file allows).
#pragma once
#include <cooperative_groups.h>

#include "conv_stream.cuh"

namespace w2s {

constexpr int kSeqMaxLayers = 32;
constexpr int kSeqThreads = 640;
constexpr int kSeqMaxCluster = 8;
constexpr int kSeqMaxTiles = 4;
constexpr int kSeqStageBytes = 16384;  // [8 chunks][128 cout][8 halfs]: one tap, half of the K range (4 K = 16 steps)
constexpr int kSeqStagesPerLayer = 14;
constexpr int kSeqCtlBytes = 1024 + 2 * 1024 + kSeqMaxTiles * 128 * 4 * 8;  // barriers, LN parameters x2, stats exchange

struct SeqArgs {
  const act_t* w[kSeqMaxLayers];      // packed [7][16][128][8] per layer
  const float* ln_w[kSeqMaxLayers];   // [128]
  const float* ln_b[kSeqMaxLayers];
  const act_t* x;        // [B, S, 128] row-major (epoch-mixer output)
  act_t* buf;            // 3 scratch tensors [B][16][S + 2 PAD][8] (block input / ping / pong)
  act_t* feat_out;       // [B, S, 128] row-major or null
  const float* head_w;   // [n_classes, 128]
  const float* head_b;
  float* logits;         // [B, S, n_classes]
  int n_classes;
  int n_blocks, n_dil;
  int S, PAD, SP;        // epochs per night, zero rows on each side (3 * max dilation), SP = S + 2 PAD
  int rpc;               // rows per CTA
  int ntiles;            // 128-row tiles per CTA (1..4)
  int AR;                // rows of the shared-memory operand per chunk (ntiles * 128 + 2 PAD)
  int n_stages;          // weight ring depth
  float ln_eps;
  int dbg;               // profiling knock-outs (tests/native/seq_probe.cu): 1 epilogue math, 2 weight copies, 4 MMAs,
                         // 8 activation copies, 16 the stage loops altogether (barriers only)
};

// Accurate GELU for the row epilogues, two elements at once: x Phi(x) with Phi from the fitted exponent
// (max abs error 2.5e-5, far below the fp16 rounding of the stored result; see gelu_fast in common.cuh).
W2S_DEVINL float2 gelu_acc2(float2 x) {
  float2 t = __fmul2_rn(x, x);
  t.x = fminf(t.x, 25.0f);
  t.y = fminf(t.y, 25.0f);
  float2 q = __ffma2_rn(t, make_float2(1.01448193e-3f, 1.01448193e-3f), make_float2(-0.10677673f, -0.10677673f));
  q = __ffma2_rn(q, t, make_float2(-2.30112048f, -2.30112048f));
  const float2 u = __fmul2_rn(x, q);
  float2 e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(u.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(u.y));
  const float2 d = __fadd2_rn(e, make_float2(1.0f, 1.0f));
  return make_float2(__fdividef(x.x, d.x), __fdividef(x.y, d.y));
}

// Profiling knock-outs of single stages are compiled in only for the stand-alone probe (tests/native/seq_probe.cu builds
// with -DW2S_SEQ_PROBE); in the library SeqArgs::dbg is ignored and the checks cost nothing.
#ifdef W2S_SEQ_PROBE
#define SEQ_DBG(p, mask) ((p).dbg & (mask))
#else
#define SEQ_DBG(p, mask) 0
#endif

W2S_DEVINL void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
W2S_DEVINL void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// generic-proxy global writes <-> async-proxy (bulk copy) reads
W2S_DEVINL void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
W2S_DEVINL void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

__global__ void __launch_bounds__(kSeqThreads, 1) seq_mixer_kernel(const SeqArgs p) {
  namespace cg = cooperative_groups;
  const cg::cluster_group cluster = cg::this_cluster();
  const int NC = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.x / NC;  // night
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r0 = rank * p.rpc;
  const int nrows = min(p.rpc, p.S - r0);  // >= 1 by construction of the launch
  const int n_layers = p.n_blocks * p.n_dil;
  const int NS = p.n_stages;
  const uint32_t tmem_cols = p.ntiles == 1 ? 128u : p.ntiles == 2 ? 256u : 512u;

  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem;                                   // [16][AR][16 B]
  uint8_t* sW = sA + (size_t)16 * p.AR * 16;            // [NS][16 KB]
  uint8_t* sCtl = sW + (size_t)NS * kSeqStageBytes;
  uint64_t* w_full = reinterpret_cast<uint64_t*>(sCtl);  // [NS <= 16]
  uint64_t* w_empty = w_full + 16;                        // [NS]
  uint64_t* a_full = w_empty + 16;                        // [2]  chunks 0..7 / 8..15
  uint64_t* acc_full = a_full + 2;                        // [1]
  uint64_t* ln_full = acc_full + 1;                       // [2]  LayerNorm parameters of layer l in sLn[l & 1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ln_full + 2);
  float* sLn = reinterpret_cast<float*>(sCtl + 1024);     // [2][256]: weight[128], bias[128]
  float2* sExch = reinterpret_cast<float2*>(sCtl + 1024 + 2048);  // [tile][row 128][quarter 4] (mean, M2) of 32 columns

  const size_t buf_elems = (size_t)16 * p.SP * 8;  // halfs per night per scratch tensor
  const size_t B_nights = gridDim.x / NC;
  act_t* bufX = p.buf + (size_t)b * buf_elems;
  act_t* bufP0 = p.buf + (B_nights + b) * buf_elems;
  act_t* bufP1 = p.buf + (2 * B_nights + b) * buf_elems;

  // ---------------- setup ----------------
  if (warp == 2) tmem_alloc(tmem_slot, tmem_cols);
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    mbar_init(&a_full[0], 1);
    mbar_init(&a_full[1], 1);
    mbar_init(acc_full, 1);
    mbar_init(&ln_full[0], 32);
    mbar_init(&ln_full[1], 32);
    fence_mbar_init();
  }
  // zero rows around the night in all three scratch tensors (first / last CTA of the cluster), then this CTA's rows of
  // x: row-major -> chunk-major
  {
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    const int npad = 16 * p.PAD;
    if (rank == 0) {
      for (int i = tid; i < 3 * npad; i += kSeqThreads) {
        const int t = i / npad, j = i - t * npad;
        const int c = j / p.PAD, row = j - c * p.PAD;
        act_t* dst = (t == 0 ? bufX : t == 1 ? bufP0 : bufP1);
        *reinterpret_cast<uint4*>(dst + ((size_t)c * p.SP + row) * 8) = z;
      }
    }
    if (rank == NC - 1) {
      for (int i = tid; i < 3 * npad; i += kSeqThreads) {
        const int t = i / npad, j = i - t * npad;
        const int c = j / p.PAD, row = j - c * p.PAD;
        act_t* dst = (t == 0 ? bufX : t == 1 ? bufP0 : bufP1);
        *reinterpret_cast<uint4*>(dst + ((size_t)c * p.SP + p.PAD + p.S + row) * 8) = z;
      }
    }
    const act_t* xb = p.x + ((size_t)b * p.S + r0) * 128;
    for (int i = tid; i < nrows * 16; i += kSeqThreads) {
      const int c = i & 15, row = i >> 4;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(xb + (size_t)row * 128) + c);
      *reinterpret_cast<uint4*>(bufX + ((size_t)c * p.SP + p.PAD + r0 + row) * 8) = v;
    }
  }
  fence_proxy_async_all();
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  cluster_arrive();
  cluster_wait();

  // Layer schedule shared by all roles: layer l = (block bl, dilation index k); input / output scratch tensors.
  //   k == 0: reads X; else reads P[(k-1)&1].  Writes P[k&1], except the last layer of a block, which adds X (the block
  //   input, same rows, same thread) and writes the block output back into X.
  auto in_buf = [&](int k) -> act_t* { return k == 0 ? bufX : (((k - 1) & 1) ? bufP1 : bufP0); };
  auto out_buf = [&](int k) -> act_t* { return k == p.n_dil - 1 ? bufX : ((k & 1) ? bufP1 : bufP0); };
  const bool stage_loops = !SEQ_DBG(p, 16);

  if (warp == 0) {
    // ---------------- weight producer: free-running ring, one 16 KB stage per (tap, K half) ----------------
    int s = 0;
    uint32_t ph = 0;
    for (int l = 0; l < n_layers; ++l) {
      const uint8_t* wl = reinterpret_cast<const uint8_t*>(p.w[l]);
      for (int st = 0; stage_loops && st < kSeqStagesPerLayer; ++st) {
        mbar_wait(&w_empty[s], ph ^ 1);
        if (elect_one()) {
          if (SEQ_DBG(p, 2)) {
            mbar_arrive(&w_full[s]);
          } else {
            mbar_arrive_expect_tx(&w_full[s], kSeqStageBytes);
            bulk_g2s(sW + (size_t)s * kSeqStageBytes, wl + (size_t)st * kSeqStageBytes, kSeqStageBytes, &w_full[s]);
          }
        }
        __syncwarp();
        if (++s == NS) {
          s = 0;
          ph ^= 1;
        }
      }
      // layer boundary: this warp only has to take part in the barrier.  It arrives as soon as its loads are issued and
      // collects that phase one layer late, so the ring keeps running ahead into the next layer's weights while the
      // other roles are still in the epilogue / barrier of this one.
      __syncwarp();
      if (l > 0) cluster_wait();
      cluster_arrive();
    }
    cluster_wait();
  } else if (warp == 1) {
    // ---------------- activation producer: 16 chunk ranges (own rows + halo) per layer ----------------
    for (int l = 0; l < n_layers; ++l) {
      const int k = l % p.n_dil;
      const int d = 1 << k;
      const uint8_t* src = reinterpret_cast<const uint8_t*>(in_buf(k));
      const uint32_t nld = (uint32_t)(nrows + 6 * d) * 16;
      if (!SEQ_DBG(p, 128)) fence_proxy_async_all();
      if (stage_loops && elect_one()) {
        for (int h = 0; h < 2; ++h) {
          if (SEQ_DBG(p, 8)) {
            mbar_arrive(&a_full[h]);
            continue;
          }
          mbar_arrive_expect_tx(&a_full[h], 8 * nld);
          for (int c = 8 * h; c < 8 * h + 8; ++c)
            bulk_g2s(sA + (size_t)c * p.AR * 16, src + ((size_t)c * p.SP + p.PAD + r0 - 3 * d) * 16, nld, &a_full[h]);
        }
      }
      __syncwarp();
      cluster_arrive();
      cluster_wait();
    }
  } else if (warp == 2) {
    // ---------------- MMA issuer ----------------
    constexpr uint32_t IDESC = umma_idesc_f16(128, 128, false);
    const uint32_t a_base = smem_u32(sA);
    const uint32_t lbo_a = (uint32_t)p.AR * 16;
    const int ntm = SEQ_DBG(p, 4) ? 0 : p.ntiles;
    int s = 0;
    uint32_t ph = 0;
    for (int l = 0; l < n_layers; ++l) {
      const int d = 1 << (l % p.n_dil);
      for (int st = 0; stage_loops && st < kSeqStagesPerLayer; ++st) {
        const int t = st >> 1, h = st & 1;  // packed weights: [tap][chunk]; a stage = chunks 8h .. 8h+7 of tap t
        if (t == 0) mbar_wait(&a_full[h], (uint32_t)(l & 1));
        mbar_wait(&w_full[s], ph);
        tc_fence_after_sync();
        if (elect_one()) {
          const uint32_t w_base = smem_u32(sW + (size_t)s * kSeqStageBytes);
          for (int j = 0; j < ntm; ++j) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint32_t a_off = ((uint32_t)(8 * h + 2 * kk) * p.AR + (uint32_t)(j * 128 + t * d)) * 16;
              umma_f16(tmem_base + j * 128, umma_smem_desc(a_base + a_off, lbo_a, 128),
                       umma_smem_desc(w_base + kk * 4096, 128 * 16, 128), IDESC, (st > 0 || kk > 0) ? 1u : 0u);
            }
          }
          umma_commit(&w_empty[s]);
        }
        __syncwarp();
        if (++s == NS) {
          s = 0;
          ph ^= 1;
        }
      }
      if (elect_one()) umma_commit(acc_full);
      __syncwarp();
      cluster_arrive();
      cluster_wait();
    }
  } else if (warp == 3) {
    // ---------------- LayerNorm parameters of the layer -> shared memory (double buffered) ----------------
    for (int l = 0; l < n_layers; ++l) {
      // buffer l & 1 was last read by the epilogue of layer l - 2, which ended before the barrier of layer l - 1
      float* dst = sLn + (l & 1) * 256;
      const float4 wv = __ldg(reinterpret_cast<const float4*>(p.ln_w[l]) + lane);
      const float4 bv = __ldg(reinterpret_cast<const float4*>(p.ln_b[l]) + lane);
      reinterpret_cast<float4*>(dst)[lane] = wv;
      reinterpret_cast<float4*>(dst + 128)[lane] = bv;
      mbar_arrive(&ln_full[l & 1]);
      __syncwarp();
      cluster_arrive();
      cluster_wait();
    }
  } else {
    // ---------------- epilogue: warp e = (lane quadrant e % 4, column quarter e / 4); thread = one epoch ----------------
    const int e = warp - 4;
    const int quad = e & 3, part = e >> 2;
    for (int l = 0; l < n_layers; ++l) {
      const int k = l % p.n_dil;
      const bool last_of_block = (k == p.n_dil - 1);
      const bool last = (l == n_layers - 1);
      mbar_wait(&ln_full[l & 1], (uint32_t)((l >> 1) & 1));
      mbar_wait(acc_full, (uint32_t)(l & 1));
      tc_fence_after_sync();
      const float* lnw = sLn + (l & 1) * 256 + part * 32;
      const float* lnb = lnw + 128;
      uint8_t* outb = reinterpret_cast<uint8_t*>(out_buf(k));
      const uint8_t* resb = reinterpret_cast<const uint8_t*>(bufX);
      for (int tile = 0; tile < p.ntiles; ++tile) {
        if (tile * 128 + quad * 32 >= nrows) break;  // warp-uniform; the partner warps (same quadrant) take the same path
        if (SEQ_DBG(p, 1)) continue;
        const int row = tile * 128 + quad * 32 + lane;  // row inside this CTA
        const bool valid = row < nrows;
        const size_t grow = (size_t)p.PAD + r0 + row;   // row in the padded scratch tensors
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + tile * 128 + part * 32, r);
        tmem_ld_wait();
        // ConvLayerNorm over the 128 channels of the row (utils.py:17-21, two-pass variance): (mean, M2) of this
        // thread's 32 columns, combined with the other quarters' (Chan): mean = avg m_i, M2 = sum M2_i + 32 sum (m_i - mean)^2
        float2 acc = make_float2(0.0f, 0.0f), acc1 = make_float2(0.0f, 0.0f);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          acc = __fadd2_rn(acc, make_float2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])));
          acc1 = __fadd2_rn(acc1, make_float2(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])));
        }
        const float m_loc = ((acc.x + acc.y) + (acc1.x + acc1.y)) * (1.0f / 32.0f);
        acc = acc1 = make_float2(0.0f, 0.0f);
        {
          const float2 nm = make_float2(-m_loc, -m_loc);
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float2 d0 = __fadd2_rn(make_float2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), nm);
            const float2 d1 = __fadd2_rn(make_float2(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])), nm);
            acc = __ffma2_rn(d0, d0, acc);
            acc1 = __ffma2_rn(d1, d1, acc1);
          }
        }
        const float m2_loc = (acc.x + acc.y) + (acc1.x + acc1.y);
        float2* ex = sExch + ((size_t)tile * 128 + quad * 32 + lane) * 4;
        ex[part] = make_float2(m_loc, m2_loc);
        named_bar_sync(1 + quad, 128);  // the four warps of this quadrant
        const float4 e01 = *reinterpret_cast<const float4*>(ex), e23 = *reinterpret_cast<const float4*>(ex + 2);
        const float mean = 0.25f * ((e01.x + e01.z) + (e23.x + e23.z));
        const float d0 = e01.x - mean, d1 = e01.z - mean, d2 = e23.x - mean, d3 = e23.z - mean;
        const float var = ((e01.y + e01.w) + (e23.y + e23.w) + 32.0f * ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3))) * (1.0f / 128.0f);
        const float rstd = rsqrtf(var + p.ln_eps);
        const float2 rs2 = make_float2(rstd, rstd);
        const float2 nm2 = make_float2(-mean * rstd, -mean * rstd);
        float lg[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) lg[c] = 0.0f;
        if (valid) {
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {  // 8 channels = one 16-byte chunk
            const int ch = part * 4 + c8;    // chunk index in the row
            const float4 w0 = *reinterpret_cast<const float4*>(lnw + c8 * 8), w1 = *reinterpret_cast<const float4*>(lnw + c8 * 8 + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(lnb + c8 * 8), b1 = *reinterpret_cast<const float4*>(lnb + c8 * 8 + 4);
            const float2 ww[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y), make_float2(w1.z, w1.w)};
            const float2 bb[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
            float2 v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 a = make_float2(__uint_as_float(r[c8 * 8 + 2 * q]), __uint_as_float(r[c8 * 8 + 2 * q + 1]));
              const float2 xh = __ffma2_rn(a, rs2, nm2);
              v[q] = __ffma2_rn(xh, ww[q], bb[q]);
              if (!SEQ_DBG(p, 256)) v[q] = gelu_acc2(v[q]);
            }
            if (last_of_block) {  // + block input (DilatedConvBlock.forward: act(out + x)), same row, read by its writer
              const uint4 rz = __ldcg(reinterpret_cast<const uint4*>(resb + ((size_t)ch * p.SP + grow) * 16));
              const uint32_t rr[4] = {rz.x, rz.y, rz.z, rz.w};
#pragma unroll
              for (int q = 0; q < 4; ++q) v[q] = gelu_acc2(__fadd2_rn(v[q], unpack_h2(rr[q])));
            }
            const uint4 o = make_uint4(pack_h2(v[0].x, v[0].y), pack_h2(v[1].x, v[1].y), pack_h2(v[2].x, v[2].y),
                                       pack_h2(v[3].x, v[3].y));
            if (!last) {
              if (!SEQ_DBG(p, 32)) *reinterpret_cast<uint4*>(outb + ((size_t)ch * p.SP + grow) * 16) = o;
            } else {
              if (p.feat_out != nullptr)
                *(reinterpret_cast<uint4*>(p.feat_out + ((size_t)b * p.S + r0 + row) * 128) + ch) = o;
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                if (c < p.n_classes) {
                  const float4 h0 = __ldg(reinterpret_cast<const float4*>(p.head_w + c * 128) + 2 * ch);
                  const float4 h1 = __ldg(reinterpret_cast<const float4*>(p.head_w + c * 128) + 2 * ch + 1);
                  float a = lg[c];
                  a = fmaf(v[0].x, h0.x, a);
                  a = fmaf(v[0].y, h0.y, a);
                  a = fmaf(v[1].x, h0.z, a);
                  a = fmaf(v[1].y, h0.w, a);
                  a = fmaf(v[2].x, h1.x, a);
                  a = fmaf(v[2].y, h1.y, a);
                  a = fmaf(v[3].x, h1.z, a);
                  a = fmaf(v[3].y, h1.w, a);
                  lg[c] = a;
                }
              }
            }
          }
        }
        if (last) {
          // classifier: the quarters' partial dot products meet in shared memory (the operand buffer is free: every MMA
          // of the last layer has completed); quarter 0 adds them up with the bias and writes the logits
          float* lx = reinterpret_cast<float*>(sA) + (size_t)row * 32;
          if (part > 0) {
#pragma unroll
            for (int c = 0; c < 8; ++c) lx[part * 8 + c] = lg[c];
          }
          named_bar_sync(1 + quad, 128);
          if (part == 0 && valid) {
            float* lrow = p.logits + ((size_t)b * p.S + r0 + row) * p.n_classes;
#pragma unroll
            for (int c = 0; c < 8; ++c)
              if (c < p.n_classes) lrow[c] = ((lg[c] + lx[8 + c]) + (lx[16 + c] + lx[24 + c])) + __ldg(p.head_b + c);
          }
        }
      }
      tc_fence_before_sync();
      if (!SEQ_DBG(p, 64)) fence_proxy_async_all();  // this layer's rows are read by bulk copies of the whole cluster after the barrier
      __syncwarp();
      cluster_arrive();
      cluster_wait();
    }
  }

  // ---------------- teardown ----------------
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, tmem_cols);
}

// Launch geometry of the fused kernel for B nights of S epochs; false = this shape is served by the layer-per-launch path.
struct SeqGeom {
  int nc, rpc, ntiles, PAD, SP, AR, n_stages;
  size_t smem;
};
inline bool seq_geometry_for(int S, int n_dil, int nc, SeqGeom& g) {
  g.PAD = 3 << (n_dil - 1);
  g.SP = S + 2 * g.PAD;
  g.rpc = (S + nc - 1) / nc;
  if (g.rpc > 128 * kSeqMaxTiles) return false;
  g.nc = (S + g.rpc - 1) / g.rpc;  // every CTA owns at least one row
  g.ntiles = (g.rpc + 127) / 128;
  g.AR = g.ntiles * 128 + 2 * g.PAD;
  const size_t a_bytes = (size_t)16 * g.AR * 16;
  const size_t budget = 232448;
  if (a_bytes + kSeqCtlBytes + 2 * kSeqStageBytes > budget) return false;
  size_t ns = (budget - a_bytes - kSeqCtlBytes) / kSeqStageBytes;
  if (ns > 8) ns = 8;
  g.n_stages = (int)ns;
  g.smem = a_bytes + ns * kSeqStageBytes + kSeqCtlBytes;
  return true;
}
inline cudaError_t seq_set_smem_attr(size_t smem) {
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(seq_mixer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  return cudaSuccess;
}
inline void seq_fill_launch(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int B, int nc, size_t smem, cudaStream_t stream) {
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(B * nc), 1, 1);
  cfg.blockDim = dim3(kSeqThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)nc;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
}
// Clusters of `nc` CTAs that can be resident at once (1 CTA per SM; the hardware places a cluster inside one GPC, so
// this is well below sm_count / nc for large clusters: 15 of size 8, 26 of size 5, 33 of size 4 on a B200).
inline int seq_max_active_clusters(int nc, size_t smem) {
  static int cache[kSeqMaxCluster + 1];
  static size_t cache_smem[kSeqMaxCluster + 1];
  if (cache[nc] != 0 && cache_smem[nc] == smem) return cache[nc];
  if (seq_set_smem_attr(smem) != cudaSuccess) return -1;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  seq_fill_launch(cfg, attr, 64, nc, smem, nullptr);
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, seq_mixer_kernel, &cfg) != cudaSuccess || n < 1) {
    cudaGetLastError();
    n = -1;
  }
  cache[nc] = n;
  cache_smem[nc] = smem;
  return n;
}
// Cluster size for B nights: the per-CTA time of a layer grows with the tiles per CTA (MMA and epilogue work) on top
// of a fixed barrier / hand-over cost, and the kernel takes ceil(B / resident clusters) waves.  nc_force > 0 pins it.
inline bool seq_fused_geometry(int B, int S, int n_dil, SeqGeom& best, int nc_force = 0) {
  if (n_dil < 2 || n_dil > 7 || S < 1 || B < 1) return false;
  double best_cost = 1e30;
  bool found = false;
  for (int nc = 1; nc <= kSeqMaxCluster; ++nc) {
    if (nc_force > 0 && nc != nc_force) continue;
    SeqGeom g;
    if (!seq_geometry_for(S, n_dil, nc, g) || g.nc != nc) continue;
    const int active = seq_max_active_clusters(nc, g.smem);
    if (active < 1) continue;
    const int waves = (B + active - 1) / active;
    const double cost = waves * (3.0 + 2.6 * g.ntiles + 0.15 * nc);  // us per layer: fixed + per tile + barrier fan-in
    if (cost < best_cost) {
      best_cost = cost;
      best = g;
      found = true;
    }
  }
  return found;
}
// Independent of the cluster size the launch picks
inline size_t seq_fused_workspace_bytes(int B, int S, int n_dil) {
  if (n_dil < 2 || n_dil > 7 || S < 1 || B < 1) return 0;
  const int PAD = 3 << (n_dil - 1);
  return (size_t)3 * B * 16 * (S + 2 * PAD) * 16;
}

inline cudaError_t launch_seq_mixer(SeqArgs a, const SeqGeom& g, int B, cudaStream_t stream) {
  cudaError_t e = seq_set_smem_attr(g.smem);
  if (e != cudaSuccess) return e;
  a.PAD = g.PAD; a.SP = g.SP; a.rpc = g.rpc; a.ntiles = g.ntiles; a.AR = g.AR; a.n_stages = g.n_stages;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  seq_fill_launch(cfg, attr, B, g.nc, g.smem, stream);
  return cudaLaunchKernelEx(&cfg, seq_mixer_kernel, a);
}

}  // namespace w2s
