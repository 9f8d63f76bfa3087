// Train-mode BatchNorm2d for the forward path: batch statistics over the folded agent-batch,
// the running-statistics update, and the normalise (+residual) (+ReLU) pass. Replaces nn.BatchNorm2d in training mode
// inside conv2DBatchNormRelu / deconv2DBatchNormRelu (ptsemseg/models/utils.py:110-114,152-164) and the resnet18 trunk
// (backbone.py:63-96) as Trainer_*.train() runs them after model.train() (trainer.py:659-669).
//
// The conv kernels are reused unchanged: the layer's conv runs with scale = 1, shift = conv bias, no ReLU and writes the
// pre-normalisation map z; then
//   1. bn_stats_kernel     per-channel sum(z), sum(z^2) over all N*H*W pixels (fp32 per thread, fp64 across threads)
//   2. bn_finalize_kernel  mean, biased variance -> scale = gamma / sqrt(var + eps), shift = beta - mean * scale;
//                          running_mean / running_var (unbiased variance, momentum) and num_batches_tracked updated
//                          IN PLACE in the module's own buffers, like nn.BatchNorm2d does
//   3. bn_apply_kernel     z <- act(z * scale + shift (+ residual)), in place, re-split into the storage planes
// All three are HBM-bound passes over the map (algorithmic bytes: 1 read, then 1 read + 1 write).
#include "common.cuh"

namespace w2c {
namespace {

constexpr int kStatThreads = 256;

// One thread per (pixel slice, 8-channel group). Block = 256 threads = (256 / groups) pixel lanes x groups; every
// thread walks pixels pixel_lane, pixel_lane + lanes * gridDim, ... accumulating 8 channel sums in fp32 (at most a few
// thousand terms each), then the block reduces over its pixel lanes in fp64 and adds to the global fp64 totals.
__global__ void __launch_bounds__(kStatThreads) bn_stats_kernel(const __nv_bfloat16* __restrict__ z, size_t n_px, int c,
                                                                int cstride, int coffset, int act,
                                                                double* __restrict__ sums) {
  extern __shared__ double s_red[];   // [2][kStatThreads][8] would be 32 KB: reduce one 8-vector at a time instead
  const int groups = c / 8;
  const int lanes = kStatThreads / groups;           // pixel lanes per block (groups divides 256 for c = 64..2048)
  const int g = threadIdx.x % groups, lane = threadIdx.x / groups;
  const bool f16 = act_is_f16(act);
  const int planes = act_planes(act);
  const size_t pix_elems = static_cast<size_t>(cstride) * planes;
  float s1[8], s2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s1[e] = s2[e] = 0.f;
  if (lane < lanes) {
    auto load_px = [&](size_t px, float (&v)[8]) {
      const __nv_bfloat16* p = z + px * pix_elems + coffset + g * 8;
      const uint4 hv = __ldg(reinterpret_cast<const uint4*>(p));
      const uint32_t* hb = reinterpret_cast<const uint32_t*>(&hv);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_act2(hb[e], f16);
        v[2 * e] = f.x, v[2 * e + 1] = f.y;
      }
      if (planes == 2) {
        const uint4 lv = __ldg(reinterpret_cast<const uint4*>(p + cstride));
        const uint32_t* lb = reinterpret_cast<const uint32_t*>(&lv);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack_act2(lb[e], f16);
          v[2 * e] += f.x, v[2 * e + 1] += f.y;
        }
      }
    };
    // four pixels per trip: four independent 16-byte loads in flight per thread (one per trip left the pass at a
    // seventh of the HBM rate - ncu, profiles/r2_train_step.md)
    size_t px = static_cast<size_t>(blockIdx.x) * lanes + lane;
    const size_t stride = static_cast<size_t>(gridDim.x) * lanes;
    if (planes == 1) {
      // one-plane storages: eight raw loads in flight, unpacked one pixel at a time (profiles/r2_bn_bwd_ncu.md)
      for (; px + 7 * stride < n_px; px += 8 * stride) {
        uint4 raw[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) raw[u] = __ldg(reinterpret_cast<const uint4*>(z + (px + u * stride) * pix_elems + coffset + g * 8));
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint32_t* hb = reinterpret_cast<const uint32_t*>(&raw[u]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = unpack_act2(hb[e], f16);
            s1[2 * e] += f.x, s2[2 * e] = fmaf(f.x, f.x, s2[2 * e]);
            s1[2 * e + 1] += f.y, s2[2 * e + 1] = fmaf(f.y, f.y, s2[2 * e + 1]);
          }
        }
      }
    }
    for (; px + 3 * stride < n_px; px += 4 * stride) {
      float v[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) load_px(px + u * stride, v[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int e = 0; e < 8; ++e) s1[e] += v[u][e], s2[e] = fmaf(v[u][e], v[u][e], s2[e]);
    }
    for (; px < n_px; px += stride) {
      float v[8];
      load_px(px, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) s1[e] += v[e], s2[e] = fmaf(v[e], v[e], s2[e]);
    }
  }
  // block reduction over the pixel lanes, in fp64: s_red[lane][g][16]
  double* mine = s_red + static_cast<size_t>(threadIdx.x) * 16;
#pragma unroll
  for (int e = 0; e < 8; ++e) mine[e] = s1[e], mine[8 + e] = s2[e];
  __syncthreads();
  if (lane == 0) {
    double t[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) t[e] = 0.0;
    for (int l = 0; l < lanes; ++l) {
      const double* o = s_red + static_cast<size_t>(l * groups + g) * 16;
#pragma unroll
      for (int e = 0; e < 16; ++e) t[e] += o[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      atomicAdd(&sums[g * 8 + e], t[e]);
      atomicAdd(&sums[c + g * 8 + e], t[8 + e]);
    }
  }
}

__global__ void bn_finalize_kernel(double* __restrict__ sums, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   long long* __restrict__ num_batches_tracked, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ stats, int c) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch < c) {
    const double mean = sums[ch] / count;
    double var = sums[c + ch] / count - mean * mean;   // biased variance: what the normalisation uses
    if (var < 0.0) var = 0.0;
    const float sc = (gamma ? gamma[ch] : 1.f) * static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    scale[ch] = sc;
    shift[ch] = (beta ? beta[ch] : 0.f) - static_cast<float>(mean) * sc;
    if (stats) {   // what the backward pass normalises with: xhat = (z - mean) * invstd
      stats[ch] = static_cast<float>(mean);
      stats[c + ch] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    }
    if (running_mean) {
      // nn.BatchNorm2d: running = (1 - momentum) * running + momentum * batch, the variance UNBIASED (n / (n - 1))
      const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * static_cast<float>(mean);
      running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * static_cast<float>(unbiased);
    }
    sums[ch] = 0.0, sums[c + ch] = 0.0;   // ready for the next forward: the program never re-zeroes them
  }
  if (ch == 0 && num_batches_tracked) *num_batches_tracked += 1;
}

// U items of one thread's grid-stride walk per trip (one-plane storages): all loads first, raw; returns the first
// index it did not handle.
template <int U, bool RES>
__device__ __forceinline__ size_t bn_apply_span(const __nv_bfloat16* z, __nv_bfloat16* y, const __nv_bfloat16* __restrict__ res,
                                                const float* __restrict__ scale, const float* __restrict__ shift, size_t idx,
                                                size_t stride, size_t total, int groups, size_t pix_elems,
                                                size_t y_pix_elems, int coffset, int y_coffset, bool f16, int relu) {
  const int g = idx % groups;   // the stride is a multiple of groups: g is fixed per thread
  const float4 sa = *reinterpret_cast<const float4*>(scale + g * 8), sb = *reinterpret_cast<const float4*>(scale + g * 8 + 4);
  const float4 ha = *reinterpret_cast<const float4*>(shift + g * 8), hb4 = *reinterpret_cast<const float4*>(shift + g * 8 + 4);
  const float sc[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
  const float sh[8] = {ha.x, ha.y, ha.z, ha.w, hb4.x, hb4.y, hb4.z, hb4.w};
  const size_t pstride = stride / groups;   // (no 64-bit division per item)
  size_t px0 = idx / groups;
  for (; idx + (U - 1) * stride < total; idx += U * stride, px0 += U * pstride) {
    uint4 zr[U], rr[RES ? U : 1];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t px = px0 + u * pstride;
      zr[u] = *reinterpret_cast<const uint4*>(z + px * pix_elems + coffset + g * 8);
      if (RES) rr[u] = *reinterpret_cast<const uint4*>(res + px * y_pix_elems + y_coffset + g * 8);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t px = px0 + u * pstride;
      const uint32_t* zb = reinterpret_cast<const uint32_t*>(&zr[u]);
      const uint32_t* rb = reinterpret_cast<const uint32_t*>(&rr[RES ? u : 0]);
      uint4 hv;
      uint32_t* hw = reinterpret_cast<uint32_t*>(&hv);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_act2(zb[e], f16);
        float a = fmaf(f.x, sc[2 * e], sh[2 * e]), b = fmaf(f.y, sc[2 * e + 1], sh[2 * e + 1]);
        if (RES) {
          const float2 r = unpack_act2(rb[e], f16);
          a += r.x, b += r.y;
        }
        if (relu) a = fmaxf(a, 0.f), b = fmaxf(b, 0.f);
        uint32_t lo;
        split_act2(a, b, f16, hw[e], lo);
      }
      *reinterpret_cast<uint4*>(y + px * y_pix_elems + y_coffset + g * 8) = hv;
    }
  }
  return idx;
}

// y <- act(z * scale + shift (+ residual)) (y may be z: in place); one thread per (pixel, 8-channel group)
__global__ void __launch_bounds__(256, 3) bn_apply_kernel(const __nv_bfloat16* z, __nv_bfloat16* y,
                                                       const __nv_bfloat16* __restrict__ res,
                                                       const float* __restrict__ scale, const float* __restrict__ shift,
                                                       size_t n_px, int c, int cstride, int coffset, int y_cstride,
                                                       int y_coffset, int act, int relu) {
  const int groups = c / 8;
  const size_t total = n_px * groups;
  const bool f16 = act_is_f16(act);
  const int planes = act_planes(act);
  const size_t pix_elems = static_cast<size_t>(cstride) * planes;
  const size_t y_pix_elems = static_cast<size_t>(y_cstride) * planes;
  size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;   // a multiple of groups: g is fixed per thread
  if (planes == 1) {
    // one-plane storages: the 16-byte loads of eight pixels (four with a residual) are issued back to back and kept raw
    // until they are consumed: one load per thread and trip left the pass at 4.3 of the 6.5 TB/s (ncu launch list)
    if (res)
      idx = bn_apply_span<4, true>(z, y, res, scale, shift, idx, stride, total, groups, pix_elems, y_pix_elems, coffset,
                                   y_coffset, f16, relu);
    else
      idx = bn_apply_span<8, false>(z, y, res, scale, shift, idx, stride, total, groups, pix_elems, y_pix_elems, coffset,
                                    y_coffset, f16, relu);
  }
  for (; idx < total; idx += stride) {
    const int g = idx % groups;
    const size_t px = idx / groups;
    const __nv_bfloat16* p = z + px * pix_elems + coffset + g * 8;
    __nv_bfloat16* q_out = y + px * y_pix_elems + y_coffset + g * 8;
    auto load8 = [&](const __nv_bfloat16* q, float (&v)[8], int cstride) {
      const uint4 hv = *reinterpret_cast<const uint4*>(q);
      const uint32_t* hb = reinterpret_cast<const uint32_t*>(&hv);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_act2(hb[e], f16);
        v[2 * e] = f.x, v[2 * e + 1] = f.y;
      }
      if (planes == 2) {
        const uint4 lv = *reinterpret_cast<const uint4*>(q + cstride);
        const uint32_t* lb = reinterpret_cast<const uint32_t*>(&lv);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack_act2(lb[e], f16);
          v[2 * e] += f.x, v[2 * e + 1] += f.y;
        }
      }
    };
    float v[8];
    load8(p, v, cstride);
    const float4 sa = *reinterpret_cast<const float4*>(scale + g * 8), sb = *reinterpret_cast<const float4*>(scale + g * 8 + 4);
    const float4 ha = *reinterpret_cast<const float4*>(shift + g * 8), hb4 = *reinterpret_cast<const float4*>(shift + g * 8 + 4);
    const float sc[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
    const float sh[8] = {ha.x, ha.y, ha.z, ha.w, hb4.x, hb4.y, hb4.z, hb4.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = fmaf(v[e], sc[e], sh[e]);
    if (res) {
      float r[8];
      load8(res + px * y_pix_elems + y_coffset + g * 8, r, y_cstride);   // the residual is laid out like the output
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += r[e];
    }
    if (relu) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
    }
    uint4 hv, lv;
    uint32_t* hw = reinterpret_cast<uint32_t*>(&hv);
    uint32_t* lw = reinterpret_cast<uint32_t*>(&lv);
#pragma unroll
    for (int e = 0; e < 4; ++e) split_act2(v[2 * e], v[2 * e + 1], f16, hw[e], lw[e]);
    *reinterpret_cast<uint4*>(q_out) = hv;
    if (planes == 2) *reinterpret_cast<uint4*>(q_out + y_cstride) = lv;
  }
}

// ---- the same for an fp32 NCHW map (the logits layer: deconv12 is conv + BatchNorm + ReLU too, backbone.py:124, and
// its 11 channels are written in the reference's NCHW layout)
__global__ void __launch_bounds__(256) bn_stats_nchw_kernel(const float* __restrict__ z, int n, int c, size_t hw,
                                                            int chunks, double* __restrict__ sums) {
  __shared__ double s1[256], s2[256];
  const int ch = blockIdx.x / chunks, part = blockIdx.x % chunks;
  float a = 0.f, b = 0.f;
  for (int img = 0; img < n; ++img) {
    const float* p = z + (static_cast<size_t>(img) * c + ch) * hw;
    for (size_t i = static_cast<size_t>(part) * 256 + threadIdx.x; i < hw; i += static_cast<size_t>(chunks) * 256) {
      const float v = p[i];
      a += v, b = fmaf(v, v, b);
    }
  }
  s1[threadIdx.x] = a, s2[threadIdx.x] = b;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) s1[threadIdx.x] += s1[threadIdx.x + st], s2[threadIdx.x] += s2[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(&sums[ch], s1[0]), atomicAdd(&sums[c + ch], s2[0]);
}

__global__ void __launch_bounds__(256) bn_apply_nchw_kernel(const float* z, float* y, const float* __restrict__ scale,
                                                            const float* __restrict__ shift, int c, size_t hw,
                                                            size_t total, int relu) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>((i / hw) % c);
    float v = fmaf(z[i], scale[ch], shift[ch]);
    y[i] = relu ? fmaxf(v, 0.f) : v;
  }
}

}  // namespace
}  // namespace w2c

using namespace w2c;

extern "C" int w2c_bn_train_nchw_fwd(float* z, int32_t n, int32_t c, int64_t hw, int32_t relu, const float* gamma,
                                     const float* beta, float eps, float momentum, float* running_mean,
                                     float* running_var, int64_t* num_batches_tracked, double* sums_ws, float* scale_ws,
                                     float* shift_ws, float* y_out, float* stats_out, w2c_stream_t stream) {
  W2C_CHECK_ARG(z && sums_ws && scale_ws && shift_ws && n > 0 && c > 0 && hw > 0, "bn_train_nchw: bad arguments");
  W2C_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "bn_train_nchw: running_mean and running_var go together");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int chunks = static_cast<int>((hw + 256 * 16 - 1) / (256 * 16));
  const int cap = (device_sm_count() * 8 + c - 1) / c;
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  bn_stats_nchw_kernel<<<c * chunks, 256, 0, s>>>(z, n, c, static_cast<size_t>(hw), chunks, sums_ws);
  W2C_CHECK_LAUNCH("bn_stats_nchw_kernel");
  bn_finalize_kernel<<<(c + 127) / 128, 128, 0, s>>>(sums_ws, static_cast<double>(n) * static_cast<double>(hw), gamma, beta,
                                                     eps, momentum, running_mean, running_var,
                                                     reinterpret_cast<long long*>(num_batches_tracked), scale_ws, shift_ws,
                                                     stats_out, c);
  W2C_CHECK_LAUNCH("bn_finalize_kernel");
  const size_t total = static_cast<size_t>(n) * c * static_cast<size_t>(hw);
  const size_t blocks = (total + 255) / 256;
  const size_t capb = static_cast<size_t>(device_sm_count()) * 32;
  bn_apply_nchw_kernel<<<static_cast<int>(blocks < capb ? blocks : capb), 256, 0, s>>>(z, y_out ? y_out : z, scale_ws, shift_ws, c,
                                                                                       static_cast<size_t>(hw), total, relu);
  W2C_CHECK_LAUNCH("bn_apply_nchw_kernel");
  return W2C_OK;
}

static int bn_train_impl(void* z, const void* residual, int64_t n_px, int32_t c, int32_t cstride, int32_t coffset,
                         int32_t act, int32_t relu, const float* gamma, const float* beta, float eps, float momentum,
                         float* running_mean, float* running_var, int64_t* num_batches_tracked, double* sums_ws,
                         float* scale_ws, float* shift_ws, void* y_out, int32_t y_cstride, int32_t y_coffset,
                         float* stats_out, bool sums_ready, w2c_stream_t stream) {
  W2C_CHECK_ARG(z && sums_ws && scale_ws && shift_ws, "bn_train: null pointer argument");
  W2C_CHECK_ARG(act_valid(act), "bn_train: bad act %d", act);
  W2C_CHECK_ARG(n_px > 0 && c > 0 && c % 8 == 0 && 256 % (c / 8) == 0 && c <= 2048,
                "bn_train: n_px=%lld c=%d (c / 8 must divide 256)", static_cast<long long>(n_px), c);
  const int cs = cstride > 0 ? cstride : c;
  W2C_CHECK_ARG(coffset >= 0 && coffset + c <= cs && cs % 8 == 0 && coffset % 8 == 0, "bn_train: channel slice out of range");
  W2C_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "bn_train: running_mean and running_var go together");
  const int ycs = y_cstride > 0 ? y_cstride : c;
  W2C_CHECK_ARG(!y_out || (y_coffset >= 0 && y_coffset + c <= ycs && ycs % 8 == 0 && y_coffset % 8 == 0),
                "bn_train: output channel slice out of range");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int groups = c / 8, lanes = kStatThreads / groups;
  // at least 8 pixels per thread (four loads in flight, two trips): every CTA ends in a shared-memory reduction and 2c
  // fp64 atomics, a ~60 us floor per launch when 1184 CTAs each brought one pixel per thread
  long long want = (n_px + lanes * 8 - 1) / (lanes * 8);
  static DeviceOnce attr;
  const size_t smem = static_cast<size_t>(kStatThreads) * 16 * sizeof(double);   // 32 KB
  if (int rc = attr.ensure([=] {
        return cudaFuncSetAttribute(bn_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      }, "bn_stats_kernel"))
    return rc;
  int occ = 1;   // one wave of resident CTAs at most: a partial second wave doubled the time of the largest maps
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bn_stats_kernel, kStatThreads, smem);
  const int cap = device_sm_count() * (occ > 0 ? occ : 1);
  const int grid = static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
  if (!sums_ready) {   // (else the conv that wrote z already added its sums: w2c_conv_args.bn_sums)
    bn_stats_kernel<<<grid, kStatThreads, smem, s>>>(static_cast<const __nv_bfloat16*>(z), static_cast<size_t>(n_px), c, cs,
                                                     coffset, act, sums_ws);
    W2C_CHECK_LAUNCH("bn_stats_kernel");
  }
  bn_finalize_kernel<<<(c + 127) / 128, 128, 0, s>>>(sums_ws, static_cast<double>(n_px), gamma, beta, eps, momentum,
                                                     running_mean, running_var,
                                                     reinterpret_cast<long long*>(num_batches_tracked), scale_ws,
                                                     shift_ws, stats_out, c);
  W2C_CHECK_LAUNCH("bn_finalize_kernel");
  const size_t total = static_cast<size_t>(n_px) * groups;
  const size_t blocks = (total + 256 * 8 - 1) / (256 * 8);   // eight items per thread: one trip of the raw-load loop
  int occ2 = 1;   // one wave of resident CTAs (the grid stride must stay a multiple of groups: 256 is)
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, bn_apply_kernel, 256, 0);
  const size_t cap2 = static_cast<size_t>(device_sm_count()) * (occ2 > 0 ? occ2 : 1);
  const int grid2 = static_cast<int>(blocks < cap2 ? (blocks ? blocks : 1) : cap2);
  bn_apply_kernel<<<grid2, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(z),
                                        static_cast<__nv_bfloat16*>(y_out ? y_out : z),
                                        static_cast<const __nv_bfloat16*>(residual),
                                        scale_ws, shift_ws, static_cast<size_t>(n_px), c, cs, coffset,
                                        y_out ? ycs : cs, y_out ? y_coffset : coffset, act, relu);
  W2C_CHECK_LAUNCH("bn_apply_kernel");
  return W2C_OK;
}

extern "C" int w2c_bn_train_fwd(void* z, const void* residual, int64_t n_px, int32_t c, int32_t cstride, int32_t coffset,
                                int32_t act, int32_t relu, const float* gamma, const float* beta, float eps,
                                float momentum, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                                double* sums_ws, float* scale_ws, float* shift_ws, void* y_out, int32_t y_cstride,
                                int32_t y_coffset, float* stats_out, w2c_stream_t stream) {
  return bn_train_impl(z, residual, n_px, c, cstride, coffset, act, relu, gamma, beta, eps, momentum, running_mean, running_var,
                       num_batches_tracked, sums_ws, scale_ws, shift_ws, y_out, y_cstride, y_coffset, stats_out, false, stream);
}

extern "C" int w2c_bn_train_from_sums_fwd(void* z, const void* residual, int64_t n_px, int32_t c, int32_t cstride,
                                          int32_t coffset, int32_t act, int32_t relu, const float* gamma, const float* beta,
                                          float eps, float momentum, float* running_mean, float* running_var,
                                          int64_t* num_batches_tracked, double* sums_ws, float* scale_ws, float* shift_ws,
                                          void* y_out, int32_t y_cstride, int32_t y_coffset, float* stats_out,
                                          w2c_stream_t stream) {
  return bn_train_impl(z, residual, n_px, c, cstride, coffset, act, relu, gamma, beta, eps, momentum, running_mean, running_var,
                       num_batches_tracked, sums_ws, scale_ws, shift_ws, y_out, y_cstride, y_coffset, stats_out, true, stream);
}
