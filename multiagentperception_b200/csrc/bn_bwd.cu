// Backward of the train-mode conv -> BatchNorm2d (-> + residual) (-> ReLU) unit (conv2DBatchNormRelu /
// deconv2DBatchNormRelu, ptsemseg/models/utils.py:87-120,148-168; the BasicBlock tail of the resnet18 trunk,
// backbone.py:63-96) between the gradient of the layer OUTPUT and the gradient of the raw conv output z: what
// autograd runs for nn.ReLU + nn.BatchNorm2d in training mode under loss.backward() (ptsemseg/trainer.py:668-670).
//
//   du     = dy * [y > 0]                                  (ReLU; y is the stored forward output)     -> dres = du
//   dbeta  = sum du,  dgamma = sum du * xhat,              xhat = (z - mean) * invstd   (batch statistics of the forward)
//   dz     = gamma * invstd * (du - dbeta / M - xhat * dgamma / M)
// Without BatchNorm (simple_decoder.pred, backbone.py:150-154): dz = du, dbias = sum du.
// Three launches like the forward: per-channel sums (fp32 per thread, fp64 across threads), finalize, apply - all
// HBM-bound passes (algorithmic bytes: dy + y + z read twice, dz written once).
#include "common.cuh"

namespace w2c {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ void load8(const __nv_bfloat16* q, int cstride, int planes, bool f16, float (&v)[8]) {
  const uint4 hv = __ldg(reinterpret_cast<const uint4*>(q));
  const uint32_t* hb = reinterpret_cast<const uint32_t*>(&hv);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f = unpack_act2(hb[e], f16);
    v[2 * e] = f.x, v[2 * e + 1] = f.y;
  }
  if (planes == 2) {
    const uint4 lv = __ldg(reinterpret_cast<const uint4*>(q + cstride));
    const uint32_t* lb = reinterpret_cast<const uint32_t*>(&lv);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack_act2(lb[e], f16);
      v[2 * e] += f.x, v[2 * e + 1] += f.y;
    }
  }
}

__device__ __forceinline__ void store8(__nv_bfloat16* q, int cstride, int planes, bool f16, const float (&v)[8]) {
  uint4 hv, lv;
  uint32_t* hw = reinterpret_cast<uint32_t*>(&hv);
  uint32_t* lw = reinterpret_cast<uint32_t*>(&lv);
#pragma unroll
  for (int e = 0; e < 4; ++e) split_act2(v[2 * e], v[2 * e + 1], f16, hw[e], lw[e]);
  *reinterpret_cast<uint4*>(q) = hv;
  if (planes == 2) *reinterpret_cast<uint4*>(q + cstride) = lv;
}

struct BwdMaps {
  const __nv_bfloat16* dy;
  const __nv_bfloat16* y;
  const __nv_bfloat16* z;
  __nv_bfloat16* dz;
  __nv_bfloat16* dres;
  int dy_cs, dy_co, y_cs, y_co, z_cs, z_co, dz_cs, dz_co, dres_cs, dres_co;
  int act_f, act_g, relu;
  size_t n_px;
  int c;
  // ReLU mask without reading y: y > 0  <=>  z * scale + shift > 0 with the forward's own scale / shift (the forward
  // evaluates exactly this fmaf in fp32 and a positive fp32 never rounds to a zero bf16). NULL: read y (residual units).
  const float* fwd_scale;
  const float* fwd_shift;
};

// du = dy masked by the unit's ReLU; zv receives the raw conv output when it is needed (statistics or the mask)
__device__ __forceinline__ void masked_du(const BwdMaps& m, size_t px, int g, bool need_z, const float* __restrict__ sc,
                                          const float* __restrict__ sh, float (&du)[8], float (&zv)[8]) {
  const bool f16f = act_is_f16(m.act_f), f16g = act_is_f16(m.act_g);
  const int pf = act_planes(m.act_f), pg = act_planes(m.act_g);
  load8(m.dy + px * (static_cast<size_t>(m.dy_cs) * pg) + m.dy_co + g * 8, m.dy_cs, pg, f16g, du);
  const bool mask_from_z = m.relu && m.fwd_scale != nullptr;
  if (need_z || mask_from_z) load8(m.z + px * (static_cast<size_t>(m.z_cs) * pf) + m.z_co + g * 8, m.z_cs, pf, f16f, zv);
  if (m.relu) {
    if (mask_from_z) {
      // sc / sh: this group's 8 forward scale / shift values in shared memory (two 16-byte loads each)
      const float4 a0 = *reinterpret_cast<const float4*>(sc), a1 = *reinterpret_cast<const float4*>(sc + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(sh), b1 = *reinterpret_cast<const float4*>(sh + 4);
      const float s8[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float h8[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) du[e] = fmaf(zv[e], s8[e], h8[e]) > 0.f ? du[e] : 0.f;
    } else {
      float yv[8];
      load8(m.y + px * (static_cast<size_t>(m.y_cs) * pf) + m.y_co + g * 8, m.y_cs, pf, f16f, yv);
#pragma unroll
      for (int e = 0; e < 8; ++e) du[e] = yv[e] > 0.f ? du[e] : 0.f;
    }
  }
}

// ---- one-plane storages: the 16-byte loads of U pixels are issued back to back and kept RAW (4 registers each) until
// they are consumed one pixel at a time. With two pixels per trip unpacked on arrival the passes ran at 2.6 (reduce) and
// 3.5 TB/s (apply) of the 6.5 the part delivers (ncu launch list of the step, profiles/r2_train_step_launches_v4.md):
// 48 KB / 32 KB of loads in flight per SM; now 128 KB (two CTAs x 256 threads x 16 loads). NEED_Y: the ReLU mask comes from the stored output (residual units), else from z.
constexpr int kRawU = 8;    // pixels in flight per thread (dy + z: 64 registers of raw loads; two 256-thread CTAs per SM)
constexpr int kRawUY = 5;   // with y as well (60 registers)
constexpr int kRawUA = 5;   // apply pass: 40 registers of per-channel constants next to the raw loads
template <bool NEED_Y>
struct RawPx {
  uint4 dy, z, y;
};
template <bool NEED_Y>
__device__ __forceinline__ void issue_px(const BwdMaps& m, size_t px, int g, bool ld_z, RawPx<NEED_Y>& r) {
  r.dy = __ldg(reinterpret_cast<const uint4*>(m.dy + px * static_cast<size_t>(m.dy_cs) + m.dy_co + g * 8));
  if (ld_z) r.z = __ldg(reinterpret_cast<const uint4*>(m.z + px * static_cast<size_t>(m.z_cs) + m.z_co + g * 8));
  if (NEED_Y) r.y = __ldg(reinterpret_cast<const uint4*>(m.y + px * static_cast<size_t>(m.y_cs) + m.y_co + g * 8));
}
__device__ __forceinline__ void unpack8(const uint4& q, bool f16, float (&v)[8]) {
  const uint32_t* b = reinterpret_cast<const uint32_t*>(&q);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f = unpack_act2(b[e], f16);
    v[2 * e] = f.x, v[2 * e + 1] = f.y;
  }
}
// du = dy masked by the unit's ReLU, zv = z (when loaded): masked_du() on the raw loads
// s8 / h8: this thread's eight forward scale / shift values IN REGISTERS (ncu on the shared-memory version: the
// short-scoreboard stall - the per-pixel constant loads - was the top stall reason, 4.9-5.9 warps per issue, at 16
// resident warps per SM)
template <bool NEED_Y>
__device__ __forceinline__ void consume_px(const BwdMaps& m, const RawPx<NEED_Y>& r, bool ld_z, const float (&s8)[8],
                                           const float (&h8)[8], float (&du)[8], float (&zv)[8]) {
  unpack8(r.dy, act_is_f16(m.act_g), du);
  if (ld_z) unpack8(r.z, act_is_f16(m.act_f), zv);
  if (NEED_Y) {
    float yv[8];
    unpack8(r.y, act_is_f16(m.act_f), yv);
#pragma unroll
    for (int e = 0; e < 8; ++e) du[e] = yv[e] > 0.f ? du[e] : 0.f;
  } else if (m.relu) {
#pragma unroll
    for (int e = 0; e < 8; ++e) du[e] = fmaf(zv[e], s8[e], h8[e]) > 0.f ? du[e] : 0.f;
  }
}
__device__ __forceinline__ void load_const8(const float* __restrict__ p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
}

// sums[ch] += sum du, sums[c + ch] += sum du * xhat
// dynamic shared memory: [reduction scratch: kThreads * 16 doubles][per-channel constants: mean, inv, scale, shift: 4c floats]
__global__ void __launch_bounds__(kThreads, 2) bn_bwd_reduce_kernel(const BwdMaps m, const float* __restrict__ stats,
                                                                    double* __restrict__ sums) {
  extern __shared__ double s_red[];
  float* s_c = reinterpret_cast<float*>(s_red + static_cast<size_t>(kThreads) * 16);
  const int groups = m.c / 8;
  const int lanes = kThreads / groups;
  const int g = threadIdx.x % groups, lane = threadIdx.x / groups;
  for (int i = threadIdx.x; i < m.c; i += kThreads) {
    s_c[i] = stats ? stats[i] : 0.f;
    s_c[m.c + i] = stats ? stats[m.c + i] : 0.f;
    s_c[2 * m.c + i] = m.fwd_scale ? m.fwd_scale[i] : 0.f;
    s_c[3 * m.c + i] = m.fwd_shift ? m.fwd_shift[i] : 0.f;
  }
  __syncthreads();
  const float* sc = s_c + 2 * m.c + g * 8;
  const float* sh = s_c + 3 * m.c + g * 8;
  float s1[8], s2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s1[e] = s2[e] = 0.f;
  if (lane < lanes) {
    size_t px = static_cast<size_t>(blockIdx.x) * lanes + lane;
    const size_t stride = static_cast<size_t>(gridDim.x) * lanes;
    const bool bn = stats != nullptr;
    // s2 collects sum du * z; sum du * xhat = inv * (sum du * z - mean * sum du) follows once per thread, in fp64,
    // below - no per-channel constant in the loop
    auto accumulate = [&](const float (&du)[8], const float (&zv)[8]) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        s1[e] += du[e];
        if (bn) s2[e] = fmaf(du[e], zv[e], s2[e]);
      }
    };
    if (act_planes(m.act_f) == 1 && act_planes(m.act_g) == 1) {
      const bool need_y = m.relu && m.fwd_scale == nullptr;
      const bool ld_z = bn || (m.relu && !need_y);
      float s8[8], h8[8];
      load_const8(sc, s8), load_const8(sh, h8);
      if (!need_y) {
        for (; px + (kRawU - 1) * stride < m.n_px; px += kRawU * stride) {
          RawPx<false> r[kRawU];
#pragma unroll
          for (int u = 0; u < kRawU; ++u) issue_px(m, px + u * stride, g, ld_z, r[u]);
#pragma unroll
          for (int u = 0; u < kRawU; ++u) {
            float du[8], zv[8];
            consume_px(m, r[u], ld_z, s8, h8, du, zv);
            accumulate(du, zv);
          }
        }
      } else {
        for (; px + (kRawUY - 1) * stride < m.n_px; px += kRawUY * stride) {
          RawPx<true> r[kRawUY];
#pragma unroll
          for (int u = 0; u < kRawUY; ++u) issue_px(m, px + u * stride, g, ld_z, r[u]);
#pragma unroll
          for (int u = 0; u < kRawUY; ++u) {
            float du[8], zv[8];
            consume_px(m, r[u], ld_z, s8, h8, du, zv);
            accumulate(du, zv);
          }
        }
      }
    }
    // two pixels per trip (independent loads in flight): the two-plane storages, and the tails of the loops above
    for (; px + stride < m.n_px; px += 2 * stride) {
      float du[2][8], zv[2][8];
#pragma unroll
      for (int u = 0; u < 2; ++u) masked_du(m, px + u * stride, g, bn, sc, sh, du[u], zv[u]);
#pragma unroll
      for (int u = 0; u < 2; ++u) accumulate(du[u], zv[u]);
    }
    for (; px < m.n_px; px += stride) {
      float du[8], zv[8];
      masked_du(m, px, g, bn, sc, sh, du, zv);
      accumulate(du, zv);
    }
  }
  double* mine = s_red + static_cast<size_t>(threadIdx.x) * 16;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const double mean = s_c[g * 8 + e], inv = s_c[m.c + g * 8 + e];
    mine[e] = s1[e];
    mine[8 + e] = inv * (static_cast<double>(s2[e]) - mean * static_cast<double>(s1[e]));
  }
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
      atomicAdd(&sums[m.c + g * 8 + e], t[8 + e]);
    }
  }
}

// parameter gradients (accumulated, like autograd's .grad) + the per-channel coefficients of the apply pass
__global__ void bn_bwd_finalize_kernel(double* __restrict__ sums, double count, const float* __restrict__ gamma,
                                       const float* __restrict__ stats, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, float* __restrict__ coef, int c) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const double sb = sums[ch], sg = sums[c + ch];
  if (dbeta) dbeta[ch] += static_cast<float>(sb);
  if (dgamma && stats) dgamma[ch] += static_cast<float>(sg);
  // dz = k * (du - a - xhat * b), xhat = (z - mean) * inv   ->   dz = A * du + B * z + C
  const double a = sb / count, b = sg / count;
  const double k = stats ? static_cast<double>(gamma ? gamma[ch] : 1.f) * stats[c + ch] : 1.0;
  const double inv = stats ? stats[c + ch] : 0.0, mean = stats ? stats[ch] : 0.0;
  coef[ch] = static_cast<float>(k);
  coef[c + ch] = static_cast<float>(stats ? -k * b * inv : 0.0);
  coef[2 * c + ch] = static_cast<float>(stats ? -k * a + k * b * inv * mean : 0.0);
  sums[ch] = 0.0, sums[c + ch] = 0.0;
}

// dynamic shared memory: per-channel constants scale, shift, A, B, C: 5c floats
__global__ void __launch_bounds__(256, 2) bn_bwd_apply_kernel(const BwdMaps m, const float* __restrict__ stats,
                                                              const float* __restrict__ coef) {
  extern __shared__ float s_k[];
  for (int i = threadIdx.x; i < m.c; i += blockDim.x) {
    s_k[i] = m.fwd_scale ? m.fwd_scale[i] : 0.f;
    s_k[m.c + i] = m.fwd_shift ? m.fwd_shift[i] : 0.f;
    s_k[2 * m.c + i] = coef[i], s_k[3 * m.c + i] = coef[m.c + i], s_k[4 * m.c + i] = coef[2 * m.c + i];
  }
  __syncthreads();
  const int groups = m.c / 8;
  const size_t total = m.n_px * groups;
  const bool f16g = act_is_f16(m.act_g);
  const int pg = act_planes(m.act_g);
  const size_t first = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;   // a multiple of groups: g is fixed per thread
  const int g = first % groups;
  const float* kc = s_k + g * 8;
  float s8[8], h8[8], ka[8], kb[8], kk[8];   // this thread's channel group: forward scale / shift, A, B, C
  load_const8(kc, s8), load_const8(kc + m.c, h8);
  load_const8(kc + 2 * m.c, ka), load_const8(kc + 3 * m.c, kb), load_const8(kc + 4 * m.c, kk);
  auto finish = [&](size_t px, float (&du)[8], const float (&zv)[8]) {
    if (m.dres) store8(m.dres + px * (static_cast<size_t>(m.dres_cs) * pg) + m.dres_co + g * 8, m.dres_cs, pg, f16g, du);
    if (stats) {
#pragma unroll
      for (int e = 0; e < 8; ++e) du[e] = fmaf(ka[e], du[e], fmaf(kb[e], zv[e], kk[e]));
    }
    store8(m.dz + px * (static_cast<size_t>(m.dz_cs) * pg) + m.dz_co + g * 8, m.dz_cs, pg, f16g, du);
  };
  // (the grid stride is a multiple of groups: a thread keeps its channel group and walks pixels - no 64-bit division
  // per item, which cost more than the arithmetic of the pass)
  size_t px = first / groups;
  const size_t pstride = stride / groups;
  if (act_planes(m.act_f) == 1 && pg == 1) {
    // one-plane storages: the loads of kRawU (kRawUY with y) pixels in flight per thread, see RawPx
    const bool need_y = m.relu && m.fwd_scale == nullptr;
    const bool ld_z = stats != nullptr || (m.relu && !need_y);
    if (!need_y) {
      for (; px + (kRawUA - 1) * pstride < m.n_px; px += kRawUA * pstride) {
        RawPx<false> r[kRawUA];
#pragma unroll
        for (int u = 0; u < kRawUA; ++u) issue_px(m, px + u * pstride, g, ld_z, r[u]);
#pragma unroll
        for (int u = 0; u < kRawUA; ++u) {
          float du[8], zv[8];
          consume_px(m, r[u], ld_z, s8, h8, du, zv);
          finish(px + u * pstride, du, zv);
        }
      }
    } else {
      for (; px + (kRawUY - 1) * pstride < m.n_px; px += kRawUY * pstride) {
        RawPx<true> r[kRawUY];
#pragma unroll
        for (int u = 0; u < kRawUY; ++u) issue_px(m, px + u * pstride, g, ld_z, r[u]);
#pragma unroll
        for (int u = 0; u < kRawUY; ++u) {
          float du[8], zv[8];
          consume_px(m, r[u], ld_z, s8, h8, du, zv);
          finish(px + u * pstride, du, zv);
        }
      }
    }
  }
  for (; px < m.n_px; px += pstride) {
    float du[8], zv[8];
    masked_du(m, px, g, stats != nullptr, kc, kc + m.c, du, zv);
    if (m.dres) store8(m.dres + px * (static_cast<size_t>(m.dres_cs) * pg) + m.dres_co + g * 8, m.dres_cs, pg, f16g, du);
    if (stats) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 a = *reinterpret_cast<const float4*>(kc + 2 * m.c + 4 * h);
        const float4 b = *reinterpret_cast<const float4*>(kc + 3 * m.c + 4 * h);
        const float4 c = *reinterpret_cast<const float4*>(kc + 4 * m.c + 4 * h);
        du[4 * h + 0] = fmaf(a.x, du[4 * h + 0], fmaf(b.x, zv[4 * h + 0], c.x));
        du[4 * h + 1] = fmaf(a.y, du[4 * h + 1], fmaf(b.y, zv[4 * h + 1], c.y));
        du[4 * h + 2] = fmaf(a.z, du[4 * h + 2], fmaf(b.z, zv[4 * h + 2], c.z));
        du[4 * h + 3] = fmaf(a.w, du[4 * h + 3], fmaf(b.w, zv[4 * h + 3], c.w));
      }
    }
    store8(m.dz + px * (static_cast<size_t>(m.dz_cs) * pg) + m.dz_co + g * 8, m.dz_cs, pg, f16g, du);
  }
}

// ---- the logits layer: fp32 NCHW maps in (dy, y, z), NHWC gradient map out (channels c .. c_pad-1 zero-filled)
__global__ void __launch_bounds__(256) bn_bwd_reduce_nchw_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                                 const float* __restrict__ z, int n, int c, size_t hw,
                                                                 int chunks, int relu, const float* __restrict__ stats,
                                                                 double* __restrict__ sums) {
  __shared__ double s1[256], s2[256];
  const int ch = blockIdx.x / chunks, part = blockIdx.x % chunks;
  const float mean = stats ? stats[ch] : 0.f, inv = stats ? stats[c + ch] : 0.f;
  float a = 0.f, b = 0.f;
  for (int img = 0; img < n; ++img) {
    const size_t base = (static_cast<size_t>(img) * c + ch) * hw;
    for (size_t i = static_cast<size_t>(part) * 256 + threadIdx.x; i < hw; i += static_cast<size_t>(chunks) * 256) {
      float du = dy[base + i];
      if (relu && !(y[base + i] > 0.f)) du = 0.f;
      a += du;
      if (stats) b = fmaf(du, (z[base + i] - mean) * inv, b);
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

// one thread per (pixel, 8-channel group of the padded NHWC output)
__global__ void __launch_bounds__(256) bn_bwd_apply_nchw_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                                const float* __restrict__ z, __nv_bfloat16* __restrict__ dz,
                                                                int n, int c, size_t hw, int c_pad, int dz_cs, int dz_co,
                                                                int act_g, int relu, const float* __restrict__ stats,
                                                                const float* __restrict__ coef) {
  const int groups = c_pad / 8;
  const size_t total = static_cast<size_t>(n) * hw * groups;
  const bool f16g = act_is_f16(act_g);
  const int pg = act_planes(act_g);
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    // pixel fastest within a group so that the NCHW reads of a warp are contiguous
    const size_t px_all = idx % (static_cast<size_t>(n) * hw);
    const int g = static_cast<int>(idx / (static_cast<size_t>(n) * hw));
    const int img = static_cast<int>(px_all / hw);
    const size_t i = px_all % hw;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int ch = g * 8 + e;
      float du = 0.f;
      if (ch < c) {
        const size_t off = (static_cast<size_t>(img) * c + ch) * hw + i;
        du = dy[off];
        if (relu && !(y[off] > 0.f)) du = 0.f;
        if (stats) du = fmaf(coef[ch], du, fmaf(coef[c + ch], z[off], coef[2 * c + ch]));
      }
      v[e] = du;
    }
    store8(dz + px_all * (static_cast<size_t>(dz_cs) * pg) + dz_co + g * 8, dz_cs, pg, f16g, v);
  }
}

}  // namespace
}  // namespace w2c

using namespace w2c;

extern "C" int w2c_bn_train_bwd(const w2c_bn_bwd_args* args, w2c_stream_t stream) {
  if (!args) return set_error(W2C_ERR_INVALID, "bn_bwd: args is NULL");
  const w2c_bn_bwd_args& a = *args;
  W2C_CHECK_ARG(a.dy && a.dz && a.sums_ws && a.coef_ws, "bn_bwd: null pointer argument");
  W2C_CHECK_ARG(!a.relu || a.y || a.fwd_scale, "bn_bwd: the ReLU mask needs the forward output y (or fwd_scale / fwd_shift)");
  W2C_CHECK_ARG(!a.stats || a.z, "bn_bwd: BatchNorm needs the raw conv output z");
  W2C_CHECK_ARG(act_valid(a.act_f) && act_valid(a.act_g), "bn_bwd: bad act");
  W2C_CHECK_ARG(a.n_px > 0 && a.c > 0 && a.c % 8 == 0 && 256 % (a.c / 8) == 0 && a.c <= 2048,
                "bn_bwd: n_px=%lld c=%d (c / 8 must divide 256)", static_cast<long long>(a.n_px), a.c);
  BwdMaps m{};
  m.dy = static_cast<const __nv_bfloat16*>(a.dy), m.y = static_cast<const __nv_bfloat16*>(a.y);
  m.z = static_cast<const __nv_bfloat16*>(a.z), m.dz = static_cast<__nv_bfloat16*>(a.dz);
  m.dres = static_cast<__nv_bfloat16*>(a.dres);
  auto cs = [&](int v) { return v > 0 ? v : a.c; };
  m.dy_cs = cs(a.dy_cstride), m.dy_co = a.dy_coffset, m.y_cs = cs(a.y_cstride), m.y_co = a.y_coffset;
  m.z_cs = cs(a.z_cstride), m.z_co = a.z_coffset, m.dz_cs = cs(a.dz_cstride), m.dz_co = a.dz_coffset;
  m.dres_cs = cs(a.dres_cstride), m.dres_co = a.dres_coffset;
  m.act_f = a.act_f, m.act_g = a.act_g, m.relu = a.relu, m.n_px = static_cast<size_t>(a.n_px), m.c = a.c;
  m.fwd_scale = a.fwd_scale, m.fwd_shift = a.fwd_shift;
  W2C_CHECK_ARG((a.fwd_scale == nullptr) == (a.fwd_shift == nullptr), "bn_bwd: fwd_scale and fwd_shift go together");
  W2C_CHECK_ARG(!a.fwd_scale || a.z, "bn_bwd: the mask from z needs z");
  const int strides[5] = {m.dy_cs, m.y_cs, m.z_cs, m.dz_cs, m.dres_cs}, offs[5] = {m.dy_co, m.y_co, m.z_co, m.dz_co, m.dres_co};
  for (int i = 0; i < 5; ++i)
    W2C_CHECK_ARG(offs[i] >= 0 && offs[i] + a.c <= strides[i] && strides[i] % 8 == 0 && offs[i] % 8 == 0,
                  "bn_bwd: channel slice %d out of range", i);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int groups = a.c / 8, lanes = kThreads / groups;
  static DeviceOnce attr;
  const size_t smem = static_cast<size_t>(kThreads) * 16 * sizeof(double) + static_cast<size_t>(4) * a.c * sizeof(float);
  const size_t smem_max = static_cast<size_t>(kThreads) * 16 * sizeof(double) + static_cast<size_t>(4) * 2048 * sizeof(float);
  if (int rc = attr.ensure([=] {
        return cudaFuncSetAttribute(bn_bwd_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_max));
      }, "bn_bwd_reduce_kernel"))
    return rc;
  // at least 8 pixels per thread: every CTA ends in a shared-memory reduction and 2c fp64 atomics (a ~60 us floor per
  // launch when 1184 CTAs each brought one pixel per thread); at most ONE wave of resident CTAs (no tail)
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bn_bwd_reduce_kernel, kThreads, smem);
  long long want = (a.n_px + lanes * 8 - 1) / (lanes * 8);
  const int cap = device_sm_count() * (occ > 0 ? occ : 1);
  const int grid = static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
  bn_bwd_reduce_kernel<<<grid, kThreads, smem, s>>>(m, a.stats, a.sums_ws);
  W2C_CHECK_LAUNCH("bn_bwd_reduce_kernel");
  bn_bwd_finalize_kernel<<<(a.c + 127) / 128, 128, 0, s>>>(a.sums_ws, static_cast<double>(a.n_px), a.gamma, a.stats,
                                                           a.dgamma, a.dbeta, a.coef_ws, a.c);
  W2C_CHECK_LAUNCH("bn_bwd_finalize_kernel");
  // >= kRawU items per thread (one full trip of the raw-load loop), one wave of resident CTAs at most
  const size_t total = static_cast<size_t>(a.n_px) * groups;
  const size_t blocks = (total + 256 * kRawU - 1) / (256 * kRawU);
  const size_t smem2 = static_cast<size_t>(5) * a.c * sizeof(float);
  int occ2 = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, bn_bwd_apply_kernel, 256, smem2);
  const size_t cap2 = static_cast<size_t>(device_sm_count()) * (occ2 > 0 ? occ2 : 1);
  const int grid2 = static_cast<int>(blocks < cap2 ? (blocks ? blocks : 1) : cap2);
  bn_bwd_apply_kernel<<<grid2, 256, smem2, s>>>(m, a.stats, a.coef_ws);
  W2C_CHECK_LAUNCH("bn_bwd_apply_kernel");
  return W2C_OK;
}

extern "C" int w2c_bn_train_nchw_bwd(const float* dy, const float* y, const float* z, void* dz, int32_t n, int32_t c,
                                     int64_t hw, int32_t c_pad, int32_t dz_cstride, int32_t dz_coffset, int32_t act_g,
                                     int32_t relu, const float* gamma, const float* stats, float* dgamma, float* dbeta,
                                     double* sums_ws, float* coef_ws, w2c_stream_t stream) {
  W2C_CHECK_ARG(dy && dz && sums_ws && coef_ws && n > 0 && c > 0 && hw > 0, "bn_bwd_nchw: bad arguments");
  W2C_CHECK_ARG(!relu || y, "bn_bwd_nchw: the ReLU mask needs the forward output y");
  W2C_CHECK_ARG(!stats || z, "bn_bwd_nchw: BatchNorm needs the raw conv output z");
  W2C_CHECK_ARG(act_valid(act_g) && c_pad >= c && c_pad % 8 == 0, "bn_bwd_nchw: act_g=%d c_pad=%d", act_g, c_pad);
  const int cs = dz_cstride > 0 ? dz_cstride : c_pad;
  W2C_CHECK_ARG(dz_coffset >= 0 && dz_coffset + c_pad <= cs && cs % 8 == 0 && dz_coffset % 8 == 0,
                "bn_bwd_nchw: output channel slice out of range");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int chunks = static_cast<int>((hw + 256 * 16 - 1) / (256 * 16));
  const int cap = (device_sm_count() * 8 + c - 1) / c;
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  bn_bwd_reduce_nchw_kernel<<<c * chunks, 256, 0, s>>>(dy, y, z, n, c, static_cast<size_t>(hw), chunks, relu, stats, sums_ws);
  W2C_CHECK_LAUNCH("bn_bwd_reduce_nchw_kernel");
  bn_bwd_finalize_kernel<<<(c + 127) / 128, 128, 0, s>>>(sums_ws, static_cast<double>(n) * static_cast<double>(hw), gamma,
                                                         stats, dgamma, dbeta, coef_ws, c);
  W2C_CHECK_LAUNCH("bn_bwd_finalize_kernel");
  const size_t total = static_cast<size_t>(n) * static_cast<size_t>(hw) * (c_pad / 8);
  const size_t blocks = (total + 255) / 256;
  const size_t capb = static_cast<size_t>(device_sm_count()) * 32;
  bn_bwd_apply_nchw_kernel<<<static_cast<int>(blocks < capb ? blocks : capb), 256, 0, s>>>(
      dy, y, z, static_cast<__nv_bfloat16*>(dz), n, c, static_cast<size_t>(hw), c_pad, cs, dz_coffset, act_g, relu, stats,
      coef_ws);
  W2C_CHECK_LAUNCH("bn_bwd_apply_nchw_kernel");
  return W2C_OK;
}
