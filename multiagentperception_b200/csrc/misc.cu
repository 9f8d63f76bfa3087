// Small HBM-bound kernels around the tensor-core convs: weight packing, BN folding, the 3-channel stems that read
// the caller's fp32 NCHW batch, max-pool, bilinear up-sampling, layout conversion, and library bookkeeping.
#include <cstring>

#include "common.cuh"

namespace w2c {

// ------------------------------------------------------------------------------------------ bookkeeping
static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

namespace {

inline int grid_for(size_t total, int threads, int max_blocks = 148 * 32) {
  size_t b = (total + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > static_cast<size_t>(max_blocks)) b = max_blocks;
  return static_cast<int>(b);
}

// ------------------------------------------------------------------------------------------ weight packing
__global__ void pack_weight_kernel(const float* __restrict__ w, int cout, int cin_real, int cin, int ntaps,
                                   int transposed, int planes, int cout_pad, __nv_bfloat16* __restrict__ out, bool f16) {
  const size_t ktot = static_cast<size_t>(ntaps) * cin;
  const size_t total = static_cast<size_t>(cout_pad) * ktot;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int co = idx / ktot;
    const int k = idx % ktot;
    const int tap = k / cin, ci = k % cin;
    float v = 0.f;
    if (co < cout && ci < cin_real) {
      // Conv2d.weight [co][ci][tap];  ConvTranspose2d.weight [ci][co][tap]
      v = transposed ? w[(static_cast<size_t>(ci) * cout + co) * ntaps + tap]
                     : w[(static_cast<size_t>(co) * cin_real + ci) * ntaps + tap];
    }
    const __nv_bfloat16 hi = float_to_elem(v, f16);
    out[idx] = hi;
    if (planes == 2) out[total + idx] = float_to_elem(v - elem_to_float(hi, f16), f16);
  }
}

__global__ void fold_bn_kernel(const float* bias, const float* gamma, const float* beta, const float* mean,
                               const float* var, float eps, int cout, float* scale, float* shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cout) return;
  const float b = bias ? bias[c] : 0.f;
  if (gamma) {
    // same association as F.batch_norm's eval formula: (x - mean) / sqrt(var + eps) * gamma + beta
    const float s = gamma[c] / sqrtf(var[c] + eps);
    scale[c] = s;
    shift[c] = beta[c] + (b - mean[c]) * s;
  } else {
    scale[c] = 1.f;
    shift[c] = b;
  }
}

// ------------------------------------------------------------------------------------------ 3-channel stems
// Conv2d(3->cout, k3 s1 p1) + affine + ReLU. One thread per output pixel; the 27 inputs live in registers, the
// weights in shared memory as [27][cout] so the per-k reads are warp-wide broadcasts of float4.
template <int GROUP>
__global__ void __launch_bounds__(256) stem3x3_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ scale,
                                                      const float* __restrict__ shift, __nv_bfloat16* __restrict__ y,
                                                      int b_sz, int n_agents, int c_total, int c_first, int h,
                                                      int wpx, int cout, int act) {
  extern __shared__ float sm[];
  float* s_w = sm;                // [27][cout]
  float* s_scale = sm + 27 * cout;
  float* s_shift = s_scale + cout;
  for (int i = threadIdx.x; i < 27 * cout; i += blockDim.x) {
    const int k = i / cout, co = i % cout;
    s_w[i] = w[co * 27 + k];
  }
  for (int i = threadIdx.x; i < cout; i += blockDim.x) s_scale[i] = scale[i], s_shift[i] = shift[i];
  __syncthreads();

  const size_t plane = static_cast<size_t>(h) * wpx;
  const size_t total = static_cast<size_t>(b_sz) * n_agents * plane;
  const int planes = act_planes(act);
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ow = idx % wpx;
    const int oh = (idx / wpx) % h;
    const int img = idx / plane;  // agent-major: img = agent * b_sz + batch
    const int agent = img / b_sz, bat = img % b_sz;
    const float* xin = x + (static_cast<size_t>(bat) * c_total + c_first + 3 * agent) * plane;
    float in[27];
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int ih = oh + kh - 1, iw = ow + kw - 1;
          in[ci * 9 + kh * 3 + kw] =
              (ih >= 0 && ih < h && iw >= 0 && iw < wpx) ? __ldg(xin + ci * plane + static_cast<size_t>(ih) * wpx + iw) : 0.f;
        }
    __nv_bfloat16* ypix = y + idx * (static_cast<size_t>(cout) * planes);
    for (int g = 0; g < cout; g += GROUP) {
      float acc[GROUP];
#pragma unroll
      for (int j = 0; j < GROUP; ++j) acc[j] = 0.f;
#pragma unroll
      for (int k = 0; k < 27; ++k) {
        const float4* wr = reinterpret_cast<const float4*>(s_w + k * cout + g);
#pragma unroll
        for (int j4 = 0; j4 < GROUP / 4; ++j4) {
          const float4 wv = wr[j4];
          acc[4 * j4 + 0] = fmaf(in[k], wv.x, acc[4 * j4 + 0]);
          acc[4 * j4 + 1] = fmaf(in[k], wv.y, acc[4 * j4 + 1]);
          acc[4 * j4 + 2] = fmaf(in[k], wv.z, acc[4 * j4 + 2]);
          acc[4 * j4 + 3] = fmaf(in[k], wv.w, acc[4 * j4 + 3]);
        }
      }
#pragma unroll
      for (int j8 = 0; j8 < GROUP / 8; ++j8) {
        uint4 hv, lv;
        __nv_bfloat162* hb = reinterpret_cast<__nv_bfloat162*>(&hv);
        __nv_bfloat162* lb = reinterpret_cast<__nv_bfloat162*>(&lv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = g + j8 * 8 + 2 * j;
          const float a = fmaxf(fmaf(acc[j8 * 8 + 2 * j], s_scale[c], s_shift[c]), 0.f);
          const float b = fmaxf(fmaf(acc[j8 * 8 + 2 * j + 1], s_scale[c + 1], s_shift[c + 1]), 0.f);
          hb[j] = __floats2bfloat162_rn(a, b);
          const float2 hf = __bfloat1622float2(hb[j]);
          lb[j] = __floats2bfloat162_rn(a - hf.x, b - hf.y);
        }
        *reinterpret_cast<uint4*>(ypix + g + j8 * 8) = hv;
        if (planes == 2) *reinterpret_cast<uint4*>(ypix + cout + g + j8 * 8) = lv;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ pool / upsample
// One thread per (output pixel, 8-channel group): nine 16-byte loads, packed bf16 max (exact), one 16-byte store.
// (The first version - one thread per element, nine 2-byte loads - took 0.46 ms per 40 frames against a 0.065 ms HBM
// floor.) In the hi|lo storage the value is hi + lo: the maximum is taken on the sum and its (hi, lo) pair kept.
__global__ void __launch_bounds__(256) maxpool3x3s2_kernel(const __nv_bfloat16* __restrict__ x,
                                                           __nv_bfloat16* __restrict__ y, int n, int h, int wpx, int c,
                                                           int act) {
  const int ho = h / 2, wo = wpx / 2;
  const int planes = act_planes(act);
  const int c8 = c / 8;
  const size_t total = static_cast<size_t>(n) * ho * wo * c8;
  const size_t pixs = static_cast<size_t>(c) * planes;  // elements per pixel
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cg = idx % c8;
    size_t t = idx / c8;
    const int ow = t % wo;
    t /= wo;
    const int oh = t % ho;
    const int img = t / ho;
    __nv_bfloat162 best[4], best_lo[4];
    float bv[8];
    bool first = true;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = oh * 2 + kh - 1;
      if (ih < 0 || ih >= h) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int iw = ow * 2 + kw - 1;
        if (iw < 0 || iw >= wpx) continue;
        const __nv_bfloat16* pix = x + ((static_cast<size_t>(img) * h + ih) * wpx + iw) * pixs + cg * 8;
        const uint4 hv = __ldg(reinterpret_cast<const uint4*>(pix));
        const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&hv);
        if (planes == 1 && act_is_f16(act)) {
          const __half2* hh = reinterpret_cast<const __half2*>(&hv);
          __half2* bh = reinterpret_cast<__half2*>(best);
#pragma unroll
          for (int e = 0; e < 4; ++e) bh[e] = first ? hh[e] : __hmax2(bh[e], hh[e]);
        } else if (planes == 1) {
#pragma unroll
          for (int e = 0; e < 4; ++e) best[e] = first ? hb[e] : __hmax2(best[e], hb[e]);
        } else {
          const uint4 lv = __ldg(reinterpret_cast<const uint4*>(pix + c));
          const __nv_bfloat162* lb = reinterpret_cast<const __nv_bfloat162*>(&lv);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 hf = unpack_act2(*reinterpret_cast<const uint32_t*>(&hb[e]), act_is_f16(act));
            const float2 lf = unpack_act2(*reinterpret_cast<const uint32_t*>(&lb[e]), act_is_f16(act));
            const float v0 = hf.x + lf.x, v1 = hf.y + lf.y;
            const bool t0 = first || v0 > bv[2 * e], t1 = first || v1 > bv[2 * e + 1];
            if (t0) bv[2 * e] = v0;
            if (t1) bv[2 * e + 1] = v1;
            best[e] = __halves2bfloat162(t0 ? __low2bfloat16(hb[e]) : __low2bfloat16(best[e]),
                                         t1 ? __high2bfloat16(hb[e]) : __high2bfloat16(best[e]));
            best_lo[e] = __halves2bfloat162(t0 ? __low2bfloat16(lb[e]) : __low2bfloat16(best_lo[e]),
                                            t1 ? __high2bfloat16(lb[e]) : __high2bfloat16(best_lo[e]));
          }
        }
        first = false;
      }
    }
    __nv_bfloat16* dst = y + (idx / c8) * pixs + cg * 8;
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(best);
    if (planes == 2) *reinterpret_cast<uint4*>(dst + c) = *reinterpret_cast<const uint4*>(best_lo);
  }
}

// Four consecutive output columns per thread (one 16-byte store); the row interpolation terms are shared.
__global__ void __launch_bounds__(256) bilinear_up_kernel(const float* __restrict__ x, float* __restrict__ y, int nc,
                                                         int h, int wpx, int factor) {
  const int ho = h * factor, wo = wpx * factor;  // wo % 4 == 0 is checked on the host
  const int wo4 = wo / 4;
  const float rs = 1.f / static_cast<float>(factor);
  const size_t total = static_cast<size_t>(nc) * ho * wo4;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ow0 = (idx % wo4) * 4;
    const int oh = (idx / wo4) % ho;
    const size_t pl = idx / (static_cast<size_t>(ho) * wo4);
    float sh = (oh + 0.5f) * rs - 0.5f;
    sh = sh < 0.f ? 0.f : sh;
    const int h0 = static_cast<int>(sh);
    const int h1 = h0 + (h0 < h - 1 ? 1 : 0);
    const float lh = sh - h0;
    const float* r0 = x + pl * static_cast<size_t>(h) * wpx + static_cast<size_t>(h0) * wpx;
    const float* r1 = x + pl * static_cast<size_t>(h) * wpx + static_cast<size_t>(h1) * wpx;
    float o[4];
    // the four columns usually fall into one source cell (always for factors that are multiples of 8: the cell
    // boundaries sit at ow = factor/2 mod factor): one set of four corner loads then serves all of them
    float swj[4];
    int w0j[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float sw = (ow0 + j + 0.5f) * rs - 0.5f;
      swj[j] = sw < 0.f ? 0.f : sw;
      w0j[j] = static_cast<int>(swj[j]);
    }
    const bool one_cell = w0j[0] == w0j[3];
    float c00 = 0.f, c01 = 0.f, c10 = 0.f, c11 = 0.f;
    if (one_cell) {
      const int w0 = w0j[0], w1 = w0 + (w0 < wpx - 1 ? 1 : 0);
      c00 = __ldg(r0 + w0), c01 = __ldg(r0 + w1), c10 = __ldg(r1 + w0), c11 = __ldg(r1 + w1);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int w0 = w0j[j];
      const float lw = swj[j] - w0;
      float v00 = c00, v01 = c01, v10 = c10, v11 = c11;
      if (!one_cell) {
        const int w1 = w0 + (w0 < wpx - 1 ? 1 : 0);
        v00 = __ldg(r0 + w0), v01 = __ldg(r0 + w1), v10 = __ldg(r1 + w0), v11 = __ldg(r1 + w1);
      }
      // same evaluation order as ATen's upsample_bilinear2d: h0lambda*(w0lambda*v00 + w1lambda*v01) + h1lambda*(...)
      o[j] = (1.f - lh) * ((1.f - lw) * v00 + lw * v01) + lh * ((1.f - lw) * v10 + lw * v11);
    }
    *reinterpret_cast<float4*>(y + (pl * ho + oh) * static_cast<size_t>(wo) + ow0) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// The same up-sampling with the arg-max over the channels taken per output pixel (first maximal index) and ONLY the
// uint8 label written: the evaluation loop of a simple_decoder model needs neither the small logits up-sampled in HBM
// (461 MB per 40 frames at 512x512) nor a second pass over them.
__global__ void __launch_bounds__(256) bilinear_argmax_kernel(const float* __restrict__ x, uint8_t* __restrict__ labels,
                                                             int n, int c, int h, int wpx, int factor) {
  const int ho = h * factor, wo = wpx * factor;
  const float rs = 1.f / static_cast<float>(factor);
  const size_t total = static_cast<size_t>(n) * ho * wo;
  const size_t plane = static_cast<size_t>(h) * wpx;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ow = idx % wo;
    const int oh = (idx / wo) % ho;
    const size_t img = idx / (static_cast<size_t>(ho) * wo);
    float sh = (oh + 0.5f) * rs - 0.5f, sw = (ow + 0.5f) * rs - 0.5f;
    sh = sh < 0.f ? 0.f : sh;
    sw = sw < 0.f ? 0.f : sw;
    const int h0 = static_cast<int>(sh), w0 = static_cast<int>(sw);
    const int h1 = h0 + (h0 < h - 1 ? 1 : 0), w1 = w0 + (w0 < wpx - 1 ? 1 : 0);
    const float lh = sh - h0, lw = sw - w0;
    const float* xp = x + img * c * plane;
    float best = 0.f;
    int arg = 0;
    for (int k = 0; k < c; ++k, xp += plane) {
      const float v00 = __ldg(xp + h0 * wpx + w0), v01 = __ldg(xp + h0 * wpx + w1);
      const float v10 = __ldg(xp + h1 * wpx + w0), v11 = __ldg(xp + h1 * wpx + w1);
      const float v = (1.f - lh) * ((1.f - lw) * v00 + lw * v01) + lh * ((1.f - lw) * v10 + lw * v11);  // as above
      if (k == 0 || v > best) best = v, arg = k;
    }
    labels[idx] = static_cast<uint8_t>(arg);
  }
}

// ------------------------------------------------------------------------------------------ layout helpers
__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int n, int h, int wpx,
                                    int c, int cstride, int coffset, int act) {
  const int planes = act_planes(act);
  const size_t total = static_cast<size_t>(n) * c * h * wpx;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ow = idx % wpx;
    size_t t = idx / wpx;
    const int oh = t % h;
    t /= h;
    const int ch = t % c;
    const int img = t / c;
    const __nv_bfloat16* pix = x + ((static_cast<size_t>(img) * h + oh) * wpx + ow) * (static_cast<size_t>(cstride) * planes) + coffset;
    y[idx] = act_load(pix, ch, cstride, act);
  }
}
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int n, int h, int wpx,
                                    int c, int cstride, int coffset, int act) {
  const int planes = act_planes(act);
  const size_t total = static_cast<size_t>(n) * c * h * wpx;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ch = idx % c;
    size_t t = idx / c;
    const int ow = t % wpx;
    t /= wpx;
    const int oh = t % h;
    const int img = t / h;
    const float v = x[((static_cast<size_t>(img) * c + ch) * h + oh) * wpx + ow];
    __nv_bfloat16* pix = y + ((static_cast<size_t>(img) * h + oh) * wpx + ow) * (static_cast<size_t>(cstride) * planes) + coffset;
    act_store(pix, ch, cstride, act, v);
  }
}


// ------------------------------------------------------------------------------------------ indexed image copy
// dst image (g*b + i), channels [dst_coffset, +c)  <-  src image (sel[g]*b + i), channels [src_coffset, +c), NHWC in
// units of 8 channels (16 bytes). sel lives on the DEVICE, so a captured program stays static while the host redraws
// the selection per forward (the random-selection baselines, agent.py:447-452,934-947). sel == NULL: identity.
__global__ void __launch_bounds__(256) gather_images_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst,
                                                            const int* __restrict__ sel, int n_groups, int b,
                                                            size_t px, int c8, int planes, int scs8, int sco8,
                                                            int dcs8, int dco8) {
  const size_t per_img = px * planes * c8;
  const size_t total = static_cast<size_t>(n_groups) * b * per_img;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cc = idx % c8;
    size_t t = idx / c8;
    const int pl = t % planes;
    t /= planes;
    const size_t pix = t % px;
    const size_t img = t / px;  // destination image
    const int g = img / b, i = img % b;
    const size_t simg = static_cast<size_t>(sel ? sel[g] : g) * b + i;
    dst[(img * px + pix) * planes * dcs8 + pl * dcs8 + dco8 + cc] =
        src[(simg * px + pix) * planes * scs8 + pl * scs8 + sco8 + cc];
  }
}

// ------------------------------------------------------------------------------------------ eval-loop glue
// labels[n][p] = argmax_c logits[n][c][p], first maximal index (torch.max(1)[1], trainer.py:804); one thread per
// pixel, coalesced plane reads.
__global__ void __launch_bounds__(256) argmax_labels_kernel(const float* __restrict__ logits,
                                                            uint8_t* __restrict__ labels, int n, int c, size_t hw) {
  const size_t total = static_cast<size_t>(n) * hw;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t img = i / hw, p = i % hw;
    const float* src = logits + img * c * hw + p;
    float best = src[0];
    int arg = 0;
    for (int k = 1; k < c; ++k) {
      const float v = src[k * hw];
      if (v > best) best = v, arg = k;
    }
    labels[i] = static_cast<uint8_t>(arg);
  }
}

// runningScore._fast_hist (metrics.py:99-104): hist[n_class * gt + pred] += 1 over pixels with 0 <= gt < n_class.
// Per-CTA shared-memory histogram, flushed with 64-bit global atomics.
template <typename GT>
__global__ void __launch_bounds__(256) confusion_kernel(const uint8_t* __restrict__ pred, const GT* __restrict__ gt,
                                                        size_t total, int n_class,
                                                        unsigned long long* __restrict__ hist) {
  extern __shared__ unsigned int s_hist[];
  const int bins = n_class * n_class;
  for (int i = threadIdx.x; i < bins; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const long long g = static_cast<long long>(gt[i]);
    const int pr = pred[i];
    if (g >= 0 && g < n_class && pr < n_class) atomicAdd(&s_hist[static_cast<int>(g) * n_class + pr], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bins; i += blockDim.x)
    if (s_hist[i]) atomicAdd(&hist[i], static_cast<unsigned long long>(s_hist[i]));
}

// runningScore.update_div (metrics.py:70-97): the same histogram routed per IMAGE into one of two matrices by a
// per-image flag (1 = "normal" image -> hist_pos, 0 = noisy -> hist_neg). One CTA never straddles two images.
template <typename GT>
__global__ void __launch_bounds__(256) confusion_div_kernel(const uint8_t* __restrict__ pred, const GT* __restrict__ gt,
                                                            const uint8_t* __restrict__ img_flag, int n_img,
                                                            size_t px_per_img, int ctas_per_img, int n_class,
                                                            unsigned long long* __restrict__ hist_pos,
                                                            unsigned long long* __restrict__ hist_neg) {
  extern __shared__ unsigned int s_hist[];
  const int bins = n_class * n_class;
  for (int i = threadIdx.x; i < bins; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  const int img = blockIdx.x / ctas_per_img, part = blockIdx.x % ctas_per_img;
  if (img >= n_img) return;
  const size_t base = static_cast<size_t>(img) * px_per_img;
  for (size_t i = static_cast<size_t>(part) * blockDim.x + threadIdx.x; i < px_per_img;
       i += static_cast<size_t>(ctas_per_img) * blockDim.x) {
    const long long g = static_cast<long long>(gt[base + i]);
    const int pr = pred[base + i];
    if (g >= 0 && g < n_class && pr < n_class) atomicAdd(&s_hist[static_cast<int>(g) * n_class + pr], 1u);
  }
  __syncthreads();
  unsigned long long* hist = img_flag[img] ? hist_pos : hist_neg;
  for (int i = threadIdx.x; i < bins; i += blockDim.x)
    if (s_hist[i]) atomicAdd(&hist[i], static_cast<unsigned long long>(s_hist[i]));
}

// runningScore.update_selection (metrics.py:23-68) on the device. counters = {total_agent, correct_when2com,
// correct_who2com}.
//   mode 0 'mimo'      action int64 [B][N], commun_label int64 [B][2][N]  (row 0: needs communication, row 1: with whom)
//   mode 1 'when2com'  action int64 [B] (arg-max link), commun_label int64 [B] in -1..N-2
//   mode 2 'when2com'  action fp32 [B][N] (thresholded weights, 'activated'), commun_label int64 [B]
__global__ void selection_kernel(const void* __restrict__ action, const long long* __restrict__ label, int b_sz, int n,
                                 int mode, unsigned long long* __restrict__ counters) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long when_ok = 0, who_ok = 0, agents = 0;
  if (mode == 0) {
    if (i < b_sz * n) {
      const int b = i / n, j = i % n;
      const long long need = label[(static_cast<size_t>(b) * 2 + 0) * n + j];
      const long long who = label[(static_cast<size_t>(b) * 2 + 1) * n + j];
      const long long act = static_cast<const long long*>(action)[i];
      agents = 1;
      // when2com: (action != own id) == uint8(need)   (commun_label[:,0,:].type(ByteTensor))
      when_ok = (static_cast<unsigned char>(act != j) == static_cast<unsigned char>(need)) ? 1 : 0;
      who_ok = (act == who * need + static_cast<long long>(j) * (1 - need)) ? 1 : 0;
    }
  } else if (i < b_sz) {
    const long long lab = label[i] + 1;  // -1,0,1,.. -> 0,1,2,..
    agents = 1;
    if (mode == 1) {
      const long long act = static_cast<const long long*>(action)[i];
      when_ok = ((act == 0) == (lab == 0)) ? 1 : 0;
      who_ok = (act == lab) ? 1 : 0;
    } else {
      const float* row = static_cast<const float*>(action) + static_cast<size_t>(i) * n;
      bool pred_comm = false;
      for (int l = 0; l < n; ++l)
        if (row[l] > 0.2f) {
          if (l == lab) ++who_ok;
          if (l != 0) pred_comm = true;
        }
      // the reference compares an int8 "communicates" flag with the (label == 0) mask as written (metrics.py:44)
      when_ok = (pred_comm == (lab == 0)) ? 1 : 0;
    }
  }
  if (agents) atomicAdd(&counters[0], agents);
  if (when_ok) atomicAdd(&counters[1], when_ok);
  if (who_ok) atomicAdd(&counters[2], who_ok);
}

}  // namespace
}  // namespace w2c

using namespace w2c;

extern "C" {

int w2c_version(void) { return 100; }
const char* w2c_last_error(void) { return g_err; }
uint64_t w2c_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int32_t w2c_cout_pad(int32_t cout) { return (cout + 15) / 16 * 16; }
size_t w2c_packed_weight_bytes(int32_t cout, int32_t cin, int32_t ntaps, int32_t act) {
  return static_cast<size_t>(w2c_cout_pad(cout)) * ntaps * cin * 2 * (act_planes(act));
}

int w2c_pack_conv_weight(const float* w, int32_t cout, int32_t cin_real, int32_t cin, int32_t ntaps,
                         int32_t transposed, int32_t act, void* packed, w2c_stream_t stream) {
  W2C_CHECK_ARG(w && packed, "pack: null pointer");
  W2C_CHECK_ARG(cout > 0 && cin_real > 0 && cin >= cin_real && cin % 64 == 0, "pack: bad channels %d/%d/%d", cout,
                cin_real, cin);
  W2C_CHECK_ARG(ntaps == 9 || ntaps == 1, "pack: ntaps must be 9 or 1");
  const int planes = act_planes(act);
  const int cout_pad = w2c_cout_pad(cout);
  const size_t total = static_cast<size_t>(cout_pad) * ntaps * cin;
  pack_weight_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, cout, cin_real, cin, ntaps, transposed, planes, cout_pad, static_cast<__nv_bfloat16*>(packed),
      act_is_f16(act));
  W2C_CHECK_LAUNCH("pack_weight_kernel");
  return W2C_OK;
}


int w2c_fold_bn(const float* conv_bias, const float* gamma, const float* beta, const float* mean, const float* var,
                float eps, int32_t cout, float* scale, float* shift, w2c_stream_t stream) {
  W2C_CHECK_ARG(scale && shift && cout > 0, "fold_bn: bad arguments");
  const bool any = gamma || beta || mean || var;
  W2C_CHECK_ARG(!any || (gamma && beta && mean && var), "fold_bn: BN tensors must be all present or all NULL");
  fold_bn_kernel<<<ceil_div(cout, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(conv_bias, gamma, beta, mean, var,
                                                                                     eps, cout, scale, shift);
  W2C_CHECK_LAUNCH("fold_bn_kernel");
  return W2C_OK;
}

int w2c_stem_conv3x3_fwd(const float* x, const float* w, const float* scale, const float* shift, void* y, int32_t b,
                         int32_t n_agents, int32_t c_total, int32_t c_first, int32_t h, int32_t w_px, int32_t cout,
                         int32_t act, int32_t n_split, w2c_stream_t stream) {
  W2C_CHECK_ARG(n_split == 1 || (n_split == 2 && cout == 128), "stem3x3: n_split=%d needs cout=128", n_split);
  W2C_CHECK_ARG(x && w && scale && shift && y, "stem3x3: null pointer");
  W2C_CHECK_ARG(b > 0 && n_agents > 0 && h > 0 && w_px > 0, "stem3x3: bad extent");
  W2C_CHECK_ARG(c_first >= 0 && c_first + 3 * n_agents <= c_total, "stem3x3: channel window [%d, %d) outside %d",
                c_first, c_first + 3 * n_agents, c_total);
  W2C_CHECK_ARG(cout % 32 == 0 && cout <= 128, "stem3x3: cout=%d must be a multiple of 32 and <= 128", cout);
  if (cout == 64 || cout == 128)
    return stem3x3_tc_forward(x, w, scale, shift, y, b, n_agents, c_total, c_first, h, w_px, cout, act, n_split,
                              static_cast<cudaStream_t>(stream));
  const size_t total = static_cast<size_t>(b) * n_agents * h * w_px;
  const size_t smem = (27 * cout + 2 * cout) * sizeof(float);
  stem3x3_kernel<32><<<grid_for(total, 256, 148 * 8), 256, smem, static_cast<cudaStream_t>(stream)>>>(
      x, w, scale, shift, static_cast<__nv_bfloat16*>(y), b, n_agents, c_total, c_first, h, w_px, cout, act);
  W2C_CHECK_LAUNCH("stem3x3_kernel");
  return W2C_OK;
}

int w2c_stem_conv3x3_u8_fwd(const uint8_t* frames, const float* lut, const float* w, const float* scale,
                            const float* shift, void* y, int32_t b, int32_t n_agents, int32_t agents_total,
                            int32_t agent_first, int32_t h, int32_t w_px, int32_t cout, int32_t act, int32_t n_split,
                            w2c_stream_t stream) {
  W2C_CHECK_ARG(n_split == 1 || (n_split == 2 && cout == 128), "stem3x3_u8: n_split=%d needs cout=128", n_split);
  W2C_CHECK_ARG(frames && lut && w && scale && shift && y, "stem3x3_u8: null pointer");
  W2C_CHECK_ARG(b > 0 && n_agents > 0 && h > 0 && w_px > 0, "stem3x3_u8: bad extent");
  W2C_CHECK_ARG(agent_first >= 0 && agent_first + n_agents <= agents_total, "stem3x3_u8: agents [%d, %d) outside %d",
                agent_first, agent_first + n_agents, agents_total);
  W2C_CHECK_ARG(cout == 64 || cout == 128, "stem3x3_u8: cout=%d (64 or 128)", cout);
  return stem3x3_tc_u8_forward(frames, lut, w, scale, shift, y, b, n_agents, agents_total, agent_first, h, w_px, cout,
                               act, n_split, static_cast<cudaStream_t>(stream));
}

int w2c_argmax_labels_fwd(const float* logits, uint8_t* labels, int32_t n, int32_t c, int64_t hw,
                          w2c_stream_t stream) {
  W2C_CHECK_ARG(logits && labels && n > 0 && hw > 0, "argmax_labels: bad arguments");
  W2C_CHECK_ARG(c > 0 && c <= 256, "argmax_labels: c=%d does not fit a uint8 label", c);
  const size_t total = static_cast<size_t>(n) * hw;
  argmax_labels_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, labels, n, c,
                                                                                           static_cast<size_t>(hw));
  W2C_CHECK_LAUNCH("argmax_labels_kernel");
  return W2C_OK;
}

int w2c_confusion_update(const uint8_t* pred, const void* gt, int32_t gt_dtype, int64_t count, int32_t n_class,
                         int64_t* hist, w2c_stream_t stream) {
  W2C_CHECK_ARG(pred && gt && hist && count > 0, "confusion: bad arguments");
  W2C_CHECK_ARG(n_class > 0 && n_class <= 64, "confusion: n_class=%d (1..64)", n_class);
  W2C_CHECK_ARG(gt_dtype == W2C_GT_U8 || gt_dtype == W2C_GT_I64, "confusion: gt_dtype=%d", gt_dtype);
  const size_t smem = static_cast<size_t>(n_class) * n_class * sizeof(unsigned int);
  const int grid = grid_for(static_cast<size_t>(count), 256 * 16, 148 * 8);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  unsigned long long* h = reinterpret_cast<unsigned long long*>(hist);
  if (gt_dtype == W2C_GT_U8)
    confusion_kernel<uint8_t><<<grid, 256, smem, s>>>(pred, static_cast<const uint8_t*>(gt), count, n_class, h);
  else
    confusion_kernel<long long><<<grid, 256, smem, s>>>(pred, static_cast<const long long*>(gt), count, n_class, h);
  W2C_CHECK_LAUNCH("confusion_kernel");
  return W2C_OK;
}

int w2c_confusion_update_div(const uint8_t* pred, const void* gt, int32_t gt_dtype, const uint8_t* img_flag,
                             int32_t n_img, int64_t px_per_img, int32_t n_class, int64_t* hist_pos, int64_t* hist_neg,
                             w2c_stream_t stream) {
  W2C_CHECK_ARG(pred && gt && img_flag && hist_pos && hist_neg && n_img > 0 && px_per_img > 0,
                "confusion_div: bad arguments");
  W2C_CHECK_ARG(n_class > 0 && n_class <= 64, "confusion_div: n_class=%d (1..64)", n_class);
  W2C_CHECK_ARG(gt_dtype == W2C_GT_U8 || gt_dtype == W2C_GT_I64, "confusion_div: gt_dtype=%d", gt_dtype);
  const size_t smem = static_cast<size_t>(n_class) * n_class * sizeof(unsigned int);
  int ctas_per_img = static_cast<int>((px_per_img + 256 * 16 - 1) / (256 * 16));
  const int cap = (device_sm_count() * 8 + n_img - 1) / n_img;
  if (ctas_per_img > cap) ctas_per_img = cap;
  if (ctas_per_img < 1) ctas_per_img = 1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  unsigned long long* hp = reinterpret_cast<unsigned long long*>(hist_pos);
  unsigned long long* hn = reinterpret_cast<unsigned long long*>(hist_neg);
  if (gt_dtype == W2C_GT_U8)
    confusion_div_kernel<uint8_t><<<n_img * ctas_per_img, 256, smem, s>>>(
        pred, static_cast<const uint8_t*>(gt), img_flag, n_img, static_cast<size_t>(px_per_img), ctas_per_img, n_class, hp, hn);
  else
    confusion_div_kernel<long long><<<n_img * ctas_per_img, 256, smem, s>>>(
        pred, static_cast<const long long*>(gt), img_flag, n_img, static_cast<size_t>(px_per_img), ctas_per_img, n_class, hp, hn);
  W2C_CHECK_LAUNCH("confusion_div_kernel");
  return W2C_OK;
}

int w2c_selection_update(const void* action, const int64_t* commun_label, int32_t b_sz, int32_t n, int32_t mode,
                         int64_t* counters, w2c_stream_t stream) {
  W2C_CHECK_ARG(action && commun_label && counters && b_sz > 0 && n > 0, "selection: bad arguments");
  W2C_CHECK_ARG(mode >= 0 && mode <= 2, "selection: mode=%d (0 mimo, 1 when2com arg-max, 2 when2com weights)", mode);
  const int total = mode == 0 ? b_sz * n : b_sz;
  selection_kernel<<<(total + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      action, reinterpret_cast<const long long*>(commun_label), b_sz, n, mode,
      reinterpret_cast<unsigned long long*>(counters));
  W2C_CHECK_LAUNCH("selection_kernel");
  return W2C_OK;
}

int w2c_gather_images_fwd(const void* src, void* dst, const int32_t* sel, int32_t n_groups, int32_t b, int32_t h,
                          int32_t w_px, int32_t c, int32_t src_cstride, int32_t src_coffset, int32_t dst_cstride,
                          int32_t dst_coffset, int32_t act, w2c_stream_t stream) {
  W2C_CHECK_ARG(src && dst && n_groups > 0 && b > 0 && h > 0 && w_px > 0 && c > 0, "gather_images: bad arguments");
  W2C_CHECK_ARG(c % 8 == 0 && src_cstride % 8 == 0 && src_coffset % 8 == 0 && dst_cstride % 8 == 0 && dst_coffset % 8 == 0,
                "gather_images: channel counts, strides and offsets must be multiples of 8");
  W2C_CHECK_ARG(src_coffset + c <= src_cstride && dst_coffset + c <= dst_cstride, "gather_images: slice out of range");
  const int planes = act_planes(act);
  const size_t px = static_cast<size_t>(h) * w_px;
  const size_t total = static_cast<size_t>(n_groups) * b * px * planes * (c / 8);
  gather_images_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(src), static_cast<uint4*>(dst), sel, n_groups, b, px, c / 8, planes, src_cstride / 8,
      src_coffset / 8, dst_cstride / 8, dst_coffset / 8);
  W2C_CHECK_LAUNCH("gather_images_kernel");
  return W2C_OK;
}

int w2c_stem_conv7x7s2_fwd(const float* x, const float* w, const float* scale, const float* shift, void* y, int32_t b,
                           int32_t n_agents, int32_t c_total, int32_t c_first, int32_t h, int32_t w_px, int32_t cout,
                           int32_t act, int32_t n_split, w2c_stream_t stream) {
  W2C_CHECK_ARG(x && w && scale && shift && y, "stem7x7: null pointer");
  W2C_CHECK_ARG(c_first >= 0 && c_first + 3 * n_agents <= c_total, "stem7x7: channel window outside the input");
  W2C_CHECK_ARG(b > 0 && n_agents > 0 && h > 0 && w_px > 0 && h % 2 == 0 && w_px % 2 == 0, "stem7x7: bad extent");
  W2C_CHECK_ARG((cout == 64 && n_split == 1) || (cout == 128 && (n_split == 1 || n_split == 2)),
                "stem7x7: cout=%d n_split=%d (64, or 128 as one or two maps)", cout, n_split);
  return stem7x7_tc_forward(x, nullptr, w, scale, shift, y, b, n_agents, c_total, c_first, h, w_px, cout, act, n_split,
                            0, static_cast<cudaStream_t>(stream));
}

// The two first layers WITHOUT the ReLU (and normally with scale = 1, shift = conv bias): the raw conv output the
// train-mode BatchNorm (csrc/bn_train.cu) takes its batch statistics from.
int w2c_stem_conv3x3_raw_fwd(const float* x, const float* w, const float* scale, const float* shift, void* y, int32_t b,
                             int32_t n_agents, int32_t c_total, int32_t c_first, int32_t h, int32_t w_px, int32_t cout,
                             int32_t act, int32_t n_split, w2c_stream_t stream) {
  W2C_CHECK_ARG(x && w && scale && shift && y, "stem3x3_raw: null pointer");
  W2C_CHECK_ARG(act_valid(act), "stem3x3_raw: bad act %d", act);
  W2C_CHECK_ARG(b > 0 && n_agents > 0 && h > 0 && w_px > 0, "stem3x3_raw: bad extent");
  W2C_CHECK_ARG(c_first >= 0 && c_first + 3 * n_agents <= c_total, "stem3x3_raw: channel window outside the input");
  W2C_CHECK_ARG(cout == 64 || cout == 128, "stem3x3_raw: cout=%d (64 or 128)", cout);
  return stem3x3_tc_forward(x, w, scale, shift, y, b, n_agents, c_total, c_first, h, w_px, cout, act | 0x100, n_split,
                            static_cast<cudaStream_t>(stream));
}

int w2c_stem_conv7x7s2_raw_fwd(const float* x, const float* w, const float* scale, const float* shift, void* y,
                               int32_t b, int32_t n_agents, int32_t c_total, int32_t c_first, int32_t h, int32_t w_px,
                               int32_t cout, int32_t act, int32_t n_split, w2c_stream_t stream) {
  W2C_CHECK_ARG(x && w && scale && shift && y, "stem7x7_raw: null pointer");
  W2C_CHECK_ARG(act_valid(act), "stem7x7_raw: bad act %d", act);
  W2C_CHECK_ARG(c_first >= 0 && c_first + 3 * n_agents <= c_total, "stem7x7_raw: channel window outside the input");
  W2C_CHECK_ARG(b > 0 && n_agents > 0 && h > 0 && w_px > 0 && h % 2 == 0 && w_px % 2 == 0, "stem7x7_raw: bad extent");
  return stem7x7_tc_forward(x, nullptr, w, scale, shift, y, b, n_agents, c_total, c_first, h, w_px, cout, act | 0x100,
                            n_split, 0, static_cast<cudaStream_t>(stream));
}

int w2c_stem_conv7x7s2_u8_fwd(const uint8_t* frames, const float* lut, const float* w, const float* scale,
                              const float* shift, void* y, int32_t b, int32_t n_agents, int32_t agents_total,
                              int32_t agent_first, int32_t h, int32_t w_px, int32_t cout, int32_t act, int32_t n_split,
                              w2c_stream_t stream) {
  W2C_CHECK_ARG(frames && lut && w && scale && shift && y, "stem7x7_u8: null pointer");
  W2C_CHECK_ARG(agent_first >= 0 && agent_first + n_agents <= agents_total, "stem7x7_u8: agent window outside the input");
  W2C_CHECK_ARG(b > 0 && n_agents > 0 && h > 0 && w_px > 0 && h % 2 == 0 && w_px % 2 == 0, "stem7x7_u8: bad extent");
  W2C_CHECK_ARG((cout == 64 && n_split == 1) || (cout == 128 && (n_split == 1 || n_split == 2)),
                "stem7x7_u8: cout=%d n_split=%d", cout, n_split);
  return stem7x7_tc_forward(frames, lut, w, scale, shift, y, b, n_agents, agents_total, agent_first, h, w_px, cout, act,
                            n_split, 1, static_cast<cudaStream_t>(stream));
}

int w2c_maxpool3x3s2_fwd(const void* x, void* y, int32_t n, int32_t h, int32_t w_px, int32_t c, int32_t act,
                         w2c_stream_t stream) {
  W2C_CHECK_ARG(x && y && n > 0 && h > 0 && w_px > 0 && c > 0 && h % 2 == 0 && w_px % 2 == 0, "maxpool: bad arguments");
  W2C_CHECK_ARG(c % 8 == 0, "maxpool: c=%d must be a multiple of 8", c);
  const size_t total = static_cast<size_t>(n) * (h / 2) * (w_px / 2) * (c / 8);
  maxpool3x3s2_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), n, h, w_px, c, act);
  W2C_CHECK_LAUNCH("maxpool3x3s2_kernel");
  return W2C_OK;
}

int w2c_bilinear_up_fwd(const float* x, float* y, int32_t n, int32_t c, int32_t h, int32_t w_px, int32_t factor,
                        w2c_stream_t stream) {
  W2C_CHECK_ARG(x && y && n > 0 && c > 0 && h > 0 && w_px > 0 && factor >= 1, "bilinear: bad arguments");
  W2C_CHECK_ARG((w_px * factor) % 4 == 0, "bilinear: output width %d must be a multiple of 4", w_px * factor);
  const size_t total = static_cast<size_t>(n) * c * h * factor * (w_px * factor / 4);
  bilinear_up_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, n * c, h, w_px, factor);
  W2C_CHECK_LAUNCH("bilinear_up_kernel");
  return W2C_OK;
}

int w2c_bilinear_argmax_fwd(const float* x, uint8_t* labels, int32_t n, int32_t c, int32_t h, int32_t w_px,
                            int32_t factor, w2c_stream_t stream) {
  W2C_CHECK_ARG(x && labels && n > 0 && c > 0 && c <= 256 && h > 0 && w_px > 0 && factor >= 1, "bilinear_argmax: bad arguments");
  const size_t total = static_cast<size_t>(n) * h * factor * w_px * factor;
  bilinear_argmax_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, labels, n, c, h, w_px,
                                                                                             factor);
  W2C_CHECK_LAUNCH("bilinear_argmax_kernel");
  return W2C_OK;
}

int w2c_nhwc_to_nchw_f32(const void* x, float* y, int32_t n, int32_t h, int32_t w_px, int32_t c, int32_t cstride,
                         int32_t coffset, int32_t act, w2c_stream_t stream) {
  W2C_CHECK_ARG(x && y && n > 0 && h > 0 && w_px > 0 && c > 0, "nhwc_to_nchw: bad arguments");
  if (cstride <= 0) cstride = c;
  const size_t total = static_cast<size_t>(n) * c * h * w_px;
  nhwc_to_nchw_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), y, n, h, w_px, c, cstride, coffset, act);
  W2C_CHECK_LAUNCH("nhwc_to_nchw_kernel");
  return W2C_OK;
}

int w2c_nchw_f32_to_nhwc(const float* x, void* y, int32_t n, int32_t h, int32_t w_px, int32_t c, int32_t cstride,
                         int32_t coffset, int32_t act, w2c_stream_t stream) {
  W2C_CHECK_ARG(x && y && n > 0 && h > 0 && w_px > 0 && c > 0, "nchw_to_nhwc: bad arguments");
  if (cstride <= 0) cstride = c;
  const size_t total = static_cast<size_t>(n) * c * h * w_px;
  nchw_to_nhwc_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<__nv_bfloat16*>(y), n, h, w_px, c, cstride, coffset, act);
  W2C_CHECK_LAUNCH("nchw_to_nhwc_kernel");
  return W2C_OK;
}

}  // extern "C"
