// Key / query heads (km_generator / linear, agent.py:145-178): flatten -> Linear(n_feat,256) -> ReLU ->
// Linear(256,128) -> ReLU -> Linear(128,out).  M = agents*scenes rows (tens), so this is weight-bandwidth work
// (fc0 is 256 x n_feat fp32 = 4 MB at 512x512): two launches for up to two heads reading the same feature map
// (key_net and query_net, agent.py:1132-1147, are independent: run side by side they halve the latency-bound time),
//   fc0_kernel     2 output neurons per CTA, 256 threads stride K with coalesced weight and activation reads,
//                  8 rows of M per pass, block reduction through warp shuffles + one smem hop; blockIdx.y = head;
//   fc12_kernel    one CTA per (row, slab of 128 outputs): fc1 (256 -> 128, recomputed per slab: 32 K MACs) and
//                  fc2 (128 -> out) fused, hidden vectors in shared memory, one warp per output neuron with lanes
//                  over K and a shuffle reduction (the single-CTA-per-row version took 78 us for the 1024-wide keys).
#include "common.cuh"

namespace w2c {
namespace {

constexpr int kRows = 8;     // rows of M per pass in fc0
constexpr int kNeurons = 2;  // output neurons per CTA in fc0
constexpr int kSlab = 128;   // outputs of fc2 per CTA in fc12

struct Heads {
  w2c_mlp_head h[2];
  int slabs0;  // fc12 slabs belonging to head 0
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// feat: NHWC policy map [m][n_feat/256 pixels][planes*256]; W: [256][n_feat] (NHWC flatten order); out [m][256]
__global__ void __launch_bounds__(256) fc0_kernel(const __nv_bfloat16* __restrict__ feat, int act,
                                                  const __grid_constant__ Heads hd, float* __restrict__ ws, int m,
                                                  int n_feat) {
  __shared__ float red[8][kNeurons * kRows];
  const float* __restrict__ W = hd.h[blockIdx.y].w0;
  const float* __restrict__ bias = hd.h[blockIdx.y].b0;
  float* __restrict__ out = ws + static_cast<size_t>(blockIdx.y) * m * 256;
  const int j0 = blockIdx.x * kNeurons;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int planes = act_planes(act);
  const size_t row_elems = static_cast<size_t>(n_feat) * planes;
  const float* w0 = W + static_cast<size_t>(j0) * n_feat;
  const float* w1 = w0 + n_feat;
  for (int m0 = 0; m0 < m; m0 += kRows) {
    float acc[kNeurons][kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) acc[0][r] = acc[1][r] = 0.f;
    // 8 consecutive k per thread and iteration: one 16-byte activation load per row (plus one for the lo plane),
    // two 32-byte weight loads per neuron (n_feat % 256 == 0, so a group of 8 never straddles a pixel)
    for (int k = tid * 8; k < n_feat; k += 256 * 8) {
      float wa[8], wb[8];
      {
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(w0 + k)), a1 = __ldg(reinterpret_cast<const float4*>(w0 + k + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(w1 + k)), b1 = __ldg(reinterpret_cast<const float4*>(w1 + k + 4));
        wa[0] = a0.x, wa[1] = a0.y, wa[2] = a0.z, wa[3] = a0.w, wa[4] = a1.x, wa[5] = a1.y, wa[6] = a1.z, wa[7] = a1.w;
        wb[0] = b0.x, wb[1] = b0.y, wb[2] = b0.z, wb[3] = b0.w, wb[4] = b1.x, wb[5] = b1.y, wb[6] = b1.z, wb[7] = b1.w;
      }
      const size_t off = static_cast<size_t>(k >> 8) * (256 * planes) + (k & 255);  // pixel * pixel-stride + channel
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        if (m0 + r < m) {
          const __nv_bfloat16* p = feat + static_cast<size_t>(m0 + r) * row_elems + off;
          const uint4 hv = __ldg(reinterpret_cast<const uint4*>(p));
          const uint32_t* hb = reinterpret_cast<const uint32_t*>(&hv);
          float x[8];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = unpack_act2(hb[e], act_is_f16(act));
            x[2 * e] = f.x, x[2 * e + 1] = f.y;
          }
          if (planes == 2) {
            const uint4 lv = __ldg(reinterpret_cast<const uint4*>(p + 256));
            const uint32_t* lb = reinterpret_cast<const uint32_t*>(&lv);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = unpack_act2(lb[e], act_is_f16(act));
              x[2 * e] += f.x, x[2 * e + 1] += f.y;
            }
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            acc[0][r] = fmaf(x[e], wa[e], acc[0][r]);
            acc[1][r] = fmaf(x[e], wb[e], acc[1][r]);
          }
        }
      }
    }
#pragma unroll
    for (int n = 0; n < kNeurons; ++n)
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        const float s = warp_sum(acc[n][r]);
        if (lane == 0) red[warp][n * kRows + r] = s;
      }
    __syncthreads();
    if (tid < kNeurons * kRows) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][tid];
      const int n = tid / kRows, r = tid % kRows;
      if (m0 + r < m) out[static_cast<size_t>(m0 + r) * 256 + j0 + n] = fmaxf(s + bias[j0 + n], 0.f);
    }
    __syncthreads();
  }
}

// one CTA per row: h1 = relu(W1 h0 + b1) (128), out = W2 h1 + b2 (out_dim)
__global__ void __launch_bounds__(256) fc12_kernel(const float* __restrict__ ws, const __grid_constant__ Heads hd,
                                                   int m) {
  __shared__ float s_h0[256];
  __shared__ float s_h1[128];
  const int row = blockIdx.x;
  const int head = static_cast<int>(blockIdx.y) >= hd.slabs0 ? 1 : 0;
  const int slab = blockIdx.y - (head ? hd.slabs0 : 0);
  const float* __restrict__ h0 = ws + static_cast<size_t>(head) * m * 256;
  const float* __restrict__ W1 = hd.h[head].w1;
  const float* __restrict__ b1 = hd.h[head].b1;
  const float* __restrict__ W2 = hd.h[head].w2;
  const float* __restrict__ b2 = hd.h[head].b2;
  float* __restrict__ out = hd.h[head].out;
  const int out_dim = hd.h[head].out_dim;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  s_h0[tid] = h0[static_cast<size_t>(row) * 256 + tid];
  __syncthreads();
  for (int j = warp; j < 128; j += 8) {
    const float* wr = W1 + j * 256;
    float acc = 0.f;
#pragma unroll
    for (int k = lane; k < 256; k += 32) acc = fmaf(__ldg(wr + k), s_h0[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) s_h1[j] = fmaxf(acc + b1[j], 0.f);
  }
  __syncthreads();
  const int j_end = min(out_dim, (slab + 1) * kSlab);
  for (int j = slab * kSlab + warp; j < j_end; j += 8) {
    const float* wr = W2 + static_cast<size_t>(j) * 128;
    float acc = 0.f;
#pragma unroll
    for (int k = lane; k < 128; k += 32) acc = fmaf(__ldg(wr + k), s_h1[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) out[static_cast<size_t>(row) * out_dim + j] = acc + b2[j];
  }
}

}  // namespace
}  // namespace w2c

using namespace w2c;

extern "C" int w2c_kq_mlp_heads_fwd(const void* feat, int32_t act, int32_t m, int32_t n_feat,
                                    const w2c_mlp_head* heads, int32_t n_heads, float* ws, w2c_stream_t stream) {
  W2C_CHECK_ARG(feat && heads && ws, "kq_mlp: null pointer");
  W2C_CHECK_ARG(n_heads == 1 || n_heads == 2, "kq_mlp: n_heads=%d (1 or 2)", n_heads);
  W2C_CHECK_ARG(m > 0 && n_feat > 0 && n_feat % 256 == 0, "kq_mlp: bad sizes m=%d n_feat=%d", m, n_feat);
  Heads hd{};
  int slabs = 0;
  for (int i = 0; i < n_heads; ++i) {
    const w2c_mlp_head& h = heads[i];
    W2C_CHECK_ARG(h.w0 && h.b0 && h.w1 && h.b1 && h.w2 && h.b2 && h.out && h.out_dim > 0,
                  "kq_mlp: head %d has a null pointer or out_dim=%d", i, h.out_dim);
    hd.h[i] = h;
    if (i == 0) hd.slabs0 = ceil_div(h.out_dim, kSlab);
    slabs += ceil_div(h.out_dim, kSlab);
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  fc0_kernel<<<dim3(256 / kNeurons, n_heads), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(feat), act, hd, ws, m,
                                                            n_feat);
  W2C_CHECK_LAUNCH("fc0_kernel");
  fc12_kernel<<<dim3(m, slabs), 256, 0, s>>>(ws, hd, m);
  W2C_CHECK_LAUNCH("fc12_kernel");
  return W2C_OK;
}

extern "C" int w2c_kq_mlp_fwd(const void* feat, int32_t act, int32_t m, int32_t n_feat, const float* w0,
                              const float* b0, const float* w1, const float* b1, const float* w2, const float* b2,
                              int32_t out_dim, float* out, float* ws, w2c_stream_t stream) {
  W2C_CHECK_ARG(feat && w0 && b0 && w1 && b1 && w2 && b2 && out && ws, "kq_mlp: null pointer");
  W2C_CHECK_ARG(out_dim > 0, "kq_mlp: out_dim=%d", out_dim);
  const w2c_mlp_head h = {w0, b0, w1, b1, w2, b2, out, out_dim};
  return w2c_kq_mlp_heads_fwd(feat, act, m, n_feat, &h, 1, ws, stream);
}
