// Backward of the communication-graph attention + fusion (csrc/attn.cu) in the differentiable (softmax / sparsemax)
// mode the trainers use (forward(training=True), ptsemseg/models/agent.py:1170-1179; loss.backward(),
// ptsemseg/trainer.py:668-670): what autograd runs for MIMOGeneralDotProductAttention.forward (agent.py:252-286),
// GeneralDotProductAttention / ScaledDotProductAttention (agent.py:194-213,345-368).
//
//   forward:  qt_j = Wq q_j + bq;  S_ij = <k_i, qt_j> / T;  P_:j = softmax_i S_:j;  F_j = sum_i P_ij V_i
//   given dF_j:
//     dV_i  = sum_j P_ij dF_j                                          attn_bwd_values_kernel   (HBM-bound)
//     dP_ij = <dF_j, V_i>                                              attn_bwd_dots_kernel     (HBM-bound)
//     dS_ij = P_ij (dP_ij - sum_i' P_i'j dP_i'j)      [sparsemax: on the support, minus the support mean]
//     dk_i  = sum_j dS_ij qt_j / T;  dqt_j = sum_i dS_ij k_i / T
//     dq_j  = Wq^T dqt_j;  dWq += dqt_j q_j^T;  dbq += dqt_j           attn_bwd_scores_kernel   (one CTA per scene)
// P is the un-biased probability matrix the forward fused with (coef_out of w2c_attn_fuse_fwd in SOFTMAX mode).
#include "common.cuh"

namespace w2c {
namespace {

constexpr int kMaxAgents = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

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

// grid (slabs, b_sz): dP[b][i][j] += sum over this slab's (pixel, 8-channel group) items of V_i * dF_j
__global__ void __launch_bounds__(256) attn_bwd_dots_kernel(const w2c_attn_bwd_args a, float* __restrict__ dP) {
  __shared__ float s_red[8][kMaxAgents * kMaxAgents];
  const int scene = blockIdx.y;
  const int groups = a.c / 8;
  const long long items = static_cast<long long>(a.hw) * groups;
  const bool f16f = act_is_f16(a.act_f), f16g = act_is_f16(a.act_g);
  const int pf = act_planes(a.act_f), pg = act_planes(a.act_g);
  const __nv_bfloat16* val = static_cast<const __nv_bfloat16*>(a.val);
  const __nv_bfloat16* df = static_cast<const __nv_bfloat16*>(a.dfused);
  const int df_cs = a.dfused_cstride > 0 ? a.dfused_cstride : a.c;
  float acc[kMaxAgents][kMaxAgents];
#pragma unroll
  for (int i = 0; i < kMaxAgents; ++i)
#pragma unroll
    for (int j = 0; j < kMaxAgents; ++j) acc[i][j] = 0.f;
  for (long long it = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; it < items;
       it += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(it % groups);
    const size_t px = static_cast<size_t>(it / groups);
    float v[kMaxAgents][8];
#pragma unroll
    for (int i = 0; i < kMaxAgents; ++i)
      if (i < a.n_k)
        load8(val + ((static_cast<size_t>(i) * a.b_sz + scene) * a.hw + px) * (static_cast<size_t>(a.c) * pf) + g * 8, a.c, pf,
              f16f, v[i]);
#pragma unroll
    for (int j = 0; j < kMaxAgents; ++j)
      if (j < a.n_q) {
        float d[8];
        load8(df + ((static_cast<size_t>(j) * a.b_sz + scene) * a.hw + px) * (static_cast<size_t>(df_cs) * pg) +
                  a.dfused_coffset + g * 8, df_cs, pg, f16g, d);
#pragma unroll
        for (int i = 0; i < kMaxAgents; ++i)
          if (i < a.n_k) {
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[i][j] = fmaf(v[i][e], d[e], acc[i][j]);
          }
      }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < kMaxAgents; ++i)
#pragma unroll
    for (int j = 0; j < kMaxAgents; ++j) {
      const float s = warp_sum(acc[i][j]);
      if (lane == 0) s_red[warp][i * kMaxAgents + j] = s;
    }
  __syncthreads();
  if (threadIdx.x < kMaxAgents * kMaxAgents) {
    const int i = threadIdx.x / kMaxAgents, j = threadIdx.x % kMaxAgents;
    if (i < a.n_k && j < a.n_q) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += s_red[w][threadIdx.x];
      atomicAdd(&dP[(static_cast<size_t>(scene) * a.n_k + i) * a.n_q + j], s);
    }
  }
}

// one CTA per scene. shared: qt [n_q][k_dim], dqt [n_q][k_dim], q [n_q][q_dim], dS [n_k][n_q]
__global__ void __launch_bounds__(256) attn_bwd_scores_kernel(const w2c_attn_bwd_args a, float* __restrict__ dP) {
  extern __shared__ float sm[];
  float* s_qt = sm;
  float* s_dqt = s_qt + a.n_q * a.k_dim;
  float* s_q = s_dqt + a.n_q * a.k_dim;
  float* s_dS = s_q + a.n_q * a.q_dim;
  const int scene = blockIdx.x;
  const float inv_t = 1.f / a.temperature;
  for (int i = threadIdx.x; i < a.n_q * a.q_dim; i += blockDim.x) {
    const int j = i / a.q_dim, e = i % a.q_dim;
    s_q[i] = a.queries[(static_cast<size_t>(j) * a.b_sz + scene) * a.q_dim + e];
  }
  // dS from P and dP (one thread per query column); dP is left zeroed for the next backward
  if (threadIdx.x < a.n_q) {
    const int j = threadIdx.x;
    float pv[kMaxAgents], dp[kMaxAgents];
    float dot = 0.f, cnt = 0.f, sum_supp = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxAgents; ++i)
      if (i < a.n_k) {
        const size_t idx = (static_cast<size_t>(scene) * a.n_k + i) * a.n_q + j;
        pv[i] = a.prob[idx], dp[i] = dP[idx];
        dP[idx] = 0.f;
        dot = fmaf(pv[i], dp[i], dot);
        if (pv[i] > 0.f) cnt += 1.f, sum_supp += dp[i];
      }
#pragma unroll
    for (int i = 0; i < kMaxAgents; ++i)
      if (i < a.n_k) {
        float ds;
        if (a.sparse)
          ds = pv[i] > 0.f ? dp[i] - sum_supp / cnt : 0.f;   // sparsemax Jacobian: diag(s) - s s^T / |S|
        else
          ds = pv[i] * (dp[i] - dot);
        s_dS[i * a.n_q + j] = ds * inv_t;
      }
  }
  __syncthreads();
  // qt = Wq q + bq (recomputed) and dqt_j = sum_i dS_ij k_i, one thread per key dimension
  for (int d = threadIdx.x; d < a.k_dim; d += blockDim.x) {
    float acc[kMaxAgents], dq[kMaxAgents];
#pragma unroll
    for (int j = 0; j < kMaxAgents; ++j) acc[j] = dq[j] = 0.f;
    if (a.wq) {
      const float* wr = a.wq + static_cast<size_t>(d) * a.q_dim;
      for (int e = 0; e < a.q_dim; ++e) {
        const float wv = __ldg(wr + e);
#pragma unroll
        for (int j = 0; j < kMaxAgents; ++j)
          if (j < a.n_q) acc[j] = fmaf(wv, s_q[j * a.q_dim + e], acc[j]);
      }
      const float bias = a.bq ? a.bq[d] : 0.f;
#pragma unroll
      for (int j = 0; j < kMaxAgents; ++j) acc[j] += bias;
    } else {
#pragma unroll
      for (int j = 0; j < kMaxAgents; ++j)
        if (j < a.n_q) acc[j] = s_q[j * a.q_dim + d];
    }
#pragma unroll
    for (int i = 0; i < kMaxAgents; ++i)
      if (i < a.n_k) {
        const float kv = a.keys[(static_cast<size_t>(i) * a.b_sz + scene) * a.k_dim + d];
        float dk = 0.f;
#pragma unroll
        for (int j = 0; j < kMaxAgents; ++j)
          if (j < a.n_q) {
            const float ds = s_dS[i * a.n_q + j];
            dk = fmaf(ds, acc[j], dk);
            dq[j] = fmaf(ds, kv, dq[j]);
          }
        a.dkeys[(static_cast<size_t>(i) * a.b_sz + scene) * a.k_dim + d] = dk;
      }
    float db = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxAgents; ++j)
      if (j < a.n_q) {
        s_qt[j * a.k_dim + d] = acc[j];
        s_dqt[j * a.k_dim + d] = dq[j];
        db += dq[j];
      }
    if (a.wq && a.dbq) atomicAdd(&a.dbq[d], db);
  }
  __syncthreads();
  if (!a.wq) {
    // no projection: the query is the key-space vector itself
    for (int i = threadIdx.x; i < a.n_q * a.q_dim; i += blockDim.x) {
      const int j = i / a.q_dim, e = i % a.q_dim;
      if (a.dqueries) a.dqueries[(static_cast<size_t>(j) * a.b_sz + scene) * a.q_dim + e] = s_dqt[j * a.k_dim + e];
    }
    return;
  }
  // dq_j[e] = sum_d Wq[d][e] dqt_j[d]  (one warp per (j, e), lanes over d)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int it = warp; it < a.n_q * a.q_dim; it += nwarps) {
    const int j = it / a.q_dim, e = it % a.q_dim;
    float acc = 0.f;
    for (int d = lane; d < a.k_dim; d += 32) acc = fmaf(__ldg(a.wq + static_cast<size_t>(d) * a.q_dim + e), s_dqt[j * a.k_dim + d], acc);
    acc = warp_sum(acc);
    if (lane == 0 && a.dqueries) a.dqueries[(static_cast<size_t>(j) * a.b_sz + scene) * a.q_dim + e] = acc;
  }
  // dWq[d][e] += sum_j dqt_j[d] q_j[e]
  if (a.dwq)
    for (int it = threadIdx.x; it < a.k_dim * a.q_dim; it += blockDim.x) {
      const int d = it / a.q_dim, e = it % a.q_dim;
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < kMaxAgents; ++j)
        if (j < a.n_q) acc = fmaf(s_dqt[j * a.k_dim + d], s_q[j * a.q_dim + e], acc);
      atomicAdd(&a.dwq[it], acc);
    }
}

// dV_i = sum_j P_ij dF_j; one thread per (image of val, pixel, 8-channel group)
__global__ void __launch_bounds__(256) attn_bwd_values_kernel(const w2c_attn_bwd_args a) {
  const int groups = a.c / 8;
  const size_t per_img = static_cast<size_t>(a.hw) * groups;
  const size_t total = per_img * a.n_k * a.b_sz;
  const bool f16g = act_is_f16(a.act_g);
  const int pg = act_planes(a.act_g);
  const __nv_bfloat16* df = static_cast<const __nv_bfloat16*>(a.dfused);
  __nv_bfloat16* dv = static_cast<__nv_bfloat16*>(a.dval);
  const int df_cs = a.dfused_cstride > 0 ? a.dfused_cstride : a.c;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t img = idx / per_img;               // = i * b_sz + scene
    const size_t rem = idx % per_img;
    const int g = static_cast<int>(rem % groups);
    const size_t px = rem / groups;
    const int i = static_cast<int>(img / a.b_sz), scene = static_cast<int>(img % a.b_sz);
    float out[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) out[e] = 0.f;
    for (int j = 0; j < a.n_q; ++j) {
      const float pij = a.prob[(static_cast<size_t>(scene) * a.n_k + i) * a.n_q + j];
      if (pij == 0.f) continue;
      float d[8];
      load8(df + ((static_cast<size_t>(j) * a.b_sz + scene) * a.hw + px) * (static_cast<size_t>(df_cs) * pg) +
                a.dfused_coffset + g * 8, df_cs, pg, f16g, d);
#pragma unroll
      for (int e = 0; e < 8; ++e) out[e] = fmaf(pij, d[e], out[e]);
    }
    if (a.dval_accumulate) {
      float old[8];
      load8(dv + (img * a.hw + px) * (static_cast<size_t>(a.c) * pg) + g * 8, a.c, pg, f16g, old);
#pragma unroll
      for (int e = 0; e < 8; ++e) out[e] += old[e];
    }
    uint4 hv, lv;
    uint32_t* hw = reinterpret_cast<uint32_t*>(&hv);
    uint32_t* lw = reinterpret_cast<uint32_t*>(&lv);
#pragma unroll
    for (int e = 0; e < 4; ++e) split_act2(out[2 * e], out[2 * e + 1], f16g, hw[e], lw[e]);
    __nv_bfloat16* q = dv + (img * a.hw + px) * (static_cast<size_t>(a.c) * pg) + g * 8;
    *reinterpret_cast<uint4*>(q) = hv;
    if (pg == 2) *reinterpret_cast<uint4*>(q + a.c) = lv;
  }
}

}  // namespace
}  // namespace w2c

using namespace w2c;

extern "C" int w2c_attn_fuse_bwd(const w2c_attn_bwd_args* args, w2c_stream_t stream) {
  if (!args) return set_error(W2C_ERR_INVALID, "attn_bwd: args is NULL");
  const w2c_attn_bwd_args& a = *args;
  W2C_CHECK_ARG(a.keys && a.queries && a.val && a.dfused && a.prob && a.dval && a.dkeys && a.dp_ws,
                "attn_bwd: null pointer argument");
  W2C_CHECK_ARG(a.n_k >= 1 && a.n_k <= kMaxAgents && a.n_q >= 1 && a.n_q <= kMaxAgents && a.b_sz > 0,
                "attn_bwd: n_k=%d n_q=%d b_sz=%d", a.n_k, a.n_q, a.b_sz);
  W2C_CHECK_ARG(act_valid(a.act_f) && act_valid(a.act_g) && a.c > 0 && a.c % 8 == 0 && a.hw > 0, "attn_bwd: bad map geometry");
  W2C_CHECK_ARG(a.wq || a.q_dim == a.k_dim, "attn_bwd: without a projection q_dim must equal k_dim");
  W2C_CHECK_ARG(a.temperature > 0.f, "attn_bwd: temperature");
  const int df_cs = a.dfused_cstride > 0 ? a.dfused_cstride : a.c;
  W2C_CHECK_ARG(a.dfused_coffset >= 0 && a.dfused_coffset + a.c <= df_cs && df_cs % 8 == 0 && a.dfused_coffset % 8 == 0,
                "attn_bwd: dfused channel slice out of range");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long items = static_cast<long long>(a.hw) * (a.c / 8);
  int slabs = static_cast<int>((items + 256 * 8 - 1) / (256 * 8));
  const int cap = ceil_div(device_sm_count() * 4, a.b_sz);
  if (slabs > cap) slabs = cap;
  if (slabs < 1) slabs = 1;
  attn_bwd_dots_kernel<<<dim3(slabs, a.b_sz), 256, 0, s>>>(a, a.dp_ws);
  W2C_CHECK_LAUNCH("attn_bwd_dots_kernel");
  const size_t smem = (static_cast<size_t>(a.n_q) * (2 * a.k_dim + a.q_dim) + a.n_k * a.n_q) * sizeof(float);
  W2C_CHECK_ARG(smem <= 200 * 1024, "attn_bwd: k_dim=%d too large for the score kernel's shared memory", a.k_dim);
  static DeviceOnce attr;
  if (int rc = attr.ensure([] {
        return cudaFuncSetAttribute(attn_bwd_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      }, "attn_bwd_scores_kernel"))
    return rc;
  attn_bwd_scores_kernel<<<a.b_sz, 256, smem, s>>>(a, a.dp_ws);
  W2C_CHECK_LAUNCH("attn_bwd_scores_kernel");
  const size_t total = static_cast<size_t>(a.hw) * (a.c / 8) * a.n_k * a.b_sz;
  const size_t blocks = (total + 255) / 256;
  const size_t capb = static_cast<size_t>(device_sm_count()) * 32;
  attn_bwd_values_kernel<<<static_cast<int>(blocks < capb ? blocks : capb), 256, 0, s>>>(a);
  W2C_CHECK_LAUNCH("attn_bwd_values_kernel");
  return W2C_OK;
}
