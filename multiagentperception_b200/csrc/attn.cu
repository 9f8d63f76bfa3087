// The communication-graph attention + feature fusion kernel.
//
// attn_fuse_kernel is ONE launch for: query projection (W*q+b), key.query scores, softmax / sparsemax over the
// supporting agents, the +0.001*I bias, the activated / argmax re-selection, action + connection bookkeeping,
// and the attention-weighted sum of the agents' feature maps.
// Each CTA owns a slab of PIX consecutive pixels of one scene. It is warp-specialised:
//   * lane 0 of warp 0 first issues one bulk async copy (cp.async.bulk -> smem, mbarrier completion) per
//     supporting agent for the slab's feature rows, so the HBM reads are in flight
//   * while they land, all warps compute the (<= 8 x 8) score matrix of the scene: every dot product is reduced
//     across a warp with shuffles (no shared-memory reduction tree), the softmax columns by one thread each
//   * then all warps wait on the mbarrier and stream the weighted sums out with 128-bit stores.
// The score math is recomputed per CTA (a few hundred kFLOP) instead of a separate launch + global round trip.
#include "common.cuh"
#include "ptx.cuh"

namespace w2c {
namespace {

constexpr int kMaxAgents = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct AttnParams {
  w2c_attn_args a;
  int pix_per_cta;
  int slabs;  // CTAs per scene
};

// Row (agent i, scene) of a per-agent array that may live in an all-gathered buffer: agents are grouped
// agents_per_rank at a time, one group per rank segment, segments rank_stride elements apart.
__device__ __forceinline__ size_t agent_row(const w2c_attn_args& a, int i, int scene, int64_t rank_stride) {
  const int apr = a.agents_per_rank > 0 ? a.agents_per_rank : a.n_k;
  return static_cast<size_t>(i / apr) * static_cast<size_t>(rank_stride) +
         static_cast<size_t>((i % apr) * a.b_sz + scene);
}

// shared memory: [mbarrier 8B pad to 16][qt: n_q*k_dim f32][q: n_q*q_dim f32][S/P/coef: 3*64 f32][V slab]
__global__ void __launch_bounds__(256) attn_fuse_kernel(const AttnParams p) {
  const w2c_attn_args& a = p.a;
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  float* s_qt = reinterpret_cast<float*>(smem + 16);
  float* s_q = s_qt + a.n_q * a.k_dim;
  float* s_S = s_q + a.n_q * a.q_dim;     // [n_k][n_q] raw scores -> probabilities
  float* s_P = s_S + kMaxAgents * kMaxAgents;   // biased probabilities (prob_out)
  float* s_C = s_P + kMaxAgents * kMaxAgents;   // fusion coefficients
  size_t v_off = 16 + (static_cast<size_t>(a.n_q) * (a.k_dim + a.q_dim) + 3 * kMaxAgents * kMaxAgents) * sizeof(float);
  v_off = (v_off + 127) & ~static_cast<size_t>(127);
  __nv_bfloat16* s_V = reinterpret_cast<__nv_bfloat16*>(smem + v_off);

  const int scene = blockIdx.x / p.slabs;
  const int slab = blockIdx.x % p.slabs;
  const int pix0 = slab * p.pix_per_cta;
  const int npix = min(p.pix_per_cta, a.hw - pix0);
  const int planes = act_planes(a.act);
  const int vpix = a.c * planes;  // elements per pixel of val
  const uint32_t slab_bytes = static_cast<uint32_t>(npix) * vpix * 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;

  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // producer: one bulk copy per supporting agent (image index = agent * b_sz + scene)
    ptx::mbar_arrive_expect_tx(bar, slab_bytes * a.n_k);
    for (int i = 0; i < a.n_k; ++i) {
      const size_t img = agent_row(a, i, scene, 0);
      const size_t seg = static_cast<size_t>(i / (a.agents_per_rank > 0 ? a.agents_per_rank : a.n_k)) *
                         static_cast<size_t>(a.val_rank_stride);
      const __nv_bfloat16* src = static_cast<const __nv_bfloat16*>(a.val) + seg + (img * a.hw + pix0) * vpix;
      ptx::bulk_load_1d(s_V + static_cast<size_t>(i) * p.pix_per_cta * vpix, src, slab_bytes, bar);
    }
  }

  // ---- scores (all warps; overlaps the copies above)
  for (int i = threadIdx.x; i < a.n_q * a.q_dim; i += blockDim.x) {
    const int j = i / a.q_dim, e = i % a.q_dim;
    const int apr = a.agents_per_rank > 0 ? a.agents_per_rank : a.n_k;
    s_q[i] = a.queries[static_cast<size_t>(j / apr) * a.queries_rank_stride +
                       static_cast<size_t>((j % apr) * a.b_sz + scene) * a.q_dim + e];
  }
  __syncthreads();
  if (a.wq) {
    // one thread per key dimension d: the row W[d,:] is read once and applied to all (<= 8) queries
    for (int d = threadIdx.x; d < a.k_dim; d += blockDim.x) {
      const float* wr = a.wq + static_cast<size_t>(d) * a.q_dim;
      float acc[kMaxAgents];
#pragma unroll
      for (int j = 0; j < kMaxAgents; ++j) acc[j] = 0.f;
      if ((a.q_dim & 3) == 0) {
        // 16-byte weight loads (the row W[d,:] is contiguous and 16-byte aligned when q_dim % 4 == 0)
        for (int e = 0; e < a.q_dim; e += 4) {
          const float4 wv = __ldg(reinterpret_cast<const float4*>(wr + e));
#pragma unroll
          for (int j = 0; j < kMaxAgents; ++j)
            if (j < a.n_q) {
              const float* qj = s_q + j * a.q_dim + e;
              acc[j] = fmaf(wv.x, qj[0], acc[j]);
              acc[j] = fmaf(wv.y, qj[1], acc[j]);
              acc[j] = fmaf(wv.z, qj[2], acc[j]);
              acc[j] = fmaf(wv.w, qj[3], acc[j]);
            }
        }
      } else {
        for (int e = 0; e < a.q_dim; ++e) {
          const float wv = __ldg(wr + e);
#pragma unroll
          for (int j = 0; j < kMaxAgents; ++j)
            if (j < a.n_q) acc[j] = fmaf(wv, s_q[j * a.q_dim + e], acc[j]);
        }
      }
      const float bias = a.bq ? a.bq[d] : 0.f;
#pragma unroll
      for (int j = 0; j < kMaxAgents; ++j)
        if (j < a.n_q) s_qt[j * a.k_dim + d] = acc[j] + bias;
    }
  } else {
    for (int i = threadIdx.x; i < a.n_q * a.k_dim; i += blockDim.x) s_qt[i] = s_q[i];
  }
  __syncthreads();
  for (int pair = warp; pair < a.n_k * a.n_q; pair += nwarps) {
    const int i = pair / a.n_q, j = pair % a.n_q;
    const int apr = a.agents_per_rank > 0 ? a.agents_per_rank : a.n_k;
    const float* kr = a.keys + static_cast<size_t>(i / apr) * a.keys_rank_stride +
                      static_cast<size_t>((i % apr) * a.b_sz + scene) * a.k_dim;
    float acc = 0.f;
    for (int d = lane; d < a.k_dim; d += 32) acc = fmaf(__ldg(kr + d), s_qt[j * a.k_dim + d], acc);
    acc = warp_sum(acc);
    if (lane == 0) s_S[i * a.n_q + j] = acc / a.temperature;
  }
  __syncthreads();
  if (threadIdx.x < a.n_q) {
    const int j = threadIdx.x;
    float z[kMaxAgents], pr[kMaxAgents];
    int idx[kMaxAgents];
    int n = 0;
    for (int i = 0; i < a.n_k; ++i)
      if (!(a.mask_self && i == j)) z[n] = s_S[i * a.n_q + j], idx[n] = i, ++n;
    float mx = -INFINITY;
    for (int t = 0; t < n; ++t) mx = fmaxf(mx, z[t]);
    if (!a.sparse) {
      float sum = 0.f;
      for (int t = 0; t < n; ++t) pr[t] = expf(z[t] - mx), sum += pr[t];
      for (int t = 0; t < n; ++t) pr[t] = pr[t] / sum;
    } else {
      // sparsemax (utils.py:834-877): shift by max, sort descending, support = {k : 1 + k z_(k) > cumsum_k}
      float zs[kMaxAgents];
      for (int t = 0; t < n; ++t) z[t] -= mx, zs[t] = z[t];
      for (int u = 1; u < n; ++u) {
        const float key = zs[u];
        int v = u - 1;
        while (v >= 0 && zs[v] < key) zs[v + 1] = zs[v], --v;
        zs[v + 1] = key;
      }
      float cum = 0.f, ssum = 0.f, kmax = 0.f;
      for (int t = 0; t < n; ++t) {
        cum += zs[t];
        const float r = static_cast<float>(t + 1);
        if (1.f + r * zs[t] > cum) kmax = fmaxf(kmax, r), ssum += zs[t];
      }
      const float tau = (ssum - 1.f) / kmax;
      for (int t = 0; t < n; ++t) pr[t] = fmaxf(0.f, z[t] - tau);
    }
    for (int i = 0; i < a.n_k; ++i) s_S[i * a.n_q + j] = 0.f;
    for (int t = 0; t < n; ++t) s_S[idx[t] * a.n_q + j] = pr[t];
    // biased matrix, coefficients, action
    int arg_p = 0;
    float best_p = -INFINITY;
    for (int i = 0; i < a.n_k; ++i) {
      const float pb = s_S[i * a.n_q + j] + (i == j ? a.diag_bias : 0.f);
      s_P[i * a.n_q + j] = pb;
      if (pb > best_p) best_p = pb, arg_p = i;
    }
    int arg_c = 0;
    float best_c = -INFINITY;
    int connect = 0;
    for (int i = 0; i < a.n_k; ++i) {
      float cf;
      if (a.mode == W2C_FUSE_SOFTMAX)
        cf = s_S[i * a.n_q + j];
      else if (a.mode == W2C_FUSE_ACTIVATED)
        cf = s_P[i * a.n_q + j] > a.thresh ? s_P[i * a.n_q + j] : 0.f;
      else
        cf = i == arg_p ? 1.f : 0.f;
      s_C[i * a.n_q + j] = cf;
      if (cf > best_c) best_c = cf, arg_c = i;
      if (i != j && cf != 0.f) ++connect;
    }
    if (slab == 0) {
      for (int i = 0; i < a.n_k; ++i) {
        const size_t o = (static_cast<size_t>(scene) * a.n_k + i) * a.n_q + j;
        a.prob_out[o] = s_P[i * a.n_q + j];
        if (a.coef_out) a.coef_out[o] = s_C[i * a.n_q + j];
      }
      if (a.action) a.action[static_cast<size_t>(scene) * a.n_q + j] = a.mode == W2C_FUSE_SOFTMAX ? arg_p : arg_c;
      if (a.connect && a.mode != W2C_FUSE_SOFTMAX && connect) atomicAdd(a.connect, connect);
    }
  }
  __syncthreads();

  // ---- weighted fusion of the slab
  ptx::mbar_wait(bar, 0);
  const int groups = a.c / 8;  // 8-channel (16 B) groups per pixel plane
  const int fpix = a.fused_cstride * planes;
  for (int t = threadIdx.x; t < npix * groups; t += blockDim.x) {
    const int px = t / groups, g = t % groups;
    float v[kMaxAgents][8];
#pragma unroll
    for (int i = 0; i < kMaxAgents; ++i) {
      if (i < a.n_k) {
        const __nv_bfloat16* src = s_V + (static_cast<size_t>(i) * p.pix_per_cta + px) * vpix + g * 8;
        const uint4 hv = *reinterpret_cast<const uint4*>(src);
        const uint32_t* hb = reinterpret_cast<const uint32_t*>(&hv);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack_act2(hb[e], act_is_f16(a.act));
          v[i][2 * e] = f.x, v[i][2 * e + 1] = f.y;
        }
        if (planes == 2) {
          const uint4 lv = *reinterpret_cast<const uint4*>(src + a.c);
          const uint32_t* lb = reinterpret_cast<const uint32_t*>(&lv);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = unpack_act2(lb[e], act_is_f16(a.act));
            v[i][2 * e] += f.x, v[i][2 * e + 1] += f.y;
          }
        }
      }
    }
    const int j_end = a.q_count > 0 ? a.q_first + a.q_count : a.n_q;
    for (int j = a.q_first; j < j_end; ++j) {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = 0.f;
#pragma unroll
      for (int i = 0; i < kMaxAgents; ++i) {
        if (i < a.n_k) {
          const float cf = s_C[i * a.n_q + j];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = fmaf(cf, v[i][e], o[e]);
        }
      }
      uint4 hv, lv;
      uint32_t* hb = reinterpret_cast<uint32_t*>(&hv);
      uint32_t* lb = reinterpret_cast<uint32_t*>(&lv);
#pragma unroll
      for (int e = 0; e < 4; ++e) split_act2(o[2 * e], o[2 * e + 1], act_is_f16(a.act), hb[e], lb[e]);
      __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(a.fused) +
                           (static_cast<size_t>((j - a.q_first) * a.b_sz + scene) * a.hw + pix0 + px) * fpix + a.fused_coffset + g * 8;
      *reinterpret_cast<uint4*>(dst) = hv;
      if (planes == 2) *reinterpret_cast<uint4*>(dst + a.fused_cstride) = lv;
    }
  }
}

}  // namespace
}  // namespace w2c

using namespace w2c;

extern "C" int w2c_attn_fuse_fwd(const w2c_attn_args* args, w2c_stream_t stream) {
  W2C_CHECK_ARG(args, "attn: args is NULL");
  const w2c_attn_args& a = *args;
  W2C_CHECK_ARG(a.keys && a.queries && a.val && a.fused && a.prob_out, "attn: null pointer");
  W2C_CHECK_ARG(a.b_sz > 0 && a.n_k > 0 && a.n_q > 0 && a.n_k <= kMaxAgents && a.n_q <= kMaxAgents,
                "attn: agent counts must be in [1, 8] (n_k=%d n_q=%d)", a.n_k, a.n_q);
  W2C_CHECK_ARG(a.k_dim > 0 && a.q_dim > 0 && (a.wq || a.q_dim == a.k_dim), "attn: query/key sizes %d/%d need wq",
                a.q_dim, a.k_dim);
  W2C_CHECK_ARG(a.hw > 0 && a.c > 0 && a.c % 8 == 0, "attn: c=%d must be a multiple of 8", a.c);
  W2C_CHECK_ARG(a.mode >= W2C_FUSE_SOFTMAX && a.mode <= W2C_FUSE_ARGMAX, "attn: bad mode %d", a.mode);
  W2C_CHECK_ARG(a.temperature != 0.f, "attn: temperature must be non-zero");
  // 16-byte vector loads of the projection rows and bulk copies / 16-byte stores of the feature maps
  W2C_CHECK_ARG(!a.wq || a.q_dim % 4 != 0 || reinterpret_cast<uintptr_t>(a.wq) % 16 == 0,
                "attn: wq must be 16-byte aligned when q_dim is a multiple of 4");
  W2C_CHECK_ARG(reinterpret_cast<uintptr_t>(a.val) % 16 == 0 && reinterpret_cast<uintptr_t>(a.fused) % 16 == 0,
                "attn: val and fused must be 16-byte aligned");
  W2C_CHECK_ARG(a.agents_per_rank == 0 || (a.val_rank_stride * 2) % 16 == 0,
                "attn: val_rank_stride must keep every rank's feature maps 16-byte aligned");
  W2C_CHECK_ARG(!a.mask_self || a.n_k > 1, "attn: mask_self needs at least two supporting agents");
  W2C_CHECK_ARG(a.q_first >= 0 && a.q_count >= 0 && a.q_first + a.q_count <= a.n_q, "attn: fused query window "
                "[%d, %d) outside n_q=%d", a.q_first, a.q_first + a.q_count, a.n_q);
  W2C_CHECK_ARG(a.agents_per_rank >= 0 && (a.agents_per_rank == 0 || a.n_k % a.agents_per_rank == 0),
                "attn: agents_per_rank=%d must divide n_k=%d", a.agents_per_rank, a.n_k);
  AttnParams p;
  p.a = a;
  if (p.a.fused_cstride <= 0) p.a.fused_cstride = a.c;
  W2C_CHECK_ARG(p.a.fused_cstride % 8 == 0 && p.a.fused_coffset % 8 == 0 && p.a.fused_coffset + a.c <= p.a.fused_cstride,
                "attn: fused slice out of range");
  const int planes = act_planes(a.act);
  const size_t head = ((16 + (static_cast<size_t>(a.n_q) * (a.k_dim + a.q_dim) + 3 * kMaxAgents * kMaxAgents) * 4) + 127) &
                      ~static_cast<size_t>(127);
  // slab size: as many pixels as fit ~96 KB of feature rows, at least 1, and 16-byte granular copies
  const size_t per_pix = static_cast<size_t>(a.n_k) * a.c * planes * 2;
  int pix = static_cast<int>((96 * 1024) / per_pix);
  if (pix < 1) pix = 1;
  if (pix > a.hw) pix = a.hw;
  // every CTA recomputes its scene's score matrix, so fewer / fatter CTAs are better as long as the SMs are covered
  while (pix > 4 && static_cast<long long>(a.b_sz) * ceil_div(a.hw, pix) < 120) pix = (pix + 1) / 2;
  p.pix_per_cta = pix;
  p.slabs = ceil_div(a.hw, pix);
  const size_t smem = head + per_pix * pix;
  W2C_CHECK_ARG(smem <= 200 * 1024, "attn: shared memory request %zu too large", smem);
  static DeviceOnce attr;  // per device, not per process; opted in to the largest request the check above allows
  if (int rc = attr.ensure([] {
        return cudaFuncSetAttribute(attn_fuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      }, "attn_fuse_kernel"))
    return rc;
  attn_fuse_kernel<<<a.b_sz * p.slabs, 256, smem, static_cast<cudaStream_t>(stream)>>>(p);
  W2C_CHECK_LAUNCH("attn_fuse_kernel");
  return W2C_OK;
}
