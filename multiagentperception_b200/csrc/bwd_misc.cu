// Backward of the small stages of the path (loss.backward() of ptsemseg/trainer.py:668-670):
//   * the key / query heads (km_generator / linear, ptsemseg/models/agent.py:145-178): three Linear layers with ReLU
//   * the weight gradient of the 3-channel first layers (n_segnet_encoder.conv1 3x3 s1, backbone.py:19,42; the resnet18
//     stem 7x7 s2, backbone.py:63-69), which read the caller's fp32 NCHW views directly
//   * the operand packing of the data-gradient convs (a conv's data gradient is a conv / transposed conv of dL/dy with
//     the SAME weight tensor re-indexed, so it runs on the forward tensor-core kernels)
//   * MaxPool2d(3, 2, 1) and the x32 bilinear up-sampling of the resnet18 + simple_decoder pair (backbone.py:72-96,
//     156-164), the stride-2 1x1 shortcut's zero-interleave, gradient accumulation
// Everything here is CUDA-core work: tens of rows (M = agents x scenes) against weight matrices, or HBM-bound passes.
#include "common.cuh"

namespace w2c {
namespace {

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

__device__ __forceinline__ void store8(__nv_bfloat16* q, int cstride, int planes, bool f16, const float (&v)[8]) {
  uint4 hv, lv;
  uint32_t* hw = reinterpret_cast<uint32_t*>(&hv);
  uint32_t* lw = reinterpret_cast<uint32_t*>(&lv);
#pragma unroll
  for (int e = 0; e < 4; ++e) split_act2(v[2 * e], v[2 * e + 1], f16, hw[e], lw[e]);
  *reinterpret_cast<uint4*>(q) = hv;
  if (planes == 2) *reinterpret_cast<uint4*>(q + cstride) = lv;
}

int grid_for(size_t total, int threads, int per_sm = 16) {
  const size_t want = (total + threads - 1) / threads;
  const size_t cap = static_cast<size_t>(device_sm_count()) * per_sm;
  return static_cast<int>(want < cap ? (want ? want : 1) : cap);
}

// ------------------------------------------------------------------------------------------------ key / query heads
struct HeadsBwd {
  w2c_mlp_head h[2];
  w2c_mlp_head_grad g[2];
  int n_heads;
};

// One CTA per (row, head): recompute h1 = relu(W1 h0 + b1), then dh1 = (dout W2) * [h1 > 0], dh0 = (dh1 W1) * [h0 > 0].
// ws per head: [h1: m*128][dh1: m*128][dh0: m*256]
__global__ void __launch_bounds__(256) mlp_bwd_hidden_kernel(const HeadsBwd hd, const float* __restrict__ ws_fwd,
                                                             float* __restrict__ ws, int m) {
  __shared__ float s_h0[256], s_h1[128], s_dh1[128];
  __shared__ float s_dout[1024];
  const int row = blockIdx.x, head = blockIdx.y;
  const w2c_mlp_head& h = hd.h[head];
  const float* __restrict__ dout = hd.g[head].dout + static_cast<size_t>(row) * h.out_dim;
  const float* __restrict__ h0 = ws_fwd + (static_cast<size_t>(head) * m + row) * 256;
  float* w_h1 = ws + static_cast<size_t>(head) * m * 512 + static_cast<size_t>(row) * 128;
  float* w_dh1 = ws + static_cast<size_t>(head) * m * 512 + static_cast<size_t>(m) * 128 + static_cast<size_t>(row) * 128;
  float* w_dh0 = ws + static_cast<size_t>(head) * m * 512 + static_cast<size_t>(m) * 256 + static_cast<size_t>(row) * 256;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  s_h0[tid] = h0[tid];
  __syncthreads();
  for (int j = warp; j < 128; j += 8) {
    const float* wr = h.w1 + j * 256;
    float acc = 0.f;
#pragma unroll
    for (int k = lane; k < 256; k += 32) acc = fmaf(__ldg(wr + k), s_h0[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) s_h1[j] = fmaxf(acc + h.b1[j], 0.f);
  }
  // dh1[k] = sum_o dout[o] W2[o][k], out_dim staged through shared memory 1024 at a time
  float acc1 = 0.f;
  for (int o0 = 0; o0 < h.out_dim; o0 += 1024) {
    __syncthreads();
    for (int o = tid; o < 1024 && o0 + o < h.out_dim; o += 256) s_dout[o] = dout[o0 + o];
    __syncthreads();
    if (tid < 128) {
      const int n = min(1024, h.out_dim - o0);
      for (int o = 0; o < n; ++o) acc1 = fmaf(s_dout[o], __ldg(h.w2 + static_cast<size_t>(o0 + o) * 128 + tid), acc1);
    }
  }
  __syncthreads();
  if (tid < 128) {
    const float d = s_h1[tid] > 0.f ? acc1 : 0.f;
    s_dh1[tid] = d;
    w_h1[tid] = s_h1[tid];
    w_dh1[tid] = d;
  }
  __syncthreads();
  float acc0 = 0.f;
#pragma unroll 4
  for (int j = 0; j < 128; ++j) acc0 = fmaf(s_dh1[j], __ldg(h.w1 + j * 256 + tid), acc0);
  w_dh0[tid] = s_h0[tid] > 0.f ? acc0 : 0.f;
}

// dW[o][k] += sum_r A[r][o] * B[r][k];  db[o] += sum_r A[r][o]   (A: [m][no], B: [m][nk], both fp32)
__global__ void __launch_bounds__(256) outer_sum_kernel(const float* __restrict__ A, const float* __restrict__ B, int m,
                                                        int no, int nk, float* __restrict__ dW, float* __restrict__ db) {
  const size_t total = static_cast<size_t>(no) * nk;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int o = static_cast<int>(idx / nk), k = static_cast<int>(idx % nk);
    float acc = 0.f, bs = 0.f;
    for (int r = 0; r < m; ++r) {
      const float av = __ldg(A + static_cast<size_t>(r) * no + o);
      acc = fmaf(av, __ldg(B + static_cast<size_t>(r) * nk + k), acc);
      bs += av;
    }
    dW[idx] += acc;
    if (db && k == 0) db[o] += bs;
  }
}

// fc0 weight gradient: dW0[j][k] += sum_r dh0[r][j] x[r][k] (x = the NHWC policy map, k in NHWC flatten order).
// One thread per (8 consecutive k, 8 consecutive j), grid (k blocks, 32 j groups, heads): the 8 activations of each row
// are loaded once and used for 8 rows of W0.
__global__ void __launch_bounds__(128) mlp_bwd_fc0_wgrad_kernel(const HeadsBwd hd, const __nv_bfloat16* __restrict__ feat,
                                                                int act_f, const float* __restrict__ ws, int m, int n_feat) {
  const int k0 = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (k0 >= n_feat) return;
  const int j0 = blockIdx.y * 8, hh = blockIdx.z;
  const int pf = act_planes(act_f);
  const size_t off_f = static_cast<size_t>(k0 >> 8) * (256 * pf) + (k0 & 255);
  const float* dh0 = ws + static_cast<size_t>(hh) * m * 512 + static_cast<size_t>(m) * 256;
  float gw[8][8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) gw[j][e] = 0.f;
  for (int r = 0; r < m; ++r) {
    float x[8];
    load8(feat + static_cast<size_t>(r) * (static_cast<size_t>(n_feat) * pf) + off_f, 256, pf, act_is_f16(act_f), x);
    const float4 da = __ldg(reinterpret_cast<const float4*>(dh0 + static_cast<size_t>(r) * 256 + j0));
    const float4 db = __ldg(reinterpret_cast<const float4*>(dh0 + static_cast<size_t>(r) * 256 + j0 + 4));
    const float d[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 8; ++e) gw[j][e] = fmaf(d[j], x[e], gw[j][e]);
  }
  float* dW0 = hd.g[hh].dw0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float4* gp = reinterpret_cast<float4*>(dW0 + static_cast<size_t>(j0 + j) * n_feat + k0);
    float4 ga = gp[0], gb = gp[1];
    ga.x += gw[j][0], ga.y += gw[j][1], ga.z += gw[j][2], ga.w += gw[j][3];
    gb.x += gw[j][4], gb.y += gw[j][5], gb.z += gw[j][6], gb.w += gw[j][7];
    gp[0] = ga, gp[1] = gb;
  }
}

// fc0 data gradient: dx[r][k] = sum_heads sum_j dh0[r][j] W0[j][k]. One thread per (row, 8 consecutive k); the row's
// dh0 vectors sit in shared memory.
__global__ void __launch_bounds__(128) mlp_bwd_fc0_dgrad_kernel(const HeadsBwd hd, const float* __restrict__ ws, int m,
                                                                int n_feat, __nv_bfloat16* __restrict__ dfeat, int act_g) {
  __shared__ float s_dh0[2][256];
  const int r = blockIdx.y;
  for (int i = threadIdx.x; i < hd.n_heads * 256; i += blockDim.x)
    s_dh0[i / 256][i % 256] = ws[static_cast<size_t>(i / 256) * m * 512 + static_cast<size_t>(m) * 256 +
                                 static_cast<size_t>(r) * 256 + (i % 256)];
  __syncthreads();
  const int k0 = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (k0 >= n_feat) return;
  const int pg = act_planes(act_g);
  float dx[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) dx[e] = 0.f;
  for (int hh = 0; hh < hd.n_heads; ++hh) {
    const float* __restrict__ W0 = hd.h[hh].w0;
#pragma unroll 4
    for (int j = 0; j < 256; ++j) {
      const float4 wa = __ldg(reinterpret_cast<const float4*>(W0 + static_cast<size_t>(j) * n_feat + k0));
      const float4 wb = __ldg(reinterpret_cast<const float4*>(W0 + static_cast<size_t>(j) * n_feat + k0 + 4));
      const float d = s_dh0[hh][j];
      dx[0] = fmaf(d, wa.x, dx[0]), dx[1] = fmaf(d, wa.y, dx[1]), dx[2] = fmaf(d, wa.z, dx[2]), dx[3] = fmaf(d, wa.w, dx[3]);
      dx[4] = fmaf(d, wb.x, dx[4]), dx[5] = fmaf(d, wb.y, dx[5]), dx[6] = fmaf(d, wb.z, dx[6]), dx[7] = fmaf(d, wb.w, dx[7]);
    }
  }
  const size_t off_g = static_cast<size_t>(k0 >> 8) * (256 * pg) + (k0 & 255);
  store8(dfeat + static_cast<size_t>(r) * (static_cast<size_t>(n_feat) * pg) + off_g, 256, pg, act_is_f16(act_g), dx);
}

// ------------------------------------------------------------------------------------------------ first-layer wgrad
// dW[co][ci][kh][kw] += sum over pixels dz[pix][co] * x[ci][pix*stride + k - pad].  Persistent CTAs walk tiles of
// TP x TP output pixels of one image: the dz tile (TP*TP x 64) and the input patch are staged in shared memory. A thread
// owns one (8-channel group, input channel, filter row) and keeps the KS taps of that row x 8 channels in registers
// across ALL its tiles: per pixel two 16-byte loads of dz and KS patch words feed 8 * KS FMAs (7x7: 56 FMAs per 9
// loads; the one-tap-per-thread version did 8 per 3 and ran at the shared-memory rate, 0.67 ms per 10 frames). With
// fewer items than compute threads (3x3: 72) the tile's pixels are dealt over `parts` partitions. The last warp(s) of
// the CTA do no arithmetic: they stage the NEXT tile into the other half of a double buffer while the others compute
// (staging and arithmetic took about the same time per tile and two 126-register CTAs per SM did not overlap them).
// One atomicAdd per (co, k) and thread at the end.
template <int KS, int STRIDE>
struct StemWgradGeom {
  static constexpr int TP = 8;                             // output pixels per tile edge
  static constexpr int PAD = KS / 2;
  static constexpr int PW = (TP - 1) * STRIDE + KS;        // input patch edge
  static constexpr int KK = 3 * KS * KS;
  static constexpr int kPatchWords = (3 * PW * PW + 3) & ~3;
  static size_t smem_bytes(int cout) { return 2 * (static_cast<size_t>(kPatchWords) + TP * TP * cout) * sizeof(float); }
};

template <int KS, int STRIDE>
__global__ void __launch_bounds__(256, 2) stem_wgrad_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dz,
                                                         float* __restrict__ dw, int b, int n_agents, int c_total,
                                                         int c_first, int h, int w, int cout, int dz_cs, int dz_co,
                                                         int act_g, int tiles_w, int tiles_h, int n_tiles) {
  using G = StemWgradGeom<KS, STRIDE>;
  constexpr int TP = G::TP, PAD = G::PAD, PW = G::PW, KK = G::KK;
  extern __shared__ float s_mem[];
  const int buf_words = G::kPatchWords + TP * TP * cout;  // one buffer: patch [3][PW][PW], then dz [TP*TP][cout]
  const int ho = h / STRIDE, wo = w / STRIDE;
  const bool f16g = act_is_f16(act_g);
  const int pg = act_planes(act_g);
  const int n_items = (cout / 8) * 3 * KS;                // (group, ci, kh), kh fastest; cout <= 64: <= 168
  // pixel partitions: three warps (two when the items fill the rest) are left to stage - with ONE loader warp the
  // 3x3 layer waited on it (16 dependent-latency trips per tile: 2.1 ms per launch instead of 0.66)
  const int parts = 160 / n_items > 1 ? 160 / n_items : 1;
  const int first_loader = (parts * n_items + 31) & ~31;
  const int n_loaders = 256 - first_loader;
  const int item = threadIdx.x % n_items, part = threadIdx.x / n_items;
  const bool active = part < parts;
  const int g = item / (3 * KS), ci = (item / KS) % 3, kh = item % KS;

  auto stage = [&](int tile, float* buf, int t0, int nt) {
    int t = tile;
    const int tx = t % tiles_w;
    t /= tiles_w;
    const int ty = t % tiles_h;
    const int img = t / tiles_h;                          // agent-major image: agent * b + scene
    const int agent = img / b, scene = img % b;
    const int ox0 = tx * TP, oy0 = ty * TP;
    const int ix0 = ox0 * STRIDE - PAD, iy0 = oy0 * STRIDE - PAD;
    const float* xin = x + (static_cast<size_t>(scene) * c_total + c_first + 3 * agent) * h * w;
    float* s_patch = buf;
    float* s_dz = buf + G::kPatchWords;
    // every load of this thread is issued before the first one is used (clamped addresses, no branches around the
    // loads: behind `if (inside)` they waited for each other, 5 dependent latencies per tile on a loader thread)
    constexpr int MAXP = (3 * PW * PW + 63) / 64, MAXD = (TP * TP * 8 + 63) / 64;   // nt >= 64, cout <= 64
    constexpr int PR = 8;   // patch words per round (registers: the 7x7 patch is 21 words per loader thread)
    auto patch_round = [&](int k0, float (&pv)[PR]) {
#pragma unroll
      for (int k = 0; k < PR; ++k) {
        const int i = t0 + (k0 + k) * nt;
        const int cc = i / (PW * PW), py = (i / PW) % PW, px = i % PW;
        const int iy = iy0 + py, ix = ix0 + px;
        const bool ok = k0 + k < MAXP && i < 3 * PW * PW && iy >= 0 && iy < h && ix >= 0 && ix < w;
        const float v = __ldg(xin + (static_cast<size_t>(ok ? cc : 0) * h + (ok ? iy : 0)) * w + (ok ? ix : 0));
        pv[k] = ok ? v : 0.f;
      }
    };
    auto patch_store = [&](int k0, const float (&pv)[PR]) {
#pragma unroll
      for (int k = 0; k < PR; ++k) {
        const int i = t0 + (k0 + k) * nt;
        if (k0 + k < MAXP && i < 3 * PW * PW) s_patch[i] = pv[k];
      }
    };
    float pv[PR];
    patch_round(0, pv);
    const int n_dz = TP * TP * (cout / 8);
    // dz: the hi plane of every item, then (two-plane storages) a second round adds the lo plane - the same thread
    // owns the same items, and one round of raw loads in registers at a time keeps the kernel at two CTAs per SM
    for (int pl = 0; pl < pg; ++pl) {
      uint4 dr[MAXD];
#pragma unroll
      for (int k = 0; k < MAXD; ++k) {
        const int i = t0 + k * nt;
        const int p = i / (cout / 8), gg = i % (cout / 8);
        const int oy = oy0 + p / TP, ox = ox0 + p % TP;
        const bool ok = i < n_dz && oy < ho && ox < wo;
        dr[k] = __ldg(reinterpret_cast<const uint4*>(
            dz + ((static_cast<size_t>(img) * ho + (ok ? oy : 0)) * wo + (ok ? ox : 0)) * (static_cast<size_t>(dz_cs) * pg) +
            pl * dz_cs + dz_co + (ok ? gg : 0) * 8));
      }
      if (pl == 0) patch_store(0, pv);
#pragma unroll
      for (int k = 0; k < MAXD; ++k) {
        const int i = t0 + k * nt;
        if (i >= n_dz) continue;
        const int p = i / (cout / 8), gg = i % (cout / 8);
        const int oy = oy0 + p / TP, ox = ox0 + p % TP;
        const bool ok = oy < ho && ox < wo;
        float v[8];
        const uint32_t* hb = reinterpret_cast<const uint32_t*>(&dr[k]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack_act2(hb[e], f16g);
          v[2 * e] = ok ? f.x : 0.f, v[2 * e + 1] = ok ? f.y : 0.f;
        }
        float4* d4 = reinterpret_cast<float4*>(s_dz + p * cout + gg * 8);
        if (pl == 0) {
          d4[0] = make_float4(v[0], v[1], v[2], v[3]);
          d4[1] = make_float4(v[4], v[5], v[6], v[7]);
        } else {
          const float4 a = d4[0], b = d4[1];
          d4[0] = make_float4(a.x + v[0], a.y + v[1], a.z + v[2], a.w + v[3]);
          d4[1] = make_float4(b.x + v[4], b.y + v[5], b.z + v[6], b.w + v[7]);
        }
      }
    }
#pragma unroll
    for (int k0 = PR; k0 < MAXP; k0 += PR) {   // the rest of a large patch
      patch_round(k0, pv);
      patch_store(k0, pv);
    }
  };

  // One named barrier per tile, hit from both role branches (256 arrivals): buffer it & 1 is staged and everybody is
  // done with the other one. The roles are separate loops so that the loaders do not carry the accumulators (and the
  // others not the raw loads) in their registers.
  auto tile_barrier = [] { asm volatile("bar.sync 1, 256;" ::: "memory"); };
  if (static_cast<int>(blockIdx.x) < n_tiles) stage(blockIdx.x, s_mem, threadIdx.x, 256);
  if (static_cast<int>(threadIdx.x) >= first_loader) {
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      tile_barrier();
      const int next = tile + gridDim.x;
      if (next < n_tiles) stage(next, s_mem + ((it + 1) & 1) * buf_words, threadIdx.x - first_loader, n_loaders);
    }
    return;
  }
  float acc[KS][8];
#pragma unroll
  for (int k = 0; k < KS; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[k][e] = 0.f;
  int it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    tile_barrier();
    const float* cur = s_mem + (it & 1) * buf_words;
    if (active) {
      const float* pp = cur + (ci * PW + kh) * PW;
      const float* dp = cur + G::kPatchWords + g * 8;
#pragma unroll(KS == 7 ? 1 : 2)
      for (int p = part; p < TP * TP; p += parts) {
        const float* px = pp + ((p / TP) * PW + (p % TP)) * STRIDE;
        const float4 d0 = *reinterpret_cast<const float4*>(dp + p * cout);
        const float4 d1 = *reinterpret_cast<const float4*>(dp + p * cout + 4);
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          const float xv = px[k];
          acc[k][0] = fmaf(d0.x, xv, acc[k][0]), acc[k][1] = fmaf(d0.y, xv, acc[k][1]);
          acc[k][2] = fmaf(d0.z, xv, acc[k][2]), acc[k][3] = fmaf(d0.w, xv, acc[k][3]);
          acc[k][4] = fmaf(d1.x, xv, acc[k][4]), acc[k][5] = fmaf(d1.y, xv, acc[k][5]);
          acc[k][6] = fmaf(d1.z, xv, acc[k][6]), acc[k][7] = fmaf(d1.w, xv, acc[k][7]);
        }
      }
    }
  }
  if (active) {
#pragma unroll
    for (int k = 0; k < KS; ++k)
#pragma unroll
      for (int e = 0; e < 8; ++e) atomicAdd(dw + (g * 8 + e) * KK + (ci * KS + kh) * KS + k, acc[k][e]);
  }
}

// ------------------------------------------------------------------------------------------------ packing for dgrad
__global__ void pack_weight_ex_kernel(const float* __restrict__ w, int cout, int cin_real, int cin, int ntaps,
                                      int transposed, int flip, int planes, int cout_pad, __nv_bfloat16* __restrict__ out,
                                      bool f16) {
  const size_t ktot = static_cast<size_t>(ntaps) * cin;
  const size_t total = static_cast<size_t>(cout_pad) * ktot;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int co = idx / ktot;
    const int k = idx % ktot;
    const int tap = k / cin, ci = k % cin;
    const int st = flip ? ntaps - 1 - tap : tap;
    float v = 0.f;
    if (co < cout && ci < cin_real)
      v = transposed ? w[(static_cast<size_t>(ci) * cout + co) * ntaps + st] : w[(static_cast<size_t>(co) * cin_real + ci) * ntaps + st];
    const __nv_bfloat16 hi = float_to_elem(v, f16);
    out[idx] = hi;
    if (planes == 2) out[total + idx] = float_to_elem(v - elem_to_float(hi, f16), f16);
  }
}

// ---- batched setup launches (w2c_pack_conv_weights_batch / w2c_fold_bn_batch): the items ride in the kernel parameters,
// CTAs are dealt to items in proportion to their size (first_block: prefix sums), 8 elements per thread
constexpr int kBatchItems = 48;
constexpr int kPackPerBlock = 256 * 8;
struct PackBatch {
  w2c_pack_item it[kBatchItems];
  int first_block[kBatchItems + 1];
  int n;
};
__global__ void __launch_bounds__(256) pack_weight_batch_kernel(const __grid_constant__ PackBatch b, int planes, bool f16) {
  int i = 0;
  while (i + 1 < b.n && static_cast<int>(blockIdx.x) >= b.first_block[i + 1]) ++i;
  const w2c_pack_item& p = b.it[i];
  const int cout_pad = (p.cout + 15) / 16 * 16;   // w2c_cout_pad
  const int ktot = p.ntaps * p.cin;
  const int total = cout_pad * ktot;          // < 2^31, checked on the host
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(p.packed);
  const int base = (static_cast<int>(blockIdx.x) - b.first_block[i]) * kPackPerBlock;
#pragma unroll 1
  for (int j = 0; j < 8; ++j) {
    const int idx = base + j * 256 + threadIdx.x;
    if (idx >= total) break;
    const int co = idx / ktot, k = idx - co * ktot;
    const int tap = k / p.cin, ci = k - tap * p.cin;
    const int st = p.flip ? p.ntaps - 1 - tap : tap;
    float v = 0.f;
    if (co < p.cout && ci < p.cin_real)
      v = p.transposed ? p.w[(static_cast<size_t>(ci) * p.cout + co) * p.ntaps + st]
                       : p.w[(static_cast<size_t>(co) * p.cin_real + ci) * p.ntaps + st];
    const __nv_bfloat16 hi = float_to_elem(v, f16);
    out[idx] = hi;
    if (planes == 2) out[static_cast<size_t>(total) + idx] = float_to_elem(v - elem_to_float(hi, f16), f16);
  }
}

struct FoldBatch {
  w2c_fold_item it[kBatchItems];
  int n;
};
__global__ void __launch_bounds__(128) fold_bn_batch_kernel(const __grid_constant__ FoldBatch b) {
  const w2c_fold_item& p = b.it[blockIdx.y];
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.cout) return;
  const float bias = p.conv_bias ? p.conv_bias[c] : 0.f;
  if (p.gamma) {   // the arithmetic of fold_bn_kernel (misc.cu), term for term
    const float s = p.gamma[c] / sqrtf(p.var[c] + p.eps);
    p.scale[c] = s;
    p.shift[c] = p.beta[c] + (bias - p.mean[c]) * s;
  } else {
    p.scale[c] = 1.f;
    p.shift[c] = bias;
  }
}

// ------------------------------------------------------------------------------------------------ resnet / simple_decoder
// MaxPool2d(3, 2, 1) backward: every input pixel collects the gradient of the output windows whose FIRST maximum (scan
// order kh, kw: the element torch's max_pool2d_with_indices records) it is. One thread per (input pixel, 8 channels).
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                          __nv_bfloat16* __restrict__ dx, int n, int h, int w, int c,
                                                          int act_f, int act_g) {
  const int groups = c / 8, ho = h / 2, wo = w / 2;
  const size_t total = static_cast<size_t>(n) * h * w * groups;
  const bool f16f = act_is_f16(act_f), f16g = act_is_f16(act_g);
  const int pf = act_planes(act_f), pg = act_planes(act_g);
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    size_t t = idx;
    const int g = t % groups;
    t /= groups;
    const int ix = t % w;
    t /= w;
    const int iy = t % h;
    const int img = static_cast<int>(t / h);
    float out[8], mine[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) out[e] = 0.f;
    load8(x + ((static_cast<size_t>(img) * h + iy) * w + ix) * (static_cast<size_t>(c) * pf) + g * 8, c, pf, f16f, mine);
    // output windows containing (iy, ix): oy with 2*oy - 1 <= iy <= 2*oy + 1
    for (int oy = (iy) / 2; oy <= (iy + 1) / 2; ++oy) {
      if (oy >= ho) continue;
      for (int ox = (ix) / 2; ox <= (ix + 1) / 2; ++ox) {
        if (ox >= wo) continue;
        // is (iy, ix) the first maximum of window (oy, ox)?
        bool first[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) first[e] = true;
        for (int kh = 0; kh < 3; ++kh)
          for (int kw = 0; kw < 3; ++kw) {
            const int yy = 2 * oy - 1 + kh, xx = 2 * ox - 1 + kw;
            if (yy < 0 || yy >= h || xx < 0 || xx >= w || (yy == iy && xx == ix)) continue;
            float o[8];
            load8(x + ((static_cast<size_t>(img) * h + yy) * w + xx) * (static_cast<size_t>(c) * pf) + g * 8, c, pf, f16f, o);
            const bool before = yy < iy || (yy == iy && xx < ix);
#pragma unroll
            for (int e = 0; e < 8; ++e) first[e] = first[e] && (before ? o[e] < mine[e] : o[e] <= mine[e]);
          }
        float d[8];
        load8(dy + ((static_cast<size_t>(img) * ho + oy) * wo + ox) * (static_cast<size_t>(c) * pg) + g * 8, c, pg, f16g, d);
#pragma unroll
        for (int e = 0; e < 8; ++e) out[e] += first[e] ? d[e] : 0.f;
      }
    }
    store8(dx + ((static_cast<size_t>(img) * h + iy) * w + ix) * (static_cast<size_t>(c) * pg) + g * 8, c, pg, f16g, out);
  }
}

// Tiled version for c % 64 == 0 (the resnet18 trunk: 64 channels). The per-pixel kernel above issues up to 33
// dependent 16-byte loads per thread (0.85 ms per 10 frames @256x256x64, ncu: 0.18 TB/s). Here a CTA owns 8x8 input
// pixels x 64 channels: (1) the 11x11 pixel window (one-pixel halo above / left, two below / right) goes to shared
// memory as fp32 with independent coalesced loads, out-of-image pixels as -inf; (2) every one of the 5x5 pooling
// windows that touch the tile finds the scan-order index of its FIRST maximum once per channel (4 bits each, eight
// channels per word); (3) an input pixel sums dy over the <= 4 windows whose index points at it, in the (oy, ox) order
// of the kernel above, so both produce the same bits.
constexpr int kMpT = 4;                       // pooling windows per tile edge -> 2 * kMpT input pixels
constexpr int kMpIn = 2 * kMpT + 3;           // staged pixels per edge
constexpr int kMpWin = kMpT + 1;              // windows per edge that touch the tile
__global__ void __launch_bounds__(256) maxpool_bwd_tiled_kernel(const __nv_bfloat16* __restrict__ x,
                                                                const __nv_bfloat16* __restrict__ dy,
                                                                __nv_bfloat16* __restrict__ dx, int h, int w, int c,
                                                                int act_f, int act_g, int tiles_w, int tiles_h) {
  __shared__ __align__(16) float s_x[kMpIn * kMpIn][64];
  __shared__ uint32_t s_idx[kMpWin * kMpWin][8];
  const int ho = h / 2, wo = w / 2;
  const bool f16f = act_is_f16(act_f), f16g = act_is_f16(act_g);
  const int pf = act_planes(act_f), pg = act_planes(act_g);
  int t = blockIdx.x;
  const int tx = t % tiles_w;
  t /= tiles_w;
  const int ty = t % tiles_h;
  const int img = t / tiles_h;
  const int c0 = blockIdx.y * 64;
  const int oy0 = ty * kMpT, ox0 = tx * kMpT;          // first window of the tile
  const int iy0 = 2 * oy0 - 1, ix0 = 2 * ox0 - 1;      // first staged pixel
  const float ninf = __int_as_float(0xff800000);
  for (int i = threadIdx.x; i < kMpIn * kMpIn * 8; i += blockDim.x) {
    const int g = i & 7, p = i >> 3;
    const int iy = iy0 + p / kMpIn, ix = ix0 + p % kMpIn;
    float v[8];
    if (iy >= 0 && iy < h && ix >= 0 && ix < w)
      load8(x + ((static_cast<size_t>(img) * h + iy) * w + ix) * (static_cast<size_t>(c) * pf) + c0 + g * 8, c, pf, f16f, v);
    else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = ninf;
    }
    float4* d4 = reinterpret_cast<float4*>(&s_x[p][g * 8]);
    d4[0] = make_float4(v[0], v[1], v[2], v[3]);
    d4[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kMpWin * kMpWin * 8; i += blockDim.x) {
    const int g = i & 7, wi = i >> 3;
    const int wy = wi / kMpWin, wx = wi % kMpWin;
    float best[8];
    uint32_t arg = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) best[e] = ninf;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float4* s4 = reinterpret_cast<const float4*>(&s_x[(2 * wy + k / 3) * kMpIn + 2 * wx + k % 3][g * 8]);
      const float4 a = s4[0], b = s4[1];
      const float o[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (k == 0 || o[e] > best[e]) best[e] = o[e], arg = (arg & ~(15u << (4 * e))) | (static_cast<uint32_t>(k) << (4 * e));
    }
    s_idx[wi][g] = arg;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * kMpT * kMpT * 8; i += blockDim.x) {
    const int g = i & 7, p = i >> 3;
    const int ly = p / (2 * kMpT), lx = p % (2 * kMpT);
    const int iy = 2 * oy0 + ly, ix = 2 * ox0 + lx;
    if (iy >= h || ix >= w) continue;
    float out[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) out[e] = 0.f;
    for (int oy = iy / 2; oy <= (iy + 1) / 2; ++oy) {
      if (oy >= ho) continue;
      for (int ox = ix / 2; ox <= (ix + 1) / 2; ++ox) {
        if (ox >= wo) continue;
        const uint32_t k = static_cast<uint32_t>((iy - (2 * oy - 1)) * 3 + (ix - (2 * ox - 1)));
        const uint32_t arg = s_idx[(oy - oy0) * kMpWin + (ox - ox0)][g];
        float d[8];
        load8(dy + ((static_cast<size_t>(img) * ho + oy) * wo + ox) * (static_cast<size_t>(c) * pg) + c0 + g * 8, c, pg, f16g, d);
#pragma unroll
        for (int e = 0; e < 8; ++e) out[e] += ((arg >> (4 * e)) & 15u) == k ? d[e] : 0.f;
      }
    }
    store8(dx + ((static_cast<size_t>(img) * h + iy) * w + ix) * (static_cast<size_t>(c) * pg) + c0 + g * 8, c, pg, f16g, out);
  }
}

// Adjoint of F.interpolate(bilinear, align_corners=False) by an integer factor: dx[n][c][iy][ix] = sum of dy over the
// output pixels that read input (iy, ix), with the forward weights. One thread per input element.
__global__ void __launch_bounds__(256) bilinear_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int nc,
                                                           int h, int w, int f) {
  const size_t total = static_cast<size_t>(nc) * h * w;
  const int ho = h * f, wo = w * f;
  const float inv = 1.f / static_cast<float>(f);
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ix = idx % w, iy = (idx / w) % h;
    const size_t plane = idx / (static_cast<size_t>(w) * h);
    const float* src = dy + plane * ho * wo;
    // output rows whose source coordinate sy = max((oy + 0.5) / f - 0.5, 0) has floor iy - 1 or iy
    const int oy_lo = max(0, (iy - 1) * f), oy_hi = min(ho - 1, (iy + 1) * f + f);
    const int ox_lo = max(0, (ix - 1) * f), ox_hi = min(wo - 1, (ix + 1) * f + f);
    float acc = 0.f;
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      const float sy = fmaxf((oy + 0.5f) * inv - 0.5f, 0.f);
      const int y0 = static_cast<int>(sy);
      const int y1 = min(y0 + 1, h - 1);
      const float ly = sy - y0;
      float wy = 0.f;
      if (y0 == iy) wy += 1.f - ly;
      if (y1 == iy) wy += ly;
      if (wy == 0.f) continue;
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        const float sx = fmaxf((ox + 0.5f) * inv - 0.5f, 0.f);
        const int x0 = static_cast<int>(sx);
        const int x1 = min(x0 + 1, w - 1);
        const float lx = sx - x0;
        float wx = 0.f;
        if (x0 == ix) wx += 1.f - lx;
        if (x1 == ix) wx += lx;
        if (wx != 0.f) acc = fmaf(wy * wx, src[static_cast<size_t>(oy) * wo + ox], acc);
      }
    }
    dx[idx] = acc;
  }
}

// Row version for the x32 up-sampling of simple_decoder (backbone.py:160): one thread per INPUT element leaves 28 k
// threads walking ~64 x 97 output pixels each (0.78 ms per 10 frames, ncu: 0.27 TB/s). The adjoint is separable: one CTA
// per (plane, input row) first folds the ~2f output rows that read the row into one weighted row in shared memory
// (coalesced, independent loads), then one warp per input column folds the ~2f columns of that row.
__device__ __forceinline__ float bilinear_w(int o, int i, int n_in, float inv) {
  const float s = fmaxf((o + 0.5f) * inv - 0.5f, 0.f);
  const int i0 = static_cast<int>(s);
  const int i1 = min(i0 + 1, n_in - 1);
  const float l = s - i0;
  float wgt = 0.f;
  if (i0 == i) wgt += 1.f - l;
  if (i1 == i) wgt += l;
  return wgt;
}
__global__ void __launch_bounds__(256) bilinear_bwd_rows_kernel(const float* __restrict__ dy, float* __restrict__ dx, int h,
                                                                int w, int f) {
  extern __shared__ float s_row[];            // [wo]
  const int ho = h * f, wo = w * f;
  const float inv = 1.f / static_cast<float>(f);
  const int iy = blockIdx.x % h;
  const size_t plane = blockIdx.x / h;
  const float* src = dy + plane * ho * wo;
  const int oy_lo = max(0, (iy - 1) * f), oy_hi = min(ho - 1, (iy + 1) * f + f);
  for (int ox = threadIdx.x; ox < wo; ox += blockDim.x) {
    float acc = 0.f;
#pragma unroll 8
    for (int oy = oy_lo; oy <= oy_hi; ++oy) acc = fmaf(bilinear_w(oy, iy, h, inv), __ldg(src + static_cast<size_t>(oy) * wo + ox), acc);
    s_row[ox] = acc;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int ix = warp; ix < w; ix += blockDim.x >> 5) {
    const int ox_lo = max(0, (ix - 1) * f), ox_hi = min(wo - 1, (ix + 1) * f + f);
    float acc = 0.f;
    for (int ox = ox_lo + lane; ox <= ox_hi; ox += 32) acc = fmaf(bilinear_w(ox, ix, w, inv), s_row[ox], acc);
    acc = warp_sum(acc);
    if (lane == 0) dx[(plane * h + iy) * w + ix] = acc;
  }
}

// dst[n][2i][2j] = src[n][i][j], zeros elsewhere (+ optional accumulate of `add`): the data gradient of a stride-2 1x1
// conv after the 1x1 stride-1 conv with the transposed weight ran on the small grid. One thread per (dst pixel, 8 ch).
__global__ void __launch_bounds__(256) upsample_zero_kernel(const __nv_bfloat16* __restrict__ src, const __nv_bfloat16* add,
                                                            __nv_bfloat16* dst, int n, int h, int w, int c, int act) {
  const int groups = c / 8;
  const size_t total = static_cast<size_t>(n) * h * w * groups;
  const bool f16 = act_is_f16(act);
  const int pl = act_planes(act);
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    size_t t = idx;
    const int g = t % groups;
    t /= groups;
    const int x = t % w;
    t /= w;
    const int y = t % h;
    const int img = static_cast<int>(t / h);
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    if (!(x & 1) && !(y & 1))
      load8(src + ((static_cast<size_t>(img) * (h / 2) + y / 2) * (w / 2) + x / 2) * (static_cast<size_t>(c) * pl) + g * 8, c, pl, f16, v);
    const size_t off = ((static_cast<size_t>(img) * h + y) * w + x) * (static_cast<size_t>(c) * pl) + g * 8;
    if (add) {
      float o[8];
      load8(add + off, c, pl, f16, o);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += o[e];
    }
    store8(dst + off, c, pl, f16, v);
  }
}

// dst (channel slice) = a (+ b): gradient accumulation / concat-slice extraction on NHWC maps
__global__ void __launch_bounds__(256) grad_add_kernel(const __nv_bfloat16* __restrict__ a, int a_cs, int a_co,
                                                       const __nv_bfloat16* b, int b_cs, int b_co, __nv_bfloat16* dst,
                                                       int d_cs, int d_co, size_t n_px, int c, int act) {
  const int groups = c / 8;
  const size_t total = n_px * groups;
  const bool f16 = act_is_f16(act);
  const int pl = act_planes(act);
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int g = idx % groups;
    const size_t px = idx / groups;
    float v[8];
    load8(a + px * (static_cast<size_t>(a_cs) * pl) + a_co + g * 8, a_cs, pl, f16, v);
    if (b) {
      float o[8];
      load8(b + px * (static_cast<size_t>(b_cs) * pl) + b_co + g * 8, b_cs, pl, f16, o);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += o[e];
    }
    store8(dst + px * (static_cast<size_t>(d_cs) * pl) + d_co + g * 8, d_cs, pl, f16, v);
  }
}

}  // namespace
}  // namespace w2c

using namespace w2c;

extern "C" int w2c_kq_mlp_heads_bwd(const void* feat, int32_t act_f, int32_t m, int32_t n_feat, const w2c_mlp_head* heads,
                                    const w2c_mlp_head_grad* grads, int32_t n_heads, const float* ws_fwd, void* dfeat,
                                    int32_t act_g, float* ws, w2c_stream_t stream) {
  W2C_CHECK_ARG(feat && heads && grads && ws_fwd && dfeat && ws, "kq_mlp_bwd: null pointer");
  W2C_CHECK_ARG(n_heads == 1 || n_heads == 2, "kq_mlp_bwd: n_heads=%d (1 or 2)", n_heads);
  W2C_CHECK_ARG(m > 0 && n_feat > 0 && n_feat % 256 == 0, "kq_mlp_bwd: bad sizes m=%d n_feat=%d", m, n_feat);
  W2C_CHECK_ARG(act_valid(act_f) && act_valid(act_g), "kq_mlp_bwd: bad act");
  HeadsBwd hd{};
  hd.n_heads = n_heads;
  for (int i = 0; i < n_heads; ++i) {
    const w2c_mlp_head& h = heads[i];
    const w2c_mlp_head_grad& g = grads[i];
    W2C_CHECK_ARG(h.w0 && h.b0 && h.w1 && h.b1 && h.w2 && h.b2 && h.out_dim > 0, "kq_mlp_bwd: head %d weights", i);
    W2C_CHECK_ARG(g.dout && g.dw0 && g.db0 && g.dw1 && g.db1 && g.dw2 && g.db2, "kq_mlp_bwd: head %d gradients", i);
    hd.h[i] = h, hd.g[i] = g;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  mlp_bwd_hidden_kernel<<<dim3(m, n_heads), 256, 0, s>>>(hd, ws_fwd, ws, m);
  W2C_CHECK_LAUNCH("mlp_bwd_hidden_kernel");
  for (int i = 0; i < n_heads; ++i) {
    float* base = ws + static_cast<size_t>(i) * m * 512;
    const float* h1 = base;
    const float* dh1 = base + static_cast<size_t>(m) * 128;
    const float* dh0 = base + static_cast<size_t>(m) * 256;
    const float* h0 = ws_fwd + static_cast<size_t>(i) * m * 256;
    outer_sum_kernel<<<grid_for(static_cast<size_t>(hd.h[i].out_dim) * 128, 256), 256, 0, s>>>(hd.g[i].dout, h1, m, hd.h[i].out_dim,
                                                                                            128, hd.g[i].dw2, hd.g[i].db2);
    W2C_CHECK_LAUNCH("outer_sum_kernel");
    outer_sum_kernel<<<grid_for(128 * 256, 256), 256, 0, s>>>(dh1, h0, m, 128, 256, hd.g[i].dw1, hd.g[i].db1);
    W2C_CHECK_LAUNCH("outer_sum_kernel");
    outer_sum_kernel<<<1, 256, 0, s>>>(dh0, dh0, m, 256, 1, ws + static_cast<size_t>(n_heads) * m * 512, hd.g[i].db0);
    W2C_CHECK_LAUNCH("outer_sum_kernel");
  }
  mlp_bwd_fc0_wgrad_kernel<<<dim3(ceil_div(n_feat / 8, 128), 32, n_heads), 128, 0, s>>>(
      hd, static_cast<const __nv_bfloat16*>(feat), act_f, ws, m, n_feat);
  W2C_CHECK_LAUNCH("mlp_bwd_fc0_wgrad_kernel");
  mlp_bwd_fc0_dgrad_kernel<<<dim3(ceil_div(n_feat / 8, 128), m), 128, 0, s>>>(hd, ws, m, n_feat,
                                                                             static_cast<__nv_bfloat16*>(dfeat), act_g);
  W2C_CHECK_LAUNCH("mlp_bwd_fc0_dgrad_kernel");
  return W2C_OK;
}

extern "C" int w2c_stem_conv_wgrad(const float* x, const void* dz, float* dw, int32_t ksize, int32_t b, int32_t n_agents,
                                   int32_t c_total, int32_t c_first, int32_t h, int32_t w_px, int32_t cout,
                                   int32_t dz_cstride, int32_t dz_coffset, int32_t act_g, w2c_stream_t stream) {
  W2C_CHECK_ARG(x && dz && dw, "stem_wgrad: null pointer");
  W2C_CHECK_ARG(ksize == 3 || ksize == 7, "stem_wgrad: ksize=%d (3: 3x3 s1, 7: 7x7 s2)", ksize);
  W2C_CHECK_ARG(b > 0 && n_agents > 0 && h > 0 && w_px > 0 && cout > 0 && cout % 8 == 0 && cout <= 64, "stem_wgrad: bad sizes (cout <= 64)");
  W2C_CHECK_ARG(c_first >= 0 && c_first + 3 * n_agents <= c_total, "stem_wgrad: channel range");
  W2C_CHECK_ARG(act_valid(act_g), "stem_wgrad: bad act");
  const int cs = dz_cstride > 0 ? dz_cstride : cout;
  W2C_CHECK_ARG(dz_coffset >= 0 && dz_coffset + cout <= cs && cs % 8 == 0 && dz_coffset % 8 == 0, "stem_wgrad: dz slice");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int stride = ksize == 7 ? 2 : 1;
  W2C_CHECK_ARG(h % stride == 0 && w_px % stride == 0, "stem_wgrad: H, W must be even for the 7x7 s2 stem");
  const int ho = h / stride, wo = w_px / stride;
  const int tiles_w = ceil_div(wo, 8), tiles_h = ceil_div(ho, 8);
  const size_t smem = ksize == 3 ? StemWgradGeom<3, 1>::smem_bytes(cout) : StemWgradGeom<7, 2>::smem_bytes(cout);
  const long long tiles = static_cast<long long>(tiles_w) * tiles_h * b * n_agents;
  W2C_CHECK_ARG(tiles < (1ll << 31), "stem_wgrad: too many tiles");
  const int cap = device_sm_count() * 2;   // two resident CTAs per SM (registers / 43 KB of shared memory each)
  const int blocks = static_cast<int>(tiles < cap ? tiles : cap);
  if (ksize == 3)
    stem_wgrad_kernel<3, 1><<<blocks, 256, smem, s>>>(x, static_cast<const __nv_bfloat16*>(dz), dw, b, n_agents, c_total, c_first, h,
                                                      w_px, cout, cs, dz_coffset, act_g, tiles_w, tiles_h, static_cast<int>(tiles));
  else
    stem_wgrad_kernel<7, 2><<<blocks, 256, smem, s>>>(x, static_cast<const __nv_bfloat16*>(dz), dw, b, n_agents, c_total, c_first, h,
                                                      w_px, cout, cs, dz_coffset, act_g, tiles_w, tiles_h, static_cast<int>(tiles));
  W2C_CHECK_LAUNCH("stem_wgrad_kernel");
  return W2C_OK;
}

extern "C" int w2c_pack_conv_weight_ex(const float* w, int32_t cout, int32_t cin_real, int32_t cin, int32_t ntaps,
                                       int32_t transposed, int32_t flip, int32_t act, void* packed, w2c_stream_t stream) {
  W2C_CHECK_ARG(w && packed, "pack_conv_weight_ex: null pointer");
  W2C_CHECK_ARG(cout > 0 && cin_real > 0 && cin >= cin_real && cin % 64 == 0 && (ntaps == 9 || ntaps == 1),
                "pack_conv_weight_ex: cout=%d cin_real=%d cin=%d ntaps=%d", cout, cin_real, cin, ntaps);
  W2C_CHECK_ARG(act_valid(act), "pack_conv_weight_ex: bad act %d", act);
  const int cout_pad = w2c_cout_pad(cout);
  const size_t total = static_cast<size_t>(cout_pad) * ntaps * cin;
  pack_weight_ex_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, cout, cin_real, cin, ntaps, transposed, flip, act_planes(act), cout_pad, static_cast<__nv_bfloat16*>(packed),
      act_is_f16(act));
  W2C_CHECK_LAUNCH("pack_weight_ex_kernel");
  return W2C_OK;
}

extern "C" int w2c_pack_conv_weights_batch(const w2c_pack_item* items, int32_t n, int32_t act, w2c_stream_t stream) {
  W2C_CHECK_ARG(items && n > 0, "pack_conv_weights_batch: no items");
  W2C_CHECK_ARG(act_valid(act), "pack_conv_weights_batch: bad act %d", act);
  for (int i0 = 0; i0 < n; i0 += kBatchItems) {
    PackBatch b{};
    b.n = n - i0 < kBatchItems ? n - i0 : kBatchItems;
    int blocks = 0;
    for (int i = 0; i < b.n; ++i) {
      const w2c_pack_item& p = items[i0 + i];
      W2C_CHECK_ARG(p.w && p.packed, "pack_conv_weights_batch: item %d: null pointer", i0 + i);
      W2C_CHECK_ARG(p.cout > 0 && p.cin_real > 0 && p.cin >= p.cin_real && p.cin % 64 == 0 && (p.ntaps == 9 || p.ntaps == 1),
                    "pack_conv_weights_batch: item %d: cout=%d cin_real=%d cin=%d ntaps=%d", i0 + i, p.cout, p.cin_real, p.cin, p.ntaps);
      const long long total = static_cast<long long>(w2c_cout_pad(p.cout)) * p.ntaps * p.cin;
      W2C_CHECK_ARG(total < (1ll << 31), "pack_conv_weights_batch: item %d too large", i0 + i);
      b.it[i] = p;
      b.first_block[i] = blocks;
      blocks += static_cast<int>((total + kPackPerBlock - 1) / kPackPerBlock);
    }
    b.first_block[b.n] = blocks;
    pack_weight_batch_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(b, act_planes(act), act_is_f16(act));
    W2C_CHECK_LAUNCH("pack_weight_batch_kernel");
  }
  return W2C_OK;
}

extern "C" int w2c_fold_bn_batch(const w2c_fold_item* items, int32_t n, w2c_stream_t stream) {
  W2C_CHECK_ARG(items && n > 0, "fold_bn_batch: no items");
  for (int i0 = 0; i0 < n; i0 += kBatchItems) {
    FoldBatch b{};
    b.n = n - i0 < kBatchItems ? n - i0 : kBatchItems;
    int cmax = 0;
    for (int i = 0; i < b.n; ++i) {
      const w2c_fold_item& p = items[i0 + i];
      W2C_CHECK_ARG(p.scale && p.shift && p.cout > 0, "fold_bn_batch: item %d: bad arguments", i0 + i);
      const bool any = p.gamma || p.beta || p.mean || p.var;
      W2C_CHECK_ARG(!any || (p.gamma && p.beta && p.mean && p.var), "fold_bn_batch: item %d: BN tensors must be all present or all NULL", i0 + i);
      b.it[i] = p;
      cmax = p.cout > cmax ? p.cout : cmax;
    }
    fold_bn_batch_kernel<<<dim3(ceil_div(cmax, 128), b.n), 128, 0, static_cast<cudaStream_t>(stream)>>>(b);
    W2C_CHECK_LAUNCH("fold_bn_batch_kernel");
  }
  return W2C_OK;
}

extern "C" int w2c_maxpool3x3s2_bwd(const void* x, const void* dy, void* dx, int32_t n, int32_t h, int32_t w_px, int32_t c,
                                    int32_t act_f, int32_t act_g, w2c_stream_t stream) {
  W2C_CHECK_ARG(x && dy && dx && n > 0 && h > 0 && w_px > 0 && h % 2 == 0 && w_px % 2 == 0 && c % 8 == 0, "maxpool_bwd: bad arguments");
  W2C_CHECK_ARG(act_valid(act_f) && act_valid(act_g), "maxpool_bwd: bad act");
  if (c % 64 == 0) {
    const int tiles_w = ceil_div(w_px / 2, kMpT), tiles_h = ceil_div(h / 2, kMpT);
    const long long tiles = static_cast<long long>(tiles_w) * tiles_h * n;
    W2C_CHECK_ARG(tiles < (1ll << 31) && c / 64 < 65536, "maxpool_bwd: too many tiles");
    maxpool_bwd_tiled_kernel<<<dim3(static_cast<unsigned>(tiles), c / 64), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy), static_cast<__nv_bfloat16*>(dx), h, w_px, c,
        act_f, act_g, tiles_w, tiles_h);
    W2C_CHECK_LAUNCH("maxpool_bwd_tiled_kernel");
    return W2C_OK;
  }
  const size_t total = static_cast<size_t>(n) * h * w_px * (c / 8);
  maxpool_bwd_kernel<<<grid_for(total, 256, 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy), static_cast<__nv_bfloat16*>(dx), n, h, w_px, c,
      act_f, act_g);
  W2C_CHECK_LAUNCH("maxpool_bwd_kernel");
  return W2C_OK;
}

extern "C" int w2c_bilinear_up_bwd(const float* dy, float* dx, int32_t n, int32_t c, int32_t h, int32_t w_px, int32_t factor,
                                   w2c_stream_t stream) {
  W2C_CHECK_ARG(dy && dx && n > 0 && c > 0 && h > 0 && w_px > 0 && factor >= 1, "bilinear_bwd: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t row_bytes = static_cast<size_t>(w_px) * factor * sizeof(float);
  if (factor >= 4 && row_bytes <= 40 * 1024 && static_cast<long long>(n) * c * h < (1ll << 31)) {
    bilinear_bwd_rows_kernel<<<n * c * h, 256, row_bytes, s>>>(dy, dx, h, w_px, factor);
    W2C_CHECK_LAUNCH("bilinear_bwd_rows_kernel");
    return W2C_OK;
  }
  const size_t total = static_cast<size_t>(n) * c * h * w_px;
  bilinear_bwd_kernel<<<grid_for(total, 256, 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, dx, n * c, h, w_px, factor);
  W2C_CHECK_LAUNCH("bilinear_bwd_kernel");
  return W2C_OK;
}

extern "C" int w2c_upsample_zero2(const void* src, const void* add, void* dst, int32_t n, int32_t h, int32_t w_px, int32_t c,
                                  int32_t act, w2c_stream_t stream) {
  W2C_CHECK_ARG(src && dst && n > 0 && h > 0 && w_px > 0 && h % 2 == 0 && w_px % 2 == 0 && c % 8 == 0 && act_valid(act),
                "upsample_zero2: bad arguments");
  const size_t total = static_cast<size_t>(n) * h * w_px * (c / 8);
  upsample_zero_kernel<<<grid_for(total, 256, 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(src), static_cast<const __nv_bfloat16*>(add), static_cast<__nv_bfloat16*>(dst), n, h, w_px, c,
      act);
  W2C_CHECK_LAUNCH("upsample_zero_kernel");
  return W2C_OK;
}

extern "C" int w2c_grad_add(const void* a, int32_t a_cstride, int32_t a_coffset, const void* b, int32_t b_cstride,
                            int32_t b_coffset, void* dst, int32_t d_cstride, int32_t d_coffset, int64_t n_px, int32_t c,
                            int32_t act, w2c_stream_t stream) {
  W2C_CHECK_ARG(a && dst && n_px > 0 && c > 0 && c % 8 == 0 && act_valid(act), "grad_add: bad arguments");
  const int acs = a_cstride > 0 ? a_cstride : c, bcs = b_cstride > 0 ? b_cstride : c, dcs = d_cstride > 0 ? d_cstride : c;
  W2C_CHECK_ARG(a_coffset + c <= acs && d_coffset + c <= dcs && (!b || b_coffset + c <= bcs), "grad_add: channel slices");
  W2C_CHECK_ARG(acs % 8 == 0 && bcs % 8 == 0 && dcs % 8 == 0 && a_coffset % 8 == 0 && b_coffset % 8 == 0 && d_coffset % 8 == 0,
                "grad_add: strides / offsets must be multiples of 8");
  const size_t total = static_cast<size_t>(n_px) * (c / 8);
  grad_add_kernel<<<grid_for(total, 256, 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(a), acs, a_coffset, static_cast<const __nv_bfloat16*>(b), bcs, b_coffset,
      static_cast<__nv_bfloat16*>(dst), dcs, d_coffset, static_cast<size_t>(n_px), c, act);
  W2C_CHECK_LAUNCH("grad_add_kernel");
  return W2C_OK;
}

// ------------------------------------------------------------------------------------------------ loss
// cross_entropy2d (ptsemseg/loss/loss.py:5-18: F.cross_entropy over the pixels of fp32 NCHW logits, ignore_index,
// mean over the counted pixels) with its gradient in the same pass: one thread per pixel keeps the <= 32 logits in
// registers, writes softmax - onehot (zero rows for ignored pixels) and adds its loss term and its count to fp64 totals.
// torch's nll_loss 2-D kernels run this reduction in ONE block: 4.1 ms per 10 frames, 11 % of a training step.
namespace w2c {
namespace {
constexpr int kMaxClasses = 32;

__global__ void __launch_bounds__(256) xent2d_kernel(const float* __restrict__ logits, const long long* __restrict__ target,
                                                     int n, int c, size_t hw, long long ignore_index,
                                                     float* __restrict__ dlogits, double* __restrict__ totals) {
  __shared__ double s_loss[8], s_cnt[8];
  const size_t total = static_cast<size_t>(n) * hw;
  double loss = 0.0, cnt = 0.0;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t img = idx / hw, i = idx % hw;
    const float* p = logits + img * c * hw + i;
    float v[kMaxClasses];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k)
      if (k < c) {
        v[k] = p[static_cast<size_t>(k) * hw];
        mx = fmaxf(mx, v[k]);
      }
    float se = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k)
      if (k < c) {
        v[k] = __expf(v[k] - mx);
        se += v[k];
      }
    const long long t = target[idx];
    const bool counted = t != ignore_index && t >= 0 && t < c;
    const float inv = counted ? 1.f / se : 0.f;
    float* d = dlogits + img * c * hw + i;
    float pt = 1.f;
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k)
      if (k < c) {
        const float sm = v[k] * inv;
        if (k == t) pt = sm;
        d[static_cast<size_t>(k) * hw] = counted ? sm - (k == t ? 1.f : 0.f) : 0.f;
      }
    if (counted) loss -= static_cast<double>(__logf(fmaxf(pt, 1e-38f))), cnt += 1.0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    loss += __shfl_xor_sync(0xffffffffu, loss, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) s_loss[warp] = loss, s_cnt[warp] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < 8; ++w) a += s_loss[w], b += s_cnt[w];
    atomicAdd(&totals[0], a);
    atomicAdd(&totals[1], b);
  }
}
}  // namespace
}  // namespace w2c

extern "C" int w2c_cross_entropy2d(const float* logits, const int64_t* target, int32_t n, int32_t c, int64_t hw,
                                   int64_t ignore_index, float* dlogits, double* totals, w2c_stream_t stream) {
  W2C_CHECK_ARG(logits && target && dlogits && totals && n > 0 && hw > 0, "cross_entropy2d: bad arguments");
  W2C_CHECK_ARG(c > 0 && c <= kMaxClasses, "cross_entropy2d: %d classes (at most %d)", c, kMaxClasses);
  const size_t total = static_cast<size_t>(n) * static_cast<size_t>(hw);
  xent2d_kernel<<<grid_for(total, 256, 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, reinterpret_cast<const long long*>(target), n, c, static_cast<size_t>(hw), ignore_index, dlogits, totals);
  W2C_CHECK_LAUNCH("xent2d_kernel");
  return W2C_OK;
}
