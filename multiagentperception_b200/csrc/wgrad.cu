// Weight gradient of the conv / transposed-conv layers on tcgen05 tensor cores (sm_100a): the second half of SURVEY.md
// section 8 f-1 (`loss.backward()` of Trainer_*.train(), ptsemseg/trainer.py:668-670, through conv2DBatchNormRelu /
// deconv2DBatchNormRelu, ptsemseg/models/utils.py:87-120,148-168).
//
//   G[cp][tap][cq] += sum over pixels p of  P[p][cp] * Q[s*p + tap - 1][cq]
//
// with P the operand living on the SMALL pixel grid and Q the one on the (s x larger) grid the taps walk over:
//   Conv2d k3/k1, stride s   P = dL/dy (h/s x w/s, cout channels), Q = x (h x w, cin)    -> G = dW [cout][tap][cin]
//   ConvTranspose2d k3 s2    P = x (h x w, cin),  Q = dL/dy (2h x 2w, cout)              -> G = dW [cin][tap][cout]
// (oy = 2*iy - 1 + kh in the transposed conv, so both are the same contraction with the roles swapped.)
//
// This is a GEMM whose K dimension is the PIXEL axis, while both operands are stored NHWC, i.e. contiguous along
// their M / N (channel) axis: the TMA boxes [64 pixels][64 channels] (128-byte rows, 128B swizzle) are fed to the MMA
// as MN-major operands (instruction-descriptor bits 15 / 16, leading-byte-offset = distance between 64-channel
// groups, stride-byte-offset = distance between 8-pixel K atoms). One CTA owns an output tile
//   M = 128 P-channels  x  N = (3 taps of one filter row) x 64 Q-channels = 192 fp32 TMEM columns
// (two such M halves - 256 P-channels, 384 columns - when the layer has them: the kernel is bound by the L2 -> SM
// traffic of its operand tiles, ncu: 35 % tensor-pipe active at 8.1 TB/s of tile loads with one half, and the second
// half re-uses the Q tiles) and a contiguous range of pixel blocks (split-K); the three taps of a filter row are three
// 64-wide N groups of ONE MMA (LBO walks from tap tile to tap tile). Partial sums are added to G with fp32 atomics.
//   warp 0 (one lane)  TMA producer: per pixel block the P tile (1-2 channel groups x planes) and the 3 shifted Q tiles
//   warp 1 (one lane)  tcgen05.mma M128 x N192 x K16, 4 per block and plane pair (lo*hi, hi*lo, hi*hi)
//   warps 2..5         epilogue: tcgen05.ld, red.global.add.v4.f32 (16-byte reductions: the scalar form issued 24.5 k
//                      L2 atomics per CTA, a fixed ~60 us per launch whatever the layer - ncu launch lists of the step)
// Two-plane storages (value = hi + lo) run the three-pass product of the forward kernels. P and Q must share the element
// type: the instruction descriptor has a format field per operand, but f16 x bf16 faults on B200 (tested).
#include "common.cuh"
#include "ptx.cuh"

namespace w2c {

int encode_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, CUtensorMapL2promotion promo);  // conv_tc.cu

namespace {

constexpr int kPixBlock = 64;                  // K per stage: 64 pixels
constexpr int kTileBytes = kPixBlock * 128;    // one [64 px][64 ch] box: 8 KB
constexpr int kThreads = 192;

struct WgradParams {
  CUtensorMap p_map;
  CUtensorMap q_map[4];
  float* g;
  int cp, cq;
  int p_c0, q_c0;        // channel offset of the hi plane inside a pixel
  int p_lo, q_lo;        // + this = the lo plane (the map's channel stride)
  int npass;             // 1 or 3
  int f16_p, f16_q;
  int ntaps, tpg;        // taps in total (9 / 1), taps per CTA (3 / 1)
  int8_t tap_dw[9], tap_dh[9], tap_map[9];
  int tw, th, tn, tiles_w, tiles_h;
  int kblocks, splits;
  int m_tiles, q_chunks, tap_groups;
  int m_groups;          // 64-channel groups of P a tile really holds (1 when cp == 64)
};

// MH: M halves per CTA (1: 128 P-channels, 2: 256)
template <int PLANES, int MH>
struct WgSmem {
  static constexpr int kStages = PLANES == 1 ? (MH == 1 ? 5 : 4) : 2;
  static constexpr int kPBytes = PLANES * 2 * MH * kTileBytes;
  static constexpr int kQBytes = PLANES * 3 * kTileBytes;
  static constexpr int kTmemCols = MH == 1 ? 256 : 512;
  static constexpr int kStageBytes = kPBytes + kQBytes;
  static constexpr int kBarOff = kStages * kStageBytes;
  static constexpr int kTmemPtrOff = kBarOff + (2 * kStages + 1) * 8;
  static constexpr int kTotal = kTmemPtrOff + 8;
  static constexpr int kDynamicBytes = kTotal + 1024;
};

// MN-major operand tile, 128-byte rows (64 channels) x K pixel rows, TMA 128B swizzle: 8-row atoms of 1024 B along K
// (SBO), 64-channel groups `lbo` bytes apart along M / N.
__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// kind::f16, fp32 accumulate, A and B both MN-major, element type per operand (0 = f16, 1 = bf16)
__device__ __forceinline__ uint32_t make_idesc_mn(uint32_t m, uint32_t n, bool f16_a, bool f16_b) {
  return (1u << 4) | ((f16_a ? 0u : 1u) << 7) | ((f16_b ? 0u : 1u) << 10) | (1u << 15) | (1u << 16) | ((n >> 3) << 17) |
         ((m >> 4) << 24);
}

template <int PLANES, int MH>
__global__ void __launch_bounds__(kThreads, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
  using L = WgSmem<PLANES, MH>;
  constexpr int kTmemCols = L::kTmemCols;
  constexpr int STAGES = L::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kTmemPtrOff);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- output tile and pixel-block range of this CTA
  int t = blockIdx.x;
  const int tap_group = t % p.tap_groups;
  t /= p.tap_groups;
  const int q_chunk = t % p.q_chunks;
  const int m_tile = t / p.q_chunks;
  const int split = blockIdx.y;
  const int kb0 = static_cast<int>(static_cast<long long>(p.kblocks) * split / p.splits);
  const int kb1 = static_cast<int>(static_cast<long long>(p.kblocks) * (split + 1) / p.splits);
  const int cp0 = m_tile * (128 * MH), cq0 = q_chunk * 64, tap0 = tap_group * p.tpg;
  const int n_cols = p.tpg * 64;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&p.p_map);
    ptx::prefetch_tensormap(&p.q_map[0]);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_barrier_init();
  } else if (warp == 1) {
    ptx::tmem_alloc(tmem_ptr, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (ptx::elect_one_sync()) {
      // ===================== TMA producer =====================
      const uint32_t stage_tx = static_cast<uint32_t>(PLANES * (p.m_groups + p.tpg) * kTileBytes);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        int r = kb;
        const int w0 = (r % p.tiles_w) * p.tw;
        r /= p.tiles_w;
        const int h0 = (r % p.tiles_h) * p.th;
        const int i0 = (r / p.tiles_h) * p.tn;
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        ptx::mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
        uint8_t* st = smem + stage * L::kStageBytes;
#pragma unroll
        for (int pl = 0; pl < PLANES; ++pl) {
          for (int g = 0; g < p.m_groups; ++g)
            ptx::tma_load_4d(&p.p_map, &full_bar[stage], st + (pl * 2 * MH + g) * kTileBytes,
                             p.p_c0 + pl * p.p_lo + cp0 + g * 64, w0, h0, i0);
          for (int tp = 0; tp < p.tpg; ++tp) {
            const int tap = tap0 + tp;
            ptx::tma_load_4d(&p.q_map[p.tap_map[tap]], &full_bar[stage], st + L::kPBytes + (pl * 3 + tp) * kTileBytes,
                             p.q_c0 + pl * p.q_lo + cq0, w0 + p.tap_dw[tap], h0 + p.tap_dh[tap], i0);
          }
        }
        if (++stage == STAGES) stage = 0, phase ^= 1;
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one_sync()) {
      // ===================== MMA issuer =====================
      const uint32_t idesc = make_idesc_mn(128, static_cast<uint32_t>(n_cols), p.f16_p != 0, p.f16_q != 0);
      const uint32_t smem_base = ptx::smem_u32(smem);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t accumulate = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint32_t st = smem_base + stage * L::kStageBytes;
        for (int pass = 0; pass < p.npass; ++pass) {
          // corrections first (lo*hi, hi*lo), then hi*hi: the order of the forward kernels
          const int pl_p = (p.npass == 3 && pass == 0) ? 1 : 0;
          const int pl_q = (p.npass == 3 && pass == 1) ? 1 : 0;
          const uint32_t a0 = st + pl_p * 2 * MH * kTileBytes;
          const uint32_t b0 = st + L::kPBytes + pl_q * 3 * kTileBytes;
#pragma unroll
          for (int ks = 0; ks < kPixBlock / 16; ++ks) {
            // 16 pixels = two 8-row atoms = 2048 B further along K
            const uint64_t b_desc = make_sw128_mnmajor_desc(b0 + ks * 2048, kTileBytes);
#pragma unroll
            for (int mh = 0; mh < MH; ++mh)   // the M halves share the Q tiles; each has its own 192 TMEM columns
              ptx::umma_bf16(tmem_base + mh * 192, make_sw128_mnmajor_desc(a0 + mh * 2 * kTileBytes + ks * 2048, kTileBytes),
                             b_desc, idesc, accumulate);
            accumulate = 1;
          }
        }
        ptx::umma_commit(&empty_bar[stage]);
        if (++stage == STAGES) stage = 0, phase ^= 1;
      }
      ptx::umma_commit(tmem_full_bar);
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tc_fence_after();
#pragma unroll 1
    for (int mh = 0; mh < MH; ++mh) {
      const int ch = cp0 + mh * 128 + row;
      const bool row_ok = ch < p.cp;
      float* grow = p.g + (static_cast<size_t>(ch) * p.ntaps + tap0) * p.cq + cq0;
#pragma unroll 1
      for (int c0 = 0; c0 < n_cols; c0 += 32) {
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + mh * 192 + c0, r);
        ptx::tmem_ld_wait();
        if (row_ok) {
          float* dst = grow + static_cast<size_t>(c0 >> 6) * p.cq + (c0 & 63);
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            ptx::red_add_v4_f32(dst + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                __uint_as_float(r[j + 3]));
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

template <int PLANES, int MH>
int launch(const WgradParams& p, cudaStream_t stream) {
  using L = WgSmem<PLANES, MH>;
  static_assert(L::kDynamicBytes <= 227 * 1024, "wgrad stages exceed the shared memory of an SM");
  static DeviceOnce attr_set;
  if (int rc = attr_set.ensure([] {
        return cudaFuncSetAttribute(wgrad_kernel<PLANES, MH>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynamicBytes);
      }, "wgrad_kernel"))
    return rc;
  dim3 grid(p.m_tiles * p.q_chunks * p.tap_groups, p.splits, 1);
  wgrad_kernel<PLANES, MH><<<grid, kThreads, L::kDynamicBytes, stream>>>(p);
  W2C_CHECK_LAUNCH("wgrad_kernel");
  return W2C_OK;
}

}  // namespace
}  // namespace w2c

using namespace w2c;

extern "C" int w2c_conv_wgrad(const w2c_wgrad_args* args, w2c_stream_t stream) {
  if (!args) return set_error(W2C_ERR_INVALID, "wgrad: args is NULL");
  const w2c_wgrad_args& a = *args;
  W2C_CHECK_ARG(a.x && a.dy && a.dw, "wgrad: null pointer argument");
  W2C_CHECK_ARG(reinterpret_cast<uintptr_t>(a.dw) % 16 == 0, "wgrad: dw must be 16-byte aligned (vector reductions)");
  W2C_CHECK_ARG(a.n > 0 && a.h_in > 0 && a.w_in > 0, "wgrad: bad image extent %dx%dx%d", a.n, a.h_in, a.w_in);
  W2C_CHECK_ARG(a.cin > 0 && a.cin % 64 == 0 && a.cout > 0 && a.cout % 64 == 0,
                "wgrad: cin=%d and cout=%d must be multiples of 64 (pad the maps)", a.cin, a.cout);
  // (tcgen05.mma kind::f16 with an f16 A and a bf16 B operand - separate format fields in the instruction descriptor -
  // raises an illegal-instruction fault on B200: both operands must be of one element type)
  W2C_CHECK_ARG(act_valid(a.act_x) && act_valid(a.act_dy) && act_planes(a.act_x) == act_planes(a.act_dy) &&
                    act_is_f16(a.act_x) == act_is_f16(a.act_dy),
                "wgrad: act_x=%d / act_dy=%d must agree in plane count and element type", a.act_x, a.act_dy);
  W2C_CHECK_ARG(a.passes == 0 || a.passes == 1 || a.passes == 3, "wgrad: passes=%d", a.passes);
  const int planes = act_planes(a.act_x);
  const int x_cs = a.x_cstride > 0 ? a.x_cstride : a.cin, dy_cs = a.dy_cstride > 0 ? a.dy_cstride : a.cout;
  W2C_CHECK_ARG(a.x_coffset >= 0 && a.x_coffset + a.cin <= x_cs && a.dy_coffset >= 0 && a.dy_coffset + a.cout <= dy_cs,
                "wgrad: channel slice out of range");
  W2C_CHECK_ARG(x_cs % 8 == 0 && dy_cs % 8 == 0 && a.x_coffset % 8 == 0 && a.dy_coffset % 8 == 0,
                "wgrad: channel strides / offsets must be multiples of 8");

  int s = 1, ntaps = 9;
  switch (a.kind) {
    case W2C_CONV3X3_S1: break;
    case W2C_CONV3X3_S2: s = 2; break;
    case W2C_DECONV3X3_S2: s = 2; break;
    case W2C_CONV1X1_S1: ntaps = 1; break;
    case W2C_CONV1X1_S2: s = 2, ntaps = 1; break;
    default: return set_error(W2C_ERR_INVALID, "wgrad: unknown kind %d", a.kind);
  }
  const bool deconv = a.kind == W2C_DECONV3X3_S2;
  if (s == 2 && !deconv) W2C_CHECK_ARG(a.h_in % 2 == 0 && a.w_in % 2 == 0, "wgrad s2: h_in, w_in must be even");
  // P: small grid, Q: the grid the taps walk over
  const __nv_bfloat16 *pbase, *qbase;
  int hp, wp, hq, wq, p_cs, q_cs;
  WgradParams p{};
  if (!deconv) {
    hp = a.h_in / s, wp = a.w_in / s, hq = a.h_in, wq = a.w_in;
    pbase = static_cast<const __nv_bfloat16*>(a.dy), qbase = static_cast<const __nv_bfloat16*>(a.x);
    p.cp = a.cout, p.cq = a.cin, p_cs = dy_cs, q_cs = x_cs, p.p_c0 = a.dy_coffset, p.q_c0 = a.x_coffset;
    p.f16_p = act_is_f16(a.act_dy), p.f16_q = act_is_f16(a.act_x);
  } else {
    hp = a.h_in, wp = a.w_in, hq = 2 * a.h_in, wq = 2 * a.w_in;
    pbase = static_cast<const __nv_bfloat16*>(a.x), qbase = static_cast<const __nv_bfloat16*>(a.dy);
    p.cp = a.cin, p.cq = a.cout, p_cs = x_cs, q_cs = dy_cs, p.p_c0 = a.x_coffset, p.q_c0 = a.dy_coffset;
    p.f16_p = act_is_f16(a.act_x), p.f16_q = act_is_f16(a.act_dy);
  }
  p.g = a.dw;
  p.p_lo = p_cs, p.q_lo = q_cs;
  p.npass = act_passes(a.act_x, a.passes);
  p.ntaps = ntaps, p.tpg = ntaps == 9 ? 3 : 1;
  p.tap_groups = ntaps / p.tpg;
  for (int kh = 0; kh < (ntaps == 9 ? 3 : 1); ++kh)
    for (int kw = 0; kw < (ntaps == 9 ? 3 : 1); ++kw) {
      const int dr = ntaps == 9 ? kh - 1 : 0, dc = ntaps == 9 ? kw - 1 : 0;
      const int t = ntaps == 9 ? kh * 3 + kw : 0;
      if (s == 1) {
        p.tap_dw[t] = (int8_t)dc, p.tap_dh[t] = (int8_t)dr, p.tap_map[t] = 0;
      } else {
        // Q row 2*py + dr: dr = -1 -> odd plane, index py - 1; 0 -> even plane, py; +1 -> odd plane, py
        p.tap_dw[t] = (int8_t)(dc == -1 ? -1 : 0), p.tap_dh[t] = (int8_t)(dr == -1 ? -1 : 0);
        p.tap_map[t] = (int8_t)((dr & 1) * 2 + (dc & 1));
      }
    }

  // ---- pixel blocks of 64 P pixels
  int tw = wp >= 16 ? 16 : pow2_ceil(wp);
  int th = pow2_ceil(hp);
  if (th > kPixBlock / tw) th = kPixBlock / tw;
  const int tn = kPixBlock / (tw * th);
  p.tw = tw, p.th = th, p.tn = tn;
  p.tiles_w = ceil_div(wp, tw), p.tiles_h = ceil_div(hp, th);
  const int tiles_img = ceil_div(a.n, tn);
  const long long kblocks = static_cast<long long>(p.tiles_w) * p.tiles_h * tiles_img;
  W2C_CHECK_ARG(kblocks < (1ll << 30), "wgrad: too many pixel blocks");
  p.kblocks = static_cast<int>(kblocks);
  W2C_CHECK_ARG(p.cp % 128 == 0 || p.cp == 64, "wgrad: %d channels on the dense operand (64 or a multiple of 128)", p.cp);
  // two M halves per CTA where the layer has 256 P-channels to give AND every CTA still gets a long pixel loop (the
  // epilogue adds 2 x 24.5 k atomics per CTA: measured on B200, 512 -> 512 at 64x64 gains 19 %, at 16x16 it loses 50 %)
  int mh = 1;
  if (p.cp % 256 == 0) {
    const int tiles2 = (p.cp / 256) * (p.cq / 64) * p.tap_groups;
    int splits2 = ceil_div(2 * device_sm_count(), tiles2);
    if (splits2 > p.kblocks) splits2 = p.kblocks;
    if (p.kblocks / splits2 >= 16) mh = 2;
  }
  p.m_tiles = ceil_div(p.cp, 128 * mh), p.q_chunks = p.cq / 64;
  p.m_groups = p.cp >= 128 ? 2 * mh : 1;
  const int out_tiles = p.m_tiles * p.q_chunks * p.tap_groups;
  int splits = ceil_div(2 * device_sm_count(), out_tiles);
  if (splits > p.kblocks) splits = p.kblocks;
  if (splits < 1) splits = 1;
  p.splits = splits;

  const cuuint64_t esz = 2;
  const cuuint32_t box[4] = {64u, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tn};
  {
    const cuuint64_t pix = static_cast<cuuint64_t>(p_cs) * planes;
    const cuuint64_t dims[4] = {pix, (cuuint64_t)wp, (cuuint64_t)hp, (cuuint64_t)a.n};
    const cuuint64_t str[3] = {pix * esz, (cuuint64_t)wp * pix * esz, (cuuint64_t)hp * wp * pix * esz};
    if (int rc = encode_map(&p.p_map, pbase, 4, dims, str, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return rc;
  }
  {
    const cuuint64_t pix = static_cast<cuuint64_t>(q_cs) * planes;
    if (s == 1) {
      const cuuint64_t dims[4] = {pix, (cuuint64_t)wq, (cuuint64_t)hq, (cuuint64_t)a.n};
      const cuuint64_t str[3] = {pix * esz, (cuuint64_t)wq * pix * esz, (cuuint64_t)hq * wq * pix * esz};
      if (int rc = encode_map(&p.q_map[0], qbase, 4, dims, str, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return rc;
      for (int i = 1; i < 4; ++i) p.q_map[i] = p.q_map[0];
    } else {
      for (int ph = 0; ph < 2; ++ph)
        for (int pw = 0; pw < 2; ++pw) {
          const cuuint64_t dims[4] = {pix, (cuuint64_t)wq / 2, (cuuint64_t)hq / 2, (cuuint64_t)a.n};
          const cuuint64_t str[3] = {2 * pix * esz, 2 * (cuuint64_t)wq * pix * esz, (cuuint64_t)hq * wq * pix * esz};
          const __nv_bfloat16* base = qbase + (static_cast<size_t>(ph) * wq + pw) * pix;
          if (int rc = encode_map(&p.q_map[ph * 2 + pw], base, 4, dims, str, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B))
            return rc;
        }
    }
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (mh == 2) return planes == 2 ? launch<2, 2>(p, st) : launch<1, 2>(p, st);
  return planes == 2 ? launch<2, 1>(p, st) : launch<1, 1>(p, st);
}
