// Implicit-GEMM 3x3 / 1x1 convolution and 3x3 stride-2 transposed convolution on tcgen05 tensor cores (sm_100a),
// with the folded-BatchNorm affine, optional residual add and ReLU fused into the TMEM epilogue.
//
// One CTA computes a 128-pixel x BLOCK_N-channel output tile:
//   warp 0 (one lane)  TMA producer: per k-block one 4-D box load of the shifted input window
//                      [tn images x th rows x tw cols x 64 channels] (zero-filled outside the image = padding) and
//                      one 2-D box load of the [BLOCK_N x 64] weight slice, both 128B-swizzled, into a STAGES ring
//   warp 1 (one lane)  issues tcgen05.mma (M=128, N=BLOCK_N, K=16) x4 per k-block, accumulating in TMEM;
//                      tcgen05.commit releases the smem stage / signals the epilogue
//   warps 2..5         epilogue: tcgen05.ld the fp32 accumulators (one output pixel per thread), apply
//                      scale/shift (+residual) (+ReLU), store bf16 NHWC (one or two planes) or fp32 NCHW
// The K loop runs over (pass, tap, 64-channel chunk); pass > 0 only in the bf16x3 parity precision, where the
// same kernel evaluates hi*hi + hi*lo + lo*hi by re-pointing the A / B loads at the lo planes.
#include "conv_plan.cuh"
#include "ptx.cuh"

namespace w2c {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // bf16 elements = 128 bytes = one swizzle row
constexpr int kAStageBytes = kBlockM * kBlockK * 2;
constexpr int kNumThreads = 192;

struct ConvTcParams {
  CUtensorMap a_map[4];
  CUtensorMap b_map;
  ConvPlan plan;
  int tw, th, tn;                   // tile extent in the M-space (tw*th*tn == 128)
  int tiles_w, tiles_h, tiles_img;  // tiles per axis
  int n_tiles;                      // cout_pad / BLOCK_N
};

template <int BLOCK_N, int STAGES>
struct SmemLayout {
  static constexpr int kBStageBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kAOff = 0;
  static constexpr int kBOff = STAGES * kAStageBytes;
  static constexpr int kBarOff = kBOff + STAGES * kBStageBytes;  // full[STAGES], empty[STAGES], tmem_full
  static constexpr int kTmemPtrOff = kBarOff + (2 * STAGES + 1) * 8;
  static constexpr int kScaleOff = kTmemPtrOff + 8;
  static constexpr int kShiftOff = kScaleOff + BLOCK_N * 4;
  static constexpr int kTotal = kShiftOff + BLOCK_N * 4;
  static constexpr int kDynamicBytes = kTotal + 1024;  // slack to align the base to 1024 B
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(kNumThreads, 1) conv_tc_kernel(const __grid_constant__ ConvTcParams p) {
  using L = SmemLayout<BLOCK_N, STAGES>;
  constexpr int kTmemCols = BLOCK_N < 32 ? 32 : BLOCK_N;
  constexpr uint32_t kStageTx = kAStageBytes + L::kBStageBytes;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kTmemPtrOff);
  float* s_scale = reinterpret_cast<float*>(smem + L::kScaleOff);
  float* s_shift = reinterpret_cast<float*>(smem + L::kShiftOff);

  const ConvPlan& pl = p.plan;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- tile coordinates: n-tile fastest so CTAs that share the input window are co-scheduled
  const int cls = blockIdx.y;
  const int n_tile = blockIdx.x % p.n_tiles;
  int m_tile = blockIdx.x / p.n_tiles;
  const int tile_w = m_tile % p.tiles_w;
  m_tile /= p.tiles_w;
  const int tile_h = m_tile % p.tiles_h;
  const int tile_i = m_tile / p.tiles_h;
  const int w0 = tile_w * p.tw, h0 = tile_h * p.th, i0 = tile_i * p.tn;
  const int n0 = n_tile * BLOCK_N;

  const int chunks = pl.cin / kBlockK;
  const int ntaps = pl.ntaps[cls];
  const int npass = pl.npass;
  const int num_kb = npass * ntaps * chunks;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&p.b_map);
    ptx::prefetch_tensormap(&p.a_map[0]);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_barrier_init();
  } else if (warp == 1) {
    ptx::tmem_alloc(tmem_ptr, kTmemCols);
    ptx::tmem_relinquish();
  } else if (warp >= 2) {
    for (int c = threadIdx.x - 64; c < BLOCK_N; c += kNumThreads - 64) {
      const bool ok = n0 + c < pl.cout;
      s_scale[c] = ok ? pl.scale[n0 + c] : 0.f;
      s_shift[c] = ok ? pl.shift[n0 + c] : 0.f;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (ptx::elect_one_sync()) {  // one lane; see ptx::elect_one_sync
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int pass = 0; pass < npass; ++pass) {
        // three-pass order lo*hi, hi*lo, hi*hi (the same as conv_pers_v1.cu: corrections first, while the
        // accumulator is small and the tensor core's truncating adds are harmless)
        const bool a_lo = npass == 3 && pass == 0, b_lo = npass == 3 && pass == 1;
        const int a_c0 = pl.x_coffset + (a_lo ? pl.x_cstride : 0);
        const int b_row = n0 + (b_lo ? pl.cout_pad : 0);
        // canonical K order (ConvPlan::kw_major): (kw, chunk, kh) for 3x3 s1 convs, (tap, chunk) otherwise
        const int n_outer = pl.kw_major ? 3 : ntaps, n_inner = pl.kw_major ? 3 : 1;
        for (int o = 0; o < n_outer; ++o)
          for (int ch = 0; ch < chunks; ++ch)
            for (int i = 0; i < n_inner; ++i) {
              const Tap tp = pl.taps[cls][pl.kw_major ? i * 3 + o : o];
              ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
              ptx::mbar_arrive_expect_tx(&full_bar[stage], kStageTx);
              ptx::tma_load_4d(&p.a_map[tp.map], &full_bar[stage], smem + L::kAOff + stage * kAStageBytes,
                               a_c0 + ch * kBlockK, w0 + tp.dw, h0 + tp.dh, i0);
              ptx::tma_load_2d(&p.b_map, &full_bar[stage], smem + L::kBOff + stage * L::kBStageBytes,
                               tp.wtap * pl.cin + ch * kBlockK, b_row);
              if (++stage == STAGES) stage = 0, phase ^= 1;
            }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one_sync()) {  // one lane; see ptx::elect_one_sync
      // ===================== MMA issuer =====================
      const uint32_t idesc = ptx::make_idesc_16(kBlockM, BLOCK_N, act_is_f16(pl.act));
      const uint64_t a_desc0 = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + L::kAOff));
      const uint64_t b_desc0 = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + L::kBOff));
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint64_t a_desc = a_desc0 + static_cast<uint64_t>((stage * kAStageBytes) >> 4);
        const uint64_t b_desc = b_desc0 + static_cast<uint64_t>((stage * L::kBStageBytes) >> 4);
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          // advance 16 bf16 = 32 B along K inside the swizzle atom: +2 in the (addr >> 4) field
          ptx::umma_bf16(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
        }
        ptx::umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs above have read it
        if (++stage == STAGES) stage = 0, phase ^= 1;
      }
      ptx::umma_commit(tmem_full_bar);  // accumulator complete
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;
    const int lw = row % p.tw;
    const int lh = (row / p.tw) % p.th;
    const int li = row / (p.tw * p.th);
    const int mw = w0 + lw, mh = h0 + lh, img = i0 + li;
    const bool valid = mw < pl.wm && mh < pl.hm && img < pl.n_img;
    const int oh = mh * pl.out_s + pl.cls_oh[cls];
    const int ow = mw * pl.out_s + pl.cls_ow[cls];
    const size_t pix = (static_cast<size_t>(img) * pl.out_h + oh) * pl.out_w + ow;

    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tc_fence_after();

    constexpr int kChunk = BLOCK_N < 32 ? 16 : 32;
#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_N; c0 += kChunk) {
      uint32_t r[kChunk];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0;
      if constexpr (kChunk == 32)
        ptx::tmem_ld_32x32b_x32(taddr, r);
      else
        ptx::tmem_ld_32x32b_x16(taddr, r);
      ptx::tmem_ld_wait();
      if (n0 + c0 >= pl.cout) break;  // warp-uniform: padded channels only

      float v[kChunk];
#pragma unroll
      for (int j = 0; j < kChunk; ++j) v[j] = fmaf(__uint_as_float(r[j]), s_scale[c0 + j], s_shift[c0 + j]);

      if (pl.out_fmt == W2C_OUT_NCHW_F32) {
        if (valid) {
          float* y = static_cast<float*>(pl.y);
          const size_t plane = static_cast<size_t>(pl.out_h) * pl.out_w;
          const size_t base = static_cast<size_t>(img) * pl.cout * plane + static_cast<size_t>(oh) * pl.out_w + ow;
#pragma unroll
          for (int j = 0; j < kChunk; ++j) {
            const int c = n0 + c0 + j;
            if (c < pl.cout) y[base + c * plane] = pl.relu ? fmaxf(v[j], 0.f) : v[j];
          }
        }
      } else if (valid) {
        // cout % 8 == 0 (checked on the host), so 8-channel groups are either fully valid or fully padding
        __nv_bfloat16* ypix = static_cast<__nv_bfloat16*>(pl.y) + pix * pl.y_pix + pl.y_coffset + n0 + c0;
        const __nv_bfloat16* rpix =
            pl.residual ? pl.residual + pix * pl.y_pix + pl.y_coffset + n0 + c0 : nullptr;
#pragma unroll
        for (int g = 0; g < kChunk / 8; ++g) {
          if (n0 + c0 + g * 8 >= pl.cout) break;
          if (rpix) {
            const uint4 rv = *reinterpret_cast<const uint4*>(rpix + g * 8);
            const uint32_t* rb = reinterpret_cast<const uint32_t*>(&rv);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = unpack_act2(rb[j], act_is_f16(pl.act));
              v[g * 8 + 2 * j] += f.x, v[g * 8 + 2 * j + 1] += f.y;
            }
            if (act_planes(pl.act) == 2) {
              const uint4 rl = *reinterpret_cast<const uint4*>(rpix + pl.y_cstride + g * 8);
              const uint32_t* lb = reinterpret_cast<const uint32_t*>(&rl);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = unpack_act2(lb[j], act_is_f16(pl.act));
                v[g * 8 + 2 * j] += f.x, v[g * 8 + 2 * j + 1] += f.y;
              }
            }
          }
          uint4 hv, lv;
          uint32_t* hb = reinterpret_cast<uint32_t*>(&hv);
          uint32_t* lb = reinterpret_cast<uint32_t*>(&lv);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float a = v[g * 8 + 2 * j], b = v[g * 8 + 2 * j + 1];
            if (pl.relu) a = fmaxf(a, 0.f), b = fmaxf(b, 0.f);
            split_act2(a, b, act_is_f16(pl.act), hb[j], lb[j]);
          }
          *reinterpret_cast<uint4*>(ypix + g * 8) = hv;
          if (act_planes(pl.act) == 2) *reinterpret_cast<uint4*>(ypix + pl.y_cstride + g * 8) = lv;
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

// -------------------------------------------------------------------------------------------- host side

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      ptr = nullptr;
    return reinterpret_cast<EncodeTiledFn>(ptr);
  }();
  return fn;
}

}  // namespace

int encode_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, CUtensorMapL2promotion promo) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return set_error(W2C_ERR_DRIVER, "cuTensorMapEncodeTiled driver entry point not available");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides_bytes, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(W2C_ERR_DRIVER, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return W2C_OK;
}

// fp32 tensor, no swizzle (the NCHW logits written through TMA)
int encode_map_f32(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims,
                   const cuuint64_t* strides_bytes, const cuuint32_t* box) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return set_error(W2C_ERR_DRIVER, "cuTensorMapEncodeTiled driver entry point not available");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(W2C_ERR_DRIVER, "cuTensorMapEncodeTiled (f32) failed with CUresult %d", (int)r);
  return W2C_OK;
}

namespace {

int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

template <int BLOCK_N, int STAGES>
int launch(const ConvTcParams& p, cudaStream_t stream) {
  using L = SmemLayout<BLOCK_N, STAGES>;
  static DeviceOnce attr_set;  // per device, not per process (common.cuh)
  if (int rc = attr_set.ensure([] {
        return cudaFuncSetAttribute(conv_tc_kernel<BLOCK_N, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    L::kDynamicBytes);
      }, "conv_tc_kernel"))
    return rc;
  dim3 grid(p.tiles_w * p.tiles_h * p.tiles_img * p.n_tiles, p.plan.num_classes, 1);
  conv_tc_kernel<BLOCK_N, STAGES><<<grid, kNumThreads, L::kDynamicBytes, stream>>>(p);
  W2C_CHECK_LAUNCH("conv_tc_kernel");
  return W2C_OK;
}

}  // namespace

int conv_tc_forward(const w2c_conv_args& a, const ConvPlan& plan, cudaStream_t stream) {
  ConvTcParams p;
  p.plan = plan;
  const int planes = act_planes(plan.act);

  // ---- tile shape in the M-space
  int tw = plan.wm >= 16 ? 16 : pow2_ceil(plan.wm);
  int th = pow2_ceil(plan.hm);
  if (th > kBlockM / tw) th = kBlockM / tw;
  int tn = kBlockM / (tw * th);
  p.tw = tw, p.th = th, p.tn = tn;
  p.tiles_w = ceil_div(plan.wm, tw);
  p.tiles_h = ceil_div(plan.hm, th);
  p.tiles_img = ceil_div(plan.n_img, tn);

  // ---- BLOCK_N
  int bn = a.block_n;
  if (bn == 0) {
    // N = 256 halves the A-operand shared-memory reads per FLOP (SS-mode MMAs re-read the 4 KB A slice for every
    // instruction: at N <= 128 the kernel is smem-bandwidth bound, see profiles/r1_conv_sweep.md)
    if (plan.cout_pad % 256 == 0 && static_cast<long long>(plan.n_img) * plan.hm * plan.wm >= 128 * 148)
      bn = 256;
    else if (plan.cout_pad % 128 == 0)
      bn = 128;
    else if (plan.cout_pad % 64 == 0)
      bn = 64;
    else if (plan.cout_pad % 32 == 0)
      bn = 32;
    else
      bn = 16;
  }
  W2C_CHECK_ARG(bn == 16 || bn == 32 || bn == 64 || bn == 128 || bn == 256, "conv: block_n=%d not supported", bn);
  W2C_CHECK_ARG(plan.cout_pad % bn == 0, "conv: block_n=%d does not divide cout_pad=%d", bn, plan.cout_pad);
  p.n_tiles = plan.cout_pad / bn;

  // ---- input tensor maps (bf16, NHWC; dims innermost first: C, W, H, N)
  const cuuint64_t esz = 2;
  const cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tn};
  if (plan.in_s == 1) {
    const cuuint64_t dims[4] = {(cuuint64_t)plan.x_pix, (cuuint64_t)plan.in_w, (cuuint64_t)plan.in_h,
                                (cuuint64_t)plan.n_img};
    const cuuint64_t str[3] = {plan.x_pix * esz, (cuuint64_t)plan.in_w * plan.x_pix * esz,
                               (cuuint64_t)plan.in_h * plan.in_w * plan.x_pix * esz};
    int rc = encode_map(&p.a_map[0], plan.x, 4, dims, str, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (rc) return rc;
    for (int i = 1; i < 4; ++i) p.a_map[i] = p.a_map[0];
  } else {
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        const cuuint64_t dims[4] = {(cuuint64_t)plan.x_pix, (cuuint64_t)plan.in_w / 2, (cuuint64_t)plan.in_h / 2,
                                    (cuuint64_t)plan.n_img};
        const cuuint64_t str[3] = {2 * plan.x_pix * esz, 2 * (cuuint64_t)plan.in_w * plan.x_pix * esz,
                                   (cuuint64_t)plan.in_h * plan.in_w * plan.x_pix * esz};
        const __nv_bfloat16* base = plan.x + (static_cast<size_t>(ph) * plan.in_w + pw) * plan.x_pix;
        int rc = encode_map(&p.a_map[ph * 2 + pw], base, 4, dims, str, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
        if (rc) return rc;
      }
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)plan.ktot, (cuuint64_t)plan.cout_pad * planes};
    const cuuint64_t str[1] = {plan.ktot * esz};
    const cuuint32_t bbox[2] = {(cuuint32_t)kBlockK, (cuuint32_t)bn};
    int rc = encode_map(&p.b_map, plan.w, 2, dims, str, bbox, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (rc) return rc;
  }

  switch (bn) {
    case 256: return launch<256, 4>(p, stream);
    case 128: return launch<128, 3>(p, stream);
    case 64: return launch<64, 4>(p, stream);
    case 32: return launch<32, 4>(p, stream);
    default: return launch<16, 4>(p, stream);
  }
}

// ============================================================================================ SIMT cross-check
// Straightforward CUDA-core evaluation of the same plan on the same packed operands (fp32 accumulate). One thread
// per (class, pixel, output channel). Test infrastructure for full-size on-GPU comparisons; not the product path.
namespace {

__global__ void conv_simt_kernel(const ConvPlan pl) {
  const int planes = act_planes(pl.act);
  const size_t per_class = static_cast<size_t>(pl.n_img) * pl.hm * pl.wm * pl.cout;
  const size_t total = per_class * pl.num_classes;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    size_t t = idx;
    const int co = t % pl.cout;
    t /= pl.cout;
    const int mw = t % pl.wm;
    t /= pl.wm;
    const int mh = t % pl.hm;
    t /= pl.hm;
    const int img = t % pl.n_img;
    const int cls = t / pl.n_img;
    float acc = 0.f;
    for (int tp = 0; tp < pl.ntaps[cls]; ++tp) {
      const Tap tap = pl.taps[cls][tp];
      const int ih = mh * pl.in_s + tap.ih_off, iw = mw * pl.in_s + tap.iw_off;
      if (ih < 0 || ih >= pl.in_h || iw < 0 || iw >= pl.in_w) continue;
      const __nv_bfloat16* xp = pl.x + ((static_cast<size_t>(img) * pl.in_h + ih) * pl.in_w + iw) * pl.x_pix + pl.x_coffset;
      const __nv_bfloat16* wp = pl.w + static_cast<size_t>(co) * pl.ktot + tap.wtap * pl.cin;
      const __nv_bfloat16* wl = wp + static_cast<size_t>(pl.cout_pad) * pl.ktot;
      for (int ci = 0; ci < pl.cin; ++ci) {
        const bool f16 = act_is_f16(pl.act);
        const float xh = elem_to_float(xp[ci], f16);
        const float wh = elem_to_float(wp[ci], f16);
        acc = fmaf(xh, wh, acc);
        if (pl.npass == 3) {
          acc = fmaf(xh, elem_to_float(wl[ci], f16), acc);
          acc = fmaf(elem_to_float(xp[pl.x_cstride + ci], f16), wh, acc);
        }
      }
    }
    float v = fmaf(acc, pl.scale[co], pl.shift[co]);
    const int oh = mh * pl.out_s + pl.cls_oh[cls], ow = mw * pl.out_s + pl.cls_ow[cls];
    if (pl.out_fmt == W2C_OUT_NCHW_F32) {
      if (pl.relu) v = fmaxf(v, 0.f);
      static_cast<float*>(pl.y)[((static_cast<size_t>(img) * pl.cout + co) * pl.out_h + oh) * pl.out_w + ow] = v;
    } else {
      const size_t pix = (static_cast<size_t>(img) * pl.out_h + oh) * pl.out_w + ow;
      if (pl.residual) v += act_load(pl.residual + pix * pl.y_pix + pl.y_coffset, co, pl.y_cstride, pl.act);
      if (pl.relu) v = fmaxf(v, 0.f);
      act_store(static_cast<__nv_bfloat16*>(pl.y) + pix * pl.y_pix + pl.y_coffset, co, pl.y_cstride, pl.act, v);
    }
  }
}

}  // namespace

int conv_simt_forward(const ConvPlan& plan, cudaStream_t stream) {
  const size_t total = static_cast<size_t>(plan.n_img) * plan.hm * plan.wm * plan.cout * plan.num_classes;
  const int threads = 256;
  const size_t want = (total + threads - 1) / threads;
  const int blocks = static_cast<int>(want < 148 * 64 ? (want ? want : 1) : 148 * 64);
  conv_simt_kernel<<<blocks, threads, 0, stream>>>(plan);
  W2C_CHECK_LAUNCH("conv_simt_kernel");
  return W2C_OK;
}

}  // namespace w2c

namespace w2c {
// Measured per-layer choice between the persistent kernel and the one-tile-per-CTA kernel
// (profiles/r1_conv_sweep_v2_persistent.md): with narrow outputs (<= 64 channels) a stride-1 3x3 conv issues MMAs too
// short for ONE issuing thread per SM to keep the pipe busy, and three co-resident one-tile CTAs (three issuers)
// win; the same holds when there are too few tiles to give every persistent CTA at least two.
bool conv_persistent_preferred(const ConvPlan& plan) {
  const long long m_tiles = (static_cast<long long>(plan.n_img) * plan.hm * plan.wm + 127) / 128;
  // at least one 128-wide tile per SM (with the lean elect.sync issue loops the persistent kernel wins from there on;
  // the 8x8 / 4x4 policy-tail layers stay on the one-tile kernel)
  const int bn = plan.cout_pad % 128 == 0 ? 128 : 64;
  const long long tiles = m_tiles * plan.num_classes * ((plan.cout_pad + bn - 1) / bn);
  if (tiles < 148) return false;
  // narrow stride-1 convs: only with row-halo stages (12 MMAs per barrier round trip) does one issuer keep up;
  // the 11-channel logits layer additionally runs three persistent CTAs (issuers) per SM: 0.53 vs 0.98 ms
  const bool row_halo = plan.num_classes == 1 && plan.in_s == 1 && plan.ntaps[0] == 9 && plan.wm >= 16 && plan.hm >= 8;
  if (plan.num_classes == 1 && plan.in_s == 1 && plan.cout_pad <= 64) return row_halo;
  return true;
}
}  // namespace w2c

extern "C" int w2c_conv_fuses_bn_sums(const w2c_conv_args* args) {
  if (!args) return w2c::set_error(W2C_ERR_INVALID, "conv: args is NULL");
  w2c::ConvPlan plan;
  if (int rc = w2c::build_conv_plan(*args, plan)) return rc;
  const int impl = args->impl & 0xff;
  const bool pers = impl == W2C_IMPL_TC_PERSIST ||
                    (impl == W2C_IMPL_TCGEN05 && w2c::conv_persv1_supported(plan) && w2c::conv_persistent_preferred(plan));
  return pers && w2c::conv_persv1_fuses_bn_sums(*args, plan) ? 1 : 0;
}

extern "C" int w2c_conv_bnrelu_fwd(const w2c_conv_args* args, w2c_stream_t stream) {
  if (!args) return w2c::set_error(W2C_ERR_INVALID, "conv: args is NULL");
  w2c::ConvPlan plan;
  int rc = w2c::build_conv_plan(*args, plan);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int impl = args->impl & 0xff;
  if (plan.bn_sums && w2c_conv_fuses_bn_sums(args) != 1)
    return w2c::set_error(W2C_ERR_UNSUPPORTED, "conv: this launch cannot accumulate bn_sums (ask w2c_conv_fuses_bn_sums)");
  if (plan.labels && impl != W2C_IMPL_TCGEN05 && impl != W2C_IMPL_TC_PERSIST)
    return w2c::set_error(W2C_ERR_UNSUPPORTED, "conv: impl %d has no label-map epilogue", impl);
  switch (impl) {
    case W2C_IMPL_SIMT:  // CUDA-core cross-check of the tensor-core kernels (tests only)
      return w2c::conv_simt_forward(plan, s);
    case W2C_IMPL_TC_TAPS:
      return w2c::conv_tc_forward(*args, plan, s);
    case W2C_IMPL_TC_PERSIST:
      if (!w2c::conv_persv1_supported(plan))
        return w2c::set_error(W2C_ERR_UNSUPPORTED, "conv: the persistent kernel needs cout <= 512");
      return w2c::conv_persv1_forward(*args, plan, s);
    case W2C_IMPL_TCGEN05: {
      // production dispatch: the persistent kernel wherever a layer has a tile per SM (and for the fused label map,
      // which lives in its logits epilogue), the one-tile-per-CTA kernel for the sub-wave layers
      if (plan.labels) {
        if (!w2c::conv_persv1_supported(plan))
          return w2c::set_error(W2C_ERR_UNSUPPORTED, "conv: label map needs the persistent kernel (cout <= 512)");
        return w2c::conv_persv1_forward(*args, plan, s);
      }
      if (w2c::conv_persv1_supported(plan) && w2c::conv_persistent_preferred(plan))
        return w2c::conv_persv1_forward(*args, plan, s);
      return w2c::conv_tc_forward(*args, plan, s);
    }
    default:
      return w2c::set_error(W2C_ERR_INVALID, "conv: unknown impl %d (the halo / row-halo experiments live in experiments/csrc)", args->impl);
  }
}
