// Halo-reuse variant of the tcgen05 implicit-GEMM convolution (3x3 stride-1 conv and 3x3 stride-2 transposed conv).
//
// conv_tc.cu loads one shifted 128-pixel input window per filter tap, i.e. the same pixels cross L2 -> SM nine
// times; ncu shows those layers pinned on L2 bandwidth (lts throughput 75 %, 19 TB/s of xbar reads) and the
// 64/128-channel layers far from both rooflines. Here a CTA loads the input tile ONCE per 64-channel chunk, with
// its halo, as a [rows+2][PW] pixel box (PW = TW + 2), 128 B per pixel, 128B-swizzled by TMA. Because the rows
// are stored at the box pitch, the window of filter tap (dy, dx) is the SAME buffer advanced by (dy*PW + dx)
// pixels: an affine shift, so every tap's A operand is just another UMMA shared-memory descriptor start address
// (base-offset field = address bits [7,10), because the start is no longer 1024-byte aligned). The price is
// that the GEMM M index runs over PW, not TW, columns: 2 of every PW accumulator rows are wrap-around garbage
// that the epilogue drops (TW = 16, PW = 18, 7 rows -> 112 useful of 128 MMA rows).
// Weights stream through their own smem ring, one [BLOCK_N x 64] tile per (tap, chunk).
// A transposed conv computes all four output-parity classes from one halo tile into four TMEM accumulators.
//   warp 0: TMA producer   warp 1: MMA issuer + TMEM owner   warps 2-5: epilogue (same as conv_tc.cu)
#include "conv_plan.cuh"
#include "ptx.cuh"

namespace w2c {

int encode_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, CUtensorMapL2promotion promo);

namespace {

constexpr int kTW = 16, kPW = 18, kTH = 7;          // valid cols, pitch, rows per 128-row MMA sub-tile
constexpr int kBlockK = 64;
constexpr int kNumThreads = 192;
constexpr int kMaxCls = 4;

struct HaloParams {
  CUtensorMap a_map;
  CUtensorMap b_map;
  ConvPlan plan;
  int mt;                 // 128-row sub-tiles stacked vertically per CTA (1 or 2)
  int halo_rows;          // mt*kTH + 2
  int a_stage_bytes;      // halo tile bytes, padded to 1024
  int a_stages, b_stages;
  int tiles_w, tiles_h;   // per image
  int n_tiles;            // cout_pad / BLOCK_N
  int oy, ox;             // halo origin shift: 1 for conv (taps -1..1), 0 for deconv (taps 0..1)
  int base_offset_mode;   // 0 (default, verified on B200): base_offset field stays 0; 1: (addr >> 7) & 7 (WRONG results)
};

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, int mode) {
  uint64_t d = ptx::make_sw128_kmajor_desc(smem_addr);
  if (mode) d |= static_cast<uint64_t>((smem_addr >> 7) & 7u) << 49;
  return d;
}

template <int BLOCK_N>
__global__ void __launch_bounds__(kNumThreads, 1) conv_halo_kernel(const __grid_constant__ HaloParams p) {
  constexpr int kBStageBytes = BLOCK_N * kBlockK * 2;
  const ConvPlan& pl = p.plan;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + p.a_stages * p.a_stage_bytes;
  uint8_t* tail = smem_b + p.b_stages * kBStageBytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* a_empty = a_full + p.a_stages;
  uint64_t* b_full = a_empty + p.a_stages;
  uint64_t* b_empty = b_full + p.b_stages;
  uint64_t* tmem_full = b_empty + p.b_stages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* s_scale = reinterpret_cast<float*>(tmem_ptr + 2);
  float* s_shift = s_scale + BLOCK_N;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncls = pl.num_classes;
  const int tmem_cols_needed = p.mt * ncls * BLOCK_N;
  const uint32_t tmem_cols = tmem_cols_needed <= 32 ? 32 : tmem_cols_needed <= 64 ? 64 : tmem_cols_needed <= 128 ? 128
                             : tmem_cols_needed <= 256 ? 256 : 512;

  const int n_tile = blockIdx.x % p.n_tiles;
  int t = blockIdx.x / p.n_tiles;
  const int tile_w = t % p.tiles_w;
  t /= p.tiles_w;
  const int tile_h = t % p.tiles_h;
  const int img = t / p.tiles_h;
  const int w0 = tile_w * kTW, h0 = tile_h * (kTH * p.mt);
  const int n0 = n_tile * BLOCK_N;
  const int chunks = pl.cin / kBlockK;
  const int npass = pl.act == W2C_ACT_BF16X2 ? 3 : 1;
  int taps_total = 0;
  for (int c = 0; c < ncls; ++c) taps_total += pl.ntaps[c];

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&p.a_map);
    ptx::prefetch_tensormap(&p.b_map);
    for (int s = 0; s < p.a_stages; ++s) ptx::mbar_init(&a_full[s], 1), ptx::mbar_init(&a_empty[s], 1);
    for (int s = 0; s < p.b_stages; ++s) ptx::mbar_init(&b_full[s], 1), ptx::mbar_init(&b_empty[s], 1);
    ptx::mbar_init(tmem_full, 1);
    ptx::fence_barrier_init();
  } else if (warp == 1) {
    ptx::tmem_alloc(tmem_ptr, tmem_cols);
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
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      const uint32_t a_bytes = static_cast<uint32_t>(p.halo_rows) * kPW * 128u;
      for (int pass = 0; pass < npass; ++pass) {
        const int a_c0 = pl.x_coffset + (pass == 2 ? pl.x_cstride : 0);
        const int b_row = n0 + (pass == 1 ? pl.cout_pad : 0);
        for (int ch = 0; ch < chunks; ++ch) {
          ptx::mbar_wait(&a_empty[sa], pa ^ 1);
          ptx::mbar_arrive_expect_tx(&a_full[sa], a_bytes);
          ptx::tma_load_4d(&p.a_map, &a_full[sa], smem_a + sa * p.a_stage_bytes, a_c0 + ch * kBlockK, w0 - p.ox,
                           h0 - p.oy, img);
          if (++sa == p.a_stages) sa = 0, pa ^= 1;
          for (int cls = 0; cls < ncls; ++cls)
            for (int tp = 0; tp < pl.ntaps[cls]; ++tp) {
              ptx::mbar_wait(&b_empty[sb], pb ^ 1);
              ptx::mbar_arrive_expect_tx(&b_full[sb], kBStageBytes);
              ptx::tma_load_2d(&p.b_map, &b_full[sb], smem_b + sb * kBStageBytes,
                               pl.taps[cls][tp].wtap * pl.cin + ch * kBlockK, b_row);
              if (++sb == p.b_stages) sb = 0, pb ^= 1;
            }
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one_sync()) {  // one lane; see ptx::elect_one_sync
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = ptx::make_idesc_bf16(128, BLOCK_N);
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      uint32_t started = 0;  // bit (cls*2 + mt) set once that accumulator has received its first MMA
      for (int pass = 0; pass < npass; ++pass)
        for (int ch = 0; ch < chunks; ++ch) {
          ptx::mbar_wait(&a_full[sa], pa);
          ptx::tc_fence_after();
          const uint32_t a_base = ptx::smem_u32(smem_a + sa * p.a_stage_bytes);
          for (int cls = 0; cls < ncls; ++cls)
            for (int tp = 0; tp < pl.ntaps[cls]; ++tp) {
              const Tap tap = pl.taps[cls][tp];
              ptx::mbar_wait(&b_full[sb], pb);
              ptx::tc_fence_after();
              const uint64_t b_desc = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem_b + sb * kBStageBytes));
              for (int mt = 0; mt < p.mt; ++mt) {
                const uint32_t a_addr =
                    a_base + static_cast<uint32_t>(((mt * kTH + tap.dh + p.oy) * kPW + tap.dw + p.ox) * 128);
                const uint32_t acc = tmem_base + static_cast<uint32_t>((cls * p.mt + mt) * BLOCK_N);
                const uint32_t bit = 1u << (cls * 2 + mt);
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k) {
                  const uint64_t a_desc = make_desc(a_addr + k * 32, p.base_offset_mode);
                  ptx::umma_bf16(acc, a_desc, b_desc + 2 * k, idesc, ((started & bit) != 0) || k > 0);
                }
                started |= bit;
              }
              ptx::umma_commit(&b_empty[sb]);
              if (++sb == p.b_stages) sb = 0, pb ^= 1;
            }
          ptx::umma_commit(&a_empty[sa]);
          if (++sa == p.a_stages) sa = 0, pa ^= 1;
        }
      ptx::umma_commit(tmem_full);
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int r = row / kPW, c = row - r * kPW;
    ptx::mbar_wait(tmem_full, 0);
    ptx::tc_fence_after();
    constexpr int kChunk = BLOCK_N < 32 ? 16 : 32;
    for (int cls = 0; cls < ncls; ++cls)
      for (int mt = 0; mt < p.mt; ++mt) {
        const int mh = h0 + mt * kTH + r, mw = w0 + c;
        const bool valid = r < kTH && c < kTW && mh < pl.hm && mw < pl.wm;
        const int oh = mh * pl.out_s + pl.cls_oh[cls], ow = mw * pl.out_s + pl.cls_ow[cls];
        const size_t pix = (static_cast<size_t>(img) * pl.out_h + oh) * pl.out_w + ow;
#pragma unroll 1
        for (int c0 = 0; c0 < BLOCK_N; c0 += kChunk) {
          uint32_t rr[kChunk];
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                 static_cast<uint32_t>((cls * p.mt + mt) * BLOCK_N + c0);
          if constexpr (kChunk == 32)
            ptx::tmem_ld_32x32b_x32(taddr, rr);
          else
            ptx::tmem_ld_32x32b_x16(taddr, rr);
          ptx::tmem_ld_wait();
          if (n0 + c0 >= pl.cout) break;
          float v[kChunk];
#pragma unroll
          for (int j = 0; j < kChunk; ++j) v[j] = fmaf(__uint_as_float(rr[j]), s_scale[c0 + j], s_shift[c0 + j]);
          if (pl.out_fmt == W2C_OUT_NCHW_F32) {
            if (valid) {
              float* y = static_cast<float*>(pl.y);
              const size_t plane = static_cast<size_t>(pl.out_h) * pl.out_w;
              const size_t base = static_cast<size_t>(img) * pl.cout * plane + static_cast<size_t>(oh) * pl.out_w + ow;
#pragma unroll
              for (int j = 0; j < kChunk; ++j) {
                const int ch = n0 + c0 + j;
                if (ch < pl.cout) y[base + ch * plane] = pl.relu ? fmaxf(v[j], 0.f) : v[j];
              }
            }
          } else if (valid) {
            __nv_bfloat16* ypix = static_cast<__nv_bfloat16*>(pl.y) + pix * pl.y_pix + pl.y_coffset + n0 + c0;
            const __nv_bfloat16* rpix = pl.residual ? pl.residual + pix * pl.y_pix + pl.y_coffset + n0 + c0 : nullptr;
#pragma unroll
            for (int g = 0; g < kChunk / 8; ++g) {
              if (n0 + c0 + g * 8 >= pl.cout) break;
              if (rpix) {
                const uint4 rv = *reinterpret_cast<const uint4*>(rpix + g * 8);
                const __nv_bfloat162* rb = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 f = __bfloat1622float2(rb[j]);
                  v[g * 8 + 2 * j] += f.x, v[g * 8 + 2 * j + 1] += f.y;
                }
                if (pl.act == W2C_ACT_BF16X2) {
                  const uint4 rl = *reinterpret_cast<const uint4*>(rpix + pl.y_cstride + g * 8);
                  const __nv_bfloat162* lb = reinterpret_cast<const __nv_bfloat162*>(&rl);
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float2 f = __bfloat1622float2(lb[j]);
                    v[g * 8 + 2 * j] += f.x, v[g * 8 + 2 * j + 1] += f.y;
                  }
                }
              }
              uint4 hv, lv;
              __nv_bfloat162* hb = reinterpret_cast<__nv_bfloat162*>(&hv);
              __nv_bfloat162* lb = reinterpret_cast<__nv_bfloat162*>(&lv);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float a = v[g * 8 + 2 * j], b = v[g * 8 + 2 * j + 1];
                if (pl.relu) a = fmaxf(a, 0.f), b = fmaxf(b, 0.f);
                hb[j] = __floats2bfloat162_rn(a, b);
                const float2 hf = __bfloat1622float2(hb[j]);
                lb[j] = __floats2bfloat162_rn(a - hf.x, b - hf.y);
              }
              *reinterpret_cast<uint4*>(ypix + g * 8) = hv;
              if (pl.act == W2C_ACT_BF16X2) *reinterpret_cast<uint4*>(ypix + pl.y_cstride + g * 8) = lv;
            }
          }
        }
      }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, tmem_cols);
  }
}

template <int BLOCK_N>
int launch_halo(const HaloParams& p, size_t smem_bytes, cudaStream_t stream) {
  static size_t attr = 0;
  if (smem_bytes > attr) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem_bytes));
    if (e != cudaSuccess) return set_error(W2C_ERR_CUDA, "conv_halo: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr = smem_bytes;
  }
  dim3 grid(p.tiles_w * p.tiles_h * p.plan.n_img * p.n_tiles, 1, 1);
  conv_halo_kernel<BLOCK_N><<<grid, kNumThreads, smem_bytes, stream>>>(p);
  W2C_CHECK_LAUNCH("conv_halo_kernel");
  return W2C_OK;
}

}  // namespace

// Whether the halo kernel applies to this plan (3x3 s1 conv or 3x3 s2 transposed conv).
bool conv_halo_supported(const ConvPlan& plan) {
  if (plan.in_s != 1) return false;
  if (plan.num_classes == 1 && plan.ntaps[0] != 9) return false;
  return true;
}

int conv_halo_forward(const w2c_conv_args& a, const ConvPlan& plan, cudaStream_t stream) {
  HaloParams p;
  p.plan = plan;
  const int planes = plan.act == W2C_ACT_BF16X2 ? 2 : 1;
  const bool deconv = plan.num_classes == 4;
  p.oy = p.ox = deconv ? 0 : 1;

  int bn = a.block_n;
  if (bn == 0) {
    if (plan.cout_pad % 128 == 0 && !deconv)
      bn = 128;
    else if (plan.cout_pad % 64 == 0)
      bn = 64;
    else if (plan.cout_pad % 32 == 0)
      bn = 32;
    else
      bn = 16;
  }
  W2C_CHECK_ARG(bn == 16 || bn == 32 || bn == 64 || bn == 128, "conv_halo: block_n=%d not supported", bn);
  W2C_CHECK_ARG(plan.cout_pad % bn == 0, "conv_halo: block_n=%d does not divide cout_pad=%d", bn, plan.cout_pad);
  p.n_tiles = plan.cout_pad / bn;
  // sub-tiles per CTA: two when the accumulators fit TMEM and the image is tall enough to use them
  // (a transposed conv already keeps four accumulators; one sub-tile leaves TMEM for a second resident CTA)
  p.mt = (!deconv && 2 * bn <= 512 && plan.hm > kTH) ? 2 : 1;
  W2C_CHECK_ARG(plan.num_classes * p.mt * bn <= 512, "conv_halo: accumulators exceed TMEM");
  p.halo_rows = p.mt * kTH + 2;
  // the MMA reads up to 128 rows starting (2*PW + 2) pixels into the last sub-tile's window: pad the stage
  const int a_px = ((p.mt - 1) * kTH + 2) * kPW + 2 + 128;
  p.a_stage_bytes = (a_px * 128 + 1023) / 1024 * 1024;
  p.tiles_w = ceil_div(plan.wm, kTW);
  p.tiles_h = ceil_div(plan.hm, kTH * p.mt);
  const int chunks = plan.cin / kBlockK;
  const int kb_total = chunks * 9 * (planes == 2 ? 3 : 1);
  const int b_stage = bn * 128;
  // small-K layers: keep the footprint low so several CTAs share an SM (their epilogues overlap the others' MMAs)
  p.a_stages = chunks * (planes == 2 ? 3 : 1) >= 2 ? 2 : 1;
  int budget = kb_total <= 18 ? 72 * 1024 : 200 * 1024;
  int bs = (budget - p.a_stages * p.a_stage_bytes) / b_stage;
  if (bs > 9) bs = 9;
  if (bs < 2) bs = 2;
  p.b_stages = bs;
  {
    static const int mode = [] {
      const char* e = getenv("W2C_HALO_BASE_OFFSET");
      return e ? atoi(e) : 0;
    }();
    p.base_offset_mode = mode;
  }

  const cuuint64_t esz = 2;
  {
    const cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)kPW, (cuuint32_t)p.halo_rows, 1};
    const cuuint64_t dims[4] = {(cuuint64_t)plan.x_pix, (cuuint64_t)plan.in_w, (cuuint64_t)plan.in_h,
                                (cuuint64_t)plan.n_img};
    const cuuint64_t str[3] = {plan.x_pix * esz, (cuuint64_t)plan.in_w * plan.x_pix * esz,
                               (cuuint64_t)plan.in_h * plan.in_w * plan.x_pix * esz};
    int rc = encode_map(&p.a_map, plan.x, 4, dims, str, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (rc) return rc;
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)plan.ktot, (cuuint64_t)plan.cout_pad * planes};
    const cuuint64_t str[1] = {plan.ktot * esz};
    const cuuint32_t bbox[2] = {(cuuint32_t)kBlockK, (cuuint32_t)bn};
    int rc = encode_map(&p.b_map, plan.w, 2, dims, str, bbox, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (rc) return rc;
  }
  const size_t smem = static_cast<size_t>(p.a_stages) * p.a_stage_bytes + static_cast<size_t>(p.b_stages) * b_stage +
                      (2 * p.a_stages + 2 * p.b_stages + 1) * 8 + 16 + 2 * bn * 4 + 1024;
  switch (bn) {
    case 128: return launch_halo<128>(p, smem, stream);
    case 64: return launch_halo<64>(p, smem, stream);
    case 32: return launch_halo<32>(p, smem, stream);
    default: return launch_halo<16>(p, smem, stream);
  }
}

}  // namespace w2c
