// First encoder layer on the tensor cores: Conv2d(3 -> COUT, k3 s1 p1) + folded BN + ReLU, reading the caller's
// fp32 NCHW views directly and writing bf16 NHWC.
//
// K = 27 is far too small for a TMA-fed implicit GEMM, and on the CUDA cores the layer is issue-bound (~2.7 ms for
// 40 agent-frames against a 0.22 ms HBM floor). Here the 128 threads of a CTA each gather the 27 taps of one output
// pixel (coalesced 4-byte reads of three channel planes), convert to bf16 and write one 64-byte K-major row of the
// A operand straight into shared memory in the 128B-swizzle layout the UMMA descriptor expects (software im2col).
// One thread then issues two tcgen05.mma (M=128, N=COUT, K=16) — six in the bf16x3 precision — and the same 128
// threads read the accumulators back from TMEM, write [128 px][64 ch] swizzled staging tiles and one thread issues
// a TMA tensor store per tile (the direct version issued 16-byte stores at a 256-byte lane stride and ran at a
// third of the HBM rate). A CTA loops over pixel tiles with the next tile's taps prefetched into registers during
// the epilogue; four CTAs per SM overlap gather, MMA and store phases.
// COUT = 128 is two encoders' first layers fused (the image is read once).
//
// The folded BatchNorm rides in the GEMM: the weight rows are pre-multiplied by scale[co] (in fp32, before the
// bf16 split) and shift[co] is the weight of a constant-1 input in the K padding (k = 27), so the accumulator IS
// the pre-activation and the epilogue is one cvt.rn.relu.bf16x2 per channel pair. (ncu on the first version: 411 M
// warp instructions per step - 2 LDS + FMA + MAX + CVT per output - 1.24 ms against a 0.42 ms HBM floor, and 187
// registers per thread, which silently capped residency at two CTAs per SM for a grid sized for three.)
#include "common.cuh"
#include "ptx.cuh"

namespace w2c {

int encode_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, CUtensorMapL2promotion promo);

namespace {

constexpr int kTile = 128;      // pixels per tile = UMMA M
constexpr int kRowBytes = 128;  // one swizzle row; only the first 64 B (K = 32 bf16) of an operand row are used
constexpr int kStagingBytes = kTile * 64 * 2;
constexpr int kStemNoRelu = 0x100;  // flag OR-ed into the kernels' `act` argument: no ReLU in the epilogue

template <int COUT, int kNumStaging>
struct StemSmem {
  static constexpr int kA = kTile * kRowBytes;  // 16 KB per plane
  static constexpr int kB = COUT * kRowBytes;
  // the lo-plane operand tiles (bf16x3 only) sit at the end, so the bf16 launch requests less shared memory
  static constexpr int kAHi = 0, kBHi = kA;
  static constexpr int kStg = kA + kB;
  static constexpr int kBar = kStg + kNumStaging * kStagingBytes;
  static constexpr int kTmemPtr = kBar + 8;
  static constexpr int kLut = kTmemPtr + 8;  // [3][256] fp32 loader-transform table (uint8 input only)
  static constexpr int kLoBase = (kLut + 3 * 256 * 4 + 1023) / 1024 * 1024;
  static constexpr int kALo = kLoBase, kBLo = kLoBase + kA;
  static constexpr int kDynamicBf16 = kLoBase + 1024;
  static constexpr int kDynamic = kLoBase + kA + kB + 1024;
};

__device__ __forceinline__ size_t total_px(int b_sz, int n_agents, int h, int wpx) {
  return static_cast<size_t>(b_sz) * n_agents * h * wpx;
}

// byte offset of 16-byte chunk `c` of row `r` in a 128B-swizzled K-major tile
__device__ __forceinline__ uint32_t sw128(uint32_t r, uint32_t c) { return r * kRowBytes + ((c ^ (r & 7u)) << 4); }

__device__ __forceinline__ void gather_taps(const float* __restrict__ x, size_t p, size_t total, int b_sz, int c_total,
                                            int c_first, int h, int wpx, float (&in)[27]) {
  if (p >= total) {
#pragma unroll
    for (int k = 0; k < 27; ++k) in[k] = 0.f;
    return;
  }
  const size_t plane = static_cast<size_t>(h) * wpx;
  const int ow = p % wpx;
  const int oh = (p / wpx) % h;
  const int img = p / plane;  // agent-major: img = agent * b_sz + batch
  const int agent = img / b_sz, bat = img % b_sz;
  const float* xin = x + (static_cast<size_t>(bat) * c_total + c_first + 3 * agent) * plane;
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
}

// The same taps from the loader's raw frames: uint8 RGB, HWC, [b][agents_total][h][w][3]. Conv input channel ci is
// BGR channel ci = RGB byte 2-ci (the loader's img[:, :, ::-1], airsim_loader.py:521), mapped through the
// (v - mean[ci]) / 255 table (airsim_loader.py:522-525; built on the host in float64 like the loader does).
__device__ __forceinline__ void gather_taps_u8(const uint8_t* __restrict__ x, const float* __restrict__ lut, size_t p,
                                               size_t total, int b_sz, int agents_total, int agent_first, int h,
                                               int wpx, float (&in)[27]) {
  if (p >= total) {
#pragma unroll
    for (int k = 0; k < 27; ++k) in[k] = 0.f;
    return;
  }
  const size_t plane = static_cast<size_t>(h) * wpx;
  const int ow = p % wpx;
  const int oh = (p / wpx) % h;
  const int img = p / plane;
  const int agent = img / b_sz, bat = img % b_sz;
  const uint8_t* xin = x + (static_cast<size_t>(bat) * agents_total + agent_first + agent) * plane * 3;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int ih = oh + kh - 1, iw = ow + kw - 1;
      const bool ok = ih >= 0 && ih < h && iw >= 0 && iw < wpx;
      const uint8_t* px = xin + (static_cast<size_t>(ih) * wpx + iw) * 3;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) in[ci * 9 + kh * 3 + kw] = ok ? lut[ci * 256 + __ldg(px + 2 - ci)] : 0.f;
    }
}

// Output tensor maps: one dense [pixels][cs * planes] tensor per split (cs = COUT / n_split channels each). The
// fused two-encoder stem writes TWO dense 64-channel maps: interleaved in one 128-channel map, each encoder's
// stride-2 conv read 128 of every 256 bytes (0.56 ms in the net against 0.43 ms on a dense map).
struct StemMaps {
  CUtensorMap m[2];
  int cs;  // channels per split
  // direct (non-TMA) store path: base of each split's map, bytes per pixel row, pixels in total
  uint8_t* base[2];
  int pitch;
  int direct;
  size_t total;
};

template <int COUT, bool U8, int NS>
__global__ void __launch_bounds__(kTile, 4) stem3x3_tc_kernel(const __grid_constant__ StemMaps y_maps,
                                                              const void* __restrict__ x_any,
                                                              const float* __restrict__ lut_g,
                                                              const float* __restrict__ w,
                                                              const float* __restrict__ scale,
                                                              const float* __restrict__ shift, int b_sz, int n_agents,
                                                              int c_total, int c_first, int h, int wpx, int act,
                                                              int num_tiles) {
  using L = StemSmem<COUT, NS>;
  constexpr int kTmemCols = COUT < 32 ? 32 : COUT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::kBar);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kTmemPtr);
  float* s_lut = reinterpret_cast<float*>(smem + L::kLut);
  const float* x = static_cast<const float*>(x_any);
  const uint8_t* xu8 = static_cast<const uint8_t*>(x_any);
  const size_t total = total_px(b_sz, n_agents, h, wpx);
  // c_total / c_first carry agents_total / agent_first for the uint8 input
  auto gather = [&](size_t p, float (&dst)[27]) {
    if constexpr (U8)
      gather_taps_u8(xu8, s_lut, p, total, b_sz, c_total, c_first, h, wpx, dst);
    else
      gather_taps(x, p, total, b_sz, c_total, c_first, h, wpx, dst);
  };
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool relu = (act & kStemNoRelu) == 0;   // (train-mode BatchNorm wants the raw conv output, csrc/bn_train.cu)
  act &= 0xff;
  const bool x3 = act_planes(act) == 2;
  const bool f16 = act_is_f16(act);

  // ---- one-time setup: scale * weights (k < 27) and shift (k = 27) -> swizzled B tiles, barrier, TMEM
  for (int i = tid; i < COUT * 4; i += kTile) {
    const int co = i >> 2, c = i & 3;  // 16-byte chunk c = k in [8c, 8c+8)
    const float sc = scale[co];
    uint4 hv, lv;
    __nv_bfloat16* hb = reinterpret_cast<__nv_bfloat16*>(&hv);
    __nv_bfloat16* lb = reinterpret_cast<__nv_bfloat16*>(&lv);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = c * 8 + e;
      const float v = k < 27 ? w[co * 27 + k] * sc : (k == 27 ? shift[co] : 0.f);
      hb[e] = float_to_elem(v, f16);
      lb[e] = float_to_elem(v - elem_to_float(hb[e], f16), f16);  // (lo plane: two-plane formats only)
    }
    *reinterpret_cast<uint4*>(smem + L::kBHi + sw128(co, c)) = hv;
    if (x3) *reinterpret_cast<uint4*>(smem + L::kBLo + sw128(co, c)) = lv;
  }
  if constexpr (U8) {
    for (int i = tid; i < 3 * 256; i += kTile) s_lut[i] = lut_g[i];
    __syncthreads();  // the first gather below reads the table
  }
  if (tid == 0) {
    ptx::prefetch_tensormap(&y_maps.m[0]);
    ptx::mbar_init(bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(tmem_ptr, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t idesc = ptx::make_idesc_16(kTile, COUT, f16);
  const uint64_t a_hi = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + L::kAHi));
  const uint64_t a_lo = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + L::kALo));
  const uint64_t b_hi = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + L::kBHi));
  const uint64_t b_lo = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + L::kBLo));

  float in[27];
  int tile = blockIdx.x;
  if (tile < num_tiles) gather(static_cast<size_t>(tile) * kTile + tid, in);
  uint32_t phase = 0;
  int unit = 0;  // staging-buffer rotation
  const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  for (; tile < num_tiles; tile += gridDim.x) {
    // ---- software im2col: this thread's pixel -> row `tid` of the A tile(s); k = 27 is the constant 1 that
    //      carries the folded shift, k = 28..31 are zero
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint4 hv, lv;
      __nv_bfloat16* hb = reinterpret_cast<__nv_bfloat16*>(&hv);
      __nv_bfloat16* lb = reinterpret_cast<__nv_bfloat16*>(&lv);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = c * 8 + e;
        const float v = k < 27 ? in[k < 27 ? k : 0] : (k == 27 ? 1.f : 0.f);
        hb[e] = float_to_elem(v, f16);
        lb[e] = float_to_elem(v - elem_to_float(hb[e], f16), f16);  // (lo plane: two-plane formats only)
      }
      *reinterpret_cast<uint4*>(smem + L::kAHi + sw128(tid, c)) = hv;
      if (x3) *reinterpret_cast<uint4*>(smem + L::kALo + sw128(tid, c)) = lv;
    }
    ptx::fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0 && ptx::elect_one_sync()) {
      ptx::tc_fence_after();
      // K = 32 = two UMMA k-steps (32 B apart inside the swizzle row)
      ptx::umma_bf16(tmem_base, a_hi, b_hi, idesc, 0);
      ptx::umma_bf16(tmem_base, a_hi + 2, b_hi + 2, idesc, 1);
      if (x3) {
        ptx::umma_bf16(tmem_base, a_hi, b_lo, idesc, 1);
        ptx::umma_bf16(tmem_base, a_hi + 2, b_lo + 2, idesc, 1);
        ptx::umma_bf16(tmem_base, a_lo, b_hi, idesc, 1);
        ptx::umma_bf16(tmem_base, a_lo + 2, b_hi + 2, idesc, 1);
      }
      ptx::umma_commit(bar);
    }
    // prefetch the next tile's taps while the MMA runs and before the store phase
    const int next = tile + gridDim.x;
    if (next < num_tiles) gather(static_cast<size_t>(next) * kTile + tid, in);

    ptx::mbar_wait(bar, phase);
    phase ^= 1;
    ptx::tc_fence_after();
    const int planes = x3 ? 2 : 1;
#pragma unroll 1
    for (int g = 0; g < COUT / 64; ++g) {
#pragma unroll 1
      for (int pln = 0; pln < planes; ++pln, ++unit) {
        uint8_t* stg = smem + L::kStg + (unit % NS) * kStagingBytes;
        // the store that last used this staging tile has read it (TMA: bulk group; direct: every thread's loads)
        if (!y_maps.direct && warp == 0 && ptx::elect_one_sync()) ptx::bulk_wait_group_read<NS - 1>();
        __syncthreads();
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {  // 32 accumulator columns at a time (register budget: 128)
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(t_row + g * 64 + half * 32, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            uint4 pk;
            if (!x3) {
              pk.x = ptx::pack_act2(__uint_as_float(r[c4 * 8 + 0]), __uint_as_float(r[c4 * 8 + 1]), relu, f16);
              pk.y = ptx::pack_act2(__uint_as_float(r[c4 * 8 + 2]), __uint_as_float(r[c4 * 8 + 3]), relu, f16);
              pk.z = ptx::pack_act2(__uint_as_float(r[c4 * 8 + 4]), __uint_as_float(r[c4 * 8 + 5]), relu, f16);
              pk.w = ptx::pack_act2(__uint_as_float(r[c4 * 8 + 6]), __uint_as_float(r[c4 * 8 + 7]), relu, f16);
            } else {
              uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float a = __uint_as_float(r[c4 * 8 + 2 * j]), b = __uint_as_float(r[c4 * 8 + 2 * j + 1]);
                if (relu) a = fmaxf(a, 0.f), b = fmaxf(b, 0.f);
                uint32_t hi, lo;
                split_act2(a, b, f16, hi, lo);
                pw[j] = pln == 0 ? hi : lo;
              }
            }
            *reinterpret_cast<uint4*>(stg + sw128(tid, half * 4 + c4)) = pk;
          }
        }
        const int split = (g * 64) / y_maps.cs;
        const int col = pln * y_maps.cs + g * 64 - split * y_maps.cs;  // first channel of this group in its map
        if (y_maps.direct) {
          // Experiment: coalesced 16-byte stores straight from the staging tile (eight consecutive threads write one
          // pixel's 128 bytes) in place of the TMA store, to test whether the TMA store engine binds this kernel. It
          // does not: measured 1 % slower (3451 vs 3496 agent-frames/s in one run).
          __syncthreads();
          uint8_t* gdst = y_maps.base[split] + static_cast<size_t>(tile) * kTile * y_maps.pitch + col * 2;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int q = i * kTile + tid, row = q >> 3, c8 = q & 7;
            if (static_cast<size_t>(tile) * kTile + row < y_maps.total)
              *reinterpret_cast<uint4*>(gdst + static_cast<size_t>(row) * y_maps.pitch + c8 * 16) =
                  *reinterpret_cast<const uint4*>(stg + sw128(row, c8));
          }
        } else {
          ptx::fence_proxy_async();
          __syncthreads();
          if (warp == 0 && ptx::elect_one_sync()) {
            // rows of the output viewed as [total pixels][COUT * planes]; rows past the end are clipped by the TMA unit
            ptx::tma_store_2d(&y_maps.m[split], stg, col, tile * kTile);
            ptx::bulk_commit_group();
          }
        }
      }
    }
    // the next iteration's __syncthreads (after its smem writes) orders these TMEM reads before the next MMA
    ptx::tc_fence_before();
  }
  __syncwarp();
  if (warp == 0 && ptx::elect_one_sync()) ptx::bulk_wait_group<0>();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int COUT, bool U8, int NS>
int launch_stem_ns(const void* x, const float* lut, const float* w, const float* scale, const float* shift, void* y,
                   int b, int n_agents, int c_total, int c_first, int h, int wpx, int act, int n_split,
                   cudaStream_t stream) {
  using L = StemSmem<COUT, NS>;
  static DeviceOnce attr;  // per device, not per process (common.cuh)
  if (int rc = attr.ensure([] {
        return cudaFuncSetAttribute(stem3x3_tc_kernel<COUT, U8, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    L::kDynamic);
      }, "stem3x3_tc_kernel"))
    return rc;
  const size_t total = static_cast<size_t>(b) * n_agents * h * wpx;
  const int num_tiles = static_cast<int>((total + kTile - 1) / kTile);
  const int planes = act_planes(act & 0xff);
  StemMaps y_maps;
  y_maps.cs = COUT / n_split;
  y_maps.direct = 0;
  y_maps.pitch = y_maps.cs * planes * 2;
  y_maps.total = total;
  for (int sp = 0; sp < 2; ++sp) {
    const cuuint64_t dims[2] = {(cuuint64_t)y_maps.cs * planes, (cuuint64_t)total};
    const cuuint64_t str[1] = {(cuuint64_t)y_maps.cs * planes * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)kTile};
    const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(y) + (sp < n_split ? sp : 0) * total * y_maps.cs * planes;
    int rc = encode_map(&y_maps.m[sp], base, 2, dims, str, box, CU_TENSOR_MAP_L2_PROMOTION_NONE);
    if (rc) return rc;
    y_maps.base[sp] = reinterpret_cast<uint8_t*>(const_cast<__nv_bfloat16*>(base));
  }
  // resident CTAs per SM: bounded by shared memory (operand tiles + NS staging tiles; the bf16x3 precision adds
  // the lo-plane operand tiles), by TMEM (COUT columns of 512) and by registers (<= 128 per thread: four CTAs)
  const int smem_bytes = planes == 2 ? L::kDynamic : L::kDynamicBf16;
  int per_sm = (227 * 1024) / (smem_bytes + 1024);
  if (per_sm > 512 / (COUT < 32 ? 32 : COUT)) per_sm = 512 / (COUT < 32 ? 32 : COUT);
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  int grid = device_sm_count() * per_sm;
  if (grid > num_tiles) grid = num_tiles;
  stem3x3_tc_kernel<COUT, U8, NS><<<grid, kTile, smem_bytes, stream>>>(y_maps, x, lut, w, scale, shift, b, n_agents,
                                                                        c_total, c_first, h, wpx, act, num_tiles);
  W2C_CHECK_LAUNCH("stem3x3_tc_kernel");
  return W2C_OK;
}

template <int COUT, bool U8>
int launch_stem(const void* x, const float* lut, const float* w, const float* scale, const float* shift, void* y, int b,
                int n_agents, int c_total, int c_first, int h, int wpx, int act, int n_split, cudaStream_t stream) {
  // one staging tile -> four CTAs per SM at COUT = 128 (52 KB each; two tiles / three CTAs measured slower)
  return launch_stem_ns<COUT, U8, 1>(x, lut, w, scale, shift, y, b, n_agents, c_total, c_first, h, wpx, act, n_split,
                                     stream);
}


// ================================================================================================ 7x7 stride-2 stem
// resnet18's first layer (Conv2d(3 -> 64, k7 s2 p3, no bias) + BN + ReLU, backbone.py:63-75) on the tensor cores with
// the same software im2col: K = 147 taps + the constant-1 shift column = 148, padded to 160 = ten K16 steps, laid
// out as three 64-wide K chunks (three 128B-swizzled [128 px][64] A tiles, three [COUT][64] B tiles). A thread
// gathers its output pixel's taps chunk by chunk (64 fp32 in registers at a time). The CUDA-core version took 2.94 ms
// per 40 frames (17 TFLOP/s) - 63 % of the whole resnet-backbone step; COUT = 128 fuses two encoders' first layers.
template <int COUT>
struct Stem7Smem {
  static constexpr int kA = 3 * kTile * kRowBytes;  // 48 KB per plane
  static constexpr int kB = 3 * COUT * kRowBytes;
  static constexpr int kAHi = 0, kBHi = kA;
  static constexpr int kStg = kA + kB;
  static constexpr int kBar = kStg + kStagingBytes;
  static constexpr int kTmemPtr = kBar + 8;
  static constexpr int kLut = kTmemPtr + 8;
  static constexpr int kLoBase = (kLut + 3 * 256 * 4 + 1023) / 1024 * 1024;
  static constexpr int kALo = kLoBase, kBLo = kLoBase + kA;
  static constexpr int kDynamicBf16 = kLoBase + 1024;
  static constexpr int kDynamic = kLoBase + kA + kB + 1024;
};

template <int COUT, bool U8>
__global__ void __launch_bounds__(kTile, 2) stem7x7_tc_kernel(const __grid_constant__ StemMaps y_maps,
                                                              const void* __restrict__ x_any,
                                                              const float* __restrict__ lut_g,
                                                              const float* __restrict__ w,
                                                              const float* __restrict__ scale,
                                                              const float* __restrict__ shift, int b_sz, int n_agents,
                                                              int c_total, int c_first, int h, int wpx, int act,
                                                              int num_tiles) {
  using L = Stem7Smem<COUT>;
  constexpr int KT = 147;  // taps; k = 147 carries the folded shift
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::kBar);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kTmemPtr);
  float* s_lut = reinterpret_cast<float*>(smem + L::kLut);
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool relu = (act & kStemNoRelu) == 0;   // (train-mode BatchNorm wants the raw conv output, csrc/bn_train.cu)
  act &= 0xff;
  const bool x3 = act_planes(act) == 2;
  const bool f16 = act_is_f16(act);
  const int ho = h / 2, wo = wpx / 2;
  const size_t total = static_cast<size_t>(b_sz) * n_agents * ho * wo;
  const size_t plane = static_cast<size_t>(h) * wpx;

  // ---- one-time setup: scale * weights and shift -> the three swizzled B chunk tiles, barrier, TMEM
  for (int i = tid; i < COUT * 24; i += kTile) {
    const int co = i / 24, piece = i % 24;  // 16-byte piece = k in [8*piece, 8*piece + 8)
    const int ck = piece >> 3, c = piece & 7;
    const float sc = scale[co];
    uint4 hv, lv;
    __nv_bfloat16* hb = reinterpret_cast<__nv_bfloat16*>(&hv);
    __nv_bfloat16* lb = reinterpret_cast<__nv_bfloat16*>(&lv);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = piece * 8 + e;
      const float v = k < KT ? w[co * KT + k] * sc : (k == KT ? shift[co] : 0.f);
      hb[e] = float_to_elem(v, f16);
      lb[e] = float_to_elem(v - elem_to_float(hb[e], f16), f16);  // (lo plane: two-plane formats only)
    }
    *reinterpret_cast<uint4*>(smem + L::kBHi + ck * COUT * kRowBytes + sw128(co, c)) = hv;
    if (x3) *reinterpret_cast<uint4*>(smem + L::kBLo + ck * COUT * kRowBytes + sw128(co, c)) = lv;
  }
  if constexpr (U8)
    for (int i = tid; i < 3 * 256; i += kTile) s_lut[i] = lut_g[i];
  if (tid == 0) {
    ptx::prefetch_tensormap(&y_maps.m[0]);
    ptx::mbar_init(bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(tmem_ptr, COUT);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t idesc = ptx::make_idesc_16(kTile, COUT, f16);
  const uint64_t a_hi = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + L::kAHi));
  const uint64_t a_lo = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + L::kALo));
  const uint64_t b_hi = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + L::kBHi));
  const uint64_t b_lo = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + L::kBLo));
  constexpr uint32_t kAChunk16 = (kTile * kRowBytes) >> 4, kBChunk16 = (COUT * kRowBytes) >> 4;

  uint32_t phase = 0;
  const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  const int planes = x3 ? 2 : 1;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    // ---- this thread's output pixel and the validity of its 7 input rows / columns
    const size_t p = static_cast<size_t>(tile) * kTile + tid;
    const bool live = p < total;
    const int ow = live ? static_cast<int>(p % wo) : 0;
    const int oh = live ? static_cast<int>((p / wo) % ho) : 0;
    const int img = live ? static_cast<int>(p / (static_cast<size_t>(ho) * wo)) : 0;
    const int agent = img / b_sz, bat = img % b_sz;
    const int ih0 = 2 * oh - 3, iw0 = 2 * ow - 3;
    uint32_t rmask = 0, cmask = 0;
#pragma unroll
    for (int t = 0; t < 7; ++t) {
      rmask |= (live && ih0 + t >= 0 && ih0 + t < h) ? (1u << t) : 0u;
      cmask |= (iw0 + t >= 0 && iw0 + t < wpx) ? (1u << t) : 0u;
    }
    const float* xf = nullptr;
    const uint8_t* xb = nullptr;
    if constexpr (U8)  // c_total / c_first carry agents_total / agent_first
      xb = static_cast<const uint8_t*>(x_any) + (static_cast<size_t>(bat) * c_total + c_first + agent) * plane * 3;
    else
      xf = static_cast<const float*>(x_any) + (static_cast<size_t>(bat) * c_total + c_first + 3 * agent) * plane;
    const ptrdiff_t pix0 = static_cast<ptrdiff_t>(ih0) * wpx + iw0;

    // ---- software im2col, one 64-wide K chunk at a time
#pragma unroll
    for (int ck = 0; ck < 3; ++ck) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 hv, lv;
        __nv_bfloat16* hb = reinterpret_cast<__nv_bfloat16*>(&hv);
        __nv_bfloat16* lb = reinterpret_cast<__nv_bfloat16*>(&lv);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          constexpr int dummy = 0;
          (void)dummy;
          const int k = ck * 64 + c * 8 + e;  // compile-time after unrolling
          float v = 0.f;
          if (k < KT) {
            const int ci = k / 49, r = k % 49, kh = r / 7, kw = r % 7;
            if ((rmask >> kh) & (cmask >> kw) & 1u) {
              if constexpr (U8)
                v = s_lut[ci * 256 + __ldg(xb + (pix0 + kh * wpx + kw) * 3 + (2 - ci))];
              else
                v = __ldg(xf + ci * plane + pix0 + kh * wpx + kw);
            }
          } else if (k == KT) {
            v = 1.f;
          }
          hb[e] = float_to_elem(v, f16);
          lb[e] = float_to_elem(v - elem_to_float(hb[e], f16), f16);  // (lo plane: two-plane formats only)
        }
        if (ck * 64 + c * 8 < 160) {  // K is padded to 160: the last 32 columns of chunk 2 are never read
          *reinterpret_cast<uint4*>(smem + L::kAHi + ck * kTile * kRowBytes + sw128(tid, c)) = hv;
          if (x3) *reinterpret_cast<uint4*>(smem + L::kALo + ck * kTile * kRowBytes + sw128(tid, c)) = lv;
        }
      }
    }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0 && ptx::elect_one_sync()) {
      ptx::tc_fence_after();
#pragma unroll
      for (int s = 0; s < 10; ++s) {  // ten K16 steps: chunk s / 4, 32 B per step inside the swizzle row
        const uint32_t ao = (s >> 2) * kAChunk16 + 2 * (s & 3), bo = (s >> 2) * kBChunk16 + 2 * (s & 3);
        ptx::umma_bf16(tmem_base, a_hi + ao, b_hi + bo, idesc, s);
        if (x3) {
          ptx::umma_bf16(tmem_base, a_hi + ao, b_lo + bo, idesc, 1);
          ptx::umma_bf16(tmem_base, a_lo + ao, b_hi + bo, idesc, 1);
        }
      }
      ptx::umma_commit(bar);
    }
    ptx::mbar_wait(bar, phase);
    phase ^= 1;
    ptx::tc_fence_after();
    uint8_t* stg = smem + L::kStg;
#pragma unroll 1
    for (int g = 0; g < COUT / 64; ++g) {
#pragma unroll 1
      for (int pln = 0; pln < planes; ++pln) {
        if (warp == 0 && ptx::elect_one_sync()) ptx::bulk_wait_group_read<0>();  // last store has read the tile
        __syncthreads();
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(t_row + g * 64 + half * 32, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            uint4 pk;
            uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float a = __uint_as_float(r[c4 * 8 + 2 * j]), b = __uint_as_float(r[c4 * 8 + 2 * j + 1]);
              if (!x3) {
                pw[j] = ptx::pack_act2(a, b, relu, f16);
              } else {
                const float ar = relu ? fmaxf(a, 0.f) : a, br = relu ? fmaxf(b, 0.f) : b;
                uint32_t hi, lo;
                split_act2(ar, br, f16, hi, lo);
                pw[j] = pln == 0 ? hi : lo;
              }
            }
            *reinterpret_cast<uint4*>(stg + sw128(tid, half * 4 + c4)) = pk;
          }
        }
        ptx::fence_proxy_async();
        __syncthreads();
        if (warp == 0 && ptx::elect_one_sync()) {
          const int split = (g * 64) / y_maps.cs;
          ptx::tma_store_2d(&y_maps.m[split], stg, pln * y_maps.cs + g * 64 - split * y_maps.cs, tile * kTile);
          ptx::bulk_commit_group();
        }
      }
    }
    ptx::tc_fence_before();  // (the next tile's __syncthreads orders these TMEM reads before its MMAs)
  }
  __syncwarp();
  if (warp == 0 && ptx::elect_one_sync()) ptx::bulk_wait_group<0>();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, COUT);
  }
}

template <int COUT, bool U8>
int launch_stem7(const void* x, const float* lut, const float* w, const float* scale, const float* shift, void* y, int b,
                 int n_agents, int c_total, int c_first, int h, int wpx, int act, int n_split, cudaStream_t stream) {
  using L = Stem7Smem<COUT>;
  static DeviceOnce attr;
  if (int rc = attr.ensure([] {
        return cudaFuncSetAttribute(stem7x7_tc_kernel<COUT, U8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    L::kDynamic);
      }, "stem7x7_tc_kernel"))
    return rc;
  const size_t total = static_cast<size_t>(b) * n_agents * (h / 2) * (wpx / 2);
  const int num_tiles = static_cast<int>((total + kTile - 1) / kTile);
  const int planes = act_planes(act & 0xff);
  StemMaps y_maps;
  y_maps.cs = COUT / n_split;
  y_maps.direct = 0;
  y_maps.pitch = y_maps.cs * planes * 2;
  y_maps.total = total;
  for (int sp = 0; sp < 2; ++sp) {
    const cuuint64_t dims[2] = {(cuuint64_t)y_maps.cs * planes, (cuuint64_t)total};
    const cuuint64_t str[1] = {(cuuint64_t)y_maps.cs * planes * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)kTile};
    const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(y) + (sp < n_split ? sp : 0) * total * y_maps.cs * planes;
    int rc = encode_map(&y_maps.m[sp], base, 2, dims, str, box, CU_TENSOR_MAP_L2_PROMOTION_NONE);
    if (rc) return rc;
    y_maps.base[sp] = reinterpret_cast<uint8_t*>(const_cast<__nv_bfloat16*>(base));
  }
  const int smem_bytes = planes == 2 ? L::kDynamic : L::kDynamicBf16;
  int per_sm = (227 * 1024) / (smem_bytes + 1024);
  if (per_sm > 2) per_sm = 2;
  if (per_sm < 1) per_sm = 1;
  int grid = device_sm_count() * per_sm;
  if (grid > num_tiles) grid = num_tiles;
  stem7x7_tc_kernel<COUT, U8><<<grid, kTile, smem_bytes, stream>>>(y_maps, x, lut, w, scale, shift, b, n_agents, c_total,
                                                                  c_first, h, wpx, act, num_tiles);
  W2C_CHECK_LAUNCH("stem7x7_tc_kernel");
  return W2C_OK;
}

}  // namespace

int stem3x3_tc_forward(const float* x, const float* w, const float* scale, const float* shift, void* y, int b,
                       int n_agents, int c_total, int c_first, int h, int wpx, int cout, int act, int n_split,
                       cudaStream_t stream) {
  if (cout == 64 && n_split == 1)
    return launch_stem<64, false>(x, nullptr, w, scale, shift, y, b, n_agents, c_total, c_first, h, wpx, act, 1, stream);
  if (cout == 128 && (n_split == 1 || n_split == 2))
    return launch_stem<128, false>(x, nullptr, w, scale, shift, y, b, n_agents, c_total, c_first, h, wpx, act, n_split,
                                   stream);
  return set_error(W2C_ERR_UNSUPPORTED, "stem3x3_tc: cout=%d n_split=%d (64, or 128 as one or two maps)", cout, n_split);
}

int stem3x3_tc_u8_forward(const uint8_t* x, const float* lut, const float* w, const float* scale, const float* shift,
                          void* y, int b, int n_agents, int agents_total, int agent_first, int h, int wpx, int cout,
                          int act, int n_split, cudaStream_t stream) {
  if (cout == 64 && n_split == 1)
    return launch_stem<64, true>(x, lut, w, scale, shift, y, b, n_agents, agents_total, agent_first, h, wpx, act, 1,
                                 stream);
  if (cout == 128 && (n_split == 1 || n_split == 2))
    return launch_stem<128, true>(x, lut, w, scale, shift, y, b, n_agents, agents_total, agent_first, h, wpx, act,
                                  n_split, stream);
  return set_error(W2C_ERR_UNSUPPORTED, "stem3x3_tc (uint8 frames): cout=%d n_split=%d", cout, n_split);
}

int stem7x7_tc_forward(const void* x, const float* lut, const float* w, const float* scale, const float* shift, void* y,
                       int b, int n_agents, int c_total, int c_first, int h, int wpx, int cout, int act, int n_split,
                       int u8, cudaStream_t stream) {
  if (!((cout == 64 && n_split == 1) || (cout == 128 && (n_split == 1 || n_split == 2))))
    return set_error(W2C_ERR_UNSUPPORTED, "stem7x7_tc: cout=%d n_split=%d (64, or 128 as one or two maps)", cout, n_split);
  if (cout == 64)
    return u8 ? launch_stem7<64, true>(x, lut, w, scale, shift, y, b, n_agents, c_total, c_first, h, wpx, act, 1, stream)
              : launch_stem7<64, false>(x, lut, w, scale, shift, y, b, n_agents, c_total, c_first, h, wpx, act, 1, stream);
  return u8 ? launch_stem7<128, true>(x, lut, w, scale, shift, y, b, n_agents, c_total, c_first, h, wpx, act, n_split,
                                      stream)
            : launch_stem7<128, false>(x, lut, w, scale, shift, y, b, n_agents, c_total, c_first, h, wpx, act, n_split,
                                       stream);
}

}  // namespace w2c
