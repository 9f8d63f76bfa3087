// Fused head of n_segnet_encoder: conv1 (3 -> 64, 3x3 s1, BN, ReLU) + conv2 (64 -> 64, 3x3 s2, BN, ReLU), the
// divide_inputs / cat regrouping (and optionally the loader transform on raw uint8 frames) in ONE persistent
// tcgen05 kernel. Replaces backbone.py:19-20,42-43 (conv1, conv2 of n_segnet_encoder.forward) as called from
// img_encoder.forward (agent.py:56-60) on the views of agent.py:1088-1108.
//
// Why: the two layers are the HBM-bound end of the stack (26 and 115 FLOP/B against a ridge of
// 210). Unfused, conv1 writes a 64-channel full-resolution map (2.62 GB per 40 frames for the two encoders) that conv2
// reads back nine taps at a time through L2 (ncu: 4.4 GB of L2 -> SM traffic per launch, weights re-fetched per tile).
// Here a CTA owns an 8x16 tile of conv2 OUTPUT pixels per step and
//   1. stages the 19x35x3 image patch the tile needs (fp32, or uint8 frames through the loader-transform table),
//   2. builds conv1's im2col rows (K = 27 taps + the constant-1 column that carries the folded BatchNorm shift) for the
//      17x33 conv1 outputs under the tile, five M = 128 tiles, and runs them on the tensor core into five TMEM
//      accumulators,
//   3. drains those accumulators (ReLU, 16-bit) straight into SIX shared-memory planes laid out as conv2's A operand:
//      (row parity even / odd) x (column variant even / odd-left / odd-right), 16 pixels x 128 B per row, 128B-swizzled.
//      Every stride-2 tap of conv2 is then a plain K-major [128 px][64 ch] tile at a 1024-byte-aligned address (row
//      shifts are 2048 B; the one-pixel column shift of the odd columns is why that plane exists twice),
//   4. runs conv2's 36 MMAs (9 taps x 4 k-steps, weights resident in shared memory for the whole kernel) into a
//      double-buffered accumulator, and
//   5. drains it through the usual scale/shift + ReLU epilogue, a swizzled staging tile and one TMA tensor store.
// The 64-channel full-resolution map never exists in HBM; per 40 frames the kernel reads 126 MB (31 MB as uint8) and
// writes 336 MB. Accumulation order and operand rounding are those of the unfused kernels (stem_tc.cu followed by
// conv_pers_v1.cu): results are bit-identical to that path, which is how the tests pin it.
//
// warps 0-3: patch + im2col ("front")    warp 4: MMA issuer, TMEM owner    warps 5-8 / 9-12: two epilogue warpgroups
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace w2c {

int encode_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, CUtensorMapL2promotion promo);

namespace {

constexpr int kM = 128;            // UMMA M
constexpr int kRow = 128;          // bytes per operand row: 64 x 16 bit, one swizzle row
constexpr int kY1Cols = 33;        // conv1 outputs under an 8x16 conv2 tile: 17 rows x 33 columns
constexpr int kY1Px = 17 * kY1Cols;
constexpr int kMTiles = 5;         // ceil(561 / 128)
constexpr int kPatchRows = 19, kPatchCols = 35, kPatchPitch = 36;
constexpr int kFront = 128;
constexpr int kEpiGroup = 128;                             // one epilogue warpgroup (TMEM lane quadrants 0..3)
constexpr int kEpiGroups = 2;                              // group 0: conv1 tiles 0-2; group 1: conv1 tiles 3-4 + conv2
constexpr int kEpi = kEpiGroups * kEpiGroup;
constexpr int kThreads = kFront + 32 + kEpi;               // 416 (13 warps: registers are allocated as for 16)
constexpr int kTmemCols = 512;                             // 5 x 64 (conv1) + 2 x 64 (conv2) = 448 -> 512
constexpr int kAcc2Col = kMTiles * 64;

struct HeadSmem {
  static constexpr int kPlaneE = 8 * 16 * kRow;             // even-row planes: 8 rows of 16 pixels
  static constexpr int kPlaneO = 9 * 16 * kRow;             // odd-row planes: 9 rows
  static constexpr int kY1 = 0;
  static constexpr int kY1Bytes = 3 * (kPlaneE + kPlaneO);  // 104448
  static constexpr int kW2 = kY1 + kY1Bytes;                // 9 resident [64 cout][64 cin] tap tiles
  static constexpr int kW2Bytes = 9 * 64 * kRow;            // 73728
  static constexpr int kA1 = kW2 + kW2Bytes;                // im2col tile: two M tiles side by side (64 B each)
  static constexpr int kA1Bytes = kM * kRow;
  static constexpr int kW1 = kA1 + kA1Bytes;                // [64 cout][K = 32 in the first 64 B]
  static constexpr int kW1Bytes = 64 * kRow;
  static constexpr int kStg = kW1 + kW1Bytes;               // output staging [128 px][64 ch]
  static constexpr int kStgBytes = kM * kRow;
  static constexpr int kPatch = kStg + kStgBytes;           // fp32 [3][19][36]
  static constexpr int kPatchBytes = 3 * kPatchRows * kPatchPitch * 4;
  static constexpr int kLut = kPatch + kPatchBytes;         // fp32 [3][256] (uint8 input only)
  static constexpr int kBar = (kLut + 3 * 256 * 4 + 7) / 8 * 8;
  static constexpr int kNumBars = 14;
  static constexpr int kTmemPtr = kBar + kNumBars * 8;
  static constexpr int kScale = (kTmemPtr + 8 + 15) / 16 * 16;   // conv2 scale[64], shift[64] (read as float4)
  static constexpr int kTotal = kScale + 128 * 4;
  static constexpr int kDynamic = kTotal + 1024;
};
static_assert(HeadSmem::kDynamic <= 232448, "enc_head: shared memory budget exceeded");
static_assert(HeadSmem::kW2 % 1024 == 0 && HeadSmem::kA1 % 1024 == 0 && HeadSmem::kW1 % 1024 == 0 &&
                  HeadSmem::kStg % 1024 == 0 && HeadSmem::kPlaneE % 1024 == 0 && HeadSmem::kPlaneO % 1024 == 0,
              "enc_head: operand tiles must sit on 1024-byte boundaries");

// barrier indices
enum { B_W = 0, B_A1_FULL, B_A1_EMPTY, B_ACC1_FULL, B_Y1_FULL, B_ACC2_FULL0, B_ACC2_FULL1, B_ACC2_EMPTY0, B_ACC2_EMPTY1,
       B_ACC1_EMPTY0 /* .. +4: one per conv1 accumulator */ };

struct HeadParams {
  CUtensorMap w2_map;   // packed conv2 weight [rows][576], box [64][64]
  CUtensorMap y_map;    // output NHWC [n][h/2][w/2][pix], box [64 ch][16][8][1]
  const void* x;        // fp32 NCHW views (B, c_total, H, W), or uint8 frames (B, agents_total, H, W, 3)
  const float* lut;     // [3][256] loader transform (uint8 input)
  const float* w1;      // conv1 weight fp32 [64][27]
  const float* scale1;  // folded BN of conv1
  const float* shift1;
  const float* scale2;
  const float* shift2;
  int b_sz, n_agents, c_total, c_first;  // uint8 input: c_total = agents_total, c_first = agent_first
  int h, w;             // image extent
  int tiles_w, tiles_h, num_tiles;
  int act;              // output storage (and operand element type); y1 is always ONE plane of that element type
  int y_cstride, y_coffset;
  long long* dbg;       // optional [32] cycle counters of CTA 0 (w2c_debug_enc_head_timing), else NULL
};

// cycle counters of CTA 0's warp roles: compiled in only with -DW2C_HEAD_TIMING (tools/time_enc_head.py builds that
// variant); the product build carries none of it
#ifdef W2C_HEAD_TIMING
#define TICK() clock64()
#define TIMING(...) __VA_ARGS__
#else
#define TICK() 0ll
#define TIMING(...)
#endif

__device__ __forceinline__ uint32_t sw_chunk(uint32_t row, uint32_t chunk) { return row * kRow + ((chunk ^ (row & 7u)) << 4); }

// byte offset of y1 plane (row parity rp: 0 even / 1 odd; column variant cv: 0 even, 1 odd-left, 2 odd-right)
__device__ __forceinline__ uint32_t plane_off(int rp, int cv) {
  return rp == 0 ? HeadSmem::kY1 + cv * HeadSmem::kPlaneE : HeadSmem::kY1 + 3 * HeadSmem::kPlaneE + cv * HeadSmem::kPlaneO;
}

struct TileAt {
  int img, oh0, ow0;
};
__device__ __forceinline__ TileAt tile_at(const HeadParams& p, int tile) {
  TileAt t;
  t.ow0 = (tile % p.tiles_w) * 16;
  const int r = tile / p.tiles_w;
  t.oh0 = (r % p.tiles_h) * 8;
  t.img = r / p.tiles_h;
  return t;
}

// Where conv1 pixel q (0..560, even rows first) of a tile lives in the y1 planes: dst0 / dst1 = byte offset of the pixel
// row in its plane(s) with +16 * (idx & 7) folded in for the swizzle (dst1 < 0: one destination only; both < 0: no
// pixel), drc = (dr << 8) | dc with image row R = 2*oh0 - 1 + dr and column C = 2*ow0 - 1 + dc.
struct Y1Slot {
  int dst0, dst1, drc;
};
// conv1 pixel q of a tile -> (row order ro: 0-7 the even image rows, 8-16 the odd ones; column cq = image column -
// (2*ow0 - 1)). Within a row the 16 even image columns (cq odd) come first, then the 17 odd ones: consecutive lanes
// then write consecutive pixels of ONE y1 plane (distinct bank groups within every quarter warp - enumerating the
// columns in image order put an even-plane and an odd-plane pixel with the same swizzle phase next to each other: two
// wavefronts per quarter warp on every 16-byte store) and read consecutive words of one column-parity half of the patch.
__device__ __forceinline__ void y1_pixel(int q, int& ro, int& cq) {
  ro = q / kY1Cols;
  const int j = q % kY1Cols;
  cq = j < 16 ? 2 * j + 1 : 2 * (j - 16);
}
__device__ __forceinline__ Y1Slot y1_slot(int q) {
  Y1Slot s{-1, -1, 0};
  if (q < kY1Px) {
    int ro, cq;
    y1_pixel(q, ro, cq);
    const int rp = ro < 8 ? 0 : 1, pri = ro < 8 ? ro : ro - 8;
    const int dr = ro < 8 ? 2 * ro + 1 : 2 * pri;
    s.drc = (dr << 8) | cq;
    if (cq & 1) {   // even image column -> the even-column plane
      const int idx = (cq - 1) >> 1;
      s.dst0 = plane_off(rp, 0) + (pri * 16 + idx) * kRow + ((idx & 7) << 4);
    } else {        // odd image column jj = cq / 2 of 0..16: left variant holds 0..15, right variant 1..16
      const int jj = cq >> 1;
      if (jj < 16) s.dst0 = plane_off(rp, 1) + (pri * 16 + jj) * kRow + ((jj & 7) << 4);
      if (jj >= 1) {
        const int d = plane_off(rp, 2) + (pri * 16 + jj - 1) * kRow + (((jj - 1) & 7) << 4);
        if (s.dst0 < 0) s.dst0 = d; else s.dst1 = d;
      }
    }
  }
  return s;
}

// One conv1 accumulator row (64 fp32 columns at TMEM address taddr) -> ReLU -> 16 bit -> its slot(s) in the y1 planes.
__device__ __forceinline__ void drain_conv1_row(uint8_t* smem, uint32_t taddr, const Y1Slot s, int Rb, int Cb, int h,
                                                int w, bool f16, uint64_t* empty_bar) {
  uint32_t r[64];
  ptx::tmem_ld_32x32b_x32(taddr, r);
  ptx::tmem_ld_32x32b_x32(taddr + 32, r + 32);
  ptx::tmem_ld_wait();
  ptx::tc_fence_before();
  ptx::mbar_arrive(empty_bar);   // the accumulator is in registers: conv1 of the next tile may overwrite it
  if (s.dst0 < 0) return;
  const int R = Rb + (s.drc >> 8), C = Cb + (s.drc & 255);
  // conv2 zero-pads ITS input: conv1 outputs outside the image are zeros, not conv1 of the padded image
  const bool inside = R >= 0 && R < h && C >= 0 && C < w;
  // convert everything first (32 distinct registers), then issue the stores back to back: reusing four registers
  // per chunk made every conversion wait for the previous store to read them (ncu: short-scoreboard stalls)
  uint32_t pk[32];
#pragma unroll
  for (int j = 0; j < 32; ++j)
    pk[j] = inside ? ptx::pack_act2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]), true, f16) : 0u;
#pragma unroll
  for (int c8 = 0; c8 < 8; ++c8)
    *reinterpret_cast<uint4*>(smem + (static_cast<uint32_t>(s.dst0) ^ (c8 << 4))) =
        make_uint4(pk[4 * c8], pk[4 * c8 + 1], pk[4 * c8 + 2], pk[4 * c8 + 3]);
  if (s.dst1 >= 0) {
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8)
      *reinterpret_cast<uint4*>(smem + (static_cast<uint32_t>(s.dst1) ^ (c8 << 4))) =
          make_uint4(pk[4 * c8], pk[4 * c8 + 1], pk[4 * c8 + 2], pk[4 * c8 + 3]);
  }
}

template <bool U8>
__global__ void __launch_bounds__(kThreads, 1) enc_head_kernel(const __grid_constant__ HeadParams p) {
  using L = HeadSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBar);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kTmemPtr);
  float* s_scale = reinterpret_cast<float*>(smem + L::kScale);
  float* s_shift = s_scale + 64;
  float* s_patch = reinterpret_cast<float*>(smem + L::kPatch);
  float* s_lut = reinterpret_cast<float*>(smem + L::kLut);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool f16 = act_is_f16(p.act);

  // ---------------------------------------------------------------- one-time setup
  if (tid < kFront) {
    // conv1 operand B: scale1 * w1 (k < 27), shift1 at k = 27 (the weight of the constant-1 input), zeros after
    for (int i = tid; i < 64 * 4; i += kFront) {
      const int co = i >> 2, c = i & 3;
      const float sc = p.scale1[co];
      uint4 hv;
      __nv_bfloat16* hb = reinterpret_cast<__nv_bfloat16*>(&hv);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = c * 8 + e;
        const float v = k < 27 ? p.w1[co * 27 + k] * sc : (k == 27 ? p.shift1[co] : 0.f);
        hb[e] = float_to_elem(v, f16);
      }
      *reinterpret_cast<uint4*>(smem + L::kW1 + sw_chunk(co, c)) = hv;
    }
    if constexpr (U8)
      for (int i = tid; i < 3 * 256; i += kFront) s_lut[i] = p.lut[i];
  } else if (tid >= kFront + 32) {
    const int e = tid - (kFront + 32);
    if (e < 64) s_scale[e] = p.scale2[e], s_shift[e] = p.shift2[e];
  }
  if (tid == 0) {
    ptx::prefetch_tensormap(&p.w2_map);
    ptx::prefetch_tensormap(&p.y_map);
    ptx::mbar_init(&bars[B_W], 1);
    ptx::mbar_init(&bars[B_A1_FULL], kFront);
    ptx::mbar_init(&bars[B_A1_EMPTY], 1);
    ptx::mbar_init(&bars[B_ACC1_FULL], 1);
    for (int i = 0; i < kMTiles; ++i) ptx::mbar_init(&bars[B_ACC1_EMPTY0 + i], kEpiGroup);
    ptx::mbar_init(&bars[B_Y1_FULL], kEpi);
    ptx::mbar_init(&bars[B_ACC2_FULL0], 1);
    ptx::mbar_init(&bars[B_ACC2_FULL1], 1);
    ptx::mbar_init(&bars[B_ACC2_EMPTY0], kEpiGroup);
    ptx::mbar_init(&bars[B_ACC2_EMPTY1], kEpiGroup);
    ptx::fence_barrier_init();
  }
  if (warp == 4) {
    ptx::tmem_alloc(tmem_ptr, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < 4) {
    // ================================================================ front: weights, patch, im2col
    if (warp == 0 && ptx::elect_one_sync()) {
      ptx::mbar_arrive_expect_tx(&bars[B_W], L::kW2Bytes);
      for (int t = 0; t < 9; ++t) ptx::tma_load_2d(&p.w2_map, &bars[B_W], smem + L::kW2 + t * 64 * kRow, t * 64, 0);
    }
    // patch loader: thread <-> (channel, patch column), looping over the 19 patch rows with plain pointer increments
    // (uint8 frames: thread <-> byte of the 105-byte RGB row, channel = BGR index of that byte, airsim_loader.py:521)
    const size_t plane = static_cast<size_t>(p.h) * p.w;
    const int l_ch = U8 ? 2 - tid % 3 : tid / kPatchPitch;
    const int l_pc = U8 ? tid / 3 : tid % kPatchPitch;
    const bool l_on = U8 ? tid < kPatchCols * 3 : (tid < 3 * kPatchPitch && l_pc < kPatchCols);
    float pre[kPatchRows];
    auto load_patch = [&](int tile) {
      const TileAt t = tile_at(p, tile);
      const int agent = t.img / p.b_sz, bat = t.img % p.b_sz;
      const int r0 = 2 * t.oh0 - 2, C = 2 * t.ow0 - 2 + l_pc;
      const bool col_ok = l_on && C >= 0 && C < p.w;
      if constexpr (U8) {
        const uint8_t* xin = static_cast<const uint8_t*>(p.x) +
                             (static_cast<size_t>(bat) * p.c_total + p.c_first + agent) * plane * 3 +
                             static_cast<size_t>(C) * 3 + (tid % 3);
        const float* lut = s_lut + l_ch * 256;
#pragma unroll
        for (int pr = 0; pr < kPatchRows; ++pr) {
          const int R = r0 + pr;
          pre[pr] = (col_ok && R >= 0 && R < p.h) ? lut[__ldg(xin + static_cast<size_t>(R) * p.w * 3)] : 0.f;
        }
      } else {
        const float* xin = static_cast<const float*>(p.x) +
                           (static_cast<size_t>(bat) * p.c_total + p.c_first + 3 * agent + l_ch) * plane + C;
#pragma unroll
        for (int pr = 0; pr < kPatchRows; ++pr) {
          const int R = r0 + pr;
          pre[pr] = (col_ok && R >= 0 && R < p.h) ? __ldg(xin + static_cast<size_t>(R) * p.w) : 0.f;
        }
      }
    };
    // patch layout: [channel][row][column parity][18] (a row is still 36 words): for a fixed filter column the lanes of
    // a warp - consecutive pixels of one column parity, see y1_pixel - read consecutive words
    auto store_patch = [&]() {
      if (l_on) {
        float* dst = s_patch + l_ch * kPatchRows * kPatchPitch + (l_pc & 1) * (kPatchPitch / 2) + (l_pc >> 1);
#pragma unroll
        for (int pr = 0; pr < kPatchRows; ++pr) dst[pr * kPatchPitch] = pre[pr];
      }
    };
    // this thread's conv1 pixel in each of the five M tiles: word offsets of its window's filter columns 0 and 1 in
    // the patch (column 2 = column 0 + one word: same parity, next pixel pair). Rows without a pixel reuse window 0.
    auto patch_col = [](int c) { return (c & 1) * (kPatchPitch / 2) + (c >> 1); };
    int win0[kMTiles], win1[kMTiles];
#pragma unroll
    for (int mt = 0; mt < kMTiles; ++mt) {
      const int q = mt * kM + tid;
      int ro, cq;
      y1_pixel(q < kY1Px ? q : 0, ro, cq);
      const int prow = ro < 8 ? 2 * ro + 1 : 2 * (ro - 8);
      win0[mt] = prow * kPatchPitch + patch_col(cq);
      win1[mt] = prow * kPatchPitch + patch_col(cq + 1);
    }

    int tile = blockIdx.x;
    if (tile < p.num_tiles) load_patch(tile);
    uint32_t n = 0;  // A1 fill counter (three per tile)
    TIMING(long long d_wait = 0; long long d_im2col = 0; long long d_patch = 0; long long t0;)
    for (; tile < p.num_tiles; tile += gridDim.x) {
      TIMING(t0 = TICK();)
      store_patch();
      ptx::named_bar_sync(1, kFront);
      const int next = tile + gridDim.x;
      if (next < p.num_tiles) load_patch(next);   // in flight during the im2col below
      TIMING(d_patch += TICK() - t0;)
#pragma unroll
      for (int g = 0; g < 3; ++g, ++n) {
        TIMING(t0 = TICK();)
        if (n > 0) ptx::mbar_wait(&bars[B_A1_EMPTY], (n - 1) & 1);   // the MMAs that read the previous fill are done
        TIMING(d_wait += TICK() - t0;)
        TIMING(t0 = TICK();)
#pragma unroll
        for (int s = 0; s < (g < 2 ? 2 : 1); ++s) {
          const float* p0 = s_patch + win0[2 * g + s];
          const float* p1 = s_patch + win1[2 * g + s];
          float in[27];
#pragma unroll
          for (int ci = 0; ci < 3; ++ci)
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
              const int ro = (ci * kPatchRows + kh) * kPatchPitch;
              in[ci * 9 + kh * 3 + 0] = p0[ro];
              in[ci * 9 + kh * 3 + 1] = p1[ro];
              in[ci * 9 + kh * 3 + 2] = p0[ro + 1];
            }
          // (a row without a pixel computes on window 0; its accumulator row is never read)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 hv;
            uint32_t* hw = reinterpret_cast<uint32_t*>(&hv);
#pragma unroll
            for (int e2 = 0; e2 < 4; ++e2) {
              const int k = c * 8 + 2 * e2;
              const float v0 = k < 27 ? in[k < 27 ? k : 0] : (k == 27 ? 1.f : 0.f);
              const float v1 = k + 1 < 27 ? in[k + 1 < 27 ? k + 1 : 0] : (k + 1 == 27 ? 1.f : 0.f);
              hw[e2] = ptx::pack_act2(v0, v1, false, f16);
            }
            *reinterpret_cast<uint4*>(smem + L::kA1 + sw_chunk(tid, 4 * s + c)) = hv;
          }
        }
        ptx::fence_proxy_async();
        ptx::mbar_arrive(&bars[B_A1_FULL]);
        TIMING(d_im2col += TICK() - t0;)
      }
      ptx::named_bar_sync(1, kFront);   // every thread has read the patch before the next one is stored
    }
    TIMING(if (p.dbg && blockIdx.x == 0 && tid == 0) p.dbg[0] = d_patch, p.dbg[1] = d_wait, p.dbg[2] = d_im2col;)
  } else if (warp == 4) {
    // ================================================================ MMA issuer
    if (ptx::elect_one_sync()) {
      const uint32_t idesc = ptx::make_idesc_16(kM, 64, f16);
      const uint64_t a1 = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + L::kA1));
      const uint64_t w1 = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + L::kW1));
      const uint64_t w2 = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + L::kW2));
      const uint64_t y1 = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + L::kY1));
      ptx::mbar_wait(&bars[B_W], 0);
      ptx::tc_fence_after();
      uint32_t n = 0;
      // conv1 of local tile jt: five M tiles, K = 32 = two k-steps, each into its own accumulator as soon as the
      // epilogue has drained that accumulator's previous contents
      TIMING(long long m_a1 = 0; long long m_acc1 = 0; long long m_y1 = 0; long long m_acc2 = 0; long long t0;)
      auto conv1 = [&](int jt) {
#pragma unroll 1
        for (int g = 0; g < 3; ++g, ++n) {
          TIMING(t0 = TICK();)
          ptx::mbar_wait(&bars[B_A1_FULL], n & 1);
          TIMING(m_a1 += TICK() - t0;)
          for (int s = 0; s < (g < 2 ? 2 : 1); ++s) {
            const int mt = 2 * g + s;
            TIMING(t0 = TICK();)
            if (jt > 0) ptx::mbar_wait(&bars[B_ACC1_EMPTY0 + mt], (jt - 1) & 1);
            TIMING(m_acc1 += TICK() - t0;)
            ptx::tc_fence_after();
            const uint32_t d = tmem_base + mt * 64;
            ptx::umma_bf16(d, a1 + 4 * s, w1, idesc, 0);
            ptx::umma_bf16(d, a1 + 4 * s + 2, w1 + 2, idesc, 1);
          }
          ptx::umma_commit(&bars[B_A1_EMPTY]);
        }
        ptx::umma_commit(&bars[B_ACC1_FULL]);
      };
      int it = 0;
      if (blockIdx.x < p.num_tiles) conv1(0);
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        // conv1 of the NEXT tile goes first: its im2col and MMAs overlap this tile's first epilogue, and the front
        // warps are not held up behind the wait for y1 below
        if (tile + static_cast<int>(gridDim.x) < p.num_tiles) conv1(it + 1);
        // ---- conv2: 9 taps x 4 k-steps on the y1 planes
        TIMING(t0 = TICK();)
        ptx::mbar_wait(&bars[B_Y1_FULL], it & 1);
        TIMING(m_y1 += TICK() - t0;)
        const int b = it & 1;
        TIMING(t0 = TICK();)
        if (it >= 2) ptx::mbar_wait(&bars[B_ACC2_EMPTY0 + b], ((it >> 1) & 1) ^ 1);
        TIMING(m_acc2 += TICK() - t0;)
        ptx::tc_fence_after();
        const uint32_t d2 = tmem_base + kAcc2Col + b * 64;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const int rp = kh == 1 ? 0 : 1, cv = kw == 1 ? 0 : (kw == 0 ? 1 : 2);
            const uint32_t off = plane_off(rp, cv) - L::kY1 + (kh == 2 ? 16 * kRow : 0);
            const uint64_t ad = y1 + (off >> 4);
            const uint64_t bd = w2 + (((kh * 3 + kw) * 64 * kRow) >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) ptx::umma_bf16(d2, ad + 2 * k, bd + 2 * k, idesc, (kh | kw | k) != 0);
          }
        ptx::umma_commit(&bars[B_ACC2_FULL0 + b]);
      }
      TIMING(if (p.dbg && blockIdx.x == 0) p.dbg[4] = m_a1, p.dbg[5] = m_acc1, p.dbg[6] = m_y1, p.dbg[7] = m_acc2;)
    }
  } else {
    // ================================================================ epilogues (two warpgroups: warps 5-8, 9-12)
    // Everything here sits on the critical path between conv2 of one tile and conv2 of the next (y1 exists once), so
    // the five conv1 tiles are split over two warpgroups (3 + 2; the second also drains conv2), and everything that
    // does not depend on the tile is precomputed per thread.
    const int grp = (warp - 5) >> 2;            // 0, 1
    const int wq = warp & 3;                    // TMEM lane quadrant this warp may read
    const int m = wq * 32 + lane;               // accumulator row of this thread
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    const bool lead = warp == 9;                // issues the TMA stores (group 1)
    const int planes = act_planes(p.act);
    const int mt0 = grp == 0 ? 0 : 3;
    // this thread's pixel in the group's conv1 tiles (scalars, not arrays: an indexed array here ended up in local
    // memory, and with the L1 carved down to nothing by 227 KB of shared memory every access cost an L2 trip)
    const Y1Slot slot_a = y1_slot(mt0 * kM + m);
    const Y1Slot slot_b = y1_slot((mt0 + 1) * kM + m);
    const Y1Slot slot_c = grp == 0 ? y1_slot(2 * kM + m) : Y1Slot{-1, -1, 0};
    const uint32_t stg_row = L::kStg + m * kRow + ((m & 7) << 4);

    auto epilogue2 = [&](int jt, int jtile) {   // conv2 accumulator of local tile jt -> staging -> TMA store (group B)
      const int b = jt & 1;
      const TileAt t = tile_at(p, jtile);
      for (int pln = 0; pln < planes; ++pln) {
        if (lead && ptx::elect_one_sync()) ptx::bulk_wait_group_read<0>();   // the previous store has read the tile
        ptx::named_bar_sync(2, kEpiGroup);
        uint32_t r[64];
        ptx::tmem_ld_32x32b_x32(t_row + kAcc2Col + b * 64, r);
        ptx::tmem_ld_32x32b_x32(t_row + kAcc2Col + b * 64 + 32, r + 32);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          uint4 pk;
          uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
          float sc[8], sh[8];
          *reinterpret_cast<float4*>(sc) = reinterpret_cast<const float4*>(s_scale)[2 * c8];
          *reinterpret_cast<float4*>(sc + 4) = reinterpret_cast<const float4*>(s_scale)[2 * c8 + 1];
          *reinterpret_cast<float4*>(sh) = reinterpret_cast<const float4*>(s_shift)[2 * c8];
          *reinterpret_cast<float4*>(sh + 4) = reinterpret_cast<const float4*>(s_shift)[2 * c8 + 1];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = c8 * 8 + 2 * j;
            const float v0 = fmaf(__uint_as_float(r[c]), sc[2 * j], sh[2 * j]);
            const float v1 = fmaf(__uint_as_float(r[c + 1]), sc[2 * j + 1], sh[2 * j + 1]);
            if (planes == 1) {
              pw[j] = ptx::pack_act2(v0, v1, true, f16);
            } else {
              uint32_t hi, lo;
              split_act2(fmaxf(v0, 0.f), fmaxf(v1, 0.f), f16, hi, lo);
              pw[j] = pln == 0 ? hi : lo;
            }
          }
          *reinterpret_cast<uint4*>(smem + (stg_row ^ (c8 << 4))) = pk;
        }
        if (pln == planes - 1) {   // the accumulator has been read for the last time
          ptx::tc_fence_before();
          ptx::mbar_arrive(&bars[B_ACC2_EMPTY0 + b]);
        }
        ptx::fence_proxy_async();
        ptx::named_bar_sync(2, kEpiGroup);
        if (lead && ptx::elect_one_sync()) {
          ptx::tma_store_4d(&p.y_map, smem + L::kStg, p.y_coffset + pln * p.y_cstride, t.ow0, t.oh0, t.img);
          ptx::bulk_commit_group();
        }
      }
    };

    int it = 0, prev_tile = -1;
    TIMING(long long e_w1 = 0; long long e_w2 = 0; long long e_e1 = 0; long long e_e2 = 0; long long e_ld = 0;
           long long e_fence = 0; long long t0; const long long t_start = TICK();)
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const TileAt t = tile_at(p, tile);
      const int Rb = 2 * t.oh0 - 1, Cb = 2 * t.ow0 - 1;
      TIMING(t0 = TICK();)
      ptx::mbar_wait(&bars[B_ACC1_FULL], it & 1);
      TIMING(e_w1 += TICK() - t0;)
      TIMING(t0 = TICK();)
      if (it > 0) ptx::mbar_wait(&bars[B_ACC2_FULL0 + ((it - 1) & 1)], ((it - 1) >> 1) & 1);   // conv2 of the previous tile has read y1
      TIMING(e_w2 += TICK() - t0;)
      ptx::tc_fence_after();
      TIMING(t0 = TICK();)
      // ---- epilogue 1: this group's conv1 accumulators -> ReLU -> 16 bit -> the y1 planes
      drain_conv1_row(smem, t_row + mt0 * 64, slot_a, Rb, Cb, p.h, p.w, f16, &bars[B_ACC1_EMPTY0 + mt0]);
      drain_conv1_row(smem, t_row + (mt0 + 1) * 64, slot_b, Rb, Cb, p.h, p.w, f16, &bars[B_ACC1_EMPTY0 + mt0 + 1]);
      if (grp == 0) drain_conv1_row(smem, t_row + 2 * 64, slot_c, Rb, Cb, p.h, p.w, f16, &bars[B_ACC1_EMPTY0 + 2]);
      TIMING(const long long tf = TICK();)
      ptx::fence_proxy_async();
      ptx::mbar_arrive(&bars[B_Y1_FULL]);
      TIMING(e_fence += TICK() - tf;)
      TIMING(e_e1 += TICK() - t0;)
      TIMING(t0 = TICK();)
      // ---- epilogue 2 of the PREVIOUS tile (its conv2 completion was waited for above); overlaps this tile's conv2
      if (grp == 1 && it > 0) epilogue2(it - 1, prev_tile);
      TIMING(e_e2 += TICK() - t0;)
      prev_tile = tile;
    }
    TIMING(if (p.dbg && blockIdx.x == 0 && lane == 0 && (warp == 5 || warp == 9)) {
      long long* d = p.dbg + (warp == 5 ? 8 : 16);
      d[0] = e_w1, d[1] = e_w2, d[2] = e_e1, d[3] = e_e2, d[4] = TICK() - t_start, d[5] = it, d[6] = e_ld, d[7] = e_fence;
    })
    if (grp == 1 && it > 0) {
      ptx::mbar_wait(&bars[B_ACC2_FULL0 + ((it - 1) & 1)], ((it - 1) >> 1) & 1);
      ptx::tc_fence_after();
      epilogue2(it - 1, prev_tile);
    }
    __syncwarp();
    if (lead && ptx::elect_one_sync()) ptx::bulk_wait_group<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    __syncwarp();
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <bool U8>
int launch_head(const HeadParams& p, cudaStream_t stream) {
  static DeviceOnce attr;
  if (int rc = attr.ensure([] {
        return cudaFuncSetAttribute(enc_head_kernel<U8>, cudaFuncAttributeMaxDynamicSharedMemorySize, HeadSmem::kDynamic);
      }, "enc_head_kernel"))
    return rc;
  const int sms = device_sm_count();
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  enc_head_kernel<U8><<<grid, kThreads, HeadSmem::kDynamic, stream>>>(p);
  W2C_CHECK_LAUNCH("enc_head_kernel");
  return W2C_OK;
}

long long* g_head_dbg = nullptr;

}  // namespace
}  // namespace w2c

using namespace w2c;

// Debug aid (not part of the product path): cycle counters of CTA 0's warp roles are written to `buf` (device memory,
// 32 x int64) by every later w2c_enc_head_fwd launch; NULL switches it off again.
extern "C" void w2c_debug_enc_head_timing(long long* buf) { g_head_dbg = buf; }

extern "C" int w2c_enc_head_fwd(const w2c_enc_head_args* a, w2c_stream_t stream) {
  if (!a) return set_error(W2C_ERR_INVALID, "enc_head: args is NULL");
  W2C_CHECK_ARG(a->x && a->w1 && a->scale1 && a->shift1 && a->w2 && a->scale2 && a->shift2 && a->y,
                "enc_head: null pointer argument");
  W2C_CHECK_ARG(!a->x_u8 || a->lut, "enc_head: uint8 frames need the loader-transform table");
  W2C_CHECK_ARG(act_valid(a->act), "enc_head: bad act %d", a->act);
  W2C_CHECK_ARG(a->b > 0 && a->n_agents > 0 && a->h > 0 && a->w > 0 && a->h % 2 == 0 && a->w % 2 == 0,
                "enc_head: bad extent b=%d agents=%d %dx%d (H, W even)", a->b, a->n_agents, a->h, a->w);
  if (a->x_u8)
    W2C_CHECK_ARG(a->c_first >= 0 && a->c_first + a->n_agents <= a->c_total, "enc_head: agent window outside the frames");
  else
    W2C_CHECK_ARG(a->c_first >= 0 && a->c_first + 3 * a->n_agents <= a->c_total, "enc_head: channel window outside the views");
  const int planes = act_planes(a->act);
  const int y_cstride = a->y_cstride > 0 ? a->y_cstride : 64;
  W2C_CHECK_ARG(a->y_coffset >= 0 && a->y_coffset + 64 <= y_cstride && y_cstride % 8 == 0 && a->y_coffset % 8 == 0,
                "enc_head: output channel slice out of range");
  HeadParams p{};
  p.x = a->x, p.lut = a->lut, p.w1 = a->w1, p.scale1 = a->scale1, p.shift1 = a->shift1;
  p.scale2 = a->scale2, p.shift2 = a->shift2;
  p.b_sz = a->b, p.n_agents = a->n_agents, p.c_total = a->c_total, p.c_first = a->c_first;
  p.h = a->h, p.w = a->w;
  const int ho = a->h / 2, wo = a->w / 2, n_img = a->b * a->n_agents;
  p.tiles_w = ceil_div(wo, 16), p.tiles_h = ceil_div(ho, 8);
  const long long tiles = static_cast<long long>(p.tiles_w) * p.tiles_h * n_img;
  W2C_CHECK_ARG(tiles < (1ll << 31), "enc_head: too many tiles");
  p.num_tiles = static_cast<int>(tiles);
  p.act = a->act, p.y_cstride = y_cstride, p.y_coffset = a->y_coffset;
  p.dbg = g_head_dbg;
  const cuuint64_t esz = 2;
  {
    // conv2 weight as packed by w2c_pack_conv_weight (cout = 64, cin = 64, 9 taps): [planes * 64][576]; the hi plane
    const cuuint64_t dims[2] = {576, (cuuint64_t)64 * planes};
    const cuuint64_t str[1] = {576 * esz};
    const cuuint32_t box[2] = {64, 64};
    int rc = encode_map(&p.w2_map, a->w2, 2, dims, str, box, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (rc) return rc;
  }
  {
    const cuuint64_t pix = (cuuint64_t)y_cstride * planes;
    const cuuint64_t dims[4] = {pix, (cuuint64_t)wo, (cuuint64_t)ho, (cuuint64_t)n_img};
    const cuuint64_t str[3] = {pix * esz, (cuuint64_t)wo * pix * esz, (cuuint64_t)ho * wo * pix * esz};
    const cuuint32_t box[4] = {64, 16, 8, 1};
    int rc = encode_map(&p.y_map, a->y, 4, dims, str, box, CU_TENSOR_MAP_L2_PROMOTION_NONE);
    if (rc) return rc;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return a->x_u8 ? launch_head<true>(p, s) : launch_head<false>(p, s);
}
