// Persistent, warp-specialised tcgen05 implicit-GEMM convolution (the production conv kernel).
//
// Same math and operand layout as conv_tc.cu (per-tap TMA windows of the NHWC input, K-major packed weights, fp32
// accumulation in TMEM, fused scale/shift(+residual)+ReLU), restructured after the first B200 profiles showed the
// one-tile-per-CTA kernel paying ~10 us of prologue + unoverlapped epilogue per tile on the small-K layers
// (20k-80k CTAs per launch) and issuing 16-byte scattered stores:
//   * one CTA per SM loops over tiles (static round-robin, n-tile fastest so concurrent CTAs share input windows);
//     barriers, the TMEM allocation and the scale/shift tables are set up once per CTA
//   * the accumulator is double-buffered in TMEM (2 x BLOCK_N columns): the MMA warp starts tile i+1 while the
//     epilogue warps drain tile i
//   * the epilogue converts to bf16, writes a [128 pixel][64 channel] 128B-swizzled staging tile and ONE thread
//     issues a TMA tensor store per 64 channels (full-line writes, edge tiles clipped by the TMA unit); the
//     transposed conv writes through four parity-strided output tensor maps
//   warp 0: TMA producer   warp 1: MMA issuer + TMEM owner   warps 2-5: epilogue
#include "conv_plan.cuh"
#include "ptx.cuh"

namespace w2c {

int encode_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, CUtensorMapL2promotion promo);

int encode_map_f32(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims,
                   const cuuint64_t* strides_bytes, const cuuint32_t* box);

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kAStageBytes = kBlockM * kBlockK * 2;
constexpr int kEpiThreads = 128;  // one epilogue warpgroup (TMEM lane quadrants 0..3)
constexpr int kStagingBytes = kBlockM * 64 * 2;  // [128 px][64 ch] bf16
constexpr int kMaxCout = 512;

struct PersParams {
  CUtensorMap a_map[4];
  CUtensorMap b_map;
  CUtensorMap y_map[4];
  ConvPlan plan;
  int tw, th, tn;
  int tiles_w, tiles_h, tiles_img;
  int n_tiles, m_tiles, total_tiles;
  int groups;     // work items dealt to the CTAs: total_tiles, or m_tiles * n_tiles when cls_shift = 2
  int cls_shift;  // 2: a CTA runs the four output-parity classes of a transposed-conv tile back to back
  int tma_store;  // 1: NHWC output through the staging tile + TMA store; 0: direct stores
  int nchw_tma;   // 1: fp32 NCHW logits through a [cout][8][16] staging box + TMA store (y_map[0] is that fp32 map)
  int ctas_per_sm;  // persistent CTAs per SM (small-footprint instantiations: several MMA issuers per SM)
};

// GROUP = k-blocks per pipeline stage. 1: one filter tap per stage (16 KB input window + one weight tile).
// 3 ("row-halo", 3x3 stride-1 convs on 8x16 tiles): the three taps of one filter COLUMN share one input box of
// 8+2 rows; tap kh reads it kh rows further down = +kh*2048 B, still 1024-byte aligned, so it is just another UMMA
// descriptor start at full operand bandwidth. One barrier round trip and ONE tcgen05.commit then cover 12 MMAs
// instead of 4, and the activations cross L2 -> SM 3.75x instead of 9x per tile.
// EG = epilogue warpgroups. 2: warps 2-5 drain accumulator 0 (even tiles) while warps 6-9 drain accumulator 1 (odd
// tiles), each with its own staging tile: the small-K layers are epilogue-bound with one group (ncu: 64->11 logits
// layer 9 % tensor active, epilogue serialising ~2000 cycles per tile).
// RES = weight tiles kept RESIDENT in shared memory (0: weights stream through the ring with the activations).
// The 64-channel layers re-fetched their whole weight matrix (72-144 KB) from L2 for every 128-pixel tile - a third
// to two thirds of their L2 -> SM bytes, on layers ncu shows bound by bytes in flight (23 % tensor, 50 % DRAM).
// With RES = ktot/64 the producer loads every [BLOCK_N][64] weight tile once per CTA (slot = k/64) and the ring
// carries activations only.
template <int BLOCK_N, int STAGES, int GROUP, int EG, int RES>
struct PersSmem {
  static constexpr int kThreads = 64 + EG * kEpiThreads;
  static constexpr int kNumStaging = BLOCK_N >= 64 ? ((EG == 1 && (BLOCK_N == 256 || (GROUP == 3 && BLOCK_N == 128) || STAGES <= 3)) ? 1 : 2) : 0;
  static constexpr int kABoxBytes = GROUP == 3 ? (8 + 2) * 16 * 128 : kAStageBytes;
  static constexpr int kAStage = GROUP == 3 ? 20 * 1024 : kAStageBytes;
  static constexpr int kBTileBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kBStageBytes = RES > 0 ? 0 : GROUP * kBTileBytes;
  static constexpr int kAOff = 0;
  static constexpr int kBOff = STAGES * kAStage;  // weight ring, or the resident weight tiles
  static constexpr int kStgOff = kBOff + (RES > 0 ? RES * kBTileBytes : STAGES * kBStageBytes);
  // fp32 [16 ch][128 px] staging box per epilogue group for the NCHW logits layer (BLOCK_N = 16 only)
  static constexpr int kNchwOff = kStgOff + kNumStaging * kStagingBytes;
  static constexpr int kNchwBytes = BLOCK_N == 16 ? 16 * kBlockM * 4 : 0;
  static constexpr int kBarOff = kNchwOff + EG * kNchwBytes;  // full[S] empty[S] tfull[2] tempty[2] res
  static constexpr int kTmemPtrOff = kBarOff + (2 * STAGES + 5) * 8;
  static constexpr int kScaleOff = kTmemPtrOff + 8;
  static constexpr int kShiftOff = kScaleOff + kMaxCout * 4;
  static constexpr int kTotal = kShiftOff + kMaxCout * 4;
  static constexpr int kDynamicBytes = kTotal + 1024;
};

struct TileCoord {
  int cls, n0, w0, h0, i0;
};

// The it-th tile of this CTA. Groups (m-tile, n-tile; n fastest so concurrent CTAs share input windows) are dealt
// round-robin; the output-parity classes of a transposed conv run back to back INSIDE the CTA, so the input window
// of a group is fetched from HBM once (class-major order re-read the whole input per class: ncu 4x the bytes).
__device__ __forceinline__ bool tile_at(const PersParams& p, int it, int block_n, TileCoord& c) {
  const int g = blockIdx.x + (it >> p.cls_shift) * gridDim.x;
  if (g >= p.groups) return false;
  const int n_tile = g % p.n_tiles;
  int m = g / p.n_tiles;
  if (p.cls_shift) {
    c.cls = it & 3;
  } else {  // few tiles: classes stay separate work items (class slowest) so the CTAs share them evenly
    c.cls = m / p.m_tiles;
    m -= c.cls * p.m_tiles;
  }
  c.n0 = n_tile * block_n;
  c.w0 = (m % p.tiles_w) * p.tw;
  m /= p.tiles_w;
  c.h0 = (m % p.tiles_h) * p.th;
  c.i0 = (m / p.tiles_h) * p.tn;
  return true;
}

// STATS: the epilogue also sums the stored output per channel (train-mode BatchNorm, ConvPlan::bn_sums). A separate
// instantiation, so that the eval-mode kernels carry none of its registers (the two-warpgroup variants sit at their
// 168-register ceiling).
template <int BLOCK_N, int STAGES, int GROUP, int EG, int RES, bool STATS>
__global__ void __launch_bounds__(64 + EG * kEpiThreads, 1) conv_persv1_kernel(const __grid_constant__ PersParams p) {
  using L = PersSmem<BLOCK_N, STAGES, GROUP, EG, RES>;
  constexpr int kTmemCols = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
  constexpr uint32_t kStageTx = L::kABoxBytes + L::kBStageBytes;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* res_bar = tempty_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kTmemPtrOff);
  float* s_scale = reinterpret_cast<float*>(smem + L::kScaleOff);
  float* s_shift = reinterpret_cast<float*>(smem + L::kShiftOff);

  const ConvPlan& pl = p.plan;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks = pl.cin / kBlockK;
  const int npass = pl.npass;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&p.b_map);
    ptx::prefetch_tensormap(&p.a_map[0]);
    for (int s = 0; s < STAGES; ++s) ptx::mbar_init(&full_bar[s], 1), ptx::mbar_init(&empty_bar[s], 1);
    for (int s = 0; s < 2; ++s) ptx::mbar_init(&tfull_bar[s], 1), ptx::mbar_init(&tempty_bar[s], kEpiThreads);
    ptx::mbar_init(res_bar, 1);
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
      if constexpr (RES > 0) {
        ptx::mbar_arrive_expect_tx(res_bar, RES * L::kBTileBytes);
        for (int s = 0; s < RES; ++s)
          ptx::tma_load_2d(&p.b_map, res_bar, smem + L::kBOff + s * L::kBTileBytes, s * kBlockK, 0);
      }
      int stage = 0;
      uint32_t phase = 0;
      TileCoord tc;
      for (int it = 0; tile_at(p, it, BLOCK_N, tc); ++it) {
        const int ntaps = pl.ntaps[tc.cls];
        for (int pass = 0; pass < npass; ++pass) {
          // three-pass order: lo*hi, hi*lo, then hi*hi. The two correction passes run while the accumulator is still
          // small, so the tensor core's truncating fp32 adds cost them nothing; only the last pass accumulates at
          // full magnitude (hi*hi first made all three passes pay: a -1e-4 relative bias per K = 4608 layer)
          const bool a_lo = npass == 3 && pass == 0, b_lo = npass == 3 && pass == 1;
          const int a_c0 = pl.x_coffset + (a_lo ? pl.x_cstride : 0);
          const int b_row = tc.n0 + (b_lo ? pl.cout_pad : 0);
          if constexpr (GROUP == 3) {
            // 3x3 stride-1 conv: filter column kw -> one box of th+2 rows starting one row above the tile
            // (a_map[1]); its three weight tiles (kh = 0,1,2) land behind each other in the stage
            for (int kw = 0; kw < 3; ++kw) {
              int a_c = a_c0;
              int b_k = kw * pl.cin;
              for (int ch = 0; ch < chunks; ++ch, a_c += kBlockK, b_k += kBlockK) {
                ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                ptx::mbar_arrive_expect_tx(&full_bar[stage], kStageTx);
                ptx::tma_load_4d(&p.a_map[1], &full_bar[stage], smem + L::kAOff + stage * L::kAStage, a_c,
                                 tc.w0 + kw - 1, tc.h0 - 1, tc.i0);
                if constexpr (RES == 0) {
                  uint8_t* sb = smem + L::kBOff + stage * L::kBStageBytes;
#pragma unroll
                  for (int kh = 0; kh < 3; ++kh)
                    ptx::tma_load_2d(&p.b_map, &full_bar[stage], sb + kh * L::kBTileBytes, b_k + kh * 3 * pl.cin,
                                     b_row);
                }
                if (++stage == STAGES) stage = 0, phase ^= 1;
              }
            }
          } else {
            // canonical K order (ConvPlan::kw_major): (kw, chunk, kh) for 3x3 s1 convs - what the row-halo stages
            // above run - and (tap, chunk) otherwise
            const int n_outer = pl.kw_major ? 3 : ntaps, n_inner = pl.kw_major ? 3 : 1;
            for (int o = 0; o < n_outer; ++o)
              for (int ch = 0; ch < chunks; ++ch)
                for (int i = 0; i < n_inner; ++i) {
                  const Tap tp = pl.taps[tc.cls][pl.kw_major ? i * 3 + o : o];
                  ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                  ptx::mbar_arrive_expect_tx(&full_bar[stage], kStageTx);
                  ptx::tma_load_4d(&p.a_map[tp.map], &full_bar[stage], smem + L::kAOff + stage * L::kAStage,
                                   a_c0 + ch * kBlockK, tc.w0 + tp.dw, tc.h0 + tp.dh, tc.i0);
                  if constexpr (RES == 0)
                    ptx::tma_load_2d(&p.b_map, &full_bar[stage], smem + L::kBOff + stage * L::kBStageBytes,
                                     tp.wtap * pl.cin + ch * kBlockK, b_row);
                  if (++stage == STAGES) stage = 0, phase ^= 1;
                }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one_sync()) {
      // ===================== MMA issuer =====================
      const uint32_t idesc = ptx::make_idesc_16(kBlockM, BLOCK_N, act_is_f16(pl.act));
      constexpr uint32_t kBTile16 = L::kBTileBytes >> 4;
      const uint64_t a_desc0 = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + L::kAOff));
      const uint64_t b_desc0 = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + L::kBOff));
      int stage = 0;
      uint32_t phase = 0;
      // resident weights: weight-tap index of every (class, tap), 4 bits each, in one register - the issuing thread
      // must not wait on a parameter-space lookup per k-block (measured: 5-9 % slower than streaming the weights)
      uint64_t wtap_pack = 0;
      uint32_t base_pack = 0;  // first tap of each class, 8 bits each
      if constexpr (RES > 0 && GROUP == 1) {
        int nt = 0;
        for (int c = 0; c < pl.num_classes; ++c) {
          base_pack |= static_cast<uint32_t>(nt) << (8 * c);
          for (int i = 0; i < pl.ntaps[c]; ++i, ++nt) wtap_pack |= static_cast<uint64_t>(pl.taps[c][i].wtap) << (4 * nt);
        }
      }
      if constexpr (RES > 0) {
        ptx::mbar_wait(res_bar, 0);  // the resident weight tiles have landed
        ptx::tc_fence_after();
      }
      TileCoord tc;
      for (int it = 0; tile_at(p, it, BLOCK_N, tc); ++it) {
        const int acc = it & 1;
        ptx::mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        const int nouter = npass * (GROUP == 3 ? 3 : pl.ntaps[tc.cls]);
        uint32_t first = 0;
        for (int o = 0; o < nouter; ++o) {
          // resident weights: slot of (tap, chunk) = its k offset / 64
          int slot0 = 0;
          if constexpr (RES > 0)
            slot0 = (GROUP == 3 ? o : static_cast<int>((wtap_pack >> (4 * (((base_pack >> (8 * tc.cls)) & 255) + o))) & 15)) *
                    chunks;
          for (int ch = 0; ch < chunks; ++ch) {
            ptx::mbar_wait(&full_bar[stage], phase);
            ptx::tc_fence_after();
            const uint64_t a_desc = a_desc0 + static_cast<uint64_t>((stage * L::kAStage) >> 4);
            const uint64_t b_desc = RES > 0 ? b_desc0 + static_cast<uint64_t>((slot0 + ch) * kBTile16)
                                            : b_desc0 + static_cast<uint64_t>((stage * L::kBStageBytes) >> 4);
#pragma unroll
            for (int u = 0; u < GROUP; ++u) {
              // unit u of a row-halo stage: the input box u rows further down (2048 B); its weight tile is the
              // u-th of the stage, or (resident) tap kh = u of this filter column: 3 * chunks slots further
              const uint32_t b_u = RES > 0 ? u * 3 * chunks * kBTile16 : u * kBTile16;
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k)
                ptx::umma_bf16(d_tmem, a_desc + (u * (2048 >> 4) + 2 * k), b_desc + (b_u + 2 * k), idesc,
                               first | u | k);
            }
            first = 1;
            ptx::umma_commit(&empty_bar[stage]);
            if (++stage == STAGES) stage = 0, phase ^= 1;
          }
        }
        ptx::umma_commit(&tfull_bar[acc]);
      }
    }
  } else {
    // ===================== epilogue (warps 2..5, 128 threads) =====================
    const int et = threadIdx.x - 64;  // 0..127
    for (int c = et; c < kMaxCout; c += EG * kEpiThreads) {
      const bool ok = c < pl.cout;
      s_scale[c] = ok ? pl.scale[c] : 0.f;
      s_shift[c] = ok ? pl.shift[c] : 0.f;
    }
    ptx::named_bar_sync(3, EG * kEpiThreads);
    const int eg = EG == 2 ? (warp - 2) >> 2 : 0;  // this thread's epilogue group
    const int bar_id = 1 + eg;
    const bool lead_warp = ((warp - 2) & 3) == 0;  // its elected lane issues this group's TMA stores
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int lw = row % p.tw;
    const int lh = (row / p.tw) % p.th;
    const int li = row / (p.tw * p.th);
    const int planes = act_planes(pl.act);
    const bool f16 = act_is_f16(pl.act);
    int unit = 0;  // staging-buffer rotation counter (EG == 1)
    // train-mode BatchNorm statistics (pl.bn_sums): this thread's sum(z), sum(z^2) of channel 64 * k + (row & 63) over
    // rows [64 * (row >> 6), + 64) of all its tiles; fp32 over <= a few hundred 64-term partial sums, fp64 from there on
    float st1[STATS ? kMaxCout / 64 : 1], st2[STATS ? kMaxCout / 64 : 1];
#pragma unroll
    for (int k = 0; k < (STATS ? kMaxCout / 64 : 1); ++k) st1[k] = st2[k] = 0.f;
    TileCoord tc;
    for (int it = eg; tile_at(p, it, BLOCK_N, tc); it += EG) {
      const int acc = it & 1;  // == eg when EG == 2
      const int mw = tc.w0 + lw, mh = tc.h0 + lh, img = tc.i0 + li;
      const bool valid = mw < pl.wm && mh < pl.hm && img < pl.n_img;
      const int oh = mh * pl.out_s + pl.cls_oh[tc.cls];
      const int ow = mw * pl.out_s + pl.cls_ow[tc.cls];
      const size_t pix = (static_cast<size_t>(img) * pl.out_h + oh) * pl.out_w + ow;
      ptx::mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N;

      if constexpr (L::kNumStaging > 0) {
        if (p.tma_store) {
#pragma unroll 1
          for (int g = 0; g < BLOCK_N / 64; ++g) {
            const int cb = tc.n0 + g * 64;
            if (cb >= pl.cout) break;
            uint32_t r[64];
            ptx::tmem_ld_32x32b_x32(t_row + g * 64, r);
            ptx::tmem_ld_32x32b_x32(t_row + g * 64 + 32, r + 32);
            ptx::tmem_ld_wait();
            float v[64];
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] = fmaf(__uint_as_float(r[j]), s_scale[cb + j], s_shift[cb + j]);
            if (pl.residual && valid) {
              const __nv_bfloat16* rp = pl.residual + pix * pl.y_pix + pl.y_coffset + cb;
              for (int pln = 0; pln < planes; ++pln)
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                  const uint4 rv = *reinterpret_cast<const uint4*>(rp + pln * pl.y_cstride + c8 * 8);
                  const uint32_t* rb = reinterpret_cast<const uint32_t*>(&rv);
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float2 f = unpack_act2(rb[j], f16);
                    v[c8 * 8 + 2 * j] += f.x, v[c8 * 8 + 2 * j + 1] += f.y;
                  }
                }
            }
            if (pl.relu && planes == 2) {  // (one plane: the ReLU rides in the bf16 conversion below)
#pragma unroll
              for (int j = 0; j < 64; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            for (int pln = 0; pln < planes; ++pln, ++unit) {
              uint8_t* stg = smem + L::kStgOff + (EG == 2 ? eg : unit % L::kNumStaging) * kStagingBytes;
              // the TMA store that last used this staging tile must have finished reading it
              if (lead_warp && ptx::elect_one_sync())
                ptx::bulk_wait_group_read<EG == 2 ? 0 : L::kNumStaging - 1>();
              ptx::named_bar_sync(bar_id, kEpiThreads);
#pragma unroll
              for (int c8 = 0; c8 < 8; ++c8) {
                uint4 pk;
                uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
                if (planes == 1) {
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    pw[j] = ptx::pack_act2(v[c8 * 8 + 2 * j], v[c8 * 8 + 2 * j + 1], pl.relu != 0, f16);
                  // statistics are summed from this tile: rows outside the image (clipped by the TMA store) count as 0
                  if (STATS && !valid) pk = make_uint4(0u, 0u, 0u, 0u);
                } else {
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    uint32_t hi, lo;
                    split_act2(v[c8 * 8 + 2 * j], v[c8 * 8 + 2 * j + 1], f16, hi, lo);
                    pw[j] = pln == 0 ? hi : lo;
                  }
                }
                *reinterpret_cast<uint4*>(stg + row * 128 + ((c8 ^ (row & 7)) << 4)) = pk;
              }
              const int ocls = tc.cls;
              const int och = pl.y_coffset + cb + pln * pl.y_cstride;
              {
                ptx::fence_proxy_async();
                ptx::named_bar_sync(bar_id, kEpiThreads);
                if (lead_warp && ptx::elect_one_sync()) {
                  ptx::tma_store_4d(&p.y_map[ocls], stg, och, tc.w0, tc.h0, tc.i0);
                  ptx::bulk_commit_group();
                }
              }
              if constexpr (STATS) {
                // train-mode BatchNorm statistics of the values AS STORED: thread (channel c, row half) walks a column of
                // the staging tile (a warp reads 64 contiguous bytes per row: no bank conflicts) and keeps the two sums
                // in registers (selected by predicates, so the arrays stay in registers); they go to the fp64 totals
                // once, when the CTA retires. The tile is next written after a named barrier every thread only reaches
                // behind these reads.
                const int c = row & 63;
                const uint8_t* col = stg + (row >> 6) * (64 * 128) + (c & 7) * 2;
                float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
                for (int r = 0; r < 64; ++r) {
                  const float zv = elem_to_float(*reinterpret_cast<const __nv_bfloat16*>(col + r * 128 + (((c >> 3) ^ (r & 7)) << 4)), f16);
                  s1 += zv, s2 = fmaf(zv, zv, s2);
                }
                const int blk = cb >> 6;
#pragma unroll
                for (int k = 0; k < kMaxCout / 64; ++k)
                  if (blk == k) st1[k] += s1, st2[k] += s2;
              }
            }
          }
          ptx::tc_fence_before();
          ptx::mbar_arrive(&tempty_bar[acc]);
          continue;
        }
      }

      // ---- fp32 NCHW logits of a full 8x16 tile: [class][8][16] staging box, one TMA store (the per-thread stores
      //      of this layout are 64-byte segments, 11 per thread: ncu had the LSU/L1 pipe at 77 % on this layer)
      if constexpr (BLOCK_N == 16) {
        if (p.nchw_tma) {
          float* stg = reinterpret_cast<float*>(smem + L::kNchwOff + eg * L::kNchwBytes);
          uint32_t r[16];
          ptx::tmem_ld_32x32b_x16(t_row, r);
          ptx::tmem_ld_wait();
          if (lead_warp && ptx::elect_one_sync()) ptx::bulk_wait_group_read<0>();  // previous store has read the box
          ptx::named_bar_sync(bar_id, kEpiThreads);
          float best = -1.f;
          int arg = 0;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float v = fmaf(__uint_as_float(r[j]), s_scale[j], s_shift[j]);
            if (pl.relu) v = fmaxf(v, 0.f);
            if (j < pl.cout) {
              stg[j * kBlockM + row] = v;
              if (j == 0 || v > best) best = v, arg = j;
            }
          }
          if (pl.labels && valid)
            pl.labels[(static_cast<size_t>(img) * pl.out_h + oh) * pl.out_w + ow] = static_cast<uint8_t>(arg);
          ptx::fence_proxy_async();
          ptx::named_bar_sync(bar_id, kEpiThreads);
          if (lead_warp && ptx::elect_one_sync()) {
            ptx::tma_store_4d(&p.y_map[0], stg, tc.w0, tc.h0, 0, tc.i0);
            ptx::bulk_commit_group();
          }
          ptx::tc_fence_before();
          ptx::mbar_arrive(&tempty_bar[acc]);
          continue;
        }
      }

      // ---- direct-store epilogue: fp32 NCHW logits, or NHWC when cout is not a multiple of 64
      constexpr int kChunk = BLOCK_N < 32 ? 16 : 32;
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += kChunk) {
        uint32_t r[kChunk];
        if constexpr (kChunk == 32)
          ptx::tmem_ld_32x32b_x32(t_row + c0, r);
        else
          ptx::tmem_ld_32x32b_x16(t_row + c0, r);
        ptx::tmem_ld_wait();
        const int cb = tc.n0 + c0;
        if (cb >= pl.cout) break;
        float v[kChunk];
#pragma unroll
        for (int j = 0; j < kChunk; ++j)
          v[j] = fmaf(__uint_as_float(r[j]), s_scale[min(cb + j, kMaxCout - 1)], s_shift[min(cb + j, kMaxCout - 1)]);
        if (pl.out_fmt == W2C_OUT_NCHW_F32) {
          if (valid) {
            float* y = static_cast<float*>(pl.y);
            const size_t plane = static_cast<size_t>(pl.out_h) * pl.out_w;
            const size_t base = static_cast<size_t>(img) * pl.cout * plane + static_cast<size_t>(oh) * pl.out_w + ow;
            if (pl.relu) {
#pragma unroll
              for (int j = 0; j < kChunk; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (y) {
#pragma unroll
              for (int j = 0; j < kChunk; ++j)
                if (cb + j < pl.cout) y[base + (cb + j) * plane] = v[j];
            }
            if (pl.labels && cb == 0) {
              // every class of this pixel sits in this thread's registers (cout <= kChunk, checked on the host):
              // first maximal index, like torch.max(1)[1]
              float best = v[0];
              int arg = 0;
#pragma unroll
              for (int j = 1; j < kChunk; ++j)
                if (j < pl.cout && v[j] > best) best = v[j], arg = j;
              pl.labels[static_cast<size_t>(img) * plane + static_cast<size_t>(oh) * pl.out_w + ow] =
                  static_cast<uint8_t>(arg);
            }
          }
        } else if (valid) {
          __nv_bfloat16* ypix = static_cast<__nv_bfloat16*>(pl.y) + pix * pl.y_pix + pl.y_coffset + cb;
          const __nv_bfloat16* rpix = pl.residual ? pl.residual + pix * pl.y_pix + pl.y_coffset + cb : nullptr;
#pragma unroll
          for (int g = 0; g < kChunk / 8; ++g) {
            if (cb + g * 8 >= pl.cout) break;
            if (rpix) {
              for (int pln = 0; pln < planes; ++pln) {
                const uint4 rv = *reinterpret_cast<const uint4*>(rpix + pln * pl.y_cstride + g * 8);
                const uint32_t* rb = reinterpret_cast<const uint32_t*>(&rv);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 f = unpack_act2(rb[j], f16);
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
              split_act2(a, b, f16, hb[j], lb[j]);
            }
            *reinterpret_cast<uint4*>(ypix + g * 8) = hv;
            if (planes == 2) *reinterpret_cast<uint4*>(ypix + pl.y_cstride + g * 8) = lv;
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tempty_bar[acc]);
    }
    if constexpr (STATS) {
#pragma unroll
      for (int k = 0; k < kMaxCout / 64; ++k)
        if (k * 64 < pl.cout) {
          atomicAdd(&pl.bn_sums[k * 64 + (row & 63)], static_cast<double>(st1[k]));
          atomicAdd(&pl.bn_sums[pl.cout + k * 64 + (row & 63)], static_cast<double>(st2[k]));
        }
    }
    __syncwarp();
    if (lead_warp && ptx::elect_one_sync()) ptx::bulk_wait_group<0>();  // all TMA stores landed before the CTA retires
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
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}

template <int BLOCK_N, int STAGES, int GROUP = 1, int EG = 1, int RES = 0>
int launch_persv1(const PersParams& p, cudaStream_t stream) {
  using L = PersSmem<BLOCK_N, STAGES, GROUP, EG, RES>;
  static_assert(L::kDynamicBytes <= 232448, "shared memory budget exceeded");
  // the opt-in is per DEVICE: one process may drive several (nn.DataParallel replicas, model.to('cuda:1'))
  const int num_sms = device_sm_count();
  const int want = num_sms * (p.ctas_per_sm > 0 ? p.ctas_per_sm : 1);
  const int grid = p.groups < want ? p.groups : want;
  if constexpr (BLOCK_N >= 64) {
    if (p.plan.bn_sums) {
      static DeviceOnce attr_stats;
      int rc = attr_stats.ensure([] {
        return cudaFuncSetAttribute(conv_persv1_kernel<BLOCK_N, STAGES, GROUP, EG, RES, true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynamicBytes);
      }, "conv_persv1_kernel");
      if (rc) return rc;
      conv_persv1_kernel<BLOCK_N, STAGES, GROUP, EG, RES, true><<<grid, L::kThreads, L::kDynamicBytes, stream>>>(p);
      W2C_CHECK_LAUNCH("conv_persv1_kernel");
      return W2C_OK;
    }
  }
  static DeviceOnce attr_set;
  int rc = attr_set.ensure([] {
    return cudaFuncSetAttribute(conv_persv1_kernel<BLOCK_N, STAGES, GROUP, EG, RES, false>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynamicBytes);
  }, "conv_persv1_kernel");
  if (rc) return rc;
  conv_persv1_kernel<BLOCK_N, STAGES, GROUP, EG, RES, false><<<grid, L::kThreads, L::kDynamicBytes, stream>>>(p);
  W2C_CHECK_LAUNCH("conv_persv1_kernel");
  return W2C_OK;
}

}  // namespace

bool conv_persv1_supported(const ConvPlan& plan) { return plan.cout <= kMaxCout; }

// The statistics live in the TMA-store epilogue: NHWC output in whole 64-channel groups, one storage plane. (With
// block_n = 0 such a layer always gets BLOCK_N >= 64.)
bool conv_persv1_fuses_bn_sums(const w2c_conv_args& a, const ConvPlan& plan) {
  return conv_persv1_supported(plan) && plan.out_fmt == W2C_OUT_NHWC && plan.cout % 64 == 0 && act_planes(plan.act) == 1 &&
         (a.block_n == 0 || a.block_n >= 64);
}

int conv_persv1_forward(const w2c_conv_args& a, const ConvPlan& plan, cudaStream_t stream) {
  PersParams p;
  p.plan = plan;
  const int planes = act_planes(plan.act);
  W2C_CHECK_ARG(plan.cout <= kMaxCout, "conv_pers: cout=%d exceeds %d", plan.cout, kMaxCout);

  int tw = plan.wm >= 16 ? 16 : pow2_ceil(plan.wm);
  int th = pow2_ceil(plan.hm);
  if (th > kBlockM / tw) th = kBlockM / tw;
  const int tn = kBlockM / (tw * th);
  p.tw = tw, p.th = th, p.tn = tn;
  p.tiles_w = ceil_div(plan.wm, tw);
  p.tiles_h = ceil_div(plan.hm, th);
  p.tiles_img = ceil_div(plan.n_img, tn);
  p.m_tiles = p.tiles_w * p.tiles_h * p.tiles_img;

  int bn = a.block_n;
  if (bn == 0) {
    if (plan.cout_pad % 256 == 0)
      bn = 256;
    else if (plan.cout_pad % 128 == 0)
      bn = 128;
    else if (plan.cout_pad % 64 == 0)
      bn = 64;
    else if (plan.cout_pad % 32 == 0)
      bn = 32;
    else
      bn = 16;
    // fewer than two waves of 256-wide tiles (the 16x16 maps: 160 tiles on 148 SMs): 128-wide tiles balance better
    // (0.058 vs 0.067 ms; one-tile kernel 0.073 - profiles/r1_conv_sweep_v7_full.md)
    if (bn == 256 && static_cast<long long>(p.m_tiles) * plan.num_classes * (plan.cout_pad / bn) < 2 * 148)
      bn = 128;
    // not enough tiles to occupy the SMs at this width: narrower tiles
    while (bn > 64 && static_cast<long long>(p.m_tiles) * plan.num_classes * (plan.cout_pad / bn) < 148)
      bn /= 2;
  }
  W2C_CHECK_ARG(bn == 16 || bn == 32 || bn == 64 || bn == 128 || bn == 256, "conv: block_n=%d not supported", bn);
  W2C_CHECK_ARG(plan.cout_pad % bn == 0, "conv: block_n=%d does not divide cout_pad=%d", bn, plan.cout_pad);
  W2C_CHECK_ARG(!plan.labels || plan.cout <= (bn < 32 ? 16 : 32), "conv: label map needs cout=%d within one %d-wide tile",
                plan.cout, bn);
  p.n_tiles = plan.cout_pad / bn;
  p.total_tiles = p.m_tiles * p.n_tiles * plan.num_classes;
  p.cls_shift = (plan.num_classes == 4 && p.m_tiles * p.n_tiles >= 8 * 148 && !((a.impl >> 8) & 128)) ? 2 : 0;
  p.groups = p.cls_shift ? p.m_tiles * p.n_tiles : p.total_tiles;
  p.tma_store = (plan.out_fmt == W2C_OUT_NHWC && plan.cout % 64 == 0 && bn >= 64) ? 1 : 0;
  W2C_CHECK_ARG(!plan.bn_sums || (p.tma_store && planes == 1), "conv: bn_sums needs the TMA-store epilogue (w2c_conv_fuses_bn_sums)");
  // row-halo stages: 3x3 stride-1 convs on full 8x16 tiles, BLOCK_N <= 128 (at 256 three weight tiles do not fit)
  const bool row_halo = !((a.impl >> 8) & 1) && plan.num_classes == 1 && plan.in_s == 1 &&
                        plan.ntaps[0] == 9 && tn == 1 && tw == 16 && th == 8 && bn <= 128;

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
    if (row_halo) {  // a_map[1]: the same tensor with a box two rows taller (the three vertical taps of a column)
      const cuuint32_t box3[4] = {(cuuint32_t)kBlockK, (cuuint32_t)tw, (cuuint32_t)(th + 2), 1};
      rc = encode_map(&p.a_map[1], plan.x, 4, dims, str, box3, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
      if (rc) return rc;
    }
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
  if (p.tma_store) {
    // output maps: [128 px][64 ch] boxes of the NHWC output; a transposed conv writes one output-parity class
    // per tile through a map that strides two pixels in H and W
    const cuuint32_t ybox[4] = {64, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tn};
    const __nv_bfloat16* y = static_cast<const __nv_bfloat16*>(plan.y);
    const int n_out_maps = plan.num_classes;
    for (int cls = 0; cls < n_out_maps; ++cls) {
      const int s = plan.out_s;
      const cuuint64_t dims[4] = {(cuuint64_t)plan.y_pix, (cuuint64_t)plan.out_w / s, (cuuint64_t)plan.out_h / s,
                                  (cuuint64_t)plan.n_img};
      const cuuint64_t str[3] = {(cuuint64_t)s * plan.y_pix * esz, (cuuint64_t)s * plan.out_w * plan.y_pix * esz,
                                 (cuuint64_t)plan.out_h * plan.out_w * plan.y_pix * esz};
      const __nv_bfloat16* base = y + (static_cast<size_t>(plan.cls_oh[cls]) * plan.out_w + plan.cls_ow[cls]) * plan.y_pix;
      int rc = encode_map(&p.y_map[cls], base, 4, dims, str, ybox, CU_TENSOR_MAP_L2_PROMOTION_NONE);
      if (rc) return rc;
    }
    for (int cls = n_out_maps; cls < 4; ++cls) p.y_map[cls] = p.y_map[0];
  } else {
    for (int cls = 0; cls < 4; ++cls) p.y_map[cls] = p.b_map;  // unused, but keep the bytes defined
  }
  p.nchw_tma = 0;
  if (!((a.impl >> 8) & 256) && plan.out_fmt == W2C_OUT_NCHW_F32 && plan.y && bn == 16 && p.n_tiles == 1 && tn == 1 &&
      tw == 16 && th == 8 && plan.num_classes == 1 && plan.out_w % 4 == 0) {
    const cuuint64_t dims[4] = {(cuuint64_t)plan.out_w, (cuuint64_t)plan.out_h, (cuuint64_t)plan.cout,
                                (cuuint64_t)plan.n_img};
    const cuuint64_t str[3] = {(cuuint64_t)plan.out_w * 4, (cuuint64_t)plan.out_h * plan.out_w * 4,
                               (cuuint64_t)plan.cout * plan.out_h * plan.out_w * 4};
    const cuuint32_t ybox[4] = {(cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)plan.cout, 1};
    int rc = encode_map_f32(&p.y_map[0], plan.y, 4, dims, str, ybox);
    if (rc) return rc;
    p.nchw_tma = 1;
  }

  // two epilogue warpgroups unless disabled (impl flag bit 1 / env) - see PersSmem
  const bool eg2 = !((a.impl >> 8) & 2);
  int cps = ((a.impl >> 8) & 8) ? 3 : ((a.impl >> 8) & 4) ? 2 : 1;
  // default for the logits layer: two CTAs per SM (with the lean elect.sync issue loops 0.50 ms, against 0.53 with
  // three and 0.72 with one; before that change three were best - profiles/r1_conv_sweep_v6_elect.md)
  if (!((a.impl >> 8) & (4 | 8 | 16)) && row_halo && bn == 16) cps = 2;
  if (!((row_halo && bn == 16) || bn == 64)) cps = 1;
  if (bn == 64 && cps > 2) cps = 2;
  p.ctas_per_sm = cps;
  // resident weights (see PersSmem): single n-tile, one bf16 plane, every [bn][64] tile of the layer fits
  const int res_slots = plan.ktot / kBlockK;
  const bool res_ok = !((a.impl >> 8) & 32) && p.n_tiles == 1 && planes == 1;
  if (row_halo) {
    switch (bn) {
      case 128:
        if (res_ok && res_slots == 9) return launch_persv1<128, 3, 3, 1, 9>(p, stream);
        return launch_persv1<128, 3, 3, 1>(p, stream);
      case 64:
        if (res_ok && res_slots == 18 && ((a.impl >> 8) & 64)) return launch_persv1<64, 3, 3, 1, 18>(p, stream);
        if (cps >= 2) return launch_persv1<64, 2, 3, 1>(p, stream);
        return eg2 ? launch_persv1<64, 4, 3, 2>(p, stream) : launch_persv1<64, 4, 3, 1>(p, stream);
      case 32: return eg2 ? launch_persv1<32, 5, 3, 2>(p, stream) : launch_persv1<32, 5, 3, 1>(p, stream);
      default:
        // 11-channel logits layer: its MMAs (N = 16) are so short that ONE issuing thread per SM is the limit
        // (ncu: 9 % tensor active, ~1800 issue cycles per tile). Small stages -> three CTAs (three issuers) per SM.
        if (cps == 3 && res_ok && res_slots == 9 && ((a.impl >> 8) & 64)) return launch_persv1<16, 2, 3, 1, 9>(p, stream);
        if (cps == 3) return launch_persv1<16, 2, 3, 1>(p, stream);
        if (cps == 2) return launch_persv1<16, 3, 3, 1>(p, stream);
        return eg2 ? launch_persv1<16, 6, 3, 2>(p, stream) : launch_persv1<16, 6, 3, 1>(p, stream);
    }
  }
  switch (bn) {
    case 256:
      return launch_persv1<256, 4, 1, 1>(p, stream);
    case 128: return eg2 ? launch_persv1<128, 5, 1, 2>(p, stream) : launch_persv1<128, 5, 1, 1>(p, stream);
    case 64:
      if (cps >= 2) return launch_persv1<64, 3, 1, 1>(p, stream);
      // (resident weights measured 5-9 % SLOWER here - 64->64 s2 conv 0.469 vs 0.431 ms, deconv 0.636 vs 0.605: the
      // per-tap slot lookup sits in the single MMA-issuing thread; opt-in only, profiles/r1_conv_sweep_v5_ab.md)
      if (res_ok && res_slots == 9 && eg2 && !plan.kw_major && ((a.impl >> 8) & 64))
        return launch_persv1<64, 7, 1, 2, 9>(p, stream);
      return eg2 ? launch_persv1<64, 6, 1, 2>(p, stream) : launch_persv1<64, 6, 1, 1>(p, stream);
    case 32: return eg2 ? launch_persv1<32, 6, 1, 2>(p, stream) : launch_persv1<32, 6, 1, 1>(p, stream);
    default: return eg2 ? launch_persv1<16, 6, 1, 2>(p, stream) : launch_persv1<16, 6, 1, 1>(p, stream);
  }
}

}  // namespace w2c
