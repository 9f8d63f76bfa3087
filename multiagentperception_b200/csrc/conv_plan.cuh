// Geometry of one fused conv / transposed-conv launch, shared by the tcgen05 kernel and the SIMT cross-check.
//
// Every supported op is expressed as a GEMM over an "M-space" of pixels (hm x wm per image):
//   conv3x3 s1 / conv1x1 s1 : M-space = output pixels = input pixels
//   conv3x3 s2 / conv1x1 s2 : M-space = output pixels (h_in/2 x w_in/2); the input is addressed through four
//                             parity planes (row parity, column parity) so every tap is a unit-stride window
//   deconv3x3 s2 (p1, op1)  : M-space = INPUT pixels; the 2h x 2w output splits into four parity classes with
//                             1 / 2 / 2 / 4 contributing taps (oh = 2*ih - 1 + kh)
// and a K loop over (pass, tap, 64-channel chunk).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace w2c {

struct Tap {
  int8_t dw, dh;          // window shift, in pixels of the TMA map the tap reads (tcgen05 path)
  int8_t map;             // which input tensor map (parity plane) the tap reads
  int8_t wtap;            // tap index into the packed weight (kh*3 + kw)
  int8_t iw_off, ih_off;  // SIMT path: iw = mw*in_s + iw_off, ih = mh*in_s + ih_off
  int8_t pad0, pad1;
};

struct ConvPlan {
  const __nv_bfloat16* x;
  const __nv_bfloat16* w;
  const __nv_bfloat16* residual;
  void* y;
  uint8_t* labels;
  const float* scale;
  const float* shift;
  int n_img, hm, wm;     // M-space
  int in_h, in_w, in_s;  // input extent; input stride multiplier
  int out_h, out_w, out_s;
  int num_classes;
  int ntaps[4];
  int cls_oh[4], cls_ow[4];
  Tap taps[4][9];
  int cin, cout, cout_pad, ktot;
  int x_pix, x_cstride, x_coffset;  // x_pix = elements per input pixel (all planes)
  int y_pix, y_cstride, y_coffset;
  int act, relu, out_fmt;
  int npass;  // MMA passes over the operand planes: 1 (hi*hi) or 3 (hi*hi + hi*lo + lo*hi)
  // K order of the accumulation. 3x3 stride-1 convs run (filter column kw, 64-channel chunk, filter row kh) in EVERY
  // tensor-core kernel - the order the row-halo stages of the persistent kernel impose - so that the result does not
  // depend on which kernel / tile width the dispatch picks for a given batch (sharded == unsharded, scene i alone
  // == scene i inside a batch, bit for bit). Everything else runs (tap, chunk).
  int kw_major;
  double* bn_sums;  // fp64 [2 * cout] += sum(z), sum(z^2) of the stored output (train-mode BatchNorm), or NULL
};

int conv_tc_forward(const w2c_conv_args& a, const ConvPlan& plan, cudaStream_t stream);
int conv_persv1_forward(const w2c_conv_args& a, const ConvPlan& plan, cudaStream_t stream);
bool conv_persistent_preferred(const ConvPlan& plan);
bool conv_persv1_supported(const ConvPlan& plan);
bool conv_persv1_fuses_bn_sums(const w2c_conv_args& a, const ConvPlan& plan);
int conv_simt_forward(const ConvPlan& plan, cudaStream_t stream);

inline int build_conv_plan(const w2c_conv_args& a, ConvPlan& p) {
  W2C_CHECK_ARG(a.x && a.w && a.scale && a.shift && (a.y || a.labels), "conv: null pointer argument");
  W2C_CHECK_ARG(!a.labels || (a.out_fmt == W2C_OUT_NCHW_F32 && a.cout <= 32),
                "conv: a label map needs the fp32 NCHW logits layout and cout <= 32 (got out_fmt=%d cout=%d)",
                a.out_fmt, a.cout);
  W2C_CHECK_ARG(a.n > 0 && a.h_in > 0 && a.w_in > 0, "conv: bad image extent %dx%dx%d", a.n, a.h_in, a.w_in);
  W2C_CHECK_ARG(a.cin > 0 && a.cin % 64 == 0, "conv: cin=%d must be a positive multiple of 64", a.cin);
  W2C_CHECK_ARG(a.cout > 0, "conv: cout=%d", a.cout);
  W2C_CHECK_ARG(act_valid(a.act), "conv: bad act %d", a.act);
  W2C_CHECK_ARG(a.passes == 0 || a.passes == 1 || a.passes == 3, "conv: passes=%d (0 = default, 1 or 3)", a.passes);
  W2C_CHECK_ARG(a.out_fmt == W2C_OUT_NHWC || a.out_fmt == W2C_OUT_NCHW_F32, "conv: bad out_fmt %d", a.out_fmt);
  const int planes = act_planes(a.act);
  p = ConvPlan{};
  p.x = static_cast<const __nv_bfloat16*>(a.x);
  p.w = static_cast<const __nv_bfloat16*>(a.w);
  p.residual = static_cast<const __nv_bfloat16*>(a.residual);
  p.y = a.y;
  p.labels = a.labels;
  p.scale = a.scale;
  p.shift = a.shift;
  p.n_img = a.n;
  p.in_h = a.h_in;
  p.in_w = a.w_in;
  p.cin = a.cin;
  p.cout = a.cout;
  p.cout_pad = w2c_cout_pad(a.cout);
  p.x_cstride = a.x_cstride > 0 ? a.x_cstride : a.cin;
  p.x_coffset = a.x_coffset;
  p.x_pix = p.x_cstride * planes;
  p.y_cstride = a.y_cstride > 0 ? a.y_cstride : a.cout;
  p.y_coffset = a.y_coffset;
  p.y_pix = p.y_cstride * planes;
  p.act = a.act;
  p.npass = act_passes(a.act, a.passes);
  p.relu = a.relu;
  p.out_fmt = a.out_fmt;
  p.bn_sums = a.bn_sums;
  W2C_CHECK_ARG(p.x_coffset >= 0 && p.x_coffset + p.cin <= p.x_cstride, "conv: input channel slice out of range");
  W2C_CHECK_ARG(p.x_cstride % 8 == 0 && p.x_coffset % 8 == 0, "conv: input channel stride/offset must be /8");
  if (a.out_fmt == W2C_OUT_NHWC) {
    W2C_CHECK_ARG(p.y_coffset >= 0 && p.y_coffset + p.cout <= p.y_cstride, "conv: output slice out of range");
    W2C_CHECK_ARG(p.y_cstride % 8 == 0 && p.y_coffset % 8 == 0 && p.cout % 8 == 0,
                  "conv: NHWC output needs cout, stride and offset divisible by 8");
  } else {
    W2C_CHECK_ARG(a.residual == nullptr, "conv: residual is not supported with NCHW fp32 output");
  }

  auto tap = [](int dw, int dh, int map, int wtap, int iw_off, int ih_off) {
    Tap t{};
    t.dw = (int8_t)dw, t.dh = (int8_t)dh, t.map = (int8_t)map, t.wtap = (int8_t)wtap;
    t.iw_off = (int8_t)iw_off, t.ih_off = (int8_t)ih_off;
    return t;
  };
  p.num_classes = 1;
  p.in_s = 1;
  p.out_s = 1;
  switch (a.kind) {
    case W2C_CONV3X3_S1:
      p.hm = a.h_in, p.wm = a.w_in;
      p.ktot = 9 * a.cin;
      p.ntaps[0] = 9;
      p.kw_major = 1;
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) p.taps[0][kh * 3 + kw] = tap(kw - 1, kh - 1, 0, kh * 3 + kw, kw - 1, kh - 1);
      break;
    case W2C_CONV1X1_S1:
      p.hm = a.h_in, p.wm = a.w_in;
      p.ktot = a.cin;
      p.ntaps[0] = 1;
      p.taps[0][0] = tap(0, 0, 0, 0, 0, 0);
      break;
    case W2C_CONV3X3_S2:
    case W2C_CONV1X1_S2:
      W2C_CHECK_ARG(a.h_in % 2 == 0 && a.w_in % 2 == 0, "conv s2: h_in, w_in must be even (got %dx%d)", a.h_in,
                    a.w_in);
      p.hm = a.h_in / 2, p.wm = a.w_in / 2;
      p.in_s = 2;
      if (a.kind == W2C_CONV1X1_S2) {
        p.ktot = a.cin;
        p.ntaps[0] = 1;
        p.taps[0][0] = tap(0, 0, 0, 0, 0, 0);
      } else {
        p.ktot = 9 * a.cin;
        p.ntaps[0] = 9;
        // input row 2*mh + dr, dr in {-1,0,1}: dr = -1 -> odd plane, index mh-1; 0 -> even plane, mh; +1 -> odd, mh
        for (int kh = 0; kh < 3; ++kh)
          for (int kw = 0; kw < 3; ++kw) {
            const int dr = kh - 1, dc = kw - 1;
            const int ph = dr & 1, pw = dc & 1;
            p.taps[0][kh * 3 + kw] = tap(dc == -1 ? -1 : 0, dr == -1 ? -1 : 0, ph * 2 + pw, kh * 3 + kw, dc, dr);
          }
      }
      break;
    case W2C_DECONV3X3_S2: {
      p.hm = a.h_in, p.wm = a.w_in;
      p.out_s = 2;
      p.ktot = 9 * a.cin;
      p.num_classes = 4;
      // oh = 2*ih - 1 + kh.  Even oh = 2a: kh = 1, ih = a.  Odd oh = 2a+1: kh = 0 -> ih = a+1; kh = 2 -> ih = a.
      const int nk[2] = {1, 2};
      const int kk[2][2] = {{1, 0}, {0, 2}};
      const int dd[2][2] = {{0, 0}, {1, 0}};
      for (int ph = 0; ph < 2; ++ph)
        for (int pw = 0; pw < 2; ++pw) {
          const int cls = ph * 2 + pw;
          p.cls_oh[cls] = ph, p.cls_ow[cls] = pw;
          int t = 0;
          for (int i = 0; i < nk[ph]; ++i)
            for (int j = 0; j < nk[pw]; ++j)
              p.taps[cls][t++] = tap(dd[pw][j], dd[ph][i], 0, kk[ph][i] * 3 + kk[pw][j], dd[pw][j], dd[ph][i]);
          p.ntaps[cls] = t;
        }
      break;
    }
    default:
      return set_error(W2C_ERR_INVALID, "conv: unknown kind %d", a.kind);
  }
  p.out_h = p.hm * p.out_s;
  p.out_w = p.wm * p.out_s;
  return W2C_OK;
}

}  // namespace w2c
