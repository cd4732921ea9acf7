// Layer table, packed-weight layout and shared elementwise kernels of the VGG16 U-Net
// (VGG.py:13-203).  Used by both convolution back ends (CUDA-core fp32 and tcgen05).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace ha {

struct ConvSpec { int cin, cout, has_bias; };
// order of HaVggStateDict: VGG.py:23-29 (encoder), :31-56 (decoders), :62-81 (confidence heads)
static const ConvSpec kVggConvs[HA_VGG_N_CONV] = {
    {3, 64, 1},    {64, 64, 1},   {64, 128, 1},  {128, 128, 1}, {128, 256, 1}, {256, 256, 1}, {256, 256, 1},
    {384, 128, 0}, {128, 128, 0}, {192, 64, 0},  {64, 64, 0},   {128, 32, 0},  {32, 16, 0},
    {256, 1, 0},   {128, 1, 0},   {64, 1, 0},    {16, 1, 0}};
enum { L_CONV0 = 0, L_CONV2, L_CONV5, L_CONV7, L_CONV10, L_CONV12, L_CONV14, L_DEC1A, L_DEC1B, L_DEC2A, L_DEC2B,
       L_DEC3A, L_DEC3B, L_CONF0, L_CONF1, L_CONF2, L_CONF3 };

// Packed weights: per conv, three views are stored back to back (256-byte aligned):
//   f32  : [9][Cin][Cout]            fp32, Cout contiguous   (CUDA-core path, confidence heads)
//   hi/lo: [9][CoutPad][CinPad]      fp16, Cin contiguous (K-major B operand of tcgen05.mma);
//          w = hi + lo * 2^-11, lo stored pre-scaled by 2^11 so it keeps 11 significant bits
//   bias : [Cout] fp32 (zeros when the conv has none)
struct PackedConv { size_t f32, hi, lo, bias; int cin_pad, cout_pad; };
struct PackedLayout { PackedConv c[HA_VGG_N_CONV]; size_t total; };

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline PackedLayout vgg_packed_layout() {
  PackedLayout L;
  size_t off = 0;
  for (int i = 0; i < HA_VGG_N_CONV; ++i) {
    const ConvSpec& s = kVggConvs[i];
    PackedConv& p = L.c[i];
    p.cin_pad = (int)align_up(s.cin, 64);
    p.cout_pad = (int)align_up(s.cout, 16);
    p.f32 = off; off = align_up(off + (size_t)9 * s.cin * s.cout * 4, 256);
    p.hi = off; off = align_up(off + (size_t)9 * p.cout_pad * p.cin_pad * 2, 256);
    p.lo = off; off = align_up(off + (size_t)9 * p.cout_pad * p.cin_pad * 2, 256);
    p.bias = off; off = align_up(off + (size_t)s.cout * 4, 256);
  }
  L.total = off;
  return L;
}

// sequential 256-byte-aligned carve-up of the caller's workspace (dry = size query)
struct Arena {
  char* base; size_t off; size_t cap; bool dry;
  void* take(size_t bytes) {
    off = align_up(off, 256);
    void* p = dry ? nullptr : base + off;
    off += bytes;
    return p;
  }
};

struct TcOut {   // where the tensor-core conv epilogue writes (any subset)
  __half* act_full = nullptr; int af_pitch = 0, af_coff = 0;   // relu(v), same resolution
  __half* act_pool = nullptr; int ap_pitch = 0, ap_coff = 0;   // relu(maxpool2x2(v))
  __half* act_up = nullptr; int au_pitch = 0, au_coff = 0;     // relu(nearest x2 upsample of feat)
  float* feat = nullptr; int feat_pooled = 0;                  // raw fp32 v or maxpool2x2(v)
};

__global__ void pack_conv_kernel(const float* __restrict__ w, const float* __restrict__ bias, int cin, int cout,
                                 int cin_pad, int cout_pad, float* __restrict__ f32, __half* __restrict__ hi,
                                 __half* __restrict__ lo, float* __restrict__ bias_out);
__global__ void split_act_kernel(const float* __restrict__ in, __half* __restrict__ out, int C, size_t n_px);
int conv_simt(const float* in, int in_pitch, int in_coff, int cin, const float* w, const float* bias, float* out,
              int out_pitch, int out_coff, int cout, int B, int H, int W, int relu_out, cudaStream_t st);
int conv_tc(const __half* in, int in_pitch, int in_coff, int cin, const char* packed, const PackedConv& pc, int cout,
            bool has_bias, const TcOut& o, int B, int H, int W, bool split, cudaStream_t st, int n_taps = 9, bool pair = true);
// Activations the backward pass needs, where vgg_forward_tc laid them out in the caller's workspace (train mode)
struct TcSaved {
  __half *a1, *cat3, *cat2, *a5, *cat1, *a10, *a12, *d1, *d2, *d3;   // post-ReLU inputs of the convolutions, hi/lo planes
  float *x2, *x7, *x14;        // raw conv outputs in front of the three max-pools (the argmax of the unpooling), fp32 NHWC
};
TcSaved vgg_tc_carve(Arena& ar, int B, int H, int W, int n_levels, bool train);
TcSaved vgg_train_saved(char* ws, int B, int H, int W, int n_levels);
// `saved` != nullptr = train mode: conv2 / conv7 / conv14 also keep their raw outputs and x15 is pooled by a kernel of its own
// `g2s`: the decoder schedule of VGGUnet_G2S (VGG.py:275-345) on the folded [2H, W/2] maps
int vgg_forward_tc(const char* packed, const PackedLayout& L, const float* img, int B, int H, int W, int n_levels,
                   int precision, float* const* out_feat, Arena& ar, cudaStream_t st, TcSaved* saved = nullptr, bool g2s = false);

constexpr float kLoScale = 2048.f;          // 2^11
constexpr float kLoInvScale = 1.f / 2048.f;

}  // namespace ha
