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

constexpr float kLoScale = 2048.f;          // 2^11
constexpr float kLoInvScale = 1.f / 2048.f;

}  // namespace ha
