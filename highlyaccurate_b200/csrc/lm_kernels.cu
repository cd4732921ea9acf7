// Fused Levenberg-Marquardt step for satellite->ground pose refinement (sm_100a).
//
// One launch per (iteration, level) does, for every sample of the batch, what the reference
// does with ~60 eager launches and ~300 B of traffic per feature element:
//   geometry  : models_kitti.py:700-801 (grd2cam2world2sat) / models_ford.py:173-264
//   sampler   : jacobian.py:138-205 (4-tap bilinear gather + d/dx, d/dy, chained with duv/dpose)
//   masking   : models_kitti.py:927-929, :1191-1199 (geometric mask, bottom half only)
//   LM update : models_kitti.py:939-1041 / models_ford.py:380-466
// Nothing is materialised: each ground pixel's feature vector is read once (128-bit loads,
// NHWC), the four satellite taps come through L1/L2, and the per-sample normal equations are
// reduced in registers -> shuffles -> shared memory -> a deterministic two-stage cross-CTA
// combine whose last-arriving CTA solves the damped system and updates the pose in place.
//
// Algebra: with s_c, a_c = ds_c/dx, b_c = ds_c/dy per channel and the
// channel-independent 2x3 matrix D_p = d(u,v)/d(pose), the kernel accumulates
//   JtJ = sum_p w_p D_p^T [[Saa,Sab],[Sab,Sbb]] D_p,  Jts = sum_p w_p D_p^T [Ssa,Ssb],
//   Jtg = sum_p w_p D_p^T [Sga,Sgb],  SS = sum s^2, GG = sum g^2, SG = sum s g
// and finalises  H = JtJ/ns^2,  grad = Jts/ns^2 - Jtg/(ns*ng)  with ns = max(|s|,1e-6),
// ng = max(|g|,1e-6)  — identical to normalising s, J and g first (models_kitti.py:982-992).
#include <math.h>

#include "common.cuh"

namespace ha {

constexpr int kLmThreads = 256;
constexpr int kLmWarps = kLmThreads / 32;
constexpr int kLmAcc = 16;              // per-sample reduced scalars
constexpr int kLmMaxCtasPerSample = 256;

struct LmStepArgs {
  const float* sat;        // [B][A][A][C]
  const float* grd;        // [B][H][W][C]
  const float* sat_scale;  // [B] or null
  const float* grd_scale;  // [B] or null
  const float* conf;       // [B][H][W] or null
  const float4* table;     // [H][W] (x,y,z,mask)
  const float* extr;       // [B][12] or null
  float* pose;             // [B][3]
  const float* reset_uv;   // [2][B] or null
  float* stats;            // [B][HA_STATS] or null
  float* traj;             // &traj[0][it][lv][0] or null
  uint32_t* status;
  double* partial;         // [B][kLmMaxCtasPerSample][kLmAcc]
  uint32_t* ticket;        // [B]
  int traj_stride;         // floats between consecutive samples in traj
  int B, A, H, W;
  int px_per_cta;          // bottom-half pixels handled by one CTA
  int dof, using_weight, use_hessian;
  float rot, lat, lon;     // rotation_range (deg), shift_range_lat / lon (m)
  float mpp, inv_mpp, center;  // satellite metres per pixel, fp32(1/mpp), A/2
  float damping[3];
};

// Per-sample constants of the warp, evaluated in the reference's fp32 operation order.
struct KittiPose {
  float c, s, Tx, Tz;                 // R(theta) and T = -R*T0
  float jux, juy, jvx, jvy;           // d(u,v)/dsu, d(u,v)/dsv  (pixel independent)
  float kms, kmc, kc, tdx, tdz;       // dR entries and -dR*T0
  float inv_mpp;
};

__device__ __forceinline__ KittiPose kitti_pose(const LmStepArgs& a, float su, float sv, float th) {
  KittiPose k;
  const float pi_f = 3.14159265358979323846f;
  float heading = __fmul_rn(__fdiv_rn(__fmul_rn(th, a.rot), 180.f), pi_f);   // models_kitti.py:719
  float shu = __fmul_rn(su, a.lon);                                            // :720
  float shv = __fmul_rn(sv, a.lat);                                            // :721
  sincosf(heading, &k.s, &k.c);
  // T = sum(-R * T0), T0 = (shv, h, -shu)   (:734-737)
  k.Tx = __fadd_rn(__fmul_rn(-k.c, shv), __fmul_rn(k.s, -shu));
  k.Tz = __fadd_rn(__fmul_rn(-k.s, shv), __fmul_rn(-k.c, -shu));
  float kk = (float)((double)a.rot / 180.0 * 3.14159265358979323846);          // python double -> fp32 scalar
  k.kms = __fmul_rn(kk, -k.s);
  k.kmc = __fmul_rn(kk, -k.c);
  k.kc = __fmul_rn(kk, k.c);
  // -dR * T0 : x row (-kms, 0, -kmc), z row (-kc, 0, -kms)
  k.tdx = __fadd_rn(__fmul_rn(-k.kms, shv), __fmul_rn(-k.kmc, -shu));
  k.tdz = __fadd_rn(__fmul_rn(-k.kc, shv), __fmul_rn(-k.kms, -shu));
  k.inv_mpp = a.inv_mpp;
  // d xyz / d su = -R * (0,0,-lon) -> (x: -s*lon, z: c*lon); uv = (z, x)/mpp
  k.jux = __fmul_rn(k.inv_mpp, __fmul_rn(-k.c, -a.lon));
  k.juy = __fmul_rn(k.inv_mpp, __fmul_rn(k.s, -a.lon));
  // d xyz / d sv = -R * (lat,0,0) -> (x: -c*lat, z: -s*lat)
  k.jvx = __fmul_rn(k.inv_mpp, __fmul_rn(-k.s, a.lat));
  k.jvy = __fmul_rn(k.inv_mpp, __fmul_rn(-k.c, a.lat));
  return k;
}

struct FordPose {
  float c, s, um, vm;
  float R[9], T[3];
  float jux, juy, jvx, jvy;
  float kms, kmc, kc;
};

__device__ __forceinline__ FordPose ford_pose(const LmStepArgs& a, int b, float su, float sv, float th) {
  FordPose f;
  const float pi_f = 3.14159265358979323846f;
#pragma unroll
  for (int i = 0; i < 9; ++i) f.R[i] = a.extr[b * 12 + i];
#pragma unroll
  for (int i = 0; i < 3; ++i) f.T[i] = a.extr[b * 12 + 9 + i];
  f.um = __fmul_rn(a.lat, su);                                                 // models_ford.py:212
  f.vm = __fmul_rn(a.lon, sv);                                                 // :213
  float yaw = __fmul_rn(__fdiv_rn(__fmul_rn(th, a.rot), 180.f), pi_f);         // :216
  sincosf(yaw, &f.s, &f.c);
  float kk = (float)((double)a.rot / 180.0 * 3.14159265358979323846);
  f.kms = __fmul_rn(kk, -f.s);
  f.kmc = __fmul_rn(kk, -f.c);
  f.kc = __fmul_rn(kk, f.c);
  // dXw/dsu = Rw * lat*(0,-1,0) = (s*(-lat), c*(-lat)); Xs = (y, -x); then / mpp   (:234-253)
  float dxu = __fmul_rn(f.s, -a.lat), dyu = __fmul_rn(f.c, -a.lat);
  float dxv = __fmul_rn(f.c, a.lon), dyv = __fmul_rn(-f.s, a.lon);
  f.jux = __fdiv_rn(dyu, a.mpp);
  f.juy = __fdiv_rn(-dxu, a.mpp);
  f.jvx = __fdiv_rn(dyv, a.mpp);
  f.jvy = __fdiv_rn(-dxv, a.mpp);
  return f;
}

struct PixelWarp {
  float u, v;        // satellite pixel coordinates (x = column, y = row)
  float jtx, jty;    // d(u,v)/dtheta
};

__device__ __forceinline__ PixelWarp warp_kitti(const KittiPose& k, const LmStepArgs& a, float4 p) {
  PixelWarp w;
  // xyz = R p + T  (:754), uv = (z, x)/mpp + A/2  (:763-765)
  float x = __fadd_rn(__fadd_rn(__fmul_rn(k.c, p.x), __fmul_rn(-k.s, p.z)), k.Tx);
  float z = __fadd_rn(__fadd_rn(__fmul_rn(k.s, p.x), __fmul_rn(k.c, p.z)), k.Tz);
  w.u = __fadd_rn(__fdiv_rn(z, a.mpp), a.center);
  w.v = __fadd_rn(__fdiv_rn(x, a.mpp), a.center);
  // d xyz / d theta = dR p - dR T0  (:786-790)
  float dx = __fadd_rn(__fadd_rn(__fmul_rn(k.kms, p.x), __fmul_rn(k.kmc, p.z)), k.tdx);
  float dz = __fadd_rn(__fadd_rn(__fmul_rn(k.kc, p.x), __fmul_rn(k.kms, p.z)), k.tdz);
  w.jtx = __fmul_rn(k.inv_mpp, dz);
  w.jty = __fmul_rn(k.inv_mpp, dx);
  return w;
}

__device__ __forceinline__ PixelWarp warp_ford(const FordPose& f, const LmStepArgs& a, float4 p) {
  PixelWarp w;
  // Xb = R_FL Xc + T_FL  (models_ford.py:209)
  float xb = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(f.R[0], p.x), __fmul_rn(f.R[1], p.y)), __fmul_rn(f.R[2], p.z)), f.T[0]);
  float yb = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(f.R[3], p.x), __fmul_rn(f.R[4], p.y)), __fmul_rn(f.R[5], p.z)), f.T[1]);
  float X = __fadd_rn(xb, f.vm);
  float Y = __fadd_rn(yb, -f.um);
  float xw = __fadd_rn(__fmul_rn(f.c, X), __fmul_rn(f.s, Y));      // Xw = Rw (Xb + Tw)  (:223)
  float yw = __fadd_rn(__fmul_rn(-f.s, X), __fmul_rn(f.c, Y));
  w.u = __fadd_rn(__fdiv_rn(yw, a.mpp), a.center);                  // Xs = (yw, -xw)  (:226-231)
  w.v = __fadd_rn(__fdiv_rn(-xw, a.mpp), a.center);
  float dxw = __fadd_rn(__fmul_rn(f.kms, X), __fmul_rn(f.kc, Y));   // dRw (Xb + Tw)  (:240-246)
  float dyw = __fadd_rn(__fmul_rn(f.kmc, X), __fmul_rn(f.kms, Y));
  w.jtx = __fdiv_rn(dyw, a.mpp);
  w.jty = __fdiv_rn(-dxw, a.mpp);
  return w;
}

// Bilinear taps of jacobian.py:147-193 (clamped corners, inclusive range mask).
struct Taps {
  int o_nw, o_ne, o_sw, o_se;          // texel offsets (in pixels) into the sample's sat map
  float w_nw, w_ne, w_sw, w_se;        // value weights
  float ax_n, ax_s;                    // d/dx weights: -ax_n*nw + ax_n*ne - ax_s*sw + ax_s*se
  float ay_w, ay_e;                    // d/dy weights: -ay_w*nw - ay_e*ne + ay_w*sw + ay_e*se
  bool inr;
};

__device__ __forceinline__ Taps make_taps(float x, float y, int A) {
  Taps t;
  const float hi = (float)(A - 1);
  t.inr = (x >= 0.f) && (x <= hi) && (y >= 0.f) && (y <= hi);
  float x0 = floorf(x), y0 = floorf(y);
  float xw = fminf(fmaxf(x0, 0.f), hi), xe = fminf(fmaxf(x0 + 1.f, 0.f), hi);
  float yn = fminf(fmaxf(y0, 0.f), hi), ys = fminf(fmaxf(y0 + 1.f, 0.f), hi);
  float ex = xe - x, wx = x - xw, sy = ys - y, ny = y - yn;
  t.w_nw = ex * sy; t.w_ne = wx * sy; t.w_sw = ex * ny; t.w_se = wx * ny;
  t.ax_n = sy; t.ax_s = ny; t.ay_w = ex; t.ay_e = wx;
  int ixw = (int)xw, ixe = (int)xe, iyn = (int)yn, iys = (int)ys;
  t.o_nw = iyn * A + ixw; t.o_ne = iyn * A + ixe; t.o_sw = iys * A + ixw; t.o_se = iys * A + ixe;
  return t;
}

template <int GEOM, int C>
__global__ void __launch_bounds__(kLmThreads) lm_step_kernel(const LmStepArgs a) {
  constexpr int LPP = (C / 4 >= 32) ? 32 : C / 4;  // lanes per pixel
  constexpr int V = C / (4 * LPP);                 // float4 per lane per pixel
  constexpr int PPW = 32 / LPP;                    // pixels per warp per iteration
  static_assert(C % 4 == 0 && V >= 1 && LPP * V * 4 == C, "channel count");

  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LPP, cl = lane % LPP;     // pixel slot within the warp, channel lane
  const int P = (a.H - a.H / 2) * a.W;             // bottom-half pixels (models_kitti.py:1195-1199)
  const int q_begin = blockIdx.x * a.px_per_cta;
  const int q_end = min(P, q_begin + a.px_per_cta);

  const float su = a.pose[b * 3 + 0], sv = a.pose[b * 3 + 1], th = a.pose[b * 3 + 2];
  KittiPose kp;
  FordPose fp;
  float jux, juy, jvx, jvy;
  if (GEOM == HA_GEOM_KITTI) { kp = kitti_pose(a, su, sv, th); jux = kp.jux; juy = kp.juy; jvx = kp.jvx; jvy = kp.jvy; }
  else { fp = ford_pose(a, b, su, sv, th); jux = fp.jux; juy = fp.juy; jvx = fp.jvx; jvy = fp.jvy; }

  const size_t px_base = (size_t)b * a.H * a.W + (size_t)(a.H / 2) * a.W;   // first bottom-half pixel of sample b
  const float4* grd = reinterpret_cast<const float4*>(a.grd) + px_base * (C / 4);
  const float4* sat = reinterpret_cast<const float4*>(a.sat) + (size_t)b * a.A * a.A * (C / 4);
  const float4* tab = a.table + (size_t)(a.H / 2) * a.W;
  const float* conf = a.conf ? a.conf + px_base : nullptr;

  float h00 = 0, h01 = 0, h02 = 0, h11 = 0, h12 = 0, h22 = 0;
  float bs0 = 0, bs1 = 0, bs2 = 0, bg0 = 0, bg1 = 0, bg2 = 0;
  float SS = 0, GG = 0, SG = 0, cnt = 0;

  for (int q = q_begin + warp * PPW + sub; q < q_end; q += kLmWarps * PPW) {
    const float4 p = __ldg(tab + q);
    if (p.w == 0.f) continue;                         // geometric mask: s, J and g all vanish
    float4 g[V];
#pragma unroll
    for (int i = 0; i < V; ++i) g[i] = ldg_nc_stream(grd + (size_t)q * (C / 4) + cl + i * LPP);
    PixelWarp w = (GEOM == HA_GEOM_KITTI) ? warp_kitti(kp, a, p) : warp_ford(fp, a, p);
    Taps t = make_taps(w.u, w.v, a.A);
    float gg = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) gg += g[i].x * g[i].x + g[i].y * g[i].y + g[i].z * g[i].z + g[i].w * g[i].w;
    GG += gg;
    if (!t.inr) continue;                             // sampler mask: s = 0, J = 0, r = -g~
    float saa = 0, sab = 0, sbb = 0, ssa = 0, ssb = 0, sga = 0, sgb = 0, ss = 0, sg = 0;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int co = cl + i * LPP;
      const float4 nw = ldg_nc(sat + (size_t)t.o_nw * (C / 4) + co);
      const float4 ne = ldg_nc(sat + (size_t)t.o_ne * (C / 4) + co);
      const float4 sw = ldg_nc(sat + (size_t)t.o_sw * (C / 4) + co);
      const float4 se = ldg_nc(sat + (size_t)t.o_se * (C / 4) + co);
#define HA_CH(f)                                                                            \
      {                                                                                     \
        float s_ = nw.f * t.w_nw + ne.f * t.w_ne + sw.f * t.w_sw + se.f * t.w_se;           \
        float a_ = (ne.f - nw.f) * t.ax_n + (se.f - sw.f) * t.ax_s;                         \
        float b_ = (sw.f - nw.f) * t.ay_w + (se.f - ne.f) * t.ay_e;                         \
        float g_ = g[i].f;                                                                  \
        saa += a_ * a_; sab += a_ * b_; sbb += b_ * b_;                                     \
        ssa += s_ * a_; ssb += s_ * b_; sga += g_ * a_; sgb += g_ * b_;                     \
        ss += s_ * s_; sg += s_ * g_;                                                       \
      }
      HA_CH(x) HA_CH(y) HA_CH(z) HA_CH(w)
#undef HA_CH
    }
    const float om = (a.using_weight && conf) ? __ldg(conf + q) : 1.f;   // models_kitti.py:994-998
    // D = [ (jux,juy); (jvx,jvy); (jtx,jty) ]; J_k = a*D_kx + b*D_ky
    const float d0x = jux, d0y = juy, d1x = jvx, d1y = jvy, d2x = w.jtx, d2y = w.jty;
    const float e0x = saa * d0x + sab * d0y, e0y = sab * d0x + sbb * d0y;   // G * D_0
    const float e1x = saa * d1x + sab * d1y, e1y = sab * d1x + sbb * d1y;
    const float e2x = saa * d2x + sab * d2y, e2y = sab * d2x + sbb * d2y;
    h00 += om * (d0x * e0x + d0y * e0y);
    h01 += om * (d0x * e1x + d0y * e1y);
    h02 += om * (d0x * e2x + d0y * e2y);
    h11 += om * (d1x * e1x + d1y * e1y);
    h12 += om * (d1x * e2x + d1y * e2y);
    h22 += om * (d2x * e2x + d2y * e2y);
    bs0 += om * (ssa * d0x + ssb * d0y);
    bs1 += om * (ssa * d1x + ssb * d1y);
    bs2 += om * (ssa * d2x + ssb * d2y);
    bg0 += om * (sga * d0x + sgb * d0y);
    bg1 += om * (sga * d1x + sgb * d1y);
    bg2 += om * (sga * d2x + sgb * d2y);
    SS += ss; SG += sg;
    if (cl == 0) cnt += 1.f;
  }

  // ---- CTA reduction: lanes -> warp (fp64 shuffles) -> shared -> one partial row per CTA
  __shared__ double red[kLmWarps][kLmAcc];
  __shared__ bool is_last;
  {
    double v[kLmAcc] = {h00, h01, h02, h11, h12, h22, bs0, bs1, bs2, bg0, bg1, bg2, SS, GG, SG, cnt};
#pragma unroll
    for (int i = 0; i < kLmAcc; ++i) {
      double r = warp_sum(v[i]);
      if (lane == 0) red[warp][i] = r;
    }
  }
  __syncthreads();
  double* part = a.partial + ((size_t)b * kLmMaxCtasPerSample + blockIdx.x) * kLmAcc;
  if (threadIdx.x < kLmAcc) {
    double r = 0;
#pragma unroll
    for (int w = 0; w < kLmWarps; ++w) r += red[w][threadIdx.x];
    part[threadIdx.x] = r;
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t prev = atomicAdd(a.ticket + b, 1u);
    is_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;

  // ---- last CTA of this sample: ordered combine of the partials, damped solve, pose update
  __threadfence();
  __shared__ double tot[kLmAcc];
  if (threadIdx.x < kLmAcc) {
    const volatile double* pp = a.partial + (size_t)b * kLmMaxCtasPerSample * kLmAcc + threadIdx.x;
    double r = 0;
    for (unsigned c = 0; c < gridDim.x; ++c) r += pp[(size_t)c * kLmAcc];
    tot[threadIdx.x] = r;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  a.ticket[b] = 0;   // ready for the next step on this stream

  const double alpha = a.sat_scale ? (double)a.sat_scale[b] : 1.0;
  const double beta = a.grd_scale ? (double)a.grd_scale[b] : 1.0;
  const double ns = fmax(alpha * sqrt(tot[12]), 1e-6);     // models_kitti.py:982-984
  const double ng = fmax(beta * sqrt(tot[13]), 1e-6);      // :987-988
  const double fs = alpha * alpha / (ns * ns), fg = alpha * beta / (ns * ng);
  double Hm[3][3] = {{tot[0] * fs, tot[1] * fs, tot[2] * fs},
                     {tot[1] * fs, tot[3] * fs, tot[4] * fs},
                     {tot[2] * fs, tot[4] * fs, tot[5] * fs}};
  double gr[3] = {tot[6] * fs - tot[9] * fg, tot[7] * fs - tot[10] * fg, tot[8] * fs - tot[11] * fg};
  const double res_sq = alpha * alpha * tot[12] / (ns * ns) + beta * beta * tot[13] / (ng * ng) -
                        2.0 * alpha * beta * tot[14] / (ns * ng);

  // DOF selection (models_kitti.py:954-957): 3 -> (0,1,2), 2 -> (0,1), 1 -> (2)
  const int n = a.dof;
  const int i0 = (n == 1) ? 2 : 0;
  double Am[3][3], rhs[3], delta[3] = {0, 0, 0};
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) Am[i][j] = Hm[i0 + i][i0 + j];
    const double lam = (double)a.damping[i];
    Am[i][i] += a.use_hessian ? lam * Hm[i0 + i][i0 + i] : lam;    // :1005-1012 (column-wise lambda on a diagonal)
    rhs[i] = gr[i0 + i];
  }
  if (n == 1) {
    delta[0] = -rhs[0] / Am[0][0];
  } else if (n == 2) {
    const double det = Am[0][0] * Am[1][1] - Am[0][1] * Am[1][0];
    delta[0] = -(Am[1][1] * rhs[0] - Am[0][1] * rhs[1]) / det;
    delta[1] = -(-Am[1][0] * rhs[0] + Am[0][0] * rhs[1]) / det;
  } else {
    const double c00 = Am[1][1] * Am[2][2] - Am[1][2] * Am[2][1];
    const double c01 = Am[1][2] * Am[2][0] - Am[1][0] * Am[2][2];
    const double c02 = Am[1][0] * Am[2][1] - Am[1][1] * Am[2][0];
    const double det = Am[0][0] * c00 + Am[0][1] * c01 + Am[0][2] * c02;
    const double c10 = Am[0][2] * Am[2][1] - Am[0][1] * Am[2][2];
    const double c11 = Am[0][0] * Am[2][2] - Am[0][2] * Am[2][0];
    const double c12 = Am[0][1] * Am[2][0] - Am[0][0] * Am[2][1];
    const double c20 = Am[0][1] * Am[1][2] - Am[0][2] * Am[1][1];
    const double c21 = Am[0][2] * Am[1][0] - Am[0][0] * Am[1][2];
    const double c22 = Am[0][0] * Am[1][1] - Am[0][1] * Am[1][0];
    // inverse = adj / det, adj[i][j] = cofactor[j][i]
    delta[0] = -(c00 * rhs[0] + c10 * rhs[1] + c20 * rhs[2]) / det;
    delta[1] = -(c01 * rhs[0] + c11 * rhs[1] + c21 * rhs[2]) / det;
    delta[2] = -(c02 * rhs[0] + c12 * rhs[1] + c22 * rhs[2]) / det;
  }
  float nsu = su, nsv = sv, nth = th;
  uint32_t st = 0;
  if (n == 3) {
    nsu = su + (float)delta[0]; nsv = sv + (float)delta[1]; nth = th + (float)delta[2];
    // models_kitti.py:1028-1033: shifts outside (-2.5, 2.5) (or NaN) are re-drawn
    if (!(nsu > -2.5f && nsu < 2.5f)) { nsu = a.reset_uv[b]; st |= HA_STATUS_RESET; }
    if (!(nsv > -2.5f && nsv < 2.5f)) { nsv = a.reset_uv[a.B + b]; st |= HA_STATUS_RESET; }
  } else if (n == 2) {
    nsu = su + (float)delta[0]; nsv = sv + (float)delta[1];
  } else {
    nth = th + (float)delta[0];
  }
  if (isnan(nsu) || isnan(nsv) || isnan(nth)) st |= HA_STATUS_NAN_POSE;
  if (tot[15] == 0.0) st |= HA_STATUS_NO_INRANGE;
  if (st) atomicOr(a.status, st);
  a.pose[b * 3 + 0] = nsu; a.pose[b * 3 + 1] = nsv; a.pose[b * 3 + 2] = nth;
  if (a.traj) {
    float* tr = a.traj + (size_t)b * a.traj_stride;
    tr[0] = nsu; tr[1] = nsv; tr[2] = nth;
  }
  if (a.stats) {
    float* s = a.stats + (size_t)b * HA_STATS;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) s[HA_STAT_H + i * 3 + j] = (float)Hm[i][j];
    for (int i = 0; i < 3; ++i) s[HA_STAT_GRAD + i] = (float)gr[i];
    s[HA_STAT_SAT_NORM] = (float)ns; s[HA_STAT_GRD_NORM] = (float)ng; s[HA_STAT_RES_SQ] = (float)res_sq;
    for (int i = 0; i < 3; ++i) s[HA_STAT_DELTA + i] = (i < n) ? (float)delta[i] : 0.f;
    s[HA_STAT_N_INRANGE] = (float)tot[15];
    for (int i = HA_STAT_N_INRANGE + 1; i < HA_STATS; ++i) s[i] = 0.f;
  }
}

template <int GEOM>
static int launch_by_channels(int C, dim3 grid, cudaStream_t st, const LmStepArgs& a) {
  switch (C) {
    case 256: lm_step_kernel<GEOM, 256><<<grid, kLmThreads, 0, st>>>(a); break;
    case 128: lm_step_kernel<GEOM, 128><<<grid, kLmThreads, 0, st>>>(a); break;
    case 64: lm_step_kernel<GEOM, 64><<<grid, kLmThreads, 0, st>>>(a); break;
    case 32: lm_step_kernel<GEOM, 32><<<grid, kLmThreads, 0, st>>>(a); break;
    case 16: lm_step_kernel<GEOM, 16><<<grid, kLmThreads, 0, st>>>(a); break;
    default: return HA_EINVAL;
  }
  count_launches(1);
  return check_launch("lm_step_kernel");
}

static size_t lm_ws_bytes(int B) {
  size_t part = (size_t)B * kLmMaxCtasPerSample * kLmAcc * sizeof(double);
  size_t tick = ((size_t)B * sizeof(uint32_t) + 255) / 256 * 256;
  return part + tick;
}

// One CTA per px_per_cta bottom-half pixels; enough CTAs to fill 148 SMs several times over
// while keeping at least a few pixels per warp.
static int choose_ctas_per_sample(int B, int P, int C) {
  const int lpp = (C / 4 >= 32) ? 32 : C / 4;
  const int ppw = 32 / lpp;
  const int min_px = kLmWarps * ppw * 4;                       // >= 4 iterations per warp
  int want = (kNumSMs * 8 + B - 1) / B;                        // ~8 CTAs per SM over the batch
  int max_by_px = (P + min_px - 1) / min_px;
  int n = want < max_by_px ? want : max_by_px;
  if (n < 1) n = 1;
  if (n > kLmMaxCtasPerSample) n = kLmMaxCtasPerSample;
  return n;
}

static int lm_step_impl(const HaLmParams* p, int level, const HaLevel* sat, const HaLevel* grd, const float* grd_conf,
                        const float* ground_table, const float* extrinsics, float* pose, const float* reset_uv,
                        float* stats, float* traj_step, int traj_stride, uint32_t* status, void* ws, size_t ws_bytes,
                        int B, cudaStream_t st) {
  if (!p || !sat || !grd || !pose || !status || !ws || !ground_table) return HA_EINVAL;
  if (level < 0 || level >= HA_MAX_LEVELS) return HA_EINVAL;
  if (sat->C != grd->C || sat->H != sat->W || (grd->H & 1)) return HA_EINVAL;
  if (p->dof < 1 || p->dof > 3) return HA_EINVAL;
  if (p->dof == 3 && !reset_uv) return HA_EINVAL;
  if (p->using_weight && !grd_conf) return HA_EINVAL;
  if (p->geometry == HA_GEOM_FORD && !extrinsics) return HA_EINVAL;
  if (ws_bytes < lm_ws_bytes(B)) return HA_ENOSPACE;
  if (((uintptr_t)sat->data | (uintptr_t)grd->data | (uintptr_t)ground_table) & 15) return HA_EINVAL;

  LmStepArgs a;
  a.sat = sat->data; a.grd = grd->data; a.sat_scale = sat->scale; a.grd_scale = grd->scale;
  a.conf = grd_conf; a.table = reinterpret_cast<const float4*>(ground_table); a.extr = extrinsics;
  a.pose = pose; a.reset_uv = reset_uv; a.stats = stats; a.traj = traj_step; a.traj_stride = traj_stride;
  a.status = status;
  a.partial = reinterpret_cast<double*>(ws);
  a.ticket = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(ws) + (size_t)B * kLmMaxCtasPerSample * kLmAcc * sizeof(double));
  a.B = B; a.A = sat->H; a.H = grd->H; a.W = grd->W;
  a.dof = p->dof; a.using_weight = p->using_weight; a.use_hessian = p->use_hessian;
  a.rot = p->rotation_range; a.lat = p->shift_range_lat; a.lon = p->shift_range_lon;
  a.mpp = p->meter_per_pixel[level]; a.inv_mpp = p->inv_meter_per_pixel[level]; a.center = p->sat_center[level];
  for (int i = 0; i < 3; ++i) a.damping[i] = p->damping[i];
  const int P = (grd->H - grd->H / 2) * grd->W;
  const int nc = choose_ctas_per_sample(B, P, grd->C);
  a.px_per_cta = (P + nc - 1) / nc;
  dim3 grid((P + a.px_per_cta - 1) / a.px_per_cta, B);
  if (p->geometry == HA_GEOM_KITTI) return launch_by_channels<HA_GEOM_KITTI>(grd->C, grid, st, a);
  if (p->geometry == HA_GEOM_FORD) return launch_by_channels<HA_GEOM_FORD>(grd->C, grid, st, a);
  return HA_EINVAL;
}

__global__ void zero_u32_kernel(uint32_t* p, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0;
}

}  // namespace ha

// ------------------------------------------------------------------------------------ C ABI
extern "C" size_t ha_lm_workspace_bytes(int B) { return B > 0 ? ha::lm_ws_bytes(B) : 0; }

extern "C" int ha_lm_step(const HaLmParams* p, int level, const HaLevel* sat, const HaLevel* grd, const float* grd_conf,
                          const float* ground_table, const float* extrinsics, float* pose, const float* reset_uv,
                          float* stats, uint32_t* status, void* ws, size_t ws_bytes, void* stream) {
  if (!p) return HA_EINVAL;
  const int B = p->batch;
  if (B <= 0) return HA_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // tickets must start at zero: the first use of a fresh workspace zeroes them here (cheap, async)
  uint32_t* ticket = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(ws) +
                                                 (size_t)B * ha::kLmMaxCtasPerSample * ha::kLmAcc * sizeof(double));
  if (ws_bytes < ha::lm_ws_bytes(B)) return HA_ENOSPACE;
  ha::zero_u32_kernel<<<(B + 255) / 256, 256, 0, st>>>(ticket, B);
  ha::count_launches(1);
  return ha::lm_step_impl(p, level, sat, grd, grd_conf, ground_table, extrinsics, pose, reset_uv, stats, nullptr, 0,
                          status, ws, ws_bytes, B, st);
}

extern "C" int ha_lm_run(const HaLmParams* p, const HaLevel* sat, const HaLevel* grd, const float* const* grd_conf,
                         const float* const* ground_tables, const float* extrinsics, float* pose,
                         const float* reset_uv, float* traj, float* stats, uint32_t* status, void* ws, size_t ws_bytes,
                         void* stream) {
  if (!p || !sat || !grd || !ground_tables || !pose) return HA_EINVAL;
  const int B = p->batch, L = p->n_levels, N = p->n_iters;
  if (B <= 0 || L < 1 || L > HA_MAX_LEVELS || N < 1) return HA_EINVAL;
  if (ws_bytes < ha::lm_ws_bytes(B)) return HA_ENOSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint32_t* ticket = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(ws) +
                                                 (size_t)B * ha::kLmMaxCtasPerSample * ha::kLmAcc * sizeof(double));
  ha::zero_u32_kernel<<<(B + 255) / 256, 256, 0, st>>>(ticket, B);
  ha::count_launches(1);
  int k = 0;
  const int outer = p->level_first ? L : N, inner = p->level_first ? N : L;
  for (int o = 0; o < outer; ++o) {
    for (int i = 0; i < inner; ++i, ++k) {
      const int it = p->level_first ? i : o, lv = p->level_first ? o : i;
      const float* ruv = (p->dof == 3 && reset_uv) ? reset_uv + (size_t)k * 2 * B : nullptr;
      float* tr = traj ? traj + ((size_t)it * L + lv) * 3 : nullptr;
      float* stp = stats ? stats + ((size_t)it * L + lv) * B * HA_STATS : nullptr;
      int rc = ha::lm_step_impl(p, lv, sat + lv, grd + lv, grd_conf ? grd_conf[lv] : nullptr, ground_tables[lv],
                                extrinsics, pose, ruv, stp, tr, N * L * 3, status, ws, ws_bytes, B, st);
      if (rc != HA_OK) return rc;
    }
  }
  return HA_OK;
}
