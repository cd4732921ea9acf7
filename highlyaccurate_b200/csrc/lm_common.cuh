// Shared device code of the LM step kernels (forward: lm_kernels.cu, backward: lm_backward.cu): launch constants, the
// kernel argument block, the per-sample pose constants and the per-pixel warp of the three geometries, evaluated in the
// reference's fp32 operation order (models_kitti.py:700-801, models_ford.py:173-264, models_kitti.py:54-160), and the
// per-pixel sampler scalars (jacobian.py:138-205).
#pragma once
#include <math.h>

#include "common.cuh"

namespace ha {

#ifndef HA_LM_MIN_CTAS
#define HA_LM_MIN_CTAS 4
#endif
constexpr int kLmThreads = 128;
constexpr int kLmWarps = kLmThreads / 32;
constexpr int kLmAcc = 16;              // per-sample reduced scalars
constexpr int kLmMaxCtasPerSample = 256;
constexpr int kLmZeroBytes = 2048;       // zero vector for masked ground pixels (C <= 256 fp32, both halves)

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
  const float4* zeros;     // >= 1 KB of zeros (read in place of masked ground pixels)
  double* gg_cache;        // [B] sum g^2 of this level (written by FULL launches, read by the others) or null
  unsigned long long* step_word;  // batch-level arrival word of the current step: (samples finished << 32) | samples with an in-range point
  int traj_stride;         // floats between consecutive samples in traj
  int B, A, H, W;
  int grd_C;               // channel count (the forward kernels take it as a template parameter; the backward kernel reads it)
  int px_per_cta;          // bottom-half pixels handled by one CTA
  int dof, using_weight, use_hessian;
  float rot, lat, lon;     // rotation_range (deg), shift_range_lat / lon (m)
  float mpp, inv_mpp, center;  // satellite metres per pixel, fp32(1/mpp), A/2
  int ori_h, ori_w;        // G2SP: size of the ground IMAGE the camera matrix refers to (models_kitti.py:111-114)
  int variant;             // HaLmParams.kernel_variant
  float damping[3];
  int row0;                // first ground row of the residual: H/2 (proj 'geo', models_kitti.py:1194-1199) or 0 (other proj)
  int optimizer;           // HA_OPT_*
  int adam_t;              // HA_OPT_ADAM: t of this step (iter * args.level + level, models_kitti.py:1241)
  float adam_b1, adam_b2;  // HA_OPT_ADAM: args.beta1 / args.beta2
  float* adam_mv;          // HA_OPT_ADAM: [B][6] first / second moments, carried between the steps of a run
  int g2sp_nn;             // HA_GEOM_G2SP_NN: the in-plane warp of models_kitti.py:289-331 instead of the camera projection
  int* done;               // chained launches (ha_lm_run): [B] steps finished per sample; null = plain stream order
  int step_index;          // chained launches: index of this step in the run (it waits for done[b] >= step_index)
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

// G2SP (models_kitti.py:86-160): P = K_l [R(-heading) | T] and the three dP/dpose, per sample.
struct G2spPose {
  float P[3][4];        // projection of (X, 0, Z, 1); column 1 multiplies Y = 0 and is dropped
  float dPt[3][2];      // dP/dtheta columns 0 (X) and 2 (Z); its last column is 0
  float du[3], dv[3];   // dP/dsu, dP/dsv: only the last column is non-zero -> duv1/dshift are per-sample constants
  float nn_c, nn_s, nn_tx, nn_ty, nn_ju, nn_jv, nn_k;   // proj 'nn' (inplane_grd_to_map): R(theta), T in pixels, d/dsu, d/dsv, k
};

__device__ __forceinline__ G2spPose g2sp_pose(const LmStepArgs& a, int b, float su, float sv, float th) {
  G2spPose g;
  const float pi_f = 3.14159265358979323846f;
  if (a.g2sp_nn) {
    // models_kitti.py:294-303: T = (-lon su, lat sv) / mpp pixels, R = [[cos, -sin], [sin, cos]] of +heading
    const float heading_nn = __fmul_rn(__fdiv_rn(__fmul_rn(th, a.rot), 180.f), pi_f);
    sincosf(heading_nn, &g.nn_s, &g.nn_c);
    g.nn_tx = -__fdiv_rn(__fmul_rn(a.lon, su), a.mpp);
    g.nn_ty = __fdiv_rn(__fmul_rn(a.lat, sv), a.mpp);
    g.nn_ju = -__fdiv_rn(a.lon, a.mpp);                                                  // :316-318
    g.nn_jv = __fdiv_rn(a.lat, a.mpp);                                                   // :319-321
    g.nn_k = (float)((double)a.rot / 180.0 * 3.14159265358979323846);                    // :322
    return g;
  }
  const float shu = __fmul_rn(a.lon, su), shv = __fmul_rn(a.lat, sv);                  // :92-93
  const float heading = __fmul_rn(__fdiv_rn(__fmul_rn(th, a.rot), 180.f), pi_f);       // :94
  float sn, cs;
  sincosf(-heading, &sn, &cs);                                                          // :96-97
  float k[3][3];
#pragma unroll
  for (int i = 0; i < 9; ++i) k[i / 3][i % 3] = a.extr[b * 9 + i];
  // camera_k rows scaled to this level: row 0 * grd_W / ori_grdW, row 1 * grd_H / ori_grdH  (:111-114)
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    k[0][j] = __fdiv_rn(__fmul_rn(k[0][j], (float)a.W), (float)a.ori_w);
    k[1][j] = __fdiv_rn(__fmul_rn(k[1][j], (float)a.H), (float)a.ori_h);
  }
  const float R[3][3] = {{cs, 0.f, -sn}, {0.f, 1.f, 0.f}, {sn, 0.f, cs}};
  const float T[3] = {shv, 1.65f, -shu};                                                // :103-105
  const float kk = (float)((double)a.rot / 180.0 * 3.14159265358979323846);
  const float dR[3][3] = {{__fmul_rn(kk, sn), 0.f, __fmul_rn(kk, cs)}, {0.f, 0.f, 0.f}, {__fmul_rn(kk, -cs), 0.f, __fmul_rn(kk, sn)}};
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = k[r][0] * R[0][c] + k[r][1] * R[1][c] + k[r][2] * R[2][c];       // P = K [R | T]  (:116)
      g.P[r][c] = v;
    }
    g.P[r][3] = k[r][0] * T[0] + k[r][1] * T[1] + k[r][2] * T[2];
    g.dPt[r][0] = k[r][0] * dR[0][0] + k[r][1] * dR[1][0] + k[r][2] * dR[2][0];          // dP/dtheta = K [dR | 0]  (:133)
    g.dPt[r][1] = k[r][0] * dR[0][2] + k[r][1] * dR[1][2] + k[r][2] * dR[2][2];
    g.du[r] = k[r][2] * (-a.lon);                                                        // K [0 | lon*(0,0,-1)]  (:127,131)
    g.dv[r] = k[r][0] * a.lat;                                                           // K [0 | lat*(1,0,0)]   (:128,132)
  }
  return g;
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


// Per-pixel scalars, evaluated ONCE per pixel by one lane (phase A) and broadcast by shuffles to
// the lanes that share the pixel's channels (phase B).  Taps follow jacobian.py:147-193 (clamped
// corners, inclusive range mask); for a pixel outside the satellite map or behind the camera the
// weights are zeroed, which reproduces `* mask` (models_kitti.py:927-929) with no control flow.
struct PixelScalars {
  float ex, wx, sy, ny;        // xe-x, x-xw, ys-y, y-yn   (0 when the sample point is masked)
  float tx, ty;                // d(u,v)/dtheta
  float om;                    // LM weight (grd_conf or 1)
  float valid;                 // 1 when the sample point is inside the satellite map and in front of the camera
  int off_n, off_s, east;      // float4 offsets of the north / south tap rows into the sample's map, and of +1 texel
  int goff;                    // float4 offset of the streamed pixel, or -1: read zeros (masked / past the end)
  float d0x, d0y, d1x, d1y;    // G2SP only: d(u,v)/dsu and d(u,v)/dsv vary per pixel (quotient rule)
};

template <int GEOM>
__device__ __forceinline__ PixelScalars pixel_scalars(const LmStepArgs& a, const KittiPose& kp, const FordPose& fp,
                                                      const G2spPose& gp, const float4 tab_px, const float* conf, int q,
                                                      int q_end, int c4) {
  PixelScalars r;
  r.ex = r.wx = r.sy = r.ny = 0.f; r.tx = r.ty = 0.f; r.om = 1.f; r.valid = 0.f;
  r.off_n = r.off_s = r.east = 0; r.goff = -1;
  r.d0x = r.d0y = r.d1x = r.d1y = 0.f;
  if (q >= q_end) return r;
  float x, y;
  int IW, IH;
  if (GEOM == HA_GEOM_G2SP && a.g2sp_nn) {
    // proj 'nn' (models_kitti.py:289-331): uv = R (p - A/2) + T + A/2 on the square ground feature map, mask all ones
    const int i = q / a.A, j = q - i * a.A;
    const float half = a.center;
    const float u2 = (float)j - half, v2 = (float)i - half;
    x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(gp.nn_c, u2), __fmul_rn(-gp.nn_s, v2)), gp.nn_tx), half);
    y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(gp.nn_s, u2), __fmul_rn(gp.nn_c, v2)), gp.nn_ty), half);
    IW = a.W; IH = a.H;
    if (!((x >= 0.f) && (x <= (float)(IW - 1)) && (y >= 0.f) && (y <= (float)(IH - 1)))) return r;
    r.goff = q * c4;
    r.d0x = gp.nn_ju; r.d0y = 0.f; r.d1x = 0.f; r.d1y = gp.nn_jv;
    // dR = k [[-sin, -cos], [cos, -sin]]  (:322-327)
    r.tx = __fmul_rn(gp.nn_k, __fadd_rn(__fmul_rn(-gp.nn_s, u2), __fmul_rn(-gp.nn_c, v2)));
    r.ty = __fmul_rn(gp.nn_k, __fadd_rn(__fmul_rn(gp.nn_c, u2), __fmul_rn(-gp.nn_s, v2)));
  } else if (GEOM == HA_GEOM_G2SP) {
    // satellite pixel (row i, col j) -> ground-plane point (X south, 0, Z east)  (models_kitti.py:54-84)
    const int i = q / a.A, j = q - i * a.A;
    const float X = __fmul_rn(a.mpp, (float)(i - (int)a.center)), Z = __fmul_rn(a.mpp, (float)(j - (int)a.center));
    float uv1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) uv1[k] = gp.P[k][0] * X + gp.P[k][2] * Z + gp.P[k][3];
    if (!(uv1[2] > 1e-6f)) return r;                           // behind the camera: Jacobian zeroed (:123,146-148)
    const float w = uv1[2];                                    // == max(w, 1e-6) here
    x = __fdiv_rn(uv1[0], w); y = __fdiv_rn(uv1[1], w);
    IW = a.W; IH = a.H;
    if (!((x >= 0.f) && (x <= (float)(IW - 1)) && (y >= 0.f) && (y <= (float)(IH - 1)))) return r;
    r.goff = q * c4;                                           // only visible pixels read their satellite vector
    const float w2 = w * w;
    float dt1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) dt1[k] = gp.dPt[k][0] * X + gp.dPt[k][1] * Z;
    // quotient rule  d(uv) = d(uv1_xy)/w - uv1_xy * d(w)/w^2   (:139-144)
    r.d0x = gp.du[0] / w - uv1[0] * gp.du[2] / w2; r.d0y = gp.du[1] / w - uv1[1] * gp.du[2] / w2;
    r.d1x = gp.dv[0] / w - uv1[0] * gp.dv[2] / w2; r.d1y = gp.dv[1] / w - uv1[1] * gp.dv[2] / w2;
    r.tx = dt1[0] / w - uv1[0] * dt1[2] / w2; r.ty = dt1[1] / w - uv1[1] * dt1[2] / w2;
  } else {
    const float4 p = tab_px;                                   // ground-plane point (x, y, z, mask) of this pixel
    if (p.w == 0.f) return r;                                  // geometric mask: s, J and g all vanish
    r.goff = q * c4;
    const PixelWarp w = (GEOM == HA_GEOM_KITTI) ? warp_kitti(kp, a, p) : warp_ford(fp, a, p);
    x = w.u; y = w.v; IW = IH = a.A;
    if (!((x >= 0.f) && (x <= (float)(IW - 1)) && (y >= 0.f) && (y <= (float)(IH - 1))))
      return r;                                                // sampler mask: s = 0, J = 0, r = -g~ (taps read texel 0, weights 0)
    r.tx = w.jtx; r.ty = w.jty;
  }
  // bilinear taps of jacobian.py:147-193: four independently clamped corners
  const float hx = (float)(IW - 1), hy = (float)(IH - 1);
  const float x0 = floorf(x), y0 = floorf(y);
  const float xw = fminf(fmaxf(x0, 0.f), hx), xe = fminf(fmaxf(x0 + 1.f, 0.f), hx);
  const float yn = fminf(fmaxf(y0, 0.f), hy), ys = fminf(fmaxf(y0 + 1.f, 0.f), hy);
  const int ixw = (int)xw, ixe = (int)xe, iyn = (int)yn, iys = (int)ys;
  r.off_n = (iyn * IW + ixw) * c4; r.off_s = (iys * IW + ixw) * c4; r.east = (ixe - ixw) * c4;
  r.ex = xe - x; r.wx = x - xw; r.sy = ys - y; r.ny = y - yn; r.valid = 1.f;
  if (a.using_weight && conf) {
    if (GEOM == HA_GEOM_G2SP) {
      // W = grd_conf warped with the same sampler (models_kitti.py:280-282, :361-362)
      const int e = ixe - ixw;
      r.om = conf[iyn * IW + ixw] * (r.ex * r.sy) + conf[iyn * IW + ixw + e] * (r.wx * r.sy) +
             conf[iys * IW + ixw] * (r.ex * r.ny) + conf[iys * IW + ixw + e] * (r.wx * r.ny);
    } else {
      r.om = __ldg(conf + q);                                  // models_kitti.py:994-998
    }
  }
  return r;
}


}  // namespace ha
