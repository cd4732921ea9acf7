// Backward of ONE fused LM step (S2GP geometries, 3 degrees of freedom): the adjoint of what lm_step_*_kernel
// computes, for training through the unrolled loop (train_kitti.py:365 `loss.backward()` differentiates
// models_kitti.py:1176-1260 / models_ford.py:652-800: project_map_to_grd -> jacobian.grid_sample -> LM_update).
//
// Forward of a step (per sample; q = bottom-half ground pixel, c = channel; lm_kernels.cu header):
//   s_c, a_c = ds_c/dx, b_c = ds_c/dy   bilinear sample of the satellite features at uv(q, pose) and its analytic
//                                       derivative (jacobian.py:174-193);  g_c = ground feature
//   J_c = D_q^T (a_c, b_c),  D_q = d(u,v)/d(pose)  (2 x 3);  fs = 1/ns^2, fg = 1/(ns ng), ns = max(|s|, 1e-6), ng likewise
//   H = fs sum w J_c J_c^T,  grad = sum w J_c (fs s_c - fg g_c),  delta = -(H + lambda (.) M)^-1 grad,  pose' = pose + delta
// Adjoint, given gpose' (3 per sample):
//   bbar = -(H + lambda M)^-T dbar,  Abar = bbar delta^T,  Hbar = Abar (+ use_hessian terms),  lambdabar_j = Abar_jj M_jj
//   Jbar_c = w [ fs (Hbar + Hbar^T) J_c + bbar (fs s_c - fg g_c) ]
//   sbar_c = w fs (bbar . J_c) + 2 s_c SSbar,   gbar_c = -w fg (bbar . J_c) + 2 g_c GGbar     (SSbar, GGbar: through ns, ng)
//   (abar_c, bbar_c) = D_q Jbar_c,  Dbar_q += (a_c, b_c) Jbar_c^T
//   taps:  nwbar = sbar sy ex - abar sy - bbar ex, ... (scatter-add into the satellite gradient: atomics)
//   uv:    xbar = sum_c sbar_c a_c + bbar_c m_c,  ybar = sum_c sbar_c b_c + abar_c m_c,  m_c = d2s/dxdy = (se - sw) - (ne - nw)
//          (the "second-order terms through uv": the analytic Jacobian itself depends on the sample position)
//   pose:  gpose += D_q^T (xbar, ybar) + <Dbar_q, dD_q/dpose>;  D's constant columns rotate with theta
//          (d(jx, jy)/dtheta = k (jy, -jx)), its theta column is k (v - c, -(u - c)).
// The per-sample totals the adjoint needs (H, grad, ns, ng, delta, the J^T g part of grad, which shifts were re-drawn)
// are the forward's HA_STATS record of the step; the features are re-sampled here (nothing was materialised).
// One warp per pixel, lanes across channels; first slice of SURVEY.md section 8 f-1 — correctness first, not tuned.
#include "lm_common.cuh"

namespace ha {

constexpr int kBwdAcc = 8;       // per-sample reduced scalars: SX, SY, ST, Da0, Db0, Da1, Db1, (unused)

struct LmBwdArgs {
  LmStepArgs f;                  // forward arguments: features, table, extrinsics, geometry constants; f.pose = pose_in
  const float* stats;            // [B][HA_STATS] forward diagnostics of this step
  const float* gpose_out;        // [B][3] adjoint of the step's output pose
  float* gpose_in;               // [B][3] adjoint of the step's input pose (written)
  float* gsat;                   // [B][A][A][C]  += (atomics: several ground pixels share a texel)
  float* ggrd;                   // [B][H][W][C]  += (one owner per pixel)
  float* glambda;                // [B][3] += adjoint of the damping columns
  double* bpartial;              // [B][kLmMaxCtasPerSample][kBwdAcc]
  uint32_t* bticket;             // [B]
};

struct BwdConsts {               // per-sample constants of the adjoint, computed once per CTA
  float Hs[3][3];                // fs (Hbar + Hbar^T)
  float bbar[3];
  float fs, fg, ss2, gg2;        // 2 SSbar, 2 GGbar
  float d0x, d0y, d1x, d1y, kk;
  float direct[3];               // adjoint that passes straight through pose' = pose + delta (0 where re-drawn)
  float lam_bar[3];
};

template <int GEOM>
__global__ void __launch_bounds__(kLmThreads) lm_step_backward_kernel(const LmBwdArgs g) {
  const LmStepArgs& a = g.f;
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int C = a.grd_C, C4 = C / 4;
  const int P = (a.H - a.H / 2) * a.W;
  const int q_begin = blockIdx.x * a.px_per_cta;
  const int q_end = min(P, q_begin + a.px_per_cta);

  const float su = a.pose[b * 3 + 0], sv = a.pose[b * 3 + 1], th = a.pose[b * 3 + 2];
  KittiPose kp;
  FordPose fp;
  G2spPose gq;
  if (GEOM == HA_GEOM_KITTI) kp = kitti_pose(a, su, sv, th);
  else fp = ford_pose(a, b, su, sv, th);

  __shared__ BwdConsts K;
  if (threadIdx.x == 0) {
    const float* st = g.stats + (size_t)b * HA_STATS;
    double H[3][3], grad[3], dl[3], G3[3], lam[3];
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) H[i][j] = st[HA_STAT_H + i * 3 + j];
      grad[i] = st[HA_STAT_GRAD + i]; dl[i] = st[HA_STAT_DELTA + i]; G3[i] = st[HA_STAT_JTG + i]; lam[i] = a.damping[i];
    }
    const double ns = st[HA_STAT_SAT_NORM], ng = st[HA_STAT_GRD_NORM];
    const int reset = (int)st[HA_STAT_RESET_MASK];
    double dbar[3];
    for (int i = 0; i < 3; ++i) {
      const bool redrawn = i < 2 && ((reset >> i) & 1);          // torch.where(in range, new, rand): no gradient through a re-draw
      dbar[i] = redrawn ? 0.0 : (double)g.gpose_out[b * 3 + i];
      K.direct[i] = (float)dbar[i];
    }
    // A = H + lambda (.) M, M = I or diag(H) (models_kitti.py:1005-1012); symmetric
    double A[3][3];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) A[i][j] = H[i][j] + (i == j ? (a.use_hessian ? lam[j] * H[j][j] : lam[j]) : 0.0);
    const double c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1], c01 = A[1][2] * A[2][0] - A[1][0] * A[2][2];
    const double c02 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
    const double det = A[0][0] * c00 + A[0][1] * c01 + A[0][2] * c02;
    const double c10 = A[0][2] * A[2][1] - A[0][1] * A[2][2], c11 = A[0][0] * A[2][2] - A[0][2] * A[2][0];
    const double c12 = A[0][1] * A[2][0] - A[0][0] * A[2][1];
    const double c20 = A[0][1] * A[1][2] - A[0][2] * A[1][1], c21 = A[0][2] * A[1][0] - A[0][0] * A[1][2];
    const double c22 = A[0][0] * A[1][1] - A[0][1] * A[1][0];
    // y = A^-T dbar (adj / det with adj[i][j] = cofactor[j][i]; transpose -> cofactor[i][j])
    const double cof[3][3] = {{c00, c01, c02}, {c10, c11, c12}, {c20, c21, c22}};
    double bb[3];
    for (int i = 0; i < 3; ++i) bb[i] = -(cof[i][0] * dbar[0] + cof[i][1] * dbar[1] + cof[i][2] * dbar[2]) / det;
    double Hb[3][3];                                              // Hbar
    double hh = 0.0;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        const double Ab = bb[i] * dl[j];                          // Abar = bbar delta^T
        Hb[i][j] = Ab * ((i == j && a.use_hessian) ? 1.0 + lam[j] : 1.0);
        hh += Hb[i][j] * H[i][j];
        if (i == j) K.lam_bar[j] = (float)(Ab * (a.use_hessian ? H[j][j] : 1.0));
      }
    const double fs = 1.0 / (ns * ns), fg = 1.0 / (ns * ng);
    const double bg = bb[0] * grad[0] + bb[1] * grad[1] + bb[2] * grad[2];
    const double bG = bb[0] * G3[0] + bb[1] * G3[1] + bb[2] * G3[2];
    const double nsbar = -(2.0 * hh + 2.0 * bg + bG) / ns, ngbar = bG / ng;
    K.ss2 = ns > 1e-6 ? (float)(nsbar / ns) : 0.f;               // 2 * SSbar = 2 * nsbar / (2 ns)
    K.gg2 = ng > 1e-6 ? (float)(ngbar / ng) : 0.f;
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) K.Hs[i][j] = (float)(fs * (Hb[i][j] + Hb[j][i]));
      K.bbar[i] = (float)bb[i];
    }
    K.fs = (float)fs; K.fg = (float)fg;
    K.d0x = GEOM == HA_GEOM_KITTI ? kp.jux : fp.jux; K.d0y = GEOM == HA_GEOM_KITTI ? kp.juy : fp.juy;
    K.d1x = GEOM == HA_GEOM_KITTI ? kp.jvx : fp.jvx; K.d1y = GEOM == HA_GEOM_KITTI ? kp.jvy : fp.jvy;
    K.kk = (float)((double)a.rot / 180.0 * 3.14159265358979323846);
  }
  __syncthreads();

  const size_t px_base = (size_t)b * a.H * a.W + (size_t)(a.H / 2) * a.W;
  const float4* grd = reinterpret_cast<const float4*>(a.grd) + px_base * C4;
  float4* ggrd = reinterpret_cast<float4*>(g.ggrd) + px_base * C4;
  const float4* sat = reinterpret_cast<const float4*>(a.sat) + (size_t)b * a.A * a.A * C4;
  float4* gsat = reinterpret_cast<float4*>(g.gsat) + (size_t)b * a.A * a.A * C4;
  const float4* tab = a.table + (size_t)(a.H / 2) * a.W;
  const float* conf = nullptr;                                    // using_weight is not part of this slice (host refuses it)

  float acc[kBwdAcc] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int q = q_begin + warp; q < q_end; q += kLmWarps) {
    const float4 tab_px = __ldg(tab + q);
    const PixelScalars ps = pixel_scalars<GEOM>(a, kp, fp, gq, tab_px, conf, q, q_end, C4);
    if (ps.goff < 0) continue;                                    // geometric mask: nothing of this pixel enters the step
    const bool taps = ps.valid != 0.f && (ps.ex + ps.wx == 1.f) && (ps.sy + ps.ny == 1.f);
    const float wx = ps.wx, ny = ps.ny, ex = 1.f - wx, sy = 1.f - ny, tx = ps.tx, ty = ps.ty, om = ps.om;
    float X = 0.f, Y = 0.f, Dtx = 0.f, Dty = 0.f, Da0 = 0.f, Db0 = 0.f, Da1 = 0.f, Db1 = 0.f;
    for (int c4 = lane; c4 < C4; c4 += 32) {
      const float4 gv = __ldg(grd + (size_t)q * C4 + c4);
      float4 gg = make_float4(K.gg2 * gv.x, K.gg2 * gv.y, K.gg2 * gv.z, K.gg2 * gv.w);       // through |g|
      if (taps) {
        const float4 nw4 = __ldg(sat + ps.off_n + c4), ne4 = __ldg(sat + ps.off_n + ps.east + c4);
        const float4 sw4 = __ldg(sat + ps.off_s + c4), se4 = __ldg(sat + ps.off_s + ps.east + c4);
        const float nwv[4] = {nw4.x, nw4.y, nw4.z, nw4.w}, nev[4] = {ne4.x, ne4.y, ne4.z, ne4.w};
        const float swv[4] = {sw4.x, sw4.y, sw4.z, sw4.w}, sev[4] = {se4.x, se4.y, se4.z, se4.w};
        const float gvv[4] = {gv.x, gv.y, gv.z, gv.w};
        float tnw[4], tne[4], tsw[4], tse[4], gb[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float dn = nev[e] - nwv[e], ds = sev[e] - swv[e];
          const float top = fmaf(dn, wx, nwv[e]), bot = fmaf(ds, wx, swv[e]);
          const float bb_ = bot - top, m = ds - dn;
          const float aa = fmaf(m, ny, dn), s = fmaf(bb_, ny, top);
          const float j0 = K.d0x * aa + K.d0y * bb_, j1 = K.d1x * aa + K.d1y * bb_, j2 = tx * aa + ty * bb_;
          const float rc = K.fs * s - K.fg * gvv[e];
          const float J0 = om * (K.Hs[0][0] * j0 + K.Hs[0][1] * j1 + K.Hs[0][2] * j2 + K.bbar[0] * rc);
          const float J1 = om * (K.Hs[1][0] * j0 + K.Hs[1][1] * j1 + K.Hs[1][2] * j2 + K.bbar[1] * rc);
          const float J2 = om * (K.Hs[2][0] * j0 + K.Hs[2][1] * j1 + K.Hs[2][2] * j2 + K.bbar[2] * rc);
          const float bj = K.bbar[0] * j0 + K.bbar[1] * j1 + K.bbar[2] * j2;
          const float sb = om * K.fs * bj + K.ss2 * s;
          gb[e] = -om * K.fg * bj;
          const float ab = K.d0x * J0 + K.d1x * J1 + tx * J2, bbb = K.d0y * J0 + K.d1y * J1 + ty * J2;
          X += sb * aa + bbb * m; Y += sb * bb_ + ab * m;
          Dtx += aa * J2; Dty += bb_ * J2; Da0 += aa * J0; Db0 += bb_ * J0; Da1 += aa * J1; Db1 += bb_ * J1;
          tnw[e] = sb * sy * ex - ab * sy - bbb * ex; tne[e] = sb * sy * wx + ab * sy - bbb * wx;
          tsw[e] = sb * ny * ex - ab * ny + bbb * ex; tse[e] = sb * ny * wx + ab * ny + bbb * wx;
        }
        gg.x += gb[0]; gg.y += gb[1]; gg.z += gb[2]; gg.w += gb[3];
        atomicAdd(gsat + ps.off_n + c4, make_float4(tnw[0], tnw[1], tnw[2], tnw[3]));
        atomicAdd(gsat + ps.off_n + ps.east + c4, make_float4(tne[0], tne[1], tne[2], tne[3]));
        atomicAdd(gsat + ps.off_s + c4, make_float4(tsw[0], tsw[1], tsw[2], tsw[3]));
        atomicAdd(gsat + ps.off_s + ps.east + c4, make_float4(tse[0], tse[1], tse[2], tse[3]));
      }
      float4 cur = ggrd[(size_t)q * C4 + c4];
      cur.x += gg.x; cur.y += gg.y; cur.z += gg.z; cur.w += gg.w;
      ggrd[(size_t)q * C4 + c4] = cur;
    }
    // d(u,v)/dtheta = k (v - c, -(u - c)): the adjoint of that column flows back into (u, v)
    const float xt = X - K.kk * Dty, yt = Y + K.kk * Dtx;
    acc[0] += xt; acc[1] += yt; acc[2] += tx * xt + ty * yt;
    acc[3] += Da0; acc[4] += Db0; acc[5] += Da1; acc[6] += Db1;
  }

  // ---- CTA reduction -> partial row; the last CTA of the sample combines in index order and writes gpose_in
  __shared__ double red[kLmWarps][kBwdAcc];
  __shared__ bool is_last;
#pragma unroll
  for (int i = 0; i < kBwdAcc; ++i) {
    const double r = warp_sum((double)acc[i]);
    if (lane == 0) red[warp][i] = r;
  }
  __syncthreads();
  double* part = g.bpartial + ((size_t)b * kLmMaxCtasPerSample + blockIdx.x) * kBwdAcc;
  if (threadIdx.x < kBwdAcc) {
    double r = 0;
    for (int w = 0; w < kLmWarps; ++w) r += red[w][threadIdx.x];
    part[threadIdx.x] = r;
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t prev = atomicAdd(g.bticket + b, 1u);
    is_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  __shared__ double tot[kBwdAcc];
  if (threadIdx.x < kBwdAcc) {
    const volatile double* pp = g.bpartial + (size_t)b * kLmMaxCtasPerSample * kBwdAcc + threadIdx.x;
    double r = 0;
    for (unsigned c = 0; c < gridDim.x; ++c) r += pp[(size_t)c * kBwdAcc];
    tot[threadIdx.x] = r;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  g.bticket[b] = 0;
  const double SX = tot[0], SY = tot[1], ST = tot[2];
  const double p0 = K.d0x * SX + K.d0y * SY, p1 = K.d1x * SX + K.d1y * SY;
  // the constant columns of D rotate with theta: d(jx, jy)/dtheta = k (jy, -jx)
  const double p2 = ST + (double)K.kk * (tot[3] * K.d0y - tot[4] * K.d0x + tot[5] * K.d1y - tot[6] * K.d1x);
  g.gpose_in[b * 3 + 0] = (float)(K.direct[0] + p0);
  g.gpose_in[b * 3 + 1] = (float)(K.direct[1] + p1);
  g.gpose_in[b * 3 + 2] = (float)(K.direct[2] + p2);
  for (int j = 0; j < 3; ++j) g.glambda[b * 3 + j] += K.lam_bar[j];
}

__global__ void zero_words_kernel(uint32_t* p, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = 0;
}

static size_t lm_bwd_ws_bytes(int B) {
  return (size_t)B * kLmMaxCtasPerSample * kBwdAcc * sizeof(double) + ((size_t)B * 4 + 255) / 256 * 256;
}

}  // namespace ha

extern "C" size_t ha_lm_backward_workspace_bytes(int B) { return B > 0 ? ha::lm_bwd_ws_bytes(B) : 0; }

extern "C" int ha_lm_step_backward(const HaLmParams* p, int level, const HaLevel* sat, const HaLevel* grd,
                                   const float* ground_table, const float* extrinsics, const float* pose_in,
                                   const float* stats, const float* gpose_out, float* gpose_in, float* gsat, float* ggrd,
                                   float* glambda, void* ws, size_t ws_bytes, void* stream) {
  using namespace ha;
  if (!p || !sat || !grd || !ground_table || !pose_in || !stats || !gpose_out || !gpose_in || !gsat || !ggrd || !glambda || !ws)
    return HA_EINVAL;
  const int B = p->batch;
  if (B <= 0 || level < 0 || level >= HA_MAX_LEVELS) return HA_EINVAL;
  // this slice: the S2GP geometries, all three degrees of freedom, unweighted residuals, unit feature scales
  if (p->geometry != HA_GEOM_KITTI && p->geometry != HA_GEOM_FORD) return HA_EINVAL;
  if (p->dof != 3 || p->using_weight || sat->scale || grd->scale) return HA_EINVAL;
  if (p->geometry == HA_GEOM_FORD && !extrinsics) return HA_EINVAL;
  if (sat->C != grd->C || sat->H != sat->W || (grd->H & 1) || (grd->C % 4)) return HA_EINVAL;
  if (p->optimizer != HA_OPT_LM || p->full_height) return HA_EINVAL;     // the adjoint exists for the default update rule and crop
  if (ws_bytes < lm_bwd_ws_bytes(B)) return HA_ENOSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LmBwdArgs g;
  LmStepArgs& a = g.f;
  a.sat = sat->data; a.grd = grd->data; a.sat_scale = nullptr; a.grd_scale = nullptr; a.conf = nullptr;
  a.table = reinterpret_cast<const float4*>(ground_table); a.extr = extrinsics;
  a.pose = const_cast<float*>(pose_in); a.reset_uv = nullptr; a.stats = nullptr; a.traj = nullptr; a.traj_stride = 0;
  a.status = nullptr; a.partial = nullptr; a.ticket = nullptr; a.zeros = nullptr; a.gg_cache = nullptr; a.step_word = nullptr;
  a.B = B; a.A = sat->H; a.H = grd->H; a.W = grd->W; a.grd_C = grd->C;
  a.dof = 3; a.using_weight = 0; a.use_hessian = p->use_hessian;
  a.rot = p->rotation_range; a.lat = p->shift_range_lat; a.lon = p->shift_range_lon;
  a.mpp = p->meter_per_pixel[level]; a.inv_mpp = p->inv_meter_per_pixel[level]; a.center = p->sat_center[level];
  for (int i = 0; i < 3; ++i) a.damping[i] = p->damping[i];
  a.ori_h = p->ori_grd_h; a.ori_w = p->ori_grd_w; a.variant = 0;
  a.row0 = grd->H / 2; a.optimizer = HA_OPT_LM; a.adam_t = 0; a.adam_b1 = a.adam_b2 = 0.f; a.adam_mv = nullptr; a.g2sp_nn = 0;
  const int P = (grd->H - grd->H / 2) * grd->W;
  int ctas = (kNumSMs * 8 + B - 1) / B;
  if (ctas > kLmMaxCtasPerSample) ctas = kLmMaxCtasPerSample;
  int per = (P + ctas - 1) / ctas;
  per = (per + kLmWarps - 1) / kLmWarps * kLmWarps;
  if (per < kLmWarps) per = kLmWarps;
  a.px_per_cta = per;
  g.stats = stats; g.gpose_out = gpose_out; g.gpose_in = gpose_in; g.gsat = gsat; g.ggrd = ggrd; g.glambda = glambda;
  g.bpartial = reinterpret_cast<double*>(ws);
  g.bticket = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(ws) + (size_t)B * kLmMaxCtasPerSample * kBwdAcc * sizeof(double));
  zero_words_kernel<<<(B + 255) / 256, 256, 0, st>>>(g.bticket, B);
  dim3 grid((P + per - 1) / per, B);
  if (p->geometry == HA_GEOM_KITTI) lm_step_backward_kernel<HA_GEOM_KITTI><<<grid, kLmThreads, 0, st>>>(g);
  else lm_step_backward_kernel<HA_GEOM_FORD><<<grid, kLmThreads, 0, st>>>(g);
  count_launches(2);
  return check_launch("lm_step_backward_kernel");
}

// ------------------------------------------------------------------------------------ pose loss (loss_func, method 0)
// models_ford.py:1041-1093 with loss_method 0 (shared by KITTI through models_kitti.py:16): direct supervision of the
// pose trajectory.  err[n][l][k] = mean_b |traj[b][n][l][k] - gt[b][k]|  (k over the engine's (shift_u, shift_v, theta)),
// loss = mean_{n,l} sum_k coe[k] err[n][l][k].  One launch forward, one backward; the rest of the reference's 13-tuple
// (decreases, last-step values) are differences / slices of `err`.
namespace ha {

__global__ void pose_loss_kernel(const float* __restrict__ traj, const float* __restrict__ gt, int B, int NL, float c0, float c1,
                                 float c2, float* __restrict__ err, float* __restrict__ loss) {
  // one warp per (n, l, k); the block's first thread then folds the NL * 3 means into the loss (deterministic order)
  __shared__ float e_s[HA_MAX_LEVELS * 64 * 3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int item = warp; item < NL * 3; item += nwarps) {
    const int k = item % 3, nl = item / 3;
    double acc = 0.0;
    for (int b = lane; b < B; b += 32) acc += fabs((double)traj[((size_t)b * NL + nl) * 3 + k] - (double)gt[b * 3 + k]);
    acc = warp_sum(acc);
    if (lane == 0) { const float m = (float)(acc / B); err[item] = m; e_s[item] = m; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int nl = 0; nl < NL; ++nl) t += (double)c0 * e_s[nl * 3] + (double)c1 * e_s[nl * 3 + 1] + (double)c2 * e_s[nl * 3 + 2];
    *loss = (float)(t / NL);
  }
}

__global__ void pose_loss_backward_kernel(const float* __restrict__ traj, const float* __restrict__ gt, int B, int NL, float c0,
                                          float c1, float c2, const float* __restrict__ gerr, const float* __restrict__ gloss,
                                          float* __restrict__ gtraj) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * NL * 3) return;
  const int k = i % 3, nl = (i / 3) % NL, b = i / (3 * NL);
  const float d = traj[i] - gt[b * 3 + k];
  const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);        // torch.abs: zero sub-gradient at 0
  const float coe = k == 0 ? c0 : (k == 1 ? c1 : c2);
  const float up = (gerr ? gerr[nl * 3 + k] : 0.f) + (gloss ? gloss[0] * coe / NL : 0.f);
  gtraj[i] = up * sgn / B;
}

}  // namespace ha

extern "C" int ha_pose_loss(const float* traj, const float* gt, int B, int n_iters, int n_levels, const float* coe3_host,
                            float* err, float* loss, void* stream) {
  if (!traj || !gt || !coe3_host || !err || !loss || B <= 0 || n_iters <= 0 || n_levels <= 0 || n_levels > HA_MAX_LEVELS) return HA_EINVAL;
  if (n_iters > 64) return HA_EINVAL;
  ha::pose_loss_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(traj, gt, B, n_iters * n_levels, coe3_host[0],
                                                                               coe3_host[1], coe3_host[2], err, loss);
  ha::count_launches(1);
  return ha::check_launch("pose_loss_kernel");
}

extern "C" int ha_pose_loss_backward(const float* traj, const float* gt, int B, int n_iters, int n_levels, const float* coe3_host,
                                     const float* gerr, const float* gloss, float* gtraj, void* stream) {
  if (!traj || !gt || !coe3_host || !gtraj || B <= 0 || n_iters <= 0 || n_levels <= 0) return HA_EINVAL;
  const int n = B * n_iters * n_levels * 3;
  ha::pose_loss_backward_kernel<<<(n + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      traj, gt, B, n_iters * n_levels, coe3_host[0], coe3_host[1], coe3_host[2], gerr, gloss, gtraj);
  ha::count_launches(1);
  return ha::check_launch("pose_loss_backward_kernel");
}
