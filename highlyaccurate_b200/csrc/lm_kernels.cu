// Fused Levenberg-Marquardt step for satellite->ground pose refinement (sm_100a).
//
// One launch per (iteration, level) does, for every sample of the batch, what the reference
// does with ~60 eager launches and ~300 B of traffic per feature element:
//   geometry  : models_kitti.py:700-801 (grd2cam2world2sat) / models_ford.py:173-264
//   sampler   : jacobian.py:138-205 (4-tap bilinear gather + d/dx, d/dy, chained with duv/dpose)
//   masking   : models_kitti.py:927-929, :1191-1199 (geometric mask, bottom half only)
//   LM update : models_kitti.py:939-1041 / models_ford.py:380-466
// Nothing is materialised: each ground pixel's feature vector is read once (NHWC; bulk async
// copies into a per-warp shared-memory ring in lm_step_v4_kernel, 128-bit register loads in
// lm_step_kernel), the four satellite taps come through L1/L2, and the per-sample normal
// equations are reduced in registers -> shuffles -> shared memory -> a deterministic two-stage
// cross-CTA combine whose last-arriving CTA solves the damped system and updates the pose in place.
// Two kernels share the geometry, the record of per-pixel scalars and the reduce-and-solve tail:
// lm_step_v4_kernel (S2GP geometries, the default) and lm_step_kernel (G2SP, and HaLmParams.kernel_variant = 1).
//
// Algebra: with s_c, a_c = ds_c/dx, b_c = ds_c/dy per channel and the
// channel-independent 2x3 matrix D_p = d(u,v)/d(pose), the kernel accumulates
//   JtJ = sum_p w_p D_p^T [[Saa,Sab],[Sab,Sbb]] D_p,  Jts = sum_p w_p D_p^T [Ssa,Ssb],
//   Jtg = sum_p w_p D_p^T [Sga,Sgb],  SS = sum s^2, GG = sum g^2, SG = sum s g
// and finalises  H = JtJ/ns^2,  grad = Jts/ns^2 - Jtg/(ns*ng)  with ns = max(|s|,1e-6),
// ng = max(|g|,1e-6)  — identical to normalising s, J and g first (models_kitti.py:982-992).
#include <math.h>

#include <algorithm>
#include <atomic>

#include "lm_common.cuh"

namespace ha {

// ---- packed fp32x2 arithmetic (sm_100 FFMA2 / FADD2 / FMUL2): the kernel is issue-bound, and one
// packed instruction does the work of two scalar ones on a channel pair.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f32x2 dup2(float v) { return pk2(v, v); }
__device__ __forceinline__ float sum2(f32x2 v) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); return lo + hi; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ void acc2(f32x2& acc, f32x2 a, f32x2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }
__device__ __forceinline__ void inc2(f32x2& acc, f32x2 a) { asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc) : "l"(a)); }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

struct V4 { f32x2 lo, hi; };   // one 128-bit load = channels (c, c+1) and (c+2, c+3)
__device__ __forceinline__ V4 ld_stream(const float4* p) {   // ground features: read once, keep out of L1
  V4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.b64 {%0, %1}, [%2];" : "=l"(r.lo), "=l"(r.hi) : "l"(p));
  return r;
}
__device__ __forceinline__ V4 ld_cached(const float4* p) {   // satellite taps: reused by neighbouring pixels
  V4 r;
  asm volatile("ld.global.nc.v2.b64 {%0, %1}, [%2];" : "=l"(r.lo), "=l"(r.hi) : "l"(p));
  return r;
}

__device__ __forceinline__ V4 ld_ring(uint32_t saddr) {      // ground features staged in shared memory by the bulk-copy ring
  V4 r;
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(r.lo), "=l"(r.hi) : "r"(saddr));
  return r;
}

struct PixelLoads {            // one 128-bit channel slice of one pixel: ground vector + the four taps
  V4 g, nw, ne, sw, se;
};

// What the reduce-and-solve tail needs beyond the level's LmStepArgs: which of the sample's CTAs this is and the
// per-step outputs, built from the launch (block indices, the argument block: all constants).
struct LmTail {
  int cta_x, n_ctas;               // this CTA among the n_ctas CTAs of the sample
  const float* reset_uv;           // [2][B] draws of this step or null
  float* stats;                    // [B][HA_STATS] of this step or null
  float* traj;                     // &traj[0][it][lv][0] or null
  unsigned long long* step_word;   // batch-level arrival word of this step
  int adam_t;                      // HA_OPT_ADAM: t of this step
  int* done;                       // chained launches (ha_lm_run): [B] steps finished per sample, else null
  int done_val;                    // value released into done[b] once the pose of this step is written
};
__device__ __forceinline__ LmTail lm_tail_of_launch(const LmStepArgs& a) {
  LmTail t;
  t.cta_x = blockIdx.x; t.n_ctas = gridDim.x; t.reset_uv = a.reset_uv; t.stats = a.stats; t.traj = a.traj;
  t.step_word = a.step_word; t.adam_t = a.adam_t; t.done = a.done; t.done_val = a.step_index + 1;
  return t;
}

// CTA reduction + (last CTA of the sample) ordered combine, damped solve and pose update; shared by the step kernels.
// v[]: this lane's sixteen running sums {Gaa, Gab, Gbb, Bx, By, Ctt, Sa, Sb, St, Ga, Gb, Gt, SS, GG, SG, count}.
// Chained launches (ha_lm_run) overlap consecutive steps, so whatever an earlier step wrote (pose, |g|^2 cache, Adam
// moments) is read with ld.global.cg after the sample's `done` flag was acquired: there is no kernel boundary in between.
template <int GEOM, bool FULL>
__device__ __forceinline__ void lm_reduce_and_solve(const LmStepArgs& a, const LmTail& tc, const int b, const double (&v)[kLmAcc],
                                                    const KittiPose& kp, const FordPose& fp, const float su, const float sv,
                                                    const float th) {
  constexpr bool G2SP = (GEOM == HA_GEOM_G2SP);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // ---- CTA reduction: lanes -> warp (fp64 shuffles) -> shared -> one partial row per CTA
  __shared__ double red[kLmWarps][kLmAcc];
  __shared__ bool is_last;
  {
#pragma unroll
    for (int i = 0; i < kLmAcc; ++i) {
      double r = warp_sum(v[i]);
      if (lane == 0) red[warp][i] = r;
    }
  }
  __syncthreads();
  double* part = a.partial + ((size_t)b * kLmMaxCtasPerSample + tc.cta_x) * kLmAcc;
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
    is_last = (prev == (uint32_t)tc.n_ctas - 1u);
  }
  __syncthreads();
  if (!is_last) return;

  // ---- last CTA of this sample: ordered combine of the partials, damped solve, pose update
  // (all threads take part: sixteen threads adding up to 256 partial rows one dependent L2 load after the other were
  // ~10 us of a small batch's ~45-us step, which is bound by exactly this per-sample chain.  Fixed order: eight strided
  // subsets, then the subsets in index order — bit-deterministic for a given CTA split.)
  __threadfence();
  __shared__ double tot8[kLmThreads / kLmAcc][kLmAcc];
  __shared__ double tot[kLmAcc];
  {
    static_assert(kLmThreads % kLmAcc == 0, "subsets");
    const int i = threadIdx.x % kLmAcc, j = threadIdx.x / kLmAcc;
    const double* pp = a.partial + (size_t)b * kLmMaxCtasPerSample * kLmAcc + i;
    double r = 0;
    for (int c = j; c < tc.n_ctas; c += kLmThreads / kLmAcc) r += __ldcg(pp + (size_t)c * kLmAcc);
    tot8[j][i] = r;
  }
  __syncthreads();
  if (threadIdx.x < kLmAcc) {
    double r = 0;
#pragma unroll
    for (int j = 0; j < kLmThreads / kLmAcc; ++j) r += tot8[j][threadIdx.x];
    tot[threadIdx.x] = r;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  a.ticket[b] = 0;   // ready for the next step on this stream
  double Hm[3][3], gr[3], ns = 0.0, ng = 0.0, res_sq = 0.0, jtg_f[3] = {0.0, 0.0, 0.0};
  int reset_mask = 0;
  if (G2SP) {
    // models_kitti.py:333-379: r = grd_proj - sat with the L2-normalised features (VGG.py:172-175) and no
    // further normalisation; the sampled (ground) and streamed (satellite) pyramids carry their own scales
    const double as_ = a.grd_scale ? (double)a.grd_scale[b] : 1.0, bg_ = a.sat_scale ? (double)a.sat_scale[b] : 1.0;
    const double f2 = as_ * as_, fsg = as_ * bg_;
    const double h[6] = {tot[0] * f2, tot[1] * f2, tot[2] * f2, tot[3] * f2, tot[4] * f2, tot[5] * f2};
    Hm[0][0] = h[0]; Hm[0][1] = Hm[1][0] = h[1]; Hm[0][2] = Hm[2][0] = h[2];
    Hm[1][1] = h[3]; Hm[1][2] = Hm[2][1] = h[4]; Hm[2][2] = h[5];
    for (int i = 0; i < 3; ++i) gr[i] = tot[6 + i] * f2 - tot[9 + i] * fsg;
  } else {
    if (FULL) { if (a.gg_cache) a.gg_cache[b] = tot[13]; }
    else tot[13] = __ldcg(a.gg_cache + b);           // sum g^2 over the unmasked bottom half, from this level's first visit

    // assemble J^T W J, J^T W s, J^T W g from the split sums with the per-sample constant rows of D
    const double d0x = (GEOM == HA_GEOM_KITTI) ? kp.jux : fp.jux, d0y = (GEOM == HA_GEOM_KITTI) ? kp.juy : fp.juy;
    const double d1x = (GEOM == HA_GEOM_KITTI) ? kp.jvx : fp.jvx, d1y = (GEOM == HA_GEOM_KITTI) ? kp.jvy : fp.jvy;
    const double Gaa = tot[0], Gab = tot[1], Gbb = tot[2], Bx = tot[3], By = tot[4], Ctt = tot[5];
    const double e0x = Gaa * d0x + Gab * d0y, e0y = Gab * d0x + Gbb * d0y;    // sum(G) * D_0
    const double e1x = Gaa * d1x + Gab * d1y, e1y = Gab * d1x + Gbb * d1y;
    const double JtJ[6] = {d0x * e0x + d0y * e0y, d0x * e1x + d0y * e1y, d0x * Bx + d0y * By,
                           d1x * e1x + d1y * e1y, d1x * Bx + d1y * By, Ctt};
    const double Jts[3] = {d0x * tot[6] + d0y * tot[7], d1x * tot[6] + d1y * tot[7], tot[8]};
    const double Jtg[3] = {d0x * tot[9] + d0y * tot[10], d1x * tot[9] + d1y * tot[10], tot[11]};

    const double alpha = a.sat_scale ? (double)a.sat_scale[b] : 1.0;
    const double beta = a.grd_scale ? (double)a.grd_scale[b] : 1.0;
    double fs, fg;
    if (a.optimizer == HA_OPT_SGD || a.optimizer == HA_OPT_ADAM) {
      // SGD_update / ADAM_update (models_kitti.py:1056-1124): r = s - g on the L2-normalised features as they are
      // (no renormalisation), gradient = sum 2 r J over the whole residual
      ns = alpha * sqrt(tot[12]); ng = beta * sqrt(tot[13]);
      fs = 2.0 * alpha * alpha; fg = 2.0 * alpha * beta;
    } else if (a.optimizer == HA_OPT_GN) {
      // GN_update (models_ford.py:549-566): s / ||s|| without the 1e-6 clamp, g taken as it is
      ns = alpha * sqrt(tot[12]); ng = beta * sqrt(tot[13]);
      fs = alpha * alpha / (ns * ns); fg = alpha * beta / ns;
    } else {
      ns = fmax(alpha * sqrt(tot[12]), 1e-6);          // models_kitti.py:982-984
      ng = fmax(beta * sqrt(tot[13]), 1e-6);           // :987-988
      fs = alpha * alpha / (ns * ns); fg = alpha * beta / (ns * ng);
    }
    Hm[0][0] = JtJ[0] * fs; Hm[0][1] = Hm[1][0] = JtJ[1] * fs; Hm[0][2] = Hm[2][0] = JtJ[2] * fs;
    Hm[1][1] = JtJ[3] * fs; Hm[1][2] = Hm[2][1] = JtJ[4] * fs; Hm[2][2] = JtJ[5] * fs;
    for (int i = 0; i < 3; ++i) { gr[i] = Jts[i] * fs - Jtg[i] * fg; jtg_f[i] = Jtg[i] * fg; }
    res_sq = FULL ? alpha * alpha * tot[12] / (ns * ns) + beta * beta * tot[13] / (ng * ng) -
                        2.0 * alpha * beta * tot[14] / (ns * ng) : 0.0;     // diagnostic, FULL launches only
  }

  // DOF selection (models_kitti.py:954-957): 3 -> (0,1,2), 2 -> (0,1), 1 -> (2)
  const int n = a.dof;
  const int i0 = (n == 1) ? 2 : 0;
  const bool first_order = !G2SP && (a.optimizer == HA_OPT_SGD || a.optimizer == HA_OPT_ADAM);
  double Am[3][3], rhs[3], delta[3] = {0, 0, 0};
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) Am[i][j] = Hm[i0 + i][i0 + j];
    const double lam = (!G2SP && a.optimizer == HA_OPT_GN) ? 0.0 : (double)a.damping[i];   // GN: inverse(Hessian), models_ford.py:577
    Am[i][i] += (a.use_hessian && a.optimizer != HA_OPT_GN) ? lam * Hm[i0 + i][i0 + i] : lam;    // :1005-1012 (column-wise lambda on a diagonal)
    rhs[i] = gr[i0 + i];
  }
  float nsu = su, nsv = sv, nth = th;
  uint32_t st = 0;
  if (first_order) {
    // pose -= 0.01 * step, all three components whatever the ranges, no reset (models_kitti.py:1080-1083, :1119-1123);
    // fp32 arithmetic in the reference's order (python-double hyper-parameters become fp32 scalars)
    float stepv[3] = {(float)gr[0], (float)gr[1], (float)gr[2]};
    if (a.optimizer == HA_OPT_ADAM) {
      float* mv = a.adam_mv + (size_t)b * 6;
      const double b1 = (double)a.adam_b1, b2 = (double)a.adam_b2;
      const float c1 = (float)(1.0 - pow(b1, (double)(tc.adam_t + 1))), c2 = (float)(1.0 - pow(b2, (double)(tc.adam_t + 1)));
      for (int i = 0; i < 3; ++i) {
        const float m0 = tc.adam_t == 0 ? 0.f : __ldcg(mv + i), v0 = tc.adam_t == 0 ? 0.f : __ldcg(mv + 3 + i);   // :1242-1244
        const float m = __fadd_rn(__fmul_rn((float)b1, m0), __fmul_rn((float)(1.0 - b1), stepv[i]));
        const float v = __fadd_rn(__fmul_rn((float)b2, v0), __fmul_rn((float)(1.0 - b2), __fmul_rn(stepv[i], stepv[i])));
        mv[i] = m; mv[3 + i] = v;
        stepv[i] = __fdiv_rn(__fdiv_rn(m, c1), __fadd_rn(sqrtf(__fdiv_rn(v, c2)), 1e-8f));
      }
    }
    for (int i = 0; i < 3; ++i) delta[i] = -(double)__fmul_rn(0.01f, stepv[i]);
    nsu = __fadd_rn(su, (float)delta[0]); nsv = __fadd_rn(sv, (float)delta[1]); nth = __fadd_rn(th, (float)delta[2]);
  } else if (n == 1) {
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
  if (first_order) {
    // pose already updated above
  } else if (n == 3) {
    nsu = su + (float)delta[0]; nsv = sv + (float)delta[1]; nth = th + (float)delta[2];
    // models_kitti.py:1028-1033: shifts outside (-2.5, 2.5) (or NaN) are re-drawn (S2GP models only)
    if (!G2SP) {
      if (!(nsu > -2.5f && nsu < 2.5f)) { nsu = tc.reset_uv[b]; st |= HA_STATUS_RESET; reset_mask |= 1; }
      if (!(nsv > -2.5f && nsv < 2.5f)) { nsv = tc.reset_uv[a.B + b]; st |= HA_STATUS_RESET; reset_mask |= 2; }
    }
  } else if (n == 2) {
    nsu = su + (float)delta[0]; nsv = sv + (float)delta[1];
  } else {
    nth = th + (float)delta[0];
  }
  if (isnan(nsu) || isnan(nsv) || isnan(nth)) st |= HA_STATUS_NAN_POSE;
  const bool any_inrange = tot[15] != 0.0;
  if (!any_inrange) st |= HA_STATUS_SAMPLE_EMPTY;
  // jacobian.py:172 asserts that SOME sample point of the whole batch is in range: the last sample to finish this step
  // sees how many samples had one (the word is reset for the next step of the stream; chained launches have one word per step)
  const unsigned long long prev = atomicAdd(tc.step_word, (1ull << 32) | (any_inrange ? 1ull : 0ull));
  if ((unsigned)(prev >> 32) == (unsigned)a.B - 1u) {
    if ((unsigned)(prev & 0xffffffffull) + (any_inrange ? 1u : 0u) == 0u) st |= HA_STATUS_NO_INRANGE;
    *tc.step_word = 0ull;
  }
  if (st) atomicOr(a.status, st);
  a.pose[b * 3 + 0] = nsu; a.pose[b * 3 + 1] = nsv; a.pose[b * 3 + 2] = nth;
  if (tc.traj) {
    float* tr = tc.traj + (size_t)b * a.traj_stride;
    tr[0] = nsu; tr[1] = nsv; tr[2] = nth;
  }
  if (tc.done) {
    // chained launches: the pose of this step is out; the CTAs of the sample's next step (already resident) may go on
    __threadfence();
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(tc.done + b), "r"(tc.done_val) : "memory");
  }
  if (tc.stats) {
    float* s = tc.stats + (size_t)b * HA_STATS;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) s[HA_STAT_H + i * 3 + j] = (float)Hm[i][j];
    for (int i = 0; i < 3; ++i) s[HA_STAT_GRAD + i] = (float)gr[i];
    s[HA_STAT_SAT_NORM] = (float)ns; s[HA_STAT_GRD_NORM] = (float)ng; s[HA_STAT_RES_SQ] = (float)res_sq;
    for (int i = 0; i < 3; ++i) s[HA_STAT_DELTA + i] = (i < n || first_order) ? (float)delta[i] : 0.f;
    s[HA_STAT_N_INRANGE] = (float)tot[15];
    // saved for the backward pass (ha_lm_step_backward): the J~^T W g~ part of grad and which shifts were re-drawn
    for (int i = 0; i < 3; ++i) s[HA_STAT_JTG + i] = G2SP ? 0.f : (float)jtg_f[i];
    s[HA_STAT_RESET_MASK] = (float)reset_mask;
    s[HA_STATS - 1] = 0.f;
  }
}

template <int GEOM, int C, bool FULL>
__global__ void __launch_bounds__(kLmThreads, HA_LM_MIN_CTAS) lm_step_kernel(const LmStepArgs a) {
  constexpr int LPP = C / 8;                       // lanes per pixel: every lane owns 2 x 4 channels
  constexpr int PPW = 32 / LPP;                    // pixels processed together by one warp
  constexpr int IPG = 32 / PPW;                    // iterations per 32-pixel group
  constexpr int C4 = C / 4;
  static_assert(C % 8 == 0 && LPP >= 1 && LPP <= 32 && PPW * LPP == 32, "channel count");

  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform (see lm_v4_body)
  const int sub = lane / LPP, cl = lane % LPP;     // pixel slot within the warp, channel lane
  constexpr bool G2SP = (GEOM == HA_GEOM_G2SP);
  // S2GP: the residual lives on the bottom half of the ground image (models_kitti.py:1195-1199), the ground
  // features are streamed and the satellite map is gathered.  G2SP: the residual lives on the whole satellite
  // map (:333-379), the satellite features are streamed and the ground features are gathered.
  const int P = G2SP ? a.A * a.A : (a.H - a.row0) * a.W;
  const int q_begin = blockIdx.x * a.px_per_cta;
  const int q_end = min(P, q_begin + a.px_per_cta);

  const float su = a.pose[b * 3 + 0], sv = a.pose[b * 3 + 1], th = a.pose[b * 3 + 2];
  KittiPose kp;
  FordPose fp;
  G2spPose gq;
  if (GEOM == HA_GEOM_KITTI) kp = kitti_pose(a, su, sv, th);
  else if (GEOM == HA_GEOM_FORD) fp = ford_pose(a, b, su, sv, th);
  else gq = g2sp_pose(a, b, su, sv, th);

  // first streamed pixel of sample b, in pixels of the streamed tensor
  const size_t px_base = G2SP ? (size_t)b * a.A * a.A : (size_t)b * a.H * a.W + (size_t)a.row0 * a.W;
  // half 0 = channels [4*cl, 4*cl+4), half 1 = channels [C/2 + 4*cl, ...): each half of a pixel is one
  // contiguous 16*LPP-byte run across the pixel's lanes
  const float4* grd = reinterpret_cast<const float4*>(G2SP ? a.sat : a.grd) + px_base * C4 + cl;              // streamed
  const float4* sat = reinterpret_cast<const float4*>(G2SP ? a.grd : a.sat) +
                      (G2SP ? (size_t)b * a.H * a.W : (size_t)b * a.A * a.A) * C4 + cl;                        // gathered
  const float4* tab = G2SP ? nullptr : a.table + (size_t)a.row0 * a.W;
  const float* conf = a.conf ? (G2SP ? a.conf + (size_t)b * a.H * a.W : a.conf + px_base) : nullptr;

  // per-warp staging of the per-pixel scalars: phase A writes 32 pixels, phase B broadcasts them
  __shared__ __align__(16) float4 ps_s[kLmWarps][32][G2SP ? 4 : 3];

  // running sums over this lane's pixels and channels (two partial sums per register pair)
  f32x2 A_aa = 0, A_ab = 0, A_bb = 0, B_x = 0, B_y = 0, C_tt = 0;
  f32x2 S_a = 0, S_b = 0, S_t = 0, G_a = 0, G_b = 0, G_t = 0, SS = 0, GG = 0, SG = 0;
  float cnt = 0.f;
  // per-pixel channel sums (reset at the first half, consumed at the second)
  f32x2 p_aa = 0, p_ab = 0, p_bb = 0, p_sa = 0, p_sb = 0, p_ga = 0, p_gb = 0;
  float4 sc0 = make_float4(0, 0, 0, 0), sc1 = sc0;   // (ex, wx, sy, ny), (tx, ty, om, valid) of the current pixel
  float4 sc2 = sc0;                                  // G2SP: (d0x, d0y, d1x, d1y)
  const float4* zeros = a.zeros;

  auto accumulate = [&](const PixelLoads& L) {
    const f32x2 ex2 = dup2(sc0.x), wx2 = dup2(sc0.y), sy2 = dup2(sc0.z), ny2 = dup2(sc0.w);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const f32x2 nw = h ? L.nw.hi : L.nw.lo, ne = h ? L.ne.hi : L.ne.lo;
      const f32x2 sw = h ? L.sw.hi : L.sw.lo, se = h ? L.se.hi : L.se.lo;
      const f32x2 g = h ? L.g.hi : L.g.lo;
      const f32x2 top = fma2(ne, wx2, mul2(nw, ex2));          // north row interpolated in x  (jacobian.py:174-186, factored)
      const f32x2 bot = fma2(se, wx2, mul2(sw, ex2));          // south row interpolated in x
      const f32x2 s = fma2(bot, ny2, mul2(top, sy2));
      const f32x2 da = fma2(sub2(se, sw), ny2, mul2(sub2(ne, nw), sy2));   // d/dx  (:190-191)
      const f32x2 db = sub2(bot, top);                         // d/dy = ex (sw - nw) + wx (se - ne)  (:192-193)
      acc2(p_aa, da, da); acc2(p_ab, da, db); acc2(p_bb, db, db);
      acc2(p_sa, s, da); acc2(p_sb, s, db); acc2(p_ga, g, da); acc2(p_gb, g, db);
      if (!G2SP) acc2(SS, s, s);
      if (FULL && !G2SP) { acc2(SG, s, g); acc2(GG, g, g); }          // |g|^2 is pose independent: cached after the first visit
    }
  };
  auto finish_pixel = [&]() {
    if (a.using_weight) {
      const f32x2 om2 = dup2(sc1.z);
      p_aa = mul2(p_aa, om2); p_ab = mul2(p_ab, om2); p_bb = mul2(p_bb, om2);
      p_sa = mul2(p_sa, om2); p_sb = mul2(p_sb, om2); p_ga = mul2(p_ga, om2); p_gb = mul2(p_gb, om2);
    }
    if (G2SP) {
      // all three rows of D = d(u,v)/d(pose) vary per pixel: accumulate J^T W J (A_aa.. reused as H00,H01,H02,
      // H11,H12,H22), J^T W s (S_*) and J^T W g (G_*) directly; there is no renormalisation in G2SP (:354)
      const f32x2 d0x = dup2(sc2.x), d0y = dup2(sc2.y), d1x = dup2(sc2.z), d1y = dup2(sc2.w), d2x = dup2(sc1.x), d2y = dup2(sc1.y);
      const f32x2 e0x = fma2(p_ab, d0y, mul2(p_aa, d0x)), e0y = fma2(p_bb, d0y, mul2(p_ab, d0x));   // G D_0
      const f32x2 e1x = fma2(p_ab, d1y, mul2(p_aa, d1x)), e1y = fma2(p_bb, d1y, mul2(p_ab, d1x));
      const f32x2 e2x = fma2(p_ab, d2y, mul2(p_aa, d2x)), e2y = fma2(p_bb, d2y, mul2(p_ab, d2x));
      acc2(A_aa, d0x, e0x); acc2(A_aa, d0y, e0y); acc2(A_ab, d0x, e1x); acc2(A_ab, d0y, e1y);
      acc2(A_bb, d0x, e2x); acc2(A_bb, d0y, e2y); acc2(B_x, d1x, e1x); acc2(B_x, d1y, e1y);
      acc2(B_y, d1x, e2x); acc2(B_y, d1y, e2y); acc2(C_tt, d2x, e2x); acc2(C_tt, d2y, e2y);
      // J^T s and J^T g kept apart: the lazy L2-norm scales of the two pyramids differ (applied by the last CTA)
      acc2(S_a, d0x, p_sa); acc2(S_a, d0y, p_sb); acc2(S_b, d1x, p_sa); acc2(S_b, d1y, p_sb); acc2(S_t, d2x, p_sa); acc2(S_t, d2y, p_sb);
      acc2(G_a, d0x, p_ga); acc2(G_a, d0y, p_gb); acc2(G_b, d1x, p_ga); acc2(G_b, d1y, p_gb); acc2(G_t, d2x, p_ga); acc2(G_t, d2y, p_gb);
      cnt += sc1.w;
      p_aa = p_ab = p_bb = p_sa = p_sb = p_ga = p_gb = 0ull;
      return;
    }
    // d(u,v)/dsu and /dsv are per-sample constants, only d/dtheta = (tx, ty) varies per pixel, so
    // J^T J splits into sum(G), sum(G t), sum(t^T G t) and the last CTA applies the constant rows.
    const f32x2 tx2 = dup2(sc1.x), ty2 = dup2(sc1.y);
    inc2(A_aa, p_aa); inc2(A_ab, p_ab); inc2(A_bb, p_bb);
    const f32x2 gx = fma2(p_ab, ty2, mul2(p_aa, tx2)), gy = fma2(p_bb, ty2, mul2(p_ab, tx2));
    inc2(B_x, gx); inc2(B_y, gy);
    acc2(C_tt, tx2, gx); acc2(C_tt, ty2, gy);
    inc2(S_a, p_sa); inc2(S_b, p_sb); acc2(S_t, p_sa, tx2); acc2(S_t, p_sb, ty2);
    inc2(G_a, p_ga); inc2(G_b, p_gb); acc2(G_t, p_ga, tx2); acc2(G_t, p_gb, ty2);
    cnt += sc1.w;                                  // every lane of the pixel counts it: divided by LPP below
    p_aa = p_ab = p_bb = p_sa = p_sb = p_ga = p_gb = 0ull;
  };

  // Flat software pipeline: every pixel is two half-iterations (channel halves); the loads of one
  // half are issued before the arithmetic of the previous one (ping-pong buffers A / B), so each
  // lane always has 5-10 independent 128-bit loads in flight.
  const int n_groups = (q_end - q_begin + 31) / 32;
  const int my_groups = (n_groups > warp) ? (n_groups - warp + kLmWarps - 1) / kLmWarps : 0;
  const int T = my_groups * IPG;                   // pixel-iterations of this warp
  PixelLoads bufA, bufB;
  float4 nsc0 = sc0, nsc1 = sc1, nsc2 = sc2;       // scalars of the pixel whose loads are in flight

  // addresses of the pixel whose loads are being issued (set by prepare(), used by load_half())
  const float4 *gp = zeros, *s_n = sat, *s_s = sat;
  int east = 0;
  auto prepare = [&](int t) {                                     // per-pixel scalars + addresses of pixel-iteration t
    const int it = t % IPG;
    const int gbase = q_begin + (warp + (t / IPG) * kLmWarps) * 32;
    if (it == 0) {                                               // phase A: one lane per pixel, 32 pixels at once
      __syncwarp();
      const float4 tab_px = (!G2SP && gbase + lane < q_end) ? __ldg(tab + gbase + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
      const PixelScalars ps = pixel_scalars<GEOM>(a, kp, fp, gq, tab_px, conf, gbase + lane, q_end, C4);
      if (G2SP) ps_s[warp][lane][G2SP ? 3 : 0] = make_float4(ps.d0x, ps.d0y, ps.d1x, ps.d1y);
      ps_s[warp][lane][0] = make_float4(ps.ex, ps.wx, ps.sy, ps.ny);
      ps_s[warp][lane][1] = make_float4(ps.tx, ps.ty, ps.om, ps.valid);
      ps_s[warp][lane][2] = make_float4(__int_as_float(ps.off_n), __int_as_float(ps.off_s), __int_as_float(ps.east),
                                        __int_as_float(ps.goff));
      __syncwarp();
    }
    const int src = it * PPW + sub;
    nsc0 = ps_s[warp][src][0];
    nsc1 = ps_s[warp][src][1];
    if (G2SP) nsc2 = ps_s[warp][src][G2SP ? 3 : 0];
    const float4 o = ps_s[warp][src][2];
    const int goff = __float_as_int(o.w);
    east = __float_as_int(o.z);
    gp = goff >= 0 ? grd + goff : zeros;                         // masked pixels read a zero vector: no predication
    s_n = sat + __float_as_int(o.x);
    s_s = sat + __float_as_int(o.y);
  };
  auto load_half = [&](PixelLoads& L, int h) {
    L.g = ld_stream(gp + h * LPP);
    L.nw = ld_cached(s_n + h * LPP); L.ne = ld_cached(s_n + east + h * LPP);
    L.sw = ld_cached(s_s + h * LPP); L.se = ld_cached(s_s + east + h * LPP);
  };

  // Software pipeline at half-pixel granularity: while one half's five 128-bit loads are reduced, the
  // other half's (and the next pixel's first half's) are in flight.
  if (T > 0) { prepare(0); load_half(bufA, 0); load_half(bufB, 1); }
  for (int t = 0; t < T; ++t) {
    sc0 = nsc0; sc1 = nsc1; sc2 = nsc2;
    const bool more = t + 1 < T;
    if (more) prepare(t + 1);
    accumulate(bufA);
    if (more) load_half(bufA, 0);
    accumulate(bufB);
    finish_pixel();
    if (more) load_half(bufB, 1);
  }

  {
    double v[kLmAcc] = {sum2(A_aa), sum2(A_ab), sum2(A_bb), sum2(B_x), sum2(B_y), sum2(C_tt), sum2(S_a), sum2(S_b),
                        sum2(S_t), sum2(G_a), sum2(G_b), sum2(G_t), sum2(SS), sum2(GG), sum2(SG), cnt / (float)LPP};
    lm_reduce_and_solve<GEOM, FULL>(a, lm_tail_of_launch(a), b, v, kp, fp, su, sv, th);
  }
}

// ------------------------------------------------------------------------------------------------------------
// v4 step kernel (S2GP geometries).  What changed against the kernel above and why (B200 measurements):
//  * HBM latency x bandwidth needs ~40 KB of streamed loads in flight per SM; register-staged loads hold 16 KB
//    (2 x 512 B per warp), which pinned v3 at 0.4 of the HBM roofline.  Here every warp owns a ring of NSLOT
//    2-KB chunks in shared memory; lane 0 keeps NSLOT-1 chunks of the ground stream in flight with 1-D bulk async
//    copies (cp.async.bulk -> SASS UBLKCP, completion on a per-slot mbarrier) and re-arms a slot as soon as the warp
//    has reduced it.  Producer and consumer of a ring are the same warp: no cross-warp synchronisation at all.
//  * The FMA pipe issues one packed FFMA2 per two cycles per SM sub-partition, which capped v3 at ~0.85 of the
//    roofline even at 100 % pipe utilisation (19 packed ops per channel pair + 19 per pixel and lane).  Here a lane
//    owns 16 channels of a pixel (the per-pixel work is amortised over twice the channels) and the interpolation
//    uses wx + ex = 1, ny + sy = 1 (16 ops per pair).  Pixels where that identity fails — a corner clamped at the
//    last row / column (jacobian.py:147-166: all weights vanish there) — and masked pixels read their taps from a
//    zero vector instead, which reproduces the reference's zeros exactly.
//  * ncu showed v3 (and a first cut of this kernel) spending more than half of their issue slots on index
//    arithmetic.  Phase A therefore leaves finished 64-bit tap addresses in the per-pixel record, all loop state
//    (ring slot, record pointer, stream position) advances incrementally, and every load uses an immediate offset.
//  * The satellite taps stay on the L1/L2 path (gathers with heavy reuse between neighbouring pixels).
extern __shared__ __align__(128) uint8_t lm_dyn_smem[];

constexpr int kLmIterBytes = 2048;   // one warp pixel-iteration streams PPW pixels x C channels x 4 B = 2 KB for every C
constexpr int kLmChainMaxSteps = 128;  // chained launches: arrival words in the workspace (1 KB)
constexpr int kLmRecBytes = 48;      // per-pixel record: (wx, ny, tx, ty) | (&nw, &sw) | (east bytes, has ground, weight, -)

// Ring slot layout.  A lane reads its 16 bytes of channel quarter k of its pixel with one LDS.128, and a quarter-warp (one
// shared-memory wavefront) holds 8 / LPP pixels whose quarters sit 4C bytes apart: for C <= 64 that is a multiple of
// 128 B (C = 16: of 64 B), so the lanes of a wavefront hit the same banks (ncu, C = 64: 17 M bank conflicts per launch,
// 8 wavefronts per LDS.128 instead of 4).  The chunk is therefore copied as NS sub-copies, sub-copy j shifted by j * C
// bytes, and a quarter-warp takes its pixels from all NS sub-copies: every wavefront covers 128 distinct bytes mod 128.
// Measured on B200 (C = 64, B = 256): the conflicts disappear but the launch is 4 % SLOWER (763 vs 734 us) — the kernel is
// bound by per-warp latency, not by the shared-memory pipe, and the elected lane now issues two copies per chunk.  The
// split therefore stays off (HA_LM_RING_SPLIT = 1 builds it).
#ifndef HA_LM_RING_SPLIT
#define HA_LM_RING_SPLIT 0
#endif
template <int C> constexpr int lm_ring_ns() { return !HA_LM_RING_SPLIT || C > 64 ? 1 : (C == 64 ? 2 : 4); }
constexpr int kLmSlotBytes = kLmIterBytes + 128;          // room for the (NS - 1) * C <= 96 bytes of shift, 128-byte aligned

template <int NSLOT>
constexpr int lm_ring_bytes() {     // rings, mbarriers, 1 KB of zeros, per-warp pixel records
  return kLmWarps * NSLOT * kLmSlotBytes + kLmWarps * NSLOT * 8 + 1024 + kLmWarps * 32 * kLmRecBytes;
}

// Loop-invariant addresses the compiler would otherwise re-derive from %tid / %ctaid inside the loop (it treats them as
// "cheap to rematerialise" under register pressure; ncu showed ~100 such instructions per iteration): make them opaque.
#define HA_KEEP32(x) asm volatile("" : "+r"(x))
#define HA_KEEP64(x) asm volatile("" : "+l"(x))

// once per kernel: the per-warp mbarriers and the zero vector (followed by __syncthreads in the caller)
template <int NSLOT>
__device__ __forceinline__ void lm_v4_smem_init() {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t dyn = smem_u32(lm_dyn_smem);
  const uint32_t bar_w = dyn + kLmWarps * NSLOT * kLmSlotBytes + warp * (NSLOT * 8);
  const uint32_t zero_s = dyn + kLmWarps * NSLOT * kLmSlotBytes + kLmWarps * NSLOT * 8;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_w + s * 8));
    fence_barrier_init();
  }
  asm volatile("st.shared.v2.b32 [%0], {%1, %1};" ::"r"(zero_s + threadIdx.x * 8), "r"(0) : "memory");
}

// One CTA's share of one LM step of sample b: pixels [cta_x * px_per_cta, ...) of the residual.
template <int GEOM, int C, bool FULL, int NSLOT, int PF, bool WEIGHTED, int UNR, bool SCAL>
__device__ __forceinline__ void lm_v4_body(const LmStepArgs& a, const int b, const int cta_x) {
  constexpr int LPP = C / 16;                      // lanes per pixel: every lane owns 4 x 4 channels
  constexpr int PPW = 32 / LPP;                    // pixels processed together by one warp
  constexpr int IPG = LPP;                         // pixel-iterations per 32-pixel group
  constexpr int C4 = C / 4;
  constexpr int NS = lm_ring_ns<C>();              // sub-copies per chunk
  constexpr int PPS = PPW / NS;                    // pixels per sub-copy
  static_assert(C % 16 == 0 && LPP >= 1 && LPP <= 32 && PPW * LPP == 32, "channel count");
  static_assert(GEOM != HA_GEOM_G2SP, "G2SP streams only the visible satellite pixels: it stays on lm_step_kernel");
  static_assert(8 * C <= kLmZeroBytes, "the global zero vector covers two texels (west + east taps), the shared one a ground pixel");

  // the warp index through a shuffle: the compiler then knows that everything derived from it (trip counts, ring and
  // stream positions, the producer's bookkeeping) is warp-uniform and moves it to the uniform datapath: 158 -> 142
  // registers, 17 -> 6 BSSY / BSYNC pairs and 10 fewer branches per two pixel-iterations (C = 64)
  const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int cl = lane % LPP;                       // channel lane
  int sub = lane / LPP;                            // pixel slot within the warp
  if constexpr (NS > 1) {
    // quarter-warp qw holds PPQ pixels; R of them come from each sub-copy (see the ring slot layout above)
    constexpr int PPQ = 8 / LPP, R = PPQ / NS;
    static_assert(PPQ % NS == 0 && R >= 1, "sub-copy split");
    const int i = sub % PPQ, qw = sub / PPQ;
    sub = (i / R) * PPS + qw * R + (i % R);
  }
  const int P = (a.H - a.row0) * a.W;              // the residual lives on the bottom half (models_kitti.py:1195-1199; row0 = H/2)
  const int q_begin = cta_x * a.px_per_cta;
  const int q_end = min(P, q_begin + a.px_per_cta);

  // through L2: with chained launches the previous step's CTA wrote it while this kernel was already running
  const float su = __ldcg(a.pose + b * 3 + 0), sv = __ldcg(a.pose + b * 3 + 1), th = __ldcg(a.pose + b * 3 + 2);
  KittiPose kp;
  FordPose fp;
  G2spPose gq;                                     // unused (pixel_scalars signature)
  if (GEOM == HA_GEOM_KITTI) kp = kitti_pose(a, su, sv, th);
  else fp = ford_pose(a, b, su, sv, th);

  const size_t px_base = (size_t)b * a.H * a.W + (size_t)a.row0 * a.W;
  const float4* grd0 = reinterpret_cast<const float4*>(a.grd) + px_base * C4;              // streamed (bulk copies)
  const char* sat_b = reinterpret_cast<const char*>(a.sat) + (size_t)b * a.A * a.A * C * 4;  // gathered
  const float4* tab = a.table + (size_t)a.row0 * a.W;
  const float* conf = a.conf ? a.conf + px_base : nullptr;

  // dynamic shared memory: [warp][NSLOT] ring slots, [warp][NSLOT] mbarriers, 1 KB of zeros (masked ground pixels),
  // [warp][32] pixel records: phase A (one lane per pixel) writes 32 of them, the pixel's channel lanes read them back
  const uint32_t dyn = smem_u32(lm_dyn_smem);
  const uint32_t ring_w = dyn + warp * (NSLOT * kLmSlotBytes);
  const uint32_t bar_w = dyn + kLmWarps * NSLOT * kLmSlotBytes + warp * (NSLOT * 8);
  const uint32_t zero_s = dyn + kLmWarps * NSLOT * kLmSlotBytes + kLmWarps * NSLOT * 8;
  const uint32_t ps_w = zero_s + 1024 + warp * (32 * kLmRecBytes);
  uint32_t ring_lane = ring_w + sub * (C * 4) + (sub / PPS) * C + cl * 16;   // this lane's slice of slot 0
  uint32_t zero_lane = zero_s + cl * 16;
  uint32_t ps_lane = ps_w + sub * kLmRecBytes;
  uint32_t ps_wr = ps_w + lane * kLmRecBytes;                      // the record this lane writes in phase A
  const uint32_t cl16 = cl * 16;
  uint32_t ring_bar = ring_w, bar_keep = bar_w;                    // opaque copies used inside the loop
  uint64_t grd_w = reinterpret_cast<uint64_t>(grd0 + (size_t)(q_begin + warp * 32) * C4);   // first chunk of this warp
  HA_KEEP32(ring_lane); HA_KEEP32(zero_lane); HA_KEEP32(ps_lane); HA_KEEP32(ps_wr); HA_KEEP32(ring_bar); HA_KEEP32(bar_keep);
  HA_KEEP64(grd_w);

  const int n_groups = (q_end - q_begin + 31) / 32;
  const int my_groups = (n_groups > warp) ? (n_groups - warp + kLmWarps - 1) / kLmWarps : 0;
  const int T = my_groups * IPG;                   // pixel-iterations of this warp

  // ---- ring producer (lane 0).  Chunk u = pixel-iteration u of this warp; `iss_px` is its first pixel.
  // `iss_left` = pixels from the chunk's first pixel to q_end (<= 0: nothing to copy), `iss_off` = its byte offset from grd_w.
  int iss_left = q_end - (q_begin + warp * 32), iss_it = 0, to_issue = T;
  uint32_t iss_off = 0;
  auto issue_next = [&](uint32_t slot) {           // arm `slot` with the next chunk of the stream, then advance
    uint32_t leader = 0;
    if (iss_left > 0)                              // warp-uniform; the warp is converged here (after __syncwarp)
      asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(leader));
    if (leader) {
      const uint32_t bar = bar_keep + slot * 8, npx = (uint32_t)min(iss_left, PPW);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(npx * (C * 4)) : "memory");
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        if (j == 0 || npx > (uint32_t)(j * PPS)) {
          const uint32_t bytes = (NS == 1 ? npx : min(npx - j * PPS, (uint32_t)PPS)) * (C * 4);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(ring_bar + slot * kLmSlotBytes + j * (PPS * C * 4 + C)), "l"(grd_w + iss_off + j * (PPS * C * 4)),
                         "r"(bytes), "r"(bar) : "memory");
        }
      }
    }
    iss_off += kLmIterBytes; iss_left -= PPW; --to_issue;
    if (++iss_it == IPG) { iss_it = 0; iss_off += (kLmWarps - 1) * 32 * C * 4; iss_left -= (kLmWarps - 1) * 32; }
  };
  for (int j = 0; j < NSLOT && to_issue > 0; ++j) issue_next(j);

  // running sums over this lane's pixels and channels (two partial sums per register pair)
  f32x2 A_aa = 0, A_ab = 0, A_bb = 0, B_x = 0, B_y = 0, C_tt = 0;
  f32x2 S_a = 0, S_b = 0, S_t = 0, G_a = 0, G_b = 0, G_t = 0, SS = 0, GG = 0, SG = 0;
  float cnt = 0.f;                                 // in-range pixels, counted by the phase-A lanes
  f32x2 p_aa = 0, p_ab = 0, p_bb = 0, p_sa = 0, p_sb = 0, p_ga = 0, p_gb = 0;   // per-pixel channel sums
  float4 sc = make_float4(0, 0, 0, 0);             // (wx, ny, tx, ty) of the pixel being reduced
  float om = 1.f;
  float rs[12] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // SCAL: the twelve pose sums as scalars

  auto accumulate = [&](const PixelLoads& L) {
    const f32x2 wx2 = dup2(sc.x), ny2 = dup2(sc.y);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const f32x2 nw = h ? L.nw.hi : L.nw.lo, ne = h ? L.ne.hi : L.ne.lo;
      const f32x2 sw = h ? L.sw.hi : L.sw.lo, se = h ? L.se.hi : L.se.lo;
      const f32x2 g = h ? L.g.hi : L.g.lo;
      const f32x2 dn = sub2(ne, nw), ds = sub2(se, sw);          // east - west on the north / south row
      const f32x2 top = fma2(dn, wx2, nw), bot = fma2(ds, wx2, sw);   // rows interpolated in x (ex = 1 - wx)
      const f32x2 db = sub2(bot, top);                           // d/dy  (jacobian.py:192-193)
      const f32x2 da = fma2(sub2(ds, dn), ny2, dn);              // d/dx = sy dn + ny ds  (:190-191, sy = 1 - ny)
      const f32x2 s = fma2(db, ny2, top);                        // sy top + ny bot  (:174-186)
      acc2(p_aa, da, da); acc2(p_ab, da, db); acc2(p_bb, db, db);
      acc2(p_sa, s, da); acc2(p_sb, s, db); acc2(p_ga, g, da); acc2(p_gb, g, db);
      acc2(SS, s, s);
      if (FULL) { acc2(SG, s, g); acc2(GG, g, g); }              // |g|^2 is pose independent: cached after the first visit
    }
  };
  auto finish_pixel = [&]() {
    if (WEIGHTED) {                                  // compile-time: a run-time branch here costs ~30 register moves per pixel
      const f32x2 om2 = dup2(om);
      p_aa = mul2(p_aa, om2); p_ab = mul2(p_ab, om2); p_bb = mul2(p_bb, om2);
      p_sa = mul2(p_sa, om2); p_sb = mul2(p_sb, om2); p_ga = mul2(p_ga, om2); p_gb = mul2(p_gb, om2);
    }
    // d(u,v)/dsu and /dsv are per-sample constants, only d/dtheta = (tx, ty) varies per pixel, so
    // J^T J splits into sum(G), sum(G t), sum(t^T G t) and the last CTA applies the constant rows.
    if (SCAL) {
      // the channel-pair halves are added first and the per-pixel fold runs on scalars: 26 one-cycle operations instead of
      // 19 packed ones (two FMA-pipe cycles each), and the twelve running sums take 12 registers instead of 24
      const float aa = sum2(p_aa), ab = sum2(p_ab), bb = sum2(p_bb), sa = sum2(p_sa), sb = sum2(p_sb), ga = sum2(p_ga), gb = sum2(p_gb);
      const float tx = sc.z, ty = sc.w;
      const float gx = fmaf(ab, ty, aa * tx), gy = fmaf(bb, ty, ab * tx);
      rs[0] += aa; rs[1] += ab; rs[2] += bb; rs[3] += gx; rs[4] += gy; rs[5] = fmaf(ty, gy, fmaf(tx, gx, rs[5]));
      rs[6] += sa; rs[7] += sb; rs[8] = fmaf(sb, ty, fmaf(sa, tx, rs[8]));
      rs[9] += ga; rs[10] += gb; rs[11] = fmaf(gb, ty, fmaf(ga, tx, rs[11]));
      p_aa = p_ab = p_bb = p_sa = p_sb = p_ga = p_gb = 0ull;
      return;
    }
    const f32x2 tx2 = dup2(sc.z), ty2 = dup2(sc.w);
    inc2(A_aa, p_aa); inc2(A_ab, p_ab); inc2(A_bb, p_bb);
    const f32x2 gx = fma2(p_ab, ty2, mul2(p_aa, tx2)), gy = fma2(p_bb, ty2, mul2(p_ab, tx2));
    inc2(B_x, gx); inc2(B_y, gy);
    acc2(C_tt, tx2, gx); acc2(C_tt, ty2, gy);
    inc2(S_a, p_sa); inc2(S_b, p_sb); acc2(S_t, p_sa, tx2); acc2(S_t, p_sb, ty2);
    inc2(G_a, p_ga); inc2(G_b, p_gb); acc2(G_t, p_ga, tx2); acc2(G_t, p_gb, ty2);
    p_aa = p_ab = p_bb = p_sa = p_sb = p_ga = p_gb = 0ull;
  };

  // ---- consumer state, advanced incrementally by prepare()
  int nxt_px = q_begin + warp * 32, nxt_it = 0;    // first pixel / position in its group of the NEXT pixel-iteration
  uint32_t slot = 0, parity = 0;                   // ring slot and mbarrier phase of the next pixel-iteration
  uint32_t ps_rd = ps_lane;                        // record of the next pixel-iteration
  uint32_t ps_cur = ps_lane;                       // record of the pixel-iteration whose loads were issued last
  const char *p_nw = nullptr, *p_sw = nullptr;     // this lane's 16 bytes of the north-west / south-west taps
  uint32_t g_addr = zero_lane;                     // this lane's 16 bytes of the ground vector (shared memory)

  float4 tab_next = make_float4(0.f, 0.f, 0.f, 0.f);   // table entry of this lane's pixel in the group phase A handles next
  auto load_tab = [&](int px0) { tab_next = (px0 + lane < q_end) ? __ldg(tab + px0 + lane) : make_float4(0.f, 0.f, 0.f, 0.f); };
  auto phase_a = [&](int px0) {                    // one lane per pixel, 32 pixels [px0, px0 + 32) at once
    __syncwarp();
    const PixelScalars ps = pixel_scalars<GEOM>(a, kp, fp, gq, tab_next, conf, px0 + lane, q_end, C4);
    load_tab(px0 + kLmWarps * 32);                 // table entry for the phase A after this one: its latency is off the path
    // taps are read unless the sample point is masked or a corner was clamped (then all weights vanish)
    const bool taps = ps.valid != 0.f && (ps.ex + ps.wx == 1.f) && (ps.sy + ps.ny == 1.f);
    const char* pn = taps ? sat_b + (size_t)ps.off_n * 16 : reinterpret_cast<const char*>(a.zeros);
    const char* psw = taps ? sat_b + (size_t)ps.off_s * 16 : reinterpret_cast<const char*>(a.zeros);
    const uint32_t rec = ps_wr;
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(rec), "f"(ps.wx), "f"(ps.ny), "f"(ps.tx), "f"(ps.ty) : "memory");
    asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(rec + 16), "l"(pn), "l"(psw) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rec + 32), "r"(taps ? ps.east * 16 : 0),
                 "r"(ps.goff >= 0 ? 1 : 0), "r"(__float_as_int(ps.om)), "r"(0) : "memory");
    cnt += ps.valid;
    if (PF != 0 && taps) {
      // The pixel's two tap rows (nw|ne and sw|se are adjacent texels) are prefetched now: first touches of a satellite
      // texel miss to DRAM (~1000 cycles), far more than the two-quarter lead of the tap loads themselves.  All but the
      // first pixel-iteration of the group get at least one iteration of lead.  (Running phase A a whole group ahead
      // with double-buffered records measured slower: 794 vs 763 us at C = 64, B = 256.)
      constexpr int kLines = (2 * C * 4 + 127) / 128;
#pragma unroll
      for (int k = 0; k < kLines; ++k) {
        if (PF == 1) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pn + k * 128));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(psw + k * 128));
        } else {
          asm volatile("prefetch.global.L1 [%0];" ::"l"(pn + k * 128));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(psw + k * 128));
        }
      }
    }
    __syncwarp();
  };

  auto prepare = [&]() {                           // addresses of the next pixel-iteration's loads
    if (nxt_px < q_end) {                          // its chunk has landed (chunks past the end are never armed)
      const uint32_t bar = bar_keep + slot * 8;
      asm volatile(
          "{\n"
          ".reg .pred p;\n"
          "LM_WAIT:\n"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
          "@p bra LM_DONE;\n"
          "bra LM_WAIT;\n"
          "LM_DONE:\n"
          "}\n" ::"r"(bar), "r"(parity) : "memory");
    }
    if (nxt_it == 0) {                             // entering a 32-pixel group: phase A
      phase_a(nxt_px);
      ps_rd = ps_lane;
    }
    uint64_t an, as_;
    uint32_t has_g;
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(an), "=l"(as_) : "r"(ps_rd + 16));
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(has_g) : "r"(ps_rd + 36));
    // the east taps sit one texel (4 C bytes) after the west ones: taps are read only when xe - xw == 1 (phase A), and
    // the zero vector that stands in otherwise is two texels long, so east is an immediate offset, not a pointer
    p_nw = reinterpret_cast<const char*>(an) + cl16; p_sw = reinterpret_cast<const char*>(as_) + cl16;
    g_addr = has_g ? ring_lane + slot * kLmSlotBytes : zero_lane;  // masked ground pixels read zeros
    ps_cur = ps_rd;
    // advance to the pixel-iteration after this one
    ps_rd += PPW * kLmRecBytes; nxt_px += PPW;
    if (++nxt_it == IPG) { nxt_it = 0; nxt_px += (kLmWarps - 1) * 32; }
    if (++slot == NSLOT) { slot = 0; parity ^= 1; }
  };
  auto load_quarter = [&](PixelLoads& L, int k) {  // k-th channel quarter: immediate offsets k * C bytes
    L.g = ld_ring(g_addr + k * C);
    L.nw = ld_cached(reinterpret_cast<const float4*>(p_nw + k * C)); L.ne = ld_cached(reinterpret_cast<const float4*>(p_nw + k * C + 4 * C));
    L.sw = ld_cached(reinterpret_cast<const float4*>(p_sw + k * C)); L.se = ld_cached(reinterpret_cast<const float4*>(p_sw + k * C + 4 * C));
  };

  // Software pipeline at quarter-pixel granularity (ping-pong buffers A / B): the five 128-bit loads of a quarter
  // are issued two quarters before they are reduced.
  PixelLoads bufA, bufB;
  uint32_t slot_cur = 0;
  if (T > 0) {
    load_tab(nxt_px);
    prepare(); load_quarter(bufA, 0); load_quarter(bufB, 1);
  }
  // (Peeling the last iteration so that `more` is a compile-time constant removes ~30 instructions per iteration
  // but measured 6 % slower — 813 vs 765 us at C = 64, B = 256 — so the loop keeps the run-time predicate.)
#pragma unroll UNR
  for (int t = 0; t < T; ++t) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sc.x), "=f"(sc.y), "=f"(sc.z), "=f"(sc.w) : "r"(ps_cur));
    if (WEIGHTED) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(om) : "r"(ps_cur + 40));
    const bool more = t + 1 < T;
    accumulate(bufA); load_quarter(bufA, 2);
    accumulate(bufB); load_quarter(bufB, 3);
    const uint32_t slot_done = slot_cur;
    if (more) { slot_cur = slot; prepare(); }
    accumulate(bufA);
    if (more) load_quarter(bufA, 0);
    accumulate(bufB);
    finish_pixel();
    // this pixel-iteration's chunk has been reduced: every lane's reads of the slot have returned (their values were
    // just consumed), so an elected lane re-arms the slot with the chunk NSLOT ahead
    __syncwarp();
    if (to_issue > 0) issue_next(slot_done);
    if (more) load_quarter(bufB, 1);
  }

  {
    double v[kLmAcc] = {sum2(A_aa), sum2(A_ab), sum2(A_bb), sum2(B_x), sum2(B_y), sum2(C_tt), sum2(S_a), sum2(S_b),
                        sum2(S_t), sum2(G_a), sum2(G_b), sum2(G_t), sum2(SS), sum2(GG), sum2(SG), cnt};
    if (SCAL) {
#pragma unroll
      for (int i = 0; i < 12; ++i) v[i] = rs[i];
    }
    lm_reduce_and_solve<GEOM, FULL>(a, lm_tail_of_launch(a), b, v, kp, fp, su, sv, th);
  }
}

// Chained launches (ha_lm_run).  Per-step launches leave the machine idle between the first CTA that finishes a step
// and the last one (B = 32: 15 launches of 40-110 us, 0.44 of the HBM roofline against 0.61 at B = 256), although step
// k + 1 of a sample needs nothing but ITS OWN pose of step k.  So every step kernel of a run is launched with
// programmatic stream serialization and signals `launch_dependents` at once: the next step's CTAs become resident as soon
// as every CTA of this step has started and SM slots free up.  They never execute griddepcontrol.wait; a CTA of
// sample b waits for `done[b] >= step` instead, which the CTA that solves the sample's previous step releases after
// writing the pose.  A kernel starts only when all CTAs of its predecessor are resident, and those wait only on CTAs
// of still earlier kernels that are resident too, so the chain cannot deadlock; a waiter gives up after ~1 s
// (HA_STATUS_TIMEOUT) rather than hang the GPU.
__device__ __forceinline__ void lm_chain_enter(const LmStepArgs& a, int b) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (a.done != nullptr && threadIdx.x == 0) {
    int seen, spins = 0;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(a.done + b) : "memory");
      if (seen >= a.step_index) break;
      __nanosleep(64);
      if (++spins > (1 << 21)) { atomicOr(a.status, HA_STATUS_TIMEOUT); break; }   // ~1 s
    }
  }
}

template <int GEOM, int C, bool FULL, int NSLOT, int MINB, int PF, bool WEIGHTED, int UNR = 1, bool SCAL = false>
__global__ void __launch_bounds__(kLmThreads, MINB) lm_step_v4_kernel(const LmStepArgs a) {
  lm_chain_enter(a, blockIdx.y);
  lm_v4_smem_init<NSLOT>();
  __syncthreads();
  lm_v4_body<GEOM, C, FULL, NSLOT, PF, WEIGHTED, UNR, SCAL>(a, blockIdx.y, blockIdx.x);
}

// Kernel selection: HaLmParams.kernel_variant 0 (default) = lm_step_v4_kernel with a 4-slot x 2 KB ring per warp, tap rows
// prefetched to L1, 4 CTAs per SM (128 registers), pixel loop unrolled by two; ha_lm_run chains its launches (above).  1 = lm_step_kernel
// (register-staged ground stream; the validation twin, and always the kernel for G2SP); 2 = the default kernel without
// chaining (one stream-ordered launch per step: the A/B twin of the chain).  Other ring depths (3, 5, 6, 8 slots),
// L2-only prefetch, no prefetch, 2 and 3 CTAs per SM and the non-unrolled loop were measured and dropped.
template <typename K>
static int lm_configure_smem(K kern, int smem, int minb) {
  // function attributes are per (device function, device): one bit per device ordinal, per instantiation
  static std::atomic<unsigned long long> configured{0};
  int dev = 0;
  HA_CUDA_TRY(cudaGetDevice(&dev));
  const unsigned long long bit = 1ull << (dev & 63);
  if (!(configured.load(std::memory_order_acquire) & bit)) {
    HA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    // shared-memory carve-out: exactly what `minb` resident CTAs need (ring + static + 1 KB reserved each); the rest of
    // the 228 KB stays L1 for the satellite taps
    cudaFuncAttributes fa;
    HA_CUDA_TRY(cudaFuncGetAttributes(&fa, kern));
    const int want = minb * (smem + (int)fa.sharedSizeBytes + 1024);
    const int pct = want >= 228 * 1024 ? 100 : (want * 100 + 228 * 1024 - 1) / (228 * 1024);
    HA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    configured.fetch_or(bit, std::memory_order_release);
  }
  return HA_OK;
}

template <int GEOM, int C, bool FULL, int NSLOT, int MINB, int PF, bool WEIGHTED, int UNR, bool SCAL = false>
static int launch_v4w(dim3 grid, cudaStream_t st, const LmStepArgs& a) {
  auto kern = lm_step_v4_kernel<GEOM, C, FULL, NSLOT, MINB, PF, WEIGHTED, UNR, SCAL>;
  constexpr int smem = lm_ring_bytes<NSLOT>();
  const int rc = lm_configure_smem(kern, smem, MINB);
  if (rc != HA_OK) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(kLmThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  // steps after the first of a chained run may start before their predecessor has finished (lm_chain_enter)
  if (a.done != nullptr && a.step_index > 0) { cfg.attrs = attr; cfg.numAttrs = 1; }
  HA_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, a));
  return HA_OK;
}

template <int GEOM, int C, bool FULL, int NSLOT, int MINB, int PF, int UNR = 1, bool SCAL = false>
static int launch_v4(dim3 grid, cudaStream_t st, const LmStepArgs& a) {
  return a.using_weight ? launch_v4w<GEOM, C, FULL, NSLOT, MINB, PF, true, UNR, SCAL>(grid, st, a)
                        : launch_v4w<GEOM, C, FULL, NSLOT, MINB, PF, false, UNR, SCAL>(grid, st, a);
}

template <int GEOM, int C, bool FULL>
static int launch_variant(dim3 grid, cudaStream_t st, const LmStepArgs& a) {
  if constexpr (GEOM == HA_GEOM_G2SP) {
    lm_step_kernel<GEOM, C, FULL><<<grid, kLmThreads, 0, st>>>(a);
    return HA_OK;
  } else {
    if (a.variant == 1) { lm_step_kernel<GEOM, C, FULL><<<grid, kLmThreads, 0, st>>>(a); return HA_OK; }
    // Measured on B200 with the warp-uniform bookkeeping (whole chained loop, fraction of the HBM peak, B = 256 / 128 / 32):
    // 3 CTAs / SM (142 registers) 0.750 / 0.719 / 0.611; 4 CTAs / SM (128 registers, no spills in the unweighted kernels,
    // 8-16 bytes in the weighted FULL ones) 0.760 / 0.757 / 0.632; 4 CTAs / SM with a 3-slot ring 0.725 / 0.743 / 0.636.
    // Tap-row prefetch at 4 CTAs / SM (B = 256 / 32, two rounds in one process): to L1 0.762, 0.696 / 0.620, 0.635; to L2
    // only 0.692, 0.673 / 0.620, 0.632; none 0.678, 0.656 / 0.596, 0.600.
#ifdef HA_LM_DEV_VARIANTS      // A/B builds only (tools/lm_profile_batches.py)
    if (a.variant == 3) return launch_v4<GEOM, C, FULL, 4, 4, 1, 2, true>(grid, st, a);   // tap rows prefetched to L2 only
    if (a.variant == 4) return launch_v4<GEOM, C, FULL, 4, 4, 0, 2, true>(grid, st, a);   // no tap prefetch
#endif
    return launch_v4<GEOM, C, FULL, 4, 4, 2, 2, true>(grid, st, a);
  }
}

template <int GEOM, bool FULL>
static int launch_by_channels(int C, dim3 grid, cudaStream_t st, const LmStepArgs& a) {
  int rc;
  switch (C) {
    case 256: rc = launch_variant<GEOM, 256, FULL>(grid, st, a); break;
    case 128: rc = launch_variant<GEOM, 128, FULL>(grid, st, a); break;
    case 64: rc = launch_variant<GEOM, 64, FULL>(grid, st, a); break;
    case 32: rc = launch_variant<GEOM, 32, FULL>(grid, st, a); break;
    case 16: rc = launch_variant<GEOM, 16, FULL>(grid, st, a); break;
    default: return HA_EINVAL;
  }
  if (rc != HA_OK) return rc;
  count_launches(1);
  return check_launch("lm_step_kernel");
}

// Workspace layout: [partials B x 256 x 16 fp64][tickets B x u32, padded to 256 B][done B x s32, padded to 256 B]
// [step region 1 KB: 64 step words (u64) for chained launches][zero vector 2 KB]
// [|g|^2 cache HA_MAX_LEVELS x B fp64][Adam moments B x 6 fp32].  Tickets ... zero vector are contiguous: one kernel clears them.
struct LmWs {
  double* partial; uint32_t* ticket; int* done; unsigned long long* step_word; const float4* zeros;
  double* gg; float* adam_mv;
  size_t clear_words;      // u32 words from `ticket` that must be zero when a run starts
  size_t total;
};
static LmWs lm_ws_carve(void* ws, int B) {
  LmWs w;
  char* p = reinterpret_cast<char*>(ws);
  const size_t part = (size_t)B * kLmMaxCtasPerSample * kLmAcc * sizeof(double);
  const size_t tick = ((size_t)B * sizeof(uint32_t) + 255) / 256 * 256;
  const size_t step = 1024;
  static_assert(kLmChainMaxSteps * 8 <= step, "step words");
  w.partial = reinterpret_cast<double*>(p);
  w.ticket = reinterpret_cast<uint32_t*>(p + part);
  w.done = reinterpret_cast<int*>(p + part + tick);
  w.step_word = reinterpret_cast<unsigned long long*>(p + part + 2 * tick);
  w.zeros = reinterpret_cast<const float4*>(p + part + 2 * tick + step);
  w.gg = reinterpret_cast<double*>(p + part + 2 * tick + step + kLmZeroBytes);
  w.clear_words = (2 * tick + step + kLmZeroBytes) / 4;
  const size_t gg = (size_t)HA_MAX_LEVELS * B * sizeof(double);
  w.adam_mv = reinterpret_cast<float*>(p + part + 2 * tick + step + kLmZeroBytes + gg);     // [B][6], HA_OPT_ADAM only
  w.total = part + 2 * tick + step + kLmZeroBytes + gg + (size_t)B * 6 * sizeof(float);
  return w;
}
static size_t lm_ws_bytes(int B) { return lm_ws_carve(nullptr, B).total; }

// Work split: a CTA takes px_per_cta consecutive residual pixels of one sample, a multiple of 32 * warps so that every warp
// gets the same number of 32-pixel groups (no barrier skew); never more than kLmMaxCtasPerSample CTAs per sample.
// The grid runs in rounds of `slots` = 148 SMs x resident CTAs, and a launch costs about rounds x (units per CTA + a fixed
// per-CTA overhead): B200, B = 256 showed 3.46 / 4.04 / 4.04 waves with the old "about 12 CTAs per SM" rule, i.e. a fifth
// round that was 4 % full (profiles/r02_lm_full.csv: launch__waves_per_multiprocessor).  Pick the split that minimises that cost.
//
// Chained launches (ha_lm_run) have no rounds: the next step's CTAs fill the slots as they free up, and what counts is
// (a) the sample's critical path (15 x (one CTA share + the fixed latency between two steps of a sample: reduction,
// ticket, solve, release / acquire, prologue, first ring fill)), which small batches are bound by, against (b) the
// per-CTA cost that fine splits multiply, the tail of one share at the end of the run, and L2: with few CTAs per
// sample more samples are resident at once and their satellite footprints stop fitting.  Measured on B200 (whole
// loop, KITTI pyramid, fraction of the HBM peak; shares in 128-pixel units at levels 0 / 1 / 2):
//   B = 32 : (1,5,13) 0.533  (1,4,10) 0.562  (1,3,8) 0.570  (1,3,6) 0.567  (1,2,4) 0.546  (2,8,16) 0.422
//   B = 64 : (1,5,13) 0.630  (1,4,10) 0.623  (1,3,8) 0.614  (4,11,16) 0.571  (8,32,32) 0.347
//   B = 128: (4,8,22) 0.671  (4,11,16) 0.671  (4,11,32) 0.668  (2,5,10) 0.648  (8,32,32) 0.562
//   B = 256: (8,32,32) 0.699  (8,32,52) 0.696  (8,32,64) 0.695  (4,32,37) 0.693  (8,32,16) 0.690  (2,13,10) 0.664
//            (8,64,37) 0.613 and (16,32,37) 0.652: ONE CTA per sample at a level puts the whole share on the critical path
// i.e. the best split keeps about 2.5 CTAs per slot in a step, at least two CTAs per sample, shares of at most 64 units.
static int choose_px_per_cta(int B, int P, int resident_per_sm, bool chained) {
  const int unit = kLmWarps * 32;
  const int units = (P + unit - 1) / unit;                     // whole-CTA units in one sample
  const long long slots = (long long)kNumSMs * resident_per_sm;
  double best_cost = 1e300;
  int best_upc = units;
  for (int upc = 1; upc <= units; ++upc) {
    const int n = (units + upc - 1) / upc;
    if (n > kLmMaxCtasPerSample) continue;
    const long long total = (long long)n * B;
    const long long rounds = (total + slots - 1) / slots;
    const double cost = (double)rounds * ((double)upc + 0.5);   // 0.5 unit: prologue (barriers, pose constants) + reduction tail
    if (cost < best_cost - 1e-9) { best_cost = cost; best_upc = upc; }
  }
  if (chained) {
    const long long n_target = std::max<long long>(2, (5 * slots / 2 + B - 1) / B);
    best_upc = (int)std::min<long long>(64, std::max<long long>(1, (units + n_target - 1) / n_target));
    while ((units + best_upc - 1) / best_upc > kLmMaxCtasPerSample) ++best_upc;
  }
  return best_upc * unit;
}

// Validates one (level, step) and fills its argument block and grid; shared by the per-step launches and the loop kernel.
static int lm_step_args(const HaLmParams* p, int level, const HaLevel* sat, const HaLevel* grd, const float* grd_conf,
                        const float* ground_table, const float* extrinsics, float* pose, const float* reset_uv,
                        float* stats, float* traj_step, int traj_stride, uint32_t* status, void* ws, size_t ws_bytes,
                        int B, int iter, int chain_step, LmStepArgs& a, dim3& grid) {
  const bool g2sp_nn = p && p->geometry == HA_GEOM_G2SP_NN;
  const bool g2sp = p && (p->geometry == HA_GEOM_G2SP || g2sp_nn);
  if (!p || !sat || !grd || !pose || !status || !ws || (!ground_table && !g2sp)) return HA_EINVAL;
  if (level < 0 || level >= HA_MAX_LEVELS) return HA_EINVAL;
  if (sat->C != grd->C || sat->H != sat->W || (grd->H & 1)) return HA_EINVAL;
  if (p->dof < 1 || p->dof > 3) return HA_EINVAL;
#ifdef HA_LM_DEV_VARIANTS
  if (p->kernel_variant < 0 || p->kernel_variant > 4 || p->reserved != 0) return HA_EINVAL;
#else
  if (p->kernel_variant < 0 || p->kernel_variant > 2 || p->reserved != 0) return HA_EINVAL;
#endif
  if (p->optimizer < HA_OPT_LM || p->optimizer > HA_OPT_GN || (p->full_height != 0 && p->full_height != 1)) return HA_EINVAL;
  const bool first_order = p->optimizer == HA_OPT_SGD || p->optimizer == HA_OPT_ADAM;
  // the ablation update rules exist for the models that define them: SGD / ADAM in LM_S2GP, GN in LM_S2GP_Ford; they
  // work on the L2-normalised features (SGD / ADAM: both branches, GN: the ground branch): HaLevel.scale must carry the
  // U-Net's 1 / ||x|| there (NULL = the data is already normalised)
  if (p->optimizer != HA_OPT_LM && g2sp) return HA_EINVAL;
  if (first_order && (p->using_weight || p->dof != 3)) return HA_EINVAL;
  if (p->optimizer == HA_OPT_GN && p->dof != 3) return HA_EINVAL;
  if (p->dof == 3 && !reset_uv && !g2sp && !first_order) return HA_EINVAL;
  if (g2sp && p->dof != 3) return HA_EINVAL;
  if (g2sp && !g2sp_nn && (!extrinsics || p->ori_grd_h <= 0 || p->ori_grd_w <= 0)) return HA_EINVAL;
  if (p->using_weight && !grd_conf) return HA_EINVAL;
  if (p->geometry == HA_GEOM_FORD && !extrinsics) return HA_EINVAL;
  if (ws_bytes < lm_ws_bytes(B)) return HA_ENOSPACE;
  if (((uintptr_t)sat->data | (uintptr_t)grd->data | (uintptr_t)ground_table) & 15) return HA_EINVAL;
  if (sat->H * sat->W >= (1 << 24) || grd->H * grd->W >= (1 << 24)) return HA_EINVAL;

  a = LmStepArgs{};
  a.sat = sat->data; a.grd = grd->data; a.sat_scale = sat->scale; a.grd_scale = grd->scale;
  a.conf = grd_conf; a.table = reinterpret_cast<const float4*>(ground_table); a.extr = extrinsics;
  a.pose = pose; a.reset_uv = reset_uv; a.stats = stats; a.traj = traj_step; a.traj_stride = traj_stride;
  a.status = status;
  const LmWs w = lm_ws_carve(ws, B);
  a.partial = w.partial; a.ticket = w.ticket; a.zeros = w.zeros;
  // chained run (ha_lm_run): the steps overlap, so each has its own arrival word and waits on the samples' done flags
  a.step_word = w.step_word + (chain_step >= 0 ? chain_step : 0);
  a.done = chain_step >= 0 ? w.done : nullptr;
  a.step_index = chain_step >= 0 ? chain_step : 0;
  a.gg_cache = w.gg + (size_t)level * B;
  a.B = B; a.A = sat->H; a.H = grd->H; a.W = grd->W;
  a.dof = p->dof; a.using_weight = p->using_weight; a.use_hessian = p->use_hessian;
  a.rot = p->rotation_range; a.lat = p->shift_range_lat; a.lon = p->shift_range_lon;
  a.mpp = p->meter_per_pixel[level]; a.inv_mpp = p->inv_meter_per_pixel[level]; a.center = p->sat_center[level];
  for (int i = 0; i < 3; ++i) a.damping[i] = p->damping[i];
  a.ori_h = p->ori_grd_h; a.ori_w = p->ori_grd_w;
  a.variant = p->kernel_variant;
  a.row0 = p->full_height ? 0 : grd->H / 2;
  a.optimizer = p->optimizer;
  a.adam_t = iter * p->adam_level_mult + level;
  a.adam_b1 = p->adam_beta1; a.adam_b2 = p->adam_beta2;
  a.adam_mv = w.adam_mv;
  a.g2sp_nn = g2sp_nn ? 1 : 0;
  const int P = g2sp ? sat->H * sat->W : (grd->H - a.row0) * grd->W;
  a.px_per_cta = choose_px_per_cta(B, P, 4, chain_step >= 0);   // resident CTAs per SM of the kernel that runs
  a.grd_C = grd->C;
  grid = dim3((P + a.px_per_cta - 1) / a.px_per_cta, B);
  return HA_OK;
}

static int lm_step_impl(const HaLmParams* p, int level, const HaLevel* sat, const HaLevel* grd, const float* grd_conf,
                        const float* ground_table, const float* extrinsics, float* pose, const float* reset_uv,
                        float* stats, float* traj_step, int traj_stride, uint32_t* status, void* ws, size_t ws_bytes,
                        int B, bool full, int iter, int chain_step, cudaStream_t st) {
  LmStepArgs a;
  dim3 grid;
  const int rc = lm_step_args(p, level, sat, grd, grd_conf, ground_table, extrinsics, pose, reset_uv, stats, traj_step,
                              traj_stride, status, ws, ws_bytes, B, iter, chain_step, a, grid);
  if (rc != HA_OK) return rc;
  const bool g2sp = p->geometry == HA_GEOM_G2SP || p->geometry == HA_GEOM_G2SP_NN;
  if (p->geometry == HA_GEOM_KITTI)
    return full ? launch_by_channels<HA_GEOM_KITTI, true>(grd->C, grid, st, a) : launch_by_channels<HA_GEOM_KITTI, false>(grd->C, grid, st, a);
  if (p->geometry == HA_GEOM_FORD)
    return full ? launch_by_channels<HA_GEOM_FORD, true>(grd->C, grid, st, a) : launch_by_channels<HA_GEOM_FORD, false>(grd->C, grid, st, a);
  if (g2sp) return launch_by_channels<HA_GEOM_G2SP, false>(grd->C, grid, st, a);
  return HA_EINVAL;
}

// start of a run: clears the tickets, the step word and the zero vector (contiguous) and the caller's status word
__global__ void lm_begin_kernel(uint32_t* p, int n_words, uint32_t* status) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += gridDim.x * blockDim.x) p[i] = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) *status = 0;
}
static void lm_begin(void* ws, int B, uint32_t* status, cudaStream_t st) {
  const LmWs w = lm_ws_carve(ws, B);
  lm_begin_kernel<<<(int)((w.clear_words + 255) / 256), 256, 0, st>>>(w.ticket, (int)w.clear_words, status);
  count_launches(1);
}

// ------------------------------------------------------------------------------------------------------------
// Optimizer 'NN' (LM_S2GP.NN_update, models_kitti.py:1043-1054; RNNs.NNrefine): the learned update works on the
// MATERIALISED residual, so this one ablation writes it out: out[b][q][c] = relu(s_c(uv(q, pose)) - g_c) over the
// residual pixels (the leading ReLU of NNrefine.linear_k is applied here), s = the satellite features warped with the
// same sampler as the LM step, both sides L2-normalised (HaLevel.scale) and masked like models_kitti.py:927,1191.
// One warp per pixel, lanes across channel quads.
template <int GEOM>
__global__ void __launch_bounds__(256) lm_residual_kernel(const LmStepArgs a, float* __restrict__ out, int C, int relu) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int P = (a.H - a.row0) * a.W;
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= P) return;
  const float su = a.pose[b * 3 + 0], sv = a.pose[b * 3 + 1], th = a.pose[b * 3 + 2];
  KittiPose kp;
  FordPose fp;
  G2spPose gq;
  if (GEOM == HA_GEOM_KITTI) kp = kitti_pose(a, su, sv, th);
  else fp = ford_pose(a, b, su, sv, th);
  const int C4 = C / 4;
  const float4* tab = a.table + (size_t)a.row0 * a.W;
  const PixelScalars ps = pixel_scalars<GEOM>(a, kp, fp, gq, tab[q], nullptr, q, P, C4);
  const float4* sat = reinterpret_cast<const float4*>(a.sat) + (size_t)b * a.A * a.A * C4;
  const float4* grd = reinterpret_cast<const float4*>(a.grd) + ((size_t)b * a.H * a.W + (size_t)a.row0 * a.W) * C4;
  const float alpha = a.sat_scale ? a.sat_scale[b] : 1.f, beta = a.grd_scale ? a.grd_scale[b] : 1.f;
  float4* o = reinterpret_cast<float4*>(out) + ((size_t)b * P + q) * C4;
  for (int c = lane; c < C4; c += 32) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ps.valid != 0.f) {
      const float4 nw = sat[ps.off_n + c], ne = sat[ps.off_n + ps.east + c], sw = sat[ps.off_s + c], se = sat[ps.off_s + ps.east + c];
      // jacobian.py:174-186: out = sum of the four taps times their (clamped-corner) weights
      const float wnw = ps.ex * ps.sy, wne = ps.wx * ps.sy, wsw = ps.ex * ps.ny, wse = ps.wx * ps.ny;
      s.x = nw.x * wnw + ne.x * wne + sw.x * wsw + se.x * wse; s.y = nw.y * wnw + ne.y * wne + sw.y * wsw + se.y * wse;
      s.z = nw.z * wnw + ne.z * wne + sw.z * wsw + se.z * wse; s.w = nw.w * wnw + ne.w * wne + sw.w * wsw + se.w * wse;
    }
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ps.goff >= 0) g = grd[ps.goff + c];
    float4 r = make_float4(alpha * s.x - beta * g.x, alpha * s.y - beta * g.y, alpha * s.z - beta * g.z, alpha * s.w - beta * g.w);
    if (relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
    o[c] = r;
  }
}

// NNrefine tail (RNNs.py:118-126) + NN_update (models_kitti.py:1045-1053): x = mean over the pixels of the 64-channel conv
// output, y = tanh(W1 relu(W0 relu(x) + b0) + b1), pose += y (all three components, no reset).  One CTA per sample.
__global__ void __launch_bounds__(256) nn_pose_update_kernel(const float* __restrict__ x, int P, const float* __restrict__ w0,
                                                             const float* __restrict__ b0, const float* __restrict__ w1,
                                                             const float* __restrict__ b1, float* __restrict__ pose,
                                                             float* __restrict__ traj, int traj_stride, uint32_t* status) {
  __shared__ double part[4][64];
  __shared__ float mean[64], hid[16];
  const int b = blockIdx.x, c = threadIdx.x & 63, slice = threadIdx.x >> 6;
  const float* xb = x + (size_t)b * P * 64;
  double acc = 0.0;
  for (int p = slice; p < P; p += 4) acc += (double)xb[(size_t)p * 64 + c];
  part[slice][c] = acc;
  __syncthreads();
  if (threadIdx.x < 64) mean[c] = fmaxf((float)((part[0][c] + part[1][c] + part[2][c] + part[3][c]) / (double)P), 0.f);   // mapping[0]: ReLU
  __syncthreads();
  if (threadIdx.x < 16) {
    float h = b0[threadIdx.x];
    for (int k = 0; k < 64; ++k) h += w0[threadIdx.x * 64 + k] * mean[k];
    hid[threadIdx.x] = fmaxf(h, 0.f);
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float y = b1[threadIdx.x];
    for (int k = 0; k < 16; ++k) y += w1[threadIdx.x * 16 + k] * hid[k];
    const float np = pose[b * 3 + threadIdx.x] + tanhf(y);
    pose[b * 3 + threadIdx.x] = np;
    if (traj) traj[(size_t)b * traj_stride + threadIdx.x] = np;
    if (np != np) atomicOr(status, HA_STATUS_NAN_POSE);
  }
}

}  // namespace ha

// ------------------------------------------------------------------------------------ C ABI
extern "C" int ha_lm_residual(const HaLmParams* p, int level, const HaLevel* sat, const HaLevel* grd, const float* ground_table,
                              const float* extrinsics, const float* pose, int relu, float* out, void* stream) {
  using namespace ha;
  if (!p || !sat || !grd || !ground_table || !pose || !out || level < 0 || level >= HA_MAX_LEVELS) return HA_EINVAL;
  if (p->geometry != HA_GEOM_KITTI && p->geometry != HA_GEOM_FORD) return HA_EINVAL;
  if (sat->C != grd->C || (grd->C % 4) || sat->H != sat->W || (grd->H & 1) || p->batch <= 0) return HA_EINVAL;
  if (p->geometry == HA_GEOM_FORD && !extrinsics) return HA_EINVAL;
  LmStepArgs a{};
  a.sat = sat->data; a.grd = grd->data; a.sat_scale = sat->scale; a.grd_scale = grd->scale;
  a.table = reinterpret_cast<const float4*>(ground_table); a.extr = extrinsics; a.pose = const_cast<float*>(pose);
  a.B = p->batch; a.A = sat->H; a.H = grd->H; a.W = grd->W; a.grd_C = grd->C;
  a.rot = p->rotation_range; a.lat = p->shift_range_lat; a.lon = p->shift_range_lon;
  a.mpp = p->meter_per_pixel[level]; a.inv_mpp = p->inv_meter_per_pixel[level]; a.center = p->sat_center[level];
  a.row0 = p->full_height ? 0 : grd->H / 2;
  const int P = (grd->H - a.row0) * grd->W;
  const dim3 grid((P + 7) / 8, p->batch);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (p->geometry == HA_GEOM_KITTI) lm_residual_kernel<HA_GEOM_KITTI><<<grid, 256, 0, st>>>(a, out, grd->C, relu);
  else lm_residual_kernel<HA_GEOM_FORD><<<grid, 256, 0, st>>>(a, out, grd->C, relu);
  count_launches(1);
  return check_launch("lm_residual_kernel");
}

extern "C" int ha_nn_pose_update(const float* x, int B, int n_px, const float* w0, const float* b0, const float* w1, const float* b1,
                                 float* pose, float* traj_step, int traj_stride, uint32_t* status, void* stream) {
  if (!x || !w0 || !b0 || !w1 || !b1 || !pose || !status || B <= 0 || n_px <= 0) return HA_EINVAL;
  ha::nn_pose_update_kernel<<<B, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, n_px, w0, b0, w1, b1, pose, traj_step, traj_stride,
                                                                                   status);
  ha::count_launches(1);
  return ha::check_launch("nn_pose_update_kernel");
}

extern "C" size_t ha_lm_workspace_bytes(int B) { return B > 0 ? ha::lm_ws_bytes(B) : 0; }

extern "C" int ha_lm_step(const HaLmParams* p, int level, const HaLevel* sat, const HaLevel* grd, const float* grd_conf,
                          const float* ground_table, const float* extrinsics, float* pose, const float* reset_uv,
                          float* stats, uint32_t* status, void* ws, size_t ws_bytes, void* stream) {
  if (!p) return HA_EINVAL;
  const int B = p->batch;
  if (B <= 0) return HA_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!ws || !status) return HA_EINVAL;
  if (ws_bytes < ha::lm_ws_bytes(B)) return HA_ENOSPACE;
  ha::lm_begin(ws, B, status, st);        // tickets / step word / zero vector / *status start at zero (cheap, async)
  return ha::lm_step_impl(p, level, sat, grd, grd_conf, ground_table, extrinsics, pose, reset_uv, stats, nullptr, 0,
                          status, ws, ws_bytes, B, /*full=*/true, p->adam_iter, /*chain_step=*/-1, st);
}

extern "C" int ha_lm_run(const HaLmParams* p, const HaLevel* sat, const HaLevel* grd, const float* const* grd_conf,
                         const float* const* ground_tables, const float* extrinsics, float* pose,
                         const float* reset_uv, float* traj, float* stats, uint32_t* status, void* ws, size_t ws_bytes,
                         void* stream) {
  if (!p || !sat || !grd || !ground_tables || !pose || !ws || !status) return HA_EINVAL;
  const int B = p->batch, L = p->n_levels, N = p->n_iters;
  if (B <= 0 || L < 1 || L > HA_MAX_LEVELS || N < 1) return HA_EINVAL;
  if (ws_bytes < ha::lm_ws_bytes(B)) return HA_ENOSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ha::lm_begin(ws, B, status, st);
  // the default step kernel of the S2GP geometries chains its launches (lm_chain_enter): one arrival word per step
  const bool chain = (p->kernel_variant == 0 || p->kernel_variant >= 3) && (p->geometry == HA_GEOM_KITTI || p->geometry == HA_GEOM_FORD) &&
                     (long long)N * L <= ha::kLmChainMaxSteps;
  int k = 0;
  const int outer = p->level_first ? L : N, inner = p->level_first ? N : L;
  for (int o = 0; o < outer; ++o) {
    for (int i = 0; i < inner; ++i, ++k) {
      const int it = p->level_first ? i : o, lv = p->level_first ? o : i;
      const float* ruv = (p->dof == 3 && reset_uv) ? reset_uv + (size_t)k * 2 * B : nullptr;   // SGD / ADAM draw nothing: reset_uv NULL
      float* tr = traj ? traj + ((size_t)it * L + lv) * 3 : nullptr;
      float* stp = stats ? stats + ((size_t)it * L + lv) * B * HA_STATS : nullptr;
      // the first visit of a level also reduces |g|^2 (pose independent) and caches it; later visits skip it
      const bool full = (it == 0) || stats != nullptr;
      int rc = ha::lm_step_impl(p, lv, sat + lv, grd + lv, grd_conf ? grd_conf[lv] : nullptr, ground_tables[lv],
                                extrinsics, pose, ruv, stp, tr, N * L * 3, status, ws, ws_bytes, B, full, it, chain ? k : -1, st);
      if (rc != HA_OK) return rc;
    }
  }
  return HA_OK;
}
