// csrc/ba_kernels.cu — sliding-window bundle adjustment on sm_100a: the whole g2o-style
// Levenberg-Marquardt loop of one LocalmapOptimization call runs inside ONE persistent kernel.
//
// Reference behaviour reproduced:
//   /root/reference/src/g2o_optimization.cc:20-177 (LocalmapOptimization call sequence)
//   g2o EdgeSE3ProjectXYZ / RobustKernelHuber / BlockSolver Schur / OptimizationAlgorithmLevenberg
//   (upstream g2o, not vendored by the reference; semantics in SURVEY.md §8c.1).
//
// Formulation: matrix-free in the observations.  Nothing per-observation is ever
// written to HBM: each phase re-derives residual / Huber weight / 2x6 and 2x3 Jacobians from the
// 24 B/observation SoA input (uv + camera index), because on B200 the fp64 flops to recompute are
// cheaper than the 144 B/observation an explicit Hpl block would cost to write and re-read twice.
//   phase LIN      warp per point: residuals, Hll/bl (warp-shuffle reduction), Dinv, and the Schur
//                  complement contributions J_i^T (w_i I - A_i B_j^T) J_j accumulated into the
//                  block-sparse reduced camera system S, right-hand side b_s
//   phase PCG      block-Jacobi preconditioned conjugate gradients on S x_p = b_s
//   phase BACKSUB  warp per point: x_l = Dinv (b_l - W^T x_p), trial state, trial robust chi2
//   decide         gain ratio, lambda update, accept / reject — on device, uniform over the scope

#include "ba_device.cuh"
#include "ba_types.h"
#include "common.cuh"
#include "kernels.h"
#include "../../include/urmvo_b200.h"

namespace urmvo {

// Phase timing (SM cycles) of window 0 as seen by thread 0 of its first CTA: a development aid read
// back by urmvo_debug_ba_timing().  0 LIN(diag) 1 LIN 2 reduce 3 PCG 4 cam update 5 BACKSUB
// 6 reduce/decide 7 classify
__device__ unsigned long long g_ba_timing[8];
#define BA_T0() const long long _t0 = clock64()
#define BA_T1(slot) do { if (timer) g_ba_timing[slot] += (unsigned long long)(clock64() - _t0); } while (0)


// Block (ci,cj) of S, ci <= cj, by binary search in the BSR row; -1 if absent.
__device__ __forceinline__ int find_block(const BAWin& W, int ci, int cj) {
  int lo = W.row_ptr[ci], hi = W.row_ptr[ci + 1] - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    const int c = W.col[mid];
    if (c == cj) return mid;
    if (c < cj) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

// Per-warp staging area, structure-of-arrays over the observations of one point.
struct WarpStage {
  double* f;  // 27 fields x kmax: Jp[12] | B[6] | A[6] | we[2] | w
  int* cf;    // kmax
  int kmax;
  __device__ __forceinline__ double& Jp(int a, int i) { return f[a * kmax + i]; }
  __device__ __forceinline__ double& B(int a, int i) { return f[(12 + a) * kmax + i]; }
  __device__ __forceinline__ double& A(int a, int i) { return f[(18 + a) * kmax + i]; }
  __device__ __forceinline__ double& we(int a, int i) { return f[(24 + a) * kmax + i]; }
  __device__ __forceinline__ double& w(int i) { return f[26 * kmax + i]; }
};
constexpr int kStageFields = 27;

__host__ __device__ inline size_t warp_stage_bytes(int kmax) {
  return (size_t)kmax * (kStageFields * sizeof(double) + sizeof(int));
}

// Per-warp work areas in dynamic shared memory: warp w owns [base + w*stride, base + (w+1)*stride).
struct WorkArea {
  double* base;
  int stride;
};

// Fixed-order sum of the warps' accumulator copies (at `off` inside each work area) -> this CTA's
// partial in BAWin::Spart.
template <class Scope>
__device__ __forceinline__ void cta_reduce_copies(const Scope& sc, const BAWin& W, WorkArea wa, int off, int len) {
  const int wpc = blockDim.x >> 5;
  __syncthreads();
  double* out = W.Spart + (size_t)sc.blk() * W.acc_len;
  for (int e = threadIdx.x; e < len; e += blockDim.x) {
    double v = wa.base[off + e];
    for (int w = 1; w < wpc; w++) v += wa.base[(size_t)w * wa.stride + off + e];
    __stcg(out + e, v);
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------- phase LIN

// DIAG = true: only diag(Hpp) (atomics into hdiag), max diag(Hll) and the robust chi2 — the
// quantities OptimizationAlgorithmLevenberg::computeLambdaInit needs at iteration 0.
//
// SMEM = true (small windows, BAWin::acc_mode 1): every warp accumulates S / b_s / b_p into its own
// shared-memory copy with plain adds (one lane per camera pair => no two lanes touch the same
// block), the copies are summed in warp order per CTA and stored as one partial per CTA in
// BAWin::Spart; the consumer sums the CTA partials in CTA order.  No atomics, bit-reproducible.
// SMEM = false: fp64 atomics into the global block-sparse S (large / sparse reduced systems).
template <bool SMEM>
__device__ __forceinline__ void acc_add(double* p, double v) {
  if (SMEM) *p += v; else atomicAdd(p, v);
}

template <bool DIAG, bool SMEM, class Scope>
__device__ void lin_phase(const Scope& sc, const BAWin& W, int cur, double lambda, bool robust,
                          double delta, WarpStage st, WorkArea wa, double& chi_acc, double& maxdiag_acc) {
  const int lane = threadIdx.x & 31;
  const int wpc = blockDim.x >> 5;
  const int gw = sc.blk() * wpc + (threadIdx.x >> 5);
  const int gstride = sc.nblk() * wpc;
  const double* __restrict__ camRt = W.camRt[cur];
  const double* __restrict__ pts = W.pts[cur];
  const double K[4] = {W.intr[0], W.intr[1], W.intr[2], W.intr[3]};
  const int acc_len = DIAG ? W.Ncf * 6 : W.acc_len;
  const int acc_off = kStageFields * st.kmax;  // the accumulator copy follows the staging fields
  double* acc = wa.base + (size_t)(threadIdx.x >> 5) * wa.stride + acc_off;  // this warp's copy
  double* accS = SMEM ? acc : W.S;
  double* accbs = SMEM ? acc + (size_t)W.nblk * 36 : W.bs;
  double* accbp = SMEM ? acc + (size_t)W.nblk * 36 + W.Ncf * 6 : W.bp;
  double* acchd = SMEM ? acc : W.hdiag;
  if (SMEM) {
    for (int e = lane; e < acc_len; e += 32) acc[e] = 0.0;
    __syncwarp();
  }

  for (int l = gw; l < W.Np; l += gstride) {
    const int ps = W.pt_start[l], k = W.pt_start[l + 1] - ps;
    const double X[3] = {pts[l * 3], pts[l * 3 + 1], pts[l * 3 + 2]};
    double h[6] = {0, 0, 0, 0, 0, 0}, bl[3] = {0, 0, 0};
    for (int i = lane; i < k; i += 32) {
      const int o = ps + i;
      int cf = -1;
      if (!W.level[o]) {
        const int c = W.ocam[o];
        const double* Rt = camRt + (size_t)c * 12;
        double pc[3], pz[3], e0, e1, w;
        map_point(Rt, X, pc);
        const double2 uv = *reinterpret_cast<const double2*>(W.uv + (size_t)o * 2);
        const double e2 = edge_error(pc, uv.x, uv.y, K, e0, e1, pz);
        chi_acc += huber_rho(e2, delta, robust, w);
        double Jx[6], B[6];
        edge_jac_point(Rt, pz, K, Jx);
#pragma unroll
        for (int a = 0; a < 6; a++) B[a] = w * Jx[a];
        h[0] += B[0] * Jx[0] + B[3] * Jx[3];
        h[1] += B[0] * Jx[1] + B[3] * Jx[4];
        h[2] += B[0] * Jx[2] + B[3] * Jx[5];
        h[3] += B[1] * Jx[1] + B[4] * Jx[4];
        h[4] += B[1] * Jx[2] + B[4] * Jx[5];
        h[5] += B[2] * Jx[2] + B[5] * Jx[5];
#pragma unroll
        for (int a = 0; a < 3; a++) bl[a] -= B[a] * e0 + B[3 + a] * e1;
        cf = W.cam_free[c];
        if (cf >= 0) {
          double Jp[12];
          edge_jac_pose(pz, K, Jp);
          if (DIAG) {
#pragma unroll
            for (int a = 0; a < 6; a++)
              acc_add<SMEM>(&acchd[cf * 6 + a], w * (Jp[a] * Jp[a] + Jp[6 + a] * Jp[6 + a]));
          } else {
#pragma unroll
            for (int a = 0; a < 12; a++) st.Jp(a, i) = Jp[a];
#pragma unroll
            for (int a = 0; a < 6; a++) st.B(a, i) = B[a];
            st.we(0, i) = w * e0;
            st.we(1, i) = w * e1;
            st.w(i) = w;
          }
        }
      }
      if (!DIAG) st.cf[i] = cf;
    }
#pragma unroll
    for (int a = 0; a < 6; a++) h[a] = warp_sum(h[a]);
    if (DIAG) {
      maxdiag_acc = fmax(maxdiag_acc, fmax(fabs(h[0]), fmax(fabs(h[3]), fabs(h[5]))));
      if (SMEM) __syncwarp();  // the next point's lanes add into the same warp copy (racecheck)
      continue;
    }
#pragma unroll
    for (int a = 0; a < 3; a++) bl[a] = warp_sum(bl[a]);
    // (Hll + lambda I)^-1 — BlockSolver::setLambda adds lambda to every diagonal block
    double Di[6];
    {
      const double hl[6] = {h[0] + lambda, h[1], h[2], h[3] + lambda, h[4], h[5] + lambda};
      sym3_inverse(hl, Di);
    }
    if (lane == 0) {
#pragma unroll
      for (int a = 0; a < 6; a++) W.Dinv[(size_t)l * 6 + a] = Di[a];
#pragma unroll
      for (int a = 0; a < 3; a++) W.bl[(size_t)l * 3 + a] = bl[a];
    }
    __syncwarp();
    // A_i = B_i Dinv (2x3), gradient terms: b_s += -J^T (w e + A b_l), b_p += -J^T w e
    for (int i = lane; i < k; i += 32) {
      const int cf = st.cf[i];
      if (cf < 0) continue;
      double A[6];
#pragma unroll
      for (int r = 0; r < 2; r++) {
        const double b0 = st.B(r * 3, i), b1 = st.B(r * 3 + 1, i), b2 = st.B(r * 3 + 2, i);
        A[r * 3 + 0] = b0 * Di[0] + b1 * Di[1] + b2 * Di[2];
        A[r * 3 + 1] = b0 * Di[1] + b1 * Di[3] + b2 * Di[4];
        A[r * 3 + 2] = b0 * Di[2] + b1 * Di[4] + b2 * Di[5];
      }
#pragma unroll
      for (int a = 0; a < 6; a++) st.A(a, i) = A[a];
      const double we0 = st.we(0, i), we1 = st.we(1, i);
      const double g0 = we0 + (A[0] * bl[0] + A[1] * bl[1] + A[2] * bl[2]);
      const double g1 = we1 + (A[3] * bl[0] + A[4] * bl[1] + A[5] * bl[2]);
#pragma unroll
      for (int a = 0; a < 6; a++) {
        const double j0 = st.Jp(a, i), j1 = st.Jp(6 + a, i);
        acc_add<SMEM>(&accbs[cf * 6 + a], -(j0 * g0 + j1 * g1));
        acc_add<SMEM>(&accbp[cf * 6 + a], -(j0 * we0 + j1 * we1));
      }
    }
    __syncwarp();
    // Schur pairs (i <= j): one lane per pair, 6x6 block = J_i^T M J_j,
    //   M = w_i I - A_i B_i^T (i == j: Hpp term and its Schur correction), M = -A_i B_j^T otherwise.
    // SMEM: the lane adds its block into the warp's shared-memory copy.  Atomic mode: the 32 blocks
    // of a round are transposed through shared memory so that each RED instruction covers 32
    // CONSECUTIVE doubles of one block (2 instructions per block) — the L2 atomic units work per
    // 32-byte sector, so this is ~4x fewer atomic transactions than 36 scattered REDs per lane.
    const int npairs = k * (k + 1) / 2;
    double* tile = wa.base + (size_t)(threadIdx.x >> 5) * wa.stride + acc_off;  // 32 x 37 doubles
    for (int pbase = 0; pbase < npairs; pbase += 32) {
      const int pi = pbase + lane;
      int my_blk = -1;
      if (pi < npairs) {
        const int kk = 2 * k + 1;
        int i = (int)(((float)kk - sqrtf((float)(kk * kk - 8 * pi))) * 0.5f);
        i = max(0, min(i, k - 1));
        while (i > 0 && i * k - i * (i - 1) / 2 > pi) i--;
        while ((i + 1) * k - (i + 1) * i / 2 <= pi) i++;
        const int j = i + (pi - (i * k - i * (i - 1) / 2));
        const int ci = st.cf[i], cj = st.cf[j];
        if (ci >= 0 && cj >= 0) {
          double M[4];
          {
            const double a0 = st.A(0, i), a1 = st.A(1, i), a2 = st.A(2, i);
            const double a3 = st.A(3, i), a4 = st.A(4, i), a5 = st.A(5, i);
            const double b0 = st.B(0, j), b1 = st.B(1, j), b2 = st.B(2, j);
            const double b3 = st.B(3, j), b4 = st.B(4, j), b5 = st.B(5, j);
            M[0] = -(a0 * b0 + a1 * b1 + a2 * b2);
            M[1] = -(a0 * b3 + a1 * b4 + a2 * b5);
            M[2] = -(a3 * b0 + a4 * b1 + a5 * b2);
            M[3] = -(a3 * b3 + a4 * b4 + a5 * b5);
            if (i == j) { const double w = st.w(i); M[0] += w; M[3] += w; }
          }
          double T[12];  // T = M J_j (2x6)
#pragma unroll
          for (int b = 0; b < 6; b++) {
            const double j0 = st.Jp(b, j), j1 = st.Jp(6 + b, j);
            T[b] = M[0] * j0 + M[1] * j1;
            T[6 + b] = M[2] * j0 + M[3] * j1;
          }
          const bool swap = ci > cj;
          // acc_mode 1 implies the dense upper-triangular block layout: block (a,b) = row_ptr[a] + b - a
          const int blk = SMEM ? (swap ? W.row_ptr[cj] + ci - cj : W.row_ptr[ci] + cj - ci)
                               : (swap ? find_block(W, cj, ci) : find_block(W, ci, cj));
          if (blk >= 0) {  // always, for a structure built from the same observations
            const bool same_cam_twice = !SMEM && (ci == cj) && (i != j);  // excluded on the host in acc_mode 1
            double* Sb = accS + (size_t)blk * 36;
            double* tl = tile + lane * 37;
#pragma unroll
            for (int a = 0; a < 6; a++) {
              const double j0 = st.Jp(a, i), j1 = st.Jp(6 + a, i);
#pragma unroll
              for (int b = 0; b < 6; b++) {
                const double v = j0 * T[b] + j1 * T[6 + b];
                if (SMEM) {
                  if (!swap) Sb[a * 6 + b] += v; else Sb[b * 6 + a] += v;
                } else if (same_cam_twice) {
                  atomicAdd(&Sb[a * 6 + b], v);
                  atomicAdd(&Sb[b * 6 + a], v);
                } else {
                  tl[swap ? b * 6 + a : a * 6 + b] = v;
                }
              }
            }
            if (!SMEM && !same_cam_twice) my_blk = blk;
          }
        }
      }
      if (!SMEM) {
        __syncwarp();
        const int cnt = min(32, npairs - pbase);
        for (int q = 0; q < cnt; q++) {
          const int b = __shfl_sync(0xffffffffu, my_blk, q);
          if (b < 0) continue;
          double* Sb = W.S + (size_t)b * 36;
          atomicAdd(&Sb[lane], tile[q * 37 + lane]);
          if (lane < 4) atomicAdd(&Sb[32 + lane], tile[q * 37 + 32 + lane]);
        }
        __syncwarp();
      }
    }
    __syncwarp();
  }
  if (SMEM) {
    // fixed-order sum of the warp copies -> this CTA's partial
    __syncthreads();
    cta_reduce_copies(sc, W, wa, acc_off, acc_len);
  }
}


// ------------------------------------------------------------------------------- stereo edges
//
// Windows of a STEREO camera (reference src/g2o_optimization.cc:96-118) mix EdgeSE3ProjectXYZ (2 rows)
// and EdgeStereoSE3ProjectXYZ (3 rows: u, v, u_right; Omega = I3; Huber delta sqrt(cfg.stereo_point)).
// The phases below are the one-point-per-warp phases written for THREE residual rows; a mono edge
// leaves its third row zero.  They serve accumulation modes 5 (shared-memory copies) and 6 (global
// atomics); the packed / tile modes stay mono-only.  Staging: Jp[18] | B[9] | A[9] | we[3] | w.
constexpr int kStageFieldsS = 40;
struct WarpStageS {
  double* f;
  int* cf;
  int kmax;
  __device__ __forceinline__ double& Jp(int a, int i) { return f[a * kmax + i]; }
  __device__ __forceinline__ double& B(int a, int i) { return f[(18 + a) * kmax + i]; }
  __device__ __forceinline__ double& A(int a, int i) { return f[(27 + a) * kmax + i]; }
  __device__ __forceinline__ double& we(int a, int i) { return f[(36 + a) * kmax + i]; }
  __device__ __forceinline__ double& w(int i) { return f[39 * kmax + i]; }
};

struct StereoPar { double delta_s; };

// Camera model of an edge: the reference reads fx, fy, cx, cy and BF per constraint from
// camera_list[mpc->id_camera] (src/g2o_optimization.cc:86-89, :106-113).  okind[o] = stereo bit | model << 1,
// BAWin::intr_tab holds one (fx fy cx cy bf) row per model (one row for the usual single-camera call).
__device__ __forceinline__ const double* edge_model(const BAWin& W, int o, bool& stereo) {
  const unsigned kb = W.okind[o];
  stereo = (kb & 1u) != 0;
  return W.intr_tab + 5 * (kb >> 1);
}

// residual (3 rows), Huber weight, chi2 contribution; returns rho0.  pz = (x/z, y/z, 1/z); K = the edge's model row.
__device__ __forceinline__ double edge_eval3(const BAWin& W, int o, const double* pc, const double*& K,
                                             const StereoPar& sp, double delta, bool robust, double* e, double* pz,
                                             double& w, bool& stereo) {
  K = edge_model(W, o, stereo);
  const double2 uv = *reinterpret_cast<const double2*>(W.uv + (size_t)o * 2);
  double e2 = edge_error(pc, uv.x, uv.y, K, e[0], e[1], pz);
  e[2] = 0.0;
  if (stereo) { e[2] = edge_error_right(pz, W.ur[o], K, K[4]); e2 += e[2] * e[2]; }
  return huber_rho(e2, stereo ? sp.delta_s : delta, robust, w);
}

template <bool DIAG, bool SMEM, class Scope>
__device__ void lin_phase_s(const Scope& sc, const BAWin& W, int cur, double lambda, bool robust,
                            double delta, StereoPar sp, WarpStageS st, WorkArea wa, double& chi_acc,
                            double& maxdiag_acc) {
  const int lane = threadIdx.x & 31;
  const int wpc = blockDim.x >> 5;
  const int gw = sc.blk() * wpc + (threadIdx.x >> 5);
  const int gstride = sc.nblk() * wpc;
  const double* __restrict__ camRt = W.camRt[cur];
  const double* __restrict__ pts = W.pts[cur];
  const int acc_len = DIAG ? W.Ncf * 6 : W.acc_len;
  const int acc_off = kStageFieldsS * st.kmax;
  double* acc = wa.base + (size_t)(threadIdx.x >> 5) * wa.stride + acc_off;
  double* accS = SMEM ? acc : W.S;
  double* accbs = SMEM ? acc + (size_t)W.nblk * 36 : W.bs;
  double* accbp = SMEM ? acc + (size_t)W.nblk * 36 + W.Ncf * 6 : W.bp;
  double* acchd = SMEM ? acc : W.hdiag;
  if (SMEM) {
    for (int e = lane; e < acc_len; e += 32) acc[e] = 0.0;
    __syncwarp();
  }
  for (int l = gw; l < W.Np; l += gstride) {
    const int ps = W.pt_start[l], k = W.pt_start[l + 1] - ps;
    const double X[3] = {pts[l * 3], pts[l * 3 + 1], pts[l * 3 + 2]};
    double h[6] = {0, 0, 0, 0, 0, 0}, bl[3] = {0, 0, 0};
    for (int i = lane; i < k; i += 32) {
      const int o = ps + i;
      int cf = -1;
      if (!W.level[o]) {
        const int c = W.ocam[o];
        const double* Rt = camRt + (size_t)c * 12;
        double pc[3], pz[3], e[3], w;
        bool stereo;
        const double* K;
        map_point(Rt, X, pc);
        chi_acc += edge_eval3(W, o, pc, K, sp, delta, robust, e, pz, w, stereo);
        double Jx[9], B[9];
        edge_jac_point(Rt, pz, K, Jx);
        if (stereo) edge_jac_point_right(Rt, pz, Jx, K[4], Jx + 6);
        else { Jx[6] = 0.0; Jx[7] = 0.0; Jx[8] = 0.0; }
#pragma unroll
        for (int a = 0; a < 9; a++) B[a] = w * Jx[a];
        h[0] += B[0] * Jx[0] + B[3] * Jx[3] + B[6] * Jx[6];
        h[1] += B[0] * Jx[1] + B[3] * Jx[4] + B[6] * Jx[7];
        h[2] += B[0] * Jx[2] + B[3] * Jx[5] + B[6] * Jx[8];
        h[3] += B[1] * Jx[1] + B[4] * Jx[4] + B[7] * Jx[7];
        h[4] += B[1] * Jx[2] + B[4] * Jx[5] + B[7] * Jx[8];
        h[5] += B[2] * Jx[2] + B[5] * Jx[5] + B[8] * Jx[8];
#pragma unroll
        for (int a = 0; a < 3; a++) bl[a] -= B[a] * e[0] + B[3 + a] * e[1] + B[6 + a] * e[2];
        cf = W.cam_free[c];
        if (cf >= 0) {
          double Jp[18];
          edge_jac_pose(pz, K, Jp);
          if (stereo) edge_jac_pose_right(pz, Jp, K[4], Jp + 12);
          else {
#pragma unroll
            for (int a = 0; a < 6; a++) Jp[12 + a] = 0.0;
          }
          if (DIAG) {
#pragma unroll
            for (int a = 0; a < 6; a++)
              acc_add<SMEM>(&acchd[cf * 6 + a], w * (Jp[a] * Jp[a] + Jp[6 + a] * Jp[6 + a] + Jp[12 + a] * Jp[12 + a]));
          } else {
#pragma unroll
            for (int a = 0; a < 18; a++) st.Jp(a, i) = Jp[a];
#pragma unroll
            for (int a = 0; a < 9; a++) st.B(a, i) = B[a];
#pragma unroll
            for (int a = 0; a < 3; a++) st.we(a, i) = w * e[a];
            st.w(i) = w;
          }
        }
      }
      if (!DIAG) st.cf[i] = cf;
    }
#pragma unroll
    for (int a = 0; a < 6; a++) h[a] = warp_sum(h[a]);
    if (DIAG) {
      maxdiag_acc = fmax(maxdiag_acc, fmax(fabs(h[0]), fmax(fabs(h[3]), fabs(h[5]))));
      if (SMEM) __syncwarp();
      continue;
    }
#pragma unroll
    for (int a = 0; a < 3; a++) bl[a] = warp_sum(bl[a]);
    double Di[6];
    {
      const double hl[6] = {h[0] + lambda, h[1], h[2], h[3] + lambda, h[4], h[5] + lambda};
      sym3_inverse(hl, Di);
    }
    if (lane == 0) {
#pragma unroll
      for (int a = 0; a < 6; a++) W.Dinv[(size_t)l * 6 + a] = Di[a];
#pragma unroll
      for (int a = 0; a < 3; a++) W.bl[(size_t)l * 3 + a] = bl[a];
    }
    __syncwarp();
    // A_i = B_i Dinv (3x3), gradient terms: b_s += -J^T (w e + A b_l), b_p += -J^T w e
    for (int i = lane; i < k; i += 32) {
      const int cf = st.cf[i];
      if (cf < 0) continue;
      double A[9], g[3], we[3];
#pragma unroll
      for (int r = 0; r < 3; r++) {
        const double b0 = st.B(r * 3, i), b1 = st.B(r * 3 + 1, i), b2 = st.B(r * 3 + 2, i);
        A[r * 3 + 0] = b0 * Di[0] + b1 * Di[1] + b2 * Di[2];
        A[r * 3 + 1] = b0 * Di[1] + b1 * Di[3] + b2 * Di[4];
        A[r * 3 + 2] = b0 * Di[2] + b1 * Di[4] + b2 * Di[5];
        we[r] = st.we(r, i);
        g[r] = we[r] + (A[r * 3] * bl[0] + A[r * 3 + 1] * bl[1] + A[r * 3 + 2] * bl[2]);
      }
#pragma unroll
      for (int a = 0; a < 9; a++) st.A(a, i) = A[a];
#pragma unroll
      for (int a = 0; a < 6; a++) {
        const double j0 = st.Jp(a, i), j1 = st.Jp(6 + a, i), j2 = st.Jp(12 + a, i);
        acc_add<SMEM>(&accbs[cf * 6 + a], -(j0 * g[0] + j1 * g[1] + j2 * g[2]));
        acc_add<SMEM>(&accbp[cf * 6 + a], -(j0 * we[0] + j1 * we[1] + j2 * we[2]));
      }
    }
    __syncwarp();
    // Schur pairs (i <= j): one lane per pair, block = J_i^T M J_j, M (3x3) = delta_ij w I - A_i B_j^T
    const int npairs = k * (k + 1) / 2;
    for (int pbase = 0; pbase < npairs; pbase += 32) {
      const int pi = pbase + lane;
      if (pi < npairs) {
        const int kk = 2 * k + 1;
        int i = (int)(((float)kk - sqrtf((float)(kk * kk - 8 * pi))) * 0.5f);
        i = max(0, min(i, k - 1));
        while (i > 0 && i * k - i * (i - 1) / 2 > pi) i--;
        while ((i + 1) * k - (i + 1) * i / 2 <= pi) i++;
        const int j = i + (pi - (i * k - i * (i - 1) / 2));
        const int ci = st.cf[i], cj = st.cf[j];
        if (ci >= 0 && cj >= 0) {
          double M[9];
#pragma unroll
          for (int r = 0; r < 3; r++)
#pragma unroll
            for (int q = 0; q < 3; q++)
              M[r * 3 + q] = -(st.A(r * 3, i) * st.B(q * 3, j) + st.A(r * 3 + 1, i) * st.B(q * 3 + 1, j) +
                               st.A(r * 3 + 2, i) * st.B(q * 3 + 2, j));
          if (i == j) { const double w = st.w(i); M[0] += w; M[4] += w; M[8] += w; }
          double T[18];  // T = M J_j (3x6)
#pragma unroll
          for (int b = 0; b < 6; b++) {
            const double j0 = st.Jp(b, j), j1 = st.Jp(6 + b, j), j2 = st.Jp(12 + b, j);
#pragma unroll
            for (int r = 0; r < 3; r++) T[r * 6 + b] = M[r * 3] * j0 + M[r * 3 + 1] * j1 + M[r * 3 + 2] * j2;
          }
          const bool swap = ci > cj;
          const int blk = SMEM ? (swap ? W.row_ptr[cj] + ci - cj : W.row_ptr[ci] + cj - ci)
                               : (swap ? find_block(W, cj, ci) : find_block(W, ci, cj));
          if (blk >= 0) {
            const bool same_cam_twice = !SMEM && (ci == cj) && (i != j);
            double* Sb = accS + (size_t)blk * 36;
#pragma unroll
            for (int a = 0; a < 6; a++) {
              const double j0 = st.Jp(a, i), j1 = st.Jp(6 + a, i), j2 = st.Jp(12 + a, i);
#pragma unroll
              for (int b = 0; b < 6; b++) {
                const double v = j0 * T[b] + j1 * T[6 + b] + j2 * T[12 + b];
                if (SMEM) {
                  if (!swap) Sb[a * 6 + b] += v; else Sb[b * 6 + a] += v;
                } else if (same_cam_twice) {
                  atomicAdd(&Sb[a * 6 + b], v);
                  atomicAdd(&Sb[b * 6 + a], v);
                } else {
                  atomicAdd(&Sb[swap ? b * 6 + a : a * 6 + b], v);
                }
              }
            }
          }
        }
      }
      if (SMEM) __syncwarp();  // a later round of this point may touch the same block (ordering inside the warp copy)
    }
    __syncwarp();
  }
  if (SMEM) {
    __syncthreads();
    cta_reduce_copies(sc, W, wa, acc_off, acc_len);
  }
}

template <class Scope>
__device__ void backsub_phase_s(const Scope& sc, const BAWin& W, int cur, double lambda, bool robust,
                                double delta, StereoPar sp, double& chi_acc, double& scale_acc) {
  const int lane = threadIdx.x & 31;
  const int wpc = blockDim.x >> 5;
  const int gw = sc.blk() * wpc + (threadIdx.x >> 5);
  const int gstride = sc.nblk() * wpc;
  const int tr = cur ^ 1;
  const double* __restrict__ camRt = W.camRt[cur];
  const double* __restrict__ camRtT = W.camRt[tr];
  for (int l = gw; l < W.Np; l += gstride) {
    const int ps = W.pt_start[l], k = W.pt_start[l + 1] - ps;
    const double X[3] = {W.pts[cur][l * 3], W.pts[cur][l * 3 + 1], W.pts[cur][l * 3 + 2]};
    double c3[3] = {0, 0, 0};
    for (int i = lane; i < k; i += 32) {
      const int o = ps + i;
      if (W.level[o]) continue;
      const int c = W.ocam[o];
      const int cf = W.cam_free[c];
      if (cf < 0) continue;
      const double* Rt = camRt + (size_t)c * 12;
      double pc[3], pz[3], e[3], w, Jp[18], Jx[9];
      bool stereo;
      const double* K;
      map_point(Rt, X, pc);
      edge_eval3(W, o, pc, K, sp, delta, robust, e, pz, w, stereo);
      edge_jac_pose(pz, K, Jp);
      edge_jac_point(Rt, pz, K, Jx);
      double s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
      for (int a = 0; a < 6; a++) {
        const double xa = __ldcg(W.xp + cf * 6 + a);
        s0 += Jp[a] * xa;
        s1 += Jp[6 + a] * xa;
      }
      if (stereo) {
        edge_jac_pose_right(pz, Jp, K[4], Jp + 12);
        edge_jac_point_right(Rt, pz, Jx, K[4], Jx + 6);
#pragma unroll
        for (int a = 0; a < 6; a++) s2 += Jp[12 + a] * __ldcg(W.xp + cf * 6 + a);
      } else {
        Jx[6] = 0.0; Jx[7] = 0.0; Jx[8] = 0.0;
      }
      s0 *= w; s1 *= w; s2 *= w;
#pragma unroll
      for (int a = 0; a < 3; a++) c3[a] += Jx[a] * s0 + Jx[3 + a] * s1 + Jx[6 + a] * s2;
    }
#pragma unroll
    for (int a = 0; a < 3; a++) c3[a] = warp_sum(c3[a]);
    const double* Di = W.Dinv + (size_t)l * 6;
    const double b0 = W.bl[(size_t)l * 3], b1 = W.bl[(size_t)l * 3 + 1], b2 = W.bl[(size_t)l * 3 + 2];
    const double r0 = b0 - c3[0], r1 = b1 - c3[1], r2 = b2 - c3[2];
    const double x0 = Di[0] * r0 + Di[1] * r1 + Di[2] * r2;
    const double x1 = Di[1] * r0 + Di[3] * r1 + Di[4] * r2;
    const double x2 = Di[2] * r0 + Di[4] * r1 + Di[5] * r2;
    const double Xn[3] = {X[0] + x0, X[1] + x1, X[2] + x2};
    if (lane == 0) {
      W.pts[tr][l * 3] = Xn[0]; W.pts[tr][l * 3 + 1] = Xn[1]; W.pts[tr][l * 3 + 2] = Xn[2];
      scale_acc += x0 * (lambda * x0 + b0) + x1 * (lambda * x1 + b1) + x2 * (lambda * x2 + b2);
    }
    for (int i = lane; i < k; i += 32) {
      const int o = ps + i;
      if (W.level[o]) continue;
      const double* Rt = camRtT + (size_t)W.ocam[o] * 12;
      double pc[3], pz[3], e[3], w;
      bool stereo;
      const double* K;
      map_point(Rt, Xn, pc);
      chi_acc += edge_eval3(W, o, pc, K, sp, delta, robust, e, pz, w, stereo);
    }
  }
}

// ------------------------------------------------------------------------------- phase PCG

// 6x6 SPD inverse by Cholesky (block-Jacobi preconditioner). Returns false if not positive definite.
__device__ __forceinline__ bool spd6_inverse(const double* __restrict__ A, double* __restrict__ Ainv) {
  double L[36];
#pragma unroll
  for (int i = 0; i < 6; i++) {
#pragma unroll
    for (int j = 0; j <= i; j++) {
      double s = A[i * 6 + j];
#pragma unroll
      for (int k = 0; k < j; k++) s -= L[i * 6 + k] * L[j * 6 + k];
      if (j < i) L[i * 6 + j] = s / L[j * 6 + j];
      else {
        if (!(s > 0.0)) return false;
        L[i * 6 + i] = sqrt(s);
      }
    }
  }
  // columns of the inverse: solve L L^T x = e_c
#pragma unroll
  for (int c = 0; c < 6; c++) {
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
      double s = (i == c) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < i; k++) s -= L[i * 6 + k] * y[k];
      y[i] = s / L[i * 6 + i];
    }
#pragma unroll
    for (int i = 5; i >= 0; i--) {
      double s = y[i];
#pragma unroll
      for (int k = i + 1; k < 6; k++) s -= L[k * 6 + i] * y[k];
      y[i] = s / L[i * 6 + i];
    }
#pragma unroll
    for (int i = 0; i < 6; i++) Ainv[i * 6 + c] = y[i];
  }
  return true;
}

constexpr int kPcgRowsPerWarp = 8;

// Solves S xp = bs on the block-sparse S with block-Jacobi preconditioned conjugate gradients in
// the Chronopoulos-Gear form: ONE scope-wide reduction (gamma = r.z and delta = z.Sz together) and
// one barrier (z visible) per iteration instead of two reductions and two barriers.
// Mapping: one WARP per camera block row; lane = (g, a) with g = lane / 6 in 0..4 striding over the
// blocks of the row (5 blocks in flight per warp, the row's list is latency-bound otherwise) and
// a = lane % 6 the scalar row; the five partial sums are combined by shuffles in fixed order.
// x, r, p, s of the row stay in registers (replicated over g), only z lives in global memory.
// Returns false when S is not positive definite (=> g2o's "ok2 == false").
// wsm / wcap: this warp's shared-memory work area (doubles), idle during the solve: the blocks of the
// warp's rows are copied there once (a-major, lower blocks transposed) and reused by every
// iteration; what does not fit is read from L2.
template <class Scope>
__device__ bool pcg_phase(const Scope& sc, const BAWin& W, double tol, int max_iter, double* part,
                          int& parity, double* red, int& iters_out, double* wsm, int wcap) {
  const int Ncf = W.Ncf;
  const int lane = threadIdx.x & 31;
  const int wpc = blockDim.x >> 5;
  const int gwarp = sc.blk() * wpc + (threadIdx.x >> 5);
  const int nwarp = sc.nblk() * wpc;
  const int grp = lane / 6, a = lane - grp * 6;       // lanes 30, 31: grp 5, idle in the products
  iters_out = 0;
  if (Ncf == 0) return true;
  // preconditioner: Minv_i = S_ii^-1
  double bad = 0.0;
  {
    const int gt = sc.blk() * blockDim.x + threadIdx.x, gstride = sc.nblk() * blockDim.x;
    for (int i = gt; i < Ncf; i += gstride)
      if (!spd6_inverse(W.S + (size_t)W.row_ptr[i] * 36, W.Minv + (size_t)i * 36)) bad = 1.0;
  }
  sc.sync();
  constexpr int RMAX = kPcgRowsPerWarp;  // block rows per warp, state kept in registers
  const int n_slots = (Ncf + nwarp - 1) / nwarp;
  if (n_slots > RMAX) return false;  // excluded on the host: scopes are sized so that this cannot happen
  double x[RMAX], r[RMAX], pv[RMAX], sv[RMAX], z[RMAX];
  const double* __restrict__ S = W.S;
  const double* __restrict__ Minv = W.Minv;
  const int* __restrict__ row_ptr = W.row_ptr;
  const int* __restrict__ col = W.col;
  const int* __restrict__ lrow_ptr = W.lrow_ptr;
  const int* __restrict__ lcol = W.lcol;
  const int* __restrict__ lblk = W.lblk;
  double* __restrict__ zg = W.z;
  auto precond = [&](int i, double rr) {  // z_a = sum_b Minv[a][b] r_b, r_b from lanes 0..5
    double zz = 0.0;
#pragma unroll
    for (int b2 = 0; b2 < 6; b2++) {
      const double rb = __shfl_sync(0xffffffffu, rr, b2);
      if (i < Ncf && a < 6) zz += Minv[(size_t)i * 36 + a * 6 + b2] * rb;
    }
    return zz;
  };
  // cache the blocks of the owned rows in shared memory (as many as fit): per cached block 36 doubles
  // of S (a-major, lower blocks transposed), 6 doubles of staging for z_j and the column index
  int c_off[RMAX], c_cnt[RMAX];
  {
    int used = 0;
#pragma unroll
    for (int k = 0; k < RMAX; k++) {
      c_off[k] = used; c_cnt[k] = 0;
      const int i = gwarp + k * nwarp;
      if (k < n_slots && i < Ncf && wsm) {
        const int u0 = row_ptr[i], nu = row_ptr[i + 1] - u0;
        const int l0 = lrow_ptr[i], nl = lrow_ptr[i + 1] - l0;
        const int fit = min(nu + nl, (wcap - used) / 43);
        int* cc = reinterpret_cast<int*>(wsm + used + fit * 42);
        for (int e = 0; e < fit; e++) {
          for (int q = lane; q < 36; q += 32) {
            const int qa = q / 6, qb = q - qa * 6;
            const double v = (e < nu) ? __ldcg(S + (size_t)(u0 + e) * 36 + q)
                                      : __ldcg(S + (size_t)lblk[l0 + e - nu] * 36 + qb * 6 + qa);
            wsm[used + e * 36 + q] = v;
          }
          if (lane == 0) cc[e] = (e < nu) ? col[u0 + e] : lcol[l0 + e - nu];
        }
        c_cnt[k] = fit;
        used += fit * 43;
      }
    }
    __syncwarp();
  }
  auto matvec_row = [&](int i, int k) {  // (S z)_{i,a}, identical in every lane with the same a
    double acc = 0.0;
    const int nc = c_cnt[k];
    double* cs = wsm + c_off[k];
    double* zs = cs + nc * 36;
    const int* cc = reinterpret_cast<const int*>(cs + nc * 42);
    // stage z of the cached neighbours: independent loads, one L2 round trip for the whole row
    for (int q = lane; q < nc * 6; q += 32) {
      const int e = q / 6;
      zs[q] = __ldcg(zg + cc[e] * 6 + (q - e * 6));
    }
    __syncwarp();
    if (i < Ncf && grp < 5) {
      const int u0 = row_ptr[i], nu = row_ptr[i + 1] - u0;
      const int l0 = lrow_ptr[i], nl = lrow_ptr[i + 1] - l0;
#pragma unroll 2
      for (int e = grp; e < nc; e += 5) {  // cached blocks: S and z from shared memory
        const double* Sb = cs + e * 36 + a * 6;
        const double* zj = zs + e * 6;
#pragma unroll
        for (int b2 = 0; b2 < 6; b2++) acc += Sb[b2] * zj[b2];
      }
#pragma unroll 2
      for (int e = nc + grp; e < nu + nl; e += 5) {
        if (e < nu) {
          const double* Sb = S + (size_t)(u0 + e) * 36 + a * 6;
          const double* zj = zg + col[u0 + e] * 6;
#pragma unroll
          for (int b2 = 0; b2 < 6; b2++) acc += __ldcg(Sb + b2) * __ldcg(zj + b2);
        } else {
          const int le = l0 + (e - nu);
          const double* Sb = S + (size_t)lblk[le] * 36 + a;
          const double* zj = zg + lcol[le] * 6;
#pragma unroll
          for (int b2 = 0; b2 < 6; b2++) acc += __ldcg(Sb + b2 * 6) * __ldcg(zj + b2);
        }
      }
    }
    // combine the five block groups in fixed order; every lane ends with the total of its a
    const int al = (grp < 5) ? a : 0;
    double tot = __shfl_sync(0xffffffffu, acc, al);
#pragma unroll
    for (int g2 = 1; g2 < 5; g2++) tot += __shfl_sync(0xffffffffu, acc, g2 * 6 + al);
    return tot;
  };
  // r = b, z = Minv r
#pragma unroll
  for (int k = 0; k < RMAX; k++) {
    x[k] = 0.0; pv[k] = 0.0; sv[k] = 0.0; r[k] = 0.0; z[k] = 0.0;
    if (k < n_slots) {
      const int i = gwarp + k * nwarp;
      if (i < Ncf && grp < 5) r[k] = __ldcg(W.bs + i * 6 + a);
      z[k] = precond(i, r[k]);
      if (i < Ncf && grp == 0) zg[i * 6 + a] = z[k];
    }
  }
  sc.sync();
  double gamma = 0.0, alpha = 0.0, stop = 0.0;
  bool ok = true;
  int it = 0;
  for (;; it++) {
    // w = S z ; gamma_new = r.z ; delta = z.w  (one reduction)
    double w[RMAX];
    double acc[3] = {0.0, 0.0, bad};
#pragma unroll
    for (int k = 0; k < RMAX; k++) {
      w[k] = 0.0;
      if (k < n_slots) {
        const int i = gwarp + k * nwarp;
        w[k] = matvec_row(i, k);
        if (i < Ncf && grp == 0) {
          acc[0] += r[k] * z[k];
          acc[1] += z[k] * w[k];
        }
      }
    }
    double dummy[1];
    scope_reduce<3, 0>(sc, acc, dummy, part, parity, red);  // its barrier also orders the z reads/writes
    if (acc[2] > 0.0) { ok = false; break; }
    const double gamma_new = acc[0], delta = acc[1];
    if (it == 0) {
      if (!(gamma_new > 0.0)) { ok = (gamma_new == 0.0); break; }  // b == 0 -> x = 0
      stop = tol * tol * gamma_new;
    } else if (!(gamma_new > stop)) {
      break;
    }
    if (it >= max_iter) break;
    const double beta = (it == 0) ? 0.0 : gamma_new / gamma;
    const double denom = (it == 0) ? delta : delta - beta * gamma_new / alpha;
    if (!(denom > 0.0)) { ok = false; break; }
    alpha = gamma_new / denom;
    gamma = gamma_new;
#pragma unroll
    for (int k = 0; k < RMAX; k++) {
      if (k < n_slots) {
        const int i = gwarp + k * nwarp;
        pv[k] = z[k] + beta * pv[k];
        sv[k] = w[k] + beta * sv[k];
        x[k] += alpha * pv[k];
        r[k] -= alpha * sv[k];
        z[k] = precond(i, r[k]);
        if (i < Ncf && grp == 0) zg[i * 6 + a] = z[k];
      }
    }
    sc.sync();
  }
#pragma unroll
  for (int k = 0; k < RMAX; k++) {
    if (k < n_slots) {
      const int i = gwarp + k * nwarp;
      if (i < Ncf && grp == 0) W.xp[i * 6 + a] = ok ? x[k] : 0.0;
    }
  }
  iters_out = it;
  return ok;
}

// Small windows (acc_mode >= 1): CTA 0 of the scope gathers the per-CTA partials of S / b_s / b_p,
// builds the damped dense reduced camera system in shared memory (n = 6*Ncf <= 96) and runs
// block-Jacobi PCG on it with the whole CTA: thread (chunk c, row r) multiplies a slice of row r,
// the slices are combined in chunk order, dot products are reduced in fixed order.
// sm: >= n*n + (5 + chunks)*n + 36*Ncf doubles (host: pcg_dense_doubles()).
// Writes x_p and the summed raw gradient b_p to global memory for the other CTAs.
// Tiled Cholesky of the n x n reduced camera system held in shared memory (row-major, leading dimension n, n a
// multiple of 6, lower triangle read) with the right-hand side carried along, then x = L^-T y by one warp.
// Thread (ti, tc), ti >= tc, keeps the 3 x 3 tile of rows 3 ti.., columns 3 tc.. in registers for the whole
// factorisation; n / 3 more threads carry the right-hand side as one more (1 x 3)-tiled row.  Step p: the owners of
// column p solve against the diagonal tile and publish their tiles, everybody to the right subtracts L_ip L_cp^T and
// the owner of the next diagonal tile factorises it (three reciprocal square roots of the leading minors, which do
// not depend on each other) while the others finish.  Two barriers per 3 pivots instead of one per pivot and no
// shared-memory read-modify-write of the trailing matrix: ~16 k cycles for n = 42 against ~45 k for the
// one-barrier-per-pivot LDL^T below and for the PCG.  Needs T (T + 1) / 2 + T <= blockDim.x threads (T = n / 3).
// scratch: panel (n + 3) * 3 doubles, lpp 8, invd n.  Returns false on a non-positive pivot (g2o: failed Cholesky).
__device__ bool chol_tiled_smem(double* S, int n, double* rhs, double* panel, double* lpp, double* invd,
                                double* x_out, int* s_fail) {
  const int T = n / 3, n_tiles = T * (T + 1) / 2, tid = threadIdx.x;
  const bool is_mat = tid < n_tiles, is_rhs = tid >= n_tiles && tid < n_tiles + T;
  int ti = 0, tc = 0;
  if (is_mat) {
    ti = (int)((sqrtf(8.0f * (float)tid + 1.0f) - 1.0f) * 0.5f);
    while ((ti + 1) * (ti + 2) / 2 <= tid) ti++;
    while (ti * (ti + 1) / 2 > tid) ti--;
    tc = tid - ti * (ti + 1) / 2;
  } else if (is_rhs) {
    ti = T;
    tc = tid - n_tiles;
  }
  double a[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  if (is_mat) {
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++) a[r][c] = S[(ti * 3 + r) * n + tc * 3 + c];
  } else if (is_rhs) {
#pragma unroll
    for (int c = 0; c < 3; c++) a[0][c] = rhs[tc * 3 + c];
  }
  if (tid == 0) *s_fail = 0;
  __syncthreads();
  auto factor_diag = [&](int p) {
    const double d0 = a[0][0], a10 = a[1][0], a11 = a[1][1], a20 = a[2][0], a21 = a[2][1], a22 = a[2][2];
    const double m1 = fma(a11, d0, -a10 * a10);
    const double c0 = fma(a11, a22, -a21 * a21), c1 = fma(a10, a22, -a21 * a20), c2 = fma(a10, a21, -a11 * a20);
    const double m2 = fma(d0, c0, fma(-a10, c1, a20 * c2));
    const bool bad = !(d0 > 0.0) || !(m1 > 0.0) || !(m2 > 0.0);
    const double q0 = rsqrt(bad ? 1.0 : d0), q1 = rsqrt(bad ? 1.0 : m1), q2 = rsqrt(bad ? 1.0 : m2);
    const double s0 = d0 * q0, s1 = m1 * q1;
    const double r0 = q0, r1 = q1 * s0, r2 = q2 * s1;
    const double l10 = a10 * r0, l20 = a20 * r0;
    const double l21 = (a21 - l20 * l10) * r1;
    if (bad) *s_fail = 1;
    lpp[0] = r0; lpp[1] = l10; lpp[2] = r1; lpp[3] = l20; lpp[4] = l21; lpp[5] = r2;
    invd[p * 3] = r0; invd[p * 3 + 1] = r1; invd[p * 3 + 2] = r2;
    a[1][0] = l10; a[2][0] = l20; a[2][1] = l21;
  };
  if (is_mat && ti == 0 && tc == 0) factor_diag(0);
  __syncthreads();
  for (int p = 0; p < T; p++) {
    if ((is_mat || is_rhs) && tc == p && ti > p) {
      const double r0 = lpp[0], l10 = lpp[1], r1 = lpp[2], l20 = lpp[3], l21 = lpp[4], r2 = lpp[5];
      double* dst = panel + ti * 9;
#pragma unroll
      for (int r = 0; r < 3; r++) {
        if (r > 0 && is_rhs) break;
        const double x0 = a[r][0] * r0;
        const double x1 = (a[r][1] - x0 * l10) * r1;
        const double x2 = (a[r][2] - x0 * l20 - x1 * l21) * r2;
        a[r][0] = x0; a[r][1] = x1; a[r][2] = x2;
        dst[r * 3] = x0; dst[r * 3 + 1] = x1; dst[r * 3 + 2] = x2;
      }
    }
    __syncthreads();
    if ((is_mat || is_rhs) && tc > p) {
      const double* Li = panel + ti * 9;
      const double* Lc = panel + tc * 9;
      double lc[9];
#pragma unroll
      for (int e = 0; e < 9; e++) lc[e] = Lc[e];
#pragma unroll
      for (int r = 0; r < 3; r++) {
        if (r > 0 && is_rhs) break;
        const double i0 = Li[r * 3], i1 = Li[r * 3 + 1], i2 = Li[r * 3 + 2];
#pragma unroll
        for (int c = 0; c < 3; c++) a[r][c] -= i0 * lc[c * 3] + i1 * lc[c * 3 + 1] + i2 * lc[c * 3 + 2];
      }
      if (is_mat && ti == p + 1 && tc == p + 1) factor_diag(p + 1);
    }
    __syncthreads();
  }
  if (*s_fail) return false;  // uniform: written before the last barrier
  // strict lower triangle of L back into S (row i of L in front of the diagonal), y into rhs
  if (is_mat) {
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++)
        if (ti != tc || c < r) S[(ti * 3 + r) * n + tc * 3 + c] = a[r][c];
  } else if (is_rhs) {
#pragma unroll
    for (int c = 0; c < 3; c++) rhs[tc * 3 + c] = a[0][c];
  }
  __syncthreads();
  // L^T x = y by warp 0: lane l holds y[l], y[l + 32], y[l + 64]; right-looking
  if (tid < 32) {
    const int lane = tid;
    double v[3];
#pragma unroll
    for (int sl = 0; sl < 3; sl++) v[sl] = sl * 32 + lane < n ? rhs[sl * 32 + lane] : 0.0;
#pragma unroll
    for (int jb = 2; jb >= 0; jb--) {
      const int jn = n - jb * 32 < 32 ? n - jb * 32 : 32;
      for (int jj = jn - 1; jj >= 0; jj--) {
        const int j = jb * 32 + jj;
        const double xj = __shfl_sync(0xffffffffu, v[jb], jj) * invd[j];
        if (lane == jj) v[jb] = xj;
        const double* Lj = S + (size_t)j * n;
#pragma unroll
        for (int sl = 0; sl <= jb; sl++) {
          const int i = sl * 32 + lane;
          if (i < j) v[sl] = fma(-Lj[i], xj, v[sl]);
        }
      }
    }
#pragma unroll
    for (int sl = 0; sl < 3; sl++)
      if (sl * 32 + lane < n) x_out[sl * 32 + lane] = v[sl];
  }
  __syncthreads();
  return true;
}

template <class Scope>
__device__ bool pcg_dense_smem(const Scope& sc, const BAWin& W, double lambda, double tol, int max_iter,
                               int solver, double* sm, int& iters_out) {
  const int n = W.Ncf * 6, nb = sc.nblk();
  const int tid = threadIdx.x;
  const int C = max(1, (int)blockDim.x / max(n, 1));  // column chunks
  const int Wc = (n + C - 1) / C;                      // columns per chunk
  double* Sd = sm;               // n x n, symmetric
  double* bsv = Sd + n * n;      // n
  double* pv = bsv + n;          // n
  double* rv = pv + n;           // n
  double* xv = rv + n;           // n
  double* apv = xv + n;          // n
  double* Mi = apv + n;          // Ncf x 36 (first the Cholesky factors, then the inverses)
  double* prt = Mi + W.Ncf * 36; // C x n partial products
  __shared__ int s_ok;
  __shared__ double s_red[8];
  iters_out = 0;
  if (n == 0) return true;
  // the descriptor lives in global memory: take the fields used per element into registers once
  const double* __restrict__ Spart = W.Spart;
  double* __restrict__ bp_out = W.bp;
  const int acc_len = W.acc_len, Ncf = W.Ncf;
  const int n_s = W.nblk * 36;
  for (int e = tid; e < n_s; e += blockDim.x) {
    double v = 0.0;
    for (int b = 0; b < nb; b++) v += __ldcg(Spart + (size_t)b * acc_len + e);
    const int blk = e / 36, ab = e - blk * 36, a = ab / 6, c = ab - a * 6;
    // dense upper layout (<= 16 free cameras): row ci starts at ci*Ncf - ci(ci-1)/2
    int ci = 0, start = 0;
    while (start + (Ncf - ci) <= blk) { start += Ncf - ci; ci++; }
    const int cj = ci + (blk - start);
    const int r = ci * 6 + a, q = cj * 6 + c;
    if (ci == cj) {
      if (a <= c) {  // the diagonal block is taken from its upper triangle
        const double d = (a == c) ? v + lambda : v;
        Sd[r * n + q] = d;
        Sd[q * n + r] = d;
      }
    } else {
      Sd[r * n + q] = v;
      Sd[q * n + r] = v;
    }
  }
  for (int e = tid; e < 2 * n; e += blockDim.x) {
    double v = 0.0;
    for (int b = 0; b < nb; b++) v += __ldcg(Spart + (size_t)b * acc_len + n_s + e);
    if (e < n) bsv[e] = v; else __stcg(bp_out + (e - n), v);
  }
  if (tid == 0) s_ok = 1;
  __syncthreads();
  const int Tt = n / 3;
  if (solver == 0 && n <= 96 && Tt * (Tt + 1) / 2 + Tt <= (int)blockDim.x) {
    // default: tiled Cholesky (the exact solve g2o's LinearSolverEigen performs)
    const bool ok = chol_tiled_smem(Sd, n, bsv, Mi, apv, pv, xv, &s_ok);
    for (int i = tid; i < n; i += blockDim.x) __stcg(W.xp + i, ok ? xv[i] : 0.0);
    __syncthreads();
    return ok;
  }
  if (solver != 2) {
    // Direct solve (urmvo_ba_options.dense_solver = 1): symmetric Gaussian elimination S = L D L^T on the lower triangle with
    // the right-hand side carried along as an extra column, then back substitution — the exact
    // solve g2o's LinearSolverEigen performs, one CTA barrier per pivot.  A non-positive pivot
    // fails the solve like a failed Cholesky does in g2o.
    double* invd = pv;  // 1 / d_k
    const int tx = tid & 15, ty = tid >> 4, ny = blockDim.x >> 4;
    bool ok = true;
    for (int k = 0; k < n; k++) {
      const double d = Sd[k * n + k];
      if (!(d > 0.0)) { ok = false; break; }  // uniform: every thread reads the same pivot
      const double inv = 1.0 / d;
      if (tid == 0) invd[k] = inv;
      const double bk = bsv[k];
      for (int i = k + 1 + ty; i < n; i += ny) {
        const double lik = Sd[i * n + k] * inv;
        for (int j = k + 1 + tx; j <= i; j += 16) Sd[i * n + j] -= lik * Sd[j * n + k];
        if (tx == 0) bsv[i] -= lik * bk;
      }
      __syncthreads();
    }
    if (ok) {
      for (int k = n - 1; k >= 0; k--) {
        const double xk = bsv[k] * invd[k];  // final: every update of row k has been applied
        for (int i = tid; i < k; i += blockDim.x) bsv[i] -= Sd[k * n + i] * xk;
        if (tid == 0) xv[k] = xk;
        __syncthreads();
      }
    }
    for (int i = tid; i < n; i += blockDim.x) __stcg(W.xp + i, ok ? xv[i] : 0.0);
    __syncthreads();
    return ok;
  }
  // block-Jacobi preconditioner: 6x6 Cholesky by one thread per camera, then the six columns of
  // the inverse by six threads per camera
  double* Lf = prt;  // Ncf x 36 scratch for the factors (prt is free until the first product)
  for (int i = tid; i < W.Ncf; i += blockDim.x) {
    double L[36];
    bool ok = true;
#pragma unroll
    for (int r = 0; r < 6; r++) {
#pragma unroll
      for (int c = 0; c <= r; c++) {
        double v = Sd[(i * 6 + r) * n + i * 6 + c];
#pragma unroll
        for (int k = 0; k < c; k++) v -= L[r * 6 + k] * L[c * 6 + k];
        if (c < r) L[r * 6 + c] = v * L[c * 6 + c];  // diagonal holds 1 / L_cc
        else {
          if (!(v > 0.0)) { ok = false; v = 1.0; }
          L[r * 6 + r] = 1.0 / sqrt(v);
        }
      }
    }
    if (!ok) s_ok = 0;
#pragma unroll
    for (int e = 0; e < 36; e++) Lf[i * 36 + e] = L[e];
  }
  __syncthreads();
  for (int e = tid; e < n; e += blockDim.x) {  // column `col` of inverse(S_ii)
    const int i = e / 6, col = e - i * 6;
    const double* L = Lf + i * 36;
    double y[6];
#pragma unroll
    for (int r = 0; r < 6; r++) {
      double v = (r == col) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < r; k++) v -= L[r * 6 + k] * y[k];
      y[r] = v * L[r * 6 + r];
    }
#pragma unroll
    for (int r = 5; r >= 0; r--) {
      double v = y[r];
#pragma unroll
      for (int k = r + 1; k < 6; k++) v -= L[k * 6 + r] * y[k];
      y[r] = v * L[r * 6 + r];
    }
#pragma unroll
    for (int r = 0; r < 6; r++) Mi[i * 36 + r * 6 + col] = y[r];
  }
  __syncthreads();
  if (!s_ok) return false;
  // fixed-order sum over the first n threads (<= 3 warps); result in every thread
  auto cta_dot = [&](double v) {
    v = warp_sum(v);
    __syncthreads();
    if ((tid & 31) == 0 && tid < 96) s_red[tid >> 5] = v;
    __syncthreads();
    double t = s_red[0];
    if (n > 32) t += s_red[1];
    if (n > 64) t += s_red[2];
    return t;
  };
  auto precond = [&](int row, const double* vec) {
    const int i = row / 6, a = row - i * 6;
    double zz = 0.0;
#pragma unroll
    for (int c = 0; c < 6; c++) zz += Mi[i * 36 + a * 6 + c] * vec[i * 6 + c];
    return zz;
  };
  double x = 0.0, r = 0.0, z = 0.0, p = 0.0;
  if (tid < n) {
    r = bsv[tid];
    z = precond(tid, bsv);
    p = z;
    pv[tid] = p;
  }
  double rz = cta_dot(tid < n ? r * z : 0.0);  // its barriers also publish pv
  bool ok = true;
  int it = 0;
  if (!(rz > 0.0)) {
    ok = (rz == 0.0);
  } else {
    const double stop = tol * tol * rz;
    const int row = tid % n, ch = tid / n;
    for (; it < max_iter; it++) {
      if (ch < C) {  // slice [ch*Wc, ...) of row `row`; S is symmetric, so read it column-major
        double s0 = 0.0;
        const int c1 = min(n, (ch + 1) * Wc);
        for (int c = ch * Wc; c < c1; c++) s0 += Sd[c * n + row] * pv[c];
        prt[ch * n + row] = s0;
      }
      __syncthreads();
      double ap = 0.0;
      if (tid < n) {
        for (int c = 0; c < C; c++) ap += prt[c * n + tid];
        apv[tid] = ap;
      }
      const double pap = cta_dot(tid < n ? p * ap : 0.0);
      if (!(pap > 0.0)) { ok = false; break; }
      const double alpha = rz / pap;
      if (tid < n) {
        x += alpha * p;
        r -= alpha * ap;
        rv[tid] = r;
      }
      __syncthreads();
      if (tid < n) z = precond(tid, rv);
      const double rzn = cta_dot(tid < n ? r * z : 0.0);
      if (!(rzn > stop)) { it++; break; }
      const double beta = rzn / rz;
      rz = rzn;
      if (tid < n) { p = z + beta * p; pv[tid] = p; }
      __syncthreads();
    }
  }
  if (tid < n) __stcg(W.xp + tid, ok ? x : 0.0);
  __syncthreads();
  iters_out = it;
  return ok;
}

int pcg_dense_doubles(int Ncf, int threads) {
  const int n = Ncf * 6;
  if (n == 0) return 0;
  const int C = threads / n > 1 ? threads / n : 1;
  const int extra = C * n > Ncf * 36 ? C * n : Ncf * 36;  // partial products / Cholesky scratch
  return n * n + 5 * n + 36 * Ncf + extra;
}

// ------------------------------------------------------------------------------- cameras

template <class Scope>
__device__ void refresh_camRt(const Scope& sc, const BAWin& W, int buf) {
  const int gt = sc.blk() * blockDim.x + threadIdx.x, gstride = sc.nblk() * blockDim.x;
  for (int c = gt; c < W.Nc; c += gstride) {
    const double* q = W.cam[buf] + (size_t)c * 7;
    double R[9];
    quat_to_R(q, R);
    double* o = W.camRt[buf] + (size_t)c * 12;
#pragma unroll
    for (int a = 0; a < 9; a++) o[a] = R[a];
    o[9] = q[4]; o[10] = q[5]; o[11] = q[6];
  }
}

// trial cameras = exp(x_c) * current (free cameras), copy (fixed cameras)
template <class Scope>
__device__ void cam_update(const Scope& sc, const BAWin& W, int cur) {
  const int gt = sc.blk() * blockDim.x + threadIdx.x, gstride = sc.nblk() * blockDim.x;
  const int tr = cur ^ 1;
  for (int c = gt; c < W.Nc; c += gstride) {
    const double* q = W.cam[cur] + (size_t)c * 7;
    double* qo = W.cam[tr] + (size_t)c * 7;
    const int cf = W.cam_free[c];
    if (cf >= 0) {
      double u[6];
#pragma unroll
      for (int a = 0; a < 6; a++) u[a] = __ldcg(W.xp + cf * 6 + a);
      double qn[4], tn[3];
      se3_oplus(u, q, q + 4, qn, tn);
      qo[0] = qn[0]; qo[1] = qn[1]; qo[2] = qn[2]; qo[3] = qn[3];
      qo[4] = tn[0]; qo[5] = tn[1]; qo[6] = tn[2];
    } else {
#pragma unroll
      for (int a = 0; a < 7; a++) qo[a] = q[a];
    }
    double R[9];
    quat_to_R(qo, R);
    double* o = W.camRt[tr] + (size_t)c * 12;
#pragma unroll
    for (int a = 0; a < 9; a++) o[a] = R[a];
    o[9] = qo[4]; o[10] = qo[5]; o[11] = qo[6];
  }
}

// ------------------------------------------------------------------------------- phase BACKSUB

// x_l = Dinv (b_l - sum_i B_i^T (J_i x_ci)); X' = X + x_l; trial robust chi2; landmark part of
// computeScale: sum x_l (lambda x_l + b_l).
template <class Scope>
__device__ void backsub_phase(const Scope& sc, const BAWin& W, int cur, double lambda, bool robust,
                              double delta, double& chi_acc, double& scale_acc) {
  const int lane = threadIdx.x & 31;
  const int wpc = blockDim.x >> 5;
  const int gw = sc.blk() * wpc + (threadIdx.x >> 5);
  const int gstride = sc.nblk() * wpc;
  const int tr = cur ^ 1;
  const double* __restrict__ camRt = W.camRt[cur];
  const double* __restrict__ camRtT = W.camRt[tr];
  const double K[4] = {W.intr[0], W.intr[1], W.intr[2], W.intr[3]};
  for (int l = gw; l < W.Np; l += gstride) {
    const int ps = W.pt_start[l], k = W.pt_start[l + 1] - ps;
    const double X[3] = {W.pts[cur][l * 3], W.pts[cur][l * 3 + 1], W.pts[cur][l * 3 + 2]};
    double c3[3] = {0, 0, 0};
    for (int i = lane; i < k; i += 32) {
      const int o = ps + i;
      if (W.level[o]) continue;
      const int c = W.ocam[o];
      const int cf = W.cam_free[c];
      if (cf < 0) continue;
      const double* Rt = camRt + (size_t)c * 12;
      double pc[3], pz[3], e0, e1, w, Jp[12], Jx[6];
      map_point(Rt, X, pc);
      const double2 uv = *reinterpret_cast<const double2*>(W.uv + (size_t)o * 2);
      const double e2 = edge_error(pc, uv.x, uv.y, K, e0, e1, pz);
      huber_rho(e2, delta, robust, w);
      edge_jac_pose(pz, K, Jp);
      edge_jac_point(Rt, pz, K, Jx);
      double s0 = 0, s1 = 0;
#pragma unroll
      for (int a = 0; a < 6; a++) {
        const double xa = __ldcg(W.xp + cf * 6 + a);
        s0 += Jp[a] * xa;
        s1 += Jp[6 + a] * xa;
      }
      s0 *= w; s1 *= w;
#pragma unroll
      for (int a = 0; a < 3; a++) c3[a] += Jx[a] * s0 + Jx[3 + a] * s1;
    }
#pragma unroll
    for (int a = 0; a < 3; a++) c3[a] = warp_sum(c3[a]);
    const double* Di = W.Dinv + (size_t)l * 6;
    const double b0 = W.bl[(size_t)l * 3], b1 = W.bl[(size_t)l * 3 + 1], b2 = W.bl[(size_t)l * 3 + 2];
    const double r0 = b0 - c3[0], r1 = b1 - c3[1], r2 = b2 - c3[2];
    const double x0 = Di[0] * r0 + Di[1] * r1 + Di[2] * r2;
    const double x1 = Di[1] * r0 + Di[3] * r1 + Di[4] * r2;
    const double x2 = Di[2] * r0 + Di[4] * r1 + Di[5] * r2;
    const double Xn[3] = {X[0] + x0, X[1] + x1, X[2] + x2};
    if (lane == 0) {
      W.pts[tr][l * 3] = Xn[0]; W.pts[tr][l * 3 + 1] = Xn[1]; W.pts[tr][l * 3 + 2] = Xn[2];
      scale_acc += x0 * (lambda * x0 + b0) + x1 * (lambda * x1 + b1) + x2 * (lambda * x2 + b2);
    }
    for (int i = lane; i < k; i += 32) {
      const int o = ps + i;
      if (W.level[o]) continue;
      const double* Rt = camRtT + (size_t)W.ocam[o] * 12;
      double pc[3], e0, e1, w;
      map_point(Rt, Xn, pc);
      const double2 uv = *reinterpret_cast<const double2*>(W.uv + (size_t)o * 2);
      const double e2 = edge_error(pc, uv.x, uv.y, K, e0, e1);
      chi_acc += huber_rho(e2, delta, robust, w);
    }
  }
}

// ------------------------------------------------------------------------------- packed phases
//
// acc_mode 2/3 (every point has <= 32 observations, <= 64 stored blocks, <= 16 free cameras — every
// window the reference can produce): a warp takes a GROUP of consecutive points whose observations
// fill its 32 lanes (lane = observation, so uv / camera index / level loads are coalesced and ~95 %
// of the lanes work, against ~25 % with one point per warp), per-point sums are formed in shared
// memory in observation order, and the Schur complement is accumulated S-STATIONARY: lane b owns
// reduced-system block b (and b+32) in REGISTERS for all the groups of the warp and visits, per
// point, the two observation slots of its camera pair.  No read-modify-write traffic, no atomics;
// the register blocks are flushed once per phase and summed in warp order, then CTA order.


// Build the per-(group, lane) records from the CSR arrays and the current edge levels.
template <class Scope>
__device__ void pack_records(const Scope& sc, const BAWin& W) {
  const int lane = threadIdx.x & 31;
  const int wpc = blockDim.x >> 5;
  const int gw = sc.blk() * wpc + (threadIdx.x >> 5);
  const int gstride = sc.nblk() * wpc;
  for (int g = gw; g < W.n_grp; g += gstride) {
    const int p0 = W.grp_pt[g];
    const int np = W.grp_pt[g + 1] - p0;
    const int o0 = W.pt_start[p0];
    const int nobs = W.pt_start[p0 + np] - o0;
    ObsRec r;
    r.u = 0.0; r.v = 0.0; r.pl = p0; r.c = 0; r.s0 = 0; r.s1 = 0; r.pi = 0; r.cf = -1;
    r.np = (unsigned char)np; r.lev = 1; r.valid = 0; r.pad = 0;
    if (lane < nobs) {
      const int o = o0 + lane;
      const int pl = W.opt[o];
      r.u = W.uv[(size_t)o * 2];
      r.v = W.uv[(size_t)o * 2 + 1];
      r.pl = pl;
      r.c = W.ocam[o];
      r.s0 = (unsigned char)(W.pt_start[pl] - o0);
      r.s1 = (unsigned char)(W.pt_start[pl + 1] - o0);
      r.pi = (unsigned char)(pl - p0);
      r.cf = (signed char)W.cam_free[r.c];
      r.lev = W.level[o];
      r.valid = 1;
    }
    W.rec[(size_t)g * 32 + lane] = r;
  }
}

template <bool DIAG, int NB, class Scope>
__device__ void lin_phase_packed(const Scope& sc, const BAWin& W, int cur, double lambda, bool robust,
                                 double delta, PackStage st, WorkArea wa, double& chi_acc,
                                 double& maxdiag_acc) {
  const int lane = threadIdx.x & 31;
  const int wpc = blockDim.x >> 5;
  const int gw = sc.blk() * wpc + (threadIdx.x >> 5);
  const int gstride = sc.nblk() * wpc;
  const double* __restrict__ camRt = W.camRt[cur];
  const double* __restrict__ pts = W.pts[cur];
  const double K[4] = {W.intr[0], W.intr[1], W.intr[2], W.intr[3]};
  // the descriptor lives in global memory: take its fields into registers once
  const ObsRec* rec = W.rec;
  double* __restrict__ Dinv_out = W.Dinv;
  double* __restrict__ bl_out = W.bl;
  const int Ncf = W.Ncf, n_grp = W.n_grp, nblk = W.nblk;
  // blocks owned by this lane: dense upper layout, block b = row_ptr[ci] + (cj - ci)
  int bci[NB], bcj[NB];
#pragma unroll
  for (int nb = 0; nb < NB; nb++) {
    const int blk = lane + 32 * nb;
    bci[nb] = -1; bcj[nb] = -1;
    if (blk < nblk) {
      int ci = 0;
      while (W.row_ptr[ci + 1] <= blk) ci++;
      bci[nb] = ci;
      bcj[nb] = ci + (blk - W.row_ptr[ci]);
    }
  }
  double accS[NB][36];
#pragma unroll
  for (int nb = 0; nb < NB; nb++)
#pragma unroll
    for (int e = 0; e < 36; e++) accS[nb][e] = 0.0;
  double accb[12];  // b_s | b_p of camera `lane` (DIAG: diag(Hpp) in the first 6)
#pragma unroll
  for (int e = 0; e < 12; e++) accb[e] = 0.0;
  // absent (point, camera) entries of the slot table point at the all-zero slot, so the pair loop
  // below runs without divergent branches: an absent pair adds exactly +0 to the accumulators
  for (int a = lane; a < kPackFields; a += 32) st.f[a * kPackSlots + kPackZero] = 0.0;
  int bsi[NB], bsj[NB];  // table columns of the lane's block(s); lanes without a block read column 0
  bool has[NB];
#pragma unroll
  for (int nb = 0; nb < NB; nb++) {
    has[nb] = bci[nb] >= 0;
    bsi[nb] = has[nb] ? bci[nb] : 0;
    bsj[nb] = has[nb] ? bcj[nb] : 0;
  }
  const int bcol = lane & (kPackCam - 1);  // camera column of the b_s / b_p accumulation (lanes < 16)
  const bool diag_lane = NB == 1 && has[0] && bci[0] == bcj[0];

  // software pipeline: the record of the NEXT group is fetched (one level of loads) while the
  // current group is processed, so the L2 round trip overlaps the arithmetic
  double* const rbuf = st.recbuf();
  if (gw < n_grp) rec_prefetch(rbuf, rec, gw, lane);
  for (int g = gw; g < n_grp; g += gstride) {
    const RecRegs rr = rec_take(rbuf, lane);
    const double2 uv = make_double2(__hiloint2double(rr.a.y, rr.a.x), __hiloint2double(rr.a.w, rr.a.z));
    const int pl = rr.b.x, c = rr.b.y;
    const int s0 = rr.b.z & 255, s1 = (rr.b.z >> 8) & 255, pi = (rr.b.z >> 16) & 255;
    const int cf_ld = rr.b.z >> 24;  // arithmetic shift keeps the sign of the int8
    const int np = rr.b.w & 255, lev = (rr.b.w >> 8) & 255;
    const bool valid = (rr.b.w >> 16) & 1;
    reinterpret_cast<int*>(st.slot)[lane] = 0x20202020;  // 32 * 16 bytes, every entry = kPackZero
    reinterpret_cast<int*>(st.slot)[lane + 32] = 0x20202020;
    reinterpret_cast<int*>(st.slot)[lane + 64] = 0x20202020;
    reinterpret_cast<int*>(st.slot)[lane + 96] = 0x20202020;
    __syncwarp();
    const double X[3] = {pts[pl * 3], pts[pl * 3 + 1], pts[pl * 3 + 2]};
    if (g + gstride < n_grp) rec_prefetch(rbuf, rec, g + gstride, lane);
    double hc[6] = {0, 0, 0, 0, 0, 0}, blc[3] = {0, 0, 0};
    int cf = -1;
    double B[6];
    if (valid && !lev) {
      const double* Rt = camRt + (size_t)c * 12;
      double pc[3], pz[3], e0, e1, w;
      map_point(Rt, X, pc);
      const double e2 = edge_error(pc, uv.x, uv.y, K, e0, e1, pz);
      chi_acc += huber_rho(e2, delta, robust, w);
      double Jx[6];
      edge_jac_point(Rt, pz, K, Jx);
#pragma unroll
      for (int a = 0; a < 6; a++) B[a] = w * Jx[a];
      hc[0] = B[0] * Jx[0] + B[3] * Jx[3];
      hc[1] = B[0] * Jx[1] + B[3] * Jx[4];
      hc[2] = B[0] * Jx[2] + B[3] * Jx[5];
      hc[3] = B[1] * Jx[1] + B[4] * Jx[4];
      hc[4] = B[1] * Jx[2] + B[4] * Jx[5];
      hc[5] = B[2] * Jx[2] + B[5] * Jx[5];
#pragma unroll
      for (int a = 0; a < 3; a++) blc[a] = -(B[a] * e0 + B[3 + a] * e1);
      cf = cf_ld;
      if (cf >= 0) {
        double Jp[12];
        edge_jac_pose(pz, K, Jp);
#pragma unroll
        for (int a = 0; a < 12; a++) st.Jp(a, lane) = Jp[a];
        st.w(lane) = w;
        if (!DIAG) {
#pragma unroll
          for (int a = 0; a < 6; a++) st.B(a, lane) = B[a];
          st.we(0, lane) = w * e0;
          st.we(1, lane) = w * e1;
        }
        st.slot[pi * kPackCam + cf] = (signed char)lane;
      }
    }
#pragma unroll
    for (int a = 0; a < 6; a++) st.h(a, lane) = hc[a];
    if (!DIAG) {
#pragma unroll
      for (int a = 0; a < 3; a++) st.bl(a, lane) = blc[a];
    }
    __syncwarp();
    // per-point sums in observation order (every lane for its own point)
    // (uniform trip count: lanes past the end of their range add the all-zero slot)
    double h[6] = {0, 0, 0, 0, 0, 0}, bl[3] = {0, 0, 0};
    {
      const int len = valid ? s1 - s0 : 0;
      const int maxlen = __reduce_max_sync(0xffffffffu, len);
      // unrolled by four: the shared-memory loads of four observations are in flight before the (ordered) additions
#pragma unroll 4
      for (int t = 0; t < maxlen; t++) {
        const int s = t < len ? s0 + t : kPackZero;
#pragma unroll
        for (int a = 0; a < 6; a++) h[a] += st.h(a, s);
        if (!DIAG) {
#pragma unroll
          for (int a = 0; a < 3; a++) bl[a] += st.bl(a, s);
        }
      }
    }
    if (DIAG) {
      if (valid) maxdiag_acc = fmax(maxdiag_acc, fmax(fabs(h[0]), fmax(fabs(h[3]), fabs(h[5]))));
      __syncwarp();
      for (int q = 0; q < np; q++) {
        const int s = lane < kPackCam ? st.slot[q * kPackCam + bcol] : kPackZero;
        const double w = st.w(s);
#pragma unroll
        for (int a = 0; a < 6; a++) {
          const double j0 = st.Jp(a, s), j1 = st.Jp(6 + a, s);
          accb[a] += w * (j0 * j0 + j1 * j1);
        }
      }
      __syncwarp();
      continue;
    }
    if (valid) {
      double Di[6];
      const double hl[6] = {h[0] + lambda, h[1], h[2], h[3] + lambda, h[4], h[5] + lambda};
      sym3_inverse(hl, Di);
      if (lane == s0) {
#pragma unroll
        for (int a = 0; a < 6; a++) Dinv_out[(size_t)pl * 6 + a] = Di[a];
#pragma unroll
        for (int a = 0; a < 3; a++) bl_out[(size_t)pl * 3 + a] = bl[a];
      }
      if (cf >= 0) {
        double A[6];
#pragma unroll
        for (int r = 0; r < 2; r++) {
          const double b0 = B[r * 3], b1 = B[r * 3 + 1], b2 = B[r * 3 + 2];
          A[r * 3 + 0] = b0 * Di[0] + b1 * Di[1] + b2 * Di[2];
          A[r * 3 + 1] = b0 * Di[1] + b1 * Di[3] + b2 * Di[4];
          A[r * 3 + 2] = b0 * Di[2] + b1 * Di[4] + b2 * Di[5];
        }
#pragma unroll
        for (int a = 0; a < 6; a++) st.A(a, lane) = A[a];
        st.g(0, lane) = st.we(0, lane) + (A[0] * bl[0] + A[1] * bl[1] + A[2] * bl[2]);
        st.g(1, lane) = st.we(1, lane) + (A[3] * bl[0] + A[4] * bl[1] + A[5] * bl[2]);
      }
    }
    __syncwarp();
    // S-stationary accumulation: for every point of the group, the lane's camera pair(s).  The slot
    // bytes of point q+1 are fetched while point q is processed.
    // NB == 1: the lane that owns the DIAGONAL block (ci, ci) also accumulates b_s / b_p of camera
    // ci — it already holds J_ci of the point in registers for its block, so the gradients cost no
    // extra shared-memory traffic.  NB == 2 keeps the camera = lane layout for the gradients.
    int nsi[NB], nsj[NB], nsb = kPackZero;
    {
      const signed char* sl = st.slot;
#pragma unroll
      for (int nb = 0; nb < NB; nb++) { nsi[nb] = sl[bsi[nb]]; nsj[nb] = sl[bsj[nb]]; }
      if (NB > 1) nsb = sl[bcol];
    }
    for (int q = 0; q < np; q++) {
      int csi[NB], csj[NB];
#pragma unroll
      for (int nb = 0; nb < NB; nb++) { csi[nb] = nsi[nb]; csj[nb] = nsj[nb]; }
      const int csb = nsb;
      {
        const signed char* sl = st.slot + (q + 1 < 32 ? q + 1 : 31) * kPackCam;
#pragma unroll
        for (int nb = 0; nb < NB; nb++) { nsi[nb] = sl[bsi[nb]]; nsj[nb] = sl[bsj[nb]]; }
        if (NB > 1) nsb = sl[bcol];
      }
#pragma unroll
      for (int nb = 0; nb < NB; nb++) {
        const bool ok = has[nb] && !((csi[nb] | csj[nb]) & kPackZero);
        const int si = ok ? csi[nb] : kPackZero, sj = ok ? csj[nb] : kPackZero;
        double M[4];
        {
          const double a0 = st.A(0, si), a1 = st.A(1, si), a2 = st.A(2, si);
          const double a3 = st.A(3, si), a4 = st.A(4, si), a5 = st.A(5, si);
          const double b0 = st.B(0, sj), b1 = st.B(1, sj), b2 = st.B(2, sj);
          const double b3 = st.B(3, sj), b4 = st.B(4, sj), b5 = st.B(5, sj);
          const double wd = si == sj ? st.w(si) : 0.0;
          M[0] = wd - (a0 * b0 + a1 * b1 + a2 * b2);
          M[1] = -(a0 * b3 + a1 * b4 + a2 * b5);
          M[2] = -(a3 * b0 + a4 * b1 + a5 * b2);
          M[3] = wd - (a3 * b3 + a4 * b4 + a5 * b5);
        }
        double T[12];
#pragma unroll
        for (int b = 0; b < 6; b++) {
          const double j0 = st.Jp(b, sj), j1 = st.Jp(6 + b, sj);
          T[b] = M[0] * j0 + M[1] * j1;
          T[6 + b] = M[2] * j0 + M[3] * j1;
        }
        // gradient terms ride on the diagonal lane (NB == 1): slot si when the block is diagonal
        const int sg = (NB == 1 && diag_lane) ? si : kPackZero;
        double g0 = 0.0, g1 = 0.0, w0 = 0.0, w1 = 0.0;
        if (NB == 1) { g0 = st.g(0, sg); g1 = st.g(1, sg); w0 = st.we(0, sg); w1 = st.we(1, sg); }
#pragma unroll
        for (int a = 0; a < 6; a++) {
          const double j0 = st.Jp(a, si), j1 = st.Jp(6 + a, si);
#pragma unroll
          for (int b = 0; b < 6; b++)  // two chained FMAs per element (not mul + fma + add)
            accS[nb][a * 6 + b] = fma(j1, T[6 + b], fma(j0, T[b], accS[nb][a * 6 + b]));
          if (NB == 1) {
            accb[a] = fma(-j1, g1, fma(-j0, g0, accb[a]));
            accb[6 + a] = fma(-j1, w1, fma(-j0, w0, accb[6 + a]));
          }
        }
      }
      if (NB > 1) {
        const int s = lane < kPackCam ? csb : kPackZero;
        const double g0 = st.g(0, s), g1 = st.g(1, s), w0 = st.we(0, s), w1 = st.we(1, s);
#pragma unroll
        for (int a = 0; a < 6; a++) {
          const double j0 = st.Jp(a, s), j1 = st.Jp(6 + a, s);
          accb[a] = fma(-j1, g1, fma(-j0, g0, accb[a]));
          accb[6 + a] = fma(-j1, w1, fma(-j0, w0, accb[6 + a]));
        }
      }
    }
    __syncwarp();
  }
  // flush the register accumulators into this warp's work area (aliases the staging fields)
  double* acc = wa.base + (size_t)(threadIdx.x >> 5) * wa.stride;
  __syncwarp();
  if (DIAG) {
    if (lane < W.Ncf) {
#pragma unroll
      for (int a = 0; a < 6; a++) acc[lane * 6 + a] = accb[a];
    }
    cta_reduce_copies(sc, W, wa, 0, W.Ncf * 6);
    return;
  }
#pragma unroll
  for (int nb = 0; nb < NB; nb++) {
    const int blk = lane + 32 * nb;
    if (blk < W.nblk) {
#pragma unroll
      for (int e = 0; e < 36; e++) acc[(size_t)blk * 36 + e] = accS[nb][e];
    }
  }
  if (NB == 1 ? diag_lane : lane < W.Ncf) {
    const int cb = NB == 1 ? bci[0] : lane;  // camera whose gradients this lane holds
#pragma unroll
    for (int a = 0; a < 6; a++) {
      acc[(size_t)W.nblk * 36 + cb * 6 + a] = accb[a];
      acc[(size_t)W.nblk * 36 + W.Ncf * 6 + cb * 6 + a] = accb[6 + a];
    }
  }
  cta_reduce_copies(sc, W, wa, 0, W.acc_len);
}

template <class Scope>
__device__ void backsub_phase_packed(const Scope& sc, const BAWin& W, int cur, double lambda, bool robust,
                                     double delta, PackStage st, double& chi_acc, double& scale_acc) {
  const int lane = threadIdx.x & 31;
  const int wpc = blockDim.x >> 5;
  const int gw = sc.blk() * wpc + (threadIdx.x >> 5);
  const int gstride = sc.nblk() * wpc;
  const int tr = cur ^ 1;
  const double* __restrict__ camRt = W.camRt[cur];
  const double* __restrict__ camRtT = W.camRt[tr];
  const double K[4] = {W.intr[0], W.intr[1], W.intr[2], W.intr[3]};
  const ObsRec* rec = W.rec;
  const double* __restrict__ pts_cur = W.pts[cur];
  double* __restrict__ pts_tr = W.pts[tr];
  const double* __restrict__ Dinv_in = W.Dinv;
  const double* __restrict__ bl_in = W.bl;
  const double* __restrict__ xp = W.xp;
  const int n_grp = W.n_grp;
  double* const rbuf = st.recbuf();
  if (lane < 3) st.h(lane, kPackZero) = 0.0;  // all-zero slot for the uniform per-point sums
  // the pose increments (<= 16 free cameras x 6) into the warp's staging area (the Jp fields are free here)
  double* const xs = st.f;
  for (int e = lane; e < W.Ncf * 6; e += 32) xs[e] = __ldcg(xp + e);
  __syncwarp();
  // Pipeline: record one group ahead (32 B per lane), its point index two groups ahead (4 B per lane,
  // into the slot-table area, which BACKSUB does not use) so that the point position X of the NEXT
  // group — the operand every lane needs first — can be copied into shared memory (free Jp fields)
  // while the current group is processed.  No shared memory is added: L1 capacity matters here.
  int* const plbuf = reinterpret_cast<int*>(st.slot);      // [2][32]
  double* const xb = st.f + 4 * kPackSlots + lane * 3;      // fields 4.. are free in BACKSUB
  auto pl_prefetch = [&](int par, int g2) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(plbuf + par * 32 + lane);
    const size_t src = __cvta_generic_to_global(&rec[(size_t)g2 * 32 + lane].pl);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
  };
  auto x_prefetch = [&](int npl) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const unsigned dst = (unsigned)__cvta_generic_to_shared(xb + a);
      const size_t src = __cvta_generic_to_global(pts_cur + npl * 3 + a);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
    }
  };
  int par = 0;
  if (gw < n_grp) {
    rec_prefetch(rbuf, rec, gw, lane);
    pl_prefetch(0, gw);
    if (gw + gstride < n_grp) pl_prefetch(1, gw + gstride);
    asm volatile("cp.async.wait_all;" ::: "memory");
    x_prefetch(plbuf[lane]);
  }
  for (int g = gw; g < n_grp; g += gstride, par ^= 1) {
    const RecRegs rr = rec_take(rbuf, lane);  // waits for every outstanding copy
    const double X[3] = {xb[0], xb[1], xb[2]};
    const int npl = plbuf[(par ^ 1) * 32 + lane];
    const double2 uv = make_double2(__hiloint2double(rr.a.y, rr.a.x), __hiloint2double(rr.a.w, rr.a.z));
    const int pl = rr.b.x, c = rr.b.y;
    const int s0 = rr.b.z & 255, s1 = (rr.b.z >> 8) & 255;
    const int cf_ld = rr.b.z >> 24;
    const bool valid = (rr.b.w >> 16) & 1;
    const bool active = valid && !((rr.b.w >> 8) & 255);
    if (g + gstride < n_grp) {
      rec_prefetch(rbuf, rec, g + gstride, lane);
      x_prefetch(npl);
      if (g + 2 * gstride < n_grp) pl_prefetch(par, g + 2 * gstride);
    }
    double Di[6], bb[3];
#pragma unroll
    for (int a = 0; a < 6; a++) Di[a] = Dinv_in[(size_t)pl * 6 + a];
#pragma unroll
    for (int a = 0; a < 3; a++) bb[a] = bl_in[(size_t)pl * 3 + a];
    double c3[3] = {0, 0, 0};
    if (active) {
      const int cf = cf_ld;
      if (cf >= 0) {
        const double* Rt = camRt + (size_t)c * 12;
        double pc[3], pz[3], e0, e1, w, Jp[12], Jx[6];
        map_point(Rt, X, pc);
        const double e2 = edge_error(pc, uv.x, uv.y, K, e0, e1, pz);
        huber_rho(e2, delta, robust, w);
        edge_jac_pose(pz, K, Jp);
        edge_jac_point(Rt, pz, K, Jx);
        double t0 = 0, t1 = 0;
#pragma unroll
        for (int a = 0; a < 6; a++) {
          const double xa = xs[cf * 6 + a];
          t0 += Jp[a] * xa;
          t1 += Jp[6 + a] * xa;
        }
        t0 *= w; t1 *= w;
#pragma unroll
        for (int a = 0; a < 3; a++) c3[a] = Jx[a] * t0 + Jx[3 + a] * t1;
      }
    }
    __syncwarp();
#pragma unroll
    for (int a = 0; a < 3; a++) st.h(a, lane) = c3[a];
    __syncwarp();
    double Xn[3] = {X[0], X[1], X[2]};
    double cs[3] = {0, 0, 0};
    {
      const int len = valid ? s1 - s0 : 0;
      const int maxlen = __reduce_max_sync(0xffffffffu, len);
#pragma unroll 4
      for (int t = 0; t < maxlen; t++) {
        const int s = t < len ? s0 + t : kPackZero;
#pragma unroll
        for (int a = 0; a < 3; a++) cs[a] += st.h(a, s);
      }
    }
    if (valid) {
      const double b0 = bb[0], b1 = bb[1], b2 = bb[2];
      const double r0 = b0 - cs[0], r1 = b1 - cs[1], r2 = b2 - cs[2];
      const double x0 = Di[0] * r0 + Di[1] * r1 + Di[2] * r2;
      const double x1 = Di[1] * r0 + Di[3] * r1 + Di[4] * r2;
      const double x2 = Di[2] * r0 + Di[4] * r1 + Di[5] * r2;
      Xn[0] += x0; Xn[1] += x1; Xn[2] += x2;
      if (lane == s0) {
        pts_tr[pl * 3] = Xn[0]; pts_tr[pl * 3 + 1] = Xn[1]; pts_tr[pl * 3 + 2] = Xn[2];
        scale_acc += x0 * (lambda * x0 + b0) + x1 * (lambda * x1 + b1) + x2 * (lambda * x2 + b2);
      }
    }
    if (active) {
      const double* Rt = camRtT + (size_t)c * 12;
      double pc[3], e0, e1, w;
      map_point(Rt, Xn, pc);
      const double e2 = edge_error(pc, uv.x, uv.y, K, e0, e1);
      chi_acc += huber_rho(e2, delta, robust, w);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------- LM driver

constexpr int kPartWidth = kBAPartWidth;

struct LMResult { int iters, trials, pcg_iters; double chi, lambda; };

template <class Scope>
__device__ void zero_system(const Scope& sc, const BAWin& W, bool diag_only, double lambda) {
  const int gt = sc.blk() * blockDim.x + threadIdx.x, gstride = sc.nblk() * blockDim.x;
  if (diag_only) {
    for (int i = gt; i < W.Ncf * 6; i += gstride) W.hdiag[i] = 0.0;
    return;
  }
  for (int i = gt; i < W.nblk * 36; i += gstride) W.S[i] = 0.0;
  // BlockSolver::setLambda: + lambda on the diagonal of every pose block (col[row_ptr[i]] == i);
  // disjoint from the zeroing above only after a barrier, so do it in the same thread ordering
  for (int i = gt; i < W.Ncf * 6; i += gstride) { W.bs[i] = 0.0; W.bp[i] = 0.0; }
  (void)lambda;
}

// S_ii += lambda I. Must run after zero_system is visible (the caller syncs in between).
template <class Scope>
__device__ void damp_diagonal(const Scope& sc, const BAWin& W, double lambda) {
  const int gt = sc.blk() * blockDim.x + threadIdx.x, gstride = sc.nblk() * blockDim.x;
  for (int i = gt; i < W.Ncf * 6; i += gstride) {
    const int c = i / 6, a = i - c * 6;
    atomicAdd(&W.S[(size_t)W.row_ptr[c] * 36 + a * 7], lambda);
  }
}

// SparseOptimizer::optimize(n_iter) with OptimizationAlgorithmLevenberg.
// `cur` is the buffer holding the current estimate (updated on accept); `have_trial` tells the
// caller whether buffer cur^1 / `last_eval` holds the state of the last computeActiveErrors().
// MODE = BAWin::acc_mode: 0 global atomics, 1 shared-memory RMW copies, 2 / 3 packed groups with
// 1 / 2 register-resident blocks per lane.
template <int MODE, class Scope>
__device__ LMResult lm_optimize(const Scope& sc, const BAWin& W, const BARun& run, int n_iter,
                                bool robust, int& cur, int& last_eval, WarpStage st, PackStage pst,
                                WorkArea wa, double* pcg_sm, int& parity, double* red,
                                double* chi_initial) {
  constexpr bool STEREO = MODE == 5 || MODE == 6;      // modes 1 / 0 with 3-row (stereo-capable) edges
  constexpr bool SMEM = (MODE >= 1 && MODE <= 3) || MODE == 5;
  constexpr bool PACKED = MODE == 2 || MODE == 3;
  constexpr int NB = MODE == 3 ? 2 : 1;
  const StereoPar sp = {run.delta_s};
  WarpStageS sts;
  sts.f = st.f; sts.cf = st.cf; sts.kmax = st.kmax;
  const bool timer = W.stats == run.timing_stats && sc.blk() == 0 && threadIdx.x == 0;
  LMResult res = {0, 0, 0, 0.0, 0.0};
  double lambda = 0.0, ni = 2.0;
  double currentChi = 0.0;
  double dummy[1];
  // one CTA (no scope barriers) only while its rows' blocks fit the shared-memory cache
  const int use_single_cta_pcg = (W.Ncf <= (int)(blockDim.x >> 5) * kPcgRowsPerWarp) &&
                                 ((long long)(2 * W.nblk - W.Ncf) * 36 <= (long long)(blockDim.x >> 5) * wa.stride);
  double* flags = W.part + (size_t)2 * sc.nblk() * kPartWidth;  // behind the two reduction buffers
  for (int it = 0; it < n_iter; it++) {
    if (it == 0) {
      // computeLambdaInit: tau * max diagonal entry of the (undamped) Hessian
      if (!SMEM) {
        zero_system(sc, W, true, 0.0);
        sc.sync();
      }
      double chi = 0.0, mx = 0.0;
      {
        BA_T0();
        if (PACKED) lin_phase_packed<true, NB>(sc, W, cur, 0.0, robust, run.delta, pst, wa, chi, mx);
        else if (STEREO) lin_phase_s<true, SMEM>(sc, W, cur, 0.0, robust, run.delta, sp, sts, wa, chi, mx);
        else lin_phase<true, SMEM>(sc, W, cur, 0.0, robust, run.delta, st, wa, chi, mx);
        BA_T1(0);
      }
      double s1[1] = {chi}, m1[1] = {mx};
      scope_reduce<1, 1>(sc, s1, m1, W.part, parity, red);
      double mp = 0.0;
      for (int i = threadIdx.x; i < W.Ncf * 6; i += blockDim.x) {
        double hd;
        if (SMEM) {
          hd = 0.0;
          for (int b = 0; b < sc.nblk(); b++) hd += __ldcg(W.Spart + (size_t)b * W.acc_len + i);
        } else {
          hd = __ldcg(W.hdiag + i);
        }
        mp = fmax(mp, fabs(hd));
      }
      double z1[1] = {0.0}, m2[1] = {mp};
      {  // CTA-local max (every CTA sees the same hdiag)
        CtaScope cs;
        int par2 = 0;
        scope_reduce<1, 1>(cs, z1, m2, nullptr, par2, red);
      }
      lambda = 1e-5 * fmax(m1[0], m2[0]);
      ni = 2.0;
      if (chi_initial) *chi_initial = s1[0];
    }
    double rho = 0.0;
    int qmax = 0;
    bool lambda_bad = false;
    do {
      if (!SMEM) {
        zero_system(sc, W, false, lambda);
        sc.sync();
        damp_diagonal(sc, W, lambda);
      } else if (it == 0 && qmax == 0) {
        sc.sync();  // every CTA has read the DIAG partials before Spart is overwritten
      }
      double chi = 0.0, mx = 0.0;
      {
        BA_T0();
        if (PACKED) lin_phase_packed<false, NB>(sc, W, cur, lambda, robust, run.delta, pst, wa, chi, mx);
        else if (STEREO) lin_phase_s<false, SMEM>(sc, W, cur, lambda, robust, run.delta, sp, sts, wa, chi, mx);
        else lin_phase<false, SMEM>(sc, W, cur, lambda, robust, run.delta, st, wa, chi, mx);
        BA_T1(1);
      }
      double s1[1] = {chi};
      {
        BA_T0();
        scope_reduce<1, 0>(sc, s1, dummy, W.part, parity, red);  // also publishes S, bs, bp
        BA_T1(2);
      }
      currentChi = s1[0];
      int pcg_it = 0;
      bool ok2;
      const long long _tp = clock64();
      if (SMEM) {
        if (sc.blk() == 0) {
          ok2 = pcg_dense_smem(sc, W, lambda, run.pcg_tol, run.pcg_max_iter, run.dense_solver, pcg_sm, pcg_it);
          if (threadIdx.x == 0) { __stcg(flags, ok2 ? 1.0 : 0.0); __stcg(flags + 1, (double)pcg_it); }
          // the CTA that solved also moves the (<= 16) cameras: the scope barrier below then publishes the solution
          // flags AND the trial cameras, instead of a second barrier after a camera phase spread over the scope
          __syncthreads();
          if (__ldcg(flags) > 0.5) cam_update(CtaScope(), W, cur);
        }
        sc.sync();
        ok2 = __ldcg(flags) > 0.5;
        pcg_it = (int)__ldcg(flags + 1);
      } else if (use_single_cta_pcg) {
        // small reduced system: one CTA iterates with __syncthreads only, the others wait
        if (sc.blk() == 0) {
          CtaScope cs;
          int par2 = 0;
          ok2 = pcg_phase(cs, W, run.pcg_tol, run.pcg_max_iter, nullptr, par2, red, pcg_it,
                          wa.base + (size_t)(threadIdx.x >> 5) * wa.stride, wa.stride);
          if (threadIdx.x == 0) { __stcg(flags, ok2 ? 1.0 : 0.0); __stcg(flags + 1, (double)pcg_it); }
        }
        sc.sync();
        ok2 = __ldcg(flags) > 0.5;
        pcg_it = (int)__ldcg(flags + 1);
      } else {
        ok2 = pcg_phase(sc, W, run.pcg_tol, run.pcg_max_iter, W.part, parity, red, pcg_it,
                        wa.base + (size_t)(threadIdx.x >> 5) * wa.stride, wa.stride);
        sc.sync();
      }
      res.pcg_iters += pcg_it;
      if (timer) g_ba_timing[3] += (unsigned long long)(clock64() - _tp);
      double tempChi = 1.7976931348623157e308;
      double scale = 0.0;
      if (ok2) {
        if (!SMEM) {
          BA_T0();
          cam_update(sc, W, cur);
          sc.sync();
          BA_T1(4);
        }
        double tchi = 0.0, sc_l = 0.0;
        {
          BA_T0();
          if (PACKED) backsub_phase_packed(sc, W, cur, lambda, robust, run.delta, pst, tchi, sc_l);
          else if (STEREO) backsub_phase_s(sc, W, cur, lambda, robust, run.delta, sp, tchi, sc_l);
          else backsub_phase(sc, W, cur, lambda, robust, run.delta, tchi, sc_l);
          BA_T1(5);
        }
        const long long _t6 = clock64();
        {  // pose part of computeScale: sum x (lambda x + b)
          const int gt = sc.blk() * blockDim.x + threadIdx.x, gstride = sc.nblk() * blockDim.x;
          for (int i = gt; i < W.Ncf * 6; i += gstride) {
            const double x = __ldcg(W.xp + i);
            sc_l += x * (lambda * x + __ldcg(W.bp + i));
          }
        }
        double s2[2] = {tchi, sc_l};
        scope_reduce<2, 0>(sc, s2, dummy, W.part, parity, red);
        tempChi = s2[0];
        scale = s2[1];
        last_eval = cur ^ 1;
        if (timer) g_ba_timing[6] += (unsigned long long)(clock64() - _t6);
      }
      rho = (currentChi - tempChi) / (scale + 1e-3);
      if (rho > 0 && isfinite(tempChi)) {
        double alpha = 1. - pow((2 * rho - 1), 3);
        alpha = fmin(alpha, 2. / 3.);
        lambda *= fmax(1. / 3., alpha);
        ni = 2;
        currentChi = tempChi;
        cur ^= 1;  // discardTop(): the trial state becomes the estimate
      } else {
        lambda *= ni;
        ni *= 2;  // pop(): keep `cur`
        if (!isfinite(lambda)) { lambda_bad = true; qmax++; break; }
      }
      qmax++;
    } while (rho < 0 && qmax < 10);
    res.trials += qmax;
    res.iters = it + 1;
    if (qmax == 10 || rho == 0 || lambda_bad || !isfinite(lambda)) break;
  }
  res.chi = currentChi;
  res.lambda = lambda;
  return res;
}

// ------------------------------------------------------------------------------- whole window

// setEstimate(SE3Quat(q, p).inverse()) (src/g2o_optimization.cc:45), points, levels, derived R|t.
template <class Scope>
__device__ void window_init(const Scope& sc, const BAWin& W) {
  const int gt = sc.blk() * blockDim.x + threadIdx.x, gstride = sc.nblk() * blockDim.x;
  for (int c = gt; c < W.Nc; c += gstride) {
    double q[4], t[3], qi[4], ti[3];
    const double* in = W.pose_in + (size_t)c * 7;
    q[0] = in[0]; q[1] = in[1]; q[2] = in[2]; q[3] = in[3];
    t[0] = in[4]; t[1] = in[5]; t[2] = in[6];
    quat_normalize_w(q);
    se3_inverse(q, t, qi, ti);
    double* o = W.cam[0] + (size_t)c * 7;
    o[0] = qi[0]; o[1] = qi[1]; o[2] = qi[2]; o[3] = qi[3];
    o[4] = ti[0]; o[5] = ti[1]; o[6] = ti[2];
  }
  for (int i = gt; i < W.Np * 3; i += gstride) {
    const double v = W.pts_in[i];
    W.pts[0][i] = v;
    W.pts[1][i] = v;  // points without observations are never rewritten by the packed BACKSUB
  }
  for (int o = gt; o < W.No; o += gstride) W.level[o] = 0;
  sc.sync();
  refresh_camRt(sc, W, 0);
  sc.sync();
}

// src/g2o_optimization.cc:129-135 (pass 0) / :150-154 (pass 1).  e->chi2() reads the error cached by
// the last computeActiveErrors() (state `last_eval`, which is the rejected trial when the final
// trial was rejected); isDepthPositive() re-maps with the CURRENT estimate.  Level-1 edges keep the
// classification error of the first pass.  Returns this thread's count of newly excluded edges.
template <class Scope>
__device__ double window_classify(const Scope& sc, const BAWin& W, double chi2_thr, int pass, int cur,
                                  int last_eval, bool have_eval, double chi2_thr_s = 0.0) {
  const int gt = sc.blk() * blockDim.x + threadIdx.x, gstride = sc.nblk() * blockDim.x;
  const double K[4] = {W.intr[0], W.intr[1], W.intr[2], W.intr[3]};
  double n_l1 = 0.0;
  for (int l = gt; l < W.Np; l += gstride) {
    for (int o = W.pt_start[l]; o < W.pt_start[l + 1]; o++) {
      const int c = W.ocam[o];
      double pc[3], e0, e1;
      const double2 uv = *reinterpret_cast<const double2*>(W.uv + (size_t)o * 2);
      const int lev = W.level[o];
      if (pass == 1 && lev == 1) continue;  // cached chi2 > thr: inlier[] already 0 from pass 0
      map_point(W.camRt[last_eval] + (size_t)c * 12, W.pts[last_eval] + (size_t)l * 3, pc);
      double e2 = 0.0, thr = chi2_thr;
      bool stereo = false;
      const double* Ke = W.okind ? edge_model(W, o, stereo) : K;  // stereo-capable windows: the edge's camera model
      if (stereo) {  // stereo edge: third row, threshold cfg.stereo_point (:137-143, :156-160)
        double pz[3];
        e2 = edge_error(pc, uv.x, uv.y, Ke, e0, e1, pz);
        const double er = edge_error_right(pz, W.ur[o], Ke, Ke[4]);
        e2 = have_eval ? e2 + er * er : 0.0;
        thr = chi2_thr_s;
      } else {
        e2 = have_eval ? edge_error(pc, uv.x, uv.y, Ke, e0, e1) : 0.0;
      }
      map_point(W.camRt[cur] + (size_t)c * 12, W.pts[cur] + (size_t)l * 3, pc);
      const bool depth_pos = pc[2] > 0.0;
      if (pass == 0) {
        // level 1: chi2 test failed; level 2: only the depth test failed (its cached chi2 stays
        // <= thr, so the final flag depends on the depth at the final estimate)
        const int nl = (e2 > thr) ? 1 : (!depth_pos ? 2 : 0);
        W.level[o] = (uint8_t)nl;
        W.inlier[o] = 0;
        n_l1 += nl ? 1.0 : 0.0;
      } else if (lev == 2) {
        W.inlier[o] = depth_pos ? 1 : 0;
      } else {
        W.inlier[o] = (e2 <= thr && depth_pos) ? 1 : 0;
      }
    }
  }
  return n_l1;
}

// write back T_wc = estimate().inverse() and the points (:164-176)
template <class Scope>
__device__ void window_finish(const Scope& sc, const BAWin& W, int cur) {
  const int gt = sc.blk() * blockDim.x + threadIdx.x, gstride = sc.nblk() * blockDim.x;
  for (int c = gt; c < W.Nc; c += gstride) {
    const double* in = W.cam[cur] + (size_t)c * 7;
    double qi[4], ti[3];
    se3_inverse(in, in + 4, qi, ti);
    double* o = W.pose_out + (size_t)c * 7;
    o[0] = qi[0]; o[1] = qi[1]; o[2] = qi[2]; o[3] = qi[3];
    o[4] = ti[0]; o[5] = ti[1]; o[6] = ti[2];
  }
  for (int i = gt; i < W.Np * 3; i += gstride) W.pts_out[i] = W.pts[cur][i];
}

template <int MODE, class Scope>
__device__ void solve_window(const Scope& sc, const BAWin& W, const BARun& run, WarpStage st,
                             PackStage pst, WorkArea wa, double* pcg_sm, double* red) {
  window_init(sc, W);
  if (MODE == 2 || MODE == 3) { pack_records(sc, W); sc.sync(); }
  int cur = 0, last_eval = 0, parity = 0;
  bool have_eval = false;  // has any computeActiveErrors() run? (g2o's cached _error is zero before)
  urmvo_ba_stats* stats = reinterpret_cast<urmvo_ba_stats*>(W.stats);
  const bool writer = (sc.blk() == 0 && threadIdx.x == 0);
  for (int pass = 0; pass < 2; pass++) {
    const bool robust = (pass == 0);
    double chi_init = 0.0;
    LMResult r = lm_optimize<MODE>(sc, W, run, pass == 0 ? run.it0 : run.it1, robust, cur, last_eval, st,
                                   pst, wa, pcg_sm, parity, red, &chi_init);
    if (r.iters > 0) have_eval = true;
    if (writer && stats) {
      stats->iters[pass] = r.iters;
      stats->trials[pass] = r.trials;
      stats->pcg_iters[pass] = r.pcg_iters;
      stats->chi2_final[pass] = r.chi;
      stats->lambda_final[pass] = r.lambda;
      if (pass == 0) stats->chi2_initial = chi_init;
    }
    const double n_l1 = window_classify(sc, W, run.chi2_thr, pass, cur, last_eval, have_eval, run.chi2_thr_s);
    if (pass == 0) {
      double s1[1] = {n_l1}, dummy[1];
      scope_reduce<1, 0>(sc, s1, dummy, W.part, parity, red);
      if (writer && stats) stats->n_level1 = (int)s1[0];
    }
    sc.sync();
    if ((MODE == 2 || MODE == 3) && pass == 0) { pack_records(sc, W); sc.sync(); }  // new edge levels
  }
  window_finish(sc, W, cur);
}

// Dynamic shared memory layout of one CTA:
//   [red: 16*32+16 doubles][work areas: nw * work_stride doubles][ints: nw * ints_per_warp]
// A warp's work area holds its staging fields and its accumulator copy (the packed modes alias the
// two); the dense PCG of the small-window modes aliases all work areas (idle during the solve).
struct SmemViews {
  double* red;
  WarpStage st;
  PackStage pst;
  WorkArea wa;
  double* pcg;
};

__device__ __forceinline__ SmemViews make_views(unsigned char* smem, int kmax, int work_stride, int ints_per_warp) {
  const int nw = blockDim.x >> 5, wid = threadIdx.x >> 5;
  SmemViews v;
  v.red = reinterpret_cast<double*>(smem);
  double* work0 = v.red + (16 * 32 + 16);
  int* ints0 = reinterpret_cast<int*>(work0 + (size_t)nw * work_stride);
  v.st.kmax = kmax;
  v.st.f = work0 + (size_t)wid * work_stride;
  v.st.cf = ints0 + (size_t)wid * ints_per_warp;
  v.pst.f = v.st.f;
  v.pst.slot = reinterpret_cast<signed char*>(v.st.cf);
  v.wa.base = work0;
  v.wa.stride = work_stride;
  v.pcg = work0;
  return v;
}

template <class Scope>
__device__ __forceinline__ void solve_window_dispatch(const Scope& sc, const BAWin& W, const BARun& run,
                                                      const SmemViews& v) {
  switch (W.acc_mode) {
    case 6: solve_window<6>(sc, W, run, v.st, v.pst, v.wa, v.pcg, v.red); break;
    case 5: solve_window<5>(sc, W, run, v.st, v.pst, v.wa, v.pcg, v.red); break;
    case 3: solve_window<3>(sc, W, run, v.st, v.pst, v.wa, v.pcg, v.red); break;
    case 2: solve_window<2>(sc, W, run, v.st, v.pst, v.wa, v.pcg, v.red); break;
    case 1: solve_window<1>(sc, W, run, v.st, v.pst, v.wa, v.pcg, v.red); break;
    default: solve_window<0>(sc, W, run, v.st, v.pst, v.wa, v.pcg, v.red); break;
  }
}

// Batched windows: one thread-block cluster per window (cluster dims set at launch).
// KMODE >= 0: every window of the batch uses accumulation mode KMODE (the usual case: one camera
// count per batch) — a kernel specialised for that mode gets its own register allocation instead of
// the maximum over all modes; KMODE < 0: per-window dispatch.
template <int KMODE>
__global__ void __launch_bounds__(256, 1)
ba_window_cluster_kernel(const BAWin* __restrict__ wins, BARun run, int kmax_all, int work_stride,
                         int ints_per_warp) {
  extern __shared__ __align__(16) unsigned char smem[];
  ClusterScope sc;
  const SmemViews v = make_views(smem, kmax_all, work_stride, ints_per_warp);
  const int n_clusters = gridDim.x / sc.nblk();
  const int cid = blockIdx.x / sc.nblk();
  for (int w = cid; w < run.n_win; w += n_clusters) {
    if (KMODE >= 0) solve_window<(KMODE >= 0 ? KMODE : 0)>(sc, wins[w], run, v.st, v.pst, v.wa, v.pcg, v.red);
    else solve_window_dispatch(sc, wins[w], run, v);
    sc.sync();
  }
}

// One large problem on the whole (cooperative) grid.  GMODE 0: mono edges, 6: stereo-capable edges.
template <int GMODE>
__global__ void __launch_bounds__(256, 1)
ba_window_grid_kernel(const BAWin* __restrict__ wins, BARun run, int kmax_all, int work_stride,
                      int ints_per_warp) {
  extern __shared__ __align__(16) unsigned char smem[];
  GridScope sc;
  const SmemViews v = make_views(smem, kmax_all, work_stride, ints_per_warp);
  for (int w = 0; w < run.n_win; w++) {
    solve_window<GMODE>(sc, wins[w], run, v.st, v.pst, v.wa, v.pcg, v.red);
    sc.sync();
  }
}

size_t ba_smem_bytes(int threads, int work_stride, int ints_per_warp) {
  const int nw = threads / 32;
  return (16 * 32 + 16) * sizeof(double) + (size_t)nw * work_stride * sizeof(double) +
         (size_t)nw * ints_per_warp * sizeof(int);
}

int ba_stage_doubles(int kmax, int stereo) { return (stereo ? kStageFieldsS : kStageFields) * kmax; }
int ba_tile_doubles() { return 32 * 37; }
// grid kernels: at least 23 KB per warp so that a ~60-neighbour block row fits the PCG cache
static __host__ __device__ int grid_work_stride(int kmax, int stereo = 0) {
  const int n = (stereo ? kStageFieldsS : kStageFields) * kmax + 32 * 37;
  return n > 2944 ? n : 2944;
}
int ba_pack_doubles() { return kPackFields * kPackSlots + 128; }

// ------------------------------------------------------------------------------- point-sharded BA
//
// One large problem sharded by POINT over the ranks of one node: every rank holds
// all cameras and a contiguous range of points with their observations.  The persistent kernel is
// cut at the two places where ranks must agree — the reduced camera system [S | b_s | b_p | chi2]
// and the trial cost [chi2', scale] — and NCCL sums those buffers over NVLink between the phase
// kernels (csrc/capi.cu drives the loop; every rank takes identical decisions from identical
// reduced values, so no decision is ever broadcast).  Each phase is the same device code as the
// persistent kernel, on the whole cooperative grid.
struct ShardState {
  double lambda, ni, currentChi, rho, chi_initial;
  int cur, last_eval, have_eval, parity;
  int robust, it, qmax, ok2;
  int iters, trials, pcg_iters, n_level1;
  // decisions for the host loop
  int cont_trials, terminate;
};

// scal layout inside the reduce buffer: [0] chi2 (LIN) [1] max diag(Hll) [2] trial chi2 [3] scale
__global__ void __launch_bounds__(256, 1)
k_sh_init(const BAWin* __restrict__ wins, ShardState* stt) {
  GridScope sc;
  window_init(sc, wins[0]);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    ShardState z = {};
    z.ni = 2.0;
    *stt = z;
  }
}

__global__ void __launch_bounds__(256, 1)
k_sh_begin_pass(ShardState* stt, int robust) {
  stt->robust = robust; stt->it = 0; stt->iters = 0; stt->trials = 0; stt->pcg_iters = 0;
  stt->cont_trials = 0; stt->terminate = 0; stt->qmax = 0;
}

__global__ void __launch_bounds__(256, 1)
k_sh_lin(const BAWin* __restrict__ wins, BARun run, ShardState* stt, double* scal, int kmax, int diag) {
  extern __shared__ __align__(16) unsigned char smem[];
  GridScope sc;
  const BAWin& W = wins[0];
  const SmemViews v = make_views(smem, kmax, grid_work_stride(kmax), kmax);
  int parity = stt->parity;
  zero_system(sc, W, diag != 0, 0.0);
  sc.sync();
  double chi = 0.0, mx = 0.0;
  if (diag) lin_phase<true, false>(sc, W, stt->cur, 0.0, stt->robust != 0, run.delta, v.st, v.wa, chi, mx);
  else lin_phase<false, false>(sc, W, stt->cur, stt->lambda, stt->robust != 0, run.delta, v.st, v.wa, chi, mx);
  double s1[1] = {chi}, m1[1] = {mx};
  scope_reduce<1, 1>(sc, s1, m1, W.part, parity, v.red);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    scal[0] = s1[0];
    scal[1] = m1[0];
    stt->parity = parity;
  }
}

// after the DIAG all-reduce: lambda = tau * max |H_jj| (computeLambdaInit)
__global__ void k_sh_lambda(const BAWin* __restrict__ wins, ShardState* stt, const double* scal) {
  const BAWin& W = wins[0];
  __shared__ double sm[256];
  double m = 0.0;
  for (int i = threadIdx.x; i < W.Ncf * 6; i += blockDim.x) m = fmax(m, fabs(W.hdiag[i]));
  sm[threadIdx.x] = m;
  __syncthreads();
  for (int off = blockDim.x >> 1; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + off]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    stt->lambda = 1e-5 * fmax(sm[0], scal[1]);
    stt->ni = 2.0;
    if (stt->robust) stt->chi_initial = scal[0];
  }
}

// after the [S | b_s | b_p | chi2] all-reduce: damp, PCG, trial cameras, back-substitution, trial cost
__global__ void __launch_bounds__(256, 1)
k_sh_solve(const BAWin* __restrict__ wins, BARun run, ShardState* stt, double* scal, int kmax, int rank) {
  extern __shared__ __align__(16) unsigned char smem[];
  GridScope sc;
  const BAWin& W = wins[0];
  const SmemViews v = make_views(smem, kmax, grid_work_stride(kmax), kmax);
  int parity = stt->parity;
  const int cur = stt->cur;
  const double lambda = stt->lambda;
  damp_diagonal(sc, W, lambda);
  sc.sync();
  int pcg_it = 0;
  const bool ok2 = pcg_phase(sc, W, run.pcg_tol, run.pcg_max_iter, W.part, parity, v.red, pcg_it,
                             v.wa.base + (size_t)(threadIdx.x >> 5) * v.wa.stride, v.wa.stride);
  sc.sync();
  double tchi = 0.0, sc_l = 0.0;
  if (ok2) {
    cam_update(sc, W, cur);
    sc.sync();
    backsub_phase(sc, W, cur, lambda, stt->robust != 0, run.delta, tchi, sc_l);
    if (rank == 0) {  // the pose part of computeScale is counted once
      const int gt = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
      for (int i = gt; i < W.Ncf * 6; i += gstride) {
        const double x = __ldcg(W.xp + i);
        sc_l += x * (lambda * x + __ldcg(W.bp + i));
      }
    }
  }
  double s2[2] = {tchi, sc_l}, dummy[1];
  scope_reduce<2, 0>(sc, s2, dummy, W.part, parity, v.red);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    scal[2] = s2[0];
    scal[3] = s2[1];
    stt->ok2 = ok2 ? 1 : 0;
    stt->pcg_iters += pcg_it;
    stt->parity = parity;
  }
}

// after the [chi2', scale] all-reduce: gain ratio, damping update, accept / reject (one thread)
__global__ void k_sh_decide(ShardState* stt, const double* scal, ShardState* host_copy) {
  ShardState s = *stt;
  const double currentChi = scal[0];
  const double tempChi = s.ok2 ? scal[2] : 1.7976931348623157e308;
  const double scale = s.ok2 ? scal[3] : 0.0;
  if (s.ok2) { s.last_eval = s.cur ^ 1; s.have_eval = 1; }
  const double rho = (currentChi - tempChi) / (scale + 1e-3);
  bool lambda_bad = false;
  s.currentChi = currentChi;
  if (rho > 0 && isfinite(tempChi)) {
    double alpha = 1. - pow((2 * rho - 1), 3);
    alpha = fmin(alpha, 2. / 3.);
    s.lambda *= fmax(1. / 3., alpha);
    s.ni = 2;
    s.currentChi = tempChi;
    s.cur ^= 1;
  } else {
    s.lambda *= s.ni;
    s.ni *= 2;
    if (!isfinite(s.lambda)) lambda_bad = true;
  }
  s.qmax++;
  s.trials++;
  s.rho = rho;
  s.cont_trials = (!lambda_bad && rho < 0 && s.qmax < 10) ? 1 : 0;
  if (!s.cont_trials) {
    s.iters = s.it + 1;
    s.terminate = (s.qmax == 10 || rho == 0 || lambda_bad || !isfinite(s.lambda)) ? 1 : 0;
    s.it++;
    s.qmax = 0;
  }
  *stt = s;
  *host_copy = s;
}

__global__ void __launch_bounds__(256, 1)
k_sh_classify(const BAWin* __restrict__ wins, BARun run, ShardState* stt, int pass) {
  __shared__ double red[16 * 32 + 16];
  GridScope sc;
  const BAWin& W = wins[0];
  int parity = stt->parity;
  const double n_l1 = window_classify(sc, W, run.chi2_thr, pass, stt->cur, stt->last_eval, stt->have_eval != 0);
  double s1[1] = {n_l1}, dummy[1];
  scope_reduce<1, 0>(sc, s1, dummy, W.part, parity, red);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (pass == 0) stt->n_level1 = (int)s1[0];
    stt->parity = parity;
    urmvo_ba_stats* st = reinterpret_cast<urmvo_ba_stats*>(W.stats);
    st->iters[pass] = stt->iters;
    st->trials[pass] = stt->trials;
    st->pcg_iters[pass] = stt->pcg_iters;
    st->chi2_final[pass] = stt->currentChi;
    st->lambda_final[pass] = stt->lambda;
    if (pass == 0) { st->chi2_initial = stt->chi_initial; st->n_level1 = (int)s1[0]; }
  }
}

__global__ void __launch_bounds__(256, 1)
k_sh_finish(const BAWin* __restrict__ wins, ShardState* stt) {
  GridScope sc;
  window_finish(sc, wins[0], stt->cur);
}

size_t shard_state_bytes() { return sizeof(ShardState); }

static cudaError_t coop(const void* k, int grid, int threads, void** args, size_t smem, cudaStream_t s) {
  return cudaLaunchCooperativeKernel(k, dim3((unsigned)grid), dim3((unsigned)threads), args, smem, s);
}

int shard_grid_capacity(int threads, int kmax) {
  const size_t smem = ba_smem_bytes(threads, grid_work_stride(kmax), kmax);
  int dev = 0, sms = 0, best = 1 << 30;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const void* ks[2] = {(const void*)k_sh_lin, (const void*)k_sh_solve};
  for (const void* k : ks) {
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, threads, smem) != cudaSuccess) return 0;
    best = per_sm * sms < best ? per_sm * sms : best;
  }
  return best;
}

cudaError_t launch_sh_init(const BAWin* w, void* stt, int grid, int threads, cudaStream_t s) {
  void* args[] = {(void*)&w, (void*)&stt};
  return coop((const void*)k_sh_init, grid, threads, args, 0, s);
}
cudaError_t launch_sh_begin_pass(void* stt, int robust, cudaStream_t s) {
  k_sh_begin_pass<<<1, 1, 0, s>>>((ShardState*)stt, robust);
  return cudaGetLastError();
}
cudaError_t launch_sh_lin(const BAWin* w, const BARun& run, void* stt, double* scal, int kmax, int diag,
                          int grid, int threads, cudaStream_t s) {
  BARun r = run;
  void* args[] = {(void*)&w, (void*)&r, (void*)&stt, (void*)&scal, (void*)&kmax, (void*)&diag};
  return coop((const void*)k_sh_lin, grid, threads, args, ba_smem_bytes(threads, grid_work_stride(kmax), kmax), s);
}
cudaError_t launch_sh_lambda(const BAWin* w, void* stt, const double* scal, cudaStream_t s) {
  k_sh_lambda<<<1, 256, 0, s>>>(w, (ShardState*)stt, scal);
  return cudaGetLastError();
}
cudaError_t launch_sh_solve(const BAWin* w, const BARun& run, void* stt, double* scal, int kmax, int rank,
                            int grid, int threads, cudaStream_t s) {
  BARun r = run;
  void* args[] = {(void*)&w, (void*)&r, (void*)&stt, (void*)&scal, (void*)&kmax, (void*)&rank};
  return coop((const void*)k_sh_solve, grid, threads, args, ba_smem_bytes(threads, grid_work_stride(kmax), kmax), s);
}
cudaError_t launch_sh_decide(void* stt, const double* scal, void* host_copy, cudaStream_t s) {
  k_sh_decide<<<1, 1, 0, s>>>((ShardState*)stt, scal, (ShardState*)host_copy);
  return cudaGetLastError();
}
cudaError_t launch_sh_classify(const BAWin* w, const BARun& run, void* stt, int pass, int grid, int threads,
                               cudaStream_t s) {
  BARun r = run;
  void* args[] = {(void*)&w, (void*)&r, (void*)&stt, (void*)&pass};
  return coop((const void*)k_sh_classify, grid, threads, args, 0, s);
}
cudaError_t launch_sh_finish(const BAWin* w, void* stt, int grid, int threads, cudaStream_t s) {
  void* args[] = {(void*)&w, (void*)&stt};
  return coop((const void*)k_sh_finish, grid, threads, args, 0, s);
}
// the host reads these two fields of the mirrored ShardState after every trial
void shard_flags(const void* host_copy, int* cont_trials, int* terminate) {
  const ShardState* s = (const ShardState*)host_copy;
  *cont_trials = s->cont_trials;
  *terminate = s->terminate;
}

template <int KMODE>
static cudaError_t launch_ba_cluster_t(const BAWin* wins_dev, const BARun& run, int kmax, int work_stride,
                                       int ints_per_warp, int n_clusters, int cluster_size, int threads,
                                       cudaStream_t stream) {
  const size_t smem = ba_smem_bytes(threads, work_stride, ints_per_warp);
  cudaError_t e = cudaFuncSetAttribute(ba_window_cluster_kernel<KMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (cluster_size > 8) {
    e = cudaFuncSetAttribute(ba_window_cluster_kernel<KMODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n_clusters * cluster_size));
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster_size;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, ba_window_cluster_kernel<KMODE>, wins_dev, run, kmax, work_stride, ints_per_warp);
}

// mode: the common accumulation mode of the batch (0..3) or -1 for a mixed batch
cudaError_t launch_ba_cluster(const BAWin* wins_dev, const BARun& run, int mode, int kmax, int work_stride,
                              int ints_per_warp, int n_clusters, int cluster_size, int threads,
                              cudaStream_t stream) {
  switch (mode) {
    case 6: return launch_ba_cluster_t<6>(wins_dev, run, kmax, work_stride, ints_per_warp, n_clusters, cluster_size, threads, stream);
    case 5: return launch_ba_cluster_t<5>(wins_dev, run, kmax, work_stride, ints_per_warp, n_clusters, cluster_size, threads, stream);
    case 3: return launch_ba_cluster_t<3>(wins_dev, run, kmax, work_stride, ints_per_warp, n_clusters, cluster_size, threads, stream);
    case 2: return launch_ba_cluster_t<2>(wins_dev, run, kmax, work_stride, ints_per_warp, n_clusters, cluster_size, threads, stream);
    case 1: return launch_ba_cluster_t<1>(wins_dev, run, kmax, work_stride, ints_per_warp, n_clusters, cluster_size, threads, stream);
    case 0: return launch_ba_cluster_t<0>(wins_dev, run, kmax, work_stride, ints_per_warp, n_clusters, cluster_size, threads, stream);
    default: return launch_ba_cluster_t<-1>(wins_dev, run, kmax, work_stride, ints_per_warp, n_clusters, cluster_size, threads, stream);
  }
}

int ba_grid_capacity(int threads, int kmax, int stereo) {
  const size_t smem = ba_smem_bytes(threads, grid_work_stride(kmax, stereo), kmax);
  const void* kern = stereo ? (const void*)ba_window_grid_kernel<6> : (const void*)ba_window_grid_kernel<0>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
  int per_sm = 0, dev = 0, sms = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess) return 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return per_sm * sms;
}

cudaError_t launch_ba_grid(const BAWin* wins_dev, const BARun& run, int kmax, int grid_blocks,
                           int threads, cudaStream_t stream, int stereo) {
  const int ws = grid_work_stride(kmax, stereo);
  const size_t smem = ba_smem_bytes(threads, ws, kmax);
  const void* kern = stereo ? (const void*)ba_window_grid_kernel<6> : (const void*)ba_window_grid_kernel<0>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  BARun r = run;
  int km = kmax, w2 = ws, ip = kmax;
  void* args[] = {(void*)&wins_dev, (void*)&r, (void*)&km, (void*)&w2, (void*)&ip};
  return cudaLaunchCooperativeKernel(kern, dim3((unsigned)grid_blocks), dim3((unsigned)threads), args, smem, stream);
}

cudaError_t ba_timing_read(unsigned long long* out, bool reset) {
  cudaError_t e = cudaMemcpyFromSymbol(out, g_ba_timing, sizeof(unsigned long long) * 8);
  if (e == cudaSuccess && reset) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    e = cudaMemcpyToSymbol(g_ba_timing, z, sizeof(z));
  }
  return e;
}

}  // namespace urmvo
