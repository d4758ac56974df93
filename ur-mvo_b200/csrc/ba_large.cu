// csrc/ba_large.cu — ONE large bundle adjustment (BASELINE.json configs[3] / configs[4]: 50 keyframes /
// 400k observations, 1000 cameras / 2M observations; also every rank of the point-sharded solve) as a
// sequence of phase kernels on the context stream ("tile mode", BAWin::acc_mode 4).
//
// Reference behaviour: the same LocalmapOptimization call as csrc/ba_kernels.cu
//   /root/reference/src/g2o_optimization.cc:20-177 (graph, optimize(10), re-classification, optimize(5))
//   g2o BlockSolver Schur complement + LinearSolverEigen (sparse Cholesky, :27-35) + Levenberg
//
// Why a second formulation: the persistent grid kernel of round 1 accumulated the Schur
// complement of a large window with ~36 global fp64 atomics per camera pair of every point (65 M
// atomics per pass at configs[3]) and solved the reduced camera system with 470 block-Jacobi PCG
// iterations per trial.  Here
//   * the host renumbers the points by their first free camera and cuts them into CHUNKS whose
//     reduced-system blocks fit one CTA: every thread owns ONE 6x6 block of S in registers for the whole
//     chunk (S-stationary, like the packed small-window modes), the warps stage groups of <= 32
//     observations (lane = observation, coalesced 32-byte records) into shared memory, and every warp
//     sweeps the staged points for the camera pairs it owns.  A chunk flushes each block once:
//     ~2 M atomics per pass instead of 400 M at configs[4];
//   * S is stored as a block band (cameras along a trajectory) and solved DIRECTLY by a block-banded
//     Cholesky factorisation in one CTA — what g2o's LinearSolverEigen does, so parity is exact — with
//     the trailing window of the factorisation held in registers (one 6x6 block per thread);
//   * chi2 / scale partial sums are combined by the last CTA of a kernel in CTA order (deterministic),
//     so no cooperative launch and no grid barrier is needed anywhere.
// The host enqueues whole LM iterations without synchronising; kernels of trials that the device-side
// LM state has already finished return immediately (LgState::active).

#include "ba_device.cuh"
#include "ba_large_tail.cuh"
#include "kernels.h"
#include "../../include/urmvo_b200.h"

namespace urmvo {


// development aid (urmvo_debug_lg_timing): SM cycles of thread 0 of the band solve per segment
// 0 diagonal factorisation (to barrier 1)  1 panel (to barrier 2)  2 trailing update + row load
// 3 back substitution  4 whole kernel  5 tail (x_p, scale, cameras)  6 steps
__device__ unsigned long long g_lg_timing[8];

constexpr int kLgThreads = 256;
constexpr int kLgWarps = kLgThreads / 32;
constexpr int kLgStage = kPackFields * kPackSlots;  // doubles per warp
constexpr int kLgSlotBytes = 32 * 32;               // (point in group, camera in chunk window) -> slot

// scal: [0] chi2 at the linearisation point  [1] unused  [2] trial chi2  [3] trial scale (points)
//       [8 + rank] max diag(Hll) of the rank (lambda initialisation; summed over ranks with zeros elsewhere)

// ------------------------------------------------------------------------------- small helpers

// CTA-order sum of `nv` per-CTA partials written to cpart[blockIdx.x * 4 + k]; the LAST CTA to arrive
// adds them in CTA order and writes out[k] (deterministic without a grid barrier).
__device__ __forceinline__ void finish_partials(double* cpart, unsigned int* ticket, int nv, double* out0,
                                                double* out1, double* out2, bool max1) {
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned v = atomicAdd(ticket, 1u);
    s_last = (v == gridDim.x - 1);
    if (s_last) *ticket = 0u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    double a0 = 0.0, a1 = max1 ? -1.0e300 : 0.0, a2 = 0.0;
    for (int b = lane; b < (int)gridDim.x; b += 32) {
      a0 += __ldcg(cpart + (size_t)b * 4);
      if (nv > 1) { const double x = __ldcg(cpart + (size_t)b * 4 + 1); a1 = max1 ? fmax(a1, x) : a1 + x; }
      if (nv > 2) a2 += __ldcg(cpart + (size_t)b * 4 + 2);
    }
    a0 = warp_sum(a0);
    a1 = max1 ? warp_max(a1) : warp_sum(a1);
    a2 = warp_sum(a2);
    if (lane == 0) {
      *out0 = a0;
      if (nv > 1 && out1) *out1 = a1;
      if (nv > 2 && out2) *out2 = a2;
    }
  }
}

// ------------------------------------------------------------------------------- init / records

__global__ void __launch_bounds__(kLgThreads)
k_lg_init(const BAWin* __restrict__ wins, LgState* stt) {
  const BAWin& W = wins[0];
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
  for (int c = gt; c < W.Nc; c += gstride) {  // setEstimate(SE3Quat(q, p).inverse()), g2o_optimization.cc:45
    double q[4], t[3], qi[4], ti[3];
    const double* in = W.pose_in + (size_t)c * 7;
    q[0] = in[0]; q[1] = in[1]; q[2] = in[2]; q[3] = in[3];
    t[0] = in[4]; t[1] = in[5]; t[2] = in[6];
    quat_normalize_w(q);
    se3_inverse(q, t, qi, ti);
    double* o = W.cam[0] + (size_t)c * 7;
    o[0] = qi[0]; o[1] = qi[1]; o[2] = qi[2]; o[3] = qi[3];
    o[4] = ti[0]; o[5] = ti[1]; o[6] = ti[2];
    double R[9];
    quat_to_R(qi, R);
    double* rt = W.camRt[0] + (size_t)c * 12;
#pragma unroll
    for (int a = 0; a < 9; a++) rt[a] = R[a];
    rt[9] = ti[0]; rt[10] = ti[1]; rt[11] = ti[2];
  }
  for (int i = gt; i < W.Np * 3; i += gstride) {
    const double v = W.pts_in[i];
    W.pts[0][i] = v;
    W.pts[1][i] = v;  // points without observations are never rewritten by BACKSUB
  }
  for (int o = gt; o < W.No; o += gstride) W.level[o] = 0;
  if (gt == 0) {
    LgState z = {};
    z.ni = 2.0;
    *stt = z;
    for (int k = 0; k < 4; k++) W.ticket[k] = 0u;
  }
}

__global__ void k_lg_begin_pass(LgState* stt, int robust, int n_iter) {
  stt->robust = robust; stt->it = 0; stt->n_iter = n_iter; stt->iters = 0; stt->trials = 0;
  stt->pcg_iters = 0; stt->qmax = 0; stt->active = n_iter > 0 ? 1 : 0;
}

// One 32-byte record per (group, lane); cf is the free-camera index RELATIVE to the chunk window.
__global__ void __launch_bounds__(kLgThreads)
k_lg_pack(const BAWin* __restrict__ wins) {
  const BAWin& W = wins[0];
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kLgWarps + (threadIdx.x >> 5), gstride = gridDim.x * kLgWarps;
  for (int g = gw; g < W.n_grp; g += gstride) {
    const int p0 = W.grp_pt[g];
    const int np = W.grp_pt[g + 1] - p0;
    const int o0 = W.pt_start[p0];
    const int nobs = W.pt_start[p0 + np] - o0;
    const int cbase = W.grp_cbase[g];
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
      const int cf = W.cam_free[r.c];
      r.cf = (signed char)(cf >= 0 ? cf - cbase : -1);
      r.lev = W.level[o];
      r.valid = 1;
    }
    W.rec[(size_t)g * 32 + lane] = r;
  }
}

// ------------------------------------------------------------------------------- phase LIN (tile mode)

struct RecView {
  double u, v;
  int pl, c, s0, s1, pi, cf, np, lev;
  bool valid;
};
__device__ __forceinline__ RecView load_rec(const ObsRec* rec, int g, int lane) {
  const int4* p = reinterpret_cast<const int4*>(rec + (size_t)g * 32 + lane);
  const int4 a = __ldg(p), b = __ldg(p + 1);
  RecView r;
  r.u = __hiloint2double(a.y, a.x);
  r.v = __hiloint2double(a.w, a.z);
  r.pl = b.x; r.c = b.y;
  r.s0 = b.z & 255; r.s1 = (b.z >> 8) & 255; r.pi = (b.z >> 16) & 255;
  r.cf = b.z >> 24;  // arithmetic shift keeps the sign of the int8
  r.np = b.w & 255; r.lev = (b.w >> 8) & 255;
  r.valid = (b.w >> 16) & 1;
  return r;
}

// DIAG: diag(Hpp) -> hdiag, max diag(Hll), robust chi2 (computeLambdaInit at iteration 0).
template <bool DIAG>
__global__ void __launch_bounds__(kLgThreads, 1)
k_lg_lin(const BAWin* __restrict__ wins, BARun run, const LgState* __restrict__ stt, double* scal, int rank) {
  extern __shared__ __align__(16) unsigned char smem[];
  if (!stt->active) return;
  const BAWin& W = wins[0];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  double* red = reinterpret_cast<double*>(smem);                     // 4 * 32 doubles
  double* stage0 = red + 4 * 32;
  // TWO staging buffers per warp: while the slower warps of the CTA still sweep round r (buffer r & 1), a warp that is
  // done stages its group of round r + 1 into the other buffer — one CTA barrier per round instead of two, and the
  // latency-bound staging chain of some warps overlaps the fp64-dense sweep of the others
  unsigned char* slot0 = reinterpret_cast<unsigned char*>(stage0 + (size_t)2 * kLgWarps * kLgStage);
  int* s_np = reinterpret_cast<int*>(slot0 + (size_t)2 * kLgWarps * kLgSlotBytes);  // [2][kLgWarps]
  auto stage_of = [&](int buf) {
    PackStage st;
    st.f = stage0 + (size_t)(buf * kLgWarps + wid) * kLgStage;
    st.slot = reinterpret_cast<signed char*>(slot0 + (size_t)(buf * kLgWarps + wid) * kLgSlotBytes);
    return st;
  };
  const int cur = stt->cur;
  const bool robust = stt->robust != 0;
  const double lambda = DIAG ? 0.0 : stt->lambda;
  const double delta = run.delta;
  const double* __restrict__ camRt = W.camRt[cur];
  const double* __restrict__ pts = W.pts[cur];
  const double K[4] = {W.intr[0], W.intr[1], W.intr[2], W.intr[3]};
  const ObsRec* rec = W.rec;
  double* __restrict__ Dinv_out = W.Dinv;
  double* __restrict__ bl_out = W.bl;
  double chi_acc = 0.0, maxdiag_acc = 0.0;
  // the all-zero slot of every stage (absent pairs read it: they add exactly +0)
  for (int a = lane; a < kPackFields; a += 32) {
    stage_of(0).f[a * kPackSlots + kPackZero] = 0.0;
    stage_of(1).f[a * kPackSlots + kPackZero] = 0.0;
  }
  __syncthreads();
  int rp = 0;  // buffer of the round that is swept next (runs on across the chunks)

  // ---- stage: warp w takes group gb + w of the round (lane = observation) into its buffer `buf`
  auto stage_round = [&](int gb, int g_end, int buf) {
    PackStage st = stage_of(buf);
      const int g = gb + wid;
      int np_mine = 0;
      if (g < g_end) {
        const RecView r = load_rec(rec, g, lane);
        np_mine = r.np;
#pragma unroll
        for (int q = 0; q < kLgSlotBytes / 128; q++) reinterpret_cast<int*>(st.slot)[lane + 32 * q] = 0x20202020;
        __syncwarp();
        const double X[3] = {pts[(size_t)r.pl * 3], pts[(size_t)r.pl * 3 + 1], pts[(size_t)r.pl * 3 + 2]};
        double hc[6] = {0, 0, 0, 0, 0, 0}, blc[3] = {0, 0, 0};
        int cf = -1;
        double B[6];
        if (r.valid && !r.lev) {
          const double* Rt = camRt + (size_t)r.c * 12;
          double pc[3], pz[3], e0, e1, w;
          map_point(Rt, X, pc);
          const double e2 = edge_error(pc, r.u, r.v, K, e0, e1, pz);
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
          cf = r.cf;
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
            st.slot[r.pi * 32 + cf] = (signed char)lane;
          }
        }
#pragma unroll
        for (int a = 0; a < 6; a++) st.h(a, lane) = hc[a];
        if (!DIAG) {
#pragma unroll
          for (int a = 0; a < 3; a++) st.bl(a, lane) = blc[a];
        }
        __syncwarp();
        // per-point sums in observation order (every lane for its own point; uniform trip count)
        double h[6] = {0, 0, 0, 0, 0, 0}, bl[3] = {0, 0, 0};
        {
          const int len = r.valid ? r.s1 - r.s0 : 0;
          const int maxlen = __reduce_max_sync(0xffffffffu, len);
          for (int t = 0; t < maxlen; t++) {
            const int s = t < len ? r.s0 + t : kPackZero;
#pragma unroll
            for (int a = 0; a < 6; a++) h[a] += st.h(a, s);
            if (!DIAG) {
#pragma unroll
              for (int a = 0; a < 3; a++) bl[a] += st.bl(a, s);
            }
          }
        }
        if (DIAG) {
          if (r.valid) maxdiag_acc = fmax(maxdiag_acc, fmax(fabs(h[0]), fmax(fabs(h[3]), fabs(h[5]))));
        } else if (r.valid) {
          double Di[6];
          const double hl[6] = {h[0] + lambda, h[1], h[2], h[3] + lambda, h[4], h[5] + lambda};
          sym3_inverse(hl, Di);
          if (lane == r.s0) {
#pragma unroll
            for (int a = 0; a < 6; a++) Dinv_out[(size_t)r.pl * 6 + a] = Di[a];
#pragma unroll
            for (int a = 0; a < 3; a++) bl_out[(size_t)r.pl * 3 + a] = bl[a];
          }
          if (cf >= 0) {
            double A[6];
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
              const double x0 = B[rr * 3], x1 = B[rr * 3 + 1], x2 = B[rr * 3 + 2];
              A[rr * 3 + 0] = x0 * Di[0] + x1 * Di[1] + x2 * Di[2];
              A[rr * 3 + 1] = x0 * Di[1] + x1 * Di[3] + x2 * Di[4];
              A[rr * 3 + 2] = x0 * Di[2] + x1 * Di[4] + x2 * Di[5];
            }
#pragma unroll
            for (int a = 0; a < 6; a++) st.A(a, lane) = A[a];
            st.g(0, lane) = st.we(0, lane) + (A[0] * bl[0] + A[1] * bl[1] + A[2] * bl[2]);
            st.g(1, lane) = st.we(1, lane) + (A[3] * bl[0] + A[4] * bl[1] + A[5] * bl[2]);
          }
        }
      }
      if (lane == 0) s_np[buf * kLgWarps + wid] = np_mine;
  };

  for (int ch = blockIdx.x; ch < W.n_chunk; ch += gridDim.x) {
    const int g_begin = W.chunk_grp[ch], g_end = W.chunk_grp[ch + 1];
    const int b0 = W.chunk_blk[ch], nb = W.chunk_blk[ch + 1] - b0;
    // a chunk with <= 128 blocks is swept by several REPLICAS of the block owners (warp-aligned), each
    // replica visiting every n_rep-th staged group; all replicas flush
    const int nb_pad = (nb + 31) & ~31;
    const int n_rep = nb_pad > 0 ? (kLgThreads / nb_pad > 0 ? kLgThreads / nb_pad : 1) : 1;
    const int my_rep = nb_pad > 0 ? tid / nb_pad : 0, bidx = nb_pad > 0 ? tid - my_rep * nb_pad : tid;
    const bool has = my_rep < n_rep && bidx < nb;
    int cil = 0, cjl = 0, gblk = 0;
    if (has) {
      const int d0 = W.blk_desc[(size_t)(b0 + bidx) * 2];
      gblk = W.blk_desc[(size_t)(b0 + bidx) * 2 + 1];
      cil = d0 & 255; cjl = (d0 >> 8) & 255;
    }
    const bool diag_lane = has && cil == cjl;
    double accS[36];
#pragma unroll
    for (int e = 0; e < 36; e++) accS[e] = 0.0;
    double accb[12];
#pragma unroll
    for (int e = 0; e < 12; e++) accb[e] = 0.0;

    stage_round(g_begin, g_end, rp);  // first round of the chunk (its buffer was last swept two rounds ago)
    for (int gb = g_begin; gb < g_end; gb += kLgWarps) {
      __syncthreads();  // round gb is staged by every warp; everybody has left the sweep of the round before
      // ---- sweep: every thread visits the staged points for the camera pair of ITS block
      for (int ws = 0; ws < kLgWarps; ws++) {
        const int np = s_np[rp * kLgWarps + ws];
        if (np == 0 || ws % n_rep != my_rep % n_rep) continue;
        PackStage sv;
        sv.f = stage0 + (size_t)(rp * kLgWarps + ws) * kLgStage;
        const unsigned char* sl = slot0 + (size_t)(rp * kLgWarps + ws) * kLgSlotBytes;
        for (int q = 0; q < np; q++) {
          int si = kPackZero, sj = kPackZero;
          if (has) { si = sl[q * 32 + cil]; sj = DIAG ? si : sl[q * 32 + cjl]; }
          const bool ok = has && !((si | sj) & kPackZero) && (!DIAG || diag_lane);
          if (!__any_sync(0xffffffffu, ok)) continue;
          if (!ok) continue;
          if (DIAG) {
            const double w = sv.w(si);
#pragma unroll
            for (int a = 0; a < 6; a++) {
              const double j0 = sv.Jp(a, si), j1 = sv.Jp(6 + a, si);
              accb[a] += w * (j0 * j0 + j1 * j1);
            }
            continue;
          }
          // 6x6 block J_i^T M J_j, M = delta_ij w I - A_i B_j^T (Hpp term + Schur correction)
          double M[4];
          {
            const double a0 = sv.A(0, si), a1 = sv.A(1, si), a2 = sv.A(2, si);
            const double a3 = sv.A(3, si), a4 = sv.A(4, si), a5 = sv.A(5, si);
            const double x0 = sv.B(0, sj), x1 = sv.B(1, sj), x2 = sv.B(2, sj);
            const double x3 = sv.B(3, sj), x4 = sv.B(4, sj), x5 = sv.B(5, sj);
            const double wd = si == sj ? sv.w(si) : 0.0;
            M[0] = wd - (a0 * x0 + a1 * x1 + a2 * x2);
            M[1] = -(a0 * x3 + a1 * x4 + a2 * x5);
            M[2] = -(a3 * x0 + a4 * x1 + a5 * x2);
            M[3] = wd - (a3 * x3 + a4 * x4 + a5 * x5);
          }
          double T[12];
#pragma unroll
          for (int b = 0; b < 6; b++) {
            const double j0 = sv.Jp(b, sj), j1 = sv.Jp(6 + b, sj);
            T[b] = M[0] * j0 + M[1] * j1;
            T[6 + b] = M[2] * j0 + M[3] * j1;
          }
          double g0 = 0.0, g1 = 0.0, w0 = 0.0, w1 = 0.0;
          if (diag_lane) { g0 = sv.g(0, si); g1 = sv.g(1, si); w0 = sv.we(0, si); w1 = sv.we(1, si); }
#pragma unroll
          for (int a = 0; a < 6; a++) {
            const double j0 = sv.Jp(a, si), j1 = sv.Jp(6 + a, si);
#pragma unroll
            for (int b = 0; b < 6; b++) accS[a * 6 + b] = fma(j1, T[6 + b], fma(j0, T[b], accS[a * 6 + b]));
            accb[a] = fma(-j1, g1, fma(-j0, g0, accb[a]));          // b_s (zero terms off the diagonal lanes)
            accb[6 + a] = fma(-j1, w1, fma(-j0, w0, accb[6 + a]));  // b_p
          }
        }
      }
      if (gb + kLgWarps < g_end) stage_round(gb + kLgWarps, g_end, rp ^ 1);
      rp ^= 1;
    }
    // ---- flush: each block of S once per chunk
    if (has) {
      if (DIAG) {
        if (diag_lane) {
          const int cg = W.grp_cbase[g_begin] + cil;
#pragma unroll
          for (int a = 0; a < 6; a++) atomicAdd(&W.hdiag[cg * 6 + a], accb[a]);
        }
      } else {
        double* Sb = W.S + (size_t)gblk * 36;
#pragma unroll
        for (int e = 0; e < 36; e++) atomicAdd(&Sb[e], accS[e]);
        if (diag_lane) {
          const int cg = W.grp_cbase[g_begin] + cil;
#pragma unroll
          for (int a = 0; a < 6; a++) {
            atomicAdd(&W.bs[cg * 6 + a], accb[a]);
            atomicAdd(&W.bp[cg * 6 + a], accb[6 + a]);
          }
        }
      }
    }
  }
  // ---- CTA partials of chi2 / max diag(Hll), combined in CTA order by the last CTA
  chi_acc = warp_sum(chi_acc);
  maxdiag_acc = warp_max(maxdiag_acc);
  if (lane == 0) { red[wid] = chi_acc; red[32 + wid] = maxdiag_acc; }
  __syncthreads();
  if (tid == 0) {
    double s = 0.0, m = 0.0;
    for (int w = 0; w < kLgWarps; w++) { s += red[w]; m = fmax(m, red[32 + w]); }
    W.cpart[(size_t)blockIdx.x * 4] = s;
    W.cpart[(size_t)blockIdx.x * 4 + 1] = m;
  }
  finish_partials(W.cpart, W.ticket + 0, DIAG ? 2 : 1, scal, DIAG ? scal + 8 + rank : nullptr, nullptr, true);
}

// after the DIAG all-reduce: lambda = tau * max |H_jj| over poses AND points (computeLambdaInit)
__global__ void k_lg_lambda(const BAWin* __restrict__ wins, LgState* stt, const double* scal, int world) {
  if (!stt->active) return;
  const BAWin& W = wins[0];
  __shared__ double sm[256];
  double m = 0.0;
  for (int i = threadIdx.x; i < W.Ncf * 6; i += blockDim.x) m = fmax(m, fabs(W.hdiag[i]));
  for (int i = threadIdx.x; i < world; i += blockDim.x) m = fmax(m, scal[8 + i]);
  sm[threadIdx.x] = m;
  __syncthreads();
  for (int off = blockDim.x >> 1; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + off]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    stt->lambda = 1e-5 * sm[0];
    stt->ni = 2.0;
    if (stt->robust) stt->chi_initial = scal[0];
  }
}

// ------------------------------------------------------------------------------- direct band solve
//
// (S + lambda I) x = b_s by block-banded Cholesky S = L L^T, one CTA, M = bw + 1 <= 22, M*M threads.
// Thread (r, c) of the M x M ring owns, in REGISTERS, the block (i, j) of the trailing window with
// i = r (mod M), j = c (mod M): its diagonal offset i - j = (r - c) mod M never changes, only the row
// advances by M each time a block has been eliminated.  Step k:
//   (a) the owner of (k, k) factorises it (6x6 Cholesky, reciprocal diagonal) and forward-substitutes
//       y_k; (b) the owners of (k + d, k) turn their blocks into L_{k+d,k} = A L_kk^-T, publish them in
//       shared memory, store them to the band factor in global memory and update the right-hand side;
//   (c) the owners of the trailing blocks (i, j), k < j <= i < k + M, subtract L_ik L_jk^T;
//   (d) the threads whose block was eliminated M - off steps ago load the block of the row that enters
//       the window (fetched with cp.async one generation ahead into the thread's private cell).
// Two CTA barriers per step.  Back substitution streams the band factor back through a cp.async ring.
// A non-positive pivot fails the solve (g2o: Cholesky failure => ok2 = false).

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  const size_t src = __cvta_generic_to_global(gsrc);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kBandHalf = 8;      // columns of the factor per half of the back-substitution ring
constexpr int kBandMaxM = 17;     // bw + 1 <= 17: M*M <= 289 ring threads with one register-resident block each
constexpr int kBandThreads = 352; // 10 warps of ring threads + one warp that factorises the diagonal blocks

__host__ __device__ inline int band_cells(int M) { return (M * M > 2 * kBandHalf * M ? M * M : 2 * kBandHalf * M); }
size_t band_smem_bytes(int M, int Ncf) {
  return ((size_t)band_cells(M) * 36 + (size_t)2 * M * 37 + 48 + 2 * 36 + 40 + (size_t)Ncf * 6 + 64) * sizeof(double);
}

// S is stored with a uniform row stride in tile mode: block (i, i + d) at (i * M + d) * 36.
//
// Per step k (two CTA barriers):
//   (b) 6 (M - 1) "row threads" turn the published blocks of column k into L_ik = A_ik L_kk^-T, one row
//       each (the rows of a block are independent), update the right-hand side, store the factor;
//   (c) ring threads subtract L_ik L_jk^T from their register-resident trailing blocks; the owners
//       of column k + 1 and of the diagonal block k + 2 publish theirs to shared memory; MEANWHILE the
//       diagonal warp applies the last update to diagonal block k + 1, factorises it (6x6 Cholesky,
//       reciprocal diagonal via rsqrt) and forward-substitutes y_{k+1};
//   (d) the ring row of step k takes the blocks of row k + M (fetched with cp.async a whole
//       generation ahead into the thread's private cell).
__global__ void __launch_bounds__(kBandThreads, 1)
k_lg_solve(const BAWin* __restrict__ wins, LgState* stt, int M) {
  extern __shared__ __align__(16) unsigned char smem[];
  if (!stt->active) return;
  const BAWin& W = wins[0];
  const int n = W.Ncf, T = M * M, t = threadIdx.x, lane = t & 31;
  const int n6 = n * 6;
  double* cells = reinterpret_cast<double*>(smem);   // [T][36]; back substitution: ring [2][kBandHalf][M][36]
  double* P = cells + (size_t)band_cells(M) * 36;    // [M][37] L_ik of the current step
  double* Praw = P + M * 37;                         // [M][37] blocks of the next column, before the solve
  double* Dk = Praw + M * 37;                        // 36: L_kk (strict lower) with reciprocal diagonal; +6: y_k
  double* Dnext = Dk + 48;                           // [2][36] diagonal blocks published by their owners
  double* Dtmp = Dnext + 72;                         // 36 (+4)
  double* yv = Dtmp + 40;                            // n6: right-hand side -> y -> x
  double* redv = yv + n6;                            // 64: final reduction
  __shared__ int s_fail;
  const double lambda = stt->lambda;
  const double* __restrict__ S = W.S;
  double* __restrict__ Lg = W.Lband;
  const bool ring = t < T;
  const bool diag_warp = t >= kBandThreads - 32;
  // column-major ring: the threads of one ring column (the panel of a step) are consecutive
  const int c = ring ? t / M : 0, r = ring ? t - c * M : 0;
  const int off = ring ? (r - c + M) % M : 0;
  double* cell = cells + (size_t)(ring ? t : 0) * 36;
  for (int i = t; i < n6; i += blockDim.x) yv[i] = W.bs[i];
  if (t == 0) s_fail = 0;
  // fetch of block (i, i - off): the transposed upper block (i - off, i)
  auto prefetch = [&](int i) {
    const int j = i - off;
    if (ring && j >= 0 && i < n) {
      const double* src = S + ((size_t)j * M + off) * 36;
#pragma unroll
      for (int q = 0; q < 18; q++) cp_async16(cell + q * 2, src + q * 2);
    }
    cp_async_commit();
  };
  double a[36];
  auto take = [&](int i) {  // registers <- cell (transposed), + lambda on the diagonal of a diagonal block
    const int j = i - off;
    const bool live = ring && j >= 0 && i < n;
    cp_async_wait_all();
#pragma unroll
    for (int x = 0; x < 6; x++)
#pragma unroll
      for (int y = 0; y < 6; y++) a[x * 6 + y] = live ? cell[y * 6 + x] : 0.0;
    if (live && off == 0) {
#pragma unroll
      for (int x = 0; x < 6; x++) a[x * 7] += lambda;
    }
    return live;
  };
  auto publish = [&](double* dst) {
#pragma unroll
    for (int e = 0; e < 36; e++) dst[e] = a[e];
  };
  // The diagonal warp: Dnext[kk & 1] minus the update of step kk - 1 -> Cholesky -> Dk, y_kk, factor column.
  auto factor_diag = [&](int kk, bool with_update) {
    const double* Dn = Dnext + (kk & 1) * 36;
    const double* P1 = P + 37;
    for (int e = lane; e < 36; e += 32) {
      const int x = e / 6, y = e - x * 6;
      double v = Dn[e];
      if (with_update) {
#pragma unroll
        for (int q = 0; q < 6; q++) v = fma(-P1[x * 6 + q], P1[y * 6 + q], v);
      }
      Dtmp[e] = v;
    }
    __syncwarp();
    if (lane == 0) {
      double d[36];
#pragma unroll
      for (int e = 0; e < 36; e++) d[e] = Dtmp[e];
      bool bad = false;
#pragma unroll
      for (int q = 0; q < 6; q++) {
        const double dd = d[q * 7];
        if (!(dd > 0.0)) bad = true;
        const double ri = rsqrt(bad ? 1.0 : dd);
        d[q * 7] = ri;
#pragma unroll
        for (int x = q + 1; x < 6; x++) d[x * 6 + q] *= ri;
#pragma unroll
        for (int y = q + 1; y < 6; y++)
#pragma unroll
          for (int x = y; x < 6; x++) d[x * 6 + y] -= d[x * 6 + q] * d[y * 6 + q];
      }
      if (bad) s_fail = 1;
      double yk[6];
#pragma unroll
      for (int x = 0; x < 6; x++) {
        double v = yv[kk * 6 + x];
#pragma unroll
        for (int y = 0; y < x; y++) v -= d[x * 6 + y] * yk[y];
        yk[x] = v * d[x * 7];
      }
#pragma unroll
      for (int x = 0; x < 6; x++) { yv[kk * 6 + x] = yk[x]; Dk[36 + x] = yk[x]; }
      double2* Lk = reinterpret_cast<double2*>(Lg + (size_t)kk * M * 36);
#pragma unroll
      for (int x = 0; x < 6; x++)
#pragma unroll
        for (int y = 0; y < 6; y += 2) {
          const double v0 = y <= x ? d[x * 6 + y] : 0.0, v1 = y + 1 <= x ? d[x * 6 + y + 1] : 0.0;
          Dk[x * 6 + y] = v0; Dk[x * 6 + y + 1] = v1;
          Lk[(x * 6 + y) >> 1] = make_double2(v0, v1);
        }
    }
    __syncwarp();
  };
  const long long t_start = clock64();
  long long t_seg[3] = {0, 0, 0}, t_upd = 0, t_take = 0;
  int i_cur = r;                 // row of the block this thread holds
  prefetch(i_cur);
  bool have = take(i_cur);
  int pending = i_cur + M;       // row of the next block of this thread
  bool need_fetch = true;        // its fetch is issued a barrier after the cell was read
  if (have && off == 0 && i_cur < 2) publish(Dnext + i_cur * 36);
  if (have && off > 0 && i_cur - off == 0) publish(Praw + off * 37);
  __syncthreads();
  prefetch(pending);             // every cell has been read
  need_fetch = false;
  if (diag_warp) factor_diag(0, false);
  __syncthreads();

  for (int k = 0; k < n; k++) {
    const int j_cur = i_cur - off;
    const long long t0 = clock64();
    if (s_fail) break;  // uniform: written before the last barrier
    // (b) panel, one row per thread: L_ik[x][:] = A_ik[x][:] L_kk^-T, y_i[x] -= L_ik[x][:] y_k
    if (t < 6 * (M - 1)) {
      const int d = 1 + t / 6, x = t - (d - 1) * 6;
      if (k + d < n) {
        const double* src = Praw + d * 37 + x * 6;
        double row[6];
#pragma unroll
        for (int q = 0; q < 6; q++) row[q] = src[q];
#pragma unroll
        for (int q = 0; q < 6; q++) {
          double v = row[q];
#pragma unroll
          for (int p2 = 0; p2 < 6; p2++)
            if (p2 < q) v = fma(-row[p2], Dk[q * 6 + p2], v);
          row[q] = v * Dk[q * 7];
        }
        double* dst = P + d * 37 + x * 6;
        double yy = yv[(k + d) * 6 + x];
#pragma unroll
        for (int q = 0; q < 6; q++) { dst[q] = row[q]; yy = fma(-row[q], Dk[36 + q], yy); }
        yv[(k + d) * 6 + x] = yy;
        double2* Lk = reinterpret_cast<double2*>(Lg + ((size_t)k * M + d) * 36 + x * 6);
        Lk[0] = make_double2(row[0], row[1]);
        Lk[1] = make_double2(row[2], row[3]);
        Lk[2] = make_double2(row[4], row[5]);
      }
    }
    if (have && j_cur == k) have = false;  // column k is eliminated (panel rows / diagonal warp)
    __syncthreads();
    const long long t1 = clock64();
    // (c) trailing update A_ij -= L_ik L_jk^T; the diagonal warp factorises block k + 1 meanwhile
    if (have && j_cur > k) {
      const double* Pi = P + (i_cur - k) * 37;
      const double* Pj = P + (j_cur - k) * 37;
#pragma unroll
      for (int hh = 0; hh < 2; hh++) {
        double lj[18];
#pragma unroll
        for (int e = 0; e < 18; e++) lj[e] = Pj[hh * 18 + e];
#pragma unroll
        for (int x = 0; x < 6; x++) {
          double li[6];
#pragma unroll
          for (int q = 0; q < 6; q++) li[q] = Pi[x * 6 + q];
#pragma unroll
          for (int y = 0; y < 3; y++) {
            double v = a[x * 6 + hh * 3 + y];
#pragma unroll
            for (int q = 0; q < 6; q++) v = fma(-li[q], lj[y * 6 + q], v);
            a[x * 6 + hh * 3 + y] = v;
          }
        }
      }
      if (off == 0 && i_cur == k + 2) publish(Dnext + (i_cur & 1) * 36);  // complete up to the update of step k
      if (off > 0 && j_cur == k + 1) publish(Praw + off * 37);            // the panel of the next step
    } else if (diag_warp && k + 1 < n) {
      const long long td = clock64();
      factor_diag(k + 1, M > 1);
      if (lane == 0) t_seg[2] += clock64() - td;
    }
    const long long t1b = clock64();
    // (d) row k + M enters the window: its blocks go to ring row k mod M.  The take comes BEFORE the
    // fetches of this step are issued: the cp.async scoreboard is per warp, so a wait behind another
    // lane's fresh fetch would wait for that fetch (one L2 round trip per step).
    const bool fetch_now = need_fetch;
    need_fetch = false;
    if (ring && pending == k + M) {
      have = take(pending);
      i_cur = pending;
      pending += M;
      need_fetch = true;
      if (have && off == 0 && i_cur == k + 2) publish(Dnext + (i_cur & 1) * 36);  // M == 2
      if (have && off > 0 && i_cur - off == k + 1) publish(Praw + off * 37);      // off == M - 1
    }
    if (fetch_now) prefetch(pending);
    const long long t1c = clock64();
    __syncthreads();
    t_seg[0] += t1 - t0; t_seg[1] += clock64() - t1;
    if (t == 0) { t_upd += t1b - t1; t_take += t1c - t1b; }
  }
  __syncthreads();
  const long long t_fact = clock64();
  if (s_fail) {
    cp_async_wait_all();
    for (int i = t; i < n6; i += blockDim.x) W.xp[i] = 0.0;
    if (t == 0) { stt->ok2 = 0; stt->scale_pose = 0.0; }
    return;
  }
  cp_async_wait_all();
  __syncthreads();
  // ---- back substitution L^T x = y.  The factor comes back through a two-half ring in shared memory:
  // while warp 0 consumes the kBandHalf columns of one half (lane = (d mod 5, q) multiplies, the five
  // groups are combined in fixed order by shuffles, every lane solves the 6x6 triangle redundantly),
  // the other warps fetch the next kBandHalf columns into the other half.
  {
    const int col_d = M * 36;  // doubles per column
    auto fetch_batch = [&](int b, int first, int nthr) {  // columns k_hi(b) .. k_lo(b) -> half b & 1
      const int k_hi = n - 1 - b * kBandHalf;
      if (k_hi < 0) return;
      const int k_lo = k_hi - kBandHalf + 1 > 0 ? k_hi - kBandHalf + 1 : 0;
      const int chunks = (k_hi - k_lo + 1) * M * 18;
      double* dst = cells + (size_t)(b & 1) * kBandHalf * col_d;
      for (int ch = first; ch < chunks; ch += nthr) {
        const int kc = ch / (M * 18), w2 = ch - kc * (M * 18);
        cp_async16(dst + (size_t)kc * col_d + w2 * 2, Lg + (size_t)(k_hi - kc) * col_d + w2 * 2);
      }
    };
    fetch_batch(0, t, blockDim.x);
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();
    const int g = lane / 6, q = lane - g * 6;
    double xl[6] = {0, 0, 0, 0, 0, 0};  // x_{k+1}
    const int n_batch = (n + kBandHalf - 1) / kBandHalf;
    for (int b = 0; b < n_batch; b++) {
      if (t >= 32) {
        fetch_batch(b + 1, t - 32, blockDim.x - 32);
        cp_async_commit();
      } else {
        const int k_hi = n - 1 - b * kBandHalf;
        const int k_lo = k_hi - kBandHalf + 1 > 0 ? k_hi - kBandHalf + 1 : 0;
        const double* half = cells + (size_t)(b & 1) * kBandHalf * col_d;
        for (int k = k_hi; k >= k_lo; k--) {
          const double* col = half + (size_t)(k_hi - k) * col_d;
          double v = 0.0;
          if (g < 5) {
            double vv[4] = {0, 0, 0, 0};
#pragma unroll
            for (int u = 0; u < 4; u++) {  // d = 1 + g + 5u <= 16: independent chains, operands loaded up front
              const int d = 1 + g + 5 * u;
              if (d < M && k + d < n) {
                const double* Lb = col + d * 36 + q;
                const double* xv = yv + (k + d) * 6;
                double lb[6], xx[6];
#pragma unroll
                for (int x = 0; x < 6; x++) { lb[x] = Lb[x * 6]; xx[x] = (u == 0 && g == 0) ? xl[x] : xv[x]; }
#pragma unroll
                for (int x = 0; x < 6; x++) vv[u] = fma(lb[x], xx[x], vv[u]);
              }
            }
            v = (vv[0] + vv[1]) + (vv[2] + vv[3]);
          }
          double tot = __shfl_sync(0xffffffffu, v, q);
#pragma unroll
          for (int g2 = 1; g2 < 5; g2++) tot += __shfl_sync(0xffffffffu, v, g2 * 6 + q);
          const double rq = yv[k * 6 + q] - tot;
          double rr[6];
#pragma unroll
          for (int x = 0; x < 6; x++) rr[x] = __shfl_sync(0xffffffffu, rq, x);
#pragma unroll
          for (int x = 5; x >= 0; x--) {
            double s2 = rr[x];
#pragma unroll
            for (int x2 = 5; x2 > x; x2--) s2 = fma(-col[x2 * 6 + x], xl[x2], s2);
            xl[x] = s2 * col[x * 7];
          }
          if (lane < 6) yv[k * 6 + lane] = xl[lane];
          __syncwarp();
        }
      }
      cp_async_wait_all();
      __syncthreads();
    }
  }
  const long long t_back = clock64();
  lg_solve_tail(W, stt, yv, redv, lambda);
  if (t == 0) {
    const long long t_end = clock64();
    g_lg_timing[0] += t_upd; g_lg_timing[7] += t_take; g_lg_timing[1] += t_seg[0]; g_lg_timing[2] += t_seg[1];
    g_lg_timing[3] += t_back - t_fact; g_lg_timing[4] += t_end - t_start;
    g_lg_timing[6] += n;
  }
  if (t == kBandThreads - 32) atomicAdd(&g_lg_timing[5], (unsigned long long)t_seg[2]);  // diagonal warp: factorisation time
}

// ------------------------------------------------------------------------------- phase BACKSUB

// x_l = Dinv (b_l - sum_i B_i^T (J_i x_ci)); X' = X + x_l; trial robust chi2; landmark part of
// computeScale.  Lane = observation (packed groups), per-point sums in observation order.
__global__ void __launch_bounds__(kLgThreads)
k_lg_backsub(const BAWin* __restrict__ wins, BARun run, const LgState* __restrict__ stt, double* scal) {
  if (!stt->active || !stt->ok2) return;
  __shared__ double s_c3[kLgWarps][3 * kPackSlots];
  __shared__ double red[64];
  const BAWin& W = wins[0];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int gw = blockIdx.x * kLgWarps + wid, gstride = gridDim.x * kLgWarps;
  const int cur = stt->cur, tr = cur ^ 1;
  const bool robust = stt->robust != 0;
  const double lambda = stt->lambda, delta = run.delta;
  const double* __restrict__ camRt = W.camRt[cur];
  const double* __restrict__ camRtT = W.camRt[tr];
  const double* __restrict__ pts_cur = W.pts[cur];
  double* __restrict__ pts_tr = W.pts[tr];
  const double* __restrict__ xp = W.xp;
  const double K[4] = {W.intr[0], W.intr[1], W.intr[2], W.intr[3]};
  double* c3s = s_c3[wid];
  if (lane < 3) c3s[lane * kPackSlots + kPackZero] = 0.0;
  __syncwarp();
  double chi_acc = 0.0, scale_acc = 0.0;
  for (int g = gw; g < W.n_grp; g += gstride) {
    const RecView r = load_rec(W.rec, g, lane);
    const int cbase = W.grp_cbase[g];
    const double X[3] = {pts_cur[(size_t)r.pl * 3], pts_cur[(size_t)r.pl * 3 + 1], pts_cur[(size_t)r.pl * 3 + 2]};
    double Di[6], bb[3];
#pragma unroll
    for (int a = 0; a < 6; a++) Di[a] = W.Dinv[(size_t)r.pl * 6 + a];
#pragma unroll
    for (int a = 0; a < 3; a++) bb[a] = W.bl[(size_t)r.pl * 3 + a];
    const bool active = r.valid && !r.lev;
    double c3[3] = {0, 0, 0};
    if (active && r.cf >= 0) {
      const double* Rt = camRt + (size_t)r.c * 12;
      double pc[3], pz[3], e0, e1, w, Jp[12], Jx[6];
      map_point(Rt, X, pc);
      const double e2 = edge_error(pc, r.u, r.v, K, e0, e1, pz);
      huber_rho(e2, delta, robust, w);
      edge_jac_pose(pz, K, Jp);
      edge_jac_point(Rt, pz, K, Jx);
      const double* xc = xp + (size_t)(cbase + r.cf) * 6;
      double t0 = 0, t1 = 0;
#pragma unroll
      for (int a = 0; a < 6; a++) {
        const double xa = __ldcg(xc + a);
        t0 += Jp[a] * xa;
        t1 += Jp[6 + a] * xa;
      }
      t0 *= w; t1 *= w;
#pragma unroll
      for (int a = 0; a < 3; a++) c3[a] = Jx[a] * t0 + Jx[3 + a] * t1;
    }
#pragma unroll
    for (int a = 0; a < 3; a++) c3s[a * kPackSlots + lane] = c3[a];
    __syncwarp();
    double cs[3] = {0, 0, 0};
    {
      const int len = r.valid ? r.s1 - r.s0 : 0;
      const int maxlen = __reduce_max_sync(0xffffffffu, len);
      for (int t = 0; t < maxlen; t++) {
        const int s = t < len ? r.s0 + t : kPackZero;
#pragma unroll
        for (int a = 0; a < 3; a++) cs[a] += c3s[a * kPackSlots + s];
      }
    }
    double Xn[3] = {X[0], X[1], X[2]};
    if (r.valid) {
      const double r0 = bb[0] - cs[0], r1 = bb[1] - cs[1], r2 = bb[2] - cs[2];
      const double x0 = Di[0] * r0 + Di[1] * r1 + Di[2] * r2;
      const double x1 = Di[1] * r0 + Di[3] * r1 + Di[4] * r2;
      const double x2 = Di[2] * r0 + Di[4] * r1 + Di[5] * r2;
      Xn[0] += x0; Xn[1] += x1; Xn[2] += x2;
      if (lane == r.s0) {
        pts_tr[(size_t)r.pl * 3] = Xn[0]; pts_tr[(size_t)r.pl * 3 + 1] = Xn[1]; pts_tr[(size_t)r.pl * 3 + 2] = Xn[2];
        scale_acc += x0 * (lambda * x0 + bb[0]) + x1 * (lambda * x1 + bb[1]) + x2 * (lambda * x2 + bb[2]);
      }
    }
    if (active) {
      const double* Rt = camRtT + (size_t)r.c * 12;
      double pc[3], e0, e1, w;
      map_point(Rt, Xn, pc);
      const double e2 = edge_error(pc, r.u, r.v, K, e0, e1);
      chi_acc += huber_rho(e2, delta, robust, w);
    }
    __syncwarp();
  }
  chi_acc = warp_sum(chi_acc);
  scale_acc = warp_sum(scale_acc);
  if (lane == 0) { red[wid] = chi_acc; red[32 + wid] = scale_acc; }
  __syncthreads();
  if (tid == 0) {
    double s0 = 0.0, s1 = 0.0;
    for (int w = 0; w < kLgWarps; w++) { s0 += red[w]; s1 += red[32 + w]; }
    W.cpart[(size_t)blockIdx.x * 4] = s0;
    W.cpart[(size_t)blockIdx.x * 4 + 1] = s1;
  }
  finish_partials(W.cpart, W.ticket + 1, 2, scal + 2, scal + 3, nullptr, false);
}

// after the [chi2', scale] all-reduce: gain ratio, damping update, accept / reject (one thread);
// OptimizationAlgorithmLevenberg::solve + the loop of SparseOptimizer::optimize
__global__ void k_lg_decide(LgState* stt, const double* scal, LgState* host_copy) {
  LgState s = *stt;
  if (!s.active) { *host_copy = s; return; }
  const double currentChi = scal[0];
  const double tempChi = s.ok2 ? scal[2] : 1.7976931348623157e308;
  const double scale = s.ok2 ? scal[3] + s.scale_pose : 0.0;
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
  const bool cont = !lambda_bad && rho < 0 && s.qmax < 10;
  if (!cont) {
    s.iters = s.it + 1;
    const bool terminate = s.qmax == 10 || rho == 0 || lambda_bad || !isfinite(s.lambda);
    s.it++;
    s.qmax = 0;
    if (terminate || s.it >= s.n_iter) s.active = 0;
  }
  *stt = s;
  *host_copy = s;
}

// src/g2o_optimization.cc:129-135 (pass 0) / :150-154 (pass 1): one thread per point, see
// window_classify in ba_kernels.cu for the cached-error semantics.
__global__ void __launch_bounds__(kLgThreads)
k_lg_classify(const BAWin* __restrict__ wins, BARun run, LgState* stt, int pass) {
  const BAWin& W = wins[0];
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
  const double K[4] = {W.intr[0], W.intr[1], W.intr[2], W.intr[3]};
  const int cur = stt->cur, last_eval = stt->last_eval;
  const bool have_eval = stt->have_eval != 0;
  const double chi2_thr = run.chi2_thr;
  int n_l1 = 0;
  for (int l = gt; l < W.Np; l += gstride) {
    for (int o = W.pt_start[l]; o < W.pt_start[l + 1]; o++) {
      const int c = W.ocam[o];
      double pc[3], e0, e1;
      const double2 uv = *reinterpret_cast<const double2*>(W.uv + (size_t)o * 2);
      const int lev = W.level[o];
      if (pass == 1 && lev == 1) continue;
      map_point(W.camRt[last_eval] + (size_t)c * 12, W.pts[last_eval] + (size_t)l * 3, pc);
      const double e2 = have_eval ? edge_error(pc, uv.x, uv.y, K, e0, e1) : 0.0;
      map_point(W.camRt[cur] + (size_t)c * 12, W.pts[cur] + (size_t)l * 3, pc);
      const bool depth_pos = pc[2] > 0.0;
      if (pass == 0) {
        const int nl = (e2 > chi2_thr) ? 1 : (!depth_pos ? 2 : 0);
        W.level[o] = (uint8_t)nl;
        W.inlier[o] = 0;
        n_l1 += nl ? 1 : 0;
      } else if (lev == 2) {
        W.inlier[o] = depth_pos ? 1 : 0;
      } else {
        W.inlier[o] = (e2 <= chi2_thr && depth_pos) ? 1 : 0;
      }
    }
  }
  if (pass == 0) {
    for (int o = 16; o > 0; o >>= 1) n_l1 += __shfl_xor_sync(0xffffffffu, n_l1, o);
    if ((threadIdx.x & 31) == 0 && n_l1) atomicAdd(&stt->n_level1, n_l1);
  }
}

__global__ void k_lg_end_pass(const BAWin* __restrict__ wins, const LgState* stt, int pass) {
  urmvo_ba_stats* st = reinterpret_cast<urmvo_ba_stats*>(wins[0].stats);
  st->iters[pass] = stt->iters;
  st->trials[pass] = stt->trials;
  st->pcg_iters[pass] = stt->pcg_iters;
  st->chi2_final[pass] = stt->currentChi;
  st->lambda_final[pass] = stt->lambda;
  if (pass == 0) { st->chi2_initial = stt->chi_initial; st->n_level1 = stt->n_level1; }
}

// write back T_wc = estimate().inverse() and the points (g2o_optimization.cc:164-176)
__global__ void __launch_bounds__(kLgThreads)
k_lg_finish(const BAWin* __restrict__ wins, const LgState* stt) {
  const BAWin& W = wins[0];
  const int cur = stt->cur;
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
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

// ------------------------------------------------------------------------------- host launchers

size_t lg_state_bytes() { return sizeof(LgState); }
cudaError_t lg_timing_read(unsigned long long* out, bool reset) {
  cudaError_t e = cudaMemcpyFromSymbol(out, g_lg_timing, sizeof(unsigned long long) * 8);
  if (e == cudaSuccess && reset) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    e = cudaMemcpyToSymbol(g_lg_timing, z, sizeof(z));
  }
  return e;
}
int lg_band_max_m() { return kBandMaxM; }

static size_t lg_lin_smem() {
  return (size_t)4 * 32 * sizeof(double) + (size_t)2 * kLgWarps * kLgStage * sizeof(double) +
         (size_t)2 * kLgWarps * kLgSlotBytes + 2 * kLgWarps * sizeof(int) + 32;  // two staging buffers per warp
}

cudaError_t lg_prepare(int M, int Ncf) {
  cudaError_t e = cudaFuncSetAttribute(k_lg_lin<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lg_lin_smem());
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_lg_lin<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lg_lin_smem());
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_lg_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)band_smem_bytes(M, Ncf));
}
cudaError_t launch_lg_init(const BAWin* w, void* stt, int grid, cudaStream_t s) {
  k_lg_init<<<grid, kLgThreads, 0, s>>>(w, (LgState*)stt);
  return cudaGetLastError();
}
cudaError_t launch_lg_begin_pass(void* stt, int robust, int n_iter, cudaStream_t s) {
  k_lg_begin_pass<<<1, 1, 0, s>>>((LgState*)stt, robust, n_iter);
  return cudaGetLastError();
}
cudaError_t launch_lg_pack(const BAWin* w, int grid, cudaStream_t s) {
  k_lg_pack<<<grid, kLgThreads, 0, s>>>(w);
  return cudaGetLastError();
}
cudaError_t launch_lg_lin(const BAWin* w, const BARun& run, void* stt, double* scal, int diag, int rank, int grid,
                          cudaStream_t s) {
  if (diag) k_lg_lin<true><<<grid, kLgThreads, lg_lin_smem(), s>>>(w, run, (const LgState*)stt, scal, rank);
  else k_lg_lin<false><<<grid, kLgThreads, lg_lin_smem(), s>>>(w, run, (const LgState*)stt, scal, rank);
  return cudaGetLastError();
}
cudaError_t launch_lg_lambda(const BAWin* w, void* stt, const double* scal, int world, cudaStream_t s) {
  k_lg_lambda<<<1, 256, 0, s>>>(w, (LgState*)stt, scal, world);
  return cudaGetLastError();
}
cudaError_t launch_lg_solve(const BAWin* w, void* stt, int M, int Ncf, cudaStream_t s) {
  k_lg_solve<<<1, kBandThreads, band_smem_bytes(M, Ncf), s>>>(w, (LgState*)stt, M);
  return cudaGetLastError();
}
cudaError_t launch_lg_backsub(const BAWin* w, const BARun& run, void* stt, double* scal, int grid, cudaStream_t s) {
  k_lg_backsub<<<grid, kLgThreads, 0, s>>>(w, run, (const LgState*)stt, scal);
  return cudaGetLastError();
}
cudaError_t launch_lg_decide(void* stt, const double* scal, void* host_copy, cudaStream_t s) {
  k_lg_decide<<<1, 1, 0, s>>>((LgState*)stt, scal, (LgState*)host_copy);
  return cudaGetLastError();
}
cudaError_t launch_lg_classify(const BAWin* w, const BARun& run, void* stt, int pass, int grid, cudaStream_t s) {
  k_lg_classify<<<grid, kLgThreads, 0, s>>>(w, run, (LgState*)stt, pass);
  if (cudaGetLastError() != cudaSuccess) return cudaErrorLaunchFailure;
  k_lg_end_pass<<<1, 1, 0, s>>>(w, (const LgState*)stt, pass);
  return cudaGetLastError();
}
cudaError_t launch_lg_finish(const BAWin* w, void* stt, int grid, cudaStream_t s) {
  k_lg_finish<<<grid, kLgThreads, 0, s>>>(w, (const LgState*)stt);
  return cudaGetLastError();
}
void lg_flags(const void* host_copy, int* active, int* it) {
  const LgState* s = (const LgState*)host_copy;
  *active = s->active;
  *it = s->it;
}

}  // namespace urmvo
