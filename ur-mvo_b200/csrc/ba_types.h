// csrc/ba_types.h — device-side problem descriptors shared by the kernels and the C-ABI layer.
#pragma once
#include <stdint.h>

namespace urmvo {

// Packed modes: one 32-byte record per (group, lane), rebuilt on the device by pack_records() at the
// start of a solve and after the outlier classification.  A lane reads everything that does not
// change inside an LM pass with two 16-byte loads, one group ahead of its use (the CSR arrays would
// need three dependent levels of loads).
struct alignas(16) ObsRec {
  double u, v;
  int pl;                  // point index (first point of the group for an idle lane)
  int c;                   // camera index
  unsigned char s0, s1;    // the point's observation range, relative to the group's first observation
  unsigned char pi;        // point index inside the group
  signed char cf;          // dense free-camera index or -1
  unsigned char np;        // points in the group (same in all lanes)
  unsigned char lev;       // g2o edge level (1 for an idle lane)
  unsigned char valid;     // lane carries an observation
  unsigned char pad;
};

// One BA window (one LocalmapOptimization call), all pointers are device pointers.
// HBM layout: SoA, fp64; observations sorted point-major so that one point's
// observations are contiguous; cameras as 7-double (q,t) T_cw state + a derived 12-double (R|t).
struct BAWin {
  int Nc, Ncf, Np, No;
  int nblk;         // stored 6x6 blocks of the reduced camera system (upper triangle, BSR)
  int kmax;         // max observations of one point (sizes the per-warp staging area)
  int acc_mode;     // 0: fp64 atomics into the global block-sparse S + BSR PCG (large systems)
                    // 1: warp-private shared-memory copies + dense in-smem PCG (<= 16 free cameras)
                    // 2/3: packed groups, 1/2 register-resident blocks per lane (+ dense PCG)
                    // 4: tile mode (csrc/ba_large.cu): point chunks with S-stationary register blocks,
                    //    band storage of S and a direct block-banded Cholesky solve
                    // 5 / 6: modes 1 / 0 for windows with STEREO edges (3-row residuals, okind / ur below)
  int n_grp;        // packed modes: number of point groups
  int acc_len;      // acc_mode 1: doubles per accumulator copy = nblk*36 + Ncf*12
  double intr[4];   // fx fy cx cy
  // ---- inputs (immutable during a run)
  const double* pose_in;   // Nc*7   T_wc initial estimate
  const double* pts_in;    // Np*3
  const double* uv;        // No*2
  const int* ocam;         // No     camera index
  const double* ur;        // No     u_right of a stereo edge (modes 5 / 6), else NULL
  const uint8_t* okind;    // No     bit 0: stereo edge (EdgeStereoSE3ProjectXYZ) / mono edge, bits 1-7: camera model
                           //        of the edge (row of intr_tab); NULL in mono single-camera windows
  const double* intr_tab;  // n_models rows of (fx fy cx cy bf): camera_list[mpc->id_camera]; modes 5 / 6 only
  const int* pt_start;     // Np+1   CSR over observations
  const int* opt;          // No     point index of each observation (packed modes)
  const int* grp_pt;       // n_grp+1 first point of each group: <= 32 observations and points per group
  const int* cam_free;     // Nc     dense index among free cameras or -1
  const int* row_ptr;      // Ncf+1  upper BSR of S: block row i -> [row_ptr[i], row_ptr[i+1])
  const int* col;          // nblk   block column (>= row), ascending inside a row, col[row_ptr[i]] == i
  const int* lrow_ptr;     // Ncf+1  mirror lists: block row i -> blocks (k,i) with k < i
  const int* lcol;         // k
  const int* lblk;         // index of block (k,i) in S
  // ---- state
  double* cam[2];          // Nc*7   T_cw (q,t): current / trial (roles swap on accept)
  double* camRt[2];        // Nc*12  R (row-major 9) | t (3) derived from cam[]
  double* pts[2];          // Np*3
  uint8_t* level;          // No     0 active, 1 excluded (g2o edge level)
  ObsRec* rec;             // n_grp*32 packed-mode observation records (device-built)
  // ---- linear system
  double* S;               // nblk*36 row-major blocks
  double* bs;              // Ncf*6  Schur right-hand side
  double* bp;              // Ncf*6  raw pose gradient b_p (for computeScale)
  double* hdiag;           // Ncf*6  raw diag(Hpp) (for computeLambdaInit)
  double* Minv;            // Ncf*36 block-Jacobi preconditioner
  double* xp;              // Ncf*6
  double* r;               // Ncf*6
  double* z;               // Ncf*6
  double* p;               // Ncf*6
  double* Ap;              // Ncf*6
  double* Dinv;            // Np*6   (Hll + lambda I)^-1, symmetric packed 00 01 02 11 12 22
  double* bl;              // Np*3
  double* part;            // scope reduction scratch: 2 * nblk_scope * 8
  double* Spart;           // acc_mode 1: one accumulator copy per CTA of the scope (nblk_scope * acc_len)
  // ---- tile mode (acc_mode 4): one large problem as phase kernels, csrc/ba_large.cu.  Points are
  // renumbered by their first free camera and cut into CHUNKS of consecutive points whose reduced-
  // system blocks (<= 256, inside a window of <= 32 consecutive free cameras) are owned one per thread.
  // S is stored as a full band: block (i, j), i <= j <= i + bw, at row_ptr[i] + (j - i).
  int n_chunk, bw;
  const int* chunk_grp;    // n_chunk+1  packed groups of each chunk
  const int* grp_cbase;    // n_grp      first free camera of the chunk window of a group
  const int* chunk_blk;    // n_chunk+1  blocks owned by the threads of a chunk, into blk_desc
  const int* blk_desc;     // 2 ints per block: (ci - cbase) | (cj - cbase) << 8, index of the block in S
  double* Lband;           // Ncf*(bw+1)*36 block columns of the Cholesky factor (diagonals as reciprocals)
  double* cpart;           // per-CTA partial sums of the phase kernels [grid][4]
  unsigned int* ticket;    // last-CTA tickets of the phase kernels
  // ---- outputs
  double* pose_out;        // Nc*7 T_wc
  double* pts_out;         // Np*3
  uint8_t* inlier;         // No
  void* stats;             // urmvo_ba_stats*
};

struct BARun {
  double chi2_thr;
  double delta;        // (double)(float)sqrt(chi2_thr)
  double chi2_thr_s;   // cfg.stereo_point
  double delta_s;      // (double)(float)sqrt(chi2_thr_s)
  double pcg_tol;
  int pcg_max_iter;
  int it0, it1;
  int n_win;
  int dense_solver;    // in-shared-memory systems: 0 tiled Cholesky (default), 1 one-barrier-per-pivot LDL^T, 2 block-Jacobi PCG
  const void* timing_stats;  // the window whose stats pointer equals this records phase cycles
};

// Device-side Levenberg-Marquardt state of one large problem in tile mode (csrc/ba_large.cu, ba_bcr.cu).
struct LgState {
  double lambda, ni, currentChi, rho, chi_initial, scale_pose;
  int cur, last_eval, have_eval;
  int robust, it, n_iter, qmax, ok2;
  int iters, trials, pcg_iters, n_level1;
  int active;  // the current optimize() call still has trials to run
};

// One pose-only frame.
struct PoseFrame {
  int No;
  int obs0;            // first observation in the concatenated arrays
};

}  // namespace urmvo
