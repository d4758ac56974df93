// csrc/ba_bcr.cu — direct solve of a LONG block-banded reduced camera system by block cyclic reduction
// (tile mode of csrc/ba_large.cu; BASELINE.json configs[4]: 1000 cameras along a trajectory).
//
// Reference behaviour: g2o's LinearSolverEigen factorises the reduced camera system with a sparse
// Cholesky (/root/reference/src/g2o_optimization.cc:27-35).  The sequential block-banded Cholesky of
// ba_large.cu (k_lg_solve) is that algorithm on one SM: one block column after the other, 2.4 us per
// camera, 2.4 ms for 1000 cameras — 70 % of a damped trial, and replicated on every rank of the
// point-sharded solve.  The dependency chain of a banded factorisation is as long as the trajectory;
// cyclic reduction is the same elimination in a different ORDER (every second super-block first, then
// every second of the rest, ...), which is still an exact Cholesky-type factorisation of the SPD system
// (nested dissection order instead of the natural one) but has depth log2(K) instead of K:
//
//   cameras are grouped into K super-blocks of m >= half-bandwidth cameras, so S is block TRIDIAGONAL
//   with dense mb x mb blocks (mb = 6 m).  Level l keeps the super-blocks whose index is a multiple of
//   s = 2^l; those with an odd index/s are eliminated:
//       D_k = L L^T,  Z_a = C_ak L^-T,  Z_c = C_ck L^-T,  y_k = L^-1 b_k
//       D_a -= Z_a Z_a^T,  D_c -= Z_c Z_c^T,  C_ac(new) = -Z_a Z_c^T,  b_a -= Z_a y_k,  b_c -= Z_c y_k
//   and after the last level x_0 = D_0^-1 b_0, then downwards x_k = L^-T (y_k - Z_a^T x_a - Z_c^T x_c).
//   Every coupling block is stored [surviving rows][eliminated columns], so all products are of the form
//   A B^T with the contracted index contiguous in both operands.
//
// Kernels per level: k_bcr_chol (one CTA per eliminated super-block, in shared memory), k_bcr_trsm (rows
// of the two coupling blocks against L, 4 rows per warp), k_bcr_update (48 x 48 output tiles of the
// products).  All sums have a fixed order: the solve is deterministic, every rank of the sharded
// problem gets the same bits from the same all-reduced system.  fp64 throughout.

#include "ba_device.cuh"
#include "ba_large_tail.cuh"
#include "kernels.h"

namespace urmvo {

namespace {

constexpr int kBcrThreads = 256;
constexpr int kBcrTile = 48;        // output tile of k_bcr_update (3 x 3 values per thread)
constexpr int kBcrRowsPerWarp = 4;  // vectors a warp of k_bcr_trsm carries through one substitution

__device__ __forceinline__ size_t bcr_pair(const BcrShape& sh, int level, int j) {
  return sh.off_C + (sh.coff[level] + (size_t)j) * sh.mb * sh.mb;
}

// ---- level 0: dense super-blocks from the band storage (block (i, i + d) at (i * M + d) * 36) ----
// D_k lower triangle (+ lambda on the diagonal; identity for the padding cameras of the last
// super-block), the coupling of super-blocks (j, j + 1) as [surviving rows][eliminated columns], b.
__global__ void __launch_bounds__(kBcrThreads)
k_bcr_assemble(const BAWin* __restrict__ wins, const LgState* __restrict__ stt, BcrShape sh, double* __restrict__ work) {
  if (!stt->active) return;
  const BAWin& W = wins[0];
  const int n = sh.n, m = sh.m, mb = sh.mb, M = sh.M, bw = M - 1;
  const double lambda = stt->lambda;
  const double* __restrict__ S = W.S;
  const int k = blockIdx.x >> 1, which = blockIdx.x & 1;
  if (which == 0) {
    double* D = work + sh.off_D + (size_t)k * mb * mb;
    for (int e = threadIdx.x; e < mb * mb; e += blockDim.x) {
      const int r = e / mb, c = e - r * mb;
      double v = 0.0;
      if (r >= c) {
        const int ia = k * m + r / 6, ic = k * m + c / 6;
        if (ia < n) {
          if (ia - ic <= bw) v = S[((size_t)ic * M + (ia - ic)) * 36 + (c % 6) * 6 + (r % 6)];
          if (r == c) v += lambda;
        } else if (r == c) {
          v = 1.0;
        }
      }
      D[e] = v;
    }
    double* b = work + sh.off_rhs + (size_t)k * mb;
    for (int e = threadIdx.x; e < mb; e += blockDim.x) b[e] = k * mb + e < n * 6 ? W.bs[k * mb + e] : 0.0;
    if (blockIdx.x == 0 && threadIdx.x == 0) *reinterpret_cast<int*>(work + sh.off_fail) = 0;
  } else if (k + 1 < sh.K) {
    double* C = work + bcr_pair(sh, 0, k);
    const bool transposed = (k & 1) != 0;  // odd left super-block: it is the eliminated one
    for (int e = threadIdx.x; e < mb * mb; e += blockDim.x) {
      const int r0 = e / mb, c0 = e - r0 * mb;
      const int r = transposed ? c0 : r0, c = transposed ? r0 : c0;  // (r, c) of E = S[k-th rows, (k+1)-th columns]
      const int ia = k * m + r / 6, ic = (k + 1) * m + c / 6;
      double v = 0.0;
      if (ic < n && ic - ia <= bw) v = S[((size_t)ia * M + (ic - ia)) * 36 + (r % 6) * 6 + (c % 6)];
      C[e] = v;
    }
  }
}

// ---- Cholesky of one mb x mb block in shared memory + y = L^-1 b (+ x = L^-T y for the last block) ----
// Input: lower triangle of D (row-major).  Output F[r][c] = L[max(r,c)][min(r,c)] (row j of F holds row j
// of L up to the diagonal and column j of L after it: both substitutions read rows of F), the
// reciprocal diagonal in invd, y over b.  Right-looking, one barrier per column: the column is read
// unscaled, A[i][c] -= A[i][j] A[c][j] / A[j][j].
__device__ __forceinline__ void bcr_chol_block(double* A, int lda, int mb, double* invd_s, int* fail_flag) {
  const int t = threadIdx.x, nt = blockDim.x;
  for (int j = 0; j < mb; j++) {
    const double d = A[j * lda + j];
    const bool bad = !(d > 0.0);
    const double ri = rsqrt(bad ? 1.0 : d);
    const double r2 = ri * ri;
    if (t == 0) { invd_s[j] = ri; if (bad) *fail_flag = 1; }
    // trailing lower triangle (i >= c > j): one row per warp, lanes over the columns
    const int warp = t >> 5, lane = t & 31, n_warp = nt >> 5;
    for (int i = j + 1 + warp; i < mb; i += n_warp) {
      const double f = A[i * lda + j] * r2;
      for (int c = j + 1 + lane; c <= i; c += 32) A[i * lda + c] = fma(-f, A[c * lda + j], A[i * lda + c]);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kBcrThreads)
k_bcr_chol(const LgState* __restrict__ stt, BcrShape sh, double* __restrict__ work, int level, int last) {
  extern __shared__ __align__(16) double smem_d[];
  if (!stt->active) return;
  const int mb = sh.mb, lda = mb + 1;
  const int s = 1 << level;
  const int k = last ? 0 : (2 * blockIdx.x + 1) * s;
  double* A = smem_d;                 // [mb][lda]
  double* invd_s = A + mb * lda;      // mb
  double* yv = invd_s + mb;           // mb
  int* fail = reinterpret_cast<int*>(work + sh.off_fail);
  double* D = work + sh.off_D + (size_t)k * mb * mb;
  double* b = work + sh.off_rhs + (size_t)k * mb;
  for (int e = threadIdx.x; e < mb * mb; e += blockDim.x) {
    const int r = e / mb, c = e - r * mb;
    A[r * lda + c] = D[e];
  }
  for (int e = threadIdx.x; e < mb; e += blockDim.x) yv[e] = b[e];
  __syncthreads();
  bcr_chol_block(A, lda, mb, invd_s, fail);
  // F: scaled factor, mirrored
  for (int e = threadIdx.x; e < mb * mb; e += blockDim.x) {
    const int r = e / mb, c = e - r * mb;
    const int hi = r > c ? r : c, lo = r > c ? c : r;
    D[e] = hi == lo ? 1.0 / invd_s[lo] : A[hi * lda + lo] * invd_s[lo];
  }
  for (int e = threadIdx.x; e < mb; e += blockDim.x) work[sh.off_invd + (size_t)k * mb + e] = invd_s[e];
  // y = L^-1 b by warp 0 (unscaled columns: L[i][j] = A[i][j] invd[j])
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    for (int j = 0; j < mb; j++) {
      const double yj = yv[j] * invd_s[j];
      __syncwarp();
      if (lane == 0) yv[j] = yj;
      const double f = yj * invd_s[j];
      for (int i = j + 1 + lane; i < mb; i += 32) yv[i] = fma(-A[i * lda + j], f, yv[i]);
      __syncwarp();
    }
    if (last) {  // x = L^-T y
      for (int j = mb - 1; j >= 0; j--) {
        const double xj = yv[j] * invd_s[j];
        __syncwarp();
        if (lane == 0) yv[j] = xj;
        for (int i = lane; i < j; i += 32) yv[i] = fma(-A[j * lda + i] * invd_s[i], xj, yv[i]);
        __syncwarp();
      }
    }
    double* out = last ? work + sh.off_x + (size_t)k * mb : b;
    for (int e = lane; e < mb; e += 32) out[e] = yv[e];
  }
}

// ---- Z = C L^-T for the rows of the (up to) two coupling blocks of every eliminated super-block ----
// A warp carries kBcrRowsPerWarp rows through the forward substitution L z = c; lane l holds the
// elements l, l + 32, ... of each row, row j of F (column j of L behind the diagonal) comes from shared memory.
template <int NSLOT>
__device__ __forceinline__ void bcr_trsm_rows(const double* __restrict__ F_s, int ldf, const double* __restrict__ invd_s,
                                              int mb, double* __restrict__ Zrows, int n_rows) {
  const int lane = threadIdx.x & 31;
  double v[kBcrRowsPerWarp][NSLOT];
#pragma unroll
  for (int q = 0; q < kBcrRowsPerWarp; q++)
#pragma unroll
    for (int sl = 0; sl < NSLOT; sl++) {
      const int i = sl * 32 + lane;
      v[q][sl] = (q < n_rows && i < mb) ? Zrows[(size_t)q * mb + i] : 0.0;
    }
#pragma unroll
  for (int jb = 0; jb < NSLOT; jb++) {
    const int jn = mb - jb * 32 < 32 ? mb - jb * 32 : 32;
    for (int jj = 0; jj < jn; jj++) {
      const int j = jb * 32 + jj;
      const double id = invd_s[j];
      double z[kBcrRowsPerWarp];
#pragma unroll
      for (int q = 0; q < kBcrRowsPerWarp; q++) {
        z[q] = __shfl_sync(0xffffffffu, v[q][jb], jj) * id;
        if (lane == jj) v[q][jb] = z[q];
      }
      const double* Fj = F_s + (size_t)j * ldf;
#pragma unroll
      for (int sl = jb; sl < NSLOT; sl++) {
        const int i = sl * 32 + lane;
        if (i > j && i < mb) {
          const double l = Fj[i];
#pragma unroll
          for (int q = 0; q < kBcrRowsPerWarp; q++) v[q][sl] = fma(-l, z[q], v[q][sl]);
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < kBcrRowsPerWarp; q++)
#pragma unroll
    for (int sl = 0; sl < NSLOT; sl++) {
      const int i = sl * 32 + lane;
      if (q < n_rows && i < mb) Zrows[(size_t)q * mb + i] = v[q][sl];
    }
}

// grid: (CTAs per eliminated super-block, eliminated super-blocks of the level)
__global__ void __launch_bounds__(kBcrThreads)
k_bcr_trsm(const LgState* __restrict__ stt, BcrShape sh, double* __restrict__ work, int level) {
  extern __shared__ __align__(16) double smem_d[];
  if (!stt->active) return;
  const int mb = sh.mb, ldf = mb;
  const int s = 1 << level, n_act = (sh.K + s - 1) / s;
  const int u = 2 * blockIdx.y + 1, k = u * s;
  const bool has_right = u + 1 < n_act;
  double* F_s = smem_d;
  double* invd_s = F_s + (size_t)mb * ldf;
  const double* F = work + sh.off_D + (size_t)k * mb * mb;
  for (int e = threadIdx.x; e < mb * mb; e += blockDim.x) F_s[e] = F[e];
  for (int e = threadIdx.x; e < mb; e += blockDim.x) invd_s[e] = work[sh.off_invd + (size_t)k * mb + e];
  __syncthreads();
  const int warp = threadIdx.x >> 5, n_warp = blockDim.x >> 5;
  const int total = (has_right ? 2 : 1) * mb;  // rows of Z_a, then rows of Z_c
  const int per_cta = n_warp * kBcrRowsPerWarp;
  for (int r0 = (blockIdx.x * n_warp + warp) * kBcrRowsPerWarp; r0 < total; r0 += gridDim.x * per_cta) {
    // a group of rows never straddles the two blocks when mb is a multiple of kBcrRowsPerWarp; handle the general case
    int first = r0, cnt = total - r0 < kBcrRowsPerWarp ? total - r0 : kBcrRowsPerWarp;
    while (cnt > 0) {
      const int blk = first >= mb ? 1 : 0;
      const int in_blk = first - blk * mb;
      const int take = (blk == 0 && first + cnt > mb) ? mb - first : cnt;
      double* Z = work + bcr_pair(sh, level, blk == 0 ? u - 1 : u) + (size_t)in_blk * mb;
      if (mb <= 64) bcr_trsm_rows<2>(F_s, ldf, invd_s, mb, Z, take);
      else if (mb <= 96) bcr_trsm_rows<3>(F_s, ldf, invd_s, mb, Z, take);
      else bcr_trsm_rows<4>(F_s, ldf, invd_s, mb, Z, take);
      first += take;
      cnt -= take;
    }
  }
}

// ---- products of the level: C (-)= A B^T, 48 x 48 tiles, the contracted index staged whole in shared memory ----
// Work items: (a) every surviving super-block a: D_a -= Z Z^T for the eliminated neighbours on both sides (lower
// tiles only), b_a -= Z y; (b) every eliminated super-block with two neighbours: the new coupling -Z_a Z_c^T,
// stored [surviving at level + 1][eliminated at level + 1].
__device__ __forceinline__ void bcr_load_tile(double* dst, const double* __restrict__ src, int row0, int mb) {
  // dst[k][kBcrTile + 1] <- src[row0 + i][k], zero rows beyond mb
  for (int e = threadIdx.x; e < kBcrTile * mb; e += blockDim.x) {
    const int i = e / mb, k = e - i * mb;
    dst[k * (kBcrTile + 1) + i] = row0 + i < mb ? src[(size_t)(row0 + i) * mb + k] : 0.0;
  }
}

__global__ void __launch_bounds__(kBcrThreads)
k_bcr_update(const LgState* __restrict__ stt, BcrShape sh, double* __restrict__ work, int level) {
  extern __shared__ __align__(16) double smem_d[];
  if (!stt->active) return;
  const int mb = sh.mb;
  const int s = 1 << level, n_act = (sh.K + s - 1) / s;
  const int T = (mb + kBcrTile - 1) / kBcrTile, Tl = T * (T + 1) / 2;
  const int n_surv = (n_act + 1) / 2, n_el = n_act / 2;
  double* As = smem_d;
  double* Bs = As + (size_t)mb * (kBcrTile + 1);
  double* ys = Bs + (size_t)mb * (kBcrTile + 1);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  int b = blockIdx.x;
  if (b < n_surv * Tl) {
    const int f = b / Tl;
    int tl = b - f * Tl;
    int ti = 0;
    while ((ti + 1) * (ti + 2) / 2 <= tl) ti++;
    const int tj = tl - ti * (ti + 1) / 2;
    const int v = 2 * f, a = v * s;
    double acc[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    double racc = 0.0;
    for (int side = 0; side < 2; side++) {
      // side 0: eliminated neighbour on the right (pair v, rows = a), side 1: on the left (pair v - 1, rows = a)
      if (side == 0 ? v + 1 >= n_act : v < 2) continue;
      const int pair = side == 0 ? v : v - 1;
      const int kel = side == 0 ? a + s : a - s;
      const double* Z = work + bcr_pair(sh, level, pair);
      __syncthreads();
      bcr_load_tile(As, Z, ti * kBcrTile, mb);
      if (tj != ti) bcr_load_tile(Bs, Z, tj * kBcrTile, mb);
      if (tj == 0)
        for (int e = threadIdx.x; e < mb; e += blockDim.x) ys[e] = work[sh.off_rhs + (size_t)kel * mb + e];
      __syncthreads();
      const double* Bt = tj != ti ? Bs : As;
      for (int k = 0; k < mb; k++) {
        double av[3], bv[3];
#pragma unroll
        for (int q = 0; q < 3; q++) { av[q] = As[k * (kBcrTile + 1) + ty * 3 + q]; bv[q] = Bt[k * (kBcrTile + 1) + tx * 3 + q]; }
#pragma unroll
        for (int p = 0; p < 3; p++)
#pragma unroll
          for (int q = 0; q < 3; q++) acc[p][q] = fma(av[p], bv[q], acc[p][q]);
      }
      if (tj == 0 && threadIdx.x < kBcrTile)
        for (int k = 0; k < mb; k++) racc = fma(As[k * (kBcrTile + 1) + threadIdx.x], ys[k], racc);
    }
    double* D = work + sh.off_D + (size_t)a * mb * mb;
#pragma unroll
    for (int p = 0; p < 3; p++)
#pragma unroll
      for (int q = 0; q < 3; q++) {
        const int i = ti * kBcrTile + ty * 3 + p, j = tj * kBcrTile + tx * 3 + q;
        if (i < mb && j <= i) D[(size_t)i * mb + j] -= acc[p][q];
      }
    if (tj == 0 && threadIdx.x < kBcrTile) {
      const int i = ti * kBcrTile + threadIdx.x;
      if (i < mb) work[sh.off_rhs + (size_t)a * mb + i] -= racc;
    }
    return;
  }
  b -= n_surv * Tl;
  if (b >= n_el * T * T) return;
  const int e = b / (T * T);
  const int tt = b - e * T * T, ti = tt / T, tj = tt - ti * T;
  const int u = 2 * e + 1;
  if (u + 1 >= n_act) return;  // no right neighbour: nothing to couple
  const double* Za = work + bcr_pair(sh, level, u - 1);  // rows: left neighbour
  const double* Zc = work + bcr_pair(sh, level, u);      // rows: right neighbour
  const int jn = (u - 1) / 2;                            // pair index at level + 1
  const bool left_survives = (jn & 1) == 0;
  const double* Ar = left_survives ? Za : Zc;            // rows of the new block
  const double* Bc = left_survives ? Zc : Za;            // columns of the new block
  bcr_load_tile(As, Ar, ti * kBcrTile, mb);
  bcr_load_tile(Bs, Bc, tj * kBcrTile, mb);
  __syncthreads();
  double acc[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int k = 0; k < mb; k++) {
    double av[3], bv[3];
#pragma unroll
    for (int q = 0; q < 3; q++) { av[q] = As[k * (kBcrTile + 1) + ty * 3 + q]; bv[q] = Bs[k * (kBcrTile + 1) + tx * 3 + q]; }
#pragma unroll
    for (int p = 0; p < 3; p++)
#pragma unroll
      for (int q = 0; q < 3; q++) acc[p][q] = fma(av[p], bv[q], acc[p][q]);
  }
  double* C = work + bcr_pair(sh, level + 1, jn);
#pragma unroll
  for (int p = 0; p < 3; p++)
#pragma unroll
    for (int q = 0; q < 3; q++) {
      const int i = ti * kBcrTile + ty * 3 + p, j = tj * kBcrTile + tx * 3 + q;
      if (i < mb && j < mb) C[(size_t)i * mb + j] = -acc[p][q];
    }
}

// ---- downwards: x_k = L^-T (y_k - Z_a^T x_a - Z_c^T x_c) for the super-blocks eliminated at `level` ----
__global__ void __launch_bounds__(kBcrThreads)
k_bcr_back(const LgState* __restrict__ stt, BcrShape sh, double* __restrict__ work, int level) {
  extern __shared__ __align__(16) double smem_d[];
  if (!stt->active) return;
  const int mb = sh.mb;
  const int s = 1 << level, n_act = (sh.K + s - 1) / s;
  const int u = 2 * blockIdx.x + 1, k = u * s;
  const bool has_right = u + 1 < n_act;
  double* F_s = smem_d;                       // [mb][mb]
  double* tv = F_s + (size_t)mb * mb;         // mb
  double* xa = tv + mb;                       // mb
  double* xc = xa + mb;                       // mb
  double* part = xc + mb;                     // [4][mb]
  double* invd_s = part + 4 * mb;             // mb
  const double* F = work + sh.off_D + (size_t)k * mb * mb;
  for (int e = threadIdx.x; e < mb * mb; e += blockDim.x) F_s[e] = F[e];
  for (int e = threadIdx.x; e < mb; e += blockDim.x) {
    xa[e] = work[sh.off_x + (size_t)(k - s) * mb + e];
    xc[e] = has_right ? work[sh.off_x + (size_t)(k + s) * mb + e] : 0.0;
    invd_s[e] = work[sh.off_invd + (size_t)k * mb + e];
  }
  __syncthreads();
  // Z^T x: thread (g, j), g = quarter of the rows
  {
    const double* Za = work + bcr_pair(sh, level, u - 1);
    const double* Zc = work + bcr_pair(sh, level, u);
    const int g = threadIdx.x / 64, j0 = threadIdx.x - g * 64;
    for (int j = j0; j < mb; j += 64) {
      double acc = 0.0;
      for (int r = g; r < mb; r += 4) {
        acc = fma(Za[(size_t)r * mb + j], xa[r], acc);
        if (has_right) acc = fma(Zc[(size_t)r * mb + j], xc[r], acc);
      }
      part[g * mb + j] = acc;
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < mb; j += blockDim.x)
    tv[j] = work[sh.off_rhs + (size_t)k * mb + j] - ((part[j] + part[mb + j]) + (part[2 * mb + j] + part[3 * mb + j]));
  __syncthreads();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    for (int j = mb - 1; j >= 0; j--) {
      const double xj = tv[j] * invd_s[j];
      __syncwarp();
      if (lane == 0) tv[j] = xj;
      const double* Fj = F_s + (size_t)j * mb;
      for (int i = lane; i < j; i += 32) tv[i] = fma(-Fj[i], xj, tv[i]);
      __syncwarp();
    }
    for (int e = lane; e < mb; e += 32) work[sh.off_x + (size_t)k * mb + e] = tv[e];
  }
}

// ---- x -> x_p, computeScale (pose part), trial cameras; a failed factorisation fails the solve ----
__global__ void __launch_bounds__(kBcrThreads)
k_bcr_tail(const BAWin* __restrict__ wins, LgState* stt, BcrShape sh, const double* __restrict__ work) {
  extern __shared__ __align__(16) double smem_d[];
  if (!stt->active) return;
  const BAWin& W = wins[0];
  const int n6 = sh.n * 6;
  double* yv = smem_d;
  double* redv = yv + n6;
  const int fail = *reinterpret_cast<const int*>(work + sh.off_fail);
  if (fail) {  // g2o: Cholesky failure => the trial is rejected (ok2 = false)
    for (int i = threadIdx.x; i < n6; i += blockDim.x) W.xp[i] = 0.0;
    if (threadIdx.x == 0) { stt->ok2 = 0; stt->scale_pose = 0.0; }
    return;
  }
  for (int i = threadIdx.x; i < n6; i += blockDim.x) yv[i] = work[sh.off_x + i];  // super-blocks are contiguous in x
  __syncthreads();
  lg_solve_tail(W, stt, yv, redv, stt->lambda);
}

size_t chol_smem(int mb) { return ((size_t)mb * (mb + 1) + 2 * mb) * sizeof(double); }
size_t trsm_smem(int mb) { return ((size_t)mb * mb + mb) * sizeof(double); }
size_t update_smem(int mb) { return ((size_t)2 * mb * (kBcrTile + 1) + mb) * sizeof(double); }
size_t back_smem(int mb) { return ((size_t)mb * mb + 8 * mb) * sizeof(double); }
size_t tail_smem(int n) { return ((size_t)n * 6 + 64) * sizeof(double); }

}  // namespace

// Shape of the reduction for n free cameras with block half-bandwidth bw (band row stride M = bw + 1).
// Returns false when the sequential band solve is the better choice (short systems).
bool bcr_shape(int n, int bw, int force, BcrShape* out) {
  BcrShape sh = {};
  const int m0 = bw > 1 ? bw : 1;
  int K0 = (n + m0 - 1) / m0;
  int L = 0;
  while ((1 << L) < K0) L++;
  int m = m0;
  if (L >= 1) {  // one level less for at most ~12 % larger super-blocks
    const int m1 = (n + (1 << (L - 1)) - 1) / (1 << (L - 1));
    if (m1 <= m0 + (m0 / 8 > 1 ? m0 / 8 : 1)) { m = m1; L--; }
  }
  const int K = (n + m - 1) / m;
  L = 0;
  while ((1 << L) < K) L++;
  sh.n = n; sh.m = m; sh.mb = 6 * m; sh.K = K; sh.L = L; sh.M = bw + 1;
  if (sh.mb > kBcrMaxMb || L > kBcrMaxLevels - 1 || K < 2) return false;
  if (!force && n < 160) return false;
  size_t pairs = 0;
  for (int l = 0; l <= L; l++) {
    sh.coff[l] = pairs;
    const int s = 1 << l, n_act = (K + s - 1) / s;
    pairs += n_act > 1 ? n_act - 1 : 0;
  }
  const size_t bb = (size_t)sh.mb * sh.mb;
  size_t off = 0;
  sh.off_D = off; off += (size_t)K * bb;
  sh.off_C = off; off += pairs * bb;
  sh.off_rhs = off; off += (size_t)K * sh.mb;
  sh.off_x = off; off += (size_t)K * sh.mb;
  sh.off_invd = off; off += (size_t)K * sh.mb;
  sh.off_fail = off; off += 2;
  sh.total = off;
  *out = sh;
  return true;
}

cudaError_t bcr_prepare(const BcrShape& sh) {
  cudaError_t e = cudaFuncSetAttribute(k_bcr_chol, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chol_smem(sh.mb));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_bcr_trsm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trsm_smem(sh.mb));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_bcr_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)update_smem(sh.mb));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_bcr_back, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)back_smem(sh.mb));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_bcr_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tail_smem(sh.n));
}

// The whole solve as plain launches on the stream; *n_launch receives the number of kernels enqueued.
cudaError_t launch_bcr_solve(const BAWin* w, void* stt_v, const BcrShape& sh, double* work, cudaStream_t st, int* n_launch) {
  LgState* stt = (LgState*)stt_v;
  int nl = 0;
  const int mb = sh.mb, T = (mb + kBcrTile - 1) / kBcrTile;
  k_bcr_assemble<<<2 * sh.K, kBcrThreads, 0, st>>>(w, stt, sh, work);
  nl++;
  for (int l = 0; l < sh.L; l++) {
    const int s = 1 << l, n_act = (sh.K + s - 1) / s;
    const int n_el = n_act / 2, n_surv = (n_act + 1) / 2;
    if (n_el == 0) continue;
    k_bcr_chol<<<n_el, kBcrThreads, chol_smem(mb), st>>>(stt, sh, work, l, 0);
    const int rows_cta = (kBcrThreads / 32) * kBcrRowsPerWarp;
    const int ctas = (2 * mb + rows_cta - 1) / rows_cta;
    k_bcr_trsm<<<dim3(ctas, n_el), kBcrThreads, trsm_smem(mb), st>>>(stt, sh, work, l);
    k_bcr_update<<<n_surv * (T * (T + 1) / 2) + n_el * T * T, kBcrThreads, update_smem(mb), st>>>(stt, sh, work, l);
    nl += 3;
  }
  k_bcr_chol<<<1, kBcrThreads, chol_smem(mb), st>>>(stt, sh, work, 0, 1);
  nl++;
  for (int l = sh.L - 1; l >= 0; l--) {
    const int s = 1 << l, n_act = (sh.K + s - 1) / s;
    const int n_el = n_act / 2;
    if (n_el == 0) continue;
    k_bcr_back<<<n_el, kBcrThreads, back_smem(mb), st>>>(stt, sh, work, l);
    nl++;
  }
  k_bcr_tail<<<1, kBcrThreads, tail_smem(sh.n), st>>>(w, stt, sh, work);
  nl++;
  if (n_launch) *n_launch = nl;
  return cudaGetLastError();
}

}  // namespace urmvo
