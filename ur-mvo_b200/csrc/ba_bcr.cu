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
constexpr int kBcrCholThreads = 736;  // lower 3 x 3 tiles of a 108 x 108 block (666) + 36 right-hand-side tiles
constexpr int kBcrBackThreads = 512;
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
  const int k = blockIdx.x >> 3, which = (blockIdx.x >> 2) & 1, quarter = blockIdx.x & 3;
  const int e_begin = quarter * ((mb * mb + 3) / 4), e_end = min(mb * mb, e_begin + (mb * mb + 3) / 4);
  if (which == 0) {
    double* D = work + sh.off_D + (size_t)k * mb * mb;
#pragma unroll 4
    for (int e = e_begin + threadIdx.x; e < e_end; e += blockDim.x) {
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
    if (quarter == 0)
      for (int e = threadIdx.x; e < mb; e += blockDim.x) b[e] = k * mb + e < n * 6 ? W.bs[k * mb + e] : 0.0;
    if (blockIdx.x == 0 && threadIdx.x == 0) *reinterpret_cast<int*>(work + sh.off_fail) = 0;
  } else if (k + 1 < sh.K) {
    double* C = work + bcr_pair(sh, 0, k);
    const bool transposed = (k & 1) != 0;  // odd left super-block: it is the eliminated one
#pragma unroll 4
    for (int e = e_begin + threadIdx.x; e < e_end; e += blockDim.x) {
      const int r0 = e / mb, c0 = e - r0 * mb;
      const int r = transposed ? c0 : r0, c = transposed ? r0 : c0;  // (r, c) of E = S[k-th rows, (k+1)-th columns]
      const int ia = k * m + r / 6, ic = (k + 1) * m + c / 6;
      double v = 0.0;
      if (ic < n && ic - ia <= bw) v = S[((size_t)ia * M + (ic - ia)) * 36 + (r % 6) * 6 + (c % 6)];
      C[e] = v;
    }
  }
}

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  const size_t src = __cvta_generic_to_global(gsrc);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  const size_t src = __cvta_generic_to_global(gsrc);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n cp.async.wait_all;" ::: "memory");
}
// mb x mb block (contiguous, 16-byte aligned on both sides) -> shared memory, every request in flight at once
__device__ __forceinline__ void bcr_fetch_block(double* dst, const double* __restrict__ src, int mb) {
  for (int c = threadIdx.x; c < mb * mb / 2; c += blockDim.x) cp_async16(dst + 2 * c, src + 2 * c);
}

// Backward substitution L^T x = t by ONE warp: lane l holds t[l], t[l + 32], ...; row j of F (row j of L in
// front of the diagonal) comes from shared memory.  Right-looking: x_j = t_j / L_jj, t_i -= L_ji x_j (i < j).
template <int NSLOT>
__device__ __forceinline__ void bcr_backsolve_warp(const double* __restrict__ F_s, int ldf, const double* __restrict__ invd_s,
                                                   int mb, const double* __restrict__ t_in, double* __restrict__ x_out) {
  const int lane = threadIdx.x & 31;
  double v[NSLOT];
#pragma unroll
  for (int sl = 0; sl < NSLOT; sl++) v[sl] = sl * 32 + lane < mb ? t_in[sl * 32 + lane] : 0.0;
#pragma unroll
  for (int jb = NSLOT - 1; jb >= 0; jb--) {
    const int jn = mb - jb * 32 < 32 ? mb - jb * 32 : 32;
    for (int jj = jn - 1; jj >= 0; jj--) {
      const int j = jb * 32 + jj;
      const double xj = __shfl_sync(0xffffffffu, v[jb], jj) * invd_s[j];
      if (lane == jj) v[jb] = xj;
      const double* Fj = F_s + (size_t)j * ldf;
#pragma unroll
      for (int sl = 0; sl <= jb; sl++) {
        const int i = sl * 32 + lane;
        if (i < j) v[sl] = fma(-Fj[i], xj, v[sl]);
      }
    }
  }
#pragma unroll
  for (int sl = 0; sl < NSLOT; sl++)
    if (sl * 32 + lane < mb) x_out[sl * 32 + lane] = v[sl];
}
__device__ __forceinline__ void bcr_backsolve(const double* F_s, int ldf, const double* invd_s, int mb, const double* t_in,
                                              double* x_out) {
  if (mb <= 64) bcr_backsolve_warp<2>(F_s, ldf, invd_s, mb, t_in, x_out);
  else if (mb <= 96) bcr_backsolve_warp<3>(F_s, ldf, invd_s, mb, t_in, x_out);
  else bcr_backsolve_warp<4>(F_s, ldf, invd_s, mb, t_in, x_out);
}

// ---- Cholesky of one mb x mb block, y = L^-1 b (+ x = L^-T y for the last block) ----
// Input: lower triangle of D (row-major).  Output F[r][c] = L[max(r,c)][min(r,c)] (row j of F holds row j
// of L up to the diagonal and column j of L behind it: both substitutions read rows of F), the
// reciprocal diagonal in invd, y over b.
// Register-resident: thread (ti, tc), ti >= tc, owns the 3 x 3 tile of rows 3 ti.., columns 3 tc.. for the whole
// factorisation; T = mb / 3 more threads own the right-hand side as one more (1 x 3)-tiled row.  Step p:
// (1) the owner of (p, p) factorises it, (2) the owners of column p solve against it and publish their tiles,
// (3) everybody to the right subtracts L_ip L_cp^T.  Two barriers per step, mb / 3 steps.
__global__ void __launch_bounds__(kBcrCholThreads)
k_bcr_chol(const LgState* __restrict__ stt, BcrShape sh, double* __restrict__ work, int level, int last) {
  extern __shared__ __align__(16) double smem_d[];
  if (!stt->active) return;
  const int mb = sh.mb, T = mb / 3, n_tiles = T * (T + 1) / 2;
  const int s = 1 << level;
  const int k = last ? 0 : (2 * blockIdx.x + 1) * s;
  double* panel = smem_d;             // [T + 1][9]: L_ip of the step, entry T: y_p
  double* Lpp = panel + (T + 1) * 9;  // r0 l10 r1 l20 l21 r2 (reciprocal diagonal)
  double* invd_s = Lpp + 8;           // mb
  double* yv = invd_s + mb;           // mb (last block only)
  double* F_s = yv + mb;              // [mb][mb] (last block only)
  int* fail = reinterpret_cast<int*>(work + sh.off_fail);
  double* D = work + sh.off_D + (size_t)k * mb * mb;
  double* b = work + sh.off_rhs + (size_t)k * mb;
  const int tid = threadIdx.x;
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
      for (int c = 0; c < 3; c++) a[r][c] = D[(size_t)(ti * 3 + r) * mb + tc * 3 + c];
  } else if (is_rhs) {
#pragma unroll
    for (int c = 0; c < 3; c++) a[0][c] = b[tc * 3 + c];
  }
  // 3 x 3 Cholesky of the diagonal tile.  The three reciprocal square roots are taken of the leading minors
  // d0, m1 = a11 d0 - a10^2, m2 = det(tile), which do not depend on each other: one rsqrt latency
  // instead of three (1 / L11 = sqrt(d0) / sqrt(m1), 1 / L22 = sqrt(m1) / sqrt(m2)).
  auto factor_diag = [&](int p) {
    const double d0 = a[0][0], a10 = a[1][0], a11 = a[1][1], a20 = a[2][0], a21 = a[2][1], a22 = a[2][2];
    const double m1 = fma(a11, d0, -a10 * a10);
    const double c0 = fma(a11, a22, -a21 * a21), c1 = fma(a10, a22, -a21 * a20), c2 = fma(a10, a21, -a11 * a20);
    const double m2 = fma(d0, c0, fma(-a10, c1, a20 * c2));
    const bool bad = !(d0 > 0.0) || !(m1 > 0.0) || !(m2 > 0.0);
    const double q0 = rsqrt(bad ? 1.0 : d0), q1 = rsqrt(bad ? 1.0 : m1), q2 = rsqrt(bad ? 1.0 : m2);
    const double s0 = d0 * q0, s1 = m1 * q1;  // sqrt(d0), sqrt(m1)
    const double r0 = q0, r1 = q1 * s0, r2 = q2 * s1;
    const double l10 = a10 * r0, l20 = a20 * r0;
    const double l21 = (a21 - l20 * l10) * r1;
    if (bad) *fail = 1;
    Lpp[0] = r0; Lpp[1] = l10; Lpp[2] = r1; Lpp[3] = l20; Lpp[4] = l21; Lpp[5] = r2;
    invd_s[p * 3] = r0; invd_s[p * 3 + 1] = r1; invd_s[p * 3 + 2] = r2;
    a[0][0] = s0; a[1][0] = l10; a[1][1] = s1 * q0; a[2][0] = l20; a[2][1] = l21; a[2][2] = m2 * q2 * q1;
    a[0][1] = l10; a[0][2] = l20; a[1][2] = l21;  // mirrored: the tile is written out whole
  };
  if (is_mat && ti == 0 && tc == 0) factor_diag(0);
  __syncthreads();
  for (int p = 0; p < T; p++) {
    if ((is_mat || is_rhs) && tc == p && ti > p) {
      const double r0 = Lpp[0], l10 = Lpp[1], r1 = Lpp[2], l20 = Lpp[3], l21 = Lpp[4], r2 = Lpp[5];
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
      // the next diagonal tile is complete: its owner factorises it while the others finish their updates
      if (is_mat && ti == p + 1 && tc == p + 1) factor_diag(p + 1);
    }
    __syncthreads();
  }
  // F (mirrored factor), y
  if (is_mat) {
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++) {
        D[(size_t)(ti * 3 + r) * mb + tc * 3 + c] = a[r][c];
        if (ti != tc) D[(size_t)(tc * 3 + c) * mb + ti * 3 + r] = a[r][c];
        if (last) {
          F_s[(ti * 3 + r) * mb + tc * 3 + c] = a[r][c];
          if (ti != tc) F_s[(tc * 3 + c) * mb + ti * 3 + r] = a[r][c];
        }
      }
  } else if (is_rhs) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      if (last) yv[tc * 3 + c] = a[0][c];
      else b[tc * 3 + c] = a[0][c];
    }
  }
  __syncthreads();
  for (int e = tid; e < mb; e += blockDim.x) work[sh.off_invd + (size_t)k * mb + e] = invd_s[e];
  if (last && tid < 32) bcr_backsolve(F_s, mb, invd_s, mb, yv, work + sh.off_x + (size_t)k * mb);  // x = L^-T y
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
  cp_async_wait_all();  // F (fetched by the whole CTA while the rows were loaded)
  __syncthreads();
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

// grid: (CTAs per eliminated super-block, eliminated super-blocks of the level); every warp takes exactly one
// group of kBcrRowsPerWarp rows (mb is a multiple of 6, the groups are cut per coupling block)
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
  bcr_fetch_block(F_s, work + sh.off_D + (size_t)k * mb * mb, mb);
  for (int e = threadIdx.x; e < mb; e += blockDim.x) invd_s[e] = work[sh.off_invd + (size_t)k * mb + e];
  const int warp = threadIdx.x >> 5, n_warp = blockDim.x >> 5;
  const int groups_blk = (mb + kBcrRowsPerWarp - 1) / kBcrRowsPerWarp;
  const int g = blockIdx.x * n_warp + warp;  // group of rows: [0, groups_blk) in Z_a, then Z_c
  const int blk = g >= groups_blk ? 1 : 0;
  const int r0 = (g - blk * groups_blk) * kBcrRowsPerWarp;
  const bool work_here = g < (has_right ? 2 : 1) * groups_blk;
  const int take = work_here ? (mb - r0 < kBcrRowsPerWarp ? mb - r0 : kBcrRowsPerWarp) : 0;
  double* Z = work + bcr_pair(sh, level, blk == 0 || !work_here ? u - 1 : u) + (size_t)(work_here ? r0 : 0) * mb;
  if (mb <= 64) bcr_trsm_rows<2>(F_s, ldf, invd_s, mb, Z, take);
  else if (mb <= 96) bcr_trsm_rows<3>(F_s, ldf, invd_s, mb, Z, take);
  else bcr_trsm_rows<4>(F_s, ldf, invd_s, mb, Z, take);
}

// ---- products of the level: C (-)= A B^T, 48 x 48 tiles, the contracted index staged whole in shared memory ----
// Work items: (a) every surviving super-block a: D_a -= Z Z^T for the eliminated neighbours on both sides (lower
// tiles only), b_a -= Z y; (b) every eliminated super-block with two neighbours: the new coupling -Z_a Z_c^T,
// stored [surviving at level + 1][eliminated at level + 1].
// Tiles are staged row-major with an odd row stride (mb + 1): thread (ty, tx) reads rows 3 ty.. of A (two
// distinct addresses per warp: broadcast) and rows 3 tx.. of B (16 rows, 3 (mb + 1) doubles apart: conflict-free).
__device__ __forceinline__ void bcr_fetch_tile(double* dst, const double* __restrict__ src, int row0, int mb) {
  const int ld = mb + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warp = blockDim.x >> 5;
  for (int i = warp; i < kBcrTile; i += n_warp) {
    const bool in = row0 + i < mb;
    const double* row = src + (size_t)(row0 + i) * mb;
    for (int k = lane; k < mb; k += 32) {
      if (in) cp_async8(dst + i * ld + k, row + k);
      else dst[i * ld + k] = 0.0;
    }
  }
}
__device__ __forceinline__ void bcr_tile_product(const double* __restrict__ As, const double* __restrict__ Bs, int mb,
                                                 int ty, int tx, double (&acc)[3][3]) {
  const int ld = mb + 1;
  const double* a0 = As + (ty * 3) * ld;
  const double* b0 = Bs + (tx * 3) * ld;
#pragma unroll 4
  for (int k = 0; k < mb; k++) {
    double av[3], bv[3];
#pragma unroll
    for (int q = 0; q < 3; q++) { av[q] = a0[q * ld + k]; bv[q] = b0[q * ld + k]; }
#pragma unroll
    for (int p = 0; p < 3; p++)
#pragma unroll
      for (int q = 0; q < 3; q++) acc[p][q] = fma(av[p], bv[q], acc[p][q]);
  }
}

__global__ void __launch_bounds__(kBcrThreads)
k_bcr_update(const LgState* __restrict__ stt, BcrShape sh, double* __restrict__ work, int level) {
  extern __shared__ __align__(16) double smem_d[];
  if (!stt->active) return;
  const int mb = sh.mb, ld = mb + 1;
  const int s = 1 << level, n_act = (sh.K + s - 1) / s;
  const int T = (mb + kBcrTile - 1) / kBcrTile, Tl = T * (T + 1) / 2;
  const int n_surv = (n_act + 1) / 2, n_el = n_act / 2;
  double* As = smem_d;                        // [2 sides][48][ld]
  double* Bs = As + (size_t)2 * kBcrTile * ld;  // [2 sides][48][ld]
  double* ys = Bs + (size_t)2 * kBcrTile * ld;  // [2][mb]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  int b = blockIdx.x;
  if (b < n_surv * Tl) {
    const int f = b / Tl;
    int tl = b - f * Tl;
    int ti = 0;
    while ((ti + 1) * (ti + 2) / 2 <= tl) ti++;
    const int tj = tl - ti * (ti + 1) / 2;
    const int v = 2 * f, a = v * s;
    // side 0: eliminated neighbour on the right (pair v), side 1: on the left (pair v - 1); rows = a in both
    const bool have[2] = {v + 1 < n_act, v >= 2};
    for (int side = 0; side < 2; side++) {
      if (!have[side]) continue;
      const int pair = side == 0 ? v : v - 1;
      const int kel = side == 0 ? a + s : a - s;
      const double* Z = work + bcr_pair(sh, level, pair);
      bcr_fetch_tile(As + (size_t)side * kBcrTile * ld, Z, ti * kBcrTile, mb);
      if (tj != ti) bcr_fetch_tile(Bs + (size_t)side * kBcrTile * ld, Z, tj * kBcrTile, mb);
      if (tj == 0)
        for (int e = threadIdx.x; e < mb; e += blockDim.x) ys[side * mb + e] = work[sh.off_rhs + (size_t)kel * mb + e];
    }
    cp_async_wait_all();
    __syncthreads();
    double acc[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    double racc = 0.0;
    for (int side = 0; side < 2; side++) {
      if (!have[side]) continue;
      const double* At = As + (size_t)side * kBcrTile * ld;
      const double* Bt = tj != ti ? Bs + (size_t)side * kBcrTile * ld : At;
      bcr_tile_product(At, Bt, mb, ty, tx, acc);
      if (tj == 0 && threadIdx.x < kBcrTile)
        for (int k = 0; k < mb; k++) racc = fma(At[threadIdx.x * ld + k], ys[side * mb + k], racc);
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
  bcr_fetch_tile(As, left_survives ? Za : Zc, ti * kBcrTile, mb);  // rows of the new block
  bcr_fetch_tile(Bs, left_survives ? Zc : Za, tj * kBcrTile, mb);  // columns of the new block
  cp_async_wait_all();
  __syncthreads();
  double acc[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  bcr_tile_product(As, Bs, mb, ty, tx, acc);
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
__global__ void __launch_bounds__(kBcrBackThreads)
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
  double* part = xc + mb;                     // [8][mb]
  double* invd_s = part + 8 * mb;             // mb
  bcr_fetch_block(F_s, work + sh.off_D + (size_t)k * mb * mb, mb);
  for (int e = threadIdx.x; e < mb; e += blockDim.x) {
    xa[e] = work[sh.off_x + (size_t)(k - s) * mb + e];
    xc[e] = has_right ? work[sh.off_x + (size_t)(k + s) * mb + e] : 0.0;
    invd_s[e] = work[sh.off_invd + (size_t)k * mb + e];
  }
  __syncthreads();
  // Z^T x: thread (g, j) sums every 8th row, 8 loads in flight
  {
    const double* Za = work + bcr_pair(sh, level, u - 1);
    const double* Zc = work + bcr_pair(sh, level, has_right ? u : u - 1);
    const int g = threadIdx.x >> 7, j = threadIdx.x & 127;  // 4 groups of 128 columns at 512 threads
    if (j < mb) {
      double acc0 = 0.0, acc1 = 0.0;
#pragma unroll 4
      for (int r = g; r < mb; r += 4) {
        acc0 = fma(Za[(size_t)r * mb + j], xa[r], acc0);
        acc1 = fma(Zc[(size_t)r * mb + j], xc[r], acc1);  // xc = 0 without a right neighbour
      }
      part[g * mb + j] = acc0 + acc1;
    }
  }
  cp_async_wait_all();
  __syncthreads();
  for (int j = threadIdx.x; j < mb; j += blockDim.x)
    tv[j] = work[sh.off_rhs + (size_t)k * mb + j] - ((part[j] + part[mb + j]) + (part[2 * mb + j] + part[3 * mb + j]));
  __syncthreads();
  if (threadIdx.x < 32) bcr_backsolve(F_s, mb, invd_s, mb, tv, work + sh.off_x + (size_t)k * mb);
}

// ---- x -> x_p, computeScale (pose part) by CTA 0; trial cameras by the other CTAs ----
// A failed factorisation fails the solve (g2o: Cholesky failure => the trial is rejected, ok2 = false).
__global__ void __launch_bounds__(kBcrThreads)
k_bcr_tail(const BAWin* __restrict__ wins, LgState* stt, BcrShape sh, const double* __restrict__ work) {
  __shared__ double redv[32];
  if (!stt->active) return;
  const BAWin& W = wins[0];
  const int n6 = sh.n * 6;
  const int fail = *reinterpret_cast<const int*>(work + sh.off_fail);
  const double* x = work + sh.off_x;  // super-blocks are contiguous in x
  if (blockIdx.x == 0) {
    if (fail) {
      for (int i = threadIdx.x; i < n6; i += blockDim.x) W.xp[i] = 0.0;
      if (threadIdx.x == 0) { stt->ok2 = 0; stt->scale_pose = 0.0; }
      return;
    }
    lg_tail_scale(W, stt, x, redv, stt->lambda);
  } else if (!fail) {
    lg_tail_cameras(W, stt, x, (blockIdx.x - 1) * blockDim.x + threadIdx.x, (gridDim.x - 1) * blockDim.x);
  }
}

size_t chol_smem(int mb, bool last) { return ((size_t)(mb / 3 + 1) * 9 + 8 + 2 * mb + (last ? (size_t)mb * mb : 0)) * sizeof(double); }
size_t trsm_smem(int mb) { return ((size_t)mb * mb + mb) * sizeof(double); }
size_t update_smem(int mb) { return ((size_t)4 * kBcrTile * (mb + 1) + 2 * mb) * sizeof(double); }
size_t back_smem(int mb) { return ((size_t)mb * mb + 12 * mb) * sizeof(double); }
int chol_threads(int mb) { const int T = mb / 3; return ((T * (T + 1) / 2 + T + 31) / 32) * 32; }

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
  cudaError_t e = cudaFuncSetAttribute(k_bcr_chol, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chol_smem(sh.mb, true));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_bcr_trsm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trsm_smem(sh.mb));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_bcr_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)update_smem(sh.mb));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_bcr_back, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)back_smem(sh.mb));
  return e;
}

// The whole solve as plain launches on the stream; *n_launch receives the number of kernels enqueued.
cudaError_t launch_bcr_solve(const BAWin* w, void* stt_v, const BcrShape& sh, double* work, cudaStream_t st, int* n_launch) {
  LgState* stt = (LgState*)stt_v;
  int nl = 0;
  const int mb = sh.mb, T = (mb + kBcrTile - 1) / kBcrTile;
  k_bcr_assemble<<<8 * sh.K, kBcrThreads, 0, st>>>(w, stt, sh, work);
  nl++;
  for (int l = 0; l < sh.L; l++) {
    const int s = 1 << l, n_act = (sh.K + s - 1) / s;
    const int n_el = n_act / 2, n_surv = (n_act + 1) / 2;
    if (n_el == 0) continue;
    k_bcr_chol<<<n_el, chol_threads(mb), chol_smem(mb, false), st>>>(stt, sh, work, l, 0);
    const int groups = 2 * ((mb + kBcrRowsPerWarp - 1) / kBcrRowsPerWarp), warps_cta = kBcrThreads / 32;
    const int ctas = (groups + warps_cta - 1) / warps_cta;
    k_bcr_trsm<<<dim3(ctas, n_el), kBcrThreads, trsm_smem(mb), st>>>(stt, sh, work, l);
    k_bcr_update<<<n_surv * (T * (T + 1) / 2) + n_el * T * T, kBcrThreads, update_smem(mb), st>>>(stt, sh, work, l);
    nl += 3;
  }
  k_bcr_chol<<<1, chol_threads(mb), chol_smem(mb, true), st>>>(stt, sh, work, 0, 1);
  nl++;
  for (int l = sh.L - 1; l >= 0; l--) {
    const int s = 1 << l, n_act = (sh.K + s - 1) / s;
    const int n_el = n_act / 2;
    if (n_el == 0) continue;
    k_bcr_back<<<n_el, kBcrBackThreads, back_smem(mb), st>>>(stt, sh, work, l);
    nl++;
  }
  k_bcr_tail<<<1 + (sh.n + 2 + kBcrThreads - 1) / kBcrThreads * 4, kBcrThreads, 0, st>>>(w, stt, sh, work);
  nl++;
  if (n_launch) *n_launch = nl;
  return cudaGetLastError();
}

}  // namespace urmvo
