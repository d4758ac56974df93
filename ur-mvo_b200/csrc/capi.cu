// csrc/capi.cu — the C ABI declared in include/urmvo_b200.h: contexts, device-resident plans,
// host-buffer entry points.  Host code here only flattens / validates inputs, builds the sparsity
// structure of the reduced camera system, moves bytes and takes the final scalar accept/reject
// decisions of EpipolarGeometry::reconstruct; all arithmetic of the path runs in the kernels.
// There is no CPU fallback: every entry point needs a live sm_100 device.

#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/urmvo_b200.h"
#include "ba_types.h"
#include "capi_internal.h"
#include "kernels.h"

using namespace urmvo;

namespace {

thread_local std::string g_err;

inline int fail(int code, const std::string& msg) { return urmvo::set_error(code, msg); }

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// NCCL is loaded lazily with dlopen so that the single-GPU path has no NCCL dependency; inside a
// torch process this resolves to the libnccl.so.2 torch already loaded.
struct NcclApi {
  void* handle = nullptr;
  struct UniqueId { char internal[128]; };
  int (*GetUniqueId)(UniqueId*) = nullptr;
  int (*CommInitRank)(void**, int, UniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  static constexpr int kFloat64 = 8, kSum = 0, kMax = 2;
  bool load(std::string& err) {
    if (handle) return true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
    }
    if (!handle) { err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return false; }
    GetUniqueId = (decltype(GetUniqueId))dlsym(handle, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(handle, "ncclCommInitRank");
    AllReduce = (decltype(AllReduce))dlsym(handle, "ncclAllReduce");
    CommDestroy = (decltype(CommDestroy))dlsym(handle, "ncclCommDestroy");
    GetErrorString = (decltype(GetErrorString))dlsym(handle, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy || !GetErrorString) {
      err = "libnccl.so.2 lacks an expected symbol";
      return false;
    }
    return true;
  }
};
NcclApi g_nccl;

// Bump allocator over one device allocation (and a mirrored pinned host staging area).
struct Arena {
  size_t size = 0;
  size_t off = 0;
  template <class T>
  size_t take(size_t n) {
    size_t o = off;
    off = align_up(off + n * sizeof(T));
    return o;
  }
};

}  // namespace


int urmvo::host_threads() {
  static const int n = [] {
    const char* e = std::getenv("URMVO_B200_HOST_THREADS");
    int v = e ? std::atoi(e) : 0;
    if (v <= 0) v = (int)std::thread::hardware_concurrency();
    return v > 0 ? v : 1;
  }();
  return n;
}

int urmvo::set_error(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

extern "C" int urmvo_version(void) { return 100; }
extern "C" const char* urmvo_last_error(void) { return g_err.c_str(); }

extern "C" int urmvo_create(urmvo_ctx** out, int device) {
  if (!out) return fail(URMVO_ERR_ARG, "urmvo_create: null out pointer");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(URMVO_ERR_NO_DEVICE, std::string("urmvo_create: no CUDA device (") +
                                         (e != cudaSuccess ? cudaGetErrorString(e) : "count 0") +
                                         "); this library has no CPU fallback");
  if (device < 0 || device >= n) return fail(URMVO_ERR_ARG, "urmvo_create: device index out of range");
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(URMVO_ERR_NO_DEVICE, "urmvo_create: device is sm_" + std::to_string(prop.major) +
                                         std::to_string(prop.minor) + ", kernels are built for sm_100a only");
  CU_TRY(cudaSetDevice(device));
  urmvo_ctx* c = new urmvo_ctx();
  c->device = device;
  c->n_sm = prop.multiProcessorCount;
  e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete c;
    return fail(URMVO_ERR_CUDA, std::string("cudaStreamCreateWithFlags: ") + cudaGetErrorString(e));
  }
  *out = c;
  return URMVO_OK;
}

extern "C" void urmvo_destroy(urmvo_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->pinned) cudaFreeHost(c->pinned);
  if (c->ws_dev) cudaFree(c->ws_dev);
  delete c;
}

extern "C" void* urmvo_stream(urmvo_ctx* c) { return c ? (void*)c->stream : nullptr; }
extern "C" int urmvo_sync(urmvo_ctx* c) {
  if (!c) return fail(URMVO_ERR_ARG, "urmvo_sync: null context");
  CU_TRY(cudaStreamSynchronize(c->stream));
  return URMVO_OK;
}
extern "C" int64_t urmvo_launch_count(urmvo_ctx* c) { return c ? c->launches : 0; }

// =====================================================================================  BA

// A window taken from the device-resident map (urmvo_map_local_ba): the values of the plan's pose / point / uv
// inputs are gathered on the device from the map's slot-addressed arrays; only the slot lists are uploaded.
struct MapGather {
  const double* d_kf;   // device: cap_kf * 7 (T_wc)
  const double* d_pt;   // device: cap_pt * 3
  const double* d_uv;   // device: cap_obs * 2
  const int* kf_slot;   // host: Nc
  const int* pt_slot;   // host: Np
  const int* obs_slot;  // host: No
};

struct urmvo_ba_plan {
  urmvo_ctx* ctx = nullptr;
  int B = 0;
  int total_c = 0, total_p = 0, total_o = 0;
  int kmax = 1;
  int work_stride = 0, ints_per_warp = 0;
  int batch_mode = -1;  // common accumulation mode of all windows or -1
  int cluster_size = 1, threads = 256, n_clusters = 0;
  bool use_grid = false;
  int grid_blocks = 0;
  BARun run{};
  unsigned char* dev = nullptr;  // one allocation
  size_t dev_bytes = 0;
  // offsets of the pieces the host reads back
  size_t off_wins = 0, off_pose_out = 0, off_pts_out = 0, off_inlier = 0, off_stats = 0;
  // observation permutation (point-major sort) when the caller's order was not sorted
  std::vector<int> perm;  // sorted position -> caller index (empty = identity)
  // point-sharded mode: [scal(8) | S | pad | b_s | b_p | hdiag] is one contiguous all-reduce buffer
  bool borrowed_dev = false;  // dev is the context's workspace
  bool sharded = false;
  size_t off_scal = 0, off_hdiag = 0;
  size_t n_reduce_main = 0;   // doubles from scal through b_p
  int n6 = 0;
  void* shard_state = nullptr;       // device
  void* shard_state_host = nullptr;  // pinned + mapped mirror written by k_sh_decide
  bool stereo = false;               // 3-row edges present (modes 5 / 6)
  bool cluster_auto = true;          // cluster_size was chosen here: it may be halved if the GPU cannot co-schedule it
  // tile mode (csrc/ba_large.cu): one large problem as plain phase kernels, direct band solve
  bool tile = false;
  int band_m = 0;                    // bw + 1
  int ncf = 0;
  bool use_bcr = false;              // long trajectories: block cyclic reduction (csrc/ba_bcr.cu) instead of the sequential band solve
  BcrShape bcr{};
  size_t off_bcr = 0;
  size_t off_gather = 0;             // map windows: device copies of the slot lists [kf | pt | obs]
  size_t off_cam_free = 0;           // dense free-camera index per camera (-1: fixed)
  int n_solve_launches = 1;
  size_t off_hd = 0;                 // [hdiag | scal] is the all-reduce buffer of the lambda initialisation
  size_t n_reduce_diag = 0;
  std::vector<int> pt_perm;          // device point index -> caller's point index
  // phase times of the last run (ms, CUDA events on the context stream): LIN, all-reduce, solve, BACKSUB
  float phase_ms[4] = {0, 0, 0, 0};
  int n_trial_launches = 0, n_host_syncs = 0;
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
};

namespace {

struct WinHost {
  int Nc, Ncf, Np, No, nblk, kmax;
  bool dup_cam = false;  // some point is observed twice by the same camera
  int acc_mode = 0, acc_len = 0;
  std::vector<int> grp_pt;  // packed modes: point groups with <= 32 observations and <= 32 points
  std::vector<int> pt_start, cam_free, row_ptr, col, lrow_ptr, lcol, lblk;
  // tile mode (acc_mode 4, csrc/ba_large.cu)
  int bw = 0, n_chunk = 0;
  std::vector<int> pt_order;  // new point index -> caller's point index
  std::vector<int> chunk_grp, grp_cbase, chunk_blk, blk_desc;
};

// Builds CSR over points, free-camera indices and the upper-BSR structure of S for one window.
// obs_pt / obs_cam are window-local; `order` (size No) receives the point-major permutation
// (sorted position -> original index), `sorted` tells whether it is the identity.
int build_window(int Nc, const uint8_t* fixed, int Np, int No, const int32_t* obs_cam,
                 const int32_t* obs_pt, WinHost& w, std::vector<int>& order, bool& sorted,
                 const uint8_t* covis = nullptr) {
  w.Nc = Nc; w.Np = Np; w.No = No;
  w.cam_free.assign(Nc, -1);
  w.Ncf = 0;
  for (int c = 0; c < Nc; c++)
    if (!fixed[c]) w.cam_free[c] = w.Ncf++;
  w.pt_start.assign(Np + 1, 0);
  sorted = true;
  bool cams_increasing = true;  // inside every point (what the reference emits): then no duplicates
  {
    int prev_p = -1, prev_c = -1;
    for (int o = 0; o < No; o++) {
      const int p = obs_pt[o], c = obs_cam[o];
      if ((unsigned)p >= (unsigned)Np || (unsigned)c >= (unsigned)Nc)
        return fail(URMVO_ERR_ARG, "local_ba: observation index out of range");
      if (p < prev_p) sorted = false;
      if (p == prev_p && c <= prev_c) cams_increasing = false;
      prev_p = p; prev_c = c;
      w.pt_start[p + 1]++;
    }
  }
  w.kmax = 1;
  for (int l = 0; l < Np; l++) {
    w.kmax = std::max(w.kmax, w.pt_start[l + 1]);
    w.pt_start[l + 1] += w.pt_start[l];
  }
  order.clear();  // empty = identity (already point-major sorted)
  if (!sorted) {
    order.resize(No);
    std::vector<int> fill(w.pt_start.begin(), w.pt_start.end() - 1);
    for (int o = 0; o < No; o++) order[fill[obs_pt[o]]++] = o;  // stable
  }
  w.dup_cam = false;
  if (!sorted || !cams_increasing) {  // a camera seeing the same point twice forces the atomic path
    std::vector<int> seen(Nc, -1);
    for (int l = 0; l < Np && !w.dup_cam; l++)
      for (int s = w.pt_start[l]; s < w.pt_start[l + 1]; s++) {
        const int c = obs_cam[sorted ? s : order[s]];
        if (seen[c] == l) { w.dup_cam = true; break; }
        seen[c] = l;
      }
  }
  // structure of S: block (a,b), a <= b, present iff some point is seen by free cameras a and b
  const int n = w.Ncf;
  std::vector<std::vector<int>> rows(n);
  for (int i = 0; i < n; i++) rows[i].push_back(i);
  if (covis) {  // structure agreed between the ranks of a point-sharded problem (Ncf x Ncf, upper)
    for (int i = 0; i < n; i++) {
      rows[i].clear();
      rows[i].push_back(i);
      for (int j = i + 1; j < n; j++)
        if (covis[(size_t)i * n + j]) rows[i].push_back(j);
    }
  } else if (n <= 64) {
    for (int i = 0; i < n; i++) {  // dense upper triangle: no scan of the observations needed
      rows[i].resize(n - i);
      for (int j = i; j < n; j++) rows[i][j - i] = j;
    }
  } else {
    std::vector<int> cfs;
    for (int l = 0; l < Np; l++) {
      cfs.clear();
      for (int s = w.pt_start[l]; s < w.pt_start[l + 1]; s++) {
        const int cf = w.cam_free[obs_cam[order.empty() ? s : order[s]]];
        if (cf >= 0) cfs.push_back(cf);
      }
      std::sort(cfs.begin(), cfs.end());
      cfs.erase(std::unique(cfs.begin(), cfs.end()), cfs.end());
      for (size_t a = 0; a < cfs.size(); a++) {
        std::vector<int>& r = rows[cfs[a]];
        for (size_t b = a + 1; b < cfs.size(); b++) {
          auto it = std::lower_bound(r.begin(), r.end(), cfs[b]);
          if (it == r.end() || *it != cfs[b]) r.insert(it, cfs[b]);
        }
      }
    }
  }
  w.row_ptr.assign(n + 1, 0);
  for (int i = 0; i < n; i++) w.row_ptr[i + 1] = w.row_ptr[i] + (int)rows[i].size();
  w.nblk = w.row_ptr[n];
  w.col.resize(w.nblk);
  std::vector<int> lcount(n + 1, 0);
  for (int i = 0; i < n; i++)
    for (size_t k = 0; k < rows[i].size(); k++) {
      w.col[w.row_ptr[i] + k] = rows[i][k];
      if (rows[i][k] > i) lcount[rows[i][k] + 1]++;
    }
  w.lrow_ptr.assign(n + 1, 0);
  for (int i = 0; i < n; i++) w.lrow_ptr[i + 1] = w.lrow_ptr[i] + lcount[i + 1];
  w.lcol.resize(w.lrow_ptr[n]);
  w.lblk.resize(w.lrow_ptr[n]);
  std::vector<int> lfill(w.lrow_ptr.begin(), w.lrow_ptr.end() - 1);
  for (int i = 0; i < n; i++)
    for (int e = w.row_ptr[i]; e < w.row_ptr[i + 1]; e++) {
      const int j = w.col[e];
      if (j > i) { w.lcol[lfill[j]] = i; w.lblk[lfill[j]] = e; lfill[j]++; }
    }
  return URMVO_OK;
}


// Tile mode (csrc/ba_large.cu) for ONE large problem.  Renumbers the points by their first free
// camera, keeps every point's observations together (caller's relative order), stores S as a full
// block band and cuts the points into chunks whose blocks fit one CTA (<= 256 blocks inside a window
// of <= 32 consecutive free cameras).  `order` receives the observation permutation (sorted position
// -> caller index).  Returns 1 when the problem does not fit the mode (caller falls back to the
// atomic / PCG path), 0 on success, < 0 on invalid input.
int build_large(int Nc, const uint8_t* fixed, int Np, int No, const int32_t* obs_cam, const int32_t* obs_pt,
                const uint8_t* covis, int n_sm, int max_m, WinHost& w, std::vector<int>& order) {
  w.Nc = Nc; w.Np = Np; w.No = No;
  w.cam_free.assign(Nc, -1);
  w.Ncf = 0;
  for (int c = 0; c < Nc; c++)
    if (!fixed[c]) w.cam_free[c] = w.Ncf++;
  const int n = w.Ncf;
  if (n == 0 || Np == 0 || No == 0) return 1;
  std::vector<int> cnt(Np, 0), kmin(Np, n), kmax_c(Np, -1);
  for (int o = 0; o < No; o++) {
    const int p = obs_pt[o], c = obs_cam[o];
    if ((unsigned)p >= (unsigned)Np || (unsigned)c >= (unsigned)Nc)
      return fail(URMVO_ERR_ARG, "local_ba: observation index out of range");
    cnt[p]++;
    const int cf = w.cam_free[c];
    if (cf >= 0) { kmin[p] = std::min(kmin[p], cf); kmax_c[p] = std::max(kmax_c[p], cf); }
  }
  w.kmax = 1;
  for (int l = 0; l < Np; l++) w.kmax = std::max(w.kmax, cnt[l]);
  if (w.kmax > 32) return 1;
  // points by first free camera (counting sort, stable); points seen by fixed cameras only go last
  w.pt_order.resize(Np);
  {
    std::vector<int> start(n + 2, 0);
    for (int l = 0; l < Np; l++) start[kmin[l] + 1]++;
    for (int k = 0; k <= n; k++) start[k + 1] += start[k];
    for (int l = 0; l < Np; l++) w.pt_order[start[kmin[l]]++] = l;
  }
  std::vector<int> inv(Np);
  for (int q = 0; q < Np; q++) inv[w.pt_order[q]] = q;
  w.pt_start.assign(Np + 1, 0);
  for (int q = 0; q < Np; q++) w.pt_start[q + 1] = w.pt_start[q] + cnt[w.pt_order[q]];
  order.resize(No);
  {
    std::vector<int> fill(w.pt_start.begin(), w.pt_start.end() - 1);
    for (int o = 0; o < No; o++) order[fill[inv[obs_pt[o]]]++] = o;  // stable inside a point
  }
  // a camera that sees the same point twice is not supported by the one-slot-per-(point, camera) table
  {
    std::vector<int> seen(Nc, -1);
    for (int q = 0; q < Np; q++)
      for (int s = w.pt_start[q]; s < w.pt_start[q + 1]; s++) {
        const int c = obs_cam[order[s]];
        if (seen[c] == q) return 1;
        seen[c] = q;
      }
  }
  // block half-bandwidth of S
  int bw = 0;
  if (covis) {
    for (int i = 0; i < n; i++)
      for (int j = i + 1; j < n; j++)
        if (covis[(size_t)i * n + j]) bw = std::max(bw, j - i);
  } else {
    for (int l = 0; l < Np; l++)
      if (kmax_c[l] >= 0) bw = std::max(bw, kmax_c[l] - kmin[l]);
  }
  if (n >= 2) bw = std::max(bw, 1);  // the band solve looks one block column ahead
  if (bw + 1 > max_m) return 1;
  w.bw = bw;
  // uniform row stride bw + 1 (the last rows carry unused blocks): block (i, i + d) at i * (bw + 1) + d,
  // so that the band solve addresses S without an index load
  w.row_ptr.assign(n + 1, 0);
  for (int i = 0; i < n; i++) w.row_ptr[i + 1] = w.row_ptr[i] + bw + 1;
  w.nblk = w.row_ptr[n];
  w.col.resize(w.nblk);
  for (int i = 0; i < n; i++)
    for (int e = w.row_ptr[i]; e < w.row_ptr[i + 1]; e++) w.col[e] = i + (e - w.row_ptr[i]);  // >= n: padding
  w.lrow_ptr.assign(n + 1, 0);  // mirror lists are not used in tile mode
  w.lcol.clear();
  w.lblk.clear();
  // chunks of consecutive (renumbered) points
  auto window_blocks = [&](int c0, int c1) {
    long long b = 0;
    for (int ci = c0; ci <= c1; ci++) b += std::min(c1, ci + bw) - ci + 1;
    return b;
  };
  const int cap = std::max(96, (Np + n_sm - 1) / n_sm);
  w.chunk_grp.assign(1, 0);
  w.chunk_blk.assign(1, 0);
  w.grp_pt.assign(1, 0);
  w.grp_cbase.clear();
  w.blk_desc.clear();
  int q = 0;
  while (q < Np) {
    const int lq = w.pt_order[q];
    int c0 = kmin[lq], c1 = kmax_c[lq];
    const bool no_cam = c1 < 0;
    if (!no_cam && (c1 - c0 + 1 > 32 || window_blocks(c0, c1) > 256)) return 1;
    int q_end = q + 1;
    while (q_end < Np && q_end - q < cap) {
      const int l = w.pt_order[q_end];
      if ((kmax_c[l] < 0) != no_cam) break;
      if (!no_cam) {
        const int n1 = std::max(c1, kmax_c[l]);  // kmin is non-decreasing: c0 stays
        if (n1 - c0 + 1 > 32 || window_blocks(c0, n1) > 256) break;
        c1 = n1;
      }
      q_end++;
    }
    // packed groups of the chunk: <= 32 observations and <= 32 points each
    int obs = 0, pts_in = 0;
    for (int l = q; l < q_end; l++) {
      const int k = w.pt_start[l + 1] - w.pt_start[l];
      if ((obs + k > 32 || pts_in == 32) && pts_in > 0) {
        w.grp_pt.push_back(l);
        w.grp_cbase.push_back(no_cam ? 0 : c0);
        obs = 0; pts_in = 0;
      }
      obs += k;
      pts_in++;
    }
    w.grp_pt.push_back(q_end);
    w.grp_cbase.push_back(no_cam ? 0 : c0);
    w.chunk_grp.push_back((int)w.grp_cbase.size());
    if (!no_cam)
      for (int ci = c0; ci <= c1; ci++)
        for (int cj = ci; cj <= std::min(c1, ci + bw); cj++) {
          w.blk_desc.push_back((ci - c0) | ((cj - c0) << 8));
          w.blk_desc.push_back(w.row_ptr[ci] + (cj - ci));
        }
    w.chunk_blk.push_back((int)w.blk_desc.size() / 2);
    q = q_end;
  }
  w.n_chunk = (int)w.chunk_grp.size() - 1;
  w.acc_mode = 4;
  w.acc_len = 0;
  w.dup_cam = false;
  return 0;
}

}  // namespace

extern "C" void urmvo_ba_plan_destroy(urmvo_ba_plan* p) {
  if (!p) return;
  cudaSetDevice(p->ctx->device);
  if (p->dev && p->borrowed_dev) p->ctx->ws_in_use = false;
  else if (p->dev) cudaFree(p->dev);
  if (p->shard_state) cudaFree(p->shard_state);
  if (p->shard_state_host) cudaFreeHost(p->shard_state_host);
  for (cudaEvent_t e : p->ev) if (e) cudaEventDestroy(e);
  delete p;
}

static int ba_plan_create_impl_(urmvo_ctx* ctx, urmvo_ba_plan** out, int B, const int32_t* cam_off,
                               const int32_t* pt_off, const int32_t* obs_off, const double* poses,
                               const uint8_t* fixed, const double* pts, const double* uv,
                               const int32_t* cam, const int32_t* pt, const double* intr,
                               double chi2_thr, int it0, int it1, const urmvo_ba_options* opts,
                               bool sharded, const uint8_t* covis, bool borrow_ws,
                               const uint8_t* kind, double chi2_thr_stereo, const MapGather* mg, int n_models);

// no exception crosses the C ABI: allocation failures of the host-side flattening become a status
static int ba_plan_create_impl(urmvo_ctx* ctx, urmvo_ba_plan** out, int B, const int32_t* cam_off,
                               const int32_t* pt_off, const int32_t* obs_off, const double* poses,
                               const uint8_t* fixed, const double* pts, const double* uv,
                               const int32_t* cam, const int32_t* pt, const double* intr,
                               double chi2_thr, int it0, int it1, const urmvo_ba_options* opts,
                               bool sharded, const uint8_t* covis, bool borrow_ws = false,
                               const uint8_t* kind = nullptr, double chi2_thr_stereo = 0.0, const MapGather* mg = nullptr,
                               int n_models = 0) {
  try {
    return ba_plan_create_impl_(ctx, out, B, cam_off, pt_off, obs_off, poses, fixed, pts, uv, cam, pt, intr, chi2_thr, it0,
                                it1, opts, sharded, covis, borrow_ws, kind, chi2_thr_stereo, mg, n_models);
  } catch (const std::exception& e) {
    if (ctx) ctx->ws_in_use = false;
    if (out) *out = nullptr;
    return fail(URMVO_ERR_ARG, std::string("ba_plan_create: ") + e.what());
  }
}

static int ba_plan_create_impl_(urmvo_ctx* ctx, urmvo_ba_plan** out, int B, const int32_t* cam_off,
                               const int32_t* pt_off, const int32_t* obs_off, const double* poses,
                               const uint8_t* fixed, const double* pts, const double* uv,
                               const int32_t* cam, const int32_t* pt, const double* intr,
                               double chi2_thr, int it0, int it1, const urmvo_ba_options* opts,
                               bool sharded, const uint8_t* covis, bool borrow_ws,
                               const uint8_t* kind, double chi2_thr_stereo, const MapGather* mg, int n_models) {
  // kind != NULL: stereo-capable window(s): uv carries 3 values per observation (u, v, u_right), intr 5 values
  // (fx, fy, cx, cy, bf), kind[o] = 1 marks an EdgeStereoSE3ProjectXYZ (reference src/g2o_optimization.cc:96-118).
  // n_models > 0: intr is a table of n_models such rows and kind[o] = stereo bit | camera model << 1 — the reference
  // reads camera_list[mpc->id_camera] per constraint (:86-89, :106-113)
  const bool stereo = kind != nullptr;
  if (n_models < 0 || n_models > 128 || (n_models > 0 && !stereo)) return fail(URMVO_ERR_ARG, "ba_plan_create: 1..128 camera models, with per-observation kind bytes");
  const int uvs = stereo ? 3 : 2;
  if (!ctx || !out) return fail(URMVO_ERR_ARG, "ba_plan_create: null context / out");
  if (stereo && (sharded || !(chi2_thr_stereo > 0))) return fail(URMVO_ERR_ARG, "ba_plan_create: stereo edges need a positive threshold and are not supported by the point-sharded solve");
  *out = nullptr;
  if (B <= 0 || !cam_off || !pt_off || !obs_off || !fixed || !cam || !pt || !intr || (!mg && (!poses || !pts || !uv)))
    return fail(URMVO_ERR_ARG, "ba_plan_create: null or empty input");
  if (mg && (B != 1 || stereo || sharded || obs_off[1] - obs_off[0] >= 100000))
    return fail(URMVO_ERR_UNSUPPORTED, "map window: one mono window of fewer than 100000 observations");
  if (it0 < 0 || it1 < 0 || !(chi2_thr > 0)) return fail(URMVO_ERR_ARG, "ba_plan_create: bad iteration counts / threshold");
  CU_TRY(cudaSetDevice(ctx->device));
  std::vector<WinHost> wh(B);
  urmvo_ba_plan* p = new urmvo_ba_plan();
  p->ctx = ctx;
  p->B = B;
  p->total_c = cam_off[B]; p->total_p = pt_off[B]; p->total_o = obs_off[B];
  bool all_sorted = true;
  std::vector<std::vector<int>> orders(B);
  std::vector<int> rcs(B, URMVO_OK);
  std::vector<char> sorted_w(B, 1);
  std::vector<std::string> errs(B);
  for (int w = 0; w < B; w++) {
    const int Nc = cam_off[w + 1] - cam_off[w], Np = pt_off[w + 1] - pt_off[w], No = obs_off[w + 1] - obs_off[w];
    if (Nc <= 0 || Np < 0 || No < 0) { delete p; return fail(URMVO_ERR_ARG, "ba_plan_create: bad window offsets"); }
  }
  const int large_mode = opts ? opts->large_mode : 0;
  bool tile = false;
  if (B == 1 && large_mode != 1 && !stereo && (sharded || obs_off[1] - obs_off[0] >= 100000)) {
    const int rc = build_large(cam_off[1] - cam_off[0], fixed + cam_off[0], pt_off[1] - pt_off[0], obs_off[1] - obs_off[0],
                               cam + obs_off[0], pt + obs_off[0], covis, ctx->n_sm, lg_band_max_m(), wh[0], orders[0]);
    if (rc < 0) { delete p; return rc; }
    tile = rc == 0;
    if (tile) sorted_w[0] = 0;
    else wh[0] = WinHost();
  }
  if (large_mode == 2 && !tile) { delete p; return fail(URMVO_ERR_UNSUPPORTED, "ba options: large_mode 2 (tile mode) does not fit this problem"); }
  if (!tile) {  // windows are independent: flatten them on all host threads
    const int nt = std::max(1, std::min({urmvo::host_threads(), 16, B / 4}));
    std::atomic<int> next(0);
    auto work = [&]() {
      for (int w = next++; w < B; w = next++) {
        const int Nc = cam_off[w + 1] - cam_off[w], Np = pt_off[w + 1] - pt_off[w], No = obs_off[w + 1] - obs_off[w];
        bool sorted = true;
        rcs[w] = build_window(Nc, fixed + cam_off[w], Np, No, cam + obs_off[w], pt + obs_off[w], wh[w], orders[w], sorted, covis);
        if (rcs[w] != URMVO_OK) errs[w] = g_err;  // thread-local message of the worker
        sorted_w[w] = sorted ? 1 : 0;
      }
    };
    if (nt == 1) {
      work();
    } else {
      std::vector<std::thread> th;
      for (int t = 0; t < nt; t++) th.emplace_back(work);
      for (auto& t : th) t.join();
    }
  }
  int max_ncf = 0;
  long long sum_blk = 0, sum_ncf = 0;
  for (int w = 0; w < B; w++) {
    if (rcs[w] != URMVO_OK) { delete p; return fail(rcs[w], errs[w]); }
    all_sorted = all_sorted && sorted_w[w];
    p->kmax = std::max(p->kmax, wh[w].kmax);
    max_ncf = std::max(max_ncf, wh[w].Ncf);
    sum_blk += wh[w].nblk;
    sum_ncf += wh[w].Ncf;
  }
  std::vector<int> perm;
  if (!all_sorted) {
    perm.resize(p->total_o);
    for (int w = 0; w < B; w++) {
      const int No = obs_off[w + 1] - obs_off[w];
      for (int o = 0; o < No; o++) perm[obs_off[w] + o] = obs_off[w] + (orders[w].empty() ? o : orders[w][o]);
    }
  }
  if (!all_sorted) p->perm = perm;
  // ---- launch shape
  p->threads = (opts && opts->threads > 0) ? opts->threads : 256;
  if (p->threads % 32 != 0 || p->threads < 64 || p->threads > 256) { delete p; return fail(URMVO_ERR_ARG, "ba options: threads must be 64..256 and a multiple of 32"); }
  const int max_np = [&] { int m = 0; for (auto& w : wh) m = std::max(m, w.Np); return m; }();
  const long long max_no = [&] { long long m = 0; for (auto& w : wh) m = std::max<long long>(m, w.No); return m; }();
  p->use_grid = (B == 1 && max_no >= 100000) || sharded;
  p->sharded = sharded;
  // accumulation mode per window (BAWin::acc_mode): packed groups + register-resident blocks when
  // every point has <= 32 observations and S has <= 64 blocks; shared-memory RMW copies up to 16
  // free cameras; global fp64 atomics + BSR PCG otherwise
  const int force = opts ? opts->force_atomic : 0;  // 1: mode 0, 2: mode <= 1
  for (auto& w : wh) {
    if (tile) break;
    w.acc_len = w.nblk * 36 + w.Ncf * 12;
    w.acc_mode = 0;
    if (!p->use_grid && force != 1 && !w.dup_cam && w.Ncf <= 16) {
      w.acc_mode = 1;
      if (force != 2 && !stereo && w.kmax <= 32 && w.nblk <= 64) w.acc_mode = w.nblk <= 32 ? 2 : 3;
    }
    if (stereo) w.acc_mode = w.acc_mode == 1 ? 5 : 6;  // the 3-row phases exist for modes 1 and 0 only
    if (w.acc_mode >= 2) {
      w.grp_pt.clear();
      w.grp_pt.push_back(0);
      int obs = 0, pts_in = 0;
      for (int l = 0; l < w.Np; l++) {
        const int k = w.pt_start[l + 1] - w.pt_start[l];
        if (obs + k > 32 || pts_in == 32) { w.grp_pt.push_back(l); obs = 0; pts_in = 0; }
        obs += k;
        pts_in++;
      }
      w.grp_pt.push_back(w.Np);
    }
  }
  const size_t smem_budget = 200 * 1024;
  for (; !tile;) {
    const int nw = p->threads / 32;
    int stride = 0, pcg_d = 0, ints = p->kmax;
    for (auto& w : wh) {
      int need = 0;
      if (w.acc_mode == 0 || w.acc_mode == 6) need = ba_stage_doubles(p->kmax, stereo) + ba_tile_doubles();
      else if (w.acc_mode == 1 || w.acc_mode == 5) need = ba_stage_doubles(p->kmax, stereo) + w.acc_len;
      else { need = std::max(ba_pack_doubles(), w.acc_len); ints = std::max(ints, 128); }
      stride = std::max(stride, need);
      if (w.acc_mode >= 1 && w.acc_mode != 6) {
        pcg_d = std::max(pcg_d, pcg_dense_doubles(w.Ncf, p->threads));
      }
    }
    stride = std::max(stride, (pcg_d + nw - 1) / nw);
    stride = (stride + 1) & ~1;  // keep the int area 16-byte aligned
    p->work_stride = stride;
    p->ints_per_warp = ints;
    p->batch_mode = wh[0].acc_mode;
    for (auto& w : wh) if (w.acc_mode != p->batch_mode) p->batch_mode = -1;
    if (ba_smem_bytes(p->threads, stride, ints) <= smem_budget) break;
    if (p->threads > 64) { p->threads -= 32; continue; }
    // drop the most expensive small-window mode to the next cheaper one and retry
    bool changed = false;
    int worst = 0;
    for (auto& w : wh) if (w.acc_mode == 1 || w.acc_mode == 5) worst = std::max(worst, w.acc_len);
    for (auto& w : wh)
      if ((w.acc_mode == 1 || w.acc_mode == 5) && w.acc_len == worst) { w.acc_mode = w.acc_mode == 5 ? 6 : 0; changed = true; }
    if (!changed) {
      delete p;
      return fail(URMVO_ERR_UNSUPPORTED, "local_ba: a point has too many observations for the per-warp staging area");
    }
    p->threads = (opts && opts->threads > 0) ? opts->threads : 256;
  }
  int cs = (opts && opts->cluster_size > 0) ? opts->cluster_size : 0;
  if (cs == 0) {
    // one CTA per SM is resident (register budget): give every window as many SMs as the batch
    // leaves free, but at least ~128 points per CTA; a full batch (B >= #SMs) runs one window per SM,
    // which avoids all cluster barriers and the idle time during the single-CTA PCG
    const int warps = p->threads / 32;
    cs = 1;
    while (cs < 16 && cs * 2 * B <= ctx->n_sm && max_np > cs * warps * 16) cs *= 2;
  }
  if (cs != 1 && cs != 2 && cs != 4 && cs != 8 && cs != 16) { delete p; return fail(URMVO_ERR_ARG, "ba options: cluster_size must be 1,2,4,8 or 16"); }
  // the BSR PCG keeps at most 8 block rows per warp in registers: a window with more free cameras
  // than the warps of its cluster can hold runs on the whole grid instead
  if (!p->use_grid && max_ncf > (p->threads / 32) * cs * 8) {
    if (B != 1) { delete p; return fail(URMVO_ERR_UNSUPPORTED, "local_ba_batch: a window has too many free cameras for a cluster; solve it alone"); }
    p->use_grid = true;
  }
  p->cluster_size = cs;
  p->n_clusters = B;
  p->cluster_auto = !(opts && opts->cluster_size > 0);
  int nblk_scope = cs;
  p->tile = tile;
  if (tile) {
    p->use_grid = true;
    p->grid_blocks = ctx->n_sm;
    p->band_m = wh[0].bw + 1;
    p->ncf = wh[0].Ncf;
    p->pt_perm = wh[0].pt_order;
    nblk_scope = 4 * ctx->n_sm;
    const int band_solver = opts ? opts->band_solver : 0;
    p->use_bcr = band_solver != 1 && bcr_shape(p->ncf, wh[0].bw, band_solver == 2, &p->bcr);
    if (p->use_bcr && bcr_prepare(p->bcr) != cudaSuccess) { delete p; return fail(URMVO_ERR_CUDA, "ba_plan_create: cudaFuncSetAttribute failed (cyclic reduction)"); }
    // the sequential band solve keeps the whole right-hand side in shared memory; the cyclic reduction does not
    if (!p->use_bcr && band_smem_bytes(p->band_m, p->ncf) > 220 * 1024) { delete p; return fail(URMVO_ERR_UNSUPPORTED, "local_ba: too many free cameras for the direct band solve"); }
    if (lg_prepare(p->band_m, p->use_bcr ? 0 : p->ncf) != cudaSuccess) { delete p; return fail(URMVO_ERR_CUDA, "ba_plan_create: cudaFuncSetAttribute failed"); }
  } else if (p->use_grid) {
    p->grid_blocks = sharded ? shard_grid_capacity(p->threads, p->kmax) : ba_grid_capacity(p->threads, p->kmax, stereo ? 1 : 0);
    if (p->grid_blocks <= 0) { delete p; return fail(URMVO_ERR_CUDA, "ba_plan_create: occupancy query failed"); }
    nblk_scope = p->grid_blocks;
  }
  p->run.chi2_thr = chi2_thr;
  p->run.delta = (double)(float)std::sqrt(chi2_thr);  // const float thHuberMonoPoint = sqrt(cfg.mono_point)
  p->stereo = stereo;
  p->run.chi2_thr_s = stereo ? chi2_thr_stereo : chi2_thr;
  p->run.delta_s = (double)(float)std::sqrt(p->run.chi2_thr_s);  // const float thHuberStereoPoint = sqrt(cfg.stereo_point)
  p->run.pcg_tol = (opts && opts->pcg_tol > 0) ? opts->pcg_tol : 1e-10;
  p->run.dense_solver = opts ? std::min(std::max(opts->dense_solver, 0), 2) : 0;
  p->run.pcg_max_iter = (opts && opts->pcg_max_iter > 0) ? opts->pcg_max_iter : std::min(1000, std::max(60, 12 * max_ncf));
  p->run.it0 = it0; p->run.it1 = it1; p->run.n_win = B;
  p->run.timing_stats = nullptr;

  // ---- device layout
  Arena A;
  const size_t TC = p->total_c, TP = p->total_p, TO = p->total_o;
  const size_t o_pose_in = A.take<double>(TC * 7), o_pts_in = A.take<double>(TP * 3), o_uv = A.take<double>(TO * 2);
  size_t sum_grp = 0;
  for (auto& w : wh) sum_grp += w.grp_pt.size();
  const size_t o_ocam = A.take<int>(TO), o_opt = A.take<int>(TO);
  const size_t o_ur = A.take<double>(stereo ? TO : 0), o_okind = A.take<uint8_t>(stereo ? TO : 0);
  const int n_tab = std::max(n_models, 1);
  const size_t o_tab = A.take<double>(stereo ? (size_t)5 * n_tab : 0);
  const size_t o_pt_start = A.take<int>(TP + B), o_cam_free = A.take<int>(TC), o_grp = A.take<int>(sum_grp + 1);
  p->off_cam_free = o_cam_free;
  const size_t o_row_ptr = A.take<int>(sum_ncf + B), o_col = A.take<int>(sum_blk);
  const size_t o_lrow_ptr = A.take<int>(sum_ncf + B), o_lcol = A.take<int>(sum_blk), o_lblk = A.take<int>(sum_blk);
  // tile mode: chunk / block-ownership tables
  const WinHost& w0 = wh[0];
  const size_t o_chunk_grp = A.take<int>(tile ? w0.chunk_grp.size() : 0), o_grp_cbase = A.take<int>(tile ? w0.grp_cbase.size() : 0);
  const size_t o_chunk_blk = A.take<int>(tile ? w0.chunk_blk.size() : 0), o_blk_desc = A.take<int>(tile ? w0.blk_desc.size() : 0);
  p->off_gather = A.take<int>(mg ? (size_t)p->total_c + p->total_p + p->total_o : 0);  // map windows: slot lists, part of the staged upload
  p->off_wins = A.take<BAWin>(B);
  const size_t upload_end = A.off;  // everything above is filled from the host
  size_t o_cam[2], o_camRt[2], o_pts[2];
  for (int k = 0; k < 2; k++) { o_cam[k] = A.take<double>(TC * 7); o_camRt[k] = A.take<double>(TC * 12); o_pts[k] = A.take<double>(TP * 3); }
  const size_t o_level = A.take<uint8_t>(TO);
  size_t sum_rec = 0;
  for (auto& w : wh) if (w.acc_mode >= 2) sum_rec += (w.grp_pt.size() - 1) * 32;
  const size_t o_rec = A.take<ObsRec>(sum_rec);
  const size_t o_hd = A.take<double>(tile ? (size_t)w0.Ncf * 6 : 0);  // tile mode: [hdiag | scal] is reduced at the lambda initialisation
  const size_t o_scal = A.take<double>(32);  // sharded: head of the contiguous all-reduce buffer
  const size_t o_S = A.take<double>((size_t)sum_blk * 36);
  const size_t o_vec = A.take<double>((size_t)(sum_ncf + B) * 6 * 8);  // bs bp hdiag xp r z p Ap (indexed by c_ncf, which counts Ncf+1 per window)
  if (tile) {
    p->off_hd = o_hd;
    p->n_reduce_diag = (o_scal - o_hd) / sizeof(double) + 32;
  }
  if (sharded || tile) {
    p->off_scal = o_scal;
    p->n6 = wh[0].Ncf * 6;
    p->n_reduce_main = (o_vec - o_scal) / sizeof(double) + (size_t)2 * p->n6;  // scal | S | pad | b_s | b_p
    p->off_hdiag = o_vec + (size_t)2 * p->n6 * sizeof(double);
  }
  const size_t o_Minv = A.take<double>((size_t)(sum_ncf + B) * 36);
  const size_t o_Dinv = A.take<double>(TP * 6), o_bl = A.take<double>(TP * 3);
  const size_t part_per_win = (size_t)2 * nblk_scope * kBAPartWidth + 8;
  const size_t o_part = A.take<double>(part_per_win * B);
  size_t spart_total = 0;
  std::vector<size_t> spart_off(B, 0);
  for (int w = 0; w < B; w++) {
    spart_off[w] = spart_total;
    if (wh[w].acc_mode && wh[w].acc_mode != 6) spart_total += (size_t)nblk_scope * wh[w].acc_len;
  }
  const size_t o_spart = A.take<double>(spart_total);
  const size_t o_lband = A.take<double>(tile ? (size_t)w0.Ncf * (w0.bw + 1) * 36 : 0);
  const size_t o_cpart = A.take<double>(tile ? (size_t)16 * ctx->n_sm : 0);
  p->off_bcr = A.take<double>(tile && p->use_bcr ? p->bcr.total : 0);
  const size_t o_ticket = A.take<unsigned int>(tile ? 8 : 0);
  p->off_pose_out = A.take<double>(TC * 7);
  p->off_pts_out = A.take<double>(TP * 3);
  p->off_inlier = A.take<uint8_t>(TO);
  p->off_stats = A.take<urmvo_ba_stats>(B);
  p->dev_bytes = A.off;
  cudaError_t ce = cudaSuccess;
  if (borrow_ws && !ctx->ws_in_use) {
    if (ctx->ws_bytes < p->dev_bytes) {
      if (ctx->ws_dev) cudaFree(ctx->ws_dev);
      ctx->ws_dev = nullptr; ctx->ws_bytes = 0;
      const size_t want = p->dev_bytes + p->dev_bytes / 4;
      ce = cudaMalloc(&ctx->ws_dev, want);
      if (ce == cudaSuccess) ctx->ws_bytes = want;
    }
    if (ce == cudaSuccess) { p->dev = ctx->ws_dev; p->borrowed_dev = true; ctx->ws_in_use = true; }
  } else {
    ce = cudaMalloc(&p->dev, p->dev_bytes);
  }
  if (ce != cudaSuccess) { delete p; return fail(URMVO_ERR_CUDA, std::string("cudaMalloc BA plan: ") + cudaGetErrorString(ce)); }
  unsigned char* D = p->dev;

  // ---- host staging of the small index arrays + descriptors (pinned), big arrays copied directly
  const size_t idx_bytes = upload_end - o_pt_start;
  if (stereo) all_sorted = false;  // the (u, v, u_right) triples are always re-packed through the staging buffer
  if (stereo && perm.empty()) { perm.resize(TO); for (size_t o = 0; o < TO; o++) perm[o] = (int)o; }
  if (ctx->ensure_pinned(idx_bytes + (all_sorted ? 0 : TO * (sizeof(double) * 2 + 2 * sizeof(int))) + (tile ? TP * 3 * sizeof(double) : 0) +
                         (stereo ? TO * (sizeof(double) + 1) : 0))) {
    urmvo_ba_plan_destroy(p);
    return fail(URMVO_ERR_CUDA, "cudaMallocHost failed");
  }
  unsigned char* H = (unsigned char*)ctx->pinned;  // mirrors [o_pt_start, upload_end)
  auto hp = [&](size_t off) { return H + (off - o_pt_start); };
  size_t c_pt = 0, c_ncf = 0, c_blk = 0, c_grp = 0, c_rec = 0;
  BAWin* hw = (BAWin*)hp(p->off_wins);
  for (int w = 0; w < B; w++) {
    const WinHost& W = wh[w];
    const size_t c0 = cam_off[w], p0 = pt_off[w], ob0 = obs_off[w];
    int* h_pt_start = (int*)hp(o_pt_start) + c_pt;
    std::memcpy(h_pt_start, W.pt_start.data(), (W.Np + 1) * sizeof(int));
    std::memcpy((int*)hp(o_cam_free) + c0, W.cam_free.data(), W.Nc * sizeof(int));
    if (!W.grp_pt.empty()) std::memcpy((int*)hp(o_grp) + c_grp, W.grp_pt.data(), W.grp_pt.size() * sizeof(int));
    std::memcpy((int*)hp(o_row_ptr) + c_ncf, W.row_ptr.data(), (W.Ncf + 1) * sizeof(int));
    std::memcpy((int*)hp(o_lrow_ptr) + c_ncf, W.lrow_ptr.data(), (W.Ncf + 1) * sizeof(int));
    if (W.nblk) std::memcpy((int*)hp(o_col) + c_blk, W.col.data(), W.nblk * sizeof(int));
    if (!W.lcol.empty()) {
      std::memcpy((int*)hp(o_lcol) + c_blk, W.lcol.data(), W.lcol.size() * sizeof(int));
      std::memcpy((int*)hp(o_lblk) + c_blk, W.lblk.data(), W.lblk.size() * sizeof(int));
    }
    BAWin& d = hw[w];
    std::memset(&d, 0, sizeof(d));
    d.Nc = W.Nc; d.Ncf = W.Ncf; d.Np = W.Np; d.No = W.No; d.nblk = W.nblk; d.kmax = W.kmax;
    d.acc_mode = W.acc_mode; d.acc_len = W.acc_len;
    d.Spart = (double*)(D + o_spart) + spart_off[w];
    for (int k = 0; k < 4; k++) d.intr[k] = intr[k];
    d.pose_in = (const double*)(D + o_pose_in) + c0 * 7;
    d.pts_in = (const double*)(D + o_pts_in) + p0 * 3;
    d.uv = (const double*)(D + o_uv) + ob0 * 2;
    d.ocam = (const int*)(D + o_ocam) + ob0;
    d.ur = stereo ? (const double*)(D + o_ur) + ob0 : nullptr;
    d.okind = stereo ? (const uint8_t*)(D + o_okind) + ob0 : nullptr;
    d.intr_tab = stereo ? (const double*)(D + o_tab) : nullptr;
    d.pt_start = (const int*)(D + o_pt_start) + c_pt;
    d.opt = (const int*)(D + o_opt) + ob0;
    d.grp_pt = (const int*)(D + o_grp) + c_grp;
    d.n_grp = W.grp_pt.empty() ? 0 : (int)W.grp_pt.size() - 1;
    d.cam_free = (const int*)(D + o_cam_free) + c0;
    d.row_ptr = (const int*)(D + o_row_ptr) + c_ncf;
    d.col = (const int*)(D + o_col) + c_blk;
    d.lrow_ptr = (const int*)(D + o_lrow_ptr) + c_ncf;
    d.lcol = (const int*)(D + o_lcol) + c_blk;
    d.lblk = (const int*)(D + o_lblk) + c_blk;
    for (int k = 0; k < 2; k++) {
      d.cam[k] = (double*)(D + o_cam[k]) + c0 * 7;
      d.camRt[k] = (double*)(D + o_camRt[k]) + c0 * 12;
      d.pts[k] = (double*)(D + o_pts[k]) + p0 * 3;
    }
    d.level = D + o_level + ob0;
    d.rec = (ObsRec*)(D + o_rec) + c_rec;
    if (W.acc_mode >= 2) c_rec += (W.grp_pt.size() - 1) * 32;
    d.S = (double*)(D + o_S) + c_blk * 36;
    double* vec = (double*)(D + o_vec) + c_ncf * 6 * 8;
    const size_t n6 = (size_t)W.Ncf * 6;
    d.bs = vec; d.bp = vec + n6; d.hdiag = vec + 2 * n6; d.xp = vec + 3 * n6;
    d.r = vec + 4 * n6; d.z = vec + 5 * n6; d.p = vec + 6 * n6; d.Ap = vec + 7 * n6;
    d.Minv = (double*)(D + o_Minv) + c_ncf * 36;
    d.Dinv = (double*)(D + o_Dinv) + p0 * 6;
    d.bl = (double*)(D + o_bl) + p0 * 3;
    d.part = (double*)(D + o_part) + part_per_win * w;
    d.pose_out = (double*)(D + p->off_pose_out) + c0 * 7;
    d.pts_out = (double*)(D + p->off_pts_out) + p0 * 3;
    d.inlier = D + p->off_inlier + ob0;
    d.stats = (urmvo_ba_stats*)(D + p->off_stats) + w;
    if (tile) {
      std::memcpy(hp(o_chunk_grp), W.chunk_grp.data(), W.chunk_grp.size() * sizeof(int));
      std::memcpy(hp(o_grp_cbase), W.grp_cbase.data(), W.grp_cbase.size() * sizeof(int));
      std::memcpy(hp(o_chunk_blk), W.chunk_blk.data(), W.chunk_blk.size() * sizeof(int));
      if (!W.blk_desc.empty()) std::memcpy(hp(o_blk_desc), W.blk_desc.data(), W.blk_desc.size() * sizeof(int));
      d.n_chunk = W.n_chunk; d.bw = W.bw;
      d.chunk_grp = (const int*)(D + o_chunk_grp);
      d.grp_cbase = (const int*)(D + o_grp_cbase);
      d.chunk_blk = (const int*)(D + o_chunk_blk);
      d.blk_desc = (const int*)(D + o_blk_desc);
      d.Lband = (double*)(D + o_lband);
      d.cpart = (double*)(D + o_cpart);
      d.ticket = (unsigned int*)(D + o_ticket);
      d.hdiag = (double*)(D + o_hd);
    }
    if (w == 0) p->run.timing_stats = d.stats;
    c_pt += W.Np + 1;
    c_ncf += W.Ncf + 1;
    c_blk += W.nblk;
    c_grp += W.grp_pt.size();
  }
  cudaStream_t s = ctx->stream;
  auto up = [&](size_t off, const void* src, size_t bytes) {
    return bytes ? cudaMemcpyAsync(D + off, src, bytes, cudaMemcpyHostToDevice, s) : cudaSuccess;
  };
  if (mg) {
    int* hg = (int*)hp(p->off_gather);
    std::memcpy(hg, mg->kf_slot, TC * sizeof(int));
    std::memcpy(hg + TC, mg->pt_slot, TP * sizeof(int));
    std::memcpy(hg + TC + TP, mg->obs_slot, TO * sizeof(int));
  }
  cudaError_t e1 = up(o_pt_start, H, idx_bytes);
  cudaError_t e2 = mg ? cudaSuccess : up(o_pose_in, poses, TC * 7 * sizeof(double));
  cudaError_t e3;
  std::vector<int> pt_inv;
  if (tile) {  // points renumbered by first free camera
    double* hpts = (double*)(H + idx_bytes + (all_sorted ? 0 : TO * (sizeof(double) * 2 + 2 * sizeof(int))));
    pt_inv.resize(TP);
    for (size_t q = 0; q < TP; q++) {
      const int l = p->pt_perm[q];
      pt_inv[l] = (int)q;
      hpts[q * 3] = pts[(size_t)l * 3]; hpts[q * 3 + 1] = pts[(size_t)l * 3 + 1]; hpts[q * 3 + 2] = pts[(size_t)l * 3 + 2];
    }
    e3 = up(o_pts_in, hpts, TP * 3 * sizeof(double));
  } else {
    e3 = mg ? cudaSuccess : up(o_pts_in, pts, TP * 3 * sizeof(double));
  }
  cudaError_t e4, e5;
  cudaError_t e8 = cudaSuccess;
  if (mg) {  // values come from the device-resident map: upload the three slot lists and gather
    if (!all_sorted) { urmvo_ba_plan_destroy(p); return fail(URMVO_ERR_ARG, "map window: observations must be point-major"); }
    int* dg = (int*)(D + p->off_gather);  // uploaded with the index structure (e1)
    e4 = launch_map_gather((double*)(D + o_pose_in), (double*)(D + o_pts_in), (double*)(D + o_uv), mg->d_kf, mg->d_pt, mg->d_uv,
                             dg, dg + TC, dg + TC + TP, (int)TC, (int)TP, (int)TO, s);
    ctx->launches++;
    e5 = up(o_ocam, cam, TO * sizeof(int));
    e8 = up(o_opt, pt, TO * sizeof(int));
  } else if (all_sorted) {
    e4 = up(o_uv, uv, TO * 2 * sizeof(double));
    e5 = up(o_ocam, cam, TO * sizeof(int));
    e8 = up(o_opt, pt, TO * sizeof(int));
  } else {
    double* huv = (double*)(H + idx_bytes);
    int* hcam = (int*)(huv + TO * 2);
    int* hpt = hcam + TO;
    for (size_t o = 0; o < TO; o++) {
      huv[o * 2] = uv[(size_t)perm[o] * uvs];
      huv[o * 2 + 1] = uv[(size_t)perm[o] * uvs + 1];
      hcam[o] = cam[perm[o]];
      hpt[o] = tile ? pt_inv[pt[perm[o]]] : pt[perm[o]];
    }
    e4 = up(o_uv, huv, TO * 2 * sizeof(double));
    e5 = up(o_ocam, hcam, TO * sizeof(int));
    e8 = up(o_opt, hpt, TO * sizeof(int));
    if (stereo) {
      double* hur = (double*)(H + idx_bytes + TO * (sizeof(double) * 2 + 2 * sizeof(int)) + (tile ? TP * 3 * sizeof(double) : 0));
      uint8_t* hk = (uint8_t*)(hur + TO);
      bool bad_model = false;
      for (size_t o = 0; o < TO; o++) {
        hur[o] = uv[(size_t)perm[o] * 3 + 2];
        hk[o] = n_models > 0 ? kind[perm[o]] : (kind[perm[o]] ? 1 : 0);
        bad_model = bad_model || (hk[o] >> 1) >= n_tab;
      }
      if (bad_model) { urmvo_ba_plan_destroy(p); return fail(URMVO_ERR_ARG, "ba_plan_create: camera model index out of range"); }
      if (e4 == cudaSuccess) e4 = up(o_ur, hur, TO * sizeof(double));
      if (e4 == cudaSuccess) e4 = up(o_okind, hk, TO);
      if (e4 == cudaSuccess) e4 = up(o_tab, intr, (size_t)5 * n_tab * sizeof(double));
    }
  }
  cudaError_t e6 = cudaMemsetAsync(D + p->off_stats, 0, sizeof(urmvo_ba_stats) * B, s);
  if (sharded && e6 == cudaSuccess) e6 = cudaMemsetAsync(D + o_scal, 0, o_vec - o_scal + (size_t)8 * p->n6 * sizeof(double), s);
  const size_t state_bytes = std::max(shard_state_bytes(), lg_state_bytes());
  if ((sharded || tile) && e6 == cudaSuccess) e6 = cudaMalloc(&p->shard_state, state_bytes);
  if ((sharded || tile) && e6 == cudaSuccess) e6 = cudaHostAlloc(&p->shard_state_host, state_bytes, cudaHostAllocMapped);
  if (tile && e6 == cudaSuccess) e6 = cudaMemsetAsync(D + o_hd, 0, o_vec - o_hd + (size_t)8 * p->n6 * sizeof(double), s);
  if (tile)
    for (cudaEvent_t& ev : p->ev)
      if (e6 == cudaSuccess) e6 = cudaEventCreate(&ev);
  // the pinned staging buffer is reused by later calls: wait for the copies that read it
  cudaError_t e7 = cudaStreamSynchronize(s);
  for (cudaError_t e : {e1, e2, e3, e4, e5, e6, e7, e8})
    if (e != cudaSuccess) {
      urmvo_ba_plan_destroy(p);
      return fail(URMVO_ERR_CUDA, std::string("ba_plan_create upload: ") + cudaGetErrorString(e));
    }
  *out = p;
  return URMVO_OK;
}

extern "C" int urmvo_ba_plan_create(urmvo_ctx* ctx, urmvo_ba_plan** out, int B, const int32_t* cam_off,
                                    const int32_t* pt_off, const int32_t* obs_off, const double* poses,
                                    const uint8_t* fixed, const double* pts, const double* uv,
                                    const int32_t* cam, const int32_t* pt, const double* intr,
                                    double chi2_thr, int it0, int it1, const urmvo_ba_options* opts) {
  return ba_plan_create_impl(ctx, out, B, cam_off, pt_off, obs_off, poses, fixed, pts, uv, cam, pt, intr, chi2_thr,
                             it0, it1, opts, false, nullptr);
}

// ------------------------------------------------------------------ point-sharded BA over NCCL

extern "C" int urmvo_nccl_unique_id(uint8_t* id128) {
  std::string err;
  if (!id128) return fail(URMVO_ERR_ARG, "nccl_unique_id: null");
  if (!g_nccl.load(err)) return fail(URMVO_ERR_NCCL, err);
  NcclApi::UniqueId id;
  const int rc = g_nccl.GetUniqueId(&id);
  if (rc != 0) return fail(URMVO_ERR_NCCL, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(rc));
  std::memcpy(id128, id.internal, 128);
  return URMVO_OK;
}

extern "C" int urmvo_comm_init(urmvo_ctx* ctx, int rank, int world, const uint8_t* id128) {
  if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return fail(URMVO_ERR_ARG, "comm_init: bad arguments");
  std::string err;
  if (!g_nccl.load(err)) return fail(URMVO_ERR_NCCL, err);
  CU_TRY(cudaSetDevice(ctx->device));
  NcclApi::UniqueId id;
  std::memcpy(id.internal, id128, 128);
  void* comm = nullptr;
  const int rc = g_nccl.CommInitRank(&comm, world, id, rank);
  if (rc != 0) return fail(URMVO_ERR_NCCL, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(rc));
  ctx->comm = comm; ctx->rank = rank; ctx->world = world;
  return URMVO_OK;
}

extern "C" int urmvo_ba_covisibility(int Nc, const uint8_t* fixed, int Np, int No, const int32_t* cam,
                                     const int32_t* pt, uint8_t* upper) {
  if (Nc <= 0 || Np < 0 || No < 0 || !fixed || !cam || !pt || !upper) return fail(URMVO_ERR_ARG, "ba_covisibility: null input or negative size");
  try {
  std::vector<int> cf(Nc, -1);
  int n = 0;
  for (int c = 0; c < Nc; c++) if (!fixed[c]) cf[c] = n++;
  std::memset(upper, 0, (size_t)n * n);
  std::vector<std::vector<int>> per_pt(Np);
  for (int o = 0; o < No; o++) {
    if (pt[o] < 0 || pt[o] >= Np || cam[o] < 0 || cam[o] >= Nc) return fail(URMVO_ERR_ARG, "ba_covisibility: index out of range");
    if (cf[cam[o]] >= 0) per_pt[pt[o]].push_back(cf[cam[o]]);
  }
  for (auto& v : per_pt)
    for (size_t a = 0; a < v.size(); a++)
      for (size_t b = 0; b < v.size(); b++)
        if (v[a] <= v[b]) upper[(size_t)v[a] * n + v[b]] = 1;
  return n;
  } catch (const std::exception& e) {
    return fail(URMVO_ERR_ARG, std::string("ba_covisibility: ") + e.what());
  }
}

extern "C" int urmvo_sharded_ba_create(urmvo_ctx* ctx, urmvo_ba_plan** plan, int Nc, const double* poses,
                                       const uint8_t* fixed, int Np, const double* pts, int No, const double* uv,
                                       const int32_t* cam, const int32_t* pt, const double* intr, double chi2_thr,
                                       int it0, int it1, const uint8_t* covis, const urmvo_ba_options* opts) {
  if (ctx && ctx->world > 1 && !covis)
    return fail(URMVO_ERR_ARG, "sharded_ba_create: with more than one rank the OR-ed co-visibility (urmvo_ba_covisibility) is "
                               "required so that every rank builds the same structure of S");
  const int32_t co[2] = {0, Nc}, po[2] = {0, Np}, oo[2] = {0, No};
  urmvo_ba_options o = {};
  if (opts) o = *opts;
  o.force_atomic = 1;  // when the tile mode does not fit: global atomics + BSR PCG (the round-1 path)
  return ba_plan_create_impl(ctx, plan, 1, co, po, oo, poses, fixed, pts, uv, cam, pt, intr, chi2_thr, it0, it1, &o,
                             true, covis);
}


// Tile mode (csrc/ba_large.cu): the LM loop of one large problem (alone or as one rank of the
// point-sharded solve).  Whole LM iterations are enqueued without synchronising: the LM state lives on
// the device, kernels of trials that turned out not to be needed return at once; the host looks at
// the state once per batch of enqueued trials (normally once per optimize() call).  Every rank
// enqueues the same sequence because every rank sees the same reduced values.
static int run_large(urmvo_ba_plan* p) {
  urmvo_ctx* ctx = p->ctx;
  cudaStream_t s = ctx->stream;
  const BAWin* wins = (const BAWin*)(p->dev + p->off_wins);
  double* scal = (double*)(p->dev + p->off_scal);
  double* hd = (double*)(p->dev + p->off_hd);
  const int G = p->grid_blocks;
  const bool multi = ctx->comm && ctx->world > 1;
  auto allreduce = [&](double* buf, size_t n) -> int {
    if (!multi) return 0;
    const int rc = g_nccl.AllReduce(buf, buf, n, NcclApi::kFloat64, NcclApi::kSum, ctx->comm, s);
    if (rc != 0) return fail(URMVO_ERR_NCCL, std::string("ncclAllReduce: ") + g_nccl.GetErrorString(rc));
    return 0;
  };
#define LG_TRY(expr) do { cudaError_t _e = (expr); ctx->launches++; if (_e != cudaSuccess) return fail(URMVO_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } while (0)
  for (float& v : p->phase_ms) v = 0.f;
  p->n_trial_launches = 0; p->n_host_syncs = 0;
  struct Mark { int kind; };  // events are only recorded for the first trial of every batch (cheap, representative)
  LG_TRY(launch_lg_init(wins, p->shard_state, 2 * G, s));
  LG_TRY(launch_lg_pack(wins, 2 * G, s));
  for (int pass = 0; pass < 2; pass++) {
    const int n_iter = pass == 0 ? p->run.it0 : p->run.it1;
    LG_TRY(launch_lg_begin_pass(p->shard_state, pass == 0 ? 1 : 0, n_iter, s));
    if (n_iter > 0) {
      // computeLambdaInit: diag(Hpp), max diag(Hll), chi2 — one all-reduce of [hdiag | scal]
      CU_TRY(cudaMemsetAsync(hd, 0, p->n_reduce_diag * sizeof(double), s));
      LG_TRY(launch_lg_lin(wins, p->run, p->shard_state, scal, 1, ctx->rank, G, s));
      if (allreduce(hd, p->n_reduce_diag)) return URMVO_ERR_NCCL;
      LG_TRY(launch_lg_lambda(wins, p->shard_state, scal, ctx->world, s));
      int remaining = n_iter;
      for (;;) {
        for (int t = 0; t < remaining; t++) {
          const bool timed = (t == 0);
          CU_TRY(cudaMemsetAsync(scal, 0, p->n_reduce_main * sizeof(double), s));
          if (timed) CU_TRY(cudaEventRecord(p->ev[0], s));
          LG_TRY(launch_lg_lin(wins, p->run, p->shard_state, scal, 0, ctx->rank, G, s));
          if (timed) CU_TRY(cudaEventRecord(p->ev[1], s));
          if (allreduce(scal, p->n_reduce_main)) return URMVO_ERR_NCCL;
          if (timed) CU_TRY(cudaEventRecord(p->ev[2], s));
          if (p->use_bcr) {
            int nl = 0;
            LG_TRY(launch_bcr_solve(wins, p->shard_state, p->bcr, (double*)(p->dev + p->off_bcr), s, &nl));
            ctx->launches += nl - 1;
            p->n_solve_launches = nl;
          } else {
            LG_TRY(launch_lg_solve(wins, p->shard_state, p->band_m, p->ncf, s));
          }
          if (timed) CU_TRY(cudaEventRecord(p->ev[3], s));
          LG_TRY(launch_lg_backsub(wins, p->run, p->shard_state, scal, 2 * G, s));
          if (allreduce(scal + 2, 2)) return URMVO_ERR_NCCL;
          LG_TRY(launch_lg_decide(p->shard_state, scal, p->shard_state_host, s));
          if (timed) CU_TRY(cudaEventRecord(p->ev[4], s));
          p->n_trial_launches++;
        }
        CU_TRY(cudaStreamSynchronize(s));
        p->n_host_syncs++;
        for (int k = 0; k < 4; k++) {
          float ms = 0.f;
          if (cudaEventElapsedTime(&ms, p->ev[k], p->ev[k + 1]) == cudaSuccess) p->phase_ms[k] += ms;
        }
        int active = 0, it = 0;
        lg_flags(p->shard_state_host, &active, &it);
        if (!active) break;
        remaining = std::max(1, n_iter - it);
      }
    }
    LG_TRY(launch_lg_classify(wins, p->run, p->shard_state, pass, 2 * G, s));
    ctx->launches++;
    if (pass == 0) LG_TRY(launch_lg_pack(wins, 2 * G, s));  // new edge levels
  }
  LG_TRY(launch_lg_finish(wins, p->shard_state, 2 * G, s));
#undef LG_TRY
  CU_TRY(cudaStreamSynchronize(s));
  return URMVO_OK;
}

// Host-driven LM loop of the sharded problem: identical control flow on every rank, the decisions
// come back from the device after each trial (one small pinned read per trial).
extern "C" int urmvo_sharded_ba_run(urmvo_ba_plan* p) {
  if (!p || !p->sharded) return fail(URMVO_ERR_ARG, "sharded_ba_run: not a sharded plan");
  urmvo_ctx* ctx = p->ctx;
  CU_TRY(cudaSetDevice(ctx->device));
  if (p->tile) return run_large(p);
  cudaStream_t s = ctx->stream;
  const BAWin* wins = (const BAWin*)(p->dev + p->off_wins);
  double* scal = (double*)(p->dev + p->off_scal);
  double* hdiag = (double*)(p->dev + p->off_hdiag);
  const int G = p->grid_blocks, T = p->threads;
  auto allreduce = [&](double* buf, size_t n, int op) -> int {
    if (!ctx->comm || ctx->world == 1) return 0;
    const int rc = g_nccl.AllReduce(buf, buf, n, NcclApi::kFloat64, op, ctx->comm, s);
    if (rc != 0) return fail(URMVO_ERR_NCCL, std::string("ncclAllReduce: ") + g_nccl.GetErrorString(rc));
    return 0;
  };
#define SH_TRY(expr) do { cudaError_t _e = (expr); ctx->launches++; if (_e != cudaSuccess) return fail(URMVO_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } while (0)
  SH_TRY(launch_sh_init(wins, p->shard_state, G, T, s));
  for (int pass = 0; pass < 2; pass++) {
    const int n_iter = pass == 0 ? p->run.it0 : p->run.it1;
    SH_TRY(launch_sh_begin_pass(p->shard_state, pass == 0 ? 1 : 0, s));
    bool terminated = false;
    for (int it = 0; it < n_iter && !terminated; it++) {
      if (it == 0) {
        SH_TRY(launch_sh_lin(wins, p->run, p->shard_state, scal, p->kmax, 1, G, T, s));
        if (allreduce(hdiag, p->n6, NcclApi::kSum)) return URMVO_ERR_NCCL;
        if (allreduce(scal, 1, NcclApi::kSum)) return URMVO_ERR_NCCL;
        if (allreduce(scal + 1, 1, NcclApi::kMax)) return URMVO_ERR_NCCL;
        SH_TRY(launch_sh_lambda(wins, p->shard_state, scal, s));
      }
      int cont = 1;
      while (cont) {
        SH_TRY(launch_sh_lin(wins, p->run, p->shard_state, scal, p->kmax, 0, G, T, s));
        if (allreduce(scal, p->n_reduce_main, NcclApi::kSum)) return URMVO_ERR_NCCL;
        SH_TRY(launch_sh_solve(wins, p->run, p->shard_state, scal, p->kmax, ctx->rank, G, T, s));
        if (allreduce(scal + 2, 2, NcclApi::kSum)) return URMVO_ERR_NCCL;
        SH_TRY(launch_sh_decide(p->shard_state, scal, p->shard_state_host, s));
        CU_TRY(cudaStreamSynchronize(s));
        int term = 0;
        shard_flags(p->shard_state_host, &cont, &term);
        if (!cont && term) terminated = true;
      }
    }
    SH_TRY(launch_sh_classify(wins, p->run, p->shard_state, pass, G, T, s));
  }
  SH_TRY(launch_sh_finish(wins, p->shard_state, G, T, s));
#undef SH_TRY
  CU_TRY(cudaStreamSynchronize(s));
  return URMVO_OK;
}

extern "C" int urmvo_ba_plan_run(urmvo_ba_plan* p) {
  if (!p) return fail(URMVO_ERR_ARG, "ba_plan_run: null plan");
  if (p->sharded) return urmvo_sharded_ba_run(p);
  CU_TRY(cudaSetDevice(p->ctx->device));
  if (p->tile) return run_large(p);
  const BAWin* wins = (const BAWin*)(p->dev + p->off_wins);
  cudaError_t e;
  if (p->use_grid) e = launch_ba_grid(wins, p->run, p->kmax, p->grid_blocks, p->threads, p->ctx->stream, p->stereo ? 1 : 0);
  else {
    e = launch_ba_cluster(wins, p->run, p->batch_mode, p->kmax, p->work_stride, p->ints_per_warp, p->n_clusters, p->cluster_size, p->threads, p->ctx->stream);
    // a 16-CTA (non-portable) cluster with ~200 KB of shared memory per CTA may not be schedulable (MIG, partly
    // disabled GPCs, another resident context): fall back to smaller clusters instead of failing the keyframe
    while (e != cudaSuccess && p->cluster_auto && p->cluster_size > 1 &&
           (e == cudaErrorInvalidConfiguration || e == cudaErrorLaunchOutOfResources || e == cudaErrorInvalidValue ||
            e == cudaErrorCooperativeLaunchTooLarge)) {
      (void)cudaGetLastError();
      p->cluster_size /= 2;
      e = launch_ba_cluster(wins, p->run, p->batch_mode, p->kmax, p->work_stride, p->ints_per_warp, p->n_clusters, p->cluster_size, p->threads, p->ctx->stream);
    }
  }
  if (e != cudaSuccess) return fail(URMVO_ERR_CUDA, std::string("BA kernel launch: ") + cudaGetErrorString(e));
  p->ctx->launches++;
  return URMVO_OK;
}

extern "C" int urmvo_ba_plan_download(urmvo_ba_plan* p, double* poses, double* pts, uint8_t* inlier,
                                      urmvo_ba_stats* stats) {
  if (!p) return fail(URMVO_ERR_ARG, "ba_plan_download: null plan");
  CU_TRY(cudaSetDevice(p->ctx->device));
  cudaStream_t s = p->ctx->stream;
  if (poses) CU_TRY(cudaMemcpyAsync(poses, p->dev + p->off_pose_out, (size_t)p->total_c * 7 * sizeof(double), cudaMemcpyDeviceToHost, s));
  std::vector<double> ptmp;
  if (pts && p->pt_perm.empty()) CU_TRY(cudaMemcpyAsync(pts, p->dev + p->off_pts_out, (size_t)p->total_p * 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (pts && !p->pt_perm.empty()) {
    ptmp.resize((size_t)p->total_p * 3);
    CU_TRY(cudaMemcpyAsync(ptmp.data(), p->dev + p->off_pts_out, ptmp.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
  }
  std::vector<uint8_t> tmp;
  if (inlier) {
    if (p->perm.empty()) {
      CU_TRY(cudaMemcpyAsync(inlier, p->dev + p->off_inlier, (size_t)p->total_o, cudaMemcpyDeviceToHost, s));
    } else {
      tmp.resize(p->total_o);
      CU_TRY(cudaMemcpyAsync(tmp.data(), p->dev + p->off_inlier, (size_t)p->total_o, cudaMemcpyDeviceToHost, s));
    }
  }
  if (stats) CU_TRY(cudaMemcpyAsync(stats, p->dev + p->off_stats, sizeof(urmvo_ba_stats) * p->B, cudaMemcpyDeviceToHost, s));
  CU_TRY(cudaStreamSynchronize(s));
  if (inlier && !p->perm.empty())
    for (int o = 0; o < p->total_o; o++) inlier[p->perm[o]] = tmp[o];
  if (pts && !p->pt_perm.empty())
    for (int q = 0; q < p->total_p; q++) {
      const size_t l = (size_t)p->pt_perm[q];
      pts[l * 3] = ptmp[(size_t)q * 3]; pts[l * 3 + 1] = ptmp[(size_t)q * 3 + 1]; pts[l * 3 + 2] = ptmp[(size_t)q * 3 + 2];
    }
  return URMVO_OK;
}

extern "C" int urmvo_ba_plan_phase_info(urmvo_ba_plan* p, float* ms4, int32_t* info5) {
  if (!p || !ms4 || !info5) return fail(URMVO_ERR_ARG, "ba_plan_phase_info: null argument");
  for (int k = 0; k < 4; k++) ms4[k] = p->phase_ms[k];
  info5[0] = p->tile ? (p->use_bcr ? 2 + 16 * p->bcr.L + 4096 * p->bcr.K : 1) : 0;  // 1: band solve, 2 + 16 levels + 4096 super-blocks: cyclic reduction
  info5[1] = p->tile ? p->band_m - 1 : -1;
  info5[2] = p->n_trial_launches;
  info5[3] = p->n_host_syncs;
  info5[4] = (int32_t)p->n_reduce_main;
  return URMVO_OK;
}

extern "C" int urmvo_debug_lg_timing(uint64_t* cycles8, int reset) {
  if (!cycles8) return fail(URMVO_ERR_ARG, "debug_lg_timing: null output");
  unsigned long long t[8];
  CU_TRY(lg_timing_read(t, reset != 0));
  for (int i = 0; i < 8; i++) cycles8[i] = t[i];
  return URMVO_OK;
}

extern "C" int urmvo_debug_ba_timing(uint64_t* cycles8, int reset) {
  if (!cycles8) return fail(URMVO_ERR_ARG, "debug_ba_timing: null output");
  unsigned long long t[8];
  CU_TRY(ba_timing_read(t, reset != 0));
  for (int i = 0; i < 8; i++) cycles8[i] = t[i];
  return URMVO_OK;
}

extern "C" int urmvo_local_ba_batch(urmvo_ctx* ctx, int B, const int32_t* cam_off, const int32_t* pt_off,
                                    const int32_t* obs_off, double* poses, const uint8_t* fixed, double* pts,
                                    const double* uv, const int32_t* cam, const int32_t* pt, const double* intr,
                                    double chi2_thr, int it0, int it1, uint8_t* inlier, urmvo_ba_stats* stats,
                                    const urmvo_ba_options* opts) {
  // URMVO_B200_TRACE=1: host-side split of the call (structure + upload | kernel | read-back) on stderr
  static const bool trace = std::getenv("URMVO_B200_TRACE") != nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  urmvo_ba_plan* p = nullptr;
  int rc = ba_plan_create_impl(ctx, &p, B, cam_off, pt_off, obs_off, poses, fixed, pts, uv, cam, pt, intr,
                               chi2_thr, it0, it1, opts, false, nullptr, /*borrow_ws=*/true);
  if (rc != URMVO_OK) return rc;
  const auto t1 = std::chrono::steady_clock::now();
  rc = urmvo_ba_plan_run(p);
  const auto t2 = std::chrono::steady_clock::now();
  if (rc == URMVO_OK) rc = urmvo_ba_plan_download(p, poses, pts, inlier, stats);
  const auto t3 = std::chrono::steady_clock::now();
  urmvo_ba_plan_destroy(p);
  if (trace) {
    auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
      return std::chrono::duration<double, std::micro>(b - a).count();
    };
    std::fprintf(stderr, "[urmvo_b200] local_ba_batch B=%d: create+upload %.0f us, launch %.0f us, wait+download %.0f us, destroy %.0f us\n",
                 B, us(t0, t1), us(t1, t2), us(t2, t3), us(t3, std::chrono::steady_clock::now()));
  }
  return rc;
}

extern "C" int urmvo_local_ba_batch_stereo(urmvo_ctx* ctx, int B, const int32_t* cam_off, const int32_t* pt_off,
                                           const int32_t* obs_off, double* poses, const uint8_t* fixed, double* pts,
                                           const double* uv3, const uint8_t* kind, const int32_t* cam, const int32_t* pt,
                                           const double* intr5, double chi2_thr_mono, double chi2_thr_stereo, int it0,
                                           int it1, uint8_t* inlier, urmvo_ba_stats* stats, const urmvo_ba_options* opts) {
  if (!kind) return fail(URMVO_ERR_ARG, "local_ba_batch_stereo: null kind flags");
  urmvo_ba_plan* p = nullptr;
  int rc = ba_plan_create_impl(ctx, &p, B, cam_off, pt_off, obs_off, poses, fixed, pts, uv3, cam, pt, intr5, chi2_thr_mono,
                               it0, it1, opts, false, nullptr, /*borrow_ws=*/true, kind, chi2_thr_stereo);
  if (rc != URMVO_OK) return rc;
  rc = urmvo_ba_plan_run(p);
  if (rc == URMVO_OK) rc = urmvo_ba_plan_download(p, poses, pts, inlier, stats);
  urmvo_ba_plan_destroy(p);
  return rc;
}

extern "C" int urmvo_local_ba_batch_multicam(urmvo_ctx* ctx, int B, const int32_t* cam_off, const int32_t* pt_off,
                                             const int32_t* obs_off, double* poses, const uint8_t* fixed, double* pts,
                                             const double* uv3, const uint8_t* kind_model, const int32_t* cam,
                                             const int32_t* pt, int n_models, const double* intr5_tab,
                                             double chi2_thr_mono, double chi2_thr_stereo, int it0, int it1,
                                             uint8_t* inlier, urmvo_ba_stats* stats, const urmvo_ba_options* opts) {
  if (!kind_model || n_models <= 0) return fail(URMVO_ERR_ARG, "local_ba_batch_multicam: null kind / model bytes or no camera model");
  urmvo_ba_plan* p = nullptr;
  int rc = ba_plan_create_impl(ctx, &p, B, cam_off, pt_off, obs_off, poses, fixed, pts, uv3, cam, pt, intr5_tab, chi2_thr_mono,
                               it0, it1, opts, false, nullptr, /*borrow_ws=*/true, kind_model, chi2_thr_stereo, nullptr, n_models);
  if (rc != URMVO_OK) return rc;
  rc = urmvo_ba_plan_run(p);
  if (rc == URMVO_OK) rc = urmvo_ba_plan_download(p, poses, pts, inlier, stats);
  urmvo_ba_plan_destroy(p);
  return rc;
}

extern "C" int urmvo_local_ba_multicam(urmvo_ctx* ctx, int Nc, double* poses, const uint8_t* fixed, int Np, double* pts,
                                       int No, const double* uv3, const uint8_t* kind_model, const int32_t* cam,
                                       const int32_t* pt, int n_models, const double* intr5_tab, double chi2_thr_mono,
                                       double chi2_thr_stereo, int it0, int it1, uint8_t* inlier,
                                       urmvo_ba_stats* stats, const urmvo_ba_options* opts) {
  const int32_t co[2] = {0, Nc}, po[2] = {0, Np}, oo[2] = {0, No};
  return urmvo_local_ba_batch_multicam(ctx, 1, co, po, oo, poses, fixed, pts, uv3, kind_model, cam, pt, n_models,
                                       intr5_tab, chi2_thr_mono, chi2_thr_stereo, it0, it1, inlier, stats, opts);
}

extern "C" int urmvo_local_ba_stereo(urmvo_ctx* ctx, int Nc, double* poses, const uint8_t* fixed, int Np, double* pts,
                                     int No, const double* uv3, const uint8_t* kind, const int32_t* cam,
                                     const int32_t* pt, const double* intr5, double chi2_thr_mono,
                                     double chi2_thr_stereo, int it0, int it1, uint8_t* inlier, urmvo_ba_stats* stats,
                                     const urmvo_ba_options* opts) {
  const int32_t co[2] = {0, Nc}, po[2] = {0, Np}, oo[2] = {0, No};
  return urmvo_local_ba_batch_stereo(ctx, 1, co, po, oo, poses, fixed, pts, uv3, kind, cam, pt, intr5, chi2_thr_mono,
                                     chi2_thr_stereo, it0, it1, inlier, stats, opts);
}

extern "C" int urmvo_local_ba(urmvo_ctx* ctx, int Nc, double* poses, const uint8_t* fixed, int Np, double* pts,
                              int No, const double* uv, const int32_t* cam, const int32_t* pt,
                              const double* intr, double chi2_thr, int it0, int it1, uint8_t* inlier,
                              urmvo_ba_stats* stats, const urmvo_ba_options* opts) {
  const int32_t co[2] = {0, Nc}, po[2] = {0, Np}, oo[2] = {0, No};
  return urmvo_local_ba_batch(ctx, 1, co, po, oo, poses, fixed, pts, uv, cam, pt, intr, chi2_thr, it0, it1,
                              inlier, stats, opts);
}

// =====================================================================================  pose-only

struct urmvo_pose_plan {
  urmvo_ctx* ctx = nullptr;
  int B = 0, total_o = 0;
  double intr[4];
  double chi2_thr = 0, delta = 0;
  int rounds = 4, its = 10;
  unsigned char* dev = nullptr;
  bool borrowed = false;  // dev is the context's grow-only workspace (one-shot urmvo_pose_only_batch)
  size_t o_off = 0, o_pose_in = 0, o_uv = 0, o_X = 0, o_inl_in = 0, o_inl = 0, o_level = 0, o_pose_out = 0,
         o_ninl = 0, o_iters = 0;
  // stereo edges (EdgeStereoSE3ProjectXYZOnlyPose, reference src/g2o_optimization.cc:235-258)
  bool stereo = false;
  size_t o_ur = 0, o_kind = 0, o_tab = 0;
  double bf = 0, chi2_thr_s = 0, delta_s = 0;
};

extern "C" void urmvo_pose_plan_destroy(urmvo_pose_plan* p) {
  if (!p) return;
  if (p->borrowed) p->ctx->ws_in_use = false;
  else if (p->dev) { cudaSetDevice(p->ctx->device); cudaFree(p->dev); }
  delete p;
}

// uv_stride 2: mono measurements; 3 with kind != NULL: (u, v, u_right) with kind[o] = 1 marking a stereo edge
// (intr then carries bf as its fifth value).
static int pose_plan_create_impl(urmvo_ctx* ctx, urmvo_pose_plan** out, int B, const int32_t* obs_off,
                                      const double* poses, const double* uv, const double* Xw,
                                      const double* intr, double chi2_thr, int rounds, int its_per_round,
                                      const uint8_t* inlier, bool borrow_ws, int uv_stride = 2,
                                      const uint8_t* kind = nullptr, double chi2_thr_stereo = 0.0, int n_models = 0) try {
  // n_models > 0: intr is a table of n_models rows (fx fy cx cy bf) and kind[o] = stereo bit | camera model << 1
  // (the reference reads camera_list[mpc->id_camera] per constraint, src/g2o_optimization.cc:221-224, :243-250)
  if (!ctx || !out) return fail(URMVO_ERR_ARG, "pose_plan_create: null context / out");
  if (n_models < 0 || n_models > 128) return fail(URMVO_ERR_ARG, "pose_plan_create: 1..128 camera models");
  *out = nullptr;
  if (B <= 0 || !obs_off || !poses || !uv || !Xw || !intr) return fail(URMVO_ERR_ARG, "pose_plan_create: null or empty input");
  if (rounds < 0 || its_per_round < 0 || !(chi2_thr > 0)) return fail(URMVO_ERR_ARG, "pose_plan_create: bad rounds / threshold");
  const bool stereo = uv_stride == 3;
  if (stereo && (!kind || !(chi2_thr_stereo > 0))) return fail(URMVO_ERR_ARG, "pose_plan_create: stereo edges need kind flags and a positive threshold");
  for (int f = 0; f < B; f++)
    if (obs_off[f + 1] < obs_off[f]) return fail(URMVO_ERR_ARG, "pose_plan_create: obs_off must be non-decreasing");
  CU_TRY(cudaSetDevice(ctx->device));
  urmvo_pose_plan* p = new urmvo_pose_plan();
  p->ctx = ctx; p->B = B; p->total_o = obs_off[B] - obs_off[0];
  for (int k = 0; k < 4; k++) p->intr[k] = intr[k];
  p->chi2_thr = chi2_thr;
  p->delta = (double)(float)std::sqrt(chi2_thr);  // src/g2o_optimization.cc:205
  p->rounds = rounds; p->its = its_per_round;
  p->stereo = stereo;
  if (stereo) {
    p->bf = intr[4];
    p->chi2_thr_s = chi2_thr_stereo;
    p->delta_s = (double)(float)std::sqrt(chi2_thr_stereo);  // src/g2o_optimization.cc:210
  }
  const size_t TO = p->total_o;
  Arena A;
  p->o_off = A.take<int>(B + 1); p->o_pose_in = A.take<double>((size_t)B * 7);
  p->o_uv = A.take<double>(TO * 2); p->o_X = A.take<double>(TO * 3);
  p->o_ur = A.take<double>(stereo ? TO : 0); p->o_kind = A.take<uint8_t>(stereo ? TO : 0);
  p->o_tab = A.take<double>(stereo ? (size_t)5 * std::max(n_models, 1) : 0);
  p->o_inl_in = A.take<uint8_t>(TO); p->o_inl = A.take<uint8_t>(TO); p->o_level = A.take<uint8_t>(TO);
  p->o_pose_out = A.take<double>((size_t)B * 7); p->o_ninl = A.take<int>(B); p->o_iters = A.take<int>(B);
  cudaError_t ce = cudaSuccess;
  if (borrow_ws && !ctx->ws_in_use) {  // per-frame calls: no cudaMalloc / cudaFree each time
    if (ctx->ws_bytes < A.off) {
      if (ctx->ws_dev) cudaFree(ctx->ws_dev);
      ctx->ws_dev = nullptr; ctx->ws_bytes = 0;
      ce = cudaMalloc(&ctx->ws_dev, A.off + A.off / 4);
      if (ce == cudaSuccess) ctx->ws_bytes = A.off + A.off / 4;
    }
    if (ce == cudaSuccess) { p->dev = ctx->ws_dev; p->borrowed = true; ctx->ws_in_use = true; }
  } else {
    ce = cudaMalloc(&p->dev, A.off);
  }
  if (ce != cudaSuccess) { delete p; return fail(URMVO_ERR_CUDA, std::string("cudaMalloc pose plan: ") + cudaGetErrorString(ce)); }
  cudaStream_t s = ctx->stream;
  std::vector<int> off(B + 1);
  for (int f = 0; f <= B; f++) off[f] = obs_off[f] - obs_off[0];
  const size_t o0 = obs_off[0];
  cudaError_t e[6];
  e[0] = cudaMemcpyAsync(p->dev + p->o_off, off.data(), (B + 1) * sizeof(int), cudaMemcpyHostToDevice, s);
  e[1] = cudaMemcpyAsync(p->dev + p->o_pose_in, poses, (size_t)B * 7 * sizeof(double), cudaMemcpyHostToDevice, s);
  std::vector<double> uv2, ur;
  std::vector<uint8_t> kb;
  if (stereo && TO) {  // split (u, v, u_right) into the mono layout + one extra column
    uv2.resize(TO * 2); ur.resize(TO); kb.resize(TO);
    for (size_t o = 0; o < TO; o++) {
      uv2[o * 2] = uv[(o0 + o) * 3]; uv2[o * 2 + 1] = uv[(o0 + o) * 3 + 1]; ur[o] = uv[(o0 + o) * 3 + 2];
      kb[o] = n_models > 0 ? kind[o0 + o] : (kind[o0 + o] ? 1 : 0);
      if ((kb[o] >> 1) >= std::max(n_models, 1)) { urmvo_pose_plan_destroy(p); return fail(URMVO_ERR_ARG, "pose_plan_create: camera model index out of range"); }
    }
    e[2] = cudaMemcpyAsync(p->dev + p->o_uv, uv2.data(), TO * 2 * sizeof(double), cudaMemcpyHostToDevice, s);
    if (e[2] == cudaSuccess) e[2] = cudaMemcpyAsync(p->dev + p->o_ur, ur.data(), TO * sizeof(double), cudaMemcpyHostToDevice, s);
    if (e[2] == cudaSuccess) e[2] = cudaMemcpyAsync(p->dev + p->o_kind, kb.data(), TO, cudaMemcpyHostToDevice, s);
    if (e[2] == cudaSuccess) e[2] = cudaMemcpyAsync(p->dev + p->o_tab, intr, (size_t)5 * std::max(n_models, 1) * sizeof(double), cudaMemcpyHostToDevice, s);
  } else
  e[2] = TO ? cudaMemcpyAsync(p->dev + p->o_uv, uv + o0 * 2, TO * 2 * sizeof(double), cudaMemcpyHostToDevice, s) : cudaSuccess;
  e[3] = TO ? cudaMemcpyAsync(p->dev + p->o_X, Xw + o0 * 3, TO * 3 * sizeof(double), cudaMemcpyHostToDevice, s) : cudaSuccess;
  if (inlier) e[4] = TO ? cudaMemcpyAsync(p->dev + p->o_inl_in, inlier + o0, TO, cudaMemcpyHostToDevice, s) : cudaSuccess;
  else e[4] = TO ? cudaMemsetAsync(p->dev + p->o_inl_in, 1, TO, s) : cudaSuccess;
  e[5] = cudaStreamSynchronize(s);  // `off` is a stack-lifetime staging vector
  for (cudaError_t x : e)
    if (x != cudaSuccess) { urmvo_pose_plan_destroy(p); return fail(URMVO_ERR_CUDA, std::string("pose_plan_create upload: ") + cudaGetErrorString(x)); }
  *out = p;
  return URMVO_OK;
} catch (const std::exception& e) {  // no exception crosses the C ABI
  return fail(URMVO_ERR_ARG, std::string("pose_plan_create_impl: ") + e.what());
}

extern "C" int urmvo_pose_plan_create(urmvo_ctx* ctx, urmvo_pose_plan** out, int B, const int32_t* obs_off,
                                      const double* poses, const double* uv, const double* Xw,
                                      const double* intr, double chi2_thr, int rounds, int its_per_round,
                                      const uint8_t* inlier) {
  return pose_plan_create_impl(ctx, out, B, obs_off, poses, uv, Xw, intr, chi2_thr, rounds, its_per_round, inlier, false);
}

extern "C" int urmvo_pose_plan_run(urmvo_pose_plan* p) {
  if (!p) return fail(URMVO_ERR_ARG, "pose_plan_run: null plan");
  CU_TRY(cudaSetDevice(p->ctx->device));
  cudaStream_t s = p->ctx->stream;
  if (p->total_o) CU_TRY(cudaMemcpyAsync(p->dev + p->o_inl, p->dev + p->o_inl_in, p->total_o, cudaMemcpyDeviceToDevice, s));
  cudaError_t e = launch_pose_only(p->B, (const int*)(p->dev + p->o_off), (const double*)(p->dev + p->o_pose_in),
                                   (const double*)(p->dev + p->o_uv), (const double*)(p->dev + p->o_X), p->intr,
                                   p->chi2_thr, p->delta, p->rounds, p->its, p->dev + p->o_inl, p->dev + p->o_level,
                                   (double*)(p->dev + p->o_pose_out), (int*)(p->dev + p->o_ninl),
                                   (int*)(p->dev + p->o_iters), s,
                                   p->stereo ? (const double*)(p->dev + p->o_ur) : nullptr,
                                   p->stereo ? (const uint8_t*)(p->dev + p->o_kind) : nullptr,
                                   p->stereo ? (const double*)(p->dev + p->o_tab) : nullptr, p->chi2_thr_s, p->delta_s);
  if (e != cudaSuccess) return fail(URMVO_ERR_CUDA, std::string("pose kernel launch: ") + cudaGetErrorString(e));
  p->ctx->launches++;
  return URMVO_OK;
}

extern "C" int urmvo_pose_plan_download(urmvo_pose_plan* p, double* poses, uint8_t* inlier, int32_t* n_inlier,
                                        int32_t* lm_iters) {
  if (!p) return fail(URMVO_ERR_ARG, "pose_plan_download: null plan");
  CU_TRY(cudaSetDevice(p->ctx->device));
  cudaStream_t s = p->ctx->stream;
  if (poses) CU_TRY(cudaMemcpyAsync(poses, p->dev + p->o_pose_out, (size_t)p->B * 7 * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (inlier && p->total_o) CU_TRY(cudaMemcpyAsync(inlier, p->dev + p->o_inl, p->total_o, cudaMemcpyDeviceToHost, s));
  if (n_inlier) CU_TRY(cudaMemcpyAsync(n_inlier, p->dev + p->o_ninl, p->B * sizeof(int), cudaMemcpyDeviceToHost, s));
  if (lm_iters) CU_TRY(cudaMemcpyAsync(lm_iters, p->dev + p->o_iters, p->B * sizeof(int), cudaMemcpyDeviceToHost, s));
  CU_TRY(cudaStreamSynchronize(s));
  return URMVO_OK;
}

extern "C" int urmvo_pose_only_batch(urmvo_ctx* ctx, int B, const int32_t* obs_off, double* poses,
                                     const double* uv, const double* Xw, const double* intr, double chi2_thr,
                                     int rounds, int its_per_round, uint8_t* inlier, int32_t* n_inlier) {
  urmvo_pose_plan* p = nullptr;
  int rc = pose_plan_create_impl(ctx, &p, B, obs_off, poses, uv, Xw, intr, chi2_thr, rounds, its_per_round, inlier, true);
  if (rc != URMVO_OK) return rc;
  rc = urmvo_pose_plan_run(p);
  if (rc == URMVO_OK) rc = urmvo_pose_plan_download(p, poses, inlier ? inlier + obs_off[0] : nullptr, n_inlier, nullptr);
  urmvo_pose_plan_destroy(p);
  return rc;
}

extern "C" int urmvo_pose_only_batch_stereo(urmvo_ctx* ctx, int B, const int32_t* obs_off, double* poses,
                                            const double* uv3, const uint8_t* kind, const double* Xw,
                                            const double* intr5, double chi2_thr_mono, double chi2_thr_stereo,
                                            int rounds, int its_per_round, uint8_t* inlier, int32_t* n_inlier) {
  urmvo_pose_plan* p = nullptr;
  int rc = pose_plan_create_impl(ctx, &p, B, obs_off, poses, uv3, Xw, intr5, chi2_thr_mono, rounds, its_per_round, inlier,
                                 true, 3, kind, chi2_thr_stereo);
  if (rc != URMVO_OK) return rc;
  rc = urmvo_pose_plan_run(p);
  if (rc == URMVO_OK) rc = urmvo_pose_plan_download(p, poses, inlier ? inlier + obs_off[0] : nullptr, n_inlier, nullptr);
  urmvo_pose_plan_destroy(p);
  return rc;
}

extern "C" int urmvo_pose_only_batch_multicam(urmvo_ctx* ctx, int B, const int32_t* obs_off, double* poses,
                                              const double* uv3, const uint8_t* kind_model, const double* Xw,
                                              int n_models, const double* intr5_tab, double chi2_thr_mono,
                                              double chi2_thr_stereo, int rounds, int its_per_round, uint8_t* inlier,
                                              int32_t* n_inlier) {
  if (n_models <= 0) return fail(URMVO_ERR_ARG, "pose_only_batch_multicam: 1..128 camera models");
  urmvo_pose_plan* p = nullptr;
  int rc = pose_plan_create_impl(ctx, &p, B, obs_off, poses, uv3, Xw, intr5_tab, chi2_thr_mono, rounds, its_per_round, inlier,
                                 true, 3, kind_model, chi2_thr_stereo, n_models);
  if (rc != URMVO_OK) return rc;
  rc = urmvo_pose_plan_run(p);
  if (rc == URMVO_OK) rc = urmvo_pose_plan_download(p, poses, inlier ? inlier + obs_off[0] : nullptr, n_inlier, nullptr);
  urmvo_pose_plan_destroy(p);
  return rc;
}

// =====================================================================================  two-view

struct urmvo_tv_plan {
  urmvo_ctx* ctx = nullptr;
  TVBuffers b{};
  float sigma = 1.0f;
  float Kh[9];
  std::vector<int> m1, m2;
  unsigned char* dev = nullptr;
  bool borrowed = false;  // dev is the context's grow-only workspace (one-shot urmvo_two_view)
  int score_mode = 0;     // 0: the reference's symmetric point-line chi2, 1: Sampson error (fundamental model)
  bool ransac_done = false;
};

extern "C" void urmvo_tv_plan_destroy(urmvo_tv_plan* p) {
  if (!p) return;
  if (p->borrowed) p->ctx->ws_in_use = false;
  else if (p->dev) { cudaSetDevice(p->ctx->device); cudaFree(p->dev); }
  delete p;
}

static int tv_plan_create_impl(urmvo_ctx* ctx, urmvo_tv_plan** out, int n1, const float* keys1, int n2,
                               const float* keys2, const int32_t* matches12, const float* K, float sigma,
                               int n_hyp, const int32_t* sets, bool borrow_ws) {
  if (!ctx || !out) return fail(URMVO_ERR_ARG, "tv_plan_create: null context / out");
  *out = nullptr;
  if (n1 <= 0 || n2 <= 0 || !keys1 || !keys2 || !matches12 || !K || n_hyp <= 0 || !sets || !(sigma > 0))
    return fail(URMVO_ERR_ARG, "tv_plan_create: null or empty input");
  CU_TRY(cudaSetDevice(ctx->device));
  urmvo_tv_plan* p = new urmvo_tv_plan();
  p->ctx = ctx;
  p->sigma = sigma;
  std::memcpy(p->Kh, K, sizeof(p->Kh));
  for (int i = 0; i < n1; i++) {  // src/epipolar_geometry.cc:34-40
    if (matches12[i] >= 0) {
      if (matches12[i] >= n2) { delete p; return fail(URMVO_ERR_ARG, "tv_plan_create: match index out of range"); }
      p->m1.push_back(i);
      p->m2.push_back(matches12[i]);
    }
  }
  const int N = (int)p->m1.size();
  if (N < 8) { delete p; return fail(URMVO_ERR_ARG, "tv_plan_create: fewer than 8 matches"); }
  for (size_t i = 0; i < (size_t)n_hyp * 8; i++)
    if (sets[i] < 0 || sets[i] >= N) { delete p; return fail(URMVO_ERR_ARG, "tv_plan_create: sample index out of range"); }
  TVBuffers& b = p->b;
  b.n1 = n1; b.n2 = n2; b.N = N; b.n_hyp = n_hyp; b.words = (N + 31) / 32;
  Arena A;
  const size_t o_k1 = A.take<float>((size_t)n1 * 2), o_k2 = A.take<float>((size_t)n2 * 2);
  const size_t o_m1 = A.take<int>(N), o_m2 = A.take<int>(N), o_sets = A.take<int>((size_t)n_hyp * 8), o_K = A.take<float>(9);
  const size_t o_pn1 = A.take<float>((size_t)n1 * 2), o_pn2 = A.take<float>((size_t)n2 * 2), o_T1 = A.take<float>(9), o_T2 = A.take<float>(9);
  const size_t o_uv = A.take<float4>(N), o_pnm = A.take<float4>(N);
  const size_t o_models = A.take<float>((size_t)2 * n_hyp * 18), o_scores = A.take<float>((size_t)2 * n_hyp);
  const size_t o_masks = A.take<uint32_t>((size_t)2 * n_hyp * b.words);
  const size_t o_bi = A.take<int>(2), o_bs = A.take<float>(2);
  const size_t o_P3D = A.take<float>((size_t)8 * n1 * 3), o_good = A.take<uint8_t>((size_t)8 * n1), o_cos = A.take<float>((size_t)8 * N);
  const size_t o_motion = A.take<TVMotionOut>(1);
  cudaError_t ce = cudaSuccess;
  if (borrow_ws && !ctx->ws_in_use) {  // no cudaMalloc / cudaFree per call (each costs ~1-2 ms)
    if (ctx->ws_bytes < A.off) {
      if (ctx->ws_dev) cudaFree(ctx->ws_dev);
      ctx->ws_dev = nullptr; ctx->ws_bytes = 0;
      ce = cudaMalloc(&ctx->ws_dev, A.off + A.off / 4);
      if (ce == cudaSuccess) ctx->ws_bytes = A.off + A.off / 4;
    }
    if (ce == cudaSuccess) { p->dev = ctx->ws_dev; p->borrowed = true; ctx->ws_in_use = true; }
  } else {
    ce = cudaMalloc(&p->dev, A.off);
  }
  if (ce != cudaSuccess) { delete p; return fail(URMVO_ERR_CUDA, std::string("cudaMalloc tv plan: ") + cudaGetErrorString(ce)); }
  unsigned char* D = p->dev;
  b.keys1 = (float*)(D + o_k1); b.keys2 = (float*)(D + o_k2);
  b.m1 = (int*)(D + o_m1); b.m2 = (int*)(D + o_m2); b.sets = (int*)(D + o_sets); b.K = (float*)(D + o_K);
  b.pn1 = (float*)(D + o_pn1); b.pn2 = (float*)(D + o_pn2); b.T1 = (float*)(D + o_T1); b.T2 = (float*)(D + o_T2);
  b.uv = (float4*)(D + o_uv); b.pnm = (float4*)(D + o_pnm);
  b.models = (float*)(D + o_models); b.scores = (float*)(D + o_scores); b.masks = (uint32_t*)(D + o_masks);
  b.best_idx = (int*)(D + o_bi); b.best_score = (float*)(D + o_bs);
  b.P3D = (float*)(D + o_P3D); b.good = D + o_good; b.cosbuf = (float*)(D + o_cos);
  b.motion = (TVMotionOut*)(D + o_motion);
  cudaStream_t s = ctx->stream;
  cudaError_t e[7];
  e[0] = cudaMemcpyAsync(D + o_k1, keys1, (size_t)n1 * 2 * sizeof(float), cudaMemcpyHostToDevice, s);
  e[1] = cudaMemcpyAsync(D + o_k2, keys2, (size_t)n2 * 2 * sizeof(float), cudaMemcpyHostToDevice, s);
  e[2] = cudaMemcpyAsync(D + o_m1, p->m1.data(), N * sizeof(int), cudaMemcpyHostToDevice, s);
  e[3] = cudaMemcpyAsync(D + o_m2, p->m2.data(), N * sizeof(int), cudaMemcpyHostToDevice, s);
  e[4] = cudaMemcpyAsync(D + o_sets, sets, (size_t)n_hyp * 8 * sizeof(int), cudaMemcpyHostToDevice, s);
  e[5] = cudaMemcpyAsync(D + o_K, K, 9 * sizeof(float), cudaMemcpyHostToDevice, s);
  e[6] = cudaStreamSynchronize(s);
  for (cudaError_t x : e)
    if (x != cudaSuccess) { urmvo_tv_plan_destroy(p); return fail(URMVO_ERR_CUDA, std::string("tv_plan_create upload: ") + cudaGetErrorString(x)); }
  *out = p;
  return URMVO_OK;
}

extern "C" int urmvo_tv_plan_create(urmvo_ctx* ctx, urmvo_tv_plan** out, int n1, const float* keys1, int n2,
                                    const float* keys2, const int32_t* matches12, const float* K, float sigma,
                                    int n_hyp, const int32_t* sets) try {
  return tv_plan_create_impl(ctx, out, n1, keys1, n2, keys2, matches12, K, sigma, n_hyp, sets, false);
} catch (const std::exception& e) {  // no exception crosses the C ABI
  return fail(URMVO_ERR_ARG, std::string("urmvo_tv_plan_create: ") + e.what());
}

extern "C" int urmvo_tv_plan_run_ransac(urmvo_tv_plan* p) {
  if (!p) return fail(URMVO_ERR_ARG, "tv_plan_run_ransac: null plan");
  CU_TRY(cudaSetDevice(p->ctx->device));
  int nl = 0;
  cudaError_t e = launch_tv_ransac(p->b, p->sigma, p->score_mode, p->ctx->n_sm, p->ctx->stream, &nl);
  if (e != cudaSuccess) return fail(URMVO_ERR_CUDA, std::string("two-view RANSAC launch: ") + cudaGetErrorString(e));
  p->ctx->launches += nl;
  p->ransac_done = true;
  return URMVO_OK;
}

extern "C" int urmvo_tv_plan_download_hyps(urmvo_tv_plan* p, int model, float* scores, uint32_t* masks, float* models) try {
  if (!p || model < 0 || model > 1) return fail(URMVO_ERR_ARG, "tv_plan_download_hyps: bad plan / model");
  CU_TRY(cudaSetDevice(p->ctx->device));
  cudaStream_t s = p->ctx->stream;
  const TVBuffers& b = p->b;
  const size_t nh = b.n_hyp;
  if (scores) CU_TRY(cudaMemcpyAsync(scores, b.scores + model * nh, nh * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (masks) CU_TRY(cudaMemcpyAsync(masks, b.masks + model * nh * b.words, nh * b.words * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  std::vector<float> tmp;
  if (models) {
    tmp.resize(nh * 18);
    CU_TRY(cudaMemcpyAsync(tmp.data(), b.models + model * nh * 18, nh * 18 * sizeof(float), cudaMemcpyDeviceToHost, s));
  }
  CU_TRY(cudaStreamSynchronize(s));
  if (models)
    for (size_t h = 0; h < nh; h++) std::memcpy(models + h * 9, tmp.data() + h * 18, 9 * sizeof(float));
  return URMVO_OK;
} catch (const std::exception& e) {  // no exception crosses the C ABI
  return fail(URMVO_ERR_ARG, std::string("urmvo_tv_plan_download_hyps: ") + e.what());
}

extern "C" int urmvo_tv_plan_reconstruct(urmvo_tv_plan* p, float* T21, float* P3D, uint8_t* triangulated,
                                         uint8_t* mask_H, uint8_t* mask_F, urmvo_tv_stats* stats, int* success) try {
  if (!p || !T21 || !P3D || !triangulated || !success) return fail(URMVO_ERR_ARG, "tv_plan_reconstruct: null argument");
  if (!p->ransac_done) return fail(URMVO_ERR_ARG, "tv_plan_reconstruct: run_ransac has not been called");
  CU_TRY(cudaSetDevice(p->ctx->device));
  cudaStream_t s = p->ctx->stream;
  const TVBuffers& b = p->b;
  const float sigma2 = p->sigma * p->sigma;
  const float th2 = 4.0 * sigma2;  // src/epipolar_geometry.cc:489
  cudaError_t e = launch_tv_motion(b, th2, s);
  if (e != cudaSuccess) return fail(URMVO_ERR_CUDA, std::string("two-view motion launch: ") + cudaGetErrorString(e));
  p->ctx->launches++;
  TVMotionOut mo;
  int best_idx[2];
  float best_score[2];
  CU_TRY(cudaMemcpyAsync(&mo, b.motion, sizeof(mo), cudaMemcpyDeviceToHost, s));
  CU_TRY(cudaMemcpyAsync(best_idx, b.best_idx, sizeof(best_idx), cudaMemcpyDeviceToHost, s));
  CU_TRY(cudaMemcpyAsync(best_score, b.best_score, sizeof(best_score), cudaMemcpyDeviceToHost, s));
  CU_TRY(cudaStreamSynchronize(s));
  urmvo_tv_stats st;
  std::memset(&st, 0, sizeof(st));
  st.SF = best_score[0]; st.SH = best_score[1];
  st.best_F = best_idx[0]; st.best_H = best_idx[1];
  st.used_H = mo.used_H;
  st.best_motion = -1;
  std::vector<uint32_t> mw(b.words);
  for (int model = 0; model < 2; model++) {
    uint8_t* mout = model == 0 ? mask_F : mask_H;
    float* Mout = model == 0 ? st.F21 : st.H21;
    if (best_idx[model] >= 0) {
      CU_TRY(cudaMemcpyAsync(Mout, b.models + ((size_t)model * b.n_hyp + best_idx[model]) * 18, 9 * sizeof(float), cudaMemcpyDeviceToHost, s));
      if (mout) {
        CU_TRY(cudaMemcpyAsync(mw.data(), b.masks + ((size_t)model * b.n_hyp + best_idx[model]) * b.words, b.words * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaStreamSynchronize(s));
        for (int i = 0; i < b.N; i++) mout[i] = (mw[i >> 5] >> (i & 31)) & 1u;
      }
    } else if (mout) {
      std::memset(mout, 0, b.N);
    }
  }
  CU_TRY(cudaStreamSynchronize(s));
  std::memset(T21, 0, 16 * sizeof(float));
  std::memset(P3D, 0, (size_t)b.n1 * 3 * sizeof(float));
  std::memset(triangulated, 0, (size_t)b.n1);
  *success = 0;
  int winner = -1;
  float parallax[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int h = 0; h < mo.n_motion; h++) {
    st.n_good[h] = mo.n_good[h];
    // :889-895  parallax = acos(vCosParallax[idx]) * 180 / CV_PI  (float acos, double division)
    if (mo.n_good[h] > 0) parallax[h] = std::acos(mo.cos_kth[h]) * 180 / 3.1415926535897932384626433832795;
    else parallax[h] = 0;
    st.parallax[h] = parallax[h];
  }
  const float minParallax = 1.0;
  const int minTriangulated = 50;
  const int Ninl = mo.n_inl;
  if (mo.used_H == 0 && mo.n_motion == 4) {
    // _reconstruct_F acceptance, src/epipolar_geometry.cc:500-561
    const int* g = mo.n_good;
    const int maxGood = std::max(g[0], std::max(g[1], std::max(g[2], g[3])));
    const int nMinGood = std::max(static_cast<int>(0.9 * Ninl), minTriangulated);
    int nsimilar = 0;
    for (int h = 0; h < 4; h++)
      if (g[h] > 0.7 * maxGood) nsimilar++;
    if (!(maxGood < nMinGood || nsimilar > 1)) {
      for (int h = 0; h < 4; h++)
        if (maxGood == g[h]) {
          if (parallax[h] > minParallax) winner = h;
          break;
        }
    }
  } else if (mo.used_H == 1 && mo.n_motion == 8) {
    // _reconstruct_H acceptance, :695-732
    int bestGood = 0, secondBestGood = 0, bestIdx = -1;
    float bestParallax = -1;
    for (int h = 0; h < 8; h++) {
      const int nGood = mo.n_good[h];
      if (nGood > bestGood) {
        secondBestGood = bestGood;
        bestGood = nGood;
        bestIdx = h;
        bestParallax = parallax[h];
      } else if (nGood > secondBestGood) {
        secondBestGood = nGood;
      }
    }
    if (secondBestGood < 0.75 * bestGood && bestParallax >= minParallax && bestGood > minTriangulated &&
        bestGood > 0.9 * Ninl)
      winner = bestIdx;
  }
  if (winner >= 0) {
    for (int i = 0; i < 16; i++) T21[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) T21[i * 4 + j] = mo.R[winner][i * 3 + j];
      T21[i * 4 + 3] = mo.t[winner][i];
    }
    CU_TRY(cudaMemcpyAsync(P3D, b.P3D + (size_t)winner * b.n1 * 3, (size_t)b.n1 * 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(triangulated, b.good + (size_t)winner * b.n1, (size_t)b.n1, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    st.best_motion = winner;
    *success = 1;
  }
  if (stats) *stats = st;
  return URMVO_OK;
} catch (const std::exception& e) {  // no exception crosses the C ABI
  return fail(URMVO_ERR_ARG, std::string("urmvo_tv_plan_reconstruct: ") + e.what());
}

extern "C" int urmvo_tv_plan_set_score_mode(urmvo_tv_plan* p, int score_mode) {
  if (!p) return fail(URMVO_ERR_ARG, "tv_plan_set_score_mode: null plan");
  if (score_mode != URMVO_TV_SCORE_REFERENCE && score_mode != URMVO_TV_SCORE_SAMPSON)
    return fail(URMVO_ERR_ARG, "tv_plan_set_score_mode: unknown mode");
  p->score_mode = score_mode;
  p->ransac_done = false;
  return URMVO_OK;
}

extern "C" int urmvo_two_view(urmvo_ctx* ctx, int n1, const float* keys1, int n2, const float* keys2,
                              const int32_t* matches12, const float* K, float sigma, int n_hyp,
                              const int32_t* sets, float* T21, float* P3D, uint8_t* triangulated,
                              uint8_t* mask_H, uint8_t* mask_F, urmvo_tv_stats* stats, int* success) {
  return urmvo_two_view_scored(ctx, n1, keys1, n2, keys2, matches12, K, sigma, n_hyp, sets, URMVO_TV_SCORE_REFERENCE,
                               T21, P3D, triangulated, mask_H, mask_F, stats, success);
}

extern "C" int urmvo_two_view_scored(urmvo_ctx* ctx, int n1, const float* keys1, int n2, const float* keys2,
                                     const int32_t* matches12, const float* K, float sigma, int n_hyp,
                                     const int32_t* sets, int score_mode, float* T21, float* P3D,
                                     uint8_t* triangulated, uint8_t* mask_H, uint8_t* mask_F,
                                     urmvo_tv_stats* stats, int* success) {
  urmvo_tv_plan* p = nullptr;
  int rc = tv_plan_create_impl(ctx, &p, n1, keys1, n2, keys2, matches12, K, sigma, n_hyp, sets, true);
  if (rc != URMVO_OK) return rc;
  rc = urmvo_tv_plan_set_score_mode(p, score_mode);
  if (rc != URMVO_OK) { urmvo_tv_plan_destroy(p); return rc; }
  rc = urmvo_tv_plan_run_ransac(p);
  if (rc == URMVO_OK) rc = urmvo_tv_plan_reconstruct(p, T21, P3D, triangulated, mask_H, mask_F, stats, success);
  urmvo_tv_plan_destroy(p);
  return rc;
}

// =====================================================================================  device-resident map
//
// SURVEY.md §8f row 3.  The reference rebuilds the local-BA problem from shared_ptr graphs every keyframe
// (src/mapping.cc:335-535) and copies every pose, point and keypoint into fresh containers.  Here keyframe
// poses, mappoint positions and observations stay in HBM across keyframes, addressed by the caller's ids; a
// window is selected by id lists, its values are gathered on the device, and the optimised poses / points go back
// into the map on the device.  The host keeps only the INDEX structure (id -> slot, per-point observation lists).

struct urmvo_map {
  urmvo_ctx* ctx = nullptr;
  double intr[4] = {0, 0, 0, 0};
  // id -> slot: frame / mappoint ids are small consecutive integers in the reference (a flat table, no hashing on
  // the per-keyframe path); ids outside [0, kFlatIds) go through a hash map
  static constexpr int kFlatIds = 1 << 22;
  struct IdTable {
    std::vector<int> flat;
    std::unordered_map<int, int> big;
    int find(int id) const {
      if ((unsigned)id < (unsigned)kFlatIds) return (size_t)id < flat.size() ? flat[id] : -1;
      auto it = big.find(id);
      return it == big.end() ? -1 : it->second;
    }
    void set(int id, int slot) {
      if ((unsigned)id < (unsigned)kFlatIds) {
        if ((size_t)id >= flat.size()) flat.resize(std::max<size_t>((size_t)id + 1, 2 * flat.size()), -1);
        flat[id] = slot;
      } else {
        big[id] = slot;
      }
    }
  };
  IdTable kf_slot, pt_slot;
  std::vector<int> kf_ids, pt_ids;                 // slot -> id
  std::vector<std::vector<int>> pt_obs;            // point slot -> observation slots (insertion order)
  std::vector<int> obs_kf, obs_pt;                 // observation slot -> keyframe / point slot
  std::vector<uint8_t> obs_alive;
  double* d_kf = nullptr; double* d_pt = nullptr; double* d_uv = nullptr;
  size_t cap_kf = 0, cap_pt = 0, cap_obs = 0;
  std::vector<int> kf_local;                       // scratch: keyframe slot -> index in the current window or -1
  std::vector<int> w_kslots, w_pslots, w_oslots, w_cam, w_pt;  // scratch of the window assembly (no reallocation per keyframe)
};

namespace {

// grow a slot-addressed device array (width W doubles) to hold `need` slots, keeping its contents
int map_reserve(urmvo_map* m, double** arr, size_t* cap, size_t used, size_t need, int W) {
  if (need <= *cap) return URMVO_OK;
  size_t nc = std::max<size_t>(need, std::max<size_t>(2 * *cap, 1024));
  double* nd = nullptr;
  if (cudaMalloc(&nd, nc * W * sizeof(double)) != cudaSuccess) return fail(URMVO_ERR_CUDA, "map: cudaMalloc failed");
  cudaStream_t s = m->ctx->stream;
  if (*arr && used) {
    if (cudaMemcpyAsync(nd, *arr, used * W * sizeof(double), cudaMemcpyDeviceToDevice, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) {
      cudaFree(nd);
      return fail(URMVO_ERR_CUDA, "map: device copy failed");
    }
  }
  if (*arr) cudaFree(*arr);
  *arr = nd;
  *cap = nc;
  return URMVO_OK;
}

// values[n * W] into the slots `slots` of a device array: one staged upload + one scatter kernel
int map_set_rows(urmvo_map* m, double* darr, const std::vector<int>& slots, const double* vals, int W) {
  const size_t n = slots.size();
  if (!n) return URMVO_OK;
  urmvo_ctx* ctx = m->ctx;
  const size_t vb = n * W * sizeof(double), ib = n * sizeof(int);
  if (ctx->ensure_pinned(vb + ib)) return fail(URMVO_ERR_CUDA, "map: cudaMallocHost failed");
  if (ctx->ws_in_use) return fail(URMVO_ERR_ARG, "map: the context workspace is in use by another call");
  if (ctx->ws_bytes < vb + ib + 256) {
    if (ctx->ws_dev) cudaFree(ctx->ws_dev);
    ctx->ws_dev = nullptr; ctx->ws_bytes = 0;
    const size_t want = std::max(vb + ib + 256, (size_t)1 << 22);
    if (cudaMalloc(&ctx->ws_dev, want) != cudaSuccess) return fail(URMVO_ERR_CUDA, "map: cudaMalloc failed");
    ctx->ws_bytes = want;
  }
  unsigned char* P = (unsigned char*)ctx->pinned;
  std::memcpy(P, vals, vb);
  std::memcpy(P + vb, slots.data(), ib);
  cudaStream_t s = ctx->stream;
  CU_TRY(cudaMemcpyAsync(ctx->ws_dev, P, vb + ib, cudaMemcpyHostToDevice, s));
  CU_TRY(launch_map_set(darr, (const double*)ctx->ws_dev, (const int*)(ctx->ws_dev + vb), (int)n, W, s));
  ctx->launches++;
  CU_TRY(cudaStreamSynchronize(s));  // the pinned buffer and the workspace are free again
  return URMVO_OK;
}

int map_get_rows(urmvo_map* m, const double* darr, const std::vector<int>& slots, double* vals, int W) {
  const size_t n = slots.size();
  if (!n) return URMVO_OK;
  urmvo_ctx* ctx = m->ctx;
  const size_t vb = n * W * sizeof(double), ib = n * sizeof(int);
  if (ctx->ensure_pinned(vb + ib)) return fail(URMVO_ERR_CUDA, "map: cudaMallocHost failed");
  if (ctx->ws_in_use) return fail(URMVO_ERR_ARG, "map: the context workspace is in use by another call");
  if (ctx->ws_bytes < vb + ib + 256) {
    if (ctx->ws_dev) cudaFree(ctx->ws_dev);
    ctx->ws_dev = nullptr; ctx->ws_bytes = 0;
    const size_t want = std::max(vb + ib + 256, (size_t)1 << 22);
    if (cudaMalloc(&ctx->ws_dev, want) != cudaSuccess) return fail(URMVO_ERR_CUDA, "map: cudaMalloc failed");
    ctx->ws_bytes = want;
  }
  unsigned char* P = (unsigned char*)ctx->pinned;
  std::memcpy(P + vb, slots.data(), ib);
  cudaStream_t s = ctx->stream;
  CU_TRY(cudaMemcpyAsync(ctx->ws_dev + vb, P + vb, ib, cudaMemcpyHostToDevice, s));
  CU_TRY(launch_map_get((double*)ctx->ws_dev, darr, (const int*)(ctx->ws_dev + vb), (int)n, W, s));
  ctx->launches++;
  CU_TRY(cudaMemcpyAsync(P, ctx->ws_dev, vb, cudaMemcpyDeviceToHost, s));
  CU_TRY(cudaStreamSynchronize(s));
  std::memcpy(vals, P, vb);
  return URMVO_OK;
}

}  // namespace

extern "C" int urmvo_map_create(urmvo_ctx* ctx, urmvo_map** out, const double* intr) {
  try {
    if (!ctx || !out || !intr) return fail(URMVO_ERR_ARG, "map_create: null argument");
    urmvo_map* m = new urmvo_map();
    m->ctx = ctx;
    for (int k = 0; k < 4; k++) m->intr[k] = intr[k];
    *out = m;
    return URMVO_OK;
  } catch (const std::exception& e) {
    return fail(URMVO_ERR_ARG, std::string("map_create: ") + e.what());
  }
}

extern "C" void urmvo_map_destroy(urmvo_map* m) {
  if (!m) return;
  cudaSetDevice(m->ctx->device);
  if (m->d_kf) cudaFree(m->d_kf);
  if (m->d_pt) cudaFree(m->d_pt);
  if (m->d_uv) cudaFree(m->d_uv);
  delete m;
}

extern "C" int urmvo_map_set_keyframes(urmvo_map* m, int n, const int32_t* ids, const double* poses) {
  try {
    if (!m || n < 0 || (n && (!ids || !poses))) return fail(URMVO_ERR_ARG, "map_set_keyframes: bad argument");
    CU_TRY(cudaSetDevice(m->ctx->device));
    std::vector<int> slots(n);
    for (int i = 0; i < n; i++) {
      int sl = m->kf_slot.find(ids[i]);
      if (sl < 0) {
        sl = (int)m->kf_ids.size();
        m->kf_slot.set(ids[i], sl);
        m->kf_ids.push_back(ids[i]);
      }
      slots[i] = sl;
    }
    if (int rc = map_reserve(m, &m->d_kf, &m->cap_kf, std::min(m->kf_ids.size(), m->cap_kf), m->kf_ids.size(), 7)) return rc;
    return map_set_rows(m, m->d_kf, slots, poses, 7);
  } catch (const std::exception& e) {
    return fail(URMVO_ERR_ARG, std::string("map_set_keyframes: ") + e.what());
  }
}

extern "C" int urmvo_map_set_points(urmvo_map* m, int n, const int32_t* ids, const double* xyz) {
  try {
    if (!m || n < 0 || (n && (!ids || !xyz))) return fail(URMVO_ERR_ARG, "map_set_points: bad argument");
    CU_TRY(cudaSetDevice(m->ctx->device));
    std::vector<int> slots(n);
    for (int i = 0; i < n; i++) {
      int sl = m->pt_slot.find(ids[i]);
      if (sl < 0) {
        sl = (int)m->pt_ids.size();
        m->pt_slot.set(ids[i], sl);
        m->pt_ids.push_back(ids[i]);
        m->pt_obs.emplace_back();
      }
      slots[i] = sl;
    }
    if (int rc = map_reserve(m, &m->d_pt, &m->cap_pt, std::min(m->pt_ids.size(), m->cap_pt), m->pt_ids.size(), 3)) return rc;
    return map_set_rows(m, m->d_pt, slots, xyz, 3);
  } catch (const std::exception& e) {
    return fail(URMVO_ERR_ARG, std::string("map_set_points: ") + e.what());
  }
}

extern "C" int urmvo_map_add_observations(urmvo_map* m, int n, const int32_t* kf_ids, const int32_t* pt_ids, const double* uv) {
  try {
    if (!m || n < 0 || (n && (!kf_ids || !pt_ids || !uv))) return fail(URMVO_ERR_ARG, "map_add_observations: bad argument");
    CU_TRY(cudaSetDevice(m->ctx->device));
    for (int i = 0; i < n; i++)
      if (m->kf_slot.find(kf_ids[i]) < 0 || m->pt_slot.find(pt_ids[i]) < 0)
        return fail(URMVO_ERR_ARG, "map_add_observations: unknown keyframe or mappoint id");
    const size_t first = m->obs_kf.size();
    if (int rc = map_reserve(m, &m->d_uv, &m->cap_obs, std::min(first, m->cap_obs), first + n, 2)) return rc;
    std::vector<int> slots(n);
    for (int i = 0; i < n; i++) {
      const int ks = m->kf_slot.find(kf_ids[i]), ps = m->pt_slot.find(pt_ids[i]);
      slots[i] = (int)(first + i);
      m->obs_kf.push_back(ks);
      m->obs_pt.push_back(ps);
      m->obs_alive.push_back(1);
      m->pt_obs[ps].push_back(slots[i]);
    }
    return map_set_rows(m, m->d_uv, slots, uv, 2);
  } catch (const std::exception& e) {
    return fail(URMVO_ERR_ARG, std::string("map_add_observations: ") + e.what());
  }
}

extern "C" int urmvo_map_remove_observations(urmvo_map* m, int n, const int32_t* kf_ids, const int32_t* pt_ids) {
  try {
    if (!m || n < 0 || (n && (!kf_ids || !pt_ids))) return fail(URMVO_ERR_ARG, "map_remove_observations: bad argument");
    for (int i = 0; i < n; i++) {
      const int ks = m->kf_slot.find(kf_ids[i]), ps = m->pt_slot.find(pt_ids[i]);
      if (ks < 0 || ps < 0) continue;  // like erasing an absent observer: no-op
      std::vector<int>& lst = m->pt_obs[ps];
      for (size_t k = 0; k < lst.size(); k++)
        if (m->obs_kf[lst[k]] == ks && m->obs_alive[lst[k]]) {
          m->obs_alive[lst[k]] = 0;
          lst.erase(lst.begin() + k);
          break;
        }
    }
    return URMVO_OK;
  } catch (const std::exception& e) {
    return fail(URMVO_ERR_ARG, std::string("map_remove_observations: ") + e.what());
  }
}

extern "C" int urmvo_map_get_keyframes(urmvo_map* m, int n, const int32_t* ids, double* poses) {
  try {
    if (!m || n < 0 || (n && (!ids || !poses))) return fail(URMVO_ERR_ARG, "map_get_keyframes: bad argument");
    CU_TRY(cudaSetDevice(m->ctx->device));
    std::vector<int> slots(n);
    for (int i = 0; i < n; i++) {
      slots[i] = m->kf_slot.find(ids[i]);
      if (slots[i] < 0) return fail(URMVO_ERR_ARG, "map_get_keyframes: unknown id");
    }
    return map_get_rows(m, m->d_kf, slots, poses, 7);
  } catch (const std::exception& e) {
    return fail(URMVO_ERR_ARG, std::string("map_get_keyframes: ") + e.what());
  }
}

extern "C" int urmvo_map_get_points(urmvo_map* m, int n, const int32_t* ids, double* xyz) {
  try {
    if (!m || n < 0 || (n && (!ids || !xyz))) return fail(URMVO_ERR_ARG, "map_get_points: bad argument");
    CU_TRY(cudaSetDevice(m->ctx->device));
    std::vector<int> slots(n);
    for (int i = 0; i < n; i++) {
      slots[i] = m->pt_slot.find(ids[i]);
      if (slots[i] < 0) return fail(URMVO_ERR_ARG, "map_get_points: unknown id");
    }
    return map_get_rows(m, m->d_pt, slots, xyz, 3);
  } catch (const std::exception& e) {
    return fail(URMVO_ERR_ARG, std::string("map_get_points: ") + e.what());
  }
}

extern "C" int urmvo_map_local_ba(urmvo_map* m, int n_kf, const int32_t* kf_ids, const uint8_t* kf_fixed, int n_pt,
                                  const int32_t* pt_ids, double chi2_thr, int it0, int it1, const urmvo_ba_options* opts,
                                  int32_t max_obs, int32_t* n_obs, int32_t* obs_kf, int32_t* obs_pt, uint8_t* inlier,
                                  urmvo_ba_stats* stats) {
  try {
    if (!m || n_kf <= 0 || n_pt < 0 || !kf_ids || !kf_fixed || (n_pt && !pt_ids) || !n_obs)
      return fail(URMVO_ERR_ARG, "map_local_ba: bad argument");
    urmvo_ctx* ctx = m->ctx;
    CU_TRY(cudaSetDevice(ctx->device));
    *n_obs = 0;
    // ---- the window: index structure only, from the host mirror
    std::vector<int>&kslots = m->w_kslots, &pslots = m->w_pslots, &oslots = m->w_oslots, &cam = m->w_cam, &pt = m->w_pt;
    kslots.assign(n_kf, 0); pslots.clear(); oslots.clear(); cam.clear(); pt.clear();
    m->kf_local.assign(m->kf_ids.size(), -1);
    for (int i = 0; i < n_kf; i++) {
      const int sl = m->kf_slot.find(kf_ids[i]);
      if (sl < 0) return fail(URMVO_ERR_ARG, "map_local_ba: unknown keyframe id");
      if (m->kf_local[sl] >= 0) return fail(URMVO_ERR_ARG, "map_local_ba: keyframe listed twice");
      m->kf_local[sl] = i;
      kslots[i] = sl;
    }
    for (int l = 0; l < n_pt; l++) {
      const int ps = m->pt_slot.find(pt_ids[l]);
      if (ps < 0) return fail(URMVO_ERR_ARG, "map_local_ba: unknown mappoint id");
      const std::vector<int>& lst = m->pt_obs[ps];
      const size_t first = oslots.size();
      for (int os : lst) {
        const int c = m->kf_local[m->obs_kf[os]];
        if (c < 0) continue;
        oslots.push_back(os);
        cam.push_back(c);
        pt.push_back((int)pslots.size());
      }
      // mono mappoints enter the optimisation with more than one constraint (reference src/mapping.cc:463-465)
      if (oslots.size() - first < 2) { oslots.resize(first); cam.resize(first); pt.resize(first); continue; }
      pslots.push_back(ps);
    }
    const int No = (int)oslots.size(), Np = (int)pslots.size();
    if (No == 0) return URMVO_OK;  // empty graph: nothing to optimise (g2o: silent no-op)
    if (No > max_obs || !obs_kf || !obs_pt || !inlier)
      return fail(URMVO_ERR_ARG, "map_local_ba: observation outputs too small (max_obs) or null");
    const int32_t co[2] = {0, n_kf}, po[2] = {0, Np}, oo[2] = {0, No};
    MapGather mg{m->d_kf, m->d_pt, m->d_uv, kslots.data(), pslots.data(), oslots.data()};
    urmvo_ba_plan* p = nullptr;
    int rc = ba_plan_create_impl(ctx, &p, 1, co, po, oo, nullptr, kf_fixed, nullptr, nullptr, cam.data(), pt.data(), m->intr,
                                 chi2_thr, it0, it1, opts, false, nullptr, /*borrow_ws=*/true, nullptr, 0.0, &mg);
    if (rc != URMVO_OK) return rc;
    rc = urmvo_ba_plan_run(p);
    if (rc == URMVO_OK) {
      const int* dg = (const int*)(p->dev + p->off_gather);
      const cudaError_t e = launch_map_scatter(m->d_kf, m->d_pt, (const double*)(p->dev + p->off_pose_out),
                                               (const double*)(p->dev + p->off_pts_out), (const int*)(p->dev + p->off_cam_free),
                                               dg, dg + n_kf, n_kf, Np, ctx->stream);
      ctx->launches++;
      if (e != cudaSuccess) rc = fail(URMVO_ERR_CUDA, std::string("map_local_ba scatter: ") + cudaGetErrorString(e));
    }
    if (rc == URMVO_OK) rc = urmvo_ba_plan_download(p, nullptr, nullptr, inlier, stats);
    urmvo_ba_plan_destroy(p);
    if (rc != URMVO_OK) return rc;
    for (int o = 0; o < No; o++) {
      obs_kf[o] = m->kf_ids[m->obs_kf[oslots[o]]];
      obs_pt[o] = m->pt_ids[m->obs_pt[oslots[o]]];
    }
    *n_obs = No;
    return URMVO_OK;
  } catch (const std::exception& e) {
    if (m && m->ctx) m->ctx->ws_in_use = false;
    return fail(URMVO_ERR_ARG, std::string("map_local_ba: ") + e.what());
  }
}
