// C ABI of include/c2a_b200.h: model upload, host half of the motion model, batch launch of the
// persistent CCD kernel (c2a_solve.cuh), unit-test hooks.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo (see c2a_b200/build.py).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/c2a_b200.h"
#include "c2a_solve.cuh"
#include "c2a_wide.cuh"
#include "c2a_contact.cuh"
#include "c2a_translation.cuh"
#include "c2a_distance.cuh"

namespace c2a {

// ---- host side -----------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

static int fail(int code, const std::string &msg)
{
  g_err = msg;
  return code;
}
#define CUDA_TRY(x)                                                                                      \
  do {                                                                                                   \
    cudaError_t e_ = (x);                                                                                \
    if (e_ != cudaSuccess) return fail(C2A_B200_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); \
  } while (0)

// Every entry runs on its models' device and puts the caller's current device back when it returns.
struct DeviceGuard
{
  int prev = -1;
  cudaError_t err;
  explicit DeviceGuard(int device)
  {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    err = cudaSetDevice(device);
    if (err != cudaSuccess) cudaGetLastError();  // (not sticky: do not leave it for the next launch check to trip over)
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define ON_DEVICE(dev)                                                                                           \
  DeviceGuard device_guard_(dev);                                                                                \
  if (device_guard_.err != cudaSuccess) return fail(C2A_B200_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(device_guard_.err))

}  // namespace c2a

struct c2a_b200_model
{
  int device;
  int n_nodes, n_tris, depth;
  double root_ang_radius;
  double *geom, *rloc, *tris;
  c2a::NodeMeta *meta;
  int *tri_vidx;  // may be NULL
  double *obb;    // [n_nodes][OBB_STRIDE]: d(3), To(3); NULL when the model came without OBB data (C2A_Collide refuses it)
};

using namespace c2a;

// ---- host side of the motion model ----------------------------------------------------------
// CInterpMotion ctor + CInterpMotion_Linear::velocity + LinearAngularVelocity
// (/root/reference/C2A/src/InterpMotion.cpp:148-168, 486-491, 228-270) for one object:
// pose = R0(9) T0(3) R1(9) T1(3)  ->  rec = R0(9) T0(3) cv(3) axis(3) w qs(4) pad.
// Runs on the host so that acos()/sqrt() are the host libm's, like the reference's.
static void motion_record_from_pose(const double *pose, double *rec)
{
  double qs[4], qt[4], q0[4], qd[4];
  for (int i = 0; i < 12; i++) rec[i] = pose[i];
  for (int i = 0; i < 3; i++) rec[12 + i] = pose[21 + i] - pose[9 + i];
  quat_from_matrix(qs, pose);
  quat_from_matrix(qt, pose + 12);
  q0[0] = -qs[0]; q0[1] = -qs[1]; q0[2] = -qs[2]; q0[3] = qs[3];
  quat_mul(qd, q0, qt);
  const double s = 1 < qd[3] ? 1 : qd[3];
  const double sign = s < 0 ? -1 : 1;
  const double a = (fabs(s - 1) <= 1e-40 || fabs(s + 1) <= 1e-40) ? (2 * sign)
                                                                  : (sign * acos(2 * s * s - 1) / sqrt(1 - s * s));
  const double tangent[3] = {a * qd[0], a * qd[1], a * qd[2]};
  rec[18] = sqrt(tangent[0] * tangent[0] + tangent[1] * tangent[1] + tangent[2] * tangent[2]);
  const double len = tangent[0] * tangent[0] + tangent[1] * tangent[1] + tangent[2] * tangent[2];
  if (len < (double)0.00000001f) { rec[15] = 1.0; rec[16] = 0.0; rec[17] = 0.0; }
  else
  {
    const double inv = 1.0 / sqrt(len);
    rec[15] = tangent[0] * inv; rec[16] = tangent[1] * inv; rec[17] = tangent[2] * inv;
  }
  rec[19] = qs[0]; rec[20] = qs[1]; rec[21] = qs[2]; rec[22] = qs[3];
  rec[23] = 0.0;  // CInterpMotion::m_toc_delta; 0 = the batch's tol_d (what C2A_Solve sets, C2A.cpp:2384-2388)
}

// Claim order for the persistent kernel: queries expected to run long first, so that their sequential
// CA chains start early instead of forming the tail of the launch.  Results do not depend on the order.
// Cost proxy: how much of the conservative motion bound is rotation rather than closing translation,
//   key = (w1*r1 + w2*r2 + |cv1 - cv2|) / |cv1 - cv2|   (r = angular radius of the root BV)
// -- a small closing speed relative to the bound means small CA steps, i.e. many of them.  Counting sort
// into 256 log-spaced buckets, descending, stable.
static void schedule_order(const double *motions, int64_t n, double ra, double rb, int32_t *order)
{
  std::vector<uint8_t> bucket((size_t)n);
  // the keys (one pass over the motion records: 384 B apart, memory-bound) on several threads, the counting sort on one
  int n_threads = (int)std::thread::hardware_concurrency();
  if (n_threads < 1) n_threads = 1;
  if ((int64_t)n_threads > n / 65536 + 1) n_threads = (int)(n / 65536 + 1);
  std::vector<std::array<int64_t, 256>> part((size_t)n_threads);
  auto keys = [&](int t, int64_t lo, int64_t hi) {
    std::array<int64_t, 256> &c = part[(size_t)t];
    c.fill(0);
    for (int64_t i = lo; i < hi; i++)
    {
      const double *m = motions + 48 * i;
      const double dx = m[12] - m[36], dy = m[13] - m[37], dz = m[14] - m[38];
      const double dcv = sqrt(dx * dx + dy * dy + dz * dz);
      const double key = (m[18] * ra + m[42] * rb + dcv) / (dcv > 1e-300 ? dcv : 1e-300);
      double l = 24.0 * log2(key > 1.0 ? key : 1.0);
      const int b = 255 - (l < 255.0 ? (int)l : 255);  // bucket 0 = largest key
      bucket[i] = (uint8_t)b;
      c[(size_t)b]++;
    }
  };
  const int64_t per = (n + n_threads - 1) / n_threads;
  if (n_threads == 1) keys(0, 0, n);
  else
  {
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++)
    {
      const int64_t lo = t * per, hi = std::min<int64_t>(n, lo + per);
      if (lo < hi) th.emplace_back(keys, t, lo, hi); else part[(size_t)t].fill(0);
    }
    for (auto &t : th) t.join();
  }
  int64_t count[257] = {0};
  for (int t = 0; t < n_threads; t++)
    for (int b = 0; b < 256; b++) count[b + 1] += part[(size_t)t][(size_t)b];
  for (int b = 0; b < 256; b++) count[b + 1] += count[b];
  for (int64_t i = 0; i < n; i++) order[count[bucket[i]]++] = (int32_t)i;
}

static void motions_from_poses_mt(const double *poses, int64_t n, double *motions, int n_threads)
{
  auto work = [=](int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; i++)
    {
      motion_record_from_pose(poses + 48 * i, motions + 48 * i);
      motion_record_from_pose(poses + 48 * i + 24, motions + 48 * i + 24);
    }
  };
  if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
  if (n_threads < 1) n_threads = 1;
  if ((int64_t)n_threads > n / 4096 + 1) n_threads = (int)(n / 4096 + 1);
  if (n_threads == 1) { work(0, n); return; }
  std::vector<std::thread> th;
  const int64_t per = (n + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; t++)
  {
    const int64_t lo = t * per, hi = std::min<int64_t>(n, lo + per);
    if (lo < hi) th.emplace_back(work, lo, hi);
  }
  for (auto &t : th) t.join();
}

extern "C" {

const char *c2a_b200_last_error(void) { return g_err.c_str(); }
int64_t c2a_b200_launch_count(void) { return g_launches.load(); }

int c2a_b200_device_count(int32_t *count)
{
  if (!count) return fail(C2A_B200_ERR_ARG, "count is NULL");
  int n = 0;
  CUDA_TRY(cudaGetDeviceCount(&n));
  *count = n;
  return C2A_B200_OK;
}

int c2a_b200_model_upload(const c2a_b200_bvh *bvh, int32_t device, c2a_b200_model **out)
{
  if (!bvh || !out) return fail(C2A_B200_ERR_ARG, "bvh/out is NULL");
  const int n = bvh->n_nodes, nt = bvh->n_tris;
  if (n <= 0 || nt <= 0 || !bvh->R || !bvh->Tr || !bvh->l || !bvh->r || !bvh->R_loc || !bvh->ang_radius ||
      !bvh->first_child || !bvh->tris)
    return fail(C2A_B200_ERR_ARG, "empty model or NULL array");

  // validate the topology and measure the depth (bounds the traversal stack)
  int depth = 0;
  {
    std::vector<std::pair<int, int>> todo;
    todo.push_back({0, 0});
    long long visited = 0;
    while (!todo.empty())
    {
      auto [i, d] = todo.back();
      todo.pop_back();
      if (++visited > n) return fail(C2A_B200_ERR_ARG, "BVH is not a tree");
      if (d > depth) depth = d;
      const int fc = bvh->first_child[i];
      if (fc < 0)
      {
        if (-fc - 1 >= nt) return fail(C2A_B200_ERR_ARG, "leaf triangle index out of range");
      }
      else
      {
        if (fc + 1 >= n || fc <= i) return fail(C2A_B200_ERR_ARG, "child index out of range");
        todo.push_back({fc, d + 1});
        todo.push_back({fc + 1, d + 1});
      }
    }
  }

  std::vector<double> geom((size_t)n * GEOM_STRIDE);
  std::vector<NodeMeta> meta(n);
  for (int i = 0; i < n; i++)
  {
    double *g = &geom[(size_t)i * GEOM_STRIDE];
    memcpy(g, bvh->R + 9 * (size_t)i, 9 * sizeof(double));
    memcpy(g + 9, bvh->Tr + 3 * (size_t)i, 3 * sizeof(double));
    g[12] = bvh->l[2 * (size_t)i]; g[13] = bvh->l[2 * (size_t)i + 1];
    g[14] = bvh->r[i];
    g[15] = bvh->ang_radius[i];
    // PQP BV::GetSize(), RSS form; sqrt/mul/add are correctly rounded on both sides, so precomputing is exact
    volatile double l0 = g[12] * g[12], l1 = g[13] * g[13];
    volatile double s = l0 + l1;
    meta[i].size = sqrt(s) + 2 * g[14];
    meta[i].first_child = bvh->first_child[i];
    meta[i].pad = 0;
  }

  // R_loc and the triangles as padded 80-byte records (128-bit loads on the device)
  std::vector<double> rloc((size_t)n * RLOC_STRIDE, 0.0), tris((size_t)nt * TRI_STRIDE, 0.0);
  for (int i = 0; i < n; i++) memcpy(&rloc[(size_t)i * RLOC_STRIDE], bvh->R_loc + 9 * (size_t)i, 9 * sizeof(double));
  for (int i = 0; i < nt; i++) memcpy(&tris[(size_t)i * TRI_STRIDE], bvh->tris + 9 * (size_t)i, 9 * sizeof(double));

  ON_DEVICE(device);
  c2a_b200_model *m = new c2a_b200_model();
  m->device = device; m->n_nodes = n; m->n_tris = nt; m->depth = depth;
  m->root_ang_radius = bvh->ang_radius[0];
  m->geom = m->rloc = m->tris = m->obb = nullptr; m->meta = nullptr; m->tri_vidx = nullptr;
  cudaError_t e;
  if ((e = cudaMalloc(&m->geom, geom.size() * sizeof(double))) != cudaSuccess ||
      (e = cudaMalloc(&m->rloc, rloc.size() * sizeof(double))) != cudaSuccess ||
      (e = cudaMalloc(&m->meta, (size_t)n * sizeof(NodeMeta))) != cudaSuccess ||
      (e = cudaMalloc(&m->tris, tris.size() * sizeof(double))) != cudaSuccess ||
      (e = cudaMemcpy(m->geom, geom.data(), geom.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemcpy(m->rloc, rloc.data(), rloc.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemcpy(m->meta, meta.data(), (size_t)n * sizeof(NodeMeta), cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemcpy(m->tris, tris.data(), tris.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess ||
      (bvh->tri_vidx && ((e = cudaMalloc(&m->tri_vidx, (size_t)nt * 3 * sizeof(int))) != cudaSuccess ||
                         (e = cudaMemcpy(m->tri_vidx, bvh->tri_vidx, (size_t)nt * 3 * sizeof(int), cudaMemcpyHostToDevice)) != cudaSuccess)))
  {
    c2a_b200_model_free(m);
    return fail(C2A_B200_ERR_CUDA, std::string("model upload: ") + cudaGetErrorString(e));
  }
  if (bvh->obb_d && bvh->obb_To)
  {
    std::vector<double> obb((size_t)n * OBB_STRIDE, 0.0);
    for (int i = 0; i < n; i++)
    {
      memcpy(&obb[(size_t)i * OBB_STRIDE], bvh->obb_d + 3 * (size_t)i, 3 * sizeof(double));
      memcpy(&obb[(size_t)i * OBB_STRIDE + 3], bvh->obb_To + 3 * (size_t)i, 3 * sizeof(double));
    }
    if ((e = cudaMalloc(&m->obb, obb.size() * sizeof(double))) != cudaSuccess ||
        (e = cudaMemcpy(m->obb, obb.data(), obb.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess)
    {
      c2a_b200_model_free(m);
      return fail(C2A_B200_ERR_CUDA, std::string("model upload (OBB data): ") + cudaGetErrorString(e));
    }
  }
  *out = m;
  return C2A_B200_OK;
}

int c2a_b200_model_free(c2a_b200_model *m)
{
  if (!m) return C2A_B200_OK;
  DeviceGuard device_guard_(m->device);
  cudaFree(m->geom); cudaFree(m->rloc); cudaFree(m->meta); cudaFree(m->tris); cudaFree(m->tri_vidx); cudaFree(m->obb);
  delete m;
  return C2A_B200_OK;
}

int c2a_b200_model_info(const c2a_b200_model *m, int32_t *device, int32_t *n_nodes, int32_t *n_tris, int32_t *depth)
{
  if (!m) return fail(C2A_B200_ERR_ARG, "model is NULL");
  if (device) *device = m->device;
  if (n_nodes) *n_nodes = m->n_nodes;
  if (n_tris) *n_tris = m->n_tris;
  if (depth) *depth = m->depth;
  return C2A_B200_OK;
}

static int g_stats_device = -1, g_trace_device = -1, g_wide_stats_device = -1;  // the device each debug buffer lives on
static unsigned long long *g_stats_dev = nullptr;  // phase statistics (c2a_b200_phase_stats), off by default
static unsigned long long *g_trace_dev = nullptr;  // per-query claim / finish times (c2a_b200_query_trace), off by default
static unsigned long long *g_wide_stats_dev = nullptr;  // counters of c2a_wide_kernel (c2a_b200_wide_stats), off by default

// CUDA events round the kernels of this thread's last launch_batch (c2a_b200_kernel_times): measurement aid
struct KernelEvents { int device = -1; cudaEvent_t e[4]; bool recorded = false; };
static thread_local KernelEvents g_kev;

static long long env_ll(const char *name, long long dflt)
{
  const char *v = getenv(name);
  return (v && *v) ? atoll(v) : dflt;
}
static int64_t g_trace_n = 0;

// Per host thread and device: a lowest-priority side stream for the wide kernel's early launch, with the two events that
// fork it from and join it to the caller's stream.
namespace {
struct SideStream { int device = -1; cudaStream_t of = nullptr; cudaStream_t s = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
struct SideStreamSet
{
  std::vector<SideStream> v;
  ~SideStreamSet()
  {
    for (auto &c : v)
    {
      if (!c.s) continue;
      int cur = -1;
      if (cudaGetDevice(&cur) != cudaSuccess || cudaSetDevice(c.device) != cudaSuccess) { cudaGetLastError(); continue; }
      cudaStreamDestroy(c.s); cudaEventDestroy(c.fork); cudaEventDestroy(c.join);
      cudaSetDevice(cur);
    }
  }
};
thread_local SideStreamSet g_side;
// One side stream per (device, caller's stream), so that launches a thread issues on several streams do not queue their
// early wide kernels behind one another; a handful at most (a thread that cycles through more streams goes without).
SideStream *side_stream(int device, cudaStream_t of)  // the caller has made `device` current; nullptr: no early launch
{
  for (auto &c : g_side.v) if (c.device == device && c.of == of) return c.s ? &c : nullptr;
  if (g_side.v.size() >= 8) return nullptr;
  SideStream c; c.device = device; c.of = of;
  int least = 0, greatest = 0;
  if (cudaDeviceGetStreamPriorityRange(&least, &greatest) == cudaSuccess &&
      cudaStreamCreateWithPriority(&c.s, cudaStreamNonBlocking, least) == cudaSuccess)
  {
    if (cudaEventCreateWithFlags(&c.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c.join, cudaEventDisableTiming) != cudaSuccess)
    {
      if (c.fork) cudaEventDestroy(c.fork);
      cudaStreamDestroy(c.s); c.s = nullptr;
    }
  }
  if (!c.s) cudaGetLastError();
  g_side.v.push_back(c);
  return g_side.v.back().s ? &g_side.v.back() : nullptr;
}
}  // namespace

// The library's own stream-ordered memory pool, one per device: per-launch scratch (traversal stacks, hand-over lists,
// staging arenas) is kept between launches without touching the release threshold of the device's DEFAULT pool, which
// belongs to the caller.  Falls back to the default pool (threshold untouched) if a pool cannot be created.
namespace {
cudaMemPool_t device_pool(int device)
{
  static std::mutex m;
  static std::vector<std::pair<int, cudaMemPool_t>> pools;
  std::lock_guard<std::mutex> lk(m);
  for (auto &p : pools) if (p.first == device) return p.second;
  if (getenv("C2A_B200_DEFAULT_POOL")) { pools.push_back({device, nullptr}); return nullptr; }   // development aid
  cudaMemPoolProps props;
  memset(&props, 0, sizeof(props));
  props.allocType = cudaMemAllocationTypePinned;
  props.handleTypes = cudaMemHandleTypeNone;
  props.location.type = cudaMemLocationTypeDevice;
  props.location.id = device;
  cudaMemPool_t pool = nullptr;
  if (cudaMemPoolCreate(&pool, &props) == cudaSuccess)
  {
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  else { pool = nullptr; cudaGetLastError(); }
  pools.push_back({device, pool});
  return pool;
}
cudaError_t pool_malloc_(void **p, size_t bytes, int device, cudaStream_t stream)
{
  cudaMemPool_t pool = device_pool(device);
  return pool ? cudaMallocFromPoolAsync(p, bytes, pool, stream) : cudaMallocAsync(p, bytes, stream);
}
#define pool_malloc(p, bytes, device, stream) pool_malloc_((void **)(p), (bytes), (device), (stream))
}  // namespace

static cudaStream_t thread_stream(int device);  // the calling thread's persistent stream on `device` (defined with HostCtx below)

// per-device scratch: the claim counter
static int launch_batch(const c2a_b200_model *a, const c2a_b200_model *b, const double *poses, const int32_t *sa,
                        const int32_t *sb, int64_t n, double tol_d, double tol_t, const c2a_b200_results *out,
                        unsigned long long *counter, cudaStream_t stream, const double *step_in = nullptr,
                        const int32_t *order = nullptr, bool allow_early = true)
{
  BatchArgs args;
  args.A = DevModel{a->geom, a->rloc, a->meta, a->tris, a->n_nodes, a->n_tris};
  args.B = DevModel{b->geom, b->rloc, b->meta, b->tris, b->n_nodes, b->n_tris};
  args.motions = poses; args.seedA = sa; args.seedB = sb; args.n = n;
  args.tol_d = tol_d; args.tol_t = tol_t; args.out = *out; args.counter = counter;
  args.step_in = step_in;
  args.order = order;

  {
    // once per device (function attributes and memory pools are per device; the callers have made a->device current):
    // opt in to the 106 KB of dynamic shared memory per block
    static std::mutex m;
    static std::vector<int> done;
    std::lock_guard<std::mutex> lk(m);
    if (std::find(done.begin(), done.end(), a->device) == done.end())
    {
      CUDA_TRY(cudaFuncSetAttribute(c2a_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BLOCK_SMEM_BYTES));
      CUDA_TRY(cudaFuncSetAttribute(c2a_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WIDE_BLOCK_SMEM));
      if (env_ll("C2A_B200_WIDE_CARVEOUT", -1) >= 0)  // development aid (no measurable effect: profiles/experiments/README.md)
        cudaFuncSetAttribute(c2a_wide_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)env_ll("C2A_B200_WIDE_CARVEOUT", -1));
      done.push_back(a->device);
    }
  }
  int sms = 0, per_sm = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, a->device));
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, c2a_solve_kernel, BLOCK_THREADS, BLOCK_SMEM_BYTES));
  if (per_sm < 1) per_sm = 1;
  long long blocks = (long long)sms * per_sm;  // persistent: one resident wave (a multiple of the SM count)
  // small batches spread out over all warps: a warp uses ceil(n / warps) of its Q query slots, and with fewer queries
  // than warps one query per warp.  A warp that holds few queries spends its idle lanes on look-ahead, and hands its
  // queries to the wide kernel as soon as they are past their fifth CA step
  const long long need = (n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
  if (blocks > need) blocks = need;
  if (const char *cap = getenv("C2A_B200_MAX_BLOCKS"))  // development aid: profile a slice of the GPU
    if (atoll(cap) > 0 && blocks > atoll(cap)) blocks = atoll(cap);
  if (blocks < 1) blocks = 1;
  const long long warps = blocks * WARPS_PER_BLOCK;
  long long max_slots = (n + warps - 1) / warps;
  if (max_slots > Q) max_slots = Q;
  if (const long long cap = env_ll("C2A_B200_MAX_SLOTS", 0))  // development aid
    if (cap > 0 && max_slots > cap) max_slots = cap;
  if (max_slots < 1) max_slots = 1;
  args.max_slots = (int)max_slots;
  // traversal stacks: one per query slot, depth(A)+depth(B)+2 entries of 128 B
  args.stack_entries = a->depth + b->depth + 2;
  const size_t stack_bytes = (size_t)blocks * WARPS_PER_BLOCK * Q * args.stack_entries * ENTRY_DOUBLES * sizeof(double);

  // hand-over to the wide kernel (whole-query mode only): spill records | spill count, claim counter | per-warp arenas
  const bool wide = !step_in && !getenv("C2A_B200_NO_WIDE");
  int wsms_per = 0;
  if (wide) CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&wsms_per, c2a_wide_kernel, WIDE_THREADS, WIDE_BLOCK_SMEM));
  if (wsms_per < 1) wsms_per = 1;
  long long wblocks = (long long)sms * wsms_per;
  const long long wneed = (n + WIDE_WPB - 1) / WIDE_WPB;
  if (wblocks > wneed) wblocks = wneed;
  if (const char *cap = getenv("C2A_B200_MAX_BLOCKS"))
    if (atoll(cap) > 0 && wblocks > atoll(cap)) wblocks = atoll(cap);
  if (wblocks < 1) wblocks = 1;
  const long long wwarps = wblocks * WIDE_WPB;
  const int w_stack_cap = (int)std::max<long long>(env_ll("C2A_B200_WIDE_STACK", 4096), args.stack_entries + 64);
  const int w_rec_cap = (int)env_ll("C2A_B200_WIDE_RECS", 131072);
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t spill_bytes = wide ? up((size_t)SPILL_BUCKETS * n * MB_DOUBLES * sizeof(double)) : 0, wctl_bytes = wide ? up(SPILL_CTL_WORDS * sizeof(unsigned long long)) : 0;
  const size_t wstack_bytes = wide ? up((size_t)wwarps * w_stack_cap * ENTRY_DOUBLES * sizeof(double)) : 0;
  const size_t wrec_bytes = wide ? up((size_t)wwarps * w_rec_cap * 4 * sizeof(double)) : 0;
  const size_t wleaf_bytes = wide ? up((size_t)wwarps * WIDE_UL * WIDE_LEAFOUT_DOUBLES * sizeof(double)) : 0;
  const size_t waux_bytes = wide ? up((size_t)wwarps * WIDE_AUX_DOUBLES * sizeof(double)) : 0;
  double *stacks = nullptr;
  CUDA_TRY(pool_malloc(&stacks, up(stack_bytes) + spill_bytes + wctl_bytes + wstack_bytes + wrec_bytes + wleaf_bytes + waux_bytes, a->device, stream));
  struct Freer { void *p; cudaStream_t s; ~Freer() { cudaFreeAsync(p, s); } } freer{stacks, stream};
  args.stacks = stacks;
  args.spill_recs = nullptr; args.spill_count = nullptr; args.spill_cap = 0; args.spill_live = 0;
  char *extra = reinterpret_cast<char *>(stacks) + up(stack_bytes);
  if (wide)
  {
    args.spill_recs = reinterpret_cast<double *>(extra);
    args.spill_count = reinterpret_cast<unsigned long long *>(extra + spill_bytes);
    args.spill_cap = n;
    args.spill_live = (int)env_ll("C2A_B200_SPILL_LIVE", 20);
    CUDA_TRY(cudaMemsetAsync(extra, 0xff, spill_bytes, stream));   // query index -1 = "record not written yet"
    CUDA_TRY(cudaMemsetAsync(extra + spill_bytes, 0, wctl_bytes, stream));
  }
  // The wide kernel is launched twice: once beside the main kernel on a side stream (its blocks move in as the main
  // kernel's blocks retire, so the long chains of the heaviest queries overlap the main kernel's run-out), and once after
  // it for whatever is left.  Only when the main kernel fills the machine -- smaller launches end too soon to overlap.
  SideStream *side = nullptr;
  if (wide && allow_early && blocks == (long long)sms * per_sm && !getenv("C2A_B200_NO_EARLY_WIDE")) side = side_stream(a->device, stream);
  args.stats = g_stats_device == a->device ? g_stats_dev : nullptr;
  args.trace = (g_trace_dev && g_trace_device == a->device && n <= g_trace_n && !step_in) ? g_trace_dev : nullptr;
  CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream));
  if (g_kev.device != a->device)
  {
    if (g_kev.device >= 0) for (auto &ev : g_kev.e) cudaEventDestroy(ev);
    g_kev.device = -1;
    bool ok = true;
    for (auto &ev : g_kev.e) ok = ok && cudaEventCreate(&ev) == cudaSuccess;
    if (ok) g_kev.device = a->device;
  }
  const bool kev = g_kev.device == a->device;
  g_kev.recorded = false;
  if (kev) cudaEventRecord(g_kev.e[0], stream);
  if (side && (cudaEventRecord(side->fork, stream) != cudaSuccess || cudaStreamWaitEvent(side->s, side->fork, 0) != cudaSuccess))
  {
    cudaGetLastError(); side = nullptr;
  }
  c2a_solve_kernel<<<(unsigned)blocks, BLOCK_THREADS, BLOCK_SMEM_BYTES, stream>>>(args);
  if (kev) cudaEventRecord(g_kev.e[1], stream);
  g_launches.fetch_add(1);
  cudaError_t le = cudaGetLastError();
  if (le != cudaSuccess) return fail(C2A_B200_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(le));
  if (wide)
  {
    WideArgs w;
    w.A = args.A; w.B = args.B; w.motions = poses; w.seedA = sa; w.seedB = sb; w.tol_d = tol_d; w.tol_t = tol_t; w.out = *out;
    w.items = args.spill_recs; w.ctl = args.spill_count; w.items_cap = n; w.main_blocks = (unsigned)blocks;
    w.stack = reinterpret_cast<double *>(extra + spill_bytes + wctl_bytes);
    w.recs = reinterpret_cast<double *>(extra + spill_bytes + wctl_bytes + wstack_bytes);
    w.leafout = reinterpret_cast<double *>(extra + spill_bytes + wctl_bytes + wstack_bytes + wrec_bytes);
    w.aux = reinterpret_cast<double *>(extra + spill_bytes + wctl_bytes + wstack_bytes + wrec_bytes + wleaf_bytes);
    w.stack_cap = w_stack_cap; w.rec_cap = w_rec_cap;
    w.window = (int)std::min<long long>(16, std::max<long long>(1, env_ll("C2A_B200_WIDE_WINDOW", 16)));
    w.stats = g_wide_stats_device == a->device ? g_wide_stats_dev : nullptr; w.trace = args.trace;
    if (side)
    {
      w.early = 1;
      c2a_wide_kernel<<<(unsigned)wblocks, WIDE_THREADS, WIDE_BLOCK_SMEM, side->s>>>(w);
      g_launches.fetch_add(1);
      // the caller's stream goes on only after the side stream's kernel (which also keeps the scratch alive for it)
      cudaError_t je = cudaGetLastError();
      if (je == cudaSuccess) je = cudaEventRecord(side->join, side->s);
      if (je == cudaSuccess) je = cudaStreamWaitEvent(stream, side->join, 0);
      if (je != cudaSuccess)
      {
        cudaStreamSynchronize(side->s);
        return fail(C2A_B200_ERR_CUDA, std::string("wide kernel, early launch: ") + cudaGetErrorString(je));
      }
    }
    w.early = 0;
    c2a_wide_kernel<<<(unsigned)wblocks, WIDE_THREADS, WIDE_BLOCK_SMEM, stream>>>(w);
    g_launches.fetch_add(1);
    le = cudaGetLastError();
    if (le != cudaSuccess) return fail(C2A_B200_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(le));
  }
  if (kev) cudaEventRecord(g_kev.e[2], stream);
  if (!step_in)
  {
    // translation-only queries (both angular speeds < 1e-8) were skipped by the kernel above: the reference
    // switches to a different traversal for them (C2A.cpp:2391-2395).  A thread per query; threads whose
    // query has rotation return at once, so a batch without such queries pays one near-empty launch.
    TransArgs t;
    t.A = args.A; t.B = args.B; t.motions = poses; t.seedA = sa; t.seedB = sb; t.n = n; t.tol_d = tol_d;
    t.out = *out; t.order = order; t.gstack = nullptr;
    long long tb = (long long)sms * 8;
    const long long tneed = (n + 127) / 128;
    if (tb > tneed) tb = tneed;
    if (args.stack_entries <= TRANS_STACK) c2a_translation_kernel<false><<<(unsigned)tb, 128, 0, stream>>>(t);
    else
    {
      // hierarchies deeper than the local-memory stack: the same kernel with its stacks in global memory, on a small grid
      if (tb > 16) tb = 16;
      CUDA_TRY(pool_malloc(&t.gstack, (size_t)tb * 128 * args.stack_entries * TRANS_ENTRY * sizeof(double), a->device, stream));
      c2a_translation_kernel<true><<<(unsigned)tb, 128, 0, stream>>>(t);
      cudaFreeAsync(t.gstack, stream);
    }
    g_launches.fetch_add(1);
    le = cudaGetLastError();
    if (le != cudaSuccess) return fail(C2A_B200_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(le));
  }
  if (kev) { cudaEventRecord(g_kev.e[3], stream); g_kev.recorded = true; }
  return C2A_B200_OK;
}

// contact pass (c2a_contact.cuh) over n queries; all pointers device-resident
static int launch_contacts(const c2a_b200_model *a, const c2a_b200_model *b, const double *poses24, const double *threshold,
                           const double *distance, const int *collisionfree, const int *status, int64_t n, int max_contacts,
                           int *num_contact, c2a_b200_contact *contacts, unsigned long long *counter, cudaStream_t stream)
{
  ContactArgs args;
  args.A = DevModel{a->geom, a->rloc, a->meta, a->tris, a->n_nodes, a->n_tris};
  args.B = DevModel{b->geom, b->rloc, b->meta, b->tris, b->n_nodes, b->n_tris};
  args.vidxA = a->tri_vidx; args.vidxB = b->tri_vidx;
  args.poses = poses24; args.threshold = threshold; args.distance = distance; args.collisionfree = collisionfree;
  args.status = status; args.n = n; args.max_contacts = max_contacts; args.num_contact = num_contact;
  args.contacts = contacts; args.counter = counter;
  int sms = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, a->device));
  long long blocks = (long long)sms * 4;
  const long long need = (n + 127) / 128;
  if (blocks > need) blocks = need;
  if (blocks < 1) blocks = 1;
  CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream));
  args.gstack = nullptr;
  const int entries = a->depth + b->depth + 2;
  if (entries <= CONTACT_STACK) c2a_contact_kernel<false><<<(unsigned)blocks, 128, 0, stream>>>(args);
  else
  {
    // hierarchies deeper than the local-memory stack: stacks in global memory, small grid
    if (blocks > 16) blocks = 16;
    CUDA_TRY(pool_malloc(&args.gstack, (size_t)blocks * 128 * entries * CONTACT_ENTRY * sizeof(double), a->device, stream));
    c2a_contact_kernel<true><<<(unsigned)blocks, 128, 0, stream>>>(args);
    cudaFreeAsync(args.gstack, stream);
  }
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return C2A_B200_OK;
}

// host-side seed check (the device-pointer entry cannot look at its seeds: the kernels map bad ones to triangle 0)
static int check_seeds(const int32_t *seeds, int64_t n, int n_tris, const char *which)
{
  if (!seeds) return C2A_B200_OK;
  for (int64_t i = 0; i < n; i++)
    if ((uint32_t)seeds[i] >= (uint32_t)n_tris)
      return fail(C2A_B200_ERR_ARG, std::string(which) + "[" + std::to_string(i) + "] = " + std::to_string(seeds[i]) + " is not a triangle of the model (" +
                                        std::to_string(n_tris) + " triangles)");
  return C2A_B200_OK;
}

static int check_pair(const c2a_b200_model *a, const c2a_b200_model *b, int64_t n, const void *poses,
                      const c2a_b200_results *out)
{
  if (!a || !b || !out || (n > 0 && !poses)) return fail(C2A_B200_ERR_ARG, "NULL argument");
  if (n < 0) return fail(C2A_B200_ERR_ARG, "negative batch size");
  if (a->device != b->device) return fail(C2A_B200_ERR_DEVICE, "models live on different devices");
  if (a->depth + b->depth + 2 > 512)
    return fail(C2A_B200_ERR_DEPTH, "BVH depths exceed the traversal stack (" + std::to_string(a->depth) + "+" +
                                        std::to_string(b->depth) + ")");
  return C2A_B200_OK;
}

int c2a_b200_contacts_batch(const c2a_b200_model *a, const c2a_b200_model *b, const double *poses24,
                            const double *threshold, int64_t n, int32_t max_contacts, int32_t *num_contact,
                            c2a_b200_contact *contacts)
{
  if (!a || !b || n < 0 || (n > 0 && (!poses24 || !threshold))) return fail(C2A_B200_ERR_ARG, "NULL argument");
  if (contacts && max_contacts <= 0) return fail(C2A_B200_ERR_ARG, "contacts requested with max_contacts <= 0");
  if (a->device != b->device) return fail(C2A_B200_ERR_DEVICE, "models live on different devices");
  if (n == 0) return C2A_B200_OK;
  ON_DEVICE(a->device);
  cudaStream_t st = thread_stream(a->device);   // stream-ordered: no cudaMalloc / cudaFree, no device-wide synchronisation
  if (!st) return fail(C2A_B200_ERR_CUDA, "no stream");
  const size_t N = (size_t)n, cbytes = contacts ? N * (size_t)max_contacts * sizeof(c2a_b200_contact) : 0;
  char *arena = nullptr;
  const size_t o_thr = N * 192, o_nc = o_thr + ((N * 8 + 255) & ~(size_t)255), o_cnt = o_nc + ((N * 4 + 255) & ~(size_t)255),
               o_ct = o_cnt + 256, total = o_ct + cbytes;
  CUDA_TRY(pool_malloc(&arena, total, a->device, st));
  int rc = C2A_B200_OK;
  cudaError_t e;
  if ((e = cudaMemcpyAsync(arena, poses24, N * 192, cudaMemcpyHostToDevice, st)) != cudaSuccess ||
      (e = cudaMemcpyAsync(arena + o_thr, threshold, N * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess)
    rc = fail(C2A_B200_ERR_CUDA, cudaGetErrorString(e));
  if (rc == C2A_B200_OK)
    rc = launch_contacts(a, b, (const double *)arena, (const double *)(arena + o_thr), nullptr, nullptr, nullptr, n, max_contacts,
                         (int *)(arena + o_nc), contacts ? (c2a_b200_contact *)(arena + o_ct) : nullptr,
                         (unsigned long long *)(arena + o_cnt), st);
  if (rc == C2A_B200_OK && (e = cudaStreamSynchronize(st)) != cudaSuccess) rc = fail(C2A_B200_ERR_CUDA, cudaGetErrorString(e));
  if (rc == C2A_B200_OK && num_contact && (e = cudaMemcpyAsync(num_contact, arena + o_nc, N * 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess)
    rc = fail(C2A_B200_ERR_CUDA, cudaGetErrorString(e));
  if (rc == C2A_B200_OK && contacts && (e = cudaMemcpyAsync(contacts, arena + o_ct, cbytes, cudaMemcpyDeviceToHost, st)) != cudaSuccess)
    rc = fail(C2A_B200_ERR_CUDA, cudaGetErrorString(e));
  cudaStreamSynchronize(st);   // (the copies into the caller's pageable arrays have landed)
  cudaFreeAsync(arena, st);
  return rc;
}

// C2A_Distance (gate = false) and C2A_Collide's C2A_DistanceResult overload (gate = true: the same walk behind the
// box-overlap test) share everything but one template flag of the kernel
static int distance_like(bool gate, const c2a_b200_model *a, const c2a_b200_model *b, const double *poses24, const int32_t *seed_a,
                         const int32_t *seed_b, int64_t n, double rel_err, double abs_err, double *distance, double *p1p2,
                         int32_t *tri_pair, int32_t *num_bv_tests, int32_t *num_tri_tests, int32_t qsize = 2)
{
  if (qsize > 4096) return fail(C2A_B200_ERR_ARG, "qsize above 4096");
  if (!a || !b || n < 0 || (n > 0 && (!poses24 || !distance))) return fail(C2A_B200_ERR_ARG, "NULL argument");
  if (a->device != b->device) return fail(C2A_B200_ERR_DEVICE, "models live on different devices");
  if (gate && (!a->obb || !b->obb)) return fail(C2A_B200_ERR_ARG, "C2A_Collide needs models uploaded with obb_d / obb_To");
  if (n == 0) return C2A_B200_OK;
  if (int rc = check_seeds(seed_a, n, a->n_tris, "seed_a")) return rc;
  if (int rc = check_seeds(seed_b, n, b->n_tris, "seed_b")) return rc;
  ON_DEVICE(a->device);
  cudaStream_t st = thread_stream(a->device);   // stream-ordered: no cudaMalloc / cudaFree, no device-wide synchronisation
  if (!st) return fail(C2A_B200_ERR_CUDA, "no stream");
  const size_t N = (size_t)n;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_pose = take(N * 192), o_sa = seed_a ? take(N * 4) : 0, o_sb = seed_b ? take(N * 4) : 0, o_d = take(N * 8);
  const size_t o_pp = p1p2 ? take(N * 48) : 0, o_tp = tri_pair ? take(N * 8) : 0, o_nbv = num_bv_tests ? take(N * 4) : 0;
  const size_t o_ntri = num_tri_tests ? take(N * 4) : 0;
  char *arena = nullptr;
  CUDA_TRY(pool_malloc(&arena, off, a->device, st));
  int rc = C2A_B200_OK;
  cudaError_t e = cudaSuccess;
#define STEP(x) if (rc == C2A_B200_OK && (e = (x)) != cudaSuccess) rc = fail(C2A_B200_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e));
  STEP(cudaMemcpyAsync(arena + o_pose, poses24, N * 192, cudaMemcpyHostToDevice, st));
  if (seed_a) STEP(cudaMemcpyAsync(arena + o_sa, seed_a, N * 4, cudaMemcpyHostToDevice, st));
  if (seed_b) STEP(cudaMemcpyAsync(arena + o_sb, seed_b, N * 4, cudaMemcpyHostToDevice, st));
  if (rc == C2A_B200_OK)
  {
    DistanceArgs args;
    args.A = DevModel{a->geom, a->rloc, a->meta, a->tris, a->n_nodes, a->n_tris};
    args.B = DevModel{b->geom, b->rloc, b->meta, b->tris, b->n_nodes, b->n_tris};
    args.poses = (const double *)(arena + o_pose);
    args.seedA = seed_a ? (const int *)(arena + o_sa) : nullptr; args.seedB = seed_b ? (const int *)(arena + o_sb) : nullptr;
    args.n = n; args.rel_err = rel_err; args.abs_err = abs_err;
    args.distance = (double *)(arena + o_d); args.p1p2 = p1p2 ? (double *)(arena + o_pp) : nullptr;
    args.tri_pair = tri_pair ? (int *)(arena + o_tp) : nullptr;
    args.num_bv_tests = num_bv_tests ? (int *)(arena + o_nbv) : nullptr;
    args.num_tri_tests = num_tri_tests ? (int *)(arena + o_ntri) : nullptr;
    int sms = 0;
    STEP(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, a->device));
    long long blocks = (long long)sms * 8;
    const long long need = (n + 127) / 128;
    if (blocks > need) blocks = need;
    args.gstack = nullptr;
    args.obbA = a->obb; args.obbB = b->obb;
    const int entries = a->depth + b->depth + 2;
    if (qsize > 2 && !gate)
    {
      // priority-queue routine: a stack of queue frames per thread in global memory, at most ~1 GB of it
      DistanceQueueArgs qa;
      qa.d = args; qa.qsize = qsize; qa.frames = entries; qa.arena = nullptr;
      const size_t per_thread = (size_t)entries * (1 + (size_t)qsize * DQ_ENTRY) * sizeof(double);
      long long fit = (long long)(((size_t)1 << 30) / (per_thread * 128));
      if (fit < 1) fit = 1;
      if (blocks > fit) blocks = fit;
      STEP(pool_malloc(&qa.arena, (size_t)blocks * 128 * per_thread, a->device, st));
      if (rc == C2A_B200_OK) c2a_distance_queue_kernel<<<(unsigned)blocks, 128, 0, st>>>(qa);
      args.gstack = qa.arena;   // (released below)
    }
    else if (entries <= DIST_STACK)
    {
      if (gate) c2a_distance_kernel<false, true><<<(unsigned)blocks, 128, 0, st>>>(args);
      else c2a_distance_kernel<false, false><<<(unsigned)blocks, 128, 0, st>>>(args);
    }
    else
    {
      if (blocks > 16) blocks = 16;
      STEP(pool_malloc(&args.gstack, (size_t)blocks * 128 * entries * DIST_ENTRY * sizeof(double), a->device, st));
      if (rc == C2A_B200_OK)
      {
        if (gate) c2a_distance_kernel<true, true><<<(unsigned)blocks, 128, 0, st>>>(args);
        else c2a_distance_kernel<true, false><<<(unsigned)blocks, 128, 0, st>>>(args);
      }
    }
    g_launches.fetch_add(1);
    STEP(cudaGetLastError());
    STEP(cudaStreamSynchronize(st));
    if (args.gstack) cudaFreeAsync(args.gstack, st);
  }
  STEP(cudaMemcpyAsync(distance, arena + o_d, N * 8, cudaMemcpyDeviceToHost, st));
  if (p1p2) STEP(cudaMemcpyAsync(p1p2, arena + o_pp, N * 48, cudaMemcpyDeviceToHost, st));
  if (tri_pair) STEP(cudaMemcpyAsync(tri_pair, arena + o_tp, N * 8, cudaMemcpyDeviceToHost, st));
  if (num_bv_tests) STEP(cudaMemcpyAsync(num_bv_tests, arena + o_nbv, N * 4, cudaMemcpyDeviceToHost, st));
  if (num_tri_tests) STEP(cudaMemcpyAsync(num_tri_tests, arena + o_ntri, N * 4, cudaMemcpyDeviceToHost, st));
#undef STEP
  cudaStreamSynchronize(st);   // (the copies into the caller's pageable arrays have landed)
  cudaFreeAsync(arena, st);
  return rc;
}

int c2a_b200_distance_batch(const c2a_b200_model *a, const c2a_b200_model *b, const double *poses24, const int32_t *seed_a,
                            const int32_t *seed_b, int64_t n, double rel_err, double abs_err, double *distance, double *p1p2,
                            int32_t *tri_pair, int32_t *num_bv_tests, int32_t *num_tri_tests)
{
  return distance_like(false, a, b, poses24, seed_a, seed_b, n, rel_err, abs_err, distance, p1p2, tri_pair, num_bv_tests, num_tri_tests);
}

int c2a_b200_distance_queue_batch(const c2a_b200_model *a, const c2a_b200_model *b, const double *poses24, const int32_t *seed_a,
                                  const int32_t *seed_b, int64_t n, double rel_err, double abs_err, int32_t qsize, double *distance,
                                  double *p1p2, int32_t *tri_pair, int32_t *num_bv_tests, int32_t *num_tri_tests)
{
  return distance_like(false, a, b, poses24, seed_a, seed_b, n, rel_err, abs_err, distance, p1p2, tri_pair, num_bv_tests, num_tri_tests, qsize);
}

int c2a_b200_collide_distance_batch(const c2a_b200_model *a, const c2a_b200_model *b, const double *poses24, const int32_t *seed_a,
                                    const int32_t *seed_b, int64_t n, double rel_err, double abs_err, double *distance,
                                    double *p1p2, int32_t *tri_pair, int32_t *num_bv_tests, int32_t *num_tri_tests)
{
  return distance_like(true, a, b, poses24, seed_a, seed_b, n, rel_err, abs_err, distance, p1p2, tri_pair, num_bv_tests, num_tri_tests);
}

int c2a_b200_collide_batch(const c2a_b200_model *a, const c2a_b200_model *b, const double *poses24, int64_t n, int32_t flag,
                           int32_t max_pairs, int32_t *num_pairs, int32_t *pairs, int32_t *num_bv_tests, int32_t *num_tri_tests)
{
  if (!a || !b || n < 0 || (n > 0 && (!poses24 || !num_pairs))) return fail(C2A_B200_ERR_ARG, "NULL argument");
  if (flag != 1 && flag != 2) return fail(C2A_B200_ERR_ARG, "flag must be 1 (all contacts) or 2 (first contact)");
  if (max_pairs < 0 || (max_pairs > 0 && !pairs)) return fail(C2A_B200_ERR_ARG, "max_pairs > 0 needs a pairs array");
  if (a->device != b->device) return fail(C2A_B200_ERR_DEVICE, "models live on different devices");
  if (!a->obb || !b->obb) return fail(C2A_B200_ERR_ARG, "C2A_Collide needs models uploaded with obb_d / obb_To");
  if (n == 0) return C2A_B200_OK;
  ON_DEVICE(a->device);
  cudaStream_t st = thread_stream(a->device);   // stream-ordered: no cudaMalloc / cudaFree, no device-wide synchronisation
  if (!st) return fail(C2A_B200_ERR_CUDA, "no stream");
  const size_t N = (size_t)n;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_pose = take(N * 192), o_np = take(N * 4), o_pairs = max_pairs ? take(N * (size_t)max_pairs * 8) : 0;
  const size_t o_nbv = num_bv_tests ? take(N * 4) : 0, o_ntri = num_tri_tests ? take(N * 4) : 0;
  char *arena = nullptr;
  CUDA_TRY(pool_malloc(&arena, off, a->device, st));
  int rc = C2A_B200_OK;
  cudaError_t e = cudaSuccess;
#define STEP(x) if (rc == C2A_B200_OK && (e = (x)) != cudaSuccess) rc = fail(C2A_B200_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e));
  STEP(cudaMemcpyAsync(arena + o_pose, poses24, N * 192, cudaMemcpyHostToDevice, st));
  if (max_pairs) STEP(cudaMemsetAsync(arena + o_pairs, 0xff, N * (size_t)max_pairs * 8, st));   // unused entries read -1
  if (rc == C2A_B200_OK)
  {
    CollideArgs args;
    args.A = DevModel{a->geom, a->rloc, a->meta, a->tris, a->n_nodes, a->n_tris};
    args.B = DevModel{b->geom, b->rloc, b->meta, b->tris, b->n_nodes, b->n_tris};
    args.obbA = a->obb; args.obbB = b->obb;
    args.poses = (const double *)(arena + o_pose);
    args.n = n; args.flag = flag; args.max_pairs = max_pairs;
    args.num_pairs = (int *)(arena + o_np); args.pairs = max_pairs ? (int *)(arena + o_pairs) : nullptr;
    args.num_bv_tests = num_bv_tests ? (int *)(arena + o_nbv) : nullptr;
    args.num_tri_tests = num_tri_tests ? (int *)(arena + o_ntri) : nullptr;
    int sms = 0;
    STEP(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, a->device));
    long long blocks = (long long)sms * 8;
    const long long need = (n + 127) / 128;
    if (blocks > need) blocks = need;
    args.gstack = nullptr;
    const int entries = a->depth + b->depth + 2;
    if (entries <= COLL_STACK) c2a_collide_kernel<false><<<(unsigned)blocks, 128, 0, st>>>(args);
    else
    {
      if (blocks > 16) blocks = 16;
      STEP(pool_malloc(&args.gstack, (size_t)blocks * 128 * entries * COLL_ENTRY * sizeof(double), a->device, st));
      if (rc == C2A_B200_OK) c2a_collide_kernel<true><<<(unsigned)blocks, 128, 0, st>>>(args);
    }
    g_launches.fetch_add(1);
    STEP(cudaGetLastError());
    STEP(cudaStreamSynchronize(st));
    if (args.gstack) cudaFreeAsync(args.gstack, st);
  }
  STEP(cudaMemcpyAsync(num_pairs, arena + o_np, N * 4, cudaMemcpyDeviceToHost, st));
  if (max_pairs) STEP(cudaMemcpyAsync(pairs, arena + o_pairs, N * (size_t)max_pairs * 8, cudaMemcpyDeviceToHost, st));
  if (num_bv_tests) STEP(cudaMemcpyAsync(num_bv_tests, arena + o_nbv, N * 4, cudaMemcpyDeviceToHost, st));
  if (num_tri_tests) STEP(cudaMemcpyAsync(num_tri_tests, arena + o_ntri, N * 4, cudaMemcpyDeviceToHost, st));
#undef STEP
  cudaStreamSynchronize(st);   // (the copies into the caller's pageable arrays have landed)
  cudaFreeAsync(arena, st);
  return rc;
}

int c2a_b200_motions_from_poses(const double *poses, int64_t n, double *motions, int32_t n_threads)
{
  if (n < 0 || (n > 0 && (!poses || !motions))) return fail(C2A_B200_ERR_ARG, "NULL argument");
  motions_from_poses_mt(poses, n, motions, n_threads);
  return C2A_B200_OK;
}

int c2a_b200_schedule_order(const c2a_b200_model *a, const c2a_b200_model *b, const double *motions, int64_t n,
                            int32_t *order)
{
  if (!a || !b || n < 0 || (n > 0 && (!motions || !order))) return fail(C2A_B200_ERR_ARG, "NULL argument");
  if (n > 0x7fffffff) return fail(C2A_B200_ERR_ARG, "batch too large for a 32-bit order");
  schedule_order(motions, n, a->root_ang_radius, b->root_ang_radius, order);
  return C2A_B200_OK;
}

int c2a_b200_solve_batch_device(const c2a_b200_model *a, const c2a_b200_model *b, const double *poses_dev,
                                const int32_t *seed_a_dev, const int32_t *seed_b_dev, const int32_t *order_dev, int64_t n,
                                double tol_d, double tol_t, const c2a_b200_results *out_dev, void *cuda_stream)
{
  int rc = check_pair(a, b, n, poses_dev, out_dev);
  if (rc) return rc;
  if (n == 0) return C2A_B200_OK;
  ON_DEVICE(a->device);
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  unsigned long long *counter = nullptr;
  CUDA_TRY(pool_malloc(&counter, sizeof(unsigned long long), a->device, stream));
  rc = launch_batch(a, b, poses_dev, seed_a_dev, seed_b_dev, n, tol_d, tol_t, out_dev, counter, stream, nullptr, order_dev);
  cudaFreeAsync(counter, stream);
  return rc;
}

// Pinned staging memory for the host-buffer path: cudaMallocHost of a few hundred MB costs more than the
// copy it serves, so one grow-only buffer is kept per process and handed to one call at a time
// (concurrent callers fall back to a private allocation).
namespace {
std::mutex g_pin_mutex;
constexpr int PIN_DEVICES = 64;
struct PinSlot { void *buf = nullptr; size_t cap = 0; bool busy = false; };
PinSlot g_pin[PIN_DEVICES];  // one grow-only staging buffer per device (the multi-device entry runs one call per device at a time)
struct PinnedLease
{
  void *ptr = nullptr;
  int slot = -1;
  cudaError_t acquire(size_t bytes, int device)
  {
    if (device >= 0 && device < PIN_DEVICES)
    {
      std::lock_guard<std::mutex> lk(g_pin_mutex);
      PinSlot &p = g_pin[device];
      if (!p.busy)
      {
        if (p.cap < bytes)
        {
          if (p.buf) cudaFreeHost(p.buf);
          p.buf = nullptr; p.cap = 0;
          cudaError_t e = cudaMallocHost(&p.buf, bytes);
          if (e != cudaSuccess) return e;
          p.cap = bytes;
        }
        p.busy = true; slot = device; ptr = p.buf;
        return cudaSuccess;
      }
    }
    return cudaMallocHost(&ptr, bytes);
  }
  ~PinnedLease()
  {
    if (!ptr) return;
    if (slot >= 0) { std::lock_guard<std::mutex> lk(g_pin_mutex); g_pin[slot].busy = false; }
    else cudaFreeHost(ptr);
  }
};
}  // namespace

// Per host thread and device: a stream, and for small calls a grow-only device arena plus pinned staging memory, kept
// between calls -- a single C2A_Solve must not pay for a stream, two pool allocations and a dozen small copies.
namespace {
struct HostCtx
{
  int device = -1;
  cudaStream_t stream = nullptr;
  char *dev = nullptr; size_t dev_cap = 0;
  char *pin = nullptr; size_t pin_cap = 0;
};
struct HostCtxSet
{
  std::vector<HostCtx> v;
  ~HostCtxSet()
  {
    for (auto &c : v)
    {
      DeviceGuard g(c.device);
      if (g.err != cudaSuccess) continue;
      if (c.stream) cudaStreamDestroy(c.stream);
      if (c.dev) cudaFree(c.dev);
      if (c.pin) cudaFreeHost(c.pin);
    }
  }
};
thread_local HostCtxSet g_ctx;
HostCtx *host_ctx(int device)  // the caller has made `device` current
{
  for (auto &c : g_ctx.v) if (c.device == device) return &c;
  HostCtx c; c.device = device;
  if (cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
  g_ctx.v.push_back(c);
  return &g_ctx.v.back();
}
constexpr size_t SMALL_CALL_BYTES = 8u << 20;  // calls whose device arena is smaller take the persistent-arena path
}  // namespace
static cudaStream_t thread_stream(int device)
{
  HostCtx *c = host_ctx(device);
  return c ? c->stream : nullptr;
}

// wall-clock breakdown of the last host-buffer call on this thread (development aid, c2a_b200_testing.h)
static thread_local double g_host_timing[8];
static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// host-buffer path shared by the three public host entries: inputs are poses (motion constants computed
// here) or ready motion records; step_in != NULL selects single-step mode
// gather != NULL (multi-device entry): element i of this call is element gather[i] of motions / seed_a / seed_b
// (the outputs stay in call order: the caller scatters them)
static int solve_host(const c2a_b200_model *a, const c2a_b200_model *b, const double *poses, const double *motions,
                      const double *step_in, const int32_t *seed_a, const int32_t *seed_b, int64_t n, double tol_d,
                      double tol_t, const c2a_b200_results *out, const int32_t *gather = nullptr)
{
  int rc = check_pair(a, b, n, poses ? poses : motions, out);
  if (rc) return rc;
  if (n == 0) return C2A_B200_OK;
  if (!gather && ((rc = check_seeds(seed_a, n, a->n_tris, "seed_a")) || (rc = check_seeds(seed_b, n, b->n_tris, "seed_b")))) return rc;
  const bool want_contacts = !step_in && (out->num_contact || out->contacts);
  if (want_contacts && out->contacts && out->max_contacts <= 0) return fail(C2A_B200_ERR_ARG, "contacts requested with max_contacts <= 0");
  ON_DEVICE(a->device);
  const double t_begin = now_s();
  HostCtx *ctx = host_ctx(a->device);
  if (!ctx) return fail(C2A_B200_ERR_CUDA, "cudaStreamCreate failed");
  cudaStream_t stream = ctx->stream;

  // one device arena: poses | seeds | outputs | counter
  const size_t N = (size_t)n;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_pose = take(N * 48 * 8);
  const size_t o_sa = seed_a ? take(N * 4) : 0, o_sb = seed_b ? take(N * 4) : 0;
  const size_t o_status = (out->status || want_contacts) ? take(N * 4) : 0, o_cf = (out->collisionfree || want_contacts) ? take(N * 4) : 0;
  const size_t o_nca = out->num_ca ? take(N * 4) : 0, o_nbv = out->num_bv_tests ? take(N * 4) : 0;
  const size_t o_ntri = out->num_tri_tests ? take(N * 4) : 0;
  const size_t o_toc = out->toc ? take(N * 8) : 0, o_dist = (out->distance || want_contacts) ? take(N * 8) : 0;
  const size_t o_mint = out->mint ? take(N * 8) : 0, o_pp = out->p1p2 ? take(N * 48) : 0;
  const size_t o_pt = (out->pose_toc || want_contacts) ? take(N * 192) : 0;
  const size_t o_lt = out->last_tri ? take(N * 8) : 0;
  const size_t o_nc = want_contacts ? take(N * 4) : 0;
  const size_t o_ct = (want_contacts && out->contacts) ? take(N * (size_t)out->max_contacts * sizeof(c2a_b200_contact)) : 0;
  const size_t o_cnt2 = want_contacts ? take(8) : 0;
  const size_t o_step = step_in ? take(N * STEP_IN_DOUBLES * 8) : 0;
  const bool use_order = !step_in && n >= 4096 && n <= 0x7fffffff;
  const size_t o_order = use_order ? take(N * 4) : 0;
  const size_t o_cnt = take(8);
  char *arena = nullptr;
  cudaError_t e = cudaSuccess;
  const size_t in_end = (o_sb ? o_sb : (o_sa ? o_sa : o_pose)) + (((o_sb || o_sa) ? N * 4 : N * 48 * 8) + 255 & ~(size_t)255);  // poses | seeds
  const bool small = off <= SMALL_CALL_BYTES && !want_contacts && !step_in && !use_order;
  if (small)
  {
    // persistent arena and pinned staging (inputs | outputs, same layout as the device arena)
    if (ctx->dev_cap < off)
    {
      if (ctx->dev) cudaFree(ctx->dev);
      ctx->dev = nullptr; ctx->dev_cap = 0;
      if ((e = cudaMalloc(&ctx->dev, SMALL_CALL_BYTES)) == cudaSuccess) ctx->dev_cap = SMALL_CALL_BYTES;
    }
    if (e == cudaSuccess && ctx->pin_cap < off)
    {
      if (ctx->pin) cudaFreeHost(ctx->pin);
      ctx->pin = nullptr; ctx->pin_cap = 0;
      if ((e = cudaMallocHost(&ctx->pin, SMALL_CALL_BYTES)) == cudaSuccess) ctx->pin_cap = SMALL_CALL_BYTES;
    }
    arena = ctx->dev;
  }
  else e = pool_malloc(&arena, off, a->device, stream);
  if (e != cudaSuccess) return fail(C2A_B200_ERR_CUDA, std::string("device arena: ") + cudaGetErrorString(e));
  c2a_b200_results d;
  memset(&d, 0, sizeof(d));
  if (out->status || want_contacts) d.status = (int32_t *)(arena + o_status);
  if (out->collisionfree || want_contacts) d.collisionfree = (int32_t *)(arena + o_cf);
  if (out->num_ca) d.num_ca = (int32_t *)(arena + o_nca);
  if (out->num_bv_tests) d.num_bv_tests = (int32_t *)(arena + o_nbv);
  if (out->num_tri_tests) d.num_tri_tests = (int32_t *)(arena + o_ntri);
  if (out->toc) d.toc = (double *)(arena + o_toc);
  if (out->distance || want_contacts) d.distance = (double *)(arena + o_dist);
  if (out->mint) d.mint = (double *)(arena + o_mint);
  if (out->p1p2) d.p1p2 = (double *)(arena + o_pp);
  if (out->pose_toc || want_contacts) d.pose_toc = (double *)(arena + o_pt);
  if (out->last_tri) d.last_tri = (int32_t *)(arena + o_lt);

  rc = C2A_B200_OK;
#define STEP(x)                                                                              \
  if (rc == C2A_B200_OK && (e = (x)) != cudaSuccess) rc = fail(C2A_B200_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e));
  STEP(cudaMemsetAsync(arena + o_pose + ((N * 48 * 8 + 255) & ~(size_t)255), 0, off - (o_pose + ((N * 48 * 8 + 255) & ~(size_t)255)), stream));  // outputs start zeroed
  // motion constants on the host (libm acos), straight into pinned staging memory
  PinnedLease pin;
  if (!small) STEP(pin.acquire(N * 48 * 8, a->device));
  double *staging = small ? (double *)(ctx->pin + o_pose) : (double *)pin.ptr;
  const double t_alloc = now_s();
  if (rc == C2A_B200_OK)
  {
    if (poses) motions_from_poses_mt(poses, n, staging, 0);
    else if (gather) for (size_t i = 0; i < N; i++) memcpy(staging + 48 * i, motions + 48 * (size_t)gather[i], 48 * 8);
    else memcpy(staging, motions, N * 48 * 8);
  }
  const double t_motions = now_s();
  if (step_in) STEP(cudaMemcpyAsync(arena + o_step, step_in, N * STEP_IN_DOUBLES * 8, cudaMemcpyHostToDevice, stream));
  std::vector<int32_t> order_host;
  if (use_order && rc == C2A_B200_OK)
  {
    order_host.resize(N);
    schedule_order(staging, n, a->root_ang_radius, b->root_ang_radius, order_host.data());
    STEP(cudaMemcpyAsync(arena + o_order, order_host.data(), N * 4, cudaMemcpyHostToDevice, stream));
  }
  const double t_order = now_s();
  STEP(cudaMemcpyAsync(arena + o_pose, staging, N * 48 * 8, cudaMemcpyHostToDevice, stream));
  std::vector<int32_t> seeds_g;
  if (gather && (seed_a || seed_b) && rc == C2A_B200_OK)
  {
    seeds_g.resize(2 * N);
    for (size_t i = 0; i < N; i++) { seeds_g[i] = seed_a ? seed_a[gather[i]] : 0; seeds_g[N + i] = seed_b ? seed_b[gather[i]] : 0; }
    seed_a = seed_a ? seeds_g.data() : nullptr; seed_b = seed_b ? seeds_g.data() + N : nullptr;
  }
  if (small)
  {
    // (the motions went up with the copy above only up to their own end: send the seeds, staged next to them, in one more)
    if (seed_a) memcpy(ctx->pin + o_sa, seed_a, N * 4);
    if (seed_b) memcpy(ctx->pin + o_sb, seed_b, N * 4);
    if (seed_a || seed_b)
    {
      const size_t lo = seed_a ? o_sa : o_sb;
      STEP(cudaMemcpyAsync(arena + lo, ctx->pin + lo, in_end - lo, cudaMemcpyHostToDevice, stream));
    }
  }
  else
  {
    if (seed_a) STEP(cudaMemcpyAsync(arena + o_sa, seed_a, N * 4, cudaMemcpyHostToDevice, stream));
    if (seed_b) STEP(cudaMemcpyAsync(arena + o_sb, seed_b, N * 4, cudaMemcpyHostToDevice, stream));
  }
  if (out->pose_toc && !small) STEP(cudaMemsetAsync(arena + o_pt, 0, N * 192, stream));
  if (rc == C2A_B200_OK)
    rc = launch_batch(a, b, (const double *)(arena + o_pose), seed_a ? (const int32_t *)(arena + o_sa) : nullptr,
                      seed_b ? (const int32_t *)(arena + o_sb) : nullptr, n, tol_d, tol_t, &d,
                      (unsigned long long *)(arena + o_cnt), stream, step_in ? (const double *)(arena + o_step) : nullptr,
                      use_order ? (const int32_t *)(arena + o_order) : nullptr);
  if (want_contacts && rc == C2A_B200_OK)
    rc = launch_contacts(a, b, d.pose_toc, nullptr, d.distance, d.collisionfree, d.status, n, out->max_contacts,
                         (int *)(arena + o_nc), out->contacts ? (c2a_b200_contact *)(arena + o_ct) : nullptr,
                         (unsigned long long *)(arena + o_cnt2), stream);
  const double t_launched = now_s();
  if (rc == C2A_B200_OK) cudaStreamSynchronize(stream);
  const double t_kernel = now_s();
  if (want_contacts && out->num_contact) STEP(cudaMemcpyAsync(out->num_contact, arena + o_nc, N * 4, cudaMemcpyDeviceToHost, stream));
  if (want_contacts && out->contacts)
    STEP(cudaMemcpyAsync(out->contacts, arena + o_ct, N * (size_t)out->max_contacts * sizeof(c2a_b200_contact), cudaMemcpyDeviceToHost, stream));
#define BACK(field, ofs, bytes)                                                                                          \
  if (out->field)                                                                                                       \
  {                                                                                                                     \
    if (small) { if (rc == C2A_B200_OK) memcpy(out->field, ctx->pin + ofs, bytes); }                                    \
    else STEP(cudaMemcpyAsync(out->field, arena + ofs, bytes, cudaMemcpyDeviceToHost, stream));                         \
  }
  if (small)
  {
    // the whole output region in one copy (issued before the wait below would be better still; the kernels dominate)
    STEP(cudaMemcpyAsync(ctx->pin + in_end, arena + in_end, off - in_end, cudaMemcpyDeviceToHost, stream));
    STEP(cudaStreamSynchronize(stream));
  }
  BACK(status, o_status, N * 4) BACK(collisionfree, o_cf, N * 4) BACK(num_ca, o_nca, N * 4)
  BACK(num_bv_tests, o_nbv, N * 4) BACK(num_tri_tests, o_ntri, N * 4) BACK(toc, o_toc, N * 8)
  BACK(distance, o_dist, N * 8) BACK(mint, o_mint, N * 8) BACK(p1p2, o_pp, N * 48) BACK(pose_toc, o_pt, N * 192)
  BACK(last_tri, o_lt, N * 8)
#undef BACK
  STEP(cudaStreamSynchronize(stream));
#undef STEP
  const double t_back = now_s();
  if (!small)
  {
    cudaFreeAsync(arena, stream);
    cudaStreamSynchronize(stream);
  }
  const double t_end = now_s();
  g_host_timing[0] = t_alloc - t_begin; g_host_timing[1] = t_motions - t_alloc; g_host_timing[2] = t_order - t_motions;
  g_host_timing[3] = t_launched - t_order; g_host_timing[4] = t_kernel - t_launched; g_host_timing[5] = t_back - t_kernel;
  g_host_timing[6] = t_end - t_back; g_host_timing[7] = t_end - t_begin;
  return rc;
}

int c2a_b200_host_timing(double *out8)
{
  if (!out8) return fail(C2A_B200_ERR_ARG, "NULL argument");
  memcpy(out8, g_host_timing, sizeof(g_host_timing));
  return C2A_B200_OK;
}

// ---- one batch over several devices ---------------------------------------------------------
// Queries are independent: every device holds a replica of the two models and solves its own shard; there is no
// collective, the "gather" is each device's D2H into its own slice.  The shards interleave the cost-sorted claim
// order (device d takes order[d], order[d + D], ...), so that the queries expected to run long are spread over the
// devices instead of ending up in one contiguous range; one host thread + stream per device.
int c2a_b200_solve_batch_multi(const c2a_b200_model *const *a, const c2a_b200_model *const *b, int32_t n_devices,
                               const double *poses, const int32_t *seed_a, const int32_t *seed_b, int64_t n, double tol_d,
                               double tol_t, const c2a_b200_results *out)
{
  if (!a || !b || n_devices <= 0 || !out || n < 0 || (n > 0 && !poses)) return fail(C2A_B200_ERR_ARG, "NULL argument");
  if (n > 0x7fffffff) return fail(C2A_B200_ERR_ARG, "batch too large for 32-bit query indices");
  if (out->num_contact || out->contacts) return fail(C2A_B200_ERR_ARG, "the contact pass is not available for multi-device batches");
  for (int32_t d = 0; d < n_devices; d++)
  {
    if (!a[d] || !b[d]) return fail(C2A_B200_ERR_ARG, "NULL model handle");
    if (a[d]->device != b[d]->device) return fail(C2A_B200_ERR_DEVICE, "models of one shard live on different devices");
    if (a[d]->n_nodes != a[0]->n_nodes || a[d]->n_tris != a[0]->n_tris || b[d]->n_nodes != b[0]->n_nodes || b[d]->n_tris != b[0]->n_tris)
      return fail(C2A_B200_ERR_ARG, "the per-device models are not replicas of one another");
    for (int32_t e = 0; e < d; e++)
      if (a[e]->device == a[d]->device) return fail(C2A_B200_ERR_DEVICE, "two shards on the same device");
  }
  if (n == 0) return C2A_B200_OK;
  if (int rc = check_seeds(seed_a, n, a[0]->n_tris, "seed_a")) return rc;
  if (int rc = check_seeds(seed_b, n, b[0]->n_tris, "seed_b")) return rc;
  if (n_devices == 1) return c2a_b200_solve_batch(a[0], b[0], poses, seed_a, seed_b, n, tol_d, tol_t, out);
  const size_t N = (size_t)n;
  const int D = n_devices;
  const double t0 = now_s();
  // host half of the motion model once, on all cores; the claim order over the whole batch
  static thread_local std::vector<double> motions;   // grow-only: a fresh 384 MB vector costs 0.1 s of page faults per call
  if (motions.size() < N * 48) motions.resize(N * 48);
  double *const motions_ptr = motions.data();  // (the workers below are other threads: they must not name the thread_local)
  motions_from_poses_mt(poses, n, motions_ptr, 0);
  std::vector<int32_t> order(N);
  if (n >= 4096) schedule_order(motions_ptr, n, a[0]->root_ang_radius, b[0]->root_ang_radius, order.data());
  else for (size_t i = 0; i < N; i++) order[i] = (int32_t)i;
  const double t1 = now_s();
  std::vector<int> rcs(D, 0);
  std::vector<std::string> errs(D);
  std::vector<double> shard_s(D, 0.0);
  auto work = [&](int d) {
    const double ts = now_s();
    const size_t nd = (N - (size_t)d + D - 1) / D;
    std::vector<int32_t> idx(nd);
    for (size_t k = 0; k < nd; k++) idx[k] = order[(size_t)d + k * D];
    // per-shard outputs in call order, scattered to the caller's arrays below
    std::vector<int32_t> status, cf, nca, nbv, ntri, lt;
    std::vector<double> toc, dist, mint, pp, pt;
    c2a_b200_results o;
    memset(&o, 0, sizeof(o));
    if (out->status) { status.resize(nd); o.status = status.data(); }
    if (out->collisionfree) { cf.resize(nd); o.collisionfree = cf.data(); }
    if (out->num_ca) { nca.resize(nd); o.num_ca = nca.data(); }
    if (out->num_bv_tests) { nbv.resize(nd); o.num_bv_tests = nbv.data(); }
    if (out->num_tri_tests) { ntri.resize(nd); o.num_tri_tests = ntri.data(); }
    if (out->toc) { toc.resize(nd); o.toc = toc.data(); }
    if (out->distance) { dist.resize(nd); o.distance = dist.data(); }
    if (out->mint) { mint.resize(nd); o.mint = mint.data(); }
    if (out->p1p2) { pp.resize(nd * 6); o.p1p2 = pp.data(); }
    if (out->pose_toc) { pt.resize(nd * 24); o.pose_toc = pt.data(); }
    if (out->last_tri) { lt.resize(nd * 2); o.last_tri = lt.data(); }
    rcs[d] = solve_host(a[d], b[d], nullptr, motions_ptr, nullptr, seed_a, seed_b, (int64_t)nd, tol_d, tol_t, &o, idx.data());
    if (rcs[d]) { errs[d] = g_err; return; }
    for (size_t k = 0; k < nd; k++)
    {
      const size_t q = (size_t)idx[k];
      if (out->status) out->status[q] = status[k];
      if (out->collisionfree) out->collisionfree[q] = cf[k];
      if (out->num_ca) out->num_ca[q] = nca[k];
      if (out->num_bv_tests) out->num_bv_tests[q] = nbv[k];
      if (out->num_tri_tests) out->num_tri_tests[q] = ntri[k];
      if (out->toc) out->toc[q] = toc[k];
      if (out->distance) out->distance[q] = dist[k];
      if (out->mint) out->mint[q] = mint[k];
      if (out->p1p2) memcpy(out->p1p2 + 6 * q, &pp[6 * k], 48);
      if (out->pose_toc) memcpy(out->pose_toc + 24 * q, &pt[24 * k], 192);
      if (out->last_tri) memcpy(out->last_tri + 2 * q, &lt[2 * k], 8);
    }
    shard_s[d] = now_s() - ts;
  };
  std::vector<std::thread> th;
  for (int d = 0; d < D; d++) th.emplace_back(work, d);
  for (auto &t : th) t.join();
  for (int d = 0; d < D; d++)
    if (rcs[d]) return fail(rcs[d], "device shard " + std::to_string(d) + ": " + errs[d]);
  const double t2 = now_s();
  g_host_timing[0] = 0; g_host_timing[1] = t1 - t0; g_host_timing[2] = 0; g_host_timing[3] = 0;
  g_host_timing[4] = *std::max_element(shard_s.begin(), shard_s.end()); g_host_timing[5] = *std::min_element(shard_s.begin(), shard_s.end());
  g_host_timing[6] = 0; g_host_timing[7] = t2 - t0;
  return C2A_B200_OK;
}

// ---- heterogeneous batches (per-query model handles) and the swept-sphere broadphase: SURVEY section 8, config 4 ----
int c2a_b200_solve_pairs(const c2a_b200_model *const *models, int32_t n_models, const int32_t *model_a,
                         const int32_t *model_b, const double *poses, const int32_t *seed_a, const int32_t *seed_b,
                         int64_t n, double tol_d, double tol_t, const c2a_b200_results *out)
{
  if (!models || n_models <= 0 || !out || n < 0 || (n > 0 && (!poses || !model_a || !model_b)))
    return fail(C2A_B200_ERR_ARG, "NULL argument");
  if (n > 0x7fffffff) return fail(C2A_B200_ERR_ARG, "batch too large for 32-bit query indices");
  if (out->num_contact || out->contacts) return fail(C2A_B200_ERR_ARG, "the contact pass is not available for heterogeneous batches");
  for (int32_t m = 0; m < n_models; m++)
  {
    if (!models[m]) return fail(C2A_B200_ERR_ARG, "NULL model handle");
    if (models[m]->device != models[0]->device) return fail(C2A_B200_ERR_DEVICE, "models live on different devices");
  }
  if (n == 0) return C2A_B200_OK;
  // group the queries by (model_a, model_b): every group is one launch whose claim order lists its queries.  Only the
  // pairs that occur are listed (sorted keys), so a scene with thousands of models costs what its queries cost
  const size_t N = (size_t)n;
  std::vector<int64_t> key(N);
  for (size_t i = 0; i < N; i++)
  {
    if (model_a[i] < 0 || model_a[i] >= n_models || model_b[i] < 0 || model_b[i] >= n_models)
      return fail(C2A_B200_ERR_ARG, "model index out of range");
    if ((seed_a && (uint32_t)seed_a[i] >= (uint32_t)models[model_a[i]]->n_tris) || (seed_b && (uint32_t)seed_b[i] >= (uint32_t)models[model_b[i]]->n_tris))
      return fail(C2A_B200_ERR_ARG, "seed of query " + std::to_string(i) + " is not a triangle of its model");
    key[i] = (int64_t)model_a[i] * n_models + model_b[i];
  }
  std::vector<int64_t> groups(key);
  std::sort(groups.begin(), groups.end());
  groups.erase(std::unique(groups.begin(), groups.end()), groups.end());
  const size_t G = groups.size();
  std::vector<int32_t> group_of(N);
  std::vector<int64_t> start(G + 1, 0);
  for (size_t i = 0; i < N; i++)
  {
    group_of[i] = (int32_t)(std::lower_bound(groups.begin(), groups.end(), key[i]) - groups.begin());
    start[(size_t)group_of[i] + 1]++;
  }
  for (size_t g = 0; g < G; g++) start[g + 1] += start[g];
  for (size_t g = 0; g < G; g++)
  {
    int rc = check_pair(models[groups[g] / n_models], models[groups[g] % n_models], n, poses, out);
    if (rc) return rc;
  }
  ON_DEVICE(models[0]->device);
  cudaStream_t stream;
  CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));

  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_pose = take(N * 48 * 8), o_order = take(N * 4);
  const size_t o_sa = seed_a ? take(N * 4) : 0, o_sb = seed_b ? take(N * 4) : 0;
  const size_t o_zero = off;  // outputs start zeroed from here
  const size_t o_status = out->status ? take(N * 4) : 0, o_cf = out->collisionfree ? take(N * 4) : 0;
  const size_t o_nca = out->num_ca ? take(N * 4) : 0, o_nbv = out->num_bv_tests ? take(N * 4) : 0;
  const size_t o_ntri = out->num_tri_tests ? take(N * 4) : 0;
  const size_t o_toc = out->toc ? take(N * 8) : 0, o_dist = out->distance ? take(N * 8) : 0;
  const size_t o_mint = out->mint ? take(N * 8) : 0, o_pp = out->p1p2 ? take(N * 48) : 0;
  const size_t o_pt = out->pose_toc ? take(N * 192) : 0, o_lt = out->last_tri ? take(N * 8) : 0;
  const size_t o_cnt = take(G * 8);
  char *arena = nullptr;
  cudaError_t e = pool_malloc(&arena, off, models[0]->device, stream);
  if (e != cudaSuccess)
  {
    cudaStreamDestroy(stream);
    return fail(C2A_B200_ERR_CUDA, std::string("cudaMallocAsync: ") + cudaGetErrorString(e));
  }
  c2a_b200_results d;
  memset(&d, 0, sizeof(d));
  if (out->status) d.status = (int32_t *)(arena + o_status);
  if (out->collisionfree) d.collisionfree = (int32_t *)(arena + o_cf);
  if (out->num_ca) d.num_ca = (int32_t *)(arena + o_nca);
  if (out->num_bv_tests) d.num_bv_tests = (int32_t *)(arena + o_nbv);
  if (out->num_tri_tests) d.num_tri_tests = (int32_t *)(arena + o_ntri);
  if (out->toc) d.toc = (double *)(arena + o_toc);
  if (out->distance) d.distance = (double *)(arena + o_dist);
  if (out->mint) d.mint = (double *)(arena + o_mint);
  if (out->p1p2) d.p1p2 = (double *)(arena + o_pp);
  if (out->pose_toc) d.pose_toc = (double *)(arena + o_pt);
  if (out->last_tri) d.last_tri = (int32_t *)(arena + o_lt);

  int rc = C2A_B200_OK;
#define STEP(x) \
  if (rc == C2A_B200_OK && (e = (x)) != cudaSuccess) rc = fail(C2A_B200_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e));
  STEP(cudaMemsetAsync(arena + o_zero, 0, off - o_zero, stream));
  PinnedLease pin;
  STEP(pin.acquire(N * 48 * 8, models[0]->device));
  std::vector<int32_t> order(N);
  if (rc == C2A_B200_OK)
  {
    motions_from_poses_mt(poses, n, (double *)pin.ptr, 0);
    // claim order: the whole batch sorted by expected length (the root radii of the first group stand in for
    // all -- it is only a hint), then split stably by group
    std::vector<int32_t> by_cost(N);
    schedule_order((const double *)pin.ptr, n, models[model_a[0]]->root_ang_radius, models[model_b[0]]->root_ang_radius, by_cost.data());
    std::vector<int64_t> cur(start.begin(), start.end() - 1);
    for (size_t k = 0; k < N; k++)
    {
      const int32_t i = by_cost[k];
      order[(size_t)cur[(size_t)group_of[i]]++] = i;
    }
  }
  STEP(cudaMemcpyAsync(arena + o_pose, pin.ptr, N * 48 * 8, cudaMemcpyHostToDevice, stream));
  STEP(cudaMemcpyAsync(arena + o_order, order.data(), N * 4, cudaMemcpyHostToDevice, stream));
  if (seed_a) STEP(cudaMemcpyAsync(arena + o_sa, seed_a, N * 4, cudaMemcpyHostToDevice, stream));
  if (seed_b) STEP(cudaMemcpyAsync(arena + o_sb, seed_b, N * 4, cudaMemcpyHostToDevice, stream));
  // one launch per non-empty group, on side streams so that a group's tail overlaps the next group's start
  constexpr int NSIDE = 4;
  cudaStream_t side[NSIDE] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ready = nullptr, done[NSIDE] = {nullptr, nullptr, nullptr, nullptr};
  STEP(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
  STEP(cudaEventRecord(ready, stream));
  for (int k = 0; k < NSIDE; k++)
  {
    STEP(cudaStreamCreateWithFlags(&side[k], cudaStreamNonBlocking));
    STEP(cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming));
    if (rc == C2A_B200_OK) STEP(cudaStreamWaitEvent(side[k], ready, 0));
  }
  int launched = 0;
  for (size_t g = 0; g < G && rc == C2A_B200_OK; g++)
  {
    const int64_t ng = start[g + 1] - start[g];
    rc = launch_batch(models[groups[g] / n_models], models[groups[g] % n_models], (const double *)(arena + o_pose),
                      seed_a ? (const int32_t *)(arena + o_sa) : nullptr, seed_b ? (const int32_t *)(arena + o_sb) : nullptr, ng,
                      tol_d, tol_t, &d, (unsigned long long *)(arena + o_cnt) + g, side[launched % NSIDE], nullptr,
                      (const int32_t *)(arena + o_order) + start[g], /*allow_early=*/false);   // (the groups overlap one another)
    launched++;
  }
  for (int k = 0; k < NSIDE; k++)
    if (side[k] && done[k])
    {
      STEP(cudaEventRecord(done[k], side[k]));
      STEP(cudaStreamWaitEvent(stream, done[k], 0));
    }
#define BACK(field, ofs, bytes) \
  if (out->field) STEP(cudaMemcpyAsync(out->field, arena + ofs, bytes, cudaMemcpyDeviceToHost, stream));
  BACK(status, o_status, N * 4) BACK(collisionfree, o_cf, N * 4) BACK(num_ca, o_nca, N * 4)
  BACK(num_bv_tests, o_nbv, N * 4) BACK(num_tri_tests, o_ntri, N * 4) BACK(toc, o_toc, N * 8)
  BACK(distance, o_dist, N * 8) BACK(mint, o_mint, N * 8) BACK(p1p2, o_pp, N * 48) BACK(pose_toc, o_pt, N * 192)
  BACK(last_tri, o_lt, N * 8)
#undef BACK
  STEP(cudaStreamSynchronize(stream));
#undef STEP
  for (int k = 0; k < NSIDE; k++)
  {
    if (side[k]) { cudaStreamSynchronize(side[k]); cudaStreamDestroy(side[k]); }
    if (done[k]) cudaEventDestroy(done[k]);
  }
  if (ready) cudaEventDestroy(ready);
  cudaFreeAsync(arena, stream);
  cudaStreamSynchronize(stream);
  cudaStreamDestroy(stream);
  return rc;
}

// Swept-sphere broadphase: one block per instance i, its threads scan j > i.
__global__ void c2a_broadphase_kernel(const double *c0, const double *c1, const double *radius, int n, double margin,
                                      int *pairs, unsigned long long max_pairs, unsigned long long *count)
{
  const int i = blockIdx.x;
  const double ax = c0[3 * i], ay = c0[3 * i + 1], az = c0[3 * i + 2];
  const double avx = c1[3 * i] - ax, avy = c1[3 * i + 1] - ay, avz = c1[3 * i + 2] - az;
  const double ri = radius[i];
  for (int j = i + 1 + threadIdx.x; j < n; j += blockDim.x)
  {
    const double px = ax - c0[3 * j], py = ay - c0[3 * j + 1], pz = az - c0[3 * j + 2];
    const double vx = avx - (c1[3 * j] - c0[3 * j]), vy = avy - (c1[3 * j + 1] - c0[3 * j + 1]), vz = avz - (c1[3 * j + 2] - c0[3 * j + 2]);
    const double vv = vx * vx + vy * vy + vz * vz;
    double t = (vv > 0.0) ? -(px * vx + py * vy + pz * vz) / vv : 0.0;
    t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
    const double qx = px + t * vx, qy = py + t * vy, qz = pz + t * vz;
    const double reach = ri + radius[j] + margin;
    if (qx * qx + qy * qy + qz * qz <= reach * reach)
    {
      const unsigned long long k = atomicAdd(count, 1ull);
      if (k < max_pairs) { pairs[2 * k] = i; pairs[2 * k + 1] = j; }
    }
  }
}

int c2a_b200_broadphase(const double *c0, const double *c1, const double *radius, int32_t n, double margin, int32_t device,
                        int32_t *pairs, int64_t max_pairs, int64_t *n_pairs)
{
  if (n < 0 || max_pairs < 0 || !n_pairs || (n > 0 && (!c0 || !c1 || !radius)) || (max_pairs > 0 && !pairs))
    return fail(C2A_B200_ERR_ARG, "NULL argument");
  *n_pairs = 0;
  if (n < 2) return C2A_B200_OK;
  ON_DEVICE(device);
  cudaStream_t st = thread_stream(device);   // stream-ordered: no cudaMalloc / cudaFree, no device-wide synchronisation
  if (!st) return fail(C2A_B200_ERR_CUDA, "no stream");
  const size_t N = (size_t)n, o_c1 = N * 24, o_r = 2 * N * 24, o_cnt = o_r + N * 8, o_pairs = o_cnt + 8;
  char *arena = nullptr;
  CUDA_TRY(pool_malloc(&arena, o_pairs + (size_t)max_pairs * 8, device, st));
  int rc = C2A_B200_OK;
  cudaError_t e;
  if ((e = cudaMemcpyAsync(arena, c0, N * 24, cudaMemcpyHostToDevice, st)) != cudaSuccess ||
      (e = cudaMemcpyAsync(arena + o_c1, c1, N * 24, cudaMemcpyHostToDevice, st)) != cudaSuccess ||
      (e = cudaMemcpyAsync(arena + o_r, radius, N * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess ||
      (e = cudaMemsetAsync(arena + o_cnt, 0, 8, st)) != cudaSuccess)
    rc = fail(C2A_B200_ERR_CUDA, cudaGetErrorString(e));
  if (rc == C2A_B200_OK)
  {
    c2a_broadphase_kernel<<<(unsigned)(n - 1), 128, 0, st>>>((const double *)arena, (const double *)(arena + o_c1), (const double *)(arena + o_r), n,
                                                      margin, (int *)(arena + o_pairs), (unsigned long long)max_pairs,
                                                      (unsigned long long *)(arena + o_cnt));
    g_launches.fetch_add(1);
    unsigned long long cnt = 0;
    if ((e = cudaGetLastError()) != cudaSuccess || (e = cudaMemcpyAsync(&cnt, arena + o_cnt, 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
        (e = cudaStreamSynchronize(st)) != cudaSuccess)
      rc = fail(C2A_B200_ERR_CUDA, cudaGetErrorString(e));
    else
    {
      *n_pairs = (int64_t)cnt;
      const size_t w = (size_t)std::min<unsigned long long>(cnt, (unsigned long long)max_pairs);
      if (w > 0 && (e = cudaMemcpyAsync(pairs, arena + o_pairs, w * 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess)
        rc = fail(C2A_B200_ERR_CUDA, cudaGetErrorString(e));
    }
  }
  cudaStreamSynchronize(st);   // (the copies into the caller's pageable arrays have landed)
  cudaFreeAsync(arena, st);
  return rc;
}

int c2a_b200_solve_batch(const c2a_b200_model *a, const c2a_b200_model *b, const double *poses,
                         const int32_t *seed_a, const int32_t *seed_b, int64_t n, double tol_d, double tol_t,
                         const c2a_b200_results *out)
{
  if (n > 0 && !poses) return fail(C2A_B200_ERR_ARG, "NULL argument");
  return solve_host(a, b, poses, nullptr, nullptr, seed_a, seed_b, n, tol_d, tol_t, out);
}

int c2a_b200_solve_batch_motions(const c2a_b200_model *a, const c2a_b200_model *b, const double *motions,
                                 const int32_t *seed_a, const int32_t *seed_b, int64_t n, double tol_d, double tol_t,
                                 const c2a_b200_results *out)
{
  if (n > 0 && !motions) return fail(C2A_B200_ERR_ARG, "NULL argument");
  return solve_host(a, b, nullptr, motions, nullptr, seed_a, seed_b, n, tol_d, tol_t, out);
}

int c2a_b200_toc_step_batch(const c2a_b200_model *a, const c2a_b200_model *b, const double *motions,
                            const double *step_in, const int32_t *seed_a, const int32_t *seed_b, int64_t n,
                            double tol_t, double tol_d, const c2a_b200_results *out)
{
  if (n > 0 && (!motions || !step_in)) return fail(C2A_B200_ERR_ARG, "NULL argument");
  return solve_host(a, b, nullptr, motions, step_in, seed_a, seed_b, n, tol_d, tol_t, out);
}

}  // extern "C"

// ---- unit-test hooks (include/c2a_b200_testing.h) -------------------------------------------
#include "../../include/c2a_b200_testing.h"

namespace c2a {
__global__ void k_test_rect_dist(const double *R, const double *T, const double *ab, long long n, double *dist, double *S)
{
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double r[9], t[3], s[3] = {S[3 * i], S[3 * i + 1], S[3 * i + 2]};
  for (int k = 0; k < 9; k++) r[k] = R[9 * i + k];
  for (int k = 0; k < 3; k++) t[k] = T[3 * i + k];
  dist[i] = rss_rect_dist(r, t, ab[4 * i], ab[4 * i + 1], ab[4 * i + 2], ab[4 * i + 3], s);
  S[3 * i] = s[0]; S[3 * i + 1] = s[1]; S[3 * i + 2] = s[2];
}
__global__ void k_test_tri_distance(const double *R, const double *T, const double *t1, const double *t2, long long n,
                                    double *dist, double *pq)
{
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double r[9], t[3], p[3], q[3];
  for (int k = 0; k < 9; k++) r[k] = R[9 * i + k];
  for (int k = 0; k < 3; k++) t[k] = T[3 * i + k];
  dist[i] = tri_distance_nl(r, t, t1 + 9 * i, t2 + 9 * i, p, q);
  for (int k = 0; k < 3; k++) { pq[6 * i + k] = p[k]; pq[6 * i + 3 + k] = q[k]; }
}
__global__ void k_test_motion(const double *rec, const double *t, const double *ar, const double *dir, long long n, double *out)
{
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Motion m;
  motion_load(m, rec + 24 * i);
  double R[9], T[3];
  motion_pose(m, t[i], R, T);
  for (int k = 0; k < 9; k++) out[14 * i + k] = R[k];
  for (int k = 0; k < 3; k++) out[14 * i + 9 + k] = T[k];
  double n1[3] = {dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]}, n2[3] = {dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]};
  out[14 * i + 12] = motion_bound_bv(m, ar[i], n1);
  out[14 * i + 13] = motion_bound_leaf(m, ar[i], n2);
}
__global__ void k_test_sincos(const double *x, long long n, double *s, double *c)
{
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  s[i] = libm_sin(x[i]);
  c[i] = libm_cos(x[i]);
}

struct DevBuf
{
  std::vector<void *> ptrs;
  ~DevBuf() { for (void *p : ptrs) cudaFree(p); }
  template <class T> cudaError_t up(T **d, const T *h, size_t count)
  {
    cudaError_t e = cudaMalloc((void **)d, count * sizeof(T));
    if (e != cudaSuccess) return e;
    ptrs.push_back(*d);
    return h ? cudaMemcpy(*d, h, count * sizeof(T), cudaMemcpyHostToDevice) : cudaSuccess;
  }
};
}  // namespace c2a

extern "C" {

int c2a_b200_test_rect_dist(const double *R, const double *T, const double *ab, int64_t n, double *dist, double *S)
{
  DevBuf b; double *dR, *dT, *dab, *dd, *dS;
  CUDA_TRY(b.up(&dR, R, 9 * n)); CUDA_TRY(b.up(&dT, T, 3 * n)); CUDA_TRY(b.up(&dab, ab, 4 * n));
  CUDA_TRY(b.up(&dd, (const double *)nullptr, n)); CUDA_TRY(b.up(&dS, S, 3 * n));
  k_test_rect_dist<<<(unsigned)((n + 127) / 128), 128>>>(dR, dT, dab, n, dd, dS);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(dist, dd, n * 8, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(S, dS, 3 * n * 8, cudaMemcpyDeviceToHost));
  return C2A_B200_OK;
}

int c2a_b200_test_tri_distance(const double *R, const double *T, const double *t1, const double *t2, int64_t n,
                               double *dist, double *pq)
{
  DevBuf b; double *dR, *dT, *d1, *d2, *dd, *dpq;
  CUDA_TRY(b.up(&dR, R, 9 * n)); CUDA_TRY(b.up(&dT, T, 3 * n)); CUDA_TRY(b.up(&d1, t1, 9 * n)); CUDA_TRY(b.up(&d2, t2, 9 * n));
  CUDA_TRY(b.up(&dd, (const double *)nullptr, n)); CUDA_TRY(b.up(&dpq, (const double *)nullptr, 6 * n));
  k_test_tri_distance<<<(unsigned)((n + 127) / 128), 128>>>(dR, dT, d1, d2, n, dd, dpq);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(dist, dd, n * 8, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(pq, dpq, 6 * n * 8, cudaMemcpyDeviceToHost));
  return C2A_B200_OK;
}

int c2a_b200_test_motion(const double *rec, const double *t, const double *ang_radius, const double *dir, int64_t n,
                         double *out)
{
  DevBuf b; double *dr, *dt, *da, *dd, *dout;
  CUDA_TRY(b.up(&dr, rec, 24 * n)); CUDA_TRY(b.up(&dt, t, n)); CUDA_TRY(b.up(&da, ang_radius, n)); CUDA_TRY(b.up(&dd, dir, 3 * n));
  CUDA_TRY(b.up(&dout, (const double *)nullptr, 14 * n));
  k_test_motion<<<(unsigned)((n + 127) / 128), 128>>>(dr, dt, da, dd, n, dout);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(out, dout, 14 * n * 8, cudaMemcpyDeviceToHost));
  return C2A_B200_OK;
}

int c2a_b200_test_sincos(const double *x, int64_t n, double *s, double *c)
{
  DevBuf b; double *dx, *ds, *dc;
  CUDA_TRY(b.up(&dx, x, n)); CUDA_TRY(b.up(&ds, (const double *)nullptr, n)); CUDA_TRY(b.up(&dc, (const double *)nullptr, n));
  k_test_sincos<<<(unsigned)((n + 127) / 128), 128>>>(dx, n, ds, dc);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(s, ds, n * 8, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(c, dc, n * 8, cudaMemcpyDeviceToHost));
  return C2A_B200_OK;
}

// Development aid: enable (enable != 0) or disable phase statistics of the solve kernel on the current
// device and read them back: out[0..5] = {expand passes, expand lanes, leaf passes, leaf lanes, advance
// passes, advance lanes} accumulated since the last enable; out[6..8] = globaltimer ns at launch start, at
// the first failed claim (batch drained) and at the last slot retirement (one launch between enables);
// out[9..10] = 32-lane look-ahead passes and the expansion levels they committed;
// out[11..13] = warp cycles spent in EXPAND / LEAF / ADVANCE passes; out[14..19] = {passes, cycles} of the
// three phases for a query alone on its warp (out must hold 20).
// Per-query timeline of the next batches of up to n queries (development aid): out [n][2] = globaltimer ns at
// claim and at result write-out.  n > 0 arms (allocates), n == 0 with out reads back the last launch, n < 0 frees.
int c2a_b200_query_trace(int64_t n, uint64_t *out)
{
  if (n > 0)
  {
    if (g_trace_dev) cudaFree(g_trace_dev);
    g_trace_dev = nullptr; g_trace_n = 0;
    CUDA_TRY(cudaMalloc(&g_trace_dev, (size_t)n * 16));
    cudaGetDevice(&g_trace_device);
    CUDA_TRY(cudaMemset(g_trace_dev, 0, (size_t)n * 16));
    g_trace_n = n;
  }
  else if (n == 0 && out && g_trace_dev)
  {
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(out, g_trace_dev, (size_t)g_trace_n * 16, cudaMemcpyDeviceToHost));
  }
  else if (n < 0 && g_trace_dev) { cudaFree(g_trace_dev); g_trace_dev = nullptr; g_trace_n = 0; }
  return C2A_B200_OK;
}

int c2a_b200_phase_stats(int32_t enable, uint64_t *out20)
{
  uint64_t *out9 = out20;
  if (out9 && g_stats_dev)
  {
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(out9, g_stats_dev, 20 * 8, cudaMemcpyDeviceToHost));
  }
  if (enable && !g_stats_dev) { CUDA_TRY(cudaMalloc(&g_stats_dev, 24 * 8)); cudaGetDevice(&g_stats_device); }
  if (enable)
  {
    const unsigned long long init[24] = {0, 0, 0, 0, 0, 0, ~0ull, ~0ull};
    CUDA_TRY(cudaMemcpy(g_stats_dev, init, sizeof(init), cudaMemcpyHostToDevice));
  }
  if (!enable && g_stats_dev) { cudaFree(g_stats_dev); g_stats_dev = nullptr; }
  return C2A_B200_OK;
}

// Durations (ms) of the kernels of the calling thread's last batch launch -- c2a_solve_kernel, c2a_wide_kernel,
// c2a_translation_kernel -- from CUDA events on the launching stream; waits for that launch to finish.
int c2a_b200_kernel_times(double *out3)
{
  if (!out3) return fail(C2A_B200_ERR_ARG, "NULL argument");
  if (!g_kev.recorded) return fail(C2A_B200_ERR_ARG, "no batch launch on this thread yet");
  ON_DEVICE(g_kev.device);
  CUDA_TRY(cudaEventSynchronize(g_kev.e[3]));
  for (int i = 0; i < 3; i++)
  {
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, g_kev.e[i], g_kev.e[i + 1]));
    out3[i] = ms;
  }
  return C2A_B200_OK;
}

// Counters of c2a_wide_kernel (development aid): enable != 0 arms / re-zeroes, out (WIDE_NSTATS = 24 words) reads the
// counters accumulated since: steps, redone steps, rounds, leaf passes, child tests, triangle tests, events, cycles in
// EXPAND / LEAF / resolve / fold / set-up, queries, one-pair-per-round steps, first / last globaltimer ns.
int c2a_b200_wide_stats(int32_t enable, uint64_t *out16)
{
  if (out16 && g_wide_stats_dev)
  {
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(out16, g_wide_stats_dev, WIDE_NSTATS * 8, cudaMemcpyDeviceToHost));
  }
  if (enable && !g_wide_stats_dev) { CUDA_TRY(cudaMalloc(&g_wide_stats_dev, WIDE_NSTATS * 8)); cudaGetDevice(&g_wide_stats_device); }
  if (enable)
  {
    unsigned long long init[WIDE_NSTATS] = {0};
    init[WS_T_FIRST] = ~0ull;
    CUDA_TRY(cudaMemcpy(g_wide_stats_dev, init, sizeof(init), cudaMemcpyHostToDevice));
  }
  if (!enable && g_wide_stats_dev) { cudaFree(g_wide_stats_dev); g_wide_stats_dev = nullptr; }
  return C2A_B200_OK;
}

int c2a_b200_host_sincos(const double *x, int64_t n, double *s, double *c)
{
  for (int64_t i = 0; i < n; i++) { s[i] = libm_sin(x[i]); c[i] = libm_cos(x[i]); }
  return C2A_B200_OK;
}

}  // extern "C"

// ---- FP64 pipe peak probe (co-roofline denominator; MEASURED_PEAKS.json has no FP64 entry) ----
namespace c2a {
template <bool FMA>
__global__ void k_fp64_peak(double *out, int iters)
{
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; i++)
  {
    if (FMA)
    {
      a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
      a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
    }
    else
    {
      a0 = __dadd_rn(__dmul_rn(a0, m), c); a1 = __dadd_rn(__dmul_rn(a1, m), c); a2 = __dadd_rn(__dmul_rn(a2, m), c);
      a3 = __dadd_rn(__dmul_rn(a3, m), c); a4 = __dadd_rn(__dmul_rn(a4, m), c); a5 = __dadd_rn(__dmul_rn(a5, m), c);
      a6 = __dadd_rn(__dmul_rn(a6, m), c); a7 = __dadd_rn(__dmul_rn(a7, m), c);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
}  // namespace c2a

extern "C" int c2a_b200_fp64_peak(double *tflops_fma, double *tflops_mul_add)
{
  int sms = 0, dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int blocks = sms * 8, threads = 256, iters = 1 << 14;
  double *out = nullptr;
  CUDA_TRY(cudaMalloc(&out, (size_t)blocks * threads * 8));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double res[2] = {0, 0};
  for (int mode = 0; mode < 2; mode++)
  {
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++)
    {
      cudaEventRecord(e0);
      if (mode == 0) k_fp64_peak<true><<<blocks, threads>>>(out, iters);
      else k_fp64_peak<false><<<blocks, threads>>>(out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    const double flops = 2.0 * 8 * (double)iters * blocks * threads;  // both modes: one mul + one add per element-op
    res[mode] = flops / (best * 1e-3) / 1e12;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  CUDA_TRY(cudaGetLastError());
  if (tflops_fma) *tflops_fma = res[0];
  if (tflops_mul_add) *tflops_mul_add = res[1];
  return C2A_B200_OK;
}

#ifdef C2A_RD_STATS
extern "C" int c2a_b200_rd_stats(unsigned long long *out64)
{
  return cudaMemcpyFromSymbol(out64, c2a::g_rd_stats, 64 * sizeof(unsigned long long)) == cudaSuccess ? 0 : -1;
}
#endif
