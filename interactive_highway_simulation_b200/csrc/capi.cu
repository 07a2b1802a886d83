// C ABI of libmcsim_b200.so (include/mcsim.h).  CUDA runtime only — no torch types cross this boundary.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "rollout_kernel.cuh"
#include "bundle_kernel.cuh"
#include "decide_kernel.cuh"
#include "scene.hpp"

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t e_ = (expr);                                                                        \
    if (e_ != cudaSuccess) return fail(MCS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)

// `helper`: the rows of the helper-warp instantiation (carve_smem: hy hax havy | hbeh hok)
size_t rollout_smem_bytes(int n_pad, int steps, bool helper = false) {
  size_t doubles = (size_t)n_pad * (4 + 6 + 6 + 3 + 2 + (helper ? 3 : 0)) + 4 * (size_t)((steps + 2) & ~1) + 16 + 4;
  size_t ints = 32 + 32 + 8 + (size_t)n_pad;
  return doubles * 8 + ints * 4 + (helper ? 7 : 5) * (size_t)n_pad;
}

// runMTimes 0xc2a6e-0xc2b1f: the means of one group of m rollouts (p[0].ok == 0: setEgo failed, fromNSims = 0)
__host__ __device__ inline void group_partial(const mcs_status* p, int m, int steps, mcs_group_partial* out) {
  mcs_group_partial g;
  for (int t = 0; t < 7; ++t) g.f[t] = 0.0;
  g.fromNSims = 0; g.reserved = 0;
  // the scene-error mark of ANY rollout of the group travels with the partial (and through the all-gather of the sharded
  // path): mcs_combine_groups refuses such a group whichever rank simulated it
  for (int i = 0; i < m; ++i) if (p[i].overflow > g.reserved) g.reserved = p[i].overflow;
  if (m > 0 && p[0].ok) {
    int nFall = 0, nFin = 0, sumStep = 0;
    double u = 0.0, c = 0.0, uo = 0.0, co = 0.0;
    for (int i = 0; i < m; ++i) {
      nFall += p[i].fallBack != 0;
      if (p[i].finish) { nFin++; sumStep += p[i].finishStep; }
      u = u + p[i].utility; c = c + p[i].comfort; uo = uo + p[i].utilityObjects; co = co + p[i].comfortObjects;
    }
    const double M = (double)m;
    g.f[0] = (double)nFall / M; g.f[1] = (double)nFin / M; g.f[2] = u / M; g.f[3] = c / M;
    g.f[4] = (double)(sumStep <= 0 ? 1 : sumStep) / (double)(nFin * steps <= 0 ? 1 : nFin * steps);
    g.f[5] = uo / M; g.f[6] = co / M;
    g.fromNSims = m;
  }
  *out = g;
}

// runMultiThreads 0xbca08-0xbca92: plain mean, in group order, of the groups whose fromNSims > 0
__host__ __device__ inline void combine_groups(const mcs_group_partial* g, int groups, mcs_feature* out) {
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};
  int nOk = 0, m = 0;
  for (int q = 0; q < groups; ++q)
    if (g[q].fromNSims > 0) { for (int t = 0; t < 7; ++t) acc[t] = acc[t] + g[q].f[t]; nOk++; m = g[q].fromNSims; }
  const double d = (double)nOk;
  out->fallBackRate = acc[0] / d; out->successRate = acc[1] / d; out->utility = acc[2] / d; out->comfort = acc[3] / d;
  out->expectedStepRatio = acc[4] / d; out->utilityObjects = acc[5] / d; out->comfortObjects = acc[6] / d;
  out->fromNSims = nOk * m; out->reserved = 0;
  for (int q = 0; q < groups; ++q) if (g[q].reserved > out->reserved) out->reserved = g[q].reserved;
}

// one thread per (candidate, group): statuses [n_cand][groups*m] -> partials [n_cand][groups]
__global__ void group_partials_kernel(const mcs_status* __restrict__ st, int n_cand, int groups, int m, int steps,
                                      mcs_group_partial* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cand * groups) return;
  group_partial(st + (size_t)i * m, m, steps, out + i);
}

// one thread per candidate: partials [n_cand][groups] -> features
__global__ void combine_groups_kernel(const mcs_group_partial* __restrict__ g, int n_cand, int groups, mcs_feature* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cand) return;
  combine_groups(g + (size_t)c * groups, groups, out + c);
}

// Roofline denominator of the rollout kernels (the binding resource is the FP64 pipe, and
// MEASURED_PEAKS.json holds no FP64 figure): a stream of independent double-precision fused multiply-adds, eight
// accumulator chains per thread so that neither latency nor issue limits it, no memory traffic.  2 flop per DFMA.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* __restrict__ out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-9, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0, x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0, x7 = x0 + 7.0;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  const double r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (r == 12345.678) out[0] = r;      // never true for the arguments used; keeps the chains alive
}

}  // namespace

// Everything a call needs besides the scene itself lives in one grow-only context per device: creating and destroying
// scenes (one per decision in the reference's call pattern) then costs one stream-ordered allocation and one H2D copy.
struct DeviceCtx {
  int device = -1;
  int n_sm = 148;
  size_t smem_per_block = 227 * 1024;     // opt-in dynamic shared memory limit of a block
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  mcs::CandDev* d_cand = nullptr; int cand_cap = 0;
  mcs_feature* d_feat = nullptr;
  mcs_group_partial* d_part = nullptr;   // [cand_cap][MCS_GROUPS]
  mcs::CandDev* h_cand = nullptr;        // pinned
  mcs_feature* h_feat = nullptr;         // pinned
  mcs_status* d_status = nullptr; size_t status_cap = 0;
  unsigned long long* d_counter = nullptr;
  unsigned int* d_err = nullptr;         // launch-wide scene-error word (LaunchParams::err_flags)
  struct Small {                         // pinned landing zone of every small read-back: a failing call may return while
    unsigned int err;                    // copies are still in flight, and they must not land in a dead stack frame
    int32_t n_states;
    unsigned long long steps;
    mcs_status status;
    mcs::DecideOut decide;
  }* h_small = nullptr;
  void* d_decide = nullptr;              // mcs::DecideOut
  unsigned char* h_stage = nullptr; size_t stage_cap = 0;   // pinned staging of the scene blob
  std::mutex mu;
};

struct mcs_scene {
  int device = 0;
  DeviceCtx* ctx = nullptr;
  mcs::SceneHost host;
  mcs::SceneDev dev{};
  unsigned char* d_blob = nullptr;   // f | i | lane_vec in one allocation, rows dev.n_pad long (0 = not uploaded yet)
  size_t blob_cap = 0;
  int slot_table = -1;               // which thread -> slot table the blob carries: 0 rollout_kernel, 1 bundle_kernel
};

namespace {

std::mutex g_ctx_mu;
DeviceCtx* g_ctx[64] = {nullptr};

int get_ctx(int device, DeviceCtx** out) {
  if (device < 0 || device >= 64) return fail(MCS_ERR_CUDA, "no such CUDA device");
  std::lock_guard<std::mutex> lock(g_ctx_mu);
  if (!g_ctx[device]) {
    CUDA_TRY(cudaSetDevice(device));
    DeviceCtx* c = new DeviceCtx();
    c->device = device;
    cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, device);
    { int v = 0; if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) == cudaSuccess && v > 0) c->smem_per_block = (size_t)v; }
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&c->ev0)) != cudaSuccess || (e = cudaEventCreate(&c->ev1)) != cudaSuccess ||
        (e = cudaMalloc(&c->d_counter, sizeof(unsigned long long))) != cudaSuccess ||
        (e = cudaMalloc(&c->d_err, sizeof(unsigned int))) != cudaSuccess ||
        (e = cudaMallocHost(&c->h_small, sizeof(DeviceCtx::Small))) != cudaSuccess ||
        (e = cudaMalloc(&c->d_decide, sizeof(mcs::DecideOut))) != cudaSuccess) {
      delete c;
      return fail(MCS_ERR_CUDA, cudaGetErrorString(e));
    }
    // scene blobs come from the stream-ordered pool: keep freed blocks in it (the default threshold 0 hands them back to
    // the driver at every synchronisation, and the next decision's cudaMallocAsync then costs milliseconds)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long keep = ~0ULL;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    g_ctx[device] = c;
  }
  *out = g_ctx[device];
  return MCS_OK;
}

int ensure_capacity(mcs_scene* sc, int n_cand, size_t n_status) {
  DeviceCtx* c = sc->ctx;
  if (n_cand > c->cand_cap) {
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (c->d_cand) cudaFree(c->d_cand);
    if (c->d_feat) cudaFree(c->d_feat);
    if (c->d_part) cudaFree(c->d_part);
    if (c->h_cand) cudaFreeHost(c->h_cand);
    if (c->h_feat) cudaFreeHost(c->h_feat);
    c->cand_cap = 0;
    int cap = n_cand < 8 ? 8 : n_cand;
    CUDA_TRY(cudaMalloc(&c->d_cand, sizeof(mcs::CandDev) * cap));
    CUDA_TRY(cudaMalloc(&c->d_feat, sizeof(mcs_feature) * cap));
    CUDA_TRY(cudaMalloc(&c->d_part, sizeof(mcs_group_partial) * cap * MCS_GROUPS));
    CUDA_TRY(cudaMallocHost(&c->h_cand, sizeof(mcs::CandDev) * cap));
    CUDA_TRY(cudaMallocHost(&c->h_feat, sizeof(mcs_feature) * cap));
    c->cand_cap = cap;
  }
  if (n_status > c->status_cap) {
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (c->d_status) cudaFree(c->d_status);
    c->status_cap = 0;
    size_t cap = n_status < 4096 ? 4096 : n_status;
    CUDA_TRY(cudaMalloc(&c->d_status, sizeof(mcs_status) * cap));
    c->status_cap = cap;
  }
  return MCS_OK;
}

// resolves every candidate on the host (Environment::setEgoAgentIdAndDrivingBehavior) and decides which kernel
// instantiation the launch needs; *merge_out = 1 when the merging / yielding code must be compiled in
int stage_candidates(mcs_scene* sc, const mcs_candidate* cands, int n_cand, int* merge_out) {
  int merge = sc->host.needs_merge_path ? 1 : 0;
  for (int c = 0; c < n_cand; ++c) {
    mcs::CandHost h = mcs::resolve_candidate(sc->host, cands[c]);
    if (h.ok && h.gap_id >= 0 && h.gap_slot < 0)
      return fail(MCS_ERR_SCENE, "gap id names no vehicle of the scene (the reference dereferences a null agent here)");
    if (h.ok && h.bad_behavior)
      return fail(MCS_ERR_SCENE, "a non-ego vehicle has a behaviour the reference pushes no action for (only idm, idm_lc, merging)");
    if (h.gap_slot >= 0) merge = 1;
    mcs::CandDev d{};
    d.ego_slot = h.ego_slot; d.action = h.action; d.gap_id = h.gap_id; d.target_lane_id = h.target_lane_id; d.ok = h.ok;
    d.gap_slot = h.gap_slot;
    sc->ctx->h_cand[c] = d;
  }
  CUDA_TRY(cudaMemcpyAsync(sc->ctx->d_cand, sc->ctx->h_cand, sizeof(mcs::CandDev) * n_cand, cudaMemcpyHostToDevice, sc->ctx->stream));
  *merge_out = merge;
  return MCS_OK;
}

// Which vehicle slot each of the `pad` threads of a rollout handles (SceneDev::thread_slot).  Slots stay in the
// reference's visit order and the live threads follow them in increasing order (so a scan over threads is a scan over
// slots: the random-draw positions depend on it).  With spare warps (pad > slots) the table leaves gaps: the slot
// sequence is cut where the initial lane changes — vehicles of one lane run the same behaviour code (merging vehicles
// on the merging lane, yielding ones next to it, plain IDM / MOBIL elsewhere) — neighbouring pieces are merged,
// smallest combined size first, until the pieces fit the warps, and every piece starts on a warp boundary.  Divergent
// behaviour code then runs on different warps side by side instead of in series inside one warp.
// `reserve` threads at the end of the rollout carry no vehicle (the ego's helper warp of rollout_kernel<..., kHelper>).
void thread_slots(const mcs::SceneHost& h, int pad, unsigned char* out, int reserve = 0) {
  struct Piece { int first, count; };
  std::vector<Piece> pieces;
  for (int k = 0; k < h.n; ++k) {
    const int lane = h.i[(size_t)mcs::I_LANE * h.n_pad + k];
    const bool cut = pieces.empty() || lane != h.i[(size_t)mcs::I_LANE * h.n_pad + k - 1] || pieces.back().count == 32;
    if (cut) pieces.push_back(Piece{k, 1}); else pieces.back().count++;
  }
  const int warps = (pad - reserve) / 32;
  auto need = [&]() { int w = 0; for (const Piece& p : pieces) w += (p.count + 31) / 32; return w; };
  while (pieces.size() > 1 && need() > warps) {
    size_t best = 0;
    for (size_t q = 1; q + 1 < pieces.size(); ++q)
      if (pieces[q].count + pieces[q + 1].count < pieces[best].count + pieces[best + 1].count) best = q;
    pieces[best].count += pieces[best + 1].count;
    pieces.erase(pieces.begin() + (long)best + 1);
  }
  std::vector<int> slot_of(pad, -1);
  int t = 0;
  for (const Piece& p : pieces) {
    t = (t + 31) & ~31;
    for (int q = 0; q < p.count; ++q) slot_of[t++] = p.first + q;
  }
  int dead = h.n;                        // idle threads take the padding slots, any order
  for (int q = 0; q < pad; ++q) out[q] = (unsigned char)(slot_of[q] >= 0 ? slot_of[q] : dead++);
}

// Slot of every slot thread of a rollout in bundle_kernel: the live slots sorted by (initial lane, initial behaviour, ego
// last within its lane) — threads that sit side by side in a warp then run the same behaviour code — then the padding
// slots.  Any permutation is correct (the draw positions come from a table indexed by slot, not from the thread order).
void bundle_slots(const mcs::SceneHost& h, int pad, unsigned char* out) {
  std::vector<int> order;
  for (int k = 0; k < h.n; ++k) order.push_back(k);
  auto cls = [&](int k) {
    const int lane = h.i[(size_t)mcs::I_LANE * h.n_pad + k], beh = h.i[(size_t)mcs::I_BEH * h.n_pad + k];
    return (lane < 0 ? 255 : lane) * 16 + beh;
  };
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cls(a) < cls(b); });
  int q = 0;
  for (int k : order) out[q++] = (unsigned char)k;
  for (int k = h.n; q < pad; ++k) out[q++] = (unsigned char)k;
}

// Copy the scene to the device with rows `pad` slots long (pad >= host.n_pad; the kernels index rows with their block
// size as a compile-time stride).  The upload happens at the first use and again only if a later launch wants another
// row length: one blob (doubles | ints | lane vector), staged through pinned memory, one stream-ordered allocation,
// one copy.  Caller holds the context lock and has selected the device.
int upload_scene(mcs_scene* sc, int pad, int slot_table = 0) {
  if (sc->dev.n_pad == pad && sc->slot_table == slot_table) return MCS_OK;
  DeviceCtx* c = sc->ctx;
  const mcs::SceneHost& h = sc->host;
  if (pad < h.n_pad) return fail(MCS_ERR_ARG, "row length below the scene's");
  const size_t bf = (size_t)mcs::F_COUNT * pad * sizeof(double), bi = (size_t)mcs::I_COUNT * pad * sizeof(int32_t), bv = (size_t)pad;
  const size_t total = bf + bi + 2 * ((bv + 15) & ~(size_t)15);        // doubles | ints | lane vector | thread -> slot table
  CUDA_TRY(cudaStreamSynchronize(c->stream));     // the staging buffer may still feed an earlier copy
  if (total > c->stage_cap) {
    if (c->h_stage) cudaFreeHost(c->h_stage);
    c->h_stage = nullptr; c->stage_cap = 0;
    const size_t cap = total < (1u << 18) ? (1u << 18) : total;
    CUDA_TRY(cudaMallocHost(&c->h_stage, cap));
    c->stage_cap = cap;
  }
  if (pad == h.n_pad) {
    memcpy(c->h_stage, h.f.data(), bf);
    memcpy(c->h_stage + bf, h.i.data(), bi);
    memcpy(c->h_stage + bf + bi, h.lane_vec.data(), bv);
  } else {
    memset(c->h_stage, 0, bf);                    // padding slots: 0.0 / -1 / 0, as build_scene_host fills its own
    memset(c->h_stage + bf, 0xff, bi);
    memset(c->h_stage + bf + bi, 0, bv);
    for (int q = 0; q < mcs::F_COUNT; ++q)
      memcpy(c->h_stage + (size_t)q * pad * sizeof(double), h.f.data() + (size_t)q * h.n_pad, (size_t)h.n_pad * sizeof(double));
    for (int q = 0; q < mcs::I_COUNT; ++q)
      memcpy(c->h_stage + bf + (size_t)q * pad * sizeof(int32_t), h.i.data() + (size_t)q * h.n_pad, (size_t)h.n_pad * sizeof(int32_t));
    memcpy(c->h_stage + bf + bi, h.lane_vec.data(), (size_t)h.n_pad);
  }
  if (slot_table == 1) bundle_slots(h, pad, c->h_stage + bf + bi + ((bv + 15) & ~(size_t)15));
  else thread_slots(h, pad, c->h_stage + bf + bi + ((bv + 15) & ~(size_t)15), slot_table >= 2 ? 32 : 0);
  if (total > sc->blob_cap) {
    if (sc->d_blob) CUDA_TRY(cudaFreeAsync(sc->d_blob, c->stream));
    sc->d_blob = nullptr; sc->blob_cap = 0; sc->dev.n_pad = 0;
    CUDA_TRY(cudaMallocAsync(&sc->d_blob, total, c->stream));
    sc->blob_cap = total;
  }
  sc->dev.n_pad = 0;
  CUDA_TRY(cudaMemcpyAsync(sc->d_blob, c->h_stage, total, cudaMemcpyHostToDevice, c->stream));
  mcs::SceneDev& d = sc->dev;
  d.f = (const double*)sc->d_blob; d.i = (const int32_t*)(sc->d_blob + bf); d.lane_vec = sc->d_blob + bf + bi;
  d.thread_slot = d.lane_vec + ((bv + 15) & ~(size_t)15);
  d.n_pad = pad;     // the copy is ordered before every later launch on the context's stream
  sc->slot_table = slot_table;
  return MCS_OK;
}

// Threads per rollout.  One per vehicle slot is the minimum; a small launch (few rollouts per SM) is latency-bound — each
// rollout walks its barriers alone — and the phases that are task-parallel inside a rollout (the seven MOBIL sub-tasks
// per undecided vehicle, the outcome statistics) finish sooner on more threads, so the row length doubles while the
// launch still fits the SMs in one wave — 24 resident warps per SM for the 80-register lane-change instantiation, 16 for
// the 128-register merging one (scripts/kernel_time.py / threads_sweep.py, 24 vehicles x 5 candidates x 160 rollouts:
// 32 threads 0.66 ms, 64: 0.50, 128: 0.44, 256: 0.55; merging, 20 vehicles x 4 gaps x 160 rollouts: 32 threads 0.87 ms,
// 64: 0.78, 128: 1.05).  The result does not depend on it.
#ifndef MCS_WIDE_WARPS_PER_SM
#define MCS_WIDE_WARPS_PER_SM 24
#endif
#ifndef MCS_WIDE_WARPS_PER_SM_MERGE
#define MCS_WIDE_WARPS_PER_SM_MERGE 16
#endif
int rollout_threads(const mcs_scene* sc, long long rollouts, bool traj, bool merge) {
  int t = sc->host.n_pad;
  if (const char* e = getenv("MCS_ROLLOUT_THREADS")) {        // tuning / test override
    const int want = atoi(e);
    if ((want == 32 || want == 64 || want == 128 || want == 256) && want >= t) return want;
    return t;
  }
  if (traj) return t;
  const long long budget = (long long)(merge ? MCS_WIDE_WARPS_PER_SM_MERGE : MCS_WIDE_WARPS_PER_SM) * sc->ctx->n_sm;
  // (a merging launch that gets a helper warp at 128 threads may fill one wave of six-rollout blocks)
  const bool helper128 = merge && t <= 64 && sc->host.n <= 96 && rollouts <= 6LL * sc->ctx->n_sm;
  while (t < 256 && (rollouts * (2 * t / 32) <= budget || (helper128 && t < 128))) {
    // a merging rollout that already has its helper warp (128 threads, <= 96 vehicles) gains nothing from twice the threads
    // (4 gaps x 64 rollouts x 40 vehicles: 0.51 ms on 128 threads, 0.64 on 256; scripts/cfg1_variants.py)
    if (merge && t >= 128 && sc->host.n <= t - 32) break;
    t *= 2;
  }
  return t;
}

// Threads per block.  A rollout of 256 threads owns a 256-thread block.  Narrower rollouts share one and start every
// step together (rollout_kernel.cuh) — 256 threads per block; the merging instantiation, the instruction-fetch-bound one,
// 512 (measured, 4 x 512 rollouts x 40 vehicles: 128 threads 3.42 ms, 256: 2.73 ms, 512: 2.48 ms).  Named barriers limit
// a block to 15 rollouts.  Small launches do not get here with narrow rollouts any more (rollout_threads widens them),
// so sharing has no lower bound by default: scripts/share_sweep.py — it never loses, and wins up to 27 % where the old
// bound of 4 rollouts per SM kept 480 128-thread rollouts in blocks of their own (0.86 -> 0.63 ms).
#ifndef MCS_SHARE_MIN_PER_SM
#define MCS_SHARE_MIN_PER_SM 0   // rollouts per SM below which a narrow rollout still owns its block (tuning knob)
#endif
#ifndef MCS_FULL_MERGE_BLOCK
#define MCS_FULL_MERGE_BLOCK 768   // threads per block for 256-thread merging rollouts in large launches (256 = own block)
#endif
#ifndef MCS_FULL_LC_BLOCK
#define MCS_FULL_LC_BLOCK 256      // a 256-thread lane-change rollout owns its block (768: three in step, measured below)
#endif
#ifndef MCS_HALF_MERGE_BLOCK
// threads per block for 128-thread merging rollouts in large launches: six rollouts at 80 registers instead of four at 128
// (scripts/merge_width_times.py, 4 gaps x 4096 rollouts x 50 steps: exit scene of 100 vehicles 16.7 -> 15.0 ms, 120 vehicles
// x 2048 rollouts 6.74 -> 6.14 ms; 896 threads: 15.9 / 6.64; a merging scene of 100 vehicles 11.97 -> 12.19 ms)
#define MCS_HALF_MERGE_BLOCK 768
#endif
#ifndef MCS_NARROW_MERGE_WAVE_BLOCK
#define MCS_NARROW_MERGE_WAVE_BLOCK 896   // 14 rollouts of 64 threads per SM in one block (448: the same 14 in two blocks of 7)
#endif
#ifndef MCS_QUARTER_MERGE_BLOCK
#define MCS_QUARTER_MERGE_BLOCK 512   // threads per block for 64-thread merging rollouts (896 when that saves a wave, below)
#endif
template <bool kMerge, int NPS>
struct BlockThreads { static constexpr int value = NPS == 256 ? (kMerge ? MCS_FULL_MERGE_BLOCK : MCS_FULL_LC_BLOCK) : (kMerge && NPS == 128 ? MCS_HALF_MERGE_BLOCK : (kMerge && NPS >= 64 ? MCS_QUARTER_MERGE_BLOCK : 256)); };

// The ego's helper warp (rollout_kernel<..., kHelper>): merging / exit launches whose rollouts are 128 or 256 threads wide and
// leave the last warp free (64-thread rollouts share blocks of 8 or 14: two named barriers per rollout do not fit the 16 of a
// block).  MCS_NO_HELPER=1 switches it off (tests run both ways; the records are identical).
// Returns the thread -> slot table the launch wants: 0 none, 2 helper warp for the ego, 3 helper warp for the ego and the
// ramp vehicles (scenes that have any).
int wants_helper(const mcs_scene* sc, int pad, bool traj, bool merge) {
  if (!merge || traj || pad < 128 || sc->host.n > pad - 32) return 0;
  if (const char* e = getenv("MCS_NO_HELPER")) if (atoi(e)) return 0;
  const mcs::SceneHost& h = sc->host;
  for (int k = 0; k < h.n; ++k)
    if (h.i[(size_t)mcs::I_BEH * h.n_pad + k] == MCS_BEH_MERGING) return 3;
  return 2;
}

template <bool kTraj, bool kMerge, int NPS, int BT, int kHelper = 0>
int launch_bt(mcs_scene* sc, dim3 grid, size_t region, const mcs::LaunchParams& lp) {
  if constexpr (kHelper == 0 && !kTraj && kMerge && NPS >= 128) {
    if (sc->slot_table == 2) return launch_bt<kTraj, kMerge, NPS, BT, 1>(sc, grid, region, lp);
    if (sc->slot_table == 3) return launch_bt<kTraj, kMerge, NPS, BT, 2>(sc, grid, region, lp);
  }
  constexpr int G = BT / NPS;       // rollouts per thread block
  static_assert(G >= 1 && G <= 15, "one named barrier per rollout of a block");
  const size_t smem = region * G;
  if (smem > sc->ctx->smem_per_block) return fail(MCS_ERR_CAPACITY, "horizon too long: one rollout's histories do not fit shared memory");
  grid.x = (grid.x + G - 1) / G;
  CUDA_TRY(cudaFuncSetAttribute(mcs::rollout_kernel<kTraj, kMerge, NPS, BT, kHelper>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mcs::rollout_kernel<kTraj, kMerge, NPS, BT, kHelper><<<grid, dim3(BT), smem, sc->ctx->stream>>>(sc->dev, sc->ctx->d_cand, lp, (int)region);
  CUDA_TRY(cudaGetLastError());
  return MCS_OK;
}

template <bool kTraj, bool kMerge, int NPS>
int launch_np(mcs_scene* sc, dim3 grid, size_t region, const mcs::LaunchParams& lp) {
  constexpr int BT = BlockThreads<kMerge, NPS>::value;
  if (NPS == BT || kTraj) return launch_bt<kTraj, kMerge, NPS, NPS>(sc, grid, region, lp);
  // shared blocks pay off once the launch puts several rollouts on every SM
  const long long rollouts = (long long)grid.x * grid.y;
  long long threshold = (long long)sc->ctx->n_sm * MCS_SHARE_MIN_PER_SM;
  if (const char* e = getenv("MCS_SHARE_MIN_PER_SM")) threshold = (long long)sc->ctx->n_sm * atoi(e);   // tuning override
  if (rollouts < threshold) return launch_bt<kTraj, kMerge, NPS, NPS>(sc, grid, region, lp);
  // a long horizon makes the per-rollout region large (four history arrays of steps + 2 doubles): share a block only
  // while its rollouts fit the SM's shared memory
  const size_t smem_limit = sc->ctx->smem_per_block;
  if constexpr (kMerge && NPS == 128 && BT != 512) {
    // six 128-thread rollouts per block pay only when every SM gets a full block; smaller launches keep the four-rollout
    // 512-thread blocks (4 gaps x 128 rollouts x 40 vehicles, widened to 128 threads: 0.77 ms there, 1.13 ms in 768-thread blocks)
    // ... unless the four-rollout blocks would need a second wave and the six-rollout ones do not
    const long long sm = sc->ctx->n_sm;
    const bool one_wave_of_six = (rollouts + 3) / 4 > sm && (rollouts + 5) / 6 <= sm && region * (BT / NPS) <= smem_limit;
    if (!one_wave_of_six && (rollouts < sm * (BT / NPS) || region * (BT / NPS) > smem_limit)) {
      if (region * 4 <= smem_limit) return launch_bt<kTraj, kMerge, NPS, 512>(sc, grid, region, lp);
      return launch_bt<kTraj, kMerge, NPS, NPS>(sc, grid, region, lp);
    }
  }
  if (region * (BT / NPS) > smem_limit) return launch_bt<kTraj, kMerge, NPS, NPS>(sc, grid, region, lp);
  if constexpr (NPS == 256) {
    // full-width rollouts share a block only when every SM still gets a full block
    if (rollouts < (long long)sc->ctx->n_sm * (BT / NPS)) return launch_bt<kTraj, kMerge, NPS, NPS>(sc, grid, region, lp);
  }
  if constexpr (kMerge && NPS == 64) {
    // The merging instantiation is latency-bound per step: a wave of 512-thread blocks (8 rollouts, 128 registers) takes
    // about as long whether the SM holds one rollout or eight, so a launch that needs a second wave is better off in
    // 896-thread blocks (14 rollouts, 72 registers, ~1.7x the time per wave — measured 4 x 512 rollouts x 40 vehicles:
    // 2.03 -> 1.86 ms; 4 x 1024: 3.90 -> 3.42; but 4 x 256, one wave either way: 1.07 vs 1.74).
    const long long sm = sc->ctx->n_sm;
    const long long waves8 = ((rollouts + 7) / 8 + sm - 1) / sm, waves14 = ((rollouts + 13) / 14 + sm - 1) / sm;
    if (waves14 * 17 < waves8 * 10 && region * 14 <= smem_limit) return launch_bt<kTraj, kMerge, NPS, MCS_NARROW_MERGE_WAVE_BLOCK>(sc, grid, region, lp);
  }
  return launch_bt<kTraj, kMerge, NPS, BT>(sc, grid, region, lp);
}

template <bool kTraj, bool kMerge>
int launch_one(mcs_scene* sc, dim3 grid, size_t smem, const mcs::LaunchParams& lp) {
  switch (sc->dev.n_pad) {
    case 32: return launch_np<kTraj, kMerge, 32>(sc, grid, smem, lp);
    case 64: return launch_np<kTraj, kMerge, 64>(sc, grid, smem, lp);
    case 128: return launch_np<kTraj, kMerge, 128>(sc, grid, smem, lp);
    case 256: return launch_np<kTraj, kMerge, 256>(sc, grid, smem, lp);
    default: return fail(MCS_ERR_CAPACITY, "unsupported block size");
  }
}

// ---- bundle_kernel (scenes of up to 64 vehicles): rollouts per block RW (a power of two: the lanes of a warp are RW
// rollouts x 32/RW vehicle slots), slot threads per rollout NS, threads per block NS * RW <= 1024.  The wider the bundle the
// more lanes of a warp run the same code; the launch should still give every SM a block.
#ifndef MCS_BUNDLE_MAX_RW_LOG2
#define MCS_BUNDLE_MAX_RW_LOG2 4
#endif
#ifndef MCS_BUNDLE_MIN_SLOTS
#define MCS_BUNDLE_MIN_SLOTS 49152    // rollouts x vehicles from which a merging launch goes to bundle_kernel
#endif
template <bool kMerge, int MAXT>
int launch_bundle_t(mcs_scene* sc, dim3 grid, int threads, size_t smem, const mcs::LaunchParams& lp, const mcs::BundleParams& bp) {
  CUDA_TRY(cudaFuncSetAttribute(mcs::bundle_kernel<kMerge, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mcs::bundle_kernel<kMerge, MAXT><<<grid, dim3(threads), smem, sc->ctx->stream>>>(sc->dev, sc->ctx->d_cand, lp, bp);
  CUDA_TRY(cudaGetLastError());
  return MCS_OK;
}

// returns 1 when the launch went to bundle_kernel, 0 when this launch is not for it, < 0 on error
int try_launch_bundle(mcs_scene* sc, int n_cand, const mcs::LaunchParams& lp, bool merge) {
  const int n = sc->host.n;
  // capability: up to 64 vehicle slots, no trajectory output, a horizon the per-rollout pool of yielding intentions is sized
  // for (longer ones go to rollout_kernel's per-vehicle tables)
  if (n > mcs::BNP || lp.traj || lp.steps > 128) return 0;
  const long long total = (long long)lp.count * n_cand;
  const char* forced = getenv("MCS_KERNEL");               // tests run every scene through both kernels
  if (forced && !strcmp(forced, "rollout")) return 0;
  if (!(forced && !strcmp(forced, "bundle"))) {
    // Where it pays (scripts/bundle_sweep.py, B200): merging / exit scenes once the launch is past what rollout_kernel runs
    // in one wave — 40 vehicles x 4 gaps x 512 rollouts 1.60 -> 1.30 ms, x 2048 rollouts 6.7 -> 4.7 ms.  Below that a
    // rollout's own warps finish a step sooner (4 x 64 rollouts 0.75 vs 1.12 ms), and in plain lane-change scenes
    // rollout_kernel's compacted MOBIL sub-tasks win (64 vehicles x 5 x 512: 0.98 vs 1.62 ms).
    if (!merge || total * n < MCS_BUNDLE_MIN_SLOTS) return 0;
  }
  const size_t region = mcs::bundle_region_bytes(lp.steps, merge);
  int max_log2 = MCS_BUNDLE_MAX_RW_LOG2;
  if (const char* e = getenv("MCS_BUNDLE_RW_LOG2")) { const int v = atoi(e); if (v >= 0 && v <= 5) max_log2 = v; }   // tuning / tests
  int rw_log2 = -1;
  for (int l = max_log2; l >= 0; --l) {
    const int rw = 1 << l, hslots = 32 >> l, ns = (n + hslots - 1) / hslots * hslots;
    if (ns * rw > 1024) continue;
    if (mcs::BUNDLE_STATIC_BYTES + (size_t)rw * region > sc->ctx->smem_per_block) continue;
    rw_log2 = l;
    // wide enough once every SM has a block (a forced width is taken as it is)
    if (getenv("MCS_BUNDLE_RW_LOG2") || (total + rw - 1) / rw >= sc->ctx->n_sm * 3 / 4) break;
  }
  if (rw_log2 < 0) return 0;                      // horizon too long for even one rollout per block: rollout_kernel decides
  const int rw = 1 << rw_log2, hslots = 32 >> rw_log2, ns = (n + hslots - 1) / hslots * hslots;
  int rc = upload_scene(sc, mcs::BNP, 1);
  if (rc != MCS_OK) return rc;
  mcs::BundleParams bp{rw_log2, ns, (int)region, 0};
  const int threads = ns * rw;
  const size_t smem = mcs::BUNDLE_STATIC_BYTES + (size_t)rw * region;
  dim3 grid((unsigned)((lp.count + rw - 1) / rw), (unsigned)n_cand);
  if (threads <= 384) rc = merge ? launch_bundle_t<true, 384>(sc, grid, threads, smem, lp, bp) : launch_bundle_t<false, 384>(sc, grid, threads, smem, lp, bp);
  else if (threads <= 512) rc = merge ? launch_bundle_t<true, 512>(sc, grid, threads, smem, lp, bp) : launch_bundle_t<false, 512>(sc, grid, threads, smem, lp, bp);
  else if (threads <= 768) rc = merge ? launch_bundle_t<true, 768>(sc, grid, threads, smem, lp, bp) : launch_bundle_t<false, 768>(sc, grid, threads, smem, lp, bp);
  else rc = merge ? launch_bundle_t<true, 1024>(sc, grid, threads, smem, lp, bp) : launch_bundle_t<false, 1024>(sc, grid, threads, smem, lp, bp);
  return rc == MCS_OK ? 1 : rc;
}

// launch the rollout kernel for rollouts [first, first+count) of every candidate; statuses → d_status (device)
int launch_rollouts(mcs_scene* sc, int n_cand, const mcs_sim_param* sp, uint64_t seed, int64_t first, int64_t count,
                    mcs_status* d_status, double* d_traj, int32_t* d_traj_n, unsigned long long* d_counter, int merge) {
  if (sp->steps < 0 || sp->steps > 4096 || !(sp->dt > 0.0)) return fail(MCS_ERR_ARG, "bad SimParameter");
  if (count <= 0 || count > 0x7fffffffLL) return fail(MCS_ERR_ARG, "bad rollout count");
  mcs::LaunchParams lp{};
  lp.dt = sp->dt; lp.steps = sp->steps; lp.min_steps = sp->min_steps;
  lp.seed = seed; lp.first = first; lp.count = count; lp.n_cand = n_cand;
  lp.status = d_status; lp.traj = d_traj; lp.traj_n = d_traj_n; lp.vehicle_steps = d_counter;
  lp.err_flags = sc->ctx->d_err;
  CUDA_TRY(cudaMemsetAsync(sc->ctx->d_err, 0, sizeof(unsigned int), sc->ctx->stream));
  int rc = try_launch_bundle(sc, n_cand, lp, merge != 0);
  if (rc < 0) return rc;
  if (rc == 1) {
    CUDA_TRY(cudaMemcpyAsync(&sc->ctx->h_small->err, sc->ctx->d_err, sizeof(unsigned int), cudaMemcpyDeviceToHost, sc->ctx->stream));
    return MCS_OK;
  }
  {
    const int pad = rollout_threads(sc, (long long)count * n_cand, d_traj != nullptr, merge != 0);
    rc = upload_scene(sc, pad, wants_helper(sc, pad, d_traj != nullptr, merge != 0));
  }
  if (rc != MCS_OK) return rc;
  size_t smem = (rollout_smem_bytes(sc->dev.n_pad, sp->steps, sc->slot_table == 3) + 15) & ~(size_t)15;   // one rollout's region
  dim3 grid((unsigned)count, (unsigned)n_cand);
  // the trajectory-writing variant is a test / Simulation.run() path: one instantiation (the merging superset) serves it
  if (d_traj) rc = launch_one<true, true>(sc, grid, smem, lp);
  else rc = merge ? launch_one<false, true>(sc, grid, smem, lp) : launch_one<false, false>(sc, grid, smem, lp);
  if (rc != MCS_OK) return rc;
  // the launch-wide scene-error word follows the kernel on the stream; callers read it after their synchronisation
  CUDA_TRY(cudaMemcpyAsync(&sc->ctx->h_small->err, sc->ctx->d_err, sizeof(unsigned int), cudaMemcpyDeviceToHost, sc->ctx->stream));
  return MCS_OK;
}

int scene_error(unsigned int bits) {
  if (bits & 2u)
    return fail(MCS_ERR_SCENE, "the reference's behaviour is undefined for this scene (vehicle without a usable behaviour, or a "
                               "merging command without a lane on that side)");
  if (bits & 1u)
    return fail(MCS_ERR_CAPACITY, "a vehicle met more distinct yielding candidates than the device table holds");
  return MCS_OK;
}
// after the stream has been synchronised: did ANY rollout of the last launch hit a case the reference leaves undefined?
int check_launch(const mcs_scene* sc) { return scene_error(sc->ctx->h_small->err); }

// synchronise; on failure the stream is drained as far as possible before the error is reported
int sync_stream(mcs_scene* sc) {
  cudaError_t e = cudaStreamSynchronize(sc->ctx->stream);
  if (e != cudaSuccess) return fail(MCS_ERR_CUDA, std::string("cudaStreamSynchronize: ") + cudaGetErrorString(e));
  return MCS_OK;
}
// an entry point that has queued copies into caller memory must not return before they have landed or been abandoned
#define TRY_SYNCED(expr)                                               \
  do {                                                                 \
    int rc_ = (expr);                                                  \
    if (rc_ != MCS_OK) { cudaStreamSynchronize(sc->ctx->stream); return rc_; } \
  } while (0)
#define CUDA_TRY_SYNCED(expr)                                                                                   \
  do {                                                                                                          \
    cudaError_t e_ = (expr);                                                                                    \
    if (e_ != cudaSuccess) {                                                                                    \
      cudaStreamSynchronize(sc->ctx->stream);                                                                   \
      return fail(MCS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));                            \
    }                                                                                                           \
  } while (0)

}  // namespace

extern "C" {

const char* mcs_last_error(void) { return g_err.c_str(); }
const char* mcs_version(void) { return "mcsim_b200 0.1 (sm_100a)"; }

int mcs_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int mcs_fp64_peak(int device, int iters, double* tflops_out, float* ms_out) {
  if (!tflops_out || iters <= 0) return fail(MCS_ERR_ARG, "bad arguments");
  if (mcs_device_count() <= device || device < 0) return fail(MCS_ERR_CUDA, "no such CUDA device");
  DeviceCtx* c = nullptr;
  int rc = get_ctx(device, &c);
  if (rc != MCS_OK) return rc;
  std::lock_guard<std::mutex> lock(c->mu);
  CUDA_TRY(cudaSetDevice(device));
  const int blocks = c->n_sm * 8, threads = 256;            // 2048 resident threads per SM
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {                        // the first repetition warms the clocks up
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
    fp64_peak_kernel<<<blocks, threads, 0, c->stream>>>((double*)c->d_counter, iters, 0.999999, 1e-6);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    if (rep > 0 && ms < best) best = ms;
  }
  const double flop = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
  *tflops_out = flop / ((double)best * 1e-3) / 1e12;
  if (ms_out) *ms_out = best;
  return MCS_OK;
}

int mcs_scene_create(const mcs_corridor* corridors, int n_corridors, const mcs_agent* agents, int n_agents, int device,
                     mcs_scene** out) {
  if (!out) return fail(MCS_ERR_ARG, "out is null");
  *out = nullptr;
  mcs_scene* sc = new mcs_scene();
  int rc = mcs::build_scene_host(corridors, n_corridors, agents, n_agents, sc->host);
  if (rc != MCS_OK) { std::string e = sc->host.error; delete sc; return fail(rc, e); }
  if (mcs_device_count() <= device || device < 0) { delete sc; return fail(MCS_ERR_CUDA, "no such CUDA device (libmcsim_b200 has no CPU fallback)"); }
  sc->device = device;
  if ((rc = get_ctx(device, &sc->ctx)) != MCS_OK) { delete sc; return rc; }
  DeviceCtx* c = sc->ctx;
  std::lock_guard<std::mutex> lock(c->mu);
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) { delete sc; return fail(MCS_ERR_CUDA, cudaGetErrorString(e)); }
  const mcs::SceneHost& h = sc->host;
  mcs::SceneDev& d = sc->dev;
  d.n = h.n; d.n_pad = 0; d.n_lanes = h.n_lanes; d.has_exit_lanes = h.has_exit_lanes ? 1 : 0;   // rows: upload_scene, at the first use
  for (int j = 0; j <= h.n_lanes; ++j) d.lane_start[j] = h.lane_start[j];
  for (int j = 0; j < h.n_lanes; ++j) {
    const mcs::LaneHost& L = h.lanes[j];
    d.lanes[j] = mcs::LaneDev{L.leftY, L.rightY, L.centerY, L.vLimit, L.endX, L.startX, L.id, L.type, L.leftIdx, L.rightIdx, L.leftmost, 0};
  }
  mcs::lane_neighbour_bits(d.lanes, h.n_lanes);
  *out = sc;
  return MCS_OK;
}

int mcs_scene_destroy(mcs_scene* sc) {
  if (!sc) return MCS_OK;
  if (sc->ctx) {
    std::lock_guard<std::mutex> lock(sc->ctx->mu);
    cudaSetDevice(sc->device);
    if (sc->d_blob) cudaFreeAsync(sc->d_blob, sc->ctx->stream);
  }
  delete sc;
  return MCS_OK;
}

int mcs_scene_visit_order(const mcs_scene* sc, int32_t* order_out, int n) {
  if (!sc || !order_out || n < sc->host.n) return fail(MCS_ERR_ARG, "bad arguments");
  for (int k = 0; k < sc->host.n; ++k) order_out[k] = sc->host.i[(size_t)mcs::I_INPUT * sc->host.n_pad + k];
  return MCS_OK;
}

int mcs_scene_prepare_host(const mcs_corridor* corridors, int n_corridors, const mcs_agent* agents, int n_agents,
                           int32_t* agent_order_out, int32_t* corridor_order_out, int32_t* left_ids_out,
                           int32_t* right_ids_out, int32_t* first_lane_ids_out, double* priors_out) {
  mcs::SceneHost h;
  int rc = mcs::build_scene_host(corridors, n_corridors, agents, n_agents, h);
  if (rc != MCS_OK) return fail(rc, h.error);
  for (int k = 0; k < h.n; ++k) {
    const int input = h.i[(size_t)mcs::I_INPUT * h.n_pad + k];
    if (agent_order_out) agent_order_out[k] = input;
    const int lane = h.i[(size_t)mcs::I_LANE * h.n_pad + k];
    if (first_lane_ids_out) first_lane_ids_out[input] = lane >= 0 ? h.lanes[lane].id : INT32_MIN;
    if (priors_out)
      for (int q = 0; q < 3; ++q) priors_out[3 * input + q] = h.f[(size_t)(mcs::F_PRIOR_L + q) * h.n_pad + k];
  }
  for (int j = 0; j < h.n_lanes; ++j) {
    int input = -1;
    for (int c = 0; c < n_corridors; ++c) if (corridors[c].id == h.lanes[j].id) { input = c; break; }
    if (corridor_order_out) corridor_order_out[j] = input;
    if (left_ids_out && input >= 0) left_ids_out[input] = h.lanes[j].leftIdx >= 0 ? h.lanes[h.lanes[j].leftIdx].id : INT32_MIN;
    if (right_ids_out && input >= 0) right_ids_out[input] = h.lanes[j].rightIdx >= 0 ? h.lanes[h.lanes[j].rightIdx].id : INT32_MIN;
  }
  return MCS_OK;
}

int mcs_reduce_features(const mcs_status* statuses, int n_candidates, int groups, int n_per_group, int steps,
                        mcs_feature* features_out) {
  if (!statuses || !features_out || n_candidates < 0 || groups <= 0 || n_per_group <= 0) return fail(MCS_ERR_ARG, "bad arguments");
  std::vector<mcs_group_partial> part(groups);
  unsigned int bits = 0;
  for (int c = 0; c < n_candidates; ++c) {
    for (int g = 0; g < groups; ++g)
      group_partial(statuses + ((size_t)c * groups + g) * n_per_group, n_per_group, steps, &part[g]);
    combine_groups(part.data(), groups, features_out + c);
    bits |= (unsigned int)features_out[c].reserved;
    features_out[c].reserved = 0;
  }
  return scene_error(bits);     // statuses gathered from other ranks carry their scene-error marks with them
}

int mcs_group_partials(const mcs_status* statuses, int n_candidates, int groups, int n_per_group, int steps,
                       mcs_group_partial* partials_out) {
  if (!statuses || !partials_out || n_candidates < 0 || groups <= 0 || n_per_group <= 0) return fail(MCS_ERR_ARG, "bad arguments");
  for (int i = 0; i < n_candidates * groups; ++i)
    group_partial(statuses + (size_t)i * n_per_group, n_per_group, steps, partials_out + i);
  return MCS_OK;
}

int mcs_combine_groups(const mcs_group_partial* partials, int n_candidates, int groups, mcs_feature* features_out) {
  if (!partials || !features_out || n_candidates < 0 || groups <= 0) return fail(MCS_ERR_ARG, "bad arguments");
  unsigned int bits = 0;
  for (int c = 0; c < n_candidates; ++c) {
    combine_groups(partials + (size_t)c * groups, groups, features_out + c);
    bits |= (unsigned int)features_out[c].reserved;
    features_out[c].reserved = 0;
  }
  return scene_error(bits);     // a rollout of some group (simulated on whichever rank) hit an undefined case
}

int mcs_scene_sync(mcs_scene* sc) {
  if (!sc) return fail(MCS_ERR_ARG, "scene is null");
  std::lock_guard<std::mutex> lock(sc->ctx->mu);
  CUDA_TRY(cudaSetDevice(sc->device));
  CUDA_TRY(cudaStreamSynchronize(sc->ctx->stream));
  return MCS_OK;
}

// shared by mcs_rollouts_groups{,_device}: rollouts of groups [first_group, first_group+n_groups) -> partials in `d_out`
static int groups_on_device(mcs_scene* sc, const mcs_candidate* cands, int n_cand, const mcs_sim_param* sp, int n_per_group,
                            uint64_t seed, int first_group, int n_groups, mcs_group_partial* d_out) {
  if (first_group < 0 || n_groups <= 0 || first_group + n_groups > MCS_GROUPS) return fail(MCS_ERR_ARG, "bad group range");
  const int64_t count = (int64_t)n_groups * n_per_group;
  const size_t n_status = (size_t)n_cand * count;
  int rc = ensure_capacity(sc, n_cand, n_status);
  if (rc != MCS_OK) return rc;
  int merge = 0;
  if ((rc = stage_candidates(sc, cands, n_cand, &merge)) != MCS_OK) return rc;
  if ((rc = launch_rollouts(sc, n_cand, sp, seed, (int64_t)first_group * n_per_group, count, sc->ctx->d_status, nullptr, nullptr,
                            nullptr, merge)) != MCS_OK) return rc;
  const int total = n_cand * n_groups;
  group_partials_kernel<<<(total + 63) / 64, 64, 0, sc->ctx->stream>>>(sc->ctx->d_status, n_cand, n_groups, n_per_group, sp->steps,
                                                                   d_out ? d_out : sc->ctx->d_part);
  CUDA_TRY(cudaGetLastError());
  return MCS_OK;
}

int mcs_rollouts_groups(mcs_scene* sc, const mcs_candidate* cands, int n_cand, const mcs_sim_param* sp, int n_per_group,
                        uint64_t seed, int first_group, int n_groups, mcs_group_partial* partials_out) {
  if (!sc || !cands || !sp || !partials_out || n_cand <= 0 || n_per_group <= 0) return fail(MCS_ERR_ARG, "bad arguments");
  std::lock_guard<std::mutex> lock(sc->ctx->mu);
  CUDA_TRY(cudaSetDevice(sc->device));
  int rc = groups_on_device(sc, cands, n_cand, sp, n_per_group, seed, first_group, n_groups, nullptr);
  if (rc != MCS_OK) return rc;
  CUDA_TRY_SYNCED(cudaMemcpyAsync(partials_out, sc->ctx->d_part, sizeof(mcs_group_partial) * n_cand * n_groups, cudaMemcpyDeviceToHost, sc->ctx->stream));
  if ((rc = sync_stream(sc)) != MCS_OK) return rc;
  return check_launch(sc);
}

int mcs_rollouts_groups_device(mcs_scene* sc, const mcs_candidate* cands, int n_cand, const mcs_sim_param* sp, int n_per_group,
                               uint64_t seed, int first_group, int n_groups, void* d_partials_out) {
  if (!sc || !cands || !sp || !d_partials_out || n_cand <= 0 || n_per_group <= 0) return fail(MCS_ERR_ARG, "bad arguments");
  std::lock_guard<std::mutex> lock(sc->ctx->mu);
  CUDA_TRY(cudaSetDevice(sc->device));
  int rc = groups_on_device(sc, cands, n_cand, sp, n_per_group, seed, first_group, n_groups, (mcs_group_partial*)d_partials_out);
  if (rc != MCS_OK) return rc;
  if ((rc = sync_stream(sc)) != MCS_OK) return rc;
  // Sharded callers exchange the partials next (a collective every rank must enter): the mark also travels inside them
  // (mcs_group_partial::reserved) and mcs_combine_groups refuses it on every rank, so this rank's own error is not lost
  // if its caller postpones the check until after the exchange.
  return check_launch(sc);
}

int mcs_rollouts(mcs_scene* sc, const mcs_candidate* cands, int n_cand, const mcs_sim_param* sp, int n_per_group,
                 uint64_t seed, mcs_feature* features_out, mcs_status* status_out) {
  if (!sc || !cands || !sp || !features_out || n_cand <= 0 || n_per_group <= 0) return fail(MCS_ERR_ARG, "bad arguments");
  std::lock_guard<std::mutex> lock(sc->ctx->mu);
  CUDA_TRY(cudaSetDevice(sc->device));
  const int64_t count = (int64_t)MCS_GROUPS * n_per_group;
  const size_t n_status = (size_t)n_cand * count;
  int rc = ensure_capacity(sc, n_cand, n_status);
  if (rc != MCS_OK) return rc;
  int merge = 0;
  if ((rc = stage_candidates(sc, cands, n_cand, &merge)) != MCS_OK) return rc;
  if ((rc = launch_rollouts(sc, n_cand, sp, seed, 0, count, sc->ctx->d_status, nullptr, nullptr, nullptr, merge)) != MCS_OK) return rc;
  group_partials_kernel<<<(n_cand * MCS_GROUPS + 63) / 64, 64, 0, sc->ctx->stream>>>(sc->ctx->d_status, n_cand, MCS_GROUPS, n_per_group,
                                                                                sp->steps, sc->ctx->d_part);
  combine_groups_kernel<<<(n_cand + 31) / 32, 32, 0, sc->ctx->stream>>>(sc->ctx->d_part, n_cand, MCS_GROUPS, sc->ctx->d_feat);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(sc->ctx->h_feat, sc->ctx->d_feat, sizeof(mcs_feature) * n_cand, cudaMemcpyDeviceToHost, sc->ctx->stream));
  if (status_out) CUDA_TRY_SYNCED(cudaMemcpyAsync(status_out, sc->ctx->d_status, sizeof(mcs_status) * n_status, cudaMemcpyDeviceToHost, sc->ctx->stream));
  if ((rc = sync_stream(sc)) != MCS_OK) return rc;
  memcpy(features_out, sc->ctx->h_feat, sizeof(mcs_feature) * n_cand);
  for (int c = 0; c < n_cand; ++c) features_out[c].reserved = 0;
  return check_launch(sc);     // any rollout of any candidate, not only the first
}

int mcs_rollouts_range(mcs_scene* sc, const mcs_candidate* cands, int n_cand, const mcs_sim_param* sp, uint64_t seed,
                       int64_t first, int64_t count, mcs_status* status_out) {
  if (!sc || !cands || !sp || !status_out || n_cand <= 0 || count <= 0) return fail(MCS_ERR_ARG, "bad arguments");
  std::lock_guard<std::mutex> lock(sc->ctx->mu);
  CUDA_TRY(cudaSetDevice(sc->device));
  const size_t n_status = (size_t)n_cand * count;
  int rc = ensure_capacity(sc, n_cand, n_status);
  if (rc != MCS_OK) return rc;
  int merge = 0;
  if ((rc = stage_candidates(sc, cands, n_cand, &merge)) != MCS_OK) return rc;
  if ((rc = launch_rollouts(sc, n_cand, sp, seed, first, count, sc->ctx->d_status, nullptr, nullptr, nullptr, merge)) != MCS_OK) return rc;
  CUDA_TRY_SYNCED(cudaMemcpyAsync(status_out, sc->ctx->d_status, sizeof(mcs_status) * n_status, cudaMemcpyDeviceToHost, sc->ctx->stream));
  if ((rc = sync_stream(sc)) != MCS_OK) return rc;
  return check_launch(sc);
}

int mcs_rollouts_device(mcs_scene* sc, const mcs_candidate* cands, int n_cand, const mcs_sim_param* sp, uint64_t seed,
                        int64_t first, int64_t count, void* d_status_out, void* cuda_stream, float* kernel_ms_out,
                        uint64_t* vehicle_steps_out) {
  if (!sc || !cands || !sp || n_cand <= 0 || count <= 0) return fail(MCS_ERR_ARG, "bad arguments");
  std::lock_guard<std::mutex> lock(sc->ctx->mu);
  CUDA_TRY(cudaSetDevice(sc->device));
  (void)cuda_stream;  // launches go to the scene's own stream; events below are recorded on that stream
  const size_t n_status = (size_t)n_cand * count;
  int rc = ensure_capacity(sc, n_cand, d_status_out ? 0 : n_status);
  if (rc != MCS_OK) return rc;
  int merge = 0;
  if ((rc = stage_candidates(sc, cands, n_cand, &merge)) != MCS_OK) return rc;
  CUDA_TRY(cudaMemsetAsync(sc->ctx->d_counter, 0, sizeof(unsigned long long), sc->ctx->stream));
  mcs_status* dst = d_status_out ? (mcs_status*)d_status_out : sc->ctx->d_status;
  CUDA_TRY(cudaEventRecord(sc->ctx->ev0, sc->ctx->stream));
  if ((rc = launch_rollouts(sc, n_cand, sp, seed, first, count, dst, nullptr, nullptr, sc->ctx->d_counter, merge)) != MCS_OK) return rc;
  CUDA_TRY(cudaEventRecord(sc->ctx->ev1, sc->ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(&sc->ctx->h_small->steps, sc->ctx->d_counter, sizeof(unsigned long long), cudaMemcpyDeviceToHost, sc->ctx->stream));
  if ((rc = sync_stream(sc)) != MCS_OK) return rc;
  if (kernel_ms_out) CUDA_TRY(cudaEventElapsedTime(kernel_ms_out, sc->ctx->ev0, sc->ctx->ev1));
  if (vehicle_steps_out) *vehicle_steps_out = sc->ctx->h_small->steps;
  return check_launch(sc);
}

int mcs_trajectories(mcs_scene* sc, const mcs_candidate* cand, const mcs_sim_param* sp, uint64_t seed, int64_t rollout,
                     double* states_out, int32_t* n_states_out, mcs_status* status_out) {
  if (!sc || !cand || !sp || !states_out) return fail(MCS_ERR_ARG, "bad arguments");
  std::lock_guard<std::mutex> lock(sc->ctx->mu);
  CUDA_TRY(cudaSetDevice(sc->device));
  int rc = ensure_capacity(sc, 1, 1);
  if (rc != MCS_OK) return rc;
  int merge = 0;
  if ((rc = stage_candidates(sc, cand, 1, &merge)) != MCS_OK) return rc;
  const size_t n_d = (size_t)sc->host.n * (sp->steps + 1) * 4;
  // one allocation for the states and their count; released on every path out of this function
  struct Scratch {
    unsigned char* p = nullptr;
    cudaStream_t stream;
    ~Scratch() { if (p) { cudaStreamSynchronize(stream); cudaFree(p); } }
  } scratch;
  scratch.stream = sc->ctx->stream;
  CUDA_TRY(cudaMalloc(&scratch.p, n_d * sizeof(double) + 16));
  double* d_traj = (double*)scratch.p;
  int32_t* d_n = (int32_t*)(scratch.p + n_d * sizeof(double));
  CUDA_TRY(cudaMemsetAsync(scratch.p, 0, n_d * sizeof(double) + 16, sc->ctx->stream));
  if ((rc = launch_rollouts(sc, 1, sp, seed, rollout, 1, sc->ctx->d_status, d_traj, d_n, nullptr, merge)) != MCS_OK) return rc;
  DeviceCtx::Small* hs = sc->ctx->h_small;
  CUDA_TRY_SYNCED(cudaMemcpyAsync(states_out, d_traj, n_d * sizeof(double), cudaMemcpyDeviceToHost, sc->ctx->stream));
  CUDA_TRY_SYNCED(cudaMemcpyAsync(&hs->n_states, d_n, sizeof(int32_t), cudaMemcpyDeviceToHost, sc->ctx->stream));
  CUDA_TRY_SYNCED(cudaMemcpyAsync(&hs->status, sc->ctx->d_status, sizeof(mcs_status), cudaMemcpyDeviceToHost, sc->ctx->stream));
  if ((rc = sync_stream(sc)) != MCS_OK) return rc;
  if (n_states_out) *n_states_out = hs->n_states;
  if (status_out) *status_out = hs->status;
  return check_launch(sc);
}

static int run_decision(mcs_scene* sc, int32_t agent_id, mcs::DecideParams dp, int32_t gap_id, mcs::DecideOut* res) {
  if (!sc) return fail(MCS_ERR_ARG, "scene is null");
  std::lock_guard<std::mutex> lock(sc->ctx->mu);
  CUDA_TRY(cudaSetDevice(sc->device));
  const mcs::SceneHost& h = sc->host;
  dp.slot = -1; dp.gap_slot = -1;
  for (int k = 0; k < h.n; ++k) {
    const int id = h.i[(size_t)mcs::I_ID * h.n_pad + k];
    if (id == agent_id && dp.slot < 0) dp.slot = k;
    if (gap_id >= 0 && id == gap_id && dp.gap_slot < 0) dp.gap_slot = k;
  }
  if (dp.slot < 0) return fail(MCS_ERR_ARG, "No agent with id: " + std::to_string(agent_id));   // the reference traps (ud2, 0xb9a37)
  if (sc->dev.n_pad == 0) { const int rc = upload_scene(sc, h.n_pad); if (rc != MCS_OK) return rc; }   // any row length serves
  const size_t smem = rollout_smem_bytes(sc->dev.n_pad, 0);
  CUDA_TRY(cudaFuncSetAttribute(mcs::decide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mcs::decide_kernel<<<1, sc->dev.n_pad, smem, sc->ctx->stream>>>(sc->dev, dp, (mcs::DecideOut*)sc->ctx->d_decide);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(&sc->ctx->h_small->decide, sc->ctx->d_decide, sizeof *res, cudaMemcpyDeviceToHost, sc->ctx->stream));
  { const int rc = sync_stream(sc); if (rc != MCS_OK) return rc; }
  *res = sc->ctx->h_small->decide;
  if (res->status != MCS_OK)
    return fail(res->status, "the reference's behaviour is undefined here (vehicle not on a lane, or no lane on that side)");
  return MCS_OK;
}

int mcs_decide_mobil(mcs_scene* sc, int32_t agent_id, const mcs_action* idm_action, int require_rss, uint64_t seed,
                     mcs_action_with_type* out) {
  if (!idm_action || !out) return fail(MCS_ERR_ARG, "bad arguments");
  mcs::DecideParams dp{};
  dp.kind = 0; dp.ax = idm_action->ax; dp.vy = idm_action->vy; dp.require_rss = require_rss ? 1 : 0; dp.key = seed;
  mcs::DecideOut r{};
  int rc = run_decision(sc, agent_id, dp, -2, &r);
  if (rc != MCS_OK) return rc;
  out->ax = r.ax; out->vy = r.vy; out->commit = r.commit; out->commit_keep_lane_step = r.keepStep;
  out->target_lane_id = r.targetLane; out->reserved = 0;
  return MCS_OK;
}

int mcs_decide_fixed_direction(mcs_scene* sc, int32_t agent_id, int32_t action, mcs_action* out) {
  if (!out) return fail(MCS_ERR_ARG, "bad arguments");
  mcs::DecideParams dp{};
  dp.kind = 1; dp.action = action;
  mcs::DecideOut r{};
  int rc = run_decision(sc, agent_id, dp, -2, &r);
  if (rc != MCS_OK) return rc;
  out->ax = r.ax; out->vy = r.vy;
  return MCS_OK;
}

int mcs_decide_fixed_gap(mcs_scene* sc, int32_t agent_id, int32_t direction, int32_t gap_id, mcs_action* out) {
  if (!out) return fail(MCS_ERR_ARG, "bad arguments");
  mcs::DecideParams dp{};
  dp.kind = 2; dp.action = direction;
  mcs::DecideOut r{};
  int rc = run_decision(sc, agent_id, dp, gap_id, &r);
  if (rc != MCS_OK) return rc;
  out->ax = r.ax; out->vy = r.vy;
  return MCS_OK;
}

int mcs_decide_closest_gap(mcs_scene* sc, int32_t agent_id, int32_t direction, mcs_action* out) {
  if (!out) return fail(MCS_ERR_ARG, "bad arguments");
  mcs::DecideParams dp{};
  dp.kind = 3; dp.action = direction;
  mcs::DecideOut r{};
  int rc = run_decision(sc, agent_id, dp, -2, &r);
  if (rc != MCS_OK) return rc;
  out->ax = r.ax; out->vy = r.vy;
  return MCS_OK;
}

}  // extern "C"
