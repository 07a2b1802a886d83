// C ABI of libmcsim_b200.so (include/mcsim.h).  CUDA runtime only — no torch types cross this boundary.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "rollout_kernel.cuh"
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

size_t rollout_smem_bytes(int n_pad, int steps) {
  size_t doubles = (size_t)n_pad * (4 + 5 + 6) + 2 * (size_t)((steps + 2) & ~1) + 16;
  size_t ints = 16 + 16 + 8 + (size_t)n_pad;
  return doubles * 8 + ints * 4 + 5 * (size_t)n_pad;
}

// per-candidate two-level mean, fixed order: runMTimes 0xc2a6e-0xc2b1f per group, runMultiThreads 0xbca08-0xbca92
__host__ __device__ inline void reduce_one(const mcs_status* s, int groups, int m, int steps, mcs_feature* out) {
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};
  int nOk = 0;
  for (int g = 0; g < groups; ++g) {
    const mcs_status* p = s + (size_t)g * m;
    if (m <= 0 || !p[0].ok) continue;
    int nFall = 0, nFin = 0, sumStep = 0;
    double u = 0.0, c = 0.0, uo = 0.0, co = 0.0;
    for (int i = 0; i < m; ++i) {
      nFall += p[i].fallBack != 0;
      if (p[i].finish) { nFin++; sumStep += p[i].finishStep; }
      u = u + p[i].utility; c = c + p[i].comfort; uo = uo + p[i].utilityObjects; co = co + p[i].comfortObjects;
    }
    const double M = (double)m;
    acc[0] = acc[0] + (double)nFall / M;
    acc[1] = acc[1] + (double)nFin / M;
    acc[2] = acc[2] + u / M;
    acc[3] = acc[3] + c / M;
    acc[4] = acc[4] + (double)(sumStep <= 0 ? 1 : sumStep) / (double)(nFin * steps <= 0 ? 1 : nFin * steps);
    acc[5] = acc[5] + uo / M;
    acc[6] = acc[6] + co / M;
    nOk++;
  }
  const double d = (double)nOk;
  out->fallBackRate = acc[0] / d; out->successRate = acc[1] / d; out->utility = acc[2] / d; out->comfort = acc[3] / d;
  out->expectedStepRatio = acc[4] / d; out->utilityObjects = acc[5] / d; out->comfortObjects = acc[6] / d;
  out->fromNSims = nOk * m; out->reserved = 0;
}

// one thread per (candidate, group) computes the group's means; thread 0 of the candidate combines them in group order
__global__ void reduce_features_kernel(const mcs_status* __restrict__ st, int groups, int m, int steps, mcs_feature* out) {
  __shared__ double part[MCS_GROUPS][7];
  __shared__ int okf[MCS_GROUPS];
  const int c = blockIdx.x, g = threadIdx.x;
  const mcs_status* p = st + ((size_t)c * groups + g) * m;
  if (g < groups) {
    int ok = (m > 0 && p[0].ok) ? 1 : 0;
    if (ok) {
      int nFall = 0, nFin = 0, sumStep = 0;
      double u = 0.0, cf = 0.0, uo = 0.0, co = 0.0;
      for (int i = 0; i < m; ++i) {
        nFall += p[i].fallBack != 0;
        if (p[i].finish) { nFin++; sumStep += p[i].finishStep; }
        u = u + p[i].utility; cf = cf + p[i].comfort; uo = uo + p[i].utilityObjects; co = co + p[i].comfortObjects;
      }
      const double M = (double)m;
      part[g][0] = (double)nFall / M; part[g][1] = (double)nFin / M; part[g][2] = u / M; part[g][3] = cf / M;
      part[g][4] = (double)(sumStep <= 0 ? 1 : sumStep) / (double)(nFin * steps <= 0 ? 1 : nFin * steps);
      part[g][5] = uo / M; part[g][6] = co / M;
    }
    okf[g] = ok;
  }
  __syncthreads();
  if (g == 0) {
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    int nOk = 0;
    for (int q = 0; q < groups; ++q)
      if (okf[q]) { for (int t = 0; t < 7; ++t) acc[t] = acc[t] + part[q][t]; nOk++; }
    const double d = (double)nOk;
    mcs_feature f;
    f.fallBackRate = acc[0] / d; f.successRate = acc[1] / d; f.utility = acc[2] / d; f.comfort = acc[3] / d;
    f.expectedStepRatio = acc[4] / d; f.utilityObjects = acc[5] / d; f.comfortObjects = acc[6] / d;
    f.fromNSims = nOk * m; f.reserved = 0;
    out[c] = f;
  }
}

}  // namespace

struct mcs_scene {
  int device = 0;
  mcs::SceneHost host;
  mcs::SceneDev dev{};
  double* d_f = nullptr;
  int32_t* d_i = nullptr;
  uint8_t* d_vec = nullptr;
  cudaStream_t stream = nullptr;
  // reusable device / pinned staging
  mcs::CandDev* d_cand = nullptr; int cand_cap = 0;
  mcs_status* d_status = nullptr; size_t status_cap = 0;
  mcs_feature* d_feat = nullptr; int feat_cap = 0;
  unsigned long long* d_counter = nullptr;
  void* d_decide = nullptr;             // mcs::DecideOut
  mcs::CandDev* h_cand = nullptr;      // pinned
  mcs_feature* h_feat = nullptr;       // pinned
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::mutex mu;
};

namespace {

int ensure_capacity(mcs_scene* sc, int n_cand, size_t n_status) {
  if (n_cand > sc->cand_cap) {
    if (sc->d_cand) cudaFree(sc->d_cand);
    if (sc->d_feat) cudaFree(sc->d_feat);
    if (sc->h_cand) cudaFreeHost(sc->h_cand);
    if (sc->h_feat) cudaFreeHost(sc->h_feat);
    sc->cand_cap = sc->feat_cap = 0;
    int cap = n_cand < 8 ? 8 : n_cand;
    CUDA_TRY(cudaMalloc(&sc->d_cand, sizeof(mcs::CandDev) * cap));
    CUDA_TRY(cudaMalloc(&sc->d_feat, sizeof(mcs_feature) * cap));
    CUDA_TRY(cudaMallocHost(&sc->h_cand, sizeof(mcs::CandDev) * cap));
    CUDA_TRY(cudaMallocHost(&sc->h_feat, sizeof(mcs_feature) * cap));
    sc->cand_cap = sc->feat_cap = cap;
  }
  if (n_status > sc->status_cap) {
    if (sc->d_status) cudaFree(sc->d_status);
    sc->status_cap = 0;
    CUDA_TRY(cudaMalloc(&sc->d_status, sizeof(mcs_status) * n_status));
    sc->status_cap = n_status;
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
    sc->h_cand[c] = d;
  }
  CUDA_TRY(cudaMemcpyAsync(sc->d_cand, sc->h_cand, sizeof(mcs::CandDev) * n_cand, cudaMemcpyHostToDevice, sc->stream));
  *merge_out = merge;
  return MCS_OK;
}

template <bool kTraj, bool kMerge>
int launch_one(mcs_scene* sc, dim3 grid, dim3 block, size_t smem, const mcs::LaunchParams& lp) {
  CUDA_TRY(cudaFuncSetAttribute(mcs::rollout_kernel<kTraj, kMerge>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mcs::rollout_kernel<kTraj, kMerge><<<grid, block, smem, sc->stream>>>(sc->dev, sc->d_cand, lp);
  CUDA_TRY(cudaGetLastError());
  return MCS_OK;
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
  const size_t smem = rollout_smem_bytes(sc->dev.n_pad, sp->steps);
  dim3 grid((unsigned)count, (unsigned)n_cand), block((unsigned)sc->dev.n_pad);
  if (d_traj) return merge ? launch_one<true, true>(sc, grid, block, smem, lp) : launch_one<true, false>(sc, grid, block, smem, lp);
  return merge ? launch_one<false, true>(sc, grid, block, smem, lp) : launch_one<false, false>(sc, grid, block, smem, lp);
}

int check_unsupported(const mcs_status* st, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    if (st[i].overflow == 2)
      return fail(MCS_ERR_SCENE, "the reference's behaviour is undefined for this scene (vehicle without a usable behaviour, or a "
                                 "merging command without a lane on that side)");
    if (st[i].overflow == 1)
      return fail(MCS_ERR_CAPACITY, "a vehicle met more distinct yielding candidates than the device table holds");
  }
  return MCS_OK;
}

}  // namespace

extern "C" {

const char* mcs_last_error(void) { return g_err.c_str(); }
const char* mcs_version(void) { return "mcsim_b200 0.1 (sm_100a)"; }

int mcs_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
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
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) { delete sc; return fail(MCS_ERR_CUDA, cudaGetErrorString(e)); }
  const mcs::SceneHost& h = sc->host;
  auto bail = [&](const char* what, cudaError_t err) { std::string m = std::string(what) + ": " + cudaGetErrorString(err); mcs_scene_destroy(sc); return fail(MCS_ERR_CUDA, m); };
  if ((e = cudaStreamCreateWithFlags(&sc->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  if ((e = cudaMalloc(&sc->d_f, h.f.size() * sizeof(double))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaMalloc(&sc->d_i, h.i.size() * sizeof(int32_t))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaMalloc(&sc->d_vec, h.lane_vec.size())) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaMalloc(&sc->d_counter, sizeof(unsigned long long))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaMemcpyAsync(sc->d_f, h.f.data(), h.f.size() * sizeof(double), cudaMemcpyHostToDevice, sc->stream)) != cudaSuccess) return bail("H2D", e);
  if ((e = cudaMemcpyAsync(sc->d_i, h.i.data(), h.i.size() * sizeof(int32_t), cudaMemcpyHostToDevice, sc->stream)) != cudaSuccess) return bail("H2D", e);
  if ((e = cudaMemcpyAsync(sc->d_vec, h.lane_vec.data(), h.lane_vec.size(), cudaMemcpyHostToDevice, sc->stream)) != cudaSuccess) return bail("H2D", e);
  if ((e = cudaEventCreate(&sc->ev0)) != cudaSuccess) return bail("cudaEventCreate", e);
  if ((e = cudaEventCreate(&sc->ev1)) != cudaSuccess) return bail("cudaEventCreate", e);
  mcs::SceneDev& d = sc->dev;
  d.n = h.n; d.n_pad = h.n_pad; d.n_lanes = h.n_lanes; d.has_exit_lanes = h.has_exit_lanes ? 1 : 0;
  d.f = sc->d_f; d.i = sc->d_i; d.lane_vec = sc->d_vec;
  for (int j = 0; j <= h.n_lanes; ++j) d.lane_start[j] = h.lane_start[j];
  for (int j = 0; j < h.n_lanes; ++j) {
    const mcs::LaneHost& L = h.lanes[j];
    d.lanes[j] = mcs::LaneDev{L.leftY, L.rightY, L.centerY, L.vLimit, L.endX, L.startX, L.id, L.type, L.leftIdx, L.rightIdx, L.leftmost, 0};
  }
  if ((e = cudaStreamSynchronize(sc->stream)) != cudaSuccess) return bail("sync", e);
  *out = sc;
  return MCS_OK;
}

int mcs_scene_destroy(mcs_scene* sc) {
  if (!sc) return MCS_OK;
  cudaSetDevice(sc->device);
  if (sc->stream) cudaStreamSynchronize(sc->stream);
  cudaFree(sc->d_f); cudaFree(sc->d_i); cudaFree(sc->d_vec); cudaFree(sc->d_cand); cudaFree(sc->d_status);
  cudaFree(sc->d_feat); cudaFree(sc->d_counter); cudaFree(sc->d_decide);
  if (sc->h_cand) cudaFreeHost(sc->h_cand);
  if (sc->h_feat) cudaFreeHost(sc->h_feat);
  if (sc->ev0) cudaEventDestroy(sc->ev0);
  if (sc->ev1) cudaEventDestroy(sc->ev1);
  if (sc->stream) cudaStreamDestroy(sc->stream);
  delete sc;
  return MCS_OK;
}

int mcs_scene_visit_order(const mcs_scene* sc, int32_t* order_out, int n) {
  if (!sc || !order_out || n < sc->host.n) return fail(MCS_ERR_ARG, "bad arguments");
  for (int k = 0; k < sc->host.n; ++k) order_out[k] = sc->host.i[(size_t)mcs::I_INPUT * sc->host.n_pad + k];
  return MCS_OK;
}

int mcs_reduce_features(const mcs_status* statuses, int n_candidates, int groups, int n_per_group, int steps,
                        mcs_feature* features_out) {
  if (!statuses || !features_out || n_candidates < 0 || groups <= 0 || n_per_group <= 0) return fail(MCS_ERR_ARG, "bad arguments");
  for (int c = 0; c < n_candidates; ++c)
    reduce_one(statuses + (size_t)c * groups * n_per_group, groups, n_per_group, steps, features_out + c);
  return MCS_OK;
}

int mcs_rollouts(mcs_scene* sc, const mcs_candidate* cands, int n_cand, const mcs_sim_param* sp, int n_per_group,
                 uint64_t seed, mcs_feature* features_out, mcs_status* status_out) {
  if (!sc || !cands || !sp || !features_out || n_cand <= 0 || n_per_group <= 0) return fail(MCS_ERR_ARG, "bad arguments");
  std::lock_guard<std::mutex> lock(sc->mu);
  CUDA_TRY(cudaSetDevice(sc->device));
  const int64_t count = (int64_t)MCS_GROUPS * n_per_group;
  const size_t n_status = (size_t)n_cand * count;
  int rc = ensure_capacity(sc, n_cand, n_status);
  if (rc != MCS_OK) return rc;
  int merge = 0;
  if ((rc = stage_candidates(sc, cands, n_cand, &merge)) != MCS_OK) return rc;
  if ((rc = launch_rollouts(sc, n_cand, sp, seed, 0, count, sc->d_status, nullptr, nullptr, nullptr, merge)) != MCS_OK) return rc;
  reduce_features_kernel<<<n_cand, 32, 0, sc->stream>>>(sc->d_status, MCS_GROUPS, n_per_group, sp->steps, sc->d_feat);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(sc->h_feat, sc->d_feat, sizeof(mcs_feature) * n_cand, cudaMemcpyDeviceToHost, sc->stream));
  std::vector<mcs_status> first_status;
  if (status_out) CUDA_TRY(cudaMemcpyAsync(status_out, sc->d_status, sizeof(mcs_status) * n_status, cudaMemcpyDeviceToHost, sc->stream));
  else {
    // the "unsupported" marker lives in the statuses; fetch the first rollout of each candidate
    first_status.resize(n_cand);
    for (int c = 0; c < n_cand; ++c)
      CUDA_TRY(cudaMemcpyAsync(&first_status[c], sc->d_status + (size_t)c * count, sizeof(mcs_status), cudaMemcpyDeviceToHost, sc->stream));
  }
  CUDA_TRY(cudaStreamSynchronize(sc->stream));
  memcpy(features_out, sc->h_feat, sizeof(mcs_feature) * n_cand);
  return status_out ? check_unsupported(status_out, n_status) : check_unsupported(first_status.data(), first_status.size());
}

int mcs_rollouts_range(mcs_scene* sc, const mcs_candidate* cands, int n_cand, const mcs_sim_param* sp, uint64_t seed,
                       int64_t first, int64_t count, mcs_status* status_out) {
  if (!sc || !cands || !sp || !status_out || n_cand <= 0 || count <= 0) return fail(MCS_ERR_ARG, "bad arguments");
  std::lock_guard<std::mutex> lock(sc->mu);
  CUDA_TRY(cudaSetDevice(sc->device));
  const size_t n_status = (size_t)n_cand * count;
  int rc = ensure_capacity(sc, n_cand, n_status);
  if (rc != MCS_OK) return rc;
  int merge = 0;
  if ((rc = stage_candidates(sc, cands, n_cand, &merge)) != MCS_OK) return rc;
  if ((rc = launch_rollouts(sc, n_cand, sp, seed, first, count, sc->d_status, nullptr, nullptr, nullptr, merge)) != MCS_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(status_out, sc->d_status, sizeof(mcs_status) * n_status, cudaMemcpyDeviceToHost, sc->stream));
  CUDA_TRY(cudaStreamSynchronize(sc->stream));
  return check_unsupported(status_out, n_status);
}

int mcs_rollouts_device(mcs_scene* sc, const mcs_candidate* cands, int n_cand, const mcs_sim_param* sp, uint64_t seed,
                        int64_t first, int64_t count, void* d_status_out, void* cuda_stream, float* kernel_ms_out,
                        uint64_t* vehicle_steps_out) {
  if (!sc || !cands || !sp || n_cand <= 0 || count <= 0) return fail(MCS_ERR_ARG, "bad arguments");
  std::lock_guard<std::mutex> lock(sc->mu);
  CUDA_TRY(cudaSetDevice(sc->device));
  (void)cuda_stream;  // launches go to the scene's own stream; events below are recorded on that stream
  const size_t n_status = (size_t)n_cand * count;
  int rc = ensure_capacity(sc, n_cand, d_status_out ? 0 : n_status);
  if (rc != MCS_OK) return rc;
  int merge = 0;
  if ((rc = stage_candidates(sc, cands, n_cand, &merge)) != MCS_OK) return rc;
  CUDA_TRY(cudaMemsetAsync(sc->d_counter, 0, sizeof(unsigned long long), sc->stream));
  mcs_status* dst = d_status_out ? (mcs_status*)d_status_out : sc->d_status;
  CUDA_TRY(cudaEventRecord(sc->ev0, sc->stream));
  if ((rc = launch_rollouts(sc, n_cand, sp, seed, first, count, dst, nullptr, nullptr, sc->d_counter, merge)) != MCS_OK) return rc;
  CUDA_TRY(cudaEventRecord(sc->ev1, sc->stream));
  unsigned long long steps = 0;
  CUDA_TRY(cudaMemcpyAsync(&steps, sc->d_counter, sizeof steps, cudaMemcpyDeviceToHost, sc->stream));
  mcs_status probe;
  CUDA_TRY(cudaMemcpyAsync(&probe, dst, sizeof probe, cudaMemcpyDeviceToHost, sc->stream));
  CUDA_TRY(cudaStreamSynchronize(sc->stream));
  if (kernel_ms_out) CUDA_TRY(cudaEventElapsedTime(kernel_ms_out, sc->ev0, sc->ev1));
  if (vehicle_steps_out) *vehicle_steps_out = steps;
  return check_unsupported(&probe, 1);
}

int mcs_trajectories(mcs_scene* sc, const mcs_candidate* cand, const mcs_sim_param* sp, uint64_t seed, int64_t rollout,
                     double* states_out, int32_t* n_states_out, mcs_status* status_out) {
  if (!sc || !cand || !sp || !states_out) return fail(MCS_ERR_ARG, "bad arguments");
  std::lock_guard<std::mutex> lock(sc->mu);
  CUDA_TRY(cudaSetDevice(sc->device));
  int rc = ensure_capacity(sc, 1, 1);
  if (rc != MCS_OK) return rc;
  int merge = 0;
  if ((rc = stage_candidates(sc, cand, 1, &merge)) != MCS_OK) return rc;
  const size_t n_d = (size_t)sc->host.n * (sp->steps + 1) * 4;
  double* d_traj = nullptr; int32_t* d_n = nullptr;
  CUDA_TRY(cudaMalloc(&d_traj, n_d * sizeof(double)));
  CUDA_TRY(cudaMalloc(&d_n, sizeof(int32_t)));
  cudaMemsetAsync(d_traj, 0, n_d * sizeof(double), sc->stream);
  cudaMemsetAsync(d_n, 0, sizeof(int32_t), sc->stream);
  rc = launch_rollouts(sc, 1, sp, seed, rollout, 1, sc->d_status, d_traj, d_n, nullptr, merge);
  mcs_status st{};
  int32_t ns = 0;
  if (rc == MCS_OK) {
    cudaMemcpyAsync(states_out, d_traj, n_d * sizeof(double), cudaMemcpyDeviceToHost, sc->stream);
    cudaMemcpyAsync(&ns, d_n, sizeof ns, cudaMemcpyDeviceToHost, sc->stream);
    cudaMemcpyAsync(&st, sc->d_status, sizeof st, cudaMemcpyDeviceToHost, sc->stream);
  }
  cudaError_t e = cudaStreamSynchronize(sc->stream);
  cudaFree(d_traj); cudaFree(d_n);
  if (rc != MCS_OK) return rc;
  if (e != cudaSuccess) return fail(MCS_ERR_CUDA, cudaGetErrorString(e));
  if (n_states_out) *n_states_out = ns;
  if (status_out) *status_out = st;
  return check_unsupported(&st, 1);
}

static int run_decision(mcs_scene* sc, int32_t agent_id, mcs::DecideParams dp, int32_t gap_id, mcs::DecideOut* res) {
  if (!sc) return fail(MCS_ERR_ARG, "scene is null");
  std::lock_guard<std::mutex> lock(sc->mu);
  CUDA_TRY(cudaSetDevice(sc->device));
  const mcs::SceneHost& h = sc->host;
  dp.slot = -1; dp.gap_slot = -1;
  for (int k = 0; k < h.n; ++k) {
    const int id = h.i[(size_t)mcs::I_ID * h.n_pad + k];
    if (id == agent_id && dp.slot < 0) dp.slot = k;
    if (gap_id >= 0 && id == gap_id && dp.gap_slot < 0) dp.gap_slot = k;
  }
  if (dp.slot < 0) return fail(MCS_ERR_ARG, "No agent with id: " + std::to_string(agent_id));   // the reference traps (ud2, 0xb9a37)
  if (!sc->d_decide) CUDA_TRY(cudaMalloc(&sc->d_decide, sizeof(mcs::DecideOut)));
  const size_t smem = rollout_smem_bytes(sc->dev.n_pad, 0);
  CUDA_TRY(cudaFuncSetAttribute(mcs::decide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mcs::decide_kernel<<<1, sc->dev.n_pad, smem, sc->stream>>>(sc->dev, dp, (mcs::DecideOut*)sc->d_decide);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(res, sc->d_decide, sizeof *res, cudaMemcpyDeviceToHost, sc->stream));
  CUDA_TRY(cudaStreamSynchronize(sc->stream));
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
