// C ABI of the outer-environment geometry path (include/mcsim_env.h).  CUDA runtime only.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "env_kernels.cuh"
#include "walk_kernel.cuh"

namespace {

thread_local std::string g_env_err;
int efail(int code, const std::string& msg) { g_env_err = msg; return code; }
#define ENV_TRY(expr)                                                                                   \
  do {                                                                                                  \
    cudaError_t e_ = (expr);                                                                            \
    if (e_ != cudaSuccess) return efail(MCE_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)

typedef std::pair<double, double> Pt;

double dist(const Pt& p, const Pt& q) {                       // utility.py:8-9
  const double dx = p.first - q.first, dy = p.second - q.second;
  return std::sqrt(dx * dx + dy * dy);
}

double seg_point_dist(const Pt& a, const Pt& b, const Pt& p) {
  const double dx = b.first - a.first, dy = b.second - a.second;
  const double d2 = dx * dx + dy * dy;
  double t = 0.0;
  if (d2 != 0.0) {
    t = ((p.first - a.first) * dx + (p.second - a.second) * dy) / d2;
    t = t < 1.0 ? t : 1.0;
    t = t > 0.0 ? t : 0.0;
  }
  const double qx = a.first + t * dx, qy = a.second + t * dy;
  const double ex = p.first - qx, ey = p.second - qy;
  return std::sqrt(ex * ex + ey * ey);
}

// LineString.simplify(0.05) — Douglas-Peucker (type.py:261-262)
std::vector<Pt> simplify(const std::vector<Pt>& pts, double tol) {
  if (pts.size() < 3) return pts;
  size_t imax = 0;
  double dmax = -1.0;
  for (size_t i = 1; i + 1 < pts.size(); ++i) {
    const double d = seg_point_dist(pts.front(), pts.back(), pts[i]);
    if (d > dmax) { imax = i; dmax = d; }
  }
  if (dmax > tol) {
    std::vector<Pt> a = simplify(std::vector<Pt>(pts.begin(), pts.begin() + imax + 1), tol);
    std::vector<Pt> b = simplify(std::vector<Pt>(pts.begin() + imax, pts.end()), tol);
    a.pop_back();
    a.insert(a.end(), b.begin(), b.end());
    return a;
  }
  return {pts.front(), pts.back()};
}

double line_length(const std::vector<Pt>& l) {
  double run = 0.0;
  for (size_t i = 0; i + 1 < l.size(); ++i) run = run + dist(l[i + 1], l[i]);
  return run;
}

Pt extend(const Pt& last, const Pt& prev, double d) {         // utility.py:113 / 120-121
  const double k = d / dist(last, prev);
  return Pt(last.first + k * (last.first - prev.first), last.second + k * (last.second - prev.second));
}

}  // namespace

struct mce_map {
  int device = 0;
  int n_lanes = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  mce::EnvMapHead* d_head = nullptr;
  mce::EnvMapTail* d_tail = nullptr;
  mce::EnvMapBorder* d_border = nullptr;
  mce::EnvLaneInfo* d_info = nullptr;   // Corridor.id / type / v_limit (mce_map_set_lane_info), read by the builders
  // the device-side walk of the outer step (mce_walk_step): vehicles, ids, attrs [8][n], noise block, random table, result
  char* d_walk = nullptr; size_t walk_cap = 0; int walk_n = 0;
  char* h_walk = nullptr;
  // state of the last match (resident)
  int n_agents = 0, n_scenes = 0;
  size_t cap_slots = 0, cap_scenes = 0, cap_arc_on = 0, cap_q = 0;
  double *d_x = nullptr, *d_y = nullptr, *d_arc = nullptr, *d_arc_on = nullptr, *d_nbd = nullptr;
  int32_t *d_lane = nullptr, *d_order = nullptr, *d_start = nullptr, *d_nbi = nullptr;
  // query / pointwise buffers
  char* d_q = nullptr;        // inputs then outputs of one query call, 144 bytes per point (+ alignment slack)
  char* h_q = nullptr;        // pinned mirror: one H2D and one D2H copy per call
  float ms[2] = {0.f, 0.f};
};

extern "C" {

const char* mce_last_error(void) { return g_env_err.c_str(); }

int mce_map_create(const mce_corridor* corridors, int n_corridors, int device, mce_map** out) {
  if (!corridors || !out) return efail(MCE_ERR_ARG, "null argument");
  if (n_corridors < 1 || n_corridors > MCE_MAX_LANES) return efail(MCE_ERR_CAPACITY, "corridor count outside 1..MCE_MAX_LANES");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return efail(MCE_ERR_CUDA, "no CUDA device (there is no CPU fallback)");
  if (device < 0 || device >= n_dev) return efail(MCE_ERR_ARG, "device index out of range");
  std::vector<mce::EnvMapHead> head(1);
  std::vector<mce::EnvMapTail> tail(1);
  std::vector<mce::EnvMapBorder> border(1);
  mce::EnvMapBorder& B = border[0];
  std::memset(&B, 0, sizeof(B));
  mce::EnvMapHead& H = head[0];
  mce::EnvMapTail& T = tail[0];
  std::memset(&H, 0, sizeof(H));
  std::memset(&T, 0, sizeof(T));
  H.n_lanes = n_corridors;
  for (int c = 0; c < n_corridors; ++c) {
    const mce_corridor& k = corridors[c];
    if (k.n_center < 2 || k.n_center > MCE_MAX_POINTS || k.n_left < 1 || k.n_left > MCE_MAX_POINTS || k.n_right < 1 ||
        k.n_right > MCE_MAX_POINTS)
      return efail(MCE_ERR_CAPACITY, "polyline point count outside the compiled limits");
    if (k.left_index < -1 || k.left_index >= n_corridors || k.right_index < -1 || k.right_index >= n_corridors)
      return efail(MCE_ERR_ARG, "neighbour corridor index out of range");
    H.left[c] = k.left_index;
    H.right[c] = k.right_index;
    std::vector<Pt> center, shell;
    for (int i = 0; i < k.n_center; ++i) center.push_back(Pt(k.center[i][0], k.center[i][1]));
    for (int i = 0; i < k.n_left; ++i) shell.push_back(Pt(k.left[i][0], k.left[i][1]));
    for (int i = k.n_right - 1; i >= 0; --i) shell.push_back(Pt(k.right[i][0], k.right[i][1]));
    if (shell.size() > 1 && shell.front() == shell.back()) shell.pop_back();
    H.n_center[c] = k.n_center;
    H.n_shell[c] = (int)shell.size();
    for (size_t i = 0; i < center.size(); ++i) { H.center[c][i][0] = center[i].first; H.center[c][i][1] = center[i].second; }
    const Pt ext = extend(center[center.size() - 1], center[center.size() - 2], 100.0);   // type.py:154
    H.center[c][center.size()][0] = ext.first;
    H.center[c][center.size()][1] = ext.second;
    for (size_t i = 0; i < shell.size(); ++i) { H.shell[c][i][0] = shell[i].first; H.shell[c][i][1] = shell[i].second; }
    // Frenet reference line (mcs_wrapper.py:42,109,211)
    std::vector<Pt> ref;
    ref.push_back(extend(center[0], center[1], 1000.0));
    ref.insert(ref.end(), center.begin(), center.end());
    ref.push_back(extend(center[center.size() - 1], center[center.size() - 2], 1000.0));
    const std::vector<Pt> refs = simplify(ref, 0.05);
    T.n_ref[c] = (int)ref.size();
    T.n_refs[c] = (int)refs.size();
    double run = 0.0;
    for (size_t i = 0; i < ref.size(); ++i) {
      if (i) run = run + dist(ref[i - 1], ref[i]);
      T.ref[c][i][0] = ref[i].first; T.ref[c][i][1] = ref[i].second; T.ref_dis[c][i] = run;
    }
    for (size_t i = 0; i < refs.size(); ++i) { T.refs[c][i][0] = refs[i].first; T.refs[c][i][1] = refs[i].second; }
    T.refs_len[c] = line_length(refs);
    // centerSimplified (type.py:153)
    const std::vector<Pt> cs = simplify(center, 0.05);
    T.n_center[c] = k.n_center;
    T.n_cs[c] = (int)cs.size();
    run = 0.0;
    for (size_t i = 0; i < center.size(); ++i) {
      if (i) run = run + dist(center[i - 1], center[i]);
      T.cpath[c][i][0] = center[i].first; T.cpath[c][i][1] = center[i].second; T.cdis[c][i] = run;
      const size_t a = i == 0 ? 0 : i - 1, b = i == 0 ? 1 : i;                           // yaw_list, type.py:267-272
      const double yaw = std::atan2(center[b].second - center[a].second, center[b].first - center[a].first);
      T.cyaw[c][i] = yaw; T.ccos[c][i] = std::cos(yaw); T.csin[c][i] = std::sin(yaw);
    }
    for (size_t i = 0; i < cs.size(); ++i) { T.cs[c][i][0] = cs[i].first; T.cs[c][i][1] = cs[i].second; }
    T.cs_len[c] = line_length(cs);
    // leftSimplified / rightSimplified (type.py:151-152) and the centre line again, for the distance kernels
    for (int sd = 0; sd < 2; ++sd) {
      std::vector<Pt> path;
      const int np = sd == 0 ? k.n_left : k.n_right;
      if (np < 2) return efail(MCE_ERR_CAPACITY, "a border polyline needs at least two points");
      for (int i = 0; i < np; ++i) path.push_back(sd == 0 ? Pt(k.left[i][0], k.left[i][1]) : Pt(k.right[i][0], k.right[i][1]));
      const std::vector<Pt> sp = simplify(path, 0.05);
      B.n_path[sd][c] = np;
      B.n_simp[sd][c] = (int)sp.size();
      run = 0.0;
      for (size_t i = 0; i < path.size(); ++i) {
        if (i) run = run + dist(path[i - 1], path[i]);
        B.path[sd][c][i][0] = path[i].first; B.path[sd][c][i][1] = path[i].second; B.dis[sd][c][i] = run;
      }
      for (size_t i = 0; i < sp.size(); ++i) { B.simp[sd][c][i][0] = sp[i].first; B.simp[sd][c][i][1] = sp[i].second; }
      B.simp_len[sd][c] = line_length(sp);
    }
    B.n_center[c] = T.n_center[c]; B.n_cs[c] = T.n_cs[c]; B.cs_len[c] = T.cs_len[c];
    std::memcpy(B.cpath[c], T.cpath[c], sizeof(B.cpath[c])); std::memcpy(B.cdis[c], T.cdis[c], sizeof(B.cdis[c]));
    std::memcpy(B.cs[c], T.cs[c], sizeof(B.cs[c]));
  }
  ENV_TRY(cudaSetDevice(device));
  mce_map* m = new mce_map();
  m->device = device;
  m->n_lanes = n_corridors;
  ENV_TRY(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
  for (int i = 0; i < 3; ++i) ENV_TRY(cudaEventCreate(&m->ev[i]));
  ENV_TRY(cudaMalloc(&m->d_head, sizeof(mce::EnvMapHead)));
  ENV_TRY(cudaMalloc(&m->d_tail, sizeof(mce::EnvMapTail)));
  ENV_TRY(cudaMemcpyAsync(m->d_head, &H, sizeof(H), cudaMemcpyHostToDevice, m->stream));
  ENV_TRY(cudaMemcpyAsync(m->d_tail, &T, sizeof(T), cudaMemcpyHostToDevice, m->stream));
  ENV_TRY(cudaMalloc(&m->d_border, sizeof(mce::EnvMapBorder)));
  ENV_TRY(cudaMemcpyAsync(m->d_border, &B, sizeof(B), cudaMemcpyHostToDevice, m->stream));
  ENV_TRY(cudaStreamSynchronize(m->stream));
  *out = m;
  return MCE_OK;
}

int mce_map_destroy(mce_map* m) {
  if (!m) return MCE_OK;
  cudaSetDevice(m->device);
  if (m->h_walk) cudaFreeHost(m->h_walk);
  void* bufs[] = {m->d_walk, m->d_info, m->d_head, m->d_tail, m->d_border, m->d_x, m->d_y, m->d_arc, m->d_arc_on, m->d_nbd, m->d_lane, m->d_order, m->d_start,
                  m->d_nbi, m->d_q};
  if (m->h_q) cudaFreeHost(m->h_q);
  for (void* b : bufs) if (b) cudaFree(b);
  for (int i = 0; i < 3; ++i) if (m->ev[i]) cudaEventDestroy(m->ev[i]);
  if (m->stream) cudaStreamDestroy(m->stream);
  delete m;
  return MCE_OK;
}

static int ensure_match_buffers(mce_map* m, size_t slots, size_t scenes, size_t arc_on) {
  if (slots > m->cap_slots) {
    void* old[] = {m->d_x, m->d_y, m->d_arc, m->d_nbd, m->d_lane, m->d_order, m->d_nbi};
    for (void* b : old) if (b) cudaFree(b);
    m->d_x = m->d_y = m->d_arc = m->d_nbd = nullptr; m->d_lane = m->d_order = m->d_nbi = nullptr; m->cap_slots = 0;
    ENV_TRY(cudaMalloc(&m->d_x, slots * 8)); ENV_TRY(cudaMalloc(&m->d_y, slots * 8)); ENV_TRY(cudaMalloc(&m->d_arc, slots * 8));
    ENV_TRY(cudaMalloc(&m->d_nbd, slots * 8 * MCE_NEIGHBORS)); ENV_TRY(cudaMalloc(&m->d_nbi, slots * 4 * MCE_NEIGHBORS));
    ENV_TRY(cudaMalloc(&m->d_lane, slots * 4)); ENV_TRY(cudaMalloc(&m->d_order, slots * 4));
    m->cap_slots = slots;
  }
  if (scenes > m->cap_scenes) {
    if (m->d_start) cudaFree(m->d_start);
    m->d_start = nullptr; m->cap_scenes = 0;
    ENV_TRY(cudaMalloc(&m->d_start, scenes * 4 * (MCE_MAX_LANES + 1)));
    m->cap_scenes = scenes;
  }
  if (arc_on > m->cap_arc_on) {
    if (m->d_arc_on) cudaFree(m->d_arc_on);
    m->d_arc_on = nullptr; m->cap_arc_on = 0;
    ENV_TRY(cudaMalloc(&m->d_arc_on, arc_on * 8));
    m->cap_arc_on = arc_on;
  }
  return MCE_OK;
}

static int ensure_query_buffers(mce_map* m, size_t n) {
  if (n > m->cap_q) {
    if (m->d_q) cudaFree(m->d_q);
    if (m->h_q) cudaFreeHost(m->h_q);
    m->d_q = nullptr; m->h_q = nullptr; m->cap_q = 0;
    const size_t cap = n < 64 ? 64 : n;
    ENV_TRY(cudaMalloc(&m->d_q, cap * 144 + 512));
    ENV_TRY(cudaMallocHost(&m->h_q, cap * 144 + 512));
    m->cap_q = cap;
  }
  return MCE_OK;
}

extern "C++" {
namespace {
// One query call: the inputs are packed into the pinned mirror and go up in ONE copy, the outputs come back in ONE copy
// (a sequential outer loop makes thousands of one-vehicle calls; per-array copies were most of their cost).
struct Query {
  mce_map* m;
  size_t in_bytes = 0, out_bytes = 0;
  struct Out { void* user; size_t off, bytes; };
  Out outs[4];
  int n_outs = 0;
  explicit Query(mce_map* map) : m(map) {}
  template <typename T>
  T* in(const T* src, size_t count) {
    std::memcpy(m->h_q + in_bytes, src, count * sizeof(T));
    T* d = reinterpret_cast<T*>(m->d_q + in_bytes);
    in_bytes += (count * sizeof(T) + 15) & ~size_t(15);
    return d;
  }
  template <typename T>
  T* in_strided(const T* src, size_t count, size_t stride, size_t offset) {     // element i = src[i * stride + offset]
    T* h = reinterpret_cast<T*>(m->h_q + in_bytes);
    for (size_t i = 0; i < count; ++i) h[i] = src[i * stride + offset];
    T* d = reinterpret_cast<T*>(m->d_q + in_bytes);
    in_bytes += (count * sizeof(T) + 15) & ~size_t(15);
    return d;
  }
  template <typename T>
  T* out(T* user, size_t count) {                                               // declare after every in()
    outs[n_outs++] = Out{user, in_bytes + out_bytes, count * sizeof(T)};
    T* d = reinterpret_cast<T*>(m->d_q + in_bytes + out_bytes);
    out_bytes += (count * sizeof(T) + 15) & ~size_t(15);
    return d;
  }
  int upload() { ENV_TRY(cudaMemcpyAsync(m->d_q, m->h_q, in_bytes, cudaMemcpyHostToDevice, m->stream)); return MCE_OK; }
  int finish() {
    ENV_TRY(cudaGetLastError());
    ENV_TRY(cudaMemcpyAsync(m->h_q + in_bytes, m->d_q + in_bytes, out_bytes, cudaMemcpyDeviceToHost, m->stream));
    ENV_TRY(cudaStreamSynchronize(m->stream));
    for (int k = 0; k < n_outs; ++k) std::memcpy(outs[k].user, m->h_q + outs[k].off, outs[k].bytes);
    return MCE_OK;
  }
};
}  // namespace
}  // extern "C++"

int mce_env_match(mce_map* m, const double* x, const double* y, int n_agents, int n_scenes, int32_t* lane_out, double* arc_out,
                  int32_t* lane_order_out, int32_t* lane_start_out, double* arc_on_lane_out, int32_t* nb_index_out,
                  double* nb_dis_out) {
  if (!m || !x || !y) return efail(MCE_ERR_ARG, "null argument");
  if (n_agents < 1 || n_scenes < 1) return efail(MCE_ERR_ARG, "empty batch");
  if (n_agents > MCE_MAX_AGENTS) return efail(MCE_ERR_CAPACITY, "more than MCE_MAX_AGENTS vehicles in a scene");
  ENV_TRY(cudaSetDevice(m->device));
  const size_t slots = (size_t)n_agents * n_scenes;
  const size_t arc_on = arc_on_lane_out ? slots * m->n_lanes : 0;
  int rc = ensure_match_buffers(m, slots, n_scenes, arc_on);
  if (rc) return rc;
  m->n_agents = n_agents; m->n_scenes = n_scenes;
  cudaStream_t s = m->stream;
  ENV_TRY(cudaMemcpyAsync(m->d_x, x, slots * 8, cudaMemcpyHostToDevice, s));
  ENV_TRY(cudaMemcpyAsync(m->d_y, y, slots * 8, cudaMemcpyHostToDevice, s));
  const size_t dyn = (size_t)n_agents * 12 + 8;
  ENV_TRY(cudaEventRecord(m->ev[0], s));
  mce::env_match_kernel<<<n_scenes, mce::kBT, dyn, s>>>(m->d_head, m->d_x, m->d_y, n_agents, m->d_lane, m->d_arc, m->d_order,
                                                         m->d_start, arc_on_lane_out ? m->d_arc_on : nullptr);
  ENV_TRY(cudaGetLastError());
  ENV_TRY(cudaEventRecord(m->ev[1], s));
  dim3 grid((n_agents + mce::kBT - 1) / mce::kBT, n_scenes);
  mce::env_neighbors_kernel<<<grid, mce::kBT, 0, s>>>(m->d_head, n_agents, m->d_lane, m->d_arc, m->d_order, m->d_start, nullptr,
                                                      nullptr, m->d_x, m->d_y, 0, m->d_nbi, m->d_nbd);
  ENV_TRY(cudaGetLastError());
  ENV_TRY(cudaEventRecord(m->ev[2], s));
  if (lane_out) ENV_TRY(cudaMemcpyAsync(lane_out, m->d_lane, slots * 4, cudaMemcpyDeviceToHost, s));
  if (arc_out) ENV_TRY(cudaMemcpyAsync(arc_out, m->d_arc, slots * 8, cudaMemcpyDeviceToHost, s));
  if (lane_order_out) ENV_TRY(cudaMemcpyAsync(lane_order_out, m->d_order, slots * 4, cudaMemcpyDeviceToHost, s));
  if (lane_start_out)
    ENV_TRY(cudaMemcpyAsync(lane_start_out, m->d_start, (size_t)n_scenes * 4 * (MCE_MAX_LANES + 1), cudaMemcpyDeviceToHost, s));
  if (arc_on_lane_out) ENV_TRY(cudaMemcpyAsync(arc_on_lane_out, m->d_arc_on, arc_on * 8, cudaMemcpyDeviceToHost, s));
  if (nb_index_out) ENV_TRY(cudaMemcpyAsync(nb_index_out, m->d_nbi, slots * 4 * MCE_NEIGHBORS, cudaMemcpyDeviceToHost, s));
  if (nb_dis_out) ENV_TRY(cudaMemcpyAsync(nb_dis_out, m->d_nbd, slots * 8 * MCE_NEIGHBORS, cudaMemcpyDeviceToHost, s));
  ENV_TRY(cudaStreamSynchronize(s));
  ENV_TRY(cudaEventElapsedTime(&m->ms[0], m->ev[0], m->ev[1]));
  ENV_TRY(cudaEventElapsedTime(&m->ms[1], m->ev[1], m->ev[2]));
  return MCE_OK;
}

int mce_env_neighbors(mce_map* m, const int32_t* scene, const int32_t* slot, const double* qx, const double* qy, int n_queries,
                      int32_t* nb_index_out, double* nb_dis_out) {
  if (!m || !scene || !slot || !qx || !qy || !nb_index_out || !nb_dis_out) return efail(MCE_ERR_ARG, "null argument");
  if (m->n_scenes == 0) return efail(MCE_ERR_ARG, "mce_env_neighbors before mce_env_match");
  if (n_queries < 1) return efail(MCE_ERR_ARG, "no queries");
  for (int i = 0; i < n_queries; ++i)
    if (scene[i] < 0 || scene[i] >= m->n_scenes || slot[i] < 0 || slot[i] >= m->n_agents)
      return efail(MCE_ERR_ARG, "query outside the matched batch");
  ENV_TRY(cudaSetDevice(m->device));
  int rc = ensure_query_buffers(m, n_queries);
  if (rc) return rc;
  const size_t q = n_queries;
  Query Q(m);
  double* d_qx = Q.in(qx, q); double* d_qy = Q.in(qy, q);
  int32_t* d_sc = Q.in(scene, q); int32_t* d_sl = Q.in(slot, q);
  int32_t* d_idx = Q.out(nb_index_out, q * MCE_NEIGHBORS); double* d_dis = Q.out(nb_dis_out, q * MCE_NEIGHBORS);
  if ((rc = Q.upload())) return rc;
  mce::env_neighbors_kernel<<<(n_queries + mce::kBT - 1) / mce::kBT, mce::kBT, 0, m->stream>>>(
      m->d_head, m->n_agents, m->d_lane, m->d_arc, m->d_order, m->d_start, d_sc, d_sl, d_qx, d_qy, n_queries, d_idx, d_dis);
  return Q.finish();
}

static int frenet_or_center(mce_map* m, int ref_lane, bool center_line, const double* x, const double* y, int n, double* s_out,
                            double* d_out) {
  if (!m || !x || !y || !s_out || !d_out) return efail(MCE_ERR_ARG, "null argument");
  if (ref_lane < 0 || ref_lane >= m->n_lanes) return efail(MCE_ERR_ARG, "reference corridor index out of range");
  if (n < 1) return efail(MCE_ERR_ARG, "no points");
  ENV_TRY(cudaSetDevice(m->device));
  int rc = ensure_query_buffers(m, n);
  if (rc) return rc;
  const size_t q = n;
  Query Q(m);
  double* dx = Q.in(x, q); double* dy = Q.in(y, q);
  double* ds = Q.out(s_out, q); double* dd = Q.out(d_out, q);
  if ((rc = Q.upload())) return rc;
  mce::env_frenet_kernel<<<(n + mce::kBT - 1) / mce::kBT, mce::kBT, 0, m->stream>>>(m->d_tail, ref_lane, center_line, dx, dy, n, ds, dd);
  return Q.finish();
}

int mce_frenet(mce_map* m, int ref_lane, const double* x, const double* y, int n, double* s_out, double* d_out) {
  return frenet_or_center(m, ref_lane, false, x, y, n, s_out, d_out);
}

int mce_project_center(mce_map* m, int lane, const double* x, const double* y, int n, double* s_out, double* dist_out) {
  return frenet_or_center(m, lane, true, x, y, n, s_out, dist_out);
}

int mce_step_along_path(mce_map* m, const int32_t* ref_lane, const double* x, const double* y, const double* v, const double* ax,
                        const double* vy, double dt, int n, double* x_out, double* y_out, double* v_out, double* yaw_out) {
  if (!m || !ref_lane || !x || !y || !v || !ax || !vy || !x_out || !y_out || !v_out || !yaw_out) return efail(MCE_ERR_ARG, "null argument");
  if (n < 1) return efail(MCE_ERR_ARG, "no vehicles");
  for (int i = 0; i < n; ++i)
    if (ref_lane[i] < 0 || ref_lane[i] >= m->n_lanes) return efail(MCE_ERR_ARG, "reference corridor index out of range");
  ENV_TRY(cudaSetDevice(m->device));
  int rc = ensure_query_buffers(m, n);
  if (rc) return rc;
  const size_t q = n;
  Query Q(m);
  double* dx = Q.in(x, q); double* dy = Q.in(y, q); double* dv = Q.in(v, q); double* dax = Q.in(ax, q); double* dvy = Q.in(vy, q);
  int32_t* dl = Q.in(ref_lane, q);
  double* ox = Q.out(x_out, q); double* oy = Q.out(y_out, q); double* ov = Q.out(v_out, q); double* oyaw = Q.out(yaw_out, q);
  if ((rc = Q.upload())) return rc;
  mce::env_step_kernel<<<(n + mce::kBT - 1) / mce::kBT, mce::kBT, 0, m->stream>>>(m->d_tail, dl, dx, dy, dv, dax, dvy, dt, n, ox, oy, ov, oyaw);
  return Q.finish();
}

int mce_center_distance(mce_map* m, const int32_t* lane, const double* x, const double* y, int n, double* signed_out, double* to_end_out) {
  if (!m || !lane || !x || !y || !signed_out || !to_end_out) return efail(MCE_ERR_ARG, "null argument");
  if (n < 1) return efail(MCE_ERR_ARG, "no points");
  for (int i = 0; i < n; ++i)
    if (lane[i] < 0 || lane[i] >= m->n_lanes) return efail(MCE_ERR_ARG, "corridor index out of range");
  ENV_TRY(cudaSetDevice(m->device));
  int rc = ensure_query_buffers(m, n);
  if (rc) return rc;
  const size_t q = n;
  Query Q(m);
  double* dx = Q.in(x, q); double* dy = Q.in(y, q);
  int32_t* dl = Q.in(lane, q);
  double* od = Q.out(signed_out, q); double* oe = Q.out(to_end_out, q);
  if ((rc = Q.upload())) return rc;
  mce::env_center_kernel<<<(n + mce::kBT - 1) / mce::kBT, mce::kBT, 0, m->stream>>>(m->d_border, dl, dx, dy, n, od, oe);
  return Q.finish();
}

int mce_border_distance(mce_map* m, const int32_t* lane, const int32_t* side, const double* hull, int n, double* out) {
  if (!m || !lane || !side || !hull || !out) return efail(MCE_ERR_ARG, "null argument");
  if (n < 1) return efail(MCE_ERR_ARG, "no hulls");
  for (int i = 0; i < n; ++i) {
    if (lane[i] < 0 || lane[i] >= m->n_lanes) return efail(MCE_ERR_ARG, "corridor index out of range");
    if (side[i] != MCE_SIDE_LEFT && side[i] != MCE_SIDE_RIGHT) return efail(MCE_ERR_ARG, "side must be MCE_SIDE_LEFT or MCE_SIDE_RIGHT");
  }
  ENV_TRY(cudaSetDevice(m->device));
  int rc = ensure_query_buffers(m, n);
  if (rc) return rc;
  const size_t q = n;
  Query Q(m);
  double* dh = nullptr;                                      // [corner coordinate][hull]: thread i reads element i
  for (int k = 0; k < 8; ++k) {
    double* p = Q.in_strided(hull, q, 8, k);
    if (k == 0) dh = p;
  }
  const size_t row = ((q * 8 + 15) & ~size_t(15)) / 8;       // doubles between two corner-coordinate rows
  int32_t* dl = Q.in(lane, q); int32_t* dsd = Q.in(side, q);
  double* o = Q.out(out, q);
  if ((rc = Q.upload())) return rc;
  mce::env_border_kernel<<<(n + mce::kBT - 1) / mce::kBT, mce::kBT, 0, m->stream>>>(m->d_border, dl, dsd, dh, row, n, o);
  return Q.finish();
}

int mce_map_set_lane_info(mce_map* m, const int32_t* ids, const int32_t* types, const double* v_limits, int n_corridors) {
  if (!m || !ids || !types || !v_limits) return efail(MCE_ERR_ARG, "null argument");
  if (n_corridors != m->n_lanes) return efail(MCE_ERR_ARG, "one entry per corridor of the map");
  ENV_TRY(cudaSetDevice(m->device));
  mce::EnvLaneInfo info;
  std::memset(&info, 0, sizeof info);
  for (int c = 0; c < n_corridors; ++c) { info.id[c] = ids[c]; info.type[c] = types[c]; info.v_limit[c] = v_limits[c]; }
  if (!m->d_info) ENV_TRY(cudaMalloc(&m->d_info, sizeof info));
  ENV_TRY(cudaMemcpyAsync(m->d_info, &info, sizeof info, cudaMemcpyHostToDevice, m->stream));
  ENV_TRY(cudaStreamSynchronize(m->stream));
  return MCE_OK;
}

int mce_build_rollout_scene(mce_map* m, const mce_build_request* request, const int32_t* ids, const double* attrs, const double* noise,
                            int n_noise, mcs_corridor* corridors_out, int32_t* n_corridors_out, mcs_agent* agents_out,
                            int32_t* n_agents_out, int32_t* actions_out, int32_t* n_actions_out, int device, mcs_scene** scene_out) {
  if (!m || !request || !ids || !attrs) return efail(MCE_ERR_ARG, "null argument");
  if (!noise) n_noise = 0;
  if (scene_out) *scene_out = nullptr;
  if (m->n_scenes == 0) return efail(MCE_ERR_ARG, "mce_build_rollout_scene before mce_env_match");
  if (!m->d_info) return efail(MCE_ERR_ARG, "mce_build_rollout_scene before mce_map_set_lane_info");
  if (request->kind < 0 || request->kind > 2) return efail(MCE_ERR_ARG, "kind must be 0 (merging), 1 (exit) or 2 (lane change)");
  if (request->scene < 0 || request->scene >= m->n_scenes || request->ego_slot < 0 || request->ego_slot >= m->n_agents)
    return efail(MCE_ERR_ARG, "request outside the matched batch");
  if (n_noise < 0) return efail(MCE_ERR_ARG, "negative noise count");
  ENV_TRY(cudaSetDevice(m->device));
  const size_t n = (size_t)m->n_agents;
  const size_t bytes = sizeof(mce_build_request) + n * 4 + n * 64 + (size_t)n_noise * 8 + sizeof(mce::BuildOut) + 256;
  int rc = ensure_query_buffers(m, bytes / 144 + 1);
  if (rc) return rc;
  Query Q(m);
  mce_build_request* d_req = Q.in(request, 1);
  int32_t* d_ids = Q.in(ids, n);
  double* d_attrs = Q.in(attrs, 8 * n);
  double* d_noise = noise ? Q.in(noise, (size_t)n_noise) : nullptr;
  static thread_local mce::BuildOut host_out;
  mce::BuildOut* d_out = Q.out(&host_out, 1);
  if ((rc = Q.upload())) return rc;
  const size_t dyn = n * 36 + 64;
  mce::env_build_kernel<<<1, 128, dyn, m->stream>>>(m->d_head, m->d_tail, m->d_border, m->d_info, m->n_agents, m->d_lane, m->d_arc,
                                                 m->d_order, m->d_start, d_req, d_ids, d_attrs, d_noise, n_noise, d_out);
  if ((rc = Q.finish())) return rc;
  const mce::BuildOut& o = host_out;
  if (o.status == 1) return efail(MCE_ERR_CAPACITY, "the builder collected more vehicles than MCE_BUILD_MAX_AGENTS or than noise draws were passed");
  if (o.status == 2) return efail(MCE_ERR_ARG, "no corridor on the side the builder needs (the reference raises here)");
  if (o.status == 3) return efail(MCE_ERR_ARG, "the ego is on no corridor");
  if (n_corridors_out) *n_corridors_out = o.n_corridors;
  if (n_agents_out) *n_agents_out = o.n_agents;
  if (n_actions_out) *n_actions_out = o.n_actions;
  if (corridors_out) std::memcpy(corridors_out, o.corridors, sizeof(mcs_corridor) * o.n_corridors);
  if (agents_out) std::memcpy(agents_out, o.agents, sizeof(mcs_agent) * o.n_agents);
  if (actions_out) std::memcpy(actions_out, o.actions, sizeof(int32_t) * 4);
  if (scene_out) {
    // MapMultiLane + Environment ctor of the inner simulator (host bookkeeping of <= 48 vehicles: visit order, first lane
    // match, priors — scene.cpp); the scene's rows go to the device at its first launch
    const int src = mcs_scene_create(o.corridors, o.n_corridors, o.agents, o.n_agents, device, scene_out);
    if (src != MCS_OK) return efail(src == MCS_ERR_CUDA ? MCE_ERR_CUDA : (src == MCS_ERR_CAPACITY ? MCE_ERR_CAPACITY : MCE_ERR_ARG),
                                    std::string("mcs_scene_create: ") + mcs_last_error());
  }
  return MCE_OK;
}

static_assert(sizeof(mce_walk_vehicle) == sizeof(mcw::WalkVehicle) && sizeof(mce_walk_result) == sizeof(mcw::WalkResult),
              "include/mcsim_env.h and csrc/walk_kernel.cuh describe the same records");

int mce_hash_iteration_order(const int32_t* ids, int n, int32_t* order_out) {
  if (!ids || !order_out || n < 0 || n > 64) return efail(MCE_ERR_ARG, "bad arguments");
  int order[64];
  mcw::hash_iteration_order(ids, n, order);
  for (int k = 0; k < n; ++k) order_out[k] = order[k];
  return MCE_OK;
}

int mce_walk_step(mce_map* m, mce_walk_vehicle* vehicles, int upload, const int32_t* ids, const double* noise, int n_noise,
                  const double* random_first, int first, int last, double dt, uint64_t seed_base, mce_walk_result* result) {
  if (!m || !vehicles || !ids || !random_first || !result) return efail(MCE_ERR_ARG, "null argument");
  if (m->n_scenes == 0) return efail(MCE_ERR_ARG, "mce_walk_step before mce_env_match");
  if (!m->d_info) return efail(MCE_ERR_ARG, "mce_walk_step before mce_map_set_lane_info");
  const int n = m->n_agents;
  if (first < 0 || last > n || first > last || n_noise < 0 || (n_noise > 0 && !noise)) return efail(MCE_ERR_ARG, "bad range");
  ENV_TRY(cudaSetDevice(m->device));
  // layout of the walk buffer (device and pinned mirror alike): vehicles | attrs [8][n] | ids | random table | result | noise
  const size_t b_veh = sizeof(mcw::WalkVehicle) * (size_t)n, b_attr = 64 * (size_t)n, b_ids = ((size_t)n * 4 + 15) & ~size_t(15);
  const size_t o_attr = b_veh, o_ids = o_attr + b_attr, o_rnd = o_ids + b_ids, o_res = o_rnd + 800, o_noise = o_res + 64;
  const size_t total = o_noise + (size_t)(n_noise > 4096 ? n_noise : 4096) * 8;
  if (total > m->walk_cap) {
    if (m->d_walk) cudaFree(m->d_walk);
    if (m->h_walk) cudaFreeHost(m->h_walk);
    m->d_walk = nullptr; m->h_walk = nullptr; m->walk_cap = 0;
    ENV_TRY(cudaMalloc(&m->d_walk, total));
    ENV_TRY(cudaMallocHost(&m->h_walk, total));
    m->walk_cap = total;
    upload = 1;
  }
  if (m->walk_n != n) upload = 1;
  m->walk_n = n;
  cudaStream_t s = m->stream;
  if (upload) {
    std::memcpy(m->h_walk, vehicles, b_veh);
    double* at = reinterpret_cast<double*>(m->h_walk + o_attr);
    for (int i = 0; i < n; ++i) {
      const mce_walk_vehicle& v = vehicles[i];
      at[i] = v.x; at[(size_t)n + i] = v.y; at[2 * (size_t)n + i] = v.v; at[3 * (size_t)n + i] = v.vy; at[4 * (size_t)n + i] = v.a;
      at[5 * (size_t)n + i] = v.l; at[6 * (size_t)n + i] = v.w; at[7 * (size_t)n + i] = v.idm[5];
    }
    std::memcpy(m->h_walk + o_ids, ids, (size_t)n * 4);
    std::memcpy(m->h_walk + o_rnd, random_first, 800);
    ENV_TRY(cudaMemcpyAsync(m->d_walk, m->h_walk, o_res, cudaMemcpyHostToDevice, s));
  }
  if (n_noise > 0) {
    std::memcpy(m->h_walk + o_noise, noise, (size_t)n_noise * 8);
    ENV_TRY(cudaMemcpyAsync(m->d_walk + o_noise, m->h_walk + o_noise, (size_t)n_noise * 8, cudaMemcpyHostToDevice, s));
  }
  mcw::WalkParams wp;
  wp.first = first; wp.last = last; wp.scene = 0; wp.dt = dt; wp.seed_base = seed_base; wp.n_noise = n_noise;
  constexpr int decide_bytes = 11840;      // rollout_kernel's shared-memory region of a 64-slot rollout without histories (capi.cu)
  const size_t smem = ((sizeof(mcw::WalkShared) + 15) & ~size_t(15)) + decide_bytes + (size_t)n * 224 + 128;
  if (smem > 200 * 1024) return efail(MCE_ERR_CAPACITY, "too many vehicles for the device walk's shared-memory tables");
  ENV_TRY(cudaFuncSetAttribute(mcw::walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mcw::walk_kernel<<<1, 128, smem, s>>>(m->d_head, m->d_tail, m->d_border, m->d_info, n, m->d_lane, m->d_arc, m->d_order, m->d_start,
                                        reinterpret_cast<mcw::WalkVehicle*>(m->d_walk), reinterpret_cast<const int32_t*>(m->d_walk + o_ids),
                                        reinterpret_cast<double*>(m->d_walk + o_attr), reinterpret_cast<const double*>(m->d_walk + o_noise),
                                        reinterpret_cast<const double*>(m->d_walk + o_rnd), wp,
                                        reinterpret_cast<mcw::WalkResult*>(m->d_walk + o_res), decide_bytes);
  ENV_TRY(cudaGetLastError());
  ENV_TRY(cudaMemcpyAsync(m->h_walk, m->d_walk, b_veh, cudaMemcpyDeviceToHost, s));
  ENV_TRY(cudaMemcpyAsync(m->h_walk + o_res, m->d_walk + o_res, sizeof(mcw::WalkResult), cudaMemcpyDeviceToHost, s));
  ENV_TRY(cudaStreamSynchronize(s));
  std::memcpy(vehicles, m->h_walk, b_veh);
  std::memcpy(result, m->h_walk + o_res, sizeof(mce_walk_result));
  return MCE_OK;
}

int mce_last_kernel_ms(mce_map* m, float* out2) {
  if (!m || !out2) return efail(MCE_ERR_ARG, "null argument");
  out2[0] = m->ms[0]; out2[1] = m->ms[1];
  return MCE_OK;
}

}  // extern "C"
