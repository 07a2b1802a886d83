// Outer-environment geometry kernels (include/mcsim_env.h): lane match + per-lane ordering, leader / follower / side
// neighbour search, Frenet conversion, step_along_path.  sm_100a, compiled with -fmad=false: every expression is the
// IEEE-double expression the tests' CPU checker evaluates, operation for operation.
//
// Layout: one thread per vehicle slot, one block per scene for the match (the per-lane vectors are ranked inside the
// block from shared memory), one thread per query for the searches.  The map — polygon shells, centre-line polylines
// (+100 m extension point), Frenet reference lines — is a ~19 KB POD staged into shared memory by every block.
// Vehicle arrays are structure-of-arrays [scene][slot] doubles / int32: thread k reads element k (coalesced).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mcsim_env.h"
#include "../../include/mcsim.h"

namespace mce {

constexpr int kBT = 256;
constexpr int kMaxShell = 2 * MCE_MAX_POINTS;
constexpr int kMaxRef = MCE_MAX_POINTS + 2;

struct EnvMapHead {                          // staged by the match and neighbour kernels
  int n_lanes;
  int left[MCE_MAX_LANES], right[MCE_MAX_LANES];
  int n_center[MCE_MAX_LANES];               // points of centerLine; centerExtended has one more (type.py:154)
  int n_shell[MCE_MAX_LANES];
  int pad[3];
  double center[MCE_MAX_LANES][MCE_MAX_POINTS + 1][2];
  double shell[MCE_MAX_LANES][kMaxShell][2]; // left + reversed right (type.py:160-164)
};

struct EnvMapTail {                          // staged by the Frenet and step kernels
  int n_ref[MCE_MAX_LANES];                  // extend_line_both_sides(center, 1000): n_center + 2 points
  int n_refs[MCE_MAX_LANES];                 // ... simplified (path_line)
  int n_center[MCE_MAX_LANES];
  int n_cs[MCE_MAX_LANES];                   // centerSimplified.path_line
  double ref[MCE_MAX_LANES][kMaxRef][2];
  double ref_dis[MCE_MAX_LANES][kMaxRef];
  double refs[MCE_MAX_LANES][kMaxRef][2];
  double refs_len[MCE_MAX_LANES];
  double cpath[MCE_MAX_LANES][MCE_MAX_POINTS][2];
  double cdis[MCE_MAX_LANES][MCE_MAX_POINTS];
  double cs[MCE_MAX_LANES][MCE_MAX_POINTS][2];
  double cs_len[MCE_MAX_LANES];
  double cyaw[MCE_MAX_LANES][MCE_MAX_POINTS], ccos[MCE_MAX_LANES][MCE_MAX_POINTS], csin[MCE_MAX_LANES][MCE_MAX_POINTS];
};

struct EnvMapBorder {                        // staged by the centre / border distance kernels
  int n_path[2][MCE_MAX_LANES];              // [0] left, [1] right border polyline (type.py:144-151)
  int n_simp[2][MCE_MAX_LANES];              // leftSimplified / rightSimplified .path_line
  int n_center[MCE_MAX_LANES], n_cs[MCE_MAX_LANES];
  double path[2][MCE_MAX_LANES][MCE_MAX_POINTS][2];
  double dis[2][MCE_MAX_LANES][MCE_MAX_POINTS];
  double simp[2][MCE_MAX_LANES][MCE_MAX_POINTS][2];
  double simp_len[2][MCE_MAX_LANES];
  double cpath[MCE_MAX_LANES][MCE_MAX_POINTS][2];
  double cdis[MCE_MAX_LANES][MCE_MAX_POINTS];
  double cs[MCE_MAX_LANES][MCE_MAX_POINTS][2];
  double cs_len[MCE_MAX_LANES];
};

template <typename T>
__device__ __forceinline__ void stage(T* dst, const T* src) {
  static_assert(sizeof(T) % 4 == 0, "word copy");
  const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
  uint32_t* d = reinterpret_cast<uint32_t*>(dst);
  for (int i = threadIdx.x; i < (int)(sizeof(T) / 4); i += blockDim.x) d[i] = s[i];
}

// shapely: nearest point of segment AB to P
__device__ __forceinline__ void seg_point(double ax, double ay, double bx, double by, double px, double py, double& dist,
                                          double& t, double& d2) {
  const double dx = bx - ax, dy = by - ay;
  d2 = dx * dx + dy * dy;
  if (d2 == 0.0) {
    t = 0.0;
  } else {
    t = ((px - ax) * dx + (py - ay) * dy) / d2;
    t = t < 1.0 ? t : 1.0;
    t = t > 0.0 ? t : 0.0;
  }
  const double qx = ax + t * dx, qy = ay + t * dy;
  const double ex = px - qx, ey = py - qy;
  dist = sqrt(ex * ex + ey * ey);
}

// LineString.project + LineString.distance (first segment wins ties)
__device__ __forceinline__ void project_distance(const double (*line)[2], int n, double px, double py, double& s, double& dist) {
  double run = 0.0;
  bool have = false;
  s = 0.0; dist = 0.0;
  for (int k = 0; k + 1 < n; ++k) {
    double d, t, d2;
    seg_point(line[k][0], line[k][1], line[k + 1][0], line[k + 1][1], px, py, d, t, d2);
    const double seg = sqrt(d2);
    if (!have || d < dist) { dist = d; s = run + t * seg; have = true; }
    run = run + seg;
  }
}

// Point.within(Polygon): even-odd crossing, boundary points outside
__device__ __forceinline__ bool within(const double (*shell)[2], int n, double px, double py) {
  bool inside = false;
  for (int k = 0; k < n; ++k) {
    const int k2 = (k + 1 == n) ? 0 : k + 1;
    const double x1 = shell[k][0], y1 = shell[k][1], x2 = shell[k2][0], y2 = shell[k2][1];
    double d, t, d2;
    seg_point(x1, y1, x2, y2, px, py, d, t, d2);
    if (d == 0.0) return false;
    if ((y1 > py) != (y2 > py)) {
      if (px < (x2 - x1) * (py - y1) / (y2 - y1) + x1) inside = !inside;
    }
  }
  return inside;
}

// type.py:255-317 LineSimplified over (path, dis_list) with the simplified line's length
struct LineView {
  const double (*path)[2];
  const double* dis;
  int n;
  double path_length;
  __device__ __forceinline__ int search(double length) const {     // np.searchsorted(dis_list, length)
    int i = 0;
    while (i < n && dis[i] < length) ++i;
    return i;
  }
  __device__ __forceinline__ int yaw_index(double length) const {  // get_yaw_at_length :276-280
    const int i = search(length);
    return i < n ? i : n - 1;
  }
  __device__ __forceinline__ void point_at(double length, double& x, double& y) const {   // :282-295
    if (length == 0.0) { x = path[0][0]; y = path[0][1]; return; }
    if (length < path_length) {
      const int i = search(length);
      const double seg = dis[i] - dis[i - 1];
      const double r = (dis[i] - length) / seg;
      x = path[i][0] - r * (path[i][0] - path[i - 1][0]);
      y = path[i][1] - r * (path[i][1] - path[i - 1][1]);
      return;
    }
    const double end = dis[n - 1] - dis[n - 2];
    const double r = (length - path_length) / end;
    x = path[n - 1][0] + r * (path[n - 1][0] - path[n - 2][0]);
    y = path[n - 1][1] + r * (path[n - 1][1] - path[n - 2][1]);
  }
  // +1 left / -1 right: sign of cross(segment direction, p - foot) == sign of sin(arctan2(p - foot) - yaw), :297-305
  __device__ __forceinline__ int side_of(double s, double px, double py) const {
    double fx, fy;
    point_at(s, fx, fy);
    const int k = yaw_index(s);
    const int a = k == 0 ? 0 : k - 1, b = k == 0 ? 1 : k;
    const double dx = path[b][0] - path[a][0], dy = path[b][1] - path[a][1];
    return dx * (py - fy) - dy * (px - fx) > 0.0 ? 1 : -1;
  }
};

// ---- match_agents_on_corridor + agents_in_corridor_by_arc_length (environment.py:85-101, 135-143) --------------------
// grid = scenes; dynamic shared memory: n doubles (arc) + n ints (lane)
static __global__ void __launch_bounds__(kBT) env_match_kernel(const EnvMapHead* __restrict__ gmap, const double* __restrict__ x,
                                                        const double* __restrict__ y, int n, int32_t* __restrict__ lane_out,
                                                        double* __restrict__ arc_out, int32_t* __restrict__ order_out,
                                                        int32_t* __restrict__ start_out, double* __restrict__ arc_on_lane_out) {
  __shared__ EnvMapHead m;
  __shared__ int count[MCE_MAX_LANES];
  __shared__ int start[MCE_MAX_LANES + 1];
  extern __shared__ double dyn[];
  double* s_arc = dyn;
  int* s_lane = reinterpret_cast<int*>(dyn + n);
  stage(&m, gmap);
  if (threadIdx.x < MCE_MAX_LANES) count[threadIdx.x] = 0;
  __syncthreads();
  const size_t base = (size_t)blockIdx.x * n;
  for (int i = threadIdx.x; i < n; i += kBT) {
    const double px = x[base + i], py = y[base + i];
    int lane = -1;
    double arc = 0.0;
    if (px == px) {
      for (int c = 0; c < m.n_lanes; ++c) {
        double s, d;
        if (arc_on_lane_out) {
          project_distance(m.center[c], m.n_center[c], px, py, s, d);
          arc_on_lane_out[((size_t)blockIdx.x * m.n_lanes + c) * n + i] = s;
        }
        if (lane < 0 && within(m.shell[c], m.n_shell[c], px, py)) {       // corridor_match: first corridor in dict order
          lane = c;
          if (!arc_on_lane_out) project_distance(m.center[c], m.n_center[c], px, py, s, d);
          arc = s;
          if (!arc_on_lane_out) break;
        }
      }
    } else if (arc_on_lane_out) {
      for (int c = 0; c < m.n_lanes; ++c) arc_on_lane_out[((size_t)blockIdx.x * m.n_lanes + c) * n + i] = 0.0;
    }
    s_lane[i] = lane;
    s_arc[i] = arc;
    if (lane >= 0) atomicAdd(&count[lane], 1);
    if (lane_out) lane_out[base + i] = lane;
    if (arc_out) arc_out[base + i] = arc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int c = 0; c <= MCE_MAX_LANES; ++c) {
      start[c] = run;
      if (c < m.n_lanes) run += count[c];
    }
  }
  __syncthreads();
  // stable rank inside the lane: smaller arc first, equal arcs in slot (dict) order
  for (int i = threadIdx.x; i < n; i += kBT) {
    const int lane = s_lane[i];
    if (lane < 0) continue;
    const double a = s_arc[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const double b = s_arc[j];
      rank += (s_lane[j] == lane) & ((b < a) | ((b == a) & (j < i)));
    }
    order_out[base + start[lane] + rank] = i;
  }
  for (int p = start[MCE_MAX_LANES] + threadIdx.x; p < n; p += kBT) order_out[base + p] = -1;
  if (threadIdx.x <= MCE_MAX_LANES) start_out[(size_t)blockIdx.x * (MCE_MAX_LANES + 1) + threadIdx.x] = start[threadIdx.x];
}

// front_agent / following_agent / neighbor_around_agent of vehicle (scene, slot) at position (px, py) against the lanes and
// keys of the last match (environment.py:166-251); idx / dis in the order of the reference's Neighbor constructor
__device__ __forceinline__ void neighbor_row(const EnvMapHead& m, int n, const int32_t* __restrict__ lane, const double* __restrict__ arc,
                                             const int32_t* __restrict__ order, const int32_t* __restrict__ start, int scene, int slot,
                                             double px, double py, int* idx, double* dis) {
  const size_t base = (size_t)scene * n;
  const int32_t* ord = order + base;
  const double* key = arc + base;
  const int32_t* st = start + (size_t)scene * (MCE_MAX_LANES + 1);
#pragma unroll
  for (int k = 0; k < MCE_NEIGHBORS; ++k) { idx[k] = -1; dis[k] = 0.0; }
  const int c = lane[base + slot];
  if (c >= 0) {
    auto lower = [&](int beg, int end, double v) {   // first position with key >= v
      while (beg < end) { const int mid = (beg + end) >> 1; if (key[ord[mid]] < v) beg = mid + 1; else end = mid; }
      return beg;
    };
    auto upper = [&](int beg, int end, double v) {   // first position with key > v
      while (beg < end) { const int mid = (beg + end) >> 1; if (key[ord[mid]] <= v) beg = mid + 1; else end = mid; }
      return beg;
    };
    double e, d;
    project_distance(m.center[c], m.n_center[c], px, py, e, d);
    {
      const int beg = st[c], end = st[c + 1];
      const int f = upper(beg, end, e);
      if (f < end) { idx[MCE_FRONT] = ord[f]; dis[MCE_FRONT] = key[ord[f]] - e; }
      const int b = lower(beg, end, e) - 1;
      if (b >= beg) { idx[MCE_BACK] = ord[b]; dis[MCE_BACK] = key[ord[b]] - e; }
    }
#pragma unroll
    for (int sd = 0; sd < 2; ++sd) {
      const int side = sd == 0 ? m.left[c] : m.right[c];
      if (side < 0) continue;
      const int o = sd == 0 ? MCE_LEFT_FF : MCE_RIGHT_FF;     // ff, f, b, bb
      const int beg = st[side], end = st[side + 1];
      double es;
      project_distance(m.center[side], m.n_center[side] + 1, px, py, es, d);   // centerExtended
      const int f = lower(beg, end, es);
      if (f < end) {
        const double df = key[ord[f]] - es;
        idx[o + 1] = ord[f]; dis[o + 1] = df;
        const double thr = df + es;
        int p = f + 1;
        while (p < end && !(key[ord[p]] > thr)) ++p;
        if (p < end) { idx[o] = ord[p]; dis[o] = key[ord[p]] - es; }
      }
      const int b = f - 1;                                    // last position with key < es
      if (b >= beg) {
        const double db = key[ord[b]] - es;
        idx[o + 2] = ord[b]; dis[o + 2] = db;
        const double thr = db + es;
        int p = b - 1;
        while (p >= beg && !(key[ord[p]] < thr)) --p;
        if (p >= beg) { idx[o + 3] = ord[p]; dis[o + 3] = key[ord[p]] - es; }
      }
    }
  }
}

// ---- front_agent / following_agent / neighbor_around_agent (environment.py:166-251) ---------------------------------
// One thread per query.  q_scene == nullptr: every slot of every scene at the matched positions (grid.y = scene);
// otherwise a list of (scene, slot, x, y) queries against the resident vectors of the last match.
static __global__ void __launch_bounds__(kBT) env_neighbors_kernel(const EnvMapHead* __restrict__ gmap, int n,
                                                            const int32_t* __restrict__ lane, const double* __restrict__ arc,
                                                            const int32_t* __restrict__ order, const int32_t* __restrict__ start,
                                                            const int32_t* __restrict__ q_scene, const int32_t* __restrict__ q_slot,
                                                            const double* __restrict__ qx, const double* __restrict__ qy,
                                                            int n_queries, int32_t* __restrict__ nb_index, double* __restrict__ nb_dis) {
  __shared__ EnvMapHead m;
  stage(&m, gmap);
  __syncthreads();
  const int t = blockIdx.x * kBT + threadIdx.x;
  int scene, slot;
  size_t out, stride;
  double px, py;
  if (q_scene == nullptr) {
    if (t >= n) return;
    scene = blockIdx.y; slot = t;
    px = qx[(size_t)scene * n + slot]; py = qy[(size_t)scene * n + slot];
    out = (size_t)scene * MCE_NEIGHBORS * n + slot; stride = n;
  } else {
    if (t >= n_queries) return;
    scene = q_scene[t]; slot = q_slot[t];
    px = qx[t]; py = qy[t];
    out = t; stride = n_queries;
  }
  int idx[MCE_NEIGHBORS];
  double dis[MCE_NEIGHBORS];
  neighbor_row(m, n, lane, arc, order, start, scene, slot, px, py, idx, dis);
#pragma unroll
  for (int k = 0; k < MCE_NEIGHBORS; ++k) {
    nb_index[out + k * stride] = idx[k];
    nb_dis[out + k * stride] = dis[k];
  }
}

// ---- LineSimplified(extend_line_both_sides(center, 1000)).to_frenet_frame (type.py:307-317) --------------------------
// center_line: project on the corridor's own centre line instead (LineString(center).project / .distance)
static __global__ void __launch_bounds__(kBT) env_frenet_kernel(const EnvMapTail* __restrict__ gmap, int ref_lane, bool center_line,
                                                         const double* __restrict__ x, const double* __restrict__ y, int n,
                                                         double* __restrict__ s_out, double* __restrict__ d_out) {
  __shared__ EnvMapTail m;
  stage(&m, gmap);
  __syncthreads();
  const int i = blockIdx.x * kBT + threadIdx.x;
  if (i >= n) return;
  const double px = x[i], py = y[i];
  double s, dis;
  if (center_line) {
    project_distance(m.cpath[ref_lane], m.n_center[ref_lane], px, py, s, dis);
    s_out[i] = s;
    d_out[i] = dis;
    return;
  }
  project_distance(m.refs[ref_lane], m.n_refs[ref_lane], px, py, s, dis);
  double d = 0.0;
  if (dis > 0.0) {
    LineView lv{m.ref[ref_lane], m.ref_dis[ref_lane], m.n_ref[ref_lane], m.refs_len[ref_lane]};
    d = lv.side_of(s, px, py) > 0 ? dis : -dis;
  }
  s_out[i] = s;
  d_out[i] = d;
}

// utility.py:88-98 for one vehicle on corridors[l].centerSimplified; k = index of the yaw / cos / sin entry of the new position
__device__ __forceinline__ void step_along_path_fn(const EnvMapTail& m, int l, double px, double py, double v, double ax, double vy,
                                                   double dt, double& x_new, double& y_new, double& v_out, int& k) {
  LineView lv{m.cpath[l], m.cdis[l], m.n_center[l], m.cs_len[l]};
  double s, dis;
  project_distance(m.cs[l], m.n_cs[l], px, py, s, dis);
  const int sign = (dis > 0.0 && lv.side_of(s, px, py) > 0) ? 1 : -1;   // is_left_of: a point on the line is not left
  const double d = sign > 0 ? dis : -dis;
  const double s_new = s + v * dt;
  const double d_new = d + vy * dt;
  double v_new = v + ax * dt;
  v_new = v_new > 0.0 ? v_new : 0.0;
  k = lv.yaw_index(s_new);
  double fx, fy;
  lv.point_at(s_new, fx, fy);
  x_new = fx - d_new * m.csin[l][k];            // offset_point utility.py:83-85
  y_new = fy + d_new * m.ccos[l][k];
  v_out = v_new;
}

// ---- utility.py:88-98 step_along_path on corridors[ref_lane[i]].centerSimplified ------------------------------------
static __global__ void __launch_bounds__(kBT) env_step_kernel(const EnvMapTail* __restrict__ gmap, const int32_t* __restrict__ ref_lane,
                                                       const double* __restrict__ x, const double* __restrict__ y,
                                                       const double* __restrict__ v, const double* __restrict__ ax,
                                                       const double* __restrict__ vy, double dt, int n, double* __restrict__ x_out,
                                                       double* __restrict__ y_out, double* __restrict__ v_out,
                                                       double* __restrict__ yaw_out) {
  __shared__ EnvMapTail m;
  stage(&m, gmap);
  __syncthreads();
  const int i = blockIdx.x * kBT + threadIdx.x;
  if (i >= n) return;
  int k;
  step_along_path_fn(m, ref_lane[i], x[i], y[i], v[i], ax[i], vy[i], dt, x_out[i], y_out[i], v_out[i], k);
  yaw_out[i] = m.cyaw[ref_lane[i]][k];
}

__device__ __forceinline__ void center_distance_fn(const EnvMapBorder& m, int l, double px, double py, double& signed_d, double& to_end) {
  double s, dis;
  project_distance(m.cs[l], m.n_cs[l], px, py, s, dis);
  double d = 0.0;
  if (dis > 0.0) {
    LineView lv{m.cpath[l], m.cdis[l], m.n_center[l], m.cs_len[l]};
    d = lv.side_of(s, px, py) > 0 ? dis : -dis;
  }
  signed_d = d;
  to_end = m.cs_len[l] - s;
}

// ---- corridor.centerSimplified.signed_distance (type.py:307-315, behaviors.py:20) and Corridor.distance_to_corridor_end
// (type.py:180-181) of n points, each on its own corridor ---------------------------------------------------------------
static __global__ void __launch_bounds__(kBT) env_center_kernel(const EnvMapBorder* __restrict__ gmap, const int32_t* __restrict__ lane,
                                                         const double* __restrict__ x, const double* __restrict__ y, int n,
                                                         double* __restrict__ d_out, double* __restrict__ end_out) {
  __shared__ EnvMapBorder m;
  stage(&m, gmap);
  __syncthreads();
  const int i = blockIdx.x * kBT + threadIdx.x;
  if (i >= n) return;
  center_distance_fn(m, lane[i], x[i], y[i], d_out[i], end_out[i]);
}

// shapely: do the closed segments AB and CD share a point
__device__ __forceinline__ double orient(double ax, double ay, double bx, double by, double cx, double cy) {
  return (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
}
__device__ __forceinline__ bool on_box(double px, double py, double qx, double qy, double rx, double ry) {
  return fmin(px, qx) <= rx && rx <= fmax(px, qx) && fmin(py, qy) <= ry && ry <= fmax(py, qy);
}
__device__ __forceinline__ bool segments_intersect(double ax, double ay, double bx, double by, double cx, double cy, double dx,
                                                   double dy) {
  const double o1 = orient(ax, ay, bx, by, cx, cy), o2 = orient(ax, ay, bx, by, dx, dy);
  const double o3 = orient(cx, cy, dx, dy, ax, ay), o4 = orient(cx, cy, dx, dy, bx, by);
  if (((o1 > 0) != (o2 > 0)) && ((o3 > 0) != (o4 > 0)) && o1 != 0 && o2 != 0 && o3 != 0 && o4 != 0) return true;
  return (o1 == 0 && on_box(ax, ay, bx, by, cx, cy)) || (o2 == 0 && on_box(ax, ay, bx, by, dx, dy)) ||
         (o3 == 0 && on_box(cx, cy, dx, dy, ax, ay)) || (o4 == 0 && on_box(cx, cy, dx, dy, bx, by));
}
__device__ __forceinline__ double seg_seg_distance(double ax, double ay, double bx, double by, double cx, double cy, double dx,
                                                   double dy) {
  if (segments_intersect(ax, ay, bx, by, cx, cy, dx, dy)) return 0.0;
  double d, t, d2, best;
  seg_point(ax, ay, bx, by, cx, cy, best, t, d2);
  seg_point(ax, ay, bx, by, dx, dy, d, t, d2); best = d < best ? d : best;
  seg_point(cx, cy, dx, dy, ax, ay, d, t, d2); best = d < best ? d : best;
  seg_point(cx, cy, dx, dy, bx, by, d, t, d2); best = d < best ? d : best;
  return best;
}

// Corridor.distance_to_border(hull, direction) for one hull (corner order of vehicle_pose_to_rect); sd: 0 = "left", 1 = "right"
__device__ __forceinline__ double border_distance_fn(const EnvMapBorder& m, int l, int sd, const double* hx, const double* hy) {
  const double (*simp)[2] = m.simp[sd][l];
  const int ns = m.n_simp[sd][l];
  LineView lv{m.path[sd][l], m.dis[sd][l], m.n_path[sd][l], m.simp_len[sd][l]};
  bool pass = false;
  double worst = 0.0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double s, dis;
    project_distance(simp, ns, hx[k], hy[k], s, dis);
    const bool left_of = dis > 0.0 && lv.side_of(s, hx[k], hy[k]) > 0;      // is_left_of type.py:297-305
    if (sd == 0 ? left_of : !left_of) {
      pass = true;
      worst = dis > worst ? dis : worst;
    }
  }
  if (pass) return -worst;
  // Polygon(hull).distance(LineString): 0 when a vertex of the line lies inside the hull, else the nearest edge pair
  double shell[4][2];
#pragma unroll
  for (int k = 0; k < 4; ++k) { shell[k][0] = hx[k]; shell[k][1] = hy[k]; }
  for (int k = 0; k < ns; ++k)
    if (within(shell, 4, simp[k][0], simp[k][1])) return 0.0;
  double best = 0.0;
  bool have = false;
  for (int e = 0; e < 4; ++e) {
    const int e2 = (e + 1) & 3;
    for (int k = 0; k + 1 < ns; ++k) {
      const double d = seg_seg_distance(hx[e], hy[e], hx[e2], hy[e2], simp[k][0], simp[k][1], simp[k + 1][0], simp[k + 1][1]);
      if (!have || d < best) { best = d; have = true; }
    }
  }
  return best;
}



// ---- Corridor.distance_to_border(hull, direction) (type.py:183-206; environment.py:113-115 dis_to_lane_border):
// < 0 = how far the hull has passed the border on that side, otherwise Polygon(hull).distance(border line) ------------
static __global__ void __launch_bounds__(kBT) env_border_kernel(const EnvMapBorder* __restrict__ gmap, const int32_t* __restrict__ lane,
                                                         const int32_t* __restrict__ side, const double* __restrict__ hull, size_t row,
                                                         int n, double* __restrict__ out) {
  __shared__ EnvMapBorder m;
  stage(&m, gmap);
  __syncthreads();
  const int i = blockIdx.x * kBT + threadIdx.x;
  if (i >= n) return;
  double hx[4], hy[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { hx[k] = hull[(size_t)(2 * k) * row + i]; hy[k] = hull[(size_t)(2 * k + 1) * row + i]; }
  out[i] = border_distance_fn(m, lane[i], side[i], hx, hy);
}

// ---- mcs_wrapper.py:24-283 — the three builders of the inner simulator's scene, on the device ----------------------
// One block per request.  Thread 0 walks the reference's selection logic over the resident match tables (the neighbour
// slots FF / F / B / BB per side, front, front of the front, follower, the vehicles within 80 m on the second lanes, the
// candidate gaps) and lists the vehicles and corridors in the reference's append order; then every thread converts one
// listed point — vehicles at their CURRENT positions, six corner points per corridor — to the Frenet frame of the builder's
// reference line (extend_line_both_sides(center, 1000), simplified) and writes its AgentMS / CorridorMS record.
struct EnvLaneInfo { int id[MCE_MAX_LANES], type[MCE_MAX_LANES]; double v_limit[MCE_MAX_LANES]; };   // Corridor.id / .type / .v_limit

constexpr int kBuildMaxAgents = MCE_BUILD_MAX_AGENTS;
struct BuildOut {
  int32_t n_agents, n_corridors, n_actions, status;     // status: 0 ok, 1 too many vehicles, 2 no lane on the needed side, 3 ego unmatched
  int32_t actions[4];
  mcs_corridor corridors[MCE_MAX_LANES];
  mcs_agent agents[kBuildMaxAgents];
};

// LineSimplified(extend_line_both_sides(center, 1000)).to_frenet_frame (type.py:307-317) — the body of env_frenet_kernel
__device__ __forceinline__ void to_frenet(const EnvMapTail& t, int ref_lane, double px, double py, double& s, double& d) {
  double dis;
  project_distance(t.refs[ref_lane], t.n_refs[ref_lane], px, py, s, dis);
  d = 0.0;
  if (dis > 0.0) {
    LineView lv{t.ref[ref_lane], t.ref_dis[ref_lane], t.n_ref[ref_lane], t.refs_len[ref_lane]};
    d = lv.side_of(s, px, py) > 0 ? dis : -dis;
  }
}

// The body of one builder call, executed by every thread of a block (it synchronises the block).  H: the map head in
// shared memory; lane / arc / order / start: ONE scene's match tables, ax / ay: the vehicles' current positions (shared
// memory: the selection walks dependent look-ups — binary searches over the lane vectors, projections on the centre
// lines — which run at shared-memory latency there); attrs [8][n] / ids: global; out: global or shared.
struct BuildScratch {
  int item_slot[kBuildMaxAgents], item_beh[kBuildMaxAgents];
  double item_arc[kBuildMaxAgents];
  int corr_lane[MCE_MAX_LANES];
  int n_items, n_corr, n_act, status, ref_lane;
  int acts[4];
  double ego_s, x_limit;
  double cs[MCE_MAX_LANES][6], cd[MCE_MAX_LANES][6];
};

// What a caller that walks the vehicles one after the other (walk_kernel.cuh) keeps per vehicle, so that the selection does
// not project again what it projected for another vehicle: the neighbour row and the arc length on the own lane at the
// vehicle's CURRENT position; stale bit 1 (row) / 4 (arc) = the vehicle has moved since: compute, store, clear.
struct BuildCache {
  int32_t* nb_idx;      // [n][MCE_NEIGHBORS]
  double* nb_dis;       // [n][MCE_NEIGHBORS]
  double* arc_now;      // [n]
  int32_t* stale;       // [n]
};

__device__ __forceinline__ void build_scene_block(const EnvMapHead& H, const EnvMapTail* __restrict__ tail,
                                                  const EnvMapBorder* __restrict__ border, const EnvLaneInfo* __restrict__ info, int n,
                                                  const int32_t* lane, const double* arc, const int32_t* order, const int32_t* start,
                                                  const double* ax, const double* ay, const mce_build_request& rq,
                                                  const int32_t* __restrict__ ids, const double* __restrict__ attrs,
                                                  const double* __restrict__ noise, int n_noise, BuildScratch& w, BuildOut& out,
                                                  const BuildCache* cache = nullptr) {
  int (&item_slot)[kBuildMaxAgents] = w.item_slot;
  int (&item_beh)[kBuildMaxAgents] = w.item_beh;
  int (&corr_lane)[MCE_MAX_LANES] = w.corr_lane;
  int& n_items = w.n_items; int& n_corr = w.n_corr; int& n_act = w.n_act; int& status = w.status; int& ref_lane = w.ref_lane;
  int (&acts)[4] = w.acts;
  double& ego_s_sh = w.ego_s; double& x_limit_sh = w.x_limit;
  double (&cs_)[MCE_MAX_LANES][6] = w.cs;
  double (&cd_)[MCE_MAX_LANES][6] = w.cd;
  const EnvMapTail& T = *tail;
  const int ego = rq.ego_slot;
  const int ego_lane = lane[ego];
  if (threadIdx.x == 0) {
    n_items = 0; n_corr = 0; n_act = 1; acts[0] = -1; status = 0; ref_lane = ego_lane;
    auto add = [&](int slot, int beh) {
      if (n_items >= kBuildMaxAgents - 1) { status = 1; return; }
      item_slot[n_items] = slot; item_beh[n_items] = beh; n_items++;
    };
    auto corridor = [&](int l) { if (n_corr < MCE_MAX_LANES) corr_lane[n_corr++] = l; };
    auto gap = [&](int slot) { if (slot >= 0 && n_act < 4) acts[n_act++] = ids[slot]; };
    // sorted_agents_on_corridor_within_range_of_vehicle (environment.py:145-156): the vehicles matched on lane l whose arc
    // length on that lane's centre line is within 80 m of the ego's, by arc length (ties in dict order)
    auto row_of = [&](int slot, int* idx, double* dis) {     // neighbor_row at the vehicle's current position
      if (cache && !(cache->stale[slot] & 1)) {
        for (int k = 0; k < MCE_NEIGHBORS; ++k) { idx[k] = cache->nb_idx[slot * MCE_NEIGHBORS + k]; dis[k] = cache->nb_dis[slot * MCE_NEIGHBORS + k]; }
        return;
      }
      neighbor_row(H, n, lane, arc, order, start, 0, slot, ax[slot], ay[slot], idx, dis);
      if (cache) {
        for (int k = 0; k < MCE_NEIGHBORS; ++k) { cache->nb_idx[slot * MCE_NEIGHBORS + k] = idx[k]; cache->nb_dis[slot * MCE_NEIGHBORS + k] = dis[k]; }
        cache->stale[slot] &= ~1;
      }
    };
    auto within_range = [&](int l, int beh) {
      double e, d;
      project_distance(H.center[l], H.n_center[l], ax[ego], ay[ego], e, d);
      const int first = n_items;
      for (int j = 0; j < n; ++j) {
        if (lane[j] != l) continue;
        double a;
        if (cache && !(cache->stale[j] & 4)) a = cache->arc_now[j];
        else {
          project_distance(H.center[l], H.n_center[l], ax[j], ay[j], a, d);
          if (cache) { cache->arc_now[j] = a; cache->stale[j] &= ~4; }
        }
        if (!(fabs(a - e) <= 80.0)) continue;
        if (n_items >= kBuildMaxAgents - 1) { status = 1; return; }
        // insertion by arc length, stable
        int p = n_items;
        while (p > first && w.item_arc[p - 1] > a) {
          item_slot[p] = item_slot[p - 1]; item_beh[p] = item_beh[p - 1]; w.item_arc[p] = w.item_arc[p - 1]; --p;
        }
        item_slot[p] = j; item_beh[p] = beh; w.item_arc[p] = a; n_items++;
      }
    };
    if (ego_lane < 0) status = 3;
    else {
      int idx[MCE_NEIGHBORS];
      double dis[MCE_NEIGHBORS];
      row_of(ego, idx, dis);
      if (rq.kind == 0) {                                   // merging_mcs_sim_from_env, mcs_wrapper.py:24-75
        const int left = H.left[ego_lane];
        if (left < 0) status = 2;
        else {
          ref_lane = left;                                  // reference line: the main lane next to the ramp
          const int beh = rq.others_merging ? MCS_BEH_MERGING : MCS_BEH_IDM_LC;
          const int laf[5] = {MCE_LEFT_FF, MCE_LEFT_F, MCE_LEFT_B, MCE_LEFT_BB, MCE_FRONT};
          for (int q = 0; q < 5; ++q) if (idx[laf[q]] >= 0) add(idx[laf[q]], beh);
          gap(idx[MCE_LEFT_BB]); gap(idx[MCE_LEFT_B]); gap(idx[MCE_LEFT_F]); gap(idx[MCE_LEFT_FF]);
          corridor(ego_lane); corridor(left);
          const int ll = H.left[left];
          if (ll >= 0) { within_range(ll, MCS_BEH_IDM_LC); corridor(ll); }
        }
      } else {                                              // exit_ (:78-190) and lane_change_mcs_sim_from_env (:193-283)
        if (rq.kind == 1) {
          int r = H.right[ego_lane], ex = -1;
          while (r >= 0) { if (info->type[r] == MCS_LANE_EXIT) { ex = r; break; } r = H.right[r]; }
          if (ex < 0) status = 2;
          else corridor(ex);                                // plan index 0: only its far centre point is used (:113-115)
        }
        corridor(ego_lane);
        if (idx[MCE_FRONT] >= 0) {
          add(idx[MCE_FRONT], MCS_BEH_IDM_LC);
          int i2[MCE_NEIGHBORS];
          double d2[MCE_NEIGHBORS];
          const int f = idx[MCE_FRONT];
          row_of(f, i2, d2);
          if (i2[MCE_FRONT] >= 0) add(i2[MCE_FRONT], MCS_BEH_IDM_LC);
        }
        if (idx[MCE_BACK] >= 0) add(idx[MCE_BACK], MCS_BEH_IDM_LC);
        const int left = H.left[ego_lane];
        if (left >= 0) {
          for (int q = MCE_LEFT_FF; q <= MCE_LEFT_BB; ++q) if (idx[q] >= 0) add(idx[q], MCS_BEH_IDM_LC);
          corridor(left);
          const int ll = H.left[left];
          if (ll >= 0) { within_range(ll, MCS_BEH_IDM_LC); corridor(ll); }
        }
        if (rq.kind == 1) { gap(idx[MCE_RIGHT_BB]); gap(idx[MCE_RIGHT_B]); gap(idx[MCE_RIGHT_F]); gap(idx[MCE_RIGHT_FF]); }
        const int right = H.right[ego_lane];
        if (right >= 0) {
          const int beh = info->type[right] == MCS_LANE_MERGING ? MCS_BEH_MERGING : MCS_BEH_IDM_LC;
          for (int q = MCE_RIGHT_FF; q <= MCE_RIGHT_BB; ++q) if (idx[q] >= 0) add(idx[q], beh);
          corridor(right);
          const int rr = H.right[right];
          if (rr >= 0) { within_range(rr, info->type[rr] == MCS_LANE_MERGING ? MCS_BEH_MERGING : MCS_BEH_IDM_LC); corridor(rr); }
        }
      }
      if (noise != nullptr && n_items > n_noise) status = 1;
    }
  }
  __syncthreads();
  if (status != 0) {
    if (threadIdx.x == 0) { out.status = status; out.n_agents = 0; out.n_corridors = 0; out.n_actions = 0; }
    return;
  }
  // ---- the ego and the listed vehicles, one per thread
  for (int q = threadIdx.x; q <= n_items; q += blockDim.x) {
    const int slot = q == 0 ? ego : item_slot[q - 1];
    double s, d;
    to_frenet(T, ref_lane, ax[slot], ay[slot], s, d);
    if (q == 0) ego_s_sh = s;
    mcs_agent a;
    a.id = ids[slot];
    a.x = s; a.y = d;
    a.v = attrs[2 * (size_t)n + slot]; a.vy = attrs[3 * (size_t)n + slot]; a.a = attrs[4 * (size_t)n + slot];
    a.l = attrs[5 * (size_t)n + slot]; a.w = attrs[6 * (size_t)n + slot];
    a.reserved = 0;
    if (q == 0) {
      a.behavior = rq.ego_behavior; a.vd = rq.ego_vd;
      for (int t = 0; t < 6; ++t) { a.idm[t] = rq.ego_idm[t]; a.rss[t] = rq.ego_rss[t]; }
      for (int t = 0; t < 7; ++t) a.yp[t] = rq.ego_yp[t];
      a.commit = rq.ego_commit; a.commit_keep_lane_step = rq.ego_keep_step; a.target_lane_id = rq.ego_target_lane;
    } else {
      const int beh = item_beh[q - 1];
      a.behavior = beh;
      a.vd = attrs[7 * (size_t)n + slot];
      if (noise != nullptr) a.vd = a.vd + noise[q - 1];                        // vd + np.random.normal(0., 1), append order
      const bool truck = a.l > 7.0;                                            // idm_param(l), mcs_wrapper.py:8-13 (memory order)
      a.idm[0] = truck ? 0.8 : 1.6; a.idm[1] = 2.0; a.idm[2] = truck ? -1.5 : -3.0; a.idm[3] = -2.0; a.idm[4] = -8.0; a.idm[5] = 1.2;
      const bool m = beh == MCS_BEH_MERGING;                                   // rss_param(behavior == "merging"), :16-21
      a.rss[0] = 0.7; a.rss[1] = 0.4; a.rss[2] = m ? -8.0 : -6.0; a.rss[3] = m ? -10.0 : -6.0; a.rss[4] = 1.8; a.rss[5] = 1.2;
      a.yp[0] = -0.5; a.yp[1] = 1.81; a.yp[2] = -4.8; a.yp[3] = -1.1; a.yp[4] = 0.9; a.yp[5] = 0.5; a.yp[6] = 0.51;   // YieldingParam()
      a.commit = MCS_COMMIT_NOT_DECIDED; a.commit_keep_lane_step = 0; a.target_lane_id = -1;
    }
    out.agents[q] = a;
  }
  // ---- corridors: Corridor.to_frenet_corridor (type.py:207-237) — the far points take the lateral offset of the near ones
  for (int q = threadIdx.x; q < n_corr * 6; q += blockDim.x) {
    const int c = q / 6, w = q - 6 * c, l = corr_lane[c];
    const EnvMapBorder& B = *border;
    double px, py;
    if (w < 4) {
      const int sd = w >> 1, last = (w & 1) ? B.n_path[sd][l] - 1 : 0;
      px = B.path[sd][l][last][0]; py = B.path[sd][l][last][1];
    } else {
      const int last = (w & 1) ? T.n_center[l] - 1 : 0;
      px = T.cpath[l][last][0]; py = T.cpath[l][last][1];
    }
    to_frenet(T, ref_lane, px, py, cs_[c][w], cd_[c][w]);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // exit builder: the centre corridor ends where the exit has to be reached (mcs_wrapper.py:113-118)
    int first = 0;
    x_limit_sh = 0.0;
    bool limited = false;
    if (rq.kind == 1) {
      const int ex = corr_lane[0], el = corr_lane[1];
      const double vl = info->v_limit[el];
      const double by_speed = vl * vl / 6.25;
      const int lanes_between = abs(info->id[ex] - info->id[el]) - 1;
      const double by_exit = (cs_[0][5] - (double)(lanes_between * 50)) - ego_s_sh;
      const double max_length = by_speed < by_exit ? by_speed : by_exit;       // min(a, b): a unless b < a
      x_limit_sh = ego_s_sh + max_length;
      limited = true;
      first = 1;
    }
    int w = 0;
    for (int c = first; c < n_corr; ++c, ++w) {
      const int l = corr_lane[c];
      mcs_corridor o;
      o.id = info->id[l]; o.type = info->type[l]; o.v_limit = info->v_limit[l];
      double far0 = cs_[c][1], far1 = cs_[c][3], far2 = cs_[c][5];
      if (limited && c == first) {
        far0 = far0 < x_limit_sh ? far0 : x_limit_sh; far1 = far1 < x_limit_sh ? far1 : x_limit_sh; far2 = far2 < x_limit_sh ? far2 : x_limit_sh;
      }
      o.left[0] = cs_[c][0]; o.left[1] = cd_[c][0]; o.left[2] = far0; o.left[3] = cd_[c][0];
      o.right[0] = cs_[c][2]; o.right[1] = cd_[c][2]; o.right[2] = far1; o.right[3] = cd_[c][2];
      o.center[0] = cs_[c][4]; o.center[1] = cd_[c][4]; o.center[2] = far2; o.center[3] = cd_[c][4];
      out.corridors[w] = o;
    }
    out.n_corridors = w; out.n_agents = n_items + 1; out.n_actions = n_act; out.status = 0;
    for (int q = 0; q < 4; ++q) out.actions[q] = q < n_act ? acts[q] : 0;
  }
}


// one block per request: stage the map head, the request's scene tables and the current positions, then build
// Dynamic shared memory: n x (8 arc + 8 x + 8 y + 4 lane + 4 order + 4 spare) bytes + 9 lane starts.
static __global__ void __launch_bounds__(128) env_build_kernel(const EnvMapHead* __restrict__ head, const EnvMapTail* __restrict__ tail,
                                                        const EnvMapBorder* __restrict__ border, const EnvLaneInfo* __restrict__ info,
                                                        int n, const int32_t* __restrict__ g_lane, const double* __restrict__ g_arc,
                                                        const int32_t* __restrict__ g_order, const int32_t* __restrict__ g_start,
                                                        const mce_build_request* __restrict__ reqs, const int32_t* __restrict__ ids,
                                                        const double* __restrict__ attrs /* [8][n]: x y v vy a l w vd, current */,
                                                        const double* __restrict__ noise, int n_noise, BuildOut* __restrict__ outs) {
  __shared__ EnvMapHead H;
  __shared__ BuildScratch w;
  extern __shared__ double build_dyn[];
  const mce_build_request& rq = reqs[blockIdx.x];
  const int scene = rq.scene;
  double* arc = build_dyn; double* sx = arc + n; double* sy = sx + n;
  int32_t* lane = reinterpret_cast<int32_t*>(sy + n); int32_t* order = lane + n; int32_t* start = order + n;
  stage(&H, head);
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    arc[j] = g_arc[(size_t)scene * n + j]; lane[j] = g_lane[(size_t)scene * n + j]; order[j] = g_order[(size_t)scene * n + j];
    sx[j] = attrs[j]; sy[j] = attrs[(size_t)n + j];
  }
  if (threadIdx.x <= MCE_MAX_LANES) start[threadIdx.x] = g_start[(size_t)scene * (MCE_MAX_LANES + 1) + threadIdx.x];
  __syncthreads();
  build_scene_block(H, tail, border, info, n, lane, arc, order, start, sx, sy, rq, ids, attrs, noise, n_noise, w, outs[blockIdx.x]);
}

}  // namespace mce
