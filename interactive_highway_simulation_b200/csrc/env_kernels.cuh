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
__global__ void __launch_bounds__(kBT) env_match_kernel(const EnvMapHead* __restrict__ gmap, const double* __restrict__ x,
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

// ---- front_agent / following_agent / neighbor_around_agent (environment.py:166-251) ---------------------------------
// One thread per query.  q_scene == nullptr: every slot of every scene at the matched positions (grid.y = scene);
// otherwise a list of (scene, slot, x, y) queries against the resident vectors of the last match.
__global__ void __launch_bounds__(kBT) env_neighbors_kernel(const EnvMapHead* __restrict__ gmap, int n,
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
  const size_t base = (size_t)scene * n;
  const int32_t* ord = order + base;
  const double* key = arc + base;
  const int32_t* st = start + (size_t)scene * (MCE_MAX_LANES + 1);
  int idx[MCE_NEIGHBORS];
  double dis[MCE_NEIGHBORS];
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
#pragma unroll
  for (int k = 0; k < MCE_NEIGHBORS; ++k) {
    nb_index[out + k * stride] = idx[k];
    nb_dis[out + k * stride] = dis[k];
  }
}

// ---- LineSimplified(extend_line_both_sides(center, 1000)).to_frenet_frame (type.py:307-317) --------------------------
// center_line: project on the corridor's own centre line instead (LineString(center).project / .distance)
__global__ void __launch_bounds__(kBT) env_frenet_kernel(const EnvMapTail* __restrict__ gmap, int ref_lane, bool center_line,
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

// ---- utility.py:88-98 step_along_path on corridors[ref_lane[i]].centerSimplified ------------------------------------
__global__ void __launch_bounds__(kBT) env_step_kernel(const EnvMapTail* __restrict__ gmap, const int32_t* __restrict__ ref_lane,
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
  const int l = ref_lane[i];
  const double px = x[i], py = y[i];
  LineView lv{m.cpath[l], m.cdis[l], m.n_center[l], m.cs_len[l]};
  double s, dis;
  project_distance(m.cs[l], m.n_cs[l], px, py, s, dis);
  const int sign = (dis > 0.0 && lv.side_of(s, px, py) > 0) ? 1 : -1;   // is_left_of: a point on the line is not left
  const double d = sign > 0 ? dis : -dis;
  const double s_new = s + v[i] * dt;
  const double d_new = d + vy[i] * dt;
  double v_new = v[i] + ax[i] * dt;
  v_new = v_new > 0.0 ? v_new : 0.0;
  const int k = lv.yaw_index(s_new);
  double fx, fy;
  lv.point_at(s_new, fx, fy);
  x_out[i] = fx - d_new * m.csin[l][k];            // offset_point utility.py:83-85
  y_out[i] = fy + d_new * m.ccos[l][k];
  v_out[i] = v_new;
  yaw_out[i] = m.cyaw[l][k];
}

// ---- corridor.centerSimplified.signed_distance (type.py:307-315, behaviors.py:20) and Corridor.distance_to_corridor_end
// (type.py:180-181) of n points, each on its own corridor ---------------------------------------------------------------
__global__ void __launch_bounds__(kBT) env_center_kernel(const EnvMapBorder* __restrict__ gmap, const int32_t* __restrict__ lane,
                                                         const double* __restrict__ x, const double* __restrict__ y, int n,
                                                         double* __restrict__ d_out, double* __restrict__ end_out) {
  __shared__ EnvMapBorder m;
  stage(&m, gmap);
  __syncthreads();
  const int i = blockIdx.x * kBT + threadIdx.x;
  if (i >= n) return;
  const int l = lane[i];
  const double px = x[i], py = y[i];
  double s, dis;
  project_distance(m.cs[l], m.n_cs[l], px, py, s, dis);
  double d = 0.0;
  if (dis > 0.0) {
    LineView lv{m.cpath[l], m.cdis[l], m.n_center[l], m.cs_len[l]};
    d = lv.side_of(s, px, py) > 0 ? dis : -dis;
  }
  d_out[i] = d;
  end_out[i] = m.cs_len[l] - s;
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

// ---- Corridor.distance_to_border(hull, direction) (type.py:183-206; environment.py:113-115 dis_to_lane_border):
// < 0 = how far the hull has passed the border on that side, otherwise Polygon(hull).distance(border line) ------------
__global__ void __launch_bounds__(kBT) env_border_kernel(const EnvMapBorder* __restrict__ gmap, const int32_t* __restrict__ lane,
                                                         const int32_t* __restrict__ side, const double* __restrict__ hull, size_t row,
                                                         int n, double* __restrict__ out) {
  __shared__ EnvMapBorder m;
  stage(&m, gmap);
  __syncthreads();
  const int i = blockIdx.x * kBT + threadIdx.x;
  if (i >= n) return;
  const int l = lane[i], sd = side[i];                       // 0 = "left", 1 = "right"
  double hx[4], hy[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { hx[k] = hull[(size_t)(2 * k) * row + i]; hy[k] = hull[(size_t)(2 * k + 1) * row + i]; }
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
  if (pass) { out[i] = -worst; return; }
  // Polygon(hull).distance(LineString): 0 when a vertex of the line lies inside the hull, else the nearest edge pair
  double shell[4][2];
#pragma unroll
  for (int k = 0; k < 4; ++k) { shell[k][0] = hx[k]; shell[k][1] = hy[k]; }
  for (int k = 0; k < ns; ++k)
    if (within(shell, 4, simp[k][0], simp[k][1])) { out[i] = 0.0; return; }
  double best = 0.0;
  bool have = false;
  for (int e = 0; e < 4; ++e) {
    const int e2 = (e + 1) & 3;
    for (int k = 0; k + 1 < ns; ++k) {
      const double d = seg_seg_distance(hx[e], hy[e], hx[e2], hy[e2], simp[k][0], simp[k][1], simp[k + 1][0], simp[k + 1][1]);
      if (!have || d < best) { best = d; have = true; }
    }
  }
  out[i] = best;
}

}  // namespace mce
