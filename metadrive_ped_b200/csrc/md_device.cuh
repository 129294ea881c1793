// md_device.cuh — device-side arithmetic of the batched MetaDrive step (sm_100a).
//
// Every function cites the reference file:line (relative to /root/reference/metadrive) whose behaviour it
// implements; Bullet-level pieces follow the published btRaycastVehicle / btTransformUtil algorithms.
// float32 throughout; this translation unit is compiled with -fmad=false so that a*b+c rounds twice exactly like
// the CPU oracle built with -ffp-contract=off (index / flag outputs are compared bit-exactly).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/md_layout.h"
#include "../../include/md_math.h"

#define MD_PI 3.14159265358979323846f
#define MD_TWO_PI 6.28318530717958647692f
#define SUSP_REST 0.4f             // panda3d bulletVehicle.cxx create_wheel default
#define MAX_SUSP_FORCE 6000.0f     // btRaycastVehicle::btVehicleTuning default
#define GRAVITY_Z (-9.81f)         // engine/core/physics_world.py:14
#define SUSP_STIFFNESS 40.0f       // component/vehicle/base_vehicle.py:97
#define SUSP_TRAVEL_CM 15.0f       // :96
#define DAMP_RELAX 4.8f            // :666
#define DAMP_COMP 1.2f             // :667
#define ROLL_INFLUENCE 0.5f        // :670
#define SIDE_DAMPING 0.2f          // btRaycastVehicle resolveSingleBilateral contactDamping
#define LIDAR_HEIGHT 1.2f          // component/sensors/lidar.py:19
#define LINE_HALF_W 0.0375f        // LANE_LINE_WIDTH / 4, component/block/base_block.py:503

struct F3 { float x, y, z; };
__device__ __forceinline__ F3 f3(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ F3 operator+(F3 a, F3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ F3 operator-(F3 a, F3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ F3 operator*(F3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot(F3 a, F3 b) { return md_dot3(a.x, a.y, a.z, b.x, b.y, b.z); }
__device__ __forceinline__ F3 cross(F3 a, F3 b) {
    return f3(md_diff2(a.y, b.z, a.z, b.y), md_diff2(a.z, b.x, a.x, b.z), md_diff2(a.x, b.y, a.y, b.x));
}
struct M3 { float m[3][3]; };  // world = m * local
__device__ __forceinline__ M3 quat_to_m3(float w, float x, float y, float z) {
    float n = md_dot4(w, x, y, z, w, x, y, z);
    float s = n > 0.0f ? 2.0f / n : 0.0f;
    M3 r;
    r.m[0][0] = 1.0f - s * md_sum2(y, y, z, z); r.m[0][1] = s * md_diff2(x, y, w, z); r.m[0][2] = s * md_sum2(x, z, w, y);
    r.m[1][0] = s * md_sum2(x, y, w, z); r.m[1][1] = 1.0f - s * md_sum2(x, x, z, z); r.m[1][2] = s * md_diff2(y, z, w, x);
    r.m[2][0] = s * md_diff2(x, z, w, y); r.m[2][1] = s * md_sum2(y, z, w, x); r.m[2][2] = 1.0f - s * md_sum2(x, x, y, y);
    return r;
}
__device__ __forceinline__ F3 col(const M3& r, int c) { return f3(r.m[0][c], r.m[1][c], r.m[2][c]); }
__device__ __forceinline__ F3 mul(const M3& r, F3 a) {
    return f3(md_dot3(r.m[0][0], r.m[0][1], r.m[0][2], a.x, a.y, a.z), md_dot3(r.m[1][0], r.m[1][1], r.m[1][2], a.x, a.y, a.z),
              md_dot3(r.m[2][0], r.m[2][1], r.m[2][2], a.x, a.y, a.z));
}
__device__ __forceinline__ F3 tmul(const M3& r, F3 a) {
    return f3(md_dot3(r.m[0][0], r.m[1][0], r.m[2][0], a.x, a.y, a.z), md_dot3(r.m[0][1], r.m[1][1], r.m[2][1], a.x, a.y, a.z),
              md_dot3(r.m[0][2], r.m[1][2], r.m[2][2], a.x, a.y, a.z));
}
__device__ __forceinline__ float clipf(float a, float lo, float hi) { return fminf(fmaxf(a, lo), hi); }
// utils/math.py:29-42
__device__ __forceinline__ float wrap_to_pi(float x) {
    float a;
    if (fabsf(x) <= 2.0f * MD_TWO_PI) {
        // |x| <= 4*pi: the conditional +-2*pi below are exact float operations (Sterbenz), hence identical to fmodf
        a = x;
        if (a >= MD_TWO_PI) a -= MD_TWO_PI;
        if (a <= -MD_TWO_PI) a += MD_TWO_PI;
    } else {
        a = fmodf(x, MD_TWO_PI);
    }
    if (a < 0.0f) a += MD_TWO_PI;
    if (a > MD_PI) a -= MD_TWO_PI;
    return a;
}
// base_vehicle.py:990-992 + base_object.py:390-398: heading = normalised xy projection of the chassis +Y axis
__device__ __forceinline__ void heading_vec(const M3& R, float& hx, float& hy) {
    float fx = R.m[0][1], fy = R.m[1][1];
    float n = sqrtf(fx * fx + fy * fy);
    hx = fx / n;
    hy = fy / n;
}

// ------------------------------------------------------------------------------------------------ lanes
// component/lane/straight_lane.py:60-74, circular_lane.py:57-121, abs_lane.py:76-90
__device__ __forceinline__ void lane_position(const float* __restrict__ L, float lon, float lat, float& x, float& y) {
    if (L[LF_TYPE] == 0.0f) {
        float dx = L[LF_P0 + 4], dy = L[LF_P0 + 5];
        x = L[LF_P0 + 0] + lon * dx + lat * dy;
        y = L[LF_P0 + 1] + lon * dy + lat * (-dx);
    } else {
        float r = L[LF_P0 + 2], dir = L[LF_P0 + 5];
        float phi = dir * lon / r + L[LF_P0 + 3];
        float rr = r + lat * dir;
        x = L[LF_P0 + 0] + rr * md_cosf(phi);
        y = L[LF_P0 + 1] + rr * md_sinf(phi);
    }
}
__device__ __forceinline__ float lane_heading_at(const float* __restrict__ L, float lon) {
    if (L[LF_TYPE] == 0.0f) return L[LF_P0 + 6];
    float dir = L[LF_P0 + 5];
    float phi = dir * lon / L[LF_P0 + 2] + L[LF_P0 + 3];
    return phi + (MD_PI / 2.0f) * dir;
}
__device__ __forceinline__ void lane_local(const float* __restrict__ L, float px, float py, float& lon, float& lat) {
    float ddx = px - L[LF_P0 + 0], ddy = py - L[LF_P0 + 1];
    if (L[LF_TYPE] == 0.0f) {
        float dx = L[LF_P0 + 4], dy = L[LF_P0 + 5];
        lon = ddx * dx + ddy * dy;
        lat = ddx * dy + ddy * (-dx);
    } else {
        float r = L[LF_P0 + 2], sp = L[LF_P0 + 3], ep = L[LF_P0 + 4], dir = L[LF_P0 + 5];
        float abs_phase = wrap_to_pi(md_atan2f(ddy, ddx));
        float sp_w = wrap_to_pi(sp), ep_w = wrap_to_pi(ep);
        float d_s = fabsf(wrap_to_pi(abs_phase - sp_w));
        float d_e = fabsf(wrap_to_pi(abs_phase - ep_w));
        bool clockwise = dir < 0.0f;
        if (d_s > d_e) {
            float diff = clockwise ? (ep - abs_phase) : (abs_phase - ep);
            lon = wrap_to_pi(diff) * r + L[LF_LENGTH];
        } else {
            float diff = clockwise ? (sp - abs_phase) : (abs_phase - sp);
            lon = wrap_to_pi(diff) * r;
        }
        lat = dir * (sqrtf(ddx * ddx + ddy * ddy) - r);
    }
}
__device__ __forceinline__ float lane_distance(const float* __restrict__ L, float px, float py) {
    float s, r;
    lane_local(L, px, py, s, r);
    float a = s - L[LF_LENGTH], b = 0.0f - s;
    return fabsf(r) + (a > 0.0f ? a : 0.0f) + (b > 0.0f ? b : 0.0f);
}
__device__ __forceinline__ bool lane_is_previous_of(const float* __restrict__ A, const float* __restrict__ B) {
    float dx = A[LF_EX] - B[LF_SX], dy = A[LF_EY] - B[LF_SY];
    return sqrtf(dx * dx + dy * dy) < 0.1f;
}
// component/block/base_block.py:459-466: lane body = convex hull of lane.polygon.  The vertex list is stored closed
// (n edges, n + 1 rows).  Edges are tested four at a time with their loads issued together; the verdict is the AND over
// all edges, so the grouping does not change the result.
__device__ __forceinline__ bool hull_edges(const float2* __restrict__ v, int i, int n, float px, float py) {
    for (; i + 4 <= n; i += 4) {
        float2 a = __ldg(v + i), b = __ldg(v + i + 1), c = __ldg(v + i + 2), d = __ldg(v + i + 3), e = __ldg(v + i + 4);
        float c0 = (b.x - a.x) * (py - a.y) - (b.y - a.y) * (px - a.x);
        float c1 = (c.x - b.x) * (py - b.y) - (c.y - b.y) * (px - b.x);
        float c2 = (d.x - c.x) * (py - c.y) - (d.y - c.y) * (px - c.x);
        float c3 = (e.x - d.x) * (py - d.y) - (e.y - d.y) * (px - d.x);
        // edge counts as inside (Bullet's hull has a 0.04 m margin)
        if (c0 < -1e-3f || c1 < -1e-3f || c2 < -1e-3f || c3 < -1e-3f) return false;
    }
    for (; i < n; i++) {
        float2 a = __ldg(v + i), b = __ldg(v + i + 1);
        float c0 = (b.x - a.x) * (py - a.y) - (b.y - a.y) * (px - a.x);
        if (c0 < -1e-3f) return false;
    }
    return true;
}
// L = the lane's lane_f row.  For arcs the outer-chord edges (stored last) are skipped when the point lies inside the
// circle inscribed in the chord polygon: every such chord is then satisfied with >= 1 cm to spare.
__device__ __forceinline__ bool point_in_hull(const float* __restrict__ L, const float* __restrict__ hull, int n, float px,
                                              float py) {
    const float2* __restrict__ v = reinterpret_cast<const float2*>(hull);
    const int n_other = (int)L[LF_HULL_NOTHER];
    if (!hull_edges(v, 0, n_other, px, py)) return false;
    if (n_other < n) {
        const float ddx = px - L[LF_P0 + 0], ddy = py - L[LF_P0 + 1], rin = L[LF_HULL_LONG];
        if (ddx * ddx + ddy * ddy > rin * rin) return hull_edges(v, n_other, n, px, py);
    }
    return true;
}

// Conservative shortcuts around point_in_hull, given the point's lane coordinates: +1 certainly inside, -1 certainly
// outside, 0 undecided (run the exact edge loop).  Both shortcuts are implied by the exact test with a >= 1 cm margin
// (float error and the 1e-3 edge tolerance are orders of magnitude smaller), so results stay bit-identical.
__device__ __forceinline__ int hull_shortcut(const float* __restrict__ L, float px, float py, float lon, float lat) {
    const float hw = 0.5f * L[LF_WIDTH];
    if (L[LF_TYPE] == 0.0f) {
        const float lmax = L[LF_HULL_LONG];
        if (lon >= 0.02f && lon <= lmax - 0.02f && fabsf(lat) <= hw - 0.02f) return 1;
        if (lon < -0.02f || lon > lmax + 0.02f || fabsf(lat) > hw + 0.02f) return -1;
        return 0;
    }
    // arc strip: the hull's outer boundary are the chords of the 1 m samples (at most 1.3 cm inside the true arc for
    // R >= 10 m); its corners reach sqrt((R + w/2)^2 + 1) because of the 1 m tangent extensions at both ends
    const float r = L[LF_P0 + 2];
    const float ro = r + hw;
    const float sag = ro / (8.0f * r * r) * 1.25f + 0.02f;  // chord sagitta of the outer 1 m samples, with slack
    if (lon >= 0.02f && lon <= L[LF_LENGTH] - 0.02f && fabsf(lat) <= hw - sag) return 1;
    const float ddx = px - L[LF_P0 + 0], ddy = py - L[LF_P0 + 1];
    if (ddx * ddx + ddy * ddy > ro * ro + 1.0f + 0.5f * ro) return -1;  // >= 25 cm beyond the farthest hull corner
    return 0;
}

// ------------------------------------------------------------------------------------------------ overlaps
struct Rect { float cx, cy, ux, uy, hu, hv; };

__device__ __forceinline__ bool rect_rect(const Rect& a, const Rect& b) {
    float dx = b.cx - a.cx, dy = b.cy - a.cy;
    float c = a.ux * b.ux + a.uy * b.uy;
    float s = a.ux * b.uy - a.uy * b.ux;
    float ac = fabsf(c), as = fabsf(s);
    if (fabsf(dx * a.ux + dy * a.uy) > a.hu + (b.hu * ac + b.hv * as)) return false;
    if (fabsf(-dx * a.uy + dy * a.ux) > a.hv + (b.hu * as + b.hv * ac)) return false;
    if (fabsf(dx * b.ux + dy * b.uy) > b.hu + (a.hu * ac + a.hv * as)) return false;
    if (fabsf(-dx * b.uy + dy * b.ux) > b.hv + (a.hu * as + a.hv * ac)) return false;
    return true;
}
__device__ __forceinline__ bool rect_circle(const Rect& a, float px, float py, float r) {
    float dx = px - a.cx, dy = py - a.cy;
    float lx = dx * a.ux + dy * a.uy;
    float ly = -dx * a.uy + dy * a.ux;
    float qx = clipf(lx, -a.hu, a.hu), qy = clipf(ly, -a.hv, a.hv);
    float ex = lx - qx, ey = ly - qy;
    return ex * ex + ey * ey <= r * r;
}
__device__ __forceinline__ bool rect_quad(const Rect& a, const float* __restrict__ q) {
    float rx[4], ry[4];
    float vx = -a.uy, vy = a.ux;
    rx[0] = a.cx + a.ux * a.hu + vx * a.hv; ry[0] = a.cy + a.uy * a.hu + vy * a.hv;
    rx[1] = a.cx - a.ux * a.hu + vx * a.hv; ry[1] = a.cy - a.uy * a.hu + vy * a.hv;
    rx[2] = a.cx - a.ux * a.hu - vx * a.hv; ry[2] = a.cy - a.uy * a.hu - vy * a.hv;
    rx[3] = a.cx + a.ux * a.hu - vx * a.hv; ry[3] = a.cy + a.uy * a.hu - vy * a.hv;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        float ax = k == 0 ? a.ux : vx, ay = k == 0 ? a.uy : vy;
        float h = k == 0 ? a.hu : a.hv;
        float c0 = a.cx * ax + a.cy * ay;
        float lo = 1e30f, hi = -1e30f;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float p = q[2 * i] * ax + q[2 * i + 1] * ay;
            lo = fminf(lo, p); hi = fmaxf(hi, p);
        }
        if (lo > c0 + h || hi < c0 - h) return false;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int j = (i + 1) & 3;
        float ex = q[2 * j] - q[2 * i], ey = q[2 * j + 1] - q[2 * i + 1];
        float nx = ey, ny = -ex;
        float qmax = q[2 * i] * nx + q[2 * i + 1] * ny;
        float rmin = 1e30f;
#pragma unroll
        for (int k = 0; k < 4; k++) rmin = fminf(rmin, rx[k] * nx + ry[k] * ny);
        if (rmin > qmax) return false;
    }
    return true;
}
// chassis footprint: box (W/2, L/2, H/2) offset +H/2 above the body origin (base_vehicle.py:588-590)
__device__ __forceinline__ Rect vehicle_rect(const float* __restrict__ P, const float* S) {
    M3 R = quat_to_m3(S[VS_QUAT], S[VS_QUAT + 1], S[VS_QUAT + 2], S[VS_QUAT + 3]);
    float hx, hy;
    heading_vec(R, hx, hy);
    float hh = 0.5f * P[VP_HEIGHT];
    Rect r;
    r.cx = S[VS_POS + 0] + R.m[0][2] * hh;
    r.cy = S[VS_POS + 1] + R.m[1][2] * hh;
    r.ux = hx; r.uy = hy;
    r.hu = 0.5f * P[VP_LENGTH];
    r.hv = 0.5f * P[VP_WIDTH];
    return r;
}
// ------------------------------------------------------------------------------------------------ contact response
// The dynamic-world part of BulletWorld.doPhysics (engine/core/engine_core.py:350-352) that pushes bodies apart.  Bullet's
// sequential-impulse solver over GJK/EPA manifolds is not restated: this is the simple impulse model SURVEY.md 7.3 #2
// allows (DESIGN.md "Contact response"): planar contacts between chassis footprints and static obstacles, one Jacobi pass
// per sub-step from a snapshot of the post-move state, frictionless inelastic normal impulse + a position push-out of
// ERP x (depth - slop) split by inverse mass.  Every product and sum is written in one fixed order (the CPU checker in the test tree follows the same).
#define CONTACT_ERP 0.2f
#define CONTACT_SLOP 0.01f
#define CORNER_EPS 1e-4f
#define CONTACT_TIE 1e-6f
#define AXIS_MARGIN 1e-3f
struct __align__(16) CBody { float ox, oy, vx, vy, w, im, ii, pad; };   // origin, planar velocity, yaw rate, 1/m, 1/Izz

__device__ __forceinline__ void corners_inside(const Rect& in, const Rect& box, int& cnt, float& sx, float& sy) {
    const float vx = -in.uy, vy = in.ux;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float su = (k == 0 || k == 3) ? 1.0f : -1.0f, sv = k < 2 ? 1.0f : -1.0f;
        const float px = in.cx + in.ux * (su * in.hu) + vx * (sv * in.hv);
        const float py = in.cy + in.uy * (su * in.hu) + vy * (sv * in.hv);
        const float dx = px - box.cx, dy = py - box.cy;
        if (fabsf(dx * box.ux + dy * box.uy) <= box.hu + CORNER_EPS &&
            fabsf(-dx * box.uy + dy * box.ux) <= box.hv + CORNER_EPS) {
            cnt += 1; sx += px; sy += py;
        }
    }
}
// normal (A -> B), depth and contact point of two overlapping rectangles; false when they do not overlap (rect_rect's verdict)
__device__ __forceinline__ bool rr_contact(const Rect& a, const Rect& b, float& nx, float& ny, float& depth, float& px, float& py) {
    const float dx = b.cx - a.cx, dy = b.cy - a.cy;
    const float c = a.ux * b.ux + a.uy * b.uy;
    const float s = a.ux * b.uy - a.uy * b.ux;
    const float ac = fabsf(c), as = fabsf(s);
    const float s0 = dx * a.ux + dy * a.uy, ov0 = a.hu + (b.hu * ac + b.hv * as) - fabsf(s0);
    const float s1 = -dx * a.uy + dy * a.ux, ov1 = a.hv + (b.hu * as + b.hv * ac) - fabsf(s1);
    const float s2 = dx * b.ux + dy * b.uy, ov2 = b.hu + (a.hu * ac + a.hv * as) - fabsf(s2);
    const float s3 = -dx * b.uy + dy * b.ux, ov3 = b.hv + (a.hu * as + a.hv * ac) - fabsf(s3);
    if (ov0 < 0.0f || ov1 < 0.0f || ov2 < 0.0f || ov3 < 0.0f) return false;
    // minimum-translation axis; B's axes win only by a clear margin: two nearly parallel boxes overlap by almost the same
    // amount along A's and B's axis and the pick would otherwise flip on rounding noise
    float best = ov0, ax = a.ux, ay = a.uy, sg = s0;
    if (ov1 < best) { best = ov1; ax = -a.uy; ay = a.ux; sg = s1; }
    float bb = ov2, bx = b.ux, by = b.uy, bs = s2;
    if (ov3 < bb) { bb = ov3; bx = -b.uy; by = b.ux; bs = s3; }
    if (bb < best - AXIS_MARGIN) { best = bb; ax = bx; ay = by; sg = bs; }
    if (sg > -CONTACT_TIE) { nx = ax; ny = ay; } else { nx = -ax; ny = -ay; }  // centres level: "+axis" by convention
    depth = best;
    int cnt = 0; float sx = 0.0f, sy = 0.0f;
    corners_inside(b, a, cnt, sx, sy);
    corners_inside(a, b, cnt, sx, sy);
    if (cnt > 0) { px = sx / (float)cnt; py = sy / (float)cnt; }
    else { px = 0.5f * (a.cx + b.cx); py = 0.5f * (a.cy + b.cy); }
    return true;
}
// rectangle A against a circle: normal A -> circle (rect_circle's verdict)
__device__ __forceinline__ bool rc_contact(const Rect& a, float cx, float cy, float r, float& nx, float& ny, float& depth,
                                           float& px, float& py) {
    const float dx = cx - a.cx, dy = cy - a.cy;
    const float vx = -a.uy, vy = a.ux;
    const float lx = dx * a.ux + dy * a.uy;
    const float ly = -dx * a.uy + dy * a.ux;
    const float qx = clipf(lx, -a.hu, a.hu), qy = clipf(ly, -a.hv, a.hv);
    const float ex = lx - qx, ey = ly - qy;
    const float d2 = ex * ex + ey * ey;
    if (!(d2 <= r * r)) return false;
    if (d2 > 1e-12f) {
        const float dist = sqrtf(d2);
        const float e0 = ex / dist, e1 = ey / dist;
        nx = a.ux * e0 + vx * e1; ny = a.uy * e0 + vy * e1;
        depth = r - dist;
        px = a.cx + a.ux * qx + vx * qy; py = a.cy + a.uy * qx + vy * qy;
        return true;
    }
    const float fx = a.hu - fabsf(lx), fy = a.hv - fabsf(ly);
    if (fx < fy) {
        const float sg = lx >= 0.0f ? 1.0f : -1.0f;
        nx = a.ux * sg; ny = a.uy * sg; depth = fx + r;
        px = a.cx + a.ux * (sg * a.hu) + vx * ly; py = a.cy + a.uy * (sg * a.hu) + vy * ly;
    } else {
        const float sg = ly >= 0.0f ? 1.0f : -1.0f;
        nx = vx * sg; ny = vy * sg; depth = fy + r;
        px = a.cx + a.ux * lx + vx * (sg * a.hv); py = a.cy + a.uy * lx + vy * (sg * a.hv);
    }
    return true;
}
// one pair's contribution to body `me` (A when !me_is_b): d = dvx dvy dw dpx dpy
__device__ __forceinline__ void pair_impulse(const CBody& A, const CBody& B, float nx, float ny, float depth, float px, float py,
                                             bool me_is_b, float* d) {
    const float rax = px - A.ox, ray = py - A.oy, rbx = px - B.ox, rby = py - B.oy;
    const float vpax = A.vx + A.w * (-ray), vpay = A.vy + A.w * rax;
    const float vpbx = B.vx + B.w * (-rby), vpby = B.vy + B.w * rbx;
    const float vn = (vpbx - vpax) * nx + (vpby - vpay) * ny;
    const float can = rax * ny - ray * nx, cbn = rbx * ny - rby * nx;
    const float K = A.im + B.im + can * can * A.ii + cbn * cbn * B.ii;
    const float J = vn < 0.0f ? -vn / K : 0.0f;
    const float push = CONTACT_ERP * fmaxf(depth - CONTACT_SLOP, 0.0f) / (A.im + B.im);
    const float sgn = me_is_b ? 1.0f : -1.0f, cn = me_is_b ? cbn : can;
    const float mim = me_is_b ? B.im : A.im, mii = me_is_b ? B.ii : A.ii;
    const float sj = sgn * J * mim;
    d[0] += nx * sj; d[1] += ny * sj;
    d[2] += sgn * J * cn * mii;
    const float sp = sgn * push * mim;
    d[3] += nx * sp; d[4] += ny * sp;
}

// TrafficBarrier: BulletBoxShape((WIDTH/2, LENGTH/2, h/2)) (static_object/traffic_object.py:143)
__device__ __forceinline__ Rect object_rect(const float* O) {
    Rect r;
    r.cx = O[OB_X]; r.cy = O[OB_Y];
    r.ux = md_cosf(O[OB_HEADING]); r.uy = md_sinf(O[OB_HEADING]);
    r.hu = O[OB_B]; r.hv = O[OB_A];
    return r;
}

// ------------------------------------------------------------------------------------------------ ray casts
// sensors/distance_detector.py:27-85 (rayTestClosest); t in [0,1], 2 = miss, 0 if the origin is inside
__device__ __forceinline__ float ray_obb(F3 o, F3 d, F3 c, const M3& R, F3 h) {
    F3 ol = tmul(R, o - c);
    F3 dl = tmul(R, d);
    float t0 = 0.0f, t1 = 1.0f;
    const float olv[3] = {ol.x, ol.y, ol.z}, dlv[3] = {dl.x, dl.y, dl.z}, hv[3] = {h.x, h.y, h.z};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (fabsf(dlv[k]) < 1e-12f) {
            if (fabsf(olv[k]) > hv[k]) return 2.0f;
        } else {
            float inv = 1.0f / dlv[k];
            float ta = (-hv[k] - olv[k]) * inv, tb = (hv[k] - olv[k]) * inv;
            if (ta > tb) { float t = ta; ta = tb; tb = t; }
            t0 = fmaxf(t0, ta);
            t1 = fminf(t1, tb);
            if (t0 > t1) return 2.0f;
        }
    }
    return t0;
}
__device__ __forceinline__ float ray_zcyl(F3 o, F3 d, float cx, float cy, float zc, float r, float hh) {
    if (fabsf(o.z - zc) > hh) return 2.0f;
    float ox = o.x - cx, oy = o.y - cy;
    float a = d.x * d.x + d.y * d.y;
    float b = ox * d.x + oy * d.y;
    float cc = ox * ox + oy * oy - r * r;
    float disc = b * b - a * cc;
    if (disc < 0.0f) return 2.0f;
    float sq = sqrtf(disc);
    float ta = (-b - sq) / a, tb = (-b + sq) / a;
    float t0 = fmaxf(0.0f, ta), t1 = fminf(1.0f, tb);
    if (t0 > t1) return 2.0f;
    return t0;
}

// ------------------------------------------------------------------------------------------------ dynamics
struct Actuation { float steer_rad, engine, brake; };
struct SteerCS { float cs, sn; };  // cos / sin of the front-wheel steering angle, constant over the sub-steps

// btTransformUtil::integrateTransform (rotation part); q = w x y z
__device__ __forceinline__ void quat_integrate(float* q, F3 w, float dt) {
    float ang = sqrtf(dot(w, w));
    if (ang * dt > 0.25f * MD_PI) ang = 0.25f * MD_PI / dt;
    F3 axis;
    if (ang < 0.001f) axis = w * (0.5f * dt - (dt * dt * dt) * 0.020833333333f * ang * ang);
    else axis = w * (md_sinf(0.5f * ang * dt) / ang);
    float aw = md_cosf(ang * dt * 0.5f), ax = axis.x, ay = axis.y, az = axis.z;
    float bw = q[0], bx = q[1], by = q[2], bz = q[3];
    float rw = md_dot4(aw, -ax, -ay, -az, bw, bx, by, bz);
    float rx = md_dot4(aw, ax, ay, -az, bx, bw, bz, by);
    float ry = md_dot4(aw, -ax, ay, az, by, bz, bw, bx);
    float rz = md_dot4(aw, ax, -ay, az, bz, by, bx, bw);
    float n = sqrtf(md_dot4(rw, rx, ry, rz, rw, rx, ry, rz));
    const float inv_n = 1.0f / n;   /* one division for the four components */
    q[0] = rw * inv_n; q[1] = rx * inv_n; q[2] = ry * inv_n; q[3] = rz * inv_n;
}

struct Body {  // rigid-body state kept in registers across the sub-steps
    F3 pos, v, w;
    float q[4];
};

// world-frame inverse inertia applied to a torque impulse; iIx.. are the reciprocals of the principal moments
// (Bullet keeps m_invInertiaLocal and multiplies, btRigidBody::updateInertiaTensor)
__device__ __forceinline__ F3 apply_inv_inertia(const M3& R, F3 t, float iIx, float iIy, float iIz) {
    F3 tl = tmul(R, t);
    tl = f3(tl.x * iIx, tl.y * iIy, tl.z * iIz);
    return mul(R, tl);
}

// one doPhysics(dt, 1, dt) sub-step of a chassis on four ray-cast wheels over the plane z = 0
// (base_vehicle.py:577-598,632-671; engine_core.py:350-352; btRaycastVehicle::updateVehicle / updateFriction)
__device__ void vehicle_substep(const float* __restrict__ P, Body& B, Actuation act, SteerCS scs, float dt) {
    const float mass = P[VP_MASS], inv_m = 1.0f / mass;
    const float lx = P[VP_WIDTH], ly = P[VP_LENGTH], lz = P[VP_HEIGHT];
    const float Ix = 1.0f / (mass / 12.0f * (ly * ly + lz * lz)), Iy = 1.0f / (mass / 12.0f * (lx * lx + lz * lz)),
                Iz = 1.0f / (mass / 12.0f * (lx * lx + ly * ly));  // reciprocal principal moments
    F3 pos = B.pos, v = B.v, w = B.w;
    v.z += GRAVITY_Z * dt;
    float wl = sqrtf(dot(w, w));
    if (wl * dt > 0.5f * MD_PI) w = w * ((0.5f * MD_PI / dt) / wl);
    pos = pos + v * dt;
    quat_integrate(B.q, w, dt);
    M3 R = quat_to_m3(B.q[0], B.q[1], B.q[2], B.q[3]);
    {   // chassis box vs ground plane: inelastic clamp of the lowest bottom corner (DESIGN.md "Dynamics")
        float hw = 0.5f * lx, hl = 0.5f * ly;
        float lowest = 0.0f;
#pragma unroll
        for (int sx = -1; sx <= 1; sx += 2)
#pragma unroll
            for (int sy = -1; sy <= 1; sy += 2)
                lowest = fminf(lowest, pos.z + R.m[2][0] * (float)sx * hw + R.m[2][1] * (float)sy * hl);
        if (lowest < 0.0f) {
            pos.z -= lowest;
            if (v.z < 0.0f) v.z = 0.0f;
        }
    }
    F3 up = col(R, 2), right = col(R, 0);
    F3 normal = f3(0.0f, 0.0f, 1.0f);
    const float conn_z = P[VP_TIRE_R] - P[VP_CHASSIS_AXIS];
    bool on[4];
    F3 cp[4];
    float fs[4];
    int n_ground = 0;
    const float raylen = SUSP_REST + P[VP_TIRE_R];
    F3 dir = up * -1.0f;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float cxw = (i & 1) ? -P[VP_LATERAL] : P[VP_LATERAL];
        float cyw = i < 2 ? P[VP_FRONT_WB] : -P[VP_REAR_WB];
        F3 hard = pos + mul(R, f3(cxw, cyw, conn_z));
        on[i] = false; fs[i] = 0.0f; cp[i] = f3(0, 0, 0);
        if (!(dir.z < 0.0f && hard.z > 0.0f)) continue;
        float t = hard.z / (-dir.z * raylen);
        if (t > 1.0f) continue;
        on[i] = true; n_ground++;
        cp[i] = hard + dir * (raylen * t);
        float slen = t * raylen - P[VP_TIRE_R];
        slen = clipf(slen, SUSP_REST - SUSP_TRAVEL_CM * 0.01f, SUSP_REST + SUSP_TRAVEL_CM * 0.01f);
        float denom = dot(normal, dir);
        F3 rel = cp[i] - pos;
        float proj_vel = dot(normal, v + cross(w, rel));
        float rel_vel, clipped_inv;
        if (denom >= -0.1f) { rel_vel = 0.0f; clipped_inv = 10.0f; }
        else { float inv = -1.0f / denom; rel_vel = proj_vel * inv; clipped_inv = inv; }
        float force = SUSP_STIFFNESS * (SUSP_REST - slen) * clipped_inv;
        force -= (rel_vel < 0.0f ? DAMP_COMP : DAMP_RELAX) * rel_vel;
        fs[i] = fmaxf(force * mass, 0.0f);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (!on[i]) continue;
        float f = fminf(fs[i], MAX_SUSP_FORCE);
        F3 imp = normal * (f * dt);
        F3 rel = cp[i] - pos;
        v = v + imp * inv_m;
        w = w + apply_inv_inertia(R, cross(rel, imp), Ix, Iy, Iz);
    }
    if (n_ground > 0) {
        // the two front wheels share one steering angle and the rear wheels are unsteered, so there are only two
        // distinct (axle, forward) pairs; computing each once is bit-identical to computing it per wheel
        F3 axle2[2], fwd2[2];
        float side[4] = {0, 0, 0, 0}, fimp[4] = {0, 0, 0, 0}, skid[4] = {1, 1, 1, 1};
        bool sliding = false;
#pragma unroll
        for (int p = 0; p < 2; p++) {
            float cs = p == 0 ? scs.cs : 1.0f, sn = p == 0 ? scs.sn : 0.0f;
            F3 a = (right * cs + cross(up, right) * sn) + up * (dot(up, right) * (1.0f - cs));
            a = a - normal * dot(a, normal);
            a = a * (1.0f / sqrtf(dot(a, a)));
            F3 f = cross(normal, a);
            f = f * (1.0f / sqrtf(dot(f, f)));
            axle2[p] = a; fwd2[p] = f;
        }
#define AXLE(i) axle2[(i) < 2 ? 0 : 1]
#define FWD(i) fwd2[(i) < 2 ? 0 : 1]
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (!on[i]) continue;
            F3 a = AXLE(i);
            F3 rel = cp[i] - pos;
            F3 aj = tmul(R, cross(rel, a));
            float jac = inv_m + (aj.x * aj.x * Ix + aj.y * aj.y * Iy + aj.z * aj.z * Iz);
            float rel_vel = dot(a, v + cross(w, rel));
            side[i] = -SIDE_DAMPING * rel_vel / jac;
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (!on[i]) continue;
            float rolling;
            if (act.engine != 0.0f) rolling = act.engine * dt;
            else {
                float max_imp = act.brake;
                F3 rel = cp[i] - pos;
                F3 c0 = cross(rel, FWD(i));
                F3 iw = apply_inv_inertia(R, c0, Ix, Iy, Iz);
                float denom0 = inv_m + dot(FWD(i), cross(iw, rel));
                float vrel = dot(FWD(i), v + cross(w, rel));
                float j1 = -vrel / (denom0 * (float)n_ground);
                rolling = clipf(j1, -max_imp, max_imp);
            }
            fimp[i] = rolling;
            float maximp = fs[i] * dt * P[VP_FRICTION];
            float x = fimp[i] * 0.5f, y = side[i];
            float imp2 = x * x + y * y;
            if (imp2 > maximp * maximp) { sliding = true; skid[i] *= maximp / sqrtf(imp2); }
        }
        if (sliding) {
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (side[i] != 0.0f && skid[i] < 1.0f) { fimp[i] *= skid[i]; side[i] *= skid[i]; }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (!on[i]) continue;
            F3 rel = cp[i] - pos;
            if (fimp[i] != 0.0f) {
                F3 imp = FWD(i) * fimp[i];
                v = v + imp * inv_m;
                w = w + apply_inv_inertia(R, cross(rel, imp), Ix, Iy, Iz);
            }
            if (side[i] != 0.0f) {
                F3 imp = AXLE(i) * side[i];
                F3 rel2 = rel - up * (dot(up, rel) * (1.0f - ROLL_INFLUENCE));
                v = v + imp * inv_m;
                w = w + apply_inv_inertia(R, cross(rel2, imp), Ix, Iy, Iz);
            }
        }
    }
#undef AXLE
#undef FWD
    B.pos = pos; B.v = v; B.w = w;
}

// EnvInputPolicy.convert_to_continuous_action (policy/env_input_policy.py:40-48); the reference computes in float64
__device__ __forceinline__ void decode_action(const MdConfig& cfg, float& a0, float& a1) {
    if (cfg.discrete_action == 0) return;
    const double su = 2.0 / (double)(cfg.discrete_steering_dim - 1), tu = 2.0 / (double)(cfg.discrete_throttle_dim - 1);
    if (cfg.discrete_action == 2) { a0 = (float)((double)a0 * su - 1.0); a1 = (float)((double)a1 * tu - 1.0); return; }
    const int idx = (int)a0;
    a0 = (float)((double)(idx % cfg.discrete_steering_dim) * su - 1.0);
    a1 = (float)((double)(idx / cfg.discrete_steering_dim) * tu - 1.0);
}
// BaseVehicle._preprocess_action + _set_action + _apply_throttle_brake (base_vehicle.py:204-209, 447-484)
__device__ __forceinline__ float scrub(float a) {  // utils/math.py:16-26
    if (isnan(a)) return 0.0f;
    return clipf(a, -1.0f, 1.0f);
}
__device__ __forceinline__ Actuation actuate(const float* __restrict__ P, float* S, float* C, float a0, float a1) {
    a0 = scrub(a0); a1 = scrub(a1);
    C[VC_PREV_A0] = C[VC_CUR_A0]; C[VC_PREV_A1] = C[VC_CUR_A1];
    C[VC_CUR_A0] = a0; C[VC_CUR_A1] = a1;
    S[VS_STEER] = a0; S[VS_THROTTLE] = a1;
    Actuation act;
    act.steer_rad = a0 * P[VP_MAX_STEER] * (MD_PI / 180.0f);
    float speed_kmh = sqrtf(S[VS_VEL] * S[VS_VEL] + S[VS_VEL + 1] * S[VS_VEL + 1]) * 3.6f;
    if (a1 >= 0.0f) {
        act.brake = 2.0f;
        act.engine = speed_kmh > P[VP_MAX_SPEED] ? 0.0f : P[VP_ENGINE] * a1;
    } else if (P[VP_REVERSE] != 0.0f) {
        act.engine = P[VP_ENGINE] * a1; act.brake = 0.0f;
    } else {
        act.engine = 0.0f; act.brake = fabsf(a1) * P[VP_BRAKE];
    }
    return act;
}
// BaseVehicle.before_step latches (base_vehicle.py:211-232)
__device__ __forceinline__ void latch_before_step(const float* S, float* C, int* I) {
    I[VI_FLAGS] = FL_ON_LANE;
    C[VC_LAST_X] = S[VS_POS]; C[VC_LAST_Y] = S[VS_POS + 1];
    M3 R = quat_to_m3(S[VS_QUAT], S[VS_QUAT + 1], S[VS_QUAT + 2], S[VS_QUAT + 3]);
    heading_vec(R, C[VC_LAST_HX], C[VC_LAST_HY]);
}

// ------------------------------------------------------------------------------------------------ map views
// Grid records (built by the library at md_load_scene from the tables above, private to it): what a grid walk needs about
// an item sits NEXT TO the item, indexed like lgrid_items / grid_items, so that a team of lanes fetches the records of all
// its items in one round trip instead of chasing item id -> table row once per item.
//   lrec[2k], lrec[2k+1] : lane id (int bits), hull AABB xmin ymin xmax | ymax, hull offset (int bits), hull points (int bits), lane type
//   irec[2k], irec[2k+1] : the 8 floats of the line_f / quad_f row of grid item k
struct MapAccel { const float4* lrec; const float4* irec; };
struct MapView {
    const float* lane_f; const int* lane_i; const float* lane_bb; const int* road_i; const float* hull;
    const float* lines; const float* quads; const int* gs; const int* gi; const int* lgs; const int* lgi;
    const float4* lrec; const float4* irec;
    int n_lanes, n_roads, n_lines, nx, ny;
    float gx0, gy0, cell;
};
__device__ __forceinline__ MapView map_view(const MdArrays& A, int map, const MapAccel* X = nullptr) {
    const int* d = A.map_desc + (size_t)map * MAPD;
    const float* df = A.map_descf + (size_t)map * MAPDF;
    MapView m;
    m.lane_f = A.lane_f + (size_t)d[MD_LANE_OFF] * LANE_F;
    m.lane_i = A.lane_i + (size_t)d[MD_LANE_OFF] * LANE_I;
    m.lane_bb = A.lane_bb + (size_t)d[MD_LANE_OFF] * 4;
    m.road_i = A.road_i + (size_t)d[MD_ROAD_OFF] * ROAD_I;
    m.hull = A.hull_xy + (size_t)d[MD_HULL_OFF] * 2;
    m.lines = A.line_f + (size_t)d[MD_LINE_OFF] * LINE_F;
    m.quads = A.quad_f + (size_t)d[MD_QUAD_OFF] * QUAD_F;
    m.gs = A.grid_start + d[MD_GRID_OFF];
    m.gi = A.grid_items + d[MD_ITEM_OFF];
    m.lgs = A.lgrid_start + d[MD_LGRID_OFF];
    m.lgi = A.lgrid_items + d[MD_LITEM_OFF];
    m.lrec = X != nullptr ? X->lrec + 2 * (size_t)d[MD_LITEM_OFF] : nullptr;
    m.irec = X != nullptr ? X->irec + 2 * (size_t)d[MD_ITEM_OFF] : nullptr;
    m.n_lanes = d[MD_N_LANES]; m.n_roads = d[MD_N_ROADS]; m.n_lines = d[MD_N_LINES];
    m.nx = d[MD_GRID_NX]; m.ny = d[MD_GRID_NY];
    m.gx0 = df[0]; m.gy0 = df[1]; m.cell = df[2];
    return m;
}
__device__ __forceinline__ int find_road(const MapView& m, int from, int to) {
    for (int r = 0; r < m.n_roads; r++)
        if (m.road_i[r * ROAD_I + RI_FROM] == from && m.road_i[r * ROAD_I + RI_TO] == to) return r;
    return -1;
}
