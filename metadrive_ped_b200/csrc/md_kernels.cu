// md_kernels.cu — kernels + C ABI of libmdstep.so (sm_100a).  See include/mdstep.h for the boundary and
// DESIGN.md for the data layout, the launch structure and the roofline of each kernel.
//
// Launch structure of one env.step:
//   k_step_vehicles  thread per vehicle slot, a CTA owns whole envs: before_step (actuation, trigger, IDM),
//                    decision_repeat sub-steps with the state in registers and the env's footprints in shared
//                    memory for the contact pass, after_step (localisation, state check), reward / cost / done and
//                    the 19 state floats of the observation; publishes one 64-byte "body row" per vehicle.
//   k_lidar          warp per agent: the env's body rows + object rows are staged into shared memory with one
//                    cp.async.bulk (TMA 1-D bulk copy) per table completing on an mbarrier; 240 rays, 7.5 per lane,
//                    nearest hit kept in registers; coalesced store of the 240 lidar floats.
//   k_done_mask / k_reset_vehicles / k_lidar(masked)   batched auto-reset, still without a host round trip.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/mdstep.h"
#include "md_device.cuh"

// Debug builds only (make EXTRA=-DMD_PHASE_CLK, scripts/phase_clk.py): thread 0 of every CTA stamps clock64() at the phase
// boundaries of the step kernels into a device table [kernel][CTA][16]; entry 15 is the globaltimer at CTA start.
#ifdef MD_PHASE_CLK
__device__ unsigned long long* g_phase_clk = nullptr;
#define CLK_CTAS 4096
__device__ __forceinline__ void clk_mark(int kern, int phase) {
    if (g_phase_clk != nullptr && threadIdx.x == 0 && blockIdx.x < CLK_CTAS) {
        unsigned long long* row = g_phase_clk + ((size_t)kern * CLK_CTAS + blockIdx.x) * 16;
        row[phase] = (unsigned long long)clock64();
        if (phase == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); row[15] = t;
                          unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); row[14] = sm; }
        if (phase >= 13) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); row[12] = t; }
    }
}
extern "C" int md_debug_phase_clk(unsigned long long* dev_table) {
    return cudaMemcpyToSymbol(g_phase_clk, &dev_table, sizeof(dev_table)) == cudaSuccess ? 0 : -1;
}
#else
#define clk_mark(kern, phase)
#endif

#define STEP_THREADS 128
// register caps per kernel (__maxnreg__): the block size is a run-time choice (envs per CTA x slots per env), so the
// occupancy is steered through the register budget instead of __launch_bounds__
#ifndef PRE_REGS
#define PRE_REGS 64
#endif
#ifndef DYN_REGS
#define DYN_REGS 80
#endif
#ifndef POST_REGS
#define POST_REGS 128
#endif
#ifndef PRE_EPB
#define PRE_EPB 16
#endif
#ifndef POST_EPB
#define POST_EPB 16
#endif
#ifndef DYN_EPB
#define DYN_EPB 8
#endif
#define LIDAR_WARPS 8
#define MAX_LASERS 512

// ray direction tables (cos, sin per laser) live in global memory, one set per handle: every lane reads a different entry
// (constant memory would serialise that), and two handles with different laser counts must not share them

enum {
    MODE_AGENT_PRE = 1, MODE_TRIGGER = 2, MODE_IDM = 4, MODE_DYN = 8, MODE_CONTACTS = 16, MODE_POST = 32,
    MODE_OUT = 64, MODE_EXT_ACT = 128, MODE_IDM_OUT = 256, MODE_RESET = 512, MODE_REMOVE = 1024, MODE_CLEAR_FLAGS = 2048,
    MODE_MARK_DONE = 4096,  // k_post writes done_mask[env] (single agent: the agent terminated or truncated)
    MODE_RESTORE = 8192,    // k_post(MODE_RESET) first restores the masked envs from the reset snapshot
    MODE_FULL = MODE_AGENT_PRE | MODE_TRIGGER | MODE_IDM | MODE_DYN | MODE_CONTACTS | MODE_POST | MODE_OUT | MODE_REMOVE
};

struct StepOut {
    float* obs; float* reward; float* cost; uint8_t* term; uint8_t* trunc; int* info_flags; float* info_f;
};
struct Snapshot {
    const int* env_i; const float* veh_s; const float* veh_c; const int* veh_i; const float* veh_idm;
    const float* veh_navi; const float* obj_f; const int* veh_route; const int* veh_rroad;
    const float* veh_p; const int* env_trigger;   // only read by a `full` restore (after scenario-bank draws)
};

// neighbour record of one vehicle slot, shared by the threads of its env
struct __align__(16) Nb {
    float x, y, vx, vy;
    Rect r;
    int lane, alive, active, kind;
};

// entries of an IDM team's neighbour table (see pre_team_size)
#define PRE_NB_CAP 32
__host__ __device__ inline int pre_nb_cap(int S, int O) { return S + O < PRE_NB_CAP ? S + O : PRE_NB_CAP; }
struct FrontBack { int fobj[3], bobj[3]; float fdist[3], bdist[3]; bool exists[3]; };
#define OBJ_NONE (-1)

// unified view of "the k-th surrounding object" for the IDM: vehicles first (by slot), then objects
struct NbrView {
    const Nb* nb; const float* obj; int S, O, self; float px, py;
    __device__ __forceinline__ int count() const { return S + O; }
    __device__ __forceinline__ bool valid(int k) const {
        if (k < S) {
            if (k == self || !nb[k].alive) return false;
            return rect_circle(nb[k].r, px, py, 50.0f);
        }
        const float* Ob = obj + (k - S) * OBJ_F;
        if (Ob[OB_KIND] < 0.0f) return false;
        if (OB_IS_BOX(Ob[OB_KIND])) { Rect r = object_rect(Ob); return rect_circle(r, px, py, 50.0f); }
        float dx = Ob[OB_X] - px, dy = Ob[OB_Y] - py, rr = 50.0f + Ob[OB_A];
        return dx * dx + dy * dy <= rr * rr;
    }
    __device__ __forceinline__ void get(int k, float& x, float& y, float& vx, float& vy, int& lane, bool& ped) const {
        if (k < S) { x = nb[k].x; y = nb[k].y; vx = nb[k].vx; vy = nb[k].vy; lane = nb[k].lane; ped = false; }
        else {
            const float* Ob = obj + (k - S) * OBJ_F;
            x = Ob[OB_X]; y = Ob[OB_Y]; vx = Ob[OB_VX]; vy = Ob[OB_VY]; lane = (int)Ob[OB_LANE]; ped = Ob[OB_KIND] == 3.0f;
        }
    }
};

// FrontBackObjects.get_find_front_back_objs (policy/idm_policy.py:82-132)
__device__ void find_front_back(const MapView& m, const NbrView& nv, unsigned long long valid_lo, unsigned long long valid_hi,
                                int lane, float px, float py, float max_d, bool use_ref, int ref_first, int ref_n,
                                FrontBack& out) {
    int lanes[3] = {-1, lane, -1};
    if (use_ref) {
        int idx = m.lane_i[lane * LANE_I + LI_IDX];
        if (idx > 0) lanes[0] = ref_first + idx - 1;
        if (idx + 1 < ref_n) lanes[2] = ref_first + idx + 1;
    }
    for (int i = 0; i < 3; i++) {
        out.fobj[i] = out.bobj[i] = OBJ_NONE;
        out.exists[i] = lanes[i] >= 0;
        out.fdist[i] = out.bdist[i] = max_d;
        if (lanes[i] < 0) continue;
        const float* Li = m.lane_f + lanes[i] * LANE_F;
        float cur_long, lat;
        lane_local(Li, px, py, cur_long, lat);
        float left_long = Li[LF_LENGTH] - cur_long;
        bool ffound = false, bfound = false;
        const int n = nv.count();
        for (int k = 0; k < n; k++) {
            bool ok = k < 64 ? ((valid_lo >> k) & 1ull) : ((valid_hi >> (k - 64)) & 1ull);
            if (!ok) continue;
            float ox, oy, ovx, ovy; int olane; bool ped;
            nv.get(k, ox, oy, ovx, ovy, olane, ped);
            if (olane < 0) continue;
            const float* Lo = m.lane_f + olane * LANE_F;
            float lon, lt;
            if (olane == lanes[i]) {
                lane_local(Li, ox, oy, lon, lt);
                lon -= cur_long;
                if (out.fdist[i] > lon && lon > 0.0f) { out.fdist[i] = lon; out.fobj[i] = k; ffound = true; }
                if (lon < 0.0f && fabsf(lon) < out.bdist[i]) { out.bdist[i] = fabsf(lon); out.bobj[i] = k; bfound = true; }
            } else if (!ffound && lane_is_previous_of(Li, Lo)) {
                lane_local(Lo, ox, oy, lon, lt);
                lon += left_long;
                if (out.fdist[i] > lon && lon > 0.0f) { out.fdist[i] = lon; out.fobj[i] = k; }
            } else if (!bfound && lane_is_previous_of(Lo, Li)) {
                lane_local(Lo, ox, oy, lon, lt);
                lon = Lo[LF_LENGTH] - lon + cur_long;
                if (out.bdist[i] > lon) { out.bdist[i] = lon; out.bobj[i] = k; }
            }
        }
    }
}

// find_front_back with the expensive part - the lane coordinates of every surrounding object and whether its lane is one
// of the three lanes, continues or precedes them - spread over the T lanes of a sub-warp team (the search is 60 % of
// k_pre).  The object loop of the reference is ORDER dependent (an object on a successor lane only counts until one on
// the lane itself was found), so the team only fills a table (`scratch`, one float2 per neighbour index), and every lane
// then runs the ordered folds over it redundantly - all lanes of the team end with the same, exact result.
// An object's longitude is taken on ITS OWN lane whichever of the three lanes is being searched (on the lane itself the
// two coincide), so the table is filled ONCE for the three searches: one lane-row fetch and one lane_local per object.
//   scratch[ord] = (codes, longitude on its own lane) of the ord-th valid neighbour; codes = 3 bits per searched lane i (<< 3 i): bit 0 the
//   object is on lane i, bit 1 lane i is the previous of its lane, bit 2 its lane is the previous of lane i
__device__ void find_front_back_team(const MapView& m, const NbrView& nv, unsigned long long valid_lo, unsigned long long valid_hi,
                                     int lane, float px, float py, float max_d, bool use_ref, int ref_first, int ref_n,
                                     FrontBack& out, int sub, int T, unsigned team_mask, float2* scratch) {
    int lanes[3] = {-1, lane, -1};
    if (use_ref) {
        int idx = m.lane_i[lane * LANE_I + LI_IDX];
        if (idx > 0) lanes[0] = ref_first + idx - 1;
        if (idx + 1 < ref_n) lanes[2] = ref_first + idx + 1;
    }
    float cur_long[3], left_long[3], sx[3], sy[3], ex[3], ey[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        out.fobj[i] = out.bobj[i] = OBJ_NONE;
        out.exists[i] = lanes[i] >= 0;
        out.fdist[i] = out.bdist[i] = max_d;
        cur_long[i] = left_long[i] = sx[i] = sy[i] = ex[i] = ey[i] = 0.0f;
        if (lanes[i] < 0) continue;
        const float* Li = m.lane_f + lanes[i] * LANE_F;
        float lat;
        lane_local(Li, px, py, cur_long[i], lat);
        left_long[i] = Li[LF_LENGTH] - cur_long[i];
        sx[i] = Li[LF_SX]; sy[i] = Li[LF_SY]; ex[i] = Li[LF_EX]; ey[i] = Li[LF_EY];
    }
    // the table: the ord-th valid neighbour belongs to lane ord % T of the team
    int ord = 0;
    for (int half = 0; half < 2; half++)
        for (unsigned long long mk = half ? valid_hi : valid_lo; mk; mk &= mk - 1, ord++) {
            if ((ord & (T - 1)) != sub) continue;
            const int k = __ffsll((long long)mk) - 1 + 64 * half;
            float ox, oy, ovx, ovy; int olane; bool ped;
            nv.get(k, ox, oy, ovx, ovy, olane, ped);
            float2 rec = make_float2(0.0f, 0.0f);
            if (olane >= 0) {
                const float* Lo = m.lane_f + olane * LANE_F;
                const float osx = Lo[LF_SX], osy = Lo[LF_SY], oex = Lo[LF_EX], oey = Lo[LF_EY];
                int codes = 0;
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    if (lanes[i] < 0) continue;
                    int c;
                    if (olane == lanes[i]) c = 1;
                    else {   // lane_is_previous_of(Li, Lo), lane_is_previous_of(Lo, Li)
                        const float ax = ex[i] - osx, ay = ey[i] - osy, bx = oex - sx[i], by = oey - sy[i];
                        c = (sqrtf(ax * ax + ay * ay) < 0.1f ? 2 : 0) | (sqrtf(bx * bx + by * by) < 0.1f ? 4 : 0);
                    }
                    codes |= c << (3 * i);
                }
                if (codes) { float lon, lt; lane_local(Lo, ox, oy, lon, lt); rec.y = lon; }
                rec.x = (float)codes;
            }
            scratch[ord] = rec;
        }
    __syncwarp(team_mask);
    // the ordered folds (policy/idm_policy.py:100-131), identical on every lane
#pragma unroll
    for (int i = 0; i < 3; i++) {
        if (lanes[i] < 0) continue;
        bool ffound = false, bfound = false;
        int o2 = 0;
        for (int half = 0; half < 2; half++)
            for (unsigned long long mk = half ? valid_hi : valid_lo; mk; mk &= mk - 1, o2++) {
                const int k = __ffsll((long long)mk) - 1 + 64 * half;
                const float2 rec = scratch[o2];
                const int code = ((int)rec.x >> (3 * i)) & 7;
                float lon = rec.y;
                if (code & 1) {
                    lon -= cur_long[i];
                    if (out.fdist[i] > lon && lon > 0.0f) { out.fdist[i] = lon; out.fobj[i] = k; ffound = true; }
                    if (lon < 0.0f && fabsf(lon) < out.bdist[i]) { out.bdist[i] = fabsf(lon); out.bobj[i] = k; bfound = true; }
                } else if (!ffound && (code & 2)) {
                    lon += left_long[i];
                    if (out.fdist[i] > lon && lon > 0.0f) { out.fdist[i] = lon; out.fobj[i] = k; }
                } else if (!bfound && (code & 4)) {
                    float ox, oy, ovx, ovy; int olane; bool ped;
                    nv.get(k, ox, oy, ovx, ovy, olane, ped);
                    lon = m.lane_f[olane * LANE_F + LF_LENGTH] - lon + cur_long[i];
                    if (out.bdist[i] > lon) { out.bdist[i] = lon; out.bobj[i] = k; }
                }
            }
    }
    __syncwarp(team_mask);  // the table is rewritten by the team's next vehicle
}

__device__ __forceinline__ float pid(float kp, float ki, float kd, float& p_err, float& i_err, float err) {
    i_err += err;
    float d = err - p_err;
    p_err = err;
    return -kp * p_err - ki * i_err - kd * d;
}
// build-defined replacement of np_random.randint(0, n) (policy/idm_policy.py:288)
__device__ __forceinline__ int hash_randint(uint32_t slot, uint32_t counter, int n) {
    uint32_t x = slot * 0x9E3779B9u + counter * 0x85EBCA6Bu + 0x165667B1u;
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return (int)(x % (uint32_t)n);
}

// word j of the k-th row of an env's random tape (include/md_layout.h: env_tape); once the tape has wrapped it is
// perturbed with a counter hash so that a long run does not repeat itself
__device__ __forceinline__ uint32_t tape_draw(const MdConfig& cfg, const int* __restrict__ tape, int env, uint32_t ctr, int j) {
    const uint32_t L = (uint32_t)cfg.tape_len;
    uint32_t v = (uint32_t)tape[((size_t)env * L + ctr % L) * TAPE_W + j];
    const uint32_t lap = ctr / L;
    if (lap) {
        uint32_t x = (uint32_t)(env + cfg.env_base) * 0x9E3779B9u + (lap * 4u + (uint32_t)j) * 0x85EBCA6Bu + 0x165667B1u;
        x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
        v += x;
    }
    return v;
}
// word 1 as a uniform number in [0, 1): the stored float on the first lap, 24 hashed bits afterwards
__device__ __forceinline__ float tape_frac(const MdConfig& cfg, const int* __restrict__ tape, int env, uint32_t ctr) {
    const uint32_t v = tape_draw(cfg, tape, env, ctr, 1);
    if (ctr / (uint32_t)cfg.tape_len == 0) return __uint_as_float(v);
    return (float)(v >> 8) * (1.0f / 16777216.0f);
}
// yaw-only quaternion (w, z) of a body whose +Y axis (the nose) points along the lane at `lon`
__device__ __forceinline__ void yaw_quat_for_lane(const float* L, float lon, float& qw, float& qz) {
    if (L[LF_TYPE] == 0.0f) {
        const float c = L[LF_P0 + 5], s = -L[LF_P0 + 4];  // cos(yaw) = dy, sin(yaw) = -dx
        const float cw = sqrtf(fmaxf(0.5f * (1.0f + c), 0.0f)), sw = sqrtf(fmaxf(0.5f * (1.0f - c), 0.0f));
        qw = cw; qz = s < 0.0f ? -sw : sw;
    } else {
        const float yaw = lane_heading_at(L, lon) - MD_PI / 2.0f;
        qw = md_cosf(0.5f * yaw); qz = md_sinf(0.5f * yaw);
    }
}

// IDMPolicy.act (policy/idm_policy.py:235-402) for one traffic vehicle; S/I/D are this thread's register copies
// With T > 1 the T lanes of a sub-warp team run this together for the same vehicle (same inputs, same control flow, same
// results); only the front / back search actually splits its work over them.
__device__ void idm_act(const MdConfig& cfg, const MapView& m, const NbrView& nv, int g, const float* S, int* I, float* D,
                        const int* __restrict__ rroad, float& out_a0, float& out_a1, int sub = 0, int T = 1,
                        unsigned team_mask = 0u, float2* scratch = nullptr) {
    float px = S[VS_POS], py = S[VS_POS + 1];
    M3 Rg = quat_to_m3(S[VS_QUAT], S[VS_QUAT + 1], S[VS_QUAT + 2], S[VS_QUAT + 3]);
    float hx, hy;
    heading_vec(Rg, hx, hy);
    float speed_kmh = sqrtf(S[VS_VEL] * S[VS_VEL] + S[VS_VEL + 1] * S[VS_VEL + 1]) * 3.6f;
    int c0 = I[VI_CKPT0], c1 = I[VI_CKPT1];
    int cur_road = rroad[c0];
    int next_road = c1 != c0 ? rroad[c1] : -1;
    int cur_first = m.road_i[cur_road * ROAD_I + RI_FIRST], cur_n = m.road_i[cur_road * ROAD_I + RI_N];
    int veh_lane = I[VI_LANE];
#define IN_CUR(l) ((l) >= cur_first && (l) < cur_first + cur_n)
    bool success;
    int rt = I[VI_ROUTING_LANE];
    if (rt < 0) {
        rt = veh_lane;
        success = IN_CUR(rt);
    } else if (!IN_CUR(rt)) {
        success = false;
        for (int l = cur_first; l < cur_first + cur_n; l++) {
            bool conn = lane_is_previous_of(m.lane_f + rt * LANE_F, m.lane_f + l * LANE_F) ||
                        find_road(m, m.lane_i[rt * LANE_I + LI_TO], m.lane_i[l * LANE_I + LI_TO]) >= 0;
            if (conn) { rt = l; success = true; break; }
        }
    } else if (IN_CUR(veh_lane) && rt != veh_lane) {
        rt = veh_lane;
        uint32_t ctr = (uint32_t)D[VD_RNG];
        D[VD_TIMER] = (float)hash_randint((uint32_t)g, ctr, 25);
        D[VD_RNG] = (float)(ctr + 1);
        success = true;
    } else success = true;
    I[VI_ROUTING_LANE] = rt;
    if (threadIdx.x == 0) clk_mark(0, 4);

    // Lidar.get_surrounding_objects(r = 50) (component/sensors/lidar.py:170-186): membership bitmask
    unsigned long long vlo = 0ull, vhi = 0ull;
    bool has_ped = false;
    const int n = nv.count();
    for (int k = sub; k < n; k += T) {   // the membership tests are strided over the team's lanes and ORed below
        if (!nv.valid(k)) continue;
        if (k < 64) vlo |= 1ull << k; else vhi |= 1ull << (k - 64);
        if (k >= nv.S) {
            const float* Ob = nv.obj + (k - nv.S) * OBJ_F;
            if (Ob[OB_KIND] == 3.0f || Ob[OB_LANE] < 0.0f) has_ped = true;  // no `.lane` -> except path (:254-259)
        }
    }
    if (T > 1) {
        int ped = has_ped ? 1 : 0;
        for (int off = T >> 1; off > 0; off >>= 1) {
            vlo |= __shfl_xor_sync(team_mask, vlo, off);
            if (n > 64) vhi |= __shfl_xor_sync(team_mask, vhi, off);
            ped |= __shfl_xor_sync(team_mask, ped, off);
        }
        has_ped = ped != 0;
    }
    if (threadIdx.x == 0) clk_mark(0, 5);
    // the team's table holds PRE_NB_CAP neighbours; with more in range every lane runs the serial search (same result)
    const bool team_ok = T > 1 && __popcll(vlo) + __popcll(vhi) <= pre_nb_cap(nv.S, nv.O);
    int front = OBJ_NONE;
    float front_dist = 0.0f;
    int steer_lane = rt;
    FrontBack fb;
    if (has_ped) {
        front = OBJ_NONE; front_dist = 5.0f; steer_lane = rt;
    } else if (success && cfg.enable_idm_lane_change) {
        if (team_ok) find_front_back_team(m, nv, vlo, vhi, rt, px, py, 30.0f, true, cur_first, cur_n, fb, sub, T, team_mask, scratch);
        else find_front_back(m, nv, vlo, vhi, rt, px, py, 30.0f, true, cur_first, cur_n, fb);
        int next_n = next_road >= 0 ? m.road_i[next_road * ROAD_I + RI_N] : 0;
        int next_first = next_road >= 0 ? m.road_i[next_road * ROAD_I + RI_FIRST] : -1;
        int diff = next_road >= 0 ? cur_n - next_n : 0;
        int lo = 0, hi = cur_n - 1;
        int rt_idx = m.lane_i[rt * LANE_I + LI_IDX];
        bool decided = false;
        front = fb.fobj[1]; front_dist = fb.fdist[1]; steer_lane = rt;
        if (diff > 0) {
            if (lane_is_previous_of(m.lane_f + cur_first * LANE_F, m.lane_f + next_first * LANE_F)) { lo = 0; hi = next_n - 1; }
            else { lo = diff; hi = cur_n - 1; }
            if (rt_idx < lo || rt_idx > hi) {
                decided = true;
                if (rt_idx > hi) {
                    if (fb.bdist[0] < 15.0f || fb.fdist[0] < 5.0f) { D[VD_TARGET_SPEED] = 5.0f; }
                    else { D[VD_TARGET_SPEED] = 30.0f; front = fb.fobj[0]; front_dist = fb.fdist[0]; steer_lane = cur_first + rt_idx - 1; }
                } else {
                    if (fb.bdist[2] < 15.0f || fb.fdist[2] < 5.0f) { D[VD_TARGET_SPEED] = 5.0f; }
                    else { D[VD_TARGET_SPEED] = 30.0f; front = fb.fobj[2]; front_dist = fb.fdist[2]; steer_lane = cur_first + rt_idx + 1; }
                }
            }
        }
        if (!decided) {
            bool overtake = false;
            float fsp = 0.0f;
            if (fb.fobj[1] != OBJ_NONE) {
                float ox, oy, ovx, ovy; int ol; bool pd;
                nv.get(fb.fobj[1], ox, oy, ovx, ovy, ol, pd);
                fsp = sqrtf(ovx * ovx + ovy * ovy) * 3.6f;
            }
            if (fabsf(speed_kmh - 30.0f) > 3.0f && fb.fobj[1] != OBJ_NONE && fabsf(fsp - 30.0f) > 3.0f && D[VD_TIMER] > 50.0f) {
                bool has_r = false, has_l = false;
                float rs = 0.0f, ls = 0.0f;
                float ox, oy, ovx, ovy; int ol; bool pd;
                if (fb.fobj[2] != OBJ_NONE) { has_r = true; nv.get(fb.fobj[2], ox, oy, ovx, ovy, ol, pd); rs = sqrtf(ovx * ovx + ovy * ovy) * 3.6f; }
                else if (fb.exists[2] && fb.fdist[2] > 15.0f && fb.bdist[2] > 15.0f) { has_r = true; rs = 100.0f; }
                if (fb.fobj[0] != OBJ_NONE) { has_l = true; nv.get(fb.fobj[0], ox, oy, ovx, ovy, ol, pd); ls = sqrtf(ovx * ovx + ovy * ovy) * 3.6f; }
                else if (fb.exists[0] && fb.fdist[0] > 15.0f && fb.bdist[0] > 15.0f) { has_l = true; ls = 100.0f; }
                if (has_l && ls - fsp > 10.0f) {
                    int e = rt_idx - 1;
                    if (e >= lo && e <= hi) { front = fb.fobj[0]; front_dist = fb.fdist[0]; steer_lane = cur_first + e; overtake = true; }
                }
                if (!overtake && has_r && rs - fsp > 10.0f) {
                    int e = rt_idx + 1;
                    if (e >= lo && e <= hi) { front = fb.fobj[2]; front_dist = fb.fdist[2]; steer_lane = cur_first + e; overtake = true; }
                }
            }
            if (!overtake) {
                D[VD_TARGET_SPEED] = 30.0f;
                D[VD_TIMER] += 1.0f;
                front = fb.fobj[1]; front_dist = fb.fdist[1]; steer_lane = rt;
            }
        }
    } else {
        if (team_ok) find_front_back_team(m, nv, vlo, vhi, rt, px, py, 30.0f, false, 0, 0, fb, sub, T, team_mask, scratch);
        else find_front_back(m, nv, vlo, vhi, rt, px, py, 30.0f, false, 0, 0, fb);
        front = fb.fobj[1]; front_dist = fb.fdist[1]; steer_lane = rt;
    }
    if (threadIdx.x == 0) clk_mark(0, 6);
    // steering_control (:293-301)
    const float* Ls = m.lane_f + steer_lane * LANE_F;
    float lon, lat;
    lane_local(Ls, px, py, lon, lat);
    float lane_heading = lane_heading_at(Ls, lon + 1.0f);
    float v_heading = md_atan2f(hy, hx);
    float steering = pid(1.7f, 0.01f, 3.5f, D[VD_H_PERR], D[VD_H_IERR], -wrap_to_pi(lane_heading - v_heading));
    steering += pid(0.3f, 0.002f, 0.05f, D[VD_L_PERR], D[VD_L_IERR], -lat);
    // acceleration (:303-320), km/h units
    float target = D[VD_TARGET_SPEED];
    float ratio = fmaxf(speed_kmh, 0.0f) / target;
    float r2 = ratio * ratio, r4 = r2 * r2, r8 = r4 * r4;
    float acc = 1.0f - r8 * r2;
    if (front != OBJ_NONE) {
        float ox, oy, ovx, ovy; int ol; bool pd;
        nv.get(front, ox, oy, ovx, ovy, ol, pd);
        float d = front_dist;
        float dvx = S[VS_VEL] * 3.6f - ovx * 3.6f, dvy = S[VS_VEL + 1] * 3.6f - ovy * 3.6f;
        float dv = dvx * hx + dvy * hy;
        float d_star = 10.0f + speed_kmh * 1.5f + speed_kmh * dv / (2.0f * sqrtf(5.0f));
        float nz = fabsf(d) > 1e-2f ? d : (d > 0.0f ? 1e-2f : -1e-2f);
        float sd = d_star / nz;
        acc -= sd * sd;
    }
    out_a0 = steering;
    out_a1 = acc;
    if (threadIdx.x == 0) clk_mark(0, 7);
#undef IN_CUR
}

// What the candidate scan of update_localization leaves behind for one vehicle: whether any lane hull contains the centre,
// and the nearest lane (lane.distance) overall / among the current reference lanes / among the next road's lanes.  A slot
// without a lane has best = -1.  The scan of one vehicle can be spread over the lanes of a sub-warp team (k_post phase
// 2a): a (distance, lane id) lexicographic minimum reproduces the serial scan, which walks ascending lane ids with `<`.
struct __align__(16) LocScan {
    float d_any, lon_any, lat_any; int best_any;
    float d_cur, lon_cur, lat_cur; int best_cur;
    float d_next, lon_next, lat_next; int best_next;
    int on_lane, static_flags, contact_flags, pad;
};
struct LocCtx { float px, py, hx, hy; int cur_first, cur_n, next_road, nx_first, nx_n; };
__device__ __forceinline__ void loc_scan_init(LocScan& s) {
    s.d_any = s.d_cur = s.d_next = 1e30f;
    s.best_any = s.best_cur = s.best_next = -1;
    s.lon_any = s.lat_any = s.lon_cur = s.lat_cur = s.lon_next = s.lat_next = 0.0f;
    s.on_lane = 0; s.static_flags = 0; s.contact_flags = 0; s.pad = 0;
}
// one candidate lane of the grid cell under the vehicle (ray_localization, utils/pg/utils.py:151-203, answered by AABB ->
// point-in-convex-hull; lane choice of node_network_navigation.py:219-241)
// ... after the AABB test: `hull_off` / `hull_n` = the lane's convex hull in m.hull
__device__ __forceinline__ void loc_candidate_in_bb(const MapView& m, const LocCtx& c, int l, int hull_off, int hull_n, LocScan& s) {
    const float px = c.px, py = c.py;
    const float* Ll = m.lane_f + l * LANE_F;
    if (Ll[LF_TYPE] != 0.0f) {  // hull_shortcut's radial rejection, before paying for the arc coordinates (atan2)
        const float ro = Ll[LF_P0 + 2] + 0.5f * Ll[LF_WIDTH];
        const float ddx = px - Ll[LF_P0 + 0], ddy = py - Ll[LF_P0 + 1];
        if (ddx * ddx + ddy * ddy > ro * ro + 1.0f + 0.5f * ro) return;
    }
    float lon, lat;
    lane_local(Ll, px, py, lon, lat);
    const int sc = hull_shortcut(Ll, px, py, lon, lat);
    if (sc < 0) return;
    if (sc == 0 && !point_in_hull(Ll, m.hull + 2 * hull_off, hull_n, px, py)) return;
    s.on_lane = 1;
    float lh = lane_heading_at(Ll, lon);
    float cosang = md_cosf(lh) * c.hx + md_sinf(lh) * c.hy;
    if (!(cosang > 0.0f)) return;
    // lane.distance (abs_lane.py:76-82) from the local coordinates already at hand (same arithmetic as lane_distance)
    const float over = lon - Ll[LF_LENGTH], under = 0.0f - lon;
    const float dist = fabsf(lat) + (over > 0.0f ? over : 0.0f) + (under > 0.0f ? under : 0.0f);
    if (dist < s.d_any) { s.d_any = dist; s.best_any = l; s.lon_any = lon; s.lat_any = lat; }
    if (l >= c.cur_first && l < c.cur_first + c.cur_n && dist < s.d_cur) { s.d_cur = dist; s.best_cur = l; s.lon_cur = lon; s.lat_cur = lat; }
    if (c.next_road >= 0 && l >= c.nx_first && l < c.nx_first + c.nx_n && dist < s.d_next) { s.d_next = dist; s.best_next = l; s.lon_next = lon; s.lat_next = lat; }
}
__device__ __forceinline__ void loc_candidate(const MapView& m, const LocCtx& c, int l, LocScan& s) {
    const float4 bb = __ldg(reinterpret_cast<const float4*>(m.lane_bb) + l);
    if (c.px < bb.x || c.py < bb.y || c.px > bb.z || c.py > bb.w) return;
    loc_candidate_in_bb(m, c, l, m.lane_i[l * LANE_I + LI_HULL_OFF], m.lane_i[l * LANE_I + LI_HULL_N], s);
}
__device__ __forceinline__ LocCtx loc_ctx(const MapView& m, const float* S, const int* I, const int* __restrict__ rroad) {
    LocCtx c;
    c.px = S[VS_POS]; c.py = S[VS_POS + 1];
    M3 R = quat_to_m3(S[VS_QUAT], S[VS_QUAT + 1], S[VS_QUAT + 2], S[VS_QUAT + 3]);
    heading_vec(R, c.hx, c.hy);
    const int c0 = I[VI_CKPT0], c1 = I[VI_CKPT1];
    const int cur_road = rroad[c0];
    c.next_road = c1 != c0 ? rroad[c1] : -1;
    c.cur_first = m.road_i[cur_road * ROAD_I + RI_FIRST]; c.cur_n = m.road_i[cur_road * ROAD_I + RI_N];
    c.nx_first = c.next_road >= 0 ? m.road_i[c.next_road * ROAD_I + RI_FIRST] : -1;
    c.nx_n = c.next_road >= 0 ? m.road_i[c.next_road * ROAD_I + RI_N] : 0;
    return c;
}
// the candidate lanes: those binned into the grid cell under the vehicle (ascending lane id, like a full scan)
__device__ __forceinline__ void loc_cell(const MapView& m, float px, float py, int& k0, int& k1) {
    k0 = 0; k1 = 0;
    int cx = (int)floorf((px - m.gx0) / m.cell), cy = (int)floorf((py - m.gy0) / m.cell);
    if (cx >= 0 && cy >= 0 && cx < m.nx && cy < m.ny) { k0 = m.lgs[cy * m.nx + cx]; k1 = m.lgs[cy * m.nx + cx + 1]; }
}

// NodeNetworkNavigation.update_localization (component/navigation_module/node_network_navigation.py:130-304) with
// ray_localization (utils/pg/utils.py:151-203) answered by AABB -> point-in-convex-hull over the map's lane table.
// `pre` = the candidate scan if a team already ran it (k_post phase 2a), else it runs here.
__device__ void localise(const MapView& m, const float* S, int* I, const int* __restrict__ route,
                         const int* __restrict__ rroad, float* navi, const LocScan* pre) {
    float px = S[VS_POS], py = S[VS_POS + 1];
    M3 R = quat_to_m3(S[VS_QUAT], S[VS_QUAT + 1], S[VS_QUAT + 2], S[VS_QUAT + 3]);
    int c0 = I[VI_CKPT0], c1 = I[VI_CKPT1], n_ck = I[VI_ROUTE_LEN];
    int cur_road = rroad[c0];
    int next_road = c1 != c0 ? rroad[c1] : -1;
    int cur_first = m.road_i[cur_road * ROAD_I + RI_FIRST], cur_n = m.road_i[cur_road * ROAD_I + RI_N];
    int nx_first = next_road >= 0 ? m.road_i[next_road * ROAD_I + RI_FIRST] : -1;
    LocScan sc;
    if (pre != nullptr) sc = *pre;
    else {
        const LocCtx c = loc_ctx(m, S, I, rroad);
        loc_scan_init(sc);
        int k0, k1;
        loc_cell(m, px, py, k0, k1);
        for (int kk = k0; kk < k1; kk++) loc_candidate(m, c, m.lgi[kk], sc);
    }
    const bool on_lane = sc.on_lane != 0;
    const int best_any = sc.best_any, best_cur = sc.best_cur, best_next = sc.best_next;
    const float lon_any = sc.lon_any, lat_any = sc.lat_any, lon_cur = sc.lon_cur, lat_cur = sc.lat_cur,
                lon_next = sc.lon_next, lat_next = sc.lat_next;
    int lane;
    float lon, lat;
    if (best_cur >= 0) { lane = best_cur; lon = lon_cur; lat = lat_cur; }
    else if (next_road >= 0 && best_next >= 0) { lane = best_next; lon = lon_next; lat = lat_next; }
    else { lane = best_any; lon = lon_any; lat = lat_any; }
    if (on_lane) I[VI_FLAGS] |= FL_ON_LANE; else I[VI_FLAGS] &= ~FL_ON_LANE;
    const bool kept = lane < 0;
    if (kept) lane = I[VI_LANE];
    I[VI_LANE] = lane;
    const float* Lc = m.lane_f + lane * LANE_F;
    if (kept) lane_local(Lc, px, py, lon, lat);
    if (c0 != c1) {  // _update_target_checkpoints (:181-201)
        int start = m.lane_i[lane * LANE_I + LI_FROM];
        bool in_tail = false;
        int idx = -1;
        for (int k = c1; k < n_ck; k++) if (route[k] == start) { in_tail = true; break; }
        if (in_tail && lon < 5.0f) {
            for (int k = c1; k < n_ck - 1; k++) if (route[k] == start) { idx = k; break; }
            if (idx >= 0) {
                c0 = idx;
                c1 = (idx + 1 == n_ck - 1) ? idx : idx + 1;
                I[VI_CKPT0] = c0; I[VI_CKPT1] = c1;
                cur_road = rroad[c0];
                next_road = c1 != c0 ? rroad[c1] : -1;
                cur_first = m.road_i[cur_road * ROAD_I + RI_FIRST]; cur_n = m.road_i[cur_road * ROAD_I + RI_N];
                nx_first = next_road >= 0 ? m.road_i[next_road * ROAD_I + RI_FIRST] : -1;
            }
        }
    }
    float lane_w = Lc[LF_WIDTH];
    float later_middle = ((float)cur_n / 2.0f - 0.5f) * lane_w;
#pragma unroll
    for (int k = 0; k < 2; k++) {  // _get_info_for_checkpoint (:243-292)
        int ref = k == 0 ? cur_first : (next_road >= 0 ? nx_first : cur_first);
        const float* Lr = m.lane_f + ref * LANE_F;
        float ckx, cky;
        lane_position(Lr, Lr[LF_LENGTH], later_middle, ckx, cky);
        float dx = ckx - px, dy = cky - py;
        float dn = sqrtf(dx * dx + dy * dy);
        if (dn > 50.0f) { dx = dx / dn * 50.0f; dy = dy / dn * 50.0f; }
        float in_heading = dx * R.m[0][1] + dy * R.m[1][1];      // base_vehicle.py:983-988
        float in_rhs = -(dx * R.m[0][0] + dy * R.m[1][0]);
        float* o = navi + 5 * k;
        o[0] = clipf((in_heading / 50.0f + 1.0f) / 2.0f, 0.0f, 1.0f);
        o[1] = clipf((in_rhs / 50.0f + 1.0f) / 2.0f, 0.0f, 1.0f);
        float bend = 0.0f, dirv = 0.0f, angle = 0.0f;
        if (Lr[LF_TYPE] != 0.0f) {
            bend = Lr[LF_P0 + 2] / (60.0f + (float)cur_n * lane_w);
            dirv = -Lr[LF_P0 + 5];
            angle = Lr[LF_P0 + 6];
        }
        o[2] = clipf(bend, 0.0f, 1.0f);
        o[3] = clipf((dirv + 1.0f) / 2.0f, 0.0f, 1.0f);
        o[4] = clipf((angle * (180.0f / MD_PI) / 135.0f + 1.0f) / 2.0f, 0.0f, 1.0f);
    }
}

// BaseVehicle._state_check, static world part + sidewalk sweep (component/vehicle/base_vehicle.py:700-792),
// broad phase = the map's uniform grid.  Items of a cell are handled four at a time: ids first, then the first float4
// of every row (centre + extent), a bounding-circle prune, and only then the exact SAT - so that the global-load
// latencies of a batch overlap instead of chaining.  Flags are ORed, so the order of evaluation is irrelevant.
__device__ void state_check_static(const MapView& m, const Rect& r, int& flags) {
    const float rad = sqrtf(r.hu * r.hu + r.hv * r.hv);
    int x0 = (int)floorf((r.cx - rad - m.gx0) / m.cell), x1 = (int)floorf((r.cx + rad - m.gx0) / m.cell);
    int y0 = (int)floorf((r.cy - rad - m.gy0) / m.cell), y1 = (int)floorf((r.cy + rad - m.gy0) / m.cell);
    x0 = max(x0, 0); y0 = max(y0, 0); x1 = min(x1, m.nx - 1); y1 = min(y1, m.ny - 1);
    const float4* line4 = reinterpret_cast<const float4*>(m.lines);
    const float4* quad4 = reinterpret_cast<const float4*>(m.quads);
    for (int cy = y0; cy <= y1; cy++)
        for (int cx = x0; cx <= x1; cx++) {
            const int c = cy * m.nx + cx;
            const int k1 = m.gs[c + 1];
            for (int k = m.gs[c]; k < k1; k += 4) {
                int id[4];
                float4 h[4];
#pragma unroll
                for (int j = 0; j < 4; j++) id[j] = k + j < k1 ? __ldg(m.gi + k + j) : -1;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (id[j] < 0) continue;
                    h[j] = id[j] < m.n_lines ? __ldg(line4 + 2 * id[j]) : __ldg(quad4 + 2 * (id[j] - m.n_lines));
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int it = id[j];
                    if (it < 0) continue;
                    if (it < m.n_lines) {
                        const int kind = (int)h[j].w;
                        const int bit = kind == 0 ? FL_ON_WHITE : (kind == 1 ? FL_ON_YELLOW : FL_ON_BROKEN);
                        if (flags & bit) continue;
                        // conservative prune: centre distance beyond the two bounding radii (+1 cm)
                        const float dx = h[j].x - r.cx, dy = h[j].y - r.cy, rr = rad + h[j].z + LINE_HALF_W + 0.01f;
                        if (dx * dx + dy * dy > rr * rr) continue;
                        const float4 u = __ldg(line4 + 2 * it + 1);
                        Rect lr;
                        lr.cx = h[j].x; lr.cy = h[j].y; lr.ux = u.x; lr.uy = u.y; lr.hu = h[j].z; lr.hv = LINE_HALF_W;
                        if (rect_rect(r, lr)) flags |= bit;
                    } else {
                        if (flags & FL_CRASH_SIDEWALK) continue;
                        const float4 q1 = __ldg(quad4 + 2 * (it - m.n_lines) + 1);
                        float q[8] = {h[j].x, h[j].y, h[j].z, h[j].w, q1.x, q1.y, q1.z, q1.w};
                        if (rect_quad(r, q)) flags |= FL_CRASH_SIDEWALK;
                    }
                }
            }
        }
}

// team version of state_check_static: the items of the cells under the bounding circle are strided over the T lanes of a
// sub-warp team; every lane returns the flags of its share (the caller ORs them: the order of evaluation is irrelevant)
__device__ __forceinline__ int state_check_static_team(const MapView& m, const Rect& r, int sub, int T) {
    int flags = 0;
    const float rad = sqrtf(r.hu * r.hu + r.hv * r.hv);
    int x0 = (int)floorf((r.cx - rad - m.gx0) / m.cell), x1 = (int)floorf((r.cx + rad - m.gx0) / m.cell);
    int y0 = (int)floorf((r.cy - rad - m.gy0) / m.cell), y1 = (int)floorf((r.cy + rad - m.gy0) / m.cell);
    x0 = max(x0, 0); y0 = max(y0, 0); x1 = min(x1, m.nx - 1); y1 = min(y1, m.ny - 1);
    const float4* line4 = reinterpret_cast<const float4*>(m.lines);
    const float4* quad4 = reinterpret_cast<const float4*>(m.quads);
    for (int cy = y0; cy <= y1; cy++)
        for (int cx = x0; cx <= x1; cx++) {
            const int c = cy * m.nx + cx;
            const int k1 = m.gs[c + 1];
            for (int k = m.gs[c] + sub; k < k1; k += T) {
                const int it = __ldg(m.gi + k);
                if (it < m.n_lines) {
                    const float4 h = __ldg(line4 + 2 * it);
                    const int kind = (int)h.w;
                    const int bit = kind == 0 ? FL_ON_WHITE : (kind == 1 ? FL_ON_YELLOW : FL_ON_BROKEN);
                    if (flags & bit) continue;
                    const float dx = h.x - r.cx, dy = h.y - r.cy, rr = rad + h.z + LINE_HALF_W + 0.01f;
                    if (dx * dx + dy * dy > rr * rr) continue;
                    const float4 u = __ldg(line4 + 2 * it + 1);
                    Rect lr;
                    lr.cx = h.x; lr.cy = h.y; lr.ux = u.x; lr.uy = u.y; lr.hu = h.z; lr.hv = LINE_HALF_W;
                    if (rect_rect(r, lr)) flags |= bit;
                } else {
                    if (flags & FL_CRASH_SIDEWALK) continue;
                    const float4 q0 = __ldg(quad4 + 2 * (it - m.n_lines)), q1 = __ldg(quad4 + 2 * (it - m.n_lines) + 1);
                    float q[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
                    if (rect_quad(r, q)) flags |= FL_CRASH_SIDEWALK;
                }
            }
        }
    return flags;
}

// contact pairs of vehicle `slot` against the env's other bodies (engine/core/collision_callback.py:5-42 when
// latch, component/vehicle/base_vehicle.py:735-742 otherwise)
__device__ __forceinline__ int dynamic_contacts(const Nb* nb, float* obj, int S, int O, int slot, const Rect& r, bool latch,
                                                int* obj_first, bool claim_pass) {
    int flags = 0;
    for (int k = 0; k < S; k++) {
        if (k == slot || !nb[k].alive) continue;
        if (rect_rect(r, nb[k].r)) flags |= FL_CRASH_VEHICLE;
    }
    for (int k = 0; k < O; k++) {
        const float* Ob = obj + k * OBJ_F;
        if (Ob[OB_KIND] < 0.0f) continue;
        bool hit;
        if (OB_IS_BOX(Ob[OB_KIND])) { Rect ro = object_rect(Ob); hit = rect_rect(r, ro); }
        else hit = rect_circle(r, Ob[OB_X], Ob[OB_Y], Ob[OB_A]);
        if (!hit) continue;
        if (Ob[OB_KIND] == 3.0f) flags |= FL_CRASH_HUMAN;
        else if (Ob[OB_KIND] == 4.0f) flags |= FL_CRASH_BUILDING;   // collision_callback.py:39-41, base_vehicle.py:737-738
        else if (latch) {
            if (Ob[OB_CRASHED] == 0.0f) {
                if (claim_pass) atomicMin(&obj_first[k], slot);          // COST_ONCE: lowest slot takes the flag
                else if (obj_first[k] == slot) flags |= FL_CRASH_OBJECT;
            }
        } else flags |= FL_CRASH_OBJECT;
    }
    return flags;
}

#define FL_TOUCH 0x40000000  // contact_pass only, never stored: a vehicle or solid obstacle overlaps (-> contact response)
// the same contact rules over a candidate bit set (k_dyn's per-step broad phase): bit k < S = vehicle slot k, else object k - S
// `touch` (optional, the vehicle's own two words in shared memory): bit k is set for every body the vehicle overlaps - the
// pairs the contact-added callback reports (engine/core/collision_callback.py:5-42), exported by md_get_contacts.
__device__ __forceinline__ int contact_pass(const Nb* nb, float* obj, int S, int slot, const Rect& r, unsigned long long lo,
                                            unsigned long long hi, int* obj_first, bool claim_pass, bool objects_only,
                                            unsigned long long* touch = nullptr) {
    int flags = 0;
    for (int half = 0; half < 2; half++) {
        for (unsigned long long mk = half ? hi : lo; mk; mk &= mk - 1) {
            const int k = __ffsll((long long)mk) - 1 + 64 * half;
            if (k < S) {
                if (!objects_only && rect_rect(r, nb[k].r)) {
                    flags |= FL_CRASH_VEHICLE | FL_TOUCH;
                    if (touch) touch[half] |= 1ull << (k & 63);
                }
                continue;
            }
            const int ko = k - S;
            const float* Ob = obj + ko * OBJ_F;
            bool hit;
            if (OB_IS_BOX(Ob[OB_KIND])) { Rect ro = object_rect(Ob); hit = rect_rect(r, ro); }
            else hit = rect_circle(r, Ob[OB_X], Ob[OB_Y], Ob[OB_A]);
            if (!hit) continue;
            if (touch && claim_pass) touch[half] |= 1ull << (k & 63);
            if (Ob[OB_KIND] == 3.0f) { flags |= FL_CRASH_HUMAN; continue; }
            flags |= FL_TOUCH;   // a solid obstacle overlaps, whether or not its COST_ONCE flag is still to be had
            if (Ob[OB_KIND] == 4.0f) { flags |= FL_CRASH_BUILDING; continue; }   // a toll booth: no COST_ONCE latch
            if (Ob[OB_CRASHED] == 0.0f) {
                if (claim_pass) atomicMin(&obj_first[ko], slot);          // COST_ONCE: lowest slot takes the flag
                else if (obj_first[ko] == slot) flags |= FL_CRASH_OBJECT;
            }
        }
    }
    return flags;
}

// reward / cost / done + the 19 state floats of the observation for one agent
// (envs/metadrive_env.py:128-279, envs/base_env.py:586-623, obs/state_obs.py:64-151)
__device__ void agent_outputs(const MdConfig& cfg, const MapView& m, int env_step, const float* P, const float* S, float* C,
                              int* I, const int* __restrict__ rroad, const float* navi, size_t a, const StepOut& out,
                              int pass) {
    // pass: bit 0 = a step (reward / cost / done are written; the tollgate env's done_function reads the stay times),
    //       bit 1 = the tollgate env's StayTimeManager.record runs after the observation (a step or a newborn agent; not reset)
    const bool write_scalars = (pass & 1) != 0;
    float px = S[VS_POS], py = S[VS_POS + 1];
    M3 R = quat_to_m3(S[VS_QUAT], S[VS_QUAT + 1], S[VS_QUAT + 2], S[VS_QUAT + 3]);
    float hx, hy;
    heading_vec(R, hx, hy);
    float speed_kmh = sqrtf(S[VS_VEL] * S[VS_VEL] + S[VS_VEL + 1] * S[VS_VEL + 1]) * 3.6f;
    int flags = I[VI_FLAGS];
    int c0 = I[VI_CKPT0], n_ck = I[VI_ROUTE_LEN];
    // the two road chains (current road, final road of the route) are independent: their loads are issued side by side, so
    // that the thread waits for one round trip per level instead of walking one chain after the other
    const int cur_road = rroad[c0];
    const int final_road = rroad[n_ck - 2];
    const int cur_first = m.road_i[cur_road * ROAD_I + RI_FIRST], cur_n = m.road_i[cur_road * ROAD_I + RI_N];
    const int final_lane = m.road_i[final_road * ROAD_I + RI_FIRST] + m.road_i[final_road * ROAD_I + RI_N] - 1;
    const int road_neg = m.road_i[cur_road * ROAD_I + RI_NEG];
    int lane = I[VI_LANE];
    float lane_w = m.lane_f[lane * LANE_F + LF_WIDTH];
    const float fl_len = m.lane_f[final_lane * LANE_F + LF_LENGTH];
    const int cur_block = m.road_i[cur_road * ROAD_I + RI_BLOCK];
    const bool toll_now = cfg.toll_env && cur_block == '$';
    if (write_scalars) {
        I[VI_EP_LEN] += 1;
        int rl = lane;
        float positive = 1.0f;
        if (!(lane >= cur_first && lane < cur_first + cur_n)) {
            rl = cur_first;
            positive = (road_neg && !cfg.ignore_road_sign) ? -1.0f : 1.0f;
        }
        float long_last, long_now, lat_now, tmp;
        lane_local(m.lane_f + rl * LANE_F, C[VC_LAST_X], C[VC_LAST_Y], long_last, tmp);
        lane_local(m.lane_f + rl * LANE_F, px, py, long_now, lat_now);
        float lateral_factor = cfg.use_lateral_reward ? clipf(1.0f - 2.0f * fabsf(lat_now) / lane_w, 0.0f, 1.0f) : 1.0f;
        float rew = 0.0f;
        rew += cfg.driving_reward * (long_now - long_last) * lateral_factor * positive;
        // MultiAgentTollgateEnv.reward_function (marl_tollgate.py:217-226): inside the toll block a vehicle faster than its
        // lane's speed limit (pgblock/tollgate.py:19,68: 3, compared with km/h - base_vehicle.py:910-911) is paid the overspeed
        // penalty INSTEAD of the driving reward, and nobody gets a speed reward there
        if (toll_now) {
            const bool lane_toll = m.road_i[m.lane_i[lane * LANE_I + LI_ROAD] * ROAD_I + RI_BLOCK] == '$';
            if ((lane_toll ? 3.0f : 1000.0f) < speed_kmh) rew = -cfg.overspeed_penalty * speed_kmh / P[VP_MAX_SPEED];
        } else
            rew += cfg.speed_reward * (speed_kmh / P[VP_MAX_SPEED]) * positive;
        float step_reward = rew;
        float fl_long, fl_lat;
        lane_local(m.lane_f + final_lane * LANE_F, px, py, fl_long, fl_lat);
        bool arrive = (fl_len - 5.0f < fl_long && fl_long < fl_len + 5.0f) &&
                      (lane_w / 2.0f >= fl_lat && fl_lat >= (0.5f - (float)cur_n) * lane_w);
        bool outr = !(flags & FL_ON_LANE);
        if (cfg.out_of_route_done) outr = outr || (flags & FL_OUT_OF_ROUTE);
        else if (cfg.on_continuous_line_done == 5) outr = outr || (flags & (FL_ON_YELLOW | FL_CRASH_SIDEWALK));   // parking-lot env: white lines may be crossed
        else if (cfg.on_continuous_line_done == 2) outr = outr || (flags & (FL_ON_WHITE | FL_CRASH_SIDEWALK));   // bottleneck env, yellow line allowed
        else if (cfg.on_continuous_line_done >= 3)   // marl_tollgate.py:239-245: leaving the lanes is not out of road there
            outr = (flags & (cfg.on_continuous_line_done == 3 ? (FL_ON_YELLOW | FL_CRASH_SIDEWALK) : FL_CRASH_SIDEWALK)) != 0;
        else if (cfg.on_continuous_line_done) outr = outr || (flags & (FL_ON_YELLOW | FL_ON_WHITE | FL_CRASH_SIDEWALK));
        if (arrive) rew = cfg.success_reward;
        else if (outr) rew = -cfg.out_of_road_penalty;
        else if (flags & FL_CRASH_VEHICLE) rew = -cfg.crash_vehicle_penalty;
        else if (flags & FL_CRASH_OBJECT) rew = -cfg.crash_object_penalty;
        C[VC_EP_REWARD] += rew;
        bool max_step = cfg.horizon > 0 && I[VI_EP_LEN] >= cfg.horizon;
        bool done = false, stay_violation = false;
        if (arrive) done = true;
        if (outr) done = true;
        if ((flags & FL_CRASH_VEHICLE) && cfg.crash_vehicle_done) done = true;
        if ((flags & FL_CRASH_OBJECT) && cfg.crash_object_done) done = true;
        if (flags & FL_CRASH_BUILDING) done = true;
        if ((flags & FL_CRASH_HUMAN) && cfg.crash_human_done) done = true;
        if (max_step && cfg.truncate_as_terminate) done = true;
        if (cfg.is_multi_agent && !max_step) {  // MultiAgentMetaDrive.done_function (multi_agent_metadrive.py:114-128)
            int crash = flags & (FL_CRASH_VEHICLE | FL_CRASH_OBJECT | FL_CRASH_BUILDING | FL_CRASH_SIDEWALK | FL_CRASH_HUMAN);
            if (cfg.toll_env) crash = flags & FL_CRASH_VEHICLE;   // marl_tollgate.py:250 reads crash_vehicle only
            if (crash && !cfg.ma_crash_done && !(arrive || outr)) done = false;
            if (outr && !cfg.ma_out_of_road_done && !arrive) done = false;
            // marl_tollgate.py:261-266: through the toll block in less than min_pass_steps - as recorded up to the previous step
            if (cfg.toll_env && (((int)C[VC_TOLL_A] >> 2) & 1)) { done = true; stay_violation = true; }
        }
        float c = 0.0f;
        if (outr) c = cfg.out_of_road_cost;
        else if (flags & FL_CRASH_VEHICLE) c = cfg.crash_vehicle_cost;
        else if (flags & FL_CRASH_OBJECT) c = cfg.crash_object_cost;
        C[VC_TOTAL_COST] += c;
        bool trunc = max_step;
        if (cfg.horizon > 0 && env_step > 5 * cfg.horizon) {
            trunc = true;
            if (cfg.truncate_as_terminate) done = true;
        }
        if (done) I[VI_DONE] = 1;
        out.reward[a] = rew; out.cost[a] = c;
        out.term[a] = (uint8_t)(I[VI_DONE] != 0); out.trunc[a] = (uint8_t)trunc;
        out.info_flags[a] = flags | ((outr || stay_violation) ? FL_OUT_OF_ROAD : 0) | (arrive ? FL_ARRIVE : 0) | (max_step ? FL_MAX_STEP : 0);
        float4* inf = reinterpret_cast<float4*>(out.info_f + a * 8);
        inf[0] = make_float4(sqrtf(S[VS_VEL] * S[VS_VEL] + S[VS_VEL + 1] * S[VS_VEL + 1]), S[VS_STEER], S[VS_THROTTLE], C[VC_STEP_ENERGY]);
        inf[1] = make_float4(C[VC_ENERGY], step_reward, C[VC_EP_REWARD], (float)I[VI_EP_LEN]);
    }
    float* o = out.obs + a * (size_t)OBS_DIM(cfg);
    const int sd = OBS_SIDE(cfg);
    if (cfg.n_side_lasers == 0) {  // else k_linedet fills the block
        o[0] = clipf(C[VC_DIST_L] / 18.0f, 0.0f, 1.0f);
        o[1] = clipf(C[VC_DIST_R] / 18.0f, 0.0f, 1.0f);
    }
    {
        const float* Lr = m.lane_f + (cur_first + cur_n - 1) * LANE_F;
        float lx, ly;
        if (Lr[LF_TYPE] == 0.0f) {
            float ex = Lr[LF_P0 + 2] - Lr[LF_P0 + 0], ey = Lr[LF_P0 + 3] - Lr[LF_P0 + 1];
            float ln = sqrtf(ex * ex + ey * ey);
            lx = ey / ln; ly = -ex / ln;
        } else if (Lr[LF_P0 + 5] > 0.0f) { lx = px - Lr[LF_P0 + 0]; ly = py - Lr[LF_P0 + 1]; }
        else { lx = Lr[LF_P0 + 0] - px; ly = Lr[LF_P0 + 1] - py; }
        float ln = sqrtf(lx * lx + ly * ly), fn = sqrtf(hx * hx + hy * hy);
        float hd = 0.0f;
        if (ln * fn != 0.0f) hd = clipf((hx * lx + hy * ly) / (ln * fn), -1.0f, 1.0f) / 2.0f + 0.5f;
        o[sd + 0] = hd;
    }
    o[sd + 1] = clipf((speed_kmh + 1.0f) / (P[VP_MAX_SPEED] + 1.0f), 0.0f, 1.0f);
    o[sd + 2] = clipf((S[VS_STEER] / 60.0f + 1.0f) / 2.0f, 0.0f, 1.0f);
    o[sd + 3] = clipf((C[VC_CUR_A0] + 1.0f) / 2.0f, 0.0f, 1.0f);
    o[sd + 4] = clipf((C[VC_CUR_A1] + 1.0f) / 2.0f, 0.0f, 1.0f);
    {
        float lhx = C[VC_LAST_HX], lhy = C[VC_LAST_HY];
        float dotp = hx * lhx + hy * lhy, crs = hx * lhy - hy * lhx;
        float beta = dotp <= 0.0f ? 0.5f * MD_PI : md_atan2f(fabsf(crs), dotp);
        o[sd + 5] = clipf(beta / 0.1f, 0.0f, 1.0f);
    }
    if (cfg.n_lane_lasers == 0) {
        float lon, lat;
        lane_local(m.lane_f + lane * LANE_F, px, py, lon, lat);
        o[sd + 6] = clipf((lat * 2.0f / 4.5f + 1.0f) / 2.0f, 0.0f, 1.0f);
    }
    if (!cfg.toll_env) {
#pragma unroll
        for (int k = 0; k < NAVI_DIM; k++) o[OBS_EGO(cfg) + k] = navi[k];
        return;
    }
    // TollGateObservation.observe (marl_tollgate.py:92-105): the counter runs while the current road is the toll block
    const int a_ = (int)C[VC_TOLL_A];
    int in_toll = a_ >> 4, has_exit = (a_ >> 3) & 1, viol = (a_ >> 2) & 1, last = a_ & 3;
    in_toll += toll_now ? 1 : 0;
    float* to = o + OBS_DIM(cfg) - 2;
    to[0] = toll_now ? 1.0f : 0.0f;
    to[1] = (toll_now && in_toll > cfg.min_pass_steps) ? 1.0f : 0.0f;
    if (pass & 2) {
        // StayTimeManager.record (marl_tollgate.py:50-62), after the step: entry when the block changes to the toll block, exit
        // when it changes from the toll block to Merge / Split
        if (last != 0) {
            if (toll_now && last != 1) {
                C[VC_TOLL_ENTRY] = (float)(env_step + 1);
                if (has_exit) viol = cfg.min_pass_steps > 0;   // exit - entry < 0 from here on
            } else if (!toll_now && last == 1 && (cur_block == 'y' || cur_block == 'Y')) {
                has_exit = 1;
                viol = C[VC_TOLL_ENTRY] > 0.0f && (env_step - ((int)C[VC_TOLL_ENTRY] - 1)) < cfg.min_pass_steps;
            }
        }
        last = toll_now ? 1 : 2;
    }
    C[VC_TOLL_A] = (float)(16 * in_toll + 8 * has_exit + 4 * viol + last);
}

__device__ __forceinline__ void load16(float* dst, const float* src) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
    for (int k = 0; k < 4; k++) { float4 v = s4[k]; dst[4 * k] = v.x; dst[4 * k + 1] = v.y; dst[4 * k + 2] = v.z; dst[4 * k + 3] = v.w; }
}
__device__ __forceinline__ void store16(float* dst, const float* src) {
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int k = 0; k < 4; k++) d4[k] = make_float4(src[4 * k], src[4 * k + 1], src[4 * k + 2], src[4 * k + 3]);
}
__device__ __forceinline__ void load16i(int* dst, const int* src) {
    const int4* s4 = reinterpret_cast<const int4*>(src);
#pragma unroll
    for (int k = 0; k < 4; k++) { int4 v = s4[k]; dst[4 * k] = v.x; dst[4 * k + 1] = v.y; dst[4 * k + 2] = v.z; dst[4 * k + 3] = v.w; }
}
__device__ __forceinline__ void store16i(int* dst, const int* src) {
    int4* d4 = reinterpret_cast<int4*>(dst);
#pragma unroll
    for (int k = 0; k < 4; k++) d4[k] = make_int4(src[4 * k], src[4 * k + 1], src[4 * k + 2], src[4 * k + 3]);
}

__device__ __forceinline__ void fill_nb(Nb& n, const float* P, const float* S, const int* I) {
    n.x = S[VS_POS]; n.y = S[VS_POS + 1]; n.vx = S[VS_VEL]; n.vy = S[VS_VEL + 1];
    n.r = vehicle_rect(P, S);
    n.lane = I[VI_LANE]; n.alive = I[VI_ALIVE]; n.active = I[VI_ACTIVE]; n.kind = I[VI_KIND];
}

// body row for the lidar kernel (BODY_ROW floats): centre(3) half(3) R(9) alive | origin x y, velocity x y
#define BODY_ROW 20
__device__ __forceinline__ void write_body_row(float* row, const float* P, const float* S, int alive) {
    M3 R = quat_to_m3(S[VS_QUAT], S[VS_QUAT + 1], S[VS_QUAT + 2], S[VS_QUAT + 3]);
    float hh = 0.5f * P[VP_HEIGHT];
    F3 cen = f3(S[VS_POS], S[VS_POS + 1], S[VS_POS + 2]) + col(R, 2) * hh;
    float4* r4 = reinterpret_cast<float4*>(row);
    r4[0] = make_float4(cen.x, cen.y, cen.z, 0.5f * P[VP_WIDTH]);
    r4[1] = make_float4(0.5f * P[VP_LENGTH], hh, R.m[0][0], R.m[0][1]);
    r4[2] = make_float4(R.m[0][2], R.m[1][0], R.m[1][1], R.m[1][2]);
    r4[3] = make_float4(R.m[2][0], R.m[2][1], R.m[2][2], alive ? 1.0f : 0.0f);
    r4[4] = make_float4(S[VS_POS], S[VS_POS + 1], S[VS_VEL], S[VS_VEL + 1]);
}

// ================================================================================================ step kernels
// Thread mapping shared by k_pre / k_dyn / k_post: a CTA owns EPB consecutive envs; thread t handles vehicle
// (env_local = t % EPB, slot = t / EPB).  With EPB = 32 a warp is "slot s of 32 different envs": every lane of a
// warp plays the same role (slot 0 = the agent, the alive traffic is a prefix of the remaining slots), which is
// what keeps the SIMT lanes converged; warps of empty slots retire immediately.
// Shared memory per CTA: Nb[EPB*S] (env-major, so an env's records are contiguous) | object rows | COST_ONCE claims.
struct StepGeom {
    int S, O, NA, epb, le, slot, env, g;
    bool work;
    Nb* nb; float* sobj; int* obj_first;
};
__device__ __forceinline__ StepGeom step_geom(const MdConfig& cfg, int epb, unsigned char* smem_raw) {
    StepGeom G;
    G.S = cfg.slots_per_env; G.O = cfg.objs_per_env; G.NA = cfg.agents_per_env; G.epb = epb;
    G.le = threadIdx.x % epb; G.slot = threadIdx.x / epb;
    G.env = blockIdx.x * epb + G.le;
    G.work = G.slot < G.S && G.env < cfg.n_envs;
    G.g = G.env * G.S + G.slot;
    const int ofs = (G.O + 3) & ~3;
    Nb* nb_all = reinterpret_cast<Nb*>(smem_raw);
    float* obj_all = reinterpret_cast<float*>(smem_raw + sizeof(Nb) * (size_t)epb * G.S);
    int* first_all = reinterpret_cast<int*>(smem_raw + sizeof(Nb) * (size_t)epb * G.S + sizeof(float) * OBJ_F * (size_t)G.O * epb);
    G.nb = nb_all + (size_t)G.le * G.S;
    G.sobj = obj_all + (size_t)G.le * G.O * OBJ_F;
    G.obj_first = first_all + (size_t)G.le * ofs;
    return G;
}
__host__ __device__ inline size_t step_smem_bytes(int S, int O, int epb) {
    return (sizeof(Nb) * (size_t)S + sizeof(float) * OBJ_F * (size_t)O + sizeof(int) * (size_t)((O + 3) & ~3)) * epb;
}
// stage the env's object rows: the threads of one env (its S slots) stride over the O*OBJ_F floats
__device__ __forceinline__ void stage_objects(const StepGeom& G, const float* __restrict__ obj_f) {
    if (G.work)
        for (int k = G.slot; k < G.O * OBJ_F; k += G.S) G.sobj[k] = obj_f[(size_t)G.env * G.O * OBJ_F + k];
}

// ---- k_pre: engine.before_step (agent actuation, traffic trigger, IDM decisions) ------------------------------
// Same shape as k_post: a CTA owns `epb` envs and runs PRE_WORKERS threads.  Phase 1 sweeps the slot rows: neighbour
// records of the alive vehicles go to shared memory, agents are actuated on the spot, alive traffic is appended to a work
// list.  Then one thread per env runs the trigger, and phase 2 runs the IDM with one thread per listed vehicle.
#ifndef PRE_WORKERS
#define PRE_WORKERS 256
#endif
#define PRE_TEAM 4                    // lanes per vehicle in the IDM phase
#define PRE_TEAM_SCRATCH (32 * 1024)  // the teams' tables (one float2 per neighbour index and team) must fit this
// k_pre stages the envs' object rows in shared memory only while they are few: a SafeMetaDriveEnv CTA (16 envs x 60 obstacles)
// would carry 46 KB of them and fit twice per SM - a second wave for half the CTAs - where the IDM reads a handful of fields
// per object, which the read-only cache serves as well
__host__ __device__ inline bool pre_stage_objs(int O, int epb) { return sizeof(float) * OBJ_F * (size_t)O * epb <= 16 * 1024; }
__host__ __device__ inline size_t pre_base_bytes(int S, int O, int epb) {   // Nb rows | objects | list | counters | active list
    size_t b = (sizeof(Nb) + 2 * sizeof(int)) * (size_t)S * epb + (pre_stage_objs(O, epb) ? sizeof(float) * OBJ_F * (size_t)O * epb : 0) + 32;
    return (b + 15) & ~(size_t)15;
}
// ... followed by one MapView per env (the map's table pointers, read by every IDM team of the env: without it each team
// chases env row -> map descriptor -> tables itself, two dependent round trips) and then the teams' neighbour tables
__host__ __device__ inline size_t pre_views_bytes(int epb) { return (sizeof(MapView) * (size_t)epb + 15) & ~(size_t)15; }
// the teams' tables are sized for 2-lane teams (threads / 2 of them): a CTA with more active vehicles than 4-lane teams
// halves the team size instead of dropping to one thread per vehicle (which made that CTA the kernel's tail)
// A team's table has one entry per VALID neighbour (the bodies within 50 m), in their order - not one per slot / object of
// the env: a SafeMetaDriveEnv scene holds 60 obstacles and a pedestrian scene 16 more bodies, of which a vehicle sees a few.
// PRE_NB_CAP entries per team; a vehicle with more valid neighbours runs the search on one lane (idm_act).
__host__ __device__ inline int pre_team_size(int S, int O, int threads) {
    return sizeof(float2) * (size_t)pre_nb_cap(S, O) * (threads / 2) <= PRE_TEAM_SCRATCH ? PRE_TEAM : 1;
}
__host__ __device__ inline size_t pre_smem_bytes(int S, int O, int epb, int threads) {
    const int T = pre_team_size(S, O, threads);
    return pre_base_bytes(S, O, epb) + pre_views_bytes(epb) + (T > 1 ? sizeof(float2) * (size_t)pre_nb_cap(S, O) * (threads / 2) : 0);
}
__global__ void __maxnreg__(PRE_REGS)
k_pre(MdConfig cfg, MdArrays A, int mode, int epb, const float* __restrict__ actions, float* __restrict__ idm_out,
      float4* __restrict__ veh_act, int use_teams, uint32_t* __restrict__ pass_ctr, uint32_t d_bank, uint32_t d_noise,
      unsigned int* __restrict__ work_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // the first kernel of a step advances the view's pass counters (scenario draws, observation passes); they live in
    // device memory so that a whole step is a fixed launch sequence (CUDA-graph replayable).  Nobody reads them in k_pre.
    if (pass_ctr != nullptr && blockIdx.x == 0 && threadIdx.x == 0) { pass_ctr[0] += d_bank; pass_ctr[1] += d_noise; }
    if (work_count != nullptr && blockIdx.x == 0 && threadIdx.x == 0) { work_count[0] = 0u; work_count[1] = 0u; }   // k_dyn's list of this step starts empty; [1] = k_scan's cursor
    const int S = cfg.slots_per_env, O = cfg.objs_per_env, NA = cfg.agents_per_env;
    const int n_rows = epb * S, env0 = blockIdx.x * epb;
    Nb* nb_all = reinterpret_cast<Nb*>(smem_raw);
    const bool stage_objs = pre_stage_objs(O, epb);
    float* obj_all = reinterpret_cast<float*>(smem_raw + sizeof(Nb) * (size_t)n_rows);
    int* list = reinterpret_cast<int*>(smem_raw + sizeof(Nb) * (size_t)n_rows + (stage_objs ? sizeof(float) * OBJ_F * (size_t)O * epb : 0));
    int* n_list = list + n_rows;
    if (threadIdx.x == 0) *n_list = 0;
    clk_mark(0, 0);
    __syncthreads();
    const float4 idle = make_float4(0.0f, 0.0f, 2.0f, 0.0f);  // what vehicle.reset() leaves (base_vehicle.py:376): brake 2
    // ---- phase 1.  The rows are visited agents first (they take the long path: latches, actuation, two more rows), so that
    // the threads' first pass holds all of them and a second pass only short traffic rows.
    MapView* mv_all = reinterpret_cast<MapView*>(smem_raw + pre_base_bytes(S, O, epb));
    if ((int)threadIdx.x < epb && env0 + (int)threadIdx.x < cfg.n_envs)
        mv_all[threadIdx.x] = map_view(A, A.env_i[(env0 + threadIdx.x) * ENV_I + EI_MAP]);
    const int n_agent_rows = epb * NA;
    for (int i = threadIdx.x; i < n_rows; i += blockDim.x) {
        int le, slot;
        if (i < n_agent_rows) { le = i / NA; slot = i - le * NA; }
        else { const int j = i - n_agent_rows; le = j / (S - NA); slot = NA + (j - le * (S - NA)); }
        const int v = le * S + slot, env = env0 + le;
        Nb& n = nb_all[v];
        n.alive = 0; n.active = 0; n.kind = 0; n.lane = -1;
        if (env >= cfg.n_envs) continue;
        const size_t g = (size_t)env * S + slot;
        int I[VEH_I];
        load16i(I, A.veh_i + g * VEH_I);
        if (I[VI_KIND] == 0) continue;
        const bool is_agent = I[VI_KIND] == 1;
        bool dirty_i = false;
        // VehicleAgentManager.before_step, second half (manager/agent_manager.py:189-202): wrecks of finished agents
        // stay in the world as static bodies for delay_done steps, then leave
        if ((mode & MODE_AGENT_PRE) && cfg.is_multi_agent && is_agent) {
            if (I[VI_NEW]) { I[VI_NEW] = 0; dirty_i = true; }
            if (I[VI_ALIVE] && !I[VI_ACTIVE] && I[VI_DYING] > 0) {
                I[VI_DYING] -= 1;
                if (I[VI_DYING] <= 0) { I[VI_ALIVE] = 0; I[VI_STATIC] = 0; }
                dirty_i = true;
            }
        }
        n.kind = I[VI_KIND]; n.lane = I[VI_LANE]; n.active = I[VI_ACTIVE];
        if (I[VI_ALIVE]) {
            float P[VEH_P], St[VEH_S];
            load16(P, A.veh_p + g * VEH_P);
            load16(St, A.veh_s + g * VEH_S);
            fill_nb(n, P, St, I);
            // agent_manager.before_step (manager/agent_manager.py:164-202)
            if ((mode & MODE_AGENT_PRE) && is_agent && I[VI_ACTIVE]) {
                float C[VEH_C];
                load16(C, A.veh_c + g * VEH_C);
                latch_before_step(St, C, I);
                const float* a = actions + ((size_t)env * NA + slot) * 2;
                float a0 = a[0], a1 = a[1];
                decode_action(cfg, a0, a1);
                const Actuation act = actuate(P, St, C, a0, a1);
                veh_act[g] = make_float4(act.steer_rad, act.engine, act.brake, 0.0f);
                store16(A.veh_s + g * VEH_S, St);
                store16(A.veh_c + g * VEH_C, C);
                dirty_i = true;
            } else {
                veh_act[g] = idle;
                if (I[VI_KIND] == 2) list[atomicAdd(n_list, 1)] = v;
            }
        } else veh_act[g] = idle;
        if (dirty_i) store16i(A.veh_i + g * VEH_I, I);
    }
    if (stage_objs)
        for (int k = threadIdx.x; k < epb * O * OBJ_F; k += blockDim.x) {
            const int env = env0 + k / (O * OBJ_F);
            if (env < cfg.n_envs) obj_all[k] = A.obj_f[(size_t)env0 * O * OBJ_F + k];
        }
    __syncthreads();
    clk_mark(0, 1);
    // ---- PGTrafficManager.before_step: trigger (manager/traffic_manager.py:74-88), one thread per env
    for (int le = threadIdx.x; le < epb; le += blockDim.x) {
        const int env = env0 + le;
        if (env >= cfg.n_envs) continue;
        Nb* nb = nb_all + (size_t)le * S;
        if ((mode & MODE_TRIGGER) && cfg.traffic_mode != 1) {
            int nt = A.env_i[env * ENV_I + EI_NEXT_TRIGGER];
            if (nt > 0) {
                const MapView& m = mv_all[le];
                const int n_blocks = A.env_i[env * ENV_I + EI_N_BLOCKS];
                for (int s = 0; s < NA && nt > 0; s++) {
                    if (!nb[s].active || nb[s].kind != 1) continue;
                    const int road = m.lane_i[nb[s].lane * LANE_I + LI_ROAD];
                    if (road == A.env_trigger[env * TRIGGER_MAX + nt]) {
                        for (int k = 0; k < S; k++)
                            if (nb[k].kind == 2 && nb[k].alive && A.veh_i[(size_t)(env * S + k) * VEH_I + VI_TRIGGER] == nt) nb[k].active = 1;
                        nt = nt + 1 < n_blocks ? nt + 1 : 0;
                    }
                }
                A.env_i[env * ENV_I + EI_NEXT_TRIGGER] = nt;
            }
        }
        if (mode & MODE_AGENT_PRE) A.env_i[env * ENV_I + EI_STEP] += 1;
    }
    __syncthreads();
    clk_mark(0, 2);
    // ---- phase 2: IDM decisions against the pre-step world (policy/idm_policy.py:235-267), one thread per active vehicle
    if (!(mode & (MODE_IDM | MODE_IDM_OUT))) return;
    // The listed vehicles that are active (triggered) are compacted once more, and a team of PRE_TEAM lanes runs each
    // of them: a GPU holds ~150 active traffic vehicles per SM at BASELINE cfg2 - one thread per vehicle would leave
    // 5 warps per SM walking the neighbour tables one entry at a time.
    const int n_work = *n_list;
    int* alist = list + n_rows + 8;
    int* n_alist = n_list + 1;
    if (threadIdx.x == 0) *n_alist = 0;
    __syncthreads();
    for (int j = threadIdx.x; j < n_work; j += blockDim.x)
        if (nb_all[list[j]].active) alist[atomicAdd(n_alist, 1)] = list[j];
    __syncthreads();
    const int n_act = *n_alist;
    clk_mark(0, 3);
    // teams only while one round covers the CTA's active vehicles (each round is a full IDM chain) and the tables fit
    int T = use_teams ? pre_team_size(S, O, blockDim.x) : 1;
    while (T > 1 && n_act * T > (int)blockDim.x) T >>= 1;
#ifdef MD_PHASE_CLK
    if (g_phase_clk != nullptr && threadIdx.x == 0 && blockIdx.x < CLK_CTAS) {
        g_phase_clk[((size_t)0 * CLK_CTAS + blockIdx.x) * 16 + 10] = (unsigned long long)n_act;
        g_phase_clk[((size_t)0 * CLK_CTAS + blockIdx.x) * 16 + 11] = (unsigned long long)T;
    }
#endif
    const int sub = threadIdx.x & (T - 1), team = threadIdx.x / T, n_teams = blockDim.x / T;
    const unsigned team_mask = T > 1 ? (((1u << T) - 1u) << ((threadIdx.x & 31) & ~(T - 1))) : 0u;
    float2* scratch = reinterpret_cast<float2*>(smem_raw + pre_base_bytes(S, O, epb) + pre_views_bytes(epb)) + (size_t)team * pre_nb_cap(S, O);
    for (int j = team; j < n_act; j += n_teams) {
        const int v = alist[j];
        const int le = v / S, slot = v - le * S, env = env0 + le;
        const size_t g = (size_t)env * S + slot;
        float P[VEH_P], St[VEH_S], C[VEH_C], D[VEH_IDM];
        int I[VEH_I];
        load16i(I, A.veh_i + g * VEH_I);
        load16(P, A.veh_p + g * VEH_P);
        load16(St, A.veh_s + g * VEH_S);
        load16(C, A.veh_c + g * VEH_C);
        const float4* d4 = reinterpret_cast<const float4*>(A.veh_idm + g * VEH_IDM);
        const float4 d0 = d4[0], d1 = d4[1];
        D[0] = d0.x; D[1] = d0.y; D[2] = d0.z; D[3] = d0.w; D[4] = d1.x; D[5] = d1.y; D[6] = d1.z; D[7] = d1.w;
        I[VI_ACTIVE] = 1;  // possibly triggered just now
        const MapView m = mv_all[le];
        NbrView nv;
        nv.nb = nb_all + (size_t)le * S; nv.S = S; nv.O = O; nv.self = slot;
        nv.obj = stage_objs ? obj_all + (size_t)le * O * OBJ_F : A.obj_f + (size_t)env * O * OBJ_F;
        nv.px = St[VS_POS]; nv.py = St[VS_POS + 1];
        float a0, a1;
        idm_act(cfg, m, nv, (int)g + cfg.env_base * S, St, I, D, A.veh_rroad + g * ROUTE_MAX, a0, a1, sub, T, team_mask, scratch);
        if (sub != 0) continue;   // the lanes of a team hold identical results: one of them writes
        if (mode & MODE_IDM_OUT) { idm_out[2 * g] = a0; idm_out[2 * g + 1] = a1; }
        if (mode & MODE_IDM) {
            latch_before_step(St, C, I);
            const Actuation act = actuate(P, St, C, a0, a1);
            veh_act[g] = make_float4(act.steer_rad, act.engine, act.brake, 0.0f);
        }
        store16(A.veh_s + g * VEH_S, St);
        store16(A.veh_c + g * VEH_C, C);
        store16i(A.veh_i + g * VEH_I, I);
        float4* o4 = reinterpret_cast<float4*>(A.veh_idm + g * VEH_IDM);
        o4[0] = make_float4(D[0], D[1], D[2], D[3]);
        o4[1] = make_float4(D[4], D[5], D[6], D[7]);
    }
#ifdef MD_PHASE_CLK
    __syncthreads();
    clk_mark(0, 13);
#endif
}

// ---- k_dyn: engine.step = n_sub x doPhysics + contact-added callback (engine/base_engine.py:417-445) --------------
#ifndef MD_RESP_INL
#define MD_RESP_INL __forceinline__   // out of line (__noinline__) triples the spills of the sub-step loop around the call
#endif
// Contact response of vehicle `slot` over its candidate set (bit k < S = vehicle slot k, else object k - S), from the
// snapshot every vehicle of the env published after moving (nb[].r footprints, cd[] origin / velocity / inverse mass).
// The pair (i, k) is evaluated with A = the lower slot on both sides, so the two threads compute the same impulse.
__device__ MD_RESP_INL void contact_response(const Nb* nb, const CBody* cd, const float* obj, int S, int slot,
                                              unsigned long long lo, unsigned long long hi, float* d) {
    d[0] = d[1] = d[2] = d[3] = d[4] = 0.0f;
    const CBody me = cd[slot];
    const Rect mr = nb[slot].r;
    for (int half = 0; half < 2; half++) {
        for (unsigned long long mk = half ? hi : lo; mk; mk &= mk - 1) {
            const int k = __ffsll((long long)mk) - 1 + 64 * half;
            float nx, ny, depth, px, py;
            if (k < S) {
                const bool me_is_b = k < slot;
                const Rect& ra = me_is_b ? nb[k].r : mr;
                const Rect& rb = me_is_b ? mr : nb[k].r;
                if (!rr_contact(ra, rb, nx, ny, depth, px, py)) continue;
                if (me_is_b) pair_impulse(cd[k], me, nx, ny, depth, px, py, true, d);
                else pair_impulse(me, cd[k], nx, ny, depth, px, py, false, d);
                continue;
            }
            const float* Ob = obj + (k - S) * OBJ_F;
            if (Ob[OB_KIND] == 3.0f) continue;  // pedestrians neither push nor are pushed
            CBody ob;
            ob.ox = Ob[OB_X]; ob.oy = Ob[OB_Y]; ob.vx = 0.0f; ob.vy = 0.0f; ob.w = 0.0f; ob.im = 0.0f; ob.ii = 0.0f;
            bool hit;
            if (OB_IS_BOX(Ob[OB_KIND])) { const Rect ro = object_rect(Ob); hit = rr_contact(mr, ro, nx, ny, depth, px, py); }
            else hit = rc_contact(mr, Ob[OB_X], Ob[OB_Y], Ob[OB_A], nx, ny, depth, px, py);
            if (hit) pair_impulse(me, ob, nx, ny, depth, px, py, false, d);
        }
    }
}

// smem of k_dyn beyond step_smem_bytes: CBody rows | list | 4 counters | alive count per env | (16-aligned) contact words
__host__ __device__ inline size_t dyn_list_bytes(int S, int epb) { return sizeof(int) * ((size_t)epb * S + 4 + epb); }
__global__ void __maxnreg__(DYN_REGS)
k_dyn(MdConfig cfg, MdArrays A, int mode, int epb, const float4* __restrict__ veh_act, const float* __restrict__ ext_act3,
      int n_sub, uint4* __restrict__ contact_tab, int* __restrict__ work_list, unsigned int* __restrict__ work_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // A CTA owns `epb` envs but runs only about as many threads as those envs hold ALIVE vehicles (a third to a half of the
    // slot rows of a PG scene are empty or not alive, and idle threads would still pin 96 registers each): housekeeping -
    // sweeping the slot rows, staging and maintaining the objects - strides over the rows, and every thread takes its
    // vehicle from a compacted list of the alive ones.  The envs are handled in passes: a pass takes as many consecutive
    // envs as have, together, at most blockDim.x alive vehicles (normally all of them: one pass; the host sizes the CTA from
    // the scene's alive counts).  Contacts never cross envs, so the passes are independent.
    const int S = cfg.slots_per_env, O = cfg.objs_per_env, ofs = (O + 3) & ~3;
    const int W = blockDim.x, tid = threadIdx.x;
    const int env0 = blockIdx.x * epb, n_env = min(epb, cfg.n_envs - env0), n_rows = n_env * S;
    Nb* nb_all = reinterpret_cast<Nb*>(smem_raw);
    float* obj_all = reinterpret_cast<float*>(smem_raw + sizeof(Nb) * (size_t)epb * S);
    int* first_all = reinterpret_cast<int*>(smem_raw + sizeof(Nb) * (size_t)epb * S + sizeof(float) * OBJ_F * (size_t)O * epb);
    CBody* cd_all = reinterpret_cast<CBody*>(smem_raw + step_smem_bytes(S, O, epb));
    int* list = reinterpret_cast<int*>(smem_raw + step_smem_bytes(S, O, epb) + sizeof(CBody) * (size_t)epb * S);
    int* n_list = list + epb * S;
    int* env_cnt = n_list + 4;
    // contact export (md_enable_contacts): two words per slot row in shared memory, ORed by the row's own vehicle thread
    unsigned long long* touch_all = reinterpret_cast<unsigned long long*>(
        smem_raw + ((step_smem_bytes(S, O, epb) + sizeof(CBody) * (size_t)epb * S + dyn_list_bytes(S, epb) + 15) & ~(size_t)15));
    const bool contacts = (mode & MODE_CONTACTS) != 0;
    clk_mark(1, 0);
    for (int le = tid; le < epb; le += W) env_cnt[le] = 0;
    __syncthreads();
    for (int row = tid; row < n_rows; row += W) {
        const int le = row / S;
        const int4 i0 = *reinterpret_cast<const int4*>(A.veh_i + ((size_t)env0 * S + row) * VEH_I);  // kind, alive, active, trigger
        nb_all[row].alive = i0.y; nb_all[row].active = i0.z;
        if (i0.y) atomicAdd(&env_cnt[le], 1);
        if (contact_tab != nullptr) { touch_all[2 * row] = 0ull; touch_all[2 * row + 1] = 0ull; }
    }
    if (contacts)   // stage the object rows of the CTA's envs (contiguous in global memory)
        for (int k = tid; k < n_env * O * OBJ_F; k += W) obj_all[k] = A.obj_f[(size_t)env0 * O * OBJ_F + k];
    __syncthreads();
    clk_mark(1, 1);
    for (int le0 = 0; le0 < n_env;) {
        int le1 = le0, n_pass = 0;
        while (le1 < n_env && n_pass + env_cnt[le1] <= W) { n_pass += env_cnt[le1]; le1++; }
        if (le1 == le0) le1 = le0 + 1;   // cannot happen (an env holds at most S <= blockDim.x vehicles); never spin
        const int row0 = le0 * S, row1 = le1 * S;
        if (tid == 0) { n_list[0] = 0; n_list[1] = 0; }
        __syncthreads();
        // the list is filled from both ends: driven vehicles (agents, triggered traffic) from the front, parked ones (the
        // untriggered traffic of later blocks: brakes on, four wheels down - the majority in a PG scene) from the back, so
        // that a warp integrates vehicles on the same control-flow path
        for (int row = row0 + tid; row < row1; row += W)
            if (nb_all[row].alive) {
                if (nb_all[row].active) list[atomicAdd(&n_list[0], 1)] = row;
                else list[epb * S - 1 - atomicAdd(&n_list[1], 1)] = row;
            }
        __syncthreads();
        const int n_front = n_list[0], n_back = n_list[1];
        const bool work = tid < n_front + n_back;
        int le = le0, slot = 0;
        if (work) {
            const int v = tid < n_front ? list[tid] : list[epb * S - 1 - (tid - n_front)];
            le = v / S; slot = v - le * S;
        }
        const int g = (env0 + le) * S + slot;
        Nb* nb = nb_all + (size_t)le * S;
        float* sobj = obj_all + (size_t)le * O * OBJ_F;
        int* obj_first = first_all + (size_t)le * ofs;
        CBody* cd = cd_all + (size_t)le * S;
        clk_mark(1, 2);
        float P[VEH_P], St[VEH_S];
        int is_static = 1, flags = 0;
        Actuation act;
        act.steer_rad = 0.0f; act.engine = 0.0f; act.brake = 2.0f;
        if (work) {
            const int* I = A.veh_i + (size_t)g * VEH_I;
            is_static = I[VI_STATIC]; flags = I[VI_FLAGS];
            load16(P, A.veh_p + (size_t)g * VEH_P);
            load16(St, A.veh_s + (size_t)g * VEH_S);
            if (mode & MODE_EXT_ACT) {
                act.steer_rad = ext_act3[3 * (size_t)g]; act.engine = ext_act3[3 * (size_t)g + 1]; act.brake = ext_act3[3 * (size_t)g + 2];
            } else {
                float4 a = veh_act[g];
                act.steer_rad = a.x; act.engine = a.y; act.brake = a.z;
            }
        }
        Body B;
        B.pos = f3(St[VS_POS], St[VS_POS + 1], St[VS_POS + 2]);
        B.q[0] = St[VS_QUAT]; B.q[1] = St[VS_QUAT + 1]; B.q[2] = St[VS_QUAT + 2]; B.q[3] = St[VS_QUAT + 3];
        B.v = f3(St[VS_VEL], St[VS_VEL + 1], St[VS_VEL + 2]);
        B.w = f3(St[VS_ANGVEL], St[VS_ANGVEL + 1], St[VS_ANGVEL + 2]);
        const bool moves = work && !is_static;
        // Contact candidates of this step (the broad phase): bit k of (cand_lo, cand_hi) = body k (vehicle slot k < S, object
        // k - S) can come within touching distance during the n_sub sub-steps: centre distance at the start of the step <=
        // the two bounding radii + 1.5 x the distance both can travel + 0.5 m.  "Can travel" uses the speed of the env's
        // fastest body for both sides: an inelastic contact impulse hands a hit body at most the speed of the one that hit
        // it.  Every truly overlapping pair is a candidate (accelerations change a speed by < 1 m/s within a step), so the
        // exact SAT below sees the same pairs as a full scan.
        unsigned long long cand_lo = 0ull, cand_hi = 0ull;
        int block_any = 0, block_obj = 0;
        if (contacts) {
            if (work) {
                const Rect r0 = vehicle_rect(P, St);
                nb[slot].r = r0;
                nb[slot].vx = sqrtf(r0.hu * r0.hu + r0.hv * r0.hv);                                  // bounding radius
                nb[slot].vy = sqrtf(B.v.x * B.v.x + B.v.y * B.v.y + B.v.z * B.v.z);                  // speed
            }
            __syncthreads();
            bool has_obj = false;
            if (work) {
                const float T = 1.5f * cfg.dt * (float)n_sub;
                const Rect r0 = nb[slot].r;
                float vmax = 0.0f;
#ifndef MD_CAND_OWN_SPEED
                for (int k = 0; k < S; k++) if (nb[k].alive) vmax = fmaxf(vmax, nb[k].vy);
                const float my_rad = nb[slot].vx + 0.5f, my_speed = vmax;
#else
                const float my_rad = nb[slot].vx + 0.5f, my_speed = nb[slot].vy;
#endif
                for (int k = 0; k < S; k++) {
                    if (k == slot || !nb[k].alive) continue;
                    const float dx = nb[k].r.cx - r0.cx, dy = nb[k].r.cy - r0.cy;
#ifndef MD_CAND_OWN_SPEED
                    const float reach = my_rad + nb[k].vx + (my_speed + vmax) * T;
#else
                    const float reach = my_rad + nb[k].vx + (my_speed + nb[k].vy) * T;
#endif
                    if (dx * dx + dy * dy <= reach * reach) { if (k < 64) cand_lo |= 1ull << k; else cand_hi |= 1ull << (k - 64); }
                }
                for (int k = 0; k < O; k++) {
                    const float* Ob = sobj + k * OBJ_F;
                    if (Ob[OB_KIND] < 0.0f) continue;
                    const float orad = OB_IS_BOX(Ob[OB_KIND]) ? sqrtf(Ob[OB_A] * Ob[OB_A] + Ob[OB_B] * Ob[OB_B]) : Ob[OB_A];
                    const float ospeed = Ob[OB_KIND] == 3.0f ? sqrtf(Ob[OB_VX] * Ob[OB_VX] + Ob[OB_VY] * Ob[OB_VY]) : 0.0f;
                    const float dx = Ob[OB_X] - r0.cx, dy = Ob[OB_Y] - r0.cy;
                    const float reach = my_rad + orad + (my_speed + ospeed) * T;
                    if (dx * dx + dy * dy <= reach * reach) {
                        const int b = S + k;
                        if (b < 64) cand_lo |= 1ull << b; else cand_hi |= 1ull << (b - 64);
                        has_obj = true;
                    }
                }
            }
            block_any = __syncthreads_or((cand_lo | cand_hi) != 0ull);
            block_obj = __syncthreads_or(has_obj);
        }
        clk_mark(1, 3);
        SteerCS scs;
        scs.cs = md_cosf(act.steer_rad); scs.sn = md_sinf(act.steer_rad);
        const int ko0 = le0 * O, ko1 = le1 * O;   // the pass's object rows, and their (padded) COST_ONCE claim words
        for (int rep = 0; rep < n_sub; rep++) {
            if (moves) {
                vehicle_substep(P, B, act, scs, cfg.dt);
                St[VS_POS] = B.pos.x; St[VS_POS + 1] = B.pos.y; St[VS_POS + 2] = B.pos.z;
                St[VS_QUAT] = B.q[0]; St[VS_QUAT + 1] = B.q[1]; St[VS_QUAT + 2] = B.q[2]; St[VS_QUAT + 3] = B.q[3];
            }
            if (contacts && block_any) {  // a pass without candidate pairs has no contact this step: no exchange, no barriers
                __syncthreads();  // everyone finished reading the previous footprints
                if (work) {  // snapshot of the post-move state: the contact flags and the response both read it
                    nb[slot].r = vehicle_rect(P, St);
                    CBody c;
                    c.ox = B.pos.x; c.oy = B.pos.y; c.vx = B.v.x; c.vy = B.v.y; c.w = B.w.z; c.pad = 0.0f;
                    c.im = moves ? 1.0f / P[VP_MASS] : 0.0f;
                    c.ii = moves ? 1.0f / (P[VP_MASS] / 12.0f * (P[VP_WIDTH] * P[VP_WIDTH] + P[VP_LENGTH] * P[VP_LENGTH])) : 0.0f;
                    cd[slot] = c;
                }
                for (int k = ko0 + tid; k < ko1; k += W) {  // kinematic movers (traffic_participants/pedestrian.py:67-95)
                    first_all[(k / O) * ofs + k % O] = 0x7fffffff;
                    float* Ob = obj_all + (size_t)k * OBJ_F;
                    if (Ob[OB_KIND] == 3.0f) { Ob[OB_X] += Ob[OB_VX] * cfg.dt; Ob[OB_Y] += Ob[OB_VY] * cfg.dt; }
                }
                __syncthreads();
                if (cand_lo | cand_hi) {
                    const int f = contact_pass(nb, sobj, S, slot, nb[slot].r, cand_lo, cand_hi, obj_first, true, false,
                                               contact_tab != nullptr ? touch_all + 2 * (le * S + slot) : nullptr);
                    flags |= f & ~FL_TOUCH;
#ifndef MD_NO_RESPONSE
                    if (moves && (f & FL_TOUCH)) {  // something overlaps: push apart (registers only;
                        float d[5];                                           // the snapshot stays as it is for the others)
                        contact_response(nb, cd, sobj, S, slot, cand_lo, cand_hi, d);
                        B.v.x += d[0]; B.v.y += d[1]; B.w.z += d[2];
                        B.pos.x += d[3]; B.pos.y += d[4];
                        St[VS_POS] = B.pos.x; St[VS_POS + 1] = B.pos.y;
                    }
#endif
                }
                if (block_obj) {  // COST_ONCE needs the claims of every vehicle before anyone reads them
                    __syncthreads();
                    if (cand_lo | cand_hi)
                        flags |= contact_pass(nb, sobj, S, slot, nb[slot].r, cand_lo, cand_hi, obj_first, false, true) & ~FL_TOUCH;
                    __syncthreads();
                    for (int k = ko0 + tid; k < ko1; k += W)
                        if (first_all[(k / O) * ofs + k % O] != 0x7fffffff) obj_all[(size_t)k * OBJ_F + OB_CRASHED] = 1.0f;
                }
            } else if (contacts) {
                for (int k = ko0 + tid; k < ko1; k += W) {  // nobody looks: the pedestrians still walk, in the same increments
                    float* Ob = obj_all + (size_t)k * OBJ_F;
                    if (Ob[OB_KIND] == 3.0f) { Ob[OB_X] += Ob[OB_VX] * cfg.dt; Ob[OB_Y] += Ob[OB_VY] * cfg.dt; }
                }
            }
        }
#ifdef MD_PHASE_CLK
        __syncthreads();
        clk_mark(1, 4);
#endif
        if (work) {
            if (moves) {
                St[VS_VEL] = B.v.x; St[VS_VEL + 1] = B.v.y; St[VS_VEL + 2] = B.v.z;
                St[VS_ANGVEL] = B.w.x; St[VS_ANGVEL + 1] = B.w.y; St[VS_ANGVEL + 2] = B.w.z;
                store16(A.veh_s + (size_t)g * VEH_S, St);
            }
            A.veh_i[(size_t)g * VEH_I + VI_FLAGS] = flags;
        }
        if (work_list != nullptr) {
            // the vehicles engine.after_step will have work for (alive and active), appended to the step's global work list:
            // one atomic per warp.  k_scan runs a warp per entry; the order of the entries is irrelevant.
            const bool emit = work && tid < n_front;
            const unsigned em = __ballot_sync(0xffffffffu, emit);
            if (em) {
                const int lane = tid & 31, leader = __ffs(em) - 1;
                unsigned base = 0;
                if (lane == leader) base = atomicAdd(work_count, (unsigned)__popc(em));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (emit) work_list[base + __popc(em & ((1u << lane) - 1u))] = g;
            }
        }
#ifdef MD_PHASE_CLK
        if (g_phase_clk != nullptr && tid == 0 && blockIdx.x < CLK_CTAS) {
            unsigned long long* row = g_phase_clk + ((size_t)1 * CLK_CTAS + blockIdx.x) * 16;
            row[8] += 1ull; row[9] += (unsigned long long)block_any; row[10] += (unsigned long long)(n_front + n_back); row[11] += (unsigned long long)n_front;
        }
#endif
        le0 = le1;
        if (le0 < n_env) __syncthreads();   // the next pass reuses the list and its counters
    }
    if (contact_tab != nullptr) {   // one 16-byte row per slot: the bodies touched during this step (zero for empty slots)
        __syncthreads();
        for (int row = tid; row < n_rows; row += W) {
            const unsigned long long a = touch_all[2 * row], b = touch_all[2 * row + 1];
            contact_tab[(size_t)env0 * S + row] = make_uint4((unsigned)a, (unsigned)(a >> 32), (unsigned)b, (unsigned)(b >> 32));
        }
    }
    if (contacts && O > 0) {
        __syncthreads();
        for (int k = tid; k < n_env * O; k += W) {  // pedestrians turn around at the ends of their crossing (peds.py)
            float* Ob = obj_all + (size_t)k * OBJ_F;
            if (Ob[OB_KIND] != 3.0f || Ob[OB_B] <= 0.0f) continue;
            const float speed = sqrtf(Ob[OB_VX] * Ob[OB_VX] + Ob[OB_VY] * Ob[OB_VY]);
            Ob[OB_HEADING] -= speed * (cfg.dt * (float)n_sub);
            if (Ob[OB_HEADING] <= 0.0f) { Ob[OB_VX] = -Ob[OB_VX]; Ob[OB_VY] = -Ob[OB_VY]; Ob[OB_HEADING] += Ob[OB_B]; }
        }
        __syncthreads();
        for (int k = tid; k < n_env * O * OBJ_F; k += W) A.obj_f[(size_t)env0 * O * OBJ_F + k] = obj_all[k];
    }
#ifdef MD_PHASE_CLK
    __syncthreads();
    clk_mark(1, 13);
#endif
}

// ---- env.reset state restore of one slot row (snapshot -> live arrays) -----------------------------------------
// `full`: the env may hold another scenario than the snapshot's (scenario-bank draws rewrite vehicle parameters, routes
// and trigger roads), so those rows are restored as well
__device__ __forceinline__ void restore_row(const MdConfig& cfg, const MdArrays& A, const Snapshot& snap, long long g, bool full = false) {
    const int S = cfg.slots_per_env, O = cfg.objs_per_env;
    const int env = (int)(g / S), slot = (int)(g - (long long)env * S);
    const float4* s4 = reinterpret_cast<const float4*>(snap.veh_s + (size_t)g * VEH_S);
    const float4* c4 = reinterpret_cast<const float4*>(snap.veh_c + (size_t)g * VEH_C);
    const int4* i4 = reinterpret_cast<const int4*>(snap.veh_i + (size_t)g * VEH_I);
    float4* ds = reinterpret_cast<float4*>(A.veh_s + (size_t)g * VEH_S);
    float4* dc = reinterpret_cast<float4*>(A.veh_c + (size_t)g * VEH_C);
    int4* di = reinterpret_cast<int4*>(A.veh_i + (size_t)g * VEH_I);
#pragma unroll
    for (int k = 0; k < 4; k++) { ds[k] = s4[k]; dc[k] = c4[k]; di[k] = i4[k]; }
    const float4* d4 = reinterpret_cast<const float4*>(snap.veh_idm + (size_t)g * VEH_IDM);
    float4* dd = reinterpret_cast<float4*>(A.veh_idm + (size_t)g * VEH_IDM);
    dd[0] = d4[0]; dd[1] = d4[1];
    for (int k = 0; k < NAVI_DIM; k++) A.veh_navi[(size_t)g * NAVI_DIM + k] = snap.veh_navi[(size_t)g * NAVI_DIM + k];
    for (int k = slot; k < O * OBJ_F; k += S) A.obj_f[(size_t)env * O * OBJ_F + k] = snap.obj_f[(size_t)env * O * OBJ_F + k];
    if (slot == 0)
        for (int k = 0; k < ENV_I; k++) A.env_i[env * ENV_I + k] = snap.env_i[env * ENV_I + k];
    if (full) {
        const float4* p4 = reinterpret_cast<const float4*>(snap.veh_p + (size_t)g * VEH_P);
        float4* dp = reinterpret_cast<float4*>(const_cast<float*>(A.veh_p) + (size_t)g * VEH_P);
#pragma unroll
        for (int k = 0; k < VEH_P / 4; k++) dp[k] = p4[k];
        if (slot == 0) {
            int* trig = const_cast<int*>(A.env_trigger);
            for (int k = 0; k < TRIGGER_MAX; k++) trig[env * TRIGGER_MAX + k] = snap.env_trigger[env * TRIGGER_MAX + k];
        }
    }
    if (full || (cfg.is_multi_agent && slot < cfg.agents_per_env) || cfg.traffic_mode != 0) {  // respawns rewrite the slot's route
        const int4* r4 = reinterpret_cast<const int4*>(snap.veh_route + (size_t)g * ROUTE_MAX);
        const int4* q4 = reinterpret_cast<const int4*>(snap.veh_rroad + (size_t)g * ROUTE_MAX);
        int4* dr = reinterpret_cast<int4*>(A.veh_route + (size_t)g * ROUTE_MAX);
        int4* dq = reinterpret_cast<int4*>(A.veh_rroad + (size_t)g * ROUTE_MAX);
#pragma unroll
        for (int k = 0; k < ROUTE_MAX / 4; k++) { dr[k] = r4[k]; dq[k] = q4[k]; }
    }
}
// auto-reset as a row copy: the post-reset snapshot already holds the localised state, the body rows and the state part
// of the reset observation (md_reset records them after a full reset)
__global__ void k_restore_post(MdConfig cfg, MdArrays A, Snapshot post, const float* __restrict__ post_body,
                               const float* __restrict__ post_obs, float* __restrict__ body_tab, float* __restrict__ obs,
                               const uint8_t* __restrict__ env_mask) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int S = cfg.slots_per_env;
    if (g >= (long long)cfg.n_envs * S) return;
    const int env = (int)(g / S), slot = (int)(g - (long long)env * S);
    if (env_mask[env] == 0) return;
    restore_row(cfg, A, post, g);
    const float4* b4 = reinterpret_cast<const float4*>(post_body + (size_t)g * BODY_ROW);
    float4* d4 = reinterpret_cast<float4*>(body_tab + (size_t)g * BODY_ROW);
#pragma unroll
    for (int k = 0; k < BODY_ROW / 4; k++) d4[k] = b4[k];
    if (slot < cfg.agents_per_env) {
        const size_t a = (size_t)env * cfg.agents_per_env + slot;
        for (int k = 0; k < OBS_STATE(cfg); k++) obs[a * (size_t)OBS_DIM(cfg) + k] = post_obs[a * OBS_STATE(cfg) + k];
        // the tollgate env's two toll floats: no spawn road lies inside the toll block, so a reset observation carries [0, 0]
        for (int k = 0; k < OBS_TOLL(cfg); k++) obs[(a + 1) * (size_t)OBS_DIM(cfg) - 1 - k] = 0.0f;
    }
}
// ---- scenario resampling at reset (BaseEnv.reset(seed=None) -> _reset_global_seed draws a scenario, envs/base_env.py:
// 886-891).  The scenario BANK is a second handle holding one env per scenario of the library, fully reset: its
// post-reset snapshot is the source of the rows.  k_restore_bank draws a scenario for every finished env (a counter hash
// instead of numpy's generator) and copies that scenario's rows into the env - everything that differs between
// two scenarios: vehicle parameters, routes, trigger roads, state, objects, env row (map id, seed), body rows and the
// state part of the reset observation.
struct BankView {
    Snapshot post;
    const float* body; const float* obs; const float* veh_p; const int* env_trigger;
    int n;
};
struct FusedBank { BankView B; uint32_t seed; const uint32_t* pass_ctr; int on; };   // k_post's fused auto-reset (bank draws)
__global__ void k_bump(uint32_t* __restrict__ pass_ctr, uint32_t d_bank, uint32_t d_noise) { pass_ctr[0] += d_bank; pass_ctr[1] += d_noise; }
// one slot row (plus its share of the env-level rows) of a finished env <- the scenario drawn for it
__device__ __forceinline__ void restore_bank_row(const MdConfig& cfg, const MdArrays& A, const BankView& B, uint32_t seed, uint32_t pass,
                                                 float* __restrict__ body_tab, float* __restrict__ obs, int env, int slot) {
    const int S = cfg.slots_per_env, O = cfg.objs_per_env, NA = cfg.agents_per_env;
    const long long g = (long long)env * S + slot;
    // the draw: a counter hash of (seed, env, reset pass) - every thread of the env computes the same scenario
    uint32_t x = seed * 0x9E3779B9u + (uint32_t)(env + cfg.env_base) * 0x85EBCA6Bu + pass * 0xC2B2AE35u + 0x165667B1u;
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    const int scn = (int)(x % (uint32_t)B.n);
    const size_t b = (size_t)scn * S + slot;   // source row in the bank
    // The copies are latency bound (a few finished envs per step, rows cold in DRAM): every thread first issues all the
    // loads of a batch, then stores, and the env-level rows are strided over the env's threads instead of being one
    // thread's chain of dependent round trips.
#define LOAD4(buf, src, n4)                                                                            \
    {                                                                                                  \
        const int4* s4_ = reinterpret_cast<const int4*>(src);                                          \
        _Pragma("unroll") for (int k_ = 0; k_ < (n4); k_++) buf[k_] = __ldg(s4_ + k_);                 \
    }
#define STORE4(dst, buf, n4)                                                                           \
    {                                                                                                  \
        int4* d4_ = reinterpret_cast<int4*>(dst);                                                      \
        _Pragma("unroll") for (int k_ = 0; k_ < (n4); k_++) d4_[k_] = buf[k_];                         \
    }
    {
        int4 s_[VEH_S / 4], c_[VEH_C / 4], i_[VEH_I / 4], d_[VEH_IDM / 4], p_[VEH_P / 4];
        LOAD4(s_, B.post.veh_s + b * VEH_S, VEH_S / 4)
        LOAD4(c_, B.post.veh_c + b * VEH_C, VEH_C / 4)
        LOAD4(i_, B.post.veh_i + b * VEH_I, VEH_I / 4)
        LOAD4(d_, B.post.veh_idm + b * VEH_IDM, VEH_IDM / 4)
        LOAD4(p_, B.veh_p + b * VEH_P, VEH_P / 4)
        STORE4(A.veh_s + (size_t)g * VEH_S, s_, VEH_S / 4)
        STORE4(A.veh_c + (size_t)g * VEH_C, c_, VEH_C / 4)
        STORE4(A.veh_i + (size_t)g * VEH_I, i_, VEH_I / 4)
        STORE4(A.veh_idm + (size_t)g * VEH_IDM, d_, VEH_IDM / 4)
        STORE4(const_cast<float*>(A.veh_p) + (size_t)g * VEH_P, p_, VEH_P / 4)
    }
    {
        int4 r_[ROUTE_MAX / 4], q_[ROUTE_MAX / 4], y_[BODY_ROW / 4];
        float nav_[NAVI_DIM];
        LOAD4(r_, B.post.veh_route + b * ROUTE_MAX, ROUTE_MAX / 4)
        LOAD4(q_, B.post.veh_rroad + b * ROUTE_MAX, ROUTE_MAX / 4)
        LOAD4(y_, B.body + b * BODY_ROW, BODY_ROW / 4)
#pragma unroll
        for (int k = 0; k < NAVI_DIM; k++) nav_[k] = __ldg(B.post.veh_navi + b * NAVI_DIM + k);
        STORE4(A.veh_route + (size_t)g * ROUTE_MAX, r_, ROUTE_MAX / 4)
        STORE4(A.veh_rroad + (size_t)g * ROUTE_MAX, q_, ROUTE_MAX / 4)
        STORE4(body_tab + (size_t)g * BODY_ROW, y_, BODY_ROW / 4)
#pragma unroll
        for (int k = 0; k < NAVI_DIM; k++) A.veh_navi[(size_t)g * NAVI_DIM + k] = nav_[k];
    }
#undef LOAD4
#undef STORE4
    for (int k = slot; k < O * OBJ_F; k += S) A.obj_f[(size_t)env * O * OBJ_F + k] = B.post.obj_f[(size_t)scn * O * OBJ_F + k];
    int* trig = const_cast<int*>(A.env_trigger);
    for (int k = slot; k < ENV_I + TRIGGER_MAX; k += S) {   // the env row and its trigger roads
        if (k < ENV_I) A.env_i[env * ENV_I + k] = B.post.env_i[scn * ENV_I + k];
        else trig[env * TRIGGER_MAX + (k - ENV_I)] = B.env_trigger[scn * TRIGGER_MAX + (k - ENV_I)];
    }
    const int n_state = NA * OBS_STATE(cfg);                // the state part of the agents' reset observation
    for (int k = slot; k < n_state; k += S) {
        const int ag = k / OBS_STATE(cfg), col = k - ag * OBS_STATE(cfg);
        obs[((size_t)env * NA + ag) * (size_t)OBS_DIM(cfg) + col] = B.obs[((size_t)scn * NA + ag) * OBS_STATE(cfg) + col];
    }
}
__global__ void k_restore_bank(MdConfig cfg, MdArrays A, BankView B, uint32_t seed, const uint32_t* __restrict__ pass_ctr,
                               float* __restrict__ body_tab, float* __restrict__ obs, const uint8_t* __restrict__ env_mask) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int S = cfg.slots_per_env;
    if (g >= (long long)cfg.n_envs * S) return;
    const int env = (int)(g / S), slot = (int)(g - (long long)env * S);
    if (env_mask != nullptr && env_mask[env] == 0) return;
    restore_bank_row(cfg, A, B, seed, pass_ctr[0], body_tab, obs, env, slot);
}
__global__ void k_restore(MdConfig cfg, MdArrays A, Snapshot snap, const uint8_t* __restrict__ env_mask, int full) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)cfg.n_envs * cfg.slots_per_env) return;
    if (env_mask != nullptr && env_mask[g / cfg.slots_per_env] == 0) return;
    restore_row(cfg, A, snap, g, full != 0);
}

// ---- k_post: engine.after_step + _get_step_return (base_vehicle.py:234-271; envs/base_env.py:586-623) -----------
// MODE_RESET: the reset-time variant (envs/base_env.py:560-584) for the envs selected by env_mask.
// A CTA owns `epb` envs and runs POST_WORKERS threads.  Phase 1: the threads sweep the epb x S slot rows (coalesced),
// publish every alive vehicle's footprint in shared memory and append the vehicles that have after_step work (the active
// ones: the agent and the triggered traffic) to a work list.  Phase 2: one thread per listed vehicle - so the warps are
// dense with vehicles that all run the same code, instead of one thread per slot with most lanes idle.  Phase 3
// (respawn / hybrid traffic only): vehicles that left the lanes are replaced, tape rows handed out in slot order.
struct __align__(16) Fp { Rect r; int alive; int mark; };   // footprint record of one slot (32 B)
#ifndef POST_WORKERS
#define POST_WORKERS 128
#endif
#define POST_TEAM 4   // lanes per vehicle in phase 2a (2 when the CTA holds more vehicles than worker threads)
__host__ __device__ inline size_t post_scan_offset(int S, int O, int epb) {   // LocScan rows follow the older tables, 16-byte aligned
    size_t b = (sizeof(Fp) + sizeof(int)) * (size_t)S * epb + sizeof(float) * OBJ_F * (size_t)O * epb + 16 + sizeof(int) * (size_t)epb;
    return (b + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t post_smem_bytes(int S, int O, int epb) {
    return post_scan_offset(S, O, epb) + sizeof(LocScan) * (size_t)S * epb;
}
__device__ __forceinline__ int fp_contacts(const Fp* fp, const float* obj, int S, int O, int slot, const Rect& r) {
    int flags = 0;  // BaseVehicle._state_check, dynamic world part (component/vehicle/base_vehicle.py:735-742): no latch
    for (int k = 0; k < S; k++) {
        if (k == slot || !fp[k].alive) continue;
        if (rect_rect(r, fp[k].r)) flags |= FL_CRASH_VEHICLE;
    }
    for (int k = 0; k < O; k++) {
        const float* Ob = obj + k * OBJ_F;
        if (Ob[OB_KIND] < 0.0f) continue;
        bool hit;
        if (OB_IS_BOX(Ob[OB_KIND])) { Rect ro = object_rect(Ob); hit = rect_rect(r, ro); }
        else hit = rect_circle(r, Ob[OB_X], Ob[OB_Y], Ob[OB_A]);
        if (hit) flags |= Ob[OB_KIND] == 3.0f ? FL_CRASH_HUMAN : (Ob[OB_KIND] == 4.0f ? FL_CRASH_BUILDING : FL_CRASH_OBJECT);
    }
    return flags;
}
__device__ __forceinline__ int fp_contacts_team(const Fp* fp, const float* obj, int S, int O, int slot, const Rect& r, int sub, int T) {
    int flags = 0;  // fp_contacts with the bodies strided over the lanes of a team
    for (int k = sub; k < S; k += T) {
        if (k == slot || !fp[k].alive) continue;
        if (rect_rect(r, fp[k].r)) flags |= FL_CRASH_VEHICLE;
    }
    for (int k = sub; k < O; k += T) {
        const float* Ob = obj + k * OBJ_F;
        if (Ob[OB_KIND] < 0.0f) continue;
        bool hit;
        if (OB_IS_BOX(Ob[OB_KIND])) { Rect ro = object_rect(Ob); hit = rect_rect(r, ro); }
        else hit = rect_circle(r, Ob[OB_X], Ob[OB_Y], Ob[OB_A]);
        if (hit) flags |= Ob[OB_KIND] == 3.0f ? FL_CRASH_HUMAN : (Ob[OB_KIND] == 4.0f ? FL_CRASH_BUILDING : FL_CRASH_OBJECT);
    }
    return flags;
}
// BaseVehicle.after_step for one vehicle (component/vehicle/base_vehicle.py:234-271): localisation, state check
// against the static world and the env's other bodies, side distances, energy.  `fp` = the env's footprints.
__device__ __forceinline__ void after_step_vehicle(const MapView& m, const float* St, float* C, int* I,
                                                   const int* __restrict__ route, const int* __restrict__ rroad, float* navi,
                                                   const Fp* fp, const float* sobj, int S, int O, int slot, const Rect& r,
                                                   const LocScan* pre = nullptr) {
    localise(m, St, I, route, rroad, navi, pre);
    int flags = I[VI_FLAGS];
    if (pre != nullptr) flags |= pre->static_flags | pre->contact_flags;   // a team already ran the three scans (phase 2a)
    else {
        state_check_static(m, r, flags);
        flags |= fp_contacts(fp, sobj, S, O, slot, r);
    }
    I[VI_FLAGS] = flags;
    int cur_road = rroad[I[VI_CKPT0]];
    int cur_first = m.road_i[cur_road * ROAD_I + RI_FIRST], cur_n = m.road_i[cur_road * ROAD_I + RI_N];
    float lon, lat;
    lane_local(m.lane_f + cur_first * LANE_F, St[VS_POS], St[VS_POS + 1], lon, lat);
    float lane_w = m.lane_f[I[VI_LANE] * LANE_F + LF_WIDTH];
    float to_left = lat + lane_w / 2.0f;
    float to_right = lane_w * (float)cur_n - to_left;
    C[VC_DIST_L] = to_left; C[VC_DIST_R] = to_right;
    if (to_right < 0.0f || to_left < 0.0f) I[VI_FLAGS] |= FL_OUT_OF_ROUTE;
    float dx = C[VC_LAST_X] - St[VS_POS], dy = C[VC_LAST_Y] - St[VS_POS + 1];
    float dist_km = sqrtf(dx * dx + dy * dy) / 1000.0f;
    float speed_kmh = sqrtf(St[VS_VEL] * St[VS_VEL] + St[VS_VEL + 1] * St[VS_VEL + 1]) * 3.6f;
    float step_energy = 3.25f * md_expf(0.01f * speed_kmh) * dist_km / 100.0f * 1000.0f;
    C[VC_STEP_ENERGY] = step_energy;
    C[VC_ENERGY] += step_energy;
}

// ---- k_scan: the two table-walking scans of a vehicle's after_step - candidate lanes of update_localization, line /
// sidewalk items of _state_check - with a WARP per vehicle of the step's work list (k_dyn appends the alive + active
// vehicles).  A GPU holds only ~30 k such vehicles at BASELINE cfg2: as teams inside k_post's CTAs (128 registers per
// thread, 4 CTAs per SM) they took two rounds of 4-lane teams, 50 of k_post's 100 us; here every vehicle gets 32 lanes at
// once, the grid is balanced over the SMs whatever the envs hold, and the register budget is the scans' own.
#define SCAN_WARPS 8
// TS lanes per vehicle (a warp holds 32 / TS vehicles).  The loads are ordered so that the independent chains of a vehicle
// overlap: (state, checkpoints, map id) -> (map tables' offsets, route roads) -> (cell ranges of BOTH scans, road rows) ->
// (grid records of the candidates AND of the first items) -> lane rows of the surviving candidates.
template <int TS, int MB>
__global__ void __launch_bounds__(SCAN_WARPS * 32, MB)
k_scan(MdConfig cfg, MdArrays A, MapAccel X, const int* __restrict__ work_list, unsigned int* __restrict__ work_count,
       LocScan* __restrict__ scan_tab) {
    const int lane = threadIdx.x & 31, tl = threadIdx.x & (TS - 1), tbase = lane & ~(TS - 1), S = cfg.slots_per_env;
    const unsigned tmask = TS == 32 ? 0xffffffffu : (((1u << TS) - 1u) << tbase);
    clk_mark(3, 0);
    // The vehicles differ a lot in cost (a cell of an intersection holds several times the lanes and line boxes of a cell of
    // a straight road), so the warps do not stride over the list: each fetches its next 32 / TS vehicles from a cursor
    // (one atomic per warp and fetch - a few thousand per launch).
    const int n = (int)work_count[0];
#ifdef MD_PHASE_CLK
    bool first = true;
#endif
    for (;;) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(&work_count[1], 32u / TS);
        base = __shfl_sync(0xffffffffu, base, 0);
        if ((int)base >= n) break;
        const int j = (int)base + lane / TS;
        if (j >= n) continue;   // the whole team skips: the masks below only name this team's lanes
        const int g = __ldg(work_list + j), env = g / S;
        const float* St = A.veh_s + (size_t)g * VEH_S;
        float P4[4], Sv[VEH_S];
        load16(Sv, St);
        const float4 p0 = *reinterpret_cast<const float4*>(A.veh_p + (size_t)g * VEH_P);
        P4[0] = p0.x; P4[1] = p0.y; P4[2] = p0.z; P4[3] = p0.w;  // type, length, width, height
        const MapView m = map_view(A, A.env_i[env * ENV_I + EI_MAP], &X);
        const LocCtx c = loc_ctx(m, Sv, A.veh_i + (size_t)g * VEH_I, A.veh_rroad + (size_t)g * ROUTE_MAX);
        const Rect r = vehicle_rect(P4, Sv);
        // cell ranges: the lane grid cell under the centre, the static grid cells under the bounding circle (first four)
        int k0, k1;
        loc_cell(m, c.px, c.py, k0, k1);
        const float rad = sqrtf(r.hu * r.hu + r.hv * r.hv);
        int x0 = (int)floorf((r.cx - rad - m.gx0) / m.cell), x1 = (int)floorf((r.cx + rad - m.gx0) / m.cell);
        int y0 = (int)floorf((r.cy - rad - m.gy0) / m.cell), y1 = (int)floorf((r.cy + rad - m.gy0) / m.cell);
        x0 = max(x0, 0); y0 = max(y0, 0); x1 = min(x1, m.nx - 1); y1 = min(y1, m.ny - 1);
        const int nxc = x1 - x0 + 1, nc = nxc > 0 && y1 >= y0 ? nxc * (y1 - y0 + 1) : 0;
        int cs[4], cn[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            cs[q] = 0; cn[q] = 0;
            if (q < nc) {
                const int cell = (y0 + q / nxc) * m.nx + x0 + q % nxc;
                cs[q] = __ldg(m.gs + cell); cn[q] = __ldg(m.gs + cell + 1) - cs[q];
            }
        }
        LocScan sc;
        loc_scan_init(sc);
        int flags = 0;
#ifdef MD_PHASE_CLK
        if (first && cn[0] >= 0 && k1 >= 0) clk_mark(3, 1);
#endif
        // ---- candidate lanes (ascending; a survivor per lane)
        for (int kb = k0; kb < k1; kb += TS) {
            const int kk = kb + tl;
            float4 a = make_float4(0.0f, 0.0f, 0.0f, 0.0f), b = a;
            bool in = false;
            if (kk < k1) {
                a = __ldg(m.lrec + 2 * kk); b = __ldg(m.lrec + 2 * kk + 1);
                in = !(c.px < a.y || c.py < a.z || c.px > a.w || c.py > b.x);
            }
            const unsigned surv = (__ballot_sync(tmask, in) & tmask) >> tbase;
            const int src = (int)__fns(surv, 0, tl + 1);          // team lane holding the (tl + 1)-th survivor, or -1
            const int sl = src >= 0 && src < TS ? src : 0;
            const int l = __float_as_int(__shfl_sync(tmask, a.x, sl, TS));
            const int ho = __float_as_int(__shfl_sync(tmask, b.y, sl, TS)), hn = __float_as_int(__shfl_sync(tmask, b.z, sl, TS));
            if (tl < __popc(surv)) loc_candidate_in_bb(m, c, l, ho, hn, sc);
        }
#ifdef MD_PHASE_CLK
        if (first && sc.d_any >= 0.0f) clk_mark(3, 2);
#endif
        // ---- line / sidewalk items of the cells, four cells at a time, their item ranges concatenated
        for (int cb = 0; cb < nc; cb += 4) {
            if (cb > 0) {
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int ci = cb + q;
                    cs[q] = 0; cn[q] = 0;
                    if (ci < nc) {
                        const int cell = (y0 + ci / nxc) * m.nx + x0 + ci % nxc;
                        cs[q] = __ldg(m.gs + cell); cn[q] = __ldg(m.gs + cell + 1) - cs[q];
                    }
                }
            }
            const int p1 = cn[0], p2 = p1 + cn[1], p3 = p2 + cn[2], tot = p3 + cn[3];
#pragma unroll 1
            for (int i = tl; i < tot; i += TS) {
                const int k = i < p1 ? cs[0] + i : (i < p2 ? cs[1] + (i - p1) : (i < p3 ? cs[2] + (i - p2) : cs[3] + (i - p3)));
                const int it = __ldg(m.gi + k);
                const float4 h = __ldg(m.irec + 2 * k), u = __ldg(m.irec + 2 * k + 1);
                if (it < m.n_lines) {
                    const int kind = (int)h.w;
                    const int bit = kind == 0 ? FL_ON_WHITE : (kind == 1 ? FL_ON_YELLOW : FL_ON_BROKEN);
                    if (flags & bit) continue;
                    const float dx = h.x - r.cx, dy = h.y - r.cy, rr = rad + h.z + LINE_HALF_W + 0.01f;
                    if (dx * dx + dy * dy > rr * rr) continue;
                    Rect lr;
                    lr.cx = h.x; lr.cy = h.y; lr.ux = u.x; lr.uy = u.y; lr.hu = h.z; lr.hv = LINE_HALF_W;
                    if (rect_rect(r, lr)) flags |= bit;
                } else {
                    if (flags & FL_CRASH_SIDEWALK) continue;
                    float qd[8] = {h.x, h.y, h.z, h.w, u.x, u.y, u.z, u.w};
                    if (rect_quad(r, qd)) flags |= FL_CRASH_SIDEWALK;
                }
            }
        }
        sc.static_flags = flags;
#ifdef MD_PHASE_CLK
        if (first && flags >= 0) clk_mark(3, 3);
#endif
#define TEAM_MIN(D, L, LON, LAT)                                                                         \
        {                                                                                                \
            const float od_ = __shfl_xor_sync(tmask, D, off, TS), olon_ = __shfl_xor_sync(tmask, LON, off, TS), \
                        olat_ = __shfl_xor_sync(tmask, LAT, off, TS);                                    \
            const int ol_ = __shfl_xor_sync(tmask, L, off, TS);                                          \
            if (od_ < D || (od_ == D && ol_ >= 0 && (L < 0 || ol_ < L))) { D = od_; L = ol_; LON = olon_; LAT = olat_; } \
        }
#pragma unroll
        for (int off = TS >> 1; off > 0; off >>= 1) {
            TEAM_MIN(sc.d_any, sc.best_any, sc.lon_any, sc.lat_any)
            TEAM_MIN(sc.d_cur, sc.best_cur, sc.lon_cur, sc.lat_cur)
            TEAM_MIN(sc.d_next, sc.best_next, sc.lon_next, sc.lat_next)
            sc.on_lane |= __shfl_xor_sync(tmask, sc.on_lane, off, TS);
            sc.static_flags |= __shfl_xor_sync(tmask, sc.static_flags, off, TS);
        }
#undef TEAM_MIN
        if (tl == 0) scan_tab[g] = sc;
#ifdef MD_PHASE_CLK
        if (first) clk_mark(3, 4);
        first = false;
#endif
    }
#ifdef MD_PHASE_CLK
    __syncthreads();
    clk_mark(3, 13);
#endif
}

__global__ void __maxnreg__(POST_REGS)
k_post(MdConfig cfg, MdArrays A, int mode, int epb, StepOut out, float* __restrict__ body_tab,
       const uint8_t* __restrict__ env_mask, uint8_t* __restrict__ done_mask, Snapshot snap, int team_pref, MapAccel X,
       const LocScan* __restrict__ scan_tab, FusedBank FB) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = cfg.slots_per_env, O = cfg.objs_per_env, NA = cfg.agents_per_env;
    const int n_rows = epb * S, env0 = blockIdx.x * epb;
    clk_mark(2, 0);
    if (mode & MODE_RESTORE) {
        // env.reset of the masked envs, part 1: snapshot -> live arrays (what k_restore does, fused in so that an
        // auto-reset costs one launch).  CTAs without a finished env leave at once.
        int any = 0;
        for (int le = 0; le < epb; le++) any |= (env0 + le < cfg.n_envs) && env_mask[env0 + le];
        if (!any) return;
        for (int v = threadIdx.x; v < n_rows; v += blockDim.x) {
            const int le = v / S, env = env0 + le;
            if (env >= cfg.n_envs || !env_mask[env]) continue;
            restore_row(cfg, A, snap, (long long)env * S + (v - le * S));
        }
        __syncthreads();
    }
    Fp* fp_all = reinterpret_cast<Fp*>(smem_raw);
    float* obj_all = reinterpret_cast<float*>(smem_raw + sizeof(Fp) * (size_t)n_rows);
    int* list = reinterpret_cast<int*>(smem_raw + sizeof(Fp) * (size_t)n_rows + sizeof(float) * OBJ_F * (size_t)O * epb);
    int* n_list = list + n_rows;
    int* ctr_base = n_list + 4;   // per env: EI_RNG before this step's respawns
    if (threadIdx.x == 0) *n_list = 0;
    for (int le = threadIdx.x; le < epb; le += blockDim.x)
        ctr_base[le] = env0 + le < cfg.n_envs ? A.env_i[(env0 + le) * ENV_I + EI_RNG] : 0;
    __syncthreads();
    clk_mark(2, 1);
    // ---- phase 1: footprints + work list (row v = env_local * S + slot, the order of the global rows)
    for (int v = threadIdx.x; v < n_rows; v += blockDim.x) {
        const int le = v / S, slot = v - le * S, env = env0 + le;
        Fp& f = fp_all[v];
        f.alive = 0; f.mark = 0;
        if (env >= cfg.n_envs || (env_mask != nullptr && env_mask[env] == 0)) continue;
        const size_t g = (size_t)env * S + slot;
        const int4 i0 = *reinterpret_cast<const int4*>(A.veh_i + g * VEH_I);  // kind, alive, active, trigger
        const int kind = i0.x, alive = i0.y, active = i0.z;
        if (kind == 0) continue;
        const bool has_work = alive && ((mode & MODE_RESET) ? (kind == 1 || kind == 2) : ((mode & MODE_POST) && active));
        if (alive) {
            float P4[4], St[VEH_S];
            const float4 p0 = *reinterpret_cast<const float4*>(A.veh_p + g * VEH_P);
            P4[0] = p0.x; P4[1] = p0.y; P4[2] = p0.z; P4[3] = p0.w;  // type, length, width, height
            load16(St, A.veh_s + g * VEH_S);
            f.r = vehicle_rect(P4, St);
            f.alive = 1;
            if (!has_work) write_body_row(body_tab + g * BODY_ROW, P4, St, 1);
        } else body_tab[g * BODY_ROW + 15] = 0.0f;
        if (has_work) list[atomicAdd(n_list, 1)] = v;
        else if ((mode & MODE_OUT) && cfg.is_multi_agent && slot < NA) {
            const size_t a = (size_t)env * NA + slot;  // an empty or wrecked seat produces no transition
            out.reward[a] = 0.0f; out.cost[a] = 0.0f; out.term[a] = 0; out.trunc[a] = 0; out.info_flags[a] = 0;
        }
    }
    for (int k = threadIdx.x; k < epb * O * OBJ_F; k += blockDim.x) {
        const int env = env0 + k / (O * OBJ_F);
        if (env < cfg.n_envs) obj_all[k] = A.obj_f[(size_t)env0 * O * OBJ_F + k];
    }
    __syncthreads();
    clk_mark(2, 2);
    // ---- phase 2a: the three scans of a vehicle's after_step that walk tables - candidate lanes of update_localization,
    // line / sidewalk items of _state_check, the env's bodies for the post-step contacts - spread over a team of 2 - 4
    // lanes per vehicle.  A GPU holds only ~200 active vehicles per SM at BASELINE cfg2: one thread per vehicle leaves 6
    // warps per SM walking long dependent chains; the teams multiply the warps in flight and shorten the chains.
    const int n_work = *n_list;
    LocScan* scan = reinterpret_cast<LocScan*>(smem_raw + post_scan_offset(S, O, epb));
    // team size (measured, gpurun_out/tune.log: k_post at T = 1 / 2 / 4 / 8 / 16 is 0.132 / 0.107 / 0.099 / 0.103 / 0.130 ms at
    // cfg2, 0.115 / 0.105 / 0.115 / 0.136 / 0.184 at cfg3, 0.227 / 0.192 / 0.184 / 0.189 / 0.245 at cfg5): 4 lanes per
    // vehicle while the CTA's vehicles fit the worker threads, else 2 - every round of the team loop pays the dependent
    // loads of map_view / loc_ctx once, so many rounds eat the gain.  MD_POST_TEAM overrides; T = 1 = no team phase.
    int T = team_pref;
    if (T <= 0) T = n_work <= (int)blockDim.x ? POST_TEAM : POST_TEAM / 2;
    const bool use_teams = T > 1;
    if (use_teams) {
        const int sub = threadIdx.x & (T - 1), team = threadIdx.x / T, n_teams = blockDim.x / T;
        const unsigned team_mask = (T >= 32 ? 0xffffffffu : ((1u << T) - 1u)) << ((threadIdx.x & 31) & ~(T - 1));
        for (int j = team; j < n_work; j += n_teams) {
            const int v = list[j], le = v / S, slot = v - le * S, env = env0 + le;
            const size_t g = (size_t)env * S + slot;
            if (scan_tab != nullptr) {
                // k_scan already ran the candidate and static scans of this vehicle: the team adds the post-step contacts
                int cf = fp_contacts_team(fp_all + (size_t)le * S, obj_all + (size_t)le * O * OBJ_F, S, O, slot, fp_all[v].r, sub, T);
                for (int off = T >> 1; off > 0; off >>= 1) cf |= __shfl_xor_sync(team_mask, cf, off);
                if (sub == 0) { LocScan sc = scan_tab[g]; sc.contact_flags = cf; scan[j] = sc; }
                continue;
            }
            const MapView m = map_view(A, A.env_i[env * ENV_I + EI_MAP], &X);
            const LocCtx c = loc_ctx(m, A.veh_s + g * VEH_S, A.veh_i + g * VEH_I, A.veh_rroad + g * ROUTE_MAX);
            LocScan sc;
            loc_scan_init(sc);
            int k0, k1;
            loc_cell(m, c.px, c.py, k0, k1);
            if (j == team) clk_mark(2, 4);
            for (int kk = k0 + sub; kk < k1; kk += T) loc_candidate(m, c, m.lgi[kk], sc);
            if (j == team) clk_mark(2, 5);
            const Rect r = fp_all[v].r;
            sc.static_flags = state_check_static_team(m, r, sub, T);
            if (j == team) clk_mark(2, 6);
            sc.contact_flags = fp_contacts_team(fp_all + (size_t)le * S, obj_all + (size_t)le * O * OBJ_F, S, O, slot, r, sub, T);
            __syncwarp(team_mask);
            if (j == team) clk_mark(2, 7);
#define TEAM_MIN(D, L, LON, LAT)                                                                         \
            {                                                                                            \
                const float od_ = __shfl_xor_sync(team_mask, D, off), olon_ = __shfl_xor_sync(team_mask, LON, off), \
                            olat_ = __shfl_xor_sync(team_mask, LAT, off);                                \
                const int ol_ = __shfl_xor_sync(team_mask, L, off);                                      \
                if (od_ < D || (od_ == D && ol_ >= 0 && (L < 0 || ol_ < L))) { D = od_; L = ol_; LON = olon_; LAT = olat_; } \
            }
            for (int off = T >> 1; off > 0; off >>= 1) {
                TEAM_MIN(sc.d_any, sc.best_any, sc.lon_any, sc.lat_any)
                TEAM_MIN(sc.d_cur, sc.best_cur, sc.lon_cur, sc.lat_cur)
                TEAM_MIN(sc.d_next, sc.best_next, sc.lon_next, sc.lat_next)
                sc.on_lane |= __shfl_xor_sync(team_mask, sc.on_lane, off);
                sc.static_flags |= __shfl_xor_sync(team_mask, sc.static_flags, off);
                sc.contact_flags |= __shfl_xor_sync(team_mask, sc.contact_flags, off);
            }
#undef TEAM_MIN
            if (sub == 0) scan[j] = sc;
        }
    }
    __syncthreads();
    clk_mark(2, 3);
    // ---- phase 2: one thread per vehicle with work
    for (int j = threadIdx.x; j < n_work; j += blockDim.x) {
        const int v = list[j], le = v / S, slot = v - le * S, env = env0 + le;
        const size_t g = (size_t)env * S + slot;
        const Fp* fp = fp_all + (size_t)le * S;
        const float* sobj = obj_all + (size_t)le * O * OBJ_F;
        float P[VEH_P], St[VEH_S], C[VEH_C], navi[NAVI_DIM];
        int I[VEH_I];
        load16i(I, A.veh_i + g * VEH_I);
        load16(P, A.veh_p + g * VEH_P);
        load16(St, A.veh_s + g * VEH_S);
        load16(C, A.veh_c + g * VEH_C);
#pragma unroll
        for (int k = 0; k < NAVI_DIM; k++) navi[k] = A.veh_navi[g * NAVI_DIM + k];
        const bool is_agent = I[VI_KIND] == 1, is_traffic = I[VI_KIND] == 2;
        if (mode & MODE_RESET) latch_before_step(St, C, I);
        const MapView m = map_view(A, A.env_i[env * ENV_I + EI_MAP]);
        const int env_step = A.env_i[env * ENV_I + EI_STEP];
        const int* rroad = A.veh_rroad + g * ROUTE_MAX;
        if (mode & MODE_CLEAR_FLAGS) I[VI_FLAGS] = FL_ON_LANE;
        if (j == 0 && St[0] > -1e30f && C[0] > -1e30f && I[0] > -1000 && navi[0] > -1e30f) clk_mark(2, 8);
        after_step_vehicle(m, St, C, I, A.veh_route + g * ROUTE_MAX, rroad, navi, fp, sobj, S, O, slot, fp[slot].r, use_teams ? scan + j : nullptr);
        if (j == 0 && I[VI_FLAGS] > -1 && C[VC_ENERGY] > -1e30f) clk_mark(2, 9);
        if ((mode & MODE_RESET) && !I[VI_ACTIVE]) I[VI_FLAGS] = FL_ON_LANE;  // reset() ends with _init_step_info
        // traffic_manager.after_step: off-lane traffic leaves the world (manager/traffic_manager.py:94-111); in respawn /
        // hybrid mode phase 3 brings it back as a new vehicle
        if ((mode & MODE_REMOVE) && is_traffic && I[VI_ACTIVE] && !(I[VI_FLAGS] & FL_ON_LANE)) {
            I[VI_ALIVE] = 0; I[VI_ACTIVE] = 0;
            fp_all[v].mark = 2;
        }
        // every agent observes the world as it is after engine.after_step: a vehicle that finishes this step is still
        // visible to the others' lidar (the body row keeps the pre-finish alive flag; k_respawn clears it afterwards)
        const int alive_row = I[VI_ALIVE];
        if ((mode & (MODE_OUT | MODE_RESET)) && is_agent && I[VI_ACTIVE] && slot < NA) {
            const size_t a = (size_t)env * NA + slot;
            agent_outputs(cfg, m, env_step, P, St, C, I, rroad, navi, a, out, (mode & MODE_OUT) ? 3 : 0);
            if ((mode & MODE_OUT) && cfg.is_multi_agent) {
                // MultiAgentMetaDrive._after_vehicle_done -> agent_manager._finish (multi_agent_metadrive.py:153-166,
                // agent_manager.py:115-128): success leaves at once, everything else becomes a static wreck
                const int fl = out.info_flags[a];
                out.info_flags[a] = fl | FL_VALID;
                if (out.term[a] || out.trunc[a]) {
                    I[VI_ACTIVE] = 0;
                    if ((fl & FL_ARRIVE) || cfg.delay_done <= 0) { I[VI_ALIVE] = 0; }
                    else { I[VI_STATIC] = 1; I[VI_DYING] = cfg.delay_done; }
                }
            }
        }
#ifdef MD_PHASE_CLK
        if (j == 0) {
            clk_mark(2, 10);
            if (g_phase_clk != nullptr && blockIdx.x < CLK_CTAS) g_phase_clk[((size_t)2 * CLK_CTAS + blockIdx.x) * 16 + 14] = is_agent ? 1ull : 0ull;
        }
#endif
        if ((mode & MODE_MARK_DONE) && is_agent && slot == 0) {
            const size_t a = (size_t)env * NA;
            done_mask[env] = (out.term[a] || out.trunc[a]) ? 1 : 0;
        }
        store16(A.veh_c + g * VEH_C, C);
        store16i(A.veh_i + g * VEH_I, I);
#pragma unroll
        for (int k = 0; k < NAVI_DIM; k++) A.veh_navi[g * NAVI_DIM + k] = navi[k];
        write_body_row(body_tab + g * BODY_ROW, P, St, alive_row);
        if (j == 0) clk_mark(2, 11);
    }
#ifdef MD_PHASE_CLK
    __syncthreads();
    clk_mark(2, 13);
#endif
    // ---- auto-reset of the envs that finished in this step, fused in (single agent, scenario bank attached): the CTA owns
    // its envs' rows, every read of this step is behind the barrier, so the rows of the drawn scenarios can be copied in
    // right here instead of by a launch of their own (k_restore_bank: 12 us for a copy of ~1 % of the envs)
    if (FB.on) {
        __syncthreads();
        const uint32_t pass = FB.pass_ctr[0];
        for (int v = threadIdx.x; v < n_rows; v += blockDim.x) {
            const int le = v / S, env = env0 + le;
            if (env < cfg.n_envs && done_mask[env]) restore_bank_row(cfg, A, FB.B, FB.seed, pass, body_tab, out.obs, env, v - le * S);
        }
        return;   // trigger-mode worlds only (md_attach_bank): no phase 3
    }
    // ---- phase 3: respawn / hybrid traffic (manager/traffic_manager.py:112-121)
    if (!((mode & MODE_REMOVE) && cfg.traffic_mode != 0)) return;
    __syncthreads();
    for (int j = threadIdx.x; j < n_work; j += blockDim.x) {
        const int v = list[j];
        if (fp_all[v].mark != 2) continue;
        const int le = v / S, slot = v - le * S, env = env0 + le;
        const size_t g = (size_t)env * S + slot;
        const Fp* fp = fp_all + (size_t)le * S;
        const int n_places = A.env_i[env * ENV_I + EI_N_PLACES];
        if (n_places <= 0) continue;
        int rank = 0, total = 0;
        for (int k = 0; k < S; k++) {
            const int lv = fp[k].mark == 2;
            total += lv;
            if (k < slot) rank += lv;
        }
        const uint32_t ctr = (uint32_t)ctr_base[le] + (uint32_t)rank;  // EI_RNG is advanced once, by the last leaver
        float P[VEH_P], St[VEH_S], C[VEH_C], navi[NAVI_DIM];
        int I[VEH_I];
        load16(P, A.veh_p + g * VEH_P);
        const int p = (int)(tape_draw(cfg, A.env_tape, env, ctr, 0) % (uint32_t)n_places);
        const float frac = tape_frac(cfg, A.env_tape, env, ctr);
        const int timer = (int)(tape_draw(cfg, A.env_tape, env, ctr, 2) % 50u);  // LANE_CHANGE_FREQ
        const int lane = (int)A.ma_place_f[((size_t)env * cfg.ma_places + p) * 8 + 4];
        const MapView m = map_view(A, A.env_i[env * ENV_I + EI_MAP]);
        const float* L = m.lane_f + lane * LANE_F;
        const float lon = frac * L[LF_LENGTH] / 2.0f;
#pragma unroll
        for (int k = 0; k < VEH_S; k++) St[k] = 0.0f;
#pragma unroll
        for (int k = 0; k < VEH_C; k++) C[k] = 0.0f;
#pragma unroll
        for (int k = 0; k < VEH_I; k++) I[k] = 0;
#pragma unroll
        for (int k = 0; k < NAVI_DIM; k++) navi[k] = 0.0f;
        lane_position(L, lon, 0.0f, St[VS_POS], St[VS_POS + 1]);
        St[VS_POS + 2] = 0.5f * P[VP_HEIGHT];
        yaw_quat_for_lane(L, lon, St[VS_QUAT], St[VS_QUAT + 3]);
        const int* rt = A.ma_route + ((size_t)env * cfg.ma_places + p) * ROUTE_MAX;
        const int* rr = A.ma_rroad + ((size_t)env * cfg.ma_places + p) * ROUTE_MAX;
        int* vr = A.veh_route + g * ROUTE_MAX;
        int* vrr = A.veh_rroad + g * ROUTE_MAX;
        int n_ck = 0;
        for (int k = 0; k < ROUTE_MAX; k++) { const int c = rt[k]; vr[k] = c; vrr[k] = rr[k]; if (c >= 0) n_ck++; }
        I[VI_KIND] = 2; I[VI_ALIVE] = 1; I[VI_ACTIVE] = 1; I[VI_TRIGGER] = -1;
        I[VI_LANE] = lane; I[VI_SPAWN_LANE] = lane;
        I[VI_CKPT0] = 0; I[VI_CKPT1] = n_ck > 2 ? 1 : 0; I[VI_ROUTE_LEN] = n_ck;
        I[VI_ROUTING_LANE] = -1;
        float4* d4 = reinterpret_cast<float4*>(A.veh_idm + g * VEH_IDM);
        const float rng_ctr = A.veh_idm[g * VEH_IDM + VD_RNG];
        d4[0] = make_float4((float)timer, 30.0f, 0.0f, 0.0f);  // fresh IDMPolicy (policy/idm_policy.py:224-233)
        d4[1] = make_float4(0.0f, 0.0f, rng_ctr, 0.0f);
        latch_before_step(St, C, I);
        const Rect r = vehicle_rect(P, St);
        after_step_vehicle(m, St, C, I, vr, vrr, navi, fp, obj_all + (size_t)le * O * OBJ_F, S, O, slot, r);
        I[VI_FLAGS] = FL_ON_LANE;  // BaseVehicle.reset ends with _init_step_info (base_vehicle.py:379)
        store16(A.veh_s + g * VEH_S, St);
        store16(A.veh_c + g * VEH_C, C);
        store16i(A.veh_i + g * VEH_I, I);
#pragma unroll
        for (int k = 0; k < NAVI_DIM; k++) A.veh_navi[g * NAVI_DIM + k] = navi[k];
        write_body_row(body_tab + g * BODY_ROW, P, St, 1);
        if (rank == total - 1) A.env_i[env * ENV_I + EI_RNG] = (int)(ctr - (uint32_t)rank + (uint32_t)total);
    }
}

// ================================================================================================ k_lidar
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

__host__ __device__ inline size_t lidar_best_offset(int S, int O) {   // the per-ray nearest-hit words follow the older tables
    size_t b = (size_t)S * BODY_ROW * 4 + (size_t)O * OBJ_F * 4 + (sizeof(float) + sizeof(int)) * (size_t)(S + O);
    return (b + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t lidar_smem_per_warp(int S, int O, int N) {
    return lidar_best_offset(S, O) + (((size_t)N * 8 + sizeof(int) * (2 * (size_t)(S + O) + 2) + 15) & ~(size_t)15) + 16;
}

// Lidar.perceive (component/sensors/lidar.py:49-73 -> sensors/distance_detector.py:27-85), one warp per agent.
// The reference's angular mask (lidar.py:140-168) only skips rays that provably miss; here the same idea is applied body by
// body: a candidate is cast only against the rays inside the (padded) angle its bounding circle subtends.
// LidarStateObservation._add_noise_to_cloud_points (obs/state_obs.py:236-244) for one ray: Gaussian noise, clip to [0, 1],
// then dropout to 0.  Uniforms come from a counter hash of (seed, observation pass, agent, ray); Box-Muller for the normal.
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ float lidar_noise(const MdConfig& cfg, float frac, uint32_t agent, uint32_t ray, uint32_t pass) {
    const uint32_t key = mix32((uint32_t)cfg.noise_seed * 0x9E3779B9u + pass) ^ mix32(agent * 0x85EBCA6Bu + ray * 0xC2B2AE35u + 0x27D4EB2Fu);
    if (cfg.lidar_gaussian_noise > 0.0f) {
        const float u1 = ((float)(mix32(key + 1u) >> 8) + 1.0f) * (1.0f / 16777216.0f);  // (0, 1]
        const float u2 = (float)(mix32(key + 2u) >> 8) * (1.0f / 16777216.0f);           // [0, 1)
        const float z = sqrtf(-2.0f * logf(u1)) * md_cosf(MD_TWO_PI * u2);
        frac = clipf(frac + cfg.lidar_gaussian_noise * z, 0.0f, 1.0f);
    }
    if (cfg.lidar_dropout_prob > 0.0f) {
        const float u3 = (float)(mix32(key + 3u) >> 8) * (1.0f / 16777216.0f);
        if (u3 < cfg.lidar_dropout_prob) frac = 0.0f;
    }
    return frac;
}
// Multi-agent passes: the seats that observe (FL_VALID / FL_NEWBORN) are a fraction of all seats (wrecks, empty seats,
// nothing new in the respawn pass).  With a warp per SEAT every CTA of k_lidar lived as long as its one or two live warps -
// 18 waves of mostly empty CTAs; the observing seats are therefore compacted first and k_lidar takes a warp per ENTRY.
// lidar.add_others_navi (component/sensors/lidar.py:120-129): the two navigation checkpoints of every neighbour of the num_others
// block (BaseNavigation.get_checkpoints, base_navigation.py:145-152: the ends of the first lane of ITS current road and of its next
// road - the current one again on the final road - shifted by later_middle = (lanes of the current road / 2 - 0.5) x width of the
// lane it is on) through Lidar._project_to_vehicle_system (:85-91): the offset from the observer, cut to the perceive distance, in
// the observer's frame.  A kernel of its own, launched after k_lidar only when the key is on (k_lidar left the neighbour's slot in
// the first of the four floats), so that the ray casting keeps its registers: one thread per (observer, neighbour).
__global__ void k_others_navi(MdConfig cfg, MdArrays A, const float* __restrict__ body_tab, float* __restrict__ out, int out_stride,
                              int out_off, const uint8_t* __restrict__ env_mask, const int* __restrict__ agent_flags, int need_flag) {
    const int S = cfg.slots_per_env, NA = cfg.agents_per_env, K = cfg.num_others;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)cfg.n_envs * NA * K) return;
    const long long a = t / K;
    const int n = (int)(t - a * K);
    const int env = (int)(a / NA), slot = (int)(a - (long long)env * NA);
    if (env_mask != nullptr && env_mask[env] == 0) return;
    if (agent_flags != nullptr) { if (!(agent_flags[a] & need_flag)) return; }   // the observers of k_lidar, by the same rule
    else if (!A.veh_i[(size_t)(env * S + slot) * VEH_I + VI_ACTIVE]) return;
    float* o4 = out + (size_t)a * out_stride + out_off + 8 * n + 4;
    const int k = __float_as_int(o4[0]);
    if (k < 0) { o4[0] = o4[1] = o4[2] = o4[3] = 0.0f; return; }
    const float* eb = body_tab + ((size_t)env * S + slot) * BODY_ROW;
    const float ox = eb[16], oy = eb[17];
    const MapView m = map_view(A, A.env_i[env * ENV_I + EI_MAP]);
    const size_t g = (size_t)env * S + k;
    const int* Ig = A.veh_i + g * VEH_I;
    const int* route = A.veh_route + g * ROUTE_MAX;
    const int c0 = Ig[VI_CKPT0], c1 = Ig[VI_CKPT1];
    const int cur_road = find_road(m, route[c0], route[c0 + 1]);
    const int next_road = c1 != c0 ? find_road(m, route[c1], route[c1 + 1]) : -1;
    const int cur_first = m.road_i[cur_road * ROAD_I + RI_FIRST], cur_n = m.road_i[cur_road * ROAD_I + RI_N];
    const int nx_first = next_road >= 0 ? m.road_i[next_road * ROAD_I + RI_FIRST] : cur_first;
    const int lane = Ig[VI_LANE] >= 0 ? Ig[VI_LANE] : cur_first;
    const float later_middle = ((float)cur_n / 2.0f - 0.5f) * m.lane_f[lane * LANE_F + LF_WIDTH];
    const float D = cfg.lidar_dist;
    for (int j = 0; j < 2; j++) {
        const float* L = m.lane_f + (j == 0 ? cur_first : nx_first) * LANE_F;
        float ckx, cky;
        lane_position(L, L[LF_LENGTH], later_middle, ckx, cky);
        float cx = ckx - ox, cy = cky - oy;
        const float dn = sqrtf(cx * cx + cy * cy);
        if (dn > D) { cx = cx / dn * D; cy = cy / dn * D; }
        const float in_heading = cx * eb[7] + cy * eb[10];
        const float in_rhs = -(cx * eb[6] + cy * eb[9]);
        o4[2 * j] = clipf((in_heading / D + 1.0f) / 2.0f, 0.0f, 1.0f);
        o4[2 * j + 1] = clipf((in_rhs / D + 1.0f) / 2.0f, 0.0f, 1.0f);
    }
}
__global__ void k_lidar_list(long long n_agents, const int* __restrict__ agent_flags, int need_flag, int* __restrict__ list,
                             unsigned int* __restrict__ count) {
    const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = a < n_agents && (agent_flags[a] & need_flag) != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if (bal == 0u) return;
    const int lane = threadIdx.x & 31;
    unsigned int base = 0;
    if (lane == 0) base = atomicAdd(count, (unsigned int)__popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (ok) list[base + __popc(bal & ((1u << lane) - 1u))] = (int)a;
}
__global__ void __launch_bounds__(LIDAR_WARPS * 32)
k_lidar(MdConfig cfg, const float* __restrict__ body_tab, const float* __restrict__ obj_f, const int* __restrict__ veh_i,
        const float* __restrict__ veh_p, float* __restrict__ out, int out_stride, int out_off, int* __restrict__ hit_out, const uint8_t* __restrict__ env_mask,
        const int* __restrict__ agent_flags, int need_flag, const float* __restrict__ ray_cs, const uint32_t* __restrict__ pass_ctr,
        uint32_t noise_off, const int* __restrict__ list, const unsigned int* __restrict__ list_count) {
    const uint32_t noise_pass = (pass_ctr != nullptr ? pass_ctr[1] : 0u) + noise_off;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = cfg.slots_per_env, O = cfg.objs_per_env, NA = cfg.agents_per_env, N = cfg.n_lasers;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t body_bytes = (size_t)S * BODY_ROW * 4, obj_bytes = (size_t)O * OBJ_F * 4;
    const size_t per_warp = lidar_smem_per_warp(S, O, N);
    unsigned char* my = smem_raw + per_warp * warp;
    float* sbody = reinterpret_cast<float*>(my);
    float* sobj = reinterpret_cast<float*>(my + body_bytes);
    float* srad = reinterpret_cast<float*>(my + body_bytes + obj_bytes);
    int* sidx = reinterpret_cast<int*>(srad + S + O);
    uint64_t* bar = reinterpret_cast<uint64_t*>(my + per_warp - 16);

    long long a = (long long)blockIdx.x * LIDAR_WARPS + warp;
    if (list != nullptr) {   // compacted pass: entry -> seat
        if (a >= (long long)*list_count) return;
        a = list[a];
    }
    if (a >= (long long)cfg.n_envs * NA) return;
    const int env = (int)(a / NA), slot = (int)(a - (long long)env * NA);
    if (env_mask != nullptr && env_mask[env] == 0) return;
    // who observes: the active agents; in a multi-agent step the seats that produced a transition (FL_VALID: an agent that
    // finished this step still gets its last observation) or, in the respawn pass, the newborn seats (FL_NEWBORN)
    if (agent_flags != nullptr) { if (!(agent_flags[a] & need_flag)) return; }
    else if (!veh_i[(size_t)(env * S + slot) * VEH_I + VI_ACTIVE]) return;
    // sensor noise belongs to the observation (obs/state_obs.py:225-229), not to Lidar.perceive: md_lidar (out_off < 0) is clean
    const bool noisy = out_off >= 0 && (cfg.lidar_gaussian_noise > 0.0f || cfg.lidar_dropout_prob > 0.0f);

    // stage the env's body rows and object rows: one bulk async copy (TMA 1-D) each, completing on the warp's mbarrier
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0) {
        mbar_expect_tx(bar, (uint32_t)(body_bytes + obj_bytes));
        bulk_g2s(sbody, body_tab + (size_t)env * S * BODY_ROW, (uint32_t)body_bytes, bar);
        if (obj_bytes) bulk_g2s(sobj, obj_f + (size_t)env * O * OBJ_F, (uint32_t)obj_bytes, bar);
    }
    mbar_wait(bar, 0);

    // bounding radii for the conservative prune (+0.1 % + 1 mm slack); < 0 marks "skip"
    for (int k = lane; k < S; k += 32) {
        const float* b = sbody + BODY_ROW * k;
        srad[k] = (b[15] != 0.0f && k != slot) ? sqrtf(b[3] * b[3] + b[4] * b[4] + b[5] * b[5]) * 1.001f + 1e-3f : -1.0f;
    }
    for (int k = lane; k < O; k += 32) {
        const float* ob = sobj + OBJ_F * k;
        float r = OB_IS_BOX(ob[OB_KIND]) ? sqrtf(ob[OB_A] * ob[OB_A] + ob[OB_B] * ob[OB_B]) : ob[OB_A];
        srad[S + k] = ob[OB_KIND] >= 0.0f ? r * 1.001f + 1e-3f : -1.0f;
    }
    __syncwarp();

    const float* eb = sbody + BODY_ROW * slot;
    const float D = cfg.lidar_dist;
    float hx, hy;
    {
        float fx = eb[7], fy = eb[10];  // R[0][1], R[1][1]: the chassis +Y axis
        float n = sqrtf(fx * fx + fy * fy);
        hx = fx / n; hy = fy / n;
    }
    const F3 o = f3(eb[16], eb[17], LIDAR_HEIGHT);
    const int K = out_off >= 0 ? cfg.num_others : 0;
    const int KW = OBS_OTHER_W(cfg);
    float* orow = out + (size_t)a * out_stride + (out_off >= 0 ? out_off + KW * K : 0);
    if (K > 0) {
        // Lidar.get_surrounding_vehicles_info (component/sensors/lidar.py:93-138): the K nearest vehicles of the broad
        // phase set (bodies overlapping the disc of radius int(distance)), by centre distance.  Lane l holds the
        // candidates k = l, l + 32, ...; K rounds of a warp arg-min (ties: lower slot).
        const float rad = (float)(int)D;
        const float max_speed = veh_p[(size_t)(env * S + slot) * VEH_P + VP_MAX_SPEED];
        float key[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int k = lane + 32 * j;
            key[j] = 3.0e38f;
            if (k < S && k != slot) {
                const float* b = sbody + BODY_ROW * k;
                if (b[15] != 0.0f) {
                    Rect r;
                    const float fx = b[7], fy = b[10], fn = sqrtf(fx * fx + fy * fy);
                    r.cx = b[0]; r.cy = b[1]; r.ux = fx / fn; r.uy = fy / fn; r.hu = b[4]; r.hv = b[3];
                    if (rect_circle(r, o.x, o.y, rad)) {
                        const float dx = o.x - b[16], dy = o.y - b[17];
                        key[j] = dx * dx + dy * dy;
                    }
                }
            }
        }
        float* orow_k = out + (size_t)a * out_stride + out_off;
        for (int n = 0; n < K; n++) {
            float bd = key[0];
            int bk = lane;
#pragma unroll
            for (int j = 1; j < 4; j++) if (key[j] < bd) { bd = key[j]; bk = lane + 32 * j; }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float od = __shfl_xor_sync(0xffffffffu, bd, off);
                const int ok = __shfl_xor_sync(0xffffffffu, bk, off);
                if (od < bd || (od == bd && ok < bk)) { bd = od; bk = ok; }
            }
            if ((bk & 31) == lane && bd < 3.0e38f) key[bk >> 5] = 3.0e38f;  // taken
            if (lane == 0) {
                float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (bd < 3.0e38f) {
                    const float* b = sbody + BODY_ROW * bk;
                    const float dx = b[16] - o.x, dy = b[17] - o.y;
                    const float fwd = dx * eb[7] + dy * eb[10];            // base_vehicle.py:983-988
                    const float rhs = -(dx * eb[6] + dy * eb[9]);
                    const float vx = b[18] * 3.6f - eb[18] * 3.6f, vy = b[19] * 3.6f - eb[19] * 3.6f;
                    const float vf = vx * eb[7] + vy * eb[10];
                    const float vr = -(vx * eb[6] + vy * eb[9]);
                    v.x = clipf((fwd / D + 1.0f) / 2.0f, 0.0f, 1.0f);
                    v.y = clipf((rhs / D + 1.0f) / 2.0f, 0.0f, 1.0f);
                    v.z = clipf((vf / max_speed + 1.0f) / 2.0f, 0.0f, 1.0f);
                    v.w = clipf((vr / max_speed + 1.0f) / 2.0f, 0.0f, 1.0f);
                }
                orow_k[KW * n] = v.x; orow_k[KW * n + 1] = v.y; orow_k[KW * n + 2] = v.z; orow_k[KW * n + 3] = v.w;
                if (KW == 8) orow_k[8 * n + 4] = __int_as_float(bd < 3.0e38f ? bk : -1);   // the neighbour's slot, for k_others_navi
            }
        }
    }
    // candidate list: the bodies a ray of length D can reach at all (centre within D + bounding radius), in ascending
    // index order so that ties between equal hit fractions resolve as in a full scan
    int n_cand = 0;
    for (int base = 0; base < S + O; base += 32) {
        const int k = base + lane;
        bool ok = false;
        if (k < S + O) {
            const float rb = srad[k];
            if (rb >= 0.0f) {
                const float cx = k < S ? sbody[BODY_ROW * k] : sobj[OBJ_F * (k - S) + OB_X];
                const float cy = k < S ? sbody[BODY_ROW * k + 1] : sobj[OBJ_F * (k - S) + OB_Y];
                const float dx = cx - o.x, dy = cy - o.y, reach = D + rb + 0.01f;
                ok = dx * dx + dy * dy <= reach * reach;
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        if (ok) sidx[n_cand + __popc(bal & ((1u << lane) - 1u))] = k;
        n_cand += __popc(bal);
    }
    __syncwarp();
    int* hrow = hit_out ? hit_out + (size_t)a * N : nullptr;
    // Body-major casting.  A body can only be hit by the rays inside the angle its bounding circle subtends (the
    // reference's own angular mask, lidar.py:140-168, turned around): for every candidate the warp casts just those rays
    // - a few per cent of the 240 x K (ray, body) pairs, flattened into one list so that the lanes stay dense - and keeps
    // the nearest hit per ray in shared memory as a packed (fraction bits, body index) word, so that equal fractions
    // resolve to the lower index as in a full scan.
    // The range is padded by two rays and 1e-3 rad: it is a prune, the exact slab / cylinder test decides.
    unsigned long long* sbest = reinterpret_cast<unsigned long long*>(my + lidar_best_offset(S, O));
    int* s_lo = reinterpret_cast<int*>(sbest + N);     // per candidate: first ray of its range, and the running offset of
    int* s_off = s_lo + (S + O);                        // its (candidate, ray) pairs in the flattened list (n_cand + 1 entries)
    const unsigned long long none = ((unsigned long long)__float_as_uint(2.0f) << 32) | 0xffffffffull;
    for (int i = lane; i < N; i += 32) sbest[i] = none;
    const float dth = MD_TWO_PI / (float)N;
    // (a) one lane per candidate: the padded range of rays its bounding circle can meet, and the prefix sum of the counts
    int total = 0;
    for (int base = 0; base < n_cand; base += 32) {
        const int ci = base + lane;
        int i_lo = 0, count = 0;
        if (ci < n_cand) {
            const int k = sidx[ci];
            const float cx = k < S ? sbody[BODY_ROW * k] : sobj[OBJ_F * (k - S) + OB_X];
            const float cy = k < S ? sbody[BODY_ROW * k + 1] : sobj[OBJ_F * (k - S) + OB_Y];
            const float rx = cx - o.x, ry = cy - o.y;
            const float d2 = rx * rx + ry * ry;
            const float rpad = srad[k] * 1.01f + 0.05f;
            count = N;
            if (d2 > rpad * rpad) {
                const float lx = rx * hx + ry * hy, ly = -rx * hy + ry * hx;        // centre in the heading frame
                const float th = md_atan2f(ly, lx);                                   // ray i points at angle i * dth
                const float al = md_atan2f(rpad, sqrtf(d2 - rpad * rpad)) + 1e-3f;    // half angle of the bounding circle
                i_lo = (int)floorf((th - al) / dth) - 2;
                count = (int)ceilf((2.0f * al) / dth) + 5;
                if (count > N) count = N;
            }
        }
        int incl = count;   // inclusive warp scan
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += v;
        }
        if (ci < n_cand) { s_lo[ci] = i_lo; s_off[ci] = total + incl - count; }
        total += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) s_off[n_cand] = total;
    __syncwarp();
    // (b) the flattened (candidate, ray) pairs, 32 at a time: dense lanes whatever the ranges look like
    for (int p = lane; p < total; p += 32) {
        int lo = 0, hi = n_cand;            // the candidate ci with s_off[ci] <= p < s_off[ci + 1]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (s_off[mid] <= p) lo = mid; else hi = mid;
        }
        const int k = sidx[lo];
        int i = (s_lo[lo] + (p - s_off[lo])) % N;
        if (i < 0) i += N;
        const float2 cs = __ldg(reinterpret_cast<const float2*>(ray_cs) + i);
        const float ux = hx * cs.x - hy * cs.y, uy = hy * cs.x + hx * cs.y;   // unit direction (distance_detector.py:177-180)
        const F3 d = f3(ux * D, uy * D, 0.0f);
        float t;
        if (k < S) {
            const float* b = sbody + BODY_ROW * k;
            M3 R;
            R.m[0][0] = b[6]; R.m[0][1] = b[7]; R.m[0][2] = b[8];
            R.m[1][0] = b[9]; R.m[1][1] = b[10]; R.m[1][2] = b[11];
            R.m[2][0] = b[12]; R.m[2][1] = b[13]; R.m[2][2] = b[14];
            t = ray_obb(o, d, f3(b[0], b[1], b[2]), R, f3(b[3], b[4], b[5]));
        } else {
            const float* ob = sobj + OBJ_F * (k - S);
            if (OB_IS_BOX(ob[OB_KIND])) {
                const float ch = md_cosf(ob[OB_HEADING]), sh = md_sinf(ob[OB_HEADING]);
                M3 R;
                R.m[0][0] = ch; R.m[0][1] = -sh; R.m[0][2] = 0.0f; R.m[1][0] = sh; R.m[1][1] = ch; R.m[1][2] = 0.0f;
                R.m[2][0] = 0.0f; R.m[2][1] = 0.0f; R.m[2][2] = 1.0f;
                t = ray_obb(o, d, f3(ob[OB_X], ob[OB_Y], ob[OB_ZC]), R, f3(ob[OB_B], ob[OB_A], 0.5f * ob[OB_HEIGHT]));
            } else t = ray_zcyl(o, d, ob[OB_X], ob[OB_Y], ob[OB_ZC], ob[OB_A], 0.5f * ob[OB_HEIGHT]);
        }
        // two candidates may meet the same ray in one pass: the packed minimum is atomic
        if (t < 2.0f) atomicMin(&sbest[i], ((unsigned long long)__float_as_uint(t) << 32) | (unsigned)k);
    }
    __syncwarp();
    for (int i = lane; i < N; i += 32) {
        const unsigned long long w = sbest[i];
        const float best = __uint_as_float((unsigned)(w >> 32));
        float frac = best <= 1.0f ? best : 1.0f;
        if (noisy) frac = lidar_noise(cfg, frac, (uint32_t)a + (uint32_t)(cfg.env_base * NA), (uint32_t)i, noise_pass);
        orow[i] = frac;
        if (hrow) hrow[i] = best <= 1.0f ? (int)(unsigned)(w & 0xffffffffull) : -1;
    }
}

// ---- k_linedet: SideDetector / LaneLineDetector.perceive (sensors/distance_detector.py:27-85, 194-209; called from
// obs/state_obs.py:77-86, 129-140).  One warp per agent, one lane per ray: horizontal rays at z = 0.2 from the vehicle
// position, start phase 90 degrees, against the lane-line ghost boxes of the static world (half extents len/2 x 0.0375 x
// 1.0 at z = 0.5).  The ray walks the map's 8 m uniform grid (the static broad phase) cell by cell and stops as soon as
// the nearest hit lies before the cell's exit.  Pass 0 = side detector (continuous lines), pass 1 = lane-line detector
// (continuous + broken).
__global__ void __launch_bounds__(LIDAR_WARPS * 32)
k_linedet(MdConfig cfg, MdArrays A, const float* __restrict__ body_tab, float* __restrict__ out,
          const uint8_t* __restrict__ env_mask, const int* __restrict__ agent_flags, int need_flag,
          const float* __restrict__ side_cs, const float* __restrict__ lane_cs) {
    const int S = cfg.slots_per_env, NA = cfg.agents_per_env;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long a = (long long)blockIdx.x * LIDAR_WARPS + warp;
    if (a >= (long long)cfg.n_envs * NA) return;
    const int env = (int)(a / NA), slot = (int)(a - (long long)env * NA);
    if (env_mask != nullptr && env_mask[env] == 0) return;
    if (agent_flags != nullptr) { if (!(agent_flags[a] & need_flag)) return; }
    else if (!A.veh_i[(size_t)(env * S + slot) * VEH_I + VI_ACTIVE]) return;
    const float* eb = body_tab + (size_t)(env * S + slot) * BODY_ROW;
    float hx, hy;
    {
        const float fx = eb[7], fy = eb[10], n = sqrtf(fx * fx + fy * fy);
        hx = fx / n; hy = fy / n;
    }
    const F3 o = f3(eb[16], eb[17], DET_HEIGHT);
    const MapView m = map_view(A, A.env_i[env * ENV_I + EI_MAP]);
    const float4* line4 = reinterpret_cast<const float4*>(m.lines);
    float* orow = out + (size_t)a * OBS_DIM(cfg);
    for (int pass = 0; pass < 2; pass++) {
        const int n = pass == 0 ? cfg.n_side_lasers : cfg.n_lane_lasers;
        const float D = pass == 0 ? cfg.side_dist : cfg.lane_dist;
        const float* tab = pass == 0 ? side_cs : lane_cs;
        float* dst = orow + (pass == 0 ? 0 : OBS_SIDE(cfg) + 6);
        for (int i = lane; i < n; i += 32) {
            const float c = tab[2 * i], s = tab[2 * i + 1];
            const float ux = hx * c - hy * s, uy = hy * c + hx * s;
            const F3 d = f3(ux * D, uy * D, 0.0f);
            float best = 2.0f;
            // grid walk (Amanatides-Woo) in units of the ray parameter t in [0, 1]
            int cx = (int)floorf((o.x - m.gx0) / m.cell), cy = (int)floorf((o.y - m.gy0) / m.cell);
            const int sx = d.x > 0.0f ? 1 : -1, sy = d.y > 0.0f ? 1 : -1;
            const float inv_x = d.x != 0.0f ? 1.0f / d.x : 3.0e38f, inv_y = d.y != 0.0f ? 1.0f / d.y : 3.0e38f;
            float tx = d.x != 0.0f ? ((m.gx0 + (float)(cx + (sx > 0)) * m.cell) - o.x) * inv_x : 3.0e38f;
            float ty = d.y != 0.0f ? ((m.gy0 + (float)(cy + (sy > 0)) * m.cell) - o.y) * inv_y : 3.0e38f;
            const float dtx = fabsf(m.cell * inv_x), dty = fabsf(m.cell * inv_y);
            for (int guard = 0; guard < 64; guard++) {
                if (cx >= 0 && cy >= 0 && cx < m.nx && cy < m.ny) {
                    const int cidx = cy * m.nx + cx;
                    const int k1 = m.gs[cidx + 1];
                    for (int k = m.gs[cidx]; k < k1; k++) {
                        const int it = __ldg(m.gi + k);
                        if (it >= m.n_lines) continue;  // sidewalk quads live in the same grid
                        const float4 h = __ldg(line4 + 2 * it);   // cx cy half kind
                        if (pass == 0 && h.w >= 2.0f) continue;   // ContinuousLaneLine only
                        const float4 u = __ldg(line4 + 2 * it + 1);
                        M3 R;
                        R.m[0][0] = u.x; R.m[0][1] = -u.y; R.m[0][2] = 0.0f; R.m[1][0] = u.y; R.m[1][1] = u.x; R.m[1][2] = 0.0f;
                        R.m[2][0] = 0.0f; R.m[2][1] = 0.0f; R.m[2][2] = 1.0f;
                        const float t = ray_obb(o, d, f3(h.x, h.y, 0.5f), R, f3(h.z, LINE_HALF_W, 1.0f));
                        if (t < best) best = t;
                    }
                }
                const float t_exit = fminf(tx, ty);
                // a hit found so far lies inside the visited cells (+ 5 cm slack for boxes registered by their AABB)
                if (best <= t_exit || t_exit > 1.0f) break;
                if (tx < ty) { cx += sx; tx += dtx; } else { cy += sy; ty += dty; }
            }
            dst[i] = best <= 1.0f ? best : 1.0f;
        }
    }
}

// env finished <=> none of its active agents is still running (single-agent: the agent terminated or truncated)
__global__ void k_done_mask(MdConfig cfg, const uint8_t* __restrict__ term, const uint8_t* __restrict__ trunc,
                            uint8_t* __restrict__ mask) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= cfg.n_envs) return;
    bool any_running = false;
    for (int s = 0; s < cfg.agents_per_env; s++) {
        size_t a = (size_t)e * cfg.agents_per_env + s;
        if (!(term[a] || trunc[a])) any_running = true;
    }
    mask[e] = any_running ? 0 : 1;
}

// multi-agent: an env is over when no seat is active any more (after this step's respawn)
__global__ void k_done_mask_ma(MdConfig cfg, const int* __restrict__ veh_i, uint8_t* __restrict__ mask) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= cfg.n_envs) return;
    bool any_active = false;
    for (int s = 0; s < cfg.agents_per_env; s++)
        if (veh_i[(size_t)(e * cfg.slots_per_env + s) * VEH_I + VI_ACTIVE]) any_active = true;
    mask[e] = any_active ? 0 : 1;
}

// ---- k_respawn: MultiAgentMetaDrive._respawn_vehicles (envs/marl_envs/multi_agent_metadrive.py:133-135, 176-212) -----
// One warp per env, after every agent has observed.  The lanes rebuild the env's footprints in shared memory and clear
// the body rows of the vehicles that left this step; lane p tests safe place p (8 x 3 m region, spawn_manager.py:163-209)
// against all of them; lane 0 then draws place + destination from the env's random tape and runs the newborn's
// BaseVehicle.reset + after_step + first observation (agent_manager.py:136-154).  At most one respawn per env per step,
// as in the reference (every clear place is marked used by the first query of a step).
#define RESPAWN_WARPS 4
__global__ void __launch_bounds__(RESPAWN_WARPS * 32)
k_respawn(MdConfig cfg, MdArrays A, StepOut out, float* __restrict__ body_tab) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = cfg.slots_per_env, O = cfg.objs_per_env, NA = cfg.agents_per_env;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int env = blockIdx.x * RESPAWN_WARPS + warp;
    if (env >= cfg.n_envs) return;
    Fp* nb = reinterpret_cast<Fp*>(smem_raw) + (size_t)warp * S;
    int n_alive = 0, seat = 0x7fffffff;
    unsigned park_taken = 0u;   // parking-lot env: the spaces ACTIVE agents are heading for (marl_parking_lot.py:61-76)
    for (int s0 = 0; s0 < S; s0 += 32) {
        const int s = s0 + lane;
        int alive = 0, free_seat = 0;
        if (s < S) {
            const size_t g = (size_t)env * S + s;
            const int* I = A.veh_i + g * VEH_I;
            alive = I[VI_ALIVE];
            if (cfg.parking_spaces > 0 && s < NA && I[VI_ACTIVE]) {
                const int sp = (int)A.veh_c[g * VEH_C + VC_PARK];
                if (sp > 0) park_taken |= 1u << (sp - 1);
            }
            nb[s].alive = alive;
            if (alive) {
                float P[VEH_P], St[VEH_S];
                load16(P, A.veh_p + g * VEH_P);
                load16(St, A.veh_s + g * VEH_S);
                nb[s].r = vehicle_rect(P, St);
            } else if (I[VI_KIND] != 0) body_tab[g * BODY_ROW + 15] = 0.0f;  // it left the world this step (or earlier)
            free_seat = s < NA && I[VI_KIND] == 1 && !alive && !(out.info_flags[(size_t)env * NA + s] & FL_VALID);
        }
        const unsigned am = __ballot_sync(0xffffffffu, alive && s < NA);
        const unsigned fm = __ballot_sync(0xffffffffu, free_seat);
        n_alive += __popc(am);
        if (fm && seat == 0x7fffffff) seat = s0 + __ffs(fm) - 1;
    }
    __syncwarp();
    unsigned park_free = 0u;
    int n_park_free = 0;
    if (cfg.parking_spaces > 0) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) park_taken |= __shfl_xor_sync(0xffffffffu, park_taken, off);
        park_free = ((1u << cfg.parking_spaces) - 1u) & ~park_taken;
        n_park_free = __popc(park_free);
    }
    int* E = A.env_i + (size_t)env * ENV_I;
    const bool allowed = cfg.allow_respawn && !(cfg.horizon > 0 && E[EI_STEP] >= cfg.horizon) &&
                         seat != 0x7fffffff && n_alive < NA - 1;
    if (!allowed) return;
    // rect_region_detection per safe place (utils/pg/utils.py:213-256): lane p <-> place p (+32, ...)
    unsigned long long clear_mask = 0ull;
    for (int p0 = 0; p0 < cfg.ma_places && p0 < 64; p0 += 32) {
        const int p = p0 + lane;
        bool clear = false;
        if (p < cfg.ma_places) {
            const float4* Pl = reinterpret_cast<const float4*>(A.ma_place_f + ((size_t)env * cfg.ma_places + p) * 8);
            const float4 a = Pl[0], b = Pl[1];
            Rect pr;
            pr.cx = a.x; pr.cy = a.y; pr.ux = b.y; pr.uy = b.z; pr.hu = 4.0f; pr.hv = 1.5f;
            // parking-lot env: no free space, no agent from outside (get_available_respawn_places, marl_parking_lot.py:103-105)
            clear = !(cfg.parking_spaces > 0 && (int)b.w < cfg.parking_in_roads && n_park_free == 0);
            for (int k = 0; k < S && clear; k++)
                if (nb[k].alive && rect_rect(pr, nb[k].r)) clear = false;
        }
        clear_mask |= (unsigned long long)__ballot_sync(0xffffffffu, clear) << p0;
    }
    if (lane != 0 || clear_mask == 0ull) return;
    const int n_clear = __popcll(clear_mask);
    const uint32_t ctr = (uint32_t)E[EI_RNG];
    int pick = (int)(tape_draw(cfg, A.env_tape, env, ctr, 0) % (uint32_t)n_clear);
    int dsel = (int)(tape_draw(cfg, A.env_tape, env, ctr, 1) % (uint32_t)cfg.ma_dests);
    E[EI_RNG] = (int)(ctr + 1u);
    int p = 0;
    for (unsigned long long mk = clear_mask;; mk &= mk - 1) {  // the pick-th clear place
        p = __ffsll((long long)mk) - 1;
        if (pick-- == 0) break;
    }
    const float* Pl = A.ma_place_f + ((size_t)env * cfg.ma_places + p) * 8;
    int park = 0;
    if (cfg.parking_spaces > 0) {   // update_destination_for (marl_parking_lot.py:80-88)
        if ((int)Pl[7] < cfg.parking_in_roads) {   // from outside: the r-th free space
            int r = (int)(tape_draw(cfg, A.env_tape, env, ctr, 1) % (uint32_t)n_park_free);
            unsigned mk = park_free;
            while (r-- > 0) mk &= mk - 1u;
            dsel = __ffs((int)mk) - 1;
            park = dsel + 1;
        } else dsel = (int)(tape_draw(cfg, A.env_tape, env, ctr, 1) % (uint32_t)cfg.parking_in_roads);   // from a space: one of the ways out
    }
    const size_t g = (size_t)env * S + seat;
    float P[VEH_P], St[VEH_S], C[VEH_C], navi[NAVI_DIM];
    int I[VEH_I];
    load16(P, A.veh_p + g * VEH_P);
    // BaseVehicle.reset (component/vehicle/base_vehicle.py:273-381): pose on the lane, height H/2, everything else zeroed
#pragma unroll
    for (int k = 0; k < VEH_S; k++) St[k] = 0.0f;
#pragma unroll
    for (int k = 0; k < VEH_C; k++) C[k] = 0.0f;
#pragma unroll
    for (int k = 0; k < VEH_I; k++) I[k] = 0;
#pragma unroll
    for (int k = 0; k < NAVI_DIM; k++) navi[k] = 0.0f;
    St[VS_POS] = Pl[0]; St[VS_POS + 1] = Pl[1]; St[VS_POS + 2] = 0.5f * P[VP_HEIGHT];
    St[VS_QUAT] = Pl[2]; St[VS_QUAT + 3] = Pl[3];
    C[VC_PARK] = (float)park;
    const int rsel = (int)Pl[7] * cfg.ma_dests + dsel;
    const int* rt = A.ma_route + ((size_t)env * cfg.ma_roads * cfg.ma_dests + rsel) * ROUTE_MAX;
    const int* rr = A.ma_rroad + ((size_t)env * cfg.ma_roads * cfg.ma_dests + rsel) * ROUTE_MAX;
    int* vr = A.veh_route + g * ROUTE_MAX;
    int* vrr = A.veh_rroad + g * ROUTE_MAX;
    int n_ck = 0;
    for (int k = 0; k < ROUTE_MAX; k++) { const int c = rt[k]; vr[k] = c; vrr[k] = rr[k]; if (c >= 0) n_ck++; }
    I[VI_KIND] = 1; I[VI_ALIVE] = 1; I[VI_ACTIVE] = 1; I[VI_TRIGGER] = -1;
    I[VI_LANE] = (int)Pl[4]; I[VI_SPAWN_LANE] = (int)Pl[4];
    I[VI_CKPT0] = 0; I[VI_CKPT1] = n_ck > 2 ? 1 : 0; I[VI_ROUTE_LEN] = n_ck;
    I[VI_ROUTING_LANE] = -1; I[VI_NEW] = 1;
    latch_before_step(St, C, I);
    nb[seat].alive = 1; nb[seat].r = vehicle_rect(P, St);
    const MapView m = map_view(A, E[EI_MAP]);
    after_step_vehicle(m, St, C, I, vr, vrr, navi, nb, A.obj_f + (size_t)env * O * OBJ_F, S, O, seat, nb[seat].r);
    const size_t a = (size_t)env * NA + seat;
    agent_outputs(cfg, m, E[EI_STEP], P, St, C, I, vrr, navi, a, out, 2);
    // a newborn agent: reward 0, not done, first observation (multi_agent_metadrive.py:137-144)
    out.reward[a] = 0.0f; out.cost[a] = 0.0f; out.term[a] = 0; out.trunc[a] = 0;
    out.info_flags[a] = (I[VI_FLAGS] & 0x3ff) | FL_VALID | FL_NEWBORN;
    float4* inf = reinterpret_cast<float4*>(out.info_f + a * 8);
    inf[0] = make_float4(0.0f, 0.0f, 0.0f, C[VC_STEP_ENERGY]);
    inf[1] = make_float4(C[VC_ENERGY], 0.0f, 0.0f, 0.0f);
    store16(A.veh_s + g * VEH_S, St);
    store16(A.veh_c + g * VEH_C, C);
    store16i(A.veh_i + g * VEH_I, I);
#pragma unroll
    for (int k = 0; k < NAVI_DIM; k++) A.veh_navi[g * NAVI_DIM + k] = navi[k];
    write_body_row(body_tab + g * BODY_ROW, P, St, 1);
}

// ---- host-path compaction of a multi-agent step: only the observation rows of seats that produced a transition
// (FL_VALID) travel to the host.  Ordered (ascending seat) and deterministic: per-block counts, then every block sums the
// counts before it, ranks its own seats with a ballot scan and its warps copy the rows.
#define COMPACT_T 256
__global__ void __launch_bounds__(COMPACT_T)
k_valid_count(long long n_agents, const int* __restrict__ flags, unsigned int* __restrict__ block_counts) {
    const long long a = (long long)blockIdx.x * COMPACT_T + threadIdx.x;
    const int ok = a < n_agents && (flags[a] & FL_VALID) != 0;
    const int n = __syncthreads_count(ok);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = (unsigned int)n;
}
__global__ void __launch_bounds__(COMPACT_T)
k_valid_gather(long long n_agents, const int* __restrict__ flags, const unsigned int* __restrict__ block_counts, int od,
               const float* __restrict__ obs, float* __restrict__ cobs, unsigned int* __restrict__ total) {
    __shared__ unsigned int s_base, s_warp[COMPACT_T / 32];
    __shared__ int s_seat[COMPACT_T];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    unsigned int part = 0;
    for (int b = threadIdx.x; b < (int)blockIdx.x; b += COMPACT_T) part += block_counts[b];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    if (lane == 0 && part) atomicAdd(&s_base, part);
    const long long a = (long long)blockIdx.x * COMPACT_T + threadIdx.x;
    const bool ok = a < n_agents && (flags[a] & FL_VALID) != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) s_warp[warp] = (unsigned int)__popc(bal);
    __syncthreads();
    unsigned int before = 0, count = 0;
    for (int w = 0; w < COMPACT_T / 32; w++) { if (w < warp) before += s_warp[w]; count += s_warp[w]; }
    if (ok) s_seat[before + __popc(bal & ((1u << lane) - 1u))] = threadIdx.x;
    __syncthreads();
    const unsigned int base = s_base;
    for (unsigned int j = warp; j < count; j += COMPACT_T / 32) {
        const float* src = obs + ((size_t)blockIdx.x * COMPACT_T + s_seat[j]) * od;
        float* dst = cobs + (size_t)(base + j) * od;
        for (int k = lane; k < od; k += 32) dst[k] = src[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *total = base + count;
}

// ================================================================================================ host side / C ABI
// A View is a contiguous range of a handle's envs with everything a launch needs: the configuration with n_envs /
// env_base set to the range, the per-env arrays advanced to its first env (the map tables are shared), its slices of the
// snapshots and work buffers, and its own pass counters.  The device-resident entry points (md_step, md_reset ...) run on
// the view of the whole batch; the host-buffer entry points run one view per host group, each on its own stream, so that
// one group's D2H copy overlaps another group's kernels.
struct View {
    MdConfig cfg;
    MdArrays dev;
    Snapshot snap, post;
    float *post_body, *post_obs, *body_tab;
    float4* veh_act;
    uint4* contact_tab;          // optional (md_enable_contacts): per slot, the bodies touched during the last step
    uint8_t* mask;
    int* lidar_list;             // multi-agent lidar passes: the observing seats, compacted (k_lidar_list)
    unsigned int* lidar_count;
    int* work_list;              // the step's work list: slot rows (view-relative) of the alive + active vehicles (k_dyn -> k_scan)
    unsigned int* work_count;
    LocScan* scan_tab;           // k_scan's result per slot row (read by k_post)
    uint32_t* pass;              // device: [0] scenario-draw passes so far (the counter of the draw hash (seed, env, pass)),
                                 // [1] observation passes so far (the counter of the lidar noise hash)
};

// One host group: an env range stepped through host buffers.  Device output block `d_blk` and its pinned mirror `h_blk`
// share one layout - reward | cost | info_flags | info_f | terminated | truncated | valid-row count - so the scalars of a
// step come back with ONE copy; the observation rows are a slice of the handle-wide obs buffers (one more copy).
struct HostGroup {
    View v;
    int env0, n_envs;
    size_t na;                   // agents (seats) of the group
    cudaStream_t stream;
    cudaEvent_t done;
    unsigned char *d_blk, *h_blk;
    size_t blk_bytes, off_cost, off_flags, off_info_f, off_term, off_trunc, off_count;
    StepOut out;                 // device pointers into d_blk (+ the group's rows of d_obs)
    float *d_actions, *h_actions, *h_obs;
    float* d_cobs;               // multi-agent, compact mode: the valid observation rows, gathered (k_valid_gather)
    unsigned int* d_block_counts;
    // the group's whole step (H2D, kernels, D2H) as an instantiated CUDA graph: one launch per send instead of ~10 API
    // calls.  Every kernel argument is a fixed pointer or constant (the pass counters live in device memory), so the
    // captured sequence replays as is; it is re-captured when what shapes the sequence changes (`gkey`).
    cudaGraphExec_t gexec;
    uint64_t gkey;
    int glaunches;               // kernels inside the graph (md_launch_count)
    bool in_flight;              // md_host_send issued, md_host_recv pending
    bool pending_compact;        // the compact observation rows still have to be fetched (their count is known after `done`)
};
#define MAX_HOST_GROUPS 32

#define N_ARR 28
#define N_SNAP 11
struct md_sim {
    MdConfig cfg;
    int device;
    std::string err;
    MdArrays dev;           // device pointers
    int64_t rows[N_ARR];
    size_t bytes[N_ARR];
    void* snap_bufs[N_SNAP];
    // the state right after a full reset (restore + reset-time after_step), so that an auto-reset is a row copy
    void* post_bufs[N_SNAP];
    float* post_body;       // [NV, BODY_ROW]
    float* post_obs;        // [A, OBS_STATE]: the state part of the reset observation
    bool post_valid;
    float* body_tab;
    float* ray_tab;         // [2*MAX_LASERS lidar | 2*MAX_DET_LASERS side | 2*MAX_DET_LASERS lane] (cos, sin) pairs
    float4* veh_act;        // [NV] steering rad, engine force, brake: k_pre -> k_dyn
    uint4* contact_tab;     // [NV] or NULL
    uint8_t* mask;
    int* lidar_list;
    unsigned int* lidar_count;   // [1 + MAX_HOST_GROUPS]: the whole-batch view, then one per host group
    uint32_t* d_pass;            // [2 * (1 + MAX_HOST_GROUPS)] pass counters, same order
    View all;               // the whole batch
    // host-buffer path: pinned staging + device mirrors, split into host groups
    float *h_actions, *h_obs, *d_actions, *d_obs;
    uint8_t *h_mask, *d_mask_in;
    std::vector<HostGroup> groups;
    int compact;            // multi-agent: only valid observation rows travel to the host
    int64_t launches;
    int* work_list; unsigned int* work_count; LocScan* scan_tab;   // k_dyn -> k_scan -> k_post (see View)
    MapAccel accel;         // grid records derived from the map tables at md_load_scene
    // md_step_autoreset as a CUDA graph (device-resident path): captured once per set of caller pointers on an internal stream
    cudaGraphExec_t dg_exec; uint64_t dg_key[12]; int dg_launches; cudaStream_t dg_stream;
    int dg_misses, dg_hits;  // consecutive re-captures / replays: a caller that passes new buffers every step gets plain launches
    float dyn_alive;        // the scene's alive vehicles per env: sizes k_dyn's CTAs (0 = one thread per slot row)
    md_sim* bank;           // scenario bank (md_attach_bank): finished envs restart as a scenario drawn from it
    uint32_t bank_seed;
    bool bank_ever;         // a bank was attached at some point: envs may hold drawn scenarios, restores must be `full`
    bool loaded;
    // optional per-kernel timing: 6 events per md_step on the launch stream (bench.py's roofline leg)
    std::vector<cudaEvent_t> prof_ev;
    int prof_cap, prof_n;
};

static const char* kNames[N_ARR] = {"map_desc", "map_descf", "lane_f", "lane_i", "lane_bb", "road_i", "hull_xy", "line_f",
                                    "quad_f", "grid_start", "grid_items", "env_i", "env_trigger", "veh_p", "veh_s", "veh_c",
                                    "veh_i", "veh_route", "veh_idm", "veh_navi", "obj_f", "lgrid_start", "lgrid_items",
                                    "veh_rroad", "ma_place_f", "ma_route", "ma_rroad", "env_tape"};
static const int kRowBytes[N_ARR] = {MAPD * 4, MAPDF * 4, LANE_F * 4, LANE_I * 4, 16, ROAD_I * 4, 8, LINE_F * 4, QUAD_F * 4, 4, 4,
                                     ENV_I * 4, TRIGGER_MAX * 4, VEH_P * 4, VEH_S * 4, VEH_C * 4, VEH_I * 4, ROUTE_MAX * 4,
                                     VEH_IDM * 4, NAVI_DIM * 4, OBJ_F * 4, 4, 4, ROUTE_MAX * 4, 32, ROUTE_MAX * 4, ROUTE_MAX * 4, TAPE_W * 4};
// env_i veh_s veh_c veh_i veh_idm veh_navi obj_f veh_route veh_rroad veh_p env_trigger
static const int kSnapIdx[N_SNAP] = {11, 14, 15, 16, 18, 19, 20, 17, 23, 13, 12};
static const bool kPerEnv[N_ARR] = {false, false, false, false, false, false, false, false, false, false, false,
                                    true, true, true, true, true, true, true, true, true, true, false, false,
                                    true, true, true, true, true};

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            sim->err = std::string(#call) + ": " + cudaGetErrorString(e_);                         \
            return -1;                                                                             \
        }                                                                                          \
    } while (0)

static int opt_in_smem(md_sim* sim);
static void** arr_slot(MdArrays* a, int i) { return reinterpret_cast<void**>(a) + i; }
// rows of array i that belong to one env (0 = the array is shared, or a dummy row)
static int64_t rows_per_env(const md_sim* sim, int i) {
    const int64_t E = sim->cfg.n_envs;
    return (kPerEnv[i] && sim->rows[i] >= E && sim->rows[i] % E == 0) ? sim->rows[i] / E : 0;
}

static void set_snapshot_ptrs(Snapshot& sn, void* const* bufs, const md_sim* sim, int env0) {
    const void* p[N_SNAP];
    for (int k = 0; k < N_SNAP; k++)
        p[k] = (const char*)bufs[k] + (size_t)env0 * rows_per_env(sim, kSnapIdx[k]) * kRowBytes[kSnapIdx[k]];
    sn.env_i = (const int*)p[0];
    sn.veh_s = (const float*)p[1];
    sn.veh_c = (const float*)p[2];
    sn.veh_i = (const int*)p[3];
    sn.veh_idm = (const float*)p[4];
    sn.veh_navi = (const float*)p[5];
    sn.obj_f = (const float*)p[6];
    sn.veh_route = (const int*)p[7];
    sn.veh_rroad = (const int*)p[8];
    sn.veh_p = (const float*)p[9];
    sn.env_trigger = (const int*)p[10];
}

// the view of envs [env0, env0 + n) of a loaded handle; `count_slot` picks the view's own lidar-list counter
static View make_view(const md_sim* sim, int env0, int n, int count_slot) {
    View v;
    const MdConfig& c = sim->cfg;
    v.cfg = c;
    v.cfg.n_envs = n;
    v.cfg.env_base = env0;
    v.dev = sim->dev;
    for (int i = 0; i < N_ARR; i++) {
        char* base = (char*)*arr_slot(const_cast<MdArrays*>(&sim->dev), i);
        *arr_slot(&v.dev, i) = base + (size_t)env0 * rows_per_env(sim, i) * kRowBytes[i];
    }
    set_snapshot_ptrs(v.snap, sim->snap_bufs, sim, env0);
    set_snapshot_ptrs(v.post, sim->post_bufs, sim, env0);
    const size_t nv0 = (size_t)env0 * c.slots_per_env, na0 = (size_t)env0 * c.agents_per_env;
    v.post_body = sim->post_body + nv0 * BODY_ROW;
    v.post_obs = sim->post_obs + na0 * OBS_STATE(c);
    v.body_tab = sim->body_tab + nv0 * BODY_ROW;
    v.veh_act = sim->veh_act + nv0;
    v.contact_tab = sim->contact_tab ? sim->contact_tab + nv0 : nullptr;
    v.mask = sim->mask + env0;
    v.lidar_list = sim->lidar_list ? sim->lidar_list + na0 : nullptr;
    v.lidar_count = sim->lidar_count ? sim->lidar_count + count_slot : nullptr;
    v.pass = sim->d_pass + 2 * count_slot;
    v.work_list = sim->work_list + nv0;
    v.work_count = sim->work_count + 2 * count_slot;   // [0] entries, [1] k_scan's cursor
    v.scan_tab = sim->scan_tab + nv0;
    return v;
}

extern "C" int md_abi_version(void) { return MD_ABI_VERSION; }
extern "C" int md_sizeof_config(void) { return (int)sizeof(MdConfig); }
extern "C" int md_sizeof_arrays(void) { return (int)sizeof(MdArrays); }

extern "C" int md_create(const MdConfig* cfg, int device, md_sim** out) {
    if (!cfg || !out) return -2;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
        fprintf(stderr, "libmdstep: no usable CUDA device %d (count %d) - there is no CPU fallback\n", device, n);
        return -3;
    }
    md_sim* sim = new md_sim();
    sim->cfg = *cfg;
    sim->cfg.env_base = 0;
    sim->device = device;
    sim->loaded = false;
    sim->ray_tab = nullptr;
    sim->post_valid = false;
    sim->launches = 0;
    sim->bank = nullptr; sim->bank_seed = 0; sim->bank_ever = false;
    sim->lidar_list = nullptr; sim->lidar_count = nullptr; sim->d_pass = nullptr; sim->contact_tab = nullptr;
    sim->compact = 0;
    sim->prof_cap = 0;
    sim->prof_n = 0;
    memset(&sim->dev, 0, sizeof(sim->dev));
    *out = sim;
    if (cfg->n_lasers > MAX_LASERS || cfg->slots_per_env > STEP_THREADS || cfg->slots_per_env < 1 ||
        cfg->agents_per_env > cfg->slots_per_env || cfg->slots_per_env + cfg->objs_per_env > 128) {
        sim->err = "unsupported sizes: n_lasers <= 512, slots_per_env <= 128, slots + objects per env <= 128";
        return -4;
    }
    CK(cudaSetDevice(device));
    if (cfg->n_side_lasers > MAX_DET_LASERS || cfg->n_lane_lasers > MAX_DET_LASERS) {
        sim->err = "side / lane-line detector: at most 128 lasers";
        return -4;
    }
    // ray tables: lidar laser i points at heading + i*2pi/N (distance_detector.py:177-180); the detectors start at 90 degrees
    // (distance_detector.py:197, 206).  cos / sin in double, rounded to float - the same table the oracle builds.
    std::vector<float> tab(2 * MAX_LASERS + 4 * MAX_DET_LASERS, 0.0f);
    for (int i = 0; i < cfg->n_lasers; i++) {
        double a = (double)i * (2.0 * 3.14159265358979323846 / (double)cfg->n_lasers);
        tab[2 * i] = (float)cos(a);
        tab[2 * i + 1] = (float)sin(a);
    }
    for (int pass = 0; pass < 2; pass++) {
        const int n = pass == 0 ? cfg->n_side_lasers : cfg->n_lane_lasers;
        float* dt = tab.data() + 2 * MAX_LASERS + pass * 2 * MAX_DET_LASERS;
        for (int i = 0; i < n; i++) {
            double a = (double)i * (2.0 * 3.14159265358979323846 / (double)n);
            dt[2 * i] = -(float)sin(a);
            dt[2 * i + 1] = (float)cos(a);
        }
    }
    CK(cudaMalloc(&sim->ray_tab, tab.size() * sizeof(float)));
    CK(cudaMemcpy(sim->ray_tab, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}

extern "C" const char* md_last_error(const md_sim* sim) { return sim ? sim->err.c_str() : "null handle"; }
extern "C" int64_t md_launch_count(const md_sim* sim) { return sim ? sim->launches : 0; }

static void free_groups(md_sim* sim) {
    for (HostGroup& g : sim->groups) {
        if (g.in_flight) cudaStreamSynchronize(g.stream);
        if (g.gexec) cudaGraphExecDestroy(g.gexec);
        cudaFree(g.d_blk); cudaFreeHost(g.h_blk); cudaFree(g.d_cobs); cudaFree(g.d_block_counts);
        cudaEventDestroy(g.done);
        cudaStreamDestroy(g.stream);
    }
    sim->groups.clear();
}

extern "C" void md_destroy(md_sim* sim) {
    if (!sim) return;
    cudaSetDevice(sim->device);
    if (sim->loaded) {
        free_groups(sim);
        for (int i = 0; i < N_ARR; i++) cudaFree(*arr_slot(&sim->dev, i));
        for (int i = 0; i < N_SNAP; i++) { cudaFree(sim->snap_bufs[i]); cudaFree(sim->post_bufs[i]); }
        cudaFree(sim->post_body); cudaFree(sim->post_obs);
        cudaFree(sim->body_tab); cudaFree(sim->veh_act); cudaFree(sim->mask);
        cudaFree((void*)sim->accel.lrec); cudaFree((void*)sim->accel.irec);
        cudaFree(sim->work_list); cudaFree(sim->work_count); cudaFree(sim->scan_tab);
        if (sim->dg_exec) cudaGraphExecDestroy(sim->dg_exec);
        if (sim->dg_stream) cudaStreamDestroy(sim->dg_stream);
        cudaFreeHost(sim->h_actions); cudaFreeHost(sim->h_obs); cudaFreeHost(sim->h_mask);
        cudaFree(sim->d_actions); cudaFree(sim->d_obs); cudaFree(sim->d_mask_in);
    }
    cudaFree(sim->ray_tab);
    cudaFree(sim->lidar_list); cudaFree(sim->lidar_count); cudaFree(sim->d_pass); cudaFree(sim->contact_tab);
    for (cudaEvent_t e : sim->prof_ev) cudaEventDestroy(e);
    delete sim;
}

extern "C" int md_snapshot(md_sim* sim) {
    if (!sim || !sim->loaded) return -2;
    CK(cudaSetDevice(sim->device));
    sim->post_valid = false;
    for (int k = 0; k < N_SNAP; k++)
        CK(cudaMemcpy(sim->snap_bufs[k], *arr_slot(&sim->dev, kSnapIdx[k]), sim->bytes[kSnapIdx[k]], cudaMemcpyDeviceToDevice));
    return 0;
}

extern "C" int md_host_groups(md_sim* sim, int n_groups);

// grid records (MapAccel, md_device.cuh): one 32-byte record per lgrid_items / grid_items entry, same indexing
static int build_accel(md_sim* sim, const MdArrays* host) {
    const int64_t M = sim->rows[0], n_l = sim->rows[22], n_i = sim->rows[10];
    std::vector<float> lrec((size_t)(n_l ? n_l : 1) * 8, 0.0f), irec((size_t)(n_i ? n_i : 1) * 8, 0.0f);
    auto bits = [](int v) { float f; memcpy(&f, &v, 4); return f; };
    for (int64_t mp = 0; mp < M; mp++) {
        const int* d = host->map_desc + mp * MAPD;
        const int cells = d[MD_GRID_NX] * d[MD_GRID_NY];
        const int nl = host->lgrid_start[d[MD_LGRID_OFF] + cells], ni = host->grid_start[d[MD_GRID_OFF] + cells];
        for (int k = 0; k < nl; k++) {
            const int l = host->lgrid_items[d[MD_LITEM_OFF] + k];
            const float* bb = host->lane_bb + (size_t)(d[MD_LANE_OFF] + l) * 4;
            const int* li = host->lane_i + (size_t)(d[MD_LANE_OFF] + l) * LANE_I;
            float* o = lrec.data() + (size_t)(d[MD_LITEM_OFF] + k) * 8;
            o[0] = bits(l); o[1] = bb[0]; o[2] = bb[1]; o[3] = bb[2]; o[4] = bb[3];
            o[5] = bits(li[LI_HULL_OFF]); o[6] = bits(li[LI_HULL_N]);
            o[7] = host->lane_f[(size_t)(d[MD_LANE_OFF] + l) * LANE_F + LF_TYPE];
        }
        for (int k = 0; k < ni; k++) {
            const int it = host->grid_items[d[MD_ITEM_OFF] + k];
            const float* src = it < d[MD_N_LINES] ? host->line_f + (size_t)(d[MD_LINE_OFF] + it) * LINE_F
                                                  : host->quad_f + (size_t)(d[MD_QUAD_OFF] + it - d[MD_N_LINES]) * QUAD_F;
            memcpy(irec.data() + (size_t)(d[MD_ITEM_OFF] + k) * 8, src, 32);
        }
    }
    void *dl = nullptr, *di = nullptr;
    CK(cudaMalloc(&dl, lrec.size() * 4)); CK(cudaMalloc(&di, irec.size() * 4));
    CK(cudaMemcpy(dl, lrec.data(), lrec.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(di, irec.data(), irec.size() * 4, cudaMemcpyHostToDevice));
    sim->accel.lrec = (const float4*)dl; sim->accel.irec = (const float4*)di;
    return 0;
}

extern "C" int md_load_scene(md_sim* sim, const MdArrays* host, const int64_t* rows) {
    if (!sim || !host || !rows) return -2;
    if (sim->loaded) { sim->err = "scene already loaded; create a new handle"; return -5; }
    CK(cudaSetDevice(sim->device));
    const MdConfig& c = sim->cfg;
    const int64_t NV = (int64_t)c.n_envs * c.slots_per_env;
    if (rows[13] != NV || rows[14] != NV || rows[16] != NV || rows[11] != c.n_envs) {
        sim->err = "row counts do not match the configuration";
        return -6;
    }
    for (int i = 0; i < N_ARR; i++) {
        sim->rows[i] = rows[i];
        sim->bytes[i] = (size_t)rows[i] * kRowBytes[i];
        void* d = nullptr;
        CK(cudaMalloc(&d, sim->bytes[i] ? sim->bytes[i] : 16));
        const void* h = *arr_slot(const_cast<MdArrays*>(host), i);
        if (sim->bytes[i]) CK(cudaMemcpy(d, h, sim->bytes[i], cudaMemcpyHostToDevice));
        *arr_slot(&sim->dev, i) = d;
    }
    for (int k = 0; k < N_SNAP; k++) CK(cudaMalloc(&sim->snap_bufs[k], sim->bytes[kSnapIdx[k]] ? sim->bytes[kSnapIdx[k]] : 16));
    for (int k = 0; k < N_SNAP; k++) CK(cudaMalloc(&sim->post_bufs[k], sim->bytes[kSnapIdx[k]] ? sim->bytes[kSnapIdx[k]] : 16));
    CK(cudaMalloc(&sim->post_body, (size_t)NV * BODY_ROW * 4));
    CK(cudaMalloc(&sim->post_obs, (size_t)c.n_envs * c.agents_per_env * OBS_STATE(c) * 4));
    CK(cudaMalloc(&sim->body_tab, (size_t)NV * BODY_ROW * 4));
    CK(cudaMemset(sim->body_tab, 0, (size_t)NV * BODY_ROW * 4));
    CK(cudaMalloc(&sim->veh_act, (size_t)NV * sizeof(float4)));
    CK(cudaMemset(sim->veh_act, 0, (size_t)NV * sizeof(float4)));
    CK(cudaMalloc(&sim->mask, (size_t)c.n_envs));
    CK(cudaMemset(sim->mask, 0, (size_t)c.n_envs));
    const size_t NA = (size_t)c.n_envs * c.agents_per_env, od = OBS_DIM(c);
    if (c.is_multi_agent) {
        CK(cudaMalloc(&sim->lidar_list, sizeof(int) * NA));
        CK(cudaMalloc(&sim->lidar_count, sizeof(unsigned int) * (1 + MAX_HOST_GROUPS)));
        CK(cudaMemset(sim->lidar_count, 0, sizeof(unsigned int) * (1 + MAX_HOST_GROUPS)));
    }
    CK(cudaMalloc(&sim->work_list, sizeof(int) * (size_t)NV));
    CK(cudaMalloc(&sim->work_count, sizeof(unsigned int) * 2 * (1 + MAX_HOST_GROUPS)));
    CK(cudaMemset(sim->work_count, 0, sizeof(unsigned int) * 2 * (1 + MAX_HOST_GROUPS)));
    CK(cudaMalloc(&sim->scan_tab, sizeof(LocScan) * (size_t)NV));
    CK(cudaMalloc(&sim->d_pass, sizeof(uint32_t) * 2 * (1 + MAX_HOST_GROUPS)));
    CK(cudaMemset(sim->d_pass, 0, sizeof(uint32_t) * 2 * (1 + MAX_HOST_GROUPS)));
    CK(cudaMallocHost(&sim->h_actions, NA * 2 * 4)); CK(cudaMallocHost(&sim->h_obs, NA * od * 4));
    CK(cudaMallocHost(&sim->h_mask, (size_t)c.n_envs));
    CK(cudaMalloc(&sim->d_actions, NA * 2 * 4)); CK(cudaMalloc(&sim->d_obs, NA * od * 4));
    CK(cudaMalloc(&sim->d_mask_in, (size_t)c.n_envs));
    if (build_accel(sim, host)) return -1;
    {   // k_dyn runs about as many threads per CTA as its envs hold alive vehicles: 1.25 x the scene's mean (envs with more
        // take a second pass inside the kernel)
        const int* hi = host->veh_i;
        double alive = 0.0;
        for (int64_t g = 0; g < NV; g++) alive += hi[g * VEH_I + VI_ALIVE] != 0;
        sim->dyn_alive = (float)(alive / (double)c.n_envs);
    }
    sim->loaded = true;
    sim->all = make_view(sim, 0, c.n_envs, 0);
    if (opt_in_smem(sim)) return -4;
    if (md_host_groups(sim, 1)) return -1;
    return md_snapshot(sim);
}

static int find_name(const char* name) {
    for (int i = 0; i < N_ARR; i++)
        if (strcmp(name, kNames[i]) == 0) return i;
    return -1;
}
extern "C" int md_get_state(md_sim* sim, const char* name, void* host_dst, size_t bytes) {
    if (!sim || !sim->loaded) return -2;
    CK(cudaSetDevice(sim->device));
    if (strcmp(name, "body_tab") == 0) {
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(host_dst, sim->body_tab, bytes, cudaMemcpyDeviceToHost));
        return 0;
    }
    int i = find_name(name);
    if (i < 0 || bytes != sim->bytes[i]) { sim->err = std::string("md_get_state: bad name or size for ") + name; return -7; }
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(host_dst, *arr_slot(&sim->dev, i), bytes, cudaMemcpyDeviceToHost));
    return 0;
}
extern "C" int md_set_state(md_sim* sim, const char* name, const void* host_src, size_t bytes) {
    if (!sim || !sim->loaded) return -2;
    CK(cudaSetDevice(sim->device));
    int i = find_name(name);
    if (i < 0 || bytes != sim->bytes[i]) { sim->err = std::string("md_set_state: bad name or size for ") + name; return -7; }
    sim->post_valid = false;
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(*arr_slot(&sim->dev, i), host_src, bytes, cudaMemcpyHostToDevice));
    return 0;
}

// CTA geometry shared by k_pre / k_dyn / k_post: EPB envs per CTA, EPB * S threads, launch bound MAXT >= threads
struct StepLaunch { int epb, threads, blocks; size_t smem; };
// epb = 32 makes a warp "one slot index of 32 envs" (uniform roles: k_pre, k_post); k_dyn has no role divergence and
// prefers small CTAs (more resident CTAs, cheaper barriers).  MD_EPB_PRE / MD_EPB_POST / MD_EPB_DYN override (tuning).
static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}
static StepLaunch step_launch(const MdConfig& c, int epb_pref) {
    StepLaunch L;
    L.epb = epb_pref;
    while (L.epb > 1 && L.epb * c.slots_per_env > 1024) L.epb >>= 1;
    L.threads = (L.epb * c.slots_per_env + 31) & ~31;
    L.blocks = (c.n_envs + L.epb - 1) / L.epb;
    L.smem = step_smem_bytes(c.slots_per_env, c.objs_per_env, L.epb);
    return L;
}
static int epb_pre() { static int v = env_int("MD_EPB_PRE", PRE_EPB); return v; }
static int epb_post() { static int v = env_int("MD_EPB_POST", POST_EPB); return v; }
// k_dyn appends the compacted list of alive vehicles to the shared tables and runs `threads_pref` threads per CTA: about as
// many as its envs hold alive vehicles (md_load_scene sizes it from the scene; MD_DYN_THREADS overrides; 0 = one per slot row)
static StepLaunch dyn_launch_epb(const MdConfig& c, float alive_per_env, int epb) {
    StepLaunch L = step_launch(c, epb);
    static const int forced = env_int("MD_DYN_THREADS", -1);
    int t = forced >= 0 ? forced : (int)(1.25f * alive_per_env * (float)L.epb + 0.999f);
    if (t > 0) {
        if (t < c.slots_per_env) t = c.slots_per_env;   // a pass holds at least one whole env
        t = (t + 31) & ~31;
        if (t < L.threads) L.threads = t;
    }
    L.smem += sizeof(CBody) * (size_t)L.epb * c.slots_per_env + dyn_list_bytes(c.slots_per_env, L.epb);
    L.smem = ((L.smem + 15) & ~(size_t)15) + 16 * (size_t)L.epb * c.slots_per_env;   // contact export words
    return L;
}
// envs per CTA of k_dyn: 16 when such CTAs (6 warps at BASELINE cfg2) still fit the SMs in ONE wave - fewer half-empty warps
// and barriers than 8-env CTAs of 3 warps (0.095 -> 0.090 ms at cfg2) - else 8 (scenes with many objects are limited by
// shared memory, multi-agent scenes by registers: measured slower with 16).  MD_EPB_DYN overrides.
static StepLaunch dyn_launch(const MdConfig& c, float alive_per_env) {
    static const int forced = env_int("MD_EPB_DYN", 0);
    if (forced > 0) return dyn_launch_epb(c, alive_per_env, forced);
    const StepLaunch L16 = dyn_launch_epb(c, alive_per_env, 2 * DYN_EPB);
    if (L16.epb == 2 * DYN_EPB && L16.threads <= 192 && L16.smem <= 48 * 1024) {
        const int by_regs = 65536 / (L16.threads * DYN_REGS), by_smem = (int)((228 * 1024) / (L16.smem + 1024));
        const int per_sm = by_regs < by_smem ? by_regs : by_smem;
        static const int sms = [] { int d = 0, n = 148; cudaGetDevice(&d); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); return n; }();
        if (L16.blocks <= sms * per_sm) return L16;
    }
    return dyn_launch_epb(c, alive_per_env, DYN_EPB);
}
static StepLaunch pre_launch(const MdConfig& c) {  // k_pre: epb envs per CTA, a fixed number of worker threads
    StepLaunch L;
    L.epb = epb_pre();
    L.threads = env_int("MD_PRE_WORKERS", PRE_WORKERS);
    if (L.threads > 1024 || L.threads < 32) L.threads = PRE_WORKERS;
    while (L.epb > 1 && pre_smem_bytes(c.slots_per_env, c.objs_per_env, L.epb, L.threads) > 128 * 1024) L.epb >>= 1;
    L.blocks = (c.n_envs + L.epb - 1) / L.epb;
    L.smem = pre_smem_bytes(c.slots_per_env, c.objs_per_env, L.epb, L.threads);
    return L;
}
static StepLaunch post_launch(const MdConfig& c) {  // k_post: epb envs per CTA, a fixed number of worker threads
    StepLaunch L;
    L.epb = epb_post();
    while (L.epb > 1 && post_smem_bytes(c.slots_per_env, c.objs_per_env, L.epb) > 96 * 1024) L.epb >>= 1;
    L.threads = env_int("MD_POST_WORKERS", POST_WORKERS);
    if (L.threads > 1024 || L.threads < 32) L.threads = POST_WORKERS;
    L.blocks = (c.n_envs + L.epb - 1) / L.epb;
    L.smem = post_smem_bytes(c.slots_per_env, c.objs_per_env, L.epb);
    return L;
}

// The attribute belongs to the FUNCTION, not to a handle: it is only ever raised, so that a handle with a small scene loaded
// later does not take away the opt-in an earlier, still live handle with a larger scene launches with ("invalid argument"
// on the next launch of the earlier handle - seen with the single-env wrappers, which keep several handles alive).
template <typename K>
static cudaError_t allow_smem(K kernel, size_t bytes) {
    static std::vector<std::pair<const void*, size_t>> high;
    for (auto& h : high)
        if (h.first == (const void*)kernel) {
            if (bytes <= h.second) return cudaSuccess;
            h.second = bytes;
            return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        }
    high.push_back({(const void*)kernel, bytes});
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}
// The step kernels are occupancy-bound by registers, with 10 - 40 KB of shared memory per CTA: ask for the shared-memory
// carve-out that lets the register-limited number of CTAs be resident (the driver's default split left k_dyn at 6 CTAs per
// SM where the registers allow 7, i.e. a second wave for 13 % of the CTAs).  MD_DEBUG_OCC=1 prints the resulting occupancy.
template <typename K>
static cudaError_t prefer_carveout(K kernel, const char* name, const StepLaunch& L, int regs) {
    const int threads = (L.threads + 31) & ~31;
    int target = 65536 / (threads * regs);
    if (target > 2048 / threads) target = 2048 / threads;
    if (target > 32) target = 32;
    if (target < 1) target = 1;
    const size_t need = (size_t)target * (L.smem + 1024);
    int pct = (int)((100 * need + 228 * 1024 - 1) / (228 * 1024));
    if (pct > 100) pct = 100;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    static const int dbg = env_int("MD_DEBUG_OCC", 0);
    if (dbg && e == cudaSuccess) {
        int n = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, L.threads, L.smem);
        fprintf(stderr, "[mdstep] %s: %d blocks x %d threads, %zu B smem, carve-out %d %% -> %d CTAs / SM (register target %d)\n",
                name, L.blocks, L.threads, L.smem, pct, n, target);
    }
    return e;
}
// dynamic shared memory above the 48 KB default needs an opt-in per kernel
static int opt_in_smem(md_sim* sim) {
    const size_t floor48 = 48 * 1024;
    StepLaunch D = dyn_launch(sim->cfg, sim->dyn_alive), A = pre_launch(sim->cfg), B = post_launch(sim->cfg);
    if (A.smem > 200 * 1024 || B.smem > 200 * 1024 || D.smem > 200 * 1024) {
        sim->err = "slots/objects per env need more than 200 KB of shared memory per CTA";
        return -4;
    }
    CK(allow_smem(k_dyn, D.smem > floor48 ? D.smem : floor48));
    CK(allow_smem(k_pre, A.smem > floor48 ? A.smem : floor48));
    CK(allow_smem(k_post, B.smem > floor48 ? B.smem : floor48));
    CK(prefer_carveout(k_dyn, "k_dyn", D, DYN_REGS));
    CK(prefer_carveout(k_pre, "k_pre", A, PRE_REGS));
    CK(prefer_carveout(k_post, "k_post", B, POST_REGS));
    size_t lb = lidar_smem_per_warp(sim->cfg.slots_per_env, sim->cfg.objs_per_env, sim->cfg.n_lasers) * LIDAR_WARPS;
    if (lb > floor48) CK(allow_smem(k_lidar, lb));
    return 0;
}

static int launch_pre(md_sim* sim, const View& v, int mode, const float* actions, float* idm_out, cudaStream_t st,
                      uint32_t d_bank = 0, uint32_t d_noise = 0) {
    StepLaunch L = pre_launch(v.cfg);
    static const int use_teams = env_int("MD_PRE_TEAM", 1);
    k_pre<<<L.blocks, L.threads, L.smem, st>>>(v.cfg, v.dev, mode, L.epb, actions, idm_out, v.veh_act, use_teams,
                                               (d_bank | d_noise) ? v.pass : nullptr, d_bank, d_noise,
                                               (mode & MODE_AGENT_PRE) ? v.work_count : nullptr);
    sim->launches++;
    CK(cudaGetLastError());
    return 0;
}
static int launch_dyn(md_sim* sim, const View& v, int mode, const float* ext_act3, int n_sub, cudaStream_t st, bool emit_list = false) {
    StepLaunch L = dyn_launch(v.cfg, sim->dyn_alive);
    k_dyn<<<L.blocks, L.threads, L.smem, st>>>(v.cfg, v.dev, mode, L.epb, v.veh_act, ext_act3, n_sub,
                                               (mode & MODE_CONTACTS) ? v.contact_tab : nullptr,
                                               emit_list ? v.work_list : nullptr, v.work_count);
    sim->launches++;
    CK(cudaGetLastError());
    return 0;
}
static BankView bank_view(const md_sim* bk) {
    BankView B;
    B.post = bk->all.post; B.body = bk->post_body; B.obs = bk->post_obs; B.veh_p = bk->dev.veh_p;
    B.env_trigger = bk->dev.env_trigger; B.n = bk->cfg.n_envs;
    return B;
}
static int launch_post(md_sim* sim, const View& v, int mode, StepOut out, const uint8_t* mask, cudaStream_t st, bool scanned = false,
                       bool fuse_bank = false) {
    StepLaunch L = post_launch(v.cfg);
    static const int team_pref = env_int("MD_POST_TEAM", 0);   // 0 = adaptive, 1 = off, 2 / 4 / 8 / 16 / 32 lanes per vehicle
    FusedBank FB;
    memset(&FB, 0, sizeof(FB));
    if (fuse_bank) { FB.B = bank_view(sim->bank); FB.seed = sim->bank_seed; FB.pass_ctr = v.pass; FB.on = 1; }
    k_post<<<L.blocks, L.threads, L.smem, st>>>(v.cfg, v.dev, mode, L.epb, out, v.body_tab, mask, v.mask, v.snap, team_pref, sim->accel,
                                                scanned ? v.scan_tab : nullptr, FB);
    sim->launches++;
    CK(cudaGetLastError());
    return 0;
}
static int launch_restore(md_sim* sim, const View& v, const uint8_t* mask, cudaStream_t st) {
    long long nv = (long long)v.cfg.n_envs * v.cfg.slots_per_env;
    k_restore<<<(int)((nv + 255) / 256), 256, 0, st>>>(v.cfg, v.dev, v.snap, mask, sim->bank_ever ? 1 : 0);
    sim->launches++;
    CK(cudaGetLastError());
    return 0;
}
// scenario draw from the bank for the masked envs (all envs if mask == NULL): rows + body rows + state observation
static int launch_restore_bank(md_sim* sim, View& v, float* obs, const uint8_t* mask, cudaStream_t st) {
    const MdConfig& c = v.cfg;
    const md_sim* bk = sim->bank;
    const long long nv = (long long)c.n_envs * c.slots_per_env;
    const BankView B = bank_view(bk);
    k_restore_bank<<<(int)((nv + 255) / 256), 256, 0, st>>>(c, v.dev, B, sim->bank_seed, v.pass, v.body_tab, obs, mask);
    sim->launches++;
    CK(cudaGetLastError());
    return 0;
}
static int launch_lidar(md_sim* sim, const View& v, uint32_t noise_off, float* out, int stride, int off, int32_t* hit,
                        const uint8_t* mask, cudaStream_t st, const int* agent_flags = nullptr, int need_flag = 0) {
    const MdConfig& c = v.cfg;
    long long na = (long long)c.n_envs * c.agents_per_env;
    int blocks = (int)((na + LIDAR_WARPS - 1) / LIDAR_WARPS);
    size_t smem = lidar_smem_per_warp(c.slots_per_env, c.objs_per_env, c.n_lasers) * LIDAR_WARPS;
    const int* list = nullptr;
    if (agent_flags != nullptr && v.lidar_list != nullptr) {   // multi-agent pass: compact the observing seats first
        CK(cudaMemsetAsync(v.lidar_count, 0, sizeof(unsigned int), st));
        k_lidar_list<<<(int)((na + 255) / 256), 256, 0, st>>>(na, agent_flags, need_flag, v.lidar_list, v.lidar_count);
        sim->launches++;
        list = v.lidar_list;
    }
    k_lidar<<<blocks, LIDAR_WARPS * 32, smem, st>>>(c, v.body_tab, v.dev.obj_f, v.dev.veh_i, v.dev.veh_p, out, stride, off, hit, mask,
                                                    agent_flags, need_flag, sim->ray_tab, off >= 0 ? v.pass : nullptr, noise_off, list,
                                                    v.lidar_count);
    sim->launches++;
    CK(cudaGetLastError());
    if (off >= 0 && c.add_others_navi && c.num_others > 0) {   // the neighbours' checkpoints, next to their four floats
        const long long nt = na * c.num_others;
        k_others_navi<<<(int)((nt + 127) / 128), 128, 0, st>>>(c, v.dev, v.body_tab, out, stride, off, mask, agent_flags, need_flag);
        sim->launches++;
        CK(cudaGetLastError());
    }
    if (off >= 0 && (c.n_side_lasers > 0 || c.n_lane_lasers > 0)) {  // the detector blocks of the same observation rows
        k_linedet<<<blocks, LIDAR_WARPS * 32, 0, st>>>(c, v.dev, v.body_tab, out, mask, agent_flags, need_flag,
                                                       sim->ray_tab + 2 * MAX_LASERS, sim->ray_tab + 2 * MAX_LASERS + 2 * MAX_DET_LASERS);
        sim->launches++;
        CK(cudaGetLastError());
    }
    return 0;
}
static int launch_respawn(md_sim* sim, const View& v, StepOut out, cudaStream_t st) {
    const MdConfig& c = v.cfg;
    int blocks = (c.n_envs + RESPAWN_WARPS - 1) / RESPAWN_WARPS;
    size_t smem = sizeof(Fp) * (size_t)c.slots_per_env * RESPAWN_WARPS;
    k_respawn<<<blocks, RESPAWN_WARPS * 32, smem, st>>>(c, v.dev, out, v.body_tab);
    sim->launches++;
    CK(cudaGetLastError());
    return 0;
}

// env.reset of a view's masked envs (all of them if mask == NULL).  With a scenario bank attached the envs restart in a
// scenario DRAWN from the bank, exactly like the auto-reset of md_step_autoreset (BaseEnv.reset(seed=None),
// envs/base_env.py:886-891) - never a mix of the env's first scenario and its last draw.
static int reset_impl(md_sim* sim, View& v, const uint8_t* env_mask_dev, float* obs_dev, cudaStream_t st, bool whole_batch) {
    StepOut out = {obs_dev, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (sim->bank && obs_dev != nullptr) {
        if (launch_restore_bank(sim, v, obs_dev, env_mask_dev, st)) return -1;
    } else {
        if (launch_restore(sim, v, env_mask_dev, st)) return -1;
        if (launch_post(sim, v, MODE_RESET, out, env_mask_dev, st)) return -1;
        if (whole_batch && env_mask_dev == nullptr && obs_dev != nullptr) {
            // a full reset: keep the resulting state so that later auto-resets are plain row copies (k_restore_post)
            const MdConfig& c = sim->cfg;
            for (int k = 0; k < N_SNAP; k++)
                CK(cudaMemcpyAsync(sim->post_bufs[k], *arr_slot(&sim->dev, kSnapIdx[k]), sim->bytes[kSnapIdx[k]], cudaMemcpyDeviceToDevice, st));
            CK(cudaMemcpyAsync(sim->post_body, sim->body_tab, (size_t)c.n_envs * c.slots_per_env * BODY_ROW * 4, cudaMemcpyDeviceToDevice, st));
            CK(cudaMemcpy2DAsync(sim->post_obs, OBS_STATE(c) * 4, obs_dev, (size_t)OBS_DIM(c) * 4, OBS_STATE(c) * 4,
                                 (size_t)c.n_envs * c.agents_per_env, cudaMemcpyDeviceToDevice, st));
            sim->post_valid = true;
        }
    }
    if (launch_lidar(sim, v, 0, obs_dev, OBS_DIM(v.cfg), OBS_STATE(v.cfg), nullptr, env_mask_dev, st)) return -1;
    k_bump<<<1, 1, 0, st>>>(v.pass, (sim->bank && obs_dev != nullptr) ? 1u : 0u, 1u);   // this pass is spent
    sim->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int md_reset(md_sim* sim, const uint8_t* env_mask_dev, float* obs_dev, void* stream) {
    if (!sim || !sim->loaded) return -2;
    CK(cudaSetDevice(sim->device));
    return reset_impl(sim, sim->all, env_mask_dev, obs_dev, (cudaStream_t)stream, true);
}

// one env.step of a view; `fused_reset` (single agent only): finished envs are reset in place before the observation is
// taken, so that the auto-reset costs one extra launch (a row copy from the post-reset snapshot or from the scenario bank)
#define N_PROF_EV 6
static int step_impl(md_sim* sim, View& v, const float* actions_dev, StepOut out, cudaStream_t st, bool fused_reset, bool may_prof) {
    const bool prof = may_prof && sim->prof_n < sim->prof_cap;
    cudaEvent_t* ev = prof ? &sim->prof_ev[N_PROF_EV * sim->prof_n] : nullptr;
    if (prof) CK(cudaEventRecord(ev[0], st));
    // k_pre spends the step's passes up front: one scenario-draw pass if finished envs will draw from a bank, two
    // observation passes (single agent uses the first; multi-agent: transition pass + newborn pass)
    const uint32_t d_bank = (fused_reset && sim->post_valid && sim->bank) ? 1u : 0u;
    if (launch_pre(sim, v, MODE_AGENT_PRE | MODE_TRIGGER | MODE_IDM, actions_dev, nullptr, st, d_bank, 2u)) return -1;
    if (prof) CK(cudaEventRecord(ev[1], st));
    static const int use_scan = env_int("MD_SCAN", 1);   // 0: the scans run as teams inside k_post (tuning / fallback)
    if (launch_dyn(sim, v, MODE_DYN | MODE_CONTACTS, nullptr, v.cfg.decision_repeat, st, use_scan != 0)) return -1;
    if (prof) CK(cudaEventRecord(ev[2], st));
    if (use_scan) {
        // a persistent grid: 4 CTAs of 8 warps per SM are resident at 64 registers per thread, and each team strides over
        // the list.  MD_SCAN_TEAM = 8 / 16 / 32 lanes per vehicle (measured), MD_SCAN_CTAS overrides the grid.
        static const int scan_team = env_int("MD_SCAN_TEAM", 8), scan_ctas = env_int("MD_SCAN_CTAS", 148 * 4);
#define SCAN_LAUNCH(TS) k_scan<TS, 4><<<scan_ctas, SCAN_WARPS * 32, 0, st>>>(v.cfg, v.dev, sim->accel, v.work_list, v.work_count, v.scan_tab)
        // (6 CTAs per SM = 40 registers spill and are no faster; 4 / 16 / 32 lanes per vehicle are within 3 us of 8, 32 is slower)
        if (scan_team == 4) SCAN_LAUNCH(4); else if (scan_team == 16) SCAN_LAUNCH(16); else if (scan_team == 32) SCAN_LAUNCH(32); else SCAN_LAUNCH(8);
#undef SCAN_LAUNCH
        sim->launches++;
        CK(cudaGetLastError());
    }
    // with a bank attached the finished envs are restored by k_post itself (MD_FUSE_BANK=0: by a k_restore_bank launch)
    static const int fuse_bank_pref = env_int("MD_FUSE_BANK", 1);
    const bool bank_reset = fused_reset && sim->post_valid && sim->bank;
    const bool fuse_bank = bank_reset && fuse_bank_pref && !v.cfg.is_multi_agent && out.obs != nullptr;
    if (launch_post(sim, v, MODE_POST | MODE_OUT | MODE_REMOVE | (fused_reset ? MODE_MARK_DONE : 0), out, nullptr, st, use_scan != 0, fuse_bank)) return -1;
    if (prof) CK(cudaEventRecord(ev[3], st));
    if (bank_reset) {
        if (!fuse_bank && launch_restore_bank(sim, v, out.obs, v.mask, st)) return -1;
    } else if (fused_reset && sim->post_valid) {
        const long long nv = (long long)v.cfg.n_envs * v.cfg.slots_per_env;
        k_restore_post<<<(int)((nv + 255) / 256), 256, 0, st>>>(v.cfg, v.dev, v.post, v.post_body, v.post_obs, v.body_tab, out.obs, v.mask);
        sim->launches++;
        CK(cudaGetLastError());
    } else if (fused_reset) {
        StepOut ro = {out.obs, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        if (launch_post(sim, v, MODE_RESET | MODE_RESTORE, ro, v.mask, st)) return -1;
    }
    if (prof) CK(cudaEventRecord(ev[4], st));
    const int od = OBS_DIM(v.cfg), os = OBS_STATE(v.cfg);
    const uint32_t np = 0;
    if (!v.cfg.is_multi_agent) {
        if (launch_lidar(sim, v, np, out.obs, od, os, nullptr, nullptr, st)) return -1;
    } else {
        // multi-agent: everyone who produced a transition observes (incl. agents that just finished), then finished
        // vehicles leave / freeze, at most one agent per env is respawned and observes the world after that
        if (!out.info_flags) { sim->err = "multi-agent md_step needs info_flags"; return -2; }
        if (launch_lidar(sim, v, np, out.obs, od, os, nullptr, nullptr, st, out.info_flags, FL_VALID)) return -1;
        if (launch_respawn(sim, v, out, st)) return -1;
        if (v.cfg.allow_respawn && launch_lidar(sim, v, np + 1, out.obs, od, os, nullptr, nullptr, st, out.info_flags, FL_NEWBORN))
            return -1;
    }
    if (prof) { CK(cudaEventRecord(ev[5], st)); sim->prof_n++; }
    return 0;
}

static int autoreset_impl(md_sim* sim, View& v, const uint8_t* terminated_dev, const uint8_t* truncated_dev, float* obs_dev, cudaStream_t st,
                          bool whole_batch) {
    if (v.cfg.is_multi_agent)
        k_done_mask_ma<<<(v.cfg.n_envs + 255) / 256, 256, 0, st>>>(v.cfg, v.dev.veh_i, v.mask);
    else
        k_done_mask<<<(v.cfg.n_envs + 255) / 256, 256, 0, st>>>(v.cfg, terminated_dev, truncated_dev, v.mask);
    sim->launches++;
    CK(cudaGetLastError());
    return reset_impl(sim, v, v.mask, obs_dev, st, whole_batch);
}
static int step_autoreset_impl(md_sim* sim, View& v, const float* actions_dev, StepOut out, cudaStream_t st, bool whole_batch) {
    if (!v.cfg.is_multi_agent && out.term && out.trunc) return step_impl(sim, v, actions_dev, out, st, true, whole_batch);
    if (step_impl(sim, v, actions_dev, out, st, false, whole_batch)) return -1;
    return autoreset_impl(sim, v, out.term, out.trunc, out.obs, st, whole_batch);
}

extern "C" int md_step(md_sim* sim, const float* actions_dev, float* obs_dev, float* reward_dev, float* cost_dev,
                       uint8_t* terminated_dev, uint8_t* truncated_dev, int32_t* info_flags_dev, float* info_f_dev, void* stream) {
    if (!sim || !sim->loaded) return -2;
    CK(cudaSetDevice(sim->device));
    StepOut out = {obs_dev, reward_dev, cost_dev, terminated_dev, truncated_dev, info_flags_dev, info_f_dev};
    return step_impl(sim, sim->all, actions_dev, out, (cudaStream_t)stream, false, true);
}

extern "C" int md_step_autoreset(md_sim* sim, const float* actions_dev, float* obs_dev, float* reward_dev, float* cost_dev,
                                 uint8_t* terminated_dev, uint8_t* truncated_dev, int32_t* info_flags_dev, float* info_f_dev,
                                 void* stream) {
    if (!sim || !sim->loaded) return -2;
    CK(cudaSetDevice(sim->device));
    StepOut out = {obs_dev, reward_dev, cost_dev, terminated_dev, truncated_dev, info_flags_dev, info_f_dev};
    // The launch sequence of a step is fixed (counters live on the device, nothing synchronises or allocates): replay it as a
    // CUDA graph unless the per-stage profiling events are being recorded.  The graph is captured on an internal stream (the
    // caller's may be the legacy default stream, which cannot be captured) and re-captured when a caller pointer or anything
    // that shapes the sequence changes.  MD_DEV_GRAPH=0 launches kernel by kernel.
    static const int use_graph = env_int("MD_DEV_GRAPH", 1);
    if (use_graph && !(sim->prof_n < sim->prof_cap) && sim->dg_misses < 4) {
        const uint64_t key[12] = {(uint64_t)(uintptr_t)actions_dev, (uint64_t)(uintptr_t)obs_dev, (uint64_t)(uintptr_t)reward_dev,
                                  (uint64_t)(uintptr_t)cost_dev, (uint64_t)(uintptr_t)terminated_dev, (uint64_t)(uintptr_t)truncated_dev,
                                  (uint64_t)(uintptr_t)info_flags_dev, (uint64_t)(uintptr_t)info_f_dev, (uint64_t)(uintptr_t)sim->bank,
                                  (uint64_t)(uintptr_t)sim->all.contact_tab, (uint64_t)sim->bank_seed, sim->post_valid ? 1ull : 0ull};
        if (sim->dg_exec == nullptr || memcmp(key, sim->dg_key, sizeof(key)) != 0) {
            // a capture costs ~100 us: worth it only if the same buffers come back.  Four re-captures without a replay in
            // between (a caller that allocates its action tensor anew every step) switch the handle to plain launches.
            if (sim->dg_exec) { sim->dg_misses = sim->dg_hits > 0 ? 1 : sim->dg_misses + 1; sim->dg_hits = 0; }
            if (sim->dg_exec) { cudaGraphExecDestroy(sim->dg_exec); sim->dg_exec = nullptr; }
            if (!sim->dg_stream) CK(cudaStreamCreateWithFlags(&sim->dg_stream, cudaStreamNonBlocking));
            const int64_t l0 = sim->launches;
            cudaGraph_t graph = nullptr;
            CK(cudaStreamBeginCapture(sim->dg_stream, cudaStreamCaptureModeThreadLocal));
            const int rc = step_autoreset_impl(sim, sim->all, actions_dev, out, sim->dg_stream, false);
            const cudaError_t ce = cudaStreamEndCapture(sim->dg_stream, &graph);
            sim->dg_launches = (int)(sim->launches - l0);
            sim->launches = l0;
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            CK(ce);
            CK(cudaGraphInstantiate(&sim->dg_exec, graph, 0));
            CK(cudaGraphDestroy(graph));
            memcpy(sim->dg_key, key, sizeof(key));
        }
        else sim->dg_hits++;
        CK(cudaGraphLaunch(sim->dg_exec, (cudaStream_t)stream));
        sim->launches += sim->dg_launches;
        return 0;
    }
    return step_autoreset_impl(sim, sim->all, actions_dev, out, (cudaStream_t)stream, true);
}

extern "C" int md_autoreset(md_sim* sim, const uint8_t* terminated_dev, const uint8_t* truncated_dev, float* obs_dev, void* stream) {
    if (!sim || !sim->loaded) return -2;
    CK(cudaSetDevice(sim->device));
    return autoreset_impl(sim, sim->all, terminated_dev, truncated_dev, obs_dev, (cudaStream_t)stream, true);
}

extern "C" int md_attach_bank(md_sim* sim, md_sim* bank, int seed) {
    if (!sim || !sim->loaded) return -2;
    if (!bank) { sim->bank = nullptr; return 0; }
    if (!bank->loaded || !bank->post_valid) { sim->err = "the scenario bank must be loaded and fully reset (md_reset with a NULL mask)"; return -2; }
    const MdConfig &a = sim->cfg, &b = bank->cfg;
    if (bank->device != sim->device || a.slots_per_env != b.slots_per_env || a.objs_per_env != b.objs_per_env ||
        a.agents_per_env != b.agents_per_env || a.n_lasers != b.n_lasers || a.n_side_lasers != b.n_side_lasers ||
        a.n_lane_lasers != b.n_lane_lasers || a.num_others != b.num_others || a.add_others_navi != b.add_others_navi) {
        sim->err = "scenario bank: device, slots / objects / agents per env and the observation layout must match";
        return -2;
    }
    if (a.is_multi_agent || a.traffic_mode != 0 || b.traffic_mode != 0) {
        sim->err = "scenario bank: single-agent, trigger-mode worlds only (respawn-mode tables are per env)";
        return -2;
    }
    for (int i = 0; i < 11; i++)   // the shared map tables (map_desc ... grid_items) and the lane grids
        if (sim->rows[i] != bank->rows[i]) { sim->err = "scenario bank: both handles must load the same map set (same map ids)"; return -2; }
    if (sim->rows[21] != bank->rows[21] || sim->rows[22] != bank->rows[22]) {
        sim->err = "scenario bank: both handles must load the same map set (same map ids)";
        return -2;
    }
    sim->bank = bank;
    sim->bank_ever = true;
    sim->bank_seed = (uint32_t)seed;
    // every view's draw counter restarts
    CK(cudaMemset2D(sim->d_pass, 2 * sizeof(uint32_t), 0, sizeof(uint32_t), 1 + MAX_HOST_GROUPS));
    return 0;
}

// Contact export.  With it on, every step records per vehicle slot the set of bodies its chassis touched during the step's
// sub-steps - bit k < slots_per_env = vehicle slot k of the same env, bit slots_per_env + j = object j - i.e. the pairs
// the reference's contact-added callback sees (engine/core/collision_callback.py:5-42).  md_get_contacts copies the table
// ([NV, 4] uint32, 128 bits per slot) to the host; synchronous.
extern "C" int md_enable_contacts(md_sim* sim, int on) {
    if (!sim || !sim->loaded) return -2;
    CK(cudaSetDevice(sim->device));
    CK(cudaDeviceSynchronize());
    for (HostGroup& g : sim->groups)
        if (g.gexec) { cudaGraphExecDestroy(g.gexec); g.gexec = nullptr; }   // the captured launches hold the old pointer
    const size_t NV = (size_t)sim->cfg.n_envs * sim->cfg.slots_per_env;
    if (on && !sim->contact_tab) {
        CK(cudaMalloc(&sim->contact_tab, NV * sizeof(uint4)));
        CK(cudaMemset(sim->contact_tab, 0, NV * sizeof(uint4)));
    } else if (!on && sim->contact_tab) {
        CK(cudaFree(sim->contact_tab));
        sim->contact_tab = nullptr;
    }
    sim->all.contact_tab = sim->contact_tab;
    for (HostGroup& g : sim->groups)
        g.v.contact_tab = sim->contact_tab ? sim->contact_tab + (size_t)g.env0 * sim->cfg.slots_per_env : nullptr;
    return 0;
}
extern "C" int md_get_contacts(md_sim* sim, uint32_t* host_dst, size_t bytes) {
    if (!sim || !sim->loaded) return -2;
    const size_t NV = (size_t)sim->cfg.n_envs * sim->cfg.slots_per_env;
    if (!sim->contact_tab || bytes != NV * sizeof(uint4)) { sim->err = "md_get_contacts: not enabled, or a wrong size"; return -7; }
    CK(cudaSetDevice(sim->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(host_dst, sim->contact_tab, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

// per-kernel device timing of the next `max_steps` md_step calls (events on the launch stream, no sync added)
extern "C" int md_profile_begin(md_sim* sim, int max_steps) {
    if (!sim || !sim->loaded) return -2;
    CK(cudaSetDevice(sim->device));
    while ((int)sim->prof_ev.size() < N_PROF_EV * max_steps) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        sim->prof_ev.push_back(e);
    }
    sim->prof_cap = max_steps;
    sim->prof_n = 0;
    return 0;
}
// after the caller synchronised: ms[5*i + k] = duration of stage k (k_pre, k_dyn, k_post, fused reset, k_lidar) of step i
extern "C" int md_profile_end(md_sim* sim, float* ms, int cap) {
    if (!sim || !sim->loaded) return -2;
    CK(cudaSetDevice(sim->device));
    int n = sim->prof_n < cap ? sim->prof_n : cap;
    for (int i = 0; i < n; i++) {
        CK(cudaEventSynchronize(sim->prof_ev[N_PROF_EV * i + 5]));
        for (int k = 0; k < 5; k++)
            CK(cudaEventElapsedTime(&ms[5 * i + k], sim->prof_ev[N_PROF_EV * i + k], sim->prof_ev[N_PROF_EV * i + k + 1]));
    }
    sim->prof_cap = 0;
    sim->prof_n = 0;
    return n;
}

// ---- TopDownObservation (obs/top_down_obs.py:98-200; obs/top_down_obs_impl.py:19-97, 203-250, 266-428): the ego-centred
// bird's-eye RGB image of TopDownSingleFrameMetaDriveEnv (envs/top_down_env.py:7-31), [A, res, res, 3] in [0, 1].  The scene the
// reference paints with pygame - lane lines (35, 35, 35) 0.3 m wide, the ego GREEN, every other vehicle BLUE (headings under 2
// degrees snapped to 0), the window of +-max_distance turned so that the ego looks up, its left on the image's right - is
// evaluated analytically at the centre of every output pixel (the CPU restatement mdo_topdown under oracle/ states the rules; pygame is not
// on this image, its rasteriser is not pinned).  One CTA per row of 16 x 16 pixel tiles of one agent: the vehicles' rectangles
// are staged in shared memory once, the line segments near each tile (through the static grid) once per tile, then every thread
// shades its pixel.
#define TD_TILE 16
#define TD_MAX_LINES 768
#define TD_LINE_HW 0.15f
#define TD_SNAP_TAN 0.03492077f
__device__ __forceinline__ Rect td_snap(Rect r) {
    if (r.ux > 0.0f && fabsf(r.uy) <= TD_SNAP_TAN * r.ux) { r.ux = 1.0f; r.uy = 0.0f; }
    return r;
}
__device__ __forceinline__ bool td_inside(const Rect& r, float x, float y) {
    const float dx = x - r.cx, dy = y - r.cy;
    return fabsf(dx * r.ux + dy * r.uy) <= r.hu && fabsf(dy * r.ux - dx * r.uy) <= r.hv;
}
// box-filtered coverage of the pixel (side px, centre (x, y)) by a segment's 2 * TD_LINE_HW wide box: the product of the
// overlaps across and along the segment, each as a fraction of a pixel.  Most segments near a tile miss a given pixel across:
// that test comes first and costs five operations.
__device__ __forceinline__ float td_cover(float cx, float cy, float half, float ux, float uy, float x, float y, float px, float inv_px) {
    const float dx = x - cx, dy = y - cy;
    const float tp = (TD_LINE_HW + 0.5f * px) - fabsf(dy * ux - dx * uy);
    if (tp <= 0.0f) return 0.0f;
    const float ta = (half + 0.5f * px) - fabsf(dx * ux + dy * uy);
    if (ta <= 0.0f) return 0.0f;
    return fminf(tp * inv_px, 1.0f) * fminf(ta * inv_px, 1.0f);
}
// CH = 3: the RGB image of TopDownObservation.  CH = 2: the two per-frame grey channels TopDownMultiChannel stacks
// (obs/top_down_obs_multi_channel.py:101-146, 235-270): [road_network = lane lines (35) over the drivable area of the route's
// lanes (64), doubled and clipped as observe() does; traffic_flow = the other vehicles in the grey of ObjectGraphics.BLUE]
#define TD_MAX_ROUTE_LANES 160
template <int CH>
__global__ void __launch_bounds__(TD_TILE * TD_TILE)
k_topdown(MdConfig cfg, MdArrays A, MapAccel X, float* __restrict__ img, int res, float max_distance) {
    __shared__ Rect s_other[128];
    __shared__ int s_base[128], s_alive[128];
    __shared__ Rect s_ego;
    __shared__ float s_line[TD_MAX_LINES * 5];
    __shared__ int s_n;
    __shared__ int s_rl[CH == 2 ? TD_MAX_ROUTE_LANES : 1], s_rl_on[CH == 2 ? TD_MAX_ROUTE_LANES : 1];
    __shared__ float4 s_rl_bb[CH == 2 ? TD_MAX_ROUTE_LANES : 1];
    __shared__ int s_nrl;
    const int S = cfg.slots_per_env, NA = cfg.agents_per_env;
    const int a = blockIdx.x, env = a / NA, slot = a - env * NA;
    const int tiles = (res + TD_TILE - 1) / TD_TILE, ty = blockIdx.y;   // one CTA per ROW of tiles: the agent's set-up is paid once
    const int r = ty * TD_TILE + threadIdx.x / TD_TILE;
    const size_t g0 = (size_t)env * S;
    float* row = img + ((size_t)a * res * res + (size_t)r * res) * CH;
    if (!A.veh_i[(g0 + slot) * VEH_I + VI_ALIVE]) {   // an empty seat: a black image
        if (r < res)
            for (int c = threadIdx.x & (TD_TILE - 1); c < res; c += TD_TILE)
                for (int k = 0; k < CH; k++) row[CH * c + k] = 0.0f;
        return;
    }
    if (threadIdx.x == 0) s_nrl = 0;
    for (int k = threadIdx.x; k < S; k += blockDim.x) {
        const size_t g = g0 + k;
        const int alive = A.veh_i[g * VEH_I + VI_ALIVE];
        s_base[k] = alive && k != slot;
        if (alive) {
            float P[VEH_P], St[VEH_S];
            load16(P, A.veh_p + g * VEH_P);
            load16(St, A.veh_s + g * VEH_S);
            const Rect rr = vehicle_rect(P, St);
            if (k == slot) s_ego = rr;
            else s_other[k] = td_snap(rr);
        }
    }
    const float px = 2.0f * max_distance / (float)res, inv_px = 1.0f / px;
    const MapView m = map_view(A, A.env_i[env * ENV_I + EI_MAP], &X);
    __syncthreads();
    if (CH == 2) {   // the lanes of the route's roads (draw_navigation_node, top_down_obs_multi_channel.py:272-277) and their boxes
        const int n_seg = A.veh_i[(g0 + slot) * VEH_I + VI_ROUTE_LEN] - 1;
        const int* rroad = A.veh_rroad + (g0 + slot) * ROUTE_MAX;
        for (int k = threadIdx.x; k < n_seg; k += blockDim.x) {
            const int rd = rroad[k];
            if (rd < 0) continue;
            const int first = m.road_i[rd * ROAD_I + RI_FIRST], nl = m.road_i[rd * ROAD_I + RI_N];
            const int j0 = atomicAdd(&s_nrl, nl);
            for (int j = 0; j < nl && j0 + j < TD_MAX_ROUTE_LANES; j++) {
                s_rl[j0 + j] = first + j;
                s_rl_bb[j0 + j] = __ldg(reinterpret_cast<const float4*>(m.lane_bb) + first + j);
            }
        }
        __syncthreads();
    }
    const int n_rl = CH == 2 ? min(s_nrl, TD_MAX_ROUTE_LANES) : 0;
    const Rect ego = s_ego, ego_box = td_snap(ego);
    const float vc = (0.5f * (float)res - ((float)(ty * TD_TILE) + 0.5f * TD_TILE)) * px;
    const float v = (0.5f * (float)res - ((float)r + 0.5f)) * px;
    // a tile's radius: every pixel centre of it, plus a line's reach (half width + one pixel)
    const float reach = 0.70710678f * TD_TILE * px + TD_LINE_HW + px;
    for (int tx = 0; tx < tiles; tx++) {
        const float uc = ((float)(tx * TD_TILE) + 0.5f * TD_TILE - 0.5f * (float)res) * px;
        const float tcx = ego.cx + (vc * ego.ux - uc * ego.uy), tcy = ego.cy + (vc * ego.uy + uc * ego.ux);
        if (threadIdx.x == 0) s_n = 0;
        for (int k = threadIdx.x; k < S; k += blockDim.x) {   // vehicles that cannot reach into this tile are dropped here
            int near = s_base[k];
            if (near) {
                const Rect& q = s_other[k];
                const float dx = q.cx - tcx, dy = q.cy - tcy, rr = 0.70710678f * TD_TILE * px + q.hu + q.hv;
                near = dx * dx + dy * dy <= rr * rr;
            }
            s_alive[k] = near;
        }
        if (CH == 2)   // route lanes whose box comes near this tile
            for (int k = threadIdx.x; k < n_rl; k += blockDim.x) {
                const float4 bb = s_rl_bb[k];
                const float R = 0.70710678f * TD_TILE * px;
                s_rl_on[k] = !(tcx + R < bb.x || tcx - R > bb.z || tcy + R < bb.y || tcy - R > bb.w);
            }
        __syncthreads();
        int x0 = (int)floorf((tcx - reach - m.gx0) / m.cell), x1 = (int)floorf((tcx + reach - m.gx0) / m.cell);
        int y0 = (int)floorf((tcy - reach - m.gy0) / m.cell), y1 = (int)floorf((tcy + reach - m.gy0) / m.cell);
        x0 = max(x0, 0); y0 = max(y0, 0); x1 = min(x1, m.nx - 1); y1 = min(y1, m.ny - 1);
        // one warp per grid cell under the tile, its lanes over the cell's items; the item's row sits next to the item in the
        // grid records (MapAccel), so a lane waits for the cell's range and then for ONE round of loads
        const int ncx = x1 - x0 + 1, n_cells = ncx * (y1 - y0 + 1);
        for (int ci = threadIdx.x >> 5; ci < n_cells; ci += (TD_TILE * TD_TILE) >> 5) {
            const int cell = (y0 + ci / ncx) * m.nx + x0 + ci % ncx;
            const int k1 = m.gs[cell + 1];
            for (int k = m.gs[cell] + (threadIdx.x & 31); k < k1; k += 32) {
                const int it = __ldg(m.gi + k);
                const float4 h = __ldg(m.irec + 2 * k), d = __ldg(m.irec + 2 * k + 1);   // cx cy half kind | ux uy . .
                if (it >= m.n_lines) continue;
                const float dx = h.x - tcx, dy = h.y - tcy, rr = reach + h.z;
                if (dx * dx + dy * dy > rr * rr) continue;
                const int j = atomicAdd(&s_n, 1);
                if (j < TD_MAX_LINES) {
                    s_line[5 * j] = h.x; s_line[5 * j + 1] = h.y; s_line[5 * j + 2] = h.z; s_line[5 * j + 3] = d.x; s_line[5 * j + 4] = d.y;
                }
            }
        }
        __syncthreads();
        const int c = tx * TD_TILE + (threadIdx.x & (TD_TILE - 1));
        if (r < res && c < res) {
            float* o = row + CH * c;
            const float u = ((float)c + 0.5f - 0.5f * (float)res) * px;
            const float x = ego.cx + (v * ego.ux - u * ego.uy), y = ego.cy + (v * ego.uy + u * ego.ux);
            bool hit = false;
            for (int k = 0; k < S && !hit; k++) hit = s_alive[k] && td_inside(s_other[k], x, y);
            const bool on_ego = CH == 3 && !hit && td_inside(ego_box, x, y);
            float best = 0.0f;
            if (CH == 2 || (!hit && !on_ego)) {
                const int n = s_n;
                if (n <= TD_MAX_LINES) {
                    for (int k = 0; k < n; k++)
                        best = fmaxf(best, td_cover(s_line[5 * k], s_line[5 * k + 1], s_line[5 * k + 2], s_line[5 * k + 3], s_line[5 * k + 4], x, y, px, inv_px));
                } else {   // more segments around the tile than the staging area holds: walk the grid cells directly
                    for (int cy = y0; cy <= y1; cy++)
                        for (int cx = x0; cx <= x1; cx++) {
                            const int cell = cy * m.nx + cx;
                            for (int k = m.gs[cell]; k < m.gs[cell + 1]; k++) {
                                const int it = m.gi[k];
                                if (it >= m.n_lines) continue;
                                const float* Ln = m.lines + (size_t)it * LINE_F;
                                best = fmaxf(best, td_cover(Ln[LN_CX], Ln[LN_CY], Ln[LN_HALF], Ln[LN_UX], Ln[LN_UY], x, y, px, inv_px));
                            }
                        }
                }
            }
            if (CH == 3) {
                if (hit) { o[0] = 100.0f / 255.0f; o[1] = 200.0f / 255.0f; o[2] = 1.0f; }           // ObjectGraphics.BLUE
                else if (on_ego) { o[0] = 50.0f / 255.0f; o[1] = 200.0f / 255.0f; o[2] = 0.0f; }   // GREEN
                else o[0] = o[1] = o[2] = best * (35.0f / 255.0f);                                 // WorldSurface.LANE_LINE_COLOR
            } else {
                bool area = false;
                for (int k = 0; k < n_rl && !area; k++) {
                    if (!s_rl_on[k]) continue;
                    const float4 bb = s_rl_bb[k];
                    if (x < bb.x || y < bb.y || x > bb.z || y > bb.w) continue;
                    const int l = s_rl[k];
                    area = point_in_hull(m.lane_f + l * LANE_F, m.hull + 2 * m.lane_i[l * LANE_I + LI_HULL_OFF], m.lane_i[l * LANE_I + LI_HULL_N], x, y);
                }
                // the line is painted over the area; observe() doubles the road channel and clips it
                const float road = best * (35.0f / 255.0f) + (1.0f - best) * (area ? 64.0f / 255.0f : 0.0f);
                o[0] = fminf(road * 2.0f, 1.0f);
                o[1] = hit ? 176.37f / 255.0f : 0.0f;   // 0.299 * 100 + 0.587 * 200 + 0.114 * 255 (_transform, :216-227)
            }
        }
        __syncthreads();   // the next tile restages s_line / s_alive
    }
}

static int topdown_impl(md_sim* sim, float* img_dev, int resolution, float max_distance, void* stream, int channels) {
    if (!sim || !sim->loaded) return -2;
    if (img_dev == nullptr || resolution <= 0 || !(max_distance > 0.0f) || sim->cfg.slots_per_env > 128) {
        sim->err = "md_topdown: needs an output buffer, resolution > 0, max_distance > 0 and at most 128 slots per env";
        return -3;
    }
    CK(cudaSetDevice(sim->device));
    const int tiles = (resolution + TD_TILE - 1) / TD_TILE;
    dim3 grid((unsigned)(sim->cfg.n_envs * sim->cfg.agents_per_env), (unsigned)tiles);
    if (channels == 3)
        k_topdown<3><<<grid, TD_TILE * TD_TILE, 0, (cudaStream_t)stream>>>(sim->cfg, sim->dev, sim->accel, img_dev, resolution, max_distance);
    else
        k_topdown<2><<<grid, TD_TILE * TD_TILE, 0, (cudaStream_t)stream>>>(sim->cfg, sim->dev, sim->accel, img_dev, resolution, max_distance);
    sim->launches++;
    CK(cudaGetLastError());
    return 0;
}
extern "C" int md_topdown(md_sim* sim, float* img_dev, int resolution, float max_distance, void* stream) {
    return topdown_impl(sim, img_dev, resolution, max_distance, stream, 3);
}
extern "C" int md_topdown_channels(md_sim* sim, float* img_dev, int resolution, float max_distance, void* stream) {
    return topdown_impl(sim, img_dev, resolution, max_distance, stream, 2);
}

extern "C" int md_lidar(md_sim* sim, float* frac_dev, int32_t* hit_dev, void* stream) {
    if (!sim || !sim->loaded) return -2;
    CK(cudaSetDevice(sim->device));
    cudaStream_t st = (cudaStream_t)stream;
    // refresh the body rows from the current state without moving anything
    StepOut out = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (launch_post(sim, sim->all, 0, out, nullptr, st)) return -1;
    return launch_lidar(sim, sim->all, 0, frac_dev, sim->cfg.n_lasers, -1, hit_dev, nullptr, st);
}
extern "C" int md_dynamics(md_sim* sim, const float* act3_dev, int n_sub, void* stream) {
    if (!sim || !sim->loaded) return -2;
    CK(cudaSetDevice(sim->device));
    return launch_dyn(sim, sim->all, MODE_DYN | MODE_EXT_ACT, act3_dev, n_sub, (cudaStream_t)stream);
}
extern "C" int md_after_step(md_sim* sim, void* stream) {
    if (!sim || !sim->loaded) return -2;
    CK(cudaSetDevice(sim->device));
    StepOut out = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    return launch_post(sim, sim->all, MODE_POST | MODE_CLEAR_FLAGS, out, nullptr, (cudaStream_t)stream);
}
extern "C" int md_idm(md_sim* sim, float* out_actions_dev, void* stream) {
    if (!sim || !sim->loaded) return -2;
    CK(cudaSetDevice(sim->device));
    return launch_pre(sim, sim->all, MODE_IDM_OUT, nullptr, out_actions_dev, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------ host-buffer path
// (Re)partition the batch into n_groups contiguous env ranges for the host-buffer entry points.  Every group owns a
// stream, a packed output block on the device with a pinned mirror of the same layout, and its rows of the handle-wide
// action / observation staging buffers.  Earlier groups get the higher stream priority, so inside one md_step_host call
// group k's kernels finish - and its D2H copy starts - while group k+1 is still computing.
static size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }
extern "C" int md_host_groups(md_sim* sim, int n_groups) {
    if (!sim || !sim->loaded) return -2;
    CK(cudaSetDevice(sim->device));
    const MdConfig& c = sim->cfg;
    if (n_groups < 1) n_groups = 1;
    if (n_groups > MAX_HOST_GROUPS) n_groups = MAX_HOST_GROUPS;
    if (n_groups > c.n_envs) n_groups = c.n_envs;
    for (HostGroup& g : sim->groups)
        if (g.in_flight) { sim->err = "md_host_groups: a group is in flight (md_host_recv it first)"; return -8; }
    // the new groups continue the counters of group 0 (of the whole-batch view when there were no groups yet)
    uint32_t ctr[2];
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(ctr, sim->d_pass + (sim->groups.empty() ? 0 : 2), sizeof(ctr), cudaMemcpyDeviceToHost));
    free_groups(sim);
    int lo_prio = 0, hi_prio = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));   // numerically lower = higher priority
    const size_t od = OBS_DIM(c);
    sim->groups.resize(n_groups);
    for (int k = 0; k < n_groups; k++) {
        HostGroup& g = sim->groups[k];
        g.env0 = (int)((long long)c.n_envs * k / n_groups);
        g.n_envs = (int)((long long)c.n_envs * (k + 1) / n_groups) - g.env0;
        g.na = (size_t)g.n_envs * c.agents_per_env;
        g.v = make_view(sim, g.env0, g.n_envs, 1 + k);
        CK(cudaMemcpy(g.v.pass, ctr, sizeof(ctr), cudaMemcpyHostToDevice));
        int prio = hi_prio + k;
        if (prio > lo_prio) prio = lo_prio;
        CK(cudaStreamCreateWithPriority(&g.stream, cudaStreamNonBlocking, prio));
        CK(cudaEventCreateWithFlags(&g.done, cudaEventDisableTiming));
        const size_t n = g.na;
        g.off_cost = align16(n * 4);
        g.off_flags = g.off_cost + align16(n * 4);
        g.off_info_f = g.off_flags + align16(n * 4);
        g.off_term = g.off_info_f + align16(n * 32);
        g.off_trunc = g.off_term + align16(n);
        g.off_count = g.off_trunc + align16(n);
        g.blk_bytes = g.off_count + 16;
        CK(cudaMalloc(&g.d_blk, g.blk_bytes));
        CK(cudaMemset(g.d_blk, 0, g.blk_bytes));
        CK(cudaMallocHost(&g.h_blk, g.blk_bytes));
        memset(g.h_blk, 0, g.blk_bytes);
        const size_t a0 = (size_t)g.env0 * c.agents_per_env;
        g.d_actions = sim->d_actions + a0 * 2; g.h_actions = sim->h_actions + a0 * 2;
        g.h_obs = sim->h_obs + a0 * od;
        g.out.obs = sim->d_obs + a0 * od;
        g.out.reward = (float*)g.d_blk; g.out.cost = (float*)(g.d_blk + g.off_cost);
        g.out.info_flags = (int*)(g.d_blk + g.off_flags); g.out.info_f = (float*)(g.d_blk + g.off_info_f);
        g.out.term = g.d_blk + g.off_term; g.out.trunc = g.d_blk + g.off_trunc;
        g.d_cobs = nullptr; g.d_block_counts = nullptr;
        g.gexec = nullptr; g.gkey = 0; g.glaunches = 0;
        g.in_flight = false; g.pending_compact = false;
    }
    return 0;
}
extern "C" int md_host_group_count(const md_sim* sim) { return sim && sim->loaded ? (int)sim->groups.size() : 0; }

// Multi-agent handles: with compact != 0 only the observation rows of the seats that produced a transition (FL_VALID in
// info_flags, newborn seats included) are copied to the host, packed in ascending seat order at the start of the group's
// obs rows; the scalars always come back for every seat, so the host finds row j's seat as the j-th FL_VALID entry.
extern "C" int md_host_compact(md_sim* sim, int compact) {
    if (!sim || !sim->loaded) return -2;
    if (compact && !sim->cfg.is_multi_agent) { sim->err = "md_host_compact: multi-agent handles only"; return -2; }
    sim->compact = compact ? 1 : 0;
    return 0;
}

// out[0..7] = addresses of group `group`'s pinned buffers: obs rows, reward, cost, terminated, truncated, info_flags,
// info_f, actions; range[0..2] = first env, env count, valid-row count of the last received step (compact mode, else all)
extern "C" int md_host_group_views(md_sim* sim, int group, void** out8, int* range3) {
    if (!sim || !sim->loaded || group < 0 || group >= (int)sim->groups.size()) return -2;
    HostGroup& g = sim->groups[group];
    if (out8) {
        out8[0] = g.h_obs; out8[1] = g.h_blk; out8[2] = g.h_blk + g.off_cost; out8[3] = g.h_blk + g.off_term;
        out8[4] = g.h_blk + g.off_trunc; out8[5] = g.h_blk + g.off_flags; out8[6] = g.h_blk + g.off_info_f; out8[7] = g.h_actions;
    }
    if (range3) {
        range3[0] = g.env0; range3[1] = g.n_envs;
        range3[2] = sim->compact ? (int)*(unsigned int*)(g.h_blk + g.off_count) : (int)g.na;
    }
    return 0;
}
// the handle-wide view (group 0's buffers; with one group - the default - they cover the whole batch)
extern "C" int md_host_views(md_sim* sim, void** out8) { return md_host_group_views(sim, 0, out8, nullptr); }

// the stream work of one group step: H2D of its actions, the kernels, D2H of the packed scalars and of the observation rows
static int enqueue_group_step(md_sim* sim, HostGroup& g, int autoreset) {
    const size_t od = OBS_DIM(sim->cfg);
    CK(cudaMemcpyAsync(g.d_actions, g.h_actions, g.na * 2 * 4, cudaMemcpyHostToDevice, g.stream));
    // auto-reset overwrites only the observation rows of finished envs; the scalars keep the finished step's values
    if (autoreset ? step_autoreset_impl(sim, g.v, g.d_actions, g.out, g.stream, false)
                  : step_impl(sim, g.v, g.d_actions, g.out, g.stream, false, false))
        return -1;
    if (sim->compact) {
        const int nb = (int)((g.na + COMPACT_T - 1) / COMPACT_T);
        k_valid_count<<<nb, COMPACT_T, 0, g.stream>>>((long long)g.na, g.out.info_flags, g.d_block_counts);
        k_valid_gather<<<nb, COMPACT_T, 0, g.stream>>>((long long)g.na, g.out.info_flags, g.d_block_counts, (int)od, g.out.obs, g.d_cobs,
                                                       (unsigned int*)(g.d_blk + g.off_count));
        sim->launches += 2;
        CK(cudaGetLastError());
    }
    CK(cudaMemcpyAsync(g.h_blk, g.d_blk, g.blk_bytes, cudaMemcpyDeviceToHost, g.stream));
    if (!sim->compact) CK(cudaMemcpyAsync(g.h_obs, g.out.obs, g.na * od * 4, cudaMemcpyDeviceToHost, g.stream));
    return 0;
}
// enqueue one env.step of a group on its stream and return without waiting; md_host_recv completes it.
// actions == NULL: the caller already wrote them into the group's pinned action buffer.
extern "C" int md_host_send(md_sim* sim, int group, const float* actions, int autoreset) {
    if (!sim || !sim->loaded || group < 0 || group >= (int)sim->groups.size()) return -2;
    CK(cudaSetDevice(sim->device));
    HostGroup& g = sim->groups[group];
    if (g.in_flight) { sim->err = "md_host_send: the group is already in flight"; return -8; }
    if (actions && actions != g.h_actions) memcpy(g.h_actions, actions, g.na * 2 * 4);
    if (sim->compact && !g.d_cobs) {
        CK(cudaMalloc(&g.d_cobs, g.na * OBS_DIM(sim->cfg) * 4));
        CK(cudaMalloc(&g.d_block_counts, sizeof(unsigned int) * ((g.na + COMPACT_T - 1) / COMPACT_T)));
    }
    static const int use_graph = env_int("MD_HOST_GRAPH", 1);
    if (use_graph) {
        // what shapes the launch sequence of a step: auto-reset or not, compaction, the bank, whether resets are row copies
        const uint64_t key = 1ull | (autoreset ? 2ull : 0ull) | (sim->compact ? 4ull : 0ull) | (sim->post_valid ? 8ull : 0ull) |
                             ((uint64_t)(uintptr_t)sim->bank << 4);
        if (g.gexec == nullptr || g.gkey != key) {
            if (g.gexec) { cudaGraphExecDestroy(g.gexec); g.gexec = nullptr; }
            const int64_t l0 = sim->launches;
            cudaGraph_t graph = nullptr;
            CK(cudaStreamBeginCapture(g.stream, cudaStreamCaptureModeThreadLocal));
            const int rc = enqueue_group_step(sim, g, autoreset);
            const cudaError_t ce = cudaStreamEndCapture(g.stream, &graph);
            g.glaunches = (int)(sim->launches - l0);
            sim->launches = l0;
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            CK(ce);
            CK(cudaGraphInstantiate(&g.gexec, graph, 0));
            CK(cudaGraphDestroy(graph));
            g.gkey = key;
        }
        CK(cudaGraphLaunch(g.gexec, g.stream));
        sim->launches += g.glaunches;
    } else if (enqueue_group_step(sim, g, autoreset)) return -1;
    g.pending_compact = sim->compact != 0;
    CK(cudaEventRecord(g.done, g.stream));
    g.in_flight = true;
    return 0;
}
// wait for a group's step; afterwards its pinned buffers hold the results (md_host_group_views)
extern "C" int md_host_recv(md_sim* sim, int group) {
    if (!sim || !sim->loaded || group < 0 || group >= (int)sim->groups.size()) return -2;
    CK(cudaSetDevice(sim->device));
    HostGroup& g = sim->groups[group];
    if (!g.in_flight) { sim->err = "md_host_recv: nothing in flight for this group"; return -8; }
    CK(cudaEventSynchronize(g.done));
    if (g.pending_compact) {   // the number of valid rows came back with the scalars: fetch exactly those rows
        const size_t cnt = *(unsigned int*)(g.h_blk + g.off_count);
        if (cnt) {
            CK(cudaMemcpyAsync(g.h_obs, g.d_cobs, cnt * OBS_DIM(sim->cfg) * 4, cudaMemcpyDeviceToHost, g.stream));
            CK(cudaStreamSynchronize(g.stream));
        }
        g.pending_compact = false;
    }
    g.in_flight = false;
    return 0;
}

extern "C" int md_reset_host(md_sim* sim, const uint8_t* env_mask, float* obs) {
    if (!sim || !sim->loaded) return -2;
    CK(cudaSetDevice(sim->device));
    for (HostGroup& g : sim->groups)
        if (g.in_flight) { sim->err = "md_reset_host: a group is in flight (md_host_recv it first)"; return -8; }
    const MdConfig& c = sim->cfg;
    const size_t NA = (size_t)c.n_envs * c.agents_per_env, od = OBS_DIM(c);
    cudaStream_t st = sim->groups[0].stream;
    const uint8_t* dm = nullptr;
    if (env_mask) {
        memcpy(sim->h_mask, env_mask, (size_t)c.n_envs);
        CK(cudaMemcpyAsync(sim->d_mask_in, sim->h_mask, (size_t)c.n_envs, cudaMemcpyHostToDevice, st));
        dm = sim->d_mask_in;
    }
    if (reset_impl(sim, sim->all, dm, sim->d_obs, st, true)) return -1;
    CK(cudaMemcpyAsync(sim->h_obs, sim->d_obs, NA * od * 4, cudaMemcpyDeviceToHost, st));
    for (HostGroup& g : sim->groups) {   // a whole-batch pass advances every group's counters
        k_bump<<<1, 1, 0, st>>>(g.v.pass, sim->bank ? 1u : 0u, 1u);
        sim->launches++;
    }
    CK(cudaStreamSynchronize(st));
    if (obs) memcpy(obs, sim->h_obs, NA * od * 4);
    return 0;
}

// env.step of the whole batch through HOST buffers, synchronous: every group is sent, then every group is received.
// Output pointers may be NULL: the caller then reads the pinned buffers in place (md_host_group_views).
extern "C" int md_step_host(md_sim* sim, const float* actions, float* obs, float* reward, float* cost, uint8_t* terminated,
                            uint8_t* truncated, int32_t* info_flags, float* info_f, int autoreset) {
    if (!sim || !sim->loaded) return -2;
    const MdConfig& c = sim->cfg;
    const size_t od = OBS_DIM(c);
    const int G = (int)sim->groups.size();
    for (int k = 0; k < G; k++) {
        HostGroup& g = sim->groups[k];
        const size_t a0 = (size_t)g.env0 * c.agents_per_env;
        const float* a = nullptr;
        if (actions && actions != sim->h_actions) a = actions + a0 * 2;
        if (md_host_send(sim, k, a, autoreset)) return -1;
    }
    for (int k = 0; k < G; k++) {
        if (md_host_recv(sim, k)) return -1;
        HostGroup& g = sim->groups[k];
        const size_t a0 = (size_t)g.env0 * c.agents_per_env, n = g.na;
        if (obs && !sim->compact) memcpy(obs + a0 * od, g.h_obs, n * od * 4);
        if (reward) memcpy(reward + a0, g.h_blk, n * 4);
        if (cost) memcpy(cost + a0, g.h_blk + g.off_cost, n * 4);
        if (terminated) memcpy(terminated + a0, g.h_blk + g.off_term, n);
        if (truncated) memcpy(truncated + a0, g.h_blk + g.off_trunc, n);
        if (info_flags) memcpy(info_flags + a0, g.h_blk + g.off_flags, n * 4);
        if (info_f) memcpy(info_f + a0 * 8, g.h_blk + g.off_info_f, n * 32);
    }
    if (obs && sim->compact) { sim->err = "md_step_host: compact mode leaves the rows in the pinned buffers (obs must be NULL)"; return -2; }
    return 0;
}

// ---- FP32-FMA peak of this GPU, measured: the denominator of the secondary (ray-test) roofline.  Every
// thread runs 8 independent chains of explicit fmaf (FFMA whatever -fmad says); 2 flops per FMA.
__global__ void __launch_bounds__(256) k_fma_peak(float* __restrict__ out, int iters, float a, float b) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.0f, x2 = x0 + 2.0f, x3 = x0 + 3.0f, x4 = x0 + 4.0f, x5 = x0 + 5.0f, x6 = x0 + 6.0f, x7 = x0 + 7.0f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    const float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456f) out[0] = s;   // keeps the chains alive
}
extern "C" int md_fp32_peak(int device, double* tflops) {
    if (!tflops) return -2;
    if (cudaSetDevice(device) != cudaSuccess) return -3;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1;
    float* out = nullptr;
    if (cudaMalloc(&out, 16) != cudaSuccess) return -1;
    const int blocks = prop.multiProcessorCount * 8, iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 6; rep++) {   // the first repetitions warm the clocks up; the best one counts (burst figure)
        cudaEventRecord(e0, 0);
        k_fma_peak<<<blocks, 256>>>(out, iters, 0.999f, 1e-3f);
        cudaEventRecord(e1, 0);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); return -1; }
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 64.0 * (double)iters * 256.0 * (double)blocks;
        if (ms > 0.0f && flops / (ms * 1e-3) / 1e12 > best) best = flops / (ms * 1e-3) / 1e12;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return 0;
}
