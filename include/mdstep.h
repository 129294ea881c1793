/* mdstep.h — C ABI of libmdstep.so, the B200-native batched MetaDrive step.
 *
 * The reference (zhuhaozh/metadrive_ped = MetaDrive v0.4.2.2) exposes this path as a Python class surface, not
 * an FFI; the arithmetic lives behind panda3d.bullet.  Each entry point below names the reference interface it
 * replaces (paths relative to /root/reference/metadrive).  INTEGRATION.md shows the ctypes stub a maintainer of the
 * reference would add.
 *
 * Conventions: plain pointers and sizes only; all `*_dev` buffers are caller-owned CUDA device memory (e.g.
 * torch tensors' data_ptr()), contiguous, float32 / int32 / uint8; `stream` is a cudaStream_t passed as void*.
 * Calls enqueue kernels on `stream` and return without synchronising (except the *_host variants and
 * md_get_state / md_set_state, which are synchronous).  One handle per device; a handle is not thread-safe.
 * Return value: 0 on success, negative on error (md_last_error gives the text).
 */
#ifndef MDSTEP_H
#define MDSTEP_H

#include <stddef.h>
#include <stdint.h>

#include "md_layout.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct md_sim md_sim;

#define MD_ABI_VERSION 8
int md_abi_version(void);
/* sizeof(MdConfig) / sizeof(MdArrays) the library was built with: the loader compares them with its own mirror */
int md_sizeof_config(void);
int md_sizeof_arrays(void);

/* engine construction: replaces initialize_engine / BaseEngine.__init__ + PhysicsWorld
 * (engine/engine_utils.py:8-15, engine/base_engine.py:51-96, engine/core/physics_world.py:9-17) */
int md_create(const MdConfig* cfg, int device, md_sim** out);
void md_destroy(md_sim* sim);
const char* md_last_error(const md_sim* sim);

/* scene upload: replaces PGMap/BaseBlock.create_in_world + manager.reset() spawning bodies into the Bullet worlds
 * (component/pgblock/pg_block.py:248-256, component/block/base_block.py:431-519, manager/traffic_manager.py:51-72,
 * manager/object_manager.py:40-91).  `host` holds HOST pointers in the md_layout.h layouts; `rows[i]` is the row
 * count of the i-th array in MdArrays field order (28 arrays).  The library copies everything and keeps a device snapshot of the
 * mutable arrays as the reset state.  Caller keeps ownership of the host buffers. */
int md_load_scene(md_sim* sim, const MdArrays* host, const int64_t* rows);

/* env.reset(): replaces BaseEnv.reset -> engine.reset + _get_reset_return (envs/base_env.py:502-584).
 * Restores the snapshot of every env whose mask byte is non-zero (all envs if env_mask_dev == NULL), runs the
 * reset-time after_step and writes the first observation of their agents into obs_dev [A, OBS_DIM(cfg)].
 * With a scenario bank attached (md_attach_bank) the selected envs restart in a scenario DRAWN from the bank instead,
 * like BaseEnv.reset(seed=None) - the same draw an auto-reset makes. */
int md_reset(md_sim* sim, const uint8_t* env_mask_dev, float* obs_dev, void* stream);

/* env.step(): replaces BaseEnv.step = _step_simulator + _get_step_return (envs/base_env.py:426-463, 586-623):
 * engine.before_step (agent actuation, trigger, IDM), decision_repeat x doPhysics + contact callback,
 * engine.after_step (localisation, state check), reward / cost / done, LidarStateObservation.observe.
 * actions_dev [A,2]; obs_dev [A, OBS_DIM(cfg)] (19 + 4*num_others (8* with add_others_navi) + n_lasers by default); reward/cost [A] f32; terminated/truncated [A] u8;
 * info_flags [A] i32 (FL_* bits); info_f [A,8] f32 = velocity, steering, acceleration, step_energy, episode_energy,
 * step_reward, episode_reward, episode_length. */
int md_step(md_sim* sim, const float* actions_dev, float* obs_dev, float* reward_dev, float* cost_dev,
            uint8_t* terminated_dev, uint8_t* truncated_dev, int32_t* info_flags_dev, float* info_f_dev, void* stream);

/* batched auto-reset on device: every env whose agents are all terminated or truncated is restored to its snapshot
 * and its agents' observation rows are overwritten with the reset observation (the user-side loop
 * `if done: env.reset()` of examples/profile_metadrive.py:26-29, without a host round trip). */
int md_autoreset(md_sim* sim, const uint8_t* terminated_dev, const uint8_t* truncated_dev, float* obs_dev, void* stream);
/* md_step + md_autoreset in one call.  Single-agent worlds take a fused path: k_post marks the finished envs, one more
 * launch restores and re-localises them, and the lidar observes the post-reset world once (the reward / cost / done /
 * info outputs keep the finished step's values, the observation rows of finished envs hold the reset observation).
 * The launch sequence of a step is fixed (every counter lives on the device), so the call replays it as a CUDA graph while
 * the caller passes the same buffers (captured on an internal stream, launched into `stream`, which may be the legacy
 * default stream); buffers that change every step switch the handle back to plain launches.  MD_DEV_GRAPH=0 disables it. */
int md_step_autoreset(md_sim* sim, const float* actions_dev, float* obs_dev, float* reward_dev, float* cost_dev,
                      uint8_t* terminated_dev, uint8_t* truncated_dev, int32_t* info_flags_dev, float* info_f_dev, void* stream);

/* ---- the same step through HOST buffers (what the Gymnasium-surface classes call) ------------------------------------
 * The batch is partitioned into HOST GROUPS: contiguous env ranges, each with its own stream (earlier groups at higher
 * priority), a packed device output block with a pinned mirror of the same layout (reward | cost | info_flags | info_f |
 * terminated | truncated | valid-row count: ONE D2H copy for all scalars of a step) and its rows of the handle-wide pinned
 * action / observation buffers (one more copy).  One group (the default) behaves like a plain synchronous call; with
 * several groups
 *   - md_step_host overlaps group k's D2H copies with group k+1's kernels inside the one synchronous call, and
 *   - md_host_send / md_host_recv pipeline ACROSS steps (the EnvPool-style split: while the caller consumes group A's
 *     observations and picks its next actions, group B is being stepped), which hides the PCIe time behind compute.
 * Results do not depend on the partition (MdConfig.env_base keeps every counter hash on global indices).
 * md_step_host: actions [A,2] host floats or NULL (= already written into the pinned action buffers); output pointers
 * may be NULL, the results then stay in the pinned buffers (md_host_views / md_host_group_views).  Synchronous. */
int md_step_host(md_sim* sim, const float* actions, float* obs, float* reward, float* cost, uint8_t* terminated,
                 uint8_t* truncated, int32_t* info_flags, float* info_f, int autoreset);
int md_reset_host(md_sim* sim, const uint8_t* env_mask, float* obs);
/* (re)partition into n_groups (1..32) host groups; no group may be in flight */
int md_host_groups(md_sim* sim, int n_groups);
int md_host_group_count(const md_sim* sim);
/* out8 = addresses of the group's pinned buffers in the order obs rows, reward, cost, terminated, truncated, info_flags,
 * info_f, actions; range3 = first env, env count, observation rows delivered by the last received step */
int md_host_group_views(md_sim* sim, int group, void** out8, int* range3);
/* group 0's buffers (the whole batch while there is one group) */
int md_host_views(md_sim* sim, void** out8);
/* enqueue one env.step of a group (H2D actions, kernels, D2H results) and return; md_host_recv waits for it */
int md_host_send(md_sim* sim, int group, const float* actions, int autoreset);
int md_host_recv(md_sim* sim, int group);
/* multi-agent handles: compact != 0 -> only the observation rows of seats that produced a transition this step
 * (FL_VALID) are copied to the host, packed in ascending seat order at the start of the group's obs rows; the scalars
 * still come back for every seat (row j belongs to the j-th seat whose info_flags carry FL_VALID).  A roundabout env keeps
 * 41 seats for <= 40 live agents and wrecks / empty seats produce no transition: their 1036-byte rows are not sent. */
int md_host_compact(md_sim* sim, int compact);

/* isolated stages, for parity tests and per-kernel ncu captures */
/* Lidar.perceive (component/sensors/lidar.py:49-73; sensors/distance_detector.py:27-85): frac_dev [A,n_lasers] in
 * [0,1]; hit_dev [A,n_lasers] = hit vehicle slot, slots_per_env + object index, or -1 */
int md_lidar(md_sim* sim, float* frac_dev, int32_t* hit_dev, void* stream);
/* TopDownObservation.observe (obs/top_down_obs.py:98-200, 221-229; obs/top_down_obs_impl.py:19-97, 203-250, 266-428 - the
 * observation of envs/top_down_env.py:7-31 TopDownSingleFrameMetaDriveEnv): img_dev [A, resolution, resolution, 3] float32 RGB
 * in [0, 1], the window of +-max_distance metres around every agent turned so that it looks up: lane lines, the ego GREEN,
 * every other vehicle BLUE; an empty seat's image is black.  Reads the current state, changes nothing. */
int md_topdown(md_sim* sim, float* img_dev, int resolution, float max_distance, void* stream);
/* The per-frame grey channels TopDownMultiChannel stacks (obs/top_down_obs_multi_channel.py:101-146, 148-205, 216-270 - the
 * observation of envs/top_down_env.py:34-48 TopDownMetaDrive): img_dev [A, resolution, resolution, 2] float32 =
 * [road_network: lane lines over the drivable area of the lanes of the agent's route, doubled and clipped as observe() does;
 *  traffic_flow: the other vehicles].  The stacking over time (past frames, past positions) is host logic, as in the reference. */
int md_topdown_channels(md_sim* sim, float* img_dev, int resolution, float max_distance, void* stream);
/* n_sub x BulletWorld.doPhysics(dt,1,dt) for every vehicle with given actuation act3_dev [NV,3] = steering rad,
 * engine force, brake (component/vehicle/base_vehicle.py:447-484; engine/core/engine_core.py:350-352) */
int md_dynamics(md_sim* sim, const float* act3_dev, int n_sub, void* stream);
/* BaseVehicle.after_step for every active vehicle (component/vehicle/base_vehicle.py:234-253, 700-792) */
int md_after_step(md_sim* sim, void* stream);
/* IDMPolicy.act for every active traffic vehicle (policy/idm_policy.py:235-267): out_dev [NV,2] */
int md_idm(md_sim* sim, float* out_actions_dev, void* stream);

/* snapshots: BaseObject.get_state / set_state (base_class/base_object.py:435-452).  `name` is an MdArrays field
 * name ("veh_s", "veh_i", "obj_f", ...); synchronous host <-> device copy of the whole array. */
int md_get_state(md_sim* sim, const char* name, void* host_dst, size_t bytes);
int md_set_state(md_sim* sim, const char* name, const void* host_src, size_t bytes);
/* Scenario resampling at reset.  The reference draws a scenario per reset: BaseEnv.reset(seed=None) ->
 * _reset_global_seed picks current_seed in [start_seed, start_seed + num_scenarios) (envs/base_env.py:502-537, 886-891), the
 * map manager loads that map and the managers respawn their bodies.  Here `bank` is a second handle on the same device
 * holding ONE ENV PER SCENARIO of the library, fully reset; after md_attach_bank a finished env of `sim` (md_step_autoreset)
 * restarts as scenario hash(seed, env, reset pass) % n_bank by copying that scenario's rows from the bank's post-reset snapshot.
 * Both handles must have loaded the same map set (same map ids) and the same slots / objects / agents per env; single-agent,
 * trigger-mode worlds.  The bank must outlive `sim` (or be detached with bank = NULL).  Returns 0 or a negative code. */
int md_attach_bank(md_sim* sim, md_sim* bank, int seed);
/* Contact pairs of the last step: what the reference's contact-added callback sees (engine/core/collision_callback.py:5-42).
 * After md_enable_contacts(sim, 1) every step records, per vehicle slot, the bodies its chassis touched during the sub-steps:
 * 128 bits, bit k < slots_per_env = vehicle slot k of the same env, bit slots_per_env + j = object j of the env.
 * md_get_contacts copies the [NV, 4] uint32 table to the host (synchronous). */
int md_enable_contacts(md_sim* sim, int on);
int md_get_contacts(md_sim* sim, uint32_t* host_dst, size_t bytes);
/* make the current device state the snapshot md_reset restores */
int md_snapshot(md_sim* sim);
/* per-stage device timing of the next max_steps md_step / md_step_autoreset calls: six cudaEvents per call recorded on
 * the launch stream (no synchronisation added).  md_profile_end (after the caller synchronised) fills ms[5*i + k] with
 * the milliseconds of stage k (0 k_pre, 1 k_dyn, 2 k_scan + k_post, 3 fused reset, 4 k_lidar) of recorded step i and returns
 * how many steps. */
int md_profile_begin(md_sim* sim, int max_steps);
int md_profile_end(md_sim* sim, float* ms, int cap);
/* measured FP32-FMA peak of `device` in TFLOP/s (a 2 ms micro-benchmark of independent FFMA chains): the denominator of
 * the secondary, ray-test roofline bench.py reports */
int md_fp32_peak(int device, double* tflops);
/* number of kernels this handle has launched since creation (bench.py's gpu_launches) */
int64_t md_launch_count(const md_sim* sim);

#ifdef __cplusplus
}
#endif
#endif
