/* md_layout.h — array layouts (column indices) of the batched MetaDrive step.
 *
 * DATA LAYOUT ONLY: no algorithm lives here.  The same flat arrays are consumed by the CUDA library
 * (metadrive_ped_b200/csrc), by the CPU oracle (oracle/md_oracle.c) and are produced by the Python
 * scene builder (metadrive_ped_b200/scene.py), so all three agree on what a column means.
 *
 * All float arrays are float32 row-major, all int arrays int32 row-major.
 * "NV" = n_envs * slots_per_env vehicle slots (agents first in every env), "NO" = n_envs * objs_per_env.
 */
#ifndef MD_LAYOUT_H
#define MD_LAYOUT_H

/* ---- map tables (one set per distinct map, concatenated; MAPD_* gives offsets) ---------------- */
/* lane_f [Ltot, 16]  (reference: component/lane/straight_lane.py:12-47, circular_lane.py:12-51) */
#define LANE_F 16
#define LF_TYPE 0    /* 0 straight, 1 circular */
#define LF_WIDTH 1
#define LF_LENGTH 2
#define LF_P0 3      /* straight: sx sy ex ey dirx diry heading ; circular: cx cy radius start_phase end_phase direction angle */
#define LF_SX 10     /* lane.start */
#define LF_SY 11
#define LF_EX 12     /* lane.end */
#define LF_EY 13
#define LF_HULL_LONG 14 /* straight lanes: last polygon sample longitude (hull = [0, this] x [-w/2, w/2]);
                           circular lanes: radius of the circle inscribed in the outer chord polygon, minus 1 cm */
#define LF_HULL_NOTHER 15 /* number of leading hull edges that are not outer-arc chords (= all edges for straight lanes);
                             acceleration hints only - the CPU oracle tests every edge */
/* lane_i [Ltot, 8] */
#define LANE_I 8
#define LI_ROAD 0    /* road id local to the map */
#define LI_IDX 1     /* lane index inside the road, 0 = left-most */
#define LI_FROM 2    /* node ids local to the map */
#define LI_TO 3
#define LI_HULL_OFF 4 /* offset into hull_xy (local to the map), convex CCW polygon */
#define LI_HULL_N 5
#define LI_LINE_L 6  /* line type of the left / right border (0 none 1 broken 2 continuous 3 side 4 guardrail) */
#define LI_LINE_R 7
/* lane_bb [Ltot, 4]: hull AABB xmin ymin xmax ymax */
/* road_i [Rtot, 6] */
#define ROAD_I 6
#define RI_FROM 0
#define RI_TO 1
#define RI_FIRST 2   /* first lane id (local to the map) */
#define RI_N 3
#define RI_NEG 4     /* Road.is_negative_road() */
#define RI_BLOCK 5   /* ord(block id char) */
/* line_f [Stot, 8]: static-world lane-line boxes (component/block/base_block.py:468-519); two float4 per row, the
 * first one holds everything a bounding-circle test needs */
#define LINE_F 8
#define LN_CX 0
#define LN_CY 1
#define LN_HALF 2    /* half length; half width is 0.0375 */
#define LN_KIND 3    /* 0 white solid, 1 yellow solid, 2 white broken, 3 yellow broken */
#define LN_UX 4      /* unit direction */
#define LN_UY 5
/* quad_f [Qtot, 8]: sidewalk strip quads, 4 corners CCW (component/pgblock/pg_block.py:294-332) */
#define QUAD_F 8
/* map_desc [M, 16] int32 */
#define MAPD 16
#define MD_LANE_OFF 0
#define MD_N_LANES 1
#define MD_ROAD_OFF 2
#define MD_N_ROADS 3
#define MD_HULL_OFF 4
#define MD_LINE_OFF 5
#define MD_N_LINES 6
#define MD_QUAD_OFF 7
#define MD_N_QUADS 8
#define MD_GRID_OFF 9   /* offset into grid_start (cells+1 entries per map) */
#define MD_GRID_NX 10
#define MD_GRID_NY 11
#define MD_ITEM_OFF 12  /* offset into grid_items */
#define MD_MAX_LANE_NUM 13 /* lane_num of the map config (navigation normalisation) */
#define MD_LGRID_OFF 14    /* offset into lgrid_start: same cells as the static grid, items = lane ids whose hull AABB
                              touches the cell (broad phase of the lane localisation) */
#define MD_LITEM_OFF 15
/* map_descf [M, 4] float: grid origin x, y, cell size, spare */
#define MAPDF 4

/* ---- per-vehicle-slot arrays ------------------------------------------------------------------ */
/* veh_p [NV, 16]: static parameters (component/vehicle/vehicle_type.py, component/pg_space.py:226-272) */
#define VEH_P 16
#define VP_TYPE 0
#define VP_LENGTH 1
#define VP_WIDTH 2
#define VP_HEIGHT 3
#define VP_MASS 4
#define VP_TIRE_R 5
#define VP_LATERAL 6
#define VP_FRONT_WB 7
#define VP_REAR_WB 8
#define VP_CHASSIS_AXIS 9
#define VP_ENGINE 10
#define VP_BRAKE 11
#define VP_MAX_STEER 12   /* degrees */
#define VP_FRICTION 13
#define VP_MAX_SPEED 14   /* km/h */
#define VP_REVERSE 15
/* veh_s [NV, 16]: rigid-body state */
#define VEH_S 16
#define VS_POS 0     /* 3 */
#define VS_QUAT 3    /* 4: w x y z */
#define VS_VEL 7     /* 3 */
#define VS_ANGVEL 10 /* 3 */
#define VS_STEER 13  /* last applied steering action in [-1,1] */
#define VS_THROTTLE 14
#define VS_SPARE 15
/* veh_c [NV, 16]: per-step latches + episode accumulators (component/vehicle/base_vehicle.py:211-232) */
#define VEH_C 16
#define VC_LAST_X 0
#define VC_LAST_Y 1
#define VC_LAST_HX 2     /* last_heading_dir */
#define VC_LAST_HY 3
#define VC_PREV_A0 4     /* last_current_action[0] */
#define VC_PREV_A1 5
#define VC_CUR_A0 6      /* last_current_action[1] */
#define VC_CUR_A1 7
#define VC_DIST_L 8
#define VC_DIST_R 9
#define VC_ENERGY 10
#define VC_EP_REWARD 11
#define VC_STEP_ENERGY 12
#define VC_TOTAL_COST 13
/* MultiAgentTollgateEnv only (cfg.toll_env; envs/marl_envs/marl_tollgate.py:39-105): small integers kept as floats.
 *   VC_TOLL_A = 16 * in_toll_time + 8 * has_exit + 4 * violation + last_block, where in_toll_time is
 *               TollGateObservation's counter (:82-95), last_block StayTimeManager.last_block reduced to what its rules read
 *               (0 not recorded yet, 1 the toll block, 2 any other block), has_exit / violation what `exit_time` means
 *               to done_function (:261-266): both times recorded and exit - entry < min_pass_steps
 *   VC_TOLL_ENTRY = entry_time + 1 (0 = none) */
#define VC_TOLL_A 14
#define VC_PARK 14       /* parking-lot env (never a toll env): 1 + the parking space this agent is heading for, 0 = none */
#define VC_TOLL_ENTRY 15
/* veh_i [NV, 16] */
#define VEH_I 16
#define VI_KIND 0        /* 0 empty, 1 agent, 2 traffic */
#define VI_ALIVE 1       /* body is in the world (visible to lidar / collisions) */
#define VI_ACTIVE 2      /* acts and is localised this step */
#define VI_TRIGGER 3     /* block index whose trigger activates this traffic vehicle, -1 = always */
#define VI_LANE 4        /* current lane id (local to the map), -1 unknown */
#define VI_CKPT0 5
#define VI_CKPT1 6
#define VI_ROUTE_LEN 7
#define VI_FLAGS 8
#define VI_ROUTING_LANE 9 /* IDM routing_target_lane, -1 = None */
#define VI_STATIC 10
#define VI_EP_LEN 11
#define VI_DONE 12       /* agent finished (auto-reset pending) */
#define VI_SPAWN_LANE 13
#define VI_DYING 14      /* multi-agent: steps left as a static wreck before removal (agent_manager.py:260-263) */
#define VI_NEW 15        /* multi-agent: spawned during this step (agent_manager.py:136-154) */
/* veh_route [NV, 24]: checkpoint node ids, -1 padded */
#define ROUTE_MAX 24
/* veh_idm [NV, 8] (policy/idm_policy.py:224-233) */
#define VEH_IDM 8
#define VD_TIMER 0
#define VD_TARGET_SPEED 1
#define VD_H_PERR 2
#define VD_H_IERR 3
#define VD_L_PERR 4
#define VD_L_IERR 5
#define VD_RNG 6         /* counter of the build-defined hash RNG that replaces np_random.randint */
/* veh_navi [NV, 10]: navigation info (node_network_navigation.py:243-292) */
#define NAVI_DIM 10

/* flag bits in VI_FLAGS (component/vehicle/base_vehicle.py:36-62) */
#define FL_CRASH_VEHICLE 0x001
#define FL_CRASH_OBJECT 0x002
#define FL_CRASH_BUILDING 0x004
#define FL_CRASH_HUMAN 0x008
#define FL_CRASH_SIDEWALK 0x010
#define FL_ON_WHITE 0x020
#define FL_ON_YELLOW 0x040
#define FL_ON_BROKEN 0x080
#define FL_ON_LANE 0x100
#define FL_OUT_OF_ROUTE 0x200
/* extra bits only in the info word returned by md_step */
#define FL_OUT_OF_ROAD 0x400
#define FL_ARRIVE 0x800
#define FL_MAX_STEP 0x1000
#define FL_VALID 0x2000    /* multi-agent: this agent seat produced a transition this step */
#define FL_NEWBORN 0x4000  /* multi-agent: the seat was (re)spawned this step: reward 0, first observation */

/* ---- per-object arrays: obj_f [NO, 12] --------------------------------------------------------- */
#define OBJ_F 12
#define OB_KIND 0    /* -1 empty, 0 cone, 1 warning, 2 barrier, 3 pedestrian, 4 building (TollGateBuilding: a static box like
                      * the barrier, crash_building instead of crash_object, no COST_ONCE latch; buildings/tollgate_building.py) */
#define OB_IS_BOX(kind) ((kind) == 2.0f || (kind) == 4.0f)
#define OB_X 1
#define OB_Y 2
#define OB_HEADING 3
#define OB_A 4       /* radius (cylinders) or half length along heading (barrier) */
#define OB_B 5       /* half width (barrier) */
#define OB_HEIGHT 6
#define OB_ZC 7      /* z of the shape centre */
#define OB_LANE 8    /* lane id as float, -1 none */
#define OB_CRASHED 9 /* COST_ONCE latch (static_object/traffic_object.py:27) */
#define OB_VX 10     /* pedestrians: planar velocity */
#define OB_VY 11

/* ---- per-env ints: env_i [E, 8] ------------------------------------------------------------------ */
#define ENV_I 8
#define EI_MAP 0
#define EI_NEXT_TRIGGER 1  /* index of the next block whose vehicles are still waiting; 0 = none left */
#define EI_STEP 2          /* engine.episode_step */
#define EI_N_BLOCKS 3
#define EI_SEED 4
#define EI_RNG 5           /* number of respawn events that consumed a row of the env's random tape (env_tape) */
#define EI_N_PLACES 6      /* respawn / hybrid traffic mode: number of respawn lanes of this env (rows of ma_place_f in use) */
#define TAPE_W 4
/* env_trigger [E, 8]: trigger road id (local to map) per block index */
#define TRIGGER_MAX 8

/* ---- observation row (obs/state_obs.py:64-151, 185-232), N = 240, num_others = 0 ---------------- */
/* ego block = [side: 2 distances | n_side detector rays] + 6 + [lane: 1 lateral offset | n_lane detector rays]
 * (obs/state_obs.py:77-98, 129-149; sensors/distance_detector.py:194-209) */
#define OBS_SIDE(cfg) ((cfg).n_side_lasers > 0 ? (cfg).n_side_lasers : 2)
#define OBS_LANE(cfg) ((cfg).n_lane_lasers > 0 ? (cfg).n_lane_lasers : 1)
#define OBS_EGO(cfg) (OBS_SIDE(cfg) + 6 + OBS_LANE(cfg))
#define OBS_NAVI 10
/* TollGateStateObservation (marl_tollgate.py:65-76) drops the navigation block; TollGateObservation (:79-105) appends
 * [in the toll block, stayed longer than min_pass_steps] after the lidar floats */
#define OBS_STATE(cfg) (OBS_EGO(cfg) + ((cfg).toll_env ? 0 : OBS_NAVI))
#define OBS_TOLL(cfg) ((cfg).toll_env ? 2 : 0)
#define MAX_DET_LASERS 128
#define DET_HEIGHT 0.2f      /* DistanceDetector.DEFAULT_HEIGHT (sensors/distance_detector.py:92) */
#define OBS_OTHER_W(cfg) ((cfg).add_others_navi ? 8 : 4)                  /* floats per neighbour (component/sensors/lidar.py:93-138) */
#define OBS_OTHERS(cfg) (OBS_OTHER_W(cfg) * (cfg).num_others)
#define OBS_DIM(cfg) (OBS_STATE(cfg) + OBS_OTHERS(cfg) + (cfg).n_lasers + OBS_TOLL(cfg))

/* ---- configuration passed by value through the C ABI ------------------------------------------- */
typedef struct MdConfig {
    int n_envs, slots_per_env, agents_per_env, objs_per_env;
    int n_lasers, horizon, decision_repeat, traffic_mode; /* 0 trigger, 1 respawn, 2 hybrid */
    float dt, lidar_dist;
    float success_reward, out_of_road_penalty, crash_vehicle_penalty, crash_object_penalty;
    float driving_reward, speed_reward;
    float crash_vehicle_cost, crash_object_cost, out_of_road_cost;
    int use_lateral_reward, out_of_route_done, on_continuous_line_done;   /* on_continuous_line_done: 1 = yellow / white solid line or sidewalk ends the episode;
                                                                            2 = white solid line or sidewalk only (MultiAgentBottleneckEnv with cross_yellow_line_done=False);
                                                                            3 = sidewalk or yellow solid line, 4 = sidewalk only, and leaving the lanes does NOT count
                                                                            (MultiAgentTollgateEnv._is_out_of_road, marl_tollgate.py:239-245);
                                                                            5 = off the lanes, yellow solid line or sidewalk (MultiAgentParkingLotEnv, marl_parking_lot.py:252-256) */
    int crash_vehicle_done, crash_object_done, crash_human_done, truncate_as_terminate;
    int enable_idm_lane_change, is_multi_agent, delay_done, allow_respawn;
    /* multi-agent respawn tables (manager/spawn_manager.py:117-217): safe places per env, destinations, spawn roads */
    int ma_places, ma_dests, ma_roads, tape_len;
    /* MultiAgentMetaDrive.done_function overrides (envs/marl_envs/multi_agent_metadrive.py:114-128) */
    int ma_crash_done, ma_out_of_road_done;
    int num_others; /* lidar.num_others: the k nearest vehicles, 4 floats each (8 with add_others_navi), between the state and the lidar floats */
    /* side_detector / lane_line_detector (vehicle_config; 0 lasers = off, the reference's default) */
    int n_side_lasers, n_lane_lasers;
    float side_dist, lane_dist;
    /* EnvInputPolicy.convert_to_continuous_action (policy/env_input_policy.py:40-48): 0 = continuous [steer, throttle];
     * 1 = Discrete(steering_dim * throttle_dim), the index travels in actions[:, 0]; 2 = MultiDiscrete([sd, td]) */
    int discrete_action, discrete_steering_dim, discrete_throttle_dim;
    /* LidarStateObservation._add_noise_to_cloud_points (obs/state_obs.py:236-244): clip(x + N(0, sigma), 0, 1), then a
     * ray is zeroed with probability dropout_prob.  The reference draws from numpy's unseeded global generator; here the
     * draws are a counter hash of (noise_seed, observation pass, agent, ray). */
    float lidar_gaussian_noise, lidar_dropout_prob;
    int noise_seed;
    /* first env of a launch's env range within the handle.  Callers pass 0; the library sets it when it runs the step
     * over a sub-range of the envs (the host-buffer groups of md_host_groups), so that the counter hashes (scenario
     * draw, random tape laps, IDM randint, lidar noise) see the same global env / slot / agent index as a whole-batch
     * launch and the results do not depend on how the batch was split. */
    int env_base;
    /* MultiAgentBottleneckEnv / MultiAgentTollgateEnv.reward_function (envs/marl_envs/marl_bottleneck.py:89-127) drop the
     * `positive_road` sign MetaDriveEnv.reward_function applies off the reference lanes (envs/metadrive_env.py:249-266) */
    int ignore_road_sign;
    /* MultiAgentTollgateEnv (envs/marl_envs/marl_tollgate.py): toll_env = 1 switches on its observation (no navigation block,
     * two toll floats at the end), its reward (inside the toll block: -overspeed_penalty * speed / max_speed when faster than the
     * lanes' limit, no speed reward), its done_function (crash_done reads crash_vehicle only; leaving the toll block less than
     * min_pass_steps after entering it ends the episode as out_of_road) and the stay-time bookkeeping behind it */
    int toll_env, min_pass_steps;
    float overspeed_penalty;
    /* lidar.add_others_navi (component/sensors/lidar.py:120-129): every neighbour of the num_others block carries 4 more floats,
     * ITS two navigation checkpoints (BaseNavigation.get_checkpoints, base_navigation.py:145-152) in the observer's frame */
    int add_others_navi;
    /* MultiAgentParkingLotEnv (envs/marl_envs/marl_parking_lot.py:47-142): parking_spaces > 0 switches on ParkingLotSpawnManager's
     * respawn rules.  The spawn roads 0 .. parking_in_roads-1 lead INTO the lot: an agent born there is sent to a parking space no
     * active agent is heading for (destination d of its road = space d; VC_PARK of the agent = d + 1), and such a place is no
     * respawn place while no space is free; the other spawn roads are the spaces themselves, whose agents are sent to the far end of
     * one of the roads into the lot (destinations 0 .. parking_in_roads-1 of their road) */
    int parking_spaces, parking_in_roads;
} MdConfig;

/* ---- all arrays of one simulation, as plain pointers (host for the oracle, device for the library) */
typedef struct MdArrays {
    /* maps */
    const int* map_desc;      /* [M, MAPD] */
    const float* map_descf;   /* [M, MAPDF] */
    const float* lane_f;      /* [Ltot, LANE_F] */
    const int* lane_i;        /* [Ltot, LANE_I] */
    const float* lane_bb;     /* [Ltot, 4] */
    const int* road_i;        /* [Rtot, ROAD_I] */
    const float* hull_xy;     /* [Htot, 2] */
    const float* line_f;      /* [Stot, LINE_F] */
    const float* quad_f;      /* [Qtot, QUAD_F] */
    const int* grid_start;    /* per map: nx*ny+1 */
    const int* grid_items;    /* item < n_lines: line id ; else quad id + n_lines (ids local to the map) */
    /* scenario / state */
    int* env_i;               /* [E, ENV_I] */
    const int* env_trigger;   /* [E, TRIGGER_MAX] */
    const float* veh_p;       /* [NV, VEH_P] */
    float* veh_s;             /* [NV, VEH_S] */
    float* veh_c;             /* [NV, VEH_C] */
    int* veh_i;               /* [NV, VEH_I] */
    int* veh_route;           /* [NV, ROUTE_MAX] (rewritten only by a multi-agent respawn) */
    float* veh_idm;           /* [NV, VEH_IDM] */
    float* veh_navi;          /* [NV, NAVI_DIM] */
    float* obj_f;             /* [NO, OBJ_F] */
    /* derived acceleration tables (the CPU oracle ignores them and scans instead) */
    const int* lgrid_start;   /* per map: nx*ny+1 */
    const int* lgrid_items;   /* lane ids local to the map */
    int* veh_rroad;           /* [NV, ROUTE_MAX]: road id of route segment k = (route[k] -> route[k+1]), -1 padded */
    /* multi-agent respawn tables, per env */
    const float* ma_place_f;  /* [E*ma_places, 8]: x, y, quat w, quat z (yaw only), lane id, heading cos, sin, spawn-road index */
    const int* ma_route;      /* [E*ma_roads*ma_dests, ROUTE_MAX]: checkpoints from spawn road r to destination d */
    const int* ma_rroad;      /* [E*ma_roads*ma_dests, ROUTE_MAX] */
    /* [E*tape_len, TAPE_W] pre-drawn 32-bit words, one row per respawn event of an env, consumed in order.  The
     * reference draws these from numpy generators at run time (multi_agent_metadrive.py:199,
     * marl_inout_roundabout.py:138-143; traffic_manager.py:113-121, idm_policy.py:229); the host fills the tape from
     * its own generator (tests: from the reference trace).
     *   multi-agent respawn : w0 = place draw (mod clear places), w1 = destination draw (mod destinations)
     *   traffic respawn     : w0 = respawn-lane draw (mod lanes), w1 = float bits of rand() in [0,1) for the longitude,
     *                         w2 = overtake-timer draw (mod LANE_CHANGE_FREQ) */
    const int* env_tape;
} MdArrays;

#endif
