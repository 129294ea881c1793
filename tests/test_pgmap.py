"""CPU: the product's own scene generator (pgmap: BIG + PG blocks; pgspawn: ego / traffic / accident scenes) against its
goldens - the scenario libraries exported from the UNMODIFIED reference (oracle/gen_assets.py): the lane / road tables and
block metadata must come out bit for bit, the rosters (vehicle classes and sampled parameters, spawn lanes and longitudes,
routes, checkpoint indices, IDM timers, obstacles) identical, poses equal after the float32 cast the device arrays take.

The default run samples every library (a few seconds); MD_PGMAP_FULL=1 walks all 1000 + 100 + 100 scenarios (~2 min)."""
import json
import os

import numpy as np
import pytest

from metadrive_ped_b200 import pgmap, pgspawn
from metadrive_ped_b200.library import GeneratedLibrary, ScenarioLibrary

FULL = os.environ.get("MD_PGMAP_FULL", "") not in ("", "0")
LIBS = [("pg3_density0.1.npz", 3, dict(traffic_density=0.1, traffic_mode="trigger", accident_prob=0.0), 1000 if FULL else 120),
        ("safe_pg3.npz", 3, dict(traffic_density=0.05, traffic_mode="trigger", accident_prob=0.8), 100 if FULL else 40),
        ("x_respawn_density0.1.npz", "X", dict(traffic_density=0.1, traffic_mode="respawn", accident_prob=0.0), 100 if FULL else 30)]


@pytest.mark.parametrize("name,spec,kw,n", LIBS)
def test_generator_reproduces_the_reference_libraries(name, spec, kw, n):
    lib = ScenarioLibrary(name)
    step = max(1, len(lib) // n)
    seen_blocks = set()
    for i in list(range(0, len(lib), step))[:n]:
        seed = int(lib.seeds[i])
        lane_f, lane_i, road_i, meta, _ = lib._map_args(i)
        g_f, g_i, g_r, g_meta, big = pgmap.generate(seed, spec)
        assert "".join(b["id"] for b in meta["blocks"]) == "".join(b.ID for b in big.blocks), seed
        seen_blocks |= {b.ID for b in big.blocks}
        np.testing.assert_array_equal(lane_f, g_f, err_msg="lane_f of seed %d" % seed)       # float64, bit for bit
        np.testing.assert_array_equal(lane_i, g_i, err_msg="lane_i of seed %d" % seed)
        np.testing.assert_array_equal(road_i, g_r, err_msg="road_i of seed %d" % seed)
        assert meta["nodes"] == g_meta["nodes"]
        for b_ref, b_gen in zip(meta["blocks"], g_meta["blocks"]):
            for key in ("id", "trigger_road", "spawn_lanes", "negative_lanes", "respawn_roads", "sockets"):
                assert b_ref[key] == b_gen[key], (seed, b_ref["id"], key)
        if meta.get("respawn"):
            assert meta["respawn"] == pgspawn.respawn_table(big, seed), seed
        ref = lib.scenario(i, 0)
        got = pgspawn.populate(big, seed, **kw)
        for field in ("veh_static", "routes", "veh_int", "idm"):
            np.testing.assert_array_equal(getattr(ref, field), getattr(got, field), err_msg="%s of seed %d" % (field, seed))
        np.testing.assert_allclose(ref.veh_dyn, got.veh_dyn, rtol=0, atol=1e-12)
        np.testing.assert_array_equal(ref.veh_dyn.astype(np.float32), got.veh_dyn.astype(np.float32))
        assert ref.objects.shape == got.objects.shape, seed
        if len(ref.objects):
            np.testing.assert_allclose(ref.objects, got.objects, rtol=0, atol=1e-12)
    if spec == 3 and n >= 100:
        assert seen_blocks == set("ISCrRXTO"), seen_blocks


@pytest.mark.parametrize("name,spec", [("lib_yY.npz", "yY"), ("lib_SyYC.npz", "SyYC"), ("lib_StollC.npz", "S$C")])
def test_generator_reproduces_maps_with_bottleneck_and_tollgate_blocks(name, spec):
    """Merge ("y"), Split ("Y") and TollGate ("$") blocks inside BIG-generated maps (map="SyYC" ...): 8 scenarios each exported
    from the reference (oracle/gen_assets.py --map).  Lane tables bit for bit; rosters identical - which, for the tollgate map,
    requires the seed every TollGateBuilding draws from the engine's stream when the block is built."""
    lib = ScenarioLibrary(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "libs", name))
    for i in range(len(lib)):
        seed = int(lib.seeds[i])
        lane_f, lane_i, road_i, meta, _ = lib._map_args(i)
        g_f, g_i, g_r, g_meta, big = pgmap.generate(seed, spec)
        assert "".join(b["id"] for b in meta["blocks"]) == "".join(b.ID for b in big.blocks) == "I" + spec
        np.testing.assert_array_equal(lane_f, g_f)
        np.testing.assert_array_equal(lane_i, g_i)
        np.testing.assert_array_equal(road_i, g_r)
        ref = lib.scenario(i, 0)
        got = pgspawn.populate(big, seed, traffic_density=0.1, traffic_mode="trigger", accident_prob=0.0)
        for field in ("veh_static", "routes", "veh_int", "idm"):
            np.testing.assert_array_equal(getattr(ref, field), getattr(got, field), err_msg="%s of seed %d" % (field, seed))
        np.testing.assert_allclose(ref.veh_dyn, got.veh_dyn, rtol=0, atol=1e-12)
    if "$" in spec:   # the booths close the object table: one on every second lane of both toll roads
        lib = GeneratedLibrary(0, 2, map=spec)
        arrays, cfg = lib.build_world([0, 1])
        booths = arrays["obj_f"][arrays["obj_f"][:, 0] == 4.0]
        assert cfg.objs_per_env >= 2 and len(booths) == 4 and (booths[:, 5] == 5.0).all() and (booths[:, 7] == 0.0).all()
        with pytest.raises(NotImplementedError):
            GeneratedLibrary(0, 1, map="SBC")


@pytest.mark.parametrize("spaces,asset", [(8, "ma_parkinglot.npz"), (4, "ma_parkinglot_4.npz"), (12, "ma_parkinglot_12.npz")])
def test_generator_reproduces_the_parking_lot_map(spaces, asset):
    """MAParkingLotMap (envs/marl_envs/marl_parking_lot.py:144-184): first block (one lane, 20 m) -> ParkingLot (parking_space_num
    / 2 spaces a side, pgblock/parking_lot.py) -> T intersection; 106 lanes with the env's 8 spaces; lane tables, node names and the
    env's spawn roads (3 ways in + the spaces) bit for bit against the reference's exports (oracle/gen_assets.py --env ma_parkinglot
    [--parking-spaces N])."""
    import json
    from metadrive_ped_b200.ma import MultiAgentLibrary
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "libs", asset))
    conf, meta = json.loads(str(d["config"])), json.loads(str(d["meta"]))
    lane_f, lane_i, road_i, g_meta, big = pgmap.build_fixed("parkinglot", conf["lane_num"], 3.5, conf["exit_length"], spaces)
    assert [b.ID for b in big.blocks] == ["I", "P", "T"] and g_meta["nodes"] == meta["nodes"]
    np.testing.assert_array_equal(lane_f, d["lane_f"])
    np.testing.assert_array_equal(lane_i, d["lane_i"])
    np.testing.assert_array_equal(road_i, d["road_i"])
    assert len(big.blocks[1].dest_roads) == spaces and len(big.blocks[1].parking_spawn_roads) == spaces
    lib = MultiAgentLibrary("parkinglot", parking_space_num=spaces)
    np.testing.assert_array_equal(lib.spawn_roads, d["spawn_roads"])
    assert lib.max_capacity == 3 + spaces and lib.conf["parking_spaces"] == spaces


def test_generated_world_equals_exported_world():
    """What the device receives: scene.pack over generated scenarios == scene.pack over the exported ones, every array."""
    gen = GeneratedLibrary(0, 16)
    ref = ScenarioLibrary("pg3_density0.1.npz")
    a, ca = gen.build_world(list(range(16)))
    b, cb = ref.build_world(list(range(16)))
    assert ca.slots_per_env == cb.slots_per_env
    for k in a:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)


def test_block_sequences_and_parameters():
    """BASELINE cfg1 (map="S") and other block sequences / counts / densities the shipped libraries do not hold."""
    lane_f, lane_i, road_i, meta, big = pgmap.generate(0, "S")
    assert [b.ID for b in big.blocks] == ["I", "S"] and len(lane_f) == 18          # "S" has 18 lanes
    assert 40.0 <= big.blocks[1].cfg["length"] <= 80.0
    for spec, n_blocks in (("SCO", 4), ("XTrR", 5), (5, 6), (1, 2)):
        _, _, _, meta, big = pgmap.generate(7, spec)
        assert len(big.blocks) == n_blocks
        if isinstance(spec, str):
            assert "".join(b.ID for b in big.blocks[1:]) == spec
    with pytest.raises(AssertionError):
        pgmap.generate(0, "SP")                                                    # a parking lot needs a one-lane road before it
    assert [b.ID for b in pgmap.generate(0, "SP", lane_num=1)[4].blocks] == ["I", "S", "P"]
    # traffic density scales the roster (traffic_manager.py:236-238); the same seed keeps the same map
    _, _, _, _, big = pgmap.generate(11, 3)
    n = [len(pgspawn.populate(pgmap.generate(11, 3)[4], 11, d).veh_static) for d in (0.0, 0.1, 0.3)]
    assert n[0] == 1 and n[1] < n[2]
    # accident scenes: cones come in runs of 12 (3 + 6 + 3, object_manager.py:118-140), break-down scenes carry their car
    kinds = []
    for seed in range(30):
        sc = pgspawn.populate(pgmap.generate(seed, 3)[4], seed, 0.05, "trigger", 0.8)
        kinds += list(sc.objects[:, 0])
        n_break = int(np.sum((sc.veh_int[:, 0] == 2) & (sc.veh_int[:, 1] == 0)))
        assert n_break == int(np.sum(sc.objects[:, 0] == 1)), seed                  # one broken-down car per warning tripod
    assert kinds.count(0.0) % 12 == 0 and kinds.count(0.0) > 0 and kinds.count(1.0) > 0 and kinds.count(2.0) > 0


def test_generated_library_interface():
    g = GeneratedLibrary(5, 4, map="SC", traffic_density=0.2, random_lane_num=True)
    assert len(g) == 4 and g.index_of_seed(7) == 2
    with pytest.raises(KeyError):
        g.index_of_seed(99)
    arrays, cfg = g.build_world([0, 1, 2, 3, 0])
    assert cfg.n_envs == 5 and arrays["env_i"][:, 4].tolist() == [5, 6, 7, 8, 5]
    assert g.max_vehicles() <= cfg.slots_per_env
    meta = json.loads(json.dumps(g._map_args(0)[3]))
    assert [b["id"] for b in meta["blocks"]] == ["I", "S", "C"]


FIXTURES = ["cfg1_S_straight", "cfg1_S_random", "cfg1_S_discrete", "cfg2_pg3_seed3", "cfg2_pg3_seed7", "cfg2_pg3_seed11_dense",
            "cfg2_SCO_nolimit", "cfg2_pg3_seed11_others4", "cfg4_safe_seed2", "cfg4_safe_seed5", "cfg4_safe_seed40_cones",
            "cfg4_safe_seed8_bump", "cfg5_ped_X"]


@pytest.mark.parametrize("tag", FIXTURES)
def test_generator_reproduces_the_golden_episodes_scenes(tag):
    """The reference episodes under tests/golden/ carry their map and reset-time roster: map "S", "SCO", density 0.2 / 0.3,
    SafeMetaDriveEnv scenes, respawn-mode "X" - configurations no shipped library holds."""
    from tests.golden_util import load_golden
    g = load_golden(tag)
    conf = json.loads(str(g["config"]))
    seed = int(g["seed"])
    safe = tag.startswith("cfg4")
    lane_f, lane_i, road_i, meta, big = pgmap.generate(seed, conf.get("map", 3))
    np.testing.assert_array_equal(np.asarray(g["map_lane_f"], np.float64), lane_f)
    np.testing.assert_array_equal(np.asarray(g["map_lane_i"], np.int32), lane_i)
    np.testing.assert_array_equal(np.asarray(g["map_road_i"], np.int32), road_i)
    got = pgspawn.populate(big, seed, conf.get("traffic_density", 0.05 if safe else 0.1), conf.get("traffic_mode", "trigger"),
                           0.8 if safe else 0.0)   # break-down scenes carry their car (the round-2 fixtures do as well)
    n = len(got.veh_static)
    assert n == len(g["init_veh_static"])
    np.testing.assert_array_equal(np.asarray(g["init_veh_static"], np.float32), got.veh_static)
    np.testing.assert_array_equal(np.asarray(g["init_routes"], np.int32), got.routes)
    np.testing.assert_array_equal(np.asarray(g["init_veh_int"], np.int32), got.veh_int)
    np.testing.assert_array_equal(np.asarray(g["init_idm"], np.float32), got.idm)
    np.testing.assert_allclose(np.asarray(g["init_veh_dyn"], np.float64), got.veh_dyn, rtol=0, atol=1e-12)
    objs = np.asarray(g["init_objects"], np.float64).reshape(-1, g["init_objects"].shape[-1] if g["init_objects"].ndim > 1 else 8)
    objs = objs[objs[:, 0] < 3][:, :8]                         # cfg5 appends its (build-defined) pedestrians
    assert objs.shape == got.objects.shape
    if len(objs):
        np.testing.assert_allclose(objs, got.objects, rtol=0, atol=1e-12)


@pytest.mark.skipif(not os.path.isdir("/root/reference/metadrive"), reason="needs the reference checkout (build container only)")
@pytest.mark.parametrize("tag", ["cfg1_S_straight", "cfg4_safe_seed2", "cfg3_ma_intersection_others_navi"])
def test_golden_fixtures_regenerate_identically(tag, tmp_path):
    """The committed recipe reproduces the committed fixtures: oracle/gen_golden.py runs ONE FRESH PROCESS PER TAG (the
    reference leaks map constants through class attributes across env instances, see its main()) with PYTHONHASHSEED=0, multi-agent
    episodes are seeded.  Two small single-agent tags and a multi-agent one (8 agents, respawns) are regenerated here and compared
    array by array."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call([sys.executable, "-m", "oracle.gen_golden", "--only", tag, "--exact", "--out", str(tmp_path)], cwd=root,
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=dict(os.environ, PYTHONHASHSEED="0"))
    new = np.load(os.path.join(str(tmp_path), tag + ".npz"))
    old = np.load(os.path.join(root, "tests", "golden", tag + ".npz"))
    assert sorted(new.files) == sorted(old.files)
    for k in old.files:
        if old[k].dtype.kind in "US":
            assert str(old[k]) == str(new[k]), k
        else:
            np.testing.assert_array_equal(old[k], new[k], err_msg=k)


@pytest.mark.parametrize("asset", ["ma_roundabout.npz", "ma_intersection.npz", "ma_bottleneck.npz", "ma_tollgate.npz",
                                   "ma_bidirection.npz"])
def test_generated_multi_agent_maps_equal_the_exported_assets(asset):
    """MARoundaboutMap / MAIntersectionMap (exit length 60, two lanes, U-turns on the intersection), MABottleneckMap (I -> Merge
    -> Split, 4 lanes into 1) and MATollGateMap (I -> Split -> TollGate -> Merge, 3 lanes into 8) and the spawn roads,
    destination nodes, slot constants and static_default vehicle of the multi-agent envs, generated vs exported."""
    from metadrive_ped_b200.ma import MultiAgentLibrary
    gen, ref = MultiAgentLibrary(asset), MultiAgentLibrary(asset, from_asset=True)
    for k in ("lane_f", "lane_i", "road_i"):
        np.testing.assert_array_equal(getattr(gen.table, k), getattr(ref.table, k), err_msg=k)
    np.testing.assert_array_equal(gen.spawn_roads, ref.spawn_roads)
    np.testing.assert_array_equal(gen.dest_nodes, ref.dest_nodes)
    np.testing.assert_array_equal(gen.veh_static, ref.veh_static)
    assert gen.slots == ref.slots and gen.max_capacity == ref.max_capacity
    for k in ("lane_num", "exit_length", "entrance_length", "respawn_longitude", "respawn_lateral", "max_vehicle_length",
              "max_vehicle_width", "num_agents"):
        assert gen.conf[k] == ref.conf[k], k
    a, ca = gen.build_world(3, 8, seed=5)
    b, cb = ref.build_world(3, 8, seed=5)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)


@pytest.mark.parametrize("env_name,asset", [("MultiAgentRoundaboutEnv", "ma_roundabout_3lanes.npz"),
                                            ("MultiAgentIntersectionEnv", "ma_intersection_3lanes.npz"),
                                            ("MultiAgentBottleneckEnv", "ma_bottleneck_neck30.npz"),
                                            ("MultiAgentTollgateEnv", "ma_tollgate_6lanes.npz")])
def test_multi_agent_envs_honour_map_config(env_name, asset):
    """config["map_config"] of the multi-agent envs (lane_num / exit_length; bottleneck: neck_length ...; tollgate: toll_lane_num /
    toll_length), read the way the reference's MA*Map classes read it: the map and spawn roads the env generates for a NON-default
    map_config against the reference's export of the same config (oracle/gen_assets.py --env ma_* --map-config JSON)."""
    import json
    import metadrive_ped_b200.envs as E
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "libs", asset))
    ref_mc = json.loads(str(d["config"]))["map_config"]
    cls = getattr(E, env_name)
    default_mc = cls.default_config()["map_config"]
    override = {k: v for k, v in ref_mc.items() if default_mc.get(k) != v}
    assert override, "the golden must hold a non-default map"
    assert all(k in default_mc for k in override), "the env declares the reference's map_config keys"
    lib = cls._make_library(E._merge(cls.default_config(), dict(map_config=override)))
    np.testing.assert_array_equal(lib.table.lane_f, d["lane_f"])
    np.testing.assert_array_equal(lib.table.lane_i, d["lane_i"])
    np.testing.assert_array_equal(lib.table.road_i, d["road_i"])
    np.testing.assert_array_equal(lib.spawn_roads, d["spawn_roads"])
    assert lib.table.meta["nodes"] == json.loads(str(d["meta"]))["nodes"]
