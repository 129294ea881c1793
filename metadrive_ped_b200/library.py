"""Scenario library: maps + reset-time rosters exported once from the reference (oracle/gen_assets.py) and shipped
as compressed arrays, so that reset() on a box without the reference still produces the reference's scenes.

A library entry i corresponds to `env.reset(seed=seeds[i])` of the reference env the library was generated from
(envs/base_env.py:502-537): the PG map (component/algorithm/BIG.py) and what the managers spawned
(manager/traffic_manager.py:211-277, manager/object_manager.py:40-151, manager/agent_manager.py:88-113).
"""
import json
import os
from concurrent.futures import ProcessPoolExecutor

import numpy as np

from . import scene as sc
from .abi import TRAFFIC_MODES, make_config

ASSET_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


def _build_geo(args):
    lane_f, lane_i, road_i, meta, lane_num = args
    return sc.build_map_geometry(sc.MapTable(lane_f, lane_i, road_i, meta, lane_num))


class ScenarioLibrary:
    """Scenarios exported from the reference (an .npz under assets/)."""
    def __init__(self, path):
        if not os.path.isabs(path) and not os.path.exists(path):
            path = os.path.join(ASSET_DIR, path)
        d = np.load(path, allow_pickle=False)
        self.path = path
        self.seeds = d["seeds"]
        self.map_off, self.veh_off, self.obj_off = d["map_off"], d["veh_off"], d["obj_off"]
        self.meta = d["meta"]
        self.env_kind = str(d["env"])
        self.config = json.loads(str(d["config"]))
        self._a = {k: d[k] for k in ("lane_f", "lane_i", "road_i", "veh_static", "veh_dyn", "routes", "veh_int", "idm",
                                      "objects")}
        self._geo = {}

    def __len__(self):
        return len(self.seeds)

    def index_of_seed(self, seed):
        idx = np.nonzero(self.seeds == seed)[0]
        if len(idx) == 0:
            raise KeyError("scenario seed %d is not in library %s" % (seed, os.path.basename(self.path)))
        return int(idx[0])

    def _map_args(self, i):
        lo, nl, ro, nr, lane_num = (int(x) for x in self.map_off[i])
        return (np.asarray(self._a["lane_f"][lo:lo + nl], np.float64), np.asarray(self._a["lane_i"][lo:lo + nl], np.int32),
                np.asarray(self._a["road_i"][ro:ro + nr], np.int32), json.loads(str(self.meta[i])), lane_num)

    def geometries(self, indices, workers=None):
        """MapGeometry for every library index in `indices` (cached; built in parallel)."""
        todo = [i for i in dict.fromkeys(indices) if i not in self._geo]
        if todo:
            workers = workers or min(len(todo), os.cpu_count() or 1)
            if workers > 1 and len(todo) > 4:
                with ProcessPoolExecutor(workers) as ex:
                    for i, g in zip(todo, ex.map(_build_geo, [self._map_args(i) for i in todo], chunksize=8)):
                        self._geo[i] = g
            else:
                for i in todo:
                    self._geo[i] = _build_geo(self._map_args(i))
        return [self._geo[i] for i in indices]

    def scenario(self, i, map_id):
        vo, nv = (int(x) for x in self.veh_off[i])
        oo, no = (int(x) for x in self.obj_off[i])
        a = self._a
        return sc.Scenario(map_id, np.asarray(a["veh_static"][vo:vo + nv], np.float32),
                           np.asarray(a["veh_dyn"][vo:vo + nv], np.float64), np.asarray(a["routes"][vo:vo + nv], np.int32),
                           np.asarray(a["veh_int"][vo:vo + nv], np.int32), np.asarray(a["idm"][vo:vo + nv], np.float32),
                           np.asarray(a["objects"][oo:oo + no], np.float64).reshape(-1, 8), int(self.seeds[i]))

    def max_vehicles(self):
        return int(self.veh_off[:, 1].max())

    def max_objects(self):
        return int(self.obj_off[:, 1].max())

    def build_world(self, indices, slots_per_env=None, objs_per_env=None, num_pedestrians=0, seed=0, map_universe=None,
                    **cfg_kw):
        """(arrays, cfg) for one env per entry of `indices` (library indices, repeats allowed).  `num_pedestrians` > 0
        adds that many crossing pedestrians per env (peds.py, BASELINE config 5), drawn with `seed`.  `map_universe`
        (library indices) fixes the loaded map set and its numbering: two worlds built with the same universe share
        their map ids, which is what a scenario bank needs (sim.attach_bank)."""
        from . import ma, peds
        indices = [int(i) for i in indices]
        uniq = list(dict.fromkeys([int(i) for i in (map_universe or [])] + indices))
        geos = self.geometries(uniq)
        map_id = {i: k for k, i in enumerate(uniq)}
        scen_cache = {i: self.scenario(i, map_id[i]) for i in uniq}
        scenarios = [scen_cache[i] for i in indices]
        if num_pedestrians > 0:
            import dataclasses
            rng = np.random.default_rng(seed)
            with_peds = []
            for s in scenarios:
                rows = peds.place_pedestrians(geos[s.map_id], rng, num_pedestrians)
                old = np.zeros((len(s.objects), 10))
                old[:, :8] = s.objects[:, :8]
                with_peds.append(dataclasses.replace(s, objects=np.concatenate([old, rows])))
            scenarios = with_peds
        S = slots_per_env or max(4, -(-max(len(s.veh_static) for s in scenarios) // 4) * 4)
        O = objs_per_env if objs_per_env is not None else max(len(s.objects) for s in scenarios)
        respawn = self.config.get("traffic_mode", "trigger") in ("respawn", "hybrid")
        kw = {}
        if respawn:
            tape = ma.make_tape(len(indices), seed=seed + 1)
            arrays = sc.pack(geos, scenarios, S, 1, O, ma_tables_tape=tape, traffic_respawn=True)
            kw.update(traffic_mode=TRAFFIC_MODES[self.config["traffic_mode"]], tape_len=ma.TAPE_LEN,
                      ma_places=len(arrays["ma_place_f"]) // len(indices))
        else:
            arrays = sc.pack(geos, scenarios, S, 1, O)
        if self.env_kind == "safe":  # envs/safe_metadrive_env.py:10-19
            kw.update(crash_vehicle_done=0, crash_object_done=0)
        kw.update(cfg_kw)
        cfg = make_config(len(indices), S, 1, O, **kw)
        return arrays, cfg


def _generate_one(args):
    """(map tables, roster) of one seed: pgmap.BIG + pgspawn.populate (a picklable unit for the process pool)"""
    from . import pgmap, pgspawn
    seed, c = args
    lane_num, lane_width = c["lane_num"], c["lane_width"]
    if c["random_lane_width"] or c["random_lane_num"]:   # PGMapManager.add_random_to_map (manager/pg_map_manager.py:76-83)
        rng = pgmap.seeded_rng(seed)
        if c["random_lane_width"]:
            lane_width = rng.rand() * (4.5 - 3.0) + 3.0
        if c["random_lane_num"]:
            lane_num = int(rng.randint(2, 3 + 1))
    lane_f, lane_i, road_i, meta, big = pgmap.generate(seed, c["map"], lane_num, lane_width, c["exit_length"])
    sc_ = pgspawn.populate(big, seed, c["traffic_density"], c["traffic_mode"], c["accident_prob"], lane_num, lane_width,
                           include_breakdown=c["include_breakdown"], random_spawn_lane=c["random_spawn_lane_index"],
                           need_inverse_traffic=c["need_inverse_traffic"], random_traffic=c["random_traffic"],
                           random_agent_model=c["random_agent_model"], agent_model=c["agent_model"])
    meta["respawn"] = pgspawn.respawn_table(big, seed)
    return seed, (lane_f, lane_i, road_i, meta, lane_num), sc_


class GeneratedLibrary(ScenarioLibrary):
    """Scenarios GENERATED on the product side for any supported config: `pgmap` (BIG + PG blocks) builds the map of a
    seed, `pgspawn` populates it (ego, IDM traffic, accident scenes) with the reference's seeded streams.  Same interface
    as the exported libraries, which double as its goldens (tests/test_pgmap.py)."""
    DEFAULTS = dict(map=3, traffic_density=0.1, traffic_mode="trigger", accident_prob=0.0, lane_num=3, lane_width=3.5,
                    exit_length=50, random_lane_width=False, random_lane_num=False, random_spawn_lane_index=True,
                    need_inverse_traffic=False, random_traffic=False, random_agent_model=False, agent_model="default",
                    include_breakdown=True)

    def __init__(self, start_seed=0, num_scenarios=1, env_kind="metadrive", **config):
        c = dict(self.DEFAULTS)
        for k, v in config.items():
            if k not in c:
                raise KeyError(k)
            c[k] = v
        if isinstance(c["map"], str) and any(b in c["map"] for b in "BP"):
            # the generator builds these blocks (pgmap.Bidirection / ParkingLot, bit-identical lane tables) and the step path localises
            # on their shared lanes (the parking-lot env does), but the reference itself cannot drive them inside a BIG map: "SPC" fails
            # an assertion while the map is built, and the Bidirection block of "SBC" puts ONE lane for both directions on the yellow
            # centre line behind three lanes a side, without a Merge in front - no trace to pin.  TollGate blocks ('$') are stepped:
            # their booths are static boxes of the device world (object kind 4).
            raise NotImplementedError("maps with Bidirection ('B') or ParkingLot ('P') blocks can be generated (pgmap.generate) "
                                      "but not stepped: the reference cannot drive them inside a procedurally generated map")
        self.gen = c
        self.path = "<generated>"
        self.seeds = np.arange(start_seed, start_seed + num_scenarios, dtype=np.int32)
        self.env_kind = env_kind
        self.config = dict(map=c["map"], traffic_density=c["traffic_density"], traffic_mode=c["traffic_mode"],
                           accident_prob=c["accident_prob"], start_seed=start_seed, num_scenarios=num_scenarios)
        self._maps, self._rosters, self._geo = {}, {}, {}

    def ensure(self, indices):
        todo = [int(i) for i in dict.fromkeys(indices) if int(i) not in self._maps]
        if not todo:
            return
        jobs = [(int(self.seeds[i]), self.gen) for i in todo]
        workers = min(len(jobs), os.cpu_count() or 1)
        if workers > 1 and len(jobs) > 8:
            with ProcessPoolExecutor(workers) as ex:
                results = list(ex.map(_generate_one, jobs, chunksize=4))
        else:
            results = [_generate_one(j) for j in jobs]
        for i, (_, m, r) in zip(todo, results):
            self._maps[i], self._rosters[i] = m, r

    def _map_args(self, i):
        self.ensure([i])
        lane_f, lane_i, road_i, meta, lane_num = self._maps[int(i)]
        return (lane_f, lane_i, road_i, meta, lane_num)

    def geometries(self, indices, workers=None):
        self.ensure(indices)
        return super().geometries(indices, workers)

    def scenario(self, i, map_id):
        import dataclasses
        self.ensure([i])
        return dataclasses.replace(self._rosters[int(i)], map_id=map_id)

    def max_vehicles(self):
        self.ensure(range(len(self)))
        return max(len(r.veh_static) for r in self._rosters.values())

    def max_objects(self):
        self.ensure(range(len(self)))
        return max(len(r.objects) for r in self._rosters.values())
