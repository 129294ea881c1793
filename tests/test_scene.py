"""CPU: the product's own derivation of static-world geometry (metadrive_ped_b200/scene.py) against the bodies the
reference put into its Bullet worlds (recorded in the golden fixtures)."""
import json

import numpy as np
import pytest

from metadrive_ped_b200 import scene as sc
from tests.golden_util import list_golden, load_golden


@pytest.mark.parametrize("tag", list_golden("cfg2") + list_golden("cfg4"))
def test_line_boxes_match_reference(tag):
    g = load_golden(tag)
    mt = sc.MapTable(np.asarray(g["map_lane_f"], np.float64), np.asarray(g["map_lane_i"], np.int32),
                     np.asarray(g["map_road_i"], np.int32), json.loads(str(g["map_meta"])), int(g["lane_num"]))
    geo = sc.build_map_geometry(mt)
    ref = g["ref_lines"]
    mine = np.stack([geo.line_f[:, 0], geo.line_f[:, 1], np.arctan2(geo.line_f[:, 5], geo.line_f[:, 4]), geo.line_f[:, 2],
                     geo.line_f[:, 3]], 1)
    assert mine.shape == ref.shape

    def order(a):
        return a[np.lexsort((a[:, 4], a[:, 3].round(4), a[:, 1].round(4), a[:, 0].round(4)))]

    a, b = order(mine), order(ref)
    np.testing.assert_allclose(a[:, [0, 1, 3]], b[:, [0, 1, 3]], atol=1e-9)
    np.testing.assert_array_equal(a[:, 4], b[:, 4])
    dang = np.abs(((a[:, 2] - b[:, 2] + np.pi) % (2 * np.pi)) - np.pi)
    assert dang.max() < 1e-9


def test_grid_covers_every_item():
    g = load_golden("cfg2_SCO_nolimit")
    mt = sc.MapTable(np.asarray(g["map_lane_f"], np.float64), np.asarray(g["map_lane_i"], np.int32),
                     np.asarray(g["map_road_i"], np.int32), json.loads(str(g["map_meta"])), 3)
    geo = sc.build_map_geometry(mt)
    n_items = len(geo.line_f) + len(geo.quad_f)
    assert set(geo.grid_items.tolist()) == set(range(n_items))
    nx, ny = geo.grid_dims
    assert len(geo.grid_start) == nx * ny + 1 and geo.grid_start[-1] == len(geo.grid_items)
    # every line centre falls into a cell that lists it
    for it, ln in enumerate(geo.line_f):
        cx = int((ln[0] - geo.grid_origin[0]) // sc.GRID_CELL)
        cy = int((ln[1] - geo.grid_origin[1]) // sc.GRID_CELL)
        c = cy * nx + cx
        assert it in geo.grid_items[geo.grid_start[c]:geo.grid_start[c + 1]]


def test_hulls_are_convex_ccw_and_contain_centreline():
    g = load_golden("cfg2_pg3_seed3")
    mt = sc.MapTable(np.asarray(g["map_lane_f"], np.float64), np.asarray(g["map_lane_i"], np.int32),
                     np.asarray(g["map_road_i"], np.int32), json.loads(str(g["map_meta"])), 3)
    geo = sc.build_map_geometry(mt)
    for l in range(len(geo.lane_f)):
        h = geo.hull_xy[geo.lane_i[l, 4]:geo.lane_i[l, 4] + geo.lane_i[l, 5]]
        e = np.roll(h, -1, 0) - h
        crossz = e[:, 0] * np.roll(e, -1, 0)[:, 1] - e[:, 1] * np.roll(e, -1, 0)[:, 0]
        assert (crossz > 0).all()
        for lon in np.linspace(0, mt.lane_f[l, 2], 7):
            p = sc.lane_position(mt.lane_f[l], lon, 0.0)
            c = e[:, 0] * (p[1] - h[:, 1]) - e[:, 1] * (p[0] - h[:, 0])
            assert (c >= -1e-9).all()


def test_map_universe_gives_both_worlds_the_same_map_ids():
    """md_attach_bank needs the live world and the scenario bank to number their maps identically: build_world's
    `map_universe` fixes the loaded map set, whatever scenarios the envs of a world happen to play."""
    from metadrive_ped_b200.library import ScenarioLibrary
    lib = ScenarioLibrary("pg3_density0.1.npz")
    uni = list(range(12))
    S = max(4, -(-lib.max_vehicles() // 4) * 4)
    live, cfg_l = lib.build_world([7, 3, 3, 11, 0], slots_per_env=S, objs_per_env=0, map_universe=uni)
    bank, cfg_b = lib.build_world(uni, slots_per_env=S, objs_per_env=0, map_universe=uni)
    for k in ("map_desc", "map_descf", "lane_f", "lane_i", "lane_bb", "road_i", "hull_xy", "line_f", "quad_f",
              "grid_start", "grid_items", "lgrid_start", "lgrid_items"):
        np.testing.assert_array_equal(live[k], bank[k], err_msg=k)
    assert cfg_l.slots_per_env == cfg_b.slots_per_env and cfg_b.n_envs == 12
    # env e of the live world plays scenario idx[e]: its rows are the bank's rows of that scenario (env_i[:, 0] = map id)
    for e, scn in enumerate([7, 3, 3, 11, 0]):
        assert live["env_i"][e, 0] == bank["env_i"][scn, 0] == scn
        np.testing.assert_array_equal(live["veh_p"][e * S:(e + 1) * S], bank["veh_p"][scn * S:(scn + 1) * S])
        np.testing.assert_array_equal(live["veh_route"][e * S:(e + 1) * S], bank["veh_route"][scn * S:(scn + 1) * S])
