"""The bird's-eye observation (TopDownObservation, obs/top_down_obs.py; md_topdown / mdo_topdown).

CPU: what the oracle's restatement draws, checked against the scene itself - the ego's GREEN rectangle sits in the middle with
the ego's footprint, looking up; every other vehicle inside the window is a BLUE rectangle where the mirrored, heading-up view
puts it; lane lines carry the reference's (35, 35, 35); turning the whole world turns nothing in the image.
-m gpu: the CUDA image equals the oracle's bit for bit, and the TopDownSingleFrameMetaDriveEnv surface.
pygame is not on this image: the reference's rasteriser itself is not pinned."""
import numpy as np
import pytest

from tests.golden_util import golden_world, load_golden

GREEN = np.array([50, 200, 0], np.float32) / 255
BLUE = np.array([100, 200, 255], np.float32) / 255


def _oracle_after(tag, steps, replicas=1):
    from oracle.oracle import OracleSim
    g = load_golden(tag)
    arrays, cfg, _ = golden_world(g, replicas=replicas)
    sim = OracleSim(arrays, cfg)
    sim.reset_observe()
    for t in range(steps):
        a = np.tile(np.asarray(g["actions"][t], np.float32).reshape(-1, 2), (replicas, 1))
        sim.step(a)
    return g, cfg, sim


def _to_image(ego_xy, ego_h, xy, res, dist):
    """pixel (row, col) of a world point: up = the ego's heading, right = the ego's LEFT (the canvas' y axis points down)"""
    d = np.asarray(xy) - np.asarray(ego_xy)
    v = d[0] * np.cos(ego_h) + d[1] * np.sin(ego_h)
    u = -d[0] * np.sin(ego_h) + d[1] * np.cos(ego_h)
    px = 2 * dist / res
    return int(np.floor(res / 2 - v / px)), int(np.floor(u / px + res / 2))


def test_oracle_topdown_draws_the_scene(oracle_lib):
    g, cfg, sim = _oracle_after("cfg5_ped_X", 105)   # the ego is inside the X intersection's traffic here
    res, dist = 84, 30.0
    img = sim.topdown(res, dist)
    assert img.shape == (1, res, res, 3) and img.dtype == np.float32 and 0.0 <= img.min() and img.max() <= 1.0
    im = img[0]
    green = np.all(im == GREEN, axis=-1)
    blue = np.all(im == BLUE, axis=-1)
    # the ego: a ~1.85 x 4.5 m rectangle around the centre, longer than wide (it looks up), 0.714 m per pixel
    rows, cols = np.nonzero(green)
    assert green[res // 2, res // 2] and 5 <= rows.max() - rows.min() + 1 <= 8 and 2 <= cols.max() - cols.min() + 1 <= 4
    assert abs(rows.mean() - (res / 2 - 0.5)) < 1.0 and abs(cols.mean() - (res / 2 - 0.5)) < 1.0
    # every other vehicle whose centre lies inside the window is BLUE at the pixel the mirrored, heading-up view puts it
    vs, vi = sim.a["veh_s"], sim.a["veh_i"]
    q = vs[0, 3:7]
    ego_h = np.arctan2(2 * (q[0] * q[3] + q[1] * q[2]), 1 - 2 * (q[2] ** 2 + q[3] ** 2)) + np.pi / 2   # body +Y is the nose
    seen = 0
    for k in range(1, cfg.slots_per_env):
        if not vi[k, 1]:
            continue
        r, c = _to_image(vs[0, 0:2], ego_h, vs[k, 0:2], res, dist)
        if 1 <= r < res - 1 and 1 <= c < res - 1:
            assert blue[r, c], (k, r, c)
            seen += 1
    assert seen >= 3 and blue.sum() >= 10 * seen
    # lane lines: grey pixels up to the reference's LANE_LINE_COLOR, none brighter
    grey = ~green & ~blue
    assert 0.0 < im[grey].max() <= 35 / 255 + 1e-7 and (im[grey][:, 0] == im[grey][:, 1]).all()
    assert (im[grey][:, 0] > 0.02).sum() > 100


def test_oracle_topdown_is_ego_centred(oracle_lib):
    """the same scene turned by 90 degrees about the origin and shifted gives the same image (up to the last bits of the
    rotated coordinates): nothing in it depends on the world frame"""
    from oracle.oracle import OracleSim
    g, cfg, sim = _oracle_after("cfg2_pg3_seed3", 30)
    img0 = sim.topdown(64, 25.0)
    a = {k: v.copy() for k, v in sim.a.items()}
    rot = lambda x, y: (-y + 13.0, x - 7.0)
    a["veh_s"][:, 0], a["veh_s"][:, 1] = rot(sim.a["veh_s"][:, 0], sim.a["veh_s"][:, 1])
    qw, qx, qy, qz = (sim.a["veh_s"][:, 3 + i].copy() for i in range(4))
    s = np.float32(np.sqrt(0.5))   # q_new = (cos 45, 0, 0, sin 45) * q
    a["veh_s"][:, 3], a["veh_s"][:, 4], a["veh_s"][:, 5], a["veh_s"][:, 6] = s * (qw - qz), s * (qx - qy), s * (qy + qx), s * (qz + qw)
    a["line_f"][:, 0], a["line_f"][:, 1] = rot(sim.a["line_f"][:, 0], sim.a["line_f"][:, 1])
    a["line_f"][:, 4], a["line_f"][:, 5] = -sim.a["line_f"][:, 5], sim.a["line_f"][:, 4]
    img1 = OracleSim(a, cfg).topdown(64, 25.0)
    assert np.abs(img1 - img0).max() < 0.02 and (np.all(img1 == GREEN, -1) == np.all(img0 == GREEN, -1)).mean() > 0.999


def test_oracle_topdown_channels(oracle_lib):
    """the per-frame channels of the stacked observation: lines over the drivable area of the ROUTE's lanes (doubled: 2 * 64 / 255
    on the area, 2 * 35 / 255 on a fully covered line pixel), other vehicles in the grey of BLUE; the ego is not drawn"""
    g, cfg, sim = _oracle_after("cfg5_ped_X", 105)
    rgb, ch = sim.topdown(84, 30.0)[0], sim.topdown(84, 30.0, channels=2)[0]
    assert ch.shape == (84, 84, 2)
    blue = np.all(rgb == BLUE, axis=-1)
    np.testing.assert_array_equal(ch[..., 1] > 0, blue)
    assert np.allclose(ch[..., 1][blue], (0.299 * 100 + 0.587 * 200 + 0.114 * 255) / 255, atol=1e-6)
    road = ch[..., 0]
    area = np.isclose(road, 2 * 64 / 255, atol=1e-6)
    assert area[42, 42] and area.sum() > 300, "the ego stands on its route's drivable area"
    assert road.max() <= 2 * 64 / 255 + 1e-6
    # off the area the road channel is twice the RGB image's line grey (where no vehicle hides it)
    off = ~area & ~blue & ~np.all(rgb == GREEN, axis=-1) & (road < 2 * 35 / 255 + 1e-6)
    assert np.abs(road[off] - 2 * rgb[..., 0][off]).max() < 0.14   # equal off the area; on it the area shines through partial lines


def test_topdown_stack_restates_the_reference_stacking():
    """obs/top_down_obs_multi_channel.py:163-182, 229-270, 283-290: channel count, the first observation fills the traffic stack
    and clears the past positions, frames are taken frame_skip steps apart, past positions are dots behind the ego at twice
    the image scale, with the ego's left on the image's right"""
    from metadrive_ped_b200.envs import TopDownStack
    st = TopDownStack(84, 30.0, frame_stack=3, post_stack=5, frame_skip=5)
    assert st.num_stacks == 5 and st.traffic.maxlen == 11 and st.past_pos.maxlen == 21
    assert st.indices(11) == [10, 5, 0] and st.indices(3) == [2] and st.indices(7) == [6, 1]
    road = np.full((84, 84), 0.5, np.float32)
    frames = [np.full((84, 84), 0.01 * (t + 1), np.float32) for t in range(13)]
    o = st.observe(road, frames[0], (0.0, 0.0), 0.0)
    assert o.shape == (84, 84, 5) and o.dtype == np.float32
    assert (o[..., 0] == 0.5).all() and (o[..., 2:] == frames[0][..., None]).all()
    assert o[42, 42, 1] == 1.0 and o[..., 1].sum() == 1.0       # the ego's own position, then the stack is cleared
    assert len(st.past_pos) == 0 and len(st.traffic) == 11
    for t in range(1, 13):   # driving along +x at 1 m per step
        o = st.observe(road, frames[t], (float(t), 0.0), 0.0)
    # traffic now, 5 and 10 steps ago
    assert o[0, 0, 2] == frames[12][0, 0] and o[0, 0, 3] == frames[7][0, 0] and o[0, 0, 4] == frames[2][0, 0]
    # past positions 0, 5 and 10 steps ago: 0, 5, 10 m behind = 0, 14, 28 pixels below the centre (scale 84 / 30 px per m)
    rows, cols = np.nonzero(o[..., 1])
    assert sorted(rows.tolist()) == [42, 56, 70] and set(cols.tolist()) == {42}
    # a past position to the ego's LEFT shows on the image's right
    st2 = TopDownStack(84, 30.0, 3, 5, 5)
    st2.observe(road, frames[0], (0.0, 0.0), np.pi / 2)
    st2.observe(road, frames[0], (0.0, 0.0), np.pi / 2)
    for _ in range(5):
        o2 = st2.observe(road, frames[0], (2.0, 0.0), np.pi / 2)   # heading +y: the start (0, 0) is 2 m to the ego's left
    rows, cols = np.nonzero(o2[..., 1])
    assert (42, 42) in set(zip(rows.tolist(), cols.tolist())) and (42, 42 + int(2 * 84 / 30)) in set(zip(rows.tolist(), cols.tolist()))
    st.reset()
    o = st.observe(road, frames[3], (50.0, 0.0), 0.0)
    assert (o[..., 2:] == frames[3][..., None]).all() and len(st.past_pos) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("tag,steps,res,dist", [("cfg5_ped_X", 105, 84, 30.0), ("cfg4_safe_seed5", 20, 100, 50.0),
                                                ("cfg3_ma_roundabout_respawn", 60, 84, 30.0), ("cfg2_SCO_nolimit", 25, 37, 12.5)])
def test_topdown_image_matches_oracle(tag, steps, res, dist, oracle_lib):
    import torch
    from metadrive_ped_b200.sim import BatchedSim
    from oracle.oracle import OracleSim
    g = load_golden(tag)
    arrays, cfg, _ = golden_world(g, replicas=3)
    sim, orc = BatchedSim(arrays, cfg), OracleSim(arrays, cfg)
    sim.reset()
    orc.reset_observe()
    for t in range(steps):
        a = np.tile(np.asarray(g["actions"][t], np.float32).reshape(-1, 2), (cfg.n_envs, 1))
        sim.step(torch.from_numpy(a).cuda())
        orc.step(a)
    np.testing.assert_array_equal(sim.get_state("veh_s"), orc.a["veh_s"])
    got = sim.topdown(res, dist).cpu().numpy()
    want = orc.topdown(res, dist)
    assert got.shape == want.shape == (cfg.n_envs * cfg.agents_per_env, res, res, 3)
    np.testing.assert_array_equal(got, want)
    assert (want > 0).any()
    got2, want2 = sim.topdown(res, dist, channels=2).cpu().numpy(), orc.topdown(res, dist, channels=2)
    np.testing.assert_array_equal(got2, want2)
    assert (want2[..., 0] > 0.4).any(), "the route's drivable area must show"
    sim.close()


@pytest.mark.gpu
def test_topdown_env_surface():
    """envs/top_down_env.py:7-31: image observations from reset() and step(), float32 in [0, 1] or uint8"""
    from metadrive_ped_b200 import TopDownSingleFrameMetaDriveEnv
    env = TopDownSingleFrameMetaDriveEnv(dict(num_scenarios=10, start_seed=0, traffic_density=0.2))
    try:
        o, info = env.reset(seed=2)
        assert o.shape == (84, 84, 3) and o.dtype == np.float32 and env.observation_space.contains(o)
        assert np.all(o[42, 42] == GREEN)
        for _ in range(20):
            o, r, te, tr, info = env.step([0.0, 1.0])
            assert env.observation_space.contains(o) and np.all(o[42, 42] == GREEN)
    finally:
        env.close()
    from metadrive_ped_b200 import TopDownMetaDrive
    env = TopDownMetaDrive(dict(num_scenarios=10, start_seed=0, traffic_density=0.2))
    try:
        o, info = env.reset(seed=2)
        assert o.shape == (84, 84, 5) and o.dtype == np.float32 and env.observation_space.contains(o)
        assert abs(o[42, 42, 0] - 2 * 64 / 255) < 1e-6 and o[42, 42, 1] == 1.0
        for t in range(12):
            o, r, te, tr, info = env.step([0.0, 1.0])
            assert env.observation_space.contains(o)
        rows, cols = np.nonzero(o[..., 1])
        assert len(rows) == 3 and rows.min() == 42 and rows.max() > 42, "the ego's past positions trail behind it"
    finally:
        env.close()
    env = TopDownSingleFrameMetaDriveEnv(dict(norm_pixel=False, resolution_size=64, distance=20))
    try:
        o, _ = env.reset()
        assert o.shape == (64, 64, 3) and o.dtype == np.uint8 and tuple(o[32, 32]) == (50, 200, 0)
    finally:
        env.close()
