"""Small exercise of every kernel (ragged sizes, auto-reset, both observation modes, dynamic mode, obstacles,
filter) for compute-sanitizer runs:  compute-sanitizer --tool memcheck|racecheck python tools/sanitize_smoke.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from pioneer_b200 import BatchConfig, BatchedPioneerEnv, SimulationConfig, demo_obstacles
from pioneer_b200.obs_filter import MeanStdObsFilter


def run(n, steps, **kw):
    sim = SimulationConfig(gravity=kw.pop("gravity", 0.0))
    env = BatchedPioneerEnv(n, seed=1, simulation_config=sim, batch_config=BatchConfig(max_episode_steps=3, **kw))
    act = (torch.rand((n, 6), device="cuda") - 0.5) * 40
    for _ in range(steps):
        obs, rew, flg = env.step_tensor(act)
    # a rollout fragment in one launch (pnr_step_many), rows padded to a multiple of 4; RLlib's reset_at for finished rows
    T, n_pad = 5, (n + 3) // 4 * 4
    acts = (torch.rand((T, n, 6), device="cuda") - 0.5) * 40
    obs_t = torch.zeros((T, n_pad, 137), device="cuda")[:, :n]
    rew_t = torch.zeros((T, n), device="cuda")
    flg_t = torch.zeros((T, n), dtype=torch.uint8, device="cuda")
    env.step_many(acts, obs_t, rew_t, flg_t)
    if env.batch_config.obs_mode == "terminal":
        env.observe_done(flg_t[-1], obs_t[-1], torch.zeros((n_pad, 137), device="cuda")[:n])
    if env.batch_config.random_box:
        env.set_boxes(env.boxes())
    host = torch.empty((n, 6), dtype=torch.float32).pin_memory()
    env.step_host_begin(host, compact=True); env.step_host_begin(host, compact=False)
    env.step_host_end(); env.step_host_end()
    env.observe(indices=[0, n - 1])
    env.reset(indices=[n // 2])
    flt = MeanStdObsFilter(env)
    flt(obs.clone())
    flt.sync()
    torch.cuda.synchronize()
    st = env.episode_stats()
    env.close()
    return st["episodes"]


def exchange():
    """pnr_iteration_sync with two handles on one GPU playing two ranks (peer windows by pointer)."""
    import ctypes as C
    from pioneer_b200 import _cabi as c
    envs = [BatchedPioneerEnv(300, seed=2, env_id_base=300 * r, batch_config=BatchConfig(max_episode_steps=3)) for r in range(2)]
    flts = [MeanStdObsFilter(e) for e in envs]
    lib = envs[0]._lib
    ptrs = (C.c_void_p * 2)()
    for r, e in enumerate(envs):
        p = C.c_void_p()
        c.check(lib.pnr_sync_window_ptr(e._h, C.byref(p)))
        ptrs[r] = p.value
    for r, e in enumerate(envs):
        c.check(lib.pnr_sync_window_connect_ptrs(e._h, ptrs, 2, r))
    outs = [torch.zeros(8 + c.PNR_FILTER_DELTA_LEN, dtype=torch.float64, device="cuda") for _ in range(2)]
    streams = [torch.cuda.Stream() for _ in range(2)]
    for it in range(3):
        for e, f in zip(envs, flts):
            for _ in range(4):
                obs, _, _ = e.step_tensor(torch.zeros((300, 6), device="cuda"))
                f.push(obs)
        torch.cuda.synchronize()
        for r in range(2):
            with torch.cuda.stream(streams[r]):
                c.check(lib.pnr_iteration_sync(envs[r]._h, it % 2, 1, outs[r].data_ptr(), 20000, streams[r].cuda_stream))
        torch.cuda.synchronize()
        assert torch.equal(outs[0], outs[1])
    for e in envs:
        e.close()
    return int(outs[0][0].item())


if __name__ == "__main__":
    total = exchange()
    for n in (1, 33, 1000, 5000):
        total += run(n, 7)
        total += run(n, 7, obs_mode="autoreset", arith="legacy64")
    total += run(777, 4, mode="dynamic", kp=100.0, kd=10.0, torque_scale=100.0, gravity=9.81)
    total += run(777, 4, obstacles=demo_obstacles(), contact_penalty=0.5, random_box=True)
    total += run(130, 2, mode="dynamic", obs_mode="autoreset", obstacles=demo_obstacles(), contact_penalty=0.5, random_box=True,
                 kp=100.0, kd=10.0, torque_scale=100.0, gravity=9.81)
    total += run(97, 2, mode="dynamic", stepping="bullet", motor_max_force=5e3, gravity=9.81)
    print("sanitize smoke ok, episodes", total)
