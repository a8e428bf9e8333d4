"""Observation normaliser (N2): pnr_filter_apply / pnr_filter_sync against the CPU restatement of RLlib's
MeanStdFilter (oracle/filter_oracle.py; parity unpinned vs RLlib, which is not installable here)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle.filter_oracle import BatchSyncFilter, RunningStat

pytestmark = pytest.mark.gpu


def test_statistics_equal_sequential_welford_and_normalisation_matches():
    from pioneer_b200 import BatchConfig, BatchedPioneerEnv
    from pioneer_b200.obs_filter import MeanStdObsFilter
    n = 1000                                                    # ragged: 31 full tiles + one of 8 rows
    env = BatchedPioneerEnv(n, seed=3, batch_config=BatchConfig(max_episode_steps=50))
    flt = MeanStdObsFilter(env)
    ora = BatchSyncFilter(137)
    g = torch.Generator(device="cuda").manual_seed(0)
    a_max = torch.as_tensor(env.a_max).cuda()
    for it in range(4):
        for t in range(5):
            act = (torch.rand((n, 6), device="cuda", generator=g) * 2 - 1) * a_max
            obs, _, _ = env.step_tensor(act)
            raw = obs.clone()
            want = ora(raw.cpu().numpy())
            got = flt(obs)                                      # in place
            assert got.data_ptr() == obs.data_ptr()
            np.testing.assert_allclose(got.cpu().numpy(), want, rtol=0, atol=2e-5)
        flt.sync(); ora.sync()
        assert flt.n == ora.rs.n == n * 5 * (it + 1)
        np.testing.assert_allclose(flt.mean, ora.rs.M, rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(flt.var, ora.rs.var, rtol=1e-5, atol=1e-9)
    # constant columns (r_lo, r_hi, ...) have zero variance: (x - mean) / (0 + 1e-8) must stay finite and ~0
    assert torch.isfinite(got).all() and float(got[:, 18:54].abs().max()) < 1e-2
    assert float(got.abs().max()) <= 10.0
    env.close()


def test_push_only_out_of_place_and_set_stats():
    from pioneer_b200 import BatchedPioneerEnv
    from pioneer_b200.obs_filter import MeanStdObsFilter
    env = BatchedPioneerEnv(256, seed=1)
    flt = MeanStdObsFilter(env, clip=5.0)
    obs = env.reset().clone()
    keep = obs.clone()
    flt.push(obs)
    assert torch.equal(obs, keep)                               # statistics only
    flt.sync()
    rs = RunningStat(137)
    for row in keep.cpu().numpy():
        rs.push(row)
    np.testing.assert_allclose(flt.mean, rs.M, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(flt.var, rs.var, rtol=1e-5, atol=1e-9)
    out = torch.empty_like(obs)
    flt(obs, update=False, out=out)
    assert torch.equal(obs, keep) and flt.n == 256
    want = np.clip((keep.cpu().numpy() - rs.M.astype(np.float32)) * (1 / (rs.std + 1e-8)).astype(np.float32), -5, 5)
    np.testing.assert_allclose(out.cpu().numpy(), want, atol=2e-5)
    flt.set_stats(10.0, np.full(137, 2.0), np.full(137, 4.0))
    y = flt(torch.full((32, 137), 3.0, device="cuda"), update=False)
    assert torch.allclose(y, torch.full_like(y, 0.5))
    env.close()


def test_filter_large_batch_matches_torch():
    from pioneer_b200 import BatchedPioneerEnv
    from pioneer_b200.obs_filter import MeanStdObsFilter
    n = 65536
    env = BatchedPioneerEnv(n, seed=5)
    flt = MeanStdObsFilter(env)
    obs = env.step_tensor(torch.rand((n, 6), device="cuda") * 40 - 20)[0].clone()
    flt.push(obs); flt.sync()
    ref = obs.double()
    np.testing.assert_allclose(flt.mean, ref.mean(0).cpu().numpy(), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(flt.var, ref.var(0, unbiased=True).cpu().numpy(), rtol=1e-5, atol=1e-9)
    y = flt(obs.clone(), update=False)
    want = ((ref - ref.mean(0)) / (ref.std(0, unbiased=True) + 1e-8)).clamp(-10, 10)
    moving = ref.std(0, unbiased=True) > 1e-3
    assert torch.allclose(y[:, moving].double(), want[:, moving], atol=1e-4)
    env.close()


@pytest.mark.parametrize("n", [1000, 70000])
def test_fused_step_normaliser_equals_the_two_pass_path(n):
    """pnr_filter_fuse: the step kernel writes normalised observations and pushes the statistics itself; same numbers
    as pnr_step followed by pnr_filter_apply (normalised values bit for bit; statistics to float rounding)."""
    from pioneer_b200 import BatchConfig, BatchedPioneerEnv
    from pioneer_b200.obs_filter import MeanStdObsFilter
    a = BatchedPioneerEnv(n, seed=4, batch_config=BatchConfig(max_episode_steps=6))
    b = BatchedPioneerEnv(n, seed=4, batch_config=BatchConfig(max_episode_steps=6))
    fa, fb = MeanStdObsFilter(a), MeanStdObsFilter(b, fused=True)
    g = torch.Generator(device="cuda").manual_seed(1)
    a_max = torch.as_tensor(a.a_max).cuda()
    for it in range(3):
        for t in range(4):
            act = (torch.rand((n, 6), device="cuda", generator=g) * 2 - 1) * a_max
            oa, ra, fl_a = a.step_tensor(act)
            ya = fa(oa.clone())
            ob, rb, fl_b = b.step_tensor(act)                  # already normalised
            assert torch.equal(ra, rb) and torch.equal(fl_a, fl_b)
            if it == 0:                                        # same applied statistics on both sides: bit for bit
                assert torch.equal(ya, ob), (t, float((ya - ob).abs().max()))
            else:                                              # the two accumulators differ by float rounding after a sync
                assert torch.allclose(ya, ob, rtol=0, atol=2e-5), (it, t, float((ya - ob).abs().max()))
        fa.sync(); fb.sync()
        assert fa.n == fb.n == n * 4 * (it + 1)
        np.testing.assert_allclose(fb.mean, fa.mean, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(fb.var, fa.var, rtol=1e-4, atol=1e-8)
    # the env state itself is untouched by the normalisation
    sa, sb = a.state(), b.state()
    assert torch.equal(sa["r"], sb["r"]) and torch.equal(sa["potential"], sb["potential"])
    # unsupported combinations are refused loudly
    from pioneer_b200 import _cabi
    c = BatchedPioneerEnv(64, batch_config=BatchConfig(obs_mode="autoreset"))
    with pytest.raises(_cabi.PioneerB200Error):
        MeanStdObsFilter(c, fused=True)
    for e in (a, b, c):
        e.close()


@pytest.mark.parametrize("obs_mode", ["terminal", "autoreset"])
def test_fused_normaliser_in_the_dynamic_kernel_equals_the_two_pass_path(obs_mode):
    """Dynamic mode (ABA kernel): the warp that owns a tile normalises it before the bulk store.  Same numbers as the
    plain dynamic step followed by pnr_filter_apply -- bit for bit under the same applied statistics -- in both
    observation modes (the auto-reset rows are normalised like any other), ragged last tile included."""
    from pioneer_b200 import BatchConfig, BatchedPioneerEnv, SimulationConfig
    from pioneer_b200.obs_filter import MeanStdObsFilter
    n = 4096 + 13

    def make():
        return BatchedPioneerEnv(n, seed=9, simulation_config=SimulationConfig(gravity=9.81),
                                 batch_config=BatchConfig(mode="dynamic", kp=800.0, kd=200.0, torque_scale=1e4,
                                                          max_episode_steps=3, obs_mode=obs_mode))
    a, b = make(), make()
    fa, fb = MeanStdObsFilter(a), MeanStdObsFilter(b, fused=True)
    g = torch.Generator(device="cuda").manual_seed(2)
    lo, hi = torch.as_tensor(a.r_lo).cuda(), torch.as_tensor(a.r_hi).cuda()
    for it in range(3):
        for t in range(4):
            act = torch.rand((n, 6), device="cuda", generator=g) * (hi - lo) + lo
            oa, ra, fl_a = a.step_tensor(act)
            ya = fa(oa.clone())
            ob, rb, fl_b = b.step_tensor(act)
            assert torch.equal(ra, rb) and torch.equal(fl_a, fl_b)
            if it == 0:
                assert torch.equal(ya, ob), (t, float((ya - ob).abs().max()))
            else:
                assert torch.allclose(ya, ob, rtol=0, atol=2e-5), (it, t, float((ya - ob).abs().max()))
        fa.sync(); fb.sync()
        assert fa.n == fb.n == n * 4 * (it + 1)
        np.testing.assert_allclose(fb.mean, fa.mean, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(fb.var, fa.var, rtol=1e-4, atol=1e-8)
    sa, sb = a.state(), b.state()
    assert torch.equal(sa["r"], sb["r"]) and torch.equal(sa["v"], sb["v"])
    a.close(); b.close()
