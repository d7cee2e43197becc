"""g2048_gae_flat_scan -- the re-associated warp-shuffle reverse scan of the GAE recurrence (src/ppo/data_loader.py:103-130).
Opt-in companion of the bit-identical g2048_gae_flat: results must agree with the reference loop within the tolerance
BASELINE.json's north_star states for GAE and returns, 1e-5 relative in fp32."""
import numpy as np
import pytest
import torch

from oracle import c_oracle as CO

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # north_star: "GAE, returns ... within 1e-5 relative in fp32"
TAGS = ["default", "short_eps", "undiscounted", "lowlam", "open_tail", "two"]


@pytest.fixture(scope="module")
def E():
    from g2048 import engine

    return engine


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def assert_within_tolerance(got, want, what):
    """|got - want| <= RTOL * max(|want|, typical magnitude): elementwise relative error, with the buffer's RMS as the
    floor of the denominator (an advantage that cancels to ~0 out of terms of size RMS cannot carry 1e-5 of ITSELF
    in fp32 -- the reference's own loop does not either).  Also: the plain elementwise bound must hold almost everywhere."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    scale = max(float(np.sqrt(np.mean(want * want))), 1e-30)
    err = np.abs(got - want)
    bound = RTOL * np.maximum(np.abs(want), scale)
    worst = int(np.argmax(err - bound))
    assert (err <= bound).all(), f"{what}: |d|={err[worst]:.3e} at {worst}, value {want[worst]:.6e}, rms {scale:.3e}"
    plain = err <= RTOL * np.abs(want)
    assert plain.mean() >= 0.995, f"{what}: only {plain.mean():.5f} of the elements within plain {RTOL} relative"


@pytest.mark.parametrize("tag", TAGS)
def test_scan_matches_the_reference_fixtures(E, golden_ppo, tag):
    """The six (rewards, values, dones) cases whose advantages / returns were produced by the reference's own
    PPODataset._compute_gae_returns (tests/golden/make_golden_ppo.py)."""
    g = golden_ppo
    gamma, lam = g[f"gae_{tag}_params"]
    adv, ret, mom = E.gae_flat(dev(g[f"gae_{tag}_rewards"]), dev(g[f"gae_{tag}_values"]), dev(g[f"gae_{tag}_dones"].astype(np.uint8)),
                               gamma, lam, entry="g2048_gae_flat_scan")
    assert_within_tolerance(adv.cpu().numpy(), g[f"gae_{tag}_adv"], f"{tag} adv")
    assert_within_tolerance(ret.cpu().numpy(), g[f"gae_{tag}_ret"], f"{tag} ret")
    assert mom[0].item() == len(g[f"gae_{tag}_rewards"])
    E.normalize_(adv, mom, 1)
    E.normalize_(ret, mom, 3)
    np.testing.assert_allclose(adv.cpu().numpy(), g[f"gae_{tag}_adv_norm"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(ret.cpu().numpy(), g[f"gae_{tag}_ret_norm"], rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("n,done_rate", [(1, 1.0), (7, 0.0), (4095, 0.01), (4096, 0.01), (4097, 0.0), (12289, 0.0), (20000, 0.5),
                                         (20001, 1.0), (50_000, 0.0), (300_000, 1 / 300), (2_000_000, 1 / 3000), (8_400_003, 1 / 300)])
def test_scan_matches_the_bit_exact_kernel_at_scale(E, n, done_rate):
    """done_rate 0 makes every tile's aggregate non-zero: the decoupled look-back then runs over many tiles."""
    rng = np.random.default_rng(n)
    r = (rng.integers(0, 64, n) * 4 * (rng.random(n) < 0.4)).astype(np.float32)
    v = (rng.standard_normal(n) * 10).astype(np.float32)
    d = (rng.random(n) < done_rate).astype(np.uint8)
    want_a, want_r = CO.gae(r, v, d, 0.99, 0.95)
    adv, ret, mom = E.gae_flat(dev(r), dev(v), dev(d), 0.99, 0.95, entry="g2048_gae_flat_scan")
    assert_within_tolerance(adv.cpu().numpy(), want_a, "adv")
    assert_within_tolerance(ret.cpu().numpy(), want_r, "ret")
    m = mom.cpu().numpy()
    assert m[0] == n
    # a thread's 8 steps are summed in fp32 before they join the fp64 sums: 1e-6 relative (the sums only feed mean / std)
    np.testing.assert_allclose(m[1], adv.double().sum().item(), rtol=1e-6, atol=1e-3 * max(1.0, float(adv.abs().double().sum()) * 1e-4))
    np.testing.assert_allclose(m[2], (adv.double() ** 2).sum().item(), rtol=1e-6)
    np.testing.assert_allclose(m[4], (ret.double() ** 2).sum().item(), rtol=1e-6)


def test_scan_at_c4_size_with_thousand_step_episodes(E):
    """2^26 steps, episodes of ~1000 steps (where the walking kernel drops to 2.4 TB/s): against the bit-exact kernel."""
    n = 1 << 26
    gen = torch.Generator(device="cuda").manual_seed(9)
    r = (torch.randint(0, 64, (n,), device="cuda", generator=gen) * 4).float() * (torch.rand(n, device="cuda", generator=gen) < 0.4)
    v = torch.randn(n, device="cuda", generator=gen) * 10
    d = (torch.rand(n, device="cuda", generator=gen) < 1 / 1000).to(torch.uint8)
    want_a, want_r, want_m = E.gae_flat(r, v, d, 0.99, 0.95)
    adv, ret, mom = E.gae_flat(r, v, d, 0.99, 0.95, entry="g2048_gae_flat_scan")
    for got, want, what in ((adv, want_a, "adv"), (ret, want_r, "ret")):
        scale = want.double().pow(2).mean().sqrt()
        err = (got.double() - want.double()).abs()
        bound = RTOL * torch.maximum(want.double().abs(), scale)
        assert bool((err <= bound).all()), what
        assert float((err <= RTOL * want.double().abs()).double().mean()) >= 0.995, what
    torch.testing.assert_close(mom, want_m, rtol=1e-6, atol=1e-3)


def test_scan_unaligned_views_and_nonbinary_dones(E):
    rng = np.random.default_rng(53)
    n = 40_000
    r = (rng.integers(0, 64, n + 3) * 4).astype(np.float32)
    v = rng.standard_normal(n + 3).astype(np.float32)
    d = ((rng.random(n + 3) < 0.01) * rng.integers(1, 256, n + 3)).astype(np.uint8)  # any non-zero byte is a done
    want_a, want_r = CO.gae(r[3:], v[3:], d[3:], 0.99, 0.95)
    adv, ret, _ = E.gae_flat(dev(r)[3:], dev(v)[3:], dev(d)[3:], 0.99, 0.95, entry="g2048_gae_flat_scan")
    assert_within_tolerance(adv.cpu().numpy(), want_a, "adv")
    assert_within_tolerance(ret.cpu().numpy(), want_r, "ret")
