"""Sample records (g2048_pack_samples / g2048_gather_samples) and the epoch-at-once minibatch feed: same batches, bit
for bit, as the per-array gather of the flat buffer (src/ppo/data_loader.py:61-67,132-166,217-223)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    import g2048
    from g2048 import engine as E

    runner = g2048.BatchRunner(init_seed=21, act_fn=g2048.act_randomly)
    buf = g2048.RolloutBuffer(31, 16, 4)
    buf.store_flat(runner.run_flat_batch(700))
    packed = buf.get_packed()
    packed["values"].copy_(torch.randn_like(packed["values"]))
    return g2048, E, packed


def test_records_hold_the_normalised_fields(setup):
    g2048, E, packed = setup
    dones = E.meta_dones(packed["meta"])
    adv, ret, mom = E.gae_flat(packed["rewards"], packed["values"], dones, 0.99, 0.95)
    rec = E.pack_samples(packed, adv, ret, mom).cpu().numpy().view(E.SAMPLE_RECORD).reshape(-1)
    want_adv = E.normalize_(adv.clone(), mom, 1).cpu().numpy()
    want_ret = E.normalize_(ret.clone(), mom, 3).cpu().numpy()
    np.testing.assert_array_equal(rec["board"], packed["boards"].cpu().numpy().view(np.uint64))
    np.testing.assert_array_equal(rec["meta"], packed["meta"].cpu().numpy())
    for name, want in (("reward", packed["rewards"]), ("log_prob", packed["log_probs"]), ("value", packed["values"])):
        np.testing.assert_array_equal(rec[name], want.cpu().numpy())
    np.testing.assert_array_equal(rec["advantage"], want_adv)
    np.testing.assert_array_equal(rec["ret"], want_ret)
    raw = E.pack_samples(packed, adv, ret, None).cpu().numpy().view(E.SAMPLE_RECORD).reshape(-1)
    np.testing.assert_array_equal(raw["advantage"], adv.cpu().numpy())


@pytest.mark.parametrize("obs_dtype", [torch.float32, torch.bfloat16, None])
@pytest.mark.parametrize("m", [1, 333, 4096])
def test_gather_from_records_equals_gather_from_arrays(setup, obs_dtype, m):
    g2048, E, packed = setup
    n = packed["boards"].shape[0]
    dones = E.meta_dones(packed["meta"])
    adv, ret, mom = E.gae_flat(packed["rewards"], packed["values"], dones, 0.99, 0.95)
    records = E.pack_samples(packed, adv, ret, mom)
    E.normalize_(adv, mom, 1)
    E.normalize_(ret, mom, 3)
    idx = torch.randint(0, n, (m,), device="cuda")
    a = E.gather_minibatch(idx, packed, adv, ret, obs_dtype)
    b = E.gather_samples(idx, records, obs_dtype)
    assert a.keys() == b.keys()
    for k in a:
        assert torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("kw", [dict(), dict(max_samples_per_epoch=5000, shuffle_on_reset=True), dict(drop_last=False),
                                dict(obs_dtype=None)])
def test_epoch_prefetch_yields_the_same_batches(setup, kw):
    g2048, E, packed = setup
    outs, flat = [], []
    for records, prefetch in ((False, False), (True, False), (True, True)):
        torch.manual_seed(5)
        src = g2048.DevicePPOBatches(packed, 0.99, 0.95, batch_size=512, sample_records=records, epoch_prefetch=prefetch, **kw)
        epochs = []
        for _ in range(2):
            src.reset_epoch()
            epochs.append([{k: v.clone() for k, v in b.items()} for b in src])
        outs.append(epochs)
        flat.append((src.advantages.clone(), src.returns.clone()))  # produced on first use in the records modes
    for adv, ret in flat[1:]:
        assert torch.equal(adv, flat[0][0]) and torch.equal(ret, flat[0][1])
    for other in outs[1:]:
        for ea, eb in zip(outs[0], other):
            assert len(ea) == len(eb) > 0
            for ba, bb in zip(ea, eb):
                assert ba.keys() == bb.keys()
                for k in ba:
                    assert torch.equal(ba[k], bb[k]), k
