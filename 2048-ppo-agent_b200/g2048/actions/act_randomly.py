"""Uniform choice among the legal actions -- the reference's src/actions/act_randomly.py:5-56.

Called directly it runs the `g2048_act` kernel on the given key(s) and mask(s); handed to
``BatchRunner`` it selects the fused random-policy kernels (the attribute ``policy_id``).
"""
from .. import engine as E
from ._common import prepare


def act_randomly(rng_key, obs, mask, rng_mode=None):
    """(key, obs (4,4,31), mask (4,)) -> (action, log_prob, None); a leading batch axis is accepted.

    Actions: 0=Left, 1=Up, 2=Right, 3=Down.  action ~ categorical over the legal actions with
    jax.random's own draws; log_prob = log(1 / n_legal).
    """
    keys, status, batched = prepare(rng_key, obs, mask)
    actions, log_probs = E.act(E.POLICY_RANDOM, status, keys, 0, 0, E.resolve_rng_mode(rng_mode))
    if not batched:
        return actions[0], log_probs[0], None
    return actions, log_probs, None


act_randomly.policy_id = E.POLICY_RANDOM
