"""Down, Right, Up, Left priority policy -- the reference's src/actions/act_drul.py:5-49."""
from .. import engine as E
from ._common import prepare


def act_drul(rng_key, obs, mask, rng_mode=None):
    """(key, obs (4,4,31), mask (4,)) -> (action, None, None); first legal of [3, 2, 1, 0]."""
    _, status, batched = prepare(None, obs, mask)
    actions, _ = E.act(E.POLICY_DRUL, status, None, 0, 0, E.resolve_rng_mode(rng_mode))
    if not batched:
        return actions[0], None, None
    return actions, None, None


act_drul.policy_id = E.POLICY_DRUL
