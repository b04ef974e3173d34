"""CPU checks of the oracle and of host-side identities the product relies on."""
import numpy as np


def test_legacy_normal_is_affine_map_of_standard_normal():
    """np.random.normal(loc, scale) consumes the legacy stream exactly like
    standard_normal() and equals loc + scale * z with a rounded product and sum -- the
    identity behind the device-side draw transform of tlbo_rank_loss_counts_normal."""
    rng = np.random.RandomState(0)
    loc = rng.randn(7, 53)
    scale = np.abs(rng.randn(7, 53)) * 1e-2
    np.random.seed(4465)
    a = np.random.normal(np.broadcast_to(loc, (30, 7, 53)), np.broadcast_to(scale, (30, 7, 53)))
    sa = np.random.get_state()
    np.random.seed(4465)
    z = np.random.standard_normal((30, 7, 53))
    sb = np.random.get_state()
    assert (a == loc[None] + scale[None] * z).all()
    assert sa[2] == sb[2] and (sa[1] == sb[1]).all() and sa[3] == sb[3] and sa[4] == sb[4]
    # and the loop form used by the reference (one call per (sample, model))
    np.random.seed(4465)
    for s in range(30):
        for m in range(7):
            v = np.random.normal(loc[m].reshape(-1, 1), scale[m].reshape(-1, 1))
            assert (v.ravel() == a[s, m]).all()
