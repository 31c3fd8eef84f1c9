import json
import os

from internnav_b200.manifest import navdp_shapes, random_navdp_state_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_manifest_matches_reference_dump():
    """oracle/navdp_manifest.json was dumped from the reference class's state_dict(); the product-side generator must
    reproduce every key and shape (order included)."""
    with open(os.path.join(ROOT, "oracle", "navdp_manifest.json")) as fh:
        ref = json.load(fh)
    mine = navdp_shapes()
    assert list(mine.keys()) == list(ref.keys())
    for k, (shape, _) in ref.items():
        assert list(mine[k]) == shape, k


def test_random_state_dict_shapes():
    sd = random_navdp_state_dict(seed=1)
    shapes = navdp_shapes()
    assert set(sd) == set(shapes)
    assert all(tuple(sd[k].shape) == tuple(shapes[k]) for k in sd)
    assert sum(v.numel() for v in sd.values()) > 98e6
