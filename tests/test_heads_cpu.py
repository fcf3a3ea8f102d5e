"""CPU: the head modules of equi_articulated_pose_b200/heads.py carry exactly the parameters / buffers of the reference's
classes (state dicts of the fixture made from the reference's own modules load with no missing or unexpected key)."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
from tests.test_gpu_heads import CASES  # noqa: E402


@pytest.mark.parametrize("name,cls,kw,over", CASES)
def test_head_state_dict_keys_match_reference(name, cls, kw, over):
    from equi_articulated_pose_b200 import heads
    g = np.load(os.path.join(GOLD, "ref_heads_small.npz"))
    params = {"dim_in": 16, "mlp": [32, 64], "fc": [64], "k": 8, "kanchor": 60, "temperature": 3.0}
    params.update(over)
    head = getattr(heads, cls)(params, **kw)
    sd = {k[len(name) + 4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(name + "_sd_")}
    res = head.load_state_dict(sd, strict=False)
    assert not res.missing_keys and not res.unexpected_keys, res
    for k, v in head.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
