"""Host-side logic that needs no GPU: axis-label handling, dictionary defaults."""

import numpy as np
import pytest
import torch

from tomobar_b200.supp.funcs import _data_dims_swapper, _parse_device_argument


def test_axis_swapper_all_permutations():
    x = torch.arange(2 * 3 * 4).reshape(2, 3, 4)
    want = ["detY", "angles", "detX"]
    import itertools

    for perm in itertools.permutations(range(3)):
        labels = [want[p] for p in perm]          # axis i of y carries label want[perm[i]]
        y = x.permute(perm)
        back = _data_dims_swapper(y, labels, want)
        assert torch.equal(back, x)
        assert _data_dims_swapper(tuple(y.shape), labels, want) == (2, 3, 4)


def test_axis_swapper_errors():
    x = torch.zeros(2, 3, 4)
    with pytest.raises(ValueError):
        _data_dims_swapper(x, ["angles", "detX"], ["detY", "angles", "detX"])
    with pytest.raises(ValueError):
        _data_dims_swapper(x, ["angles", "detZ", "detX"], ["detY", "angles", "detX"])


def test_parse_device():
    assert _parse_device_argument(3) == ("gpu", 3)
    assert _parse_device_argument("gpu") == ("gpu", 0)
    assert _parse_device_argument("cpu") == ("cpu", -1)
    with pytest.raises(ValueError):
        _parse_device_argument("tpu")
