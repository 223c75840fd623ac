"""Axis-label handling and device-argument parsing (behaviour of tomobar/supp/funcs.py:84-206)."""

from __future__ import annotations

from typing import List, Sequence, Tuple, Union

import torch

VALID_LABELS = ("angles", "detX", "detY")


def _axes_permutation(data_axes_labels: Sequence[str], required_labels_order: Sequence[str]) -> List[int]:
    """Permutation p such that ``data.permute(p)`` has its axes in ``required_labels_order``.
    Error behaviour follows ``_swap_data_axes_to_accepted`` (funcs.py:99-141)."""
    if len(data_axes_labels) != len(required_labels_order):
        raise ValueError("Warning: The mismatch in length between provided labels and data dimensions.")
    for label in data_axes_labels:
        if label not in required_labels_order:
            raise ValueError(
                f'Axis title "{label}" is not valid, please use one of these: "angles", "detX", or "detY"'
            )
    return [list(data_axes_labels).index(label) for label in required_labels_order]


def _data_dims_swapper(data, data_axes_labels_order: Sequence[str], required_labels_order: Sequence[str]):
    """Re-orders the axes of ``data`` (a tensor, or a shape tuple) to ``required_labels_order``
    (funcs.py:190-206).  Returns a view for tensors, like the reference's ``swapaxes``."""
    perm = _axes_permutation(list(data_axes_labels_order), list(required_labels_order))
    if isinstance(data, tuple):
        return tuple(data[p] for p in perm)
    if perm == list(range(len(perm))):
        return data
    return data.permute(perm)


def _raw_buffer_view(data):
    """What the reference's GPULink hands to ASTRA for a non-contiguous (axis-swapped) view: the raw pointer with
    the view's LOGICAL shape and a dense pitch (astra_base.py:533-535), i.e. the underlying buffer re-read as if it
    were contiguous in the new shape.  Only the ``compat_view_bug`` switches use this;
    the default everywhere is the logically correct array."""
    if data.is_contiguous():
        return data
    import torch

    strides, acc = [], 1
    for d in reversed(data.shape):
        strides.append(acc)
        acc *= int(d)
    return torch.as_strided(data, tuple(data.shape), tuple(reversed(strides)), data.storage_offset())


def _parse_device_argument(device_int_or_string: Union[int, str]) -> Tuple[str, int]:
    """funcs.py:174-187."""
    if isinstance(device_int_or_string, int):
        return "gpu", device_int_or_string
    if device_int_or_string == "gpu":
        return "gpu", 0
    if device_int_or_string == "cpu":
        return "cpu", -1
    raise ValueError(
        'Unknown device {0}. Expecting either "cpu" or "gpu" strings OR the gpu device integer'.format(
            device_int_or_string
        )
    )
