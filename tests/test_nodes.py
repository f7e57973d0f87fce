"""Node schema parity with the reference (ADP:789-817, 849-870, 886-915, 1473-1483)."""
import torch
import torch.nn as nn

import comfyui_parallelanything_b200 as pa


def test_mapping_keys():
    assert set(pa.NODE_CLASS_MAPPINGS) == {"ParallelAnything", "ParallelDevice", "ParallelDeviceList"}
    assert pa.NODE_DISPLAY_NAME_MAPPINGS["ParallelDevice"] == "Parallel Device Config"
    assert pa.NODE_DISPLAY_NAME_MAPPINGS["ParallelDeviceList"] == "Parallel Device List (1-4x)"
    assert pa.NODE_DISPLAY_NAME_MAPPINGS["ParallelAnything"] == "Parallel Anything (True Multi-GPU)"


def test_schema_equals_reference(reference):
    for name in ("ParallelDevice", "ParallelDeviceList", "ParallelAnything"):
        ours, ref = pa.NODE_CLASS_MAPPINGS[name], reference.NODE_CLASS_MAPPINGS[name]
        assert ours.INPUT_TYPES() == ref.INPUT_TYPES(), name
        for attr in ("RETURN_TYPES", "RETURN_NAMES", "FUNCTION", "CATEGORY"):
            assert getattr(ours, attr) == getattr(ref, attr), (name, attr)
        assert hasattr(ours, getattr(ours, "FUNCTION"))
    assert pa.NODE_DISPLAY_NAME_MAPPINGS == reference.NODE_DISPLAY_NAME_MAPPINGS
    assert pa.ParallelDevice.DESCRIPTION == reference.ParallelDevice.DESCRIPTION


def test_chain_building_matches_reference(reference):
    a = pa.ParallelDevice().add_device("cpu", 40)[0]
    a = pa.ParallelDevice().add_device("cpu", 60, a)[0]
    b = reference.ParallelDevice().add_device("cpu", 40)[0]
    b = reference.ParallelDevice().add_device("cpu", 60, b)[0]
    assert a == b
    assert pa.ParallelDeviceList().create_list("cpu", 50, "cpu", 0, "cpu", 25) == \
        reference.ParallelDeviceList().create_list("cpu", 50, "cpu", 0, "cpu", 25)


def test_add_device_copies_previous():
    prev = pa.ParallelDevice().add_device("cpu", 10)[0]
    new = pa.ParallelDevice().add_device("cpu", 20, prev)[0]
    assert len(prev) == 1 and len(new) == 2


def test_passthrough_on_bad_input():
    m = nn.Linear(2, 2)
    node = pa.ParallelAnything()
    assert node.setup_parallel(None, [{"device": "cpu", "percentage": 100}]) == (None,)
    assert node.setup_parallel(m, []) == (m,)
    out, = node.setup_parallel(m, [{"device": "bogus:9", "percentage": 100, "weight": 1.0}])
    assert out is m and not getattr(m, "_true_parallel_active", False)


def test_repair_stranded_is_a_noop_without_a_cuda_home():
    """C21 (ADP:932-961): only a model that ComfyUI believes lives on a CUDA device is moved back."""
    from comfyui_parallelanything_b200 import nodes

    class Patcher:
        load_device = torch.device("cpu")

    m = nn.Linear(4, 4)
    nodes._repair_stranded(Patcher(), m)
    assert m.weight.device.type == "cpu"
    Patcher.load_device = None
    nodes._repair_stranded(Patcher(), m)            # no load_device and no comfy.model_management: leave it alone
    assert m.weight.device.type == "cpu"
