"""Host-side contract of the drop-in modules (no GPU): plans are never copied or pickled, and any write to the
weights invalidates the packed copies (train_stcd.py:81-87,326 deep-copies models)."""
import copy
import ctypes as C
import io
import pickle

import torch

from stcd_b200 import siamunet


class _FakePlan:
    """Stands in for stcd_b200.plan.Plan: holds what makes a real one un-picklable (a ctypes pointer)."""

    def __init__(self):
        self._h = C.c_void_p(1234)
        self.lib = C.CDLL(None)


def _net_with_plan():
    net = siamunet.SiamUnet_diff(3, 2).eval()
    net._plans[(0, 64, 64, 2, None)] = _FakePlan()
    net._fp = net._weights_fingerprint()
    return net


def test_deepcopy_drops_plans():
    net = _net_with_plan()
    twin = copy.deepcopy(net)
    assert twin._plans == {} and twin._fp is None and twin._fp_tensors is None
    assert len(net._plans) == 1                      # the original keeps its plans
    for (k, a), (_, b) in zip(net.state_dict().items(), twin.state_dict().items()):
        assert torch.equal(a, b), k
        assert a.data_ptr() != b.data_ptr()


def test_pickle_and_torch_save():
    net = _net_with_plan()
    twin = pickle.loads(pickle.dumps(net))
    assert twin._plans == {}
    buf = io.BytesIO()
    torch.save(net, buf)
    buf.seek(0)
    again = torch.load(buf, weights_only=False)
    assert again._plans == {}
    assert torch.equal(again.conv11.weight, net.conv11.weight)


def test_fingerprint_sees_weight_writes():
    net = _net_with_plan()
    fp = net._fp
    assert net._weights_fingerprint() == fp
    with torch.no_grad():
        net.conv11.weight.mul_(1.0)                 # in-place write bumps _version
    assert net._weights_fingerprint() != fp
    net._fp = net._weights_fingerprint()
    net.bn11.load_state_dict(net.bn11.state_dict())  # a CHILD's load_state_dict copies in place
    assert net._weights_fingerprint() != net._fp
    net._fp = net._weights_fingerprint()
    net.conv12.weight.data = net.conv12.weight.data.clone()   # storage swapped under the same Parameter
    assert net._weights_fingerprint() != net._fp


def test_top_level_mutators_invalidate():
    net = _net_with_plan()
    net.load_state_dict(net.state_dict())
    assert net._plans == {} and net._fp is None
    net = _net_with_plan()
    net.float()
    assert net._plans == {}
