"""Host logic of strotss_tensorflow_b200/tf_adapter.py without a GPU: the adapter bound to tests/tf_standin.py, its DLPack
structs against torch's capsules, the output-buffer capsule + deleter round trip (with a fake library handing out host
memory), device parsing and the reference's KeyError.  The compute calls themselves run in tests/test_gpu_tf_adapter.py."""
import ctypes as C
import gc
import importlib
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def tfa():
    import tf_standin
    saved = sys.modules.get("tensorflow")
    sys.modules["tensorflow"] = tf_standin.make_module()
    import strotss_tensorflow_b200.tf_adapter as mod
    mod = importlib.reload(mod)
    try:
        yield mod
    finally:
        if saved is None:
            sys.modules.pop("tensorflow", None)
        else:
            sys.modules["tensorflow"] = saved
        importlib.reload(mod)


def test_adapter_refuses_without_tensorflow():
    import strotss_tensorflow_b200.tf_adapter as mod
    mod = importlib.reload(mod)
    assert not mod._HAVE_TF
    with pytest.raises(RuntimeError, match="TensorFlow"):
        mod.self_similarity(None, None)


def test_dlpack_structs_match_torch_capsules(tfa):
    t = torch.arange(12, dtype=torch.float32).reshape(3, 4)[1:]          # a view with a storage offset
    ptr, keep = tfa._dev_ptr(tfa.tf.Tensor(t))
    assert ptr.value == t.data_ptr()
    cap = torch.utils.dlpack.to_dlpack(t)
    m = C.cast(C.pythonapi.PyCapsule_GetPointer(cap, b"dltensor"), C.POINTER(tfa._DLManagedTensor)).contents.dl_tensor
    assert (m.ndim, m.dtype.code, m.dtype.bits, m.dtype.lanes) == (2, tfa._KDL_FLOAT, 32, 1)
    assert [m.shape[0], m.shape[1]] == [2, 4]


def test_output_buffers_enter_through_dlpack_and_are_returned(tfa, monkeypatch):
    freed = []

    class FakeLib:
        def strotss_device_alloc(self, h, nbytes, out):
            self.buf = (C.c_float * (nbytes // 4))(*range(nbytes // 4))
            out._obj.value = C.addressof(self.buf)
            return 0

        def strotss_device_free(self, h, p):
            freed.append(p.value if hasattr(p, "value") else p)
            return 0

        def strotss_last_error(self, h):
            return b""

    monkeypatch.setattr(tfa, "_KDL_CUDA", 1)                              # kDLCPU: host memory stands in for the device buffer
    lib = FakeLib()
    out = tfa._Output(lib, C.c_void_p(1), 0, (2, 3))
    t = out.to_tf()
    assert t.t.tolist() == [[0.0, 1.0, 2.0], [3.0, 4.0, 5.0]] and len(tfa._live) == 1
    del t
    gc.collect()
    assert freed == [C.addressof(lib.buf)] and len(tfa._live) == 0
    # the deleter thunk and its table survive a reload of the module (tensors created before the reload still point at them)
    thunk, table = tfa._free_output, tfa._live
    mod = importlib.reload(tfa)
    assert mod._free_output is thunk and mod._live is table


def test_device_parsing_and_distance_keyerror(tfa):
    class D:
        device = "/job:localhost/replica:0/task:0/device:GPU:3"
    assert tfa._device_index(D()) == 3
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tfa._device_index(tfa.tf.Tensor(torch.zeros(1)))
    with pytest.raises(KeyError):
        tfa.relaxed_emd(tfa.tf.Tensor(torch.zeros(2, 3)), tfa.tf.Tensor(torch.zeros(2, 3)), distance="sinkhorn")


def test_custom_gradient_of_the_standin_scales_with_upstream(tfa):
    tf = tfa.tf

    @tf.custom_gradient
    def op(a, b):
        return a * 2.0, (lambda up: (up * 2.0, None))
    a = tf.Tensor(torch.ones(3, requires_grad=True))
    with tf.GradientTape() as tape:
        loss = tf.reduce_mean(op(a, tf.Tensor(torch.ones(3)))) * 3.0
    assert torch.allclose(tape.gradient(loss, a).t, torch.full((3,), 2.0))
