"""The torch.library harness: ops are registered and their fake (shape-inference) kernels agree with
the documented output shapes - the analogue of CustomOpProp.infer_shape (choose_feat.py:55-58)."""
import pytest

torch = pytest.importorskip("torch")


def test_fake_shapes():
    import lsfa_b200.torch_ops  # noqa: F401  (registers lsfa::*)
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        flow = torch.empty(3, 2, 38, 63, device="cuda")
        key = torch.empty(3, 1024, 38, 63, device="cuda")
        mv = torch.empty(3, 600, 1000, 2, dtype=torch.int32, device="cuda")
        assert torch.ops.lsfa.grid_generator_warp(flow).shape == flow.shape
        assert torch.ops.lsfa.bilinear_sampler(key, flow).shape == key.shape
        g2 = torch.empty(3, 2, 11, 13, device="cuda")
        assert torch.ops.lsfa.bilinear_sampler(key, g2).shape == (3, 1024, 11, 13)
        f = torch.ops.lsfa.mv_pool(mv, 1.0, 0)
        assert f.shape == (3, 2, 38, 63) and f.dtype == torch.float32
        out = torch.ops.lsfa.warp_scale_aggregate(key, mv, key, key, flow, None, 3, 2, 1.0, 0)
        assert out.shape == key.shape
        kb = torch.empty(3, 38, 63, 1024, dtype=torch.bfloat16, device="cuda")
        out = torch.ops.lsfa.warp_scale_aggregate(kb, flow, kb, kb, flow, None, 3, 0, 1.0, 2)
        assert out.shape == kb.shape and out.dtype == torch.bfloat16
        with pytest.raises(Exception):
            torch.ops.lsfa.grid_generator_warp(torch.empty(3, 3, 4, 4, device="cuda"))
