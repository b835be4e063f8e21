"""dist.FlatGradBucket: every parameter gradient is a view of the one flat buffer WITH the parameter's own strides, so a
channels-last convolution weight (Trainer(amp=True)) accumulates into a channels-last gradient (CPU, no GPU needed)."""
import torch

from agent0_b200.dist import FlatGradBucket


def test_gradient_views_follow_the_parameter_layout_and_share_the_flat_buffer():
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(4, 8, 3).to(memory_format=torch.channels_last)
    lin = torch.nn.Linear(8 * 7 * 7, 3)
    params = list(conv.parameters()) + list(lin.parameters())
    bucket = FlatGradBucket(params)
    assert bucket.flat.numel() == sum(p.numel() for p in params)
    x = torch.randn(2, 4, 9, 9).contiguous(memory_format=torch.channels_last)
    lin(conv(x).flatten(1)).sum().backward()
    off = 0
    for p in params:
        assert p.grad.stride() == p.stride() and p.grad.data_ptr() == bucket.flat.data_ptr() + 4 * off
        off += p.numel()
    # the same numbers as a plain backward pass with separate gradient tensors
    ref_conv = torch.nn.Conv2d(4, 8, 3)
    ref_conv.load_state_dict(conv.state_dict())
    ref_lin = torch.nn.Linear(8 * 7 * 7, 3)
    ref_lin.load_state_dict(lin.state_dict())
    ref_lin(ref_conv(x.contiguous()).flatten(1)).sum().backward()
    for p, q in zip(params, list(ref_conv.parameters()) + list(ref_lin.parameters())):
        assert torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-6)
    bucket.zero_()
    assert all(float(p.grad.abs().sum()) == 0.0 for p in params)
