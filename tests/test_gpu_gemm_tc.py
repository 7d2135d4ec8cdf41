"""The tcgen05 split-bf16 GEMM through the C ABI, against fp64 on the same inputs: both operand layouts, ragged edges
(TMA zero fill), the fused epilogue, persistent multi-tile scheduling and the deterministic split reduction."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops():
    from legommenders_b200 import ops as _ops
    return _ops


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


@pytest.mark.parametrize('rows,cols', [(5, 300), (129, 256), (1000, 64), (3, 4), (4100, 768)])
def test_split_planes(ops, rows, cols):
    g = torch.Generator().manual_seed(rows + cols)
    x = torch.randn(rows, cols, generator=g) * 3
    p = ops.split_planes(x.cuda())
    assert p.ld % 8 == 0 and p.ld >= cols
    hi, lo = p.hi.float().cpu(), p.lo.float().cpu()
    assert torch.equal(hi[:, :cols], x.bfloat16().float())                       # round-to-nearest hi plane, bit exact
    assert torch.equal(lo[:, :cols], (x - x.bfloat16().float()).bfloat16().float())
    assert (hi[:, cols:] == 0).all() and (lo[:, cols:] == 0).all()
    assert ((hi + lo)[:, :cols] - x).abs().max().item() <= 2.0 ** -16 * x.abs().max().item()
    t = ops.split_planes(x.cuda(), transpose=True)
    assert torch.equal(t.hi.float().cpu()[:, :rows], hi[:, :cols].t())
    assert torch.equal(t.lo.float().cpu()[:, :rows], lo[:, :cols].t())
    assert (t.hi.float().cpu()[:, rows:] == 0).all()
    p2, cs = ops.split_planes(x.cuda(), colsum=True)
    assert torch.equal(p2.hi, p.hi) and torch.equal(p2.lo, p.lo)
    assert (cs.double().cpu() - x.double().sum(0)).abs().max().item() <= 1e-5 * max(1.0, x.abs().sum(0).max().item())


@pytest.mark.parametrize('M,N,K', [(128, 128, 64), (128, 128, 256), (300, 256, 256), (1000, 768, 256), (257, 64, 300),
                                   (4000, 256, 300), (513, 32, 64), (20000, 256, 256), (256, 256, 4096), (40000, 768, 256), (40000, 256, 300), (38000, 128, 64)])
def test_tc_gemm_k_major(ops, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g)
    b = torch.randn(N, K, generator=g) / K ** 0.5
    y = ops.tc_gemm(ops.split_planes(a.cuda()), ops.split_planes(b.cuda()), False, M, N, K)
    ref = a.double() @ b.double().t()
    assert rel(y, ref) <= 3e-5


def test_tc_gemm_epilogue(ops):
    g = torch.Generator().manual_seed(1)
    M, N, K = 700, 256, 256
    a, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / 16
    bias = torch.randn(N, generator=g)
    rm = (torch.rand(M, generator=g) < 0.6).long()
    base = torch.randn(M, N, generator=g)
    ap, bp = ops.split_planes(a.cuda()), ops.split_planes(b.cuda())
    pre = a.double() @ b.double().t() + bias.double()
    y = ops.tc_gemm(ap, bp, False, M, N, K, bias=bias.cuda(), rowmask=rm.cuda(), act=1)
    assert rel(y, torch.tanh(pre) * rm.unsqueeze(-1)) <= 3e-5
    y = ops.tc_gemm(ap, bp, False, M, N, K, bias=bias.cuda(), act=2)
    assert rel(y, torch.relu(pre)) <= 3e-5
    out = base.clone().cuda()
    ops.tc_gemm(ap, bp, False, M, N, K, out=out, accumulate=True)
    assert rel(out, base.double() + a.double() @ b.double().t()) <= 3e-5


@pytest.mark.parametrize('T,N,K', [(256, 128, 128), (5000, 256, 256), (13200, 768, 256), (13200, 256, 300), (1024, 32, 64),
                                   (116160, 256, 256)])
def test_tc_gemm_mn_major_weight_gradient(ops, T, N, K):
    """dW[N,K] = dY[T,N]^T · X[T,K]: reduction over T token rows, split across CTAs, deterministic."""
    g = torch.Generator().manual_seed(T + N + K)
    dy = torch.randn(T, N, generator=g)
    x = torch.randn(T, K, generator=g)
    dyp, xp = ops.split_planes(dy.cuda()), ops.split_planes(x.cuda())
    dw = ops.tc_gemm(dyp, xp, True, N, K, T)
    ref = dy.double().t() @ x.double()
    assert rel(dw, ref) <= 3e-5
    dw2 = ops.tc_gemm(dyp, xp, True, N, K, T)
    assert torch.equal(dw, dw2)
