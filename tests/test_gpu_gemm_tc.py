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


def test_split_planes_multi_matches_single(ops):
    """All weight matrices of a step in one launch: bit-identical to the per-matrix split, pad columns zero."""
    g = torch.Generator().manual_seed(11)
    shapes = [(768, 256), (256, 256), (256, 300), (1, 4), (5, 12), (256, 256), (300, 64)]
    mats = [(torch.randn(r, c, generator=g) * 2).cuda() for r, c in shapes]
    multi = ops.split_planes_multi(mats)
    for x, p in zip(mats, multi):
        q = ops.split_planes(x)
        assert p.ld == q.ld and torch.equal(p.hi, q.hi) and torch.equal(p.lo, q.lo)
    with pytest.raises(RuntimeError):
        ops.split_planes_multi([mats[0]] * 17)


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


@pytest.mark.parametrize('M,N,K', [(700, 256, 300), (40000, 256, 256), (130, 768, 256)])
def test_tc_gemm_fused_producer_epilogue(ops, M, N, K):
    """lk_tc_gemm_ex: row mask from token ids, small-table row addends, plane output without an fp32 store, column sums,
    accumulate-without-store — the fusions the native NRMS step relies on."""
    g = torch.Generator().manual_seed(M + N)
    a, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / 16
    bias = torch.randn(N, generator=g)
    tok = torch.randint(-1, 50, (M,), generator=g)                       # -1 = not a title token
    tab0, tab1 = torch.randn(18, N, generator=g), torch.randn(3, N, generator=g)
    id0 = torch.where(torch.rand(M, generator=g) < 0.1, torch.randint(0, 18, (M,), generator=g), torch.full((M,), -1))
    id1 = torch.where(torch.rand(M, generator=g) < 0.2, torch.randint(0, 3, (M,), generator=g), torch.full((M,), -1))
    ap, bp = ops.split_planes(a.cuda()), ops.split_planes(b.cuda())
    ref = (a.double() @ b.double().t() + bias.double()) * (tok > -1).double().unsqueeze(-1)
    ref = ref + torch.where((id0 > -1).unsqueeze(-1), tab0.double()[id0.clamp(min=0)], torch.zeros((), dtype=torch.float64))
    ref = ref + torch.where((id1 > -1).unsqueeze(-1), tab1.double()[id1.clamp(min=0)], torch.zeros((), dtype=torch.float64))
    y, planes, cs = ops.tc_gemm_ex(ap, bp, M, N, K, bias=bias.cuda(), rowmask=tok.cuda(), rowmask_is_ids=True,
                                   add0=(id0.cuda(), tab0.cuda()), add1=(id1.cuda(), tab1.cuda()), want_planes=True, want_colsum=True)
    assert rel(y, ref) <= 3e-5
    assert torch.equal(planes.hi.float().cpu()[:, :N], y.cpu().bfloat16().float())          # planes are the split of what was stored
    assert torch.equal(planes.lo.float().cpu()[:, :N], (y.cpu() - y.cpu().bfloat16().float()).bfloat16().float())
    assert (cs.double().cpu() - y.double().cpu().sum(0)).abs().max().item() <= 2e-5 * max(1.0, y.abs().sum(0).max().item())
    # planes only (no fp32 result at all), and accumulate onto C without storing C
    y2, planes2, _ = ops.tc_gemm_ex(ap, bp, M, N, K, bias=bias.cuda(), store_c=False, want_planes=True)
    assert y2 is None
    full = a.double() @ b.double().t() + bias.double()
    assert rel(planes2.hi.float()[:, :N] + planes2.lo.float()[:, :N], full) <= 3e-5
    base = torch.randn(M, N, generator=g)
    c = base.clone().cuda()
    _, planes3, cs3 = ops.tc_gemm_ex(ap, bp, M, N, K, out=c, store_c=False, accumulate=True, want_planes=True, want_colsum=True)
    assert torch.equal(c.cpu(), base)                                                        # C untouched
    tot = base.double() + a.double() @ b.double().t()
    assert rel(planes3.hi.float()[:, :N] + planes3.lo.float()[:, :N], tot) <= 3e-5
    assert (cs3.double().cpu() - tot.sum(0)).abs().max().item() <= 3e-5 * max(1.0, tot.abs().sum(0).max().item())


@pytest.mark.parametrize('M,N,K', [(300, 256, 256), (40000, 256, 768), (5000, 300, 256), (129, 64, 32)])
def test_tc_gemm_input_gradient_with_untransposed_weight(ops, M, N, K):
    """dX[M,N] = dY[M,K] · W[K,N] with W's own (row-major [K,N]) planes as an MN-major B operand: no transposed weight copies."""
    g = torch.Generator().manual_seed(M + K)
    dy = torch.randn(M, K, generator=g)
    w = torch.randn(K, N, generator=g) / K ** 0.5
    y, _, _ = ops.tc_gemm_ex(ops.split_planes(dy.cuda()), ops.split_planes(w.cuda()), M, N, K, b_mn=True)
    assert rel(y, dy.double() @ w.double()) <= 3e-5


@pytest.mark.parametrize('M,N,K', [(1000, 768, 256), (300, 512, 320), (4000, 1024, 64)])
def test_tc_gemm_wide_tile(ops, M, N, K):
    """512 <= N <= 4096 with N % 256 == 0 runs on 128x256 tiles (two 96 KB stages, the whole TMEM): same answer as fp64."""
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g)
    b = torch.randn(N, K, generator=g) / K ** 0.5
    bias = torch.randn(N, generator=g)
    y = ops.tc_gemm(ops.split_planes(a.cuda()), ops.split_planes(b.cuda()), False, M, N, K, bias=bias.cuda())
    assert rel(y, a.double() @ b.double().t() + bias.double()) <= 3e-5
    # weight-gradient layout (both operands MN-major, reduction over rows) at a wide N
    dy = torch.randn(K * 8, 256, generator=g)
    x = torch.randn(K * 8, N, generator=g)
    dw = ops.tc_gemm(ops.split_planes(dy.cuda()), ops.split_planes(x.cuda()), True, 256, N, K * 8)
    assert rel(dw, dy.double().t() @ x.double()) <= 3e-5


def test_deferred_column_sums(ops):
    """Bias-type gradients left as partial sums by their producers (GEMM epilogue, plane split, per-sequence sums) and finished by
    ONE lk_colsum_finish_multi launch: equal to the direct reductions, deterministic, strided jobs and wide jobs included."""
    import ctypes
    from legommenders_b200._lib import call, ptr
    g = torch.Generator().manual_seed(3)
    M, N, K = 1000, 256, 256
    a, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / 16
    ap, bp = ops.split_planes(a.cuda()), ops.split_planes(b.cuda())
    nparts = (M + 127) // 128 * 4
    part = torch.full((nparts, N), float('nan'), device='cuda')
    ep = ops.GemmEpilogue()
    ep.colsum_part = ptr(part)
    y = torch.empty(M, N, device='cuda')
    ws = torch.empty(1 << 20, dtype=torch.uint8, device='cuda')
    call('lk_tc_gemm_ex', ptr(ap.hi), ptr(ap.lo), ap.ld, 0, ptr(bp.hi), ptr(bp.lo), bp.ld, 0, ptr(y), N, M, N, K, ctypes.addressof(ep), ptr(ws),
         ws.numel())
    # plane split with partial column sums
    x = torch.randn(777, 256, generator=g).cuda()
    hi = torch.empty(2, 777, 256, dtype=torch.bfloat16, device='cuda')
    sp = torch.full(((777 + 63) // 64, 256), float('nan'), device='cuda')
    call('lk_split_bf16_partial', ptr(x), 777, 256, 256, ptr(hi[0]), ptr(hi[1]), 256, ptr(sp))
    ref_planes = ops.split_planes(x)
    assert torch.equal(hi[0], ref_planes.hi) and torch.equal(hi[1], ref_planes.lo)
    # a strided job (columns 256.. of a wider partial array) and a wide one (4608 columns, several column blocks per CTA)
    wide = torch.randn(333, 5632, generator=g).cuda()
    outs = [torch.empty(N, device='cuda'), torch.empty(256, device='cuda'), torch.empty(4608, device='cuda'), torch.empty(256, device='cuda')]
    jobs = [(part, outs[0], False), (sp, outs[1], False), (wide[:, 256:4864], outs[2], False), (wide[:, :256], outs[3], False)]
    ops.colsum_finish_multi(jobs)
    again = [torch.empty_like(o) for o in outs]
    ops.colsum_finish_multi([(j[0], o, False) for j, o in zip(jobs, again)])
    refs = [y.double().sum(0), x.double().sum(0), wide[:, 256:4864].double().sum(0), wide[:, :256].double().sum(0)]
    for o, o2, r in zip(outs, again, refs):
        assert torch.equal(o, o2)
        assert (o.double() - r).abs().max().item() <= 2e-5 * max(1.0, r.abs().max().item())
    acc = outs[3].clone()
    ops.colsum_finish_multi([(wide[:, :256], acc, True)])
    assert torch.allclose(acc, 2 * outs[3], rtol=1e-6)


@pytest.mark.parametrize('M', [128, 100, 300, 1748, 20000, 42587])
@pytest.mark.parametrize('b_mn', [False, True])
def test_tc_chain(ops, M, b_mn):
    """lk_tc_chain: three chained 256x256 contractions with on-chip intermediates, every fused output against fp64:
    forward flavour (bias, bias, bias+tanh+row dot) with K-major weights; backward flavour (addend + column sums) with MN-major weights."""
    g = torch.Generator().manual_seed(M + int(b_mn))
    D = 256
    a = torch.randn(M, D, generator=g)
    ws = [torch.randn(D, D, generator=g) / D ** 0.5 for _ in range(3)]
    bs = [torch.randn(D, generator=g) * 0.1 for _ in range(3)]
    add = torch.randn(M, D, generator=g)
    dot = torch.randn(D, generator=g)
    ad, wd = a.double(), [w.double() for w in ws]
    mm = (lambda x, w: x @ w) if b_mn else (lambda x, w: x @ w.t())
    A = ops.split_planes(a.cuda())
    W = [ops.split_planes(w.cuda()) for w in ws]
    if not b_mn:
        r0 = mm(ad, wd[0]) + bs[0].double()
        r1 = mm(r0, wd[1]) + bs[1].double()
        r2 = torch.tanh(mm(r1, wd[2]) + bs[2].double())
        outs = ops.tc_chain(A, [dict(w=W[0], bias=bs[0].cuda(), want_planes=True),
                                dict(w=W[1], bias=bs[1].cuda(), want_f32=True, want_planes=True),
                                dict(w=W[2], bias=bs[2].cuda(), act=1, want_f32=True, dotvec=dot.cuda())], b_mn=False)
        assert rel(outs[0]['planes'].hi.float() + outs[0]['planes'].lo.float(), r0) <= 2e-5
        assert rel(outs[1]['f32'], r1) <= 2e-5
        assert rel(outs[1]['planes'].hi.float() + outs[1]['planes'].lo.float(), r1) <= 2e-5
        assert rel(outs[2]["f32"], r2) <= 6e-5
        assert rel(outs[2]["rowdot_part"].double().sum(1), r2 @ dot.double()) <= 6e-5
    else:
        r0 = mm(ad, wd[0]) + add.double()
        r1 = mm(r0, wd[1])
        r2 = mm(r1, wd[2])
        outs = ops.tc_chain(A, [dict(w=W[0], addsrc=add.cuda(), want_planes=True, want_colsum=True),
                                dict(w=W[1], want_planes=True, want_colsum=True),
                                dict(w=W[2], want_f32=True)], b_mn=True)
        assert rel(outs[0]['planes'].hi.float() + outs[0]['planes'].lo.float(), r0) <= 2e-5
        assert rel(outs[1]['planes'].hi.float() + outs[1]['planes'].lo.float(), r1) <= 2e-5
        assert rel(outs[2]["f32"], r2) <= 6e-5
        for o, r in ((outs[0], r0), (outs[1], r1)):
            cs = o['colsum_part'].double().sum(0).cpu()
            assert (cs - r.sum(0)).abs().max().item() <= 2e-5 * r.abs().sum(0).max().item()
    # one- and two-contraction chains run through the same kernel
    one = ops.tc_chain(A, [dict(w=W[0], want_f32=True) if b_mn else dict(w=W[0], bias=bs[0].cuda(), want_f32=True)], b_mn=b_mn)
    assert rel(one[0]['f32'], mm(ad, wd[0]) + (0 if b_mn else bs[0].double())) <= 2e-5
    two = ops.tc_chain(A, [dict(w=W[0]), dict(w=W[1], want_f32=True)], b_mn=b_mn)
    assert rel(two[1]['f32'], mm(mm(ad, wd[0]), wd[1])) <= 2e-5


@pytest.mark.parametrize('U,N,k', [(128, 128, 10), (300, 5000, 10), (70, 1000, 1), (4096, 40000, 16), (129, 131, 5)])
def test_sweep_topk(ops, U, N, k):
    """lk_sweep_topk (scores never materialised) against torch.topk of the fp64 score matrix: same items (ties and near-ties aside), scores
    within the split-bf16 bar; ragged user / item tiles, several item ranges."""
    g = torch.Generator().manual_seed(U + N + k)
    u = torch.randn(U, 256, generator=g)
    it = torch.randn(N, 256, generator=g)
    if N >= 1000:
        it[7] = it[3]                                   # an exact tie: the smaller item index must come first
    vals, idx = ops.sweep_topk(ops.split_planes(u.cuda()), ops.split_planes(it.cuda()), k)
    ref = u.double() @ it.double().t()
    rv, ri = torch.sort(ref, dim=1, descending=True, stable=True)
    rv, ri = rv[:, :k], ri[:, :k]
    vals, idx = vals.cpu().double(), idx.cpu()
    assert vals.shape == (U, k) and idx.dtype == torch.int64
    scale = ref.abs().max().item()
    assert (vals - rv).abs().max().item() <= 2e-5 * scale
    got_scores = torch.gather(ref, 1, idx)              # the items returned really have (almost) the k best scores
    assert (got_scores - rv).abs().max().item() <= 2e-5 * scale
    assert (idx == ri).double().mean().item() >= 0.999   # rank swaps only between scores closer than the arithmetic resolves
    for r in range(U):
        assert len(set(idx[r].tolist())) == k
    if N >= 1000:
        rows = (idx == 7).any(dim=1) & (idx == 3).any(dim=1)
        for r in torch.nonzero(rows).flatten().tolist():
            pos3, pos7 = idx[r].tolist().index(3), idx[r].tolist().index(7)
            assert pos3 < pos7
