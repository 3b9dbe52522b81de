"""GPU parity tests proper: the CUDA path, called through the C ABI (include/wsis_b200.h via wsis_b200.ops /
the drop-in spconv + pointgroup_ops packages), against the CPU oracle on the same seeded inputs, against the
committed golden fixtures (outputs of the reference itself), and -- at BASELINE.json's full size -- through
size-independent properties.

Bars: integer / index outputs bit-exact; fp32 path 1e-4 relative; bf16 tensor-core path 1e-2 relative
(BASELINE.json north_star).  Relative error is max|a-b| / max|b| over a tensor.
"""
import itertools
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
BF16_TOL = 1e-2


def _need_gpu():
    assert torch.cuda.is_available(), "these tests must run on a CUDA box (-m gpu)"


@pytest.fixture(scope="module")
def W():
    _need_gpu()
    from wsis_b200 import ops
    return ops


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def gen_coords(rng, shape, npts, bs):
    cells = np.stack(np.meshgrid(*[np.arange(s) for s in shape], indexing="ij"), -1).reshape(-1, 3)
    out = []
    for b in range(bs):
        sel = rng.permutation(len(cells))[:npts]
        out.append(np.concatenate([np.full((len(sel), 1), b), cells[sel]], 1))
    return np.concatenate(out).astype(np.int32)


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ---------------------------------------------------------------------------------------------------------
# primitives
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [0, 1, 255, 2048, 2049, 100003, 1 << 21])
def test_exclusive_scan(W, n):
    import ctypes
    from wsis_b200._lib import lib
    rng = np.random.default_rng(n)
    x = rng.integers(0, 5, n).astype(np.int32)
    d = cu(x) if n else torch.empty(0, dtype=torch.int32, device="cuda")
    out = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    ws = torch.empty(lib().value("wsis_scan_ws_bytes", n), dtype=torch.uint8, device="cuda")
    lib().call("wsis_exclusive_scan_i32", W._ptr(d), W._ptr(out), n, W._ptr(ws), W._stream())
    ref = np.concatenate([[0], np.cumsum(x, dtype=np.int64)]).astype(np.int32)
    assert np.array_equal(out.cpu().numpy(), ref)


@pytest.mark.parametrize("n,bits", [(1, 8), (1000, 5), (2048, 8), (70001, 12), (300000, 17), (50000, 32)])
def test_radix_sort_stable(W, n, bits):
    from wsis_b200._lib import lib
    rng = np.random.default_rng(n + bits)
    keys = rng.integers(0, 1 << min(bits, 31), n).astype(np.uint32)
    if bits == 32:
        keys = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    vals = np.arange(n, dtype=np.uint32)
    dk, dv = cu(keys.view(np.int32)), cu(vals.view(np.int32))
    ok, ov = torch.empty_like(dk), torch.empty_like(dv)
    ws = torch.empty(lib().value("wsis_sort_ws_bytes", n), dtype=torch.uint8, device="cuda")
    lib().call("wsis_sort_pairs_u32", W._ptr(dk), W._ptr(dv), W._ptr(ok), W._ptr(ov), n, 0, bits, W._ptr(ws), W._stream())
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(ok.cpu().numpy().view(np.uint32), keys[order])
    assert np.array_equal(ov.cpu().numpy().view(np.uint32), vals[order])


# ---------------------------------------------------------------------------------------------------------
# rulebooks: bit-exact against the oracle (= the reference CPU path, tests/test_oracle_cpu.py pins that)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("bs,ksize,dil", [(1, 3, 1), (2, 3, 1), (2, 3, 2), (1, (1, 3, 3), 1), (2, (3, 1, 3), 1)])
def test_rulebook_subm_bit_exact(W, orc, bs, ksize, dil):
    rng = np.random.default_rng(11)
    shape = [19, 18, 17]
    c = gen_coords(rng, shape, 1500, bs)
    pairs, num = orc.rulebook_subm(c, bs, shape, ksize, dil)
    rb = W.rulebook_subm(cu(c), shape, ksize, dil)
    gp, gn = rb.pairs()
    assert np.array_equal(gn.cpu().numpy(), num)
    assert np.array_equal(gp.cpu().numpy(), pairs)  # same order as the CPU reference, not only the same set


@pytest.mark.parametrize("bs,k,s,p,d", [t for t in itertools.product([1, 2], [2, 3], [1, 2, 3], [0, 1, 2], [1, 2])
                                         if not (t[2] > 1 and t[4] > 1)])
def test_rulebook_conv_bit_exact(W, orc, bs, k, s, p, d):
    """The parameter grid of the reference's own testSpConv3d (test/test_conv.py:325-343)."""
    rng = np.random.default_rng(484)
    shape = [19, 18, 17]
    c = gen_coords(rng, shape, 1000, bs)
    oc, pairs, num, oshape = orc.rulebook_conv(c, bs, shape, k, s, p, d)
    if min(oshape) <= 0:
        pytest.skip("empty output shape")
    rb, gshape = W.rulebook_conv(cu(c), shape, k, s, p, d)
    assert gshape == oshape
    gp, gn = rb.pairs()
    assert np.array_equal(rb.out_coords.cpu().numpy(), oc)
    assert np.array_equal(gn.cpu().numpy(), num)
    assert np.array_equal(gp.cpu().numpy(), pairs)
    # nbr_out is the transpose of nbr_in
    nin, nout = rb.nbr_in.cpu().numpy(), rb.nbr_out.cpu().numpy()
    ii, kk = np.nonzero(nin >= 0)
    assert np.array_equal(nout[nin[ii, kk], kk], ii)
    assert (nout >= 0).sum() == (nin >= 0).sum()


def test_rulebook_edge_cases(W, orc):
    shape = [8, 8, 8]
    empty = torch.empty((0, 4), dtype=torch.int32, device="cuda")
    rb = W.rulebook_subm(empty, shape, 3, 1)
    p, n = rb.pairs()
    assert p.shape == (27, 2, 0) and int(n.sum()) == 0
    rb, _ = W.rulebook_conv(empty, shape, 2, 2, 0, 1)
    assert rb.n_out == 0
    # a single voxel in a corner, and a full dense block (every neighbour present)
    one = np.array([[0, 0, 0, 0]], np.int32)
    pairs, num = orc.rulebook_subm(one, 1, shape, 3, 1)
    gp, gn = W.rulebook_subm(cu(one), shape, 3, 1).pairs()
    assert np.array_equal(gn.cpu().numpy(), num) and np.array_equal(gp.cpu().numpy(), pairs)
    dense = gen_coords(np.random.default_rng(0), [6, 6, 6], 216, 2)
    pairs, num = orc.rulebook_subm(dense, 2, [6, 6, 6], 3, 1)
    gp, gn = W.rulebook_subm(cu(dense), [6, 6, 6], 3, 1).pairs()
    assert np.array_equal(gn.cpu().numpy(), num) and np.array_equal(gp.cpu().numpy(), pairs)
    # odd extent: the last row/column is dropped by a k2 s2 conv (9300 -> 9200 pairs)
    odd = gen_coords(np.random.default_rng(1), [9, 7, 5], 200, 1)
    oc, pairs, num, osh = orc.rulebook_conv(odd, 1, [9, 7, 5], 2, 2, 0, 1)
    rb, gs = W.rulebook_conv(cu(odd), [9, 7, 5], 2, 2, 0, 1)
    gp, gn = rb.pairs()
    assert gs == osh and np.array_equal(rb.out_coords.cpu().numpy(), oc) and np.array_equal(gp.cpu().numpy(), pairs)
    assert int(gn.sum()) < 200


def test_rulebook_full_size_properties(W):
    """150 000-voxel shell: pair counts are known from the reference run (1 347 750 pairs,
    k2s2: 150 000 pairs), the subm rulebook is symmetric and every pair is a true neighbour."""
    from wsis_b200.synthetic import make_shell
    c, shape = make_shell()
    d = cu(c.astype(np.int32))
    rb = W.rulebook_subm(d, shape, 3, 1)
    gp, gn = rb.pairs()
    num = gn.cpu().numpy()
    assert num.sum() == 1347750 and num[13] == 150000
    assert np.array_equal(num, num[::-1])
    pairs = gp.cpu().numpy()
    for k in (0, 5, 13, 26):
        i, o = pairs[k, 0, :num[k]], pairs[k, 1, :num[k]]
        off = np.array([k // 9 - 1, (k // 3) % 3 - 1, k % 3 - 1])
        assert np.array_equal(c[i, 1:] - c[o, 1:], np.broadcast_to(off, (num[k], 3)))  # in = out + (k - 1)
        assert np.all(np.diff(i) > 0)  # ascending input rows (CPU reference order)
    rb2, oshape = W.rulebook_conv(d, shape, 2, 2, 0, 1)
    assert oshape == [200, 150, 64] and rb2.n_out == 37400
    assert int(rb2.pairs()[1].sum()) == 150000
    oc = rb2.out_coords.cpu().numpy()
    nin = rb2.nbr_in.cpu().numpy()
    ii, kk = np.nonzero(nin >= 0)
    assert np.array_equal(oc[nin[ii, kk], 1:], c[ii, 1:] // 2)
    assert np.unique(oc, axis=0).shape[0] == oc.shape[0]


# ---------------------------------------------------------------------------------------------------------
# sparse convolution
# ---------------------------------------------------------------------------------------------------------
def _conv_case(rng, npts, bs, cin, cout, shape=(19, 18, 17)):
    c = gen_coords(rng, list(shape), npts, bs)
    f = rng.uniform(-1, 1, (len(c), cin)).astype(np.float32)
    w = (rng.uniform(-1, 1, (3, 3, 3, cin, cout)) / np.sqrt(cin)).astype(np.float32)
    return c, f, w


@pytest.mark.parametrize("cin,cout", [(6, 32), (3, 5), (32, 32), (64, 32), (32, 64), (96, 96), (128, 160), (160, 160),
                                      (320, 160), (64, 48), (40, 24)])
@pytest.mark.parametrize("prec,tol", [("simt", 1e-5), ("fp32", FP32_TOL), ("bf16", BF16_TOL)])
def test_subm_conv_forward(W, orc, cin, cout, prec, tol):
    rng = np.random.default_rng(cin * 1000 + cout)
    c, f, w = _conv_case(rng, 1500, 2, cin, cout)
    pairs, num = orc.rulebook_subm(c, 2, [19, 18, 17], 3, 1)
    ref = orc.indice_conv(f, w, pairs, num, len(c))
    rb = W.rulebook_subm(cu(c), [19, 18, 17], 3, 1)
    out = W.sparse_conv(cu(f), cu(w.reshape(27, cin, cout)), rb.nbr_in, len(c), 1, precision=prec)
    assert rel(out.cpu().numpy(), ref) < tol


@pytest.mark.parametrize("ksize,dil", [(3, 2), (2, 1), (4, 1), (3, 3)])
def test_subm_conv_dilated_and_even_kernels(W, orc, ksize, dil):
    """Submanifold rulebooks with dilation > 1 or an even kernel are NOT symmetric (the padding is ks/2 whatever the
    dilation, spconv_ops.h:74-77), so the forward / wgrad map cannot be the flipped input-side map: the module path
    (SubMConv3d forward, din and dW through autograd) is checked against the oracle on those shapes."""
    import spconv
    rng = np.random.default_rng(100 * ksize + dil)
    shape = [19, 18, 17]
    c = gen_coords(rng, shape, 1200, 2)
    cin, cout, K = 32, 32, ksize ** 3
    f = rng.uniform(-1, 1, (len(c), cin)).astype(np.float32)
    w = (rng.uniform(-1, 1, (ksize, ksize, ksize, cin, cout)) / np.sqrt(cin)).astype(np.float32)
    g = rng.uniform(-1, 1, (len(c), cout)).astype(np.float32)
    pairs, num = orc.rulebook_subm(c, 2, shape, ksize, dil)
    ref = orc.indice_conv(f, w.reshape(K, cin, cout), pairs, num, len(c))
    din_ref, dw_ref = orc.indice_conv_backward(f, w.reshape(K, cin, cout), g, pairs, num)
    conv = spconv.SubMConv3d(cin, cout, ksize, padding=ksize // 2, dilation=dil, bias=False, indice_key="k").cuda()
    with torch.no_grad():
        conv.weight.copy_(cu(w))
    x = cu(f).requires_grad_(True)
    out = conv(spconv.SparseConvTensor(x, cu(c), shape, 2)).features
    assert rel(out.detach().cpu().numpy(), ref) < FP32_TOL
    out.backward(cu(g))
    assert rel(x.grad.cpu().numpy(), din_ref) < FP32_TOL
    assert rel(conv.weight.grad.cpu().numpy().reshape(K, cin, cout), dw_ref) < FP32_TOL
    with torch.no_grad():                                   # the fused inference path uses the same maps
        out2 = conv(spconv.SparseConvTensor(cu(f), cu(c), shape, 2)).features
    assert rel(out2.cpu().numpy(), ref) < FP32_TOL


@pytest.mark.parametrize("prec,tol", [("simt", 2e-6), ("fp32", FP32_TOL), ("bf16", BF16_TOL)])
def test_conv_fused_prologue_and_residual(W, orc, prec, tol):
    rng = np.random.default_rng(5)
    c, f, w = _conv_case(rng, 2000, 1, 64, 32)
    scale = rng.uniform(0.5, 1.5, 64).astype(np.float32)
    shift = rng.uniform(-0.3, 0.3, 64).astype(np.float32)
    res = rng.uniform(-1, 1, (len(c), 32)).astype(np.float32)
    pairs, num = orc.rulebook_subm(c, 1, [19, 18, 17], 3, 1)
    ref = orc.indice_conv(np.maximum(f * scale + shift, 0), w, pairs, num, len(c)) + res
    rb = W.rulebook_subm(cu(c), [19, 18, 17], 3, 1)
    out = W.sparse_conv(cu(f), cu(w.reshape(27, 64, 32)), rb.nbr_in, len(c), 1, prologue=(cu(scale), cu(shift), 1),
                        residual=cu(res), precision=prec)
    assert rel(out.cpu().numpy(), ref) < tol


@pytest.mark.parametrize("prec,tol", [("simt", 2e-6), ("fp32", FP32_TOL), ("bf16", BF16_TOL)])
def test_strided_and_inverse_conv(W, orc, prec, tol):
    """The k2 s2 conv <-> inverse couple of the UNet (sparse_unet3d.py:258-298; reference test_conv.py:443-497)."""
    rng = np.random.default_rng(6)
    shape = [20, 18, 16]
    c = gen_coords(rng, shape, 2500, 2)
    f = rng.uniform(-1, 1, (len(c), 32)).astype(np.float32)
    wd = rng.uniform(-1, 1, (8, 32, 64)).astype(np.float32) / 6
    wu = rng.uniform(-1, 1, (8, 64, 32)).astype(np.float32) / 8
    oc, pairs, num, _ = orc.rulebook_conv(c, 2, shape, 2, 2, 0, 1)
    down_ref = orc.indice_conv(f, wd, pairs, num, len(oc))
    up_ref = orc.indice_conv(down_ref, wu, pairs, num, len(c), inverse=True)
    rb, _ = W.rulebook_conv(cu(c), shape, 2, 2, 0, 1)
    down = W.sparse_conv(cu(f), cu(wd), rb.nbr_out, rb.n_out, 0, precision=prec)
    assert rel(down.cpu().numpy(), down_ref) < tol
    up = W.sparse_conv(cu(down_ref), cu(wu), rb.nbr_in, rb.n_in, 0, precision=prec)
    assert rel(up.cpu().numpy(), up_ref) < tol


@pytest.mark.parametrize("k,s,p,d", [(3, 1, 1, 1), (3, 2, 1, 1), (2, 2, 0, 1), (3, 1, 2, 2), (3, 3, 0, 1)])
def test_sparse_conv_vs_dense_conv3d(W, k, s, p, d):
    """The reference's own golden check (test_conv.py:325-382): SparseConv3d == nn.Conv3d on the densified input,
    including input and weight gradients, atol 1e-4."""
    import spconv
    rng = np.random.default_rng(484)
    shape, bs, IC, OC = [19, 18, 17], 2, 32, 48
    c = gen_coords(rng, shape, 1000, bs)
    f = rng.uniform(-1, 1, (len(c), IC)).astype(np.float32)
    w = rng.uniform(0, 1, (k, k, k, IC, OC)).astype(np.float32)
    feats = cu(f).requires_grad_(True)
    net = spconv.SparseConv3d(IC, OC, k, s, p, d, bias=False).cuda()
    net.weight.data[:] = cu(w)
    out = net(spconv.SparseConvTensor(feats, cu(c), shape, bs)).dense()
    dense = torch.zeros((bs, IC, *shape), device="cuda")
    dense[cu(c[:, 0]).long(), :, cu(c[:, 1]).long(), cu(c[:, 2]).long(), cu(c[:, 3]).long()] = cu(f)
    dense.requires_grad_(True)
    ref_net = torch.nn.Conv3d(IC, OC, k, s, p, d, bias=False).cuda()
    ref_net.weight.data[:] = cu(w).permute(4, 3, 0, 1, 2).contiguous()
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        out_ref = ref_net(dense)
        dout = cu(rng.uniform(-0.2, 0.2, tuple(out_ref.shape)).astype(np.float32))
        out.backward(dout)
        out_ref.backward(dout)
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    scale = float(out_ref.abs().max())
    assert float((out - out_ref).abs().max()) < 1e-4 * max(scale, 1.0)
    din_ref = dense.grad.permute(0, 2, 3, 4, 1)[cu(c[:, 0]).long(), cu(c[:, 1]).long(), cu(c[:, 2]).long(),
                                                 cu(c[:, 3]).long()]
    assert float((feats.grad - din_ref).abs().max()) < 1e-4 * max(float(din_ref.abs().max()), 1.0)
    dw_ref = ref_net.weight.grad.permute(2, 3, 4, 1, 0)
    assert float((net.weight.grad - dw_ref).abs().max()) < 1e-4 * max(float(dw_ref.abs().max()), 1.0)


@pytest.mark.parametrize("kind", ["subm", "conv", "inverse"])
def test_conv_backward_vs_oracle(W, orc, kind):
    """indice_conv_backward (spconv_ops.h:351-433): din and dW."""
    from spconv import ops as sops
    rng = np.random.default_rng(8)
    shape = [20, 18, 16]
    c = gen_coords(rng, shape, 2000, 2)
    if kind == "subm":
        pairs, num = orc.rulebook_subm(c, 2, shape, 3, 1)
        n_in = n_out = len(c)
        K, cin, cout = 27, 32, 64
        outids, gp, gn = sops.get_indice_pairs(cu(c), 2, shape, 3, 1, 1, 1, 0, True)
    else:
        oc, pairs, num, _ = orc.rulebook_conv(c, 2, shape, 2, 2, 0, 1)
        K = 8
        outids, gp, gn = sops.get_indice_pairs(cu(c), 2, shape, 2, 2, 0, 1, 0, False)
        if kind == "conv":
            n_in, n_out, cin, cout = len(c), len(oc), 32, 64
        else:
            n_in, n_out, cin, cout = len(oc), len(c), 64, 32
    f = rng.uniform(-1, 1, (n_in, cin)).astype(np.float32)
    w = (rng.uniform(-1, 1, (K, cin, cout)) / np.sqrt(cin)).astype(np.float32)
    g = rng.uniform(-1, 1, (n_out, cout)).astype(np.float32)
    inv = kind == "inverse"
    din_ref, dw_ref = orc.indice_conv_backward(f, w, g, pairs, num, inverse=inv)
    din, dw = sops.indice_conv_backward(cu(f), cu(w), cu(g), gp, gn, inv, kind == "subm")
    assert rel(din.cpu().numpy(), din_ref) < FP32_TOL
    assert rel(dw.cpu().numpy(), dw_ref) < FP32_TOL
    # also through foreign reference-format pairs (no neighbour maps attached)
    gp2 = gp.clone()
    out = sops.indice_conv(cu(f), cu(w), gp2, gn, n_out, inv, kind == "subm")
    assert rel(out.cpu().numpy(), orc.indice_conv(f, w, pairs, num, n_out, inverse=inv)) < FP32_TOL


def test_conv_full_size_linearity(W):
    """150k-voxel shell, C=32: conv(a*x + y) == a*conv(x) + conv(y) and the tensor-core path agrees with the
    exact-fp32 SIMT path (size-independent properties at BASELINE.json's full size)."""
    from wsis_b200.synthetic import make_shell
    c, shape = make_shell()
    rb = W.rulebook_subm(cu(c.astype(np.int32)), shape, 3, 1)
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.rand((len(c), 32), device="cuda", generator=g) - 0.5
    y = torch.rand((len(c), 32), device="cuda", generator=g) - 0.5
    w = (torch.rand((27, 32, 32), device="cuda", generator=g) - 0.5) * 0.3
    fx = W.sparse_conv(x, w, rb.nbr_in, len(c), 1, precision="fp32")
    fy = W.sparse_conv(y, w, rb.nbr_in, len(c), 1, precision="fp32")
    fxy = W.sparse_conv(2.5 * x + y, w, rb.nbr_in, len(c), 1, precision="fp32")
    s = float(fxy.abs().max())
    assert float((fxy - (2.5 * fx + fy)).abs().max()) < 2e-4 * s
    exact = W.sparse_conv(x, w, rb.nbr_in, len(c), 1, precision="simt")
    assert float((fx - exact).abs().max()) < FP32_TOL * float(exact.abs().max())
    b16 = W.sparse_conv(x, w, rb.nbr_in, len(c), 1, precision="bf16")
    assert float((b16 - exact).abs().max()) < BF16_TOL * float(exact.abs().max())


# ---------------------------------------------------------------------------------------------------------
# voxelization, segmented reductions
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,span", [(1, 4), (5000, 20), (60000, 64), (3000, 3)])
def test_voxelization_device_matches_oracle(W, orc, n, span):
    import pointgroup_ops
    rng = np.random.default_rng(n)
    coords = np.concatenate([rng.integers(0, 3, (n, 1)), rng.integers(0, span, (n, 3))], 1).astype(np.int64)
    locs, p2v, v2p = orc.voxelization_idx(coords, 3, 4)
    glocs, gp2v, gv2p = pointgroup_ops.voxelization_idx(cu(coords), 3, 4)
    assert np.array_equal(glocs.cpu().numpy(), locs)
    assert np.array_equal(gp2v.cpu().numpy(), p2v)
    assert np.array_equal(gv2p.cpu().numpy(), v2p)
    hlocs, hp2v, hv2p = pointgroup_ops.voxelization_idx(torch.from_numpy(coords), 3, 4)  # host path, same tensors
    assert np.array_equal(hlocs.numpy(), locs) and np.array_equal(hp2v.numpy(), p2v) and np.array_equal(hv2p.numpy(), v2p)
    feats = rng.uniform(-1, 1, (n, 6)).astype(np.float32)
    f = cu(feats).requires_grad_(True)
    out = pointgroup_ops.voxelization(f, gv2p, 4)
    assert rel(out.detach().cpu().numpy(), orc.voxelization(feats, v2p)) < 1e-6
    g = rng.uniform(-1, 1, out.shape).astype(np.float32)
    out.backward(cu(g))
    assert rel(f.grad.cpu().numpy(), orc.voxelization_backward(g, v2p, n)) < 1e-6


@pytest.mark.parametrize("reduce", ["sum", "mean", "max"])
@pytest.mark.parametrize("n,S,C", [(1000, 37, 32), (150000, 3000, 32), (5000, 400, 3), (4000, 100, 1), (300, 299, 64)])
def test_segment_reduce_matches_oracle(W, orc, reduce, n, S, C):
    rng = np.random.default_rng(n + S)
    ids = rng.integers(0, S, n).astype(np.int64)
    ids[:min(n, S)] = np.arange(min(n, S))  # contiguous ids 0..S-1 as asserted at scannetv2_dataset.py:424
    src = rng.uniform(-1, 1, (n, C)).astype(np.float32)
    ref = orc.scatter(src, ids, reduce, S)
    out = W.scatter(cu(src), cu(ids), 0, reduce, S)
    assert rel(out.cpu().numpy(), ref) < 2e-6
    out2 = W.scatter(cu(src), cu(ids), 0, reduce)  # dim_size inferred like torch_scatter
    assert out2.shape[0] == ids.max() + 1


def test_gather_then_pool_fused(W, orc):
    rng = np.random.default_rng(2)
    vox = rng.uniform(-1, 1, (5000, 32)).astype(np.float32)
    p2v = rng.integers(0, 5000, 8000).astype(np.int32)
    sp = rng.integers(0, 300, 8000).astype(np.int64)
    ref = orc.scatter(vox[p2v], sp, "mean", 300)
    seg = W.SegmentIndex(cu(sp), 300)
    out = W.segment_reduce(cu(vox), seg, "mean", gather=cu(p2v))
    assert rel(out.cpu().numpy(), ref) < 2e-6
    assert np.array_equal(W.gather_rows(cu(vox), cu(p2v)).cpu().numpy(), vox[p2v])


# ---------------------------------------------------------------------------------------------------------
# affinity + random walk
# ---------------------------------------------------------------------------------------------------------
def _graph(rng, S, deg):
    e = set()
    for u in range(S):
        for v in rng.choice(S, deg, replace=False):
            if u != v:
                e.add((u, int(v)))
                e.add((int(v), u))
    return np.array(sorted(e), np.int64)


def test_edge_attention_matches_oracle(W, orc):
    rng = np.random.default_rng(4)
    S = 700
    e = _graph(rng, S, 4)
    e = e[rng.permutation(len(e))]  # the kernel must not rely on sorted edges
    q, k, v, ecc = (rng.standard_normal((S, 64)).astype(np.float32) for _ in range(4))
    cen = rng.uniform(0, 5, (S, 3)).astype(np.float32)
    w1, b1 = rng.standard_normal((16, 3)).astype(np.float32), rng.standard_normal(16).astype(np.float32)
    w2, b2 = rng.standard_normal((1, 16)).astype(np.float32), rng.standard_normal(1).astype(np.float32)
    aff_ref, sp_ref = orc.edge_attention(q, k, v, ecc, cen, e[:, 0], e[:, 1], w1, b1, w2, b2)
    eu, ev = cu(e[:, 0]), cu(e[:, 1])
    pos = cu(np.concatenate([w1.ravel(), b1, w2.ravel(), b2]))
    aff, sp = W.edge_attention(cu(q), cu(k), cu(v), cu(ecc), cu(cen), eu, ev, W.SegmentIndex(eu, S), pos)
    assert rel(aff.cpu().numpy(), aff_ref) < 1e-5
    assert rel(sp.cpu().numpy(), sp_ref) < 1e-5
    sums = np.zeros(S)
    np.add.at(sums, e[:, 0], aff.cpu().numpy())
    assert np.allclose(sums[np.unique(e[:, 0])], 1.0, atol=1e-5)  # softmax rows sum to one


@pytest.mark.parametrize("iterations", [0, 1, 2])
def test_random_walk_matches_reference_numpy(W, orc, iterations):
    """Pseudo labels identical, scores to 1e-12 (float64 both sides; only the summation order differs)."""
    rng = np.random.default_rng(10 + iterations)
    S, classes = 500, 20
    e = _graph(rng, S, 4)
    aff = rng.uniform(0.01, 1, len(e)).astype(np.float32)
    seed = np.full(S, -100, np.int64)
    pick = rng.choice(S, 40, replace=False)
    seed[pick] = rng.integers(0, classes, 40)
    pred = rng.integers(0, 6, S).astype(np.int64)
    pred[pick] = np.where(rng.random(40) < 0.8, seed[pick], pred[pick])
    conf = rng.uniform(0.4, 1.0, S).astype(np.float32)
    adj = np.zeros((S, S))
    adj[e[:, 0], e[:, 1]] = 1
    final_ref, score_ref = orc.weak_label_propagation(seed, adj, conf, pred, orc.dense_affinity(e[:, 0], e[:, 1], aff, S),
                                                      iterations, classes)
    pseudo, score = W.random_walk(cu(e[:, 0]), cu(e[:, 1]), cu(aff), cu(seed), cu(pred), cu(conf), classes, iterations)
    got = pseudo.cpu().numpy().astype(np.float64)
    assert np.array_equal(got, final_ref), "differs at %s" % np.nonzero(got != final_ref)[0][:10]
    assert np.abs(score.cpu().numpy() - score_ref).max() < 1e-12
    assert (final_ref != -100).sum() > 0


@pytest.mark.parametrize("iterations", [0, 1, 3])
def test_random_walk_matches_reference_golden(W, golden_dir, iterations):
    """The device random walk against outputs of the reference's OWN method (scannetv2_dataset.py:664-735 executed by
    tests/golden/make_golden_rw.py) on a full scene's superpoint graph: labels identical, scores to 1e-12."""
    g = np.load(os.path.join(golden_dir, "rw_scene0.npz"))
    e = g["edges"].astype(np.int64)
    pseudo, score = W.random_walk(cu(e[:, 0]), cu(e[:, 1]), cu(g["aff"]), cu(g["seed_label"].astype(np.int64)),
                                  cu(g["pred"].astype(np.int64)), cu(g["conf"]), 20, iterations)
    assert np.array_equal(pseudo.cpu().numpy().astype(np.int32), g["pseudo_it%d" % iterations])
    assert np.abs(score.cpu().numpy() - g["score_it%d" % iterations]).max() < 1e-12


# ---------------------------------------------------------------------------------------------------------
# end to end against the reference's own outputs (tests/golden/make_golden.py)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["scene_b1", "scene_b2"])
@pytest.mark.parametrize("prec,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL), ("simt", 2e-5)])
def test_network_matches_reference_golden(W, golden_dir, name, prec, tol):
    """The reference Network (run on the reference's spconv CPU kernels) vs the mirror on the CUDA path."""
    import ast
    from wsis_b200 import pipeline, synthetic
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    nscene, kw = ast.literal_eval(str(gold["scene_kw"]))
    batch = synthetic.collate([synthetic.make_scene(9000 + i, **kw) for i in range(nscene)])
    net = pipeline.build_network(seed=123, device="cpu")
    g = torch.Generator().manual_seed(7)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) * 0.5 + 0.75)
            if m.affine:
                m.weight.data.copy_(torch.rand(m.weight.shape, generator=g) * 0.5 + 0.75)
                m.bias.data.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
    chk = np.array([float(p.detach().double().abs().sum()) for p in net.parameters()])
    # orthogonal_ (LAPACK QR) may round differently on another host CPU, hence 1e-5 and not bit equality
    assert np.allclose(chk, gold["param_checksum"], rtol=1e-5), "fixed-seed initialisation differs from the reference"
    net = net.cuda().eval()
    dbatch, _ = pipeline.to_device(batch)
    old = W.get_precision()
    W.set_precision(prec)
    try:
        with torch.no_grad():
            ret, aux = pipeline.forward_batch(net, dbatch)
    finally:
        W.set_precision(old)
    # integer outputs: bit exact
    assert np.array_equal(aux["voxel_locs"].cpu().numpy(), gold["voxel_locs"])
    assert np.array_equal(aux["p2v_map"].cpu().numpy(), gold["p2v"])
    assert np.array_equal(aux["v2p_map"].cpu().numpy(), gold["v2p"])
    inp = aux["input"]
    for key in ("subm1", "spconv1", "subm2", "spconv2"):
        outids, _, pairs, num, _ = inp.indice_dict[key]
        if hasattr(pairs, "tensors"):       # inference pipeline: the reference-format tensors are written on demand
            pairs, num = pairs.tensors()
        assert np.array_equal(outids.cpu().numpy(), gold["rb_%s_outids" % key])
        assert np.array_equal(num.cpu().numpy(), gold["rb_%s_num" % key])
        if key in ("subm1", "spconv1"):
            assert np.array_equal(pairs.cpu().numpy(), gold["rb_%s_pairs" % key])
    assert rel(aux["centers"].cpu().numpy(), gold["centers"]) < 1e-6
    for k in ("semantic_scores", "sp_semantic_scores", "pred_sp_offset_vectors", "pred_sp_occupancy",
              "pred_sp_ins_size", "edge_affinity", "sp_discriminative_feats"):
        assert rel(ret[k].cpu().numpy(), gold["ret_" + k]) < tol, k


def test_drop_in_module_api_unfused_equals_fused(W):
    """SparseSequential's fused BN+ReLU+conv path gives the same result as the module-by-module path."""
    import spconv
    from torch import nn
    rng = np.random.default_rng(3)
    shape = [19, 18, 17]
    c = gen_coords(rng, shape, 1500, 1)
    f = cu(rng.uniform(-1, 1, (len(c), 32)).astype(np.float32))
    torch.manual_seed(0)
    seq = spconv.SparseSequential(nn.BatchNorm1d(32, eps=1e-4), nn.ReLU(),
                                  spconv.SubMConv3d(32, 64, 3, padding=1, bias=False, indice_key="k")).cuda().eval()
    seq[0].running_mean.uniform_(-0.2, 0.2)
    seq[0].running_var.uniform_(0.7, 1.3)
    with torch.no_grad():
        fused = seq(spconv.SparseConvTensor(f, cu(c), shape, 1)).features
    x = spconv.SparseConvTensor(f, cu(c), shape, 1)
    with torch.no_grad():
        x.features = torch.relu(seq[0](x.features))
        plain = seq[2](x).features
    assert float((fused - plain).abs().max()) < FP32_TOL * float(plain.abs().max())


def test_public_api_train_mode_autograd(W, orc):
    """A SparseSequential(BatchNorm1d, ReLU, SubMConv3d) built only through the public names of the drop-in, in TRAIN
    mode with autograd (batch statistics, no fusion): forward and the gradients of the input and of every parameter
    against torch BatchNorm + the oracle's conv forward / backward (spconv_ops.h:253-433)."""
    import spconv
    rng = np.random.default_rng(77)
    shape = [19, 18, 17]
    c = gen_coords(rng, shape, 1500, 2)
    cin, cout = 32, 64
    f = rng.uniform(-1, 1, (len(c), cin)).astype(np.float32)
    g = rng.uniform(-1, 1, (len(c), cout)).astype(np.float32)
    torch.manual_seed(3)
    seq = spconv.SparseSequential(torch.nn.BatchNorm1d(cin, eps=1e-4, momentum=0.1), torch.nn.ReLU(),
                                  spconv.SubMConv3d(cin, cout, 3, padding=1, bias=False, indice_key="k")).cuda().train()
    x = cu(f).requires_grad_(True)
    out = seq(spconv.SparseConvTensor(x, cu(c), shape, 2)).features
    out.backward(cu(g))
    # reference formulation on the host: torch BN (train) + ReLU, then the oracle's indice_conv / backward
    xr = torch.from_numpy(f).requires_grad_(True)
    bn = torch.nn.BatchNorm1d(cin, eps=1e-4, momentum=0.1).train()
    with torch.no_grad():
        bn.weight.copy_(seq[0].weight.detach().cpu())
        bn.bias.copy_(seq[0].bias.detach().cpu())
    a = torch.relu(bn(xr))
    w = seq[2].weight.detach().cpu().numpy().reshape(27, cin, cout)
    pairs, num = orc.rulebook_subm(c, 2, shape, 3, 1)
    ref = orc.indice_conv(a.detach().numpy(), w, pairs, num, len(c))
    da, dw = orc.indice_conv_backward(a.detach().numpy(), w, g, pairs, num)
    a.backward(torch.from_numpy(da))
    assert rel(out.detach().cpu().numpy(), ref) < FP32_TOL
    assert rel(seq[2].weight.grad.cpu().numpy().reshape(27, cin, cout), dw) < FP32_TOL
    assert rel(x.grad.cpu().numpy(), xr.grad.numpy()) < 2e-4
    assert rel(seq[0].weight.grad.cpu().numpy(), bn.weight.grad.numpy()) < 2e-4
    assert rel(seq[0].running_mean.cpu().numpy(), bn.running_mean.numpy()) < 1e-5


# ---------------------------------------------------------------------------------------------------------
# BASELINE.json configs[0] / configs[1] at full size against the reference's own CPU kernels
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_scenes", [1, 4])
def test_full_size_parity_vs_reference_cpu_kernels(W, n_scenes):
    """One 150k-point scene and the batch of 4 (BASELINE.json configs[0], configs[1]) through pipeline.forward_batch
    on the device and through oracle/cpu_pipeline.forward on the host, whose rulebooks and convolutions are the
    UNMODIFIED reference spconv CPU kernels compiled into oracle/_ref (getIndicePair / indiceConv,
    include/spconv/spconv_ops.h:27-137, 253-349).  All nine rulebooks (conv.py:149-152) must be equal as sorted pair
    sets in coordinate space, the voxelization maps identical, and the U-Net output and every entry of the result
    dict within the fp32 contract (1e-4 of the largest magnitude).  The bf16-operand path is measured on the same
    batch and must meet its 1e-2 contract per tensor except where noted."""
    from oracle import cpu_pipeline, parity
    from wsis_b200 import pipeline, synthetic
    batch = synthetic.collate([synthetic.make_scene(2000 + i, n_points=150000) for i in range(n_scenes)])
    cpu_net = pipeline.build_network(seed=123, device="cpu").eval()
    keep = {}
    cpu_ret, _, kind = cpu_pipeline.forward(cpu_net, batch, keep=keep)
    if kind != "reference":
        pytest.skip("oracle/_ref (compiled reference spconv) is not available on this box")
    net = pipeline.build_network(seed=123, device="cuda").eval()
    dbatch, _ = pipeline.to_device(batch)
    W.set_precision("fp32")
    try:
        with torch.no_grad():
            ret, aux = pipeline.forward_batch(net, dbatch, keep_unet_features=True)
        res = parity.compare(ret, aux, cpu_ret, keep)
        assert res["rulebooks_equal"], res["rulebooks"]
        assert len(res["rulebooks"]) == 9 and res["rulebooks"]["subm1"]["pairs"] > 100000 * n_scenes
        assert res["voxelization_equal"]
        assert res["max_rel"] < FP32_TOL, res["outputs"]
        W.set_precision("bf16")
        with torch.no_grad():
            ret16, aux16 = pipeline.forward_batch(net, dbatch, keep_unet_features=True)
        res16 = parity.compare(ret16, aux16, cpu_ret, keep)
        print("bf16 end-to-end relative errors:", {k: "%.2e" % v for k, v in res16["outputs"].items()})
        # measured on the B200 (profiles/r02_parity_full_size.txt): 2-3e-3 end to end, 49 bf16-operand layers deep
        assert res16["max_rel"] < BF16_TOL, res16["outputs"]
    finally:
        W.set_precision("fp32")


# ---------------------------------------------------------------------------------------------------------
# spatial tiling (tilemap.cu) and the tiled tensor-core kernel
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,bs,shape", [(1, 1, [8, 8, 8]), (127, 1, [19, 18, 17]), (128, 2, [19, 18, 17]),
                                        (5000, 3, [40, 36, 32]), (60000, 4, [400, 300, 128])])
def test_spatial_order_is_a_local_permutation(W, n, bs, shape):
    rng = np.random.default_rng(n)
    c = np.concatenate([rng.integers(0, bs, (n, 1)), rng.integers(0, shape[0], (n, 1)), rng.integers(0, shape[1], (n, 1)),
                        rng.integers(0, min(shape[2], 4), (n, 1))], 1).astype(np.int32)   # a thin slab = a surface
    order = W.spatial_order(cu(c), shape, bs).cpu().numpy()
    n_pad = (n + 127) // 128 * 128
    assert order.shape[0] == max(n_pad, 1) or order.shape[0] == n_pad
    assert np.array_equal(np.sort(order[:n]), np.arange(n)) and np.all(order[n:n_pad] == -1)
    s = c[order[:n]]
    assert np.all(np.diff(s[:, 0]) >= 0)                                  # batch-major
    if n >= 5000:
        step = np.abs(np.diff(s[:, 1:].astype(np.int64), axis=0)).sum(1)
        rnd = np.abs(np.diff(c[:, 1:].astype(np.int64), axis=0)).sum(1)
        assert np.median(step) * 8 < np.median(rnd)                       # curve order is spatially local


def _decode_records(t, K):
    """Tile records (include/wsis_b200.h: wsis_tile_records) -> dense int32[n_pad, K] map, checking the layout."""
    rec = t.records.cpu().numpy()
    meta = t.meta.cpu().numpy()
    uidx = t.uidx.cpu().numpy()
    assert t.stride == 16 * K + 80 + 256 * K
    out = np.full((t.num_tiles * 128, K), -1, np.int32)
    for ti in range(t.num_tiles):
        r = rec[ti * t.stride:(ti + 1) * t.stride]
        valid = r[:16 * K].view(np.uint32).reshape(K, 4)
        hdr = r[16 * K:16 * K + 16].view(np.uint32)
        nU, amask, P, npack = int(hdr[0]), int(hdr[1]), int(hdr[2]), int(hdr[3])
        members = r[16 * K + 16:16 * K + 80].view(np.uint16)
        locr = r[16 * K + 80:].view(np.uint16).reshape(K, 128).astype(np.int64)   # one row per pack
        bits = np.unpackbits(valid.view(np.uint8).reshape(K, 16), axis=1, bitorder="little").astype(bool)
        assert P == bits.sum() and amask == (sum(1 << k for k in range(K) if bits[k].any()) or 1)
        assert list(meta[ti]) == [t.stride, nU, npack, P]
        # packs: every active offset is a member of exactly one pack, at most two per pack, disjoint valid slots
        seen = []
        loc = np.full((K, 128), 0xFFFF, np.int64)
        assert 1 <= npack <= max(1, int(bits.any(1).sum())) and np.all(members[npack:] == 0xFFFF)
        for i in range(npack):
            ks = [k for k in (int(members[i]) & 0xFF, int(members[i]) >> 8) if k != 0xFF]
            assert (len(ks) >= 1 or P == 0) and all(bits[k].any() for k in ks)
            if len(ks) == 2:
                assert not (bits[ks[0]] & bits[ks[1]]).any()
            seen += ks
            for k in ks:
                loc[k, bits[k]] = locr[i, bits[k]]
            union = np.zeros(128, bool)
            for k in ks:
                union |= bits[k]
            assert np.all(locr[i, ~union] == 0xFFFF)
        assert sorted(seen) == [k for k in range(K) if bits[k].any()]
        uniq = uidx[ti * t.ustride:ti * t.ustride + nU]
        assert len(np.unique(uniq)) == nU                      # the tile's source rows, each exactly once
        assert P == 0 or (loc[bits].max() < nU and len(np.unique(loc[bits])) == nU)
        for k in range(K):
            sl = np.nonzero(bits[k])[0]
            out[ti * 128 + sl, k] = uniq[loc[k, sl]]
    return out


@pytest.mark.parametrize("K", [27, 8, 1])
def test_tile_records_match_numpy(W, K):
    rng = np.random.default_rng(2 + K)
    n = 1000
    m = rng.integers(-1, n, (n, K)).astype(np.int32)
    m[rng.random((n, K)) < 0.6] = -1
    m[130:260] = -1                                          # a tile without any entry
    m[300:428] = rng.integers(0, n, (128, K))                # a full tile (worst-case record size)
    order = np.concatenate([rng.permutation(n), np.full(24, -1)]).astype(np.int32)
    for flip in (0, 1):
        t = W.TileMap(cu(m), n, flip, cu(order))
        exp = np.full((1024, K), -1, np.int32)
        exp[:n] = m[order[:n]][:, ::-1] if flip else m[order[:n]]
        assert t.num_tiles == 8 and np.array_equal(_decode_records(t, K), exp)


@pytest.mark.parametrize("npts,bs", [(100, 1), (128, 1), (3000, 2), (12000, 2)])
@pytest.mark.parametrize("prec,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL)])
def test_subm_conv_morton_tiles(W, orc, npts, bs, prec, tol):
    """The production path: Morton-ordered tiles, fused BN+ReLU prologue and residual epilogue; > 148 tiles in the
    largest case so every CTA walks several tiles through the double-buffered map/accumulator."""
    rng = np.random.default_rng(npts)
    shape = [40, 36, 32]
    c = gen_coords(rng, shape, npts, bs)
    cin, cout = 64, 32
    f = rng.uniform(-1, 1, (len(c), cin)).astype(np.float32)
    w = (rng.uniform(-1, 1, (27, cin, cout)) / np.sqrt(cin)).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, cin).astype(np.float32)
    shift = rng.uniform(-0.3, 0.3, cin).astype(np.float32)
    res = rng.uniform(-1, 1, (len(c), cout)).astype(np.float32)
    pairs, num = orc.rulebook_subm(c, bs, shape, 3, 1)
    ref = orc.indice_conv(np.maximum(f * scale + shift, 0), w, pairs, num, len(c)) + res
    rb = W.rulebook_subm(cu(c), shape, 3, 1, batch_size=bs)
    out = W.sparse_conv(cu(f), cu(w), rb.nbr_in, len(c), 1, prologue=(cu(scale), cu(shift), 1), residual=cu(res),
                        precision=prec, tiles=rb.tiles_out())
    assert rel(out.cpu().numpy(), ref) < tol
    # dgrad through the input-side tiles: din = sum_k g[nbr_in[i,k]] W[k]^T
    g = rng.uniform(-1, 1, (len(c), cout)).astype(np.float32)
    din_ref, _ = orc.indice_conv_backward(f, w, g, pairs, num)
    din = W.sparse_conv(cu(g), cu(w), rb.nbr_in, len(c), 0, True, precision=prec, tiles=rb.tiles_in())
    assert rel(din.cpu().numpy(), din_ref) < tol


@pytest.mark.parametrize("prec,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL)])
def test_strided_and_inverse_conv_morton_tiles(W, orc, prec, tol):
    rng = np.random.default_rng(16)
    shape = [44, 38, 30]
    c = gen_coords(rng, shape, 9000, 2)
    f = rng.uniform(-1, 1, (len(c), 32)).astype(np.float32)
    wd = rng.uniform(-1, 1, (8, 32, 64)).astype(np.float32) / 6
    wu = rng.uniform(-1, 1, (8, 64, 32)).astype(np.float32) / 8
    oc, pairs, num, _ = orc.rulebook_conv(c, 2, shape, 2, 2, 0, 1)
    down_ref = orc.indice_conv(f, wd, pairs, num, len(oc))
    up_ref = orc.indice_conv(down_ref, wu, pairs, num, len(c), inverse=True)
    rb, _ = W.rulebook_conv(cu(c), shape, 2, 2, 0, 1, batch_size=2)
    down = W.sparse_conv(cu(f), cu(wd), rb.nbr_out, rb.n_out, 0, precision=prec, tiles=rb.tiles_out())
    assert rel(down.cpu().numpy(), down_ref) < tol
    up = W.sparse_conv(cu(down_ref), cu(wu), rb.nbr_in, rb.n_in, 0, precision=prec, tiles=rb.tiles_in())
    assert rel(up.cpu().numpy(), up_ref) < tol


def test_conv_tile_without_any_neighbour(W):
    """A whole tile whose map rows are all -1 must still produce zeros (+ residual), not stale accumulator data."""
    rng = np.random.default_rng(4)
    n, K, cin, cout = 384, 27, 32, 32
    m = np.full((n, K), -1, np.int32)
    m[128:256, 13] = np.arange(128, 256)                       # only the middle tile has neighbours (identity)
    f = rng.uniform(-1, 1, (n, cin)).astype(np.float32)
    w = rng.uniform(-1, 1, (K, cin, cout)).astype(np.float32)
    res = rng.uniform(-1, 1, (n, cout)).astype(np.float32)
    out = W.sparse_conv(cu(f), cu(w), cu(m), n, 0, residual=cu(res), precision="fp32").cpu().numpy()
    exp = res.copy()
    exp[128:256] += f[128:256] @ w[13]
    assert rel(out, exp) < FP32_TOL
    assert np.array_equal(out[:128], res[:128]) and np.array_equal(out[256:], res[256:])


@pytest.mark.parametrize("cin,cout", [(6, 32), (32, 32), (64, 48), (100, 16)])
@pytest.mark.parametrize("prec,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL)])
def test_conv_row_cache_overflow_and_narrow_rows(W, cin, cout, prec, tol):
    """Tiles that read far more distinct source rows than a row-cache buffer holds (a random map: up to 128*K distinct
    rows per tile) take the direct-fetch path for the overflow; widths that are not a multiple of 32 (or of 4) are
    zero-padded inside the kernel.  Checked against a direct numpy evaluation of the definition."""
    rng = np.random.default_rng(cin * 1000 + cout)
    n, K = 700, 27
    m = rng.integers(0, n, (n, K)).astype(np.int32)
    m[rng.random((n, K)) < 0.35] = -1
    m[256:384] = rng.integers(0, n, (128, K))                  # a full tile: 3456 entries
    f = rng.uniform(-1, 1, (n, cin)).astype(np.float32)
    w = (rng.uniform(-1, 1, (K, cin, cout)) / np.sqrt(cin)).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, cin).astype(np.float32)
    shift = rng.uniform(-0.3, 0.3, cin).astype(np.float32)
    g = np.maximum(f * scale + shift, 0).astype(np.float64)
    exp = np.zeros((n, cout))
    for k in range(K):
        ok = m[:, k] >= 0
        exp[ok] += g[m[ok, k]] @ w[k].astype(np.float64)
    out = W.sparse_conv(cu(f), cu(w), cu(m), n, 0, prologue=(cu(scale), cu(shift), 1), precision=prec).cpu().numpy()
    assert rel(out, exp) < tol


def test_s3dis_shaped_room_full_size_properties(W):
    """BASELINE.json configs[3]: one S3DIS-shaped room (1 M points, 5 cm voxels, ~6k superpoints) through the whole
    hot path.  At this size the CPU oracle is replaced by size-independent properties: voxelization is a partition of
    the points, the submanifold rulebook is symmetric, the tensor-core conv agrees with the exact-fp32 kernel and is
    linear, superpoint pooling agrees with an index_add formulation, and the network outputs are finite."""
    import pointgroup_ops
    from wsis_b200 import pipeline, synthetic
    room = synthetic.make_room_s3dis(7, n_points=1000000)
    batch = synthetic.collate([room])
    dbatch, _ = pipeline.to_device(batch)
    n = batch["locs"].shape[0]
    locs, p2v, v2p = pointgroup_ops.voxelization_idx(dbatch["locs"], 1, 4)
    m = locs.shape[0]
    assert m == len(np.unique(room["locs"], axis=0))
    assert int(v2p[:, 0].sum()) == n and int(p2v.max()) == m - 1
    assert torch.equal(locs[p2v.long()], dbatch["locs"])                   # every point lies in the voxel it maps to
    rb = W.rulebook_subm(locs.int(), batch["spatial_shape"], 3, 1, batch_size=1)
    valid = (rb.nbr_in >= 0)
    num = valid.sum(0).cpu().numpy()
    assert num[13] == m and np.array_equal(num, num[::-1])
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.rand((m, 32), device="cuda", generator=g) - 0.5
    y = torch.rand((m, 32), device="cuda", generator=g) - 0.5
    w = (torch.rand((27, 32, 32), device="cuda", generator=g) - 0.5) * 0.3
    tiles = rb.tiles_out()
    fx = W.sparse_conv(x, w, rb.nbr_in, m, 1, precision="fp32", tiles=tiles)
    exact = W.sparse_conv(x, w, rb.nbr_in, m, 1, precision="simt")
    scale = float(exact.abs().max())
    assert float((fx - exact).abs().max()) < FP32_TOL * scale
    fy = W.sparse_conv(y, w, rb.nbr_in, m, 1, precision="fp32", tiles=tiles)
    fxy = W.sparse_conv(2.0 * x - y, w, rb.nbr_in, m, 1, precision="fp32", tiles=tiles)
    assert float((fxy - (2.0 * fx - fy)).abs().max()) < 2e-4 * scale
    sp = dbatch["superpoint"]
    S = batch["num_superpoints"]
    pts = torch.rand((n, 32), device="cuda", generator=g)
    pooled = W.segment_reduce(pts, W.SegmentIndex(sp, S), "mean")
    ref = torch.zeros((S, 32), device="cuda").index_add_(0, sp, pts) / torch.bincount(sp, minlength=S).clamp(min=1)[:, None]
    assert float((pooled - ref).abs().max()) < 1e-5
    net = pipeline.build_network(seed=123, device="cuda").eval()
    with torch.no_grad():
        ret, _ = pipeline.forward_batch(net, dbatch)
    assert ret["semantic_scores"].shape == (n, 20) and ret["edge_affinity"].shape[0] == batch["edge_u_list"].shape[0]
    for k, v in ret.items():
        assert bool(torch.isfinite(v).all()), k


@pytest.mark.parametrize("filter_free,tol", [(False, 2e-5), (True, 1e-4)])
@pytest.mark.parametrize("S,E,layernorm", [(700, 6000, True), (50, 40, True), (300, 2500, False), (33, 1, True),
                                           (3000, 40000, True)])
def test_ecc_gru_fused_matches_module(W, orc, S, E, layernorm, filter_free, tol):
    """The inference ECC-GRU kernels against the module-by-module torch formulation of spg_modules.py:152-185 /
    226-253 (NNConv mean aggregation + GRUCellEx), which the golden network test pins to the reference:
    filter_free=False: csrc/ecc.cu streams the materialised [E,1024] filters (one kernel per step);
    filter_free=True:  csrc/ecc_umma.cu regenerates the filters on the tensor cores inside every step (bf16 hi/mid
    split operands, the fp32 contract) -- no [E,1024] tensor exists.  Includes superpoints without in-edges, unsorted
    targets, duplicate edges and edge counts that are not multiples of the 128-edge tile."""
    from wsis_b200 import model as M
    torch.manual_seed(S + E)
    rng = np.random.default_rng(S * 7 + E)
    fnet = M.create_fnet([13, 32, 128, 64, 32 * 32], True, True, 2)
    cell = M.GRUCellEx(32, 32, bias=True, layernorm=layernorm, ingate=True)
    mod = M.RNNGraphConvModule(cell, fnet, 32, nrepeats=7, cat_all=True).cuda().eval()
    src = rng.integers(0, S, E)
    tgt = rng.integers(0, max(S - 3, 1), E)                    # the last superpoints never receive a message
    edge_index = cu(np.stack([src, tgt]).astype(np.int64).reshape(2, E))
    feats = cu(rng.standard_normal((E, 13)).astype(np.float32))
    mod.set_info(M.GraphInfo(edge_index, feats))
    hx = cu(rng.standard_normal((S, 32)).astype(np.float32))
    ref = mod(hx).detach()                                      # grad mode: the torch formulation
    old = W.ECC_FUSED_FILTERS
    W.ECC_FUSED_FILTERS = filter_free
    try:
        assert W.ecc_fnet_supported(fnet)
        with torch.no_grad():
            out = mod(hx)                                       # inference: the fused kernels
    finally:
        W.ECC_FUSED_FILTERS = old
    assert out.shape == ref.shape == (S, 32 * 8)
    assert torch.equal(out[:, :32], hx)
    assert rel(out.cpu().numpy(), ref.cpu().numpy()) < tol
    # and against the float64 oracle, step by step from the kernel's own previous state
    ig = cell._modules["ig"]
    with torch.no_grad():
        w = fnet.cuda()(feats).cpu().numpy()
    o = out.cpu().numpy()
    for r in range(7):
        exp = orc.ecc_gru_step(o[:, 32 * r:32 * r + 32], w, src, tgt, ig.weight.detach().cpu().numpy(),
                               ig.bias.detach().cpu().numpy(), cell.weight_ih.detach().cpu().numpy(),
                               cell.weight_hh.detach().cpu().numpy(), cell.bias_ih.detach().cpu().numpy(),
                               cell.bias_hh.detach().cpu().numpy(), layernorm=layernorm)
        assert rel(o[:, 32 * r + 32:32 * r + 64], exp) < tol


@pytest.mark.parametrize("C,cout,n", [(32, 20, 70001), (64, 20, 3000), (64, 3, 3000), (64, 1, 1), (64, 7, 257), (32, 32, 129)])
def test_fused_mlp_head_matches_torch(W, C, cout, n):
    """csrc/heads.cu: Linear -> BatchNorm1d(eval) -> ReLU -> Linear (backbone_3D_WSIS.py:57-62, 71-104) in one kernel,
    with and without the fused voxel -> point gather, against the torch modules."""
    torch.manual_seed(C + cout)
    head = torch.nn.Sequential(torch.nn.Linear(C, C), torch.nn.BatchNorm1d(C, eps=1e-4), torch.nn.ReLU(inplace=True),
                               torch.nn.Linear(C, cout)).cuda().eval()
    with torch.no_grad():
        head[1].running_mean.uniform_(-0.5, 0.5)
        head[1].running_var.uniform_(0.5, 2.0)
        head[1].weight.uniform_(0.5, 1.5)
        head[1].bias.uniform_(-0.5, 0.5)
        x = torch.randn(n, C, device="cuda")
        assert W.mlp_head_supported(head, x)
        ref = head(x.clone())
        assert rel(W.mlp_head(head, x).cpu().numpy(), ref.cpu().numpy()) < 1e-5
        idx = torch.randint(0, n, (2 * n + 3,), device="cuda", dtype=torch.int32)
        got = W.mlp_head(head, x, gather=idx)
        assert rel(got.cpu().numpy(), ref[idx.long()].cpu().numpy()) < 1e-5
        # parameter edits through load_state_dict are seen (cache epoch)
        sd = {k: v.clone() for k, v in head.state_dict().items()}
        sd["3.bias"] += 1.0
        head.load_state_dict(sd)
        W.invalidate_caches()
        assert rel(W.mlp_head(head, x).cpu().numpy(), (ref + 1.0).cpu().numpy()) < 1e-5


@pytest.mark.parametrize("seed", [2000, 2001])
def test_final_instance_masks_identical_to_reference_cpu_path(W, seed):
    """BASELINE.json north_star: "final instance masks identical on fixed seeds".  One scene goes through the whole
    chain twice -- network on the CUDA path vs network on the reference's CPU kernels (oracle/cpu_pipeline), same
    weights -- and each set of outputs through the instance clustering (test_scannetv2.py:203-262 argmax + :281-455
    clustering_in_graph, pinned to the reference's own function by tests/test_cpu.py): instance masks, labels and the
    superpoint-level semantic argmax must be identical, confidences equal to 1e-5."""
    from oracle import cpu_pipeline
    from wsis_b200 import cluster, pipeline, synthetic
    sc = synthetic.make_scene(seed, n_points=16000, room=(4.0, 3.0, 2.6), n_boxes=6)   # ~1000 superpoints
    batch = synthetic.collate([sc])
    cpu_net = pipeline.build_network(seed=123, device="cpu").eval()
    ref, _, _ = cpu_pipeline.forward(cpu_net, batch)
    net = pipeline.build_network(seed=123, device="cuda").eval()
    with torch.no_grad():
        got, _ = pipeline.forward_batch(net, pipeline.to_device(batch)[0])
    torch.cuda.synchronize()
    nbrs = cluster.neighbors_from_edges(sc["edges"], sc["num_superpoints"])
    out = []
    for r in (ref, {k: v.cpu() for k, v in got.items()}):
        sem = r["sp_semantic_scores"].max(1)[1].numpy()                         # test_scannetv2.py:206-207
        out.append((sem,) + cluster.clustering_in_graph(sc["xyz"], sc["superpoint"], nbrs, sem,
                                                        r["pred_sp_offset_vectors"].numpy(), r["pred_sp_occupancy"].numpy(),
                                                        r["pred_sp_ins_size"].numpy()))
    (sem_a, conf_a, lab_a, mask_a), (sem_b, conf_b, lab_b, mask_b) = out
    assert np.array_equal(sem_a, sem_b)
    assert mask_a.shape == mask_b.shape and np.array_equal(mask_a, mask_b)
    assert np.array_equal(lab_a, lab_b)
    assert np.allclose(conf_a, conf_b, rtol=1e-5, atol=0)


@pytest.mark.parametrize("name", ["cluster_scene0", "cluster_scene1"])
def test_device_clustering_matches_reference_golden(W, name):
    """csrc/cluster.cu (one warp walks the superpoint graph, point-sized work in parallel kernels) against the outputs of
    the reference's OWN clustering_in_graph (test_scannetv2.py:281-455, executed by tests/golden/make_golden_cluster.py):
    identical instance masks and labels, confidences to 1e-6 (the group means are summed in member order instead of
    numpy's pairwise float32 order)."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tests", "golden"))
    import make_golden_cluster as mg
    from wsis_b200 import cluster
    gold = np.load(os.path.join(root, "tests", "golden", name + ".npz"))
    case, nbrs = mg.make_case(int(gold["seed"]), int(gold["n_points"]))
    S = len(case["sem"])
    edges = np.array([(s, n) for s in range(S) for n in nbrs[s]], dtype=np.int64).reshape(-1, 2)
    csr = cluster.neighbors_csr_device(cu(edges), S)
    off = csr[0].cpu().numpy()
    assert all(csr[1][off[s]:off[s + 1]].cpu().tolist() == list(nbrs[s]) for s in range(0, S, 97))
    conf, label, point_inst, inst_of_sp = cluster.clustering_in_graph_device(
        cu(case["xyz"]), cu(case["superpoint"].astype(np.int64)), csr, cu(case["sem"].astype(np.int64)), cu(case["off"]),
        cu(case["occ"]), cu(case["size"]), num_superpoints=S)
    I = conf.shape[0]
    masks = cluster.dense_masks(point_inst, I).cpu().numpy()
    assert np.array_equal(np.packbits(masks.astype(bool), axis=1), gold["masks"])
    assert np.array_equal(label.cpu().numpy(), gold["label_id"])
    assert np.allclose(conf.cpu().numpy(), gold["conf"], rtol=1e-6, atol=0)


def test_device_clustering_on_network_outputs_equals_host(W):
    """Network (CUDA path) -> argmax -> instance clustering entirely on the device, against the host restatement fed
    with the same network outputs: identical masks and labels on a full-size scene."""
    from wsis_b200 import cluster, pipeline, synthetic
    sc = synthetic.make_scene(2003, n_points=60000)
    batch = synthetic.collate([sc])
    net = pipeline.build_network(seed=123, device="cuda").eval()
    db = pipeline.to_device(batch)[0]
    with torch.no_grad():
        ret, aux = pipeline.forward_batch(net, db)
    S = sc["num_superpoints"]
    sem = ret["sp_semantic_scores"].max(1)[1]
    csr = cluster.neighbors_csr_device(cu(sc["edges"]), S)
    conf, label, point_inst, _ = cluster.clustering_in_graph_device(
        cu(sc["xyz"]), db["superpoint"], csr, sem, ret["pred_sp_offset_vectors"], ret["pred_sp_occupancy"],
        ret["pred_sp_ins_size"], num_superpoints=S)
    nbrs = cluster.neighbors_from_edges(sc["edges"], S)
    hconf, hlabel, hmasks = cluster.clustering_in_graph(sc["xyz"], sc["superpoint"], nbrs, sem.cpu().numpy(),
                                                        ret["pred_sp_offset_vectors"].cpu().numpy(),
                                                        ret["pred_sp_occupancy"].cpu().numpy(),
                                                        ret["pred_sp_ins_size"].cpu().numpy())
    I = conf.shape[0]
    assert I == len(hconf) and I > 10
    host_inst = np.where(hmasks.any(0), hmasks.argmax(0), -1)
    assert np.array_equal(point_inst.cpu().numpy(), host_inst)
    assert np.array_equal(label.cpu().numpy(), hlabel)
    assert np.allclose(conf.cpu().numpy(), hconf, rtol=1e-5, atol=0)


def test_batch_stream_and_result_fetcher_equal_the_direct_path(W):
    """pipeline.BatchStream (H2D on a copy stream into persistent staging sets, overlapped with the previous batch's
    compute) + ResultFetcher (async D2H into pinned buffers) give bit-identical results to to_device + forward_batch, for
    batches of different sizes that recycle the staging sets, also when a second loader inherits the first one's sets."""
    from wsis_b200 import pipeline, synthetic
    net = pipeline.build_network(seed=123, device="cuda").eval()
    sizes = [9000, 14000, 6000, 14000, 9000]
    host = [pipeline.pin_batch(synthetic.collate([synthetic.make_scene(2100 + i, n_points=n)])) for i, n in enumerate(sizes)]
    ref = []
    for b in host:
        with torch.no_grad():
            ret, _ = pipeline.forward_batch(net, pipeline.to_device(b)[0])
        ref.append({k: ret[k].cpu().clone() for k in ("edge_affinity", "sp_semantic_scores", "semantic_scores")})
    fetch = pipeline.ResultFetcher(keys=("edge_affinity", "sp_semantic_scores"))
    first = pipeline.BatchStream(host[:2])
    got = []
    for stream in (first, pipeline.BatchStream(host[2:], copy_stream=first.copy_stream, staging=first.staging)):
        for db, nb in stream:
            assert nb > 0
            with torch.no_grad():
                ret, _ = pipeline.forward_batch(net, db)
            out, _ = fetch.fetch(ret)
            fetch.wait()
            got.append({"edge_affinity": out["edge_affinity"].clone(), "sp_semantic_scores": out["sp_semantic_scores"].clone(),
                        "semantic_scores": ret["semantic_scores"].cpu()})
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        for k in b:
            assert torch.equal(a[k], b[k]), k


def test_batch_stream_with_geometry_prefetch_equals_the_direct_path(W):
    """BatchStream(prepare=True): the coordinate-only part of batch i+1 (prepare_geometry: voxelization maps, all nine
    rulebooks, tile records, segment indices) is built on the loader's side stream while batch i computes, and
    forward_batch consumes it instead of building rulebooks lazily inside the convs; the results must be bit-identical
    to the lazy path, for batches of different sizes (the full-size version of this loop is tools/stream_debug.py: it
    found a barrier phase-aliasing bug in ecc_messages_kernel that only co-resident CTAs of a second stream expose)."""
    from wsis_b200 import pipeline, synthetic
    net = pipeline.build_network(seed=123, device="cuda").eval()
    sizes = [9000, 14000, 6000, 14000, 9000, 11000]
    host = [pipeline.pin_batch(synthetic.collate([synthetic.make_scene(2200 + i, n_points=n)])) for i, n in enumerate(sizes)]
    keys = ("edge_affinity", "sp_semantic_scores", "semantic_scores", "pred_sp_offset_vectors")
    ref = []
    for b in host:
        with torch.no_grad():
            ret, _ = pipeline.forward_batch(net, pipeline.to_device(b)[0])
        ref.append({k: ret[k].cpu().clone() for k in keys})
    got = []
    for db, _ in pipeline.BatchStream(host, prepare=True):
        assert "_geometry" in db and len(db["_geometry"]["indice_dict"]) == 9
        with torch.no_grad():
            ret, _ = pipeline.forward_batch(net, db)
        got.append({k: ret[k].cpu().clone() for k in keys})
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        for k in keys:
            assert torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("n,cin,cout,linear", [(70001, 64, 32, False), (5000, 320, 160, False), (129, 128, 64, False),
                                               (12345, 64, 64, True), (1, 64, 64, True)])
def test_dense_rows_on_the_conv_kernel_matches_torch(W, n, cin, cout, linear):
    """A 1x1 submanifold conv (conv.py:113-119: torch.mm) / a bias-free Linear over rows as a sparse conv with the identity
    rulebook on the tensor-core kernel: within the fp32 contract of a float64 matmul."""
    torch.manual_seed(n + cin)
    x = torch.randn(n, cin, device="cuda")
    w = torch.randn((cout, cin) if linear else (cin, cout), device="cuda") / cin ** 0.5
    ref = (x.double() @ (w.double().t() if linear else w.double())).float()
    holder = torch.empty(1, device="cuda")
    for _ in range(2):                                       # second call: cached identity tiles and packed weights
        got = W.dense_rows(x, w, packed=W.PackedWeights(), holder=holder, linear_layout=linear)
        assert got is not None and rel(got.cpu().numpy(), ref.cpu().numpy()) < FP32_TOL
    assert W.dense_rows(x, w, precision="simt", linear_layout=linear) is None


def test_segment_index_validation_raises_on_out_of_range_ids(W):
    """A wrong segment count must not silently merge rows into the last segment (torch_scatter raises too)."""
    ids = cu(np.array([0, 1, 5, 2], dtype=np.int64))
    x = cu(np.ones((4, 3), np.float32))
    with pytest.raises(RuntimeError):
        W.scatter(x, ids, reduce="sum", dim_size=4)
    with pytest.raises(RuntimeError):
        W.SegmentIndex(ids, 3).validate()
    assert W.scatter(x, ids, reduce="sum", dim_size=6).shape == (6, 3)
