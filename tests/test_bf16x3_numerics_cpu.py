"""CPU test (-m "not gpu") of the arithmetic the tensor-core kernels use -- not of the kernels: the "bf16x3" formulation of
DESIGN.md section 5 (x = hi + lo, hi = bf16(x), lo = bf16(x - hi); x.w evaluated as hi.hi + hi.lo + lo.hi with fp32
accumulation; the 3-way split of the training forward adds a third piece) emulated with torch.bfloat16 on the model's own
layer shapes, next to a single-pass TF32 emulation and plain fp32, all against float64.  It states in numbers why `dtype` in
bench.py reads "f32 via bf16x3": the formulation is a way of doing fp32 contractions on kind::f16 MMAs, two orders of
magnitude closer to fp32 than TF32, and inside the tolerance the parity tests write down."""
import torch


def _split(x, pieces):
    out, rest = [], x
    for _ in range(pieces):
        p = rest.bfloat16().float()
        out.append(p)
        rest = rest - p
    return out


def _bf16_contraction(x, w, pieces):
    """sum of the piece products whose order (i + j) stays below `pieces` (2 pieces -> hi.hi + hi.lo + lo.hi)"""
    xs, ws = _split(x, pieces), _split(w, pieces)
    acc = torch.zeros(x.shape[0], w.shape[1])
    for i, xi in enumerate(xs):
        for j, wj in enumerate(ws):
            if i + j < pieces:
                acc = acc + xi @ wj  # bf16 x bf16 products are exact in fp32; accumulation in fp32 like TMEM
    return acc


def _tf32(x):
    """round to nearest-even on a 10-bit mantissa"""
    i = x.contiguous().view(torch.int32)
    i = (i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF
    return i.view(torch.float32)


def test_bf16x3_is_an_fp32_class_contraction():
    g = torch.Generator().manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    rel = {}
    for cin, cout in ((64, 128), (128, 256), (256, 128), (128, 128)):       # the contractions of detector / descriptor
        x = torch.relu(torch.randn((4096, cin), generator=g))               # post-ReLU activations
        w = torch.randn((cin, cout), generator=g) * (2.6 / cin) ** 0.5      # the initialiser's scale
        want = x.double() @ w.double()
        scale = want.abs().max()
        err = lambda y: ((y.double() - want).abs().max() / scale).item()
        rel[(cin, cout)] = dict(fp32=err(x @ w), bf16x3=err(_bf16_contraction(x, w, 2)), bf16_3way=err(_bf16_contraction(x, w, 3)),
                                tf32=err(_tf32(x) @ _tf32(w)), bf16=err(x.bfloat16().float() @ w.bfloat16().float()))
    for shape, e in rel.items():
        assert e["fp32"] < 2e-6, (shape, e)
        assert e["bf16x3"] < 2e-5, (shape, e)                  # O(2^-16) dropped terms: inside the 2e-4 the parity tests allow
        assert e["bf16_3way"] < 4 * e["fp32"] + 1e-6, (shape, e)  # the training forward's 3-way split is fp32-accurate
        assert e["tf32"] > 20 * e["bf16x3"], (shape, e)        # single-pass TF32 is far outside it ...
        assert e["bf16"] > 100 * e["bf16x3"], (shape, e)       # ... and plain bf16 further still
